"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(idoc_b200._lib -> libidoc_b200.so), against (i) golden vectors produced by executing the reference
sources and (ii) the NumPy oracle on seeded random inputs.  Tolerances are north_star's:
rel 1e-10 in fp64, rel 1e-4 in fp32 (max-norm relative to the oracle array's max-norm)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-4}
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lqr_*.npz")))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ib():
    import idoc_b200
    idoc_b200._lib.require_device()
    return idoc_b200


def _params(ib, z):
    return ib.lqr.Params(z["x0"], ib.lqr.LQR(**{k: z[k] for k in ib.lqr.LQR._fields}))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_fwd_matches_reference_golden(ib, path):
    z = np.load(path)
    p = _params(ib, z)
    T = z["A"].shape[0]
    solver = ib.lqr.build(T)
    s = solver.direct(p)
    tol = 1e-10
    assert rel(s.X, z["X"]) < tol and rel(s.U, z["U"]) < tol and rel(s.Nu, z["Nu"]) < tol
    gains, ec = ib.lqr.backward(p.lqr, T, return_expected_change=True)
    assert rel(gains.K, z["K"]) < tol and rel(gains.k, z["k"]) < tol
    want = z["dC"] + z["dc"]
    assert abs(ec(1.0) - want) <= 1e-9 * max(abs(z["dC"]), abs(z["dc"]))
    r = solver.kkt(ib.typs.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), p)
    assert rel(r.X, z["kkt_X"]) < tol and rel(r.U, z["kkt_U"]) < tol and rel(r.Nu, z["kkt_Nu"]) < tol


@pytest.mark.parametrize("path", [p for p in GOLD if "shift" not in p],
                         ids=[os.path.basename(p)[:-4] for p in GOLD if "shift" not in p])
def test_vjp_matches_reference_kkt_dense_solve(ib, path):
    z = np.load(path)
    p = _params(ib, z)
    solver = ib.lqr.build(z["A"].shape[0])
    s, pull = solver.vjp(p)
    for pre, ct in (("g_", ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"])),
                    ("g0_", ib.typs.State(z["X"], z["U"], None))):
        g = pull(ct)
        assert rel(g.x0, z[pre + "x0"]) < 1e-10
        for name in ib.lqr.LQR._fields:
            want = z[pre + name]
            got = getattr(g.lqr, name)
            assert got.shape == want.shape
            assert rel(got, want) < 1e-10, name


def test_shift_status_flag(ib):
    path = [c for c in GOLD if "shift_n4m2" in c][0]
    z = np.load(path)
    p = ib.lqr.DeviceProblem.from_params(_params(ib, z))
    plan = ib.lqr.DevicePlan(p.B, p.T, p.n, p.m, p.dtype, vjp=False)
    plan.fwd(p)
    assert int(plan.status.numpy()[0]) & 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(8, 4, 25, 37), (4, 2, 50, 5), (3, 2, 5, 64), (2, 1, 9, 33), (5, 3, 7, 3),
                                   (3, 3, 6, 10), (4, 1, 11, 17), (8, 4, 3, 1), (8, 4, 1, 9), (8, 4, 2, 16), (2, 1, 1, 5),
                                   (16, 8, 1, 3), (16, 8, 2, 4), (32, 16, 1, 2),
                                   # BASELINE config 5 shape at its full horizon, config 3's shape time-varying, config 2's horizon
                                   (32, 16, 200, 4), (16, 8, 200, 3), (8, 4, 100, 24),
                                   # even horizons: two timesteps per request; multiples of 4: four (fp32 adjoint-forward sweep)
                                   (8, 4, 10, 13), (8, 4, 12, 9), (8, 4, 4, 8)])
def test_random_batch_vs_oracle(ib, dtype, shape):
    from oracle import lqr_np as O
    n, m, T, B = shape
    rng = np.random.default_rng(1000 + n * 100 + m * 10 + T)
    p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)  # fp32: oracle runs in fp64 on the rounded inputs
    so, go = O.direct(p, return_gains=True)
    solver = ib.lqr.build(T)
    pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
    s, pull = solver.vjp(pp)
    tol = TOL[dtype]
    assert s.X.dtype == dtype
    assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol
    gains = ib.lqr.backward(pp.lqr, T)
    assert rel(gains.K, go.K) < tol and rel(gains.k, go.k) < tol
    for nub in (True, False):
        w = O.State(rng.standard_normal(so.X.shape), rng.standard_normal(so.U.shape),
                    rng.standard_normal(so.Nu.shape) if nub else np.zeros_like(so.Nu))
        w = O.State(*[a.astype(dtype) for a in w])
        g_o = O.vjp_exact(p, so, w)
        g = pull(ib.typs.State(w.X, w.U, w.Nu if nub else None))
        vt = tol  # north_star's bar holds for every gradient leaf (measured: <= 5e-15 / 5e-6, tests/checks/parity_report.py)
        assert rel(g.x0, g_o.x0) < vt
        for name in ib.lqr.LQR._fields:
            assert rel(getattr(g.lqr, name), getattr(g_o.lqr, name)) < vt, (name, nub)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(8, 4, 10, 21), (8, 4, 12, 700), (16, 8, 6, 5), (32, 16, 4, 3), (10, 3, 5, 4), (3, 2, 5, 40),
                                   (5, 3, 7, 9), (2, 1, 9, 33), (64, 8, 3, 2), (8, 4, 1, 5)])
def test_reduce_batch_vjp(ib, shape, dtype):
    """batch-summed gradients (reduce_batch = 1: vectors-only sweeps + the batch contraction of csrc/grad_reduce.cuh) on the
    tuned, mid-size and generic kernels, edge tiles (n, m not multiples of 4), several batch chunks (B = 700), several
    tile groups (n = 32, 64): every leaf against the oracle's per-problem gradients summed over the batch, within
    north_star's bar; bit-identical from run to run (no atomics)."""
    from oracle import lqr_np as O
    n, m, T, B = shape
    rng = np.random.default_rng(5)
    p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
    so = O.direct(p)
    w = O.State(*[rng.standard_normal(a.shape).astype(dtype) for a in so])
    g_o = O.vjp_exact(p, so, w)
    L = ib._lib
    dp = ib.lqr.DeviceProblem.from_params(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, reduce_batch=True)
    plan.fwd(dp)
    ct = [L.DeviceArray.from_numpy(a) for a in w]
    plan.vjp(dp, *ct)
    tol = TOL[dtype]
    first = plan.grad_buffer.numpy().copy()
    for name in ib.lqr.LQR._fields:
        got = plan.grads[name].numpy()
        want = getattr(g_o.lqr, name).sum(axis=0)
        assert got.shape == want.shape
        # relative to the largest summand magnitude: a sum over B terms of mixed sign can cancel
        scale = max(np.max(np.abs(want)), np.max(np.abs(getattr(g_o.lqr, name))))
        assert np.max(np.abs(got - want)) <= tol * scale, name  # (T = 1: gQ, gq are exactly zero on both sides)
    assert rel(plan.grads["x0"].numpy(), g_o.x0) < tol
    plan.vjp(dp, *ct)
    assert np.array_equal(plan.grad_buffer.numpy(), first)  # deterministic summation order


@pytest.mark.parametrize("ramp", [True, False])
def test_host_pipeline_matches_oracle(ib, ramp):
    """The pinned-host chunk pipeline (bench.py's e2e leg): every chunk of the ramped schedule lands in the right rows of
    the host arrays; solution and gradients (cotangent (X, U, 0)) against the oracle."""
    from oracle import lqr_np as O
    from idoc_b200 import host
    n, m, T, B, chunk = 8, 4, 6, 64, 16
    rng = np.random.default_rng(11)
    p = O.random_params(rng, n, m, T, batch=(B,))
    so = O.direct(p)
    g_o = O.vjp_exact(p, so, O.State(so.X, so.U, np.zeros_like(so.Nu)))
    hp, hs, hg = host.pinned_like_problem(B, T, n, m, np.float64)
    for k in ib.lqr.LEAVES:
        hp[k].a[...] = getattr(p.lqr, k)
    hp["x0"].a[...] = p.x0
    pipe = host.HostPipeline(B, T, n, m, np.float64, chunk=chunk, nslots=2, ramp=ramp)
    assert sum(pipe.sizes) == B and (pipe.sizes[:3] == [2, 4, 10]) == ramp
    for _ in range(2):  # second pass reuses every slot
        pipe.solve_and_vjp(hp, hs, hg)
        ib._lib.sync()
    assert rel(hs["X"].a, so.X) < 1e-10 and rel(hs["U"].a, so.U) < 1e-10 and rel(hs["Nu"].a, so.Nu) < 1e-10
    for k in ib.lqr.LEAVES:
        assert rel(hg[k].a, getattr(g_o.lqr, k)) < 1e-10, k
    assert rel(hg["x0"].a, g_o.x0) < 1e-10
    # batch-summed gradients through the same pipeline: one summed buffer per chunk leaves the device
    hp2, hs2, hg2 = host.pinned_like_problem(B, T, n, m, np.float64, reduce_batch=True)
    for k in ib.lqr.LEAVES + ("x0",):
        hp2[k].a[...] = hp[k].a
    pipe2 = host.HostPipeline(B, T, n, m, np.float64, chunk=chunk, nslots=2, ramp=ramp, reduce_batch=True)
    for _ in range(2):
        pipe2.solve_and_vjp(hp2, hs2, hg2)
    assert rel(hs2["X"].a, so.X) < 1e-10
    for k in ib.lqr.LEAVES:
        want = getattr(g_o.lqr, k)
        assert np.max(np.abs(hg2[k].a - want.sum(axis=0))) < 1e-10 * np.max(np.abs(want)), k
    assert rel(hg2["x0"].a, g_o.x0) < 1e-10 and pipe2.d2h_bytes < pipe.d2h_bytes / 5
    if ramp:
        return
    # the same pipeline behind ONE C-ABI call per batch (idoc_lqr_host_solve_vjp), ragged last chunk (B = 64, chunk = 24),
    # per-problem and batch-summed gradients, blocking and enqueue-only (wait = 0, two batches in flight) calls
    for reduce in (False, True):
        hp3, hs3, hg3 = host.pinned_like_problem(B, T, n, m, np.float64, reduce_batch=reduce)
        for k in ib.lqr.LEAVES + ("x0",):
            hp3[k].a[...] = hp[k].a
        nat = host.NativeHostPipeline(B, T, n, m, np.float64, chunk=24, nslots=2, reduce_batch=reduce)
        ms = nat.solve_and_vjp(hp3, hs3, hg3)
        assert ms > 0 and nat.h2d_bytes == pipe.h2d_bytes
        for rep in range(3):
            for k in hg3:
                hg3[k].a[...] = 0
            nat.solve_and_vjp(hp3, hs3, hg3, wait=False)
        nat.finish()
        assert rel(hs3["X"].a, so.X) < 1e-10 and rel(hs3["U"].a, so.U) < 1e-10 and rel(hs3["Nu"].a, so.Nu) < 1e-10
        for k in ib.lqr.LEAVES:
            want = getattr(g_o.lqr, k)
            want = want.sum(axis=0) if reduce else want
            assert np.max(np.abs(hg3[k].a - want)) < 1e-10 * np.max(np.abs(getattr(g_o.lqr, k))), (k, reduce)
        assert rel(hg3["x0"].a, g_o.x0) < 1e-10
        nat.close()


def test_errors(ib):
    L = ib._lib
    rng = np.random.default_rng(0)
    from oracle import lqr_np as O
    p = O.random_params(rng, 130, 5, 4, batch=(2,))
    with pytest.raises(L.IdocError):
        ib.lqr.build(4).direct(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))  # n > 128: unsupported -> loud error


@pytest.mark.parametrize("dtype,tol,kkt_tol", [(np.float64, 1e-10, 1e-10), (np.float32, 1e-4, 1e-5)])
def test_full_size_kkt_property(ib, dtype, tol, kkt_tol):
    """BASELINE config 2 at full size (B=65536, n=8, m=4, T=100), both dtypes: the device solution of
    device-generated problems satisfies the KKT conditions (the reference's own acceptance test,
    idoc/utils.py:16-27: mean |residual| < 1e-5; asserted at 1e-10 in fp64) and a sampled subset matches the oracle
    (solution AND every gradient leaf) within north_star's bar."""
    from oracle import lqr_np as O
    L = ib._lib
    B, T, n, m = 65536, 100, 8, 4
    dp = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=0)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True)
    plan.fwd(dp)
    rX, rU, rNu = (L.DeviceArray(a.shape, dtype) for a in (plan.X, plan.U, plan.Nu))
    plan.kkt(dp, plan.X, plan.U, plan.Nu, rX, rU, rNu)
    for r in (rX, rU, rNu):
        a = r.numpy()
        assert np.isfinite(a).all() and np.mean(np.abs(a)) < kkt_tol
    assert not plan.status.numpy().any()
    par = dp.to_params()
    idx = np.array([0, 1, 7, 8, 4095, 32768, 65535])
    sub = O.Params(par.x0[idx], O.LQR(*[a[idx] for a in par.lqr]))
    so = O.direct(sub)
    X, U, Nu = plan.X.numpy(), plan.U.numpy(), plan.Nu.numpy()
    assert rel(X[idx], so.X) < tol and rel(U[idx], so.U) < tol and rel(Nu[idx], so.Nu) < tol
    # implicit VJP with the reference tests' cotangent (X, U, 0) on the whole batch, sampled problems against the oracle
    plan.vjp(dp, plan.X, plan.U, None)
    g_o = O.vjp_exact(sub, so, O.State(X[idx].astype(np.float64), U[idx].astype(np.float64), np.zeros_like(so.Nu)))
    assert rel(plan.grads["x0"].numpy()[idx], g_o.x0) < tol
    for name in ib.lqr.LEAVES:
        assert rel(plan.grads[name].numpy()[idx], getattr(g_o.lqr, name)) < tol, name


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-4)])
def test_full_size_vjp_adjoint_identity(ib, dtype, tol):
    """BASELINE config 2 at full size, fwd + implicit VJP: the solution is AFFINE in (q, r, d, x0), so for any perturbation
    delta of those leaves and any cotangent w,  <w, s(theta + delta) - s(theta)> = <vjp(w), delta>  exactly (no step-size
    error): the transpose identity of the implicit VJP, checked per problem on all 65536 problems (even horizon: the paired
    requests of the streaming sweeps)."""
    L = ib._lib
    B, T, n, m = 65536, 100, 8, 4
    dp = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=3)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True)
    plan.fwd(dp)
    s0 = [a.numpy().astype(np.float64) for a in (plan.X, plan.U, plan.Nu)]
    rng = np.random.default_rng(0)
    w = [rng.standard_normal(a.shape).astype(dtype) for a in s0]
    plan.vjp(dp, *[L.DeviceArray.from_numpy(a) for a in w])
    keys = ("q", "r", "d", "x0")
    delta = {k: rng.standard_normal(dp.shapes[k]).astype(dtype) for k in keys}
    rhs = sum((plan.grads[k].numpy().astype(np.float64) * delta[k]).reshape(B, -1).sum(axis=1) for k in keys)
    for k in keys:  # theta + delta, in place on the device copy of the problem
        dp.arr[k].copy_from(dp.arr[k].numpy() + delta[k])
    plan.fwd(dp)
    s1 = [a.numpy().astype(np.float64) for a in (plan.X, plan.U, plan.Nu)]
    lhs = sum((wi.astype(np.float64) * (b - a)).reshape(B, -1).sum(axis=1) for wi, a, b in zip(w, s0, s1))
    assert not plan.status.numpy().any()
    scale = np.abs(lhs).max()
    assert np.isfinite(lhs).all() and np.abs(lhs - rhs).max() < tol * scale, (np.abs(lhs - rhs).max(), scale)


@pytest.mark.parametrize("shape", [(8, 4, 6, 9), (16, 8, 5, 4), (7, 5, 4, 5), (3, 2, 5, 6)])
@pytest.mark.parametrize("where", ["A", "q", "R"])
def test_nonfinite_inputs_are_reported_and_propagated(ib, shape, where):
    """The reference propagates NaN from its inputs into X, U, Nu.  The kernels' pivot clamps must not turn such a problem
    into finite garbage: the problem gets IDOC_STATUS_NONFINITE, its outputs are NaN, its neighbours are untouched, and
    the reference-API call warns (tuned, mid-size and generic kernels)."""
    from oracle import lqr_np as O
    n, m, T, B = shape
    rng = np.random.default_rng(21)
    p = O.random_params(rng, n, m, T, batch=(B,))
    so = O.direct(p)
    bad = 2
    getattr(p.lqr, where)[bad, T // 2, 0] = np.nan
    pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
    with pytest.warns(ib._lib.IdocNumericalWarning):
        s = ib.lqr.build(T).direct(pp)
    assert np.isnan(s.X[bad]).all() and np.isnan(s.U[bad]).all() and np.isnan(s.Nu[bad]).all()
    ok = np.arange(B) != bad
    assert rel(s.X[ok], so.X[ok]) < 1e-10 and rel(s.U[ok], so.U[ok]) < 1e-10 and rel(s.Nu[ok], so.Nu[ok]) < 1e-10
    dp = ib.lqr.DeviceProblem.from_params(pp)
    plan = ib.lqr.DevicePlan(B, T, n, m, np.float64, vjp=False)
    plan.fwd(dp)
    st = plan.status.numpy()
    assert (st[bad] & ib._lib.STATUS_NONFINITE) and not (st[ok] & ib._lib.STATUS_NONFINITE).any()


@pytest.mark.parametrize("shape", [(16, 8, 6, 3), (5, 2, 6, 4)])
def test_tilqr_failed_factorisation_is_reported(ib, shape):
    """tilqr factors R + B'PB without a shift (sym_pos=True, idoc/tilqr.py:72): an indefinite R makes the reference's
    Cholesky solve fail (NaN).  The device path flags the problem (CLAMP | NONFINITE), returns NaN for it and warns."""
    n, m, T, B = shape
    rng = np.random.default_rng(3)
    A = np.stack([0.5 * np.linalg.qr(rng.standard_normal((n, n)))[0] for _ in range(B)])
    R = np.broadcast_to(np.eye(m), (B, m, m)).copy()
    R[1] = -50.0 * np.eye(m)  # indefinite: R + B'PB has negative pivots
    th = ib.tilqr.Params(rng.standard_normal((B, n)),
                         ib.tilqr.TILQR(Q=np.broadcast_to(np.eye(n), (B, n, n)).copy(), Qf=np.broadcast_to(np.eye(n), (B, n, n)).copy(),
                                        R=R, A=A, B=0.1 * rng.standard_normal((B, n, m))))
    with pytest.warns(ib._lib.IdocNumericalWarning):
        s = ib.tilqr.build(T).direct(th)
    assert np.isnan(s.X[1]).all() and np.isnan(s.U[1]).all() and np.isnan(s.Nu[1]).all()
    assert np.isfinite(s.X[[0, 2]]).all()
