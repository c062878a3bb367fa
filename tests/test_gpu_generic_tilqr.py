"""GPU parity tests for the shape-generic kernels (any n, m <= 32) and the time-invariant solver (tilqr)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = {np.float64: 1e-10, np.float32: 1e-4}  # north_star's bars, every output and every gradient leaf
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ib():
    import idoc_b200
    idoc_b200._lib.require_device()
    return idoc_b200


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(7, 5, 6, 5), (10, 3, 30, 3), (16, 8, 12, 6), (1, 1, 4, 3), (32, 16, 5, 2), (10, 10, 20, 1),
                                   (90, 2, 5, 2), (64, 8, 4, 1), (8, 8, 7, 5), (16, 16, 6, 3), (24, 8, 6, 3), (32, 8, 5, 3)])
def test_generic_lqr_vs_oracle(ib, dtype, shape):
    from oracle import lqr_np as O
    n, m, T, B = shape
    assert ib._lib.lib.idoc_lqr_supported(1, n, m) == 2
    rng = np.random.default_rng(7 + n + m)
    p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
    so, go = O.direct(p, return_gains=True)
    pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
    s, pull = ib.lqr.build(T).vjp(pp)
    tol = TOL[dtype]
    assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol
    gains = ib.lqr.backward(pp.lqr, T)
    assert rel(gains.K, go.K) < tol and rel(gains.k, go.k) < tol
    for nub in (True, False):
        w = O.State(rng.standard_normal(so.X.shape).astype(dtype), rng.standard_normal(so.U.shape).astype(dtype),
                    (rng.standard_normal(so.Nu.shape) if nub else np.zeros(so.Nu.shape)).astype(dtype))
        g_o = O.vjp_exact(p, so, w)
        g = pull(ib.typs.State(w.X, w.U, w.Nu if nub else None))
        assert rel(g.x0, g_o.x0) < tol
        for name in ib.lqr.LQR._fields:
            assert rel(getattr(g.lqr, name), getattr(g_o.lqr, name)) < tol, (name, nub)


def test_generic_equals_tuned_on_8x4(ib, monkeypatch):
    from oracle import lqr_np as O
    rng = np.random.default_rng(3)
    p = O.random_params(rng, 8, 4, 9, batch=(11,))
    pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
    s1 = ib.lqr.build(9).direct(pp)
    monkeypatch.setenv("IDOC_FORCE_GENERIC", "1")
    assert ib._lib.lib.idoc_lqr_supported(1, 8, 4) == 2
    s2 = ib.lqr.build(9).direct(pp)
    assert rel(s2.X, s1.X) < 1e-12 and rel(s2.Nu, s1.Nu) < 1e-12


def test_generic_shift_golden(ib, monkeypatch):
    monkeypatch.setenv("IDOC_FORCE_GENERIC", "1")
    for name in ("lqr_shift_n4m2T10", "lqr_shift_n3m3T6"):
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        p = ib.lqr.Params(z["x0"], ib.lqr.LQR(**{k: z[k] for k in ib.lqr.LQR._fields}))
        s = ib.lqr.build(z["A"].shape[0]).direct(p)
        assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10


TCASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tilqr_*.npz")))


@pytest.mark.parametrize("path", TCASES, ids=[os.path.basename(p)[:-4] for p in TCASES])
def test_tilqr_matches_reference_golden(ib, path):
    z = np.load(path)
    T = z["X"].shape[0]
    th = ib.tilqr.Params(z["x0"], ib.tilqr.TILQR(Q=z["Q"], Qf=z["Qf"], R=z["R"], A=z["A"], B=z["B"]))
    solver = ib.tilqr.build(T)
    s, pull = solver.vjp(th)
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    ib.utils.check_kkt(solver.kkt, s, th, thres=1e-10)   # tests/test_tilqr.py:50
    r = solver.kkt(ib.typs.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), th)
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10
    g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
    assert rel(g.x0, z["g_x0"]) < 1e-10
    for name in ib.tilqr.TILQR._fields:
        assert rel(getattr(g.lqr, name), z["g_" + name]) < 1e-10, name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tilqr_config3_shape_vs_oracle(ib, dtype):
    """BASELINE config 3 shape (n=16, m=8, T=200), B = 64, both dtypes, solution and every gradient leaf at north_star's bar."""
    from oracle import tilqr_np as TO
    from oracle import lqr_np as O
    n, m, T, B = 16, 8, 200, 64
    rng = np.random.default_rng(16)
    W = rng.standard_normal((B, n, n))
    A = np.stack([O.init_stable(rng, n) for _ in range(B)])
    th = TO.Params(rng.standard_normal((B, n)),
                   TO.TILQR(Q=np.eye(n) + 0.2 * W @ np.swapaxes(W, -1, -2) / n, Qf=np.broadcast_to(np.eye(n), (B, n, n)).copy(),
                            R=np.broadcast_to(0.01 * np.eye(m), (B, m, m)).copy(), A=A, B=rng.standard_normal((B, n, m))))
    th = TO.Params(th.x0.astype(dtype), TO.TILQR(*[z.astype(dtype) for z in th.lqr]))
    so = TO.direct(th, T)
    s, pull = ib.tilqr.build(T).vjp(ib.tilqr.Params(th.x0, ib.tilqr.TILQR(*th.lqr)))
    tol = TOL[dtype]
    assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol
    w = TO.State(so.X.astype(dtype), so.U.astype(dtype), np.zeros_like(so.Nu).astype(dtype))
    g = pull(ib.typs.State(w.X, w.U, None))
    g_o = TO.vjp_exact(th, so, w, T)
    assert rel(g.x0, g_o.x0) < tol
    for name in ib.tilqr.TILQR._fields:
        want = getattr(g_o.lqr, name)
        if np.max(np.abs(want)) < 1e-30:  # Qf: the state has decayed to ~1e-70 by t = 200 (underflows in fp32)
            assert np.max(np.abs(getattr(g.lqr, name))) < 1e-30
            continue
        assert rel(getattr(g.lqr, name), want) < tol, name


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-4)])
@pytest.mark.parametrize("shape", [(16, 8, 9, 7), (32, 16, 6, 3), (8, 8, 6, 5), (16, 16, 5, 3), (24, 8, 5, 2), (32, 8, 5, 3)])
def test_mid_riccati_equals_generic_and_handles_shift(ib, dtype, tol, shape, monkeypatch):
    """The warp-per-problem register-tiled Riccati sweep (csrc/mid.cuh, the default at the shapes of csrc/mid_shapes.def) against the
    shape-generic kernel (IDOC_NO_MID=1) on the same problems, including eigen-shifted ones (singular R, M = 0)."""
    from oracle import lqr_np as O
    n, m, T, B = shape
    rng = np.random.default_rng(n + m)
    p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
    lq = p.lqr._asdict() if hasattr(p.lqr, "_asdict") else dict(zip(ib.lqr.LEAVES, p.lqr))
    # problem 0: rank-deficient R and B -> Guu singular at every step -> the eigen-shift path
    lq["R"][0] = 0.0
    lq["M"][0] = 0.0
    lq["B"][0, :, :, 0] = 0.0
    params = ib.lqr.Params(p.x0, ib.lqr.LQR(**lq))
    out = {}
    for no_mid in ("0", "1"):
        monkeypatch.setenv("IDOC_NO_MID", no_mid)
        gains, ec = ib.lqr.backward(params.lqr, T, return_expected_change=True)
        s, pull = ib.lqr.build(T).vjp(params)
        grads = [pull(ib.typs.State(s.X, s.U, None)), pull(ib.typs.State(s.X, s.U, s.Nu))]
        out[no_mid] = (gains, ec(1.0), s, grads)
    (g1, e1, s1, gr1), (g0, e0, s0, gr0) = out["0"], out["1"]
    for ga, gb in zip(gr1, gr0):  # rollout + both adjoint sweeps, with and without a Nu cotangent
        assert rel(ga.x0[1:], gb.x0[1:]) < tol
        for k in ib.lqr.LEAVES:
            assert rel(getattr(ga.lqr, k)[1:], getattr(gb.lqr, k)[1:]) < tol, k
    assert rel(g1.K[1:], g0.K[1:]) < tol and rel(g1.k[1:], g0.k[1:]) < tol
    assert rel(s1.X[1:], s0.X[1:]) < tol and rel(s1.U[1:], s0.U[1:]) < tol and rel(s1.Nu[1:], s0.Nu[1:]) < tol
    assert rel(e1[1:], e0[1:]) < tol
    if dtype == np.float64:  # the shifted problem: same regularised gains (conditioning 1e8: looser)
        assert np.isfinite(g1.K[0]).all() and rel(g1.K[0], g0.K[0]) < 1e-4


def _mma_mode(monkeypatch, mode):
    """"1": tensor-core sweep with its defaults (two problems per warp at n = 16); "g1": the other geometry -- one problem per
    warp at n = 16, the other warps-per-problem setting at n = 32 (IDOC_MID_MMA_NW 1 <-> 2); "0": the register-tile sweep"""
    monkeypatch.setenv("IDOC_MID_MMA", "0" if mode == "0" else "1")
    if mode == "g1":
        monkeypatch.setenv("IDOC_MID_MMA_G", "1")
        monkeypatch.setenv("IDOC_MID_MMA_NW", os.environ.get("IDOC_TEST_OTHER_NW", "2"))
    else:
        monkeypatch.delenv("IDOC_MID_MMA_G", raising=False)
        monkeypatch.delenv("IDOC_MID_MMA_NW", raising=False)


@pytest.mark.parametrize("shape", [(32, 16, 40, 3), (32, 8, 12, 2), (16, 8, 50, 5), (16, 16, 9, 2)])
@pytest.mark.parametrize("ti", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_tensor_core_riccati_sweep_vs_oracle(ib, shape, ti, dtype, monkeypatch):
    """The tensor-core Riccati sweep (csrc/mid_mma.cuh, n in {16, 32}: 3xTF32 mma.sync in fp32, DMMA m8n8k4 in fp64) against
    the fp64 oracle at north_star's bars (1e-4 / 1e-10), for the time-varying solver (gains, solution, every gradient leaf,
    eigen-shifted problem included) and for tilqr, with two problems per warp and with one; and against the register-tile
    sweep (IDOC_MID_MMA=0) it replaces."""
    from oracle import lqr_np as O
    from oracle import tilqr_np as TO
    n, m, T, B = shape
    rng = np.random.default_rng(100 + n + m)
    f32 = dtype == np.float32
    tol, same = (1e-4, 1e-5) if f32 else (1e-10, 1e-11)
    out = {}
    if not ti:
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
        so, go = O.direct(p, return_gains=True)
        w = O.State(rng.standard_normal(so.X.shape).astype(dtype), rng.standard_normal(so.U.shape).astype(dtype),
                    rng.standard_normal(so.Nu.shape).astype(dtype))
        g_o = O.vjp_exact(p, so, w)
        pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
        for mma in ("1", "g1", "0"):
            _mma_mode(monkeypatch, mma)
            s, pull = ib.lqr.build(T).vjp(pp)
            gains = ib.lqr.backward(pp.lqr, T)
            g = pull(ib.typs.State(w.X, w.U, w.Nu))
            assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol, mma
            assert rel(gains.K, go.K) < tol and rel(gains.k, go.k) < tol, mma
            for name in ib.lqr.LQR._fields:
                assert rel(getattr(g.lqr, name), getattr(g_o.lqr, name)) < tol, (name, mma)
            out[mma] = s
        assert rel(out["1"].X, out["0"].X) < same and rel(out["g1"].X, out["0"].X) < same
        # eigen-shift path (singular Guu): same regularised solve as the FFMA sweep
        lq = dict(zip(ib.lqr.LEAVES, [z.copy() for z in p.lqr]))
        lq["R"][0] = 0.0
        lq["M"][0] = 0.0
        lq["B"][0, :, :, 0] = 0.0
        sh = {}
        for mma in ("1", "g1", "0"):
            _mma_mode(monkeypatch, mma)
            dp = ib.lqr.DeviceProblem.from_params(ib.lqr.Params(p.x0, ib.lqr.LQR(**lq)))
            plan = ib.lqr.DevicePlan(B, T, n, m, dtype, gains=True, vjp=False)
            plan.fwd(dp)
            sh[mma] = (plan.status.numpy().copy(), plan.X.numpy().copy())
        for mma in ("1", "g1"):
            assert (sh[mma][0][0] & 1) and (sh["0"][0][0] & 1) and np.isfinite(sh[mma][1]).all()
            assert np.array_equal(sh[mma][0], sh["0"][0]) and rel(sh[mma][1][1:], sh["0"][1][1:]) < same
    else:
        W = rng.standard_normal((B, n, n))
        A = np.stack([O.init_stable(rng, n) for _ in range(B)])
        th = TO.Params(rng.standard_normal((B, n)).astype(dtype),
                       TO.TILQR(*[z.astype(dtype) for z in (np.eye(n) + 0.2 * W @ np.swapaxes(W, -1, -2) / n,
                                                                   np.broadcast_to(np.eye(n), (B, n, n)).copy(),
                                                                   np.broadcast_to(0.01 * np.eye(m), (B, m, m)).copy(), A,
                                                                   rng.standard_normal((B, n, m)))]))
        so = TO.direct(th, T)
        w = TO.State(so.X, so.U, np.zeros_like(so.Nu))
        g_o = TO.vjp_exact(th, so, w, T)
        # fp32 error budget of the reference ALGORITHM on these inputs (R = 0.01 I against B'PB ~ 10: the gradient of R loses
        # three digits to cancellation): the oracle itself executed in float32 against its float64 run.  The device result
        # must be within north_star's 1e-4 or, where the reference arithmetic itself cannot do better, within 4x that budget.
        budget = dict(Q=0.0, R=0.0, A=0.0, B=0.0)
        if f32:
            with O.compute_dtype(np.float32):
                s32 = TO.direct(th, T)
                g32 = TO.vjp_exact(th, s32, w, T)
            budget = {name: rel(getattr(g32.lqr, name), getattr(g_o.lqr, name)) for name in ("Q", "R", "A", "B")}
        for mma in ("1", "g1", "0"):
            _mma_mode(monkeypatch, mma)
            s, pull = ib.tilqr.build(T).vjp(ib.tilqr.Params(th.x0, ib.tilqr.TILQR(*th.lqr)))
            g = pull(ib.typs.State(so.X.astype(dtype), so.U.astype(dtype), None))
            assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol, mma
            for name in ("Q", "R", "A", "B"):
                assert rel(getattr(g.lqr, name), getattr(g_o.lqr, name)) < max(tol, 4 * budget[name]), (name, mma, budget)
