"""GPU parity tests for the shared-control batch iLQR (idoc_b200.bilqr) against golden vectors produced by running
the reference's bilqr.py (its own test size batch_size=1, n=m=10, T=20, and a 3-system case with the exact VJP)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bilqr_*.npz")))
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))
FIELDS = ("Q", "q", "Qf", "R", "r", "A", "B")


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_bilqr_matches_reference_golden(path):
    import idoc_b200 as ib
    ib._lib.require_device()
    z = np.load(path)
    T, bs, n = z["X"].shape
    m = z["U"].shape[1]
    cs = ib.bilqr.Problem(family="relu_ax", horizon=T, state_dim=n, control_dim=m)
    solver = ib.bilqr.build(cs, maxiter=int(z["maxiter"]), thres=1e-8, line_search=ib.make_line_search())
    params = ib.bilqr.Params(z["x0"], tuple(z["th_" + k] for k in FIELDS))
    s, pull = solver.vjp(ib.typs.State(z["X0"], z["U0"], np.zeros_like(z["X0"])), params)
    assert s.X.shape == z["X"].shape
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    ib.utils.check_kkt(solver.kkt, s, params, thres=1e-9)                     # tests/test_bilqr.py:121
    r = solver.kkt(ib.typs.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), params)
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10 and rel(r.Nu, z["kkt_Nu"]) < 1e-10
    if "g_x0" in z.files:
        g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
        assert rel(g.x0, z["g_x0"]) < 1e-10
        for k, got in zip(FIELDS, g.theta):
            assert rel(got, z["g_" + k]) < 1e-10, k


# ------------------------------------------------------------------ fused shared-control kernels (csrc/bilqr_fused.cu)
def _pendulum_case(bs, T, seed):
    rng = np.random.default_rng(seed)
    theta = np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.01, 100.0, 10.0])
    x0 = np.stack([np.pi + 0.4 * rng.standard_normal(bs), 0.2 * rng.standard_normal(bs)], axis=1)
    return theta, x0


def _relu_theta(rng, n, m):
    sym = lambda z: z @ z.T / z.shape[0] + np.eye(z.shape[0])
    return (sym(rng.standard_normal((n, n))), 0.1 * rng.standard_normal(n), sym(rng.standard_normal((n, n))),
            sym(rng.standard_normal((m, m))), 0.1 * rng.standard_normal(m), 0.5 * rng.standard_normal((n, n)) / np.sqrt(n),
            rng.standard_normal((n, m)))


def _lift_cases():
    rng = np.random.default_rng(5)
    th, x0 = _pendulum_case(2, 30, 1)
    yield "pendulum_x2", "pendulum", 2, 1, 30, th, x0
    th, x0 = _pendulum_case(4, 30, 2)
    yield "pendulum_x4", "pendulum", 2, 1, 30, th, x0
    th = np.array([0.02, 1.0, 0.1, 0.5, 9.81, 0.1, 1.0, 0.1, 0.1, 0.01, 10.0, 100.0, 10.0, 10.0])
    x0 = np.stack([0.1 * rng.standard_normal(2), np.pi + 0.2 * rng.standard_normal(2), 0.1 * rng.standard_normal(2),
                   0.1 * rng.standard_normal(2)], -1)
    yield "cartpole_x2", "cartpole", 4, 1, 25, th, x0
    yield "relu_rnn_3_1_x2", "relu_rnn", 3, 1, 12, _relu_theta(rng, 3, 1), rng.standard_normal((2, 3))
    yield "relu_ax_3_2_x2", "relu_ax", 3, 2, 12, _relu_theta(rng, 3, 2), rng.standard_normal((2, 3))


LIFTS = list(_lift_cases())


def _flat(theta):
    return np.concatenate([np.ravel(z) for z in theta]) if isinstance(theta, tuple) else np.asarray(theta)


@pytest.mark.parametrize("case", LIFTS, ids=[c[0] for c in LIFTS])
def test_fused_shared_control_matches_multikernel_loop(case, monkeypatch):
    """One-kernel bilqr (SharedControl lift inside csrc/ilqr_fused.cuh) against the multi-kernel loop that the reference
    goldens above pin (IDOC_ILQR_FUSED=0): same iterates and iteration counts, same co-states, same implicit gradients."""
    import idoc_b200 as ib
    ib._lib.require_device()
    _, family, n, m, T, theta, x0 = case
    bs = x0.shape[0]
    cs = ib.bilqr.Problem(family=family, horizon=T, state_dim=n, control_dim=m)
    solver = ib.bilqr.build(cs, maxiter=60, thres=1e-10, line_search=ib.make_line_search())
    init = ib.typs.State(np.repeat(x0[None], T, axis=0), np.zeros((T, m)), np.zeros((T, bs, n)))
    params = ib.bilqr.Params(x0, theta)
    rng = np.random.default_rng(3)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("IDOC_ILQR_FUSED", fused)
        s, info = solver.direct(init, params, return_info=True)
        _, pull = solver.vjp(init, params)
        out[fused] = (s, info, [pull(ib.typs.State(w[0], w[1], w[2])) for w in
                                [(s.X, s.U, None), (rng.standard_normal(s.X.shape), rng.standard_normal(s.U.shape),
                                                    rng.standard_normal(s.Nu.shape))]])
        rng = np.random.default_rng(3)
    (s1, i1, g1), (s0, i0, g0) = out["1"], out["0"]
    assert int(i1["iterations"]) == int(i0["iterations"]) and int(i1["iterations"]) > 1
    assert rel(s1.X, s0.X) < 1e-10 and rel(s1.U, s0.U) < 1e-10 and rel(s1.Nu, s0.Nu) < 1e-10
    r1, r0 = solver.kkt(s1, params), solver.kkt(s0, params)
    assert max(np.abs(r1.X).max(), np.abs(r1.Nu).max()) < 1e-9 and np.abs(r1.U - r0.U).max() <= 1e-9 * max(1.0, np.abs(r0.U).max())
    for a, b in zip(g1, g0):
        assert rel(a.x0, b.x0) < 1e-9 and rel(_flat(a.theta), _flat(b.theta)) < 1e-9


def _pend_cost(theta, x0, U):
    """total cost of the shared-control pendulum problem (csrc/ilqr.cuh Pendulum; summed over the systems, bilqr.py:139-143)"""
    h, gl, b, c, w1, w2, wu, f1, f2 = theta
    x = x0.copy()
    tot = 0.0
    for u in U[:, 0]:
        tot += 0.5 * np.sum(w1 * (x[:, 0] - np.pi) ** 2 + w2 * x[:, 1] ** 2 + wu * u * u)
        x = np.stack([x[:, 0] + h * x[:, 1], x[:, 1] + h * (-gl * np.sin(x[:, 0]) - b * x[:, 1] + c * u)], axis=1)
    return tot + 0.5 * np.sum(f1 * (x[:, 0] - np.pi) ** 2 + f2 * x[:, 1] ** 2)


@pytest.mark.parametrize("bs", [2, 4])
def test_control_limited_bilqr_pendulum(bs):
    """Control limits on the SHARED control of bilqr (no reference counterpart; where north_star places the box step):
    limits respected and binding, wide limits = the unlimited solver, KKT with the sign conditions of the active bounds,
    the optimum of a bounded quasi-Newton solve of the same problem, implicit gradient vs central differences."""
    from scipy.optimize import minimize
    import idoc_b200 as ib
    ib._lib.require_device()
    T, umax = 40, (3.0 if bs == 2 else 4.0)
    rng = np.random.default_rng(20 + bs)
    # one control for several pendulums: a compromise, so the control cost is raised until the optimum is not bang-bang
    # (L-BFGS-B on this problem: 10 of 40 steps on a bound with 2 systems, 15 with 4)
    theta = np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.2, 20.0, 2.0])
    x0 = np.stack([np.pi + 0.3 + 0.1 * rng.standard_normal(bs), 0.2 * rng.standard_normal(bs)], axis=1)
    cs = ib.bilqr.Problem(family="pendulum", horizon=T, state_dim=2, control_dim=1)
    # a fixed number of full steps (thres < 0 never stops early): the reference's stopping rule watches the COST, which stalls at
    # fp64 resolution while the controls are still 1e-7 from stationarity -- too loose for finite differences of the solver
    kw = dict(maxiter=80, thres=-1.0, line_search=None)
    boxed = ib.bilqr.build(cs, control_limits=(-umax, umax), **kw)
    wide = ib.bilqr.build(cs, control_limits=(-1e9, 1e9), **kw)
    plain = ib.bilqr.build(cs, **kw)

    def init(th):
        X0 = np.zeros((T, bs, 2))
        x = x0.copy()
        for t in range(T):
            x = np.stack([x[:, 0] + th[0] * x[:, 1], x[:, 1] + th[0] * (-th[1] * np.sin(x[:, 0]) - th[2] * x[:, 1])], axis=1)
            X0[t] = x
        return ib.typs.State(X0, np.zeros((T, 1)), np.zeros((T, bs, 2)))

    params = ib.bilqr.Params(x0, theta)
    sw, sp = wide.direct(init(theta), params), plain.direct(init(theta), params)
    assert rel(sw.X, sp.X) < 1e-12 and rel(sw.U, sp.U) < 1e-12 and rel(sw.Nu, sp.Nu) < 1e-12
    assert np.abs(sp.U).max() > umax  # the limits bind
    s, pull = boxed.vjp(init(theta), params)
    assert s.X.shape == (T, bs, 2) and s.U.shape == (T, 1)
    assert s.U.min() >= -umax and s.U.max() <= umax
    at_lo, at_hi = s.U <= -umax, s.U >= umax
    free = ~(at_lo | at_hi)
    assert (at_lo | at_hi).sum() >= 2 and free.sum() >= T // 2
    r = boxed.kkt(s, params)
    assert np.abs(r.X).max() < 1e-8 and np.abs(r.Nu).max() < 1e-8 and np.abs(r.U[free]).max() < 1e-8
    assert (r.U[at_lo] >= -1e-9).all() and (r.U[at_hi] <= 1e-9).all()
    # independent optimiser on the same bounded problem (non-convex: started from the device's solution, it must stay there
    # and find no lower cost; started from U = 0 it may end in another basin, never in a better one than 1e-6 below)
    opts = dict(maxiter=20000, maxfun=400000, ftol=1e-16, gtol=1e-10)
    f = lambda z: _pend_cost(theta, x0, z.reshape(T, 1))
    c_dev = f(s.U.ravel())
    res = minimize(f, s.U.ravel(), method="L-BFGS-B", bounds=[(-umax, umax)] * T, options=opts)
    assert res.fun >= c_dev * (1 - 1e-9) and np.abs(res.x - s.U.ravel()).max() < 1e-4, (c_dev, res.fun, np.abs(res.x - s.U.ravel()).max())
    cold = minimize(f, np.zeros(T), method="L-BFGS-B", bounds=[(-umax, umax)] * T, options=opts)
    if np.abs(cold.x - s.U.ravel()).max() < 1e-2:  # same basin: same optimum
        assert abs(c_dev - cold.fun) <= 1e-7 * abs(cold.fun), (c_dev, cold.fun)

    def loss(th):
        z = boxed.direct(init(th), ib.bilqr.Params(x0, th))
        return 0.5 * (z.X ** 2).sum() + 0.5 * (z.U ** 2).sum()

    g = pull(ib.typs.State(s.X, s.U, None))
    gth = np.asarray(g.theta)
    for e in range(9):
        d = np.zeros_like(theta)
        d[e] = 1e-5 * abs(theta[e])
        fd = (loss(theta + d) - loss(theta - d)) / (2 * d[e])
        assert np.isclose(gth[e], fd, rtol=1e-5, atol=1e-7 * np.abs(gth).max()), (e, gth[e], fd)
    for i in range(bs):  # x0 gradient of one system
        d = np.zeros_like(x0)
        d[i, 0] = 1e-5
        lp = 0.5 * sum((z ** 2).sum() for z in boxed.direct(init(theta), ib.bilqr.Params(x0 + d, theta))[:2])
        lm = 0.5 * sum((z ** 2).sum() for z in boxed.direct(init(theta), ib.bilqr.Params(x0 - d, theta))[:2])
        assert np.isclose(g.x0[i, 0], (lp - lm) / 2e-5, rtol=1e-5, atol=1e-7 * np.abs(g.x0).max()), (i, g.x0[i, 0], (lp - lm) / 2e-5)


def test_control_limited_bilqr_needs_a_compiled_lift():
    import idoc_b200 as ib
    ib._lib.require_device()
    cs = ib.bilqr.Problem(family="relu_ax", horizon=5, state_dim=10, control_dim=10)
    solver = ib.bilqr.build(cs, control_limits=(-1.0, 1.0))
    with pytest.raises(ib._lib.IdocError, match="no fused shared-control kernel"):
        solver.direct(ib.typs.State(np.zeros((5, 2, 10)), np.zeros((5, 10)), np.zeros((5, 2, 10))),
                      ib.bilqr.Params(np.zeros((2, 10)), tuple(np.zeros(s) for s in [(10, 10), (10,), (10, 10), (10, 10), (10,), (10, 10), (10, 10)])))
