"""CPU tests of the oracle's autodiff families (oracle/ad_np.py, oracle/ilqr_np.py: Pendulum, Cartpole, Unicycle): jets against
central differences, the iLQR loop reaches a KKT point, the exact implicit VJP against finite differences through the
converged solve, and fp32 mode of the LQR restatement (the rounding budget used by the fp32 parity tests)."""
import numpy as np
import pytest

from oracle import ad_np as AD
from oracle import ilqr_np as IO
from oracle import lqr_np as O

CASES = {
    "pendulum": (IO.Pendulum, np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.01, 100.0, 10.0]), np.array([0.3, -0.1])),
    "cartpole": (IO.Cartpole, np.array([0.02, 1.0, 0.1, 0.5, 9.81, 0.1, 1.0, 0.1, 0.1, 0.01, 10.0, 100.0, 10.0, 10.0]),
                 np.array([0.05, np.pi + 0.15, 0.02, -0.03])),
    "unicycle": (IO.Unicycle, np.array([0.1, 1.0, 0.5, 1.0, 0.1, 0.1, 0.1, 10.0, 1.0]), np.array([0.2, -0.1, 0.3])),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_jets_match_central_differences(name):
    P, th, x = CASES[name]
    rng = np.random.default_rng(0)
    u = 0.3 * rng.standard_normal(P.m)
    lx, lu, lxx, luu, lxu, fx, fu = P.derivs(x, u, th)
    h = 1e-6
    z = np.concatenate((x, u))
    grad = np.concatenate((lx, lu))
    J = np.concatenate((fx, fu), axis=1)
    H = np.block([[lxx, lxu], [lxu.T, luu]])
    for i in range(z.size):
        e = np.zeros(z.size); e[i] = h
        zp, zm = z + e, z - e
        cp, cm = P.cost(zp[:P.n], zp[P.n:], th), P.cost(zm[:P.n], zm[P.n:], th)
        assert abs((cp - cm) / (2 * h) - grad[i]) < 1e-6 * max(1.0, abs(grad[i]))
        fp_, fm_ = P.dynamics(zp[:P.n], zp[P.n:], th), P.dynamics(zm[:P.n], zm[P.n:], th)
        assert np.allclose((fp_ - fm_) / (2 * h), J[:, i], atol=1e-6)
        gp = np.concatenate(P.derivs(zp[:P.n], zp[P.n:], th)[:2])
        gm = np.concatenate(P.derivs(zm[:P.n], zm[P.n:], th)[:2])
        assert np.allclose((gp - gm) / (2 * h), H[:, i], atol=1e-5)
    assert np.allclose(H, H.T)


@pytest.mark.parametrize("name,T", [("pendulum", 30), ("unicycle", 20)])
def test_ad_family_ilqr_kkt_and_exact_vjp(name, T):
    P, th, x0 = CASES[name]
    U0 = np.zeros((T, P.m))
    solve = lambda th_, x0_: IO.ilqr(P, th_, x0_, IO.simulate(P, th_, x0_, U0)[0], U0, maxiter=300, thres=1e-14, line_search=True)[0]
    s = solve(th, x0)
    r = IO.kkt(P, th, x0, s)
    assert max(np.abs(z).max() for z in r) < 1e-7
    gx0, gth = IO.vjp_exact_ad(P, th, x0, s, O.State(s.X, s.U, np.zeros_like(s.Nu)))
    loss = lambda z: 0.5 * np.sum(z.X ** 2) + 0.5 * np.sum(z.U ** 2)
    for e in range(P.theta_dim):
        d = np.zeros(P.theta_dim); d[e] = 1e-5 * max(abs(th[e]), 1e-2)
        fd = (loss(solve(th + d, x0)) - loss(solve(th - d, x0))) / (2 * d[e])
        assert abs(gth[e] - fd) < 2e-4 * max(abs(fd), 1e-3 * np.abs(gth).max()), (e, gth[e], fd)
    for e in range(P.n):
        d = np.zeros(P.n); d[e] = 1e-6
        fd = (loss(solve(th, x0 + d)) - loss(solve(th, x0 - d))) / 2e-6
        assert abs(gx0[e] - fd) < 2e-4 * max(abs(fd), 1e-3 * np.abs(gx0).max()), (e, gx0[e], fd)


def test_fp32_mode_of_the_lqr_restatement_gives_a_rounding_budget():
    rng = np.random.default_rng(1)
    p = O.random_params(rng, 8, 4, 40, batch=(3,), dtype=np.float32)
    s64 = O.direct(p)
    with O.compute_dtype(np.float32):
        s32 = O.direct(p)
    assert s32.X.dtype == np.float32 and s64.X.dtype == np.float64
    err = np.max(np.abs(s32.X - s64.X)) / np.max(np.abs(s64.X))
    assert 1e-9 < err < 1e-4
    assert O.direct(p).X.dtype == np.float64  # the context manager restores fp64
