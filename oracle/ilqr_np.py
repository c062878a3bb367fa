"""CPU ORACLE (test infrastructure, not product code) — NumPy fp64 restatement of the reference's iLQR outer
loop, `/root/reference/idoc/ilqr.py:127-268` and `idoc/line_search.py:7-42`, for problem families whose
derivatives are written out by hand (the reference obtains them with jax.grad / jacfwd).  Pinned against golden
vectors produced by executing the unmodified reference sources with forward-mode duals standing in for JAX
autodiff (tests/golden/make_golden.py: ilqr_*.npz).  Single problem per call (loop over a batch in the tests)."""
from typing import NamedTuple

import numpy as np

from . import lqr_np as O


class ReluRnn:
    """tests/test_ilqr.py:14-64: f = A relu(x) + B u + 0.5; l = 1/2 x'Qx + q'x + 1e-4 u'Ru - 1e-4 (1'u)(1'x) + 1e-4 r'u."""
    FIELDS = ("Q", "q", "Qf", "R", "r", "A", "B")

    class Theta(NamedTuple):
        Q: np.ndarray
        q: np.ndarray
        Qf: np.ndarray
        R: np.ndarray
        r: np.ndarray
        A: np.ndarray
        B: np.ndarray

    @staticmethod
    def dynamics(x, u, th):
        return th.A @ np.maximum(x, 0.0) + th.B @ u + 0.5

    @staticmethod
    def cost(x, u, th):
        return (0.5 * x @ th.Q @ x + th.q @ x + 1e-4 * u @ th.R @ u - 1e-4 * np.sum(u) * np.sum(x) + 1e-4 * th.r @ u)

    @staticmethod
    def costf(x, th):
        return 0.5 * x @ th.Qf @ x

    @staticmethod
    def derivs(x, u, th):
        n, m = x.size, u.size
        lxx = 0.5 * (th.Q + th.Q.T)
        luu = 1e-4 * (th.R + th.R.T)
        lxu = -1e-4 * np.ones((n, m))
        lx = lxx @ x + th.q - 1e-4 * np.sum(u)
        lu = luu @ u - 1e-4 * np.sum(x) + 1e-4 * th.r
        fx = th.A * (x > 0.0)[None, :]
        return lx, lu, lxx, luu, lxu, fx, th.B

    @staticmethod
    def final_derivs(x, th):
        H = 0.5 * (th.Qf + th.Qf.T)
        return H @ x, H


def make_lqr_approx(P, th, x0, T, local=True):
    """ilqr.py:127-158."""
    def approx(X, U):
        sX = np.concatenate((x0[None], X[:-1]))
        leaves = {k: [] for k in ("Q", "q", "R", "r", "M", "A", "B", "d")}
        for t in range(T):
            x, u = sX[t], U[t]
            lx, lu, lxx, luu, lxu, fx, fu = P.derivs(x, u, th)
            if not local:  # ilqr.py:139-142
                q = lx - lxx @ x - lxu @ u
                r = lu - luu @ u - lxu.T @ x
                d = P.dynamics(x, u, th) - fx @ x - fu @ u
            else:          # ilqr.py:143-144
                q, r, d = lx, lu, np.zeros(x.size)
            for k, v in zip(("Q", "q", "R", "r", "M", "A", "B", "d"), (lxx, q, luu, r, lxu, fx, fu, d)):
                leaves[k].append(v)
        xf = X[-1]
        qf, Qf = P.final_derivs(xf, th)
        if not local:
            qf = qf - Qf @ xf
        L = {k: np.stack(v) for k, v in leaves.items()}
        return O.LQR(Q=L["Q"], q=L["q"], Qf=Qf, qf=qf, M=L["M"], R=L["R"], r=L["r"], A=L["A"], B=L["B"], d=L["d"])
    return approx


def simulate(P, th, x0, U):
    """ilqr.py:161-176."""
    x, c, X = x0, 0.0, []
    for t in range(U.shape[0]):
        c = c + P.cost(x, U[t], th)
        x = P.dynamics(x, U[t], th)
        X.append(x)
    return np.stack(X), c + P.costf(x, th)


def update(P, th, x0, X, U, gains, alpha):
    """ilqr.py:196-219."""
    sX = np.concatenate((x0[None], X[:-1]))
    xh, l, Xn, Un = x0, 0.0, [], []
    for t in range(U.shape[0]):
        du = gains.K[t] @ (xh - sX[t]) + alpha * gains.k[t]
        uh = U[t] + du
        l = l + P.cost(xh, uh, th)
        xh = P.dynamics(xh, uh, th)
        Xn.append(xh)
        Un.append(uh)
    return np.stack(Xn), np.stack(Un), l + P.costf(xh, th)


def backtrack(f, c_initial, dC, dc, lower=1e-3, upper=2.0, rho=0.5, maxiter=8):
    """line_search.py:9-40 (returns the last tried point even when every trial fails)."""
    alpha, carry_on, out = 1.0, True, None
    for _ in range(maxiter):
        if not carry_on:
            break
        X, U, c = f(alpha)
        ec = alpha ** 2 * dC + alpha * dc
        p = abs((c_initial - c + 1e-10) / (1e-10 + ec))
        carry_on = (p <= lower) or (p >= upper)
        out = (X, U, c)
        alpha *= rho
    return out + (not carry_on,)


def ilqr(P, th, x0, Xinit, Uinit, maxiter=100, thres=1e-8, line_search=True, ls_kw=None):
    """ilqr.py:221-268.  Returns (State, iterations)."""
    T = Uinit.shape[0]
    local = make_lqr_approx(P, th, x0, T, local=True)
    glob = make_lqr_approx(P, th, x0, T, local=False)
    _, c_old = simulate(P, th, x0, Uinit)
    X, U, it, carry_on = Xinit, Uinit, 0, True
    while carry_on and it < maxiter:
        p = local(X, U)
        gains, (dC, dc) = O.backward(p, T, return_expected_change=True)
        f = lambda a: update(P, th, x0, X, U, gains, a)
        if line_search:
            nX, nU, nc, _ = backtrack(f, c_old, dC, dc, **(ls_kw or {}))
        else:
            nX, nU, nc = f(1.0)
        carry_on = abs((c_old - nc) / c_old) > thres
        X, U, c_old, it = nX, nU, nc, it + 1
    pg = glob(X, U)
    Nu = O.adjoint(X, U, pg, T)
    return O.State(X, U, Nu), it


def kkt(P, th, x0, s):
    """ilqr.py:192-194."""
    T = s.U.shape[0]
    return O.kkt(s, O.Params(x0, make_lqr_approx(P, th, x0, T, local=False)(s.X, s.U)))


def vjp_exact_relu(th, x0, s, w):
    """Exact implicit VJP for the ReluRnn family (f_xx = 0 a.e.): adjoint LQR on the global approximation, then the
    leaf cotangents pulled back through theta -> leaves (Q_t = sym Q, q_t = q, R_t = 1e-4 (R+R'), r_t = 1e-4 r,
    A_t = A diag(relu'(x_t)), B_t = B, Qf_leaf = sym Qf)."""
    T = s.U.shape[0]
    pg = make_lqr_approx(ReluRnn, th, x0, T, local=False)(s.X, s.U)
    g = O.vjp_exact(O.Params(x0, pg), s, w)
    sX = np.concatenate((x0[None], s.X[:-1]))
    sym = lambda Z: 0.5 * (Z + np.swapaxes(Z, -1, -2))
    gth = ReluRnn.Theta(Q=sym(g.lqr.Q).sum(0), q=g.lqr.q.sum(0), Qf=sym(g.lqr.Qf), R=1e-4 * (g.lqr.R + np.swapaxes(g.lqr.R, -1, -2)).sum(0),
                        r=1e-4 * g.lqr.r.sum(0), A=(g.lqr.A * (sX > 0.0)[:, None, :]).sum(0), B=g.lqr.B.sum(0))
    return g.x0, gth
