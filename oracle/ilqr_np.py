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


# ---------------------------------------------------------------------------------------- families by autodiff
# BASELINE config 4's systems (pendulum, cart-pole) and the user-family example (unicycle) have no counterpart in the
# reference (its iLQR tests use ReluRnn only, tests/test_ilqr.py:42-56), so the family DEFINITIONS below are ours — the
# same formulas as idoc_b200/csrc/ilqr.cuh (PendulumF, CartpoleF) and csrc/families/unicycle.cuh — while everything
# done WITH them is the reference's algorithm: make_lqr_approx (ilqr.py:127-158) differentiates cost and dynamics
# (here with oracle/ad_np.py jets instead of jax.grad / jacfwd), and the loop above is ilqr.py:221-268.
from . import ad_np as AD  # noqa: E402


class AdFamily:
    """A problem family given by dyn(th, x, u) -> [n], cost(th, x, u), costf(th, x) over generic scalars (floats or Jets)."""

    def __init__(self, name, n, m, theta_dim, dyn, cost, costf):
        self.name, self.n, self.m, self.theta_dim = name, n, m, theta_dim
        self._dyn, self._cost, self._costf = dyn, cost, costf

    def dynamics(self, x, u, th):
        return np.array([float(v) for v in self._dyn(list(th), list(x), list(u))])

    def cost(self, x, u, th):
        return float(self._cost(list(th), list(x), list(u)))

    def costf(self, x, th):
        return float(self._costf(list(th), list(x)))

    def _jets(self, x, u, th, with_theta=False):
        vals = list(x) + list(u) + (list(th) if with_theta else [])
        z = AD.Jet.seed(vals)
        n, m = self.n, self.m
        thj = z[n + m:] if with_theta else list(th)
        l = self._cost(thj, z[:n], z[n:n + m])
        f = self._dyn(thj, z[:n], z[n:n + m])
        return l, f

    def derivs(self, x, u, th):
        """l_x, l_u, l_xx, l_uu, l_xu, f_x, f_u  (jax.grad / jacfwd / jacobian of ilqr.py:132-138)."""
        n = self.n
        l, f = self._jets(x, u, th)
        fx = np.stack([fi.g[:n] for fi in f])
        fu = np.stack([fi.g[n:] for fi in f])
        return l.g[:n], l.g[n:], l.H[:n, :n], l.H[n:, n:], l.H[:n, n:], fx, fu

    def final_derivs(self, x, th):
        z = AD.Jet.seed(list(x))
        lf = self._costf(list(th), z)
        return lf.g.copy(), lf.H.copy()

    def curvature(self, x, u, th, nu):
        """sum_i nu_i Hess f_i: the dynamics part of the Hessian of the Lagrangian -> (Qc, Rc, Mc)."""
        n = self.n
        _, f = self._jets(x, u, th)
        H = sum(nu[i] * f[i].H for i in range(n))
        return H[:n, :n], H[n:, n:], H[:n, n:]


def vjp_exact_ad(P: AdFamily, th, x0, s, w):
    """Exact implicit VJP at a KKT point s for an AdFamily: the linear system custom_root(kkt) poses (ilqr.py:270-272) has
    the FULL Hessian of the Lagrangian as its quadratic blocks (kkt = lqr.kkt on the global approximation, differentiated
    through make_lqr_approx), so it is the adjoint LQR of oracle/lqr_np.vjp_exact on
    Q' = l_xx + nu.f_xx, R' = l_uu + nu.f_uu, M' = l_xu + nu.f_xu; theta receives
    d/dtheta [ sux.l_x + uU.l_u + nu.(f_x sux + f_u uU) + uNu.f ] per step + d/dtheta [ uX_T . lf_x ]  (mixed second
    derivatives from jets seeded in (x, u, theta)).  Returns (gx0, gtheta)."""
    T, n, m = s.U.shape[0], P.n, P.m
    pg = make_lqr_approx(P, th, x0, T, local=False)(s.X, s.U)
    sX = np.concatenate((x0[None], s.X[:-1]))
    Q, R, M = pg.Q.copy(), pg.R.copy(), pg.M.copy()
    for t in range(T):
        Qc, Rc, Mc = P.curvature(sX[t], s.U[t], th, s.Nu[t])
        Q[t] += Qc
        R[t] += Rc
        M[t] += Mc
    g = O.vjp_exact(O.Params(x0, pg._replace(Q=Q, R=R, M=M)), s, w)
    sux, uU, uNu, uXT = g.lqr.q, g.lqr.r, g.lqr.d, g.lqr.qf
    gth = np.zeros(P.theta_dim)
    for t in range(T):
        l, f = P._jets(sX[t], s.U[t], th, with_theta=True)
        dz = np.concatenate((sux[t], uU[t]))
        gth += dz @ l.H[:n + m, n + m:]
        for i in range(n):
            gth += s.Nu[t, i] * (dz @ f[i].H[:n + m, n + m:]) + uNu[t, i] * f[i].g[n + m:]
    z = AD.Jet.seed(list(s.X[-1]) + list(th))
    lf = P._costf(z[n:], z[:n])
    gth += uXT @ lf.H[:n, n:]
    return g.x0, gth


_PI = 3.14159265358979323846


def _pend_dyn(th, x, u):
    return [x[0] + th[0] * x[1], x[1] + th[0] * (th[3] * u[0] - th[1] * AD.sin(x[0]) - th[2] * x[1])]


def _pend_cost(th, x, u):
    e = x[0] - _PI
    return 0.5 * (th[4] * e * e + th[5] * x[1] * x[1] + th[6] * u[0] * u[0])


def _pend_costf(th, x):
    e = x[0] - _PI
    return 0.5 * (th[7] * e * e + th[8] * x[1] * x[1])


def _cart_dyn(th, x, u):
    sn, cs = AD.sin(x[1]), AD.cos(x[1])
    pdd = (u[0] + th[2] * sn * (th[3] * x[3] * x[3] + th[4] * cs)) / (th[1] + th[2] * sn * sn)
    tdd = -(pdd * cs + th[4] * sn) / th[3]
    return [x[0] + th[0] * x[2], x[1] + th[0] * x[3], x[2] + th[0] * pdd, x[3] + th[0] * tdd]


def _cart_cost(th, x, u):
    e = x[1] - _PI
    return 0.5 * (th[5] * x[0] * x[0] + th[6] * e * e + th[7] * x[2] * x[2] + th[8] * x[3] * x[3] + th[9] * u[0] * u[0])


def _cart_costf(th, x):
    e = x[1] - _PI
    return 0.5 * (th[10] * x[0] * x[0] + th[11] * e * e + th[12] * x[2] * x[2] + th[13] * x[3] * x[3])


def _uni_dyn(th, x, u):
    return [x[0] + th[0] * u[0] * AD.cos(x[2]), x[1] + th[0] * u[0] * AD.sin(x[2]), x[2] + th[0] * u[1]]


def _uni_cost(th, x, u):
    ex, ey = x[0] - th[1], x[1] - th[2]
    return 0.5 * (th[3] * (ex * ex + ey * ey) + th[4] * x[2] * x[2] + th[5] * u[0] * u[0] + th[6] * u[1] * u[1])


def _uni_costf(th, x):
    ex, ey = x[0] - th[1], x[1] - th[2]
    return 0.5 * (th[7] * (ex * ex + ey * ey) + th[8] * x[2] * x[2])


# theta = [h, g/l, b, c, w1, w2, wu, f1, f2]
Pendulum = AdFamily("pendulum", 2, 1, 9, _pend_dyn, _pend_cost, _pend_costf)
# theta = [h, m_c, m_p, l, g, w_p, w_th, w_v, w_om, w_u, f_p, f_th, f_v, f_om]
Cartpole = AdFamily("cartpole", 4, 1, 14, _cart_dyn, _cart_cost, _cart_costf)
# theta = [h, goal_x, goal_y, w_p, w_phi, w_v, w_om, f_p, f_phi]
Unicycle = AdFamily("unicycle", 3, 2, 9, _uni_dyn, _uni_cost, _uni_costf)
