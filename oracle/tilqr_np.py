"""CPU ORACLE (test infrastructure, not product code) — NumPy fp64 restatement of the reference's
time-invariant LQR, `/root/reference/idoc/tilqr.py`.  Pinned against golden vectors produced by executing the
unmodified reference source on the NumPy jax stand-in (tests/golden/make_golden.py: tilqr_*.npz); the exact
implicit VJP is the time-summed adjoint-LQR VJP of oracle/lqr_np.py on the broadcast problem, checked against
the dense solve assembled from the reference's own kkt.  Leading batch axes allowed on every leaf."""
from typing import NamedTuple

import numpy as np

from . import lqr_np as O

State = O.State


class TILQR(NamedTuple):  # tilqr.py:12-19
    Q: np.ndarray
    Qf: np.ndarray
    R: np.ndarray
    A: np.ndarray
    B: np.ndarray

    def symm(self):  # tilqr.py:21-25
        return TILQR(Q=O._sym(self.Q), Qf=O._sym(self.Qf), R=O._sym(self.R), A=self.A, B=self.B)


class Params(NamedTuple):  # tilqr.py:28-32
    x0: np.ndarray
    lqr: TILQR


def direct(theta: Params, horizon: int) -> State:
    """tilqr.py:56-99."""
    x0 = np.asarray(theta.x0, O._COMPUTE[-1])
    lq = TILQR(*[np.asarray(z, O._COMPUTE[-1]) for z in theta.lqr]).symm()  # tilqr.py:57
    A, B, Q, R, Qf = lq.A, lq.B, lq.Q, lq.R, lq.Qf
    AT, BT = O._T(A), O._T(B)
    P = Qf
    Ks = []
    for _ in range(horizon):  # tilqr.py:74-81
        K = -np.linalg.solve(R + BT @ P @ B, BT @ O._T(P) @ A)  # gain(), tilqr.py:71-72
        P = Q + AT @ P @ A + AT @ P @ B @ K  # tilqr.py:77
        Ks.append(K)
    Ks = Ks[::-1]
    x = x0
    Xs, Us = [], []
    for t in range(horizon):  # tilqr.py:62-69
        u = O._mv(Ks[t], x)
        x = O._mv(A, x) + O._mv(B, u)
        Xs.append(x)
        Us.append(u)
    X, U = np.stack(Xs, axis=-2), np.stack(Us, axis=-2)
    nu = O._mv(Qf, X[..., -1, :])  # tilqr.py:92
    Nus = [nu]
    for t in range(horizon - 1, 0, -1):  # tilqr.py:86-96
        nu = O._mv(AT, nu) + O._mv(Q, X[..., t - 1, :])
        Nus.append(nu)
    return State(X, U, np.stack(Nus[::-1], axis=-2))


def kkt(s: State, theta: Params) -> State:
    """tilqr.py:38-54 (row-vector form there; same residuals)."""
    s = State(*[np.asarray(z, O._COMPUTE[-1]) for z in s])
    x0 = np.asarray(theta.x0, O._COMPUTE[-1])
    lq = TILQR(*[np.asarray(z, O._COMPUTE[-1]) for z in theta.lqr]).symm()
    X, U, Nu = s
    A, B, Q, Qf, R = lq.A, lq.B, lq.Q, lq.Qf, lq.R
    bc = lambda Mx: Mx[..., None, :, :]
    dLdX = np.concatenate((O._mv(bc(Q), X[..., :-1, :]) + O._mtv(bc(A), Nu[..., 1:, :]) - Nu[..., :-1, :],
                           (O._mv(Qf, X[..., -1, :]) - Nu[..., -1, :])[..., None, :]), axis=-2)
    dLdU = O._mtv(bc(B), Nu) + O._mv(bc(R), U)
    sX = np.concatenate((x0[..., None, :], X[..., :-1, :]), axis=-2)
    dLdNu = O._mv(bc(A), sX) + O._mv(bc(B), U) - X
    return State(dLdX, dLdU, dLdNu)


def expand(theta: Params, horizon: int) -> O.Params:
    """The same problem as a time-varying LQR (q = r = d = M = 0)."""
    lq = TILQR(*[np.asarray(z, O._COMPUTE[-1]) for z in theta.lqr])
    n, m = lq.A.shape[-1], lq.B.shape[-1]
    batch = lq.A.shape[:-2]
    rep = lambda z: np.repeat(z[..., None, :, :], horizon, axis=-3)
    zeros = lambda *s_: np.zeros(batch + s_)
    return O.Params(np.asarray(theta.x0, O._COMPUTE[-1]),
                    O.LQR(Q=rep(lq.Q), q=zeros(horizon, n), Qf=lq.Qf, qf=zeros(n), M=zeros(horizon, n, m), R=rep(lq.R),
                          r=zeros(horizon, m), A=rep(lq.A), B=rep(lq.B), d=zeros(horizon, n)))


def vjp_exact(theta: Params, s: State, w: State, horizon: int) -> Params:
    g = O.vjp_exact(expand(theta, horizon), s, w)
    return Params(g.x0, TILQR(Q=g.lqr.Q.sum(axis=-3), Qf=g.lqr.Qf, R=g.lqr.R.sum(axis=-3), A=g.lqr.A.sum(axis=-3),
                              B=g.lqr.B.sum(axis=-3)))
