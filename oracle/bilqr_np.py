"""CPU ORACLE (test infrastructure, not product code) — the reference's batch iLQR with shared controls,
`/root/reference/idoc/bilqr.py:19-221`, restated through the block-diagonal lift used by oracle/blqr_np.py: the `bs`
systems (same theta, different x0) are one iLQR problem on the stacked state, cost summed, one control sequence.
Pinned against golden vectors produced by executing the unmodified reference bilqr.py (tests/golden/bilqr_*.npz)."""
from typing import NamedTuple

import numpy as np

from . import ilqr_np as IO
from . import lqr_np as O


class ReluAx:
    """tests/test_bilqr.py:39-83: f = relu(A x) + B u + 0.5; cost weights 1e-6."""
    FIELDS = IO.ReluRnn.FIELDS
    Theta = IO.ReluRnn.Theta
    W = 1e-6

    @staticmethod
    def dynamics(x, u, th):
        return np.maximum(th.A @ x, 0.0) + th.B @ u + 0.5

    @classmethod
    def cost(cls, x, u, th):
        return 0.5 * x @ th.Q @ x + th.q @ x + cls.W * (u @ th.R @ u - np.sum(u) * np.sum(x) + th.r @ u)

    @staticmethod
    def costf(x, th):
        return 0.5 * x @ th.Qf @ x

    @classmethod
    def derivs(cls, x, u, th):
        n, m = x.size, u.size
        lxx = 0.5 * (th.Q + th.Q.T)
        luu = cls.W * (th.R + th.R.T)
        lxu = -cls.W * np.ones((n, m))
        lx = lxx @ x + th.q - cls.W * np.sum(u)
        lu = luu @ u - cls.W * np.sum(x) + cls.W * th.r
        fx = th.A * ((th.A @ x) > 0.0)[:, None]
        return lx, lu, lxx, luu, lxu, fx, th.B

    final_derivs = staticmethod(IO.ReluRnn.final_derivs)


def lifted(P, bs, n):
    """The problem family on the stacked state of bs systems sharing the control (bilqr.py:52-91, 93-117)."""
    sl = lambda x: [x[b * n:(b + 1) * n] for b in range(bs)]

    class Lifted:
        @staticmethod
        def dynamics(x, u, th):
            return np.concatenate([P.dynamics(xb, u, th) for xb in sl(x)])

        @staticmethod
        def cost(x, u, th):
            return sum(P.cost(xb, u, th) for xb in sl(x))

        @staticmethod
        def costf(x, th):
            return sum(P.costf(xb, th) for xb in sl(x))

        @staticmethod
        def derivs(x, u, th):
            N, m = bs * n, u.size
            lx, lu, lxx, luu = np.zeros(N), np.zeros(m), np.zeros((N, N)), np.zeros((m, m))
            lxu, fx, fu = np.zeros((N, m)), np.zeros((N, N)), np.zeros((N, m))
            for b, xb in enumerate(sl(x)):
                a, c, A2, R2, M2, F, G = P.derivs(xb, u, th)
                r = slice(b * n, (b + 1) * n)
                lx[r], lxx[r, r], lxu[r], fx[r, r], fu[r] = a, A2, M2, F, G
                lu += c
                luu += R2
            return lx, lu, lxx, luu, lxu, fx, fu

        @staticmethod
        def final_derivs(x, th):
            N = bs * n
            g, H = np.zeros(N), np.zeros((N, N))
            for b, xb in enumerate(sl(x)):
                r = slice(b * n, (b + 1) * n)
                g[r], H[r, r] = P.final_derivs(xb, th)
            return g, H

    return Lifted


def bilqr(P, th, x0, Xinit, Uinit, **kw):
    """bilqr.py:170-213: x0[bs,n], Xinit[T,bs,n] -> State(X[T,bs,n], U[T,m], Nu[T,bs,n]), iterations."""
    bs, n = x0.shape
    T = Uinit.shape[0]
    s, it = IO.ilqr(lifted(P, bs, n), th, x0.reshape(-1), Xinit.reshape(T, bs * n), Uinit, **kw)
    return O.State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n)), it


def kkt(P, th, x0, s):
    bs, n = x0.shape
    T = s.U.shape[0]
    r = IO.kkt(lifted(P, bs, n), th, x0.reshape(-1), O.State(s.X.reshape(T, -1), s.U, s.Nu.reshape(T, -1)))
    return O.State(r.X.reshape(T, bs, n), r.U, r.Nu.reshape(T, bs, n))


def vjp_exact_relu_ax(th, x0, s, w):
    """Exact implicit VJP for ReluAx with shared theta: adjoint LQR on the lifted global approximation, leaf cotangents
    pulled back through theta -> leaves and summed over the systems."""
    bs, n = x0.shape
    T, m = s.U.shape
    L = lifted(ReluAx, bs, n)
    fl = lambda z: z.reshape(T, bs * n)
    sl_ = O.State(fl(s.X), s.U, fl(s.Nu))
    pg = IO.make_lqr_approx(L, th, x0.reshape(-1), T, local=False)(sl_.X, sl_.U)
    g = O.vjp_exact(O.Params(x0.reshape(-1), pg), sl_, O.State(fl(w.X), w.U, fl(w.Nu)))
    sX = np.concatenate((x0.reshape(1, -1), sl_.X[:-1]))
    sym = lambda Z: 0.5 * (Z + np.swapaxes(Z, -1, -2))
    gQ, gq, gQf, gA, gB = np.zeros((n, n)), np.zeros(n), np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, m))
    for b in range(bs):
        r = slice(b * n, (b + 1) * n)
        gQ += sym(g.lqr.Q[:, r, r]).sum(0)
        gq += g.lqr.q[:, r].sum(0)
        gQf += sym(g.lqr.Qf[r, r])
        gate = (np.einsum("ij,tj->ti", th.A, sX[:, r]) > 0.0)           # relu'(A x_t)  [T, n]
        gA += (g.lqr.A[:, r, r] * gate[:, :, None]).sum(0)
        gB += g.lqr.B[:, r, :].sum(0)
    gR = bs * ReluAx.W * (g.lqr.R + np.swapaxes(g.lqr.R, -1, -2)).sum(0)
    gr = bs * ReluAx.W * g.lqr.r.sum(0)
    return g.x0.reshape(bs, n), ReluAx.Theta(Q=gQ, q=gq, Qf=gQf, R=gR, r=gr, A=gA, B=gB)
