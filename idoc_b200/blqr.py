"""Differentiable batch LQR with SHARED controls — host-side mirror of the reference module idoc/blqr.py.

`batch_size` systems are driven by one control sequence U[T, m] (idoc/blqr.py:1, :63-125).  The reference carries
a dense [B, n, B, n] value tensor through einsums; that recursion is exactly the LQR recursion of the block-diagonal
lift (state n' = B*n, A' = blockdiag(A_b), Q' = blockdiag(Q_b), B' = vstack(B_b), M' = vstack(M_b), shared R, r), so
this module lifts on the host and runs the LQR kernels (tuned or shape-generic, n' <= 128) on the device; gradients of
the lifted leaves are un-lifted (diagonal blocks / row slices).  Shapes follow the reference (time axis first):
Q[T,B,n,n] q[T,B,n] Qf[B,n,n] qf[B,n] M[T,B,n,m] R[T,m,m] r[T,m] A[T,B,n,n] B[T,B,n,m] d[T,B,n] x0[B,n]."""
from typing import NamedTuple

import numpy as np

from . import lqr as _lqr
from . import typs


class Gains(NamedTuple):
    """LQR gains (idoc/blqr.py:17-21): K[T, B, m, n], k[T, m]."""
    K: np.ndarray
    k: np.ndarray


class BLQR(NamedTuple):
    """LQR specs (idoc/blqr.py:24-36)."""
    Q: np.ndarray
    q: np.ndarray
    Qf: np.ndarray
    qf: np.ndarray
    M: np.ndarray
    R: np.ndarray
    r: np.ndarray
    A: np.ndarray
    B: np.ndarray
    d: np.ndarray

    def symm(self):  # idoc/blqr.py:38-53
        sw = lambda z: 0.5 * (z + np.swapaxes(z, -1, -2))
        return self._replace(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R))


class Params(NamedTuple):
    """LQR parameters (idoc/blqr.py:56-60)."""
    x0: np.ndarray
    blqr: BLQR


def _bd(Z):
    Z = np.asarray(Z)
    bs, n = Z.shape[-3], Z.shape[-1]
    out = np.zeros(Z.shape[:-3] + (bs * n, bs * n), Z.dtype)
    for b in range(bs):
        out[..., b * n:(b + 1) * n, b * n:(b + 1) * n] = Z[..., b, :, :]
    return out


def lift(params: Params) -> _lqr.Params:
    """The equivalent single LQR with state dimension batch_size * n."""
    l = BLQR(*[np.asarray(z) for z in params.blqr])
    T, bs, n, m = l.B.shape
    cat = lambda Z: Z.reshape(Z.shape[:-2] + (bs * Z.shape[-1],))
    vst = lambda Z: Z.reshape(Z.shape[:-3] + (bs * n, Z.shape[-1]))
    return _lqr.Params(cat(np.asarray(params.x0)),
                       _lqr.LQR(Q=_bd(l.Q), q=cat(l.q), Qf=_bd(l.Qf), qf=cat(l.qf), M=vst(l.M), R=l.R, r=l.r, A=_bd(l.A),
                                B=vst(l.B), d=cat(l.d)))


def _dims(params: Params):
    return np.asarray(params.blqr.B).shape  # T, bs, n, m


def _unlift_grads(g: _lqr.Params, bs: int, n: int) -> Params:
    l = g.lqr
    T, m = l.r.shape
    blk = lambda Z: np.stack([Z[..., b * n:(b + 1) * n, b * n:(b + 1) * n] for b in range(bs)], axis=-3)
    return Params(g.x0.reshape(bs, n),
                  BLQR(Q=blk(l.Q), q=l.q.reshape(T, bs, n), Qf=blk(l.Qf), qf=l.qf.reshape(bs, n), M=l.M.reshape(T, bs, n, m),
                       R=l.R, r=l.r, A=blk(l.A), B=l.B.reshape(T, bs, n, m), d=l.d.reshape(T, bs, n)))


def backward(lqr: BLQR, horizon: int, *, return_expected_change: bool = False):
    """LQR backward pass (idoc/blqr.py:128-157)."""
    T, bs, n, m = np.asarray(lqr.B).shape
    lifted = lift(Params(np.zeros((bs, n), np.asarray(lqr.A).dtype), lqr)).lqr
    out = _lqr.backward(lifted, horizon, return_expected_change=return_expected_change)
    gains = out[0] if return_expected_change else out
    K = np.stack([gains.K[:, :, b * n:(b + 1) * n] for b in range(bs)], axis=1)
    g = Gains(K, gains.k)
    return (g, out[1]) if return_expected_change else g


def kkt(s: typs.State, params: Params) -> typs.State:
    """LQR KKT conditions (idoc/blqr.py:185-226)."""
    T, bs, n, m = _dims(params)
    fl = lambda z: np.asarray(z).reshape(T, bs * n)
    r = _lqr.kkt(typs.State(fl(s.X), np.asarray(s.U), fl(s.Nu)), lift(params))
    return typs.State(r.X.reshape(T, bs, n), r.U, r.Nu.reshape(T, bs, n))


def build(horizon: int) -> typs.Solver:
    """Build a LQR differentiable solver (idoc/blqr.py:245-261)."""
    inner = _lqr.build(horizon)

    def direct(params: Params) -> typs.State:
        T, bs, n, m = _dims(params)
        s = inner.direct(lift(params))
        return typs.State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n))   # blqr.py:253-256

    def vjp(params: Params):
        T, bs, n, m = _dims(params)
        s, pull = inner.vjp(lift(params))

        def pullback(ct: typs.State) -> Params:
            fl = lambda z: None if z is None else np.asarray(z).reshape(T, bs * n)
            return _unlift_grads(pull(typs.State(fl(ct.X), np.asarray(ct.U), fl(ct.Nu))), bs, n)

        return typs.State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n)), pullback

    return typs.Solver(direct=direct, implicit=direct, kkt=kkt, vjp=vjp)
