"""CPU ORACLE (test infrastructure, not product code) — the reference's shared-control batch LQR,
`/root/reference/idoc/blqr.py`, restated through its block-diagonal lift: `batch_size` systems sharing one control
sequence are ONE LQR with state n' = batch_size*n, A' = blockdiag(A_b), Q' = blockdiag(Q_b), B' = vstack(B_b),
M' = vstack(M_b), shared R, r (the einsums of blqr.py:88-121 compute exactly this).  Pinned against golden vectors
produced by executing the unmodified reference blqr.py (tests/golden/blqr_*.npz)."""
from typing import NamedTuple

import numpy as np

from . import lqr_np as O


class BLQR(NamedTuple):  # blqr.py:24-36  (time axis first, then the system axis)
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


class Params(NamedTuple):  # blqr.py:56-60
    x0: np.ndarray
    blqr: BLQR


def _bd(Z):
    """[..., bs, n, n] -> [..., bs*n, bs*n] block diagonal."""
    bs, n = Z.shape[-3], Z.shape[-1]
    out = np.zeros(Z.shape[:-3] + (bs * n, bs * n), Z.dtype)
    for b in range(bs):
        out[..., b * n:(b + 1) * n, b * n:(b + 1) * n] = Z[..., b, :, :]
    return out


def lift(p: Params) -> O.Params:
    l = p.blqr
    T, bs, n, m = l.B.shape
    cat = lambda Z: Z.reshape(Z.shape[:-2] + (bs * Z.shape[-1],))          # [.., bs, n] -> [.., bs*n]
    vst = lambda Z: Z.reshape(Z.shape[:-3] + (bs * n, Z.shape[-1]))         # [.., bs, n, m] -> [.., bs*n, m]
    return O.Params(cat(p.x0), O.LQR(Q=_bd(l.Q), q=cat(l.q), Qf=_bd(l.Qf), qf=cat(l.qf), M=vst(l.M), R=l.R, r=l.r,
                                     A=_bd(l.A), B=vst(l.B), d=cat(l.d)))


def unlift_state(s: O.State, bs: int, n: int) -> O.State:
    T = s.U.shape[0]
    return O.State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n))


def unlift_grads(g: O.Params, bs: int, n: int) -> Params:
    l = g.lqr
    T, m = l.r.shape
    blk = lambda Z: np.stack([Z[..., b * n:(b + 1) * n, b * n:(b + 1) * n] for b in range(bs)], axis=-3)
    return Params(g.x0.reshape(bs, n),
                  BLQR(Q=blk(l.Q), q=l.q.reshape(T, bs, n), Qf=blk(l.Qf), qf=l.qf.reshape(bs, n), M=l.M.reshape(T, bs, n, m),
                       R=l.R, r=l.r, A=blk(l.A), B=l.B.reshape(T, bs, n, m), d=l.d.reshape(T, bs, n)))


def direct(p: Params, return_gains=False):
    T, bs, n, m = p.blqr.B.shape
    s, gains = O.direct(lift(p), T, return_gains=True)
    s = unlift_state(s, bs, n)
    if not return_gains:
        return s
    K = np.stack([gains.K[:, :, b * n:(b + 1) * n] for b in range(bs)], axis=1)   # blqr.Gains.K[T, bs, m, n]
    return s, O.Gains(K, gains.k)


def kkt(s: O.State, p: Params) -> O.State:
    T, bs, n, m = p.blqr.B.shape
    r = O.kkt(O.State(s.X.reshape(T, bs * n), s.U, s.Nu.reshape(T, bs * n)), lift(p))
    return unlift_state(r, bs, n)


def vjp_exact(p: Params, s: O.State, w: O.State) -> Params:
    T, bs, n, m = p.blqr.B.shape
    fl = lambda z: z.reshape(T, bs * n)
    g = O.vjp_exact(lift(p), O.State(fl(s.X), s.U, fl(s.Nu)), O.State(fl(w.X), w.U, fl(w.Nu)))
    return unlift_grads(g, bs, n)
