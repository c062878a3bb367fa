"""Differentiable batch LQR with SHARED controls — host-side mirror of the reference module idoc/blqr.py.

`batch_size` systems are driven by one control sequence U[T, m] (idoc/blqr.py:1, :63-125).  The reference carries a dense
[B, n, B, n] value tensor through einsums; that recursion is the LQR recursion of the block-diagonal lift (state n' = B*n,
A' = blockdiag(A_b), Q' = blockdiag(Q_b), B' = vstack(B_b), M' = vstack(M_b), shared R, r), whose value matrix is
block-diagonal minus a low-rank term.  The device solver (csrc/blqr.cuh behind idoc_blqr_fwd / idoc_blqr_vjp) carries exactly
that structure: B diagonal n x n blocks + the accumulated rank-m-per-step correction, so work and memory are LINEAR in the
number of systems (any batch_size; per-system n <= 16, m <= 8).  `lift` is kept for the KKT diagnostic of small problems
and for tests.  Shapes follow the reference (time axis first):
Q[T,B,n,n] q[T,B,n] Qf[B,n,n] qf[B,n] M[T,B,n,m] R[T,m,m] r[T,m] A[T,B,n,n] B[T,B,n,m] d[T,B,n] x0[B,n]."""
from typing import NamedTuple

import numpy as np

from . import _lib as L
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


FIELDS = BLQR._fields


class _Dev:
    """device copies of a blqr problem + outputs + workspace for one (T, bs, n, m, dtype)"""

    def __init__(self, params: Params, want_vjp: bool):
        L.require_device()
        l = BLQR(*[np.asarray(z) for z in params.blqr])
        self.T, self.bs, self.n, self.m = l.B.shape
        T, bs, n, m = self.T, self.bs, self.n, self.m
        self.dtype = np.dtype(l.A.dtype if l.A.dtype in (np.float32, np.float64) else np.float64)
        self.code = L.dtype_code(self.dtype)
        dt = self.dtype
        self.arr = {k: L.DeviceArray.from_numpy(np.ascontiguousarray(getattr(l, k), dt)) for k in FIELDS}
        self.x0 = L.DeviceArray.from_numpy(np.ascontiguousarray(np.asarray(params.x0), dt).reshape(bs, n))
        D = lambda *s: L.DeviceArray(s, dt)
        self.X, self.U, self.Nu = D(T, bs, n), D(T, m), D(T, bs, n)
        self.K, self.k, self.dCdc = D(T, bs, m, n), D(T, m), D(2)
        self.status = L.DeviceArray((1,), np.int32)
        self.ws_bytes = L.lib.idoc_blqr_ws_bytes(self.code, 1, T, bs, n, m)
        self.ws = L.DeviceArray((self.ws_bytes,), np.uint8)
        if want_vjp:
            self.vws_bytes = L.lib.idoc_blqr_vjp_ws_bytes(self.code, 1, T, bs, n, m)
            self.vws = L.DeviceArray((self.vws_bytes,), np.uint8)

    def fwd(self, rollout=True):
        a = self.arr
        L.check(L.lib.idoc_blqr_fwd(None, self.code, 1, self.T, self.bs, self.n, self.m, *[a[k].ptr for k in FIELDS], self.x0.ptr,
                                    self.X.ptr if rollout else None, self.U.ptr if rollout else None,
                                    self.Nu.ptr if rollout else None, self.K.ptr, self.k.ptr, self.dCdc.ptr, self.status.ptr,
                                    self.ws.ptr, self.ws_bytes), "blqr_fwd")
        L.report_status(self.status.numpy(), "blqr")

    def vjp(self, ct: typs.State) -> Params:
        T, bs, n, m, dt = self.T, self.bs, self.n, self.m, self.dtype
        a = self.arr
        Xb = L.DeviceArray.from_numpy(np.ascontiguousarray(np.asarray(ct.X, dt)).reshape(T, bs, n))
        Ub = L.DeviceArray.from_numpy(np.ascontiguousarray(np.asarray(ct.U, dt)).reshape(T, m))
        Nub = None if ct.Nu is None else L.DeviceArray.from_numpy(np.ascontiguousarray(np.asarray(ct.Nu, dt)).reshape(T, bs, n))
        g = {k: L.DeviceArray(a[k].shape, dt) for k in FIELDS}
        gx0 = L.DeviceArray((bs, n), dt)
        L.check(L.lib.idoc_blqr_vjp(None, self.code, 1, T, bs, n, m, a["Q"].ptr, a["Qf"].ptr, a["M"].ptr, a["R"].ptr, a["A"].ptr,
                                    a["B"].ptr, self.x0.ptr, self.X.ptr, self.U.ptr, self.Nu.ptr, Xb.ptr, Ub.ptr, L.ptr(Nub),
                                    *[g[k].ptr for k in FIELDS], gx0.ptr, self.vws.ptr, self.vws_bytes), "blqr_vjp")
        return Params(gx0.numpy(), BLQR(*[g[k].numpy() for k in FIELDS]))


def backward(lqr: BLQR, horizon: int, *, return_expected_change: bool = False):
    """LQR backward pass (idoc/blqr.py:128-157)."""
    T, bs, n, m = np.asarray(lqr.B).shape
    assert T == horizon
    dev = _Dev(Params(np.zeros((bs, n), np.asarray(lqr.A).dtype), lqr), False)
    dev.fwd(rollout=False)
    g = Gains(dev.K.numpy(), dev.k.numpy())
    if not return_expected_change:
        return g
    dCdc = dev.dCdc.numpy()
    return g, (lambda alpha: (alpha ** 2) * dCdc[0] + alpha * dCdc[1])  # idoc/blqr.py:154-155


def kkt(s: typs.State, params: Params) -> typs.State:
    """LQR KKT conditions (idoc/blqr.py:185-226).  Evaluated by the LQR KKT kernel with the systems as a batch of problems
    that see the shared U and a 1/batch_size share of (R, r): dLdX and dLdNu are per system, dLdU is the sum over the
    systems of the per-system control residuals."""
    T, bs, n, m = _dims(params)
    l = BLQR(*[np.asarray(z) for z in params.blqr])
    sw = lambda z: np.ascontiguousarray(np.swapaxes(z, 0, 1))  # [T, bs, ...] -> [bs, T, ...]
    rep = lambda z: np.broadcast_to(z[None] / bs, (bs,) + z.shape).copy()
    per = _lqr.Params(np.asarray(params.x0), _lqr.LQR(Q=sw(l.Q), q=sw(l.q), Qf=l.Qf, qf=l.qf, M=sw(l.M), R=rep(l.R), r=rep(l.r),
                                                        A=sw(l.A), B=sw(l.B), d=sw(l.d)))
    U = np.asarray(s.U)
    r = _lqr.kkt(typs.State(sw(np.asarray(s.X)), np.broadcast_to(U[None], (bs,) + U.shape).copy(), sw(np.asarray(s.Nu))), per)
    return typs.State(np.swapaxes(r.X, 0, 1), r.U.sum(axis=0), np.swapaxes(r.Nu, 0, 1))


def build(horizon: int) -> typs.Solver:
    """Build a LQR differentiable solver (idoc/blqr.py:245-261)."""

    def direct(params: Params) -> typs.State:
        dev = _Dev(params, False)
        assert dev.T == horizon
        dev.fwd()
        return typs.State(dev.X.numpy(), dev.U.numpy(), dev.Nu.numpy())   # blqr.py:253-256

    def vjp(params: Params):
        dev = _Dev(params, True)
        assert dev.T == horizon
        dev.fwd()
        s = typs.State(dev.X.numpy(), dev.U.numpy(), dev.Nu.numpy())
        return s, dev.vjp

    return typs.Solver(direct=direct, implicit=direct, kkt=kkt, vjp=vjp)
