"""Differentiable time-invariant LQR — host-side mirror of the reference module idoc/tilqr.py.

Same pytrees (`TILQR(Q, Qf, R, A, B)`, `Params(x0, lqr)`, idoc/tilqr.py:12-32) and `build(horizon)` ->
`typs.Solver`; leaves may carry leading batch axes.  The solve runs in the shape-generic kernels with the
time-invariant flag (csrc/generic.cuh) behind `idoc_tilqr_fwd` / `idoc_tilqr_vjp` (include/idoc_b200.h).
"""
from typing import NamedTuple

import numpy as np

from . import _lib as L
from . import lqr as _lqr
from . import typs


class TILQR(NamedTuple):
    """time-invariant LQR specs (idoc/tilqr.py:12-19)."""
    Q: np.ndarray
    Qf: np.ndarray
    R: np.ndarray
    A: np.ndarray
    B: np.ndarray

    def symm(self):  # idoc/tilqr.py:21-25
        sw = lambda z: 0.5 * (z + np.swapaxes(z, -1, -2))
        return TILQR(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R), A=self.A, B=self.B)


class Params(NamedTuple):
    """time-invariant LQR parameters (idoc/tilqr.py:28-32)."""
    x0: np.ndarray
    lqr: TILQR


LEAVES = ("Q", "Qf", "R", "A", "B")


class DevicePlan:
    """Device-resident time-invariant problems + outputs/workspaces for a fixed (B, T, n, m, dtype)."""

    def __init__(self, B, T, n, m, dtype, vjp=True):
        self.B, self.T, self.n, self.m, self.dtype = int(B), int(T), int(n), int(m), np.dtype(dtype)
        self.code = L.dtype_code(dtype)
        D = lambda *s: L.DeviceArray(s, self.dtype)
        self.shapes = dict(Q=(B, n, n), Qf=(B, n, n), R=(B, m, m), A=(B, n, n), B=(B, n, m), x0=(B, n))
        self.arr = {k: D(*s) for k, s in self.shapes.items()}
        self.X, self.U, self.Nu = D(B, T, n), D(B, T, m), D(B, T, n)
        self.status = L.DeviceArray((B,), np.int32)
        self.ws_bytes = L.lib.idoc_lqr_ws_bytes(self.code, B, T, n, m)
        if self.ws_bytes == 0:
            raise L.IdocError(f"unsupported shape n={n}, m={m}")
        self.ws = L.DeviceArray((self.ws_bytes,), np.uint8)
        self.grads = None
        if vjp:
            self.ws2_bytes = L.lib.idoc_lqr_vjp_ws_bytes(self.code, B, T, n, m)
            self.ws2 = L.DeviceArray((self.ws2_bytes,), np.uint8)
            self.grads = {k: D(*s) for k, s in self.shapes.items()}

    def load(self, theta: Params):
        for k in LEAVES:
            self.arr[k].copy_from(np.asarray(getattr(theta.lqr, k), self.dtype).reshape(self.shapes[k]))
        self.arr["x0"].copy_from(np.asarray(theta.x0, self.dtype).reshape(self.shapes["x0"]))

    def fwd(self, stream=None):
        a = self.arr
        L.check(L.lib.idoc_tilqr_fwd(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                     a["Q"].ptr, a["Qf"].ptr, a["R"].ptr, a["A"].ptr, a["B"].ptr, a["x0"].ptr,
                                     self.X.ptr, self.U.ptr, self.Nu.ptr, self.status.ptr, self.ws.ptr, self.ws_bytes),
                "tilqr_fwd")

    def vjp(self, Xb, Ub, Nub=None, stream=None):
        a, g = self.arr, self.grads
        L.check(L.lib.idoc_tilqr_vjp(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                     a["A"].ptr, a["B"].ptr, a["x0"].ptr, self.X.ptr, self.U.ptr, self.Nu.ptr, self.ws.ptr,
                                     Xb.ptr, Ub.ptr, L.ptr(Nub), g["Q"].ptr, g["Qf"].ptr, g["R"].ptr, g["A"].ptr,
                                     g["B"].ptr, g["x0"].ptr, self.ws2.ptr, self.ws2_bytes), "tilqr_vjp")


def _expand(theta: Params, horizon: int) -> _lqr.Params:
    """The same problem as a time-varying LQR (q = r = d = M = 0), for the KKT diagnostic."""
    lq = theta.lqr
    A = np.asarray(lq.A)
    n, m = A.shape[-1], np.asarray(lq.B).shape[-1]
    batch = A.shape[:-2]
    rep = lambda z: np.repeat(np.asarray(z)[..., None, :, :], horizon, axis=-3)
    zeros = lambda *s: np.zeros(batch + s, A.dtype)
    return _lqr.Params(np.asarray(theta.x0), _lqr.LQR(Q=rep(lq.Q), q=zeros(horizon, n), Qf=np.asarray(lq.Qf), qf=zeros(n),
                                                       M=zeros(horizon, n, m), R=rep(lq.R), r=zeros(horizon, m),
                                                       A=rep(lq.A), B=rep(lq.B), d=zeros(horizon, n)))


def build(horizon: int) -> typs.Solver:
    """Build a time-invariant LQR differentiable solver (idoc/tilqr.py:35-104)."""

    def _solve(theta: Params, want_vjp):
        L.require_device()
        A = np.asarray(theta.lqr.A)
        batch, n, m = A.shape[:-2], A.shape[-1], np.asarray(theta.lqr.B).shape[-1]
        B = int(np.prod(batch)) if batch else 1
        plan = DevicePlan(B, horizon, n, m, A.dtype, vjp=want_vjp)
        plan.load(theta)
        plan.fwd()
        L.report_status(plan.status.numpy(), "tilqr.direct")
        un = lambda z, tail: z.reshape(tuple(batch) + z.shape[-tail:])
        return plan, batch, typs.State(un(plan.X.numpy(), 2), un(plan.U.numpy(), 2), un(plan.Nu.numpy(), 2))

    def direct(theta: Params) -> typs.State:
        return _solve(theta, False)[2]

    def kkt(s: typs.State, theta: Params) -> typs.State:  # idoc/tilqr.py:38-54
        return _lqr.kkt(s, _expand(theta, horizon))

    def vjp(theta: Params):
        plan, batch, s = _solve(theta, True)

        def pullback(ct: typs.State) -> Params:
            dt = plan.dtype
            Xb = L.DeviceArray.from_numpy(np.asarray(ct.X, dt).reshape(plan.X.shape))
            Ub = L.DeviceArray.from_numpy(np.asarray(ct.U, dt).reshape(plan.U.shape))
            Nub = None if ct.Nu is None else L.DeviceArray.from_numpy(np.asarray(ct.Nu, dt).reshape(plan.Nu.shape))
            plan.vjp(Xb, Ub, Nub)
            tails = dict(Q=2, Qf=2, R=2, A=2, B=2, x0=1)
            g = {k: a.numpy().reshape(tuple(batch) + a.shape[-tails[k]:]) for k, a in plan.grads.items()}
            return Params(g["x0"], TILQR(*[g[k] for k in LEAVES]))

        return s, pullback

    return typs.Solver(direct=direct, implicit=direct, kkt=kkt, vjp=vjp)
