"""Differentiable time-varying LQR — host-side mirror of the reference module idoc/lqr.py.

Same pytrees (`Gains`, `LQR`, `Params`, field order as idoc/lqr.py:16-59), same entry points
(`backward`, `forward`-through-`direct`, `kkt`, `build(horizon)` -> `typs.Solver`), but every leaf may carry
leading batch axes (what `jax.vmap` of the reference would see) and all arithmetic runs in the sm_100a
kernels behind the C ABI (include/idoc_b200.h).  Host arrays are NumPy; `DeviceProblem`/`DevicePlan`
keep everything resident in HBM for callers that iterate (bench.py, iLQR).
"""
from typing import NamedTuple, Optional

import numpy as np

from . import _lib as L
from . import typs


class Gains(NamedTuple):
    """LQR gains (idoc/lqr.py:16-20)."""
    K: np.ndarray
    k: np.ndarray


class LQR(NamedTuple):
    """LQR specs (idoc/lqr.py:23-35)."""
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

    def symm(self):  # idoc/lqr.py:37-52
        sw = lambda z: 0.5 * (z + np.swapaxes(z, -1, -2))
        return self._replace(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R))


class Params(NamedTuple):
    """LQR parameters (idoc/lqr.py:55-59)."""
    x0: np.ndarray
    lqr: LQR


LEAVES = ("Q", "q", "Qf", "qf", "M", "R", "r", "A", "B", "d")


def _dims(params: Params):
    A, Bm = np.asarray(params.lqr.A), np.asarray(params.lqr.B)
    T, n, m = A.shape[-3], A.shape[-1], Bm.shape[-1]
    batch = A.shape[:-3]
    return batch, T, n, m


def leaf_shapes(B, T, n, m):
    return dict(Q=(B, T, n, n), q=(B, T, n), Qf=(B, n, n), qf=(B, n), M=(B, T, n, m), R=(B, T, m, m), r=(B, T, m),
                A=(B, T, n, n), B=(B, T, n, m), d=(B, T, n), x0=(B, n))


class DeviceProblem:
    """A batch of LQR problems resident in device memory (one DeviceArray per pytree leaf)."""

    def __init__(self, B, T, n, m, dtype):
        self.B, self.T, self.n, self.m, self.dtype = int(B), int(T), int(n), int(m), np.dtype(dtype)
        self.code = L.dtype_code(dtype)
        if not L.lib.idoc_lqr_supported(self.code, self.n, self.m):
            raise L.IdocError(f"unsupported shape n={n}, m={m}: tuned kernels are listed in csrc/shapes.def and "
                              "csrc/mid_shapes.def, the shape-generic kernels cover 1 <= n <= 128, 1 <= m <= 32")
        self.shapes = leaf_shapes(self.B, self.T, self.n, self.m)
        self.arr = {k: L.DeviceArray(s, self.dtype) for k, s in self.shapes.items()}

    @classmethod
    def from_params(cls, params: Params, dtype=None):
        batch, T, n, m = _dims(params)
        B = int(np.prod(batch)) if batch else 1
        dtype = np.dtype(dtype if dtype is not None else np.asarray(params.lqr.A).dtype)
        self = cls(B, T, n, m, dtype)
        self.batch = batch
        for k in LEAVES:
            self.arr[k].copy_from(np.asarray(getattr(params.lqr, k), dtype=dtype).reshape(self.shapes[k]))
        self.arr["x0"].copy_from(np.asarray(params.x0, dtype=dtype).reshape(self.shapes["x0"]))
        return self

    @classmethod
    def generate(cls, B, T, n, m, dtype, seed=0, stream=None):
        """Synthetic random stabilisable problems, generated on the device."""
        self = cls(B, T, n, m, dtype)
        self.batch = (B,)
        a = self.arr
        L.check(L.lib.idoc_lqr_generate(stream.ptr if stream else None, self.code, B, T, n, m, seed,
                                        *[a[k].ptr for k in LEAVES], a["x0"].ptr), "lqr_generate")
        return self

    def to_params(self) -> Params:
        return Params(self.arr["x0"].numpy(), LQR(*[self.arr[k].numpy() for k in LEAVES]))

    def nbytes(self):
        return sum(a.nbytes for a in self.arr.values())


class DevicePlan:
    """Pre-allocated outputs + workspaces for fwd / vjp on a fixed (B, T, n, m, dtype)."""

    def __init__(self, B, T, n, m, dtype, gains=False, vjp=True, reduce_batch=False, status=True):
        self.B, self.T, self.n, self.m, self.dtype = int(B), int(T), int(n), int(m), np.dtype(dtype)
        self.code = L.dtype_code(dtype)
        D = lambda *s: L.DeviceArray(s, self.dtype)
        self.X, self.U, self.Nu = D(B, T, n), D(B, T, m), D(B, T, n)
        self.K = D(B, T, m, n) if gains else None
        self.k = D(B, T, m) if gains else None
        self.dCdc = D(B, 2)
        self.status = L.DeviceArray((B,), np.int32) if status else None
        self.ws_bytes = L.lib.idoc_lqr_ws_bytes(self.code, B, T, n, m)
        self.ws = L.DeviceArray((self.ws_bytes,), np.uint8)
        self.reduce_batch = bool(reduce_batch)
        self.grads = None
        self.grad_buffer = None
        if vjp:
            wsb = L.lib.idoc_lqr_vjp_reduce_ws_bytes if reduce_batch else L.lib.idoc_lqr_vjp_ws_bytes
            self.ws2_bytes = wsb(self.code, B, T, n, m)
            self.ws2 = L.DeviceArray((self.ws2_bytes,), np.uint8)
            shp = leaf_shapes(B, T, n, m)
            if reduce_batch:
                # batch-summed gradients: the ten leaves (shapes without B) are views of ONE contiguous buffer, in pytree
                # order, each 16-byte aligned: that buffer is what a single NCCL all-reduce combines (idoc_b200.dist)
                shp = {k: (s[1:] if k != "x0" else s) for k, s in shp.items()}
                isz = self.dtype.itemsize
                offs, o = {}, 0
                for k in LEAVES:
                    offs[k] = o
                    o += -(-int(np.prod(shp[k])) * isz // 16) * 16
                self.grad_buffer = L.DeviceArray((o // isz,), self.dtype, zero=True)
                self.grads = {k: self.grad_buffer.view(shp[k], offs[k]) for k in LEAVES}
                self.grads["x0"] = L.DeviceArray(shp["x0"], self.dtype)
            else:
                self.grads = {k: L.DeviceArray(s, self.dtype) for k, s in shp.items()}

    def fwd(self, p: DeviceProblem, stream=None):
        a = p.arr
        L.check(L.lib.idoc_lqr_fwd(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                   *[a[k].ptr for k in LEAVES], a["x0"].ptr,
                                   self.X.ptr, self.U.ptr, self.Nu.ptr, L.ptr(self.K), L.ptr(self.k),
                                   self.dCdc.ptr, L.ptr(self.status), self.ws.ptr, self.ws_bytes), "lqr_fwd")

    def backward(self, p: DeviceProblem, stream=None):
        a = p.arr
        L.check(L.lib.idoc_lqr_backward(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                        *[a[k].ptr for k in LEAVES], L.ptr(self.K), L.ptr(self.k),
                                        self.dCdc.ptr, L.ptr(self.status), self.ws.ptr, self.ws_bytes), "lqr_backward")

    def vjp(self, p: DeviceProblem, Xb, Ub, Nub=None, stream=None):
        """Cotangents Xb, Ub, Nub are DeviceArrays ([B,T,n], [B,T,m], [B,T,n] or None)."""
        a, g = p.arr, self.grads
        L.check(L.lib.idoc_lqr_vjp(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                   a["M"].ptr, a["A"].ptr, a["B"].ptr, a["x0"].ptr,
                                   self.X.ptr, self.U.ptr, self.Nu.ptr, self.ws.ptr,
                                   Xb.ptr, Ub.ptr, L.ptr(Nub),
                                   *[g[k].ptr for k in LEAVES], g["x0"].ptr,
                                   self.ws2.ptr, self.ws2_bytes, int(self.reduce_batch)), "lqr_vjp")

    def kkt(self, p: DeviceProblem, X, U, Nu, rX, rU, rNu, stream=None):
        a = p.arr
        L.check(L.lib.idoc_lqr_kkt(stream.ptr if stream else None, self.code, self.B, self.T, self.n, self.m,
                                   *[a[k].ptr for k in LEAVES], a["x0"].ptr, X.ptr, U.ptr, Nu.ptr,
                                   rX.ptr, rU.ptr, rNu.ptr), "lqr_kkt")

    def state(self) -> typs.State:
        return typs.State(self.X.numpy(), self.U.numpy(), self.Nu.numpy())


class _PlanPool:
    """Device buffers of the reference-API calls (`build().direct / vjp`, `kkt`, `backward`) are pooled by
    (B, T, n, m, dtype, vjp, gains): a second call of the same shape reuses its DeviceProblem + DevicePlan instead of
    cudaMalloc-ing ~25 arrays per call.  A (problem, plan) pair handed to a `vjp` pullback stays checked out until the
    pullback is garbage collected (its residuals live in the plan)."""

    def __init__(self, max_free_per_key=2):
        self.free, self.max_free = {}, max_free_per_key

    def acquire(self, B, T, n, m, dtype, vjp, gains=False):
        key = (int(B), int(T), int(n), int(m), np.dtype(dtype).str, bool(vjp), bool(gains))
        lst = self.free.get(key)
        if lst:
            return key, lst.pop()
        return key, (DeviceProblem(B, T, n, m, dtype), DevicePlan(B, T, n, m, dtype, gains=gains, vjp=vjp))

    def release(self, key, pair):
        lst = self.free.setdefault(key, [])
        if len(lst) < self.max_free:
            lst.append(pair)

    def clear(self):
        self.free.clear()


_POOL = _PlanPool()


def _load(p: "DeviceProblem", params: "Params"):
    dt = p.dtype
    for k in LEAVES:
        p.arr[k].copy_from(np.asarray(getattr(params.lqr, k), dtype=dt).reshape(p.shapes[k]))
    p.arr["x0"].copy_from(np.asarray(params.x0, dtype=dt).reshape(p.shapes["x0"]))


def _unbatch(batch, arr, tail):
    return arr.reshape(tuple(batch) + tuple(arr.shape[-tail:]))


def backward(lqr: LQR, horizon: int, *, return_expected_change: bool = False):
    """LQR backward pass (idoc/lqr.py:62-107): returns Gains, optionally the expected_change callable."""
    L.require_device()
    A = np.asarray(lqr.A)
    batch, n = A.shape[:-3], A.shape[-1]
    params = Params(np.zeros(tuple(batch) + (n,), A.dtype), lqr)
    batch, T, n, m = _dims(params)
    B = int(np.prod(batch)) if batch else 1
    assert horizon == T
    key, (p, plan) = _POOL.acquire(B, T, n, m, A.dtype, False, gains=True)
    _load(p, params)
    plan.backward(p)
    gains = Gains(_unbatch(batch, plan.K.numpy(), 3), _unbatch(batch, plan.k.numpy(), 2))
    dCdc = plan.dCdc.numpy().reshape(tuple(batch) + (2,))
    _POOL.release(key, (p, plan))
    if not return_expected_change:
        return gains

    def expected_change(alpha):  # idoc/lqr.py:104-105
        return (alpha ** 2) * dCdc[..., 0] + alpha * dCdc[..., 1]

    return gains, expected_change


def kkt(s: typs.State, params: Params) -> typs.State:
    """LQR KKT conditions (idoc/lqr.py:128-150)."""
    L.require_device()
    batch, T, n, m = _dims(params)
    B = int(np.prod(batch)) if batch else 1
    dt = np.dtype(np.asarray(params.lqr.A).dtype)
    key, (p, plan) = _POOL.acquire(B, T, n, m, dt, False)
    _load(p, params)
    X = L.DeviceArray.from_numpy(np.asarray(s.X, dt).reshape(plan.X.shape))
    U = L.DeviceArray.from_numpy(np.asarray(s.U, dt).reshape(plan.U.shape))
    Nu = L.DeviceArray.from_numpy(np.asarray(s.Nu, dt).reshape(plan.Nu.shape))
    plan.kkt(p, X, U, Nu, plan.X, plan.U, plan.Nu)
    r = plan.state()
    _POOL.release(key, (p, plan))
    return typs.State(*[_unbatch(batch, z, 2) for z in r])


def build(horizon: int) -> typs.Solver:
    """Build a differentiable LQR solver (idoc/lqr.py:167-180).

    direct(params) / implicit(params) -> State(X, U, Nu);  kkt(state, params) -> residual State;
    vjp(params) -> (State, pullback) with pullback(State cotangent) -> Params cotangent (exact
    implicit-function gradient; Nu cotangent may be None).  Device buffers are pooled per shape (_PlanPool)."""
    import weakref

    def _solve(params: Params, want_vjp):
        L.require_device()
        batch, T, n, m = _dims(params)
        B = int(np.prod(batch)) if batch else 1
        if T != horizon:
            raise ValueError(f"horizon {horizon} != leading time axis {T}")
        dt = np.dtype(np.asarray(params.lqr.A).dtype)
        key, (p, plan) = _POOL.acquire(B, T, n, m, dt, want_vjp)
        p.batch = batch
        _load(p, params)
        plan.fwd(p)
        st = plan.state()
        L.report_status(plan.status.numpy(), "lqr.direct")
        return key, p, plan, typs.State(*[_unbatch(batch, z, 2) for z in st])

    def direct(params: Params) -> typs.State:
        key, p, plan, s = _solve(params, False)
        _POOL.release(key, (p, plan))
        return s

    def vjp(params: Params):
        key, p, plan, s = _solve(params, True)

        def pullback(ct: typs.State) -> Params:
            dt = p.dtype
            Xb = L.DeviceArray.from_numpy(np.asarray(ct.X, dt).reshape(plan.X.shape))
            Ub = L.DeviceArray.from_numpy(np.asarray(ct.U, dt).reshape(plan.U.shape))
            Nub = None if ct.Nu is None else L.DeviceArray.from_numpy(np.asarray(ct.Nu, dt).reshape(plan.Nu.shape))
            plan.vjp(p, Xb, Ub, Nub)
            g = {k: a.numpy() for k, a in plan.grads.items()}
            tails = dict(Q=3, q=2, Qf=2, qf=1, M=3, R=3, r=2, A=3, B=3, d=2, x0=1)
            g = {k: _unbatch(p.batch, v, tails[k]) for k, v in g.items()}
            return Params(g["x0"], LQR(*[g[k] for k in LEAVES]))

        weakref.finalize(pullback, _POOL.release, key, (p, plan))  # the residuals live in the plan while the pullback does
        return s, pullback

    return typs.Solver(direct=direct, implicit=direct, kkt=kkt, vjp=vjp)
