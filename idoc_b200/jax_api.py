"""JAX front end of idoc_b200: the reference's solver objects on jax arrays, backed by the XLA typed-FFI handlers of
ffi/idoc_b200_ffi.cc (which forward to the C ABI of include/idoc_b200.h).

    import idoc_b200.jax_api as idoc                      # instead of `import idoc`
    solver = idoc.lqr.build(T)                            # idoc/lqr.py:167-180
    s = solver.implicit(params)                           # State(X, U, Nu); params = lqr.Params(x0, lqr.LQR(...)) of jax arrays
    g = jax.grad(lambda p: loss(solver.implicit(p)))(params)   # implicit gradient: ONE adjoint LQR solve, not jaxopt's CG

Same pytrees (NamedTuples, field order of idoc/lqr.py:16-59, idoc/tilqr.py:12-32, idoc/ilqr.py:120-124, idoc/typs.py:6-16);
leading batch axes are native (and `jax.vmap` maps onto them: ffi_call(vmap_method="broadcast_all")).  `direct` and
`implicit` are the same custom_vjp function: the reference's `direct` differentiates by unrolling, `implicit` by CG on the
KKT system; both are replaced by the exact implicit-function VJP.

This module needs jax >= 0.4.38 (jax.ffi) and the compiled shim ffi/libidoc_b200_ffi.so (build line in the .cc header).
Neither exists in the build image of this repository (no jaxlib, no network): importing it there raises ImportError with
that explanation; nothing else in the package imports it.  Written against the documented jax.ffi API; UNTESTED against a
real jaxlib until a JAX-equipped box exists.
"""
import ctypes
import os
from typing import Any, NamedTuple

import numpy as np

try:
    import jax
    import jax.numpy as jnp
    from jax import ffi as _ffi
except Exception as e:  # pragma: no cover - no jax in this image
    raise ImportError("idoc_b200.jax_api needs jax >= 0.4.38 with jax.ffi (not installed here); the NumPy front end "
                      "(idoc_b200.lqr / tilqr / ilqr) works without it") from e

from . import _lib as L
from . import typs
from .line_search import LineSearch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(os.path.dirname(_HERE), "ffi", "libidoc_b200_ffi.so")
if not os.path.exists(_SHIM):  # pragma: no cover
    raise ImportError(f"{_SHIM} is missing: compile ffi/idoc_b200_ffi.cc against jaxlib's headers (see its header comment)")
_shim = ctypes.CDLL(_SHIM)
for _name in ("lqr_fwd", "lqr_backward", "lqr_vjp", "lqr_kkt", "tilqr_fwd", "tilqr_vjp", "ilqr_solve", "ilqr_solve_box",
              "ilqr_vjp", "ilqr_vjp_box", "ilqr_linearize", "blqr_fwd", "blqr_vjp", "bilqr_solve_box", "bilqr_vjp_box"):
    _ffi.register_ffi_target("idoc_" + _name, _ffi.pycapsule(getattr(_shim, f"idoc_{_name}_ffi")), platform="CUDA")

State = typs.State
Solver = typs.Solver


def _sds(shape, dtype):
    return jax.ShapeDtypeStruct(tuple(int(s) for s in shape), dtype)


def _code(dtype):
    return L.dtype_code(np.dtype(dtype))


# ---------------------------------------------------------------------------------------------- lqr
class _LqrModule:
    class Gains(NamedTuple):  # idoc/lqr.py:16-20
        K: Any
        k: Any

    class LQR(NamedTuple):  # idoc/lqr.py:23-35
        Q: Any
        q: Any
        Qf: Any
        qf: Any
        M: Any
        R: Any
        r: Any
        A: Any
        B: Any
        d: Any

        def symm(self):  # idoc/lqr.py:37-52
            sw = lambda z: 0.5 * (z + jnp.swapaxes(z, -1, -2))
            return self._replace(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R))

    class Params(NamedTuple):  # idoc/lqr.py:55-59
        x0: Any
        lqr: Any

    @staticmethod
    def _dims(lqr):
        A, Bm = lqr.A, lqr.B
        batch, T, n, m = A.shape[:-3], A.shape[-3], A.shape[-1], Bm.shape[-1]
        return batch, int(np.prod(batch)) if batch else 1, T, n, m

    @classmethod
    def _fwd(cls, params):
        lq = params.lqr
        batch, B, T, n, m = cls._dims(lq)
        dt = lq.A.dtype
        ws = L.lib.idoc_lqr_ws_bytes(_code(dt), B, T, n, m)
        out = _ffi.ffi_call("idoc_lqr_fwd",
                            (_sds(batch + (T, n), dt), _sds(batch + (T, m), dt), _sds(batch + (T, n), dt),
                             _sds(batch + (T, m, n), dt), _sds(batch + (T, m), dt), _sds(batch + (2,), dt),
                             _sds(batch, jnp.int32), _sds((ws,), jnp.uint8)),
                            vmap_method="broadcast_all")(lq.Q, lq.q, lq.Qf, lq.qf, lq.M, lq.R, lq.r, lq.A, lq.B, lq.d, params.x0)
        X, U, Nu, K, k, dCdc, status, wsbuf = out
        return State(X, U, Nu), cls.Gains(K, k), dCdc, status, wsbuf

    @classmethod
    def backward(cls, lqr, horizon, *, return_expected_change=False):
        """idoc/lqr.py:62-107."""
        batch, B, T, n, m = cls._dims(lqr)
        assert T == horizon
        dt = lqr.A.dtype
        ws = L.lib.idoc_lqr_ws_bytes(_code(dt), B, T, n, m)
        K, k, dCdc, _, _ = _ffi.ffi_call("idoc_lqr_backward",
                                         (_sds(batch + (T, m, n), dt), _sds(batch + (T, m), dt), _sds(batch + (2,), dt),
                                          _sds(batch, jnp.int32), _sds((ws,), jnp.uint8)),
                                         vmap_method="broadcast_all")(lqr.Q, lqr.q, lqr.Qf, lqr.qf, lqr.M, lqr.R, lqr.r, lqr.A,
                                                                      lqr.B, lqr.d)
        gains = cls.Gains(K, k)
        if not return_expected_change:
            return gains
        return gains, (lambda alpha: alpha ** 2 * dCdc[..., 0] + alpha * dCdc[..., 1])  # idoc/lqr.py:104-105

    @classmethod
    def kkt(cls, s, params):
        """idoc/lqr.py:128-150."""
        lq = params.lqr
        dt = lq.A.dtype
        rX, rU, rNu = _ffi.ffi_call("idoc_lqr_kkt", (_sds(s.X.shape, dt), _sds(s.U.shape, dt), _sds(s.Nu.shape, dt)),
                                    vmap_method="broadcast_all")(lq.Q, lq.q, lq.Qf, lq.qf, lq.M, lq.R, lq.r, lq.A, lq.B, lq.d,
                                                                 params.x0, s.X, s.U, s.Nu)
        return State(rX, rU, rNu)

    @classmethod
    def build(cls, horizon):
        """idoc/lqr.py:167-180."""

        @jax.custom_vjp
        def implicit(params):
            return cls._fwd(params)[0]

        def fwd(params):
            s, _, _, _, wsbuf = cls._fwd(params)
            return s, (params, s, wsbuf)

        def bwd(res, ct):
            params, s, wsbuf = res
            lq = params.lqr
            batch, B, T, n, m = cls._dims(lq)
            dt = lq.A.dtype
            ws2 = L.lib.idoc_lqr_vjp_ws_bytes(_code(dt), B, T, n, m)
            shapes = [lq.Q.shape, lq.q.shape, lq.Qf.shape, lq.qf.shape, lq.M.shape, lq.R.shape, lq.r.shape, lq.A.shape,
                      lq.B.shape, lq.d.shape, params.x0.shape]
            out = _ffi.ffi_call("idoc_lqr_vjp", tuple(_sds(sh, dt) for sh in shapes) + (_sds((ws2,), jnp.uint8),),
                                vmap_method="broadcast_all")(lq.M, lq.A, lq.B, params.x0, s.X, s.U, s.Nu, wsbuf, ct.X, ct.U, ct.Nu,
                                                             has_nub=True, reduce_batch=False)
            gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0 = out[:11]
            return (cls.Params(gx0, cls.LQR(gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd)),)

        implicit.defvjp(fwd, bwd)
        return Solver(direct=implicit, implicit=implicit, kkt=cls.kkt)


lqr = _LqrModule


# ---------------------------------------------------------------------------------------------- tilqr
class _TilqrModule:
    class TILQR(NamedTuple):  # idoc/tilqr.py:12-19
        Q: Any
        Qf: Any
        R: Any
        A: Any
        B: Any

        def symm(self):  # idoc/tilqr.py:21-25
            sw = lambda z: 0.5 * (z + jnp.swapaxes(z, -1, -2))
            return self._replace(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R))

    class Params(NamedTuple):  # idoc/tilqr.py:28-32
        x0: Any
        lqr: Any

    @classmethod
    def build(cls, horizon):
        """idoc/tilqr.py:35-104."""
        T = int(horizon)

        def _dims(lq):
            batch, n, m = lq.A.shape[:-2], lq.A.shape[-1], lq.B.shape[-1]
            return batch, int(np.prod(batch)) if batch else 1, n, m

        def _fwd(theta):
            lq = theta.lqr
            batch, B, n, m = _dims(lq)
            dt = lq.A.dtype
            ws = L.lib.idoc_lqr_ws_bytes(_code(dt), B, T, n, m)
            X, U, Nu, status, wsbuf = _ffi.ffi_call(
                "idoc_tilqr_fwd", (_sds(batch + (T, n), dt), _sds(batch + (T, m), dt), _sds(batch + (T, n), dt),
                                   _sds(batch, jnp.int32), _sds((ws,), jnp.uint8)),
                vmap_method="broadcast_all")(lq.Q, lq.Qf, lq.R, lq.A, lq.B, theta.x0, horizon=T)
            return State(X, U, Nu), wsbuf

        @jax.custom_vjp
        def implicit(theta):
            return _fwd(theta)[0]

        def fwd(theta):
            s, wsbuf = _fwd(theta)
            return s, (theta, s, wsbuf)

        def bwd(res, ct):
            theta, s, wsbuf = res
            lq = theta.lqr
            batch, B, n, m = _dims(lq)
            dt = lq.A.dtype
            ws2 = L.lib.idoc_lqr_vjp_ws_bytes(_code(dt), B, T, n, m)
            out = _ffi.ffi_call("idoc_tilqr_vjp",
                                (_sds(lq.Q.shape, dt), _sds(lq.Qf.shape, dt), _sds(lq.R.shape, dt), _sds(lq.A.shape, dt),
                                 _sds(lq.B.shape, dt), _sds(theta.x0.shape, dt), _sds((ws2,), jnp.uint8)),
                                vmap_method="broadcast_all")(lq.A, lq.B, theta.x0, s.X, s.U, s.Nu, wsbuf, ct.X, ct.U, ct.Nu,
                                                             horizon=T, has_nub=True)
            gQ, gQf, gR, gA, gB, gx0 = out[:6]
            return (cls.Params(gx0, cls.TILQR(gQ, gQf, gR, gA, gB)),)

        implicit.defvjp(fwd, bwd)

        def kkt(s, theta):  # idoc/tilqr.py:38-54: the same residuals as lqr.kkt on the broadcast problem
            lq = theta.lqr
            n, m = lq.A.shape[-1], lq.B.shape[-1]
            batch = lq.A.shape[:-2]
            rep = lambda z: jnp.repeat(z[..., None, :, :], T, axis=-3)
            zeros = lambda *sh: jnp.zeros(batch + sh, lq.A.dtype)
            full = lqr.Params(theta.x0, lqr.LQR(Q=rep(lq.Q), q=zeros(T, n), Qf=lq.Qf, qf=zeros(n), M=zeros(T, n, m), R=rep(lq.R),
                                                r=zeros(T, m), A=rep(lq.A), B=rep(lq.B), d=zeros(T, n)))
            return lqr.kkt(s, full)

        return Solver(direct=implicit, implicit=implicit, kkt=kkt)


tilqr = _TilqrModule


# ---------------------------------------------------------------------------------------------- ilqr
class _IlqrModule:
    class Params(NamedTuple):  # idoc/ilqr.py:120-124
        x0: Any
        theta: Any

    @staticmethod
    def _flat(theta, nb):
        leaves = jax.tree_util.tree_leaves(theta)
        return jnp.concatenate([z.reshape(z.shape[:nb] + (-1,)) for z in leaves], axis=-1)

    @staticmethod
    def _limits(control_limits, dt, batch, m):
        if control_limits is None:  # placeholders (has_limits=False)
            return jnp.zeros((1,), dt), jnp.zeros((1,), dt)
        return (jnp.broadcast_to(jnp.asarray(control_limits[0], dt), batch + (m,)),
                jnp.broadcast_to(jnp.asarray(control_limits[1], dt), batch + (m,)))

    @classmethod
    def build(cls, cs, *, maxiter=100, thres=1e-8, line_search=None, unroll=False, jit=True, verbose=False, systems=1,
              control_limits=None):
        """idoc/ilqr.py:179-278 for a compiled problem family (`cs`: idoc_b200.ilqr.Problem); `unroll`, `jit`, `verbose` are
        tracing knobs of the reference: accepted, ignored.  control_limits=(lo, hi): the control-limited extension."""
        from .ilqr import FAMILIES
        pid, T, nsys, m = FAMILIES[cs.family], cs.horizon, cs.state_dim, cs.control_dim
        n = nsys * systems
        lifted = systems > 1 and bool(L.lib.idoc_bilqr_box_supported(pid, nsys, m, systems))
        if control_limits is not None and systems > 1 and not lifted:
            raise L.IdocError(f"control_limits with systems={systems}: no fused shared-control kernel for {cs.family!r}, n={nsys}, m={m}")
        ls = line_search if line_search is not None else LineSearch()
        attrs = dict(maxiter=int(maxiter), thres=float(thres), use_line_search=line_search is not None, ls_lower=float(ls.lower),
                     ls_upper=float(ls.upper), ls_rho=float(ls.rho), ls_maxiter=int(ls.maxiter))

        def _solve(init, params):
            x0 = params.x0
            batch = x0.shape[:-1]
            B = int(np.prod(batch)) if batch else 1
            dt = x0.dtype
            th = cls._flat(params.theta, len(batch))
            ws = L.lib.idoc_ilqr_ws_bytes(_code(dt), B, T, nsys, m, systems)
            res = (_sds(batch + (T, n), dt), _sds(batch + (T, m), dt), _sds(batch + (T, n), dt), _sds(batch, jnp.int32),
                   _sds(batch, jnp.int32), _sds((ws,), jnp.uint8))
            if lifted:  # shared-control systems with a one-kernel lift (csrc/bilqr_fused.cu), limits optional
                lo, hi = cls._limits(control_limits, dt, batch, m)
                out = _ffi.ffi_call("idoc_bilqr_solve_box", res, vmap_method="broadcast_all")(
                    th, x0, lo, hi, init.X, init.U, problem=pid, systems=systems, has_limits=control_limits is not None, **attrs)
            elif control_limits is None:
                out = _ffi.ffi_call("idoc_ilqr_solve", res, vmap_method="broadcast_all")(th, x0, init.X, init.U, problem=pid,
                                                                                      systems=systems, **attrs)
            else:
                lo = jnp.broadcast_to(jnp.asarray(control_limits[0], dt), batch + (m,))
                hi = jnp.broadcast_to(jnp.asarray(control_limits[1], dt), batch + (m,))
                out = _ffi.ffi_call("idoc_ilqr_solve_box", res, vmap_method="broadcast_all")(th, x0, lo, hi, init.X, init.U,
                                                                                          problem=pid, **attrs)
            return State(out[0], out[1], out[2])

        @jax.custom_vjp
        def implicit(init, params):
            return _solve(init, params)

        def fwd(init, params):
            s = _solve(init, params)
            return s, (init, params, s)

        def bwd(res, ct):
            init, params, s = res
            x0 = params.x0
            batch = x0.shape[:-1]
            B = int(np.prod(batch)) if batch else 1
            dt = x0.dtype
            th = cls._flat(params.theta, len(batch))
            ws = L.lib.idoc_ilqr_vjp_ws_bytes(_code(dt), B, T, nsys, m, systems)
            res_t = (_sds(th.shape, dt), _sds(x0.shape, dt), _sds((ws,), jnp.uint8))
            if lifted:
                lo, hi = cls._limits(control_limits, dt, batch, m)
                gth, gx0, _ = _ffi.ffi_call("idoc_bilqr_vjp_box", res_t, vmap_method="broadcast_all")(
                    th, x0, lo, hi, s.X, s.U, s.Nu, ct.X, ct.U, ct.Nu, problem=pid, systems=systems,
                    has_limits=control_limits is not None, has_nub=True)
            elif control_limits is None:
                gth, gx0, _ = _ffi.ffi_call("idoc_ilqr_vjp", res_t, vmap_method="broadcast_all")(
                    th, x0, s.X, s.U, s.Nu, ct.X, ct.U, ct.Nu, problem=pid, systems=systems, has_nub=True)
            else:
                lo = jnp.broadcast_to(jnp.asarray(control_limits[0], dt), batch + (m,))
                hi = jnp.broadcast_to(jnp.asarray(control_limits[1], dt), batch + (m,))
                gth, gx0, _ = _ffi.ffi_call("idoc_ilqr_vjp_box", res_t, vmap_method="broadcast_all")(
                    th, x0, lo, hi, s.X, s.U, s.Nu, ct.X, ct.U, ct.Nu, problem=pid, has_nub=True)
            # un-flatten the theta cotangent into the caller's pytree
            leaves, treedef = jax.tree_util.tree_flatten(params.theta)
            outs, o = [], 0
            for z in leaves:
                k = int(np.prod(z.shape[len(batch):])) if z.ndim > len(batch) else 1
                outs.append(gth[..., o:o + k].reshape(z.shape))
                o += k
            zero_init = jax.tree_util.tree_map(jnp.zeros_like, init)  # the solution does not depend on the initial guess
            return zero_init, cls.Params(gx0, jax.tree_util.tree_unflatten(treedef, outs))

        implicit.defvjp(fwd, bwd)

        def kkt(s, params):  # idoc/ilqr.py:192-194: lqr.kkt on the global approximation at s
            x0 = params.x0
            batch = x0.shape[:-1]
            dt = x0.dtype
            th = cls._flat(params.theta, len(batch))
            shp = dict(Q=(T, n, n), q=(T, n), Qf=(n, n), qf=(n,), M=(T, n, m), R=(T, m, m), r=(T, m), A=(T, n, n), B=(T, n, m), d=(T, n))
            order = ("Q", "q", "Qf", "qf", "M", "R", "r", "A", "B", "d")
            lv = _ffi.ffi_call("idoc_ilqr_linearize", tuple(_sds(batch + shp[k], dt) for k in order),
                               vmap_method="broadcast_all")(th, x0, s.X, s.U, s.Nu, problem=pid, systems=systems, mode=1)
            return lqr.kkt(s, lqr.Params(x0, lqr.LQR(*lv)))

        return Solver(direct=implicit, implicit=implicit, kkt=kkt)


ilqr = _IlqrModule


# ---------------------------------------------------------------------------------------------- blqr (shared controls)
class _BlqrModule:
    class Gains(NamedTuple):  # idoc/blqr.py:17-21
        K: Any
        k: Any

    class BLQR(NamedTuple):  # idoc/blqr.py:24-36
        Q: Any
        q: Any
        Qf: Any
        qf: Any
        M: Any
        R: Any
        r: Any
        A: Any
        B: Any
        d: Any

        def symm(self):  # idoc/blqr.py:38-53
            sw = lambda z: 0.5 * (z + jnp.swapaxes(z, -1, -2))
            return self._replace(Q=sw(self.Q), Qf=sw(self.Qf), R=sw(self.R))

    class Params(NamedTuple):  # idoc/blqr.py:56-60
        x0: Any
        blqr: Any

    @staticmethod
    def _dims(l):  # A[..., T, bs, n, n], B[..., T, bs, n, m]
        A, Bm = l.A, l.B
        lead, T, bs, n, m = A.shape[:-4], A.shape[-4], A.shape[-3], A.shape[-1], Bm.shape[-1]
        return lead, int(np.prod(lead)) if lead else 1, T, bs, n, m

    @classmethod
    def _fwd(cls, params):
        l = params.blqr
        lead, P, T, bs, n, m = cls._dims(l)
        dt = l.A.dtype
        ws = L.lib.idoc_blqr_ws_bytes(_code(dt), P, T, bs, n, m)
        out = _ffi.ffi_call("idoc_blqr_fwd",
                            (_sds(lead + (T, bs, n), dt), _sds(lead + (T, m), dt), _sds(lead + (T, bs, n), dt),
                             _sds(lead + (T, bs, m, n), dt), _sds(lead + (T, m), dt), _sds(lead + (2,), dt),
                             _sds(lead, jnp.int32), _sds((ws,), jnp.uint8)),
                            vmap_method="broadcast_all")(l.Q, l.q, l.Qf, l.qf, l.M, l.R, l.r, l.A, l.B, l.d, params.x0)
        X, U, Nu, K, k, dCdc, _, _ = out
        return State(X, U, Nu), cls.Gains(K, k), dCdc

    @classmethod
    def backward(cls, blqr, horizon, *, return_expected_change=False):
        """idoc/blqr.py:128-157."""
        lead, P, T, bs, n, m = cls._dims(blqr)
        assert T == horizon
        _, gains, dCdc = cls._fwd(cls.Params(jnp.zeros(lead + (bs, n), blqr.A.dtype), blqr))
        if not return_expected_change:
            return gains
        return gains, (lambda alpha: alpha ** 2 * dCdc[..., 0] + alpha * dCdc[..., 1])  # idoc/blqr.py:154-155

    @classmethod
    def kkt(cls, s, params):
        """idoc/blqr.py:185-226: the LQR KKT kernel on the systems as a batch that sees the shared U and a 1/bs share of (R, r);
        dLdU is the sum of the per-system control residuals."""
        l = params.blqr
        bs = l.A.shape[-3]
        sw = lambda z, k: jnp.moveaxis(z, -k, -k - 1)  # [..., T, bs, tail...] -> [..., bs, T, tail...]
        rep = lambda z, k: jnp.broadcast_to(jnp.expand_dims(z, -k - 1) / bs, z.shape[:-k] + (bs,) + z.shape[-k:])
        per = lqr.Params(params.x0, lqr.LQR(Q=sw(l.Q, 3), q=sw(l.q, 2), Qf=l.Qf, qf=l.qf, M=sw(l.M, 3), R=rep(l.R, 3), r=rep(l.r, 2),
                                            A=sw(l.A, 3), B=sw(l.B, 3), d=sw(l.d, 2)))
        U = jnp.broadcast_to(jnp.expand_dims(s.U, -3), s.U.shape[:-2] + (bs,) + s.U.shape[-2:])
        r = lqr.kkt(State(sw(s.X, 2), U, sw(s.Nu, 2)), per)
        return State(jnp.moveaxis(r.X, -3, -2), r.U.sum(axis=-3), jnp.moveaxis(r.Nu, -3, -2))

    @classmethod
    def build(cls, horizon):
        """idoc/blqr.py:245-261."""

        @jax.custom_vjp
        def implicit(params):
            return cls._fwd(params)[0]

        def fwd(params):
            s = cls._fwd(params)[0]
            return s, (params, s)

        def bwd(res, ct):
            params, s = res
            l = params.blqr
            lead, P, T, bs, n, m = cls._dims(l)
            dt = l.A.dtype
            ws = L.lib.idoc_blqr_vjp_ws_bytes(_code(dt), P, T, bs, n, m)
            shapes = [l.Q.shape, l.q.shape, l.Qf.shape, l.qf.shape, l.M.shape, l.R.shape, l.r.shape, l.A.shape, l.B.shape,
                      l.d.shape, params.x0.shape]
            out = _ffi.ffi_call("idoc_blqr_vjp", tuple(_sds(sh, dt) for sh in shapes) + (_sds((ws,), jnp.uint8),),
                                vmap_method="broadcast_all")(l.Q, l.Qf, l.M, l.R, l.A, l.B, params.x0, s.X, s.U, s.Nu, ct.X, ct.U,
                                                             ct.Nu, has_nub=True)
            return (cls.Params(out[10], cls.BLQR(*out[:10])),)

        implicit.defvjp(fwd, bwd)
        return Solver(direct=implicit, implicit=implicit, kkt=cls.kkt)


blqr = _BlqrModule


# ---------------------------------------------------------------------------------------------- bilqr (shared controls, iLQR)
class _BilqrModule:
    class Params(NamedTuple):  # idoc/bilqr.py:45-49: x0[bs, n], theta shared by the systems
        x0: Any
        theta: Any

    @classmethod
    def build(cls, cs, *, maxiter=100, thres=1e-8, line_search=None, unroll=False, jit=True, verbose=False, control_limits=None):
        """idoc/bilqr.py:120-221: State(X[T, bs, n], U[T, m], Nu[T, bs, n]); the systems are the block-diagonal lift of
        ilqr.build(systems=bs) (one kernel where csrc/bilqr_fused.cu has the lift, else the multi-kernel loop)."""
        T, n = cs.horizon, cs.state_dim
        cache = {}

        def inner(bs):
            if bs not in cache:
                cache[bs] = ilqr.build(cs, maxiter=maxiter, thres=thres, line_search=line_search, systems=bs,
                                       control_limits=control_limits)
            return cache[bs]

        lift = lambda s, bs: State(s.X.reshape(T, bs * n), s.U, s.Nu.reshape(T, bs * n))
        unlift = lambda s, bs: State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n))

        def implicit(init, params):  # reshapes are differentiable: jax.grad flows through ilqr's custom_vjp
            bs = params.x0.shape[0]
            return unlift(inner(bs).implicit(lift(init, bs), ilqr.Params(params.x0.reshape(-1), params.theta)), bs)

        def kkt(s, params):
            bs = params.x0.shape[0]
            return unlift(inner(bs).kkt(lift(s, bs), ilqr.Params(params.x0.reshape(-1), params.theta)), bs)

        return Solver(direct=implicit, implicit=implicit, kkt=kkt)


bilqr = _BilqrModule
