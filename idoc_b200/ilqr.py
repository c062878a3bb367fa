"""Differentiable iLQR — host-side mirror of the reference module idoc/ilqr.py.

`Problem` keeps the reference's fields (idoc/ilqr.py:94-117) but, because a CUDA kernel cannot call Python
callables, the cost/dynamics are named by `family` (a compiled device functor, csrc/ilqr.cuh):
  "relu_rnn"  tests/test_ilqr.py:14-64, theta = NamedTuple/tuple (Q, q, Qf, R, r, A, B)
  "relu_ax"   tests/test_bilqr.py:39-83 (f = relu(A x) + B u + 0.5, cost weights 1e-6), same theta layout
  "pendulum"  damped pendulum swing-up, theta = [h, g/l, b, c, w1, w2, wu, f1, f2]
  "cartpole"  cart-pole swing-up (n=4, m=1), theta = [h, m_c, m_p, l, g, w_p, w_th, w_v, w_om, w_u, f_p, f_th, f_v, f_om];
              written once over the scalar type and differentiated on the device by forward-mode AD (csrc/ad.cuh), the
              counterpart of the reference's jax autodiff in make_lqr_approx; "pendulum_ad" = the pendulum through that path
  <name>      any family defined AT RUN TIME from source text with define_family(name, source) (nvcc builds a small plugin
              library in a few seconds; no rebuild of the product library), or dropped into csrc/families/<file>.cuh (dynamics, cost, costf over a generic scalar, registered with
              IDOC_FAMILY(name, Struct)) and compiled in by build(): e.g. "unicycle" (n=3, m=2, 9 parameters); FAMILIES lists them
`build(problem, maxiter, thres, line_search, ...)` returns `typs.Solver(direct(init, params), implicit, kkt, vjp)`
like idoc/ilqr.py:179-278; the whole linearise -> Riccati -> line-search loop runs on the device
(idoc_ilqr_solve), the implicit VJP is one adjoint LQR solve (idoc_ilqr_vjp)."""
from dataclasses import dataclass
from typing import Any, NamedTuple, Optional

import os

import numpy as np

from . import _lib as L
from . import lqr as _lqr
from . import typs
from .line_search import LineSearch

FAMILIES = {"relu_rnn": 0, "pendulum": 1, "relu_ax": 2, "cartpole": 3, "pendulum_ad": 4}
# user-defined families compiled into the library (csrc/families/*.cuh: cost, costf, dynamics over a generic scalar)
FAMILIES.update({name: info[0] for name, info in L.user_families().items()})
PLUGINS = {}  # family name -> ctypes handle of its run-time plugin library (define_family)


def define_family(name: str, source: str, *, nvcc: str = "nvcc", verbose: bool = False) -> str:
    """Define an iLQR problem family AT RUN TIME from CUDA source text -- the counterpart of handing Python callables to the
    reference's ilqr.Problem (idoc/ilqr.py:94-117), which a kernel cannot call.  `source` declares a struct with
    `static constexpr int N, M, TD` and the three functions `dyn(th, x, u, f)`, `cost(th, x, u)`, `costf(th, x)` templated
    over the scalar type (see csrc/families/unicycle.cuh; `msin`, `mcos`, `mexp`, `msqrt`, `mtanh` are the differentiable
    elementary functions of csrc/ad.cuh) and ends with `IDOC_FAMILY(<name>, <Struct>)`.  The device forward-mode AD supplies
    every derivative.  nvcc compiles csrc/ilqr_api.cu for this one family into a small plugin library (a few seconds; cached
    by the hash of the source under idoc_b200/csrc/build/plugins/), which exports the same idoc_ilqr_* entry points and calls
    the LQR kernels of libidoc_b200.so.  Afterwards `Problem(family=name, ...)` works like a built-in family: fused solve, fused
    implicit VJP, kkt, control limits, fp32 and fp64.  Needs the CUDA toolkit at run time (the same one that built the library)."""
    import hashlib
    import re
    import subprocess
    m = re.search(r"IDOC_FAMILY\(\s*(\w+)\s*,\s*(\w+)\s*\)", source)
    if not m or m.group(1) != name:
        raise L.IdocError(f"define_family: the source must end with IDOC_FAMILY({name}, <Struct>)")
    if name in FAMILIES and name not in PLUGINS:
        raise L.IdocError(f"define_family: {name!r} is a family compiled into the library")
    struct = m.group(2)
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(here, "csrc")
    lib_mtime = str(os.path.getmtime(L.LIB_PATH))
    tag = hashlib.sha256((source + lib_mtime).encode()).hexdigest()[:16]
    d = os.path.join(csrc, "build", "plugins", f"{name}_{tag}")
    so = os.path.join(d, f"libidoc_family_{name}.so")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        fam = os.path.join(d, "family.cuh")
        open(fam, "w").write("#pragma once\n" + source + "\n")
        inc = os.path.join(d, "families.inc")
        pid = 16  # IDOC_PROBLEM_USER0: the plugin holds exactly one family
        open(inc, "w").write(f'#if defined(IDOC_FAMILY_INCLUDES)\n#include "{fam}"\n#elif defined(IDOC_FAMILY_DISPATCH_DEF)\n'
                             f"#define IDOC_DISPATCH_USER(CALL) IDOC_USER_FAMILY_CALL(CALL, {pid}, {struct})\n"
                             f'#elif defined(IDOC_USER_FAMILY)\nIDOC_USER_FAMILY({pid}, "{name}", {struct})\n#endif\n')
        obj = os.path.join(d, "ilqr_api.o")
        cmds = [[nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                 "-Xcompiler", "-fPIC", "-I", csrc, "-DIDOC_PLUGIN_ONLY", f'-DIDOC_FAMILIES_INC="{inc}"', "-c",
                 os.path.join(csrc, "ilqr_api.cu"), "-o", obj],
                [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so + ".tmp", obj, "-L", here,
                 "-l:libidoc_b200.so", "-Xlinker", "-Bsymbolic", "-Xlinker", f"-rpath={here}"]]
        for cmd in cmds:
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose:
                print(" ".join(cmd), r.stdout, r.stderr, sep="\n")
            if r.returncode != 0:
                raise L.IdocError(f"define_family({name!r}): nvcc failed:\n{r.stderr[-4000:]}")
        os.replace(so + ".tmp", so)
    h = L.load_ilqr_plugin(so)
    if h.idoc_ilqr_family_count() != 1:
        raise L.IdocError(f"define_family({name!r}): the plugin does not hold exactly one family")
    PLUGINS[name] = h
    FAMILIES[name] = 16
    return so


@dataclass
class Problem:
    """iLQR Problem (idoc/ilqr.py:94-117); `family` names the compiled cost/dynamics."""
    family: str
    horizon: int
    state_dim: int
    control_dim: int
    cost: Any = None       # kept for signature compatibility; must be None (Python callables cannot run on the device)
    costf: Any = None
    dynamics: Any = None


class Params(NamedTuple):
    """iLQR Params (idoc/ilqr.py:120-124)."""
    x0: np.ndarray
    theta: Any


def _flat_theta(theta, batch_shape, dtype):
    leaves = list(theta) if isinstance(theta, (tuple, list)) else [theta]
    nb = len(batch_shape)
    flat = [np.asarray(z, dtype).reshape(tuple(np.asarray(z).shape[:nb]) + (-1,)) for z in leaves]
    return np.ascontiguousarray(np.concatenate(flat, axis=-1)), [np.asarray(z).shape[nb:] for z in leaves]


def _unflat_theta(theta_like, flat, shapes, batch_shape):
    out, o = [], 0
    for sh in shapes:
        k = int(np.prod(sh)) if len(sh) else 1
        out.append(flat[..., o:o + k].reshape(tuple(batch_shape) + tuple(sh)))
        o += k
    if isinstance(theta_like, tuple) and hasattr(theta_like, "_fields"):
        return type(theta_like)(*out)
    if isinstance(theta_like, (tuple, list)):
        return type(theta_like)(out)
    return out[0]


def build(cs: Problem, *, maxiter: int = 100, thres: float = 1e-8, line_search: Optional[LineSearch] = None,
          unroll: bool = False, jit: bool = True, verbose: bool = False, systems: int = 1,
          control_limits=None) -> typs.Solver:
    """Build iLQR solver (idoc/ilqr.py:179-278).  `unroll`, `jit`, `verbose` are JAX tracing knobs: accepted, ignored.

    `control_limits=(lo, hi)` (scalars or arrays broadcastable to the batch shape + (m,)) solves the control-limited
    problem lo <= u_t <= hi (systems with a fused kernel; per-step box QP, a clamp when m = 1): an extension with no
    reference counterpart (BASELINE config 4); the implicit gradients treat the steps whose control sits on a bound as an
    active set.  With `systems` > 1 (bilqr) the limits act on the SHARED control: idoc_bilqr_solve_box / _vjp_box, for the
    (family, n, m, systems) combinations compiled in csrc/bilqr_fused.cu (which also serve the unlimited solve, fused)."""
    if cs.cost is not None or cs.costf is not None or cs.dynamics is not None:
        raise L.IdocError("idoc_b200.ilqr.Problem takes a compiled `family`, not Python callables (see module docstring)")
    if cs.family not in FAMILIES:
        raise L.IdocError(f"unknown problem family {cs.family!r}; compiled: {sorted(FAMILIES)}")
    pid, T, nsys, m = FAMILIES[cs.family], cs.horizon, cs.state_dim, cs.control_dim
    IL = PLUGINS.get(cs.family, L.lib)  # the library holding this family's idoc_ilqr_* entry points
    chk = lambda rc, what: L.check_with(IL, rc, what)
    n = nsys * systems  # lifted state (systems > 1: shared-control batch, idoc/bilqr.py)
    ls = line_search

    lifted_fused = systems > 1 and IL is L.lib and bool(L.lib.idoc_bilqr_box_supported(pid, nsys, m, systems))
    if control_limits is not None and systems != 1 and not lifted_fused:
        raise L.IdocError(f"control_limits with systems={systems}: no fused shared-control kernel compiled for family "
                          f"{cs.family!r}, n={nsys}, m={m} (csrc/bilqr_fused.cu lists them)")

    def use_lifted():  # IDOC_ILQR_FUSED=0 forces the multi-kernel loop where it can serve the call (no limits)
        return lifted_fused and (control_limits is not None or os.environ.get("IDOC_ILQR_FUSED", "1") != "0")

    def _limits(c):
        lo, hi = control_limits
        shp = tuple(c["batch"]) + (m,)
        mk = lambda z: L.DeviceArray.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(z, c["dt"]), shp)).reshape(c["B"], m))
        return mk(lo), mk(hi)

    def _setup(params: Params):
        L.require_device()
        x0 = np.asarray(params.x0)
        dt = x0.dtype if x0.dtype in (np.float32, np.float64) else np.dtype(np.float64)
        batch = x0.shape[:-1]
        B = int(np.prod(batch)) if batch else 1
        th, shapes = _flat_theta(params.theta, batch, dt)
        th = th.reshape(B, -1)
        tdim = IL.idoc_ilqr_theta_dim(pid, nsys, m)
        if th.shape[1] != tdim:
            raise L.IdocError(f"theta has {th.shape[1]} scalars per problem, family {cs.family!r} needs {tdim}")
        code = L.dtype_code(dt)
        return dict(dt=dt, batch=batch, B=B, code=code, tdim=tdim, shapes=shapes,
                    theta=L.DeviceArray.from_numpy(th), x0=L.DeviceArray.from_numpy(x0.reshape(B, n).astype(dt)))

    def _dev_state(c, s: typs.State):
        dt, B = c["dt"], c["B"]
        D = lambda z, k: L.DeviceArray.from_numpy(np.asarray(z, dt).reshape(B, T, k))
        return D(s.X, n), D(s.U, m), D(s.Nu, n)

    def _un(c, arr, k):
        return arr.numpy().reshape(tuple(c["batch"]) + (T, k))

    def ilqr(init: typs.State, params: Params, return_info: bool = False):
        c = _setup(params)
        X, U, Nu = _dev_state(c, init)
        ws_bytes = IL.idoc_ilqr_ws_bytes(c["code"], c["B"], T, nsys, m, systems)
        ws = L.DeviceArray((ws_bytes,), np.uint8)
        iters = L.DeviceArray((c["B"],), np.int32)
        lsf = L.DeviceArray((c["B"],), np.int32)
        use_ls = 0 if ls is None else 1
        lsp = ls if ls is not None else LineSearch()
        if use_lifted():
            lo, hi = _limits(c) if control_limits is not None else (None, None)
            L.check(L.lib.idoc_bilqr_solve_box(None, c["code"], pid, c["B"], T, nsys, m, systems, c["theta"].ptr, c["tdim"],
                                               c["x0"].ptr, L.ptr(lo), L.ptr(hi), X.ptr, U.ptr, Nu.ptr, maxiter, float(thres),
                                               use_ls, float(lsp.lower), float(lsp.upper), float(lsp.rho), int(lsp.maxiter),
                                               iters.ptr, lsf.ptr, ws.ptr, ws_bytes), "bilqr_solve_box")
        elif control_limits is None:
            chk(IL.idoc_ilqr_solve(None, c["code"], pid, c["B"], T, nsys, m, systems, c["theta"].ptr, c["tdim"], c["x0"].ptr,
                                          X.ptr, U.ptr, Nu.ptr, maxiter, float(thres), use_ls, float(lsp.lower), float(lsp.upper),
                                          float(lsp.rho), int(lsp.maxiter), iters.ptr, lsf.ptr, ws.ptr, ws_bytes, 1), "ilqr_solve")
        else:
            lo, hi = _limits(c)
            chk(IL.idoc_ilqr_solve_box(None, c["code"], pid, c["B"], T, nsys, m, c["theta"].ptr, c["tdim"], c["x0"].ptr,
                                              lo.ptr, hi.ptr, X.ptr, U.ptr, Nu.ptr, maxiter, float(thres), use_ls,
                                              float(lsp.lower), float(lsp.upper), float(lsp.rho), int(lsp.maxiter), iters.ptr,
                                              lsf.ptr, ws.ptr, ws_bytes), "ilqr_solve_box")
        s = typs.State(_un(c, X, n), _un(c, U, m), _un(c, Nu, n))
        if return_info:
            return s, dict(iterations=iters.numpy().reshape(c["batch"]), line_search_failures=lsf.numpy().reshape(c["batch"]))
        return s

    def _leaves(c, X, U, Nu, mode):
        B, dt = c["B"], c["dt"]
        shp = _lqr.leaf_shapes(B, T, n, m)
        lv = {k: L.DeviceArray(shp[k], dt) for k in _lqr.LEAVES}
        chk(IL.idoc_ilqr_linearize(None, c["code"], pid, B, T, nsys, m, systems, c["theta"].ptr, c["tdim"], c["x0"].ptr, X.ptr,
                                          U.ptr, L.ptr(Nu), mode, *[lv[k].ptr for k in _lqr.LEAVES]), "ilqr_linearize")
        return lv

    def kkt(s: typs.State, params: Params) -> typs.State:
        """KKT residual of the nonlinear problem = lqr.kkt on the global approximation (idoc/ilqr.py:192-194)."""
        c = _setup(params)
        X, U, Nu = _dev_state(c, s)
        lv = _leaves(c, X, U, None, 1)
        B, dt = c["B"], c["dt"]
        rX, rU, rNu = L.DeviceArray((B, T, n), dt), L.DeviceArray((B, T, m), dt), L.DeviceArray((B, T, n), dt)
        L.check(L.lib.idoc_lqr_kkt(None, c["code"], B, T, n, m, *[lv[k].ptr for k in _lqr.LEAVES], c["x0"].ptr,
                                   X.ptr, U.ptr, Nu.ptr, rX.ptr, rU.ptr, rNu.ptr), "lqr_kkt")
        return typs.State(_un(c, rX, n), _un(c, rU, m), _un(c, rNu, n))

    def lqr_approx(s: typs.State, params: Params, local: bool = True) -> _lqr.LQR:
        """make_lqr_approx(cs, params, local)(X, U) (idoc/ilqr.py:127-158)."""
        c = _setup(params)
        X, U, Nu = _dev_state(c, s)
        lv = _leaves(c, X, U, None, 0 if local else 1)
        tails = dict(Q=3, q=2, Qf=2, qf=1, M=3, R=3, r=2, A=3, B=3, d=2)
        return _lqr.LQR(*[lv[k].numpy().reshape(tuple(c["batch"]) + lv[k].shape[-tails[k]:]) for k in _lqr.LEAVES])

    def vjp(init: typs.State, params: Params):
        s = ilqr(init, params)
        c = _setup(params)
        X, U, Nu = _dev_state(c, s)

        def pullback(ct: typs.State) -> Params:
            dt, B = c["dt"], c["B"]
            Xb = L.DeviceArray.from_numpy(np.asarray(ct.X, dt).reshape(B, T, n))
            Ub = L.DeviceArray.from_numpy(np.asarray(ct.U, dt).reshape(B, T, m))
            Nub = None if ct.Nu is None else L.DeviceArray.from_numpy(np.asarray(ct.Nu, dt).reshape(B, T, n))
            gth, gx0 = L.DeviceArray((B, c["tdim"]), dt), L.DeviceArray((B, n), dt)
            wsb = IL.idoc_ilqr_vjp_ws_bytes(c["code"], B, T, nsys, m, systems)
            ws = L.DeviceArray((wsb,), np.uint8)
            if use_lifted():
                lo, hi = _limits(c) if control_limits is not None else (None, None)
                L.check(L.lib.idoc_bilqr_vjp_box(None, c["code"], pid, B, T, nsys, m, systems, c["theta"].ptr, c["tdim"],
                                                 c["x0"].ptr, L.ptr(lo), L.ptr(hi), X.ptr, U.ptr, Nu.ptr, Xb.ptr, Ub.ptr,
                                                 L.ptr(Nub), gth.ptr, gx0.ptr, ws.ptr, wsb), "bilqr_vjp_box")
            elif control_limits is None:
                chk(IL.idoc_ilqr_vjp(None, c["code"], pid, B, T, nsys, m, systems, c["theta"].ptr, c["tdim"], c["x0"].ptr, X.ptr,
                                            U.ptr, Nu.ptr, Xb.ptr, Ub.ptr, L.ptr(Nub), gth.ptr, gx0.ptr, ws.ptr, wsb), "ilqr_vjp")
            else:
                lo, hi = _limits(c)
                chk(IL.idoc_ilqr_vjp_box(None, c["code"], pid, B, T, nsys, m, c["theta"].ptr, c["tdim"], c["x0"].ptr, lo.ptr,
                                                hi.ptr, X.ptr, U.ptr, Nu.ptr, Xb.ptr, Ub.ptr, L.ptr(Nub), gth.ptr, gx0.ptr, ws.ptr,
                                                wsb), "ilqr_vjp_box")
            g = gth.numpy().reshape(tuple(c["batch"]) + (c["tdim"],))
            if not np.isfinite(g).all():  # the implicit-function gradient exists at a stationary point of a well-posed problem only
                import warnings
                bad = int((~np.isfinite(g.reshape(-1, c["tdim"])).all(axis=1)).sum())
                warnings.warn(f"ilqr implicit VJP: non-finite parameter gradient for {bad} of {c['B']} problem(s) (the solve did not "
                              "reach a stationary point, or the Hessian of the Lagrangian there is indefinite)", L.IdocNumericalWarning,
                              stacklevel=2)
            return Params(gx0.numpy().reshape(tuple(c["batch"]) + (n,)), _unflat_theta(params.theta, g, c["shapes"], c["batch"]))

        return s, pullback

    solver = typs.Solver(direct=ilqr, implicit=ilqr, kkt=kkt, vjp=vjp)
    solver.lqr_approx = lqr_approx
    return solver
