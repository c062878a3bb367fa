"""Differentiable batch iLQR with SHARED controls — host-side mirror of the reference module idoc/bilqr.py.

`batch_size` copies of one system (same theta, different initial states x0[B, n]) are driven by one control sequence
U[T, m]; the cost is summed over the copies (idoc/bilqr.py:93-117, 136-167).  On the device this is the iLQR loop
of csrc/ilqr.cuh with `systems = batch_size`: the kernels assemble the block-diagonal lift per timestep (the same
lift that idoc_b200.blqr uses for idoc/blqr.py) and run the Riccati sweep on it.  Shapes follow the reference:
x0[B, n], State(X[T, B, n], U[T, m], Nu[T, B, n])."""
from typing import Any, NamedTuple, Optional

import numpy as np

from . import ilqr as _ilqr
from . import typs
from .line_search import LineSearch

Problem = _ilqr.Problem  # idoc/bilqr.py:19-42 (same fields; `family` names the compiled cost/dynamics)


class Params(NamedTuple):
    """iLQR Params (idoc/bilqr.py:45-49): x0[B, n], theta shared by the B systems."""
    x0: np.ndarray
    theta: Any


def build(cs: Problem, *, maxiter: int = 100, thres: float = 1e-8, line_search: Optional[LineSearch] = None,
          unroll: bool = False, jit: bool = True, verbose: bool = False, control_limits=None) -> typs.Solver:
    """Build iLQR solver (idoc/bilqr.py:120-221).  `control_limits=(lo, hi)` (scalars or [m]) bounds the shared control
    lo <= u_t <= hi: no reference counterpart (see idoc_b200.ilqr.build); needs a fused shared-control kernel for the
    (family, n, m, batch_size) at hand (csrc/bilqr_fused.cu)."""
    T, n = cs.horizon, cs.state_dim
    cache = {}

    def _inner(bs):
        if bs not in cache:
            cache[bs] = _ilqr.build(cs, maxiter=maxiter, thres=thres, line_search=line_search, systems=bs,
                                    control_limits=control_limits)
        return cache[bs]

    def _lift_state(s: typs.State, bs):
        return typs.State(np.asarray(s.X).reshape(T, bs * n), np.asarray(s.U), np.asarray(s.Nu).reshape(T, bs * n))

    def _unlift_state(s: typs.State, bs):
        return typs.State(s.X.reshape(T, bs, n), s.U, s.Nu.reshape(T, bs, n))

    def _lp(params: Params):
        x0 = np.asarray(params.x0)
        return x0.shape[0], _ilqr.Params(x0.reshape(-1), params.theta)

    def ilqr(init: typs.State, params: Params, return_info: bool = False):
        bs, lp = _lp(params)
        out = _inner(bs).direct(_lift_state(init, bs), lp, return_info=return_info)
        if return_info:
            return _unlift_state(out[0], bs), out[1]
        return _unlift_state(out, bs)

    def kkt(s: typs.State, params: Params) -> typs.State:
        bs, lp = _lp(params)
        return _unlift_state(_inner(bs).kkt(_lift_state(s, bs), lp), bs)

    def vjp(init: typs.State, params: Params):
        bs, lp = _lp(params)
        s, pull = _inner(bs).vjp(_lift_state(init, bs), lp)

        def pullback(ct: typs.State) -> Params:
            fl = lambda z: None if z is None else np.asarray(z).reshape(T, bs * n)
            g = pull(typs.State(fl(ct.X), np.asarray(ct.U), fl(ct.Nu)))
            return Params(g.x0.reshape(bs, n), g.theta)

        return _unlift_state(s, bs), pullback

    return typs.Solver(direct=ilqr, implicit=ilqr, kkt=kkt, vjp=vjp)
