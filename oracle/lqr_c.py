"""ctypes wrapper of the C oracle (oracle/lqr_c.c -> oracle/liblqr_oracle.so).  Test / CPU-baseline
infrastructure only (see the header of lqr_c.c)."""
import ctypes as C
import os

import numpy as np

from . import lqr_np as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = C.CDLL(os.path.join(_HERE, "liblqr_oracle.so"))
_P = C.c_void_p
for suf in ("f64", "f32"):
    f = getattr(_lib, "oracle_lqr_fwd_" + suf)
    f.restype = None
    f.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int] + [_P] * 16
    f = getattr(_lib, "oracle_lqr_vjp_" + suf)
    f.restype = None
    f.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int] + [_P] * 24
    getattr(_lib, "oracle_threads_" + suf).restype = C.c_int
    getattr(_lib, "oracle_set_threads_" + suf).argtypes = [C.c_int]
    getattr(_lib, "oracle_set_threads_" + suf).restype = None


def threads():
    return _lib.oracle_threads_f64()


def set_threads(n):
    """OpenMP threads of the following calls (both precisions share one runtime)."""
    _lib.oracle_set_threads_f64(int(n))


def _suf(dt):
    return "f64" if np.dtype(dt) == np.float64 else "f32"


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def fwd(p: O.Params, out=None):
    dt = np.asarray(p.lqr.A).dtype
    l = O.LQR(*[_c(a, dt) for a in p.lqr])
    x0 = _c(p.x0, dt)
    B, T, n = l.A.shape[0], l.A.shape[1], l.A.shape[2]
    m = l.B.shape[3]
    if out is None:
        out = (np.empty((B, T, n), dt), np.empty((B, T, m), dt), np.empty((B, T, n), dt),
               np.empty((B, T, m, n), dt), np.empty((B, T, m), dt))
    X, U, Nu, K, k = out
    getattr(_lib, "oracle_lqr_fwd_" + _suf(dt))(
        B, T, n, m, *[a.ctypes.data for a in (l.Q, l.q, l.Qf, l.qf, l.M, l.R, l.r, l.A, l.B, l.d, x0, X, U, Nu, K, k)])
    return O.State(X, U, Nu), K, k


def vjp(p: O.Params, s: O.State, w: O.State, out=None):
    dt = np.asarray(p.lqr.A).dtype
    l = O.LQR(*[_c(a, dt) for a in p.lqr])
    x0 = _c(p.x0, dt)
    B, T, n = l.A.shape[0], l.A.shape[1], l.A.shape[2]
    m = l.B.shape[3]
    if out is None:
        out = O.Params(np.empty_like(x0), O.LQR(*[np.empty_like(a) for a in l]))
    g = out
    Nub = None if w.Nu is None else _c(w.Nu, dt).ctypes.data
    args = [l.Q, l.Qf, l.M, l.R, l.A, l.B, x0, _c(s.X, dt), _c(s.U, dt), _c(s.Nu, dt), _c(w.X, dt), _c(w.U, dt)]
    getattr(_lib, "oracle_lqr_vjp_" + _suf(dt))(
        B, T, n, m, *[a.ctypes.data for a in args], Nub,
        *[a.ctypes.data for a in (g.lqr.Q, g.lqr.q, g.lqr.Qf, g.lqr.qf, g.lqr.M, g.lqr.R, g.lqr.r, g.lqr.A, g.lqr.B,
                                  g.lqr.d, g.x0)])
    return g
