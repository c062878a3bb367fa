"""jax.scipy.linalg stand-in (SciPy LAPACK: gesv for general, posv for sym_pos)."""
import numpy as _np
import scipy.linalg as _sl

block_diag = _sl.block_diag


def solve(a, b, sym_pos=False, assume_a=None, **kw):
    a = _np.asarray(a)
    b = _np.asarray(b)
    if sym_pos or assume_a == "pos":
        return _sl.solve(a, b, assume_a="pos")
    return _sl.solve(a, b)
