"""jax.numpy stand-in: plain NumPy (fp64)."""
from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import linalg, ndarray  # noqa: F401

abs = _np.abs  # noqa: A001
sum = _np.sum  # noqa: A001
max = _np.max  # noqa: A001
min = _np.min  # noqa: A001
