"""jax.random stand-in. NOT jax's threefry stream — a deterministic NumPy equivalent."""
import numpy as _np


def PRNGKey(seed):
    return _np.random.SeedSequence(int(seed))


def split(key, num=2):
    return tuple(key.spawn(num))


def normal(key, shape=(), dtype=None):
    return _np.random.default_rng(key).standard_normal(shape)
