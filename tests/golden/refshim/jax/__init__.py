"""NumPy stand-in for the slice of `jax` the reference idoc sources touch.
Test infrastructure (golden-vector generation) only — see ../README.md."""
import numpy as _np
from . import numpy  # noqa: F401  (jax.numpy)
from . import lax  # noqa: F401
from . import scipy  # noqa: F401
from . import random  # noqa: F401
from . import nn  # noqa: F401
from . import flatten_util  # noqa: F401
from ._tree import tree_map, tree_leaves, tree_unflatten_like
from ._ad import grad, jacfwd, jacobian, jacrev, value_and_grad  # noqa: F401


class Array:  # scipy's array-api-compat probes jax.Array when a module named jax is loaded
    pass


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


class _Ops:
    @staticmethod
    def index_update(x, idx, y):
        x = _np.array(x, copy=True)
        x[idx] = y
        return x


ops = _Ops()


def jit(f, *a, **k):
    return f


def vmap(f, in_axes=0, out_axes=0):
    """Python-loop vmap: maps over axis `in_axes` of each positional arg, stacks outputs."""

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        size = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaves = tree_leaves(a)
                size = leaves[0].shape[ax]
                break
        outs = []
        for i in range(size):
            call = []
            for a, ax in zip(args, axes):
                if ax is None:
                    call.append(a)
                else:
                    call.append(tree_map(lambda z: _np.take(z, i, axis=ax) if not hasattr(z, "_dual_take") else z._dual_take(i, ax), a))
            outs.append(f(*call))
        return _stack_trees(outs)

    return mapped


def _stack_trees(outs):
    first = outs[0]
    if isinstance(first, tuple):
        cols = [_stack_trees([o[j] for o in outs]) for j in range(len(first))]
        if hasattr(first, "_fields"):
            return type(first)(*cols)
        return tuple(cols)
    if isinstance(first, list):
        return [_stack_trees([o[j] for o in outs]) for j in range(len(first))]
    from ._ad import Dual, dual_stack

    if any(isinstance(o, Dual) for o in outs):
        return dual_stack(outs)
    return _np.stack([_np.asarray(o) for o in outs], axis=0)
