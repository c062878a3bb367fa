"""jax.lax stand-in: scan as a Python loop over the leading axis."""
import numpy as _np
from ._tree import tree_map, tree_leaves


def scan(f, init, xs, length=None):
    import jax

    n = tree_leaves(xs)[0].shape[0] if xs is not None else length
    carry = init
    ys = []
    for i in range(n):
        x = tree_map(lambda z: z[i], xs) if xs is not None else None
        carry, y = f(carry, x)
        ys.append(y)
    return carry, jax._stack_trees(ys)
