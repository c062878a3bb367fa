import numpy as _np
from ._tree import tree_leaves, tree_unflatten_like


def ravel_pytree(t):
    leaves = [_np.asarray(l, dtype=float) for l in tree_leaves(t)]
    shapes = [l.shape for l in leaves]
    sizes = [l.size for l in leaves]
    flat = _np.concatenate([l.ravel() for l in leaves]) if leaves else _np.zeros(0)

    def unravel(v):
        out, o = [], 0
        for sh, sz in zip(shapes, sizes):
            out.append(_np.asarray(v[o:o + sz]).reshape(sh))
            o += sz
        return tree_unflatten_like(t, out)

    return flat, unravel
