import numpy as _np


def _is_leaf(x):
    return not isinstance(x, (tuple, list, dict))


def tree_leaves(t):
    if t is None:
        return []
    if _is_leaf(t):
        return [t]
    if isinstance(t, dict):
        out = []
        for k in sorted(t):
            out += tree_leaves(t[k])
        return out
    out = []
    for c in t:
        out += tree_leaves(c)
    return out


def tree_map(f, t, *rest):
    if t is None:
        return None
    if _is_leaf(t):
        return f(t, *rest)
    if isinstance(t, dict):
        return {k: tree_map(f, t[k], *[r[k] for r in rest]) for k in t}
    cols = [tree_map(f, c, *[r[i] for r in rest]) for i, c in enumerate(t)]
    if hasattr(t, "_fields"):
        return type(t)(*cols)
    return type(t)(cols)


def tree_unflatten_like(t, leaves):
    it = iter(leaves)
    return tree_map(lambda _: next(it), t)
