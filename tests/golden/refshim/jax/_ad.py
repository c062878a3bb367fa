"""Forward-mode dual numbers on NumPy arrays: enough of grad/jacfwd/jacobian for the
reference's `make_lqr_approx` (idoc/ilqr.py:127-158) on the callables its tests use.
Nested differentiation (jacfwd(grad(f))) works through perturbation levels.
Test infrastructure only."""
import numpy as np

_LEVEL = [0]


def _lvl(x):
    return x.level if isinstance(x, Dual) else 0


def _val(x, level):
    return x.val if isinstance(x, Dual) and x.level == level else x


def _tan(x, level):
    if isinstance(x, Dual) and x.level == level:
        return x.tan
    return None


def _shape(x):
    return x.shape if isinstance(x, Dual) else np.shape(x)


def _zeros_like(x):
    if isinstance(x, Dual):
        return Dual(_zeros_like(x.val), _zeros_like(x.tan), x.level)
    return np.zeros(np.shape(x))


class Dual:
    __array_priority__ = 1000.0

    def __init__(self, val, tan, level):
        self.val, self.tan, self.level = val, tan, level

    # ---- structure ----
    @property
    def shape(self):
        return _shape(self.val)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def T(self):
        return Dual(self.val.T, self.tan.T, self.level)

    def transpose(self, *axes):
        return Dual(self.val.transpose(*axes), self.tan.transpose(*axes), self.level)

    def reshape(self, *s):
        return Dual(self.val.reshape(*s), self.tan.reshape(*s), self.level)

    def __getitem__(self, idx):
        return Dual(self.val[idx], self.tan[idx], self.level)

    def _dual_take(self, i, ax):
        return Dual(_take(self.val, i, ax), _take(self.tan, i, ax), self.level)

    def sum(self, axis=None):
        return Dual(_sum(self.val, axis), _sum(self.tan, axis), self.level)

    def __len__(self):
        return self.shape[0]

    # ---- arithmetic ----
    def __neg__(self):
        return Dual(-self.val, -self.tan, self.level)

    def __add__(self, o):
        return _binary(self, o, lambda a, b: a + b, lambda a, b, ta, tb: _add(ta, tb))

    __radd__ = __add__

    def __sub__(self, o):
        return _binary(self, o, lambda a, b: a - b, lambda a, b, ta, tb: _add(ta, None if tb is None else -tb))

    def __rsub__(self, o):
        return _binary(o, self, lambda a, b: a - b, lambda a, b, ta, tb: _add(ta, None if tb is None else -tb))

    def __mul__(self, o):
        return _binary(self, o, lambda a, b: a * b,
                       lambda a, b, ta, tb: _add(None if ta is None else ta * b, None if tb is None else a * tb))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _binary(self, o, lambda a, b: a / b,
                       lambda a, b, ta, tb: _add(None if ta is None else ta / b,
                                                 None if tb is None else -(a / (b * b)) * tb))

    def __rtruediv__(self, o):
        return _binary(o, self, lambda a, b: a / b,
                       lambda a, b, ta, tb: _add(None if ta is None else ta / b,
                                                 None if tb is None else -(a / (b * b)) * tb))

    def __pow__(self, p):
        assert not isinstance(p, Dual)
        return _unary(self, lambda a: a ** p, lambda a: p * a ** (p - 1))

    def __matmul__(self, o):
        return _binary(self, o, lambda a, b: a @ b,
                       lambda a, b, ta, tb: _add(None if ta is None else ta @ b, None if tb is None else a @ tb))

    def __rmatmul__(self, o):
        return _binary(o, self, lambda a, b: a @ b,
                       lambda a, b, ta, tb: _add(None if ta is None else ta @ b, None if tb is None else a @ tb))

    # ---- numpy protocols ----
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        f = _UFUNCS.get(ufunc)
        if f is None:
            raise NotImplementedError(f"refshim dual: ufunc {ufunc}")
        return f(*inputs)

    def __array_function__(self, func, types, args, kwargs):
        f = _FUNCS.get(func)
        if f is None:
            raise NotImplementedError(f"refshim dual: function {func}")
        return f(*args, **kwargs)


def _take(x, i, ax):
    if isinstance(x, Dual):
        return x._dual_take(i, ax)
    return np.take(x, i, axis=ax)


def _sum(x, axis=None):
    if isinstance(x, Dual):
        return x.sum(axis)
    return np.sum(x, axis=axis)


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a + b


def _binary(a, b, f, jvp):
    level = max(_lvl(a), _lvl(b))
    av, bv = _val(a, level), _val(b, level)
    ta, tb = _tan(a, level), _tan(b, level)
    out = f(av, bv)
    t = jvp(av, bv, ta, tb)
    if t is None:
        t = _zeros_like(out)
    return Dual(out, t, level)


def _unary(a, f, df):
    return Dual(f(a.val), df(a.val) * a.tan, a.level)


def _maximum(a, b):
    level = max(_lvl(a), _lvl(b))
    av, bv = _val(a, level), _val(b, level)
    ta, tb = _tan(a, level), _tan(b, level)
    pa, pb = _primal(av), _primal(bv)
    mask = pa >= pb
    out = _where(mask, av, bv)
    ta = _zeros_like(out) if ta is None else ta
    tb = _zeros_like(out) if tb is None else tb
    return Dual(out, _where(mask, ta, tb), level)


def _primal(x):
    while isinstance(x, Dual):
        x = x.val
    return np.asarray(x)


def _where(mask, a, b):
    if isinstance(a, Dual) or isinstance(b, Dual):
        level = max(_lvl(a), _lvl(b))
        av, bv = _val(a, level), _val(b, level)
        ta, tb = _tan(a, level), _tan(b, level)
        out = _where(mask, av, bv)
        ta = _zeros_like(out) if ta is None else ta
        tb = _zeros_like(out) if tb is None else tb
        return Dual(out, _where(mask, ta, tb), level)
    return np.where(mask, a, b)


def _dot(a, b):
    # the reference only uses 1-D·1-D, 2-D·1-D and 2-D·2-D dots, for which dot == matmul
    if not isinstance(a, Dual) and not isinstance(b, Dual):
        return np.dot(a, b)
    return _binary(a, b, lambda x, y: _dot(x, y) if (isinstance(x, Dual) or isinstance(y, Dual)) else np.dot(x, y),
                   lambda x, y, tx, ty: _add(None if tx is None else _dot(tx, y), None if ty is None else _dot(x, ty)))


def _matmul(a, b):
    if isinstance(a, Dual):
        return a.__matmul__(b)
    return b.__rmatmul__(a)


def _einsum(spec, a, b):
    """two-operand einsum (bilinear)"""
    if not isinstance(a, Dual) and not isinstance(b, Dual):
        return np.einsum(spec, a, b)
    return _binary(a, b, lambda x, y: _einsum(spec, x, y),
                   lambda x, y, tx, ty: _add(None if tx is None else _einsum(spec, tx, y),
                                             None if ty is None else _einsum(spec, x, ty)))


def dual_stack(xs, axis=0):
    level = max(_lvl(x) for x in xs)
    vals = [_val(x, level) for x in xs]
    tans = [_tan(x, level) if _tan(x, level) is not None else _zeros_like(_val(x, level)) for x in xs]
    st = (lambda zs: dual_stack(zs, axis) if any(isinstance(z, Dual) for z in zs) else np.stack(zs, axis=axis))
    return Dual(st(vals), st(tans), level)


def _concatenate(xs, axis=0):
    level = max(_lvl(x) for x in xs)
    vals = [_val(x, level) for x in xs]
    tans = [_tan(x, level) if _tan(x, level) is not None else _zeros_like(_val(x, level)) for x in xs]
    cc = (lambda zs: _concatenate(zs, axis) if any(isinstance(z, Dual) for z in zs) else np.concatenate(zs, axis=axis))
    return Dual(cc(vals), cc(tans), level)


_UFUNCS = {
    np.add: lambda a, b: Dual.__add__(a, b) if isinstance(a, Dual) else Dual.__radd__(b, a),
    np.subtract: lambda a, b: Dual.__sub__(a, b) if isinstance(a, Dual) else Dual.__rsub__(b, a),
    np.multiply: lambda a, b: Dual.__mul__(a, b) if isinstance(a, Dual) else Dual.__rmul__(b, a),
    np.true_divide: lambda a, b: Dual.__truediv__(a, b) if isinstance(a, Dual) else Dual.__rtruediv__(b, a),
    np.negative: lambda a: -a,
    np.matmul: _matmul,
    np.maximum: _maximum,
    np.sin: lambda a: _unary(a, np.sin if not isinstance(a.val, Dual) else _usin, np.cos if not isinstance(a.val, Dual) else _ucos),
    np.cos: lambda a: _unary(a, np.cos if not isinstance(a.val, Dual) else _ucos,
                             (lambda z: -np.sin(z)) if not isinstance(a.val, Dual) else (lambda z: -_usin(z))),
    np.exp: lambda a: _unary(a, np.exp if not isinstance(a.val, Dual) else _uexp, np.exp if not isinstance(a.val, Dual) else _uexp),
    np.tanh: lambda a: _unary(a, _utanh, lambda z: 1.0 - _utanh(z) * _utanh(z)),
    np.square: lambda a: a * a,
}


def _usin(z):
    return np.sin(z)


def _ucos(z):
    return np.cos(z)


def _uexp(z):
    return np.exp(z)


def _utanh(z):
    return np.tanh(z)


_FUNCS = {
    np.dot: _dot,
    np.einsum: _einsum,
    np.sum: lambda a, axis=None, **k: _sum(a, axis),
    np.where: _where,
    np.stack: lambda xs, axis=0: dual_stack(list(xs), axis),
    np.concatenate: lambda xs, axis=0: _concatenate(list(xs), axis),
    np.transpose: lambda a, axes=None: a.transpose(*(axes or ())),
    np.shape: lambda a: a.shape,
    np.zeros_like: lambda a, **k: np.zeros(a.shape),
    np.ones_like: lambda a, **k: np.ones(a.shape),
}


# ------------------------------------------------------------------ transforms
def _jac_one(f, argnum, args):
    x = args[argnum]
    xs = _shape(x)
    size = int(np.prod(xs)) if len(xs) else 1
    _LEVEL[0] += 1
    level = _LEVEL[0]
    cols = []
    out_proto = None
    try:
        for i in range(size):
            e = np.zeros(size)
            e[i] = 1.0
            e = e.reshape(xs)
            call = list(args)
            call[argnum] = Dual(x, e, level)
            out = f(*call)
            if isinstance(out, Dual) and out.level == level:
                cols.append(out.tan)
                out_proto = out.val
            else:
                cols.append(None)
                out_proto = out
    finally:
        _LEVEL[0] -= 1
    cols = [c if c is not None else _zeros_like(out_proto) for c in cols]
    if any(isinstance(c, Dual) for c in cols):
        st = dual_stack(cols, axis=-1) if _shape(cols[0]) != () else dual_stack(cols, axis=0)
        return st.reshape(tuple(_shape(out_proto)) + tuple(xs))
    st = np.stack([np.asarray(c, dtype=float) for c in cols], axis=-1)
    return st.reshape(tuple(np.shape(out_proto)) + tuple(xs))


def jacfwd(f, argnums=0):
    def jf(*args):
        if isinstance(argnums, (tuple, list)):
            return tuple(_jac_one(f, a, args) for a in argnums)
        return _jac_one(f, argnums, args)

    return jf


jacobian = jacfwd
jacrev = jacfwd
grad = jacfwd


def value_and_grad(f, argnums=0):
    g = grad(f, argnums)
    return lambda *a: (f(*a), g(*a))
