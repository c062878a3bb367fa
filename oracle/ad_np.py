"""CPU ORACLE (test infrastructure, not product code) — second-order forward-mode jets in NumPy.

The reference linearises / quadratises arbitrary Python cost and dynamics with `jax.grad`, `jax.jacfwd` and
`jax.jacobian` (`/root/reference/idoc/ilqr.py:132-138`, `:152-153`).  jax is not installable here, so the oracle's
problem families that have no hand-written derivatives (pendulum, cart-pole, unicycle: BASELINE config 4) get theirs
from this small jet arithmetic: a `Jet` carries value, gradient and Hessian with respect to Z seeded inputs, exact to
rounding (no step size).  It is the NumPy counterpart of the device autodiff in `idoc_b200/csrc/ad.cuh`, written
independently (dense Hessians, operator overloading) so that the two can be compared.
"""
import numpy as np


class Jet:
    __slots__ = ("v", "g", "H")

    def __init__(self, v, g, H):
        self.v, self.g, self.H = float(v), g, H

    @staticmethod
    def seed(values):
        """Independent variables z_0..z_{Z-1}: Jet i has gradient e_i and zero Hessian."""
        Z = len(values)
        eye = np.eye(Z)
        return [Jet(values[i], eye[i].copy(), np.zeros((Z, Z))) for i in range(Z)]

    def _lift(self, o):
        return o if isinstance(o, Jet) else Jet(o, np.zeros_like(self.g), np.zeros_like(self.H))

    def _unary(self, f, f1, f2):
        return Jet(f, f1 * self.g, f1 * self.H + f2 * np.outer(self.g, self.g))

    def __add__(self, o):
        o = self._lift(o)
        return Jet(self.v + o.v, self.g + o.g, self.H + o.H)

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.g, -self.H)

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) + (-self)

    def __mul__(self, o):
        o = self._lift(o)
        og = np.outer(self.g, o.g)
        return Jet(self.v * o.v, self.v * o.g + o.v * self.g, self.v * o.H + o.v * self.H + og + og.T)

    __rmul__ = __mul__

    def recip(self):
        r = 1.0 / self.v
        return self._unary(r, -r * r, 2.0 * r * r * r)

    def __truediv__(self, o):
        return self * self._lift(o).recip()

    def __rtruediv__(self, o):
        return self._lift(o) * self.recip()


def sin(x):
    if isinstance(x, Jet):
        s, c = np.sin(x.v), np.cos(x.v)
        return x._unary(s, c, -s)
    return np.sin(x)


def cos(x):
    if isinstance(x, Jet):
        s, c = np.sin(x.v), np.cos(x.v)
        return x._unary(c, -s, -c)
    return np.cos(x)


def relu(x):
    if isinstance(x, Jet):
        on = 1.0 if x.v > 0.0 else 0.0
        return x._unary(on * x.v, on, 0.0)
    return np.maximum(x, 0.0)
