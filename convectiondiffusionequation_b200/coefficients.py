"""Coefficient functions of the driver (run.py:24-30) without NGSolve.

The reference builds `exact`, `coeff` and the enrichment functions as NGSolve CoefficientFunction
trees from `x`, `y`, `exp`.  The kernels need them in closed form, so this module offers the one
family run.py uses -- the exponential boundary-layer profile

    L(s) = s + (exp(beta (s-1)/eps) - exp(-beta/eps)) / (exp(-beta/eps) - 1)        run.py:26-27

of one coordinate, closed under the operations run.py applies (scalar multiples, sums, the
product of an x-profile with a y-profile).  Anything else raises: there is no silent fallback.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Coordinate:
    axis: int


x = Coordinate(0)
y = Coordinate(1)


class LayerExpr:
    """c0 + c1 Lx + c2 Ly + c3 Lx Ly with Lx = L(kx; x), Ly = L(ky; y)."""

    def __init__(self, c=(0.0, 0.0, 0.0, 0.0), kx=None, ky=None):
        self.c = tuple(float(v) for v in c)
        self.kx = kx
        self.ky = ky

    # -- algebra ---------------------------------------------------------------------------
    @staticmethod
    def _merge(a, b, name):
        if a is None:
            return b
        if b is None or a == b:
            return a
        raise ValueError(f"two different layer steepnesses along {name} in one expression are not supported")

    def __add__(self, other):
        if isinstance(other, (int, float)):
            return LayerExpr((self.c[0] + other,) + self.c[1:], self.kx, self.ky)
        if isinstance(other, LayerExpr):
            return LayerExpr(tuple(a + b for a, b in zip(self.c, other.c)),
                             self._merge(self.kx, other.kx, "x"), self._merge(self.ky, other.ky, "y"))
        return NotImplemented

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return LayerExpr(tuple(v * other for v in self.c), self.kx, self.ky)
        if isinstance(other, LayerExpr):
            a, b = self.c, other.c
            if a[3] or b[3] or (a[1] and b[1]) or (a[2] and b[2]):
                raise ValueError("only products of an x-profile with a y-profile are supported")
            c = (a[0] * b[0], a[0] * b[1] + a[1] * b[0], a[0] * b[2] + a[2] * b[0], a[1] * b[2] + a[2] * b[1])
            return LayerExpr(c, self._merge(self.kx, other.kx, "x"), self._merge(self.ky, other.ky, "y"))
        return NotImplemented

    __rmul__ = __mul__

    # -- description for the C ABI ------------------------------------------------------------
    def desc(self):
        return self.c, (0.0 if self.kx is None else float(self.kx), 0.0 if self.ky is None else float(self.ky))

    def single_axis(self):
        """(axis, k) if the expression is exactly one profile L(k; x) or L(k; y) (an enrichment function)."""
        if self.c == (0.0, 1.0, 0.0, 0.0):
            return 0, float(self.kx)
        if self.c == (0.0, 0.0, 1.0, 0.0):
            return 1, float(self.ky)
        raise ValueError("enrichment functions must be a single boundary-layer profile of x or of y")

    # -- host evaluation (plots, checks; not used by the kernels) ---------------------------------
    def __call__(self, xs, ys):
        def L(k, s):
            if k is None:
                return np.zeros_like(s)
            e0 = np.exp(-k)
            return s + (np.exp(-k * (1.0 - s)) - e0) / (e0 - 1.0)
        lx, ly = L(self.kx, np.asarray(xs, float)), L(self.ky, np.asarray(ys, float))
        return self.c[0] + self.c[1] * lx + self.c[2] * ly + self.c[3] * lx * ly


def boundary_layer(coord: Coordinate, beta: float, eps: float) -> LayerExpr:
    """p(x) / q(y) of run.py:26-27."""
    k = float(beta) / float(eps)
    if coord.axis == 0:
        return LayerExpr((0.0, 1.0, 0.0, 0.0), kx=k)
    return LayerExpr((0.0, 0.0, 1.0, 0.0), ky=k)
