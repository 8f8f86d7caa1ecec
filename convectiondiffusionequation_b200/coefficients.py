"""Coefficient functions of the driver (run.py:24-30) without NGSolve.

The reference builds `exact`, `coeff` and the enrichment functions as NGSolve CoefficientFunction
trees from `x`, `y`, `exp`.  The kernels need them in closed form, so this module offers the one
family run.py uses -- the exponential boundary-layer profile

    L(s) = s + (exp(beta (s-1)/eps) - exp(-beta/eps)) / (exp(-beta/eps) - 1)        run.py:26-27

of one coordinate, closed under the operations run.py applies (scalar multiples, sums, the
product of an x-profile with a y-profile): the fast path of the kernels, evaluated through the
cancellation-free distances 1 - x, 1 - y.

Anything else -- `coeff` and `exact` as ARBITRARY expressions of x and y, like the reference's
CoefficientFunctions -- goes through `SymExpr`: a sympy expression (or anything sympify accepts, in the
symbols `sx`, `sy` exported here) compiled to the postfix program that the kernels interpret
(csrc/ehdg_expr.h, ehdg_set_expression).  The enrichment functions stay in the layer family.
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


# ------------------------------------------------------------------------------------------------------------
# generic expressions: sympy -> postfix program of csrc/ehdg_expr.h
# ------------------------------------------------------------------------------------------------------------
OPS = dict(CONST=1, X=2, Y=3, ADD=4, SUB=5, MUL=6, DIV=7, NEG=8, EXP=9, LOG=10, SIN=11, COS=12, TAN=13, SQRT=14, POWI=15, POW=16,
           TANH=17, ABS=18, ATAN=19, SIGN=20)
EXPR_MAX = 128


def _symbols():
    import sympy
    return sympy.Symbol("x", real=True), sympy.Symbol("y", real=True)


class SymExpr:
    """A coefficient function given as a sympy expression of the symbols `sx`, `sy` (x and y of the unit square).
    `program()` is what the C ABI takes; `__call__` evaluates on the host with numpy (oracle, plots);
    `diff(axis)` is the derivative as a new SymExpr (the reference's `.Diff(x)`, enrichment_proxy.py:18-19)."""

    def __init__(self, expr):
        import sympy
        sx_, sy_ = _symbols()
        self.expr = sympy.sympify(expr, locals={"x": sx_, "y": sy_})
        extra = self.expr.free_symbols - {sx_, sy_}
        if extra:
            raise ValueError(f"unknown symbols in the coefficient expression: {sorted(map(str, extra))}")
        self._fn = None

    def diff(self, axis: int) -> "SymExpr":
        import sympy
        return SymExpr(sympy.diff(self.expr, _symbols()[axis]))

    def __call__(self, xs, ys):
        import sympy
        if self._fn is None:
            self._fn = sympy.lambdify(_symbols(), self.expr, "numpy")
        return np.broadcast_to(np.asarray(self._fn(np.asarray(xs, float), np.asarray(ys, float)), dtype=float), np.shape(xs)).copy()

    def program(self):
        """(ops i32[n], args f64[n]) in postfix order"""
        import sympy
        sx_, sy_ = _symbols()
        ops, args = [], []

        def emit(op, arg=0.0):
            ops.append(OPS[op])
            args.append(float(arg))

        def walk(e):
            if e == sx_:
                return emit("X")
            if e == sy_:
                return emit("Y")
            if e.is_Number or e.is_NumberSymbol:
                return emit("CONST", float(e))
            if e.is_Add:
                terms = e.as_ordered_terms()
                walk(terms[0])
                for t in terms[1:]:
                    if t.could_extract_minus_sign():
                        walk(-t)
                        emit("SUB")
                    else:
                        walk(t)
                        emit("ADD")
                return
            if e.is_Mul:
                num, den = e.as_numer_denom()
                if den != 1:
                    walk(num)
                    walk(den)
                    return emit("DIV")
                coeff, rest = e.as_coeff_Mul()
                if coeff == -1:
                    walk(rest)
                    return emit("NEG")
                fs = e.as_ordered_factors()
                walk(fs[0])
                for f in fs[1:]:
                    walk(f)
                    emit("MUL")
                return
            if e.is_Pow:
                b, ex = e.as_base_exp()
                if ex.is_Integer and abs(int(ex)) <= 64:
                    walk(b)
                    return emit("POWI", int(ex))
                if ex == sympy.Rational(1, 2):
                    walk(b)
                    return emit("SQRT")
                walk(b)
                walk(ex)
                return emit("POW")
            one = {sympy.exp: "EXP", sympy.log: "LOG", sympy.sin: "SIN", sympy.cos: "COS", sympy.tan: "TAN", sympy.tanh: "TANH",
                   sympy.Abs: "ABS", sympy.atan: "ATAN", sympy.sign: "SIGN"}
            for fn, op in one.items():
                if isinstance(e, fn):
                    walk(e.args[0])
                    return emit(op)
            if isinstance(e, sympy.sinh) or isinstance(e, sympy.cosh):
                return walk(e.rewrite(sympy.exp))
            raise ValueError(f"unsupported construct in a coefficient expression: {type(e).__name__} in {e}")

        walk(self.expr)
        if len(ops) > EXPR_MAX:
            raise ValueError(f"the expression compiles to {len(ops)} instructions, the kernels take {EXPR_MAX}")
        return np.asarray(ops, dtype=np.int32), np.asarray(args, dtype=np.float64)


def interpret(ops, args, xs, ys):
    """host twin of csrc/ehdg_expr.h::expr_eval (numpy, vectorised over points): used by the CPU suite to check the compiler"""
    xs, ys = np.asarray(xs, float), np.asarray(ys, float)
    inv = {v: k for k, v in OPS.items()}
    st = []
    for op, a in zip(ops, args):
        name = inv[int(op)]
        if name == "CONST":
            st.append(np.full_like(xs, a))
        elif name == "X":
            st.append(xs.copy())
        elif name == "Y":
            st.append(ys.copy())
        elif name in ("ADD", "SUB", "MUL", "DIV", "POW"):
            b = st.pop()
            t = st.pop()
            st.append({"ADD": t + b, "SUB": t - b, "MUL": t * b, "DIV": t / b, "POW": np.power(t, b)}[name])
        elif name == "NEG":
            st.append(-st.pop())
        elif name == "POWI":
            st.append(st.pop() ** int(a))
        else:
            f = dict(EXP=np.exp, LOG=np.log, SIN=np.sin, COS=np.cos, TAN=np.tan, SQRT=np.sqrt, TANH=np.tanh, ABS=np.abs, ATAN=np.arctan, SIGN=np.sign)[name]
            st.append(f(st.pop()))
    assert len(st) == 1
    return st[0]
