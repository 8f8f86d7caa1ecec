"""Generic coefficient expressions: the sympy -> postfix compiler against sympy's own evaluation, through the host
twin of the device interpreter (the device side is checked in tests/test_gpu_symexpr.py)."""
import numpy as np
import pytest

from convectiondiffusionequation_b200.coefficients import EXPR_MAX, SymExpr, interpret

EXPRS = ["sin(pi*x)*sin(pi*y)", "x**2*y - 3*x*y**3 + 2", "exp(-(x-0.3)**2/0.01) + y/(1+x)", "sqrt(x+1)*cos(3*y) - tanh(x-y)",
         "(x+y)**-2 + abs(x-0.5)", "1.5", "x", "-x*y", "log(2+x)*atan(y)", "2**x + y**2.5"]


@pytest.mark.parametrize("src", EXPRS)
def test_program_matches_sympy(src):
    e = SymExpr(src)
    ops, args = e.program()
    assert 0 < len(ops) <= EXPR_MAX
    rng = np.random.default_rng(0)
    xs, ys = rng.random(200), rng.random(200)
    assert np.allclose(interpret(ops, args, xs, ys), e(xs, ys), rtol=1e-13, atol=1e-13)
    for ax in (0, 1):
        d = e.diff(ax)
        o2, a2 = d.program()
        assert np.allclose(interpret(o2, a2, xs, ys), d(xs, ys), rtol=1e-12, atol=1e-12)


def test_rejects_unknown_symbols_and_constructs():
    with pytest.raises(ValueError):
        SymExpr("x + z")
    with pytest.raises(ValueError):
        SymExpr("Heaviside(x)").program()
