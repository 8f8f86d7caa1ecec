"""Generic coefficient functions on the device (reference run.py:26-30, :44): `coeff` and `exact` as arbitrary
expressions of x and y, compiled to the postfix program the kernels interpret, against the oracle evaluating the same sympy
expressions with numpy."""
import numpy as np
import pytest
import sympy

from convectiondiffusionequation_b200 import SymExpr
from convectiondiffusionequation_b200.edg_path import EDGPipeline
from convectiondiffusionequation_b200.hot_path import EHDGPipeline, ProblemSpec
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from oracle.edg_oracle import EDGOracle

pytestmark = pytest.mark.gpu
X, Y = sympy.Symbol("x", real=True), sympy.Symbol("y", real=True)


def _problem(u, eps, beta):
    """manufactured problem: coeff = -eps Lap u + b . grad u; (oracle Problem, product coeff / exact)"""
    f = -eps * (sympy.diff(u, X, 2) + sympy.diff(u, Y, 2)) + beta[0] * sympy.diff(u, X) + beta[1] * sympy.diff(u, Y)
    cf, ex = SymExpr(f), SymExpr(u)
    gx, gy = ex.diff(0), ex.diff(1)
    prob = orc.Problem(beta, eps, coeff=lambda dx, dy: cf(1.0 - dx, 1.0 - dy), exact=lambda dx, dy: ex(1.0 - dx, 1.0 - dy),
                       exact_grad=lambda dx, dy: (gx(1.0 - dx, 1.0 - dy), gy(1.0 - dx, 1.0 - dy)))
    return prob, cf, ex


@pytest.mark.parametrize("order,kind", [(2, "structured"), (3, "unstructured"), (4, "unstructured")])
def test_hdg_with_symbolic_coefficients(order, kind):
    eps, beta = 0.05, (1.0, 0.5)
    prob, cf, ex = _problem(sympy.sin(sympy.pi * X) * sympy.sin(sympy.pi * Y), eps, beta)      # vanishes on the boundary
    mesh = unit_square_mesh(6, kind=kind)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    lams = o.alpha_all()
    l2_o, h1_o = o.l2_error(o.solve_condensed(lams), h1=True)
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=10, coeff=cf, exact=ex)
    P = EHDGPipeline(mesh, spec).setup().assemble().solve(restart=60, max_iters=8000, rtol=1e-11)
    assert P.gmres_info.converged == 1, (P.gmres_info.iters, P.gmres_info.rel_residual)
    l2, h1 = P.errors()
    assert abs(l2 - l2_o) < 1e-8 * l2_o and abs(h1 - h1_o) < 1e-7 * h1_o, (l2, l2_o, h1, h1_o)
    assert l2 < 5e-2 / order ** 2                                # and it is the right answer, not just the same one
    P.close()


def test_lifting_with_symbolic_boundary_data():
    eps, beta, order = 0.05, (1.0, 0.5), 3
    prob, cf, ex = _problem(sympy.exp(X) * sympy.cos(Y) + X * Y, eps, beta)                      # non-zero on the whole boundary
    mesh = unit_square_mesh(5, kind="unstructured")
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    lams = o.alpha_all()
    l2_o = o.l2_error(o.solve_condensed_lifted(lams))
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=10, coeff=cf, exact=ex)
    P = EHDGPipeline(mesh, spec, dirichlet_lifting=True).setup().assemble().solve(restart=60, max_iters=8000, rtol=1e-11)
    l2, _ = P.errors()
    # the error itself is 3e-6 of a solution of size 3: 1e-8 of it is below what a 1e-11 residual resolves
    assert abs(l2 - l2_o) < max(1e-8 * l2_o, 1e-10), (l2, l2_o)
    assert l2 < 1e-4
    P.close()


def test_edg_with_symbolic_coefficients():
    eps, beta, order = 0.05, (1.0, 0.5), 2
    prob, cf, ex = _problem(sympy.sin(sympy.pi * X) * sympy.sin(sympy.pi * Y), eps, beta)
    mesh = unit_square_mesh(6, kind="structured")
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=10, coeff=cf, exact=ex)
    P = EDGPipeline(mesh, spec).setup().alpha()
    l2_o = o.l2_error(o.solve(P.lam.cpu().numpy()))
    P.assemble().gmres(restart=60, max_iters=20000, rtol=1e-10, method="gmres").backsolve()
    l2, _ = P.errors()
    assert abs(l2 - l2_o) < 1e-8 * l2_o, (l2, l2_o)
    P.close()


def test_malformed_program_is_rejected():
    import ctypes as C
    from convectiondiffusionequation_b200 import _lib
    mesh = unit_square_mesh(2)
    P = EHDGPipeline(mesh, ProblemSpec(order=1, epsilon=0.1, beta=(1.0, 0.0)))
    ops = np.array([4], dtype=np.int32)              # ADD on an empty stack
    args = np.zeros(1)
    rc = P.lib.ehdg_set_expression(P.ctx, 0, ops.ctypes.data_as(C.c_void_p), args.ctypes.data_as(C.c_void_p), 1)
    assert rc == -1 and b"malformed" in P.lib.ehdg_last_error()
    P.close()
