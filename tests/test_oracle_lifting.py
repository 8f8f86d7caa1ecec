"""Inhomogeneous Dirichlet lifting (reference Python/debug_ehdg.py:210-214) in the oracle: the reference's literal sequence on the
uncondensed system and the condensed variant agree, and shifting the exact solution by a constant (which every order reproduces,
and which is non-zero on the whole boundary) leaves the error where it was."""
import numpy as np
import pytest

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc


def _problem(shift, eps, beta):
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    base = prob.exact
    prob.exact = lambda dx, dy: shift + base(dx, dy)
    prob.exact_desc = ((shift, 0.0, 0.0, 1.0), prob.exact_desc[1])
    return prob


@pytest.mark.parametrize("order,kind", [(1, "structured"), (2, "unstructured"), (3, "structured")])
def test_lifting_condensed_equals_full_and_constant_shift_is_reproduced(order, kind):
    mesh = unit_square_mesh(4, kind=kind)
    eps, beta = 0.05, (1.0, 0.5)
    o0 = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, _problem(0.0, eps, beta))
    o1 = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, _problem(1.5, eps, beta))
    lams = o0.alpha_all()
    e0 = o0.l2_error(o0.solve_condensed(lams))
    xf, xc = o1.solve_full_lifted(lams), o1.solve_condensed_lifted(lams)
    assert np.abs(xf - xc).max() < 1e-10 * np.abs(xf).max()
    e1 = o1.l2_error(xc)
    assert abs(e1 - e0) < 1e-9 * e0
    # without the lifting the shifted problem is solved with the wrong boundary data
    assert o1.l2_error(o1.solve_condensed(lams)) > 10 * e0
    # homogeneous data: the lifted solve is the plain one
    assert np.abs(o0.solve_condensed_lifted(lams) - o0.solve_condensed(lams)).max() < 1e-12
