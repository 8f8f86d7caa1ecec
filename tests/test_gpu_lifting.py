"""Inhomogeneous Dirichlet lifting on the device (reference Python/debug_ehdg.py:210-214) against the oracle."""
import numpy as np
import pytest
import torch

from convectiondiffusionequation_b200 import boundary_layer, x, y
from convectiondiffusionequation_b200.coefficients import LayerExpr
from convectiondiffusionequation_b200.hot_path import EHDGPipeline, ProblemSpec
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order,kind,shift", [(1, "structured", 1.5), (3, "unstructured", -0.75), (4, "structured", 2.0)])
def test_lifted_solution_matches_the_oracle(order, kind, shift):
    mesh = unit_square_mesh(5, kind=kind)
    eps, beta = 0.05, (1.0, 0.5)
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    base = prob.exact
    prob.exact = lambda dx, dy: shift + base(dx, dy)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    lams = o.alpha_all()
    xo = o.solve_condensed_lifted(lams)
    l2_o, h1_o = o.l2_error(xo), None
    p, q = boundary_layer(x, beta[0], eps), boundary_layer(y, beta[1], eps)
    exact = p * q + LayerExpr((shift, 0.0, 0.0, 0.0))
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=10, coeff=beta[1] * p + beta[0] * q, exact=exact)
    P = EHDGPipeline(mesh, spec, dirichlet_lifting=True).setup().assemble().solve(restart=60, max_iters=8000, rtol=1e-13)
    assert P.gmres_info.converged == 1
    # projected boundary data: the oracle's coefficients facet by facet
    ub = P.ubnd.cpu().numpy().reshape(mesh.nf, order + 1)
    for f, c in o.dirichlet_values().items():
        assert np.abs(ub[f] - c).max() < 1e-12 * max(1.0, np.abs(c).max())
    assert np.all(ub[mesh.facet2el[:, 1] >= 0] == 0.0)
    l2, _ = P.errors()
    assert abs(l2 - l2_o) < 1e-8 * l2_o, (l2, l2_o)
    # and the same pipeline without the lifting solves a different (wrong-data) problem
    P0 = EHDGPipeline(mesh, spec).setup().assemble().solve(restart=60, max_iters=8000, rtol=1e-13)
    assert P0.errors()[0] > 10 * l2
    P.close()
    P0.close()
