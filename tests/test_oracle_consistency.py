"""Self-consistency of the CPU oracle: these replace the golden matrices
the reference does not hold."""
import numpy as np
import pytest

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc


def _oracle(n, order, kind="structured", **kw):
    mesh = unit_square_mesh(n, kind=kind)
    return orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, orc.run_py_problem(**kw))


@pytest.mark.parametrize("order,kind", [(1, "structured"), (2, "unstructured"), (3, "structured")])
def test_condensed_equals_full_direct_solve(order, kind):
    """static condensation + back-solve == the reference's uncondensed PARDISO-style solve (:385-387)"""
    o = _oracle(3, order, kind, enriched=False)
    lams = o.alpha_all()
    xf, xc = o.solve_full(lams), o.solve_condensed(lams)
    assert np.abs(xf - xc).max() < 1e-10 * np.abs(xf).max()


def test_schur_of_global_equals_scatter_of_local():
    o = _oracle(3, 2, enriched=False)
    lams = o.alpha_all()
    M, rhs = o.assemble_full(lams)
    S, g = o.assemble_condensed(lams)
    act = o.active
    nV = o.offF
    v = np.arange(nV)
    f = np.nonzero(act & (np.arange(o.ndof) >= nV))[0]
    Md = M.toarray()
    Sg = Md[np.ix_(f, f)] - Md[np.ix_(f, v)] @ np.linalg.solve(Md[np.ix_(v, v)], Md[np.ix_(v, f)])
    assert np.abs(Sg - S.toarray()[np.ix_(f, f)]).max() < 1e-11 * np.abs(Sg).max()


def test_enriched_solution_has_smaller_error_and_function_level_agreement():
    """with enrichment the discrete system is nearly singular (near-polynomial q(y)); the solution as a
    FUNCTION is still well defined: both solves give the same L2 error."""
    o = _oracle(4, 2, eps=0.01)
    o.prune()
    lams = o.alpha_all()
    ef, ec = o.l2_error(o.solve_full(lams)), o.l2_error(o.solve_condensed(lams))
    assert abs(ef - ec) < 1e-5 * ef


def test_l2_error_decreases_with_refinement_and_order():
    errs = {}
    for n in (2, 4):
        for order in (1, 2):
            o = _oracle(n, order, enriched=False, eps=0.1, beta=(1.0, 0.5))
            errs[(n, order)] = o.l2_error(o.solve_condensed(o.alpha_all()), order=12)
    assert errs[(4, 1)] < 0.45 * errs[(2, 1)]
    assert errs[(4, 2)] < 0.25 * errs[(2, 2)]
    assert errs[(4, 2)] < errs[(4, 1)]


def test_dof_numbering_and_free_dofs():
    o = _oracle(3, 2)
    assert o.ndof == o.ne * o.n_p + o.nf * 3 + sum(int(m.sum()) for m in o.el_mark) + sum(int(m.sum()) for m in o.fac_mark)
    bnd = np.nonzero(o.facet2el[:, 1] < 0)[0]
    for f in bnd:
        assert not o.free[o.facet_dofs(f)].any()
    d = o.el_dofs(0)
    assert len(d) == o.nloc
