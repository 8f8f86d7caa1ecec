"""Self-consistency of the EDG oracle (oracle/edg_oracle.py; reference convection_diffusion.py:42-219, debug_edg.py:191-192):
these tests pin the oracle the CUDA path (tests/test_gpu_edg.py) is checked against."""
import json
import os

import numpy as np
import pytest

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from oracle.edg_oracle import EDGOracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _poly_problem(beta, eps, bonus=10):
    """u = x (1 - x) y (1 - y) (degree 4, zero on the boundary), coeff = -eps Lap u + b . grad u; callables take the
    distances (dx, dy) = (1 - x, 1 - y) like every oracle coefficient"""
    u = lambda dx, dy: (1 - dx) * dx * (1 - dy) * dy
    ux = lambda dx, dy: (2 * dx - 1) * (1 - dy) * dy                   # d/dx = -d/d(dx)
    uy = lambda dx, dy: (1 - dx) * dx * (2 * dy - 1)
    lap = lambda dx, dy: -2 * (1 - dy) * dy - 2 * (1 - dx) * dx
    coeff = lambda dx, dy: -eps * lap(dx, dy) + beta[0] * ux(dx, dy) + beta[1] * uy(dx, dy)
    return orc.Problem(beta, eps, coeff=coeff, exact=u, exact_grad=lambda dx, dy: (ux(dx, dy), uy(dx, dy)), bonus_int=bonus)


def _oracle(mesh, order, prob, theta=1e-5):
    return EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob, theta=theta)


def test_alpha_pencil_known_answers():
    """the EDG alpha is the same pencil (m, a + f f^T) as the EHDG one: known answers of tests/golden/alpha_kat.json"""
    kat = json.load(open(os.path.join(HERE, "golden", "alpha_kat.json")))
    mesh = unit_square_mesh(2, kind="structured")                      # right-isosceles triangles
    o = _oracle(mesh, 1, orc.run_py_problem(eps=1e-2, enriched=False))
    lam = o.alpha_all()
    assert np.allclose(lam[1:], 4.0, rtol=1e-12)                       # P1 right-isosceles (csv rows 0 / 11)
    assert lam[0] > 0.0                                                # element 0 loses its constant mode (`dof > 0`, sic): still a valid pencil
    for p in (2, 3):
        oe = _oracle(mesh, p, orc.run_py_problem(eps=1e-2, enriched=False))
        le = oe.alpha_all()
        lh = oe.h.alpha_all()
        assert np.allclose(le[1:], lh[1:], rtol=1e-10)                 # bonus order changes nothing for polynomials
    assert kat                                                          # the golden file is the shared reference of both paths


def test_diffusion_form_is_symmetric_and_convection_free_of_boundary_flux():
    mesh = unit_square_mesh(3, kind="unstructured")
    o = _oracle(mesh, 2, _poly_problem((0.0, 0.0), 1.0))
    A, _ = o.assemble(o.alpha_all())
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    # with convection the form loses its symmetry but every row of constants still tests to zero mass flux inside:
    ob = _oracle(mesh, 2, _poly_problem((2.0, 1.0), 1e-12))
    Ab, _ = ob.assemble(np.zeros(mesh.ne))
    const = np.zeros(ob.ndof)
    const[:mesh.ne] = 1.0                                              # u = 1 (phi_00 = 1 is the first dof of every element)
    r = Ab @ const
    # -int b u . grad v + sum_int b.n u_up [v] with u = 1: for v = 1 on one element this is the inflow/outflow through its
    # INTERIOR facets only (the reference has no boundary term); summed over all elements the interior fluxes cancel
    assert abs(r[:mesh.ne].sum()) < 1e-12


@pytest.mark.parametrize("beta,eps,kind", [((0.0, 0.0), 1.0, "structured"), ((2.0, 1.0), 0.5, "unstructured")])
def test_polynomial_solution_is_reproduced(beta, eps, kind):
    """SIP + upwind DG is consistent for u in H^2 with u = 0 on the boundary, and u = x(1-x)y(1-y) lies in P4"""
    mesh = unit_square_mesh(3, kind=kind)
    o = _oracle(mesh, 4, _poly_problem(beta, eps))
    x = o.solve()
    assert o.l2_error(x) < 1e-11


@pytest.mark.parametrize("beta,eps,kind,order", [((2.0, 1.0), 0.5, "structured", 2), ((-1.0, 0.7), 0.1, "unstructured", 3)])
def test_inhomogeneous_boundary_data_are_consistent_up_to_the_outflow_term(beta, eps, kind, order):
    """debug_edg.py:191-192 puts the Nitsche terms and the INFLOW data on the right-hand side; the reference's convection form
    (:181-182, skeleton integrals only) has no outflow term  int_{dOmega, b.n > 0} b.n u v  on the left, so the method is
    consistent for non-vanishing boundary values exactly up to that term.  Check: with it added, u = 1 + x - 2 y + x y (in P2)
    is reproduced to round-off; without the boundary data the error is O(1)."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    u = lambda dx, dy: 1.0 + (1 - dx) - 2.0 * (1 - dy) + (1 - dx) * (1 - dy)
    ux = lambda dx, dy: 1.0 + (1 - dy)
    uy = lambda dx, dy: -2.0 + (1 - dx)
    coeff = lambda dx, dy: beta[0] * ux(dx, dy) + beta[1] * uy(dx, dy)          # Lap u = 0
    prob = orc.Problem(beta, eps, coeff=coeff, exact=u, exact_grad=lambda dx, dy: (ux(dx, dy), uy(dx, dy)), bonus_int=10)
    mesh = unit_square_mesh(3, kind=kind)
    o = _oracle(mesh, order, prob)
    alpha = o.alpha_all()
    A, F = o.assemble(alpha, boundary_data=True)
    rows, cols, vals = [], [], []
    for f in np.nonzero(mesh.facet2el[:, 1] < 0)[0]:
        e1 = mesh.facet2el[f, 0]
        lam1, we, n, _ = o.facet_points(e1, f, 2 * order + 10)
        bn = float(np.asarray(beta) @ n)
        if bn > 0:
            u1 = o.shapes(e1, lam1)[0]
            d = o.el_dofs(e1)
            M = bn * np.einsum("in,jn,n->ij", u1, u1, we)
            rr, cc = np.meshgrid(d, d, indexing="ij")
            rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(M.ravel())
    Aout = sps.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=A.shape)
    x = spla.spsolve((A + Aout).tocsc(), F)
    assert o.l2_error(x) < 1e-11
    assert o.l2_error(o.solve(alpha, boundary_data=True)) < 5e-3               # the reference's own (slightly inconsistent) answer
    assert o.l2_error(o.solve(alpha, boundary_data=False)) > 1e-1              # without the terms the data are taken as zero


def test_dg_error_decays_with_the_order():
    mesh = unit_square_mesh(4, kind="unstructured")
    prob = orc.run_py_problem(beta=(2.0, 1.0), eps=0.5, enriched=False)
    errs = [(_o := _oracle(mesh, p, prob)).l2_error(_o.solve()) for p in (1, 2, 3)]
    assert errs[0] > 3 * errs[1] > 9 * errs[2]


def test_edg_equals_dg_when_everything_is_pruned_and_enrichment_helps_otherwise():
    mesh = unit_square_mesh(4, kind="structured")
    eps, beta = 2e-2, (2.0, 1.0)
    dg = _oracle(mesh, 2, orc.run_py_problem(beta=beta, eps=eps, enriched=False))
    e_dg = dg.l2_error(dg.solve())
    # theta = 2: every enrichment dof fails the factor test (factor <= 1 < theta) and is deactivated
    off = _oracle(mesh, 2, orc.run_py_problem(beta=beta, eps=eps, enriched=True), theta=2.0)
    off.prune()
    assert not off.active[off.ne * off.n_p:].any()
    x = off.solve()
    assert abs(off.l2_error(x) - e_dg) < 1e-10 * e_dg
    # the reference's theta keeps the layer functions on the strip: the boundary layer is resolved better
    on = _oracle(mesh, 2, orc.run_py_problem(beta=beta, eps=eps, enriched=True))
    on.prune()
    assert on.active[on.ne * on.n_p:].any()
    e_edg = on.l2_error(on.solve())
    assert e_edg < 0.7 * e_dg
