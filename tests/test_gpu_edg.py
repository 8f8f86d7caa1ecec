"""EDG path on the device (reference Python/convection_diffusion.py:42-219) against oracle/edg_oracle.py through the C ABI:
pruning decisions, alpha, sparsity pattern (bit-exact in NGSolve's numbering), matrix values and right-hand side (1e-11),
discrete solution through the L2 error (1e-8)."""
import numpy as np
import pytest
import scipy.sparse as sps
import torch

from convectiondiffusionequation_b200.edg_path import EDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from oracle.edg_oracle import EDGOracle
from tests.common import make_pair

pytestmark = pytest.mark.gpu


def _pair(n, order, eps, beta, enriched, kind="structured", theta=1e-5):
    mesh = unit_square_mesh(n, kind=kind)
    _, spec = make_pair(mesh, order, beta=beta, eps=eps, enriched=enriched)
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=enriched)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob, theta=theta)
    o.prune()
    return mesh, spec, o


def _oracle_active_bits(o):
    act = np.zeros(o.ne, dtype=np.uint8)
    for i in range(o.nenr):
        for el in np.nonzero(o.h.el_mark[i])[0]:
            if o.active[o.offQ[i] + o.q_num[i][el]]:
                act[el] |= 1 << i
    return act


def _row_to_oracle_dof(o, act):
    """oracle (NGSolve-numbered) dof of every row of the element-major system"""
    out = []
    for el in range(o.ne):
        d = o.el_dofs(el)
        out += list(d[:o.n_p])
        for i in range(o.nenr):
            if (act[el] >> i) & 1:
                out.append(d[o.n_p + i])
    return np.asarray(out, dtype=np.int64)


CASES = [(4, 1, 1e-2, (2.0, 1.0), False, "structured"), (4, 3, 1e-3, (2.0, 0.001), False, "unstructured"),
         (4, 2, 2e-2, (2.0, 1.0), True, "structured"), (5, 2, 1e-2, (-1.0, 2.0), True, "unstructured"),
         (4, 4, 1e-2, (2.0, 0.001), True, "structured")]


@pytest.mark.parametrize("n,order,eps,beta,enriched,kind", CASES)
def test_edg_matrix_rhs_and_pattern(n, order, eps, beta, enriched, kind):
    mesh, spec, o = _pair(n, order, eps, beta, enriched, kind)
    P = EDGPipeline(mesh, spec).setup(theta=1e-5)
    act = _oracle_active_bits(o)
    assert np.array_equal(P.elem_active.cpu().numpy(), act)                      # pruning decisions (:83-123)
    rng = np.random.default_rng(7)
    alpha_stab = 5.0 + 20.0 * rng.random(mesh.ne)                                # the forms are linear in alpha_stab
    P.lam = torch.from_numpy(alpha_stab).to(P.dev)
    P.status = torch.zeros(mesh.ne, dtype=torch.int32, device=P.dev)
    P.csr_symbolic()
    n_rows, nnz = P.csr_sizes.n_rows, P.csr_sizes.nnz
    P.vals = torch.empty(nnz, dtype=torch.float64, device=P.dev)
    P.rhs = torch.empty(n_rows, dtype=torch.float64, device=P.dev)
    P.csr_numeric()
    rp, ci, rs = (t.cpu().numpy() for t in P.csr_arrays())
    A_dev = sps.csr_matrix((P.vals.cpu().numpy(), ci, rp), shape=(n_rows, n_rows))
    A_or, F_or = o.assemble(alpha_stab)
    perm = _row_to_oracle_dof(o, act)
    assert len(perm) == n_rows == int(o.active.sum())
    A_ref = A_or.tocsr()[perm][:, perm].tocsr()
    A_ref.sort_indices()
    # pattern: in NGSolve's numbering, bit-exact
    inv = np.empty(o.ndof, dtype=np.int64)
    inv[:] = -1
    inv[perm] = np.arange(n_rows)
    act_idx = np.nonzero(o.active)[0]
    A_ng = A_or.tocsr()[act_idx][:, act_idx].tocsr()
    A_ng.sort_indices()
    D_ng = A_dev.tocoo()
    remap = np.searchsorted(act_idx, perm)                                       # row of the device system -> index among the active oracle dofs
    D_ng = sps.csr_matrix((D_ng.data, (remap[D_ng.row], remap[D_ng.col])), shape=A_ng.shape)
    D_ng.sort_indices()
    assert np.array_equal(D_ng.indptr, A_ng.indptr) and np.array_equal(D_ng.indices, A_ng.indices)
    scale = abs(A_ref).max()
    assert abs(A_dev - A_ref).max() < 1e-11 * scale
    assert np.abs(P.rhs.cpu().numpy() - F_or[perm]).max() < 1e-12 * max(1.0, np.abs(F_or).max())
    assert int((P.status != 0).sum()) == 0
    P.close()


@pytest.mark.parametrize("n,order,eps,beta,enriched,kind", CASES)
def test_edg_alpha(n, order, eps, beta, enriched, kind):
    mesh, spec, o = _pair(n, order, eps, beta, enriched, kind)
    P = EDGPipeline(mesh, spec).setup(theta=1e-5).alpha()
    lam = P.lam.cpu().numpy()
    ref = o.alpha_all()
    marked = o.h.el_marked_any if enriched else np.zeros(mesh.ne, bool)
    assert np.max(np.abs(lam[~marked] - ref[~marked]) / ref[~marked]) < 1e-10
    if marked.any():
        st = P.status.cpu().numpy()
        ok = marked & (st == 0)
        assert ok.sum() >= 0.5 * marked.sum()
        assert np.max(np.abs(lam[ok] - ref[ok]) / ref[ok]) < 1e-6           # a + f f^T of an enriched element is ill-conditioned
    P.close()


@pytest.mark.parametrize("n,order,eps,beta,enriched,kind", [(6, 2, 1e-2, (2.0, 1.0), False, "structured"),
                                                            (6, 3, 2e-2, (2.0, 0.5), False, "unstructured"),
                                                            (6, 2, 2e-2, (2.0, 1.0), True, "structured")])
def test_edg_solution(n, order, eps, beta, enriched, kind):
    mesh, spec, o = _pair(n, order, eps, beta, enriched, kind)
    P = EDGPipeline(mesh, spec).setup(theta=1e-5).alpha()
    alpha_stab = P.lam.cpu().numpy()
    x = o.solve(alpha_stab)                                                      # same stabilisation on both sides
    l2_ref = o.l2_error(x)
    P.assemble().gmres(restart=60, max_iters=20000, rtol=1e-10, method="gmres").backsolve()      # floor of the line blocks: a few 1e-10 at p = 3
    assert P.gmres_info.converged == 1, (P.gmres_info.iters, P.gmres_info.rel_residual)
    l2, _ = P.errors()
    assert abs(l2 - l2_ref) < 1e-8 * l2_ref, (l2, l2_ref)
    P.close()


@pytest.mark.parametrize("order,kind,shift,beta", [(2, "structured", 1.5, (2.0, 1.0)), (3, "unstructured", -0.75, (-1.0, 0.5))])
def test_edg_inhomogeneous_boundary_data(order, kind, shift, beta):
    """debug_edg.py:191-192: Nitsche + inflow terms of the right-hand side for boundary values that do not vanish; the
    right-hand side (1e-11) and the solution (L2 error, 1e-8) against the oracle"""
    from convectiondiffusionequation_b200 import boundary_layer, x, y
    from convectiondiffusionequation_b200.coefficients import LayerExpr
    from convectiondiffusionequation_b200.hot_path import ProblemSpec
    mesh = unit_square_mesh(5, kind=kind)
    eps = 0.05
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    base = prob.exact
    prob.exact = lambda dx, dy: shift + base(dx, dy)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    p, q = boundary_layer(x, beta[0], eps), boundary_layer(y, beta[1], eps)
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=10, coeff=beta[1] * p + beta[0] * q,
                       exact=p * q + LayerExpr((shift, 0.0, 0.0, 0.0)))
    P = EDGPipeline(mesh, spec, boundary_data=True).setup(theta=1e-5).alpha()
    alpha_stab = P.lam.cpu().numpy()
    _, F = o.assemble(alpha_stab, boundary_data=True)
    _, F0 = o.assemble(alpha_stab, boundary_data=False)
    P.assemble()
    rhs = P.rhs.cpu().numpy()[:o.ndof]                       # no enrichment: rows are the oracle's V dofs in element order
    perm = _row_to_oracle_dof(o, np.zeros(o.ne, dtype=np.uint8))
    assert np.abs(rhs - F[perm]).max() < 1e-11 * np.abs(F).max()
    assert np.abs(F - F0).max() > 1e-3 * np.abs(F).max()      # the boundary terms are not negligible here
    x_o = o.solve(alpha_stab, boundary_data=True)
    l2_ref = o.l2_error(x_o)
    P.gmres(restart=60, max_iters=20000, rtol=1e-10, method="gmres").backsolve()
    assert P.gmres_info.converged == 1
    l2, _ = P.errors()
    # the discrete solutions agree to 1e-8 of their size, 1 + |shift| (the solver tolerance: floor of the line blocks a few
    # 1e-10); the L2 ERROR is 3e-3 .. 4e-2 of that, so the bound is stated on the scale of the solution
    assert abs(l2 - l2_ref) < 1e-8 * (1.0 + abs(shift)), (l2, l2_ref)
    # the driver key: config['dirichlet_lifting'] switches the terms on in _solveEDG
    Q = EDGPipeline(mesh, spec).setup(theta=1e-5).alpha().assemble()
    assert np.abs(Q.rhs.cpu().numpy()[:o.ndof] - F0[perm]).max() < 1e-11 * np.abs(F0).max()
    P.close()
    Q.close()


def test_solveEDG_driver_tables():
    from convectiondiffusionequation_b200 import run
    from convectiondiffusionequation_b200.convection_diffusion import Convection_Diffusion
    cfg = dict(run.config)
    cfg.update({'order': [1, 2], 'mesh_size': [0.25], 'quiet': True})
    res, alphas = Convection_Diffusion(cfg)._solveEDG()
    assert list(res.columns) == ['Order', 'DOFs', 'Mesh Size', 'Error', 'Bonus Int', 'Type']
    assert list(res['Type']) == ['edg', 'edg'] and np.all(np.isfinite(res['Error'].to_numpy(dtype=float)))
    assert list(alphas.columns) == ['alpha', 'type', 'mesh_size', 'order'] and len(alphas) == 2 * 32
    plain = dict(cfg)
    plain.update({'enrich_functions': [], 'enrich_domain_ind': []})
    res2, _ = Convection_Diffusion(plain)._solveEDG()
    assert list(res2['Type']) == ['dg', 'dg']
