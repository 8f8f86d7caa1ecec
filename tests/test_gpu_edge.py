"""Edge cases of the CUDA path: degenerate meshes, argument errors, one enrichment, high orders, determinism,
tensor path vs quadrature path on an enriched mesh."""
import ctypes as C

import numpy as np
import pytest

from tests.common import make_pair, rel

pytestmark = pytest.mark.gpu


def _pipe(mesh, spec, **kw):
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    return EHDGPipeline(mesh, spec, **kw)


def test_single_triangle_has_no_free_facet_dofs():
    from convectiondiffusionequation_b200.mesh import single_triangle
    mesh = single_triangle([[0, 0], [1, 0], [0.2, 0.9]])
    o, spec = make_pair(mesh, 3, eps=1e-2, enriched=False)
    P = _pipe(mesh, spec).setup().assemble().solve(restart=10, max_iters=10, rtol=1e-12)
    assert P.csr_sizes.n_rows == 0 and P.csr_sizes.nnz == 0 and P.gmres_info.converged == 1
    # all facets Dirichlet: u_h = A_VV^-1 f_V
    z = P.z.cpu().numpy()
    assert np.array_equal(P.u.cpu().numpy(), z)
    x = o.solve_condensed(o.alpha_all())
    assert rel(z, x[o.vol_dofs(0)]) < 1e-11
    l2, _ = P.errors()
    assert abs(l2 - o.l2_error(x)) < 1e-10 * l2


def test_two_element_mesh_and_one_enrichment():
    from convectiondiffusionequation_b200 import boundary_layer, x, y
    from convectiondiffusionequation_b200.hot_path import ProblemSpec
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    from oracle import ehdg_oracle as orc
    mesh = unit_square_mesh(1)
    beta, eps = (2.0, 1.0), 0.05
    p_, q_ = boundary_layer(x, beta[0], eps), boundary_layer(y, beta[1], eps)
    spec = ProblemSpec(order=2, epsilon=eps, beta=beta, coeff=beta[1] * p_ + beta[0] * q_, exact=p_ * q_,
                       enrich_functions=(p_,), enrich_domain_ind=(lambda X, Y, h: X > 1 - h,))
    P = _pipe(mesh, spec).setup()
    assert P.sizes.nb == 4 and P.sizes.nvmax == 7                      # one enrichment slot
    assert int(P.elem_mask.sum()) == 2 and P.plan_sizes.n_gen == 2    # every element touches x = 1
    P.assemble().solve(restart=20, max_iters=200, rtol=1e-12)
    assert P.csr_sizes.n_rows >= 3
    # oracle with the same single enrichment
    prob = orc.run_py_problem(beta=beta, eps=eps)
    prob.enrich, prob.indicators, prob.enrich_desc = prob.enrich[:1], prob.indicators[:1], prob.enrich_desc[:1]
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, 2, prob)
    o.prune()
    lam = o.alpha_all()
    assert np.max(np.abs(P.lam.cpu().numpy() - lam) / lam) < 1e-8
    l2, _ = P.errors()
    assert abs(l2 - o.l2_error(o.solve_condensed(lam))) < 1e-6 * l2


def test_argument_errors_are_reported_not_crashed():
    from convectiondiffusionequation_b200 import _lib
    from convectiondiffusionequation_b200.hot_path import ProblemSpec
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    lib = _lib.load()
    for bad in (dict(order=7), dict(order=0), dict(order=2, epsilon=0.0)):
        p = _lib.Problem()
        p.order, p.epsilon = bad.get("order", 2), bad.get("epsilon", 0.1)
        ctx = C.c_void_p()
        assert lib.ehdg_create(C.byref(p), C.byref(ctx)) == -1 and b"invalid problem" in lib.ehdg_last_error()
    assert lib.ehdg_get_sizes(None, None) == -1
    P = _pipe(unit_square_mesh(2), ProblemSpec(order=1, epsilon=0.1, beta=(1.0, 0.0))).setup()
    with pytest.raises(_lib.EhdgError):
        _lib.check(lib.ehdg_alpha(P.ctx, P.plan, C.byref(P.cmesh), None, None, None, None), "ehdg_alpha")
    with pytest.raises(ValueError):
        from convectiondiffusionequation_b200 import boundary_layer, x
        f = boundary_layer(x, 1.0, 0.1)
        ProblemSpec(order=1, epsilon=0.1, beta=(1, 0), enrich_functions=(f, f, f)).c_struct()


@pytest.mark.parametrize("order", [5, 6])
def test_high_order_full_solve(order):
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    mesh = unit_square_mesh(2, kind="unstructured")
    o, spec = make_pair(mesh, order, beta=(1.0, 0.5), eps=0.05, enriched=False)
    P = _pipe(mesh, spec).setup().assemble().solve(restart=80, max_iters=4000, rtol=1e-11)      # line blocks without row exchanges: ~1e-11 attainable at p = 6
    assert P.gmres_info.converged == 1, (P.gmres_info.iters, P.gmres_info.rel_residual, P.gmres_info.rel_residual_unscaled)
    lam = o.alpha_all()
    x = o.solve_condensed(lam)
    l2, h1 = P.errors()
    l2o, h1o = o.l2_error(x, h1=True)
    assert abs(l2 - l2o) < 1e-7 * l2o and abs(h1 - h1o) < 1e-6 * h1o


def test_assembly_and_solve_are_bit_reproducible():
    import torch
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    mesh = unit_square_mesh(12, kind="unstructured")
    _, spec = make_pair(mesh, 3, beta=(2.0, 1.0), eps=2e-2, enriched=True)
    outs = []
    for _ in range(2):
        P = _pipe(mesh, spec).setup().assemble().csr_build()
        vals = P.vals.clone()
        P.gmres(restart=40, max_iters=300, rtol=1e-10)
        outs.append((P.S.clone(), vals, P.rhs.clone(), P.xhat.clone(), P.lam.clone()))
        P.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)                       # no atomics anywhere on the value path


def test_tensor_path_equals_quadrature_path_on_an_enriched_mesh():
    """same enriched problem through the default plan (tensor path for the bulk) and with every element forced
    through the quadrature kernels: identical condensed systems"""
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    mesh = unit_square_mesh(10)
    _, spec = make_pair(mesh, 4, beta=(2.0, 1.0), eps=1e-2, enriched=True)
    Pd = _pipe(mesh, spec).setup().assemble().csr_build()
    Pq = _pipe(mesh, spec, force_general=True).setup()
    assert Pq.plan_sizes.n_poly == 0 and Pd.plan_sizes.n_poly > 0
    Pq.alpha()
    st = (Pd.status == 0) & (Pq.status == 0)
    assert float(((Pd.lam - Pq.lam).abs() / Pq.lam)[st].max()) < 1e-9
    Pq.lam.copy_(Pd.lam)
    Pq.assemble_condense().csr_build()
    assert Pd.csr_sizes.n_rows == Pq.csr_sizes.n_rows and Pd.csr_sizes.nnz == Pq.csr_sizes.nnz
    scale = float(Pq.vals.abs().max())
    assert float((Pd.vals - Pq.vals).abs().max()) < 1e-9 * scale
    assert float((Pd.rhs - Pq.rhs).norm() / Pq.rhs.norm()) < 1e-8


@pytest.mark.parametrize("order,enriched", [(1, False), (2, False), (3, False), (4, False), (5, False), (6, False),
                                            (1, True), (4, True), (6, True)])
def test_facet_block_spmv_every_variant(order, enriched):
    """the facet-block SpMV has one instantiation per block height (2, 3, 4, 6, 9 slots; 16 or 32 lanes per facet; one or two
    column chunks): each against scipy on the CSR arrays the library exposes, on a mesh with boundary facets and ragged rows"""
    import numpy as np
    import scipy.sparse as sps
    import torch
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    from tests.common import make_pair
    mesh = unit_square_mesh(7, kind="unstructured")
    _, spec = make_pair(mesh, order, beta=(2.0, 1.0), eps=1e-2, enriched=enriched)
    P = EHDGPipeline(mesh, spec).setup().assemble().csr_build()
    rp, ci, rs, dm = [t.cpu().numpy() for t in P.csr_arrays()]
    n, nnz = P.csr_sizes.n_rows, P.csr_sizes.nnz
    A = sps.csr_matrix((P.vals.cpu().numpy()[:nnz], ci, rp), shape=(n, n))
    rng = np.random.default_rng(order)
    x = rng.standard_normal(n)
    y = P.spmv(torch.from_numpy(x).to(P.dev)).cpu().numpy()
    ref = A @ x
    assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
    P.close()
