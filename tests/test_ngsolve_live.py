"""Live comparison with NGSolve (SURVEY §4(v), §8c check 6) -- runs only where `import ngsolve` works, skips elsewhere.

This is the one route from "parity unpinned" to pinned: the same arrays go into an NGSolve mesh, the reference's own HDG forms
(Python/convection_diffusion.py:357-371 without enrichment) are assembled with `condense=True`, and alpha_K, the sparsity pattern
of the condensed matrix (bit-exact) and its entries (1e-11) are compared with the device's.  It has NEVER been executed in the
build container (no ngsolve, no network); it restates the NGSolve calls of the reference (`debug_ehdg.py:206-207` is the COO hook)
and is expected to need small fixes the first time a box with NGSolve runs it.  Result is logged either way."""
import numpy as np
import pytest

ngsolve = pytest.importorskip("ngsolve", reason="NGSolve is not installed: the live parity suite is skipped (oracle parity stays unpinned)")
pytestmark = pytest.mark.gpu


def _ngmesh(mesh):
    from netgen.meshing import Element1D, Element2D, FaceDescriptor, Mesh as NGMesh, MeshPoint, Pnt
    m = NGMesh(dim=2)
    pids = [m.Add(MeshPoint(Pnt(float(x), float(y), 0.0))) for x, y in mesh.verts]
    m.Add(FaceDescriptor(surfnr=1, domin=1, bc=1))
    for t in mesh.tris:
        m.Add(Element2D(1, [pids[int(v)] for v in t]))
    for f in np.nonzero(mesh.facet2el[:, 1] < 0)[0]:
        a, b = mesh.facet_verts[f]
        m.Add(Element1D([pids[int(a)], pids[int(b)]], index=1))
    return ngsolve.Mesh(m)


@pytest.mark.parametrize("n,order", [(2, 1), (2, 3), (16, 2)])
def test_alpha_and_condensed_matrix_against_live_ngsolve(n, order):
    from ngsolve import (L2, BilinearForm, CoefficientFunction, FacetFESpace, FESpace, IfPos, SymbolicBFI, TaskManager, dx, grad,
                         specialcf)
    import scipy.sparse as sps
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    from tests.common import make_pair
    from tests.test_gpu_parity import _rows_to_ngsolve
    mesh = unit_square_mesh(n)
    eps, beta = 1e-2, (2.0, 0.001)
    o, spec = make_pair(mesh, order, beta=beta, eps=eps, enriched=False)
    P = EHDGPipeline(mesh, spec).setup().assemble().csr_build()
    ngm = _ngmesh(mesh)
    V = L2(ngm, order=order)
    F = FacetFESpace(ngm, order=order, dirichlet=".*")
    fes = FESpace([V, F])
    (u, uhat), (v, vhat) = fes.TnT()
    h = specialcf.mesh_size
    nrm = specialcf.normal(2)
    # alpha_K as in convection_diffusion.py:314-351
    a_diff = SymbolicBFI(grad(u) * grad(v))
    f_int = ngsolve.SymbolicLFI(h ** ((-2 - ngm.dim) / 2) * v)
    m_int = SymbolicBFI(h * (grad(u) * nrm) * (grad(v) * nrm), element_boundary=True)
    lam = np.zeros(mesh.ne)
    for el in fes.Elements():
        a = a_diff.CalcElementMatrix(el.GetFE(), el.GetTrafo()).NumPy().copy()
        fv = f_int.CalcElementVector(el.GetFE(), el.GetTrafo()).NumPy()
        a += np.outer(fv, fv)
        mm = m_int.CalcElementMatrix(el.GetFE(), el.GetTrafo()).NumPy()
        imp = [i for i, d in enumerate(el.dofs) if d > 0]
        lam[el.nr] = np.max(np.linalg.eig(np.linalg.pinv(a[np.ix_(imp, imp)]) @ mm[np.ix_(imp, imp)])[0]).real
    lam_dev = P.lam.cpu().numpy()
    assert np.max(np.abs(lam_dev - lam) / lam) < 1e-10
    # HDG forms :357-371 with condense=True
    alpha = CoefficientFunction(list(20.0 * lam))           # P0 field, element-wise
    b = CoefficientFunction(beta)
    dS = dx(element_boundary=True, bonus_intorder=10)
    dy = dx(bonus_intorder=10)
    ju, jv = u - uhat, v - vhat
    diffusion = grad(u) * grad(v) * dy + alpha * order ** 2 / h * ju * jv * dS + (-grad(u) * nrm * jv - grad(v) * nrm * ju) * dS
    uup = IfPos(b * nrm, u, uhat)
    convection = -b * u * grad(v) * dy + b * nrm * uup * jv * dS + IfPos(b * nrm, (uhat - u) * vhat, 0) * dS
    acd = BilinearForm(fes, condense=True)
    acd += eps * diffusion + convection
    with TaskManager():
        acd.Assemble()
    rows, cols, vals = acd.mat.COO()
    A_ng = sps.csr_matrix((np.asarray(vals), (np.asarray(rows), np.asarray(cols))), shape=(fes.ndof, fes.ndof))
    rp, ci, rs, dm = [t.cpu().numpy() for t in P.csr_arrays()]
    nrow = P.csr_sizes.n_rows
    A_dev = sps.csr_matrix((P.vals.cpu().numpy()[:P.csr_sizes.nnz], ci, rp), shape=(nrow, nrow))
    ng_rows = _rows_to_ngsolve(o, rs, dm, order)             # device row -> NGSolve dof number (oracle numbering = A.3 / A.4)
    perm = np.argsort(ng_rows)
    Ap = A_dev[perm][:, perm].tocsr()
    Ap.sort_indices()
    free = np.sort(ng_rows)
    Sub = A_ng[free][:, free].tocsr()
    Sub.sort_indices()
    assert np.array_equal(Ap.indptr, Sub.indptr) and np.array_equal(Ap.indices, Sub.indices)     # pattern: bit-exact
    assert np.abs(Ap.data - Sub.data).max() < 1e-11 * np.abs(Sub.data).max()
    print(f"live NGSolve parity ran: n={n} order={order} alpha and condensed matrix agree")
