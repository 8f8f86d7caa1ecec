"""Host build of the EDG element routine (csrc/edg_element.h, groundwork for SURVEY.md 8(f) row 1) against the EDG oracle:
every block row of the global matrix and the right-hand side, plain DG and enriched, all vertex orderings."""
import ctypes as C

import numpy as np
import pytest

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from oracle.edg_oracle import EDGOracle
from tests.hostcheck.binding import c_problem, dptr, iptr, lib

L = lib()


@pytest.mark.parametrize("order,eps,beta,enriched,kind", [(1, 1e-2, (2.0, 1.0), False, "structured"),
                                                          (3, 1e-3, (2.0, 0.001), False, "unstructured"),
                                                          (2, 2e-2, (2.0, 1.0), True, "structured"),
                                                          (2, 1e-2, (-1.0, 2.0), True, "unstructured")])
def test_edg_block_rows_match_the_oracle(order, eps, beta, enriched, kind):
    mesh = unit_square_mesh(4, kind=kind)
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=enriched)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    o.prune()
    rng = np.random.default_rng(3)
    alpha_stab = 5.0 + 20.0 * rng.random(mesh.ne)                      # any positive field: the forms are linear in it
    A, F = o.assemble(alpha_stab)
    A = A.tocsr()
    cp = c_problem(order, prob)
    nenr = len(prob.enrich_desc)
    nv = o.n_p + nenr
    # active enrichment bits per element: marked and not pruned
    act = np.zeros(mesh.ne, dtype=np.uint8)
    for i in range(nenr):
        for el in np.nonzero(o.h.el_mark[i])[0]:
            d = o.offQ[i] + o.q_num[i][el]
            if o.active[d]:
                act[el] |= 1 << i
    tris = np.ascontiguousarray(mesh.tris, np.int32)
    e2f = np.ascontiguousarray(mesh.el2facet, np.int32)
    f2e = np.ascontiguousarray(mesh.facet2el, np.int32)
    scale = abs(A).max()
    for el in range(mesh.ne):
        D, O, Fl = np.zeros((nv, nv)), np.zeros((3, nv, nv)), np.zeros(nv)
        L.hc_edg_element(C.byref(cp), dptr(mesh.verts), iptr(tris), iptr(e2f), iptr(f2e), act.ctypes.data_as(C.c_void_p),
                         dptr(alpha_stab), C.c_int(el), dptr(D), dptr(O), dptr(Fl))
        d_own = o.el_dofs(el)
        keep = np.array([bool(d >= 0 and (i < o.n_p or (act[el] >> (i - o.n_p)) & 1)) for i, d in enumerate(d_own)], dtype=bool)
        ref = A[d_own[keep]][:, d_own[keep]].toarray()
        assert np.abs(D[np.ix_(keep, keep)] - ref).max() < 1e-11 * scale, ("diag", el)
        assert np.abs(Fl[keep] - F[d_own[keep]]).max() < 1e-12 * max(1.0, np.abs(F).max())
        assert np.all(D[~keep] == 0.0) and np.all(D[:, ~keep] == 0.0)
        for e in range(3):
            f = mesh.el2facet[el, e]
            nb = mesh.facet2el[f]
            other = nb[1] if nb[0] == el else nb[0]
            if other < 0:
                assert np.all(O[e] == 0.0)
                continue
            d_nb = o.el_dofs(other)
            keep_nb = np.array([bool(d >= 0 and (i < o.n_p or (act[other] >> (i - o.n_p)) & 1)) for i, d in enumerate(d_nb)], dtype=bool)
            ref = A[d_own[keep]][:, d_nb[keep_nb]].toarray()
            assert np.abs(O[e][np.ix_(keep, keep_nb)] - ref).max() < 1e-11 * scale, ("off", el, e)


@pytest.mark.parametrize("order,eps,beta,theta", [(2, 2e-2, (2.0, 1.0), 1e-5), (1, 1e-2, (2.0, 0.001), 1e-5), (3, 1e-1, (2.0, 1.0), 0.2)])
def test_edg_pruning_and_alpha_match_the_oracle(order, eps, beta, theta):
    mesh = unit_square_mesh(4, kind="structured")
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=True)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob, theta=theta)
    o.prune()
    cp = c_problem(order, prob)
    nenr = len(prob.enrich_desc)
    act = np.zeros(mesh.ne, dtype=np.int64)
    n_marked = 0
    for el in range(mesh.ne):
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        elmark = sum(1 << i for i in range(nenr) if o.h.el_mark[i][el])
        if not elmark:
            continue
        n_marked += 1
        fac = np.zeros(nenr)
        drop = L.hc_edg_prune(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_int(elmark), C.c_double(theta), dptr(fac))
        got = [fac[i] for i in range(nenr) if (elmark >> i) & 1]
        ref = o.factors[el]
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            # the factor is a difference of O(1) terms: absolute agreement, and the same side of theta unless it sits on it
            assert abs(a - b) < 1e-6 * max(1.0, abs(b)) or abs(b - theta) < 1e-6
        for i in range(nenr):
            if (elmark >> i) & 1:
                kept_ref = bool(o.active[o.offQ[i] + o.q_num[i][el]])
                kept = not ((drop >> i) & 1)
                assert kept == kept_ref or min(abs(f - theta) for f in ref) < 1e-6
                if kept_ref:
                    act[el] |= 1 << i
    assert n_marked >= 4
    # alpha: generalised eigenvalue of (m, a + f f^T) with the bonus rule on the kept dofs (reference :129-163)
    lam = o.alpha_all()
    checked = 0
    for el in range(1, mesh.ne):            # element 0: `dof > 0` drops its constant mode (exclude_const)
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        st = C.c_int(0)
        got = L.hc_alpha_bonus(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_int(int(act[el])), C.c_int(0), C.byref(st))
        if st.value == 0:
            tol = 1e-10 if act[el] == 0 else 1e-5       # enriched pencils are ill-conditioned
            assert abs(got - lam[el]) < tol * lam[el], (el, got, lam[el])
            checked += 1
    assert checked >= mesh.ne // 2


@pytest.mark.parametrize("order,enriched,kind", [(1, False, "unstructured"), (2, True, "structured"), (2, True, "unstructured")])
def test_edg_block_csr_equals_the_oracle_matrix(order, enriched, kind):
    """element-major block rows (edg_block_neighbours / edg_fill_row): the whole matrix against the oracle's, pattern bit-exact
    under the permutation to NGSolve's numbering, values and right-hand side to round-off, and the solve reproduces the error"""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    mesh = unit_square_mesh(4, kind=kind)
    prob = orc.run_py_problem(beta=(2.0, 1.0), eps=2e-2, enriched=enriched)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    o.prune()
    alpha_stab = o.alpha_all()
    A, F = o.assemble(alpha_stab)
    nenr = len(prob.enrich_desc)
    act = np.zeros(mesh.ne, dtype=np.uint8)
    for i in range(nenr):
        for el in np.nonzero(o.h.el_mark[i])[0]:
            if o.active[o.offQ[i] + o.q_num[i][el]]:
                act[el] |= 1 << i
    cp = c_problem(order, prob)
    tris = np.ascontiguousarray(mesh.tris, np.int32)
    e2f = np.ascontiguousarray(mesh.el2facet, np.int32)
    f2e = np.ascontiguousarray(mesh.facet2el, np.int32)
    nr, nz = C.c_longlong(0), C.c_longlong(0)
    actp = act.ctypes.data_as(C.c_void_p)
    L.hc_edg_csr_sizes(C.byref(cp), C.c_int(mesh.ne), iptr(e2f), iptr(f2e), actp, C.byref(nr), C.byref(nz))
    n, nnz = nr.value, nz.value
    assert n == int(o.active.sum())
    rowptr, colind = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int32)
    vals, rhs = np.zeros(nnz), np.zeros(n)
    L.hc_edg_csr(C.byref(cp), C.c_int(mesh.ne), dptr(mesh.verts), iptr(tris), iptr(e2f), iptr(f2e), actp, dptr(alpha_stab),
                 rowptr.ctypes.data_as(C.c_void_p), iptr(colind), dptr(vals), dptr(rhs))
    assert rowptr[-1] == nnz
    B = sps.csr_matrix((vals, colind, rowptr), shape=(n, n))
    for r in range(n):
        assert np.all(np.diff(colind[rowptr[r]:rowptr[r + 1]]) > 0)       # columns ascending in every row
    # our row -> NGSolve dof: element-major over the active local slots
    ng = []
    for el in range(mesh.ne):
        d = o.el_dofs(el)
        for k, dof in enumerate(d):
            if dof >= 0 and (k < o.n_p or (act[el] >> (k - o.n_p)) & 1):
                ng.append(dof)
    ng = np.array(ng)
    assert len(ng) == n and np.array_equal(np.sort(ng), np.nonzero(o.active)[0])
    perm = np.argsort(ng)
    Bp = B[perm][:, perm].tocsr()
    Bp.sort_indices()
    actd = np.nonzero(o.active)[0]
    Ao = A[actd][:, actd].tocsr()
    Ao.sort_indices()
    assert np.array_equal(Bp.indptr, Ao.indptr) and np.array_equal(Bp.indices, Ao.indices)      # pattern: bit-exact
    assert np.abs(Bp.data - Ao.data).max() < 1e-11 * np.abs(Ao.data).max()
    assert np.abs(rhs[perm] - F[actd]).max() < 1e-12 * max(1.0, np.abs(F).max())
    x = np.zeros(o.ndof)
    x[ng] = spla.spsolve(B.tocsc(), rhs)
    e_ref = o.l2_error(o.solve(alpha_stab))
    assert abs(o.l2_error(x) - e_ref) < 1e-8 * e_ref
