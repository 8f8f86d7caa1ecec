"""CPU checks of the product's element math: the device routines of csrc/ (quadrature path and
tensor path) compiled for the host by tests/hostcheck and compared with the numpy oracle.  This is
host-logic coverage; the GPU build of the same source is covered by the `-m gpu` parity tests."""
import ctypes as C

import numpy as np
import pytest
import scipy.special as sp

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from tests.common import rel
from tests.hostcheck.binding import c_problem, dptr, iptr, lib

L = lib()
L.hc_fast_alpha.restype = C.c_double


@pytest.mark.parametrize("order", [0, 1, 2, 7, 8, 14, 18, 22, 60])
def test_quadrature_rules_match_scipy(order):
    n = order // 2 + 1
    x, y, w = np.zeros(n * n), np.zeros(n * n), np.zeros(n * n)
    assert L.hc_trig_rule(order, dptr(x), dptr(y), dptr(w)) == n * n
    pts, W = orc.trig_rule(order)
    assert np.abs(x - pts[:, 0]).max() < 1e-15 and np.abs(y - pts[:, 1]).max() < 1e-15
    assert np.abs(w - W).max() < 1e-13 * W.max()
    t, ws = np.zeros(n), np.zeros(n)
    assert L.hc_seg_rule(order, dptr(t), dptr(ws)) == n
    tt, ww = orc.seg_rule(order)
    assert np.abs(t - tt).max() < 1e-15 and np.abs(ws - ww).max() < 1e-13 * ww.max()
    assert abs(w.sum() - 0.5) < 1e-14 and abs(ws.sum() - 1.0) < 1e-14


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_dubiner_recurrence_matches_symbolic_basis(p):
    rng = np.random.default_rng(p)
    X = rng.uniform(0, 1, 20)
    Y = rng.uniform(0, 1, 20) * (1 - X)
    n_p = (p + 1) * (p + 2) // 2
    v, dX, dY = np.zeros((20, n_p)), np.zeros((20, n_p)), np.zeros((20, n_p))
    L.hc_dubiner(p, 20, dptr(X), dptr(Y), dptr(v), dptr(dX), dptr(dY))
    vo, dXo, dYo = orc.dubiner_eval(p, X, Y)
    assert np.abs(v.T - vo).max() < 1e-12 and np.abs(dX.T - dXo).max() < 1e-11 and np.abs(dY.T - dYo).max() < 1e-11


def _alpha_cond(o, el):
    """condition number of a + f f^T on the dofs the reference keeps (conv_diff.py:336-342)"""
    a, f, _ = o.alpha_matrices(el)
    a = a + np.outer(f, f)
    imp = [i for i, d in enumerate(o.el_dofs(el)) if d > 0 and o.active[d] and a[i, i] != 0.0]
    return np.linalg.cond(a[np.ix_(imp, imp)])


def _masks(o, el):
    dofs = o.el_dofs(el)
    volmask, elmark, facmask = 0, 0, np.zeros(3, np.int32)
    for i in range(o.nenr):
        d = dofs[o.loc_Q(i)]
        if d >= 0:
            elmark |= 1 << i
            if o.active[d]:
                volmask |= 1 << i
        for e in range(3):
            if dofs[o.loc_QF(i, e)] >= 0:
                facmask[e] |= 1 << i
    return volmask, elmark, facmask


@pytest.mark.parametrize("n,order,eps,beta,kind", [(3, 2, 1e-2, (2.0, 1.0), "structured"), (2, 4, 1e-6, (2.0, 0.001), "structured"),
                                                   (3, 3, 1e-4, (2.0, 0.001), "unstructured")])
def test_quadrature_path_element_blocks(n, order, eps, beta, kind):
    mesh = unit_square_mesh(n, kind=kind)
    prob = orc.run_py_problem(beta=beta, eps=eps)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    o.prune()
    cp = c_problem(order, prob)
    sz = np.zeros(4, np.int32)
    L.hc_sizes(C.byref(cp), iptr(sz))
    n_p, nb, nv, nf = [int(s) for s in sz]
    vidx = list(o.loc_V()) + [o.loc_Q(i) for i in range(o.nenr)]
    fidx = sum([list(o.loc_F(e)) + [o.loc_QF(i, e) for i in range(o.nenr)] for e in range(3)], [])
    for el in range(mesh.ne):
        volmask, elmark, facmask = _masks(o, el)
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        st = C.c_int(0)
        lam = L.hc_alpha(C.byref(cp), dptr(mesh.verts), iptr(tri), volmask, int(el == 0), C.byref(st))
        lam_o = o.alpha_lambda(el)
        if st.value == 0:
            assert abs(lam - lam_o) < max(1e-10, 1e-13 * _alpha_cond(o, el)) * lam_o
        ldA = nv + nf + 1
        S, g, W, z = np.zeros((nf, nf)), np.zeros(nf), np.zeros((nv, nf)), np.zeros(nv)
        Af = np.zeros(nv * ldA + nf * nv + nf * nf)
        L.hc_assemble(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_double(lam_o), volmask, iptr(facmask), dptr(S), dptr(g),
                      dptr(W), dptr(z), dptr(Af))
        A, F = o.element_matrix(el, lam_o)
        vact = np.array([True] * n_p + [bool((volmask >> i) & 1) for i in range(o.nenr)])
        fact = np.concatenate([[True] * (order + 1) + [bool((facmask[e] >> i) & 1) for i in range(o.nenr)] for e in range(3)])
        act = np.concatenate([vact, fact])
        Ao = A[np.ix_(vidx + fidx, vidx + fidx)].copy()
        Ao[~act, :] = 0
        Ao[:, ~act] = 0
        Aug = Af[:nv * ldA].reshape(nv, ldA)
        Avv = Aug[:, :nv].copy()
        Avv[~vact, ~vact] = 0
        mine = np.block([[Avv, Aug[:, nv:nv + nf]], [Af[nv * ldA:nv * ldA + nf * nv].reshape(nf, nv), Af[nv * ldA + nf * nv:].reshape(nf, nf)]])
        assert rel(mine, Ao) < 1e-12
        assert rel(Aug[:, -1], F[vidx] * vact) < 1e-12
        iv, jf = np.nonzero(vact)[0], np.nonzero(fact)[0]
        Avv_o = Ao[np.ix_(iv, iv)]
        cond = np.linalg.cond(Avv_o)
        So = Ao[np.ix_(nv + jf, nv + jf)] - Ao[np.ix_(nv + jf, iv)] @ np.linalg.solve(Avv_o, Ao[np.ix_(iv, nv + jf)])
        assert rel(S[np.ix_(jf, jf)], So) < max(1e-11, 1e-14 * cond)


def test_pruning_factors_and_decisions():
    mesh = unit_square_mesh(4)
    prob = orc.run_py_problem(eps=0.01)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, 2, prob)
    o.prune()
    cp = c_problem(2, prob)
    assert (o.free & ~o.active).sum() > 0              # the pruning path is live on this mesh
    for el, lst in o.prune_factors.items():
        _, elmark, facmask = _masks(o, el)
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        vd, fd, fac = C.c_int(0), np.zeros(3, np.int32), np.zeros(8)
        L.hc_prune(C.byref(cp), dptr(mesh.verts), iptr(tri), elmark, iptr(facmask), C.c_double(1e-5), C.byref(vd), iptr(fd), dptr(fac))
        got = {}
        for i in range(o.nenr):
            got[o.loc_Q(i)] = (fac[4 * i], (vd.value >> i) & 1)
            for e in range(3):
                got[o.loc_QF(i, e)] = (fac[4 * i + 1 + e], (fd[e] >> i) & 1)
        for idx, val in lst:
            if idx in got:
                if val < 0:                      # psi vanishes on that facet: skipped on both sides (factor -3)
                    assert got[idx] == (val, 0)
                    continue
                assert abs(got[idx][0] - val) < 1e-7 + 1e-9 * val
                assert got[idx][1] == int(val <= 1e-5)


@pytest.mark.parametrize("order,eps,kind", [(1, 1e-2, "structured"), (2, 1e-2, "unstructured"), (3, 1e-4, "unstructured"),
                                            (4, 1e-6, "unstructured"), (5, 1e-3, "unstructured"), (6, 1e-2, "structured")])
def test_tensor_path_matches_quadrature_oracle(order, eps, kind):
    mesh = unit_square_mesh(2, kind=kind)
    prob = orc.run_py_problem(eps=eps, enriched=False)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    cp = c_problem(order, prob)
    n_p, nf = o.n_p, 3 * (order + 1)
    for el in range(mesh.ne):
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        st = C.c_int(0)
        lam = L.hc_fast_alpha(C.byref(cp), dptr(mesh.verts), iptr(tri), C.byref(st))
        lam_o = o.alpha_lambda(el)
        assert st.value == 0 and abs(lam - lam_o) < 1e-12 * lam_o
        S, g, W, z = np.zeros((nf, nf)), np.zeros(nf), np.zeros((n_p, nf)), np.zeros(n_p)
        assert L.hc_fast_assemble(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_double(lam_o), dptr(S), dptr(g), dptr(W), dptr(z)) == 0
        A, F = o.element_matrix(el, lam_o)
        iv, jf = np.arange(n_p), np.arange(n_p, n_p + nf)
        cond = np.linalg.cond(A[np.ix_(iv, iv)])
        tol = max(1e-11, 1e-14 * cond)
        Wo = np.linalg.solve(A[np.ix_(iv, iv)], A[np.ix_(iv, jf)])
        zo = np.linalg.solve(A[np.ix_(iv, iv)], F[iv])
        assert rel(S, A[np.ix_(jf, jf)] - A[np.ix_(jf, iv)] @ Wo) < tol
        assert rel(W, Wo) < tol and rel(z, zo) < 10 * tol and rel(g, -A[np.ix_(jf, iv)] @ zo) < 10 * tol


def test_tensor_path_all_vertex_order_classes():
    """random vertex renumbering and local rotations exercise all 6 sort classes of the Dubiner frame"""
    import convectiondiffusionequation_b200.mesh as M
    base = unit_square_mesh(2, kind="unstructured")
    rng = np.random.default_rng(1)
    perm = rng.permutation(base.nv)
    verts = np.zeros_like(base.verts)
    verts[perm] = base.verts
    tris = np.array([np.roll(t, r) for t, r in zip(perm[base.tris], rng.integers(0, 3, size=base.ne))])
    old = M.normalise_triangles
    M.normalise_triangles = lambda v, t: np.asarray(t, np.int32)
    try:
        mesh = M.build_topology(verts, tris, 0.5, "perm")
    finally:
        M.normalise_triangles = old
    classes = {tuple(np.argsort(t)) for t in mesh.tris}
    assert len(classes) >= 4
    prob = orc.run_py_problem(eps=1e-3, enriched=False)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, 3, prob)
    cp = c_problem(3, prob)
    for el in range(mesh.ne):
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        lam_o = o.alpha_lambda(el)
        S, g, W, z = np.zeros((12, 12)), np.zeros(12), np.zeros((10, 12)), np.zeros(10)
        L.hc_fast_assemble(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_double(lam_o), dptr(S), dptr(g), dptr(W), dptr(z))
        c = o.condense(el, lam_o)
        A, F = c["A"], c["F"]
        iv, jf = np.arange(10), np.arange(10, 22)
        Wo = np.linalg.solve(A[np.ix_(iv, iv)], A[np.ix_(iv, jf)])
        assert rel(S, A[np.ix_(jf, jf)] - A[np.ix_(jf, iv)] @ Wo) < 1e-11
        assert rel(g, -A[np.ix_(jf, iv)] @ np.linalg.solve(A[np.ix_(iv, iv)], F[iv])) < 1e-10


@pytest.mark.parametrize("eps,beta", [(1e-6, (2.0, 1.0)), (5e-3, (2.0, 1.0)), (2e-3, (0.5, 2.0))])
def test_tensor_path_moment_rhs_matches_quadrature_oracle(eps, beta):
    """Where exp(-k d) < 2e-22 on the whole element (k d_min > 50) the device takes f_V from six moment vectors instead of
    the quadrature loop (csrc/ehdg_fast.h); the oracle always integrates (convection_diffusion.py:380-383).  The
    intermediate k = beta/eps of the second and third case put elements with 50 < k d < 746 on the moment path,
    where exp() has not underflowed yet but cannot change a double next to the polynomial part."""
    import convectiondiffusionequation_b200.mesh as M
    from convectiondiffusionequation_b200.hot_path import LAYER_DROP
    base = unit_square_mesh(4, kind="unstructured")
    rng = np.random.default_rng(7)
    perm = rng.permutation(base.nv)
    verts = np.zeros_like(base.verts)
    verts[perm] = base.verts
    tris = np.array([np.roll(t, r) for t, r in zip(perm[base.tris], rng.integers(0, 3, size=base.ne))])
    old = M.normalise_triangles
    M.normalise_triangles = lambda v, t: np.asarray(t, np.int32)
    try:
        mesh = M.build_topology(verts, tris, 0.25, "perm")
    finally:
        M.normalise_triangles = old
    order = 3
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    cp = c_problem(order, prob)
    n_p, nf = o.n_p, 3 * (order + 1)
    dmin = (1.0 - mesh.verts[mesh.tris]).min(axis=1)                     # [ne, 2]
    moment = (dmin[:, 0] * beta[0] / eps > LAYER_DROP) & (dmin[:, 1] * beta[1] / eps > LAYER_DROP)
    classes = {tuple(np.argsort(t)) for t in mesh.tris[moment]}
    assert moment.sum() >= 8 and (~moment).sum() >= 4 and len(classes) >= 3
    if eps > 1e-6:      # some of them sit where exp() is far from underflow
        kd = np.minimum(dmin[:, 0] * beta[0] / eps, dmin[:, 1] * beta[1] / eps)
        assert ((kd > LAYER_DROP) & (kd < 746.0)).sum() >= 4
    for el in range(mesh.ne):
        tri = np.ascontiguousarray(mesh.tris[el], np.int32)
        lam_o = o.alpha_lambda(el)
        S, g, W, z = np.zeros((nf, nf)), np.zeros(nf), np.zeros((n_p, nf)), np.zeros(n_p)
        assert L.hc_fast_assemble(C.byref(cp), dptr(mesh.verts), iptr(tri), C.c_double(lam_o), dptr(S), dptr(g), dptr(W), dptr(z)) == 0
        A, F = o.element_matrix(el, lam_o)
        iv, jf = np.arange(n_p), np.arange(n_p, n_p + nf)
        tol = max(1e-11, 1e-14 * np.linalg.cond(A[np.ix_(iv, iv)]))
        zo = np.linalg.solve(A[np.ix_(iv, iv)], F[iv])
        assert rel(z, zo) < 10 * tol and rel(g, -A[np.ix_(jf, iv)] @ zo) < 10 * tol, (el, bool(moment[el]))
