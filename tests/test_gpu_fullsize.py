"""Parity at BASELINE.json's full sizes through size-independent properties and sampled oracle checks
(the numpy oracle cannot reach these sizes; the C oracle checks a strided sample)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BETA = (2.0, 0.001)


def _pipe(n, kind, order, eps, enriched=True):
    from bench import make_spec
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    mesh = unit_square_mesh(n, kind=kind)
    return mesh, EHDGPipeline(mesh, make_spec(order, eps, enriched=enriched))


@pytest.mark.parametrize("n,kind,order,eps", [(708, "structured", 4, 1e-6), (1414, "unstructured", 4, 1e-6), (224, "structured", 3, 1e-4)])
def test_sampled_oracle_parity_at_full_size(n, kind, order, eps):
    import torch
    from oracle import ehdg_cpu
    from oracle import ehdg_oracle as orc
    mesh, P = _pipe(n, kind, order, eps)
    P.setup().assemble()
    torch.cuda.synchronize()
    prob = orc.run_py_problem(beta=BETA, eps=eps)
    # marking: vectorised restatement of mark_elements / mark_element_bnd
    em, fm = ehdg_cpu.marks(mesh, prob)
    assert np.array_equal(P.elem_mask.cpu().numpy(), em)
    assert np.array_equal(P.facet_mask.cpu().numpy()[:mesh.nf], fm)
    slot = P.eslot().cpu().numpy()
    poly = (slot & 0x40000000) == 0
    assert poly.sum() == P.plan_sizes.n_poly and (~poly).sum() == P.plan_sizes.n_gen
    # general path = exactly the elements with an enriched facet
    assert np.array_equal(~poly, (fm[mesh.el2facet] != 0).any(axis=1))
    # strided sample of polynomial-path elements against the C oracle (alpha, S_K, g_K)
    els = np.nonzero(poly)[0][:: max(1, int(poly.sum()) // 1500)].astype(np.int32)
    co = ehdg_cpu.COracle(order, prob)
    ref = co.batch(mesh.verts, mesh.tris, els, np.zeros(len(els), np.uint8), np.zeros((len(els), 3), np.uint8))
    lam = P.lam.cpu().numpy()
    assert np.max(np.abs(lam[els] - ref["lam"]) / ref["lam"]) < 1e-11
    ref = co.batch(mesh.verts, mesh.tris, els, np.zeros(len(els), np.uint8), np.zeros((len(els), 3), np.uint8), lam=lam[els])
    nfp, nb, nfl = P.sizes.nf_poly, P.sizes.nb, P.sizes.nfl
    keep = np.concatenate([e * nb + np.arange(nfl) for e in range(3)])
    idx = torch.as_tensor(slot[els].astype(np.int64), device=P.dev)
    S = P.S[: P.plan_sizes.n_poly * nfp * nfp].view(-1, nfp, nfp)[idx].cpu().numpy()
    g = P.g[: P.plan_sizes.n_poly * nfp].view(-1, nfp)[idx].cpu().numpy()
    So = ref["S"][:, keep][:, :, keep]
    go = ref["g"][:, keep]
    errS = np.linalg.norm((S - So).reshape(len(els), -1), axis=1) / np.linalg.norm(So.reshape(len(els), -1), axis=1)
    assert errS.max() < 1e-10, errS.max()       # eps = 1e-6: cond(A_VV) ~ 1e4 on these elements
    gn = np.linalg.norm(go, axis=1)
    errg = np.linalg.norm(g - go, axis=1)[gn > 0] / gn[gn > 0]
    assert errg.size == 0 or errg.max() < 1e-8
    # pruning decisions, independently of the device: the numpy oracle's literal loop (conv_diff.py:286-310) on a strided
    # sample of marked elements and on their marked facet neighbours (a facet dof is switched off by either side)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    ea = P.elem_active.cpu().numpy()
    fa = P.facet_active.cpu().numpy()[:mesh.nf]
    marked = np.nonzero(o.el_marked_any)[0]
    sample = marked[:: max(1, len(marked) // 100)]
    need = set(int(e) for e in sample)
    for el in sample:
        for f in mesh.el2facet[el]:
            for nb in mesh.facet2el[f]:
                if nb >= 0 and o.el_marked_any[nb]:
                    need.add(int(nb))
    dec = {el: set(o.prune_element(el)[0]) for el in sorted(need)}
    checked = 0
    for el in sample:
        for i in range(o.nenr):
            if o.el_mark[i][el]:
                d = o.offQ[i] + o.q_num[i][el]
                assert ((ea[el] >> i) & 1) == (0 if d in dec[int(el)] else 1), ("volume enrichment", el, i)
                checked += 1
            for f in mesh.el2facet[el]:
                if not o.fac_mark[i][f]:
                    continue
                d = o.offQF[i] + o.qf_num[i][f]
                off = (not o.free[d]) or any(d in dec[int(nb)] for nb in mesh.facet2el[f] if nb >= 0 and o.el_marked_any[nb])
                assert ((fa[f] >> i) & 1) == (0 if off else 1), ("facet enrichment", el, f, i)
                checked += 1
    assert checked > 0
    # enriched-path sample: S_K with the device's alpha, element by element, to max(1e-11, 100 eps_mach cond(A_VV)) with the
    # condition number of the oracle's own A_VV (active volume dofs)
    gel = np.nonzero(~poly)[0][:: max(1, int((~poly).sum()) // 300)].astype(np.int32)
    refg = co.batch(mesh.verts, mesh.tris, gel, ea[gel], fm[mesh.el2facet[gel]], lam=lam[gel], want_full=True)
    nfg = P.sizes.nfmax
    nvg = P.sizes.nvmax
    idxg = torch.as_tensor((slot[gel] & 0x3fffffff).astype(np.int64), device=P.dev)
    Sg = P.S[P.plan_sizes.n_poly * nfp * nfp:].view(-1, nfg, nfg)[idxg].cpu().numpy()
    err = np.linalg.norm((Sg - refg["S"]).reshape(len(gel), -1), axis=1) / np.linalg.norm(refg["S"].reshape(len(gel), -1), axis=1)
    Nloc = nvg + nfg
    worst = 0.0
    stg = P.status.cpu().numpy()
    n_dropped = 0
    for t, el in enumerate(gel):
        if stg[el] & 8:              # a dependent volume enrichment left this element's space (ehdg_set_drop_tol): S_K is that of
            n_dropped += 1           # the reduced space, the oracle's of the singular one
            continue
        Afull = refg["Afull"][t][: Nloc * Nloc].reshape(Nloc, Nloc)
        act = [i for i in range(nvg) if i < P.sizes.n_p or (ea[el] >> (i - P.sizes.n_p)) & 1]
        cond = np.linalg.cond(Afull[np.ix_(act, act)])
        bound = max(1e-11, 100.0 * 2.2e-16 * cond)
        worst = max(worst, err[t] / bound)
        assert err[t] <= bound, (int(el), err[t], cond)
    assert n_dropped < len(gel)
    assert np.all(np.isfinite(P.S.cpu().numpy() if mesh.ne < 200000 else Sg))


def test_fast_and_quadrature_paths_agree_at_scale():
    """100k elements through both product paths: tensor path == quadrature path (same alpha)."""
    import torch
    mesh, Pf = _pipe(224, "unstructured", 4, 1e-6, enriched=False)
    Pf.setup().assemble()
    from bench import make_spec
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    Pq = EHDGPipeline(mesh, make_spec(4, 1e-6, enriched=False), force_general=True)
    Pq.setup().alpha()
    lam_f, lam_q = Pf.lam, Pq.lam
    assert float(((lam_f - lam_q).abs() / lam_q).max()) < 1e-11
    Pq.lam.copy_(lam_f)
    Pq.assemble_condense()
    nf = Pf.sizes.nf_poly
    Sf = Pf.S.view(-1, nf * nf)
    Sq = Pq.S.view(-1, nf * nf)
    err = ((Sf - Sq).norm(dim=1) / Sq.norm(dim=1)).max()
    assert float(err) < 1e-10
    gerr = (Pf.g - Pq.g).norm() / Pq.g.norm()
    assert float(gerr) < 1e-9
    assert int(Pf.status.abs().sum()) == 0 and int(Pq.status.abs().sum()) == 0


def test_condensed_system_properties_at_1m():
    """CSR invariants at the 1M-element size: sorted columns, exact nnz formula, conservation of the scattered blocks,
    linearity of the SpMV, back-solve of the zero facet solution returns z."""
    import torch
    mesh, P = _pipe(708, "structured", 4, 1e-6, enriched=False)
    P.setup().assemble().csr_build()
    n, nnz = P.csr_sizes.n_rows, P.csr_sizes.nnz
    N = 708
    assert n == (mesh.nf - 4 * N) * 5
    rp, ci, rs, dm = P.csr_arrays()
    assert int(rp[-1]) == nnz and bool((rp[1:] > rp[:-1]).all())
    # columns strictly ascending inside every row
    d = ci[1:].to(torch.int64) - ci[:-1].to(torch.int64)
    row_end = torch.zeros(nnz, dtype=torch.bool, device=P.dev)
    row_end[(rp[1:-1] - 1).long()] = True
    assert bool((d[~row_end[:-1]] > 0).all())
    # sum of all matrix entries == sum over elements of the interior-facet part of S_K (each entry lands exactly once)
    nf = P.sizes.nf_poly
    interior = torch.as_tensor(mesh.facet2el[:, 1] >= 0, device=P.dev)
    e2f = torch.as_tensor(mesh.el2facet.astype(np.int64), device=P.dev)
    act = interior[e2f].repeat_interleave(5, dim=1).to(torch.float64)          # [ne, 15] active local facet dofs
    S = P.S.view(-1, nf, nf)
    total_blocks = torch.einsum("ek,ekl,el->", act, S, act)
    total_csr = P.vals[:nnz].sum()
    assert abs(float(total_blocks - total_csr)) <= 1e-9 * float(P.vals[:nnz].abs().sum())
    gsum = (P.g.view(-1, nf) * act).sum()
    assert abs(float(gsum - P.rhs[:n].sum())) <= 1e-9 * float(P.rhs[:n].abs().sum() + 1e-300)
    # SpMV linearity
    x = torch.randn(n, dtype=torch.float64, device=P.dev, generator=torch.Generator(device=P.dev).manual_seed(1))
    y = torch.randn(n, dtype=torch.float64, device=P.dev, generator=torch.Generator(device=P.dev).manual_seed(2))
    lhs = P.spmv(2.0 * x - 3.0 * y)
    rhs = 2.0 * P.spmv(x) - 3.0 * P.spmv(y)
    assert float((lhs - rhs).norm() / rhs.norm()) < 1e-13
    # back-solve with a zero facet solution returns z
    P.xhat = torch.zeros(n, dtype=torch.float64, device=P.dev)
    P.backsolve()
    assert torch.equal(P.u, P.z)


@pytest.mark.parametrize("n,kind,order,eps", [(708, "structured", 4, 1e-6), (1414, "unstructured", 4, 1e-6)])
def test_fused_store_equals_two_pass_fill_at_full_size(n, kind, order, eps):
    """a5-a8 at BASELINE configs 3 and 4: the value array and the right-hand side written by the element kernels
    (ehdg_assemble_condense_csr) are identical to assemble_condense + csr_numeric, entry by entry; every entry is written
    (the array starts as NaN); and a checksum of checksums of A x through two different routes agrees."""
    import torch
    mesh, P = _pipe(n, kind, order, eps)
    P.setup()
    P.csr_pattern()
    P.alpha().assemble_condense()
    P.csr_numeric()
    nnz, nr = P.csr_sizes.nnz, P.csr_sizes.n_rows
    v0 = P.vals[:nnz].clone()
    b0 = P.rhs[:nr].clone()
    P.S = P.g = None
    torch.cuda.empty_cache()
    P.vals.fill_(float("nan"))
    P.rhs.fill_(float("nan"))
    P.assemble_condense_csr()
    torch.cuda.synchronize()
    assert not torch.isnan(P.vals[:nnz]).any()
    assert torch.equal(P.vals[:nnz], v0) and torch.equal(P.rhs[:nr], b0)
    # checksum of checksums: sum_r (A x)_r through the facet-block SpMV against sum_nnz a_rc x_c through the exported CSR arrays
    # (fills the lazily written column indices on the way), both in fp64 with different reduction orders
    g = torch.Generator(device=P.dev).manual_seed(7)
    x = torch.rand(nr, dtype=torch.float64, device=P.dev, generator=g)
    y = P.spmv(x)
    rp, ci, _, _ = P.csr_arrays()
    assert int(rp[-1]) == nnz
    tot, scale = 0.0, 0.0
    for a in range(0, nnz, 1 << 27):
        b = min(nnz, a + (1 << 27))
        t = P.vals[a:b] * x[ci[a:b].long()]
        tot += float(t.sum())
        scale += float(t.abs().sum())
    assert abs(float(y.sum()) - tot) < 1e-12 * scale
    P.close()
