// EDG (enriched interior-penalty DG) element routine -- GROUNDWORK for SURVEY.md 8(f) row 1, not wired into libehdg yet.
//
// Restates the forms of the reference's _solveEDG (Python/convection_diffusion.py:165-198) for ONE block row of the
// global matrix: the diagonal block of element K and the blocks that couple K's test functions to the trial functions
// of its (up to three) facet neighbours, plus K's right-hand side.  Every interior facet is visited from both sides,
// each side writing its own block row: no atomics, a deterministic result.  Written in the lock-step style of
// ehdg_general.h (lanes of a Coop stride over points and matrix entries), so the same source is checked on the host
// against oracle/edg_oracle.py (tests/test_hostcheck_edg.py) before it meets a GPU.
//
// Conventions (those of the oracle; NGSolve decides them, see its header): on an interior facet the normal, the mesh
// size h = |det J| / |e| and alpha_stab are taken from the FIRST element of the facet (facet2el[f][0]); facet points are
// parametrised from the lower to the higher global vertex number, which both neighbours agree on.
//   eps [ int grad u . grad v + sum_int ( pen [u][v] - {dn u}[v] - {dn v}[u] ) + sum_bnd ( pen u v - dn u v - dn v u ) ]
//   right-hand side: int coeff v (+ with pb.edg_bnd_data the boundary data of Python/debug_edg.py:191-192)
//   - int (b u) . grad v + sum_int (b.n) u_up [v],        pen = 4 alpha_stab p^2 / h^2,  [w] = w_first - w_second.
#pragma once
#include "ehdg_general.h"

namespace ehdg {

// doubles of workspace edg_element_blocks needs
EHDG_HD int edg_ws_doubles(const Problem& pb, int nq, int nqf) {
    const int vol = 3 * pb.nvmax * pad_odd(nq) + nq;
    const int fac = 4 * pb.nvmax * pad_odd(nqf) + nqf;
    return ((vol > fac ? vol : fac) + 1) & ~1;
}

// barycentrics of the point at parameter t (from the lower to the higher global vertex) of local edge e of an element
EHDG_HD void edg_facet_point(const ElemGeom& g, int e, double t, double* l) {
    int a, b;
    local_edge(e, a, b);
    const int lo = g.vn[a] < g.vn[b] ? a : b, hi = g.vn[a] < g.vn[b] ? b : a;
    l[0] = l[1] = l[2] = 0.0;
    l[lo] = 1.0 - t;
    l[hi] = t;
}

// D [nv x nv] (row-major, test x trial), O [3][nv x nv] (neighbour across local edge e; zero on the boundary), F [nv];
// nv = pb.nvmax = n_p + n_enrich, enrichment rows/columns of inactive functions stay zero.
// elem_active: u8 per element, bit k = enrichment k active on that element (after pruning); may be null.
EHDG_HD void edg_element_blocks(const Coop& co, const Problem& pb, const RuleSet& rs, const double* verts, const int* tris,
                                const int* el2facet, const int* facet2el, const unsigned char* elem_active,
                                const double* alpha_stab, int el, double* ws, double* D, double* O, double* F) {
    const int nv = pb.nvmax, nq = rs.nq, nqf = rs.nqf;
    const int ldq = pad_odd(nq), ldf = pad_odd(nqf);
    ElemGeom g;
    elem_geom(verts, tris + 3 * el, g);
    const int volmask = elem_active ? elem_active[el] : 0;
    for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
        D[idx] = 0.0;
        O[idx] = O[nv * nv + idx] = O[2 * nv * nv + idx] = 0.0;
    }
    for (int i = co.lane; i < nv; i += co.nl) F[i] = 0.0;
    // ---- volume: eps grad u . grad v - (b u) . grad v ; rhs
    {
        double* Tp = ws;
        double* Tgx = Tp + nv * ldq;
        double* Tgy = Tgx + nv * ldq;
        double* wq = Tgy + nv * ldq;
        for (int q = co.lane; q < nq; q += co.nl) {
            vol_shapes(pb, g, volmask, rs.vx[q], rs.vy[q], Tp + q, Tgx + q, Tgy + q, ldq);
            wq[q] = rs.vw[q] * g.adet;
        }
        co.sync();
        for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
            const int i = idx / nv, j = idx % nv;
            const double *gxi = Tgx + i * ldq, *gyi = Tgy + i * ldq, *gxj = Tgx + j * ldq, *gyj = Tgy + j * ldq, *pj = Tp + j * ldq;
            double acc = 0.0;
            for (int q = 0; q < nq; ++q)
                acc += wq[q] * (pb.eps * (gxi[q] * gxj[q] + gyi[q] * gyj[q]) - (pb.bx * gxi[q] + pb.by * gyi[q]) * pj[q]);
            D[idx] += acc;
        }
        for (int i = co.lane; i < nv; i += co.nl) {
            double acc = 0.0;
            for (int q = 0; q < nq; ++q) {
                const double l0 = rs.vx[q], l1 = rs.vy[q], l2 = 1.0 - l0 - l1;
                const double dx = l0 * g.omx[0] + l1 * g.omx[1] + l2 * g.omx[2];
                const double dy = l0 * g.omy[0] + l1 * g.omy[1] + l2 * g.omy[2];
                acc += wq[q] * problem_coeff(pb, dx, dy) * Tp[i * ldq + q];
            }
            F[i] = acc;
        }
        co.sync();
    }
    // ---- facets
    for (int e = 0; e < 3; ++e) {
        const int f = el2facet[3 * el + e];
        const int e1 = facet2el[2 * f], e2 = facet2el[2 * f + 1];
        const bool bnd = e2 < 0;
        const bool first = (e1 == el);
        const int other = bnd ? -1 : (first ? e2 : e1);
        EdgeGeom eg;
        edge_geom(g, e, eg);
        // normal, h and alpha_stab of the facet's first element
        double nx = eg.nx, ny = eg.ny, h1 = eg.h_e;
        ElemGeom go;
        int oe = 0, omask = 0;
        if (!bnd) {
            elem_geom(verts, tris + 3 * other, go);
            omask = elem_active ? elem_active[other] : 0;
            for (int k = 0; k < 3; ++k)
                if (el2facet[3 * other + k] == f) oe = k;
            if (!first) {
                EdgeGeom ego;
                edge_geom(go, oe, ego);
                nx = ego.nx;
                ny = ego.ny;
                h1 = ego.h_e;
            }
        }
        const double pen = 4.0 * alpha_stab[e1] / h1 * (double)(pb.p * pb.p) / h1;
        const double bn = pb.bx * nx + pb.by * ny;
        double* Up = ws;                    // own values at the facet points
        double* Ud = Up + nv * ldf;         // own derivative along the first element's normal
        double* Vp = Ud + nv * ldf;         // neighbour values
        double* Vd = Vp + nv * ldf;         // neighbour normal derivative
        double* we = Vd + nv * ldf;
        for (int q = co.lane; q < nqf; q += co.nl) {
            double l[3], ph[EHDG_MAX_NP + EHDG_MAX_ENRICH], gx[EHDG_MAX_NP + EHDG_MAX_ENRICH], gy[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            edg_facet_point(g, e, rs.st[q], l);
            vol_shapes(pb, g, volmask, l[0], l[1], ph, gx, gy, 1);
            for (int i = 0; i < nv; ++i) {
                Up[i * ldf + q] = ph[i];
                Ud[i * ldf + q] = gx[i] * nx + gy[i] * ny;
            }
            if (!bnd) {
                edg_facet_point(go, oe, rs.st[q], l);
                vol_shapes(pb, go, omask, l[0], l[1], ph, gx, gy, 1);
                for (int i = 0; i < nv; ++i) {
                    Vp[i * ldf + q] = ph[i];
                    Vd[i * ldf + q] = gx[i] * nx + gy[i] * ny;
                }
            }
            we[q] = rs.sw[q] * eg.len;
        }
        co.sync();
        if (bnd) {
            for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
                const int i = idx / nv, j = idx % nv;
                double acc = 0.0;
                for (int q = 0; q < nqf; ++q)
                    acc += we[q] * (pen * Up[i * ldf + q] * Up[j * ldf + q] - Up[i * ldf + q] * Ud[j * ldf + q] - Ud[i * ldf + q] * Up[j * ldf + q]);
                D[idx] += pb.eps * acc;
            }
            if (pb.edg_bnd_data) {
                // debug_edg.py:191-192: f += exact eps (pen v - dn v) dS + b.n IfPos(b.n, 0, -exact) v dS
                for (int i = co.lane; i < nv; i += co.nl) {
                    double acc = 0.0;
                    for (int q = 0; q < nqf; ++q) {
                        double l[3], ex, egx, egy;
                        edg_facet_point(g, e, rs.st[q], l);
                        const double dx = l[0] * g.omx[0] + l[1] * g.omx[1] + l[2] * g.omx[2];
                        const double dy = l[0] * g.omy[0] + l[1] * g.omy[1] + l[2] * g.omy[2];
                        problem_exact(pb, dx, dy, &ex, &egx, &egy);
                        double t = pb.eps * (pen * Up[i * ldf + q] - Ud[i * ldf + q]);
                        if (!(bn > 0.0)) t -= bn * Up[i * ldf + q];
                        acc += we[q] * ex * t;
                    }
                    F[i] += acc;
                }
            }
        } else {
            // own sign in the jump: + for the first element, - for the second
            const double so = first ? 1.0 : -1.0;
            const bool up_own = first ? (bn > 0.0) : !(bn > 0.0);      // u_up = first's trace iff b.n > 0
            double* On = O + (size_t)e * nv * nv;
            for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
                const int i = idx / nv, j = idx % nv;
                double dd = 0.0, oo = 0.0;
                for (int q = 0; q < nqf; ++q) {
                    const double vi = Up[i * ldf + q], dvi = Ud[i * ldf + q];
                    // trial = own (sign so), trial = neighbour (sign -so); test = own (sign so)
                    const double uj = Up[j * ldf + q], duj = Ud[j * ldf + q];
                    const double wj = Vp[j * ldf + q], dwj = Vd[j * ldf + q];
                    dd += we[q] * (pb.eps * (pen * vi * uj - 0.5 * so * vi * duj - 0.5 * so * dvi * uj) + (up_own ? bn * so * vi * uj : 0.0));
                    oo += we[q] * (pb.eps * (-pen * vi * wj - 0.5 * so * vi * dwj + 0.5 * so * dvi * wj) + (up_own ? 0.0 : bn * so * vi * wj));
                }
                D[idx] += dd;
                On[idx] += oo;
            }
        }
        co.sync();
    }
}

// EDG pruning of one marked element (reference :83-123): mass Gram  int u() v()  of the enriched volume shapes with the
// bonus rule; enrichment k (local dof n_p + k, in the order of the compound space [V, Q_0, Q_1, ...]) is dropped when
//   factor = sqrt |1 - 2 sum_j G_ij^2 / (G_ii G_jj) + sum_{j,k} G_ij G_ik G_jk / (G_ii G_jj G_kk)| <= theta
// over the dofs j, k < i that are still kept.  Unlike the EHDG loop there is no try/except around it.
// elmark: bits of the enrichments whose P0 dof exists on the element.  ws: edg_ws_doubles() + nv * nv doubles.
// factors[k] (optional): the factor of enrichment k, -1 where it has no dof.  Returns the bit mask of dropped enrichments.
EHDG_HD int edg_prune_element(const Coop& co, const Problem& pb, const RuleSet& rs, const double* verts, const int* tri,
                              int elmark, double theta, double* ws, double* factors) {
    const int nv = pb.nvmax, nq = rs.nq, ldq = pad_odd(nq);
    ElemGeom g;
    elem_geom(verts, tri, g);
    double* Tp = ws;
    double* wq = Tp + nv * ldq;
    double* G = ws + edg_ws_doubles(pb, rs.nq, rs.nqf);
    for (int q = co.lane; q < nq; q += co.nl) {
        vol_shapes(pb, g, elmark, rs.vx[q], rs.vy[q], Tp + q, (double*)0, (double*)0, ldq);
        wq[q] = rs.vw[q] * g.adet;
    }
    co.sync();
    for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
        const int i = idx / nv, j = idx % nv;
        double acc = 0.0;
        for (int q = 0; q < nq; ++q) acc += wq[q] * Tp[i * ldq + q] * Tp[j * ldq + q];
        G[idx] = acc;
    }
    co.sync();
    int drop = 0;                       // every lane runs the same sequential test
    for (int k = 0; k < pb.nenr; ++k) {
        double fac = -1.0;
        if ((elmark >> k) & 1) {
            const int i = pb.n_p + k;
            const double gii = G[i * nv + i];
            int act[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            int na = 0;
            for (int j = 0; j < pb.n_p; ++j) act[na++] = j;
            for (int kk = 0; kk < k; ++kk)
                if (((elmark >> kk) & 1) && !((drop >> kk) & 1)) act[na++] = pb.n_p + kk;
            double s1 = 0.0, s2 = 0.0;
            for (int a = 0; a < na; ++a) {
                const int j = act[a];
                s1 += G[i * nv + j] * G[i * nv + j] / gii / G[j * nv + j];
            }
            for (int a = 0; a < na; ++a)
                for (int b = 0; b < na; ++b) {
                    const int j = act[a], kk = act[b];
                    s2 += G[i * nv + j] * G[i * nv + kk] * G[j * nv + kk] / gii / G[j * nv + j] / G[kk * nv + kk];
                }
            fac = sqrt(fabs(1.0 - 2.0 * s1 + s2));
            if (fac <= theta) drop |= 1 << k;
        }
        if (factors && co.lane == 0) factors[k] = fac;
    }
    co.sync();
    return drop;
}

// ---- block-row layout of the DG matrix (dgjumps=True: an element couples to itself and its facet neighbours) ----
// Rows are element-major: row of (element el, local slot k) = row_start[el] + rank of k among the ACTIVE slots of el, where
// slot k < n_p is polynomial and slot n_p + j is enrichment j (active iff bit j of elem_active[el]).  NGSolve numbers the
// constant modes of all elements first: a permutation of this numbering.

// neighbours of el in ascending element number including el itself: nbr[0..n), edge[t] = local edge of el that leads to
// nbr[t] (-1 for el itself).  Returns n <= 4.
EHDG_HD int edg_block_neighbours(const int* el2facet, const int* facet2el, int el, int* nbr, int* edge) {
    int n = 0;
    nbr[n] = el;
    edge[n++] = -1;
    for (int e = 0; e < 3; ++e) {
        const int f = el2facet[3 * el + e];
        const int a = facet2el[2 * f], b = facet2el[2 * f + 1];
        if (b < 0) continue;
        nbr[n] = (a == el) ? b : a;
        edge[n++] = e;
    }
    for (int i = 1; i < n; ++i) {               // insertion sort on the element number
        const int v = nbr[i], w = edge[i];
        int j = i - 1;
        while (j >= 0 && nbr[j] > v) { nbr[j + 1] = nbr[j]; edge[j + 1] = edge[j]; --j; }
        nbr[j + 1] = v;
        edge[j + 1] = w;
    }
    return n;
}

EHDG_HD unsigned edg_dofmask(int n_p, int active_bits) { return ((1u << n_p) - 1u) | ((unsigned)active_bits << n_p); }

// Values of one matrix row: local slot k of element el (must be active) against the active slots of its neighbours in
// ascending element order, read from the block row (D, O) that edg_element_blocks wrote.  out needs
// sum_t popc(mask[nbr[t]]) entries; returns that count.
EHDG_HD int edg_fill_row(int nv, int n_p, const unsigned char* elem_active, int n_nbr, const int* nbr, const int* edge, int k,
                         const double* D, const double* O, double* out) {
    int pos = 0;
    for (int t = 0; t < n_nbr; ++t) {
        const unsigned m = edg_dofmask(n_p, elem_active ? elem_active[nbr[t]] : 0);
        const double* B = edge[t] < 0 ? D : O + (size_t)edge[t] * nv * nv;
        for (int l = 0; l < nv; ++l)
            if ((m >> l) & 1u) out[pos++] = B[k * nv + l];
    }
    return pos;
}

}  // namespace ehdg
