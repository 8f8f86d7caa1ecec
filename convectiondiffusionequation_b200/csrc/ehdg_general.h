// General (quadrature) per-element routines of the (E)HDG hot path: one warp per element,
// working set in shared memory.  Used for every element that touches an enriched dof
// (marked elements and their facet neighbours) and, on request, for all elements.
//
// The routines are written as lane-strided loops over a `Coop` (lane, nl, sync) so that the
// same source also compiles for the host with nl = 1 (tests/hostcheck builds it with g++ to
// check the element math against the numpy oracle without a GPU; nothing in the product
// path calls the host build).
//
// Reference semantics restated here (Python/convection_diffusion.py of the reference):
//   pruning      :270-310     alpha_K      :314-352     forms   :357-375     rhs  :380-381
//   enriched shapes: Python/enrichment_proxy.py:95-96, 133-138 (volume), :161-162 (facet)
// NGSolve conventions: SURVEY.md Appendix A; block formulas: SURVEY.md Appendix B.
#pragma once
#include "ehdg_basis.h"
#include "ehdg_expr.h"

namespace ehdg {

struct Problem {
    int p, n_p, nfl, nenr, nb, nvmax, nfmax, bonus;
    double eps, bx, by;
    double vol_drop_tol;     // selection rule for dependent VOLUME enrichment dofs in the element condensation (0 = off)
    int edg_bnd_data;        // EDG: right-hand side carries the weakly imposed Dirichlet data (debug_edg.py:191-192)
    int enr_axis[EHDG_MAX_ENRICH];
    double enr_k[EHDG_MAX_ENRICH];
    LayerExpr coeff, exact;
    // generic expression programs (device pointers, ehdg_set_expression); null = the layer family above
    const ExprProg *xcoeff, *xexact, *xexact_dx, *xexact_dy;
};

EHDG_HD double problem_coeff(const Problem& pb, double dx, double dy) {
    return pb.xcoeff ? expr_eval(pb.xcoeff, 1.0 - dx, 1.0 - dy) : layer_expr_eval(pb.coeff, dx, dy);
}
EHDG_HD void problem_exact(const Problem& pb, double dx, double dy, double* val, double* gx, double* gy) {
    if (pb.xexact) {
        const double x = 1.0 - dx, y = 1.0 - dy;
        *val = expr_eval(pb.xexact, x, y);
        *gx = pb.xexact_dx ? expr_eval(pb.xexact_dx, x, y) : 0.0;
        *gy = pb.xexact_dy ? expr_eval(pb.xexact_dy, x, y) : 0.0;
    } else {
        layer_expr_grad(pb.exact, dx, dy, val, gx, gy);
    }
}

struct RuleSet {          // pointers valid on the side (host/device) the routine runs on
    int nq, nqf;
    const double *vx, *vy, *vw;   // triangle rule (reference coords lam0, lam1; weights sum 1/2)
    const double *st, *sw;        // segment rule on [0,1]
};

struct Coop {
    int lane, nl;
    EHDG_HD void sync() const {
#ifdef __CUDA_ARCH__
        if (nl > 32) __syncthreads();      // one CTA per element
        else __syncwarp();                 // one warp per element
#endif
    }
    // number of lanes whose predicate holds (a barrier for all nl lanes on the device; the host runs one lane)
    EHDG_HD int count(bool pred) const {
#ifdef __CUDA_ARCH__
        if (nl > 32) return __syncthreads_count(pred ? 1 : 0);
        return __popc(__ballot_sync(0xffffffffu, pred));
#else
        return pred ? 1 : 0;
#endif
    }
};

// status bits per element
enum { EHDG_ST_SINGULAR = 1, EHDG_ST_NONFINITE = 2, EHDG_ST_ALPHA_DEFICIENT = 4, EHDG_ST_ENR_DROPPED = 8 };

struct ElemGeom {
    double P[3][2];
    double det, adet;
    double gl[3][2];     // physical gradients of the local barycentrics
    double omx[3], omy[3];   // 1 - x, 1 - y at the vertices (distance to the outflow sides)
    int f0, f1;          // local vertex carrying the smallest / second smallest global number
    int vn[3];
};

EHDG_HD void elem_geom(const double* verts, const int* tri, ElemGeom& g) {
    for (int v = 0; v < 3; ++v) {
        g.vn[v] = tri[v];
        g.P[v][0] = verts[2 * tri[v]];
        g.P[v][1] = verts[2 * tri[v] + 1];
        g.omx[v] = 1.0 - g.P[v][0];
        g.omy[v] = 1.0 - g.P[v][1];
    }
    const double ax = g.P[0][0] - g.P[2][0], ay = g.P[0][1] - g.P[2][1];
    const double bx = g.P[1][0] - g.P[2][0], by = g.P[1][1] - g.P[2][1];
    g.det = ax * by - bx * ay;
    g.adet = fabs(g.det);
    const double id = 1.0 / g.det;
    g.gl[0][0] = by * id;  g.gl[0][1] = -bx * id;
    g.gl[1][0] = -ay * id; g.gl[1][1] = ax * id;
    g.gl[2][0] = -g.gl[0][0] - g.gl[1][0];
    g.gl[2][1] = -g.gl[0][1] - g.gl[1][1];
    // stable argsort of the three vertex numbers
    int o0 = 0, o1 = 1, o2 = 2, t;
    if (g.vn[o1] < g.vn[o0]) { t = o0; o0 = o1; o1 = t; }
    if (g.vn[o2] < g.vn[o1]) { t = o1; o1 = o2; o2 = t; }
    if (g.vn[o1] < g.vn[o0]) { t = o0; o0 = o1; o1 = t; }
    g.f0 = o0;
    g.f1 = o1;
}

// local edge e -> (a, b): NGSolve ET_TRIG edges e0={V2,V0}, e1={V1,V2}, e2={V0,V1}
EHDG_HD void local_edge(int e, int& a, int& b) {
    a = (e == 0) ? 2 : (e == 1 ? 1 : 0);
    b = (e == 0) ? 0 : (e == 1 ? 2 : 1);
}

struct EdgeGeom {
    int a, b;
    double len, nx, ny, h_e;
    double xsign;   // xi = xsign * (2t - 1)
};

EHDG_HD void edge_geom(const ElemGeom& g, int e, EdgeGeom& eg) {
    local_edge(e, eg.a, eg.b);
    const double tx = g.P[eg.b][0] - g.P[eg.a][0], ty = g.P[eg.b][1] - g.P[eg.a][1];
    eg.len = sqrt(tx * tx + ty * ty);
    double nx = ty / eg.len, ny = -tx / eg.len;
    const int c = 3 - eg.a - eg.b;
    if (nx * (g.P[eg.a][0] - g.P[c][0]) + ny * (g.P[eg.a][1] - g.P[c][1]) < 0.0) { nx = -nx; ny = -ny; }
    eg.nx = nx;
    eg.ny = ny;
    eg.h_e = g.adet / eg.len;                       // specialcf.mesh_size on element_boundary (A.2)
    eg.xsign = (g.vn[eg.a] < g.vn[eg.b]) ? 1.0 : -1.0;
}

// Enriched volume shapes at the point with local barycentrics (l0, l1, l2):
//   out[i*ld] for i < n_p: Dubiner; for i = n_p + k: psi_k if bit k of volmask else 0.
// gx/gy may be null (values only).
EHDG_HD void vol_shapes(const Problem& pb, const ElemGeom& g, int volmask, double l0, double l1,
                        double* phi, double* gx, double* gy, int ld) {
    double v[EHDG_MAX_NP], dX[EHDG_MAX_NP], dY[EHDG_MAX_NP];
    const double l[3] = {l0, l1, 1.0 - l0 - l1};
    dubiner_eval(pb.p, l[g.f0], l[g.f1], v, dX, dY);
    const double ax = g.gl[g.f0][0], ay = g.gl[g.f0][1], bx = g.gl[g.f1][0], by = g.gl[g.f1][1];
    for (int i = 0; i < pb.n_p; ++i) {
        phi[i * ld] = v[i];
        if (gx) {
            gx[i * ld] = dX[i] * ax + dY[i] * bx;
            gy[i * ld] = dX[i] * ay + dY[i] * by;
        }
    }
    for (int k = 0; k < pb.nenr; ++k) {
        double val = 0.0, der = 0.0;
        if ((volmask >> k) & 1) {
            const double* om = pb.enr_axis[k] == 0 ? g.omx : g.omy;
            const double d = l[0] * om[0] + l[1] * om[1] + l[2] * om[2];
            layer_eval(pb.enr_k[k], d, &val, &der);
        }
        const int i = pb.n_p + k;
        phi[i * ld] = val;
        if (gx) {
            gx[i * ld] = pb.enr_axis[k] == 0 ? der : 0.0;
            gy[i * ld] = pb.enr_axis[k] == 0 ? 0.0 : der;
        }
    }
}

// Enriched facet shapes on local edge at parameter t: out[k*ld], k < p+1 Legendre in xi,
// k = p+1+j: psi_j restricted to the edge if bit j of facmask else 0.
EHDG_HD void facet_shapes(const Problem& pb, const ElemGeom& g, const EdgeGeom& eg, int facmask, double t,
                          double* fs, int ld) {
    double mu[EHDG_MAX_ORDER + 1];
    legendre_eval(pb.p, eg.xsign * (2.0 * t - 1.0), mu);
    for (int k = 0; k <= pb.p; ++k) fs[k * ld] = mu[k];
    for (int j = 0; j < pb.nenr; ++j) {
        double val = 0.0, der;
        if ((facmask >> j) & 1) {
            const double* om = pb.enr_axis[j] == 0 ? g.omx : g.omy;
            const double d = (1.0 - t) * om[eg.a] + t * om[eg.b];
            layer_eval(pb.enr_k[j], d, &val, &der);
        }
        fs[(pb.p + 1 + j) * ld] = val;
    }
}

// -------------------------------------------------------------------------------------------
// workspace sizes (doubles)
// -------------------------------------------------------------------------------------------
EHDG_HD int pad_odd(int n) { return n | 1; }

// volume points are processed in chunks of GEN_QCHUNK so that the shape tables stay small (more resident CTAs)
#define EHDG_GEN_QCHUNK 32
EHDG_HD int general_tabs_doubles(const Problem& pb, int nq, int nqf) {
    const int qc = nq < EHDG_GEN_QCHUNK ? nq : EHDG_GEN_QCHUNK;
    int tabs = 3 * pb.nvmax * pad_odd(qc) + 2 * qc;
    const int etabs = (2 * pb.nvmax + pb.nb) * pad_odd(nqf) + nqf;
    if (etabs > tabs) tabs = etabs;
    return (tabs + 1) & ~1;
}

// workspace of alpha_general (whole order-2p rule at once)
EHDG_HD int alpha_ws_doubles(const Problem& pb, int nq0, int nqf0) {
    int tabs = 3 * pb.nvmax * pad_odd(nq0) + 2 * nq0;
    const int etabs = (2 * pb.nvmax + pb.nb) * pad_odd(nqf0) + nqf0;
    if (etabs > tabs) tabs = etabs;
    tabs = (tabs + 1) & ~1;
    return tabs + 3 * pb.nvmax * pb.nvmax + 6 * pb.nvmax + 8;
}

EHDG_HD int general_ws_doubles(const Problem& pb, int nq, int nqf) {
    const int tabs = general_tabs_doubles(pb, nq, nqf);
    const int ldA = pb.nvmax + pb.nfmax + 1;
    // blocks (assemble) or a, m, work (alpha) or Gram (prune): assemble is the largest
    int blocks = pb.nvmax * ldA + pb.nfmax * pb.nvmax + pb.nfmax * pb.nfmax;
    const int alpha_blocks = 3 * pb.nvmax * pb.nvmax + 6 * pb.nvmax;
    if (alpha_blocks > blocks) blocks = alpha_blocks;
    return tabs + blocks + 8;
}

// -------------------------------------------------------------------------------------------
// a5 + a6 + a7: element blocks by quadrature, static condensation
// -------------------------------------------------------------------------------------------
// volmask : bit k set = volume enrichment k is an active dof of this element
// facmask[e]: bit j set = enriched facet dof j exists on local edge e (facet marked; pruned and
//             Dirichlet dofs are dropped later, after S_K is formed: SURVEY App. B)
// Outputs (padded layout): S[nfmax*nfmax], gvec[nfmax], W[nvmax*nfmax], z[nvmax].
// Optional Afull: the uncondensed blocks for parity tests, layout
//             [Avv | Avf | fV] (nvmax x (nvmax+nfmax+1)) then Afv (nfmax x nvmax) then Aff (nfmax x nfmax).
EHDG_HD int assemble_condense_general(const Coop& co, const Problem& pb, const RuleSet& rs,
                                      const double* verts, const int* tri, double lam, int volmask,
                                      const int* facmask, double* ws, double* S, double* gvec,
                                      double* W, double* z, double* Afull) {
    ElemGeom g;
    elem_geom(verts, tri, g);
    const int nv = pb.nvmax, nf = pb.nfmax, nb = pb.nb, nq = rs.nq, nqf = rs.nqf;
    const int qc = nq < EHDG_GEN_QCHUNK ? nq : EHDG_GEN_QCHUNK;
    const int ldq = pad_odd(qc), ldf = pad_odd(nqf), ldA = nv + nf + 1;
    const int tabs = general_tabs_doubles(pb, nq, nqf);
    double* Tphi = ws;
    double* Tgx = Tphi + nv * ldq;
    double* Tgy = Tgx + nv * ldq;
    double* wq = Tgy + nv * ldq;
    double* cq = wq + qc;
    double* Aug = ws + tabs;
    double* Afv = Aug + nv * ldA;
    double* Aff = Afv + nf * nv;
    int status = 0;

    for (int idx = co.lane; idx < nf * nf; idx += co.nl) Aff[idx] = 0.0;
    for (int idx = co.lane; idx < nv * ldA; idx += co.nl) Aug[idx] = 0.0;
    co.sync();
    // ---- volume points, one chunk of the rule at a time
    for (int q0 = 0; q0 < nq; q0 += qc) {
        const int nqc = nq - q0 < qc ? nq - q0 : qc;
        for (int q = co.lane; q < nqc; q += co.nl) {
            const double l0 = rs.vx[q0 + q], l1 = rs.vy[q0 + q], l2 = 1.0 - l0 - l1;
            vol_shapes(pb, g, volmask, l0, l1, Tphi + q, Tgx + q, Tgy + q, ldq);
            wq[q] = rs.vw[q0 + q] * g.adet;
            const double dx = l0 * g.omx[0] + l1 * g.omx[1] + l2 * g.omx[2];
            const double dy = l0 * g.omy[0] + l1 * g.omy[1] + l2 * g.omy[2];
            cq[q] = problem_coeff(pb, dx, dy);
        }
        co.sync();
        for (int idx = co.lane; idx < nv * (nv + 1); idx += co.nl) {
            const int i = idx / (nv + 1), j = idx % (nv + 1);
            const double* pi = Tphi + i * ldq;
            const double* gxi = Tgx + i * ldq;
            const double* gyi = Tgy + i * ldq;
            double acc = 0.0;
            if (j < nv) {
                const double* pj = Tphi + j * ldq;
                const double* gxj = Tgx + j * ldq;
                const double* gyj = Tgy + j * ldq;
                for (int q = 0; q < nqc; ++q)
                    acc += wq[q] * (pb.eps * (gxi[q] * gxj[q] + gyi[q] * gyj[q])
                                    - (pb.bx * gxi[q] + pb.by * gyi[q]) * pj[q]);
                Aug[i * ldA + j] += acc;
            } else {
                for (int q = 0; q < nqc; ++q) acc += wq[q] * cq[q] * pi[q];
                Aug[i * ldA + nv + nf] += acc;
            }
        }
        co.sync();
    }

    // ---- element boundary
    double* Ephi = ws;
    double* Edn = Ephi + nv * ldf;
    double* Efs = Edn + nv * ldf;
    double* we = Efs + nb * ldf;
    const double alpha = 20.0 * lam;                       // conv_diff.py:349
    for (int e = 0; e < 3; ++e) {
        EdgeGeom eg;
        edge_geom(g, e, eg);
        for (int q = co.lane; q < nqf; q += co.nl) {
            const double t = rs.st[q];
            double l[3] = {0.0, 0.0, 0.0};
            l[eg.a] = 1.0 - t;
            l[eg.b] = t;
            we[q] = rs.sw[q] * eg.len;
            double gyl[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            double ph[EHDG_MAX_NP + EHDG_MAX_ENRICH], gxl[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            vol_shapes(pb, g, volmask, l[0], l[1], ph, gxl, gyl, 1);
            for (int i = 0; i < nv; ++i) {
                Ephi[i * ldf + q] = ph[i];
                Edn[i * ldf + q] = gxl[i] * eg.nx + gyl[i] * eg.ny;
            }
            facet_shapes(pb, g, eg, facmask[e], t, Efs + q, ldf);
        }
        co.sync();
        const double bn = pb.bx * eg.nx + pb.by * eg.ny;
        const double pos = bn > 0.0 ? 1.0 : 0.0, neg = 1.0 - pos;     // IfPos(b*n, ., .)
        const double tau = pb.eps * alpha * (double)(pb.p * pb.p) / eg.h_e;
        const double cvv = tau + pos * bn;
        const double cvf = -tau + neg * bn;
        const double cfv = -tau - pos * bn - pos;
        const double cff = tau - neg * bn + pos;
        // A_VV
        for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
            const int i = idx / nv, j = idx % nv;
            const double *pi = Ephi + i * ldf, *pj = Ephi + j * ldf, *di = Edn + i * ldf, *dj = Edn + j * ldf;
            double acc = 0.0;
            for (int q = 0; q < nqf; ++q)
                acc += we[q] * (cvv * pi[q] * pj[q] - pb.eps * (pi[q] * dj[q] + di[q] * pj[q]));
            Aug[i * ldA + j] += acc;
        }
        // A_VF and A_FV
        for (int idx = co.lane; idx < nv * nb; idx += co.nl) {
            const int i = idx / nb, k = idx % nb;
            const double *pi = Ephi + i * ldf, *di = Edn + i * ldf, *fk = Efs + k * ldf;
            double m = 0.0, d = 0.0;
            for (int q = 0; q < nqf; ++q) {
                m += we[q] * pi[q] * fk[q];
                d += we[q] * di[q] * fk[q];
            }
            Aug[i * ldA + nv + e * nb + k] = cvf * m + pb.eps * d;
            Afv[(e * nb + k) * nv + i] = cfv * m + pb.eps * d;
        }
        // A_FF (block diagonal by edge)
        for (int idx = co.lane; idx < nb * nb; idx += co.nl) {
            const int k = idx / nb, l2 = idx % nb;
            const double *fk = Efs + k * ldf, *fl = Efs + l2 * ldf;
            double acc = 0.0;
            for (int q = 0; q < nqf; ++q) acc += we[q] * fk[q] * fl[q];
            Aff[(e * nb + k) * nf + e * nb + l2] = cff * acc;
        }
        co.sync();
    }
    // inactive volume enrichment slots: decoupled unit rows
    for (int k = co.lane; k < pb.nenr; k += co.nl)
        if (!((volmask >> k) & 1)) Aug[(pb.n_p + k) * ldA + pb.n_p + k] = 1.0;
    co.sync();
    if (Afull) {
        for (int idx = co.lane; idx < nv * ldA; idx += co.nl) Afull[idx] = Aug[idx];
        for (int idx = co.lane; idx < nf * nv; idx += co.nl) Afull[nv * ldA + idx] = Afv[idx];
        for (int idx = co.lane; idx < nf * nf; idx += co.nl) Afull[nv * ldA + nf * nv + idx] = Aff[idx];
        co.sync();
    }

    // ---- a6: Gaussian elimination on [A_VV | A_VF f_V].  Polynomial rows: partial pivoting among themselves (the polynomial
    // block is the plain HDG element matrix).  Enrichment rows come last and are eliminated on their diagonal, which after the
    // polynomial (and earlier enrichment) dofs have been eliminated is the share of psi_k that those dofs do not explain,
    // relative to its own size d0: below pb.vol_drop_tol the enrichment is dependent on this element (q(y) on an element of
    // height h with beta_y h / eps < 1 is a polynomial to 1e-6) and leaves the element space -- its coefficient is 0, its row
    // and column do not enter S_K.  The reference keeps such dofs (its pruning factor :296-301 is a boundary-Gram distance at
    // 1e-5) and hands PARDISO a singular block; status bit EHDG_ST_ENR_DROPPED records where the rule fired.
    const int ncol = ldA;
    double d0[EHDG_MAX_ENRICH];
    for (int k = 0; k < pb.nenr; ++k) d0[k] = fabs(Aug[(pb.n_p + k) * ldA + pb.n_p + k]);
    co.sync();
    for (int k = 0; k < nv; ++k) {
        const bool enr = k >= pb.n_p;
        int piv = k;
        double best = fabs(Aug[k * ldA + k]);
        if (!enr)
            for (int i = k + 1; i < pb.n_p; ++i) {
                const double a = fabs(Aug[i * ldA + k]);
                if (a > best) { best = a; piv = i; }
            }
        if (!(best > 0.0) || best != best) status |= EHDG_ST_SINGULAR;
        const bool drop = enr && ((volmask >> (k - pb.n_p)) & 1) && pb.vol_drop_tol > 0.0 && !(best > pb.vol_drop_tol * d0[k - pb.n_p]);
        co.sync();                      // every lane has finished scanning column k before rows move
        if (drop) {
            status |= EHDG_ST_ENR_DROPPED;
            for (int j = co.lane; j < ncol; j += co.nl) Aug[k * ldA + j] = (j == k) ? 1.0 : 0.0;     // x_k = 0
            for (int i = co.lane; i < nv; i += co.nl)
                if (i != k) Aug[i * ldA + k] = 0.0;                                                   // and nobody sees it
            co.sync();
            continue;
        }
        if (piv != k) {
            for (int j = k + co.lane; j < ncol; j += co.nl) {
                const double t = Aug[k * ldA + j];
                Aug[k * ldA + j] = Aug[piv * ldA + j];
                Aug[piv * ldA + j] = t;
            }
            co.sync();
        }
        const double ip = 1.0 / Aug[k * ldA + k];
        const int nr = nv - k - 1, nc = ncol - k - 1;
        for (int idx = co.lane; idx < nr * nc; idx += co.nl) {
            const int i = k + 1 + idx / nc, j = k + 1 + idx % nc;
            Aug[i * ldA + j] -= (Aug[i * ldA + k] * ip) * Aug[k * ldA + j];
        }
        co.sync();
    }
    // back substitution, one right-hand side per lane; solutions overwrite the RHS columns
    for (int c = co.lane; c < nf + 1; c += co.nl) {
        const int col = nv + c;
        for (int i = nv - 1; i >= 0; --i) {
            double s = Aug[i * ldA + col];
            for (int j = i + 1; j < nv; ++j) s -= Aug[i * ldA + j] * Aug[j * ldA + col];
            Aug[i * ldA + col] = s / Aug[i * ldA + i];
        }
    }
    co.sync();
    // ---- S = A_FF - A_FV W,  g = -A_FV z
    for (int idx = co.lane; idx < nf * (nf + 1); idx += co.nl) {
        const int k = idx / (nf + 1), l2 = idx % (nf + 1);
        double acc = (l2 < nf) ? Aff[k * nf + l2] : 0.0;
        for (int j = 0; j < nv; ++j) acc -= Afv[k * nv + j] * Aug[j * ldA + nv + l2];
        if (acc != acc || fabs(acc) > 1.7e308) status |= EHDG_ST_NONFINITE;
        if (l2 < nf) S[k * nf + l2] = acc;
        else gvec[k] = acc;
    }
    for (int idx = co.lane; idx < nv * (nf + 1); idx += co.nl) {
        const int i = idx / (nf + 1), c = idx % (nf + 1);
        const double val = Aug[i * ldA + nv + c];
        if (c < nf) W[i * nf + c] = val;
        else z[i] = val;
    }
    co.sync();
    return status;
}

// -------------------------------------------------------------------------------------------
// symmetric helper: largest eigenvalue of the pencil (M, A), A SPD, both n x n row-major (ld n)
// in workspace memory; destroys A and M.  Cholesky -> C = L^-1 M L^-T -> Householder
// tridiagonalisation -> bisection on the Sturm count.  `wk` needs 4 n doubles.
// -------------------------------------------------------------------------------------------
EHDG_HD double pencil_lambda_max(const Coop& co, int n, double* A, double* M, double* wk, int* deficient) {
    if (n <= 0) return 0.0;
    // Cholesky, A = L L^T (lower triangle of A overwritten by L)
    for (int k = 0; k < n; ++k) {
        const double dkk_orig = A[k * n + k];
        double d = dkk_orig;
        for (int j = 0; j < k; ++j) d -= A[k * n + j] * A[k * n + j];
        co.sync();
        double l;
        // a pivot this small means the dof is numerically in the span of the earlier ones; the
        // reference's pinv (rcond 1e-15) projects such a direction out (conv_diff.py:345)
        const bool drop = !(d > 1e-13 * fabs(dkk_orig));
        if (drop) { *deficient = 1; l = 0.0; } else l = sqrt(d);
        for (int i = k + 1 + co.lane; i < n; i += co.nl) {
            double s = A[i * n + k];
            for (int j = 0; j < k; ++j) s -= A[i * n + j] * A[k * n + j];
            A[i * n + k] = drop ? 0.0 : s / l;
        }
        co.sync();
        if (co.lane == 0) A[k * n + k] = drop ? 0.0 : l;
        co.sync();
    }
    // M <- L^-1 M (forward substitution on every column), directions dropped by the Cholesky
    // (pseudo-inverse semantics of np.linalg.pinv, conv_diff.py:345) give zero rows
    for (int c = co.lane; c < n; c += co.nl) {
        for (int i = 0; i < n; ++i) {
            double s = M[i * n + c];
            for (int j = 0; j < i; ++j) s -= A[i * n + j] * M[j * n + c];
            const double l = A[i * n + i];
            M[i * n + c] = (l > 0.0) ? s / l : 0.0;
        }
    }
    co.sync();
    // M <- M L^-T  (same substitution on every row)
    for (int r = co.lane; r < n; r += co.nl) {
        for (int i = 0; i < n; ++i) {
            double s = M[r * n + i];
            for (int j = 0; j < i; ++j) s -= A[i * n + j] * M[r * n + j];
            const double l = A[i * n + i];
            M[r * n + i] = (l > 0.0) ? s / l : 0.0;
        }
    }
    co.sync();
    // symmetrise (round-off only)
    for (int idx = co.lane; idx < n * n; idx += co.nl) {
        const int i = idx / n, j = idx % n;
        if (i < j) {
            const double s = 0.5 * (M[i * n + j] + M[j * n + i]);
            A[i * n + j] = s;   // A is free now: use it as scratch for the symmetrised copy
            A[j * n + i] = s;
        } else if (i == j) A[idx] = M[idx];
    }
    co.sync();
    double* C = A;
    double* dg = wk;            // diagonal of T
    double* of = wk + n;        // off-diagonal of T (of[k] couples k, k+1)
    double* v = wk + 2 * n;
    double* pv = wk + 3 * n;
    // Householder tridiagonalisation (lower form): at step k annihilate C[k+2.., k]
    for (int k = 0; k + 2 < n; ++k) {
        double sig = 0.0;
        for (int i = k + 1; i < n; ++i) sig += C[i * n + k] * C[i * n + k];
        const double x0 = C[(k + 1) * n + k];
        const double nrm = sqrt(sig);
        co.sync();
        if (nrm == 0.0 || sig - x0 * x0 <= 0.0) {
            if (co.lane == 0) of[k] = x0;
            co.sync();
            continue;
        }
        const double alpha = (x0 > 0.0) ? -nrm : nrm;
        const double v0 = x0 - alpha;
        const double vnorm2 = sig - x0 * x0 + v0 * v0;
        const double beta = 2.0 / vnorm2;
        for (int i = k + 1 + co.lane; i < n; i += co.nl) v[i] = (i == k + 1) ? v0 : C[i * n + k];
        co.sync();
        // p = beta * C v  (trailing block)
        for (int i = k + 1 + co.lane; i < n; i += co.nl) {
            double s = 0.0;
            for (int j = k + 1; j < n; ++j) s += C[i * n + j] * v[j];
            pv[i] = beta * s;
        }
        co.sync();
        double kk = 0.0;
        for (int i = k + 1; i < n; ++i) kk += v[i] * pv[i];
        kk *= 0.5 * beta;
        co.sync();
        for (int i = k + 1 + co.lane; i < n; i += co.nl) pv[i] -= kk * v[i];
        co.sync();
        const int nt = n - k - 1;
        for (int idx = co.lane; idx < nt * nt; idx += co.nl) {
            const int i = k + 1 + idx / nt, j = k + 1 + idx % nt;
            C[i * n + j] -= v[i] * pv[j] + pv[i] * v[j];
        }
        if (co.lane == 0) of[k] = alpha;
        co.sync();
    }
    if (n >= 2 && co.lane == 0) of[n - 2] = C[(n - 1) * n + (n - 2)];
    for (int i = co.lane; i < n; i += co.nl) dg[i] = C[i * n + i];
    co.sync();
    // Gershgorin upper bound and bisection for the largest eigenvalue
    double lo = dg[0], hi = dg[0];
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? fabs(of[i - 1]) : 0.0) + (i + 1 < n ? fabs(of[i]) : 0.0);
        if (dg[i] + r > hi) hi = dg[i] + r;
        if (dg[i] - r < lo) lo = dg[i] - r;
    }
    const double scale = fabs(hi) > fabs(lo) ? fabs(hi) : fabs(lo);
    const double tiny = 1e-300 + 1e-32 * scale * scale;
    hi += 1e-14 * scale;
    // (nl + 1)-section on the Sturm count: lane l tests the abscissa lo + (l + 1) step; with one lane this is
    // plain bisection.
    const int nsec = co.nl > 1 ? co.nl : 1;
    for (int it = 0; it < 200; ++it) {
        const double step = (hi - lo) / (double)(nsec + 1);
        const double x = lo + step * (double)((nsec > 1 ? co.lane : 0) + 1);
        if (!(lo + step > lo) || !(lo + step * (double)nsec < hi) || !(step > 0.0)) break;   // lane-independent: every lane leaves together
        // all eigenvalues below x  <=>  every pivot of the LDL^T of T - x I is negative
        int nneg = 0;
        double q = dg[0] - x;
        if (q < 0.0) ++nneg;
        for (int i = 1; i < n; ++i) {
            if (fabs(q) < tiny) q = (q < 0.0) ? -tiny : tiny;
            q = (dg[i] - x) - of[i - 1] * of[i - 1] / q;
            if (q < 0.0) ++nneg;
        }
        // first abscissa above the whole spectrum: the test is monotone in x, so it is the number of lanes still below
        // (one barrier-reduction instead of a flag array that every lane scans)
        int f = nsec;
        if (nsec > 1) {
            f = co.count(nneg != n);
        } else if (nneg == n) {
            f = 0;
        }
        const double nlo = lo + step * (double)f;
        const double nhi = (f < nsec) ? lo + step * (double)(f + 1) : hi;
        lo = nlo;
        hi = nhi;
    }
    co.sync();
    return 0.5 * (lo + hi);
}

// -------------------------------------------------------------------------------------------
// a4: lambda_K = max eig(pinv(a + f f^T) m) on the active volume-type dofs (conv_diff.py:314-351)
//   a = int grad u . grad v (order 2p), f = int h^-2 v (volume h = sqrt|detJ|),
//   m = oint h_e (d_n u)(d_n v) (order 2p, h_e = |detJ| / |e|).
// exclude_const: the reference keeps dofs with `dof > 0`, which drops the constant mode of
// element 0 only (conv_diff.py:338).
// -------------------------------------------------------------------------------------------
EHDG_HD double alpha_general(const Coop& co, const Problem& pb, const RuleSet& rs0, const double* verts,
                             const int* tri, int volmask, int exclude_const, double* ws, int* status) {
    ElemGeom g;
    elem_geom(verts, tri, g);
    const int nv = pb.nvmax, nq = rs0.nq, nqf = rs0.nqf;
    // the order-2p rule of the alpha integrators is small (nq0 <= 49): processed in one piece
    const int ldq = pad_odd(nq), ldf = pad_odd(nqf);
    int tabs = 3 * nv * ldq + 2 * nq;
    {
        const int etabs = (2 * nv + pb.nb) * ldf + nqf;
        if (etabs > tabs) tabs = etabs;
    }
    tabs = (tabs + 1) & ~1;
    double* Tphi = ws;
    double* Tgx = Tphi + nv * ldq;
    double* Tgy = Tgx + nv * ldq;
    double* wq = Tgy + nv * ldq;
    double* Am = ws + tabs;            // nv x nv
    double* Mm = Am + nv * nv;         // nv x nv
    double* Cm = Mm + nv * nv;         // compacted copies / scratch
    double* fv = Cm + nv * nv;         // nv
    double* wk = fv + nv;              // 4 nv  (+ nv spare)
    for (int q = co.lane; q < nq; q += co.nl) {
        vol_shapes(pb, g, volmask, rs0.vx[q], rs0.vy[q], Tphi + q, Tgx + q, Tgy + q, ldq);
        wq[q] = rs0.vw[q] * g.adet;
    }
    co.sync();
    for (int i = co.lane; i < nv; i += co.nl) {
        double s = 0.0;
        for (int q = 0; q < nq; ++q) s += wq[q] * Tphi[i * ldq + q];
        fv[i] = s / g.adet;                                   // h^-2 with h^2 = |detJ|
    }
    co.sync();
    for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
        const int i = idx / nv, j = idx % nv;
        double acc = 0.0;
        for (int q = 0; q < nq; ++q)
            acc += wq[q] * (Tgx[i * ldq + q] * Tgx[j * ldq + q] + Tgy[i * ldq + q] * Tgy[j * ldq + q]);
        Am[idx] = acc + fv[i] * fv[j];                        // conv_diff.py:329-331
        Mm[idx] = 0.0;
    }
    co.sync();
    double* Edn = ws;
    double* we = Edn + nv * ldf;
    for (int e = 0; e < 3; ++e) {
        EdgeGeom eg;
        edge_geom(g, e, eg);
        for (int q = co.lane; q < nqf; q += co.nl) {
            const double t = rs0.st[q];
            double l[3] = {0.0, 0.0, 0.0};
            l[eg.a] = 1.0 - t;
            l[eg.b] = t;
            double ph[EHDG_MAX_NP + EHDG_MAX_ENRICH], gxl[EHDG_MAX_NP + EHDG_MAX_ENRICH],
                gyl[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            vol_shapes(pb, g, volmask, l[0], l[1], ph, gxl, gyl, 1);
            for (int i = 0; i < nv; ++i) Edn[i * ldf + q] = gxl[i] * eg.nx + gyl[i] * eg.ny;
            we[q] = rs0.sw[q] * eg.len * eg.h_e;
        }
        co.sync();
        for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
            const int i = idx / nv, j = idx % nv;
            double acc = 0.0;
            for (int q = 0; q < nqf; ++q) acc += we[q] * Edn[i * ldf + q] * Edn[j * ldf + q];
            Mm[idx] += acc;
        }
        co.sync();
    }
    // restrict to the important dofs
    int keep[EHDG_MAX_NP + EHDG_MAX_ENRICH];
    int n = 0;
    for (int i = 0; i < nv; ++i) {
        bool ok = true;
        if (i == 0 && exclude_const) ok = false;
        if (i >= pb.n_p && !((volmask >> (i - pb.n_p)) & 1)) ok = false;
        if (ok) keep[n++] = i;
    }
    for (int idx = co.lane; idx < n * n; idx += co.nl) {
        const int i = idx / n, j = idx % n;
        Cm[idx] = Am[keep[i] * nv + keep[j]];
    }
    co.sync();
    for (int idx = co.lane; idx < n * n; idx += co.nl) Am[idx] = Cm[idx];
    co.sync();
    for (int idx = co.lane; idx < n * n; idx += co.nl) {
        const int i = idx / n, j = idx % n;
        Cm[idx] = Mm[keep[i] * nv + keep[j]];
    }
    co.sync();
    int deficient = 0;
    const double lam = pencil_lambda_max(co, n, Am, Cm, wk, &deficient);
    if (deficient) *status |= EHDG_ST_ALPHA_DEFICIENT;
    if (lam != lam) *status |= EHDG_ST_NONFINITE;
    return lam;
}

// -------------------------------------------------------------------------------------------
// a3: enrichment pruning on one marked element (conv_diff.py:270-310).
//   Gram = oint (u v + uhat vhat) with the bonus rule.  The volume-type dofs (V, Q_k) and the
//   facet-type dofs of each edge (F_e, QF_k,e) never couple, so the reference's factor
//        sqrt|1 - 2 sum_j G_ij^2/(G_ii G_jj) + sum_jk G_ij G_ik G_jk/(G_ii G_jj G_kk)|
//   over the earlier important dofs reduces to sums inside one group, visited in NGSolve's
//   local order  V, F, Q_0, QF_0[e0..e2], Q_1, QF_1[e0..e2].
//   elmark  : bit k = element marked for enrichment k (Q_k dof exists)
//   facmask : bit j = facet marked for enrichment j (QF_j dof exists), per local edge
// Outputs: vol_drop (bits of Q_k pruned), fac_drop[e] (bits of QF_j,e pruned), factors[4*nenr]
//   in visiting order (Q_k, QF_k,e0..e2), -1 where the dof does not exist.
// -------------------------------------------------------------------------------------------
EHDG_HD void prune_general(const Coop& co, const Problem& pb, const RuleSet& rs, const double* verts,
                           const int* tri, int elmark, const int* facmask, double threshold, double* ws,
                           int* vol_drop, int* fac_drop, double* factors) {
    ElemGeom g;
    elem_geom(verts, tri, g);
    const int nv = pb.nvmax, nb = pb.nb, nqf = rs.nqf, ldf = pad_odd(nqf);
    const int tabs = general_tabs_doubles(pb, rs.nq, nqf);
    double* Ephi = ws;
    double* Efs = Ephi + nv * ldf;
    double* we = Efs + nb * ldf;
    double* Gv = ws + tabs;               // nv x nv
    double* Gf = Gv + nv * nv;            // 3 x nb x nb
    for (int idx = co.lane; idx < nv * nv; idx += co.nl) Gv[idx] = 0.0;
    co.sync();
    for (int e = 0; e < 3; ++e) {
        EdgeGeom eg;
        edge_geom(g, e, eg);
        for (int q = co.lane; q < nqf; q += co.nl) {
            const double t = rs.st[q];
            double l[3] = {0.0, 0.0, 0.0};
            l[eg.a] = 1.0 - t;
            l[eg.b] = t;
            vol_shapes(pb, g, elmark, l[0], l[1], Ephi + q, (double*)0, (double*)0, ldf);
            facet_shapes(pb, g, eg, facmask[e], t, Efs + q, ldf);
            we[q] = rs.sw[q] * eg.len;
        }
        co.sync();
        for (int idx = co.lane; idx < nv * nv; idx += co.nl) {
            const int i = idx / nv, j = idx % nv;
            double acc = 0.0;
            for (int q = 0; q < nqf; ++q) acc += we[q] * Ephi[i * ldf + q] * Ephi[j * ldf + q];
            Gv[idx] += acc;
        }
        for (int idx = co.lane; idx < nb * nb; idx += co.nl) {
            const int k = idx / nb, l2 = idx % nb;
            double acc = 0.0;
            for (int q = 0; q < nqf; ++q) acc += we[q] * Efs[k * ldf + q] * Efs[l2 * ldf + q];
            Gf[e * nb * nb + idx] = acc;
        }
        co.sync();
    }
    // sequential factor tests (every lane computes the same result).  A zero Gram diagonal means the
    // enrichment function vanishes identically on that facet (psi on the outflow edge itself, where
    // d = 0 exactly in the cancellation-free form).  NGSolve evaluates psi from a rounded coordinate,
    // gets ~1e-16 rather than 0 and therefore neither divides by zero nor prunes this (Dirichlet)
    // dof; here it is skipped (factor -3) and kept out of the later active sets.
    int vdrop = 0;
    int fdrop[3] = {0, 0, 0};
    int vzero = 0;
    int fzero[3] = {0, 0, 0};
    for (int k = 0; k < pb.nenr; ++k) {
        // --- Q_k against V and the earlier surviving Q's
        double fac = -1.0;
        if ((elmark >> k) & 1) {
            const int i = pb.n_p + k;
            const double gii = Gv[i * nv + i];
            if (gii == 0.0) { vzero |= 1 << k; fac = -3.0; }
            else {
                int act[EHDG_MAX_NP + EHDG_MAX_ENRICH];
                int na = 0;
                for (int j = 0; j < pb.n_p; ++j) act[na++] = j;
                for (int kk = 0; kk < k; ++kk)
                    if (((elmark >> kk) & 1) && !(((vdrop | vzero) >> kk) & 1)) act[na++] = pb.n_p + kk;
                double s1 = 0.0, s2 = 0.0;
                for (int a = 0; a < na; ++a) {
                    const int j = act[a];
                    s1 += Gv[i * nv + j] * Gv[i * nv + j] / gii / Gv[j * nv + j];
                }
                for (int a = 0; a < na; ++a)
                    for (int b = 0; b < na; ++b) {
                        const int j = act[a], kk = act[b];
                        s2 += Gv[i * nv + j] * Gv[i * nv + kk] * Gv[j * nv + kk] / gii / Gv[j * nv + j] / Gv[kk * nv + kk];
                    }
                fac = sqrt(fabs(1.0 - 2.0 * s1 + s2));
                if (fac <= threshold) vdrop |= 1 << k;
            }
        }
        factors[4 * k] = fac;
        // --- QF_k on each edge against F_e and the earlier surviving QF's of that edge
        for (int e = 0; e < 3; ++e) {
            double ff = -1.0;
            if ((facmask[e] >> k) & 1) {
                const double* G = Gf + e * nb * nb;
                const int i = pb.p + 1 + k;
                const double gii = G[i * nb + i];
                if (gii == 0.0) { fzero[e] |= 1 << k; ff = -3.0; }
                else {
                    int act[EHDG_MAX_ORDER + 1 + EHDG_MAX_ENRICH];
                    int na = 0;
                    for (int j = 0; j <= pb.p; ++j) act[na++] = j;
                    for (int kk = 0; kk < k; ++kk)
                        if (((facmask[e] >> kk) & 1) && !(((fdrop[e] | fzero[e]) >> kk) & 1)) act[na++] = pb.p + 1 + kk;
                    double s1 = 0.0, s2 = 0.0;
                    for (int a = 0; a < na; ++a) {
                        const int j = act[a];
                        s1 += G[i * nb + j] * G[i * nb + j] / gii / G[j * nb + j];
                    }
                    for (int a = 0; a < na; ++a)
                        for (int b = 0; b < na; ++b) {
                            const int j = act[a], kk = act[b];
                            s2 += G[i * nb + j] * G[i * nb + kk] * G[j * nb + kk] / gii / G[j * nb + j] / G[kk * nb + kk];
                        }
                    ff = sqrt(fabs(1.0 - 2.0 * s1 + s2));
                    if (ff <= threshold) fdrop[e] |= 1 << k;
                }
            }
            factors[4 * k + 1 + e] = ff;
        }
    }
    *vol_drop = vdrop;
    for (int e = 0; e < 3; ++e) fac_drop[e] = fdrop[e];
    co.sync();
}

}  // namespace ehdg
