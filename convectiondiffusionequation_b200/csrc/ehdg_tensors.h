// Host-side builder of the reference tensors of the polynomial path.
//
// On an affine element with constant beta/eps every polynomial-polynomial entry of the
// (E)HDG blocks (convection_diffusion.py:357-375) is a combination of
// element-independent integrals over the vertex-sorted reference triangle (X = lam_Q0,
// Y = lam_Q1, Q0<Q1<Q2 by global vertex number) with a handful of geometric coefficients.
// The integrals are evaluated here once per (order, bonus) with the same Duffy/Gauss rules the
// quadrature path uses (exact for these polynomial integrands).
//
// sorted-frame edges: s = edge opposite Q_s:  s0 = (Q1,Q2), s1 = (Q0,Q2), s2 = (Q0,Q1),
// parameter t in [0,1] from the lower to the higher vertex, facet basis Legendre(2t-1).
#pragma once
#include <vector>

#include "ehdg_basis.h"
#include "ehdg_rules.h"

namespace ehdg {

struct TensorLayout {
    int NP, NFL, nq;
    int NFS;    // row stride of the facet tensors: 3 NFL made odd (conflict-free for a lane-per-row read)
    // offsets in doubles into `asm_tab`
    int offV;   // [5][NP][NP]    KXX, KS = KXY + KXY^T, KYY, CX, CY
    int offE;   // [3][3][NP][NP] per edge: M, DXS = DX + DX^T, DYS = DY + DY^T
    int offFT;  // [3][NP][NFS]    per edge s: MF, DXF, DYF stored as [t][i][(s,k)]: the lane that owns volume row i reads
                //                 with the odd stride 3 NFL, the lane that owns facet row (s,k) with unit stride
    int asm_doubles;
    // offsets into `alpha_tab`
    int offAK;  // [3][NP][NP]    KXX, KS, KYY
    int offAN;  // [3][3][NP][NP] per edge: NXX, NXYS, NYY   (products of derivatives on the edge)
    int alpha_doubles;
};

EHDG_HD TensorLayout tensor_layout(int p, int nq) {
    TensorLayout L;
    L.NP = (p + 1) * (p + 2) / 2;
    L.NFL = p + 1;
    L.nq = nq;
    L.NFS = (3 * L.NFL) | 1;
    L.offV = 0;
    L.offE = L.offV + 5 * L.NP * L.NP;
    L.offFT = L.offE + 9 * L.NP * L.NP;
    L.asm_doubles = L.offFT + 3 * L.NP * L.NFS;
    L.offAK = 0;
    L.offAN = 3 * L.NP * L.NP;
    L.alpha_doubles = 12 * L.NP * L.NP;
    return L;
}

// sorted-frame edge point: (X, Y) at parameter t
inline void sorted_edge_point(int s, double t, double* X, double* Y) {
    if (s == 0) { *X = 0.0; *Y = 1.0 - t; }        // Q1 -> Q2
    else if (s == 1) { *X = 1.0 - t; *Y = 0.0; }   // Q0 -> Q2
    else { *X = 1.0 - t; *Y = t; }                 // Q0 -> Q1
}

inline void build_tensors(int p, int order, std::vector<double>& asm_tab, std::vector<double>& alpha_tab) {
    const TrigRule tr = trig_rule(order);
    const SegRule sr = seg_rule(order);
    const TensorLayout L = tensor_layout(p, (int)tr.w.size());
    const int NP = L.NP, NFL = L.NFL;
    std::vector<long double> tv(5 * NP * NP, 0.0L), te(9 * NP * NP, 0.0L), tf(9 * NP * NFL, 0.0L), tn(9 * NP * NP, 0.0L);
    double v[EHDG_MAX_NP], dX[EHDG_MAX_NP], dY[EHDG_MAX_NP], mu[EHDG_MAX_ORDER + 1];
    for (size_t q = 0; q < tr.w.size(); ++q) {
        dubiner_eval(p, tr.x[q], tr.y[q], v, dX, dY);
        const long double w = tr.w[q];
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) {
                tv[(0 * NP + i) * NP + j] += w * dX[i] * dX[j];
                tv[(1 * NP + i) * NP + j] += w * (dX[i] * dY[j] + dY[i] * dX[j]);
                tv[(2 * NP + i) * NP + j] += w * dY[i] * dY[j];
                tv[(3 * NP + i) * NP + j] += w * dX[i] * v[j];
                tv[(4 * NP + i) * NP + j] += w * dY[i] * v[j];
            }
    }
    for (int s = 0; s < 3; ++s)
        for (size_t q = 0; q < sr.w.size(); ++q) {
            double X, Y;
            sorted_edge_point(s, sr.t[q], &X, &Y);
            dubiner_eval(p, X, Y, v, dX, dY);
            legendre_eval(p, 2.0 * sr.t[q] - 1.0, mu);
            const long double w = sr.w[q];
            for (int i = 0; i < NP; ++i) {
                for (int j = 0; j < NP; ++j) {
                    te[((s * 3 + 0) * NP + i) * NP + j] += w * v[i] * v[j];
                    te[((s * 3 + 1) * NP + i) * NP + j] += w * (v[i] * dX[j] + dX[i] * v[j]);
                    te[((s * 3 + 2) * NP + i) * NP + j] += w * (v[i] * dY[j] + dY[i] * v[j]);
                    tn[((s * 3 + 0) * NP + i) * NP + j] += w * dX[i] * dX[j];
                    tn[((s * 3 + 1) * NP + i) * NP + j] += w * (dX[i] * dY[j] + dY[i] * dX[j]);
                    tn[((s * 3 + 2) * NP + i) * NP + j] += w * dY[i] * dY[j];
                }
                for (int k = 0; k < NFL; ++k) {
                    tf[((s * 3 + 0) * NP + i) * NFL + k] += w * v[i] * mu[k];
                    tf[((s * 3 + 1) * NP + i) * NFL + k] += w * dX[i] * mu[k];
                    tf[((s * 3 + 2) * NP + i) * NFL + k] += w * dY[i] * mu[k];
                }
            }
        }
    asm_tab.assign(L.asm_doubles, 0.0);
    alpha_tab.assign(L.alpha_doubles, 0.0);
    for (int i = 0; i < 5 * NP * NP; ++i) asm_tab[L.offV + i] = (double)tv[i];
    for (int i = 0; i < 9 * NP * NP; ++i) asm_tab[L.offE + i] = (double)te[i];
    for (int s = 0; s < 3; ++s)
        for (int t = 0; t < 3; ++t)
            for (int j = 0; j < NP; ++j)
                for (int k = 0; k < NFL; ++k)
                    asm_tab[L.offFT + (t * NP + j) * L.NFS + s * NFL + k] = (double)tf[((s * 3 + t) * NP + j) * NFL + k];
    for (int i = 0; i < 3 * NP * NP; ++i) alpha_tab[L.offAK + i] = (double)tv[i];
    for (int i = 0; i < 9 * NP * NP; ++i) alpha_tab[L.offAN + i] = (double)tn[i];
}

// phi at the points of the triangle rule for the 9 (f0, f1) vertex-order codes (6 are valid
// permutations); layout [f0*3+f1][q][NP].  The rule lives in LOCAL reference coordinates
// (lam_0, lam_1), so the sorted barycentrics of a point depend on the class of the element.
//
// Behind it, [f0*3+f1][6][NP]: the moments  sum_q w_q m(q) phi_i(q)  of the monomials m = 1, l0, l1, l0^2, l0 l1, l1^2
// in the LOCAL barycentrics.  Where the layer functions of `coeff` have underflowed on the whole element the
// right-hand side coefficient is a quadratic polynomial and f_V is a combination of these six vectors (the rule
// integrates degree p + 2 exactly, so this is the same number up to rounding).
constexpr int EHDG_RHS_MOMENTS = 6;
inline void build_rhs_phi(int p, int order, std::vector<double>& tab) {
    const TrigRule tr = trig_rule(order);
    const int NP = (p + 1) * (p + 2) / 2, nq = (int)tr.w.size();
    tab.assign((size_t)9 * nq * NP + (size_t)9 * EHDG_RHS_MOMENTS * NP, 0.0);
    std::vector<long double> mom((size_t)9 * EHDG_RHS_MOMENTS * NP, 0.0L);
    double v[EHDG_MAX_NP], dX[EHDG_MAX_NP], dY[EHDG_MAX_NP];
    for (int f0 = 0; f0 < 3; ++f0)
        for (int f1 = 0; f1 < 3; ++f1) {
            if (f0 == f1) continue;
            for (int q = 0; q < nq; ++q) {
                const double l[3] = {tr.x[q], tr.y[q], 1.0 - tr.x[q] - tr.y[q]};
                dubiner_eval(p, l[f0], l[f1], v, dX, dY);
                for (int i = 0; i < NP; ++i) tab[((size_t)(f0 * 3 + f1) * nq + q) * NP + i] = v[i];
                const long double m[EHDG_RHS_MOMENTS] = {1.0L, l[0], l[1], (long double)l[0] * l[0], (long double)l[0] * l[1],
                                                         (long double)l[1] * l[1]};
                for (int k = 0; k < EHDG_RHS_MOMENTS; ++k)
                    for (int i = 0; i < NP; ++i)
                        mom[((size_t)(f0 * 3 + f1) * EHDG_RHS_MOMENTS + k) * NP + i] += (long double)tr.w[q] * m[k] * v[i];
            }
        }
    for (size_t i = 0; i < mom.size(); ++i) tab[(size_t)9 * nq * NP + i] = (double)mom[i];
}

}  // namespace ehdg
