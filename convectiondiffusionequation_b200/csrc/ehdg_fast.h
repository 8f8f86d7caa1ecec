// Polynomial-path element routines (no enriched dof on the element): alpha_K (a4) and
// assembly + static condensation (a5-a7) from the reference tensors of ehdg_tensors.h.
//
// Execution model: a GROUP of LG lanes (LG = 8, 16 or 32, a power of two >= max(n_p, 3(p+1)))
// owns one element; lane i keeps ROW i of the element matrices in registers.  Values cross
// lanes only through the group's shared-memory scratch and three primitives (argmax / sum /
// ballot).  The body is written in lock-step phases
//     EHDG_FOR_LANES(l, L)  ...  EHDG_END_LANES
// so that the same source runs on the device (each thread is one lane, L == 0, a phase ends
// with __syncwarp) and on the host (a phase is a loop over the LG lanes, L == l); the host
// instantiation exists only for tests/hostcheck.
#pragma once
#include <stdint.h>
#include <string.h>
#include <cmath>

#include "ehdg_general.h"
#include "ehdg_tensors.h"

namespace ehdg {

#ifdef __CUDA_ARCH__
#define EHDG_NLS(LG) 1
#define EHDG_FOR_LANES(l, L) { const int l = gc.lane; constexpr int L = 0; (void)L;
#define EHDG_END_LANES } __syncwarp();
#else
#define EHDG_NLS(LG) (LG)
#define EHDG_FOR_LANES(l, L) for (int l = 0; l < LG; ++l) { const int L = l;
#define EHDG_END_LANES }
#endif

// strict sign change between two doubles (both non-zero), on the integer pipe
EHDG_HD bool opposite_signs(double a, double b) {
#ifdef __CUDA_ARCH__
    // exact zeros are a set of measure zero in the bracket refinement; the sign bit alone decides
    return (__double2hiint(a) ^ __double2hiint(b)) < 0;
#else
    return std::signbit(a) != std::signbit(b);
#endif
}

// v[s] for s in {0,1,2} without dynamic indexing (keeps small per-edge arrays in registers)
EHDG_HD double sel3(const double* v, int s) { return s == 0 ? v[0] : (s == 1 ? v[1] : v[2]); }
EHDG_HD int sel3i(const int* v, int s) { return s == 0 ? v[0] : (s == 1 ? v[1] : v[2]); }

// 16-byte shared-memory accesses (p must be 16-byte aligned): one wavefront moves two doubles
EHDG_HD void ld2(const double* p, double& a, double& b) {
#ifdef __CUDA_ARCH__
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
}
EHDG_HD void st2(double* p, double a, double b) {
#ifdef __CUDA_ARCH__
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
#else
    p[0] = a;
    p[1] = b;
#endif
}

// Fused a5-a8 store (ehdg_assemble_condense_csr): where the condensed block of one element goes in the facet-major value
// array.  Off-diagonal facet pairs (a != b) belong to exactly one element and are written in place; the diagonal block of a
// facet has two contributors: the first element of facet2el writes it in place, the second one into a side buffer that a
// merge kernel adds afterwards (first + second = the ascending-element order of the gather in ehdg_csr_numeric).
struct ElMap {
    int64_t base[3];     // first value of the row block of local facet a; -1: no rows (boundary facet / another rank's rows)
    int32_t row0[3];     // first global row of that facet
    int32_t len[3];      // row length
    uint8_t coff[3][3];  // column offset of facet b's block inside a row of facet a; 255: facet b has no dofs
    uint8_t hi;          // bit a: this element is the second one of facet a
    uint8_t pad[2];
};
struct CsrOut {
    const ElMap* em;     // [ne]; null: the condensed blocks leave element-wise (S, g)
    double *vals, *rhs;  // value array of the pattern; right-hand side by GLOBAL row
    double *diag2, *g2;  // second contributions: [own rows][nb] and [own rows]
    int64_t row_lo;
    int nb;
    int* flags;          // [own rows], zeroed per call; null: the diagonal blocks are merged by a kernel afterwards.
                         // flags[first own row of a facet] counts the elements that have stored their share of the facet's
                         // diagonal block and right-hand side; the element that finds 1 there adds the two shares in place.
};

// exp(-50) = 1.9e-22: below this the exponential of a layer function is dropped next to its polynomial part
#define EHDG_LAYER_DROP 50.0
// orders whose Gauss-Jordan broadcasts the pivot row by warp shuffle (bit p set); see fast_assemble_element
#ifndef EHDG_GJ_SHFL_ORDERS
#define EHDG_GJ_SHFL_ORDERS 0x60     /* p = 5, 6 */
#endif
#define EHDG_GJ_SHFL_FOR(p) (((EHDG_GJ_SHFL_ORDERS >> (p)) & 1) != 0)
// Schur update of the p = 4 assemble kernel on the FP64 tensor cores (mma.sync m8n8k4).  A/B on B200, 4M elements
// (profiles/r02_kernels.md): 20.26 ms with DMMA against 19.18 ms without (bit-identical S): the 32 DMMA per warp and the
// fragment-layout A_FV build cost more than the 104 shared-memory wavefronts per element they save.  Off by default.
#ifndef EHDG_SCHUR_DMMA
#define EHDG_SCHUR_DMMA 0
#endif

// entry j of row slot ho of a lane that keeps R rows (static register indexing)
template <int R, int NC>
EHDG_HD double own_row_entry(const double (*rows)[NC], int ho, int j) {
    double v = rows[0][j];
#pragma unroll
    for (int h = 1; h < R; ++h)
        if (h == ho) v = rows[h][j];
    return v;
}

template <int LG>
struct Group {
    int lane;       // lane inside the group (device); unused on the host
    double* sh;     // group-private scratch
    // Device: where the group's lanes sit in the warp.  il = 1: LG consecutive threads (group g = threads g LG ...).
    // il = 32 / LG: interleaved, lane l of group g is thread l * il + g, so that the lanes of the same ROW of all the
    // groups of a warp are neighbours: when one lane per group stores its row (16-byte accesses are issued per
    // quarter warp) the warp needs one shared-memory wavefront instead of one per group.
    int il = 1;

    // Row holding the largest key.  Every lane offers R candidate rows (row = h * LG + lane); keys are
    // magnitudes >= 0, or -1 for "not a candidate".  They are compared through the top 32 bits of the
    // double with the low bits replaced by the row number, so the search is ONE 32-bit warp reduction
    // (redux.sync) per group; pivots that agree to ~15 bits of mantissa are tie-broken towards the
    // lower row (immaterial for partial pivoting).
    template <int ROWS>
    static EHDG_HD unsigned order_key(double key, int row) {
        if (!(key >= 0.0)) return 0u;                       // -1 / NaN: never wins against a real candidate
        unsigned long long bits;
#ifdef __CUDA_ARCH__
        bits = (unsigned long long)__double_as_longlong(key);
#else
        memcpy(&bits, &key, 8);
#endif
        const unsigned hi = (unsigned)(bits >> 32);
        return ((hi | 0x80000000u) & ~(unsigned)(ROWS - 1)) | (unsigned)(ROWS - 1 - row);
    }
    // key[L][h]: candidate of (lane, slot h); returns the winning row
    template <int R>
    EHDG_HD int argmax_row(const double (*key)[R]) const {
        constexpr int ROWS = LG * R;
#ifdef __CUDA_ARCH__
        unsigned k = 0u;
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const unsigned kh = order_key<ROWS>(key[0][h], h * LG + lane);
            k = kh > k ? kh : k;
        }
        const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(LG - 1);
        const unsigned mask = (LG == 32) ? 0xffffffffu : (((1u << LG) - 1u) << base);
        const unsigned best = __reduce_max_sync(mask, k);
        return ROWS - 1 - (int)(best & (unsigned)(ROWS - 1));
#else
        unsigned best = 0u;
        for (int l = 0; l < LG; ++l)
            for (int h = 0; h < R; ++h) {
                const unsigned k = order_key<ROWS>(key[l][h], h * LG + l);
                if (k > best) best = k;
            }
        return ROWS - 1 - (int)(best & (unsigned)(ROWS - 1));
#endif
    }
    EHDG_HD double sum(const double* x) const {
#ifdef __CUDA_ARCH__
        double s = x[0];
#pragma unroll
        for (int d = LG / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d * il);
        return s;
#else
        // same pairing as the butterfly so host and device round identically
        double t[LG];
        for (int l = 0; l < LG; ++l) t[l] = x[l];
        for (int d = LG / 2; d > 0; d >>= 1) {
            double u[LG];
            for (int l = 0; l < LG; ++l) u[l] = t[l] + t[l ^ d];
            for (int l = 0; l < LG; ++l) t[l] = u[l];
        }
        return t[0];
#endif
    }
    // first lane of the group whose flag is non-zero, LG if none
    EHDG_HD int first_set(const int* flag) const {
#ifdef __CUDA_ARCH__
        const unsigned b = __ballot_sync(0xffffffffu, flag[0] != 0);
        const int t = threadIdx.x & 31;
        if (il == 1) {
            const int base = t & ~(LG - 1);
            const unsigned m = (LG == 32) ? b : ((b >> base) & ((1u << LG) - 1u));
            return m ? __ffs((int)m) - 1 : LG;
        }
        // interleaved: the group's lanes are the bits g, g + il, g + 2 il, ...
        const int g = t % il;
        unsigned pat = 0u;
#pragma unroll
        for (int l = 0; l < LG; ++l) pat |= 1u << (l * (32 / LG));
        const unsigned m = (b >> g) & pat;
        return m ? (__ffs((int)m) - 1) / il : LG;
#else
        for (int l = 0; l < LG; ++l)
            if (flag[l]) return l;
        return LG;
#endif
    }
};

// geometry of one element in its vertex-sorted frame (computed redundantly by every lane)
struct SortedGeom {
    double adet;
    double g00, g01, g11;       // Gram of grad lam_Q0, grad lam_Q1
    double c0, c1;              // b . grad lam_Qa
    double len[3], h[3], ih[3], bn[3]; // per sorted edge s (opposite Q_s); ih = 1/h
    double ga0[3], ga1[3];      // n_s . grad lam_Q0 / Q1
    int eloc[3];                // local edge number of sorted edge s
    int f0, f1;                 // local index of Q0, Q1
    double omx[3], omy[3];      // 1-x, 1-y at the LOCAL vertices
};

EHDG_HD void sorted_geom(const double* verts, const int* tri, double bx, double by, SortedGeom& sg) {
    // vertex triples (global number, x, y, local index), sorted by global number with three compare-exchanges;
    // scalars only, so nothing is indexed dynamically (no local memory)
    int n0 = tri[0], n1 = tri[1], n2 = tri[2];
    double x0 = verts[2 * n0], y0 = verts[2 * n0 + 1];
    double x1 = verts[2 * n1], y1 = verts[2 * n1 + 1];
    double x2 = verts[2 * n2], y2 = verts[2 * n2 + 1];
    sg.omx[0] = 1.0 - x0; sg.omy[0] = 1.0 - y0;
    sg.omx[1] = 1.0 - x1; sg.omy[1] = 1.0 - y1;
    sg.omx[2] = 1.0 - x2; sg.omy[2] = 1.0 - y2;
    int i0 = 0, i1 = 1, i2 = 2;
#define EHDG_CSWAP(na, xa, ya, ia, nb, xb, yb, ib)                 \
    if (nb < na) {                                                 \
        const int tn = na; na = nb; nb = tn;                       \
        const double tx = xa; xa = xb; xb = tx;                    \
        const double ty = ya; ya = yb; yb = ty;                    \
        const int ti = ia; ia = ib; ib = ti;                       \
    }
    EHDG_CSWAP(n0, x0, y0, i0, n1, x1, y1, i1)
    EHDG_CSWAP(n1, x1, y1, i1, n2, x2, y2, i2)
    EHDG_CSWAP(n0, x0, y0, i0, n1, x1, y1, i1)
#undef EHDG_CSWAP
    sg.f0 = i0;
    sg.f1 = i1;
    // local vertex v is opposite local edge: V0 -> e1, V1 -> e0, V2 -> e2
    sg.eloc[0] = (i0 == 0) ? 1 : (i0 == 1 ? 0 : 2);
    sg.eloc[1] = (i1 == 0) ? 1 : (i1 == 1 ? 0 : 2);
    sg.eloc[2] = (i2 == 0) ? 1 : (i2 == 1 ? 0 : 2);
    const double ax = x0 - x2, ay = y0 - y2;
    const double bx_ = x1 - x2, by_ = y1 - y2;
    const double det = ax * by_ - bx_ * ay;
    const double id = 1.0 / det;
    sg.adet = fabs(det);
    double gl[3][2];
    gl[0][0] = by_ * id;  gl[0][1] = -bx_ * id;
    gl[1][0] = -ay * id;  gl[1][1] = ax * id;
    gl[2][0] = -gl[0][0] - gl[1][0];
    gl[2][1] = -gl[0][1] - gl[1][1];
    sg.g00 = gl[0][0] * gl[0][0] + gl[0][1] * gl[0][1];
    sg.g01 = gl[0][0] * gl[1][0] + gl[0][1] * gl[1][1];
    sg.g11 = gl[1][0] * gl[1][0] + gl[1][1] * gl[1][1];
    sg.c0 = bx * gl[0][0] + by * gl[0][1];
    sg.c1 = bx * gl[1][0] + by * gl[1][1];
    for (int s = 0; s < 3; ++s) {
        // outward normal of the edge opposite Q_s is -grad lam_Qs / |grad lam_Qs|, |grad lam_Qs| = 1/h_s
        const double g2 = gl[s][0] * gl[s][0] + gl[s][1] * gl[s][1];
#ifdef __CUDA_ARCH__
        const double ign = rsqrt(g2);
#else
        const double ign = 1.0 / sqrt(g2);
#endif
        const double nx = -gl[s][0] * ign, ny = -gl[s][1] * ign;
        sg.h[s] = ign;                      // = |detJ| / |e|   
        sg.ih[s] = g2 * ign;
        sg.len[s] = sg.adet * (g2 * ign);
        sg.bn[s] = bx * nx + by * ny;
        sg.ga0[s] = nx * gl[0][0] + ny * gl[0][1];
        sg.ga1[s] = nx * gl[1][0] + ny * gl[1][1];
    }
}

// -------------------------------------------------------------------------------------------
// Layout: a group of LG lanes owns one element, every lane keeps R rows of the element matrices in
// registers, row = h * LG + lane (h < R).  Two rows per lane halve the shared-memory broadcast volume
// and the pivot-search / synchronisation overhead per element (ncu: the kernels are bound by issue
// slots and the shared-memory pipe, not by latency).
// -------------------------------------------------------------------------------------------
template <int P, int RR>
struct FastDimsR {
    static constexpr int NP = (P + 1) * (P + 2) / 2;
    static constexpr int NFL = P + 1;
    static constexpr int NF = 3 * NFL;
    static constexpr int NC = NP + NF + 1;
    static constexpr int M = NP > NF ? NP : NF;
    static constexpr int R = RR;                        // rows per lane
    static constexpr int LG = R == 2 ? (M <= 8 ? 4 : (M <= 16 ? 8 : 16)) : (M <= 8 ? 8 : (M <= 16 ? 16 : 32));
    static constexpr int ROWS = LG * R;
    static constexpr int NXS = ((NF + 2) & ~1) + 2;    // row stride of the X scratch: even (16-byte rows), 2-way conflicts at most
};

// measured on B200 (tools/ab_time.py, A/B builds): two rows per lane pay off in the alpha kernel (-23 % at p = 4) but not in
// the assembly kernel (254 registers, 2 CTAs/SM: +7 %)
#ifndef EHDG_ASM_R
#define EHDG_ASM_R 1
#endif
#ifndef EHDG_ALPHA_R
#define EHDG_ALPHA_R 2
#endif
#ifndef EHDG_ALPHA_R2_MAXP
#define EHDG_ALPHA_R2_MAXP 5     /* A/B at 1M: p = 5 21.4 -> 15.9 ms with two rows per lane (254 registers), p = 6 40.4 -> 42.8 ms */
#endif
template <int P>
using FastDims = FastDimsR<P, (EHDG_ASM_R == 2 && P <= 4) ? 2 : 1>;
template <int P>
using AlphaDims = FastDimsR<P, (EHDG_ALPHA_R == 2 && P <= EHDG_ALPHA_R2_MAXP) ? 2 : 1>;

template <int P>
EHDG_HD constexpr int fast_asm_scratch_doubles(int nq) {
    typedef FastDims<P> D;
    constexpr int xs = D::NP * D::NXS, ss = D::NF * D::NF + D::NF;      // X scratch doubles as the S/g staging area
    return ((D::NC + 1) & ~1) + 2 + 12 + (((xs > ss ? xs : ss) + 1) & ~1) + ((nq + 1) & ~1);
}

// a5-a7 on one polynomial element.  Scratch `gc.sh` needs fast_asm_scratch_doubles() doubles.
// Outputs in LOCAL edge order (3 (p+1) facet dofs): S[NF*NF], gvec[NF], W[NP*NF], z[NP].
// tab: assembly tensors (shared memory on the device), phi_rhs: [9][nq][NP] (global/L1), rs: bonus rule.
template <int P>
EHDG_HD int fast_assemble_element(const Group<FastDims<P>::LG>& gc, const Problem& pb, const RuleSet& rs,
                                  const double* tab, const double* phi_rhs, const double* verts, const int* tri,
                                  double lam, bool store, double* S, double* gvec, double* W, double* z,
                                  const LayerConst& cx, const LayerConst& cy, const ElMap* em = nullptr,
                                  const CsrOut* co = nullptr) {
    typedef FastDims<P> D;
    constexpr int NP = D::NP, NFL = D::NFL, NF = D::NF, NC = D::NC, LG = D::LG, R = D::R, NXS = D::NXS;
    constexpr int NLS = EHDG_NLS(LG), NFS = NF | 1;     // NFS: row stride of the facet tensors
    const TensorLayout TL = tensor_layout(P, rs.nq);
    SortedGeom sg;
    sorted_geom(verts, tri, pb.bx, pb.by, sg);
    double* prow = gc.sh;                       // NC (padded even) + 2: pivot row and the reciprocal of the pivot
    double* geo = prow + ((NC + 1) & ~1) + 2;   // 12: per-edge coefficients of A_FV / A_FF, parked across the elimination
    double* Xs = geo + 12;                      // NP x NXS, later the staging area of S (NF x NF) and g (NF)
    double* cw = Xs + (((NP * NXS > NF * NF + NF ? NP * NXS : NF * NF + NF) + 1) & ~1);   // nq
    // per-edge coefficients
    double cvv[3], cvf[3], cfv[3], cff[3], eg0[3], eg1[3];
    for (int s = 0; s < 3; ++s) {
        const double bn = sg.bn[s];
        const double pos = bn > 0.0 ? 1.0 : 0.0, neg = 1.0 - pos;
        const double tau = pb.eps * (20.0 * lam) * (double)(P * P) * sg.ih[s];
        cvv[s] = sg.len[s] * (tau + pos * bn);
        cvf[s] = sg.len[s] * (-tau + neg * bn);
        cfv[s] = sg.len[s] * (-tau - pos * bn - pos);
        cff[s] = sg.len[s] * (tau - neg * bn + pos);
        eg0[s] = sg.len[s] * pb.eps * sg.ga0[s];
        eg1[s] = sg.len[s] * pb.eps * sg.ga1[s];
    }
    const double kxx = sg.adet * pb.eps * sg.g00, ks = sg.adet * pb.eps * sg.g01, kyy = sg.adet * pb.eps * sg.g11;
    const double ccx = -sg.adet * sg.c0, ccy = -sg.adet * sg.c1;

    double a[NLS][R][NC];      // rows of [A_VV | A_VF | f_V]          (rows < NP)
    int mycol[NLS][R];
    double myinv[NLS][R];
    int status = 0;

    // ---- right-hand side.  Where the exponentials of both layer functions of `coeff` are below 2e-22 on the whole
    // element (k d > EHDG_LAYER_DROP at the vertex nearest to the outflow side, hence at every point) they cannot change
    // a double next to the polynomial part of L(d) = (1 - d) + (E - e0) inv: |E inv| <= 2e-22 while 1 - d only gets
    // that small where E - e0 ~ k e0 (1 - d) shrinks with it.  The coefficient is then a quadratic polynomial in the
    // local barycentrics and f_V comes from the six moment vectors behind the phi table (the order 2p + bonus rule
    // integrates that product exactly, so this is the quadrature's value up to rounding).
    // Otherwise: weights at the quadrature points, cw[q] = w_q |detJ| coeff(x_q).
    // cx, cy = layer_const(pb.coeff.k[0 / 1]): the same for every element, computed once by the caller
    const bool ux = pb.coeff.c[1] != 0.0 || pb.coeff.c[3] != 0.0, uy = pb.coeff.c[2] != 0.0 || pb.coeff.c[3] != 0.0;
    const double dminx = fmin(sg.omx[0], fmin(sg.omx[1], sg.omx[2])), dminy = fmin(sg.omy[0], fmin(sg.omy[1], sg.omy[2]));
#ifdef EHDG_NO_RHS_MOMENTS
    const bool poly_rhs = false;
#else
    const bool poly_rhs = !pb.xcoeff && (!ux || cx.k * dminx > EHDG_LAYER_DROP) && (!uy || cy.k * dminy > EHDG_LAYER_DROP);     // uniform over the group
#endif
    const bool zx = cx.k * dminx > EHDG_LAYER_DROP, zy = cy.k * dminy > EHDG_LAYER_DROP;
    double cm[EHDG_RHS_MOMENTS];
    if (poly_rhs) {
        // L(d) = (1 - d) - e0 inv with d = d_2 + l0 (d_0 - d_2) + l1 (d_1 - d_2)
        const double ax = ux ? (1.0 - sg.omx[2]) - cx.e0 * cx.inv : 0.0, ay = uy ? (1.0 - sg.omy[2]) - cy.e0 * cy.inv : 0.0;
        const double bx0 = ux ? sg.omx[2] - sg.omx[0] : 0.0, bx1 = ux ? sg.omx[2] - sg.omx[1] : 0.0;
        const double by0 = uy ? sg.omy[2] - sg.omy[0] : 0.0, by1 = uy ? sg.omy[2] - sg.omy[1] : 0.0;
        const double c0 = pb.coeff.c[0], c1 = pb.coeff.c[1], c2 = pb.coeff.c[2], c3 = pb.coeff.c[3];
        cm[0] = sg.adet * (c0 + c1 * ax + c2 * ay + c3 * ax * ay);
        cm[1] = sg.adet * (c1 * bx0 + c2 * by0 + c3 * (ax * by0 + ay * bx0));
        cm[2] = sg.adet * (c1 * bx1 + c2 * by1 + c3 * (ax * by1 + ay * bx1));
        cm[3] = sg.adet * (c3 * bx0 * by0);
        cm[4] = sg.adet * (c3 * (bx0 * by1 + bx1 * by0));
        cm[5] = sg.adet * (c3 * bx1 * by1);
    } else {
        for (int m = 0; m < EHDG_RHS_MOMENTS; ++m) cm[m] = 0.0;
    }
    const int elpack = sg.eloc[0] | (sg.eloc[1] << 2) | (sg.eloc[2] << 4);     // the one piece of geometry kept in a register
    {
        EHDG_FOR_LANES(l, L)
            if (l == 0) {       // every lane holds the same values: one parks the A_FV / A_FF coefficients for after the elimination
                st2(geo + 0, cfv[0], cfv[1]);
                st2(geo + 2, cfv[2], cff[0]);
                st2(geo + 4, cff[1], cff[2]);
                st2(geo + 6, eg0[0], eg0[1]);
                st2(geo + 8, eg0[2], eg1[0]);
                st2(geo + 10, eg1[1], eg1[2]);
            }
            if (!poly_rhs)
            for (int q = l; q < rs.nq; q += LG) {
                const double l0 = rs.vx[q], l1 = rs.vy[q], l2 = 1.0 - l0 - l1;
                const double dx = l0 * sg.omx[0] + l1 * sg.omx[1] + l2 * sg.omx[2];
                const double dy = l0 * sg.omy[0] + l1 * sg.omy[1] + l2 * sg.omy[2];
                // an axis whose exponential is negligible on the whole element (uniform over the group) skips exp()
                const double lx = ux ? (zx ? (1.0 - dx) - cx.e0 * cx.inv : layer_val(cx, dx)) : 0.0;
                const double ly = uy ? (zy ? (1.0 - dy) - cy.e0 * cy.inv : layer_val(cy, dy)) : 0.0;
                cw[q] = rs.vw[q] * sg.adet * (pb.xcoeff ? expr_eval(pb.xcoeff, 1.0 - dx, 1.0 - dy)
                                                        : pb.coeff.c[0] + pb.coeff.c[1] * lx + pb.coeff.c[2] * ly + pb.coeff.c[3] * lx * ly);
            }
        EHDG_END_LANES
    }

    // ---- build the rows from the tensors
    EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const int row = h * LG + l;
            const int i = row < NP ? row : NP - 1;          // clamp: surplus rows recompute the last row
            const double* tv = tab + TL.offV + i * NP;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
#ifdef EHDG_EXP_SKIP_TABLES
                double acc = kxx * (i + 1) + ks * j + (i == j ? 10.0 * kyy + 100.0 : 0.0) + cvv[0] * 0.01;
#else
                double acc = kxx * tv[j] + ks * tv[NP * NP + j] + kyy * tv[2 * NP * NP + j]
                           + ccx * tv[3 * NP * NP + j] + ccy * tv[4 * NP * NP + j];
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const double* te = tab + TL.offE + (s * 3 * NP + i) * NP;
                    acc += cvv[s] * te[j] - eg0[s] * te[NP * NP + j] - eg1[s] * te[2 * NP * NP + j];
                }
#endif
                a[L][h][j] = acc;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const double* tf = tab + TL.offFT + i * NFS + s * NFL;      // [t][i][(s,k)]: odd lane stride NFS
#pragma unroll
                for (int k = 0; k < NFL; ++k)
                    a[L][h][NP + s * NFL + k] = cvf[s] * tf[k] + eg0[s] * tf[NP * NFS + k] + eg1[s] * tf[2 * NP * NFS + k];
            }
            mycol[L][h] = -1;
            myinv[L][h] = 1.0;
        }
        // f_V[i] = sum_q cw[q] phi_i(q), the R rows of the lane share the cw reads
        if (poly_rhs) {
            const double* mo = phi_rhs + (size_t)9 * rs.nq * NP + (size_t)(sg.f0 * 3 + sg.f1) * EHDG_RHS_MOMENTS * NP;
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int i = h * LG + l < NP ? h * LG + l : NP - 1;
                double acc = 0.0;
#pragma unroll
                for (int m = EHDG_RHS_MOMENTS - 1; m >= 0; --m) acc += cm[m] * mo[m * NP + i];     // small terms first
                a[L][h][NC - 1] = acc;
            }
        } else {
            const double* ph = phi_rhs + (size_t)(sg.f0 * 3 + sg.f1) * rs.nq * NP;
            int ii[R];
            double acc[R][4];                      // four independent accumulation chains per row
#pragma unroll
            for (int h = 0; h < R; ++h) {
                ii[h] = h * LG + l < NP ? h * LG + l : NP - 1;
                acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0;
            }
            int q = 0;
#ifdef EHDG_EXP_SKIP_RHS
            q = rs.nq & ~3;
#endif
            for (; q + 3 < rs.nq; q += 4) {
                double c0, c1, c2, c3;
                ld2(cw + q, c0, c1);
                ld2(cw + q + 2, c2, c3);
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    acc[h][0] += c0 * ph[q * NP + ii[h]];
                    acc[h][1] += c1 * ph[(q + 1) * NP + ii[h]];
                    acc[h][2] += c2 * ph[(q + 2) * NP + ii[h]];
                    acc[h][3] += c3 * ph[(q + 3) * NP + ii[h]];
                }
            }
            for (; q < rs.nq; ++q) {
#pragma unroll
                for (int h = 0; h < R; ++h) acc[h][0] += cw[q] * ph[q * NP + ii[h]];
            }
#pragma unroll
            for (int h = 0; h < R; ++h) a[L][h][NC - 1] = (acc[h][0] + acc[h][1]) + (acc[h][2] + acc[h][3]);
        }
    EHDG_END_LANES

    // ---- Gauss-Jordan with partial pivoting, rows stay where they are.  The pivot of step k+1 is searched as soon
    // as column k+1 has been updated in step k, so that the warp reduction and the reciprocal of the next pivot
    // overlap the rest of the update instead of heading the dependency chain of every step.
#ifdef EHDG_EXP_SKIP_GJ
#define EHDG_GJ_STEPS 1
#else
#define EHDG_GJ_STEPS NP
#endif
    constexpr int NCP = (NC + 1) & ~1;          // prow[NCP] = 1 / pivot
    constexpr bool GJ_SHFL = EHDG_GJ_SHFL_FOR(P);
    int rw;
    {
        double key[NLS][R];
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h) key[L][h] = (h * LG + l < NP) ? fabs(a[L][h][0]) : -1.0;
        EHDG_END_LANES
#ifdef EHDG_EXP_NOPIVOT
        rw = 0;                                 // timing experiment only: natural pivot order (wrong results)
#else
        rw = gc.template argmax_row<R>(key);
#endif
    }
    // Two ways to hand the pivot row to the other lanes, chosen per order from A/B timings on B200 (EHDG_GJ_SHFL_ORDERS):
    // by warp shuffle straight out of the owner lane's registers (no store / barrier / load round trip and no section
    // that only the owner executes: 4-5 % faster for p = 5, 6 where a group is a whole warp), or through the group's
    // shared-memory scratch with 16-byte accesses (fewer instructions per broadcast value: equal at p = 1, 1-4 % faster
    // for p = 2, 3, 4 where one load serves the two groups of a warp).
    if (GJ_SHFL) {
#ifdef __CUDA_ARCH__
    const int grp_base = (int)(threadIdx.x & 31u) & ~(LG - 1);
#define EHDG_GJ_BCAST(j) __shfl_sync(0xffffffffu, own_row_entry<R, NC>(a[0], ho, j), grp_base + lo)
#else
#define EHDG_GJ_BCAST(j) a[lo][ho][j]
#endif
#pragma unroll
    for (int k = 0; k < EHDG_GJ_STEPS; ++k) {
        const int lo = rw % LG, ho = rw / LG;
        double f[NLS][R];
        double key[NLS][R];
        EHDG_FOR_LANES(l, L)
            const double piv = EHDG_GJ_BCAST(k);
            if (!(fabs(piv) > 0.0)) status |= EHDG_ST_SINGULAR;
#ifdef __CUDA_ARCH__
            const double inv = __drcp_rn(piv);
#else
            const double inv = 1.0 / piv;
#endif
            const double x1 = EHDG_GJ_BCAST(k + 1);
            // multipliers: 0 for the pivot row (it stays unscaled; scaled once at the end) and for surplus rows
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const bool is_piv = (l == lo) && (h == ho);
                if (is_piv) {
                    mycol[L][h] = k;
                    myinv[L][h] = inv;
                }
                f[L][h] = (is_piv || h * LG + l >= NP) ? 0.0 : a[L][h][k] * inv;
            }
#pragma unroll
            for (int h = 0; h < R; ++h) {
                a[L][h][k + 1] -= f[L][h] * x1;
                key[L][h] = (h * LG + l < NP && mycol[L][h] < 0) ? fabs(a[L][h][k + 1]) : -1.0;
            }
        EHDG_END_LANES
#ifdef EHDG_EXP_NOPIVOT
        if (k + 1 < NP) rw = k + 1;
#else
        if (k + 1 < NP) rw = gc.template argmax_row<R>(key);
#endif
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int j = k + 2; j < NC; ++j) {
                const double xj = EHDG_GJ_BCAST(j);
#pragma unroll
                for (int h = 0; h < R; ++h) a[L][h][j] -= f[L][h] * xj;
            }
        EHDG_END_LANES
    }
#undef EHDG_GJ_BCAST

    } else {
#pragma unroll
    for (int k = 0; k < EHDG_GJ_STEPS; ++k) {
        const int lo = rw % LG, ho = rw / LG;
        double f[NLS][R];
        EHDG_FOR_LANES(l, L)
            if (l == lo) {
                double piv = a[L][0][k];
#pragma unroll
                for (int h = 1; h < R; ++h)
                    if (h == ho) piv = a[L][h][k];
                if (!(fabs(piv) > 0.0)) status |= EHDG_ST_SINGULAR;
#ifdef __CUDA_ARCH__
                const double inv = __drcp_rn(piv);
#else
                const double inv = 1.0 / piv;
#endif
#pragma unroll
                for (int jp = (k + 1) >> 1; jp <= (NC - 1) >> 1; ++jp) {
                    double v0 = a[L][0][2 * jp], v1 = (2 * jp + 1 < NC) ? a[L][0][2 * jp + 1] : 0.0;
#pragma unroll
                    for (int h = 1; h < R; ++h)
                        if (h == ho) {
                            v0 = a[L][h][2 * jp];
                            v1 = (2 * jp + 1 < NC) ? a[L][h][2 * jp + 1] : 0.0;
                        }
                    st2(prow + 2 * jp, v0, v1);
                }
                prow[NCP] = inv;
#pragma unroll
                for (int h = 0; h < R; ++h)
                    if (h == ho) {
                        mycol[L][h] = k;
                        myinv[L][h] = inv;
                    }
            }
        EHDG_END_LANES
        double key[NLS][R];
        EHDG_FOR_LANES(l, L)
            const double inv = prow[NCP];
            // multipliers: 0 for the pivot row (it stays unscaled; scaled once at the end) and for surplus rows
            const double x1 = prow[k + 1];
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const bool is_piv = (l == lo) && (h == ho);
                f[L][h] = (is_piv || h * LG + l >= NP) ? 0.0 : a[L][h][k] * inv;
                a[L][h][k + 1] -= f[L][h] * x1;
                key[L][h] = (h * LG + l < NP && mycol[L][h] < 0) ? fabs(a[L][h][k + 1]) : -1.0;
            }
        EHDG_END_LANES
#ifdef EHDG_EXP_NOPIVOT
        if (k + 1 < NP) rw = k + 1;
#else
        if (k + 1 < NP) rw = gc.template argmax_row<R>(key);
#endif
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int jp = (k + 2) >> 1; jp <= (NC - 1) >> 1; ++jp) {
                double x0, x1;
                ld2(prow + 2 * jp, x0, x1);
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    if (2 * jp >= k + 2) a[L][h][2 * jp] -= f[L][h] * x0;
                    if (2 * jp + 1 < NC) a[L][h][2 * jp + 1] -= f[L][h] * x1;
                }
            }
        EHDG_END_LANES
    }

    }
    // ---- X = A_VV^-1 [A_VF | f_V]: row c sits where pivot column c was chosen
    EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h)
            if (h * LG + l < NP) {
                const int c = mycol[L][h];
                const double sc = myinv[L][h];
#pragma unroll
                for (int j = 0; j <= NF; ++j) a[L][h][NP + j] *= sc;
#pragma unroll
                for (int jp = 0; jp <= NF >> 1; ++jp)
                    st2(Xs + c * NXS + 2 * jp, a[L][h][NP + 2 * jp], (2 * jp + 1 <= NF) ? a[L][h][NP + 2 * jp + 1] : 0.0);
            }
    EHDG_END_LANES

#if defined(__CUDA_ARCH__) && EHDG_SCHUR_DMMA
    if constexpr (P == 4 && R == 1 && LG == 16) {
        // ---- S = A_FF - A_FV X, g = -A_FV z on the FP64 tensor cores (mma.sync m8n8k4): per element a 16 x 16 x 16 product
        // (NF = 15 rows, NP = 15 inner, NF + 1 = 16 columns: S and g together).  A_FV is formed directly in A-fragment layout
        // from the facet tensor, X is read in B-fragment layout (one conflict-free LDS.64 per fragment instead of the 120
        // broadcast LDS.128 per warp of the row-per-lane update), the result leaves in C-fragment layout through the
        // staging area.  The two elements of a warp are processed one after the other by all 32 lanes.
        const unsigned FULL = 0xffffffffu;
        const int wl = (int)(threadIdx.x & 31u), q4 = wl >> 2, r4 = wl & 3, myg = wl >> 4;
        const int scr = fast_asm_scratch_doubles<P>(rs.nq);
        int eloc_[3];
        eloc_[0] = elpack & 3;
        eloc_[1] = (elpack >> 2) & 3;
        eloc_[2] = (elpack >> 4) & 3;
        if (store) {
            int sinv_[3];
            for (int s2 = 0; s2 < 3; ++s2) sinv_[eloc_[s2]] = s2;
            const int l = gc.lane;
            for (int col = l; col < NF; col += LG) {
                const int src = sel3i(sinv_, col / NFL) * NFL + col % NFL;
#pragma unroll
                for (int c = 0; c < NP; ++c) W[c * NF + col] = Xs[c * NXS + src];
            }
            for (int c = l; c < NP; c += LG) z[c] = Xs[c * NXS + NF];
        }
        // B fragments of both elements before the staging area (which aliases X) is written
        double bfr[2][4][2];
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
            const double* Xe = Xs + (e2 - myg) * scr;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int tc = 0; tc < 2; ++tc) {
                    const int c = 4 * ks + r4, j = 8 * tc + q4;
                    bfr[e2][ks][tc] = c < NP ? Xe[c * NXS + j] : 0.0;
                }
        }
        __syncwarp();
        const double* tfb = tab + TL.offFT;
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
            double* Se = Xs + (e2 - myg) * scr;
            const double* ge = geo + (e2 - myg) * scr;
            const int elp = __shfl_sync(FULL, elpack, e2 * 16);
            int el3[3];
            el3[0] = elp & 3;
            el3[1] = (elp >> 2) & 3;
            el3[2] = (elp >> 4) & 3;
            int bad = 0;
#pragma unroll
            for (int tr = 0; tr < 2; ++tr) {
                const int rr = 8 * tr + q4;
                const bool rv = rr < NF;
                const int r = rv ? rr : NF - 1;
                const int s = r / NFL, k = r - s * NFL;
                const double c_fv = ge[s], c_g0 = ge[6 + s], c_g1 = ge[9 + s], c_ff = ge[3 + s];
                double af[4];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int jj = 4 * ks + r4;
                    const int j = jj < NP ? jj : NP - 1;
                    const double v = c_fv * tfb[j * NFS + r] + c_g0 * tfb[(NP + j) * NFS + r] + c_g1 * tfb[(2 * NP + j) * NFS + r];
                    af[ks] = (rv && jj < NP) ? v : 0.0;
                }
                const int orow = sel3i(el3, s) * NFL + k;
                const double dia = c_ff / (2.0 * k + 1.0);
#pragma unroll
                for (int tc = 0; tc < 2; ++tc) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c0), "+d"(c1)
                                     : "d"(af[ks]), "d"(bfr[e2][ks][tc]));
                    if (rv) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int col = 8 * tc + 2 * r4 + u;
                            double val = -(u == 0 ? c0 : c1);
                            if (col == r) val += dia;
                            if (val != val) bad = 1;
                            if (col < NF) {
                                const int s2 = col / NFL;
                                Se[orow * NF + sel3i(el3, s2) * NFL + (col - s2 * NFL)] = val;
                            } else {
                                Se[NF * NF + orow] = val;
                            }
                        }
                    }
                }
            }
            if (__any_sync(FULL, bad) && myg == e2) status |= EHDG_ST_NONFINITE;
        }
        __syncwarp();
    } else
#endif
    {
    // ---- rows of A_FV and the diagonal of A_FF, r = (s, k) in the sorted frame.  Built only now, from per-edge
    // coefficients that waited in shared memory: nothing of the geometry stays live in registers across the elimination,
    // which is what lets the kernel run four CTAs per SM.
    double afv[NLS][R][NP];
    double aff[NLS][R];
    int eloc[3];
    eloc[0] = elpack & 3;
    eloc[1] = (elpack >> 2) & 3;
    eloc[2] = (elpack >> 4) & 3;
    {
        EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const int row = h * LG + l;
            const int r = row < NF ? row : NF - 1;
            const int s = r / NFL, k = r % NFL;
            const double* tf = tab + TL.offFT + r;              // [t][j][(s,k)]: unit stride across the lanes
            double cfv2[3], cff2[3], fg0[3], fg1[3];
            ld2(geo + 0, cfv2[0], cfv2[1]);
            ld2(geo + 2, cfv2[2], cff2[0]);
            ld2(geo + 4, cff2[1], cff2[2]);
            ld2(geo + 6, fg0[0], fg0[1]);
            ld2(geo + 8, fg0[2], fg1[0]);
            ld2(geo + 10, fg1[1], fg1[2]);
            const double c_fv = sel3(cfv2, s), c_g0 = sel3(fg0, s), c_g1 = sel3(fg1, s);
#pragma unroll
            for (int j = 0; j < NP; ++j)
                afv[L][h][j] = c_fv * tf[j * NFS] + c_g0 * tf[(NP + j) * NFS] + c_g1 * tf[(2 * NP + j) * NFS];
            double rk = 1.0;                                 // 1 / (2k+1) = int_0^1 mu_k^2 dt, k <= 6
#pragma unroll
            for (int kk = 1; kk < NFL; ++kk)
                if (k == kk) rk = 1.0 / (2.0 * kk + 1.0);
            aff[L][h] = sel3(cff2, s) * rk;
        }
        EHDG_END_LANES
    }

    // ---- S = A_FF - A_FV X, g = -A_FV z ; rows in the sorted frame.  W, z leave from the X scratch and S, g from a
    // staging copy with contiguous (full-sector) global stores in local edge order.
    int sinv[3];                                    // sorted edge of local edge e
    for (int s2 = 0; s2 < 3; ++s2) sinv[eloc[s2]] = s2;
    double srow[NLS][R][NF + 1];
    EHDG_FOR_LANES(l, L)
        if (store) {
            // lane l owns the output columns l, l + LG, ...: source column fixed per lane, rows by constant strides
            for (int col = l; col < NF; col += LG) {
                const int src = sel3i(sinv, col / NFL) * NFL + col % NFL;
#pragma unroll
                for (int c = 0; c < NP; ++c) W[c * NF + col] = Xs[c * NXS + src];
            }
            for (int c = l; c < NP; c += LG) z[c] = Xs[c * NXS + NF];
        }
#pragma unroll
        for (int h = 0; h < R; ++h)
#pragma unroll
            for (int j = 0; j <= NF; ++j) srow[L][h][j] = 0.0;
#ifdef EHDG_EXP_SKIP_SCHUR
#define EHDG_SCHUR_N 1
#else
#define EHDG_SCHUR_N NP
#endif
#pragma unroll
        for (int c = 0; c < EHDG_SCHUR_N; ++c) {
#pragma unroll
            for (int jp = 0; jp <= NF >> 1; ++jp) {
                double x0, x1;
                ld2(Xs + c * NXS + 2 * jp, x0, x1);
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    srow[L][h][2 * jp] -= afv[L][h][c] * x0;
                    if (2 * jp + 1 <= NF) srow[L][h][2 * jp + 1] -= afv[L][h][c] * x1;
                }
            }
        }
    EHDG_END_LANES
    double* Ss = Xs;                                // every lane is done reading X
    EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const int row = h * LG + l;
            if (row < NF) {
                const int s = row / NFL, k = row % NFL;
                const int orow = sel3i(eloc, s) * NFL + k;
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2)
#pragma unroll
                    for (int k2 = 0; k2 < NFL; ++k2) {
                        double val = srow[L][h][s2 * NFL + k2];
                        if (s2 * NFL + k2 == row) val += aff[L][h];
                        if (val != val) status |= EHDG_ST_NONFINITE;
                        Ss[orow * NF + eloc[s2] * NFL + k2] = val;
                    }
                Ss[NF * NF + orow] = srow[L][h][NF];
            }
        }
    EHDG_END_LANES
    }
    double* Ss = Xs;                                // every lane is done reading X (repeated for the copy-out below)
    EHDG_FOR_LANES(l, L)
        if (store && em) {
            // straight into the pattern: lane l owns column (b, lc) of the local block; the NFL lanes of one facet write
            // NFL consecutive doubles of a row
            for (int col = l; col < NF; col += LG) {
                const int b = col / NFL, lc = col - b * NFL;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int64_t base = em->base[a];
                    const int cf = em->coff[a][b];
                    if (base < 0 || cf == 255) continue;
                    double* dst;
                    size_t stride;
                    if (a == b && ((em->hi >> a) & 1)) {
                        dst = co->diag2 + (size_t)(em->row0[a] - co->row_lo) * co->nb + lc;
                        stride = (size_t)co->nb;
                    } else {
                        dst = co->vals + base + cf + lc;
                        stride = (size_t)em->len[a];
                    }
#pragma unroll
                    for (int i = 0; i < NFL; ++i) dst[i * stride] = Ss[(a * NFL + i) * NF + col];
                }
            }
            for (int r = l; r < NF; r += LG) {
                const int a = r / NFL, i = r - a * NFL;
                if (em->base[a] < 0) continue;
                if ((em->hi >> a) & 1) co->g2[em->row0[a] - co->row_lo + i] = Ss[NF * NF + r];
                else co->rhs[em->row0[a] + i] = Ss[NF * NF + r];
            }
        } else if (store) {
            for (int col = l; col < NF; col += LG) {
#pragma unroll
                for (int r = 0; r < NF; ++r) S[r * NF + col] = Ss[r * NF + col];
            }
            for (int r = l; r < NF; r += LG) gvec[r] = Ss[NF * NF + r];
        }
    EHDG_END_LANES
#ifdef __CUDA_ARCH__
    if (em && co->flags) {
        // In-kernel merge of the two-contributor entries.  Every lane makes its stores visible (fence), then one lane per
        // group counts this element in on each of its facets; whoever arrives second on a facet (count was 1) finds both
        // shares in memory -- first element's in the value array, second one's in the side buffer, whichever of the two it
        // is itself -- and stores first + second: the ascending-element order of the gather, independent of who merges.
        __threadfence();
        const int wl = threadIdx.x & 31;
        const int leader = (wl / LG) * LG;
        const unsigned gmask = LG >= 32 ? 0xffffffffu : (((1u << (LG & 31)) - 1u) << leader);
        const int l = gc.lane;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const bool has = store && em->base[a] >= 0;
            int old = 0;
            if (has && l == 0) old = atomicAdd(co->flags + (em->row0[a] - co->row_lo), 1);
            old = __shfl_sync(gmask, old, leader);
            if (has && old == 1) {
                __threadfence();
                const int64_t r0 = em->row0[a] - co->row_lo;
                double* v = co->vals + em->base[a] + em->coff[a][a];
                const int64_t len = em->len[a];
                for (int e = l; e < NFL * NFL; e += LG) {
                    const int i = e / NFL, c = e - i * NFL;
                    double* pv = v + i * len + c;
                    *pv = __ldcg(pv) + __ldcg(co->diag2 + (r0 + i) * co->nb + c);
                }
                for (int i = l; i < NFL; i += LG) {
                    double* pr = co->rhs + em->row0[a] + i;
                    *pr = __ldcg(pr) + __ldcg(co->g2 + r0 + i);
                }
            }
        }
    }
#endif
    return status;
}

// -------------------------------------------------------------------------------------------
// a4 on one polynomial element: lambda_max of (m, a') on the non-constant modes (the constant
// mode decouples: row 0 of m vanishes and f f^T only touches a'[0][0]).
// Scratch: fast_alpha_scratch_doubles() doubles.
// -------------------------------------------------------------------------------------------
template <int P>
EHDG_HD constexpr int fast_alpha_scratch_doubles() {
    return (((AlphaDims<P>::NP - 1) * (AlphaDims<P>::NP - 1) + 1) & ~1) + 4 * 32;     // even: groups stay 16-byte aligned
}

template <int P>
EHDG_HD double fast_alpha_element(const Group<AlphaDims<P>::LG>& gc, const double* tab, const double* verts,
                                  const int* tri, int* status_out) {
    typedef AlphaDims<P> D;
    constexpr int NP = D::NP, LG = D::LG, R = D::R, N = NP - 1;
    constexpr int NLS = EHDG_NLS(LG);
    const TensorLayout TL = tensor_layout(P, 0);
    SortedGeom sg;
    sorted_geom(verts, tri, 0.0, 0.0, sg);
    double* shT = gc.sh;                      // N x N
    double* shv = shT + ((N * N + 1) & ~1);   // 32
    double* shq = shv + 32;                   // 32
    double* shd = shq + 32;                   // 32
    double* she = shd + 32;                   // 32
    int status = 0;

    double A[NLS][R][N];   // a' rows -> Cholesky factor rows
    double M[NLS][R][N];   // m rows  -> C rows
    double eown[NLS][R];   // sub-diagonal entry T[i][i-1] owned by row i
    double invd[NLS][R];   // 1 / l_ii of the Cholesky factor

    // m = sum_s h_s len_s (dn u)(dn v) with dn = ga0 dX + ga1 dY ; the common factor |detJ| of a'
    // and m is dropped (h_s len_s = |detJ|)
    double w0[3], w1[3], w2[3];
    for (int s = 0; s < 3; ++s) {
        w0[s] = sg.ga0[s] * sg.ga0[s];
        w1[s] = sg.ga0[s] * sg.ga1[s];
        w2[s] = sg.ga1[s] * sg.ga1[s];
    }
    EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const int row = h * LG + l;
            const int i = (row < N ? row : N - 1) + 1;
            const double* tk = tab + TL.offAK + i * NP + 1;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                A[L][h][j] = sg.g00 * tk[j] + sg.g01 * tk[NP * NP + j] + sg.g11 * tk[2 * NP * NP + j];
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const double* tn = tab + TL.offAN + (s * 3 * NP + i) * NP + 1;
                    acc += w0[s] * tn[j] + w1[s] * tn[NP * NP + j] + w2[s] * tn[2 * NP * NP + j];
                }
                M[L][h][j] = acc;
            }
            eown[L][h] = 0.0;
            invd[L][h] = 0.0;
        }
    EHDG_END_LANES

    // ---- Cholesky a' = L L^T, right-looking; A[row][k] becomes l_{row,k} (k <= row)
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int kl = k % LG, kh = k / LG;         // owner of row k (compile-time after unrolling)
        EHDG_FOR_LANES(l, L)
            if (l == kl) {
                double d = 0.0;
#pragma unroll
                for (int h = 0; h < R; ++h)
                    if (h == kh) d = A[L][h][k];
                if (!(d > 0.0)) status |= EHDG_ST_ALPHA_DEFICIENT;
#ifdef __CUDA_ARCH__
                shv[0] = rsqrt(d);
#else
                shv[0] = 1.0 / sqrt(d);
#endif
#pragma unroll
                for (int h = 0; h < R; ++h)
                    if (h == kh) invd[L][h] = shv[0];
            }
        EHDG_END_LANES
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                if (row >= k && row < N) {
                    A[L][h][k] *= shv[0];
                    shq[row] = A[L][h][k];
                }
            }
        EHDG_END_LANES
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                const double f = (row > k && row < N) ? A[L][h][k] : 0.0;
#pragma unroll
                for (int j = k + 1; j < N; ++j) A[L][h][j] -= f * shq[j];
            }
        EHDG_END_LANES
    }
    // ---- two forward substitutions with a transpose in between: C = L^-1 (L^-1 m)^T.
    // Row j is broadcast UNSCALED (Y~_j = l_jj Y_j) together with 1/l_jj; the rows below subtract
    // (l_ij / l_jj) Y~_j and every row is scaled once when the pass is finished.
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int jl = j % LG, jh = j / LG;
            EHDG_FOR_LANES(l, L)
                if (l == jl) {
#pragma unroll
                    for (int cp = 0; cp <= (N - 1) >> 1; ++cp) {
                        double v0 = 0.0, v1 = 0.0;
#pragma unroll
                        for (int h = 0; h < R; ++h)
                            if (h == jh) {
                                v0 = M[L][h][2 * cp];
                                v1 = (2 * cp + 1 < N) ? M[L][h][2 * cp + 1] : 0.0;
                            }
                        st2(shT + 2 * cp, v0, v1);
                    }
#pragma unroll
                    for (int h = 0; h < R; ++h)
                        if (h == jh) shv[0] = invd[L][h];
                }
            EHDG_END_LANES
            EHDG_FOR_LANES(l, L)
                double f[R];
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    const int row = h * LG + l;
                    f[h] = (row > j && row < N) ? A[L][h][j] * shv[0] : 0.0;
                }
#pragma unroll
                for (int cp = 0; cp <= (N - 1) >> 1; ++cp) {
                    double x0, x1;
                    ld2(shT + 2 * cp, x0, x1);
#pragma unroll
                    for (int h = 0; h < R; ++h) {
                        M[L][h][2 * cp] -= f[h] * x0;
                        if (2 * cp + 1 < N) M[L][h][2 * cp + 1] -= f[h] * x1;
                    }
                }
            EHDG_END_LANES
        }
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h)
#pragma unroll
                for (int c = 0; c < N; ++c) M[L][h][c] *= invd[L][h];
        EHDG_END_LANES
        if (pass == 0) {
            EHDG_FOR_LANES(l, L)
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    const int row = h * LG + l;
                    if (row < N) {
#pragma unroll
                        for (int c = 0; c < N; ++c) shT[row * N + c] = M[L][h][c];
                    }
                }
            EHDG_END_LANES
            EHDG_FOR_LANES(l, L)
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    const int row = h * LG + l;
                    if (row < N) {
#pragma unroll
                        for (int c = 0; c < N; ++c) M[L][h][c] = shT[c * N + row];
                    }
                }
            EHDG_END_LANES
        }
    }
    // ---- Householder tridiagonalisation of C (rows in M)
#pragma unroll
    for (int k = 0; k + 2 < N; ++k) {
        const int ol = (k + 1) % LG, oh = (k + 1) / LG;       // owner of row k+1
        double part[NLS], vi[NLS][R], pi[NLS][R];
        EHDG_FOR_LANES(l, L)
            double acc = 0.0;
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                if (row > k && row < N) acc += M[L][h][k] * M[L][h][k];
                if (l == ol && h == oh) shv[0] = M[L][h][k];
            }
            part[L] = acc;
        EHDG_END_LANES
        const double sig = gc.sum(part);
        const double x0 = shv[0];
        const bool skip = !(sig > 0.0);
#ifdef __CUDA_ARCH__
        const double nrm = skip ? 0.0 : sig * rsqrt(sig);
#else
        const double nrm = sqrt(sig);
#endif
        const double alpha = (x0 > 0.0) ? -nrm : nrm;
        const double beta = skip ? 0.0 : 1.0 / (sig - x0 * alpha);     // 2 / v^T v
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                vi[L][h] = skip ? 0.0 : ((row == k + 1) ? x0 - alpha : ((row > k + 1 && row < N) ? M[L][h][k] : 0.0));
                if (row < 32) shq[row] = vi[L][h];
                if (row == k + 1) eown[L][h] = skip ? x0 : alpha;
            }
        EHDG_END_LANES
        EHDG_FOR_LANES(l, L)
            double acc2 = 0.0;
            double ps[R];
#pragma unroll
            for (int h = 0; h < R; ++h) ps[h] = 0.0;
#pragma unroll
            for (int j = k + 1; j < N; ++j) {
                const double vj = shq[j];
#pragma unroll
                for (int h = 0; h < R; ++h) ps[h] += M[L][h][j] * vj;
            }
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                pi[L][h] = (row > k && row < N) ? beta * ps[h] : 0.0;
                acc2 += vi[L][h] * pi[L][h];
            }
            part[L] = acc2;
        EHDG_END_LANES
        const double kk = 0.5 * beta * gc.sum(part);
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const int row = h * LG + l;
                pi[L][h] -= kk * vi[L][h];            // q_i
                if (row < 32) shd[row] = pi[L][h];
            }
        EHDG_END_LANES
        EHDG_FOR_LANES(l, L)
#pragma unroll
            for (int j = k + 1; j < N; ++j) {
                const double qj = shd[j], vj = shq[j];
#pragma unroll
                for (int h = 0; h < R; ++h) M[L][h][j] -= vi[L][h] * qj + pi[L][h] * vj;     // rows <= k: vi = pi = 0
            }
        EHDG_END_LANES
    }
    // ---- gather the tridiagonal matrix
    EHDG_FOR_LANES(l, L)
#pragma unroll
        for (int h = 0; h < R; ++h) {
            const int row = h * LG + l;
            double d = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j == row) d = M[L][h][j];
            if (N >= 2 && row == N - 1) eown[L][h] = M[L][h][N - 2];
            if (row < 32) {
                shd[row] = d;
                she[row] = eown[L][h];
            }
        }
    EHDG_END_LANES
    double dg[N], e2[N];
    double lo = 0.0, hi = 0.0, scale = 0.0;
    {
        // every lane reads the whole tridiagonal matrix (uniform across lanes)
        for (int i = 0; i < N; ++i) {
            const double r = (i > 0 ? fabs(she[i]) : 0.0) + (i + 1 < N ? fabs(she[i + 1]) : 0.0);
            const double up = shd[i] + r, dn = shd[i] - r;
            if (i == 0 || up > hi) hi = up;
            if (i == 0 || dn < lo) lo = dn;
        }
        scale = fabs(hi) > fabs(lo) ? fabs(hi) : fabs(lo);
        if (!(scale > 0.0)) scale = 1.0;
        const double is = 1.0 / scale;
        for (int i = 0; i < N; ++i) {
            dg[i] = shd[i] * is;
            e2[i] = (i > 0) ? (she[i] * is) * (she[i] * is) : 0.0;
        }
        lo *= is;
        hi = hi * is + 1e-15;
    }
    // ---- multisection on "all eigenvalues below x" (signs of the Sturm sequence alternate): LG + 1 sections per round
    // (LG + 1)^ROUNDS >= 4e14 sections of a bracket of width <= 2 (scaled spectrum): lambda to ~5e-15 relative to the scale
    constexpr int ROUNDS = LG >= 32 ? 10 : (LG == 16 ? 12 : (LG == 8 ? 16 : 21));
    for (int it = 0; it < ROUNDS; ++it) {
        int below[NLS];
        const double step = (hi - lo) / (double)(LG + 1);
        EHDG_FOR_LANES(l, L)
            const double x = lo + step * (double)(l + 1);
            double pm = 1.0, pc = dg[0] - x;
            bool ok = pc < 0.0;
#pragma unroll
            for (int i = 1; i < N; ++i) {
                const double pn = (dg[i] - x) * pc - e2[i] * pm;
                ok = ok && opposite_signs(pn, pc);
                pm = pc;
                pc = pn;
            }
            below[L] = ok ? 1 : 0;
        EHDG_END_LANES
        const int f = gc.first_set(below);        // first lane whose abscissa is above the spectrum
        const double nlo = lo + step * (double)f;          // x_{f-1}  (f = 0 -> lo)
        const double nhi = (f < LG) ? lo + step * (double)(f + 1) : hi;
        lo = nlo;
        hi = nhi;
    }
    *status_out = status;
    return 0.5 * (lo + hi) * scale;
}

}  // namespace ehdg
