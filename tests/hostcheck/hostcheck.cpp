// TEST-ONLY host build of the per-element device routines (g++, no CUDA): lets the CPU test
// suite compare the math of csrc/ehdg_general.h and the reference tensors with the numpy
// oracle without a GPU.  Never loaded by the product package.
#include <cstring>
#include <vector>
#include "../../include/ehdg.h"
#include "../../convectiondiffusionequation_b200/csrc/ehdg_general.h"
#include "../../convectiondiffusionequation_b200/csrc/ehdg_rules.h"
#include "../../convectiondiffusionequation_b200/csrc/ehdg_problem.h"
#include "../../convectiondiffusionequation_b200/csrc/ehdg_fast.h"
#include "../../convectiondiffusionequation_b200/csrc/edg_element.h"

using namespace ehdg;

struct HostRules {
    TrigRule tr;
    SegRule sr;
    RuleSet rs;
    explicit HostRules(int order) : tr(trig_rule(order)), sr(seg_rule(order)) {
        rs.nq = (int)tr.w.size();
        rs.nqf = (int)sr.w.size();
        rs.vx = tr.x.data(); rs.vy = tr.y.data(); rs.vw = tr.w.data();
        rs.st = sr.t.data(); rs.sw = sr.w.data();
    }
};

template <int P>
static int fast_assemble_host(const Problem& pb, const double* verts, const int* tri, double lam, double* S, double* g,
                              double* W, double* z) {
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> asm_tab, alpha_tab, phi;
    build_tensors(P, 2 * P + pb.bonus, asm_tab, alpha_tab);
    build_rhs_phi(P, 2 * P + pb.bonus, phi);
    std::vector<double> sh(fast_asm_scratch_doubles<P>(hr.rs.nq));
    Group<FastDims<P>::LG> gc{0, sh.data()};
    return fast_assemble_element<P>(gc, pb, hr.rs, asm_tab.data(), phi.data(), verts, tri, lam, true, S, g, W, z,
                                    layer_const(pb.coeff.k[0]), layer_const(pb.coeff.k[1]));
}
template <int P>
static double fast_alpha_host(const double* verts, const int* tri, int bonus, int* status) {
    std::vector<double> asm_tab, alpha_tab;
    build_tensors(P, 2 * P + bonus, asm_tab, alpha_tab);
    std::vector<double> sh(fast_alpha_scratch_doubles<P>());
    Group<AlphaDims<P>::LG> gc{0, sh.data()};
    return fast_alpha_element<P>(gc, alpha_tab.data(), verts, tri, status);
}

extern "C" {

int hc_trig_rule(int order, double* x, double* y, double* w) {
    TrigRule r = trig_rule(order);
    for (size_t i = 0; i < r.w.size(); ++i) { x[i] = r.x[i]; y[i] = r.y[i]; w[i] = r.w[i]; }
    return (int)r.w.size();
}
int hc_seg_rule(int order, double* t, double* w) {
    SegRule r = seg_rule(order);
    for (size_t i = 0; i < r.w.size(); ++i) { t[i] = r.t[i]; w[i] = r.w[i]; }
    return (int)r.w.size();
}
void hc_dubiner(int p, int n, const double* X, const double* Y, double* v, double* dX, double* dY) {
    const int np = (p + 1) * (p + 2) / 2;
    for (int i = 0; i < n; ++i) dubiner_eval(p, X[i], Y[i], v + i * np, dX + i * np, dY + i * np);
}
void hc_sizes(const ehdg_problem* ep, int* out) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    out[0] = pb.n_p; out[1] = pb.nb; out[2] = pb.nvmax; out[3] = pb.nfmax;
}
int hc_assemble(const ehdg_problem* ep, const double* verts, const int* tri, double lam, int volmask,
                const int* facmask, double* S, double* g, double* W, double* z, double* Afull) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> ws(general_ws_doubles(pb, hr.rs.nq, hr.rs.nqf));
    Coop co{0, 1};
    return assemble_condense_general(co, pb, hr.rs, verts, tri, lam, volmask, facmask, ws.data(), S, g, W, z, Afull);
}
// same with the selection rule for dependent volume enrichments at a given tolerance (0 = off)
int hc_assemble_tol(const ehdg_problem* ep, const double* verts, const int* tri, double lam, int volmask,
                    const int* facmask, double vol_drop_tol, double* S, double* g, double* W, double* z, double* Afull) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = vol_drop_tol;
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> ws(general_ws_doubles(pb, hr.rs.nq, hr.rs.nqf));
    Coop co{0, 1};
    return assemble_condense_general(co, pb, hr.rs, verts, tri, lam, volmask, facmask, ws.data(), S, g, W, z, Afull);
}
double hc_alpha(const ehdg_problem* ep, const double* verts, const int* tri, int volmask, int exclude_const,
                int* status) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hb(2 * pb.p + pb.bonus), h0(2 * pb.p);
    std::vector<double> ws(alpha_ws_doubles(pb, h0.rs.nq, h0.rs.nqf));
    Coop co{0, 1};
    *status = 0;
    return alpha_general(co, pb, h0.rs, verts, tri, volmask, exclude_const, ws.data(), status);
}
void hc_prune(const ehdg_problem* ep, const double* verts, const int* tri, int elmark, const int* facmask,
              double threshold, int* vol_drop, int* fac_drop, double* factors) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hb(2 * pb.p + pb.bonus);
    std::vector<double> ws(general_ws_doubles(pb, hb.rs.nq, hb.rs.nqf));
    Coop co{0, 1};
    prune_general(co, pb, hb.rs, verts, tri, elmark, facmask, threshold, ws.data(), vol_drop, fac_drop, factors);
}

int hc_fast_assemble(const ehdg_problem* ep, const double* verts, const int* tri, double lam, double* S, double* g,
                     double* W, double* z) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    switch (pb.p) {
        case 1: return fast_assemble_host<1>(pb, verts, tri, lam, S, g, W, z);
        case 2: return fast_assemble_host<2>(pb, verts, tri, lam, S, g, W, z);
        case 3: return fast_assemble_host<3>(pb, verts, tri, lam, S, g, W, z);
        case 4: return fast_assemble_host<4>(pb, verts, tri, lam, S, g, W, z);
        case 5: return fast_assemble_host<5>(pb, verts, tri, lam, S, g, W, z);
        case 6: return fast_assemble_host<6>(pb, verts, tri, lam, S, g, W, z);
    }
    return -1;
}
double hc_fast_alpha(const ehdg_problem* ep, const double* verts, const int* tri, int* status) {
    switch (ep->order) {
        case 1: return fast_alpha_host<1>(verts, tri, ep->bonus_int, status);
        case 2: return fast_alpha_host<2>(verts, tri, ep->bonus_int, status);
        case 3: return fast_alpha_host<3>(verts, tri, ep->bonus_int, status);
        case 4: return fast_alpha_host<4>(verts, tri, ep->bonus_int, status);
        case 5: return fast_alpha_host<5>(verts, tri, ep->bonus_int, status);
        case 6: return fast_alpha_host<6>(verts, tri, ep->bonus_int, status);
    }
    return -1.0;
}

// EDG block row of one element (csrc/edg_element.h): D [nv x nv], O [3][nv x nv], F [nv], nv = n_p + n_enrich
void hc_edg_element(const ehdg_problem* ep, const double* verts, const int* tris, const int* el2facet, const int* facet2el,
                    const unsigned char* elem_active, const double* alpha_stab, int el, double* D, double* O, double* F) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> ws(edg_ws_doubles(pb, hr.rs.nq, hr.rs.nqf));
    Coop co{0, 1};
    edg_element_blocks(co, pb, hr.rs, verts, tris, el2facet, facet2el, elem_active, alpha_stab, el, ws.data(), D, O, F);
}

int hc_edg_prune(const ehdg_problem* ep, const double* verts, const int* tri, int elmark, double theta, double* factors) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> ws(edg_ws_doubles(pb, hr.rs.nq, hr.rs.nqf) + pb.nvmax * pb.nvmax);
    Coop co{0, 1};
    return edg_prune_element(co, pb, hr.rs, verts, tri, elmark, theta, ws.data(), factors);
}
// the alpha pencil with the rule of order 2p + bonus (the EDG integrators pass bonus_intorder, reference :129-134)
double hc_alpha_bonus(const ehdg_problem* ep, const double* verts, const int* tri, int volmask, int exclude_const, int* status) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hb(2 * pb.p + pb.bonus);
    std::vector<double> ws(alpha_ws_doubles(pb, hb.rs.nq, hb.rs.nqf));
    Coop co{0, 1};
    *status = 0;
    return alpha_general(co, pb, hb.rs, verts, tri, volmask, exclude_const, ws.data(), status);
}

// Whole DG matrix in element-major block rows through edg_block_neighbours / edg_fill_row: rowptr [n_rows + 1], colind, vals
// (caller sizes them from hc_edg_csr_sizes), rhs [n_rows]
void hc_edg_csr_sizes(const ehdg_problem* ep, int ne, const int* el2facet, const int* facet2el, const unsigned char* elem_active,
                      long long* n_rows, long long* nnz) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    std::vector<int> cnt(ne);
    for (int el = 0; el < ne; ++el) cnt[el] = __builtin_popcount(edg_dofmask(pb.n_p, elem_active ? elem_active[el] : 0));
    long long r = 0, z = 0;
    for (int el = 0; el < ne; ++el) {
        int nbr[4], edge[4];
        const int n = edg_block_neighbours(el2facet, facet2el, el, nbr, edge);
        int len = 0;
        for (int t = 0; t < n; ++t) len += cnt[nbr[t]];
        r += cnt[el];
        z += (long long)cnt[el] * len;
    }
    *n_rows = r;
    *nnz = z;
}
void hc_edg_csr(const ehdg_problem* ep, int ne, const double* verts, const int* tris, const int* el2facet, const int* facet2el,
                const unsigned char* elem_active, const double* alpha_stab, long long* rowptr, int* colind, double* vals, double* rhs) {
    Problem pb = make_problem(*ep);
    pb.vol_drop_tol = 0.0;      // host checks compare block by block with the oracle: reference-literal elimination
    HostRules hr(2 * pb.p + pb.bonus);
    std::vector<double> ws(edg_ws_doubles(pb, hr.rs.nq, hr.rs.nqf));
    const int nv = pb.nvmax;
    std::vector<double> D(nv * nv), O(3 * nv * nv), F(nv);
    std::vector<int> row_start(ne + 1, 0);
    for (int el = 0; el < ne; ++el)
        row_start[el + 1] = row_start[el] + __builtin_popcount(edg_dofmask(pb.n_p, elem_active ? elem_active[el] : 0));
    Coop co{0, 1};
    long long pos = 0;
    int r = 0;
    for (int el = 0; el < ne; ++el) {
        edg_element_blocks(co, pb, hr.rs, verts, tris, el2facet, facet2el, elem_active, alpha_stab, el, ws.data(), D.data(), O.data(), F.data());
        int nbr[4], edge[4];
        const int n = edg_block_neighbours(el2facet, facet2el, el, nbr, edge);
        const unsigned m = edg_dofmask(pb.n_p, elem_active ? elem_active[el] : 0);
        for (int k = 0; k < nv; ++k) {
            if (!((m >> k) & 1u)) continue;
            rowptr[r] = pos;
            rhs[r] = F[k];
            const int cnt = edg_fill_row(nv, pb.n_p, elem_active, n, nbr, edge, k, D.data(), O.data(), vals + pos);
            int c = 0;
            for (int t = 0; t < n; ++t)
                for (int l = row_start[nbr[t]]; l < row_start[nbr[t] + 1]; ++l) colind[pos + c++] = l;
            pos += cnt;
            ++r;
        }
    }
    rowptr[r] = pos;
}
}
