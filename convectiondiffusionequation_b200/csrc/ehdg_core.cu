// libehdg core: context, marking (a1), plan, pruning (a3) and the general quadrature kernels
// for alpha (a4) and assembly + condensation (a5-a7).  See include/ehdg.h for the contract.
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ehdg_dist.h"
#include "ehdg_internal.h"
#include "ehdg_rules.h"

namespace ehdg {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

static int upload_rules(int order, DeviceRules& dr) {
    TrigRule tr = trig_rule(order);
    SegRule sr = seg_rule(order);
    const int nq = (int)tr.w.size(), nqf = (int)sr.w.size();
    std::vector<double> h(3 * nq + 2 * nqf);
    memcpy(h.data(), tr.x.data(), nq * 8);
    memcpy(h.data() + nq, tr.y.data(), nq * 8);
    memcpy(h.data() + 2 * nq, tr.w.data(), nq * 8);
    memcpy(h.data() + 3 * nq, sr.t.data(), nqf * 8);
    memcpy(h.data() + 3 * nq + nqf, sr.w.data(), nqf * 8);
    EHDG_CUDA(cudaMalloc(&dr.buf, h.size() * 8));
    EHDG_CUDA(cudaMemcpy(dr.buf, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    dr.rs.nq = nq;
    dr.rs.nqf = nqf;
    dr.rs.vx = dr.buf;
    dr.rs.vy = dr.buf + nq;
    dr.rs.vw = dr.buf + 2 * nq;
    dr.rs.st = dr.buf + 3 * nq;
    dr.rs.sw = dr.buf + 3 * nq + nqf;
    return EHDG_OK;
}

// -------------------------------------------------------------------------------------------
// a1 marking
// -------------------------------------------------------------------------------------------
__global__ void mark_elements_kernel(int ne, int nv, int nenr, const int32_t* __restrict__ tris,
                                     const uint8_t* __restrict__ vflags, uint8_t* __restrict__ elem_mask) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= ne) return;
    int m = 0;
    for (int k = 0; k < nenr; ++k) {
        const uint8_t* f = vflags + (size_t)k * nv;
        if (f[tris[3 * el]] | f[tris[3 * el + 1]] | f[tris[3 * el + 2]]) m |= 1 << k;   // any vertex
    }
    elem_mask[el] = (uint8_t)m;
}

__global__ void mark_facets_kernel(int nf, const int32_t* __restrict__ facet2el,
                                   const uint8_t* __restrict__ elem_mask, uint8_t* __restrict__ facet_mask) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    int m = 0;
    const int e0 = facet2el[2 * f], e1 = facet2el[2 * f + 1];
    if (e0 >= 0) m |= elem_mask[e0];
    if (e1 >= 0) m |= elem_mask[e1];
    facet_mask[f] = (uint8_t)m;                       // any adjacent element marked
}

// -------------------------------------------------------------------------------------------
// partitioned assembly: elements that touch a facet range (own + ghost), elements owned by it
// -------------------------------------------------------------------------------------------
__global__ void select_elements_kernel(int ne, int nf, int f_lo, int f_hi, const int32_t* __restrict__ el2facet,
                                       uint8_t* __restrict__ select, uint8_t* __restrict__ owned) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= ne) return;
    const int a = el2facet[3 * el], b = el2facet[3 * el + 1], c = el2facet[3 * el + 2];
    const bool ia = a >= f_lo && a < f_hi, ib = b >= f_lo && b < f_hi, ic = c >= f_lo && c < f_hi;
    select[el] = (ia || ib || ic) ? 1 : 0;
    if (owned) {
        const int m = min(a, min(b, c));          // the element belongs to the rank that owns its lowest-numbered facet
        owned[el] = (m >= f_lo && m < f_hi) ? 1 : 0;
    }
}

// -------------------------------------------------------------------------------------------
// plan
// -------------------------------------------------------------------------------------------
__global__ void plan_flags_kernel(int ne, const int32_t* __restrict__ el2facet,
                                  const uint8_t* __restrict__ facet_mask, int force, const uint8_t* __restrict__ select,
                                  int32_t* __restrict__ flag_gen, int32_t* __restrict__ flag_poly) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= ne) return;
    int g = force;
    if (facet_mask)
        g |= facet_mask[el2facet[3 * el]] | facet_mask[el2facet[3 * el + 1]] | facet_mask[el2facet[3 * el + 2]];
    const int on = select ? (select[el] != 0) : 1;      // elements outside the selection get no slot on either path
    flag_gen[el] = (on && g) ? 1 : 0;
    flag_poly[el] = (on && !g) ? 1 : 0;
}

__global__ void plan_scatter_kernel(int ne, const int32_t* __restrict__ flag_gen, const int32_t* __restrict__ flag_poly,
                                    const int32_t* __restrict__ scan_gen, const int32_t* __restrict__ scan_poly,
                                    int32_t* __restrict__ eslot, int32_t* __restrict__ list_poly,
                                    int32_t* __restrict__ list_gen) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= ne) return;
    if (flag_gen[el]) {
        const int gi = scan_gen[el];
        eslot[el] = EHDG_SLOT_GEN | gi;
        list_gen[gi] = el;
    } else if (flag_poly[el]) {
        const int pi = scan_poly[el];
        eslot[el] = pi;
        list_poly[pi] = el;
    } else {
        eslot[el] = EHDG_SLOT_NONE;
    }
}

// -------------------------------------------------------------------------------------------
// general kernels: one CTA of GEN_THREADS threads per element, workspace in dynamic smem
// -------------------------------------------------------------------------------------------
// A/B on B200, 4M headline mesh, 12 183 elements on this path (profiles/r02_ab_gen_threads*.txt), step in ms:
// 256 threads 41.5, 128: 39.2, 96: 39.0, 64: 38.4, 32: 38.2.  The kernels need 128 registers per thread and 26 KB of shared
// memory per element (p = 4): smaller CTAs mean more elements resident per SM and cheaper barriers, and the per-element
// phases have too little parallelism for more than two warps anyway.  64 keeps 8 warps per SM at p = 6 as well (50 KB).
#ifndef EHDG_GEN_THREADS
#define EHDG_GEN_THREADS 64
#endif
constexpr int GEN_THREADS = EHDG_GEN_THREADS;

__device__ __forceinline__ int fac_bits(const uint8_t* fm, const int32_t* el2facet, int el, int e) {
    return fm ? (int)fm[el2facet[3 * el + e]] : 0;
}

__global__ void __launch_bounds__(GEN_THREADS)
prune_kernel(Problem pb, RuleSet rs, int n_gen, const int32_t* __restrict__ list_gen,
             const double* __restrict__ verts, const int32_t* __restrict__ tris,
             const int32_t* __restrict__ el2facet, const uint8_t* __restrict__ elem_mask,
             const uint8_t* __restrict__ facet_mask, double threshold, uint8_t* elem_active,
             uint8_t* facet_active, double* factors) {
    extern __shared__ double ws[];
    const int gi = blockIdx.x;
    if (gi >= n_gen) return;
    const int el = list_gen[gi];
    const int elmark = elem_mask[el];
    if (elmark == 0) return;                      // only marked elements are visited (:285)
    Coop co{(int)threadIdx.x, GEN_THREADS};
    int facmask[3];
    for (int e = 0; e < 3; ++e) facmask[e] = fac_bits(facet_mask, el2facet, el, e);
    int vdrop, fdrop[3];
    double fac[4 * EHDG_MAX_ENRICH];
    prune_general(co, pb, rs, verts, tris + 3 * el, elmark, facmask, threshold, ws, &vdrop, fdrop, fac);
    if (threadIdx.x == 0) {
        if (vdrop) elem_active[el] = (uint8_t)(elmark & ~vdrop);
        for (int e = 0; e < 3; ++e)
            if (fdrop[e]) {
                const int f = el2facet[3 * el + e];
                unsigned int* word = reinterpret_cast<unsigned int*>(facet_active) + (f >> 2);
                atomicAnd(word, ~((unsigned int)fdrop[e] << (8 * (f & 3))));
            }
        if (factors)
            for (int i = 0; i < 4 * pb.nenr; ++i) factors[(size_t)el * 4 * pb.nenr + i] = fac[i];
    }
}

__global__ void prune_init_kernel(int ne, int nf, int nenr, const uint8_t* __restrict__ elem_mask,
                                  const uint8_t* __restrict__ facet_mask, const int32_t* __restrict__ facet2el,
                                  uint8_t* elem_active, uint8_t* facet_active, double* factors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ne) {
        elem_active[i] = elem_mask[i];
        if (factors)
            for (int k = 0; k < 4 * nenr; ++k) factors[(size_t)i * 4 * nenr + k] = -1.0;
    }
    if (i < nf) facet_active[i] = (facet2el[2 * i + 1] < 0) ? 0 : facet_mask[i];   // dirichlet=".*"
}

__global__ void __launch_bounds__(GEN_THREADS)
alpha_general_kernel(Problem pb, RuleSet rs0, int n_gen, const int32_t* __restrict__ list_gen,
                     const double* __restrict__ verts, const int32_t* __restrict__ tris,
                     const uint8_t* __restrict__ elem_active, int skip_unenriched, double* __restrict__ lam, int32_t* status) {
    extern __shared__ double ws[];
    const int gi = blockIdx.x;
    if (gi >= n_gen) return;
    const int el = list_gen[gi];
    Coop co{(int)threadIdx.x, GEN_THREADS};
    const int volmask = elem_active ? elem_active[el] : 0;
    if (skip_unenriched && volmask == 0) return;      // polynomial alpha: done by the tensor kernel (CTA-uniform exit)
    int st = 0;
    const double l = alpha_general(co, pb, rs0, verts, tris + 3 * el, volmask, el == 0, ws, &st);
    st = block_or_bits(st);
    if (threadIdx.x == 0) {
        lam[el] = l;
        if (status) status[el] = st;
    }
}

__global__ void __launch_bounds__(GEN_THREADS)
assemble_general_kernel(Problem pb, RuleSet rs, int n_gen, const int32_t* __restrict__ list_gen,
                        const double* __restrict__ verts, const int32_t* __restrict__ tris,
                        const int32_t* __restrict__ el2facet, const double* __restrict__ lam,
                        const uint8_t* __restrict__ elem_active, const uint8_t* __restrict__ facet_mask,
                        double* __restrict__ S, double* __restrict__ g, double* __restrict__ W,
                        double* __restrict__ z, int32_t* status, double* Afull) {
    extern __shared__ double ws[];
    const int gi = blockIdx.x;
    if (gi >= n_gen) return;
    const int el = list_gen[gi];
    Coop co{(int)threadIdx.x, GEN_THREADS};
    const int volmask = elem_active ? elem_active[el] : 0;
    int facmask[3];
    for (int e = 0; e < 3; ++e) facmask[e] = fac_bits(facet_mask, el2facet, el, e);
    const int nv = pb.nvmax, nf = pb.nfmax;
    const size_t afull_sz = (size_t)nv * (nv + nf + 1) + (size_t)nf * nv + (size_t)nf * nf;
    int st = assemble_condense_general(co, pb, rs, verts, tris + 3 * el, lam[el], volmask, facmask, ws,
                                       S + (size_t)gi * nf * nf, g + (size_t)gi * nf, W + (size_t)gi * nv * nf,
                                       z + (size_t)gi * nv, Afull ? Afull + (size_t)gi * afull_sz : nullptr);
    st = block_or_bits(st);
    if (threadIdx.x == 0 && status) status[el] |= st;
}

// the quadrature path on the plan's general elements; Sg .. zg point at THEIR part of the element stores
int general_assemble(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const double* lam, const uint8_t* elem_active,
                     const uint8_t* facet_mask, double* Sg, double* gg, double* Wg, double* zg, int32_t* status, double* Afull,
                     cudaStream_t st) {
    if (p->n_gen <= 0) return EHDG_OK;
    assemble_general_kernel<<<(unsigned)p->n_gen, GEN_THREADS, c->ws_doubles * 8, st>>>(
        c->pb, c->rb.rs, (int)p->n_gen, p->list_gen, m->verts, m->tris, m->el2facet, lam,
        c->pb.nenr ? elem_active : nullptr, c->pb.nenr ? facet_mask : nullptr, Sg, gg, Wg, zg, status, Afull);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

}  // namespace ehdg

using namespace ehdg;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* ehdg_version(void) { return "libehdg 0.1 (sm_100a)"; }
const char* ehdg_last_error(void) { return g_err.c_str(); }

int ehdg_create(const ehdg_problem* prob, ehdg_ctx** out) {
    if (!prob || !out) { set_error("ehdg_create: null argument"); return EHDG_ERR_ARG; }
    if (!problem_valid(*prob)) { set_error("ehdg_create: invalid problem (order 1..6, n_enrich 0..2, eps > 0)"); return EHDG_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("ehdg_create: no CUDA device (libehdg has no CPU path)");
        return EHDG_ERR_CUDA;
    }
    ehdg_ctx* c = new ehdg_ctx();
    // an error below releases whatever the context already owns
    struct Guard { ehdg_ctx* c; bool ok = false; ~Guard() { if (!ok) ehdg_destroy(c); } } guard{c};
    c->ep = *prob;
    c->pb = make_problem(*prob);
    EHDG_CUDA(cudaGetDevice(&c->device));
    EHDG_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    {
        // patterns and solver workspaces come from the stream-ordered pool: freed blocks stay cached in it, so a rebuilt
        // pattern / a second solve does not pay cudaMalloc again (milliseconds per call at these sizes)
        cudaMemPool_t pool;
        EHDG_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->device));
        uint64_t keep = ~(uint64_t)0;
        EHDG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    int rc;
    if ((rc = upload_rules(2 * c->pb.p + c->pb.bonus, c->rb))) return rc;
    if ((rc = upload_rules(2 * c->pb.p, c->r0))) return rc;
    if ((rc = upload_rules(50 + c->pb.bonus, c->rerr))) return rc;
    if ((rc = upload_rules(5 + c->pb.bonus, c->redg))) return rc;
    c->ws_doubles = general_ws_doubles(c->pb, c->rb.rs.nq, c->rb.rs.nqf);
    const int a_ws = alpha_ws_doubles(c->pb, c->r0.rs.nq, c->r0.rs.nqf);
    const int a_ws_b = alpha_ws_doubles(c->pb, c->rb.rs.nq, c->rb.rs.nqf);      // EDG: the alpha integrators take the bonus rule
    const int smem = ((c->ws_doubles > a_ws ? c->ws_doubles : a_ws) > a_ws_b ? (c->ws_doubles > a_ws ? c->ws_doubles : a_ws) : a_ws_b) * 8;
    EHDG_CUDA(cudaFuncSetAttribute(prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EHDG_CUDA(cudaFuncSetAttribute(alpha_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EHDG_CUDA(cudaFuncSetAttribute(assemble_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if ((rc = fast_build_tensors(c))) return rc;
    {
        // the quadrature-path kernels (O(sqrt(ne)) elements, latency-bound: one CTA per element) can run on a high-priority side
        // stream beside the polynomial-path kernels, which then draw their work from a counter (see fast_assemble_kernel)
        int lo = 0, hi = 0;
        EHDG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        EHDG_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
        EHDG_CUDA(cudaMalloc(&c->work_counters, 2 * sizeof(int)));
        const char* e = getenv("EHDG_OVERLAP");
        c->overlap_general = (e && e[0] == '1') ? 1 : 0;
        // A/B on B200, 4M elements (profiles/r02_ab_overlap.txt): drawn work makes the assemble kernel 1.0 ms faster (18.57 ->
        // 17.57 ms: no tail) and the alpha kernel 0.4 ms slower (uniform work, the atomics only cost); the overlap itself buys
        // nothing on top (both polynomial kernels fill the register file, the quadrature CTAs do not become co-resident).
        // Default: assemble draws (bit 0), alpha keeps its static partition (bit 1 off), no overlap.
        const char* d = getenv("EHDG_DYNAMIC");
        c->dynamic_work = d ? atoi(d) : (c->overlap_general ? 3 : 1);
    }
    EHDG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    EHDG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    guard.ok = true;
    *out = c;
    return EHDG_OK;
}

int ehdg_destroy(ehdg_ctx* c) {
    if (!c) return EHDG_OK;
    cudaFree(c->rb.buf);
    cudaFree(c->r0.buf);
    cudaFree(c->rerr.buf);
    cudaFree(c->redg.buf);
    for (int i = 0; i < 4; ++i) cudaFree(c->d_prog[i]);
    ehdg_comm_destroy(c);
    fast_free_tensors(c);
    cudaFree(c->scratch);
    cudaFree(c->work_counters);
    cudaFree(c->pool_eslot);
    cudaFree(c->pool_lists);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int i = 0; i < 4; ++i)
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    delete c;
    return EHDG_OK;
}

int ehdg_set_expression(ehdg_ctx* c, int which, const int32_t* ops, const double* args, int n) {
    if (!c || which < 0 || which > 3 || n < 0 || n > EHDG_EXPR_MAX || (n > 0 && (!ops || !args))) {
        set_error("ehdg_set_expression: bad argument (at most %d instructions)", EHDG_EXPR_MAX);
        return EHDG_ERR_ARG;
    }
    const ExprProg** slot[4] = {&c->pb.xcoeff, &c->pb.xexact, &c->pb.xexact_dx, &c->pb.xexact_dy};
    if (n == 0) { *slot[which] = nullptr; return EHDG_OK; }
    // validate the stack discipline on the host: the device interpreter does not check
    int sp = 0;
    for (int i = 0; i < n; ++i) {
        const int op = ops[i];
        const int pops = (op == EX_ADD || op == EX_SUB || op == EX_MUL || op == EX_DIV || op == EX_POW) ? 2
                         : (op == EX_CONST || op == EX_X || op == EX_Y) ? 0 : 1;
        if (op <= EX_END || op > EX_SIGN || sp < pops) { set_error("ehdg_set_expression: malformed program at instruction %d", i); return EHDG_ERR_ARG; }
        sp += 1 - pops;
        if (sp > EHDG_EXPR_STACK) { set_error("ehdg_set_expression: expression needs more than %d stack slots", EHDG_EXPR_STACK); return EHDG_ERR_ARG; }
    }
    if (sp != 1) { set_error("ehdg_set_expression: program leaves %d values on the stack", sp); return EHDG_ERR_ARG; }
    ExprProg h;
    memset(&h, 0, sizeof(h));
    h.n = n;
    for (int i = 0; i < n; ++i) { h.op[i] = ops[i]; h.arg[i] = args[i]; }
    if (!c->d_prog[which]) EHDG_CUDA(cudaMalloc(&c->d_prog[which], sizeof(ExprProg)));
    EHDG_CUDA(cudaMemcpy(c->d_prog[which], &h, sizeof(ExprProg), cudaMemcpyHostToDevice));
    *slot[which] = c->d_prog[which];
    return EHDG_OK;
}

int ehdg_edg_set_boundary_data(ehdg_ctx* c, int enable) {
    if (!c) { set_error("ehdg_edg_set_boundary_data: null context"); return EHDG_ERR_ARG; }
    c->pb.edg_bnd_data = enable ? 1 : 0;
    return EHDG_OK;
}

int ehdg_set_drop_tol(ehdg_ctx* c, double vol_drop_tol) {
    if (!c || !(vol_drop_tol >= 0.0)) { set_error("ehdg_set_drop_tol: bad argument"); return EHDG_ERR_ARG; }
    c->pb.vol_drop_tol = vol_drop_tol;
    return EHDG_OK;
}

int ehdg_kernel_timing(ehdg_ctx* c, int enable) {
    if (!c) { set_error("ehdg_kernel_timing: null argument"); return EHDG_ERR_ARG; }
    if (enable && !c->ev_k[0])
        for (int i = 0; i < 4; ++i) EHDG_CUDA(cudaEventCreate(&c->ev_k[i]));
    c->time_kernels = enable ? 1 : 0;
    return EHDG_OK;
}

int ehdg_kernel_times_ms(ehdg_ctx* c, float* alpha_ms, float* assemble_ms) {
    if (!c || !c->ev_k[0] || !alpha_ms || !assemble_ms) { set_error("ehdg_kernel_times_ms: timing not enabled"); return EHDG_ERR_ARG; }
    EHDG_CUDA(cudaEventSynchronize(c->ev_k[1]));
    EHDG_CUDA(cudaEventSynchronize(c->ev_k[3]));
    EHDG_CUDA(cudaEventElapsedTime(alpha_ms, c->ev_k[0], c->ev_k[1]));
    EHDG_CUDA(cudaEventElapsedTime(assemble_ms, c->ev_k[2], c->ev_k[3]));
    return EHDG_OK;
}

int ehdg_get_sizes(const ehdg_ctx* c, ehdg_sizes* o) {
    if (!c || !o) { set_error("ehdg_get_sizes: null argument"); return EHDG_ERR_ARG; }
    const Problem& pb = c->pb;
    o->n_p = pb.n_p;
    o->nfl = pb.nfl;
    o->nb = pb.nb;
    o->nvmax = pb.nvmax;
    o->nfmax = pb.nfmax;
    o->nf_poly = 3 * pb.nfl;
    o->nq = c->rb.rs.nq;
    o->nqf = c->rb.rs.nqf;
    o->nq0 = c->r0.rs.nq;
    o->nqf0 = c->r0.rs.nqf;
    o->nq_err = c->rerr.rs.nq;
    o->flops_poly_elem = fast_flops_per_element(c);
    o->flops_poly_alpha = fast_flops_alpha(c);
    o->flops_poly_assemble = fast_flops_assemble(c);
    // F_asm + F_cond of the quadrature algorithm on an unenriched element
    const int64_t nV = pb.n_p, ne_ = pb.nfl, nF = 3 * pb.nfl, nq = o->nq, nqf = o->nqf;
    o->flops_quad_elem = 4 * nV * nV * nq + 2 * nV * nq + 3 * nqf * (4 * nV * nV + 6 * nV * ne_ + 2 * ne_ * ne_)
                         + (2 * nV * nV * nV) / 3 + 2 * nV * nV * (nF + 1) + 2 * nF * nV * (nF + 1);
    return EHDG_OK;
}

int ehdg_mark(ehdg_ctx* c, const ehdg_mesh* m, const uint8_t* vert_flags, uint8_t* elem_mask, uint8_t* facet_mask,
              void* stream) {
    if (!c || !m || !elem_mask || !facet_mask) { set_error("ehdg_mark: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (m->ne > 0) mark_elements_kernel<<<(m->ne + 255) / 256, 256, 0, st>>>(m->ne, m->nv, c->pb.nenr, m->tris, vert_flags, elem_mask);
    EHDG_LAUNCH_CHECK();
    if (m->nf > 0) mark_facets_kernel<<<(m->nf + 255) / 256, 256, 0, st>>>(m->nf, m->facet2el, elem_mask, facet_mask);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

static int ctx_scratch(ehdg_ctx* c, size_t bytes, void** out) {
    if (bytes > c->scratch_bytes) {
        if (c->scratch) cudaFree(c->scratch);
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        EHDG_CUDA(cudaMalloc(&c->scratch, bytes));
        c->scratch_bytes = bytes;
    }
    *out = c->scratch;
    return EHDG_OK;
}

int ehdg_select_elements(ehdg_ctx* c, const ehdg_mesh* m, int32_t f_lo, int32_t f_hi, uint8_t* elem_select, uint8_t* elem_owned,
                         void* stream) {
    if (!c || !m || !elem_select) { set_error("ehdg_select_elements: null argument"); return EHDG_ERR_ARG; }
    if (m->ne > 0)
        select_elements_kernel<<<(m->ne + 255) / 256, 256, 0, (cudaStream_t)stream>>>(m->ne, m->nf, f_lo, f_hi, m->el2facet, elem_select, elem_owned);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int ehdg_plan_create_select(ehdg_ctx* c, const ehdg_mesh* m, const uint8_t* facet_mask, int force_general,
                            const uint8_t* elem_select, ehdg_plan** out, void* stream) {
    if (!c || !m || !out) { set_error("ehdg_plan_create: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    ehdg_plan* p = new ehdg_plan();
    p->ctx = c;
    p->forced = force_general ? 1 : 0;
    p->ne = m->ne;
    p->nf_poly = 3 * c->pb.nfl;
    p->nfmax = c->pb.nfmax;
    p->n_p = c->pb.n_p;
    p->nvmax = c->pb.nvmax;
    const int ne = m->ne;
    const size_t nint = (size_t)(ne + 1);
    size_t tmp_bytes = 0;
    EHDG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (int32_t*)nullptr, (int32_t*)nullptr, ne + 1, st));
    const size_t off_arr = (nint * 4 + 255) & ~(size_t)255, off_tmp = 4 * off_arr;
    void* base = nullptr;
    int rc = ctx_scratch(c, off_tmp + tmp_bytes + 256, &base);
    if (rc) return rc;
    int32_t* flag_gen = (int32_t*)base;
    int32_t* flag_poly = (int32_t*)((char*)base + off_arr);
    int32_t* scan_gen = (int32_t*)((char*)base + 2 * off_arr);
    int32_t* scan_poly = (int32_t*)((char*)base + 3 * off_arr);
    void* tmp = (char*)base + off_tmp;
    if (c->pool_eslot && c->pool_ne == ne) {             // reuse the buffers of the last destroyed plan
        p->eslot = c->pool_eslot;
        p->lists = c->pool_lists;
        c->pool_eslot = c->pool_lists = nullptr;
    } else {
        EHDG_CUDA(cudaMalloc(&p->eslot, sizeof(int32_t) * (size_t)(ne > 0 ? ne : 1)));
        EHDG_CUDA(cudaMalloc(&p->lists, sizeof(int32_t) * (size_t)(ne > 0 ? ne : 1)));
    }
    EHDG_CUDA(cudaMemsetAsync(flag_gen, 0, 2 * off_arr, st));
    if (ne > 0) plan_flags_kernel<<<(ne + 255) / 256, 256, 0, st>>>(ne, m->el2facet, c->pb.nenr ? facet_mask : nullptr, force_general ? 1 : 0,
                                                                     elem_select, flag_gen, flag_poly);
    EHDG_LAUNCH_CHECK();
    EHDG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag_gen, scan_gen, ne + 1, st));
    EHDG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag_poly, scan_poly, ne + 1, st));
    int32_t ngen = 0, npoly = 0;
    EHDG_CUDA(cudaMemcpyAsync(&ngen, scan_gen + ne, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaMemcpyAsync(&npoly, scan_poly + ne, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaStreamSynchronize(st));
    p->n_gen = ngen;
    p->n_poly = npoly;
    p->list_poly = p->lists;
    p->list_gen = p->lists + p->n_poly;
    if (ne > 0) plan_scatter_kernel<<<(ne + 255) / 256, 256, 0, st>>>(ne, flag_gen, flag_poly, scan_gen, scan_poly, p->eslot, p->list_poly, p->list_gen);
    EHDG_LAUNCH_CHECK();
    *out = p;
    return EHDG_OK;
}

int ehdg_plan_create(ehdg_ctx* c, const ehdg_mesh* m, const uint8_t* facet_mask, int force_general, ehdg_plan** out,
                     void* stream) {
    return ehdg_plan_create_select(c, m, facet_mask, force_general, nullptr, out, stream);
}

int ehdg_plan_destroy(ehdg_plan* p) {
    if (!p) return EHDG_OK;
    ehdg_ctx* c = p->ctx;
    if (c && !c->pool_eslot) {
        c->pool_eslot = p->eslot;
        c->pool_lists = p->lists;
        c->pool_ne = p->ne;
    } else {
        cudaFree(p->eslot);
        cudaFree(p->lists);
    }
    delete p;
    return EHDG_OK;
}

int ehdg_plan_get_sizes(const ehdg_plan* p, ehdg_plan_sizes* o) {
    if (!p || !o) { set_error("ehdg_plan_get_sizes: null argument"); return EHDG_ERR_ARG; }
    o->n_poly = p->n_poly;
    o->n_gen = p->n_gen;
    o->S_doubles = p->n_poly * p->nf_poly * p->nf_poly + p->n_gen * p->nfmax * p->nfmax;
    o->g_doubles = p->n_poly * p->nf_poly + p->n_gen * p->nfmax;
    o->W_doubles = p->n_poly * p->n_p * p->nf_poly + p->n_gen * p->nvmax * p->nfmax;
    o->z_doubles = p->n_poly * p->n_p + p->n_gen * p->nvmax;
    return EHDG_OK;
}

const int32_t* ehdg_plan_eslot(const ehdg_plan* p) { return p ? p->eslot : nullptr; }

int ehdg_prune(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_mask, const uint8_t* facet_mask,
               double threshold, uint8_t* elem_active, uint8_t* facet_active, double* factors, void* stream) {
    if (!c || !p || !m || !elem_mask || !facet_mask || !elem_active || !facet_active) {
        set_error("ehdg_prune: null argument");
        return EHDG_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = m->ne > m->nf ? m->ne : m->nf;
    if (n > 0) prune_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(m->ne, m->nf, c->pb.nenr, elem_mask, facet_mask, m->facet2el, elem_active, facet_active, factors);
    EHDG_LAUNCH_CHECK();
    if (c->pb.nenr > 0 && p->n_gen > 0) {
        prune_kernel<<<(unsigned)p->n_gen, GEN_THREADS, c->ws_doubles * 8, st>>>(c->pb, c->rb.rs, (int)p->n_gen, p->list_gen, m->verts, m->tris,
                                                                          m->el2facet, elem_mask, facet_mask, threshold, elem_active,
                                                                          facet_active, factors);
        EHDG_LAUNCH_CHECK();
    }
    return EHDG_OK;
}

static int alpha_impl(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_active, double* lam,
                      int32_t* status, void* stream, bool bonus_rule);

int ehdg_alpha(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_active, double* lam,
               int32_t* status, void* stream) {
    return alpha_impl(c, p, m, elem_active, lam, status, stream, false);
}

// EDG (reference :129-163): the same pencil, integrated with the bonus rule (the EDG integrators pass bonus_intorder);
// on polynomial elements both rules are exact, so the tensor kernel serves them unchanged
int ehdg_edg_alpha(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_active, double* lam,
                   int32_t* status, void* stream) {
    const int rc = alpha_impl(c, p, m, elem_active, lam, status, stream, true);
    if (rc) return rc;
    return ehdg_edg_alpha_carry(c, m->ne, lam, stream);
}

static int alpha_impl(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_active, double* lam,
                      int32_t* status, void* stream, bool bonus_rule) {
    if (!c || !p || !m || !lam) { set_error("ehdg_alpha: null argument"); return EHDG_ERR_ARG; }
    const RuleSet& rsa = bonus_rule ? c->rb.rs : c->r0.rs;
    cudaStream_t st = (cudaStream_t)stream;
    // running the O(sqrt(ne)) enriched elements on a side stream beside the bulk kernel was measured slower on B200
    // (the persistent bulk CTAs already fill every SM; +0.5 ms at 1M elements), so it is opt-in
    const bool fork = p->n_gen > 0 && p->n_poly > 0 && c->overlap_general;
    cudaStream_t sg = fork ? c->side : st;
    if (fork) {
        EHDG_CUDA(cudaEventRecord(c->ev_fork, st));
        EHDG_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    }
    if (p->n_gen > 0) {
        // the alpha integrators use the order-2p rule: a quarter of the assembly workspace, 3x the resident CTAs
        alpha_general_kernel<<<(unsigned)p->n_gen, GEN_THREADS, alpha_ws_doubles(c->pb, rsa.nq, rsa.nqf) * 8, sg>>>(c->pb, rsa, (int)p->n_gen, p->list_gen, m->verts,
                                                                                  m->tris, c->pb.nenr ? elem_active : nullptr, (c->pb.nenr && !p->forced) ? 1 : 0, lam, status);
        EHDG_LAUNCH_CHECK();
    }
    if (p->n_poly > 0) {
        int rc = fast_alpha(c, p, m, p->list_poly, (int)p->n_poly, nullptr, lam, status, st);
        if (rc) return rc;
    }
    if (p->n_gen > 0 && c->pb.nenr && !p->forced) {
        // facet neighbours of the marked elements carry no volume enrichment: their alpha is the polynomial one
        int rc = fast_alpha(c, p, m, p->list_gen, (int)p->n_gen, elem_active, lam, status, sg);
        if (rc) return rc;
    }
    if (fork) {
        EHDG_CUDA(cudaEventRecord(c->ev_join, c->side));
        EHDG_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
    }
    return EHDG_OK;
}

int ehdg_assemble_condense(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const double* lam,
                           const uint8_t* elem_active, const uint8_t* facet_mask, double* S, double* g, double* W,
                           double* z, int32_t* status, double* Afull, void* stream) {
    if (!c || !p || !m || !lam || !S || !g || !W || !z) { set_error("ehdg_assemble_condense: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const bool fork = p->n_gen > 0 && p->n_poly > 0 && c->overlap_general;
    cudaStream_t sg = fork ? c->side : st;
    if (fork) {
        EHDG_CUDA(cudaEventRecord(c->ev_fork, st));
        EHDG_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    }
    if (p->n_gen > 0) {
        const size_t nfp = p->nf_poly, np = p->n_p;
        int rc = general_assemble(c, p, m, lam, elem_active, facet_mask, S + (size_t)p->n_poly * nfp * nfp, g + (size_t)p->n_poly * nfp,
                                  W + (size_t)p->n_poly * np * nfp, z + (size_t)p->n_poly * np, status, Afull, sg);
        if (rc) return rc;
    }
    if (p->n_poly > 0) {
        int rc = fast_assemble(c, p, m, lam, S, g, W, z, status, st);
        if (rc) return rc;
    }
    if (fork) {
        EHDG_CUDA(cudaEventRecord(c->ev_join, c->side));
        EHDG_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
    }
    return EHDG_OK;
}

}  // extern "C"
