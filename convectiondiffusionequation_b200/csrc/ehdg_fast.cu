// Polynomial-path kernels: persistent CTAs, reference tensors staged once per CTA in shared
// memory, one LG-lane group per element with its matrix rows in registers (ehdg_fast.h).
#include "ehdg_fast.h"
#include "ehdg_internal.h"

namespace ehdg {

constexpr int FAST_THREADS = 128;
// Threads per CTA at high order.  The reference tensors are staged once per CTA (p = 6: 99.5 KB for the assemble kernel, 73.5 KB
// for the alpha kernel): more threads per CTA are more element groups sharing one copy of the tables, i.e. more warps per SM.
// A/B on B200, 1M elements, p = 6 (profiles/r02_ab_p6_threads*.jsonl): assemble 128 / 256 / 384 / 512 threads 51.6 / 31.0 / 21.9 /
// 21.2 ms (512 spills), alpha 128 / 256 / 384 threads 40.3 / 40.9 / 33.8 ms; p = 5: assemble 128 / 192 / 256 threads 13.4 / 13.3 /
// 13.8 ms (stays at 128).
#ifndef EHDG_ASM_THREADS_P5
#define EHDG_ASM_THREADS_P5 128
#endif
#ifndef EHDG_ASM_MINBLOCKS_P5
#define EHDG_ASM_MINBLOCKS_P5 2
#endif
#ifndef EHDG_ASM_THREADS_P6
#define EHDG_ASM_THREADS_P6 384
#endif
#ifndef EHDG_ALPHA_THREADS_P6
#define EHDG_ALPHA_THREADS_P6 384
#endif
#ifndef EHDG_ALPHA_MINBLOCKS_P6
#define EHDG_ALPHA_MINBLOCKS_P6 1
#endif
template <int P>
constexpr int asm_threads() { return P == 6 ? EHDG_ASM_THREADS_P6 : (P == 5 ? EHDG_ASM_THREADS_P5 : FAST_THREADS); }
template <int P>
constexpr int alpha_threads() { return P == 6 ? EHDG_ALPHA_THREADS_P6 : FAST_THREADS; }
#ifndef EHDG_ASM_MINBLOCKS
#define EHDG_ASM_MINBLOCKS (EHDG_ASM_R == 2 ? 2 : 4)     // 128 registers without spills since A_FV is built after the elimination
#endif
#ifndef EHDG_ALPHA_MINBLOCKS
#define EHDG_ALPHA_MINBLOCKS (EHDG_ALPHA_R == 2 ? 3 : 4)
#endif

// lane 0 of the warp draws `step` consecutive work items, all lanes get the first one
__device__ __forceinline__ int next_work(int* counter, int step) {
    int b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(counter, step);
    return __shfl_sync(0xffffffffu, b, 0);
}

template <int P>
__global__ void __launch_bounds__(asm_threads<P>(), (P <= 4 ? EHDG_ASM_MINBLOCKS : (P == 5 ? EHDG_ASM_MINBLOCKS_P5 : 1)))
fast_assemble_kernel(Problem pb, LayerConst cx, LayerConst cy, RuleSet rs, const double* __restrict__ tab_g, int tab_doubles,
                     const double* __restrict__ phi_rhs, int n, const int32_t* __restrict__ list,
                     const double* __restrict__ verts, const int32_t* __restrict__ tris,
                     const double* __restrict__ lam, double* __restrict__ S, double* __restrict__ g,
                     double* __restrict__ W, double* __restrict__ z, int32_t* status, CsrOut co, int* __restrict__ work) {
    typedef FastDims<P> D;
    constexpr int TH = asm_threads<P>();
    constexpr int LG = D::LG, GPC = TH / LG, GPW = 32 / LG;
    extern __shared__ double smem[];
    for (int i = threadIdx.x; i < tab_doubles; i += TH) smem[i] = tab_g[i];
    __syncthreads();
    const int gid = threadIdx.x / LG;
    const int scratch = fast_asm_scratch_doubles<P>(rs.nq);
    Group<LG> gc{(int)(threadIdx.x % LG), smem + ((tab_doubles + 1) & ~1) + gid * scratch};
    // cx, cy: constants of the two layer functions of `coeff`, formed once on the host; as kernel parameters they are
    // constant-bank operands and cost no registers across the elimination
    // work: every warp draws its next GPW elements from a global counter (null: static grid-stride partition).  With the
    // quadrature-path kernel running beside this one on another stream, CTAs of this kernel become resident at different
    // times; drawn work keeps them all busy to the end, a static share would make the late ones the tail.
    const int gw = gid % GPW;
    for (int base = work ? next_work(work, GPW) : blockIdx.x * GPC; base < n;
         base = work ? next_work(work, GPW) : base + gridDim.x * GPC) {
        const int mine = work ? base + gw : base + gid;
        const bool live = mine < n;
        const int idx = live ? mine : n - 1;
        const int el = list[idx];
        const int st = fast_assemble_element<P>(gc, pb, rs, smem, phi_rhs, verts, tris + 3 * el, lam[el], live,
                                                S + (size_t)idx * D::NF * D::NF, g + (size_t)idx * D::NF,
                                                W + (size_t)idx * D::NP * D::NF, z + (size_t)idx * D::NP, cx, cy,
                                                co.em ? co.em + el : nullptr, &co);
        if (st && live && status) atomicOr(status + el, st);
    }
}

template <int P>
__global__ void __launch_bounds__(alpha_threads<P>(), (P <= 4 ? EHDG_ALPHA_MINBLOCKS : (P == 6 ? EHDG_ALPHA_MINBLOCKS_P6 : 2)))
fast_alpha_kernel(const double* __restrict__ tab_g, int tab_doubles, int n, const int32_t* __restrict__ list,
                  const uint8_t* __restrict__ skip_if_set, const double* __restrict__ verts, const int32_t* __restrict__ tris,
                  double* __restrict__ lam, int32_t* status, int* __restrict__ work) {
    typedef AlphaDims<P> D;
    constexpr int TH = alpha_threads<P>();
    constexpr int LG = D::LG, GPC = TH / LG;
    extern __shared__ double smem[];
    for (int i = threadIdx.x; i < tab_doubles; i += TH) smem[i] = tab_g[i];
    __syncthreads();
    // interleaved groups (ehdg_fast.h, Group::il): lane l of group g of a warp is thread l * GPW + g
    constexpr int GPW = 32 / LG;
    const int gid = (threadIdx.x / 32) * GPW + (threadIdx.x % GPW);
    Group<LG> gc{(int)((threadIdx.x & 31) / GPW), smem + ((tab_doubles + 1) & ~1) + gid * fast_alpha_scratch_doubles<P>(), GPW};
    const int gw = threadIdx.x % GPW;
    for (int base = work ? next_work(work, GPW) : blockIdx.x * GPC; base < n;
         base = work ? next_work(work, GPW) : base + gridDim.x * GPC) {
        const int mine = work ? base + gw : base + gid;
        const bool live = mine < n;
        const int idx = live ? mine : n - 1;
        const int el = list[idx];
        int st = 0;
        const double l = fast_alpha_element<P>(gc, smem, verts, tris + 3 * el, &st);
        if (live && gc.lane == 0 && !(skip_if_set && skip_if_set[el])) {
            lam[el] = l;
            if (status) status[el] = st;
        }
    }
}

template <int P>
static int launch_alpha(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const int32_t* list, int n, const uint8_t* skip_if_set,
                        double* lam, int32_t* status, cudaStream_t st) {
    typedef AlphaDims<P> D;
    constexpr int TH = alpha_threads<P>();
    constexpr int GPC = TH / D::LG;
    const int tabd = c->tens.alpha_doubles;
    const size_t smem = (size_t)(((tabd + 1) & ~1) + GPC * fast_alpha_scratch_doubles<P>()) * 8;
    EHDG_CUDA(cudaFuncSetAttribute(fast_alpha_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    EHDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast_alpha_kernel<P>, TH, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t need = ((int64_t)n + GPC - 1) / GPC;
    const int64_t cap = (int64_t)c->sm_count * per_sm;
    const int grid = (int)(need < cap ? need : cap);
    const bool timed = c->time_kernels && list == p->list_poly;
    if (timed) EHDG_CUDA(cudaEventRecord(c->ev_k[0], st));
    int* work = nullptr;
    if ((c->dynamic_work & 2) && list == p->list_poly) {
        work = c->work_counters;
        EHDG_CUDA(cudaMemsetAsync(work, 0, sizeof(int), st));
    }
    fast_alpha_kernel<P><<<grid, TH, smem, st>>>(c->tens.alpha_tab, tabd, n, list, skip_if_set, m->verts, m->tris, lam, status, work);
    if (timed) EHDG_CUDA(cudaEventRecord(c->ev_k[1], st));
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

template <int P>
static int launch_assemble(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const double* lam, double* S, double* g,
                           double* W, double* z, int32_t* status, cudaStream_t st, const CsrOut& co) {
    typedef FastDims<P> D;
    constexpr int TH = asm_threads<P>();
    constexpr int GPC = TH / D::LG;
    const int tabd = c->tens.asm_doubles;
    const size_t smem = (size_t)(((tabd + 1) & ~1) + GPC * fast_asm_scratch_doubles<P>(c->rb.rs.nq)) * 8;
    EHDG_CUDA(cudaFuncSetAttribute(fast_assemble_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    EHDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast_assemble_kernel<P>, TH, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t need = (p->n_poly + GPC - 1) / GPC;
    const int64_t cap = (int64_t)c->sm_count * per_sm;
    const int grid = (int)(need < cap ? need : cap);
    int* work = nullptr;
    if (c->dynamic_work & 1) {
        work = c->work_counters + 1;
        EHDG_CUDA(cudaMemsetAsync(work, 0, sizeof(int), st));
    }
    if (c->time_kernels) EHDG_CUDA(cudaEventRecord(c->ev_k[2], st));
    fast_assemble_kernel<P><<<grid, TH, smem, st>>>(c->pb, layer_const(c->pb.coeff.k[0]), layer_const(c->pb.coeff.k[1]), c->rb.rs, c->tens.asm_tab, tabd, c->tens.rhs_phi, (int)p->n_poly,
                                                             p->list_poly, m->verts, m->tris, lam, S, g, W, z, status, co, work);
    if (c->time_kernels) EHDG_CUDA(cudaEventRecord(c->ev_k[3], st));
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int fast_build_tensors(ehdg_ctx* c) {
    std::vector<double> asm_tab, alpha_tab, phi;
    const int order = 2 * c->pb.p + c->pb.bonus;
    build_tensors(c->pb.p, order, asm_tab, alpha_tab);
    build_rhs_phi(c->pb.p, order, phi);
    c->tens.asm_doubles = (int)asm_tab.size();
    c->tens.alpha_doubles = (int)alpha_tab.size();
    EHDG_CUDA(cudaMalloc(&c->tens.asm_tab, asm_tab.size() * 8));
    EHDG_CUDA(cudaMalloc(&c->tens.alpha_tab, alpha_tab.size() * 8));
    EHDG_CUDA(cudaMalloc(&c->tens.rhs_phi, phi.size() * 8));
    EHDG_CUDA(cudaMemcpy(c->tens.asm_tab, asm_tab.data(), asm_tab.size() * 8, cudaMemcpyHostToDevice));
    EHDG_CUDA(cudaMemcpy(c->tens.alpha_tab, alpha_tab.data(), alpha_tab.size() * 8, cudaMemcpyHostToDevice));
    EHDG_CUDA(cudaMemcpy(c->tens.rhs_phi, phi.data(), phi.size() * 8, cudaMemcpyHostToDevice));
    return EHDG_OK;
}

void fast_free_tensors(ehdg_ctx* c) {
    cudaFree(c->tens.asm_tab);
    cudaFree(c->tens.alpha_tab);
    cudaFree(c->tens.rhs_phi);
}

#define EHDG_DISPATCH_P(p, CALL)                 \
    switch (p) {                                  \
        case 1: return CALL(1);                   \
        case 2: return CALL(2);                   \
        case 3: return CALL(3);                   \
        case 4: return CALL(4);                   \
        case 5: return CALL(5);                   \
        case 6: return CALL(6);                   \
        default: set_error("unsupported order %d", p); return EHDG_ERR_ARG; \
    }

int fast_alpha(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const int32_t* list, int n, const uint8_t* skip_if_set, double* lam,
               int32_t* status, cudaStream_t st) {
    if (n <= 0) return EHDG_OK;
#define CALL(PP) launch_alpha<PP>(c, p, m, list, n, skip_if_set, lam, status, st)
    EHDG_DISPATCH_P(c->pb.p, CALL)
#undef CALL
}

int fast_assemble(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const double* lam, double* S, double* g, double* W,
                  double* z, int32_t* status, cudaStream_t st, const CsrOut* csr_out) {
    CsrOut co{};
    if (csr_out) co = *csr_out;
#define CALL(PP) launch_assemble<PP>(c, p, m, lam, S, g, W, z, status, st, co)
    EHDG_DISPATCH_P(c->pb.p, CALL)
#undef CALL
}

// FMA = 2 flops; counts the arithmetic the polynomial path executes on the matrix data of one
// element (geometry, coefficient evaluation and pivot search excluded), see DESIGN.md.
int64_t fast_flops_assemble(const ehdg_ctx* c) {
    const int64_t p = c->pb.p, NP = c->pb.n_p, NFL = p + 1, NF = 3 * NFL, NC = NP + NF + 1, nq = c->rb.rs.nq;
    int64_t f = 0;
    f += 2 * 14 * NP * NP;          // A_VV from 5 volume + 9 edge tensors
    f += 2 * 2 * 3 * NP * NF;       // A_VF and A_FV from 3 tensors each
    f += 2 * nq * NP;               // f_V
    for (int64_t k = 0; k < NP; ++k) f += 2 * NP * (NC - 1 - k);   // Gauss-Jordan on [A_VV | A_VF f_V]
    f += 2 * NF * NP * (NF + 1);    // Schur complement and g
    return f;
}

int64_t fast_flops_alpha(const ehdg_ctx* c) {
    // tensors, Cholesky, two triangular solves, Householder, Sturm multisection
    const int64_t NP = c->pb.n_p, NF = 3 * (c->pb.p + 1), N = NP - 1;
    int64_t f = 2 * 12 * N * N;
    for (int64_t k = 0; k < N; ++k) f += 2 * (N - 1 - k) * (N - 1 - k);
    for (int64_t j = 0; j < N; ++j) f += 2 * 2 * (N - 1 - j) * N;
    for (int64_t k = 0; k + 2 < N; ++k) f += 6 * (N - 1 - k) * (N - 1 - k);
    const int64_t M = NP > NF ? NP : NF;
    const bool two = c->pb.p <= EHDG_ALPHA_R2_MAXP && EHDG_ALPHA_R == 2;
    const int64_t LG = two ? (M <= 8 ? 4 : (M <= 16 ? 8 : 16)) : (M <= 8 ? 8 : (M <= 16 ? 16 : 32));
    const int64_t rounds = LG >= 32 ? 10 : (LG == 16 ? 12 : (LG == 8 ? 16 : 21));
    f += rounds * LG * N * 5;
    return f;
}

int64_t fast_flops_per_element(const ehdg_ctx* c) { return fast_flops_alpha(c) + fast_flops_assemble(c); }

}  // namespace ehdg
