// Internal definitions shared by the libehdg translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/ehdg.h"
#include "ehdg_general.h"
#include "ehdg_problem.h"

namespace ehdg {

struct DistComm;
void set_error(const char* fmt, ...);

#define EHDG_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            ehdg::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return EHDG_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define EHDG_LAUNCH_CHECK() EHDG_CUDA(cudaGetLastError())

#ifdef __CUDACC__
// bitwise OR of a per-thread status word over the CTA (__syncthreads_or only says whether ANY bit is set anywhere)
__device__ __forceinline__ int block_or_bits(int v) {
    __shared__ int s_bits;
    if (threadIdx.x == 0) s_bits = 0;
    __syncthreads();
    if (v) atomicOr(&s_bits, v);
    __syncthreads();
    return s_bits;
}
#endif

struct DeviceRules {
    RuleSet rs;                 // device pointers
    double* buf = nullptr;
};

// reference tensors of the polynomial path (device), see ehdg_tensors.h
struct DeviceTensors {
    double* asm_tab = nullptr;     // assembly tensors
    double* alpha_tab = nullptr;   // alpha pencil tensors
    double* rhs_phi = nullptr;     // phi at the bonus-rule points, 6 vertex-order classes
    int asm_doubles = 0, alpha_doubles = 0;
};

}  // namespace ehdg

struct ehdg_ctx {
    ehdg_problem ep;
    ehdg::Problem pb;
    ehdg::DeviceRules rb, r0, rerr, redg;   // orders 2p+bonus, 2p, 50+bonus, 5+bonus (EDG error, reference :205-210)
    ehdg::DeviceTensors tens;
    int ws_doubles;                   // per-warp workspace of the general kernels
    int device;
    int sm_count;
    // grow-only scratch and a one-deep pool of plan buffers: cudaMalloc/cudaFree cost milliseconds, a plan
    // is rebuilt for every mesh upload in the end-to-end path
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    int32_t *pool_eslot = nullptr, *pool_lists = nullptr;
    int32_t pool_ne = 0;
    ehdg::DistComm* dist = nullptr;   // NCCL communicator + row partition of the distributed facet solve
    ehdg::ExprProg* d_prog[4] = {nullptr, nullptr, nullptr, nullptr};   // coeff, exact, d/dx exact, d/dy exact (ehdg_set_expression)
    void* user_ws = nullptr;          // caller-provided workspace of the Krylov solvers (ehdg_set_workspace)
    size_t user_ws_bytes = 0;
    cudaStream_t side = nullptr;      // the general (enriched) kernels run beside the polynomial ones
    int overlap_general = 0;          // 1: quadrature-path kernels on `side`, concurrently with the polynomial-path kernels
    int dynamic_work = 1;             // bit 0: the assemble kernel draws its elements from a counter (work_counters), bit 1: alpha too
    int* work_counters = nullptr;     // [2] device: alpha, assemble
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // optional CUDA-event timing of the two polynomial-path kernels on their launching stream
    int time_kernels = 0;
    cudaEvent_t ev_k[4] = {nullptr, nullptr, nullptr, nullptr};   // alpha start/stop, assemble start/stop
};

struct ehdg_plan {
    ehdg_ctx* ctx = nullptr;
    int32_t* lists = nullptr;       // [ne] backing store of list_poly | list_gen
    int32_t ne;
    int64_t n_poly, n_gen;
    int32_t* eslot = nullptr;       // [ne]
    int32_t* list_poly = nullptr;   // [n_poly]
    int32_t* list_gen = nullptr;    // [n_gen]
    int nf_poly, nfmax, n_p, nvmax;
    int forced = 0;                 // every element on the quadrature path (cross-checks)
};

#define EHDG_SLOT_GEN 0x40000000
#define EHDG_SLOT_MASK 0x3fffffff
#define EHDG_SLOT_NONE ((int32_t)0x80000000)     /* element outside the plan's selection (partitioned assembly) */

struct ehdg_csr {
    int32_t nf;
    int64_t n_rows, nnz;
    int32_t* row_start = nullptr;   // [nf+1]  first row of facet f
    int64_t* rowptr = nullptr;      // [n_rows+1]
    int colind_full = 1;            // 0: only the first row of every facet holds its column indices yet (ensure_colind)
    int32_t* colind = nullptr;      // [nnz]
    int64_t* blk_base = nullptr;    // [nf+1]  first nnz of the row block of facet f
    int32_t* nbr = nullptr;         // [nf][5] neighbour facets ascending, -1 padded
    int32_t* nbr_len = nullptr;     // [nf]    row length (sum of active dofs of the neighbours)
    uint16_t* dofmask = nullptr;    // [nf]    active slots of the facet
    int32_t* row_facet = nullptr;   // [n_rows] facet of each row
    uint8_t* strip = nullptr;       // [nf]    1 = free facet of a marked element (enriched strip), null if none
    void* elmap = nullptr;          // [ne] ElMap: scatter table of the fused a5-a8 store, built on first use
    int32_t elmap_ne = 0;
    int32_t* diag_off = nullptr;    // [nf]    column offset of the facet's own block inside its rows (with elmap)
    int64_t* diag_dst = nullptr;    // [own rows] value index of the row's piece of the diagonal block | entries << 56 (merge of the fused store)
    int32_t* line = nullptr;        // [nf]    chain of every facet for the line-block preconditioner, null until set
    int edg = 0;                    // EDG pattern: the blocks are ELEMENTS (self + facet neighbours), up to n_p + n_enrich rows each
    const double* folded_vals = nullptr;   // value array a Krylov solve has overwritten with A M^-1 (facet block-Jacobi mode)
    double* dinv = nullptr;         // [n_rows * nbmax] block-Jacobi inverses (set by ehdg_gmres)
    int nb;                         // slots per facet
    // partitioned assembly (ehdg_csr_symbolic_part): this rank holds the facets [f_lo, f_hi) = rows [row_lo, row_hi);
    // n_rows stays the GLOBAL size (column indices are global), nnz / colind / the value array cover the own rows only
    int windowed = 0;
    int32_t f_lo = 0, f_hi = 0;
    int64_t row_lo = 0, row_hi = 0;
    int64_t col_lo = 0, col_hi = 0;    // column range the own rows reference (SpMV halo extent)
    cudaStream_t stream = nullptr;     // stream of the pattern's stream-ordered allocations
    int nranks = 1;
    std::vector<int64_t> part_rows;    // [nranks+1] first row of every rank
    std::vector<int32_t> part_facets;  // [nranks+1] first facet of every rank
};

namespace ehdg {
int ensure_colind(const ehdg_csr* A, cudaStream_t st);   // ehdg_solve.cu
int build_elmap(ehdg_csr* A, const ehdg_mesh* m, cudaStream_t st);   // ehdg_fused.cu
int build_diag_dst(ehdg_csr* A, cudaStream_t st);                    // ehdg_fused.cu
struct CsrOut;   // ehdg_fast.h
// polynomial-path kernels (ehdg_fast.cu)
int fast_build_tensors(ehdg_ctx* ctx);
void fast_free_tensors(ehdg_ctx* ctx);
// list/n: elements to process; skip_if_set (optional u8 per element): elements whose byte is non-zero are left untouched
int fast_alpha(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const int32_t* list, int n, const uint8_t* skip_if_set,
               double* lam, int32_t* status, cudaStream_t st);
int fast_assemble(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const double* lam, double* S,
                  double* g, double* W, double* z, int32_t* status, cudaStream_t st, const CsrOut* csr_out = nullptr);
int64_t fast_flops_per_element(const ehdg_ctx* ctx);
int64_t fast_flops_alpha(const ehdg_ctx* ctx);
int64_t fast_flops_assemble(const ehdg_ctx* ctx);
}  // namespace ehdg
