// Line-block Jacobi: the preconditioner of the condensed facet system (a9).
//
// The facets are partitioned into CHAINS: streamline-aligned lines of the mesh (one chain per line of width ~h across
// the flow, ehdg_csr_auto_lines / ehdg_csr_set_lines) and, for enriched problems, the strip of marked elements, whose
// penalties tau ~ eps * 20 lambda_K p^2 / h are orders of magnitude above the bulk (conv_diff.py:349).  Every chain is
// one diagonal block of a block-Jacobi preconditioner and is solved EXACTLY: the facets of a chain are ordered by the
// BFS levels of the chain's facet graph (any graph is block tridiagonal in its BFS levels), consecutive levels are
// merged into blocks of at most CHAIN_MAXB dofs, and the block-tridiagonal matrix is factorised once (block Thomas with
// rank-revealing pivots) by one CTA per chain.  An application is one forward and one backward sweep of small dense
// mat-vecs per chain: ONE WARP PER CHAIN, lanes own the rows of a block, the factor blocks arrive through a ring of
// TMA bulk copies (cp.async.bulk + mbarrier) that runs ahead of the recurrence, so the sequential chain only ever waits
// for shared memory.  Facet block-Jacobi is the special case "every facet its own chain".
//
// Measured on the 64^2 .. 128^2 order-4 systems (host study, tools/precond_study.py): 7-13x fewer GMRES iterations than
// facet block-Jacobi; the count is insensitive to the restart length.
#include <algorithm>
#include <numeric>
#include <vector>

#include "ehdg_chain.h"
#include "ehdg_internal.h"

namespace ehdg {

constexpr int CHAIN_MAXB = 64;       // largest block (two rows per lane in the sweep kernel)

// ---------------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------------
__global__ void chain_pos_kernel(int nd, const int32_t* __restrict__ dof_rows, const int32_t* __restrict__ dof_blk,
                                 const ChainBlock* __restrict__ blocks, int32_t* __restrict__ sblk, int32_t* __restrict__ sloc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd || dof_rows[s] < 0) return;
    const int b = dof_blk[s];
    sblk[dof_rows[s]] = b;
    sloc[dof_rows[s]] = s - blocks[b].s0;
}

// dense blocks (column major) from the CSR rows: D_b (n x n) and U_b = A[b, b+1] (n x n_next) into DH, L_b = A[b, b-1]
// (n x n_prev) into G.  Couplings that leave the chain or skip a block do not belong to the preconditioner.
__global__ void chain_extract_kernel(int nd, const int32_t* __restrict__ dof_rows, const int32_t* __restrict__ dof_blk,
                                     const ChainBlock* __restrict__ blocks, const int32_t* __restrict__ blk_chain,
                                     const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                     const double* __restrict__ vals, const int32_t* __restrict__ sblk, const int32_t* __restrict__ sloc,
                                     double* __restrict__ DH, double* __restrict__ G, int* __restrict__ skipped) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd || dof_rows[s] < 0) return;
    const int b = dof_blk[s], r = dof_rows[s];
    const ChainBlock B = blocks[b];
    const int i = s - B.s0, n = B.n, ch = blk_chain[b];
    int nskip = 0;
    for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
        const int c = colind[q];
        const int cb = sblk[c];
        if (cb < 0 || blk_chain[cb] != ch) continue;
        const int j = sloc[c];
        if (cb == b) DH[B.offDH + (int64_t)j * n + i] = vals[q];
        else if (cb == b - 1) G[B.offG + (int64_t)j * n + i] = vals[q];
        else if (cb == b + 1) DH[B.offDH + (int64_t)n * n + (int64_t)j * n + i] = vals[q];
        else if (vals[q] != 0.0) ++nskip;     // cannot happen for BFS levels
    }
    if (nskip) atomicAdd(skipped, nskip);
}

// Block LU of one chain per CTA.  On exit: DH holds [Delta_b^-1 | H_b = Delta_b^-1 U_b] (rank-revealing: dependent dofs
// get zero rows / columns), G holds G_b = L_b Delta_{b-1}^-1.
__global__ void __launch_bounds__(128)
chain_factor_kernel(const int32_t* __restrict__ chain_ptr, const ChainBlock* __restrict__ blocks, const int32_t* __restrict__ dof_rows,
                    uint8_t* __restrict__ rowdrop, double drop_tol, double* DH, double* G, int* dropped) {
    extern __shared__ double sm[];
    double* Aw = sm;                                   // n x 2n augmented [Delta | I], row major
    __shared__ double s_scale;
    __shared__ unsigned char s_drop[CHAIN_MAXB];
    __shared__ double s_dsc[CHAIN_MAXB];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b0 = chain_ptr[blockIdx.x], b1 = chain_ptr[blockIdx.x + 1];
    for (int b = b0; b < b1; ++b) {
        const ChainBlock B = blocks[b];
        const int n = B.n;
        double* Db = DH + B.offDH;                     // column major
        const int np_ = b > b0 ? blocks[b - 1].n : 0;
        double* Uprev = b > b0 ? DH + blocks[b - 1].offDH + (int64_t)np_ * np_ : nullptr;   // np_ x n, still the original U_{b-1}
        // symmetric scaling by the ORIGINAL diagonal, d_i = |D_b[i][i]|^-1/2: the elimination then runs on a matrix whose
        // entries are relative to each dof's own size, and a pivot is the share of a dof that the dofs eliminated before it
        // (earlier blocks of the chain and earlier pivots of this block) do not explain -- whatever the row scales
        // (penalties of marked elements are 1e5 .. 1e10 times those of their neighbours)
        if (tid < n) {
            const double dg = fabs(Db[tid * n + tid]);
            s_dsc[tid] = dg > 0.0 ? rsqrt(dg) : 1.0;
        }
        __syncthreads();
        // Delta_b = D_b - G_b U_{b-1}   (G_b was completed at the end of step b-1)
        if (b > b0) {
            const double* Gb = G + B.offG;             // n x np_
            for (int idx = tid; idx < n * n; idx += nt) {
                const int i = idx % n, j = idx / n;
                double acc = Db[idx];
                for (int k = 0; k < np_; ++k) acc -= Gb[k * n + i] * Uprev[j * np_ + k];
                Aw[i * 2 * n + j] = acc * s_dsc[i] * s_dsc[j];
            }
        } else {
            for (int idx = tid; idx < n * n; idx += nt) Aw[(idx % n) * 2 * n + idx / n] = Db[idx] * s_dsc[idx % n] * s_dsc[idx / n];
        }
        for (int idx = tid; idx < n * n; idx += nt) Aw[(idx / n) * 2 * n + n + idx % n] = (idx / n == idx % n) ? 1.0 : 0.0;
        if (tid < n) s_drop[tid] = rowdrop ? rowdrop[dof_rows[B.s0 + tid]] : 0;
        __syncthreads();
        // H_{b-1} = Delta_{b-1}^-1 U_{b-1}: the original U_{b-1} is not needed any more
        if (b > b0) {
            const double* Dp = DH + blocks[b - 1].offDH;
            double* tmp = Aw + 2 * n * n;              // scratch behind the augmented matrix (sized at set-up)
            for (int idx = tid; idx < np_ * n; idx += nt) {
                const int i = idx % np_, j = idx / np_;
                double acc = 0.0;
                for (int k = 0; k < np_; ++k) acc += Dp[k * np_ + i] * Uprev[j * np_ + k];
                tmp[idx] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < np_ * n; idx += nt) Uprev[idx] = tmp[idx];
        }
        // dofs already known to be dependent inside their facet block: decouple
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx % n;
            if (s_drop[i] || s_drop[j]) Aw[i * 2 * n + j] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (tid == 0) s_scale = 1.0;           // the scaled matrix has a unit original diagonal
        __syncthreads();
        // Gauss-Jordan on the diagonal: the condensed (E)HDG form is coercive, so every Schur complement met here has a
        // positive-definite symmetric part and needs no row exchanges; what the diagonal pivot measures after the scaling above
        // is the share of dof k that the dofs eliminated before it do not explain.  Below drop_tol the dof is dependent and
        // leaves the block as test AND trial function (row k and column k), consistently with the caller's row weights.
        for (int k = 0; k < n; ++k) {
            const double piv = Aw[k * 2 * n + k];
            __syncthreads();
            if (!(fabs(piv) > drop_tol * s_scale) || s_drop[k]) {
                if (!s_drop[k] && tid == 0) { s_drop[k] = 1; atomicAdd(dropped, 1); }
                for (int i = tid; i < n; i += nt) Aw[i * 2 * n + k] = (i == k) ? 1.0 : 0.0;
                __syncthreads();
                for (int j = tid; j < 2 * n; j += nt) Aw[k * 2 * n + j] = (j == k) ? 1.0 : 0.0;
                __syncthreads();
                continue;
            }
            const double inv = 1.0 / piv;
            for (int j = tid; j < 2 * n; j += nt) Aw[k * 2 * n + j] *= inv;
            __syncthreads();
            for (int idx = tid; idx < n * 2 * n; idx += nt) {
                const int i = idx / (2 * n), j = idx % (2 * n);
                if (i == k || j == k) continue;
                Aw[idx] -= Aw[i * 2 * n + k] * Aw[k * 2 * n + j];
            }
            __syncthreads();
            for (int i = tid; i < n; i += nt)
                if (i != k) Aw[i * 2 * n + k] = 0.0;
            __syncthreads();
        }
        // Delta_b^-1 with dropped rows / columns zeroed, column major; the dropped dofs are reported per row
        if (rowdrop && tid < n && s_drop[tid]) rowdrop[dof_rows[B.s0 + tid]] = 1;
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx % n, j = idx / n;
            Db[idx] = (s_drop[i] || s_drop[j]) ? 0.0 : Aw[i * 2 * n + n + j] * s_dsc[i] * s_dsc[j];
        }
        __syncthreads();
        if (b + 1 < b1) {
            // G_{b+1} = L_{b+1} Delta_b^-1  (in place via the scratch)
            const int nn = blocks[b + 1].n;
            double* Lb = G + blocks[b + 1].offG;       // nn x n column major
            for (int idx = tid; idx < nn * n; idx += nt) {
                const int i = idx % nn, j = idx / nn;
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc += Lb[k * nn + i] * Db[j * n + k];
                Aw[idx] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < nn * n; idx += nt) Lb[idx] = Aw[idx];
            __syncthreads();
        } else if (n > 0) {
            // the last block of a chain has no upper coupling: its (padding) H region stays zero
        }
    }
}

// ---------------------------------------------------------------------------------------------
// application: one warp per chain, factor blocks through a TMA ring
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// vp[s] = v[dof_rows[s] - row0] (0 in padding slots): the chains read their right-hand side in chain order
__global__ void chain_gather_kernel(int nd, const int32_t* __restrict__ dof_rows, const double* __restrict__ v, int64_t row0,
                                    const double* __restrict__ vscale, double* __restrict__ vp) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd) return;
    const int r = dof_rows[s];
    vp[s] = r >= 0 ? (vscale ? v[r - row0] * vscale[r - row0] : v[r - row0]) : 0.0;
}
__global__ void chain_scatter_kernel(int nd, const int32_t* __restrict__ dof_rows, const double* __restrict__ xp, int64_t row0,
                                     double* __restrict__ t) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd) return;
    const int r = dof_rows[s];
    if (r >= 0) t[r - row0] = xp[s];
}

__device__ __forceinline__ int pad2i(int x) { return x < 2 ? 2 : ((x + 1) & ~1); }

// One warp per chain.  Everything the sequential recurrence touches is either in registers or arrives through the TMA ring:
// the only per-block metadata is the block size (one byte, fetched 32 blocks at a time and 32 steps ahead); positions and
// factor offsets are running sums of the sizes, the right-hand side comes in chain order with the factor block (second bulk
// copy on the stage's mbarrier), and the forward results the backward sweep needs are read two steps ahead.
template <int STAGES>
__global__ void __launch_bounds__(32)
chain_sweep_kernel(int nchains, const int32_t* __restrict__ chain_ptr, const ChainBlock* __restrict__ blocks,
                   const uint8_t* __restrict__ nsz, const double* __restrict__ G, const double* __restrict__ DH,
                   const double* __restrict__ vp, double* __restrict__ c, double* __restrict__ xp, int stage_doubles) {
    extern __shared__ __align__(16) double ring[];              // STAGES x stage_doubles
    __shared__ __align__(8) uint64_t bar[STAGES];
    const int chain = blockIdx.x, lane = threadIdx.x;
    if (chain >= nchains) return;
    const int b0 = chain_ptr[chain], b1 = chain_ptr[chain + 1], nblk = b1 - b0;
    if (nblk <= 0) return;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const ChainBlock first = blocks[b0], last = blocks[b1 - 1];
    const uint8_t* sz = nsz + b0;
    // sizes of the blocks [base, base + 64) of this chain in two registers per lane
    int base = 0;
    int ch0 = lane < nblk ? sz[lane] : 0, ch1 = 32 + lane < nblk ? sz[32 + lane] : 0;
    auto size_of = [&](int idx) -> int {          // idx uniform, base <= idx < base + 64 (or outside the chain: 0)
        if (idx < 0 || idx >= nblk) return 0;
        const int o = idx - base;
        return __shfl_sync(0xffffffffu, o < 32 ? ch0 : ch1, o & 31);
    };
    __syncwarp();
    // ---- forward: c_b = v_b - G_b c_{b-1}.  Stage k % STAGES holds [G_b | v_b].
    int64_t iss_offG = first.offG;                // issue pointer: next block to request
    int iss_k = 0, iss_s0 = first.s0, iss_np = 0;
    auto issue_fwd = [&]() {
        const int n = size_of(iss_k);
        const int gd = pad2i(n * iss_np), vd = pad2i(n);
        if (lane == 0) {
            uint64_t* br = &bar[iss_k % STAGES];
            double* dst = ring + (size_t)(iss_k % STAGES) * stage_doubles;
            mbar_expect_tx(br, (uint32_t)((gd + vd) * 8));
            bulk_g2s(dst, G + iss_offG, (uint32_t)(gd * 8), br);
            bulk_g2s(dst + gd, vp + iss_s0, (uint32_t)(vd * 8), br);
        }
        iss_offG += gd;
        iss_s0 += vd;
        iss_np = n;
        ++iss_k;
    };
    for (int q = 0; q < STAGES && q < nblk; ++q) issue_fwd();
    double cp0 = 0.0, cp1 = 0.0;           // c of the previous block: rows lane, lane + 32
    int np_ = 0, s0 = first.s0;
    for (int k = 0; k < nblk; ++k) {
        if (k - base == 32) {              // slide the size window: the new chunk is first used 32 - STAGES steps from now
            base += 32;
            ch0 = ch1;
            ch1 = base + 32 + lane < nblk ? sz[base + 32 + lane] : 0;
        }
        const int n = size_of(k);
        const int gd = pad2i(n * np_);
        const double* Gs = ring + (size_t)(k % STAGES) * stage_doubles;
        const bool r0 = lane < n, r1 = lane + 32 < n;
        mbar_wait(&bar[k % STAGES], (k / STAGES) & 1);
        double a0 = r0 ? Gs[gd + lane] : 0.0, a1 = r1 ? Gs[gd + lane + 32] : 0.0, e0 = 0.0, e1 = 0.0;
        int j = 0;
        for (; j + 1 < np_; j += 2) {
            const double x0 = j < 32 ? __shfl_sync(0xffffffffu, cp0, j) : __shfl_sync(0xffffffffu, cp1, j - 32);
            const double x1 = j + 1 < 32 ? __shfl_sync(0xffffffffu, cp0, j + 1) : __shfl_sync(0xffffffffu, cp1, j + 1 - 32);
            if (r0) { a0 -= Gs[j * n + lane] * x0; e0 -= Gs[(j + 1) * n + lane] * x1; }
            if (r1) { a1 -= Gs[j * n + lane + 32] * x0; e1 -= Gs[(j + 1) * n + lane + 32] * x1; }
        }
        if (j < np_) {
            const double x0 = j < 32 ? __shfl_sync(0xffffffffu, cp0, j) : __shfl_sync(0xffffffffu, cp1, j - 32);
            if (r0) a0 -= Gs[j * n + lane] * x0;
            if (r1) a1 -= Gs[j * n + lane + 32] * x0;
        }
        a0 += e0;
        a1 += e1;
        __syncwarp();                       // every lane is done with this stage
        if (iss_k < nblk) issue_fwd();
        if (r0) c[s0 + lane] = a0;
        if (r1) c[s0 + lane + 32] = a1;
        cp0 = r0 ? a0 : 0.0;
        cp1 = r1 ? a1 : 0.0;
        np_ = n;
        s0 += pad2i(n);
    }
    // ---- backward: x_b = Dinv_b c_b - H_b x_{b+1}; the ring sequence continues at it = nblk.  Stage holds [Dinv_b | H_b].
    // size window for the way back: blocks [base, base + 64) with base the largest multiple of 32 <= nblk - 1, minus 32
    base = ((nblk - 1) / 32) * 32 - 32;
    ch0 = (base >= 0 && base + lane < nblk) ? sz[base + lane] : 0;
    ch1 = base + 32 + lane < nblk ? sz[base + 32 + lane] : 0;
    __syncwarp();
    int64_t iss_offDH = last.offDH;
    int iss_q = 0, iss_nn = 0;             // issue pointer walks down: block nblk - 1 - iss_q
    auto issue_bwd = [&]() {
        const int kb = nblk - 1 - iss_q, it = nblk + iss_q;
        const int n = size_of(kb);
        const int dd = pad2i(n * n + n * iss_nn);
        if (iss_q > 0) iss_offDH -= dd;      // region of block kb ends where the one above starts
        if (lane == 0) {
            uint64_t* br = &bar[it % STAGES];
            mbar_expect_tx(br, (uint32_t)(dd * 8));
            bulk_g2s(ring + (size_t)(it % STAGES) * stage_doubles, DH + iss_offDH, (uint32_t)(dd * 8), br);
        }
        iss_nn = n;
        ++iss_q;
    };
    for (int q = 0; q < STAGES && q < nblk; ++q) issue_bwd();
    double xn0 = 0.0, xn1 = 0.0;           // x of the block after this one
    int nn = 0;
    s0 = last.s0;
    // c two steps ahead: cA = block k - 1, cB = block k - 2 (same lanes wrote them in the forward sweep)
    double c0 = cp0, c1 = cp1;             // c of the last block is still in registers
    int nA = size_of(nblk - 2), nB = size_of(nblk - 3);
    int sA = s0 - pad2i(nA), sB = sA - pad2i(nB);
    double cA0 = lane < nA ? c[sA + lane] : 0.0, cA1 = lane + 32 < nA ? c[sA + lane + 32] : 0.0;
    double cB0 = lane < nB ? c[sB + lane] : 0.0, cB1 = lane + 32 < nB ? c[sB + lane + 32] : 0.0;
    for (int q = 0; q < nblk; ++q) {
        const int k = nblk - 1 - q, it = nblk + q;
        if (k < base + 32 && base >= 0) {  // slide the window down as soon as block k sits in the lower chunk
            base -= 32;
            ch1 = ch0;
            ch0 = (base >= 0 && base + lane < nblk) ? sz[base + lane] : 0;
        }
        const int n = size_of(k);
        // request c of block k - 3 (needed three steps from now as cB -> cA -> c)
        const int nC = size_of(k - 3);
        const int sC = sB - pad2i(nC);
        const double cC0 = lane < nC ? c[sC + lane] : 0.0, cC1 = lane + 32 < nC ? c[sC + lane + 32] : 0.0;
        const double* Ds = ring + (size_t)(it % STAGES) * stage_doubles;
        const double* Hs = Ds + (size_t)n * n;
        const bool r0 = lane < n, r1 = lane + 32 < n;
        mbar_wait(&bar[it % STAGES], (it / STAGES) & 1);
        double a0 = 0.0, a1 = 0.0, e0 = 0.0, e1 = 0.0;
        for (int j = 0; j < n; ++j) {
            const double x0 = j < 32 ? __shfl_sync(0xffffffffu, c0, j) : __shfl_sync(0xffffffffu, c1, j - 32);
            if (r0) a0 += Ds[j * n + lane] * x0;
            if (r1) a1 += Ds[j * n + lane + 32] * x0;
        }
        for (int j = 0; j < nn; ++j) {
            const double x0 = j < 32 ? __shfl_sync(0xffffffffu, xn0, j) : __shfl_sync(0xffffffffu, xn1, j - 32);
            if (r0) e0 -= Hs[j * n + lane] * x0;
            if (r1) e1 -= Hs[j * n + lane + 32] * x0;
        }
        a0 += e0;
        a1 += e1;
        __syncwarp();
        if (iss_q < nblk) issue_bwd();
        if (r0) xp[s0 + lane] = a0;
        if (r1) xp[s0 + lane + 32] = a1;
        xn0 = r0 ? a0 : 0.0;
        xn1 = r1 ? a1 : 0.0;
        nn = n;
        s0 = sA;
        sA = sB;
        sB = sC;
        c0 = cA0; c1 = cA1;
        cA0 = cB0; cA1 = cB1;
        cB0 = cC0; cB1 = cC1;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {
struct DevFree {
    std::vector<void*> p;
    ~DevFree() { for (void* q : p) cudaFree(q); }
    template <class T> T* keep(T* q) { p.push_back((void*)q); return q; }
};
}  // namespace

int chain_setup(ehdg_csr* A, const double* vals, uint8_t* rowdrop, const int32_t* line_dev, int use_strip, int target_dofs,
                double drop_tol, int max_blocks, cudaStream_t st, ChainSolver** out) {
    *out = nullptr;
    const int nf = A->nf;
    const bool strip = use_strip && A->strip != nullptr;
    if ((!line_dev && !strip) || nf == 0) return EHDG_OK;
    if (target_dofs <= 0) target_dofs = 1;
    const int f_lo = A->windowed ? A->f_lo : 0, f_hi = A->windowed ? A->f_hi : nf;
    // ---- topology to the host: neighbour lists, rows per facet, chain ids
    std::vector<int32_t> nbr(5 * (size_t)nf), row_start((size_t)nf + 1), line((size_t)nf, -1);
    std::vector<uint8_t> stripf;
    EHDG_CUDA(cudaMemcpyAsync(nbr.data(), A->nbr, sizeof(int32_t) * 5 * (size_t)nf, cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaMemcpyAsync(row_start.data(), A->row_start, sizeof(int32_t) * ((size_t)nf + 1), cudaMemcpyDeviceToHost, st));
    if (line_dev) EHDG_CUDA(cudaMemcpyAsync(line.data(), line_dev, sizeof(int32_t) * (size_t)nf, cudaMemcpyDeviceToHost, st));
    if (strip) {
        stripf.resize(nf);
        EHDG_CUDA(cudaMemcpyAsync(stripf.data(), A->strip, (size_t)nf, cudaMemcpyDeviceToHost, st));
    }
    EHDG_CUDA(cudaStreamSynchronize(st));
    // ---- chain of every participating facet (dense ids, ascending in the order of first appearance of the key)
    //      key: strip facets share one key; line >= 0: that line; otherwise (lines given) a chain of its own
    std::vector<int64_t> key((size_t)nf, -1);
    const int64_t KEY_STRIP = (int64_t)1 << 40, KEY_SINGLE = (int64_t)1 << 41;
    size_t n_part = 0;
    for (int f = f_lo; f < f_hi; ++f) {
        if (row_start[f + 1] == row_start[f]) continue;
        if (strip && stripf[f]) key[f] = KEY_STRIP;
        else if (!line_dev) continue;
        else key[f] = line[f] >= 0 ? (int64_t)line[f] : KEY_SINGLE + f;
        ++n_part;
    }
    if (n_part == 0) return EHDG_OK;
    std::vector<int32_t> part;                 // participating facets sorted by (key, facet)
    part.reserve(n_part);
    for (int f = f_lo; f < f_hi; ++f)
        if (key[f] >= 0) part.push_back(f);
    std::stable_sort(part.begin(), part.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<int32_t> chain_first;          // index into `part` of the first facet of every chain
    for (size_t i = 0; i < part.size(); ++i)
        if (i == 0 || key[part[i]] != key[part[i - 1]]) chain_first.push_back((int32_t)i);
    const int nchains = (int)chain_first.size();
    chain_first.push_back((int32_t)part.size());
    std::vector<int32_t> chain_of((size_t)nf, -1);
    for (int ch = 0; ch < nchains; ++ch)
        for (int i = chain_first[ch]; i < chain_first[ch + 1]; ++i) chain_of[part[i]] = ch;
    // ---- per chain: BFS levels from a pseudo-peripheral facet, levels merged into blocks
    std::vector<int32_t> level((size_t)nf, -1), order, lev_of_order, queue;
    order.reserve(part.size());
    lev_of_order.reserve(part.size());
    std::vector<ChainBlock> blocks;
    std::vector<int32_t> chain_ptr{0}, dof_rows, dof_blk, blk_chain;
    int nchains_out = 0;                      // chains after cutting (segments)
    int nmax = 0;
    int64_t n_real = 0;
    auto bfs = [&](int seed, int ch, int lev0, std::vector<int32_t>& visit) -> int {      // returns the last level used
        visit.clear();
        level[seed] = lev0;
        visit.push_back(seed);
        int maxlev = lev0;
        for (size_t h = 0; h < visit.size(); ++h) {
            const int u = visit[h];
            maxlev = std::max(maxlev, (int)level[u]);
            for (int t2 = 0; t2 < 5; ++t2) {
                const int g = nbr[5 * (size_t)u + t2];
                if (g < 0 || chain_of[g] != ch || level[g] >= 0) continue;
                level[g] = level[u] + 1;
                visit.push_back(g);
            }
        }
        return maxlev;
    };
    // max_blocks < 0: automatic.  Measured on B200 (profiles/r02_line_segments.jsonl, r02d_bench_1gpu_4M.json): with 708 lines
    // (1M elements, 4.8 chains per SM) the sweep is latency-bound and cutting every line in 4 takes the iteration from 3.9 to
    // 2.3 ms for 15 % more iterations (2.6 -> 1.8 s); 16 segments stop converging.  With 1 414 lines (4M) the iteration is not
    // sweep-bound: 2 segments leave it at 8.3 ms and cost 50 % more iterations (13.6 -> 19.5 s).  With 224 lines of ~670 blocks
    // (100k) 4 segments cost 50 % more iterations and the convergence of the enriched config-2 system.
    // Rule: cut only while there are fewer chains than 8 warps per SM, into at most 4 segments of at least 350 blocks.
    const int auto_seg = nchains >= 1184 ? 1 : std::max(1, std::min(4, 2832 / std::max(1, nchains)));
    constexpr int AUTO_SEG_MIN_BLOCKS = 350;
    std::vector<int32_t> visit, visit2;
    for (int ch = 0; ch < nchains; ++ch) {
        const size_t o0 = order.size();
        int lev_base = 0;
        for (int i = chain_first[ch]; i < chain_first[ch + 1]; ++i) {
            const int seed0 = part[i];
            if (level[seed0] >= 0) continue;
            // pseudo-peripheral start: the facet found last by a BFS from the component's first facet
            bfs(seed0, ch, 0, visit);
            const int far = visit.back();
            for (int u : visit) level[u] = -1;
            int maxlev = bfs(far, ch, lev_base, visit2);
            // The strip of marked elements is ONE chain and, being never cut (below), the longest by far: two arms along the
            // outflow sides joined at the corner.  Rooting its level structure in the middle of the level range halves the
            // number of levels (the two halves advance side by side, blocks twice as large) and with it the latency of the
            // sweep; any root gives a block-tridiagonal matrix.  Taken only if the wider levels still fit one row per lane.
            // Measured (profiles/r02j / r02k bench lines): config 3 (1M) 8.1 -> 7.1 ms per iteration, 6.4 -> 5.7 s; config 2
            // through the driver 1.38 -> 1.11 s; same iteration counts (the preconditioner is the same operator).
            if (strip && key[seed0] == KEY_STRIP && maxlev - lev_base >= 64) {
                const int mid = lev_base + (maxlev - lev_base) / 2;
                int center = -1;
                for (int u : visit2)
                    if (level[u] == mid) { center = u; break; }
                std::vector<int32_t> keep_order(visit2);
                std::vector<int32_t> keep_level(visit2.size());
                for (size_t k = 0; k < visit2.size(); ++k) keep_level[k] = level[visit2[k]];
                for (int u : visit2) level[u] = -1;
                const int maxlev2 = bfs(center, ch, lev_base, visit2);
                std::vector<int> ldofs((size_t)(maxlev2 - lev_base + 1), 0);
                for (int u : visit2) ldofs[(size_t)(level[u] - lev_base)] += row_start[u + 1] - row_start[u];
                bool fits = visit2.size() == keep_order.size();
                for (int d : ldofs) fits = fits && d <= 32;       // one row per lane in the sweep; beyond that the wider steps cost
                                                                    // more than the halved length saves (4M unstructured: 25 -> 45 rows,
                                                                    // 38 -> 44 ms per iteration)
                if (fits) {
                    maxlev = maxlev2;
                } else {                                   // back to the peripheral root
                    for (int u : visit2) level[u] = -1;
                    visit2 = keep_order;
                    for (size_t k = 0; k < visit2.size(); ++k) level[visit2[k]] = keep_level[k];
                }
            }
            // BFS order is level order
            for (int u : visit2) { order.push_back(u); lev_of_order.push_back(level[u]); }
            lev_base = maxlev + 2;          // gap: components never share a block boundary coupling
        }
        // merge consecutive levels into blocks
        size_t i = o0;
        while (i < order.size()) {
            int dofs = 0;
            size_t j = i;
            int last_level = lev_of_order[i];
            while (j < order.size()) {
                size_t k = j;
                int ld = 0;
                const int lv = lev_of_order[j];
                while (k < order.size() && lev_of_order[k] == lv) { ld += row_start[order[k] + 1] - row_start[order[k]]; ++k; }
                if (ld > CHAIN_MAXB) { set_error("line preconditioner: a BFS level with %d dofs exceeds %d", ld, CHAIN_MAXB); return EHDG_ERR_ARG; }
                if (j > i && (dofs + ld > target_dofs || lv != last_level + 1)) break;
                dofs += ld;
                last_level = lv;
                j = k;
            }
            ChainBlock B{};
            B.s0 = (int32_t)dof_rows.size();
            const int b = (int)blocks.size();
            for (size_t k = i; k < j; ++k) {
                const int u = order[k];
                for (int d = row_start[u]; d < row_start[u + 1]; ++d) { dof_rows.push_back(d); dof_blk.push_back(b); }
            }
            B.n = (int32_t)dof_rows.size() - B.s0;
            n_real += B.n;
            while ((int)dof_rows.size() - B.s0 < 2 || ((dof_rows.size() - (size_t)B.s0) & 1)) { dof_rows.push_back(-1); dof_blk.push_back(b); }   // even, 16-byte slots
            nmax = std::max(nmax, (int)B.n);
            blocks.push_back(B);
            blk_chain.push_back(nchains_out);
            i = j;
        }
        // long chains are cut into segments of at most max_blocks blocks (each solved exactly, the coupling between two
        // segments is left to the Krylov iteration): the sweep is a sequential recurrence of one warp per chain, so its
        // latency is the chain length, and a GPU (or a rank of a multi-GPU run) with few chains has warp slots to spare
        {
            const int first_blk = chain_ptr.back(), nb_ch = (int)blocks.size() - first_blk;
            int nseg = (max_blocks > 0 && nb_ch > max_blocks) ? (nb_ch + max_blocks - 1) / max_blocks : 1;
            if (max_blocks < 0) nseg = std::max(1, std::min(auto_seg, nb_ch / AUTO_SEG_MIN_BLOCKS));
            // the strip of marked elements stays whole under the automatic rule: the enriched systems owe their convergence to
            // its exact solve -- config 2: whole 885 iterations; cut in 2 it stalls at 9e-8, in 3 or 4 at 1e-6 (driver tolerance
            // 1e-10; profiles/r02d / r02g / r02h bench lines).  The price: at 1M the whole strip (2 800 blocks) is the longest
            // chain by a factor of 8 and holds every sweep (config 3: 6.5 s whole against 3.3-3.8 s cut, both converging).
            if (max_blocks < 0 && strip && key[part[chain_first[ch]]] == KEY_STRIP) nseg = 1;
            for (int sg = 1; sg <= nseg; ++sg) {
                const int end_blk = first_blk + (int)(((int64_t)nb_ch * sg) / nseg);
                if (sg > 1)
                    for (int b2 = chain_ptr.back(); b2 < end_blk; ++b2) blk_chain[b2] = nchains_out;
                chain_ptr.push_back(end_blk);
                ++nchains_out;
            }
            // blocks of the first segment keep the id they were given; ids must be unique per segment
        }
    }
    const int nblk = (int)blocks.size(), nd = (int)dof_rows.size();
    if (nd == 0) return EHDG_OK;
    // ---- factor storage: G_b (n x n_prev), [Dinv_b | H_b] (n x n, n x n_next); every region padded to 16 bytes, never empty
    int64_t tG = 0, tDH = 0, stage = 0;
    auto pad2 = [](int64_t x) { return std::max<int64_t>(2, (x + 1) & ~(int64_t)1); };
    for (int b = 0; b < nblk; ++b) {
        const int ch = blk_chain[b];
        const int64_t n = blocks[b].n;
        const int64_t np_ = (b > 0 && blk_chain[b - 1] == ch) ? blocks[b - 1].n : 0;
        const int64_t nn = (b + 1 < nblk && blk_chain[b + 1] == ch) ? blocks[b + 1].n : 0;
        blocks[b].offG = tG;
        blocks[b].offDH = tDH;
        const int64_t g = pad2(n * np_), dh = pad2(n * n + n * nn);
        tG += g;
        tDH += dh;
        stage = std::max(stage, std::max(g + pad2(n), dh));      // forward stage: [G_b | v_b], backward: [Dinv_b | H_b]
    }
    ChainBlock sentinel{};
    sentinel.s0 = nd;
    sentinel.offG = tG;
    sentinel.offDH = tDH;
    blocks.push_back(sentinel);
    ChainSolver* S = new ChainSolver();
    S->n_dofs = nd;
    S->n_blocks = nblk;
    S->n_chains = nchains_out;
    S->n_rows = A->n_rows;
    S->nmax = nmax;
    S->factor_bytes = (tG + tDH) * 8;
    S->stage_bytes = (int)(stage * 8);
    {
        const int64_t own_rows = A->windowed ? A->row_hi - A->row_lo : A->n_rows;
        S->all_rows = n_real == own_rows ? 1 : 0;
    }
    auto fail = [&](int rc) { chain_free(S); return rc; };
    auto up = [&](auto** dptr, const auto& h) -> int {
        using T = typename std::remove_reference<decltype(h)>::type::value_type;
        EHDG_CUDA(cudaMalloc(dptr, sizeof(T) * std::max<size_t>(h.size(), 1)));
        EHDG_CUDA(cudaMemcpyAsync(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
        return EHDG_OK;
    };
    int rc;
    if ((rc = up(&S->chain_ptr, chain_ptr))) return fail(rc);
    if ((rc = up(&S->blocks, blocks))) return fail(rc);
    if ((rc = up(&S->dof_rows, dof_rows))) return fail(rc);
    if ((rc = up(&S->dof_blk, dof_blk))) return fail(rc);
    if ((rc = up(&S->blk_chain, blk_chain))) return fail(rc);
    {
        std::vector<uint8_t> nsz((size_t)nblk);
        for (int b = 0; b < nblk; ++b) nsz[b] = (uint8_t)blocks[b].n;
        if ((rc = up(&S->nsz, nsz))) return fail(rc);
    }
    auto dev_alloc = [&]() -> int {
        EHDG_CUDA(cudaMalloc(&S->sblk, sizeof(int32_t) * (size_t)A->n_rows));
        EHDG_CUDA(cudaMalloc(&S->sloc, sizeof(int32_t) * (size_t)A->n_rows));
        EHDG_CUDA(cudaMemsetAsync(S->sblk, 0xff, sizeof(int32_t) * (size_t)A->n_rows, st));
        EHDG_CUDA(cudaMalloc(&S->G, sizeof(double) * (size_t)tG));
        EHDG_CUDA(cudaMalloc(&S->DH, sizeof(double) * (size_t)tDH));
        EHDG_CUDA(cudaMalloc(&S->c, sizeof(double) * 3 * (size_t)nd));
        S->vp = S->c + nd;
        S->xp = S->c + 2 * (size_t)nd;
        EHDG_CUDA(cudaMemsetAsync(S->G, 0, sizeof(double) * (size_t)tG, st));
        EHDG_CUDA(cudaMemsetAsync(S->DH, 0, sizeof(double) * (size_t)tDH, st));
        return EHDG_OK;
    };
    if ((rc = dev_alloc())) return fail(rc);
    int* d_flags = nullptr;
    auto run = [&]() -> int {
        EHDG_CUDA(cudaMalloc(&d_flags, 2 * sizeof(int)));
        EHDG_CUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
        chain_pos_kernel<<<(nd + 127) / 128, 128, 0, st>>>(nd, S->dof_rows, S->dof_blk, S->blocks, S->sblk, S->sloc);
        EHDG_LAUNCH_CHECK();
        const int rce = ensure_colind(A, st);
        if (rce) return rce;
        chain_extract_kernel<<<(nd + 127) / 128, 128, 0, st>>>(nd, S->dof_rows, S->dof_blk, S->blocks, S->blk_chain, A->rowptr, A->colind,
                                                              vals, S->sblk, S->sloc, S->DH, S->G, d_flags);
        EHDG_LAUNCH_CHECK();
        const int smem = (2 * nmax * nmax + nmax * nmax) * 8;      // augmented matrix + scratch
        EHDG_CUDA(cudaFuncSetAttribute(chain_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        chain_factor_kernel<<<nchains_out, 128, smem, st>>>(S->chain_ptr, S->blocks, S->dof_rows, rowdrop, drop_tol, S->DH, S->G, d_flags + 1);
        EHDG_LAUNCH_CHECK();
        int flags[2] = {0, 0};
        EHDG_CUDA(cudaMemcpyAsync(flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        EHDG_CUDA(cudaStreamSynchronize(st));
        if (flags[0]) { set_error("line preconditioner: matrix is not block tridiagonal in the BFS blocks"); return EHDG_ERR_ARG; }
        S->dropped = flags[1];
        return EHDG_OK;
    };
    rc = run();
    cudaFree(d_flags);
    if (rc) return fail(rc);
    // ring of the sweep kernel: as many stages as fit 24 KB per warp, between 2 and 8
    int stages = (int)std::min<int64_t>(8, std::max<int64_t>(2, (24 * 1024) / std::max(1, S->stage_bytes)));
    stages = stages >= 8 ? 8 : (stages >= 4 ? 4 : 2);
    S->stages = stages;
    *out = S;
    return EHDG_OK;
}

int chain_apply(const ChainSolver* S, const double* v, double* t, int64_t row0, cudaStream_t st, const double* vscale) {
    if (!S) return EHDG_OK;
    const int sd = S->stage_bytes / 8;
    const size_t smem = (size_t)S->stages * S->stage_bytes;
    const int nd = S->n_dofs;
    chain_gather_kernel<<<(nd + 255) / 256, 256, 0, st>>>(nd, S->dof_rows, v, row0, vscale, S->vp);
#define EHDG_SWEEP(ST)                                                                                                          \
    do {                                                                                                                        \
        static size_t set_##ST = 0;                                                                                             \
        if (smem > 48 * 1024 && smem > set_##ST) {                                                                              \
            EHDG_CUDA(cudaFuncSetAttribute(chain_sweep_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
            set_##ST = smem;                                                                                                    \
        }                                                                                                                       \
        chain_sweep_kernel<ST><<<S->n_chains, 32, smem, st>>>(S->n_chains, S->chain_ptr, S->blocks, S->nsz, S->G, S->DH, S->vp,  \
                                                              S->c, S->xp, sd);                                                 \
    } while (0)
    if (S->stages == 8) EHDG_SWEEP(8);
    else if (S->stages == 4) EHDG_SWEEP(4);
    else EHDG_SWEEP(2);
#undef EHDG_SWEEP
    chain_scatter_kernel<<<(nd + 255) / 256, 256, 0, st>>>(nd, S->dof_rows, S->xp, row0, t);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

void chain_free(ChainSolver* S) {
    if (!S) return;
    cudaFree(S->chain_ptr);
    cudaFree(S->blocks);
    cudaFree(S->dof_rows);
    cudaFree(S->dof_blk);
    cudaFree(S->blk_chain);
    cudaFree(S->sblk);
    cudaFree(S->sloc);
    cudaFree(S->G);
    cudaFree(S->DH);
    cudaFree(S->nsz);
    cudaFree(S->c);
    delete S;
}

}  // namespace ehdg
