// Exact elimination of the enriched strip inside the block-Jacobi preconditioner.
//
// The facets of the marked elements carry penalties tau ~ eps * 20 lambda_K p^2 / h that are 1e3..1e5
// times those of the bulk (the reference's alpha_K on enriched elements, conv_diff.py:349).  Under facet
// block-Jacobi the modes that are continuous along the strip of marked elements get eigenvalues
// O(1/tau) and GMRES stagnates.  The strip has O(sqrt(ne)) unknowns and is a chain, so it is handled
// as ONE block of the block-Jacobi preconditioner, solved exactly: facets ordered by BFS levels of the
// strip's facet graph (any graph is block tridiagonal in its BFS levels), consecutive levels merged
// into blocks of <= STRIP_MAXB dofs, block LU (block Thomas) factorised once on the device, two
// sequential sweeps of small dense mat-vecs per application.
#include <algorithm>
#include <cub/cub.cuh>
#include <queue>
#include <vector>

#include "ehdg_internal.h"
#include "ehdg_strip.h"

namespace ehdg {

constexpr int STRIP_MAXB = 64;
constexpr int STRIP_TARGET = 40;

__global__ void strip_gather_topo_kernel(int ns, const int32_t* __restrict__ list, const int32_t* __restrict__ nbr,
                                         const int32_t* __restrict__ row_start, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const int f = list[i];
    for (int t = 0; t < 5; ++t) out[7 * i + t] = nbr[5 * f + t];
    out[7 * i + 5] = row_start[f];
    out[7 * i + 6] = row_start[f + 1] - row_start[f];
}

__global__ void strip_pos_kernel(int nd, const int32_t* __restrict__ dof_rows, const int32_t* __restrict__ dof_blk,
                                 const int32_t* __restrict__ blk_ptr, int32_t* __restrict__ sblk, int32_t* __restrict__ sloc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd) return;
    const int b = dof_blk[s];
    sblk[dof_rows[s]] = b;
    sloc[dof_rows[s]] = s - blk_ptr[b];
}

// dense blocks D_b (n_b x n_b), L_b = A[b, b-1] (n_b x n_{b-1}), U_b = A[b, b+1] (n_b x n_{b+1}) from the CSR
__global__ void strip_extract_kernel(int nd, const int32_t* __restrict__ dof_rows, const int32_t* __restrict__ dof_blk,
                                     const int32_t* __restrict__ blk_ptr, const int64_t* __restrict__ rowptr,
                                     const int32_t* __restrict__ colind, const double* __restrict__ vals,
                                     const int32_t* __restrict__ sblk, const int32_t* __restrict__ sloc,
                                     const int64_t* __restrict__ offD, const int64_t* __restrict__ offL,
                                     const int64_t* __restrict__ offU, double* __restrict__ D, double* __restrict__ Lm,
                                     double* __restrict__ Um, int* __restrict__ err) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd) return;
    const int b = dof_blk[s], i = s - blk_ptr[b], r = dof_rows[s];
    const int nb = blk_ptr[b + 1] - blk_ptr[b];
    for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
        const int c = colind[q];
        const int cb = sblk[c];
        if (cb < 0) continue;
        const int j = sloc[c];
        if (cb == b) D[offD[b] + (int64_t)i * nb + j] = vals[q];
        else if (cb == b - 1) Lm[offL[b] + (int64_t)i * (blk_ptr[b] - blk_ptr[b - 1]) + j] = vals[q];
        else if (cb == b + 1) Um[offU[b] + (int64_t)i * (blk_ptr[b + 2] - blk_ptr[b + 1]) + j] = vals[q];
        else if (vals[q] != 0.0) *err = 1;     // not block tridiagonal: cannot happen for BFS levels
    }
}

// Sequential block LU over the chain, one CTA.  On exit: D holds Delta_b^-1 (rank-revealing: dependent
// dofs get zero rows/columns), Lm holds G_b = L_b Delta_{b-1}^-1, Um holds H_b = Delta_b^-1 U_b.
__global__ void __launch_bounds__(256)
strip_factor_kernel(int nblk, const int32_t* __restrict__ blk_ptr, const int32_t* __restrict__ dof_rows,
                    const uint8_t* __restrict__ rowdrop, const int64_t* __restrict__ offD, const int64_t* __restrict__ offL,
                    const int64_t* __restrict__ offU, double* D, double* Lm, double* Um, int* dropped) {
    extern __shared__ double sm[];
    double* Aw = sm;                                   // n x 2n augmented [Delta | I]
    __shared__ int s_piv;
    __shared__ double s_best, s_scale;
    __shared__ unsigned char s_drop[STRIP_MAXB];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int b = 0; b < nblk; ++b) {
        const int n = blk_ptr[b + 1] - blk_ptr[b];
        double* Db = D + offD[b];
        // Delta_b = D_b - G_b U_{b-1}   (G_b already computed at the end of step b-1)
        if (b > 0) {
            const int np_ = blk_ptr[b] - blk_ptr[b - 1];
            const double* G = Lm + offL[b];
            const double* Uprev = Um + offU[b - 1];          // still the original U_{b-1} (np_ x n)
            for (int idx = tid; idx < n * n; idx += nt) {
                const int i = idx / n, j = idx % n;
                double acc = Db[idx];
                for (int k = 0; k < np_; ++k) acc -= G[i * np_ + k] * Uprev[k * n + j];
                Aw[i * 2 * n + j] = acc;
            }
        } else {
            for (int idx = tid; idx < n * n; idx += nt) Aw[(idx / n) * 2 * n + idx % n] = Db[idx];
        }
        for (int idx = tid; idx < n * n; idx += nt) Aw[(idx / n) * 2 * n + n + idx % n] = (idx / n == idx % n) ? 1.0 : 0.0;
        if (tid < n) s_drop[tid] = rowdrop[dof_rows[blk_ptr[b] + tid]];
        __syncthreads();
        // now H_{b-1} = Delta_{b-1}^-1 U_{b-1} may overwrite U_{b-1}: done at the end of step b-1 below (uses a copy)
        // dofs already known to be dependent inside their facet block: decouple
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx % n;
            if (s_drop[i] || s_drop[j]) Aw[i * 2 * n + j] = (i == j) ? 1.0 : 0.0;
        }
        if (tid == 0) {
            double mx = 0.0;
            for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(Aw[i * 2 * n + i]));
            s_scale = mx;
        }
        __syncthreads();
        // Gauss-Jordan with partial pivoting; a pivot below 1e-12 of the largest diagonal entry = dependent dof
        for (int k = 0; k < n; ++k) {
            if (tid == 0) {
                int piv = k;
                double best = fabs(Aw[k * 2 * n + k]);
                for (int i = k + 1; i < n; ++i)
                    if (fabs(Aw[i * 2 * n + k]) > best) { best = fabs(Aw[i * 2 * n + k]); piv = i; }
                s_piv = piv;
                s_best = best;
            }
            __syncthreads();
            if (!(s_best > 1e-12 * s_scale)) {
                // decouple dof k: column k := e_k, row k := e_k, right part row k := 0
                for (int i = tid; i < n; i += nt) Aw[i * 2 * n + k] = (i == k) ? 1.0 : 0.0;
                for (int j = tid; j < 2 * n; j += nt) Aw[k * 2 * n + j] = (j == k) ? 1.0 : 0.0;
                if (tid == 0) { s_drop[k] = 1; atomicAdd(dropped, 1); }
                __syncthreads();
                continue;
            }
            const int piv = s_piv;
            if (piv != k) {
                for (int j = tid; j < 2 * n; j += nt) {
                    const double t = Aw[k * 2 * n + j];
                    Aw[k * 2 * n + j] = Aw[piv * 2 * n + j];
                    Aw[piv * 2 * n + j] = t;
                }
                __syncthreads();
            }
            const double inv = 1.0 / Aw[k * 2 * n + k];
            __syncthreads();
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
        // Delta_b^-1 with dropped rows / columns zeroed
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx % n;
            Db[idx] = (s_drop[i] || s_drop[j]) ? 0.0 : Aw[i * 2 * n + n + j];
        }
        __syncthreads();
        if (b + 1 < nblk) {
            const int nn = blk_ptr[b + 2] - blk_ptr[b + 1];
            // G_{b+1} = L_{b+1} Delta_b^-1  (in place via the scratch Aw)
            double* Lb = Lm + offL[b + 1];
            for (int idx = tid; idx < nn * n; idx += nt) {
                const int i = idx / n, j = idx % n;
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc += Lb[i * n + k] * Db[k * n + j];
                Aw[idx] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < nn * n; idx += nt) Lb[idx] = Aw[idx];
            __syncthreads();
        }
    }
    // H_b = Delta_b^-1 U_b, all blocks (the original U_b were needed unmodified above)
    for (int b = 0; b + 1 < nblk; ++b) {
        const int n = blk_ptr[b + 1] - blk_ptr[b], nn = blk_ptr[b + 2] - blk_ptr[b + 1];
        const double* Db = D + offD[b];
        double* Ub = Um + offU[b];
        for (int idx = tid; idx < n * nn; idx += nt) {
            const int i = idx / nn, j = idx % nn;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += Db[i * n + k] * Ub[k * nn + j];
            Aw[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < n * nn; idx += nt) Ub[idx] = Aw[idx];
        __syncthreads();
    }
}

// t[strip rows] = A_SS^-1 v[strip rows]; one CTA, sequential over the chain
__global__ void __launch_bounds__(128)
strip_solve_kernel(int nblk, const int32_t* __restrict__ blk_ptr, const int32_t* __restrict__ dof_rows,
                   const int64_t* __restrict__ offD, const int64_t* __restrict__ offL, const int64_t* __restrict__ offU,
                   const double* __restrict__ Dinv, const double* __restrict__ G, const double* __restrict__ H,
                   const double* __restrict__ v, double* __restrict__ c, double* __restrict__ t) {
    __shared__ double prev[STRIP_MAXB], cur[STRIP_MAXB];
    const int tid = threadIdx.x, nt = blockDim.x;
    // forward: c_b = v_b - G_b c_{b-1}
    for (int b = 0; b < nblk; ++b) {
        const int s0 = blk_ptr[b], n = blk_ptr[b + 1] - s0;
        const int np_ = b > 0 ? s0 - blk_ptr[b - 1] : 0;
        const double* Gb = G + offL[b];
        for (int i = tid; i < n; i += nt) {
            double acc = v[dof_rows[s0 + i]];
            for (int k = 0; k < np_; ++k) acc -= Gb[i * np_ + k] * prev[k];
            cur[i] = acc;
            c[s0 + i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) prev[i] = cur[i];
        __syncthreads();
    }
    // backward: x_b = Delta_b^-1 c_b - H_b x_{b+1}
    for (int b = nblk - 1; b >= 0; --b) {
        const int s0 = blk_ptr[b], n = blk_ptr[b + 1] - s0;
        const int nn = b + 1 < nblk ? blk_ptr[b + 2] - blk_ptr[b + 1] : 0;
        const double* Db = Dinv + offD[b];
        const double* Hb = H + offU[b];
        for (int i = tid; i < n; i += nt) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += Db[i * n + k] * c[s0 + k];
            for (int k = 0; k < nn; ++k) acc -= Hb[i * nn + k] * prev[k];
            cur[i] = acc;
            t[dof_rows[s0 + i]] = acc;
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) prev[i] = cur[i];
        __syncthreads();
    }
}

struct IsSet {
    const uint8_t* flags;
    __host__ __device__ bool operator()(int i) const { return flags[i] != 0; }
};

int strip_setup(ehdg_csr* A, const double* vals, const uint8_t* rowdrop, cudaStream_t st, StripSolver** out) {
    *out = nullptr;
    const int nf = A->nf;
    if (!A->strip || nf == 0) return EHDG_OK;
    // ---- strip facets, ascending (deterministic)
    int32_t *d_list = nullptr, *d_count = nullptr;
    void* tmp = nullptr;
    size_t tb = 0;
    EHDG_CUDA(cudaMalloc(&d_list, sizeof(int32_t) * (size_t)nf));
    EHDG_CUDA(cudaMalloc(&d_count, sizeof(int32_t)));
    cub::CountingInputIterator<int> ids(0);
    IsSet pred{A->strip};
    EHDG_CUDA(cub::DeviceSelect::If(nullptr, tb, ids, d_list, d_count, nf, pred, st));
    EHDG_CUDA(cudaMalloc(&tmp, tb + 8));
    EHDG_CUDA(cub::DeviceSelect::If(tmp, tb, ids, d_list, d_count, nf, pred, st));
    int32_t ns = 0;
    EHDG_CUDA(cudaMemcpyAsync(&ns, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    cudaFree(d_count);
    if (ns == 0) { cudaFree(d_list); return EHDG_OK; }
    int32_t* d_topo = nullptr;
    EHDG_CUDA(cudaMalloc(&d_topo, sizeof(int32_t) * 7 * (size_t)ns));
    strip_gather_topo_kernel<<<(ns + 127) / 128, 128, 0, st>>>(ns, d_list, A->nbr, A->row_start, d_topo);
    EHDG_LAUNCH_CHECK();
    std::vector<int32_t> list(ns), topo(7 * (size_t)ns);
    EHDG_CUDA(cudaMemcpyAsync(list.data(), d_list, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaMemcpyAsync(topo.data(), d_topo, sizeof(int32_t) * 7 * (size_t)ns, cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_list);
    cudaFree(d_topo);
    // ---- BFS levels over the strip's facet graph (host, O(sqrt(ne)))
    auto local_of = [&](int f) -> int {
        auto it = std::lower_bound(list.begin(), list.end(), f);
        return (it != list.end() && *it == f) ? (int)(it - list.begin()) : -1;
    };
    std::vector<int> level(ns, -1), order;
    order.reserve(ns);
    int lev_base = 0;
    std::vector<int> lev_of_order;
    for (int seed = 0; seed < ns; ++seed) {
        if (level[seed] >= 0) continue;
        std::queue<int> q;
        level[seed] = lev_base;
        q.push(seed);
        int maxlev = lev_base;
        while (!q.empty()) {
            const int u = q.front();
            q.pop();
            order.push_back(u);
            lev_of_order.push_back(level[u]);
            maxlev = std::max(maxlev, level[u]);
            for (int t = 0; t < 5; ++t) {
                const int g = topo[7 * (size_t)u + t];
                if (g < 0) continue;
                const int w = local_of(g);
                if (w >= 0 && level[w] < 0) { level[w] = level[u] + 1; q.push(w); }
            }
        }
        lev_base = maxlev + 2;          // gap: components never share a block boundary coupling
    }
    // ---- merge consecutive levels into blocks
    std::vector<int32_t> blk_ptr{0}, dof_rows, dof_blk;
    {
        size_t i = 0;
        while (i < order.size()) {
            int dofs = 0;
            size_t j = i;
            int last_level = lev_of_order[i];
            while (j < order.size()) {
                // take a whole level at a time
                size_t k = j;
                int ld = 0;
                const int lv = lev_of_order[j];
                while (k < order.size() && lev_of_order[k] == lv) { ld += topo[7 * (size_t)order[k] + 6]; ++k; }
                if (ld > STRIP_MAXB) { set_error("strip solver: BFS level with %d dofs exceeds %d", ld, STRIP_MAXB); return EHDG_ERR_ARG; }
                if (j > i && (dofs + ld > STRIP_TARGET || lv != last_level + 1)) break;
                dofs += ld;
                last_level = lv;
                j = k;
            }
            const int b = (int)blk_ptr.size() - 1;
            for (size_t k = i; k < j; ++k) {
                const int u = order[k];
                for (int d = 0; d < topo[7 * (size_t)u + 6]; ++d) { dof_rows.push_back(topo[7 * (size_t)u + 5] + d); dof_blk.push_back(b); }
            }
            blk_ptr.push_back((int32_t)dof_rows.size());
            i = j;
        }
    }
    const int nblk = (int)blk_ptr.size() - 1, nd = (int)dof_rows.size();
    if (nd == 0) return EHDG_OK;
    std::vector<int64_t> offD(nblk + 1), offL(nblk + 1), offU(nblk + 1);
    int64_t tD = 0, tL = 0, tU = 0;
    for (int b = 0; b < nblk; ++b) {
        const int64_t n = blk_ptr[b + 1] - blk_ptr[b];
        offD[b] = tD; tD += n * n;
        offL[b] = tL; if (b > 0) tL += n * (blk_ptr[b] - blk_ptr[b - 1]);
        offU[b] = tU; if (b + 1 < nblk) tU += n * (blk_ptr[b + 2] - blk_ptr[b + 1]);
    }
    blk_ptr.push_back(blk_ptr.back());            // sentinel for blk_ptr[b + 2]
    StripSolver* S = new StripSolver();
    S->n_dofs = nd;
    S->n_blocks = nblk;
    S->n_rows = A->n_rows;
    auto up = [&](auto** dptr, const auto& h) -> int {
        using T = typename std::remove_reference<decltype(h)>::type::value_type;
        EHDG_CUDA(cudaMalloc(dptr, sizeof(T) * std::max<size_t>(h.size(), 1)));
        EHDG_CUDA(cudaMemcpyAsync(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
        return EHDG_OK;
    };
    int rc;
    if ((rc = up(&S->blk_ptr, blk_ptr))) return rc;
    if ((rc = up(&S->dof_rows, dof_rows))) return rc;
    if ((rc = up(&S->dof_blk, dof_blk))) return rc;
    if ((rc = up(&S->offD, offD))) return rc;
    if ((rc = up(&S->offL, offL))) return rc;
    if ((rc = up(&S->offU, offU))) return rc;
    EHDG_CUDA(cudaMalloc(&S->sblk, sizeof(int32_t) * (size_t)A->n_rows));
    EHDG_CUDA(cudaMalloc(&S->sloc, sizeof(int32_t) * (size_t)A->n_rows));
    EHDG_CUDA(cudaMemsetAsync(S->sblk, 0xff, sizeof(int32_t) * (size_t)A->n_rows, st));
    EHDG_CUDA(cudaMalloc(&S->D, sizeof(double) * (size_t)std::max<int64_t>(tD, 1)));
    EHDG_CUDA(cudaMalloc(&S->G, sizeof(double) * (size_t)std::max<int64_t>(tL, 1)));
    EHDG_CUDA(cudaMalloc(&S->H, sizeof(double) * (size_t)std::max<int64_t>(tU, 1)));
    EHDG_CUDA(cudaMalloc(&S->c, sizeof(double) * (size_t)nd));
    EHDG_CUDA(cudaMemsetAsync(S->D, 0, sizeof(double) * (size_t)std::max<int64_t>(tD, 1), st));
    EHDG_CUDA(cudaMemsetAsync(S->G, 0, sizeof(double) * (size_t)std::max<int64_t>(tL, 1), st));
    EHDG_CUDA(cudaMemsetAsync(S->H, 0, sizeof(double) * (size_t)std::max<int64_t>(tU, 1), st));
    int* d_flags = nullptr;
    EHDG_CUDA(cudaMalloc(&d_flags, 2 * sizeof(int)));
    EHDG_CUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
    strip_pos_kernel<<<(nd + 127) / 128, 128, 0, st>>>(nd, S->dof_rows, S->dof_blk, S->blk_ptr, S->sblk, S->sloc);
    EHDG_LAUNCH_CHECK();
    strip_extract_kernel<<<(nd + 127) / 128, 128, 0, st>>>(nd, S->dof_rows, S->dof_blk, S->blk_ptr, A->rowptr, A->colind, vals, S->sblk,
                                                          S->sloc, S->offD, S->offL, S->offU, S->D, S->G, S->H, d_flags);
    EHDG_LAUNCH_CHECK();
    const int smem = STRIP_MAXB * 2 * STRIP_MAXB * 8;
    EHDG_CUDA(cudaFuncSetAttribute(strip_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    strip_factor_kernel<<<1, 256, smem, st>>>(nblk, S->blk_ptr, S->dof_rows, rowdrop, S->offD, S->offL, S->offU, S->D, S->G, S->H, d_flags + 1);
    EHDG_LAUNCH_CHECK();
    int flags[2] = {0, 0};
    EHDG_CUDA(cudaMemcpyAsync(flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    EHDG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_flags);
    if (flags[0]) { set_error("strip solver: matrix is not block tridiagonal in the BFS blocks"); strip_free(S); return EHDG_ERR_ARG; }
    S->dropped = flags[1];
    *out = S;
    return EHDG_OK;
}

int strip_apply(const StripSolver* S, const double* v, double* t, cudaStream_t st) {
    if (!S) return EHDG_OK;
    strip_solve_kernel<<<1, 128, 0, st>>>(S->n_blocks, S->blk_ptr, S->dof_rows, S->offD, S->offL, S->offU, S->D, S->G, S->H, v, S->c, t);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

void strip_free(StripSolver* S) {
    if (!S) return;
    cudaFree(S->blk_ptr);
    cudaFree(S->dof_rows);
    cudaFree(S->dof_blk);
    cudaFree(S->offD);
    cudaFree(S->offL);
    cudaFree(S->offU);
    cudaFree(S->sblk);
    cudaFree(S->sloc);
    cudaFree(S->D);
    cudaFree(S->G);
    cudaFree(S->H);
    cudaFree(S->c);
    delete S;
}

}  // namespace ehdg
