// libehdg: the EDG path (enriched interior-penalty DG), reference Python/convection_diffusion.py:42-219.
//   pruning   :83-123   mass Gram of the enriched volume shapes, threshold config['theta']
//   alpha     :129-163  lambda_max of the pencil (m, a + f f^T) with the bonus rule            (ehdg_edg_alpha)
//   forms     :165-198  SIP diffusion + upwind convection over the skeleton, one block row per element (ehdg_edg_assemble)
//   error     :205-210  recombination u_h = u_0 + sum u_i psi_i, L2 error with the order 5 + bonus rule
// The matrix is element-major: block row of element K = [K | its facet neighbours] in ascending element number, rows =
// active volume slots of K (polynomial modes, then the enrichments that survived the pruning).  Every interior facet is
// integrated from both sides, each side writing its own block row: no atomics, a deterministic result.  The pattern
// (ehdg_edg_csr_symbolic, in ehdg_solve.cu) uses the container of the condensed facet system with "facet" := element, so the
// Krylov solvers, the line-block preconditioner and the SpMV of the HDG path serve it unchanged.
#include "edg_element.h"
#include "ehdg_internal.h"

namespace ehdg {

constexpr int EDG_THREADS = 128;

// i-th active volume slot of an element whose enrichment bits are `bits`
__device__ __forceinline__ int edg_active_slot(int n_p, int bits, int i) {
    if (i < n_p) return i;
    int m = bits;
    for (int t = n_p; t < i; ++t) m &= m - 1;
    return n_p + __ffs(m) - 1;
}

__global__ void __launch_bounds__(EDG_THREADS)
edg_prune_kernel(Problem pb, RuleSet rs, int nlist, const int32_t* __restrict__ list, const double* __restrict__ verts,
                 const int32_t* __restrict__ tris, const uint8_t* __restrict__ elem_mask, double theta,
                 uint8_t* __restrict__ elem_active, double* __restrict__ factors) {
    extern __shared__ double ws[];
    if ((int)blockIdx.x >= nlist) return;
    const int el = list ? list[blockIdx.x] : (int)blockIdx.x;
    const int elmark = elem_mask[el];
    if (elmark == 0) return;                                  // only marked elements carry enrichment dofs (:52-54)
    Coop co{(int)threadIdx.x, EDG_THREADS};
    double fac[EHDG_MAX_ENRICH];
    const int drop = edg_prune_element(co, pb, rs, verts, tris + 3 * el, elmark, theta, ws, fac);
    if (threadIdx.x == 0) {
        elem_active[el] = (uint8_t)(elmark & ~drop);
        if (factors)
            for (int k = 0; k < pb.nenr; ++k) factors[(size_t)el * pb.nenr + k] = fac[k];
    }
}

__global__ void edg_prune_init_kernel(int ne, int nenr, const uint8_t* __restrict__ elem_mask, uint8_t* __restrict__ elem_active,
                                      double* __restrict__ factors) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= ne) return;
    elem_active[el] = elem_mask[el];
    if (factors)
        for (int k = 0; k < nenr; ++k) factors[(size_t)el * nenr + k] = -1.0;
}

// One CTA per element: the block row (D, O[3], F) by quadrature, then its values into the pattern.
__global__ void __launch_bounds__(EDG_THREADS)
edg_assemble_kernel(Problem pb, RuleSet rs, int ne, const double* __restrict__ verts, const int32_t* __restrict__ tris,
                    const int32_t* __restrict__ el2facet, const int32_t* __restrict__ facet2el,
                    const uint8_t* __restrict__ elem_active, const double* __restrict__ alpha_stab,
                    const int32_t* __restrict__ row_start, const int64_t* __restrict__ blk_base,
                    const int32_t* __restrict__ nbr_len, int ws_doubles, double* __restrict__ vals, double* __restrict__ rhs,
                    int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int el = blockIdx.x;
    if (el >= ne) return;
    const int nv = pb.nvmax, n_p = pb.n_p;
    double* ws = sm;
    double* D = ws + ws_doubles;
    double* O = D + nv * nv;
    double* F = O + 3 * nv * nv;
    Coop co{(int)threadIdx.x, EDG_THREADS};
    edg_element_blocks(co, pb, rs, verts, tris, el2facet, facet2el, elem_active, alpha_stab, el, ws, D, O, F);
    __syncthreads();
    int nbr[4], edge[4];
    const int n_nbr = edg_block_neighbours(el2facet, facet2el, el, nbr, edge);
    const int own_bits = elem_active ? elem_active[el] : 0;
    const int rows = n_p + __popc(own_bits);
    const int len = nbr_len[el];
    const int64_t base = blk_base[el];
    const int r0 = row_start[el];
    int off = 0, bad = 0;
    for (int t = 0; t < n_nbr; ++t) {
        const int bits = elem_active ? elem_active[nbr[t]] : 0;
        const int cols = n_p + __popc(bits);
        const double* B = edge[t] < 0 ? D : O + (size_t)edge[t] * nv * nv;
        for (int idx = threadIdx.x; idx < rows * cols; idx += EDG_THREADS) {
            const int i = idx / cols, l = idx % cols;
            const double v = B[edg_active_slot(n_p, own_bits, i) * nv + edg_active_slot(n_p, bits, l)];
            if (v != v) bad = EHDG_ST_NONFINITE;
            vals[base + (int64_t)i * len + off + l] = v;
        }
        off += cols;
    }
    for (int i = threadIdx.x; i < rows; i += EDG_THREADS) rhs[r0 + i] = F[edg_active_slot(n_p, own_bits, i)];
    bad = block_or_bits(bad);
    if (threadIdx.x == 0 && status) status[el] |= bad;
}

// u [ne][nvmax] (inactive slots 0) from the solution vector in row order
__global__ void edg_unpack_kernel(int ne, int n_p, int nv, const int32_t* __restrict__ row_start, const uint8_t* __restrict__ elem_active,
                                  const double* __restrict__ x, double* __restrict__ u) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int el = (int)(t / nv), k = (int)(t % nv);
    if (el >= ne) return;
    const int bits = elem_active ? elem_active[el] : 0;
    double v = 0.0;
    if (k < n_p) v = x[row_start[el] + k];
    else if ((bits >> (k - n_p)) & 1) v = x[row_start[el] + n_p + __popc(bits & ((1 << (k - n_p)) - 1))];
    u[(size_t)el * nv + k] = v;
}

// squared L2 / H1-seminorm error of u [ne][nvmax], one warp per element (reference :205-210: order 5 + bonus)
__global__ void __launch_bounds__(128)
edg_error_kernel(Problem pb, RuleSet rs, int ne, const double* __restrict__ verts, const int32_t* __restrict__ tris,
                 const uint8_t* __restrict__ elem_active, const double* __restrict__ u, double* __restrict__ block_out) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int el = blockIdx.x * 4 + wid;
    double e2 = 0.0, h2 = 0.0;
    if (el < ne) {
        ElemGeom g;
        elem_geom(verts, tris + 3 * el, g);
        const int nv = pb.nvmax;
        const double* ul = u + (size_t)el * nv;
        const int volmask = elem_active ? elem_active[el] : 0;
        for (int q = lane; q < rs.nq; q += 32) {
            const double l0 = rs.vx[q], l1 = rs.vy[q], l2 = 1.0 - l0 - l1;
            double ph[EHDG_MAX_NP + EHDG_MAX_ENRICH], gx[EHDG_MAX_NP + EHDG_MAX_ENRICH], gy[EHDG_MAX_NP + EHDG_MAX_ENRICH];
            vol_shapes(pb, g, volmask, l0, l1, ph, gx, gy, 1);
            double uh = 0.0, ux = 0.0, uy = 0.0;
            for (int i = 0; i < nv; ++i)
                if (i < pb.n_p || ((volmask >> (i - pb.n_p)) & 1)) {
                    uh += ul[i] * ph[i];
                    ux += ul[i] * gx[i];
                    uy += ul[i] * gy[i];
                }
            const double dx = l0 * g.omx[0] + l1 * g.omx[1] + l2 * g.omx[2];
            const double dy = l0 * g.omy[0] + l1 * g.omy[1] + l2 * g.omy[2];
            double ex, egx, egy;
            problem_exact(pb, dx, dy, &ex, &egx, &egy);
            const double w = rs.vw[q] * g.adet;
            e2 += w * (uh - ex) * (uh - ex);
            h2 += w * ((ux - egx) * (ux - egx) + (uy - egy) * (uy - egy));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        e2 += __shfl_xor_sync(0xffffffffu, e2, d);
        h2 += __shfl_xor_sync(0xffffffffu, h2, d);
    }
    __shared__ double red[4][2];
    if (lane == 0) { red[wid][0] = e2; red[wid][1] = h2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_out[(size_t)blockIdx.x * 2 + 0] = red[0][0] + red[1][0] + red[2][0] + red[3][0];
        block_out[(size_t)blockIdx.x * 2 + 1] = red[0][1] + red[1][1] + red[2][1] + red[3][1];
    }
}

// out[t] = sum_b partial[b][t], t < 2, fixed order
__global__ void edg_reduce2_kernel(int nblocks, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double red[256];
    const int t = blockIdx.x;
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += partial[(size_t)b * 2 + t];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[t] = red[0];
}

// the reference carries the previous element's value over when the eigenvalue is exactly zero (:158-161): a sequential rule,
// applied only where it can matter
__global__ void edg_alpha_zero_flag_kernel(int ne, const double* __restrict__ lam, int* __restrict__ flag) {
    const int el = blockIdx.x * blockDim.x + threadIdx.x;
    if (el < ne && lam[el] == 0.0) *flag = 1;
}
__global__ void edg_alpha_carry_kernel(int ne, double* __restrict__ lam) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double csr = 0.0;
    for (int el = 0; el < ne; ++el) {
        if (lam[el] != 0.0) csr = lam[el];
        lam[el] = csr;
    }
}

}  // namespace ehdg

using namespace ehdg;

static int edg_ws(const ehdg_ctx* c) { return edg_ws_doubles(c->pb, c->rb.rs.nq, c->rb.rs.nqf); }

extern "C" {

int ehdg_edg_prune(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const uint8_t* elem_mask, double theta,
                   uint8_t* elem_active, double* factors, void* stream) {
    if (!c || !m || !elem_mask || !elem_active) { set_error("ehdg_edg_prune: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (m->ne == 0) return EHDG_OK;
    edg_prune_init_kernel<<<(m->ne + 255) / 256, 256, 0, st>>>(m->ne, c->pb.nenr, elem_mask, elem_active, factors);
    EHDG_LAUNCH_CHECK();
    if (c->pb.nenr == 0) return EHDG_OK;
    const int nlist = p ? (int)p->n_gen : m->ne;              // with a plan: only its general list (marked elements + neighbours)
    if (nlist == 0) return EHDG_OK;
    const int smem = (edg_ws(c) + c->pb.nvmax * c->pb.nvmax) * 8;
    EHDG_CUDA(cudaFuncSetAttribute(edg_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    edg_prune_kernel<<<nlist, EDG_THREADS, smem, st>>>(c->pb, c->rb.rs, nlist, p ? p->list_gen : nullptr, m->verts, m->tris, elem_mask, theta,
                                                       elem_active, factors);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int ehdg_edg_assemble(ehdg_ctx* c, const ehdg_mesh* m, const ehdg_csr* A, const double* alpha_stab, const uint8_t* elem_active,
                      double* vals, double* rhs, int32_t* status, void* stream) {
    if (!c || !m || !A || !alpha_stab || !vals || !rhs) { set_error("ehdg_edg_assemble: null argument"); return EHDG_ERR_ARG; }
    if (!A->edg || A->nf != m->ne) { set_error("ehdg_edg_assemble: the pattern is not an EDG pattern of this mesh"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (m->ne == 0) return EHDG_OK;
    const int nv = c->pb.nvmax, wsd = edg_ws(c);
    const int smem = (wsd + 4 * nv * nv + nv + 2) * 8;
    EHDG_CUDA(cudaFuncSetAttribute(edg_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    edg_assemble_kernel<<<m->ne, EDG_THREADS, smem, st>>>(c->pb, c->rb.rs, m->ne, m->verts, m->tris, m->el2facet, m->facet2el,
                                                          c->pb.nenr ? elem_active : nullptr, alpha_stab, A->row_start, A->blk_base, A->nbr_len,
                                                          wsd, vals, rhs, status);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int ehdg_edg_alpha_carry(ehdg_ctx* c, int32_t ne, double* lam, void* stream) {
    if (!c || !lam) { set_error("ehdg_edg_alpha_carry: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (ne == 0) return EHDG_OK;
    int* flag = nullptr;
    EHDG_CUDA(cudaMallocAsync(&flag, sizeof(int), st));
    cudaMemsetAsync(flag, 0, sizeof(int), st);
    edg_alpha_zero_flag_kernel<<<(ne + 255) / 256, 256, 0, st>>>(ne, lam, flag);
    int h = 0;
    const cudaError_t e1 = cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFreeAsync(flag, st);
    EHDG_CUDA(e1);
    EHDG_CUDA(e2);
    if (h) edg_alpha_carry_kernel<<<1, 32, 0, st>>>(ne, lam);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int ehdg_edg_unpack(ehdg_ctx* c, const ehdg_mesh* m, const ehdg_csr* A, const uint8_t* elem_active, const double* x, double* u,
                    void* stream) {
    if (!c || !m || !A || !x || !u) { set_error("ehdg_edg_unpack: null argument"); return EHDG_ERR_ARG; }
    if (!A->edg || A->nf != m->ne) { set_error("ehdg_edg_unpack: the pattern is not an EDG pattern of this mesh"); return EHDG_ERR_ARG; }
    if (m->ne == 0) return EHDG_OK;
    const int64_t n = (int64_t)m->ne * c->pb.nvmax;
    edg_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->ne, c->pb.n_p, c->pb.nvmax, A->row_start,
                                                                                     c->pb.nenr ? elem_active : nullptr, x, u);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int ehdg_edg_l2_error(ehdg_ctx* c, const ehdg_mesh* m, const uint8_t* elem_active, const double* u, double* err2, void* stream) {
    if (!c || !m || !u || !err2) { set_error("ehdg_edg_l2_error: null argument"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nblocks = (m->ne + 3) / 4;
    if (nblocks == 0) { EHDG_CUDA(cudaMemsetAsync(err2, 0, 16, st)); return EHDG_OK; }
    double* partial = nullptr;
    EHDG_CUDA(cudaMallocAsync(&partial, sizeof(double) * 2 * (size_t)nblocks, st));
    edg_error_kernel<<<nblocks, 128, 0, st>>>(c->pb, c->redg.rs, m->ne, m->verts, m->tris, c->pb.nenr ? elem_active : nullptr, u, partial);
    edg_reduce2_kernel<<<2, 256, 0, st>>>(nblocks, partial, err2);
    const cudaError_t e1 = cudaGetLastError();
    cudaFreeAsync(partial, st);
    EHDG_CUDA(e1);
    return EHDG_OK;
}

}  // extern "C"
