// Fused a5-a8: the element kernels store their condensed blocks straight into the facet-major value array of the pattern.
//
// The separate numeric fill (ehdg_csr_numeric) reads every S_K back from HBM and writes the values: 13.9 GB and 9.2 ms at
// 4M elements, a fifth of the step.  It is not needed: in the condensed matrix the block of a facet pair (f, g), f != g, has
// exactly ONE contributing element (two facets share at most one triangle), so its owner can write it in place; only the
// diagonal blocks (f, f) and the right-hand side have two contributors.  There the first element of facet2el writes in
// place, the second one into a side buffer, and a merge kernel adds first + second -- the ascending-element order of the
// gather, so the value array is identical to ehdg_csr_numeric's (tests/test_gpu_fused.py) and no atomics are involved.
#include "ehdg_fast.h"
#include "ehdg_internal.h"

namespace ehdg {

// one thread per element: where its block goes
__global__ void elmap_kernel(int ne, int f_lo, int f_hi, const int32_t* __restrict__ el2facet, const int32_t* __restrict__ facet2el,
                             const int32_t* __restrict__ row_start, const int64_t* __restrict__ blk_base,
                             const int32_t* __restrict__ nbr, const int32_t* __restrict__ nbr_len,
                             const uint16_t* __restrict__ dofmask, ElMap* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    ElMap m;
    int f[3];
    for (int a = 0; a < 3; ++a) f[a] = el2facet[3 * e + a];
    m.hi = 0;
    m.pad[0] = m.pad[1] = 0;
    for (int a = 0; a < 3; ++a) {
        const int fa = f[a];
        const int r0 = row_start[fa], rows = row_start[fa + 1] - r0;
        const bool mine = rows > 0 && fa >= f_lo && fa < f_hi;
        m.base[a] = mine ? blk_base[fa] : -1;
        m.row0[a] = r0;
        m.len[a] = mine ? nbr_len[fa] : 0;
        if (facet2el[2 * fa + 1] == e) m.hi |= (uint8_t)(1u << a);
        for (int b = 0; b < 3; ++b) {
            int off = 0, found = 255;
            for (int t = 0; t < 5; ++t) {
                const int g = nbr[5 * fa + t];
                if (g < 0) break;
                if (g == f[b]) { found = dofmask[g] ? off : 255; break; }
                off += __popc((unsigned)dofmask[g]);
            }
            m.coff[a][b] = (uint8_t)found;
        }
    }
    out[e] = m;
}

// general-path elements: one thread per (element, local facet slot row) of the element-wise S_K the quadrature kernel wrote
__global__ void scatter_general_kernel(int n_gen, int nb, const int32_t* __restrict__ list_gen, const int32_t* __restrict__ el2facet,
                                       const uint16_t* __restrict__ dofmask, const ElMap* __restrict__ em,
                                       const double* __restrict__ Sg, const double* __restrict__ gg, CsrOut co) {
    const int nfg = 3 * nb;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_gen * nfg) return;
    const int idx = (int)(t / nfg), rr = (int)(t - (int64_t)idx * nfg);
    const int e = list_gen[idx];
    const int a = rr / nb, k = rr - a * nb;
    const ElMap m = em[e];
    if (m.base[a] < 0) return;
    const unsigned mf = dofmask[el2facet[3 * e + a]];
    if (!((mf >> k) & 1u)) return;
    const int i = __popc(mf & ((1u << k) - 1u));                    // row of slot k inside the facet's row block
    const double* srow = Sg + (size_t)idx * nfg * nfg + (size_t)rr * nfg;
    const bool hi = (m.hi >> a) & 1;
    for (int b = 0; b < 3; ++b) {
        const int cf = m.coff[a][b];
        if (cf == 255) continue;
        unsigned mg = dofmask[el2facet[3 * e + b]];
        double* dst = (a == b && hi) ? co.diag2 + (size_t)(m.row0[a] + i - co.row_lo) * co.nb
                                     : co.vals + m.base[a] + (int64_t)i * m.len[a] + cf;
        int j = 0;
        while (mg) {
            const int l = __ffs(mg) - 1;
            mg &= mg - 1;
            dst[j++] = srow[b * nb + l];
        }
    }
    const double gval = gg[(size_t)idx * nfg + rr];
    if (hi) co.g2[m.row0[a] + i - co.row_lo] = gval;
    else co.rhs[m.row0[a] + i] = gval;
    // in-kernel merge protocol (CsrOut::flags): this element's share of facet a is in memory once the kernel has finished, and
    // the polynomial kernel that reads the counter starts after it -- where in this kernel the count happens does not matter
    if (co.flags && k == 0) atomicAdd(co.flags + (m.row0[a] - co.row_lo), 1);
}

// facets whose TWO elements are on the quadrature path: nobody merges them in the polynomial kernel.  One thread per
// (general element, local facet); the second element of the facet does the work.
__global__ void merge_gengen_kernel(int n_gen, int nb, const int32_t* __restrict__ list_gen, const int32_t* __restrict__ el2facet,
                                    const int32_t* __restrict__ facet2el, const int32_t* __restrict__ eslot,
                                    const uint16_t* __restrict__ dofmask, const ElMap* __restrict__ em, CsrOut co) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_gen) return;
    const int e = list_gen[t / 3], a = t % 3;
    const ElMap m = em[e];
    if (m.base[a] < 0 || !((m.hi >> a) & 1)) return;
    const int f = el2facet[3 * e + a];
    const int other = facet2el[2 * f];
    if (!(eslot[other] & EHDG_SLOT_GEN)) return;
    const int rows = __popc((unsigned)dofmask[f]);
    const int64_t r0 = m.row0[a] - co.row_lo;
    double* v = co.vals + m.base[a] + m.coff[a][a];
    for (int i = 0; i < rows; ++i) {
        for (int c = 0; c < rows; ++c) v[(int64_t)i * m.len[a] + c] += co.diag2[(r0 + i) * co.nb + c];
        co.rhs[m.row0[a] + i] += co.g2[r0 + i];
    }
}

// column offset of a facet's own (diagonal) block inside its rows; one thread per facet
__global__ void diag_off_kernel(int nf, const int32_t* __restrict__ nbr, const uint16_t* __restrict__ dofmask, int32_t* __restrict__ doff) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    int off = 0;
    for (int t = 0; t < 5; ++t) {
        const int g = nbr[5 * f + t];
        if (g < 0 || g == f) break;
        off += __popc((unsigned)dofmask[g]);
    }
    doff[f] = off;
}

// per own row: where its piece of the diagonal block starts in the value array, and how many entries it has (top byte)
__global__ void diag_dst_kernel(int64_t row_lo, int64_t row_hi, const int32_t* __restrict__ row_facet, const int32_t* __restrict__ row_start,
                                const int64_t* __restrict__ rowptr, const int32_t* __restrict__ doff, int64_t* __restrict__ dst) {
    const int64_t r = row_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= row_hi) return;
    const int f = row_facet[r];
    const int64_t rows = row_start[f + 1] - row_start[f];
    dst[r - row_lo] = (rowptr[r] + doff[f]) | (rows << 56);
}

// first + second contribution of the diagonal blocks and of the right-hand side; one thread per own row, one 8-byte index
// load per row (diag_dst), the side buffer read contiguously
__global__ void merge_diag_kernel(int64_t row_lo, int64_t n_own, int nb, const int64_t* __restrict__ dst,
                                  const double* __restrict__ diag2, const double* __restrict__ g2, double* __restrict__ vals,
                                  double* __restrict__ rhs) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_own) return;
    const int64_t d = dst[t];
    const int rows = (int)(d >> 56);
    double* v = vals + (d & (((int64_t)1 << 56) - 1));
    const double* src = diag2 + t * nb;
    for (int l = 0; l < rows; ++l) v[l] += src[l];
    rhs[row_lo + t] += g2[t];
}

// scatter table + diagonal offsets of a pattern (built with the pattern by ehdg_csr_symbolic, on the pattern's stream)
int build_elmap(ehdg_csr* A, const ehdg_mesh* m, cudaStream_t st) {
    const int fa = A->windowed ? A->f_lo : 0, fb = A->windowed ? A->f_hi : A->nf;
    if (A->elmap) EHDG_CUDA(cudaFreeAsync(A->elmap, st));
    if (A->diag_off) EHDG_CUDA(cudaFreeAsync(A->diag_off, st));
    A->elmap = nullptr;
    A->diag_off = nullptr;
    EHDG_CUDA(cudaMallocAsync(&A->elmap, sizeof(ElMap) * (size_t)(m->ne > 0 ? m->ne : 1), st));
    EHDG_CUDA(cudaMallocAsync(&A->diag_off, sizeof(int32_t) * (size_t)(A->nf > 0 ? A->nf : 1), st));
    A->elmap_ne = m->ne;
    if (m->ne > 0)
        elmap_kernel<<<(m->ne + 127) / 128, 128, 0, st>>>(m->ne, fa, fb, m->el2facet, m->facet2el, A->row_start, A->blk_base, A->nbr,
                                                          A->nbr_len, A->dofmask, (ElMap*)A->elmap);
    EHDG_LAUNCH_CHECK();
    if (A->nf > 0) diag_off_kernel<<<(A->nf + 255) / 256, 256, 0, st>>>(A->nf, A->nbr, A->dofmask, A->diag_off);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

// needs rowptr / row_facet: called once the pattern is complete (first fused store)
int build_diag_dst(ehdg_csr* A, cudaStream_t st) {
    const int64_t rlo = A->windowed ? A->row_lo : 0, rhi = A->windowed ? A->row_hi : A->n_rows;
    const int64_t nown = rhi > rlo ? rhi - rlo : 0;
    if (A->diag_dst) EHDG_CUDA(cudaFreeAsync(A->diag_dst, st));
    A->diag_dst = nullptr;
    EHDG_CUDA(cudaMallocAsync(&A->diag_dst, sizeof(int64_t) * (size_t)(nown > 0 ? nown : 1), st));
    if (nown > 0)
        diag_dst_kernel<<<(unsigned)((nown + 255) / 256), 256, 0, st>>>(rlo, rhi, A->row_facet, A->row_start, A->rowptr, A->diag_off, A->diag_dst);
    EHDG_LAUNCH_CHECK();
    return EHDG_OK;
}

int general_assemble(ehdg_ctx* c, const ehdg_plan* p, const ehdg_mesh* m, const double* lam, const uint8_t* elem_active,
                     const uint8_t* facet_mask, double* Sg, double* gg, double* Wg, double* zg, int32_t* status, double* Afull,
                     cudaStream_t st);   // ehdg_core.cu

}  // namespace ehdg

using namespace ehdg;

extern "C" int ehdg_assemble_condense_csr(ehdg_ctx* c, const ehdg_plan* p, ehdg_csr* A, const ehdg_mesh* m, const double* lam,
                                          const uint8_t* elem_active, const uint8_t* facet_mask, double* vals, double* rhs,
                                          double* W, double* z, int32_t* status, void* stream) {
    if (!c || !p || !A || !m || !lam || !vals || !rhs || !W || !z) { set_error("ehdg_assemble_condense_csr: null argument"); return EHDG_ERR_ARG; }
    if (A->edg) { set_error("ehdg_assemble_condense_csr: the pattern is an EDG pattern"); return EHDG_ERR_ARG; }
    if (A->nf != m->nf) { set_error("ehdg_assemble_condense_csr: pattern and mesh do not match"); return EHDG_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rlo = A->windowed ? A->row_lo : 0, rhi = A->windowed ? A->row_hi : A->n_rows;
    if (A->folded_vals == vals) A->folded_vals = nullptr;
    if (!A->elmap || A->elmap_ne != m->ne) {
        const int rcm = build_elmap(A, m, st);
        if (rcm) return rcm;
    }
    if (!A->diag_dst) {
        const int rcd = build_diag_dst(A, st);
        if (rcd) return rcd;
    }
    const int nb = c->pb.nb;
    const int64_t nown = rhi > rlo ? rhi - rlo : 0;
    double *diag2 = nullptr, *g2 = nullptr, *Sg = nullptr, *gg = nullptr;
    // the two scratch buffers go back to the pool on every path out of this function
    struct AsyncFree { double*& p; cudaStream_t st; ~AsyncFree() { if (p) cudaFreeAsync(p, st); } };
    AsyncFree free_diag2{diag2, st}, free_sg{Sg, st};
    const bool fork = p->n_gen > 0 && p->n_poly > 0 && c->overlap_general;
    // Merge of the two-contributor entries: by merge_diag_kernel afterwards (default) or inside the element kernel
    // (EHDG_INKERNEL_MERGE=1; fence + arrival counter per facet, the second arrival adds the two shares in place).  A/B on
    // B200, 4M elements (profiles/r02_ab_inkernel_merge.txt): the in-kernel variant is bit-identical and takes the element
    // kernel from 17.3 to 25.5 ms -- the fences and the L2-only read-modify-writes cost 8 ms where the merge kernel costs 2.3.
    static const bool want_inkernel = [] { const char* e = getenv("EHDG_INKERNEL_MERGE"); return e && e[0] == '1'; }();
    const bool inkernel = want_inkernel && !fork;
    // side buffer [own rows][nb] + [own rows], then (in-kernel merge) the arrival counters, one int per own row
    const size_t side_doubles = (size_t)(nown * (nb + 1) + 1);
    const size_t flag_bytes = inkernel ? sizeof(int) * (size_t)(nown + 1) : 0;
    EHDG_CUDA(cudaMallocAsync(&diag2, sizeof(double) * side_doubles + flag_bytes, st));
    g2 = diag2 + nown * nb;
    int* flags = inkernel ? (int*)(diag2 + side_doubles) : nullptr;
    if (inkernel) EHDG_CUDA(cudaMemsetAsync(flags, 0, flag_bytes, st));
    else EHDG_CUDA(cudaMemsetAsync(diag2, 0, sizeof(double) * side_doubles, st));
    CsrOut co{(const ElMap*)A->elmap, vals, rhs, diag2, g2, rlo, nb, flags};
    int rc = EHDG_OK;
    cudaStream_t sg = fork ? c->side : st;
    if (fork) {
        EHDG_CUDA(cudaEventRecord(c->ev_fork, st));
        EHDG_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    }
    if (p->n_gen > 0) {
        const size_t nfg = (size_t)p->nfmax;
        if (cudaMallocAsync(&Sg, sizeof(double) * (size_t)p->n_gen * (nfg * nfg + nfg), st) != cudaSuccess) {
            Sg = nullptr;
            set_error("ehdg_assemble_condense_csr: out of memory for the element blocks of the quadrature path");
            return EHDG_ERR_CUDA;
        }
        gg = Sg + (size_t)p->n_gen * nfg * nfg;
        rc = general_assemble(c, p, m, lam, elem_active, facet_mask, Sg, gg, W + (size_t)p->n_poly * p->n_p * p->nf_poly,
                              z + (size_t)p->n_poly * p->n_p, status, nullptr, sg);
        if (rc == EHDG_OK) {
            const int64_t nthr = p->n_gen * (int64_t)nfg;
            scatter_general_kernel<<<(unsigned)((nthr + 127) / 128), 128, 0, sg>>>((int)p->n_gen, nb, p->list_gen, m->el2facet, A->dofmask,
                                                                                   (const ElMap*)A->elmap, Sg, gg, co);
            if (cudaGetLastError() != cudaSuccess) rc = EHDG_ERR_CUDA;
            if (rc == EHDG_OK && inkernel) {
                merge_gengen_kernel<<<(unsigned)((3 * p->n_gen + 127) / 128), 128, 0, sg>>>((int)p->n_gen, nb, p->list_gen, m->el2facet, m->facet2el,
                                                                                        p->eslot, A->dofmask, (const ElMap*)A->elmap, co);
                if (cudaGetLastError() != cudaSuccess) rc = EHDG_ERR_CUDA;
            }
        }
    }
    if (rc == EHDG_OK && p->n_poly > 0) rc = fast_assemble(c, p, m, lam, nullptr, nullptr, W, z, status, st, &co);
    if (fork) {
        cudaEventRecord(c->ev_join, c->side);
        cudaStreamWaitEvent(st, c->ev_join, 0);
    }
    if (rc == EHDG_OK && nown > 0 && !inkernel) {
        merge_diag_kernel<<<(unsigned)((nown + 255) / 256), 256, 0, st>>>(rlo, nown, nb, A->diag_dst, diag2, g2, vals, rhs);
        if (cudaGetLastError() != cudaSuccess) { set_error("ehdg_assemble_condense_csr: merge launch failed"); rc = EHDG_ERR_CUDA; }
    }
    return rc;
}
