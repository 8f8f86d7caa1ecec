// Line-block preconditioner of the condensed facet system (see ehdg_chain.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct ehdg_csr;

namespace ehdg {

// per-block metadata, read one step ahead of the recurrence (32 bytes, one broadcast load)
struct ChainBlock {
    int32_t s0;       // first chain dof of the block
    int32_t n;        // dofs of the block
    int64_t offG;     // G_b  (n x n_prev, column major)          in ChainSolver::G
    int64_t offDH;    // [Delta_b^-1 (n x n) | H_b (n x n_next)]  in ChainSolver::DH, column major
    int64_t pad;
};

struct ChainSolver {
    int n_dofs = 0 /* chain positions incl. padding to even block sizes */, n_blocks = 0, n_chains = 0, dropped = 0, nmax = 0;
    int all_rows = 0;              // every row of the (rank's) system lies in a chain
    int64_t n_rows = 0;
    int64_t factor_bytes = 0;      // bytes of G + DH (what one application streams)
    int stage_bytes = 0, stages = 0;
    int32_t *chain_ptr = nullptr;  // [n_chains+1] block range of every chain
    ChainBlock* blocks = nullptr;  // [n_blocks+1]
    int32_t *dof_rows = nullptr, *dof_blk = nullptr, *sblk = nullptr, *sloc = nullptr, *blk_chain = nullptr;
    uint8_t* nsz = nullptr;        // [n_blocks] dofs of every block (the only metadata the sweep reads per block)
    double *G = nullptr, *DH = nullptr, *c = nullptr, *vp = nullptr, *xp = nullptr;   // vp, xp: vectors in chain order (inside c's allocation)
};

// Chains of facets: facet f belongs to chain line[f] (host array of nf ints; < 0: the facet forms a chain of its own);
// strip facets (A->strip) form one extra chain when use_strip != 0.  line == nullptr: only the strip chain.
// Only facets inside the matrix window [f_lo, f_hi) take part.  vals: the ORIGINAL matrix values (before any fold);
// rowdrop u8[n_rows] (optional, in / out): dofs found dependent inside their facet block on entry; the dofs whose pivot in
// the chain factorisation falls below drop_tol of their block's largest diagonal entry are added.  *out stays null when there
// is no chain.
// max_blocks > 0: chains longer than that are cut into equal segments (more, shorter recurrences).
int chain_setup(ehdg_csr* A, const double* vals, uint8_t* rowdrop, const int32_t* line_dev, int use_strip, int target_dofs,
                double drop_tol, int max_blocks, cudaStream_t st, ChainSolver** out);
// t[chain rows] = M_chain^-1 (vscale .* v)[chain rows]   (other entries of t untouched); v, vscale, t indexed from row `row0`
int chain_apply(const ChainSolver* S, const double* v, double* t, int64_t row0, cudaStream_t st, const double* vscale = nullptr);
void chain_free(ChainSolver* S);

}  // namespace ehdg
