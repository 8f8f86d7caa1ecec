// Strip block of the block-Jacobi preconditioner (see ehdg_strip.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct ehdg_csr;

namespace ehdg {

struct StripSolver {
    int n_dofs = 0, n_blocks = 0, dropped = 0;
    int64_t n_rows = 0;
    int32_t *blk_ptr = nullptr, *dof_rows = nullptr, *dof_blk = nullptr, *sblk = nullptr, *sloc = nullptr;
    int64_t *offD = nullptr, *offL = nullptr, *offU = nullptr;
    double *D = nullptr, *G = nullptr, *H = nullptr, *c = nullptr;
};

// vals: the ORIGINAL matrix values (before the block-Jacobi fold); rowdrop u8[n_rows]: dofs found
// dependent inside their facet block.  *out stays null when there is no strip.
int strip_setup(ehdg_csr* A, const double* vals, const uint8_t* rowdrop, cudaStream_t st, StripSolver** out);
// t[strip rows] = A_SS^-1 v[strip rows]   (other entries of t untouched)
int strip_apply(const StripSolver* S, const double* v, double* t, cudaStream_t st);
void strip_free(StripSolver* S);

}  // namespace ehdg
