// NCCL plumbing of the distributed facet solve: row-partitioned Krylov vectors, SpMV halo exchange with the
// two neighbouring ranks, dot products reduced with one ncclAllReduce per batch.  libnccl is resolved at run
// time (dlopen of the libnccl.so.2 that torch has already loaded), so libehdg.so has no link-time dependency.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace ehdg {

struct DistComm {
    void* lib = nullptr;
    void* comm = nullptr;
    int rank = 0, nranks = 1;
    std::vector<int64_t> row_begin;          // [nranks+1] first row of each rank
    std::vector<int64_t> col_lo, col_hi;     // [nranks]  column range [lo, hi) referenced by the rows of each rank
    // resolved entry points (signatures of nccl.h, 2.x)
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ ...) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    bool active() const { return comm != nullptr && nranks > 1 && !row_begin.empty(); }
    // in-place sum of `count` doubles over all ranks
    int allreduce_sum(double* buf, size_t count, cudaStream_t st) const;
    // xg is a full-length vector whose rows [row_begin[rank], row_begin[rank+1]) are current: fetch the halo rows
    // width > 1: every row carries `width` doubles (the block-Jacobi inverses)
    // own (optional): the rank's own rows as a separate array (row r at own[(r - row_begin[rank]) * width]); sent from there
    int halo_exchange(double* xg, cudaStream_t st, int width = 1, const double* own = nullptr) const;
    // every rank contributes its own row slice of x; afterwards all ranks hold the whole vector
    int allgather_rows(double* x, cudaStream_t st) const;
};

}  // namespace ehdg
