#include <dlfcn.h>
#include <cstring>

#include "ehdg_dist.h"
#include "ehdg_internal.h"

namespace ehdg {

// nccl.h constants (NCCL 2.x ABI): ncclFloat64 = 8, ncclSum = 0
constexpr int NCCL_F64 = 8;
constexpr int NCCL_SUM = 0;
struct UniqueId { char internal[128]; };

#define EHDG_NCCL(call)                                                                         \
    do {                                                                                        \
        const int r__ = (call);                                                                 \
        if (r__ != 0) {                                                                         \
            set_error("%s:%d %s: NCCL error %d (%s)", __FILE__, __LINE__, #call, r__,            \
                      GetErrorString ? GetErrorString(r__) : "?");                               \
            return EHDG_ERR_NCCL;                                                               \
        }                                                                                       \
    } while (0)

int DistComm::allreduce_sum(double* buf, size_t count, cudaStream_t st) const {
    EHDG_NCCL(AllReduce(buf, buf, count, NCCL_F64, NCCL_SUM, comm, st));
    return EHDG_OK;
}

// Inside an open ncclGroupStart() every call is issued; the first failure is remembered, the group is always closed
// (an early return would leave the thread's group nesting open and no later collective of this communicator would launch).
#define EHDG_NCCL_IN_GROUP(call)                                                                \
    do {                                                                                        \
        const int r__ = (call);                                                                 \
        if (r__ != 0 && first_err == 0) { first_err = r__; first_what = #call; }                 \
    } while (0)

int DistComm::halo_exchange(double* xg, cudaStream_t st, int width, const double* own) const {
    const size_t w = (size_t)width;
    const int64_t r0 = row_begin[rank], r1 = row_begin[rank + 1];
    int first_err = 0;
    const char* first_what = "";
    EHDG_NCCL(GroupStart());
    if (rank > 0) {
        // lower neighbour needs my first rows up to its col_hi; I need its last rows from my col_lo
        const int64_t send_hi = col_hi[rank - 1] < r1 ? col_hi[rank - 1] : r1;
        if (send_hi > r0) EHDG_NCCL_IN_GROUP(Send(own ? own : xg + r0 * w, (size_t)(send_hi - r0) * w, NCCL_F64, rank - 1, comm, st));
        if (col_lo[rank] < r0) EHDG_NCCL_IN_GROUP(Recv(xg + col_lo[rank] * w, (size_t)(r0 - col_lo[rank]) * w, NCCL_F64, rank - 1, comm, st));
    }
    if (rank + 1 < nranks) {
        const int64_t send_lo = col_lo[rank + 1] > r0 ? col_lo[rank + 1] : r0;
        if (send_lo < r1) EHDG_NCCL_IN_GROUP(Send(own ? own + (send_lo - r0) * w : xg + send_lo * w, (size_t)(r1 - send_lo) * w, NCCL_F64, rank + 1, comm, st));
        if (col_hi[rank] > r1) EHDG_NCCL_IN_GROUP(Recv(xg + r1 * w, (size_t)(col_hi[rank] - r1) * w, NCCL_F64, rank + 1, comm, st));
    }
    const int rg = GroupEnd();
    if (first_err == 0 && rg != 0) { first_err = rg; first_what = "GroupEnd()"; }
    if (first_err) {
        set_error("halo_exchange: %s: NCCL error %d (%s)", first_what, first_err, GetErrorString ? GetErrorString(first_err) : "?");
        return EHDG_ERR_NCCL;
    }
    return EHDG_OK;
}

int DistComm::allgather_rows(double* x, cudaStream_t st) const {
    int first_err = 0;
    const char* first_what = "";
    EHDG_NCCL(GroupStart());
    for (int r = 0; r < nranks; ++r) {
        const int64_t cnt = row_begin[r + 1] - row_begin[r];
        if (cnt > 0) EHDG_NCCL_IN_GROUP(Broadcast(x + row_begin[r], x + row_begin[r], (size_t)cnt, NCCL_F64, r, comm, st));
    }
    const int rg = GroupEnd();
    if (first_err == 0 && rg != 0) { first_err = rg; first_what = "GroupEnd()"; }
    if (first_err) {
        set_error("allgather_rows: %s: NCCL error %d (%s)", first_what, first_err, GetErrorString ? GetErrorString(first_err) : "?");
        return EHDG_ERR_NCCL;
    }
    return EHDG_OK;
}

static int load_nccl(DistComm* d) {
    if (d->lib) return EHDG_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        d->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (d->lib) break;
    }
    if (!d->lib) {
        const char* env = getenv("EHDG_NCCL_LIB");
        if (env) d->lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!d->lib) { set_error("cannot load libnccl.so.2 (import torch first or set EHDG_NCCL_LIB): %s", dlerror()); return EHDG_ERR_NCCL; }
#define SYM(field, name)                                                                     \
    *(void**)(&d->field) = dlsym(d->lib, name);                                              \
    if (!d->field) { set_error("libnccl lacks %s", name); return EHDG_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(Broadcast, "ncclBroadcast")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return EHDG_OK;
}

}  // namespace ehdg

using namespace ehdg;

extern "C" {

int ehdg_comm_unique_id(char* out128) {
    if (!out128) { set_error("ehdg_comm_unique_id: null argument"); return EHDG_ERR_ARG; }
    DistComm tmp;
    int rc = load_nccl(&tmp);
    if (rc) return rc;
    UniqueId id;
    const int r = tmp.GetUniqueId(&id);
    dlclose(tmp.lib);                     // the library stays mapped through torch's own handle (and ehdg_comm_init's)
    if (r != 0) { set_error("ncclGetUniqueId failed (%d)", r); return EHDG_ERR_NCCL; }
    memcpy(out128, id.internal, 128);
    return EHDG_OK;
}

int ehdg_comm_init(ehdg_ctx* c, const char* id128, int nranks, int rank) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("ehdg_comm_init: bad argument"); return EHDG_ERR_ARG; }
    if (!c->dist) c->dist = new DistComm();
    DistComm* d = c->dist;
    int rc = load_nccl(d);
    if (rc) return rc;
    if (d->comm) {                        // re-initialisation: the previous communicator is released first
        d->CommDestroy(d->comm);
        d->comm = nullptr;
        d->row_begin.clear();
    }
    UniqueId id;
    memcpy(id.internal, id128, 128);
    typedef int (*init_fn)(void**, int, UniqueId, int);
    const int r = ((init_fn)d->CommInitRank)(&d->comm, nranks, id, rank);
    if (r != 0) { set_error("ncclCommInitRank failed (%d: %s)", r, d->GetErrorString(r)); return EHDG_ERR_NCCL; }
    d->rank = rank;
    d->nranks = nranks;
    return EHDG_OK;
}

int ehdg_comm_set_partition(ehdg_ctx* c, const int64_t* row_begin, const int64_t* col_lo, const int64_t* col_hi) {
    if (!c || !c->dist || !row_begin || !col_lo || !col_hi) { set_error("ehdg_comm_set_partition: communicator not initialised"); return EHDG_ERR_ARG; }
    DistComm* d = c->dist;
    d->row_begin.assign(row_begin, row_begin + d->nranks + 1);
    d->col_lo.assign(col_lo, col_lo + d->nranks);
    d->col_hi.assign(col_hi, col_hi + d->nranks);
    for (int r = 0; r < d->nranks; ++r) {
        // the halo of a rank must lie inside its two neighbours (banded facet numbering)
        const int64_t lo_ok = r > 0 ? d->row_begin[r - 1] : d->row_begin[0];
        const int64_t hi_ok = r + 1 < d->nranks ? d->row_begin[r + 2] : d->row_begin[d->nranks];
        if (d->col_lo[r] < lo_ok || d->col_hi[r] > hi_ok) {
            set_error("rank %d references columns [%lld, %lld) outside its neighbours' rows", r, (long long)d->col_lo[r], (long long)d->col_hi[r]);
            d->row_begin.clear();
            return EHDG_ERR_ARG;
        }
    }
    return EHDG_OK;
}

int ehdg_comm_destroy(ehdg_ctx* c) {
    if (!c || !c->dist) return EHDG_OK;
    if (c->dist->comm) c->dist->CommDestroy(c->dist->comm);
    if (c->dist->lib) dlclose(c->dist->lib);
    delete c->dist;
    c->dist = nullptr;
    return EHDG_OK;
}

}  // extern "C"
