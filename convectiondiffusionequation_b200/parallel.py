"""Row partition of the condensed facet system for the multi-GPU solve (one process per GPU).

The facet-major numbering follows the facets' vertex pairs, so on the (jittered) lattice meshes the matrix is
banded: a contiguous block of rows references columns of the two neighbouring blocks at most.  Each rank
iterates on its own rows; per SpMV it receives the halo rows from its two neighbours (ncclSend/Recv inside
libehdg) and every batch of dot products is one ncclAllReduce.  This module holds the host logic: splitting
the rows at facet boundaries and finding each rank's halo extent.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def split_rows_at_facets(row_start: np.ndarray, nranks: int) -> np.ndarray:
    """row_begin[nranks+1]: rows split into nranks nearly equal contiguous ranges that never cut a facet block.
    row_start[f] = first row of facet f, row_start[nf] = number of rows."""
    row_start = np.asarray(row_start, dtype=np.int64)
    n = int(row_start[-1])
    targets = (np.arange(1, nranks, dtype=np.int64) * n) // nranks
    cut_facets = np.searchsorted(row_start, targets, side="left")
    row_begin = np.concatenate([[0], row_start[cut_facets], [n]]).astype(np.int64)
    return np.maximum.accumulate(row_begin)


def halo_extents(rowptr, colind, row_begin) -> tuple[np.ndarray, np.ndarray]:
    """(col_lo, col_hi)[nranks]: smallest column and one past the largest column referenced by each rank's rows.
    rowptr / colind may be numpy arrays or torch tensors (any device)."""
    lo, hi = [], []
    for r in range(len(row_begin) - 1):
        a, b = int(row_begin[r]), int(row_begin[r + 1])
        if b <= a:
            lo.append(a)
            hi.append(a)
            continue
        z0, z1 = int(rowptr[a]), int(rowptr[b])
        seg = colind[z0:z1]
        lo.append(min(int(seg.min()), a))
        hi.append(max(int(seg.max()) + 1, b))
    return np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64)


def check_neighbour_halo(row_begin, col_lo, col_hi) -> bool:
    """True when every rank's halo lies inside the rows of rank-1 and rank+1 (what ehdg_comm_set_partition requires)."""
    nr = len(row_begin) - 1
    for r in range(nr):
        lo_ok = row_begin[r - 1] if r > 0 else row_begin[0]
        hi_ok = row_begin[r + 2] if r + 1 < nr else row_begin[nr]
        if col_lo[r] < lo_ok or col_hi[r] > hi_ok:
            return False
    return True


def init_communicator(pipe, group=None):
    """Creates the NCCL communicator of `pipe`'s context over torch.distributed's default (or given) group:
    rank 0 draws the ncclUniqueId, torch broadcasts its 128 bytes."""
    import torch
    import torch.distributed as dist
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(pipe.lib.ehdg_comm_unique_id(buf), "ehdg_comm_unique_id")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dev = pipe.dev if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = t.to(dev)
    dist.broadcast(t, src=0, group=group)
    raw = bytes(t.cpu().numpy().tobytes())
    _lib.check(pipe.lib.ehdg_comm_init(pipe.ctx, raw, world, rank), "ehdg_comm_init")
    return rank, world


def set_partition(pipe, nranks):
    """Splits the rows of pipe's CSR system over nranks and hands the partition to the library."""
    from . import _lib
    from .hot_path import _device_view
    import torch
    n, nnz, nf = pipe.csr_sizes.n_rows, pipe.csr_sizes.nnz, pipe.mesh_host.nf
    rs = _device_view(pipe.lib.ehdg_csr_row_start(pipe.csr), nf + 1, torch.int32, pipe.dev).cpu().numpy()
    row_begin = split_rows_at_facets(rs, nranks)
    rp = _device_view(pipe.lib.ehdg_csr_rowptr(pipe.csr), n + 1, torch.int64, pipe.dev)
    ci = _device_view(pipe.lib.ehdg_csr_colind(pipe.csr), max(nnz, 1), torch.int32, pipe.dev)
    col_lo, col_hi = halo_extents(rp, ci, row_begin)
    if not check_neighbour_halo(row_begin, col_lo, col_hi):
        raise _lib.EhdgError("the halo of a rank reaches beyond its neighbouring ranks: too many ranks for this mesh")
    p64 = lambda a: a.ctypes.data_as(C.c_void_p)
    rb, lo, hi = (np.ascontiguousarray(a, dtype=np.int64) for a in (row_begin, col_lo, col_hi))
    _lib.check(pipe.lib.ehdg_comm_set_partition(pipe.ctx, p64(rb), p64(lo), p64(hi)), "ehdg_comm_set_partition")
    return rb, lo, hi
