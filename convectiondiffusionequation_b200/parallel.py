"""Row partition of the condensed facet system for the multi-GPU path (one process per GPU).

Partitioned assembly: a rank owns a contiguous facet range = a contiguous row range of the condensed
system.  It assembles and condenses the elements that touch its facets (its own elements plus one layer of
redundantly computed ghosts, so no S_K crosses GPUs), fills its own CSR rows, and iterates on them; the halo of x
travels per SpMV, the block-Jacobi inverses of the halo facets once, dot products as one all-reduce per batch.
`partitioned_assemble` / `partitioned_solve` below drive that; `set_partition` is the older replicated-matrix mode.

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


# ------------------------------------------------------------------------------------------------------------
# partitioned assembly: element selection, ownership, halo extents
# ------------------------------------------------------------------------------------------------------------
def select_elements(facet2el, f_lo: int, f_hi: int, ne: int):
    """u8 [ne]: 1 for every element adjacent to a facet of [f_lo, f_hi) (torch tensor or numpy array in, same kind out)."""
    import torch
    if isinstance(facet2el, np.ndarray):
        sel = np.zeros(ne, dtype=np.uint8)
        adj = facet2el[f_lo:f_hi].reshape(-1)
        sel[adj[adj >= 0]] = 1
        return sel
    sel = torch.zeros(ne, dtype=torch.uint8, device=facet2el.device)
    adj = facet2el[f_lo:f_hi].reshape(-1).long()
    sel[adj[adj >= 0]] = 1
    return sel


def owned_elements(el2facet, f_lo: int, f_hi: int):
    """u8 [ne]: an element belongs to the rank that owns its lowest-numbered facet (a partition of the elements; the
    owner always has the element selected)."""
    import torch
    if isinstance(el2facet, np.ndarray):
        m = el2facet.min(axis=1)
        return ((m >= f_lo) & (m < f_hi)).astype(np.uint8)
    m = el2facet.min(dim=1).values
    return ((m >= f_lo) & (m < f_hi)).to(torch.uint8)


def gather_extents(lo: int, hi: int, group=None, device="cpu"):
    """all ranks' (col_lo, col_hi) from each rank's own pair"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([int(lo), int(hi)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    arr = torch.stack(out).cpu().numpy()
    return arr[:, 0].copy(), arr[:, 1].copy()


def partitioned_assemble(pipe, rank: int, world: int, group=None, exchange=True, assemble=True, fused=True):
    """mark + prune (replicated: O(sqrt(ne)) work), then this rank's share: pattern of its rows, alpha + element blocks +
    condensation of the elements that touch its facets, numeric fill of its rows, and the communicator's partition.
    exchange=False stops before the collective part (single-process checks of one rank's share); assemble=False stops after
    the per-mesh set-up (marks, pruning, pattern, selection, plan, value arrays) and leaves a4-a8 to the caller; fused=False
    keeps the element blocks S_K and fills the rows with the separate gather."""
    import torch
    import torch.distributed as dist
    from . import _lib
    from .hot_path import _device_view
    import time
    tm = {}

    def lap(name, t0):
        torch.cuda.synchronize(pipe.dev)
        tm[name] = (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t0 = time.perf_counter()
    pipe.setup()                                             # a1, a3 on the full mesh
    t0 = lap("mark_prune_ms", t0)
    pipe.csr_symbolic(world, rank)
    t0 = lap("csr_symbolic_ms", t0)
    f_lo, f_hi = int(pipe.facet_begin[rank]), int(pipe.facet_begin[rank + 1])
    sel, pipe.elem_owned = pipe.select_elements(f_lo, f_hi)        # same rule as select_elements / owned_elements above
    pipe.make_plan(sel)
    t0 = lap("select_plan_ms", t0)
    pipe.csr_alloc_values()
    if assemble and fused and not pipe.dirichlet_lifting:
        pipe.assemble_csr()                                  # a4-a8 on the selection, own rows written by the element kernels
        t0 = lap("alpha_assemble_condense_csr_ms", t0)
    elif assemble:
        pipe.assemble()                                      # a4-a7 on the selection
        t0 = lap("alpha_assemble_condense_ms", t0)
        pipe.csr_numeric()                                   # a8, own rows
        t0 = lap("csr_numeric_ms", t0)
    pipe.part_timing = tm
    rb = pipe.row_begin
    lo, hi = pipe.own_extent                                 # computed on the device with the pattern (ehdg_csr_get_extent)
    if not exchange:
        return pipe
    if world > 1:
        dev = pipe.dev if dist.get_backend(group) == "nccl" else "cpu"
        col_lo, col_hi = gather_extents(lo, hi, group, dev)
    else:
        col_lo, col_hi = np.asarray([lo], dtype=np.int64), np.asarray([hi], dtype=np.int64)
    if not check_neighbour_halo(rb, col_lo, col_hi):
        raise _lib.EhdgError("the halo of a rank reaches beyond its neighbouring ranks: too many ranks for this mesh")
    if world > 1:
        p64 = lambda a: a.ctypes.data_as(C.c_void_p)
        rbc, loc, hic = (np.ascontiguousarray(a, dtype=np.int64) for a in (rb, col_lo, col_hi))
        _lib.check(pipe.lib.ehdg_comm_set_partition(pipe.ctx, p64(rbc), p64(loc), p64(hic)), "ehdg_comm_set_partition")
    pipe.col_lo, pipe.col_hi = col_lo, col_hi
    return pipe


def partitioned_solve(pipe, group=None, **gm):
    """Krylov solve on the partitioned system, back-solve on the rank's elements, (L2, H1) error over all ranks"""
    import torch
    import torch.distributed as dist
    if gm.get("method") == "auto":
        pipe.solve_auto(rtol=gm.get("rtol", 1e-8), max_iters=gm.get("max_iters", 8000), precond=gm.get("precond", "line"),
                        drop_tol=gm.get("drop_tol", 0.0), strip_block=gm.get("strip_block", True))
        pipe.backsolve()
    else:
        pipe.gmres(**gm).backsolve()
    e = pipe.error_sums(pipe.elem_owned)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) != "nccl":
            e = e.cpu()
        dist.all_reduce(e, group=group)
    e = e.cpu().numpy()
    return float(np.sqrt(e[0])), float(np.sqrt(e[1]))
