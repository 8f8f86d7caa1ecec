"""Host logic of the multi-GPU facet solve, world_size 2 over gloo on the CPU: row split at facet boundaries,
halo extents, and the halo-exchange + row-local SpMV protocol libehdg runs over NCCL (emulated with numpy
vectors and torch.distributed send/recv)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from convectiondiffusionequation_b200.mesh import unit_square_mesh
from convectiondiffusionequation_b200.parallel import check_neighbour_halo, halo_extents, split_rows_at_facets


def _facet_system(n, nb=3, seed=0):
    """a random matrix with the facet-major block pattern of the condensed system (nb dofs per interior facet)"""
    mesh = unit_square_mesh(n)
    interior = mesh.facet2el[:, 1] >= 0
    nbf = np.where(interior, nb, 0)
    row_start = np.concatenate([[0], np.cumsum(nbf)]).astype(np.int64)
    rows, cols = [], []
    for f in np.nonzero(interior)[0]:
        nbr = np.unique(np.concatenate([mesh.el2facet[e] for e in mesh.facet2el[f] if e >= 0]))
        nbr = nbr[interior[nbr]]
        r = row_start[f] + np.arange(nb)
        c = np.concatenate([row_start[g] + np.arange(nb) for g in nbr])
        rows.append(np.repeat(r, len(c)))
        cols.append(np.tile(c, nb))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    rng = np.random.default_rng(seed)
    A = sps.csr_matrix((rng.standard_normal(len(rows)), (rows, cols)), shape=(row_start[-1], row_start[-1]))
    A.sort_indices()
    return row_start, A


def test_split_and_halo_extents():
    row_start, A = _facet_system(12)
    for nranks in (1, 2, 3, 4):
        rb = split_rows_at_facets(row_start, nranks)
        assert rb[0] == 0 and rb[-1] == A.shape[0] and np.all(np.diff(rb) >= 0)
        assert set(rb).issubset(set(row_start))                       # never cuts a facet block
        sizes = np.diff(rb)
        assert sizes.max() - sizes.min() <= 3 * 5                      # balanced up to a few facet blocks
        lo, hi = halo_extents(A.indptr, A.indices, rb)
        for r in range(nranks):
            sub = A[rb[r]:rb[r + 1]]
            assert lo[r] == min(sub.indices.min(), rb[r]) and hi[r] == max(sub.indices.max() + 1, rb[r + 1])
        assert check_neighbour_halo(rb, lo, hi)
    # with too many ranks the band no longer fits the neighbours
    rb = split_rows_at_facets(row_start, 64)
    lo, hi = halo_extents(A.indptr, A.indices, rb)
    assert not check_neighbour_halo(rb, lo, hi)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    row_start, A = _facet_system(10, seed=3)
    n = A.shape[0]
    rb = split_rows_at_facets(row_start, world)
    lo, hi = halo_extents(A.indptr, A.indices, rb)
    r0, r1 = int(rb[rank]), int(rb[rank + 1])
    x_full = np.linspace(-1.0, 1.0, n)
    xg = torch.zeros(n, dtype=torch.float64)
    xg[r0:r1] = torch.from_numpy(x_full[r0:r1])                     # own rows only
    # the exchange of DistComm::halo_exchange (ehdg_dist.cu), with gloo send/recv
    ops = []
    if rank > 0:
        send_hi = min(int(hi[rank - 1]), r1)
        if send_hi > r0:
            ops.append(dist.P2POp(dist.isend, xg[r0:send_hi].clone(), rank - 1))
        if lo[rank] < r0:
            ops.append(dist.P2POp(dist.irecv, xg[int(lo[rank]):r0], rank - 1))
    if rank + 1 < world:
        send_lo = max(int(lo[rank + 1]), r0)
        if send_lo < r1:
            ops.append(dist.P2POp(dist.isend, xg[send_lo:r1].clone(), rank + 1))
        if hi[rank] > r1:
            ops.append(dist.P2POp(dist.irecv, xg[r1:int(hi[rank])], rank + 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    y_local = A[r0:r1] @ xg.numpy()
    # dot product with one all-reduce
    part = torch.tensor([float(y_local @ y_local)], dtype=torch.float64)
    dist.all_reduce(part)
    y_ref = A @ x_full
    ok = np.allclose(y_local, y_ref[r0:r1], rtol=1e-13, atol=1e-13) and abs(part.item() - float(y_ref @ y_ref)) < 1e-9 * float(y_ref @ y_ref)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_halo_exchange_protocol_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
