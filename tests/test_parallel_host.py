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


# ------------------------------------------------------------------------------------------------------------
# partitioned assembly: element selection, ownership, extents all-gathered over gloo, and the
# protocol of the partitioned solve set-up: each rank fills its own rows from its own + ghost elements only,
# block-Jacobi inverses of the halo facets arrive from the neighbours
# ------------------------------------------------------------------------------------------------------------
def _element_system(n, nb=2, seed=0):
    """random element blocks S_K (3 nb x 3 nb) on the facets of each element: the condensed system is their sum"""
    mesh = unit_square_mesh(n)
    interior = mesh.facet2el[:, 1] >= 0
    nbf = np.where(interior, nb, 0)
    row_start = np.concatenate([[0], np.cumsum(nbf)]).astype(np.int64)
    rng = np.random.default_rng(seed)
    SK = rng.standard_normal((mesh.ne, 3 * nb, 3 * nb))

    def assemble(elements, rows_lo=0, rows_hi=None):
        rows_hi = row_start[-1] if rows_hi is None else rows_hi
        A = sps.lil_matrix((row_start[-1], row_start[-1]))
        for el in elements:
            fs = mesh.el2facet[el]
            for a in range(3):
                if not interior[fs[a]]:
                    continue
                for i in range(nb):
                    r = row_start[fs[a]] + i
                    if not (rows_lo <= r < rows_hi):
                        continue
                    for b in range(3):
                        if interior[fs[b]]:
                            for j in range(nb):
                                A[r, row_start[fs[b]] + j] += SK[el, a * nb + i, b * nb + j]
        return A.tocsr()
    return mesh, row_start, assemble


def test_selection_and_ownership_partition_the_mesh():
    from convectiondiffusionequation_b200.parallel import owned_elements, select_elements
    mesh, row_start, assemble = _element_system(8)
    A = assemble(range(mesh.ne))
    for world in (2, 3, 4):
        rb = split_rows_at_facets(row_start, world)
        fb = np.searchsorted(row_start, rb, side="left")
        fb[-1] = mesh.nf
        owned = np.zeros(mesh.ne, dtype=int)
        for r in range(world):
            sel = select_elements(mesh.facet2el, int(fb[r]), int(fb[r + 1]), mesh.ne)
            own = owned_elements(mesh.el2facet, int(fb[r]), int(fb[r + 1]))
            assert np.all(sel[own == 1] == 1)                      # the owner has its elements selected
            owned += own
            # the rank's rows need nothing but the selected elements
            Ar = assemble(np.nonzero(sel)[0], int(rb[r]), int(rb[r + 1]))
            assert abs(Ar[rb[r]:rb[r + 1]] - A[rb[r]:rb[r + 1]]).max() == 0.0
            assert sel.sum() < mesh.ne
        assert np.all(owned == 1)                                  # every element has exactly one owner
    # torch tensors in, torch tensors out
    t = select_elements(torch.from_numpy(mesh.facet2el), 3, 40, mesh.ne)
    assert torch.equal(t, torch.from_numpy(select_elements(mesh.facet2el, 3, 40, mesh.ne)))
    t = owned_elements(torch.from_numpy(mesh.el2facet), 3, 40)
    assert torch.equal(t, torch.from_numpy(owned_elements(mesh.el2facet, 3, 40)))


def _part_worker(rank, world, port, out):
    from convectiondiffusionequation_b200.parallel import gather_extents, select_elements
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = 2
    mesh, row_start, assemble = _element_system(10, nb=nb, seed=5)
    rb = split_rows_at_facets(row_start, world)
    fb = np.searchsorted(row_start, rb, side="left")
    fb[-1] = mesh.nf
    r0, r1 = int(rb[rank]), int(rb[rank + 1])
    sel = select_elements(mesh.facet2el, int(fb[rank]), int(fb[rank + 1]), mesh.ne)
    A_own = assemble(np.nonzero(sel)[0], r0, r1)[r0:r1]            # this rank's rows from its own + ghost elements
    lo, hi = min(int(A_own.indices.min()), r0), max(int(A_own.indices.max()) + 1, r1)
    col_lo, col_hi = gather_extents(lo, hi)
    ok = check_neighbour_halo(rb, col_lo, col_hi)
    # block-Jacobi inverses: own facets computed, halo facets received (DistComm::halo_exchange with width nb)
    n = int(row_start[-1])
    dinv = torch.zeros(n, nb, dtype=torch.float64)
    for r in range(r0, r1, nb):
        D = A_own[r - r0:r - r0 + nb, r:r + nb].toarray()
        dinv[r:r + nb] = torch.from_numpy(np.linalg.inv(D))
    ops = []
    if rank > 0:
        send_hi = min(int(col_hi[rank - 1]), r1)
        if send_hi > r0:
            ops.append(dist.P2POp(dist.isend, dinv[r0:send_hi].clone(), rank - 1))
        if col_lo[rank] < r0:
            ops.append(dist.P2POp(dist.irecv, dinv[int(col_lo[rank]):r0], rank - 1))
    if rank + 1 < world:
        send_lo = max(int(col_lo[rank + 1]), r0)
        if send_lo < r1:
            ops.append(dist.P2POp(dist.isend, dinv[send_lo:r1].clone(), rank + 1))
        if col_hi[rank] > r1:
            ops.append(dist.P2POp(dist.irecv, dinv[r1:int(col_hi[rank])], rank + 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    # fold: own rows of A M^-1 against the global computation
    A = assemble(range(mesh.ne))
    Minv = sps.block_diag([np.linalg.inv(A[r:r + nb, r:r + nb].toarray()) for r in range(0, n, nb)]).tocsr()
    ref = (A @ Minv)[r0:r1].toarray()
    Ml = sps.lil_matrix((n, n))
    for r in range(int(col_lo[rank]), int(col_hi[rank]), nb):
        Ml[r:r + nb, r:r + nb] = dinv[r:r + nb].numpy()
    got = (A_own @ Ml.tocsr()).toarray()
    out[rank] = bool(ok and np.allclose(got, ref, rtol=1e-12, atol=1e-12))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_partitioned_setup_protocol_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_part_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
