"""General meshes in the partitioned path (host logic, CPU): a Delaunay mesh with randomly numbered vertices has no banded facet numbering --
the halo of a contiguous row range reaches far beyond the neighbouring ranges -- and gets one from mesh.stripe_order."""
import numpy as np
import scipy.sparse as sps
from scipy.spatial import Delaunay

from convectiondiffusionequation_b200 import parallel
from convectiondiffusionequation_b200.mesh import from_arrays, stripe_order


def _facet_system(mesh, nb=3):
    """pattern of the condensed facet system in facet-major numbering (nb rows per interior facet)"""
    interior = mesh.facet2el[:, 1] >= 0
    nrows_f = np.where(interior, nb, 0)
    row_start = np.concatenate([[0], np.cumsum(nrows_f)]).astype(np.int64)
    rows, cols = [], []
    for a in range(3):
        for b in range(3):
            fa, fb = mesh.el2facet[:, a], mesh.el2facet[:, b]
            ok = interior[fa] & interior[fb]
            rows.append(row_start[fa[ok]])
            cols.append(row_start[fb[ok]])
    n = int(row_start[-1])
    A = sps.coo_matrix((np.ones(sum(len(r) for r in rows)), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    return row_start, A


def _halo_ok(mesh, nranks):
    row_start, A = _facet_system(mesh)
    rb = parallel.split_rows_at_facets(row_start, nranks)
    lo, hi = parallel.halo_extents(A.indptr, A.indices, rb)
    return parallel.check_neighbour_halo(rb, lo, hi)


def test_stripe_order_makes_a_random_mesh_partitionable():
    # a quality mesh (Delaunay of a jittered grid) whose vertices come in random order, as a mesh generator may number them
    rng = np.random.default_rng(5)
    g = np.linspace(0.0, 1.0, 41)
    X, Y = np.meshgrid(g, g)
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    inner = (pts[:, 0] > 0) & (pts[:, 0] < 1) & (pts[:, 1] > 0) & (pts[:, 1] < 1)
    pts[inner] += rng.uniform(-0.3 / 40, 0.3 / 40, size=(int(inner.sum()), 2))
    pts = pts[rng.permutation(len(pts))]
    tri = Delaunay(pts)
    mesh = from_arrays(pts, tri.simplices)
    assert mesh.ne == len(tri.simplices) and (mesh.facet2el[:, 1] < 0).sum() > 0
    assert not _halo_ok(mesh, 4)                        # generator numbering: no
    striped = stripe_order(mesh)
    assert striped.ne == mesh.ne and striped.nf == mesh.nf
    assert abs(_area(striped) - _area(mesh)) < 1e-12
    assert _halo_ok(striped, 4) and _halo_ok(striped, 8)


def _area(m):
    P = m.verts[m.tris]
    return 0.5 * np.abs((P[:, 0, 0] - P[:, 2, 0]) * (P[:, 1, 1] - P[:, 2, 1]) - (P[:, 1, 0] - P[:, 2, 0]) * (P[:, 0, 1] - P[:, 2, 1])).sum()
