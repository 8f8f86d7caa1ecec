"""Synthetic unit-square triangulations and their facet topology (host side, numpy).

Replaces ``Mesh(unit_square.GenerateMesh(maxh=size))``
(reference Python/convection_diffusion.py:47, :228).  Netgen's advancing-front
meshes cannot be reproduced here, so the hot path is fed by deterministic
synthetic meshes:

* ``structured``   N x N cells, each split by its lower-left/upper-right diagonal.
* ``unstructured`` same lattice, interior vertices jittered by U(-0.25h, 0.25h)^2
  and the diagonal of every cell chosen by Bernoulli(1/2) (seeded).

Conventions (NGSolve ET_TRIG):
  local edges  e0 = {V2, V0},  e1 = {V1, V2},  e2 = {V0, V1}
  triangles are counter-clockwise and rotated so that the smallest global vertex
  number comes first, hence only two vertex-sort classes occur:
    class 0: v0 < v1 < v2      class 1: v0 < v2 < v1
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# local edge -> (local vertex a, local vertex b)   (NGSolve ET_TRIG edge table)
LOCAL_EDGES = np.array([[2, 0], [1, 2], [0, 1]], dtype=np.int32)

HEADLINE_SEED = 20261019


@dataclass
class TriMesh:
    verts: np.ndarray      # f64 [nv, 2]
    tris: np.ndarray       # i32 [ne, 3]   CCW, smallest vertex first
    el2facet: np.ndarray   # i32 [ne, 3]   facet id of local edge 0,1,2
    facet2el: np.ndarray   # i32 [nf, 2]   adjacent elements ascending, -1 if none
    facet_verts: np.ndarray  # i32 [nf, 2] vertices ascending
    facet_bnd: np.ndarray  # u8  [nf]      1 on the domain boundary
    maxh: float            # nominal mesh size handed to the enrichment indicators
    kind: str = "structured"

    @property
    def ne(self) -> int:
        return int(self.tris.shape[0])

    @property
    def nf(self) -> int:
        return int(self.facet_verts.shape[0])

    @property
    def nv(self) -> int:
        return int(self.verts.shape[0])

    def tri_class(self) -> np.ndarray:
        """0 if v0<v1<v2, 1 if v0<v2<v1 (after normalisation v0 is the minimum)."""
        return (self.tris[:, 1] > self.tris[:, 2]).astype(np.uint8)


def normalise_triangles(verts: np.ndarray, tris: np.ndarray) -> np.ndarray:
    """Make every triangle CCW and rotate it so its smallest vertex id is first."""
    tris = np.asarray(tris, dtype=np.int64).copy()
    P = verts[tris]
    det = ((P[:, 0, 0] - P[:, 2, 0]) * (P[:, 1, 1] - P[:, 2, 1])
           - (P[:, 1, 0] - P[:, 2, 0]) * (P[:, 0, 1] - P[:, 2, 1]))
    flip = det < 0
    tris[flip, 1], tris[flip, 2] = tris[flip, 2].copy(), tris[flip, 1].copy()
    k = np.argmin(tris, axis=1)
    idx = (k[:, None] + np.arange(3)[None, :]) % 3
    tris = np.take_along_axis(tris, idx, axis=1)
    return tris.astype(np.int32)


def build_topology(verts: np.ndarray, tris: np.ndarray, maxh: float, kind: str) -> TriMesh:
    tris = normalise_triangles(verts, tris)
    ne = tris.shape[0]
    nv = verts.shape[0]
    ev = tris[:, LOCAL_EDGES]                       # [ne, 3, 2]
    lo = ev.min(axis=2).astype(np.int64)
    hi = ev.max(axis=2).astype(np.int64)
    key = (lo * nv + hi).reshape(-1)
    ukey, inv = np.unique(key, return_inverse=True)  # facets numbered by (lo, hi)
    nf = ukey.shape[0]
    el2facet = inv.reshape(ne, 3).astype(np.int32)
    facet_verts = np.stack([ukey // nv, ukey % nv], axis=1).astype(np.int32)
    # adjacent elements, ascending element id
    order = np.argsort(inv, kind="stable")           # element-major within equal facets
    f_sorted = inv[order]
    el_sorted = (order // 3).astype(np.int32)
    first = np.ones(f_sorted.shape[0], dtype=bool)
    first[1:] = f_sorted[1:] != f_sorted[:-1]
    facet2el = np.full((nf, 2), -1, dtype=np.int32)
    facet2el[f_sorted[first], 0] = el_sorted[first]
    facet2el[f_sorted[~first], 1] = el_sorted[~first]
    facet_bnd = (facet2el[:, 1] < 0).astype(np.uint8)
    return TriMesh(verts=np.ascontiguousarray(verts, dtype=np.float64), tris=tris,
                   el2facet=el2facet, facet2el=facet2el, facet_verts=facet_verts,
                   facet_bnd=facet_bnd, maxh=float(maxh), kind=kind)


def unit_square_mesh(n: int, kind: str = "structured", seed: int = HEADLINE_SEED,
                     jitter: float = 0.25, maxh: float | None = None) -> TriMesh:
    """n x n x 2 triangles on [0,1]^2.  ``kind`` in {'structured', 'unstructured'}."""
    if n < 1:
        raise ValueError("need n >= 1")
    h = 1.0 / n
    g = np.arange(n + 1, dtype=np.float64) * h
    g[-1] = 1.0
    X, Y = np.meshgrid(g, g, indexing="xy")          # vertex id = i + j*(n+1)
    verts = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (i + j * (n + 1)).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    if kind == "structured":
        diag = np.zeros(n * n, dtype=bool)
    elif kind == "unstructured":
        rng = np.random.default_rng(seed)
        jit = rng.uniform(-jitter * h, jitter * h, size=verts.shape)
        interior = ((verts[:, 0] > 0.5 * h) & (verts[:, 0] < 1 - 0.5 * h)
                    & (verts[:, 1] > 0.5 * h) & (verts[:, 1] < 1 - 0.5 * h))
        verts = verts + jit * interior[:, None]
        diag = rng.random(n * n) < 0.5
    else:
        raise ValueError(f"unknown mesh kind {kind!r}")
    # diag False: split along v00-v11 ; diag True: split along v10-v01
    ta = np.where(diag[:, None], np.stack([v00, v10, v01], 1), np.stack([v00, v10, v11], 1))
    tb = np.where(diag[:, None], np.stack([v10, v11, v01], 1), np.stack([v00, v11, v01], 1))
    tris = np.empty((2 * n * n, 3), dtype=np.int64)
    tris[0::2] = ta
    tris[1::2] = tb
    return build_topology(verts, tris, maxh=h if maxh is None else maxh, kind=kind)


def mesh_for_size(size: float, kind: str = "structured", seed: int = HEADLINE_SEED) -> TriMesh:
    """Stand-in for ``unit_square.GenerateMesh(maxh=size)``: lattice with h = 1/ceil(1/size)."""
    n = max(1, int(np.ceil(1.0 / float(size) - 1e-12)))
    return unit_square_mesh(n, kind=kind, seed=seed, maxh=float(size))


def single_triangle(P) -> TriMesh:
    """One-element mesh (known-answer tests on a given triangle shape)."""
    verts = np.asarray(P, dtype=np.float64).reshape(3, 2)
    return build_topology(verts, np.array([[0, 1, 2]]), maxh=1.0, kind="single")


def from_arrays(verts, tris, maxh: float | None = None, kind: str = "imported") -> TriMesh:
    """Any conforming triangle mesh given as arrays (a Netgen / NGSolve mesh exported with `mesh.Coordinates()` and the vertex
    numbers of `mesh.Elements()`): triangles are made counter-clockwise, the facet topology is built.  maxh defaults to the
    longest edge (what the enrichment indicators of run.py:45 receive as `h`)."""
    verts = np.ascontiguousarray(verts, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    if maxh is None:
        e = verts[tris[:, [1, 2, 0]]] - verts[tris]
        maxh = float(np.sqrt((e ** 2).sum(axis=2)).max())
    return build_topology(verts, tris, maxh=maxh, kind=kind)


def stripe_order(mesh: TriMesh, direction=(0.0, 1.0)) -> TriMesh:
    """Renumbers the vertices along `direction` (default: y, across a flow in x) and rebuilds the topology.  Facets are numbered
    by their (lower, higher) vertex pair, so the facet-major rows of the condensed system become banded along that direction
    whatever numbering the mesh generator used: contiguous row ranges are then stripes of the domain, the SpMV halo of a stripe
    lies in its two neighbouring stripes (what ehdg_comm_set_partition requires), and the lines of the line-block preconditioner
    (along the flow) stay inside one rank.  This is the general-mesh route to the partitioned multi-GPU path."""
    d = np.asarray(direction, dtype=np.float64)
    key = mesh.verts @ (d / np.linalg.norm(d))
    order = np.lexsort((mesh.verts[:, 0] * d[1] - mesh.verts[:, 1] * d[0], key))      # along the stripe second
    new_id = np.empty(mesh.nv, dtype=np.int64)
    new_id[order] = np.arange(mesh.nv)
    verts = mesh.verts[order]
    tris = new_id[mesh.tris]
    # elements sorted by their lowest vertex: element-major stores follow the stripes too
    tris = tris[np.argsort(tris.min(axis=1), kind="stable")]
    return build_topology(verts, tris, maxh=mesh.maxh, kind=mesh.kind)
