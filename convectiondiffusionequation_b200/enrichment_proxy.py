"""Host-side part of the enrichment layer (reference Python/enrichment_proxy.py).

The proxy classes of the reference (EnrichmentProxy_VOL/_FAC, :83-195) only exist to state
u = u_poly + sum c_i psi_i to NGSolve's symbolic integrators; here that statement is built into
the kernels (csrc/ehdg_general.h: vol_shapes / facet_shapes).  What remains on the host is the
evaluation of the Python predicates `enr_indicator(x, y, h)` on the mesh vertices
(mark_elements, :201-208): the any-vertex / any-adjacent-element reductions run on the device
(ehdg_mark).
"""
from __future__ import annotations

import numpy as np


def vertex_flags(verts: np.ndarray, indicators, mesh_size: float) -> np.ndarray:
    """u8 [n_enrich, nv]: enr_indicator(x_v, y_v, mesh_size) per vertex, as the reference's
    mark_elements evaluates it (enrichment_proxy.py:204-206)."""
    out = np.zeros((len(indicators), verts.shape[0]), dtype=np.uint8)
    xs, ys = np.ascontiguousarray(verts[:, 0]), np.ascontiguousarray(verts[:, 1])
    for i, ind in enumerate(indicators):
        try:
            r = np.asarray(ind(xs, ys, mesh_size))
            if r.shape != xs.shape:
                raise ValueError
            out[i] = r.astype(bool)
        except Exception:
            out[i] = np.fromiter((bool(ind(float(a), float(b), mesh_size)) for a, b in zip(xs, ys)),
                                 dtype=bool, count=xs.shape[0])
    return out


def vertex_flags_device(verts_dev, indicators, mesh_size):
    """Same flags evaluated on the device copy of the vertices (torch): the predicates of run.py:45 are plain
    comparisons, which torch tensors support; returns None when a predicate cannot take tensors (the caller
    then evaluates it on the host with vertex_flags)."""
    import torch
    xs, ys = verts_dev[:, 0], verts_dev[:, 1]
    out = torch.zeros((len(indicators), verts_dev.shape[0]), dtype=torch.uint8, device=verts_dev.device)
    try:
        for i, ind in enumerate(indicators):
            r = ind(xs, ys, mesh_size)
            if not torch.is_tensor(r) or r.shape != xs.shape:
                return None
            out[i] = r.to(torch.uint8)
    except Exception:
        return None
    return out
