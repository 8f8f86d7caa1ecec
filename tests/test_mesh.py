import numpy as np
import pytest

from convectiondiffusionequation_b200.coefficients import boundary_layer, x, y
from convectiondiffusionequation_b200.enrichment_proxy import vertex_flags
from convectiondiffusionequation_b200.mesh import LOCAL_EDGES, mesh_for_size, unit_square_mesh


@pytest.mark.parametrize("n,kind", [(1, "structured"), (4, "structured"), (7, "unstructured")])
def test_topology(n, kind):
    m = unit_square_mesh(n, kind=kind)
    assert m.ne == 2 * n * n and m.nf == 3 * n * n + 2 * n and m.nv == (n + 1) ** 2
    assert int(m.facet_bnd.sum()) == 4 * n
    P = m.verts[m.tris]
    det = (P[:, 0, 0] - P[:, 2, 0]) * (P[:, 1, 1] - P[:, 2, 1]) - (P[:, 1, 0] - P[:, 2, 0]) * (P[:, 0, 1] - P[:, 2, 1])
    assert np.all(det > 0) and abs(det.sum() / 2 - 1.0) < 1e-12
    assert np.all(m.tris[:, 0] == m.tris.min(axis=1))
    for el in range(m.ne):
        for e in range(3):
            f = m.el2facet[el, e]
            assert set(m.tris[el, LOCAL_EDGES[e]]) == set(m.facet_verts[f])
            assert el in m.facet2el[f]
    assert np.all((m.facet2el[:, 0] < m.facet2el[:, 1]) | (m.facet2el[:, 1] < 0))


def test_unstructured_is_seeded_and_headline_sizes():
    a, b = unit_square_mesh(6, kind="unstructured"), unit_square_mesh(6, kind="unstructured")
    assert np.array_equal(a.verts, b.verts) and np.array_equal(a.tris, b.tris)
    assert 2 * 224 ** 2 == 100352 and 2 * 708 ** 2 == 1002528 and 2 * 1414 ** 2 == 3998792
    assert mesh_for_size(0.0625).ne == 512 and mesh_for_size(0.1).ne == 200


def test_layer_expressions_and_vertex_flags():
    beta, eps = (2.0, 0.001), 0.01
    p, q = boundary_layer(x, beta[0], eps), boundary_layer(y, beta[1], eps)
    exact, coeff = p * q, beta[1] * p + beta[0] * q
    assert exact.desc() == ((0.0, 0.0, 0.0, 1.0), (200.0, 0.1))
    assert coeff.desc() == ((0.0, 0.001, 2.0, 0.0), (200.0, 0.1))
    assert p.single_axis() == (0, 200.0) and q.single_axis() == (1, 0.1)
    xs = np.array([0.0, 0.5, 1.0])
    assert np.allclose(exact(xs, xs), p(xs, xs) * q(xs, xs)) and abs(p(np.array([1.0]), 0)[0]) < 1e-14
    with pytest.raises(ValueError):
        p * p
    m = unit_square_mesh(4)
    vf = vertex_flags(m.verts, [lambda X, Y, h: X > 1 - h, lambda X, Y, h: Y > 1 - h], m.maxh)
    assert vf.shape == (2, m.nv) and vf[0].sum() == 5 and vf[1].sum() == 5
    vf2 = vertex_flags(m.verts, [lambda X, Y, h: bool(X > 1 - h) and True], m.maxh)   # scalar-only predicate
    assert np.array_equal(vf2[0], vf[0])
