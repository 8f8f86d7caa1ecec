"""Fused a5-a8 store (ehdg_assemble_condense_csr) against assemble_condense + csr_numeric: the value array and the right-hand
side must be IDENTICAL (every entry has one writer, the two-contributor entries are added first + second like the gather), and
every entry must be written (the value array starts as NaN)."""
import numpy as np
import pytest
import torch

from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from tests.common import make_pair

pytestmark = pytest.mark.gpu


def _both(mesh, spec, force_general=False, nranks=1, rank=0):
    P = EHDGPipeline(mesh, spec, force_general=force_general).setup()
    if nranks > 1:
        P.csr_symbolic(nranks, rank)
        sel, _ = P.select_elements(int(P.facet_begin[rank]), int(P.facet_begin[rank + 1]))
        P.make_plan(sel)
    else:
        P.csr_pattern()
    P.alpha().assemble_condense()
    P.csr_alloc_values()
    P.csr_numeric()
    v0, b0, W0, z0 = P.vals.clone(), P.rhs.clone(), P.W.clone(), P.z.clone()
    P.vals.fill_(float("nan"))
    P.rhs.fill_(float("nan"))
    P.W.fill_(float("nan"))
    P.z.fill_(float("nan"))
    P.assemble_condense_csr()
    torch.cuda.synchronize()
    return P, v0, b0, W0, z0


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("kind", ["structured", "unstructured"])
def test_fused_store_equals_gather_hdg(order, kind):
    mesh = unit_square_mesh(9, kind=kind)
    _, spec = make_pair(mesh, order, eps=1e-3, enriched=False)
    P, v0, b0, W0, z0 = _both(mesh, spec)
    nnz, n = P.csr_sizes.nnz, P.csr_sizes.n_rows
    assert not torch.isnan(P.vals[:nnz]).any()
    assert torch.equal(P.vals[:nnz], v0[:nnz])
    assert torch.equal(P.rhs[:n], b0[:n])
    assert torch.equal(P.W, W0) and torch.equal(P.z, z0)
    P.close()


@pytest.mark.parametrize("order,eps,n", [(2, 1e-2, 8), (4, 1e-6, 24), (3, 1e-4, 16)])
def test_fused_store_equals_gather_enriched(order, eps, n):
    mesh = unit_square_mesh(n)
    _, spec = make_pair(mesh, order, eps=eps, enriched=True)
    P, v0, b0, W0, z0 = _both(mesh, spec)
    nnz, nr = P.csr_sizes.nnz, P.csr_sizes.n_rows
    assert P.plan_sizes.n_gen > 0 and P.plan_sizes.n_poly > 0
    assert not torch.isnan(P.vals[:nnz]).any()
    assert torch.equal(P.vals[:nnz], v0[:nnz])
    assert torch.equal(P.rhs[:nr], b0[:nr])
    assert torch.equal(P.W, W0) and torch.equal(P.z, z0)
    P.close()


def test_fused_store_all_elements_on_the_quadrature_path():
    mesh = unit_square_mesh(6, kind="unstructured")
    _, spec = make_pair(mesh, 2, eps=1e-2, enriched=True)
    P, v0, b0, _, _ = _both(mesh, spec, force_general=True)
    nnz, nr = P.csr_sizes.nnz, P.csr_sizes.n_rows
    assert torch.equal(P.vals[:nnz], v0[:nnz]) and torch.equal(P.rhs[:nr], b0[:nr])
    P.close()


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_fused_store_on_a_row_partition(rank):
    """partitioned assembly: own rows only, ghost elements contribute to own facets and skip the others"""
    mesh = unit_square_mesh(12)
    _, spec = make_pair(mesh, 3, eps=1e-4, enriched=True)
    P, v0, b0, _, _ = _both(mesh, spec, nranks=3, rank=rank)
    nnz = P.csr_sizes.nnz
    lo, hi = int(P.row_begin[rank]), int(P.row_begin[rank + 1])
    assert nnz > 0 and not torch.isnan(P.vals[:nnz]).any()
    assert torch.equal(P.vals[:nnz], v0[:nnz])
    assert torch.equal(P.rhs[lo:hi], b0[lo:hi])
    assert torch.isnan(P.rhs[:lo]).all() and torch.isnan(P.rhs[hi:P.csr_sizes.n_rows]).all()      # nothing outside the own rows
    P.close()


def test_fused_pipeline_solves_like_the_two_pass_pipeline():
    mesh = unit_square_mesh(16, kind="unstructured")
    _, spec = make_pair(mesh, 3, eps=1e-3, enriched=False)
    A = EHDGPipeline(mesh, spec).setup().assemble().csr_build()
    A.gmres(restart=40, max_iters=4000, rtol=1e-12).backsolve()
    B = EHDGPipeline(mesh, spec).setup().assemble_csr()
    B.gmres(restart=40, max_iters=4000, rtol=1e-12).backsolve()
    assert torch.equal(A.xhat, B.xhat) and torch.equal(A.u, B.u)
    assert A.errors() == B.errors()
    A.close()
    B.close()


@pytest.mark.parametrize("env", [{"EHDG_INKERNEL_MERGE": "1"}, {"EHDG_OVERLAP": "1"}, {"EHDG_DYNAMIC": "0"}, {"EHDG_DYNAMIC": "3"}])
def test_opt_in_variants_give_the_same_values(env):
    """the measured-and-rejected variants stay behind their switches (read once per process, hence the subprocess) and stay
    correct: in-kernel merge of the diagonal blocks, quadrature path on a side stream, static / drawn work distribution"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import torch\n"
        "from convectiondiffusionequation_b200.hot_path import EHDGPipeline\n"
        "from convectiondiffusionequation_b200.mesh import unit_square_mesh\n"
        "from tests.common import make_pair\n"
        "for n, order, eps, kind in ((24, 4, 1e-6, 'structured'), (17, 3, 1e-4, 'unstructured'), (40, 2, 1e-3, 'unstructured')):\n"
        "    mesh = unit_square_mesh(n, kind=kind)\n"
        "    _, spec = make_pair(mesh, order, eps=eps, enriched=True)\n"
        "    P = EHDGPipeline(mesh, spec).setup()\n"
        "    P.csr_pattern()\n"
        "    P.alpha().assemble_condense()\n"
        "    P.csr_numeric()\n"
        "    nnz, nr = P.csr_sizes.nnz, P.csr_sizes.n_rows\n"
        "    v0, b0 = P.vals[:nnz].clone(), P.rhs[:nr].clone()\n"
        "    for rep in range(3):\n"
        "        P.vals.fill_(float('nan')); P.rhs.fill_(float('nan'))\n"
        "        P.assemble_condense_csr()\n"
        "        torch.cuda.synchronize()\n"
        "        assert torch.equal(P.vals[:nnz], v0) and torch.equal(P.rhs[:nr], b0), (n, order, rep)\n"
        "    P.close()\n"
        "print('variant ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stderr[-2000:]
