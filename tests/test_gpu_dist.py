"""Two-rank NCCL run of the distributed facet solve (needs >= 2 GPUs; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_solve_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DS_N="48", DS_P="3", DS_EPS="1e-6")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tools", "dist_solve_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["converged"] and out["solution_identical_on_all_ranks"]
    assert out["rel_diff_vs_single"] < 1e-9
    assert abs(out["iters"] - out["single_iters"]) <= 0.02 * out["single_iters"] + 2
    # partitioned assembly: every rank assembled only the elements that touch its rows
    pa = out["partitioned"]
    assert pa["converged"] and pa["rel_diff_vs_single"] < 1e-9
    assert abs(pa["l2_error"] - out["l2_error"]) < 1e-9 * out["l2_error"]
    assert max(pa["elements_per_rank"]) < 0.6 * out["elements"] and sum(pa["owned_per_rank"]) == out["elements"]


@pytest.mark.parametrize("enriched", [False, True])
def test_partitioned_rows_equal_the_global_assembly(enriched):
    """One process plays every rank of a 3-way partition in turn: the CSR rows, right-hand side and element count of
    each share are bit-identical to the corresponding slice of the single-GPU assembly (no collective involved)."""
    import numpy as np
    import torch
    from convectiondiffusionequation_b200 import parallel
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    from tests.common import make_pair
    mesh = unit_square_mesh(12, kind="unstructured")
    _, spec = make_pair(mesh, 3, beta=(2.0, 1.0), eps=1e-2, enriched=enriched)
    F = EHDGPipeline(mesh, spec).setup().assemble().csr_build()
    rp, ci, rs, dm = F.csr_arrays()
    vals, rhs = F.vals.clone(), F.rhs.clone()
    n = F.csr_sizes.n_rows
    world, owned_total, rows_seen = 3, 0, 0
    for rank in range(world):
        P = EHDGPipeline(mesh, spec)
        parallel.partitioned_assemble(P, rank, world, exchange=False)
        assert P.csr_sizes.n_rows == n
        r0, r1 = int(P.row_begin[rank]), int(P.row_begin[rank + 1])
        assert r0 == int(rs[int(P.facet_begin[rank])]) and r1 == int(rs[int(P.facet_begin[rank + 1])])
        z0, z1 = int(rp[r0]), int(rp[r1])
        assert P.csr_sizes.nnz == z1 - z0
        prp, pci, prs, pdm = P.csr_arrays()
        assert torch.equal(prs, rs) and torch.equal(pdm, dm)
        assert torch.equal(prp[r0:r1 + 1] - prp[r0], rp[r0:r1 + 1] - z0)
        assert torch.equal(pci, ci[z0:z1])
        assert torch.equal(P.vals[:z1 - z0], vals[z0:z1])                 # bit-exact: same kernels, same order of additions
        assert torch.equal(P.rhs[r0:r1], rhs[r0:r1])
        n_sel = P.plan_sizes.n_poly + P.plan_sizes.n_gen
        # (with enrichment the partition is by cost: ranks without strip elements take more of the cheap ones)
        assert n_sel == int(P.elem_select.sum()) and n_sel < (0.75 if enriched else 0.6) * mesh.ne
        lo, hi = P.own_extent
        assert lo <= r0 and hi >= r1 and lo == min(int(ci[z0:z1].min()), r0)
        owned_total += int(P.elem_owned.sum())
        rows_seen += r1 - r0
        # own rows of the SpMV agree with the global operator
        x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device=F.dev)
        y_full = F.spmv(x)
        y = torch.zeros_like(x)
        P.spmv(x, y)
        assert torch.equal(y[r0:r1], y_full[r0:r1])
        P.close()
    assert owned_total == mesh.ne and rows_seen == n
    F.close()


def test_device_selection_matches_host_rule():
    import numpy as np
    import torch
    from convectiondiffusionequation_b200 import parallel
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    from tests.common import make_pair
    mesh = unit_square_mesh(9, kind="unstructured")
    _, spec = make_pair(mesh, 1, enriched=False)
    P = EHDGPipeline(mesh, spec)
    for f_lo, f_hi in [(0, mesh.nf), (0, 0), (17, 120), (mesh.nf - 30, mesh.nf)]:
        sel, own = P.select_elements(f_lo, f_hi)
        assert np.array_equal(sel.cpu().numpy(), parallel.select_elements(mesh.facet2el, f_lo, f_hi, mesh.ne))
        assert np.array_equal(own.cpu().numpy(), parallel.owned_elements(mesh.el2facet, f_lo, f_hi))
    P.close()
