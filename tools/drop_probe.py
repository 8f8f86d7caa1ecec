"""Enriched solve against the two selection tolerances (volume: vol_drop_tol, facet: drop_tol): does a larger tolerance make the
system of BASELINE config 4 (and its small reproducer N = 48, eps = 3e-5) solvable, and what error does it give?"""
import dataclasses, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n = int(os.environ.get("DP_N", "48"))
eps = float(os.environ.get("DP_EPS", "3e-5"))
kind = os.environ.get("DP_KIND", "unstructured")
restart = int(os.environ.get("DP_RESTART", "60"))
iters = int(os.environ.get("DP_ITERS", "3000"))
grid = json.loads(os.environ.get("DP_GRID", "[[1e-6,1e-6],[1e-4,1e-4],[1e-2,1e-2],[1e-1,1e-1],[1e-6,1e-2],[1e-2,1e-6]]"))
mesh = unit_square_mesh(n, kind=kind)
if not os.environ.get("DP_SKIP_HDG"):
    H = EHDGPipeline(mesh, make_spec(4, eps, enriched=False)).setup().assemble_csr()
    H.gmres(restart=restart, max_iters=iters, rtol=1e-8).backsolve()
    print(json.dumps(dict(n=n, eps=eps, type="hdg", iters=H.gmres_info.iters, conv=bool(H.gmres_info.converged), l2=H.errors()[0])), flush=True)
    H.close()
for vt, ft in grid:
    spec = dataclasses.replace(make_spec(4, eps, enriched=True), vol_drop_tol=vt)
    P = EHDGPipeline(mesh, spec).setup().assemble_csr()
    P.gmres(restart=restart, max_iters=iters, rtol=1e-8, drop_tol=ft).backsolve()
    gi = P.gmres_info
    l2, h1 = P.errors()
    print(json.dumps(dict(n=n, eps=eps, type="ehdg", vol_drop_tol=vt, drop_tol=ft, iters=gi.iters, conv=bool(gi.converged), res=gi.rel_residual,
                          sec=round(gi.total_seconds, 2), l2=l2, dropped_facet_dofs=gi.dropped_dofs,
                          elements_with_dropped_volume_enrichment=int(((P.status & 8) != 0).sum().item()),
                          xmax=float(P.xhat.abs().max().item()))), flush=True)
    P.close()
