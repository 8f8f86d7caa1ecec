"""How many iterations does (nearly) unrestarted GMRES need on the structured 1M HDG system?  (bounds what any short-recurrence
Krylov method can reach with the same block-Jacobi preconditioner)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh
n, order = int(os.environ.get("SL_N", "708")), int(os.environ.get("SL_P", "4"))
P = EHDGPipeline(unit_square_mesh(n, kind="structured"), make_spec(order, 1e-6, enriched=False)).setup().assemble()
for m in [int(v) for v in os.environ.get("SL_MS", "400,1200").split(",")]:
    P.csr_build()
    P.gmres(restart=m, max_iters=int(os.environ.get("SL_ITERS", "6000")), rtol=1e-8, method="gmres")
    gi = P.gmres_info
    print(json.dumps(dict(restart=m, iters=gi.iters, conv=gi.converged, rel=gi.rel_residual, seconds=gi.total_seconds)), flush=True)
