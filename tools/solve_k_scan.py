"""DQGMRES(k) on the structured 1M HDG system for several k: iterations, seconds, true residual."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh
n, order = int(os.environ.get("SL_N", "708")), int(os.environ.get("SL_P", "4"))
P = EHDGPipeline(unit_square_mesh(n, kind="structured"), make_spec(order, 1e-6, enriched=False)).setup().assemble()
for k in [int(v) for v in os.environ.get("SL_KS", "4,8,12,16").split(",")]:
    P.csr_build()
    P.gmres(restart=k, max_iters=60000, rtol=1e-8, method="dqgmres")
    gi = P.gmres_info
    P.backsolve()
    print(json.dumps(dict(k=k, iters=gi.iters, conv=gi.converged, rel=gi.rel_residual, seconds=gi.total_seconds,
                          ms_per_it=1e3 * gi.total_seconds / max(gi.iters, 1), l2=P.errors()[0])), flush=True)
