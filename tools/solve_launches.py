"""Short Krylov solve on the structured 1M HDG system for a kernel launch list (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, order = int(os.environ.get("SL_N", "708")), int(os.environ.get("SL_P", "4"))
P = EHDGPipeline(unit_square_mesh(n, kind="structured"), make_spec(order, 1e-6, enriched=False)).setup().assemble().csr_build()
P.gmres(restart=int(os.environ.get("SL_K", "16")), max_iters=int(os.environ.get("SL_ITERS", "40")), rtol=1e-30,
        method=os.environ.get("SL_METHOD", "dqgmres"))
torch.cuda.synchronize()
print("iters", P.gmres_info.iters, "seconds", P.gmres_info.total_seconds)
