"""Times the alpha / assemble+condense kernels of the library named by $EHDG_LIB_NAME (A/B builds)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, order, eps = int(os.environ.get("AB_N", "708")), int(os.environ.get("AB_P", "4")), 1e-6
mesh = unit_square_mesh(n, kind=os.environ.get("AB_KIND", "structured"))
P = EHDGPipeline(mesh, make_spec(order, eps)).setup()
for _ in range(3):
    P.assemble()
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
ta = tb = 0.0
K = 5
for _ in range(K):
    a0, a1, a2 = ev(), ev(), ev()
    a0.record(); P.alpha(); a1.record(); P.assemble_condense(); a2.record()
    torch.cuda.synchronize()
    ta += a0.elapsed_time(a1); tb += a1.elapsed_time(a2)
chk = float(P.S.double().abs().sum().item()), float(P.lam.sum().item())
print(json.dumps(dict(lib=os.environ.get("EHDG_LIB_NAME", "libehdg.so"), n=n, p=order, alpha_ms=ta / K, asm_ms=tb / K,
                      melem_s=mesh.ne / ((ta + tb) / K) / 1e3, checksum_S=chk[0], checksum_lam=chk[1])))
