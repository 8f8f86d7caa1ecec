"""a8 numeric fill: ms and GB/s of ehdg_csr_numeric on the named mesh (EHDG_CSR_NUMERIC_FACET=1: the warp-per-facet variant)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, order = int(os.environ.get("AB_N", "1414")), int(os.environ.get("AB_P", "4"))
mesh = unit_square_mesh(n, kind=os.environ.get("AB_KIND", "unstructured"))
P = EHDGPipeline(mesh, make_spec(order, 1e-6)).setup().assemble().csr_build()
for _ in range(3):
    P.csr_numeric()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    P.csr_numeric()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
nbytes = P.S.numel() * 8 + P.g.numel() * 8 + P.csr_sizes.nnz * 8 + P.csr_sizes.n_rows * 8
print(json.dumps(dict(variant="facet-lane-per-column" if os.environ.get("EHDG_CSR_NUMERIC_FACET") == "1" else "row", n=n, p=order, ms=ms, gbs=nbytes / ms / 1e6,
                      bytes=nbytes, checksum=float(P.vals.abs().sum().item()), rhs=float(P.rhs.abs().sum().item()))))
