"""SpMV timing on the structured 1M HDG system (library named by $EHDG_LIB_NAME)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh
n, order = int(os.environ.get("SL_N", "708")), int(os.environ.get("SL_P", "4"))
P = EHDGPipeline(unit_square_mesh(n, kind="structured"), make_spec(order, 1e-6, enriched=False)).setup().assemble().csr_build()
nr, nnz = P.csr_sizes.n_rows, P.csr_sizes.nnz
x = torch.rand(nr, dtype=torch.float64, device=P.dev); y = torch.empty_like(x)
for _ in range(5): P.spmv(x, y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): P.spmv(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(json.dumps(dict(lib=os.environ.get("EHDG_LIB_NAME", "libehdg.so"), rows=nr, nnz=nnz, spmv_ms=ms, gbs_12nnz=(12.0 * nnz + 20.0 * nr) / ms / 1e6,
                      gbs_8nnz=(8.0 * nnz + 16.0 * nr) / ms / 1e6, checksum=float(y.sum()))))
