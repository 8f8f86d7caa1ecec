"""A/B of the a5-a8 store: assemble_condense + csr_numeric (two passes over S_K) against ehdg_assemble_condense_csr (the element
kernels write the value array).  Prints ms per variant, kernel-only times of the polynomial assemble kernel, and whether the
value arrays are identical."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, order = int(os.environ.get("AB_N", "1414")), int(os.environ.get("AB_P", "4"))
mesh = unit_square_mesh(n, kind=os.environ.get("AB_KIND", "unstructured"))
P = EHDGPipeline(mesh, make_spec(order, 1e-6)).setup()
P.csr_pattern()
P.kernel_timing(True)
P.alpha()


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks = []
    a.record()
    for _ in range(reps):
        fn()
        ks.append(P.kernel_times_ms()[1])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, sum(ks) / len(ks)


def two_pass():
    P.assemble_condense()
    P.csr_numeric()


ms2, k2 = timed(two_pass)
v0, b0 = P.vals.clone(), P.rhs.clone()
P.S = P.g = None
torch.cuda.empty_cache()
P.vals.fill_(float("nan"))
ms1, k1 = timed(P.assemble_condense_csr)
nnz, nr = P.csr_sizes.nnz, P.csr_sizes.n_rows
print(json.dumps(dict(n=n, p=order, elements=mesh.ne, two_pass_ms=ms2, two_pass_assemble_kernel_ms=k2, fused_ms=ms1,
                      fused_assemble_kernel_ms=k1, identical_vals=bool(torch.equal(P.vals[:nnz], v0[:nnz])),
                      identical_rhs=bool(torch.equal(P.rhs[:nr], b0[:nr])), nnz=int(nnz))))
