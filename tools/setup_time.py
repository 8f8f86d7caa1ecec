"""Per-stage wall times of the per-mesh set-up (marks, plan, pruning, pattern, value storage), repeated: development aid for
the `setup_ms_per_mesh` figure of bench.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n = int(os.environ.get("AB_N", "1414"))
mesh = unit_square_mesh(n, kind="unstructured")
P = EHDGPipeline(mesh, make_spec(4, 1e-6))
for rep in range(6):
    t = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for name, fn in (("mark_plan", P.mark), ("prune", P.prune), ("csr_symbolic", lambda: P.csr_symbolic(1, 0)), ("alloc", P.csr_alloc_values)):
        fn()
        t[name + "_issue"] = round((time.perf_counter() - t0) * 1e3, 3)
        torch.cuda.synchronize()
        t[name] = round((time.perf_counter() - t0) * 1e3, 3)
        t0 = time.perf_counter()
    if rep >= 2:
        P.alpha(); P.assemble_condense_csr()
        torch.cuda.synchronize()
    print(json.dumps(dict(rep=rep, **t)), flush=True)
