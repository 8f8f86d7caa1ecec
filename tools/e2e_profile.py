"""Wall-clock breakdown of one end-to-end step (host mesh -> alpha table on host)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n = int(os.environ.get("AB_N", "1414"))
mesh = unit_square_mesh(n, kind="unstructured")
P = EHDGPipeline(mesh, make_spec(4, 1e-6)).setup()
P.assemble()
pinned = dict(verts=torch.from_numpy(mesh.verts).pin_memory(), tris=torch.from_numpy(mesh.tris).pin_memory(),
              el2facet=torch.from_numpy(mesh.el2facet).pin_memory(), facet2el=torch.from_numpy(mesh.facet2el).pin_memory())
lam_host = torch.empty(mesh.ne, dtype=torch.float64).pin_memory()
st_host = torch.empty(mesh.ne, dtype=torch.int32).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t = [T()]
    P.upload_mesh(pinned); t.append(T())
    P.mark(); t.append(T())
    P.prune(); t.append(T())
    P.alpha(); t.append(T())
    P.assemble_condense(); t.append(T())
    lam_host.copy_(P.lam, non_blocking=True); st_host.copy_(P.status, non_blocking=True); t.append(T())
    names = ["upload", "mark+plan", "prune", "alpha", "assemble", "d2h"]
    print(json.dumps({k: round(1e3 * (b - a), 2) for k, a, b in zip(names, t[:-1], t[1:])} | {"total": round(1e3 * (t[-1] - t[0]), 2)}))
