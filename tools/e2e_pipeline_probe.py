"""Timeline of the two-slot end-to-end pipeline of bench.py: per step, when its upload / compute / download ran on the device."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n = int(os.environ.get("AB_N", "1414"))
mesh = unit_square_mesh(n, kind="unstructured")
dev = torch.device("cuda", 0)
P = EHDGPipeline(mesh, make_spec(4, 1e-6)).setup()
P.assemble()
ne = mesh.ne
pinned = dict(verts=torch.from_numpy(mesh.verts).pin_memory(), tris=torch.from_numpy(mesh.tris).pin_memory(),
              el2facet=torch.from_numpy(mesh.el2facet).pin_memory(), facet2el=torch.from_numpy(mesh.facet2el).pin_memory())
copy_stream = torch.cuda.Stream(device=dev)
comp = torch.cuda.current_stream(dev)
E = lambda: torch.cuda.Event(enable_timing=True)
slots = [dict(lam=torch.empty(ne, dtype=torch.float64, device=dev), status=torch.zeros(ne, dtype=torch.int32, device=dev),
              up=E(), done=E(), down=E()) for _ in range(2)]
lam_h = [torch.empty(ne, dtype=torch.float64).pin_memory() for _ in range(2)]
st_h = [torch.empty(ne, dtype=torch.int32).pin_memory() for _ in range(2)]
K = 6
ev = [dict() for _ in range(K)]
host = [dict() for _ in range(K)]

def upload(i):
    s = slots[i % 2]
    with torch.cuda.stream(copy_stream):
        ev[i]["up0"] = E(); ev[i]["up0"].record(copy_stream)
        s["mesh"] = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        s["up"] = E(); s["up"].record(copy_stream); ev[i]["up1"] = s["up"]

t00 = E(); t00.record(comp)
torch.cuda.synchronize()
h0 = time.perf_counter()
upload(0)
for i in range(K):
    host[i]["top"] = time.perf_counter() - h0
    if i + 1 < K:
        upload(i + 1)
    s = slots[i % 2]
    comp.wait_event(s["up"]); comp.wait_event(s["down"]) if "down_rec" in s else None
    ev[i]["c0"] = E(); ev[i]["c0"].record(comp)
    P.upload_mesh(s["mesh"]); P.lam, P.status = s["lam"], s["status"]
    P.mark(); ev[i]["c_mark"] = E(); ev[i]["c_mark"].record(comp)
    host[i]["after_mark"] = time.perf_counter() - h0
    P.prune(); P.alpha(); P.assemble_condense()
    s["done"] = E(); s["done"].record(comp); ev[i]["c1"] = s["done"]
    host[i]["enqueued"] = time.perf_counter() - h0
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(s["done"])
        ev[i]["d0"] = E(); ev[i]["d0"].record(copy_stream)
        lam_h[i % 2].copy_(s["lam"], non_blocking=True); st_h[i % 2].copy_(s["status"], non_blocking=True)
        s["down"] = E(); s["down"].record(copy_stream); s["down_rec"] = True; ev[i]["d1"] = s["down"]
torch.cuda.synchronize()
total = time.perf_counter() - h0
for i in range(K):
    r = {k: round(t00.elapsed_time(v), 2) for k, v in ev[i].items()}
    r.update({"host_" + k: round(1e3 * v, 2) for k, v in host[i].items()})
    print(json.dumps(dict(step=i, **r)))
print("wall ms/step", 1e3 * total / K)
