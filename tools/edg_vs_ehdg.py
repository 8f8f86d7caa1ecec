"""BASELINE config 3: enriched DG (EDG) against enriched HDG (EHDG) on the same mesh, order 4, eps = 1e-6 -- set-up, assembly and
solve times, iterations, errors.  EDG_N (default 708 = 1M elements), EDG_P, EDG_EPS, EDG_KIND, EDG_ENR (1 = enriched)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.edg_path import EDGPipeline
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, p, eps = int(os.environ.get("EDG_N", "708")), int(os.environ.get("EDG_P", "4")), float(os.environ.get("EDG_EPS", "1e-6"))
kind, enr = os.environ.get("EDG_KIND", "structured"), os.environ.get("EDG_ENR", "1") == "1"
iters = int(os.environ.get("EDG_ITERS", "4000"))
mesh = unit_square_mesh(n, kind=kind)
ev = lambda: torch.cuda.Event(enable_timing=True)
out = dict(n=n, p=p, eps=eps, kind=kind, enriched=enr, elements=mesh.ne)
for name, cls in (("edg", EDGPipeline), ("ehdg", EHDGPipeline)):
    P = cls(mesh, make_spec(p, eps, enriched=enr))
    P.setup()
    if name == "edg":
        P.assemble()
    else:
        P.assemble().csr_build()
    torch.cuda.synchronize()
    a, b, c = ev(), ev(), ev()
    a.record()
    P.setup()
    b.record()
    if name == "edg":
        P.assemble()
    else:
        P.assemble().csr_build()
    c.record()
    torch.cuda.synchronize()
    rec = dict(setup_ms=a.elapsed_time(b), assemble_ms=b.elapsed_time(c), rows=int(P.csr_sizes.n_rows), nnz=int(P.csr_sizes.nnz),
               ndof=P.ndof(), status_flags=int(((P.status & ~8) != 0).sum()))
    t0 = time.perf_counter()
    try:
        P.solve_auto(rtol=1e-8, max_iters=iters)
        P.backsolve()
        l2, h1 = P.errors()
        gi = P.gmres_info
        rec.update(attempts=P.solve_attempts, iters=int(gi.iters), converged=bool(gi.converged), rel_residual=float(gi.rel_residual),
                   solve_s=float(gi.total_seconds), setup_s=float(gi.setup_seconds), l2_error=l2, h1_error=h1, chains=int(gi.line_chains),
                   max_block=int(gi.line_max_block), factor_GB=gi.precond_bytes / 1e9, dropped=int(gi.dropped_dofs))
    except Exception as exc:
        rec["error"] = str(exc)
    rec["solve_wall_s"] = time.perf_counter() - t0
    out[name] = rec
    P.close()
    del P
    torch.cuda.empty_cache()
print(json.dumps(out))
