"""Iteration counts of the facet solve versus mesh size / epsilon / sweeps (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

def run(n, order, eps, sweeps, restart=40, iters=300, enriched=False, kind="structured", method="gmres"):
    mesh = unit_square_mesh(n, kind=kind)
    P = EHDGPipeline(mesh, make_spec(order, eps, enriched=enriched))
    P.setup().assemble().csr_build()
    P.gmres(restart=restart, max_iters=iters, rtol=1e-10, jacobi_sweeps=sweeps, method=method).backsolve()
    gi = P.gmres_info
    l2, _ = P.errors()
    print(json.dumps(dict(method=method, n=n, p=order, eps=eps, enr=enriched, sweeps=sweeps, restart=restart, iters=gi.iters, conv=bool(gi.converged),
                          res=gi.rel_residual, spmv=gi.spmv_calls, sec=round(gi.total_seconds, 3), l2=l2, strip=gi.strip_dofs)), flush=True)
    P.close()

if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "4m":
        run(1414, 4, 1e-6, 8, restart=40, iters=12000, method="dqgmres", kind="unstructured")
    else:
        for sw in (4, 8, 16):
            run(708, 4, 1e-6, sw, restart=40, iters=30000 // sw, method="dqgmres")
        run(224, 3, 1e-4, 8, restart=40, iters=3000, method="dqgmres")
