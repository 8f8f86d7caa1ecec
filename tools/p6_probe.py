import sys; sys.path.insert(0,'/root/repo')
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from tests.common import make_pair
for order in (5,6):
  for prec in ("line","facet"):
    for meth,k in (("gmres",80),("gmres",30)):
        mesh = unit_square_mesh(2, kind="unstructured")
        o, spec = make_pair(mesh, order, beta=(1.0, 0.5), eps=0.05, enriched=False)
        P = EHDGPipeline(mesh, spec).setup().assemble().csr_build()
        P.gmres(restart=k, max_iters=4000, rtol=1e-12, precond=prec, method=meth)
        gi=P.gmres_info
        print(order,prec,meth,k,gi.iters,gi.converged,gi.rel_residual,gi.rel_residual_unscaled,gi.line_chains,gi.line_blocks,gi.line_max_block,gi.dropped_dofs, flush=True)
