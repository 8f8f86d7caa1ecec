"""Status of the enriched facet solve on the device at several sizes; optionally dumps the assembled system
(EP_DUMP=dir) for host-side preconditioner studies.  Development aid, not product path."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh


def run(n, order, eps, kind="structured", methods=(("dqgmres", 8), ("dqgmres", 40), ("gmres", 60)), iters=20000, dump=None, rtol=1e-8):
    mesh = unit_square_mesh(n, kind=kind)
    P = EHDGPipeline(mesh, make_spec(order, eps, enriched=True))
    P.setup().assemble()
    lam = P.lam.cpu().numpy(); st = P.status.cpu().numpy(); em = P.elem_mask.cpu().numpy()
    info = dict(n=n, p=order, eps=eps, ne=mesh.ne, marked=int((em != 0).sum()), flagged=int((st != 0).sum()),
                lam_marked_max=float(lam[em != 0].max()), lam_unmarked_max=float(lam[em == 0].max()))
    print(json.dumps(info), flush=True)
    if dump:
        P.csr_build()
        rp, ci, rs, dm = P.csr_arrays()
        np.savez_compressed(os.path.join(dump, f"ehdg_{n}_{order}_{eps}_{kind}.npz"), rowptr=rp.cpu().numpy(), colind=ci.cpu().numpy(),
                            vals=P.vals.cpu().numpy(), rhs=P.rhs.cpu().numpy(), row_start=rs.cpu().numpy(), dofmask=dm.cpu().numpy(),
                            facet_mask=P.facet_mask.cpu().numpy(), verts=mesh.verts, tris=mesh.tris, el2facet=mesh.el2facet,
                            facet2el=mesh.facet2el)
    for meth, k in methods:
        P.csr_build()
        P.gmres(restart=k, max_iters=iters, rtol=rtol, method=meth, strip_block=True).backsolve()
        if dump:
            np.savez_compressed(os.path.join(dump, f"sol_{n}_{order}_{eps}_{kind}.npz"), xhat=P.xhat.cpu().numpy(), lines=P.lines().cpu().numpy(), elem_active=P.elem_active.cpu().numpy(), status=P.status.cpu().numpy(), lam=P.lam.cpu().numpy())
        gi = P.gmres_info
        l2, h1 = P.errors()
        print(json.dumps(dict(method=meth, k=k, iters=gi.iters, conv=bool(gi.converged), res=gi.rel_residual, sec=round(gi.total_seconds, 3),
                              l2=l2, h1=h1, strip=gi.strip_dofs, blocks=gi.strip_blocks, dropped=gi.dropped_dofs)), flush=True)
    P.close()


if __name__ == "__main__":
    dump = os.environ.get("EP_DUMP")
    if dump:
        os.makedirs(dump, exist_ok=True)
    for (n, p, eps) in [(48, 4, 3e-5), (48, 4, 1.5e-5)]:
        try:
            run(n, p, eps, kind="unstructured", dump=dump, methods=(("gmres", 60),), iters=1500)
        except Exception as exc:
            print("FAILED", n, p, eps, exc, flush=True)
