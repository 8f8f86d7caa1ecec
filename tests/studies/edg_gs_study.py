"""Host study: EDG systems from the oracle, block Jacobi over streamline lines vs block Gauss-Seidel across the lines (upwind order)."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc
from oracle.edg_oracle import EDGOracle

def study(n, p, eps, kind, beta=(2.0, 0.001)):
    mesh = unit_square_mesh(n, kind=kind)
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    o = EDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, p, prob)
    o.prune()
    t0 = time.time()
    alpha = o.alpha_all()
    A, F = o.assemble(alpha)
    A = A.tocsr()
    # lines of elements: band index of the centroid across the flow (y)
    cy = mesh.verts[mesh.tris].mean(axis=1)[:, 1]
    h = np.sqrt(2.0 / mesh.ne)
    line = np.floor(cy / h + 0.25).astype(int)
    nl = line.max() + 1
    dofs = [np.concatenate([o.el_dofs(el) for el in np.nonzero(line == j)[0]]) for j in range(nl)]
    lus = [spla.splu(A[d][:, d].tocsc()) for d in dofs]
    def jac(r):
        z = np.zeros_like(r)
        for d, lu in zip(dofs, lus): z[d] = lu.solve(r[d])
        return z
    def gs(r, order):
        z = np.zeros_like(r)
        for j in order:
            d = dofs[j]
            z[d] = lus[j].solve(r[d] - A[d] @ z)
        return z
    res = {}
    for name, M in (("jacobi", jac), ("gs_up", lambda r: gs(r, range(nl))), ("gs_down", lambda r: gs(r, range(nl - 1, -1, -1)))):
        it = [0]
        def cb(x): it[0] += 1
        x, info = spla.gmres(A, F, M=spla.LinearOperator(A.shape, matvec=M), restart=60, maxiter=40, rtol=1e-10, callback=cb, callback_type='pr_norm')
        res[name] = (it[0], info, float(np.linalg.norm(F - A @ x) / np.linalg.norm(F)))
    print(n, p, eps, kind, nl, res, round(time.time() - t0, 1), flush=True)

for n in (8, 16, 32):
    study(n, 2, 1e-6, "structured")
study(16, 2, 1e-6, "unstructured")
study(24, 3, 1e-6, "unstructured")
