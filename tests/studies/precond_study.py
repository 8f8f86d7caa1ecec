"""Host-side study behind the line-block Jacobi preconditioner (csrc/ehdg_chain.cu): GMRES iteration counts of facet
block-Jacobi, line-block Jacobi (lattice rows along / across the flow) on the condensed HDG system built with the C oracle.
Pure CPU (scipy); development aid, not product path.   python tools/precond_study.py 64 4 1.1e-5 [structured|unstructured]

eps is scaled with h so that tau h / |b| matches the 1M / 4M configurations (eps = 1e-6 * 708 / N for config 3).
Recorded (order 4, rtol 1e-8, GMRES(30)):   N=64 structured: facet 1686, lines along the flow 232, lines across 883
                                           N=128 structured: lines along the flow 342 (GMRES(8): 333)
                                           N=64 unstructured: lines along the flow, bin offset 0 / 0.25 / 0.5: 166 / 162 / 346
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from oracle import ehdg_oracle as orc, ehdg_cpu


def condensed_hdg(mesh, p, eps, beta=(2.0, 0.001)):
    prob = orc.run_py_problem(beta=beta, eps=eps, enriched=False)
    co = ehdg_cpu.COracle(p, prob)
    els = np.arange(mesh.ne, dtype=np.int32)
    r = co.batch(mesh.verts, mesh.tris, els, np.zeros(mesh.ne, np.uint8), np.zeros((mesh.ne, 3), np.uint8))
    nb = p + 1
    S, g = r["S"], r["g"]
    interior = mesh.facet2el[:, 1] >= 0
    fid = -np.ones(mesh.nf, np.int64)
    fid[interior] = np.arange(interior.sum())
    n = int(interior.sum()) * nb
    rows, cols, vals = [], [], []
    rhs = np.zeros(n)
    loc = np.arange(nb)
    for a in range(3):
        fa = fid[mesh.el2facet[:, a]]
        for b in range(3):
            fb = fid[mesh.el2facet[:, b]]
            ok = (fa >= 0) & (fb >= 0)
            rows.append(((fa[ok, None, None] * nb + loc[None, :, None]) + 0 * loc[None, None, :]).ravel())
            cols.append(((fb[ok, None, None] * nb + loc[None, None, :]) + 0 * loc[None, :, None]).ravel())
            vals.append(S[ok][:, a * nb:(a + 1) * nb, b * nb:(b + 1) * nb].ravel())
        ok = fa >= 0
        np.add.at(rhs, (fa[ok, None] * nb + loc[None, :]).ravel(), g[ok][:, a * nb:(a + 1) * nb].ravel())
    A = sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    cent = np.zeros((mesh.nf, 2))
    for e, (va, vb) in enumerate([(2, 0), (1, 2), (0, 1)]):
        cent[mesh.el2facet[:, e]] = 0.5 * (mesh.verts[mesh.tris[:, va]] + mesh.verts[mesh.tris[:, vb]])
    return A, rhs, nb, cent[interior]


def main():
    N, p, eps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    kind = sys.argv[4] if len(sys.argv) > 4 else "structured"
    mesh = unit_square_mesh(N, kind=kind)
    A, b, nb, fc = condensed_hdg(mesh, p, eps)
    n = A.shape[0]
    nfi = n // nb
    h = 1.0 / N
    dofs = lambda facets: (np.asarray(facets)[:, None] * nb + np.arange(nb)[None, :]).ravel()

    def block_prec(groups):
        lus = [(gi, spla.splu(A[gi][:, gi].tocsc())) for gi in groups]

        def ap(r):
            z = np.zeros_like(r)
            for gi, lu in lus:
                z[gi] = lu.solve(r[gi])
            return z
        return spla.LinearOperator((n, n), matvec=ap)

    def run(name, M, restart):
        it = [0]
        t = time.time()
        x, info = spla.gmres(A, b, M=M, restart=restart, maxiter=20000 // restart, rtol=1e-8, atol=0,
                             callback=lambda r: it.__setitem__(0, it[0] + 1), callback_type="pr_norm")
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        print(f"{name:36s} GMRES({restart:3d}) iterations {it[0]:6d} info {info} residual {res:.2e}  {time.time() - t:.1f} s", flush=True)

    def lines(axis, off):
        rid = np.clip(np.floor(fc[:, axis] / h + off).astype(int), 0, N)
        return [dofs(np.nonzero(rid == j)[0]) for j in range(N + 1) if (rid == j).any()]

    print(f"N {N} order {p} eps {eps} {kind}: {n} rows, {A.nnz} nnz")
    for restart in (30, 8):
        if nfi <= 20000:
            run("facet block-Jacobi", block_prec([dofs([f]) for f in range(nfi)]), restart)
        for off in (0.0, 0.25, 0.5):
            run(f"lines along the flow, offset {off}", block_prec(lines(1, off)), restart)
        run("lines across the flow", block_prec(lines(0, 0.25)), restart)


if __name__ == "__main__":
    main()
