"""Host study: alternating-direction line blocks on condensed HDG systems (C oracle): z = Mx^-1 r; z += My^-1 (r - A z) with
Mx = streamline lines (the device preconditioner) and My = lines across the flow, against Mx alone.  N order eps [kind]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse.linalg as spla
from convectiondiffusionequation_b200.mesh import unit_square_mesh
from tests.studies.precond_study import condensed_hdg


def main():
    N, p, eps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    kind = sys.argv[4] if len(sys.argv) > 4 else "structured"
    mesh = unit_square_mesh(N, kind=kind)
    A, b, nb, fc = condensed_hdg(mesh, p, eps)
    n = A.shape[0]
    h = 1.0 / N
    dofs = lambda facets: (np.asarray(facets)[:, None] * nb + np.arange(nb)[None, :]).ravel()

    def lines(axis, off, width=1.0):
        rid = np.floor(fc[:, axis] / (h * width) + off).astype(int)
        return [dofs(np.nonzero(rid == j)[0]) for j in np.unique(rid)]

    def solver(groups):
        lus = [(gi, spla.splu(A[gi][:, gi].tocsc())) for gi in groups]

        def ap(r):
            z = np.zeros_like(r)
            for gi, lu in lus:
                z[gi] = lu.solve(r[gi])
            return z
        return ap

    mx, my = solver(lines(1, 0.25)), solver(lines(0, 0.25))
    my2, my4 = solver(lines(0, 0.25, 2.0)), solver(lines(0, 0.25, 4.0))
    mx2 = solver(lines(1, 0.25, 2.0))

    def run(name, ap, restart, cost):
        it = [0]
        t = time.time()
        x, info = spla.gmres(A, b, M=spla.LinearOperator((n, n), matvec=ap), restart=restart, maxiter=6000 // restart, rtol=1e-8, atol=0,
                             callback=lambda r: it.__setitem__(0, it[0] + 1), callback_type="pr_norm")
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        print(f"{name:44s} GMRES({restart:3d}) iterations {it[0]:5d} ({cost} sweeps each) info {info} residual {res:.1e} {time.time() - t:.0f} s", flush=True)

    def chain(*ms):
        def ap(r):
            z = ms[0](r)
            for m in ms[1:]:
                z = z + m(r - A @ z)
            return z
        return ap

    print(f"N {N} order {p} eps {eps} {kind}: {n} rows")
    restart = 30
    run("streamline lines x", mx, restart, 1)
    run("x, y", chain(mx, my), restart, 2)
    run("x, y, x", chain(mx, my, mx), restart, 3)
    run("x, y(2h), x", chain(mx, my2, mx), restart, 3)
    run("x, y(4h), x", chain(mx, my4, mx), restart, 3)
    run("x, x", chain(mx, mx), restart, 2)
    run("x, x(2h bands, offset), x", chain(mx, mx2, mx), restart, 3)
    run("y(4h), x", chain(my4, mx), restart, 2)


if __name__ == "__main__":
    main()
