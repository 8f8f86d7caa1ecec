"""Line-block Jacobi against facet block-Jacobi on the device: iterations, seconds, residual, error (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh


def run(n, order, eps, kind="structured", enriched=False, configs=(("line", "dqgmres", 8, 0), ("facet", "dqgmres", 8, 0)), iters=60000, rtol=1e-8,
        drop_tol=0.0, max_blocks=0):
    mesh = unit_square_mesh(n, kind=kind)
    P = EHDGPipeline(mesh, make_spec(order, eps, enriched=enriched))
    P.setup().assemble()
    for prec, meth, k, tgt in configs:
        P.csr_build()
        try:
            P.gmres(restart=k, max_iters=iters, rtol=rtol, method=meth, precond=prec, line_target_dofs=tgt, drop_tol=drop_tol, line_max_blocks=max_blocks).backsolve()
        except Exception as exc:
            print("FAILED", n, order, eps, prec, meth, k, exc, flush=True)
            continue
        gi = P.gmres_info
        l2, h1 = P.errors()
        print(json.dumps(dict(n=n, p=order, eps=eps, kind=kind, enr=enriched, precond=prec, method=meth, k=k, target=tgt, iters=gi.iters,
                              conv=bool(gi.converged), res=gi.rel_residual, res_raw=gi.rel_residual_unscaled, sec=round(gi.total_seconds, 3), setup=round(gi.setup_seconds, 3),
                              ms_per_it=round(1e3 * (gi.total_seconds - gi.setup_seconds) / max(1, gi.iters), 4), l2=l2, h1=h1,
                              chains=gi.line_chains, blocks=gi.line_blocks, maxb=gi.line_max_block, factor_GB=round(gi.precond_bytes / 1e9, 3),
                              dropped=gi.dropped_dofs, drop_tol=drop_tol, max_blocks=max_blocks, rows=int(P.csr_sizes.n_rows))), flush=True)
    P.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(16, 2, 1e-2)
        run(64, 4, 1.1e-5)
        run(64, 4, 1.1e-5, kind="unstructured")
        run(128, 4, 5.53e-6, configs=(("line", "dqgmres", 8, 0), ("line", "gmres", 30, 0), ("line", "dqgmres", 8, 32), ("facet", "dqgmres", 8, 0)))
        run(224, 3, 1e-4, configs=(("line", "dqgmres", 8, 0), ("facet", "dqgmres", 8, 0)))
    elif which == "scan":
        L = lambda m, k, t=0: ("line", m, k, t)
        run(16, 2, 1e-2, configs=(L("dqgmres", 8), L("gmres", 30)))
        run(128, 4, 5.53e-6, configs=(L("dqgmres", 8), L("dqgmres", 16), L("dqgmres", 30), L("gmres", 20), L("gmres", 30), L("gmres", 50),
                                      L("gmres", 30, 16), L("gmres", 30, 32), L("bicgstab", 1)))
        run(224, 3, 1e-4, configs=(L("gmres", 30), L("dqgmres", 16), ("facet", "dqgmres", 8, 0)))
        run(708, 4, 1e-6, configs=(L("gmres", 30), L("gmres", 30, 16), L("dqgmres", 16)))
    elif which == "1m":
        run(708, 4, 1e-6, configs=(("line", "dqgmres", 8, 0), ("line", "dqgmres", 8, 32), ("line", "gmres", 30, 0)))
    elif which == "4m":
        run(1414, 4, 1e-6, kind="unstructured", configs=(("line", "gmres", 30, 0), ("line", "gmres", 60, 0)), iters=5000)
        run(1414, 4, 1e-6, kind="unstructured", enriched=True, configs=(("line", "gmres", 30, 0),), iters=5000)
    elif which == "4mh":        # plain HDG at 4M, DQGMRES(8): A/B of the row equilibration (EHDG_NO_EQUILIBRATE=1)
        run(1414, 4, 1e-6, kind="unstructured", configs=(("line", "dqgmres", 8, 0),), iters=5000)
    elif which == "1mh":
        run(708, 4, 1e-6, configs=(("line", "dqgmres", 8, 0), ("line", "gmres", 30, 0)), iters=5000, max_blocks=-1)
    elif which == "one":
        run(708, 4, 1e-6, configs=(("line", "gmres", 30, 0),), iters=40)
    elif which == "enr3":
        for (n, p, eps) in [(16, 2, 1e-2), (16, 4, 1e-2), (32, 4, 1e-3), (32, 3, 1e-4), (64, 4, 1.1e-5), (64, 4, 1e-6), (224, 3, 1e-4)]:
            for dt in (0.0, 1e-6, 1e-10):
                run(n, p, eps, enriched=True, configs=(("line", "gmres", 30, 0),), iters=4000, drop_tol=dt)
            run(n, p, eps, enriched=True, configs=(("facet", "gmres", 60, 0),), iters=4000)
            run(n, p, eps, enriched=False, configs=(("line", "gmres", 30, 0),), iters=4000)
        run(708, 4, 1e-6, enriched=True, configs=(("line", "gmres", 30, 0),), iters=4000)
    elif which == "4m2":
        run(1414, 4, 1e-6, kind="unstructured", configs=(("line", "dqgmres", 16, 0), ("line", "gmres", 150, 0), ("line", "dqgmres", 8, 32)), iters=8000)
    elif which == "4m3":
        run(1414, 4, 1e-6, kind="unstructured", configs=(("line", "dqgmres", 8, 0),), iters=8000)
        run(1414, 4, 1e-6, kind="unstructured", enriched=True, configs=(("line", "dqgmres", 8, 0), ("line", "gmres", 150, 0)), iters=8000)
        run(708, 4, 1e-6, enriched=True, configs=(("line", "dqgmres", 8, 0),), iters=8000)
        run(224, 3, 1e-4, enriched=True, configs=(("line", "dqgmres", 8, 0),), iters=8000)
    elif which == "enru":
        for (n, p, eps) in [(64, 4, 1e-6), (64, 4, 2.2e-5), (224, 4, 1e-6), (708, 4, 1e-6)]:
            run(n, p, eps, kind="unstructured", enriched=True, configs=(("line", "gmres", 60, 0), ("facet", "gmres", 60, 0)), iters=3000)
            run(n, p, eps, kind="unstructured", enriched=False, configs=(("line", "gmres", 60, 0),), iters=3000)
    elif which == "seg":
        for mb in (0, 708, 354, 177, 88):
            run(708, 4, 1e-6, configs=(("line", "gmres", 30, 0),), iters=4000, max_blocks=mb)
        for mb in (0, 1414, 707, 354, 177):
            run(1414, 4, 1e-6, kind="unstructured", configs=(("line", "dqgmres", 8, 0),), iters=4000, max_blocks=mb)
    elif which == "enr":
        for (n, p, eps) in [(16, 2, 1e-2), (16, 4, 1e-2), (32, 4, 1e-3), (32, 3, 1e-4), (64, 4, 1.1e-5), (64, 4, 1e-6), (224, 3, 1e-4)]:
            run(n, p, eps, enriched=True, configs=(("line", "gmres", 30, 0), ("line", "gmres", 100, 0), ("facet", "gmres", 60, 0)), iters=6000)
        run(708, 4, 1e-6, enriched=True, configs=(("line", "gmres", 30, 0),), iters=6000)
