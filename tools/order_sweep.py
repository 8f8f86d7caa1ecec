"""BASELINE config 5: order sweep 1-6 x eps sweep on the 1M-element structured mesh.
Per cell: assembly+condensation elements/s, FP64 fraction of the two polynomial-path kernels (own flop count)
and of the quadrature algorithm they replace; per order: SpMV GB/s of the HDG facet system."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_spec, measured_peaks
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

N = int(os.environ.get("SWEEP_N", "708"))
mesh = unit_square_mesh(N)
hbm, _, fp64, _ = measured_peaks()
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for order in range(1, 7):
    spmv = None
    for eps in (1e-2, 1e-4, 1e-6, 1e-8):
        P = EHDGPipeline(mesh, make_spec(order, eps)).setup()
        P.kernel_timing(True)
        for _ in range(3):
            P.assemble()
        torch.cuda.synchronize()
        ts, ka, kb = [], [], []
        for _ in range(4):
            a, b = ev(), ev()
            a.record(); P.assemble(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            t1, t2 = P.kernel_times_ms(); ka.append(t1); kb.append(t2)
        ms = float(np.mean(ts)); s = P.sizes; npoly = P.plan_sizes.n_poly
        mf = P.moment_rhs_fraction()            # elements whose f_V comes from the moment vectors execute fewer flops
        f_asm = s.flops_poly_assemble - mf * (2 * s.nq * s.n_p - 12 * s.n_p)
        row = dict(order=order, eps=eps, elements=mesh.ne, enriched_path=int(P.plan_sizes.n_gen), ms=ms, melem_s=mesh.ne / ms / 1e3,
                   alpha_tf=npoly * s.flops_poly_alpha / np.mean(ka) / 1e9, asm_tf=npoly * f_asm / np.mean(kb) / 1e9, moment_rhs_fraction=mf,
                   quad_equiv_tf=mesh.ne * s.flops_quad_elem / ms / 1e9, status_flags=int((P.status != 0).sum()))
        row["alpha_frac"], row["asm_frac"], row["quad_equiv_frac"] = row["alpha_tf"] / fp64, row["asm_tf"] / fp64, row["quad_equiv_tf"] / fp64
        if spmv is None:
            Ph = EHDGPipeline(mesh, make_spec(order, eps, enriched=False)).setup().assemble().csr_build()
            n, nnz = Ph.csr_sizes.n_rows, Ph.csr_sizes.nnz
            x = torch.ones(n, dtype=torch.float64, device=Ph.dev); y = torch.empty_like(x)
            for _ in range(3):
                Ph.spmv(x, y)
            a, b = ev(), ev(); a.record()
            for _ in range(20):
                Ph.spmv(x, y)
            b.record(); torch.cuda.synchronize()
            t = a.elapsed_time(b) / 20
            spmv = dict(rows=int(n), nnz=int(nnz), ms=t, gbs=(12.0 * nnz + 20.0 * n) / t / 1e6)
            spmv["frac_of_hbm"] = spmv["gbs"] / hbm
            Ph.close(); del Ph, x, y
        row["spmv"] = spmv
        rows.append(row)
        print(json.dumps(row), flush=True)
        P.close(); del P
        torch.cuda.empty_cache()

if os.environ.get("SWEEP_MD"):
    with open(os.environ["SWEEP_MD"], "w") as fh:
        fh.write("| p | eps | step ms | Melem/s | alpha TF/s (frac) | assemble TF/s (frac) | moment-rhs share | quadrature-equivalent TF/s (frac) | "
                 "flagged elements | SpMV rows / nnz | SpMV ms | SpMV GB/s (frac) |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            sp = r["spmv"]
            fh.write(f"| {r['order']} | {r['eps']:g} | {r['ms']:.2f} | {r['melem_s']:.1f} | {r['alpha_tf']:.2f} ({100 * r['alpha_frac']:.1f} %) | "
                     f"{r['asm_tf']:.2f} ({100 * r['asm_frac']:.1f} %) | {100 * r['moment_rhs_fraction']:.0f} % | "
                     f"{r['quad_equiv_tf']:.2f} ({100 * r['quad_equiv_frac']:.1f} %) | {r['status_flags']} | {sp['rows'] / 1e6:.2f} M / {sp['nnz'] / 1e6:.0f} M | "
                     f"{sp['ms']:.3f} | {sp['gbs']:.0f} ({100 * sp['frac_of_hbm']:.0f} %) |\n")
