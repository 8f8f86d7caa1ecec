#!/usr/bin/env python
"""Benchmark of the (E)HDG hot path (driver contract: one JSON line on rank 0).

metric  : condensed EHDG elements/second (FP64) = elements whose stabilisation eigenvalue (a4),
          element blocks + right-hand side (a5, a7) and static condensation (a6) were computed,
          per second, whole job.  One "step" = one such pass over the whole synthetic mesh.
value   : device-timed (CUDA events, inputs resident in HBM).
e2e     : same metric through the public pipeline with HOST buffers: pinned mesh arrays -> device,
          mark/plan/prune, alpha, assemble+condense, per-element alpha table + status -> host.
solve   : CSR build + block-Jacobi GMRES + back-solve + L2 error, run once after the timed steps
          ("full solve time" of BASELINE.json's metric), with achieved SpMV GB/s.
--impl reference : the CPU restatement of the reference algorithm (oracle/) on the host cores,
          bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lattice N, kind, order, eps)      SURVEY.md 8(d)
    "ehdg_p4_eps1e-6_unstructured_4M": (1414, "unstructured", 4, 1e-6),
    "ehdg_p4_eps1e-6_structured_1M": (708, "structured", 4, 1e-6),
    "ehdg_p3_eps1e-4_structured_100k": (224, "structured", 3, 1e-4),
    "ehdg_p2_eps1e-2_runpy_512": (16, "structured", 2, 1e-2),
}
DEFAULT_WORKLOAD = "ehdg_p4_eps1e-6_unstructured_4M"
BETA = (2.0, 0.001)          # run.py:24


def make_spec(order, eps, enriched=True):
    from convectiondiffusionequation_b200 import boundary_layer, x, y
    from convectiondiffusionequation_b200.hot_path import ProblemSpec
    p, q = boundary_layer(x, BETA[0], eps), boundary_layer(y, BETA[1], eps)
    return ProblemSpec(order=order, epsilon=eps, beta=BETA, bonus_int=10, coeff=BETA[1] * p + BETA[0] * q, exact=p * q,
                       enrich_functions=(p, q) if enriched else (),
                       enrich_domain_ind=(lambda X, Y, h: X > 1 - h, lambda X, Y, h: Y > 1 - h) if enriched else ())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    fp64, fp64_src = 37.0, "nominal"
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
        fp64, fp64_src = float(fp["fp64_fma_tflops"]), "profiles/fp64_peak_r01.json (tools/fp64_peak.cu, measured on this pool's B200)"
    except Exception:
        pass
    return hbm, hbm_src, fp64, fp64_src


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores, bounded sample of the workload
# --------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=20.0, max_elems=None):
    from oracle import ehdg_cpu
    n, kind, order, eps = WORKLOADS[workload]
    return ehdg_cpu.time_sample(n, kind, order, eps, BETA, budget_s=budget_s, max_elems=max_elems)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference(args.workload, budget_s=min(20.0, 120.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.mean(vals)) if vals else float(info["value"])
    n, kind, order, eps = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "condensed EHDG elements/sec (FP64)", "value": v, "unit": "elements/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * info["sample_elements"] / v if v else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "lattice_n": n, "mesh": kind, "order": order, "epsilon": eps, "beta": list(BETA),
                       "bonus_int": 10},
            "cpu_baseline": {"value": v, "unit": "elements/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    _emit(line)


# --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner,
    torchrun notices) is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def kernel_source_hash():
    """sha256 over the sources of the polynomial-path kernels: a committed ncu traffic figure is only quoted for the code it
    was captured on"""
    import hashlib
    h = hashlib.sha256()
    for f in ("ehdg_fast.h", "ehdg_fast.cu", "ehdg_tensors.h", "ehdg_basis.h"):
        h.update(open(os.path.join(ROOT, "convectiondiffusionequation_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-solve", action="store_true", help="skip the one-off solves after the timed steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e-solve", action="store_true", help="skip the mesh-in / error-out leg through Convection_Diffusion (N = 1)")
    ap.add_argument("--solve-iters", type=int, default=6000)
    ap.add_argument("--solve-method", default="auto", choices=["auto", "gmres", "dqgmres", "bicgstab"],
                    help="auto = EHDGPipeline.solve_auto: DQGMRES(8), then GMRES(#lines / 10) if that did not converge")
    ap.add_argument("--solve-workload", default="same",
                    help="mesh of the one-off solve leg ('same' = the timed workload, partitioned like it at N > 1)")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--restart", type=int, default=30, help="GMRES restart / DQGMRES truncation length (tools/line_probe.py: GMRES(30) is the robust choice)")
    ap.add_argument("--sweeps", type=int, default=1, help="block-Jacobi sweeps per preconditioner application")
    ap.add_argument("--precond", default="line", choices=["line", "facet"], help="line-block or facet-block Jacobi")
    ap.add_argument("--solve-type", default="both", choices=["hdg", "ehdg", "both"],
                    help="systems of the one-off solve: the enriched one (the workload's own) and / or plain HDG on the same mesh")
    ap.add_argument("--enriched-headline-iters", type=int, default=600, help="iteration cap per attempt of the enriched 4M solve")
    ap.add_argument("--two-pass", action="store_true", help="a8 as a separate gather over the element blocks S_K (round-2a path) instead of the fused store")
    ap.add_argument("--replicas", action="store_true", help="N > 1: also time one private full mesh per rank (weak-scaled replicas)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from convectiondiffusionequation_b200 import build, parallel
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline
    from convectiondiffusionequation_b200.mesh import unit_square_mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()

    n, kind, order, eps = WORKLOADS[args.workload]
    # strong scaling: ONE mesh of the named size (same seed on every rank); at N > 1 the facet rows are split into N
    # contiguous ranges and a rank assembles, condenses and fills the rows of the elements that touch its range
    mesh = unit_square_mesh(n, kind=kind, seed=20261019)
    spec = make_spec(order, eps)
    P = EHDGPipeline(mesh, spec, device=dev)
    if world > 1:
        parallel.init_communicator(P)
        parallel.gather_extents(0, 0, None, dev)                 # NCCL connects its all-gather channels on first use
    ne = mesh.ne
    sizes = P.sizes
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def rank_setup(Q):
        """per-mesh set-up of this rank's share: marks, pruning, pattern of the own rows, element selection, plan, values"""
        if world > 1:
            parallel.partitioned_assemble(Q, rank, world, assemble=False)
        else:
            Q.setup()
            Q.csr_symbolic(1, 0)
            Q.csr_alloc_values()
        return Q

    def step(Q):
        """the metric's unit of work on this rank's share: a4 (alpha), a5-a7 (blocks, rhs, condensation), a8 (numeric fill)"""
        Q.alpha()
        if args.two_pass:
            Q.assemble_condense()
            Q.csr_numeric()
        else:
            Q.assemble_condense_csr()                            # the element kernels store straight into the value array

    rank_setup(P)                                                # first call pays the allocations (cached by the pools afterwards)
    barrier()
    t_s0 = ev(); t_s1 = ev()
    t_s0.record()
    rank_setup(P)
    t_s1.record()
    torch.cuda.synchronize(dev)
    setup_ms = t_s0.elapsed_time(t_s1)
    ps = P.plan_sizes
    n_share = int(ps.n_poly + ps.n_gen)

    # L2 flush between timed iterations: the outputs of one step (S, W: > 10 GB at 4M) exceed the 126 MB L2;
    # for small shares an explicit flush buffer is written.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    need_flush = ((ps.S_doubles if args.two_pass else int(P.csr_sizes.nnz)) + ps.W_doubles) * 8 < (512 << 20)

    for _ in range(max(args.warmup, 3)):
        step(P)
    barrier()
    # ---- timed region: K steps, device time per stage, clocks sampled meanwhile
    t_alpha, t_asm, t_csr = 0.0, 0.0, 0.0
    k_alpha, k_asm = [], []          # the two polynomial-path kernels alone (events recorded inside libehdg)
    P.kernel_timing(True)
    with ClockSampler(local) as clocks:
        barrier()
        stage = []
        for _ in range(args.steps):
            if need_flush:
                flush.fill_(1)
            a0, a1, a2, a3 = ev(), ev(), ev(), ev()
            a0.record()
            P.alpha()
            a1.record()
            if args.two_pass:
                P.assemble_condense()
                a2.record()
                P.csr_numeric()
            else:
                P.assemble_condense_csr()
                a2.record()
            a3.record()
            stage.append((a0, a1, a2, a3))
            if P.plan_sizes.n_poly > 0:
                ka, kb = P.kernel_times_ms()      # waits for this step's kernels: the next step starts after it anyway
                k_alpha.append(ka)
                k_asm.append(kb)
        barrier()
    for a0, a1, a2, a3 in stage:
        t_alpha += a0.elapsed_time(a1)
        t_asm += a1.elapsed_time(a2)
        t_csr += a2.elapsed_time(a3)
    ms_step = (t_alpha + t_asm + t_csr) / args.steps
    tmax = torch.tensor([ms_step, setup_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step_max, setup_ms_max = float(tmax[0].item()), float(tmax[1].item())
    value = ne / (ms_step_max * 1e-3)                 # global elements / slowest rank's time for its share
    status_bad = int(((P.status & ~8) != 0).sum().item())           # singular / non-finite / deficient-alpha flags
    status_dropped = int(((P.status & 8) != 0).sum().item())         # elements where a dependent volume enrichment left the space
    moment_frac = P.moment_rhs_fraction()
    share = torch.tensor([n_share, int(P.csr_sizes.nnz)], dtype=torch.int64, device=dev)
    shares = [torch.zeros_like(share) for _ in range(world)]
    if world > 1:
        dist.all_gather(shares, share)
    else:
        shares = [share]

    # ---- e2e: host buffers in, alpha table + status out, every step
    pinned = dict(verts=torch.from_numpy(mesh.verts).pin_memory(), tris=torch.from_numpy(mesh.tris).pin_memory(),
                  el2facet=torch.from_numpy(mesh.el2facet).pin_memory(), facet2el=torch.from_numpy(mesh.facet2el).pin_memory())

    # Two-deep pipeline through the public API: a copy stream uploads the mesh of step i+1 and downloads the results of
    # step i-1 while the compute stream works on step i.  Every step still moves its own inputs host->device and its own
    # results device->host inside the timed region; only the copies overlap the kernels of the neighbouring steps.
    copy_stream = torch.cuda.Stream(device=dev)
    comp = torch.cuda.current_stream(dev)
    slots = [dict(), dict()]
    for sl_ in slots:
        sl_["lam"] = torch.empty(ne, dtype=torch.float64, device=dev)
        sl_["status"] = torch.zeros(ne, dtype=torch.int32, device=dev)
        sl_["up"], sl_["done"], sl_["down"] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    lam_hosts = [torch.empty(ne, dtype=torch.float64).pin_memory() for _ in range(2)]
    st_hosts = [torch.empty(ne, dtype=torch.int32).pin_memory() for _ in range(2)]

    def e2e_upload(i):
        sl_ = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            sl_["mesh"] = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            sl_["up"].record(copy_stream)

    def e2e_compute(i):
        sl_ = slots[i % 2]
        comp.wait_event(sl_["up"])
        comp.wait_event(sl_["down"])                       # the slot's previous results have left the device
        P.upload_mesh(sl_["mesh"])                         # device tensors: adopts them, no further copy
        P.lam, P.status = sl_["lam"], sl_["status"]
        rank_setup(P)
        step(P)
        sl_["done"].record(comp)

    def e2e_download(i):
        sl_ = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl_["done"])
            lam_hosts[i % 2].copy_(sl_["lam"], non_blocking=True)
            st_hosts[i % 2].copy_(sl_["status"], non_blocking=True)
            sl_["down"].record(copy_stream)

    def e2e_run(k):
        e2e_upload(0)
        for i in range(k):
            if i + 1 < k:
                e2e_upload(i + 1)
            e2e_compute(i)
            e2e_download(i)
        copy_stream.synchronize()
        torch.cuda.synchronize(dev)

    # Warm-up: the end-to-end loop is the host-heavy part of the bench (per-mesh set-up with host round trips, pinned copies).
    # On some boxes of the pool its first repetitions run at a third of the steady speed while the host wakes up from the
    # GPU-bound phase before it (measured: 135, 55, 44 ms per step in three back-to-back repetitions; 44.4, 44.4, 44.5 on
    # others), so it is warmed with 18 steps before the timed repetitions (a fixed count: every rank runs the same number of
    # set-ups, which contain collectives at N > 1).
    for _ in range(3):
        e2e_run(6)
    e2e_steps = max(2, min(args.steps, 6))
    e2e_reps = []                      # wall clock with host copies: seven repetitions, the median is reported (single
    for _ in range(7):                 # repetitions are hit by host-side hiccups of 2-4x on some boxes)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        e2e_reps.append((time.perf_counter() - t0) / e2e_steps)
    e2e_s = sorted(e2e_reps)[len(e2e_reps) // 2]
    lam_host = lam_hosts[(e2e_steps - 1) % 2]
    sel_host = P.elem_select.cpu() if getattr(P, "elem_select", None) is not None else None
    lam_chk = lam_host if sel_host is None else lam_host[sel_host.bool()]
    assert bool(torch.isfinite(lam_chk).all()) and float(lam_chk.min()) > 0.0, "e2e: alpha table did not arrive"
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = ne / float(te.item())
    h2d = P.h2d_bytes + (P.vert_flags.numel() if hasattr(P, "vert_flags") else 0)
    d2h = ne * 12

    # ---- optional: weak-scaled replicas (one private full mesh per rank, no partition), the round-1 number
    replica_value = None
    if args.replicas and world > 1:
        R = EHDGPipeline(unit_square_mesh(n, kind=kind, seed=20261019 + rank), spec, device=dev)
        R.setup()
        for _ in range(3):
            R.alpha(); R.assemble_condense()
        barrier()
        r0_, r1_ = ev(), ev()
        r0_.record()
        for _ in range(args.steps):
            R.alpha(); R.assemble_condense()
        r1_.record()
        barrier()
        tr_ = torch.tensor([r0_.elapsed_time(r1_) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(tr_, op=dist.ReduceOp.MAX)
        replica_value = world * ne / (float(tr_.item()) * 1e-3)
        R.close()
        del R
        torch.cuda.empty_cache()

    # ---- one-off solves (a9-a10) on the timed pipeline's own system (already assembled and filled: a8 is in the step) and,
    # with --solve-type both, on the plain HDG system of the same mesh.  N > 1: the rows stay partitioned, the Krylov
    # iteration runs over NCCL (SpMV halo exchange + all-reduced dot products).
    solves = []
    if not args.no_solve:
        swl = args.workload if args.solve_workload == "same" else args.solve_workload
        sn, skind, sorder, seps = WORKLOADS[swl]
        legs = [("ehdg", swl), ("hdg", swl)] if args.solve_type == "both" else [(args.solve_type, swl)]
        if args.solve_type == "both" and swl == DEFAULT_WORKLOAD and world == 1:
            legs.append(("ehdg", "ehdg_p4_eps1e-6_structured_1M"))       # BASELINE config 3: the largest size at which the enriched system converges
        for styp, lwl in legs:
            sn, skind, sorder, seps = WORKLOADS[lwl]
            if styp == "ehdg" and lwl == args.workload:
                Ps = P
                P.upload_mesh()                                          # back to the pipeline's own mesh tensors
                rank_setup(P)
                step(P)
            else:
                smesh = mesh if lwl == args.workload else unit_square_mesh(sn, kind=skind)
                Ps = EHDGPipeline(smesh, make_spec(sorder, seps, enriched=(styp == "ehdg")), device=dev)
                if world > 1:
                    parallel.init_communicator(Ps)
                rank_setup(Ps)
                step(Ps)
            swl_leg = lwl
            smesh = Ps.mesh_host
            barrier()
            if rank == 0:
                print(f"[bench] solve leg: {styp} on {swl_leg}", file=sys.stderr, flush=True)
            nrows, nnz = Ps.csr_sizes.n_rows, Ps.csr_sizes.nnz            # nnz: this rank's rows
            own_rows = nrows if world == 1 else int(Ps.row_begin[rank + 1] - Ps.row_begin[rank])
            # SpMV bandwidth, 20 launches (own rows at N > 1, no exchange).  Two byte counts: the CSR-equivalent 12 nnz + 20 rows
            # of SURVEY 8d, and what the facet-block layout really moves (8 B per value, ONE 4-byte column index per block
            # column, 16 B of block metadata per facet, x and y once): the HBM fraction is quoted on the latter.
            xs = torch.ones(max(nrows, 1), dtype=torch.float64, device=dev)
            ys = torch.empty_like(xs)
            for _ in range(3):
                Ps.spmv(xs, ys)
            b0, b1 = ev(), ev()
            b0.record()
            for _ in range(20):
                Ps.spmv(xs, ys)
            b1.record()
            torch.cuda.synchronize(dev)
            spmv_ms = b0.elapsed_time(b1) / 20
            nb_rows = sizes.nfl if styp == "hdg" else None
            rp, _, rs_, dm_ = Ps.csr_arrays()
            f_act = (rs_[1:] - rs_[:-1]) > 0
            if world > 1:
                f_idx = torch.arange(f_act.numel(), device=dev)
                f_act &= (f_idx >= int(Ps.facet_begin[rank])) & (f_idx < int(Ps.facet_begin[rank + 1]))
            first_rows = rs_[:-1][f_act].long()
            block_cols = int((rp[first_rows + 1] - rp[first_rows]).sum().item())
            n_fac = int(f_act.sum().item())
            del rp, rs_, dm_, first_rows, f_act
            bytes_csr = 12.0 * nnz + 20.0 * own_rows
            bytes_actual = 8.0 * nnz + 4.0 * block_cols + 16.0 * n_fac + 16.0 * own_rows
            spmv_gbs = bytes_csr / (spmv_ms * 1e-3) / 1e9
            spmv_gbs_actual = bytes_actual / (spmv_ms * 1e-3) / 1e9
            del xs, ys
            barrier()
            s2, s3 = ev(), ev()
            s2.record()
            sweeps = max(1, args.sweeps)
            # the enriched 4M headline system does not converge with this preconditioner family: the attempt is
            # capped and recorded as it ends; every other leg gets the full budget
            it_cap = min(args.solve_iters, args.enriched_headline_iters) if (styp == "ehdg" and swl_leg == DEFAULT_WORKLOAD) else args.solve_iters
            kw = dict(restart=args.restart, max_iters=it_cap, rtol=args.rtol, jacobi_sweeps=sweeps, method=args.solve_method,
                      precond=args.precond, strip_block=True)
            err = None
            try:
                if world > 1:
                    l2, h1 = parallel.partitioned_solve(Ps, **kw)
                else:
                    if args.solve_method == "auto":
                        Ps.solve_auto(rtol=args.rtol, max_iters=it_cap, precond=args.precond)
                    else:
                        Ps.gmres(**kw)
                    Ps.backsolve()
                    l2, h1 = Ps.errors()
            except Exception as exc:                       # reported in the line, the assembly numbers stand
                l2 = h1 = None
                err = str(exc)
            s3.record()
            torch.cuda.synchronize(dev)
            gi = getattr(Ps, "gmres_info", None)
            sv = {"type": styp, "ranks": world, "workload": swl_leg, "method": args.solve_method, "precond": args.precond,
                  "elements": int(smesh.ne), "rows": int(nrows), "nnz_this_rank": int(nnz), "error": err,
                  "gmres_backsolve_error_ms": s2.elapsed_time(s3), "restart": args.restart, "jacobi_sweeps": sweeps, "rtol": args.rtol,
                  "l2_error": l2, "h1_error": h1, "spmv_ms": spmv_ms, "spmv_gbs_csr_equivalent": spmv_gbs, "spmv_gbs_actual": spmv_gbs_actual,
                  "spmv_bytes_actual": bytes_actual, "spmv_bytes_csr_equivalent": bytes_csr}
            if gi is not None and err is None:
                it_s = float(gi.total_seconds) - float(gi.setup_seconds)
                sv.update({"attempts": getattr(Ps, "solve_attempts", None) if args.solve_method == "auto" else None,
                           "gmres_iters": int(gi.iters), "gmres_converged": bool(gi.converged), "gmres_rel_residual": float(gi.rel_residual),
                           "gmres_seconds": float(gi.total_seconds), "precond_setup_seconds": float(gi.setup_seconds),
                           "ms_per_iteration": 1e3 * it_s / max(1, int(gi.iters)), "spmv_calls": int(gi.spmv_calls),
                           "dropped_dofs": int(gi.dropped_dofs), "strip_dofs": int(gi.strip_dofs), "line_chains": int(gi.line_chains),
                           "line_blocks": int(gi.line_blocks), "line_max_block": int(gi.line_max_block),
                           "precond_factor_bytes_this_rank": float(gi.precond_bytes),
                           "precond_apply_gbs_floor": (float(gi.precond_bytes) / 1e9) / max(1e-9, (it_s / max(1, int(gi.iters)))),
                           "note": "enriched systems at this size are singular to working precision in the reference itself (DESIGN 2.3, 6); "
                                   "the residual is that of the assembled system, dependent dofs dropped" if styp == "ehdg" else None})
            solves.append(sv)
            if Ps is not P:
                Ps.close()
                del Ps
                torch.cuda.empty_cache()

    # ---- e2e_solve (N = 1): the reference-facing call, mesh in -> L2 error out, through Convection_Diffusion._solveEHDG
    e2e_solve = None
    if world == 1 and not args.no_e2e_solve:
        try:
            e2e_solve = run_e2e_solve(dev)
        except Exception as exc:
            e2e_solve = {"error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, hbm_src, fp64, fp64_src = measured_peaks()
    # dominant kernel: fast_assemble_kernel<P> (one launch per step; fast_alpha_kernel<P> second)
    n_poly, n_gen = int(ps.n_poly), int(ps.n_gen)
    f_asm_quad, f_alpha = int(sizes.flops_poly_assemble), int(sizes.flops_poly_alpha)
    # elements whose right-hand side comes from the six moment vectors (exp underflowed on the whole element) spend
    # 12 n_p flops on it instead of 2 nq n_p: the reported flop count is the one executed
    f_asm = f_asm_quad - moment_frac * (2 * int(sizes.nq) * int(sizes.n_p) - 12 * int(sizes.n_p))
    asm_ms = float(np.mean(k_asm)) if k_asm else None
    alpha_ms = float(np.mean(k_alpha)) if k_alpha else None
    achieved_tf = n_poly * f_asm / (asm_ms * 1e-3) / 1e12 if asm_ms else None
    alpha_tf = n_poly * f_alpha / (alpha_ms * 1e-3) / 1e12 if alpha_ms else None
    step_tf = n_poly * (f_asm + f_alpha) / ((t_alpha + t_asm) / args.steps * 1e-3) / 1e12
    traffic, traffic_note = None, "no ncu capture of this kernel source on file"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_assemble_traffic.json")))
        if tr.get("workload") == args.workload and tr.get("kernel_source_sha256_16") == kernel_source_hash() and world == 1:
            traffic, traffic_note = tr["dram_bytes_per_launch"], f"ncu --set full capture {tr.get('captured')} of this kernel source (hash matches)"
        elif tr.get("kernel_source_sha256_16") != kernel_source_hash():
            traffic_note = "profiles/r02_assemble_traffic.json was captured on different kernel sources: not quoted"
    except Exception:
        pass
    own_rows_step = int(P.csr_sizes.n_rows if world == 1 else P.row_begin[rank + 1] - P.row_begin[rank])
    store_doubles = (ps.S_doubles + ps.g_doubles) if args.two_pass else (int(P.csr_sizes.nnz) + own_rows_step)
    out_bytes = (store_doubles + ps.W_doubles + ps.z_doubles + n_share) * 8 + n_share * 60
    line = {
        "metric": "condensed EHDG elements/sec (FP64)", "value": value, "unit": "elements/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "lattice_n": n, "mesh": kind, "order": order, "epsilon": eps, "beta": list(BETA),
                   "bonus_int": 10, "elements": ne, "facets": mesh.nf, "partition": "one mesh, facet rows split into N contiguous ranges; "
                   "a rank computes the elements touching its range (one redundant ghost layer, no collective in the step)",
                   "elements_per_rank": [int(t[0]) for t in shares], "nnz_per_rank": [int(t[1]) for t in shares],
                   "enriched_path_elements_rank0": n_gen,
                   "step": "alpha (a4) + blocks, rhs, condensation (a5-a7) + numeric CSR fill (a8) of the rank's share",
                   "a8_store": "two passes: element blocks S_K to HBM, then the gather ehdg_csr_numeric" if args.two_pass else
                               "fused: the element kernels write the CSR value array (ehdg_assemble_condense_csr), diagonal blocks merged first + second",
                   "l2_flush": "explicit 256 MiB write between steps" if need_flush else "step outputs (values, W) exceed L2",
                   "elements_with_status_flags_rank0": status_bad, "elements_with_dropped_enrichment_rank0": status_dropped},
        "stage_ms": ({"alpha": t_alpha / args.steps, "assemble_condense": t_asm / args.steps, "csr_numeric": t_csr / args.steps} if args.two_pass
                     else {"alpha": t_alpha / args.steps, "assemble_condense_csr": t_asm / args.steps}),
        "setup_ms_per_mesh": {"max_over_ranks": setup_ms_max, "rank0": setup_ms,
                              "what": "marks, pruning, pattern of the own rows, element selection, plan (outside the step: once per mesh)",
                              "stages_ms_rank0": {k: round(v, 3) for k, v in getattr(P, "part_timing", {}).items()}},
        "roofline": {"bound": "fp64", "kernel": f"fast_assemble_kernel<{order}>", "achieved": achieved_tf, "peak": fp64, "unit": "TFLOP/s",
                     "frac": (achieved_tf / fp64) if achieved_tf else None, "traffic": traffic, "traffic_note": traffic_note, "kernel_ms": asm_ms,
                     "peak_source": fp64_src, "flops_per_element": f_asm, "moment_rhs_fraction": moment_frac,
                     "note": "FP64-FMA-bound kernel (neither hbm nor tensor); flops = tensor algorithm actually executed, "
                             "no credit for the quadrature work it avoids",
                     "algorithmic_hbm_bytes_per_launch": int((store_doubles + ps.W_doubles + ps.z_doubles) * 8 + n_poly * (80 if args.two_pass else 144)),
                     "alpha_kernel": {"kernel": f"fast_alpha_kernel<{order}>", "kernel_ms": alpha_ms, "achieved": alpha_tf,
                                      "frac": (alpha_tf / fp64) if alpha_tf else None, "flops_per_element": f_alpha},
                     "step_a4_a7": {"achieved": step_tf, "frac": step_tf / fp64, "flops_per_element": f_asm + f_alpha},
                     "quadrature_algorithm_flops_per_element": int(sizes.flops_quad_elem),
                     "hbm_gbs_of_outputs": out_bytes / (ms_step * 1e-3) / 1e9, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src},
        "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "repetitions_ms_per_step": [round(t * 1e3, 3) for t in e2e_reps],
                "what": "per step: pinned mesh arrays -> device, per-mesh set-up, the step, alpha table + status -> host (per rank at N > 1)",
                "pipeline": "2 slots: the copy stream moves step i+1's mesh in and step i-1's alpha table + status out while step i computes"},
        # two-pass: alpha, assemble, gather (+ alpha_general, alpha of the neighbours, assemble_general);
        # fused: alpha, assemble, diagonal merge (+ the same three and the scatter of the quadrature-path blocks)
        "gpu_launches": int(args.steps * (3 + ((3 if args.two_pass else 4) if n_gen else 0))),
        "clocks": clocks.summary(),
    }
    if replica_value is not None:
        line["replicas_weak_value"] = replica_value
    for sv in solves:
        sv["spmv_frac_of_hbm_peak"] = sv["spmv_gbs_actual"] / hbm
        sv["spmv_frac_of_hbm_peak_csr_equivalent_bytes"] = sv["spmv_gbs_csr_equivalent"] / hbm
    if solves:
        line["solve"] = solves[0]
        if len(solves) > 1:
            line["solve_hdg"] = solves[1]
        if len(solves) > 2:
            line["solve_config3_enriched_1M"] = solves[2]
    if e2e_solve is not None:
        line["e2e_solve"] = e2e_solve
    if not args.no_cpu and world == 1:
        try:
            info = cpu_reference(args.workload, budget_s=15.0)
            line["cpu_baseline"] = {"value": info["value"], "unit": "elements/s", "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"]}
        except Exception as exc:          # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e_solve(dev, workload="ehdg_p3_eps1e-4_structured_100k"):
    """BASELINE config 2 through the reference-facing class: host config in, error table out (mesh generation, upload, a1-a10)."""
    import torch
    from convectiondiffusionequation_b200 import boundary_layer, x, y
    from convectiondiffusionequation_b200.convection_diffusion import Convection_Diffusion
    n, kind, order, eps = WORKLOADS[workload]
    p, q = boundary_layer(x, BETA[0], eps), boundary_layer(y, BETA[1], eps)
    out = {"workload": workload}
    for typ, enr in (("ehdg", True), ("hdg", False)):
        cfg = {'order': [order], 'beta': BETA, 'mesh_size': [1.0 / n], 'epsilon': eps, 'exact': p * q, 'coeff': BETA[1] * p + BETA[0] * q,
               'bonus_int': 10, 'theta': 1e-5, 'enrich_functions': [p, q] if enr else [],
               'enrich_domain_ind': [lambda X, Y, h: X > 1 - h, lambda X, Y, h: Y > 1 - h] if enr else [], 'mesh_kind': kind,
               'device': str(dev), 'raise_on_nonconvergence': False, 'quiet': True}
        CT = Convection_Diffusion(cfg)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res, _ = CT._solveEHDG()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        si = CT.solver_info[-1]
        out[typ] = {"wall_s": dt, "elements": 2 * n * n, "elements_per_s": 2 * n * n / dt, "l2_error": float(res['Error'].iloc[-1]),
                    "iters": si["iters"], "converged": si["converged"], "rel_residual": si["rel_residual"], "stages_s": si.get("stages_s")}
    return out


if __name__ == "__main__":
    main()
