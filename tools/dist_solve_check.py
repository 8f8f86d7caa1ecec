"""Multi-GPU facet solve check (launch with torchrun, one rank per GPU): every rank assembles the same HDG
system, the Krylov iteration runs row-partitioned over NCCL (halo exchange + all-reduced dots) and the result is
compared with the single-GPU solve of rank 0's own copy.  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import make_spec
from convectiondiffusionequation_b200 import parallel
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("DS_N", "64"))
    order, eps = int(os.environ.get("DS_P", "3")), float(os.environ.get("DS_EPS", "1e-6"))
    method, k = os.environ.get("DS_METHOD", "gmres"), int(os.environ.get("DS_K", "40"))
    precond = os.environ.get("DS_PRECOND", "line")       # of the partitioned solve; the replicated-matrix mode has facet blocks only
    iters = int(os.environ.get("DS_ITERS", "20000"))
    mesh = unit_square_mesh(n, kind=os.environ.get("DS_KIND", "unstructured"))
    spec = make_spec(order, eps, enriched=False)
    # single-GPU reference (every rank computes it; cheap at the check size)
    ref = None
    if os.environ.get("DS_REF", "1") == "1":
        Pr = EHDGPipeline(mesh, spec, device=dev).setup().assemble().csr_build()
        Pr.gmres(restart=k, max_iters=iters, rtol=1e-10, method=method, precond="facet")
        ref = (Pr.xhat.clone(), Pr.gmres_info.iters, Pr.gmres_info.total_seconds, Pr.gmres_info.rel_residual)
        Pr.close()
    P = EHDGPipeline(mesh, spec, device=dev).setup().assemble().csr_build()
    parallel.init_communicator(P)
    rb, lo, hi = parallel.set_partition(P, world)
    dist.barrier()
    P.gmres(restart=k, max_iters=iters, rtol=1e-10, method=method, strip_block=False, precond="facet")
    gi = P.gmres_info
    out = dict(world=world, n=n, p=order, eps=eps, method=method, rows=int(P.csr_sizes.n_rows), iters=int(gi.iters), converged=bool(gi.converged),
               rel_residual=float(gi.rel_residual), seconds=float(gi.total_seconds), rows_per_rank=[int(b - a) for a, b in zip(rb[:-1], rb[1:])],
               halo_rows=[int((hi[r] - lo[r]) - (rb[r + 1] - rb[r])) for r in range(world)])
    if ref is not None:
        x1, it1, t1, res1 = ref
        out.update(single_iters=int(it1), single_seconds=float(t1), single_rel_residual=float(res1),
                   rel_diff_vs_single=float((P.xhat - x1).norm() / x1.norm()))
    # all ranks hold the same full solution
    chk = torch.tensor([float(P.xhat.double().sum())], dtype=torch.float64, device=dev)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    out["solution_identical_on_all_ranks"] = bool(all(float(t) == float(lst[0]) for t in lst))
    P.backsolve()
    l2, _ = P.errors()
    out["l2_error"] = l2
    out["elements"] = int(mesh.ne)
    # partitioned assembly: every rank assembles and condenses only the elements that touch its rows
    if os.environ.get("DS_PART", "1") == "1":
        P.close()
        P = EHDGPipeline(mesh, spec, device=dev)
        parallel.init_communicator(P)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        dist.barrier(); t0.record()
        parallel.partitioned_assemble(P, rank, world)
        t1.record(); torch.cuda.synchronize()
        l2p, h1p = parallel.partitioned_solve(P, restart=k, max_iters=iters, rtol=1e-10, method=method, precond=precond)
        gi = P.gmres_info
        cnt = torch.tensor([int(P.plan_sizes.n_poly + P.plan_sizes.n_gen), int(P.elem_owned.sum()), int(P.csr_sizes.nnz)],
                           dtype=torch.int64, device=dev)
        lst = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(lst, cnt)
        pa = dict(precond=precond, chains=int(gi.line_chains), iters=int(gi.iters), converged=bool(gi.converged), rel_residual=float(gi.rel_residual), seconds=float(gi.total_seconds),
                  l2_error=l2p, h1_error=h1p, setup_assemble_ms=float(t0.elapsed_time(t1)),
                  elements_per_rank=[int(t[0]) for t in lst], owned_per_rank=[int(t[1]) for t in lst], nnz_per_rank=[int(t[2]) for t in lst])
        if ref is not None:
            pa["rel_diff_vs_single"] = float((P.xhat - ref[0]).norm() / ref[0].norm())
        out["partitioned"] = pa
    if rank == 0:
        print(json.dumps(out), flush=True)
    P.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
