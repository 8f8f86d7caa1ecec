"""Two-rank NCCL run of the distributed facet solve (needs >= 2 GPUs; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_solve_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DS_N="48", DS_P="3", DS_EPS="1e-6")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tools", "dist_solve_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["converged"] and out["solution_identical_on_all_ranks"]
    assert out["rel_diff_vs_single"] < 1e-9
    assert abs(out["iters"] - out["single_iters"]) <= 0.02 * out["single_iters"] + 2
