"""bench.py's output contract, checked on the CPU: the argument surface the driver uses, the reference arm's line shape, and the
last recorded GPU line (profiles/r02l_bench_1gpu_4M.json) against the key set the task's measurement section asks for."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_recorded_line_has_every_contract_key():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02l_bench_1gpu_4M.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["config"]["workload"] == "ehdg_p4_eps1e-6_unstructured_4M" and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1 and line["warmup"] >= 3
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] is not None
    c = line["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] == "port" and c["cores"] >= 1
    e = line["e2e"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]      # not a copy of the device-timed value
    assert line["gpu_launches"] > 0 and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = elements / step time
    assert abs(line["value"] - line["config"]["elements"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]


def test_reference_arm_line_on_a_bounded_sample():
    """`bench.py --impl reference` runs the C/OpenMP port of the path on the host cores (no GPU): same metric / unit / config,
    impl = reference, e2e with zero copy bytes"""
    env = dict(os.environ, EHDG_BENCH_REF_BUDGET_S="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "ehdg_p3_eps1e-4_structured_100k"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "condensed EHDG elements/sec (FP64)" and line["unit"] == "elements/s"
    assert line["config"]["workload"] == "ehdg_p3_eps1e-4_structured_100k" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
