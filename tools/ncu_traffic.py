"""Reads an `ncu --set full` report of the default bench command and writes profiles/r02_assemble_traffic.json: DRAM bytes per
launch of the dominant kernel (fast_assemble_kernel<P>) plus the pipe / stall figures quoted in DESIGN.md, stamped with the hash
of the kernel sources it was captured on (bench.py quotes `roofline.traffic` only while that hash matches).
    python tools/ncu_traffic.py gpurun_out/asm_r02.ncu-rep ehdg_p4_eps1e-6_unstructured_4M"""
import csv, io, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if "fast_assemble_kernel" in r[col["Kernel Name"]]]
    if not sel:
        raise SystemExit("no fast_assemble_kernel launch in the report")

    def metric(name, scale_units=False):
        vals = []
        for r in sel:
            v = float(r[col[name]].replace(",", ""))
            if scale_units:
                v *= UNIT.get(units[col[name]], 1.0)
            vals.append(v)
        return sum(vals) / len(vals)

    rec = {"workload": workload, "kernel": sel[0][col["Kernel Name"]][:60], "launches_captured": len(sel),
           "captured": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "kernel_source_sha256_16": kernel_source_hash(),
           "dram_bytes_per_launch": metric("dram__bytes_read.sum", True) + metric("dram__bytes_write.sum", True),
           "dram_read_bytes": metric("dram__bytes_read.sum", True), "dram_write_bytes": metric("dram__bytes_write.sum", True),
           "gpu_time_ms_under_ncu": metric("gpu__time_duration.sum") / 1e6 if units[col["gpu__time_duration.sum"]] in ("nsecond", "ns") else metric("gpu__time_duration.sum"),
           "registers_per_thread": metric("launch__registers_per_thread")}
    for k in ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"):
        if k in col:
            rec[k] = metric(k)
    json.dump(rec, open(os.path.join(ROOT, "profiles", "r02_assemble_traffic.json"), "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
