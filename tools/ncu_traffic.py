#!/usr/bin/env python
"""DRAM traffic per GEMM launch from an `ncu --set full` capture of one step (tools/step_for_ncu.py), written with the
identity of the build it measured so that bench.py can refuse a stale file:

    python tools/ncu_traffic.py gpurun_out/r2_gemm_step.ncu-rep profiles/r2_gemm_dram_traffic.json
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * mul


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    gemm = [r for r in data if "gemm_kernel" in r[idx["Kernel Name"]]]
    rd = [to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) for r in gemm]
    wr = [to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in gemm]
    dur = [float(r[idx["gpu__time_duration.sum"]]) for r in gemm]
    with open(os.path.join(ROOT, "cross-scale-mae_b200", "csrc", "gemm_tcgen05.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    doc = {"source": f"{os.path.basename(rep)}: ncu --set full --clock-control none over ONE un-graphed step of the "
                     f"default bench workload (tools/step_for_ncu.py), cold caches per pass",
           "gemm_source_sha256": sha, "launches": len(gemm), "launches_per_step": len(gemm), "unit": "bytes",
           "avg_bytes_per_launch": (sum(rd) + sum(wr)) / max(1, len(gemm)),
           "dram_read_per_step": sum(rd), "dram_write_per_step": sum(wr),
           "sum_duration_us_under_ncu": sum(dur) if units[idx["gpu__time_duration.sum"]] == "us" else sum(dur) / 1e3}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    main()
