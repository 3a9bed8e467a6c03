#!/usr/bin/env python
"""DRAM traffic per GEMM launch of ONE un-graphed step (tools/step_for_ncu.py) from an ncu CSV log
(`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_kernel --csv`), written
with the identity of the build it measured so that bench.py can refuse a stale file:

    python tools/ncu_traffic.py gpurun_out/r2d_gemm_traffic.csv profiles/r2_gemm_dram_traffic.json
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MUL = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    log, out = sys.argv[1], sys.argv[2]
    rows = list(csv.DictReader(l for l in open(log) if l.startswith('"')))
    per = {}
    for r in rows:
        if "gemm_kernel" not in r["Kernel Name"]:
            continue
        d = per.setdefault(r["ID"], {"kernel": r["Kernel Name"].split("(")[0].replace("void <unnamed>::", ""),
                                     "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * MUL[r["Metric Unit"]]
    launches = list(per.values())
    rd = sum(d["dram__bytes_read.sum"] for d in launches)
    wr = sum(d["dram__bytes_write.sum"] for d in launches)
    us = sum(d["gpu__time_duration.sum"] for d in launches)
    with open(os.path.join(ROOT, "cross-scale-mae_b200", "csrc", "gemm_tcgen05.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    by_kernel = {}
    for d in launches:
        k = by_kernel.setdefault(d["kernel"] + " grid=" + d["grid"], {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        k["us"] += d["gpu__time_duration.sum"]
    doc = {"source": f"{os.path.basename(log)}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                     f"--clock-control none over every GEMM launch of ONE un-graphed step of the default bench workload "
                     f"(tools/step_for_ncu.py; caches flushed per launch by ncu); the --set full view of the same kernels is "
                     f"profiles/*_ncu_full_gemm_summary.csv",
           "gemm_source_sha256": sha, "launches": len(launches), "launches_per_step": len(launches), "unit": "bytes",
           "avg_bytes_per_launch": (rd + wr) / max(1, len(launches)), "dram_read_per_step": rd, "dram_write_per_step": wr,
           "sum_duration_us_under_ncu": us, "by_kernel": by_kernel}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps({k: v for k, v in doc.items() if k != "by_kernel"}))


if __name__ == "__main__":
    main()
