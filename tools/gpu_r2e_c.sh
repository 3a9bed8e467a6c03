#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/opbench.py --cublas --json gpurun_out/r2e_opbench_cublas.json > gpurun_out/r2e_opbench.txt 2>&1
tail -5 gpurun_out/r2e_opbench.txt
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2e_bench_base_nocpu.json 2> gpurun_out/err.log || tail -5 gpurun_out/err.log
cut -c1-300 gpurun_out/r2e_bench_base_nocpu.json
