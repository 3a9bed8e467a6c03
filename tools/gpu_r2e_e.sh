#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2e_pytest_gpu.txt
timeout 300 python tools/opbench.py --only attn 2>&1 | tail -4
timeout 300 python bench.py --quick --steps 30 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
timeout 300 python bench.py --quick --arch large --steps 20 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
timeout 300 python bench.py --quick --arch large --batch 16 --input-size 448 --steps 15 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
