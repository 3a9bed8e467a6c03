#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -6
CSMAE_LIB=cross-scale-mae_b200/csmae_b200/lib/libcsmae_b200_timing.so python tools/attn_phase.py 128 197 16 32 | tail -14
timeout 300 python tools/opbench.py --only attn 2>&1 | tail -5
timeout 300 python bench.py --quick --steps 30 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
timeout 300 python bench.py --quick --arch large --batch 16 --input-size 448 --steps 15 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
