#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | grep -v "^$" | tail -3
timeout 300 python tools/opbench.py --only attn 2>&1 | tail -4
timeout 300 python tools/opbench.py --only gemm 2>&1 | grep "enc.fc1\|dec.fc1\|dec.proj" 
timeout 300 python bench.py --quick --steps 30 --warmup 5 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
