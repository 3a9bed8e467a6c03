#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear" --timeout 600 2>&1 | grep -v "^$" | tail -3
timeout 600 python tools/opbench.py --only gemm 2>&1 | grep "linear_fwd" 
for i in 1 2; do timeout 300 python bench.py --quick --steps 40 --warmup 5 2>/dev/null | cut -c1-120; done
