#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" --timeout 300 2>&1 | grep -v "^$" | tail -4
CSMAE_LIB=cross-scale-mae_b200/csmae_b200/lib/libcsmae_b200_timing.so timeout 120 python tools/attn_phase.py 128 197 16 32 2>&1 | tail -14
timeout 300 python tools/opbench.py --only attn 2>&1 | tail -5
