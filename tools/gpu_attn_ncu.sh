#!/bin/bash
# ncu --set full with source counters of the decoder-shape attention kernels (one launch each)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_tc --launch-skip 2 -c 1 \
    -f -o gpurun_out/r2f_attn_bwd_dec python tools/attn_tc_prof.py 128 197 16 32 bwd > gpurun_out/prof1.log 2>&1
ls -la gpurun_out/*.ncu-rep
