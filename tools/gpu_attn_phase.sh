#!/bin/bash
# Phase breakdown of the attention kernels.  Build the timing variant of the library first (here, before gpurun -- the
# .so travels with the snapshot):
#   CSM_NVCC_EXTRA=-DCSM_ATTN_TIMING python -c "import sys; sys.path.insert(0, 'cross-scale-mae_b200'); \
#       from csmae_b200 import build as b; b.LIB_PATH = b.LIB_PATH.replace('.so', '_timing.so'); b.STAMP += '.timing'; \
#       b.build(force=True)"; python __graft_entry__.py      # (the second call restores the objects of the product build)
export CSMAE_LIB=cross-scale-mae_b200/csmae_b200/lib/libcsmae_b200_timing.so
python tools/attn_phase.py 128 197 16 32
python tools/attn_phase.py 128 50 12 64
python tools/attn_phase.py 32 785 16 32
