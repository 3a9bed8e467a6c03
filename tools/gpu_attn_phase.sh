#!/bin/bash
# phase breakdown of the attention backward (timing build travels as lib/libcsmae_b200_timing.so)
export CSMAE_LIB=cross-scale-mae_b200/csmae_b200/lib/libcsmae_b200_timing.so
python tools/attn_phase.py 128 197 16 32
python tools/attn_phase.py 128 50 12 64
python tools/attn_phase.py 32 785 16 32
