"""Per-phase cycle breakdown of the tcgen05 attention backward kernel (development build:
    CSM_NVCC_EXTRA=-DCSM_ATTN_TIMING python -m csmae_b200.build --force   (tools/gpu_attn_phase.sh does both)
    python tools/attn_phase.py B S H d"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-scale-mae_b200"))
import torch
from csmae_b200 import _native as nat

B, S, H, d = (int(x) for x in sys.argv[1:5])
Dm = H * d
torch.manual_seed(0)
qkv = torch.randn(B * S, 3 * Dm, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, Dm, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device="cuda")
d_out = torch.randn(B * S, Dm, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
delta = torch.empty(B * H * S, device="cuda")
lib = nat.load()
lib.csm_attn_phase_read.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    nat.call("csm_attention_fwd", qkv, out, lse, B, S, H, d)
lib.csm_attn_phase_read(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nat.call("csm_attention_fwd", qkv, out, lse, B, S, H, d)
e1.record()
lib.csm_attn_phase_read(buf)
n = max(buf[15], 1)
print(f"forward B={B} S={S} H={H} d={d}: {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented), {n} iterations over the sampled warps")
for i, nm in enumerate(["wait S", "tcgen05.ld + row max", "exp2 / sum / pack / store P", "whole iteration", "drain: wait P.V",
                        "drain: O read-back", "tcgen05.st wait"]):
    print(f"  {nm:28s} {buf[i] / n:8.0f} clk/iteration")
for _ in range(3):
    nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dqkv, None, B, S, H, d)
lib.csm_attn_phase_read(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dqkv, None, B, S, H, d)
e1.record()
lib.csm_attn_phase_read(buf)
n = max(buf[15], 1)
names = ["wait row statistics", "wait S,dP", "tcgen05.ld+release", "exp/dS math", "wait grads(n-1)", "st.shared P,dS",
         "read-back dQ/dK/dV", "fence+arrive", "MMA: wait P,dS", "MMA: issue grads", "MMA: wait sdp_free", "MMA: issue S,dP"]
print(f"B={B} S={S} H={H} d={d}: {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented), {n} sub-blocks over the sampled warps")
tot = sum(buf[i] for i in range(8))
for i, nm in enumerate(names):
    print(f"  {nm:22s} {buf[i] / n:8.0f} clk/sub-block" + (f"  ({100.0 * buf[i] / tot:4.1f}%)" if i < 8 else ""))
print(f"  softmax warp total     {tot / n:8.0f} clk/sub-block")
