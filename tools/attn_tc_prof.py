"""Runs the tcgen05 attention kernels a few times on one shape (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-scale-mae_b200"))
import torch
from csmae_b200 import _native as nat

B, S, H, d = (int(x) for x in sys.argv[1:5])
mode = sys.argv[5] if len(sys.argv) > 5 else "fwd"
Dm = H * d
torch.manual_seed(0)
qkv = torch.randn(B * S, 3 * Dm, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, Dm, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device="cuda")
d_out = torch.randn(B * S, Dm, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
delta = torch.empty(B * H * S, device="cuda")
for _ in range(3):
    nat.call("csm_attention_fwd_tc", qkv, out, lse, B, S, H, d, 0)
    if mode == "bwd":
        nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dqkv, None, B, S, H, d)
torch.cuda.synchronize()
