import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/cross-scale-mae_b200")
import torch, csmae_b200
from csmae_b200 import DevicePrefetcher
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = csmae_b200.mae_vit_base_patch16(input_size=224, device=str(dev)).to(dev).train()
decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.ndim == 1 or n.endswith(".bias"))]
no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.ndim == 1 or n.endswith(".bias"))]
opt = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}], lr=1.5e-4, betas=(0.9, 0.95), fused=True)
x1 = torch.randn(64, 3, 224, 224, device=dev); x2 = torch.randn(64, 3, 224, 224, device=dev)
h1, h2 = x1.cpu().pin_memory(), x2.cpu().pin_memory()
def step(a, b):
    opt.zero_grad(set_to_none=True)
    loss, _, _ = model(a, b, 0.75)
    loss.backward()
    opt.step()
    return loss
for _ in range(6): step(x1, x2)
def timed(fn, k=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("resident, no sync      ", timed(lambda: step(x1, x2)))
print("resident, item per step", timed(lambda: step(x1, x2).item()))
def e2e(k):
    for a, b in DevicePrefetcher(((h1, h2) for _ in range(k)), dev):
        step(a, b).item()
e2e(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); e2e(20); torch.cuda.synchronize()
print("prefetched h2d, item   ", (time.perf_counter() - t0) / 20 * 1e3)
def e2e_nosync(k):
    for a, b in DevicePrefetcher(((h1, h2) for _ in range(k)), dev):
        step(a, b)
torch.cuda.synchronize(); t0 = time.perf_counter(); e2e_nosync(20); torch.cuda.synchronize()
print("prefetched h2d, no sync", (time.perf_counter() - t0) / 20 * 1e3)
# host time from step start to forward-graph launch
import csmae_b200.engine as E
