"""Where does the host-fed (e2e) step lose time?  Development probe, not a benchmark of record.

Prints: raw H2D bandwidth of one pinned 38.5 MB batch alone and while the training step runs; the e2e loop with
per-step host timestamps (spiky or uniform?); the e2e loop with the copy issued on the compute stream instead.
"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cross-scale-mae_b200"))
import torch, csmae_b200
from csmae_b200 import DevicePrefetcher
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = csmae_b200.mae_vit_base_patch16(input_size=224, device=str(dev)).to(dev).train()
opt = csmae_b200.FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1.5e-4, betas=(0.9, 0.95), model=model)
x1 = torch.randn(64, 3, 224, 224, device=dev); x2 = torch.randn(64, 3, 224, 224, device=dev)
h1, h2 = x1.cpu().pin_memory(), x2.cpu().pin_memory()
print("pinned:", h1.is_pinned(), h2.is_pinned())
def step(a, b):
    opt.zero_grad(set_to_none=True)
    loss, _, _ = model(a, b, 0.75)
    loss.backward()
    opt.step()
    return loss
for _ in range(6): step(x1, x2)
torch.cuda.synchronize()

def h2d_times(n, busy):
    side = torch.cuda.Stream()
    d = torch.empty_like(x1)
    out = []
    for _ in range(n):
        if busy:
            step(x1, x2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(); d.copy_(h1, non_blocking=True); e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out
import subprocess, threading
MODE = sys.argv[1] if len(sys.argv) > 1 else "none"
proc = None
stop = threading.Event()
if MODE == "smi":
    proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                             "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
elif MODE.startswith("nvml"):
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    full = MODE == "nvml_full"
    def poll():
        n = 0
        while not stop.is_set():
            t0 = time.perf_counter()
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
            if full:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)
            dt = time.perf_counter() - t0
            n += 1
            if n <= 3: print(f"nvml query took {dt*1e3:.2f} ms")
            stop.wait(0.2)
    threading.Thread(target=poll, daemon=True).start()
print("sampler mode:", MODE)
time.sleep(1.0)
for busy in (False, True):
    t = h2d_times(12, busy)
    print(f"H2D 38.5 MB, compute {'busy' if busy else 'idle'}: min {min(t):.2f} med {statistics.median(t):.2f} max {max(t):.2f} ms "
          f"({38.535 / statistics.median(t):.1f} GB/s)")

def e2e_trace(k, lag, prefetch=True):
    stamps = []
    prev = None
    it = DevicePrefetcher(((h1, h2) for _ in range(k)), dev) if prefetch else \
        ((h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)) for _ in range(k))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b in it:
        ta = time.perf_counter()
        l = step(a, b)
        tb = time.perf_counter()
        if lag:
            if prev is not None: prev.item()
            prev = l
        else:
            l.item()
        stamps.append((ta - t0, tb - ta, time.perf_counter() - tb))
        t0 = time.perf_counter()
    if prev is not None: prev.item()
    torch.cuda.synchronize()
    return stamps
for name, lag, pf in (("prefetch lag1", 1, True), ("prefetch immediate", 0, True)):
    for rep in range(4):
        e2e_trace(3, lag, pf)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = e2e_trace(20, lag, pf)
        tot = (time.perf_counter() - t0) / 20 * 1e3
        print(f"{name}: {tot:.2f} ms/step; per-step [iter-overhead, enqueue, read-wait] ms:",
              " ".join(f"[{a*1e3:.1f},{b*1e3:.1f},{c*1e3:.1f}]" for a, b, c in st[:20]) if tot > 16 else "")
stop.set()
if proc is not None:
    proc.terminate(); proc.wait()
