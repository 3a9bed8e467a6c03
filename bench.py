#!/usr/bin/env python
"""Benchmark of the Cross-Scale MAE pretraining hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch base|large] [--batch B] [--input-size S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

A step = one full training step of MAE_ViT_MsLdCeCd on one synthetic two-scale batch:
zero_grad -> forward(imgs1, imgs2, 0.75) -> backward -> AdamW.step (betas 0.9/0.95, wd 0.05, as
main_pretrain.py:426-427).  `value` times it with the batch already resident in HBM; `e2e` times the
same step through the public nn.Module call with the batch in pinned HOST memory (H2D copy and the
loss read-back inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "cross-scale-mae_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "pretrain images/sec (two-scale)"
MASK_RATIO = 0.75
ARCH = {
    "base": dict(dim_model=768, encoder_num_layers=12, encoder_num_heads=12, decoder_embed_dim=512,
                 decoder_num_layers=8, decoder_num_heads=16),
    "large": dict(dim_model=1024, encoder_num_layers=24, encoder_num_heads=16, decoder_embed_dim=512,
                  decoder_num_layers=8, decoder_num_heads=16),
}


# ------------------------------------------------------------------------------------------------
# algorithmic FLOPs: 2*M*N*K per GEMM, 4*S^2*D per attention layer; bwd = 2x fwd.
# Patch-embed is counted on the kept patches only (the build skips the 75 % the reference wastes).
# ------------------------------------------------------------------------------------------------
def flops_per_image(arch, size, patch=16, mask_ratio=MASK_RATIO, hidden=2048):
    a = ARCH[arch]
    D, Le, Dd, Ld = a["dim_model"], a["encoder_num_layers"], a["decoder_embed_dim"], a["decoder_num_layers"]
    L = (size // patch) ** 2
    keep = int(L * (1 - mask_ratio))
    Se, Sd, P = keep + 1, L + 1, patch * patch * 3
    embed = 2 * Se * P * D
    enc_lin, enc_att = Le * 24 * D * D * Se, Le * 4 * Se * Se * D
    dec_embed = 2 * Se * D * Dd
    dec_lin, dec_att = Ld * 24 * Dd * Dd * Sd, Ld * 4 * Sd * Sd * Dd
    dec_pred = 2 * Sd * Dd * P
    predictor = 2 * Sd * Dd * hidden * 2
    gemm_fwd = 2 * (embed + enc_lin + dec_embed + dec_lin + dec_pred) + predictor
    attn_fwd = 2 * (enc_att + dec_att)
    return dict(gemm_fwd=gemm_fwd, attn_fwd=attn_fwd, fwd=gemm_fwd + attn_fwd, fwd_bwd=3 * (gemm_fwd + attn_fwd))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"], hbm_gbs=d["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs: ONE long-running `nvidia-smi -lms 200`
    (the recipe's clocks line, B200_PROFILING.md) started before and killed after.  (Spawning nvidia-smi per sample
    re-initialises NVML every time, which stalls the driver for ~0.1 s and showed up as +5..8 ms per step in the
    host-fed e2e region.)"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None
        self.reader = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.reader = threading.Thread(target=self._read, daemon=True)
        self.reader.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()            # the exact process started above
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.reader is not None:
                self.reader.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle/restatement.py, pinned against the real reference by
# tests/test_oracle_golden.py) on the host cores.  /root/reference itself does not exist on the GPU box.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_time(arch, size, batch, steps, warmup):
    from oracle import restatement as R
    torch.set_num_threads(os.cpu_count() or 1)
    a = ARCH[arch]
    sd = R.make_state_dict(**a, input_size=size, patch_size=16, seed=0)
    leaves = {k: v.clone().requires_grad_(k not in R.FROZEN_KEYS) for k, v in sd.items()}
    opt = torch.optim.AdamW([v for v in leaves.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.95),
                            weight_decay=0.05)
    g = torch.Generator().manual_seed(1000)
    L = (size // 16) ** 2
    x1, x2 = torch.randn(batch, 3, size, size, generator=g), torch.randn(batch, 3, size, size, generator=g)
    rm, rv = torch.zeros(L), torch.ones(L)
    times = []
    for i in range(warmup + steps):
        n1, n2 = torch.rand(batch, L, generator=g), torch.rand(batch, L, generator=g)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = R.cross_scale_forward(leaves, x1, x2, n1, n2, MASK_RATIO, a["encoder_num_heads"], a["decoder_num_heads"],
                                    running=(rm, rv))
        out["loss"].backward()
        opt.step()
        out["loss"].item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu_batch = args.cpu_batch
    sec = cpu_reference_step_time(args.arch, args.input_size, cpu_batch, args.steps, max(1, min(args.warmup, 1)))
    ips = cpu_batch / sec
    cores = torch.get_num_threads()
    sample = (f"{args.steps} timed steps of fwd+bwd+AdamW, each on a {cpu_batch}-image sample of the workload's batch "
              f"(images/s is batch-insensitive on CPU), fp32, all host threads, oracle/restatement.py")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch, max(1, args.gpus)),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, batch, world):
    return {"workload": f"MAE_ViT_MsLdCeCd vit_{args.arch}_patch16 two-scale ({args.input_size}+{args.input_size}) "
                        f"bs={batch}/GPU mask_ratio=0.75, fwd+bwd+AdamW",
            "global_batch": batch * world, "input_size": args.input_size,
            "parallelism": (f"dp{world} ({'torch DDP' if getattr(args, 'torch_ddp', False) else 'engine-overlapped NCCL all-reduce'})"
                            if world > 1 else "single"),
            "optimizer": "torch.optim.AdamW(fused=True)" if getattr(args, "torch_adamw", False) else "csmae_b200.FusedAdamW",
            "l2_policy": "per-step working set (GBs of activations + 1.4 GB weights/grads/moments) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="csmae_b200", choices=["csmae_b200", "reference"])
    ap.add_argument("--arch", default="base", choices=["base", "large"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 64 base / 32 large)")
    ap.add_argument("--input-size", type=int, default=224)
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-ddp", action="store_true", help="N>1: wrap with torch's DistributedDataParallel")
    ap.add_argument("--torch-adamw", action="store_true", help="torch.optim.AdamW(fused=True) instead of FusedAdamW")
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel CUDA-event breakdown")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 64 if args.arch == "base" else 32
    if args.impl == "reference":
        run_reference(args)
        return

    import csmae_b200
    from csmae_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    torch.manual_seed(0)
    ctor = csmae_b200.mae_vit_base_patch16 if args.arch == "base" else csmae_b200.mae_vit_large_patch16
    model = ctor(input_size=args.input_size, device=str(dev)).to(dev).train()
    step_model = model
    if world > 1:
        # same constructor call as main_pretrain.py:417-421; csmae_b200's wrapper lets the engine all-reduce the
        # flat gradient buffer in segments overlapped with the backward (--torch-ddp: torch's own wrapper)
        wrapper = torch.nn.parallel.DistributedDataParallel if args.torch_ddp else csmae_b200.DistributedDataParallel
        step_model = wrapper(model, device_ids=[local_rank], find_unused_parameters=True)
    decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.ndim == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.ndim == 1 or n.endswith(".bias"))]
    groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}]
    if args.torch_adamw:
        opt = torch.optim.AdamW(groups, lr=1.5e-4, betas=(0.9, 0.95), fused=True)
    else:
        # row f1: one kernel for the whole AdamW step, which also rewrites the bf16 shadow weights
        opt = csmae_b200.FusedAdamW(groups, lr=1.5e-4, betas=(0.9, 0.95), model=model)

    B, S = args.batch, args.input_size
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    imgs1 = torch.randn(B, 3, S, S, device=dev, generator=g)
    imgs2 = torch.randn(B, 3, S, S, device=dev, generator=g)
    host1 = imgs1.cpu().pin_memory()
    host2 = imgs2.cpu().pin_memory()
    torch.manual_seed(1 + rank)      # masking noise: per-rank stream (main_pretrain.py:368-369)

    def step(x1, x2):
        opt.zero_grad(set_to_none=True)
        loss, _, _ = step_model(x1, x2, MASK_RATIO)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / k

    for _ in range(max(args.warmup, 3)):
        loss = step(imgs1, imgs2)
    assert torch.isfinite(loss).item(), "non-finite loss"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count
    ms_step = timed(lambda: step(imgs1, imgs2), args.steps)
    launches = _native.launch_count - launches0

    # fwd+bwd only (the BASELINE metric's second figure)
    def fwd_bwd():
        opt.zero_grad(set_to_none=True)
        l_, _, _ = step_model(imgs1, imgs2, MASK_RATIO)
        l_.backward()
    ms_fwd_bwd = timed(fwd_bwd, max(3, args.steps // 2))

    # host-side cost of enqueueing one step (no synchronisation inside): if this approaches ms_per_step the
    # path is launch-bound on the CPU
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step(imgs1, imgs2)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    torch.cuda.synchronize()

    # end to end through the public API: every step's batch starts in pinned HOST memory, is copied to the
    # device (csmae_b200.DevicePrefetcher: the copy of step i+1 runs on a side stream while step i trains),
    # trains, and the loss is read back to the host
    from csmae_b200 import DevicePrefetcher

    class HostBatches:                      # a re-iterable "loader" of k pinned host batches
        k = 0

        def __iter__(self):
            return ((host1, host2) for _ in range(self.k))
    loader = HostBatches()
    prefetcher = DevicePrefetcher(loader, dev)          # one instance, as a training script keeps it across epochs

    def e2e_run(k, lag):
        # every step's loss is read back to the host; with lag=1 the read of step i is issued after step i+1 has
        # been enqueued (asynchronous logging: the GPU never waits for the host between steps)
        last, prev = None, None
        loader.k = k
        for x1, x2 in prefetcher:
            loss_t = step(x1, x2)
            if lag:
                if prev is not None:
                    last = prev.item()
                prev = loss_t
            else:
                last = loss_t.item()
        if prev is not None:
            last = prev.item()
        return last

    def e2e_timed(lag):
        e2e_run(3, lag)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps, lag)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms
    ms_e2e_sync = e2e_timed(0)       # loss.item() right after every step, as engine_pretrain.py:55 does
    ms_e2e = e2e_timed(1)            # same reads, one step late
    sampler.stop()

    # per-kernel breakdown of one step with CUDA events on the launching stream (outside the timed region)
    model._engine.use_graphs = False          # the breakdown times each C-ABI call eagerly
    breakdown = kernel_breakdown(_native, lambda: step(imgs1, imgs2))     # every rank: the step all-reduces

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fl = flops_per_image(args.arch, S)
    peaks = measured_peaks()
    ips = B * world / (ms_step * 1e-3)
    ips_e2e = B * world / (ms_e2e * 1e-3)
    # DRAM bytes per GEMM launch from the committed ncu --set full capture of the same shapes (ViT-B only)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1e_gemm_dram_traffic.json")
    if args.arch == "base" and args.input_size == 224 and B == 64 and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["avg_bytes_per_launch"]
    gemm_ms = sum(v["ms"] for k, v in breakdown.items() if k.startswith("csm_linear"))
    gemm_n = sum(v["n"] for k, v in breakdown.items() if k.startswith("csm_linear"))
    step_ms_prof = sum(v["ms"] for v in breakdown.values())
    gemm_tflops = 3 * fl["gemm_fwd"] * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05; csm_linear_fwd/dgrad/wgrad, all launches of one step)",
                "achieved": gemm_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": gemm_tflops / peaks["tflops_sustained"], "traffic": traffic,
                "launches_per_step": gemm_n, "avg_launch_us": gemm_ms * 1e3 / gemm_n if gemm_n else None,
                "flops_per_launch_avg": 3 * fl["gemm_fwd"] * B / gemm_n if gemm_n else None,
                "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
                "share_of_step": gemm_ms / step_ms_prof if step_ms_prof else None,
                "step_frac_of_flop_roofline": ips / world * fl["fwd_bwd"] / (peaks["tflops_sustained"] * 1e12)}
    line = {"metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B, world),
            "fwd_bwd_ms": ms_fwd_bwd, "host_enqueue_ms_per_step": host_ms, "gpu_launches": launches,
            "e2e": {"value": ips_e2e, "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(host1.numel() * 4 * 2), "d2h_bytes_per_step": 4,
                    "loss_read": "every step, issued after the next step is enqueued (1-step lag)",
                    "immediate_read_ms_per_step": ms_e2e_sync,
                    "immediate_read_value": B * world / (ms_e2e_sync * 1e-3)},
            "clocks": sampler.summary(), "roofline": roofline,
            "flops_per_image": fl, "kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(
                breakdown.items(), key=lambda kv: -kv[1]["ms"])}}
    if world == 1 and not args.no_cpu_baseline:
        sec = cpu_reference_step_time(args.arch, S, args.cpu_batch, 2, 1)
        line["cpu_baseline"] = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"2 timed steps of fwd+bwd+AdamW at batch {args.cpu_batch}, "
                                                          f"fp32, oracle/restatement.py (reference algorithm)"}
    if args.profile_kernels:
        for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"# {k:28s} {v['ms']:9.3f} ms  {v['n']:5d} launches", file=sys.stderr)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_breakdown(native, fn):
    """Times every C-ABI call of one step with CUDA events recorded on the launching stream."""
    events = []
    orig = native.call

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *a)
        e1.record()
        events.append((name, e0, e1))
        return r
    import csmae_b200.engine as eng
    native.call = wrapped
    eng.call = wrapped
    try:
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
    finally:
        native.call = orig
        eng.call = orig
    out = {}
    for name, e0, e1 in events:
        d = out.setdefault(name, {"ms": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
    return out


if __name__ == "__main__":
    main()
