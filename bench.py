#!/usr/bin/env python
"""Benchmark of the Cross-Scale MAE pretraining hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch base|large] [--batch B] [--input-size S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the VERBATIM reference classes (oracle/_ref) on the host CPU cores

A step = one full training step of MAE_ViT_MsLdCeCd on one synthetic two-scale batch:
zero_grad -> forward(imgs1, imgs2, 0.75) -> backward -> AdamW.step (betas 0.9/0.95, wd 0.05, as
main_pretrain.py:426-427).  `value` times it with the batch already resident in HBM; `e2e` times the
same step through the public nn.Module call with the batch in pinned HOST memory (H2D copy and an immediate
loss.item() read-back inside the timed region, as engine_pretrain.py:49-55 does); `e2e.stock_engine` is the
UNMODIFIED reference loop engine_pretrain.train_one_epoch driving this module.  One JSON line on stdout (rank 0);
the default run appends `extra_configs` (ViT-L/224 bs 32 and ViT-L/448 bs 16, BASELINE.json configs[2] and [4]).
"""
import argparse
import contextlib
import hashlib
import io
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
GEMM_SOURCE = os.path.join(ROOT, "cross-scale-mae_b200", "csrc", "gemm_tcgen05.cu")


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


def gemm_traffic(launches_per_step):
    """DRAM bytes per GEMM launch from the committed `ncu --set full` capture of THIS build's GEMM (tools/ncu_traffic.py
    writes the file with the sha256 of gemm_tcgen05.cu and the launch count it saw): anything else is not a
    measurement of the benched binary and is reported as null with the reason."""
    path = os.path.join(ROOT, "profiles", "r2_gemm_dram_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu --set full capture committed for this build"
    with open(path) as f:
        t = json.load(f)
    with open(GEMM_SOURCE, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    if t.get("gemm_source_sha256") != sha:
        return None, "profiles/r2_gemm_dram_traffic.json was captured on another build of gemm_tcgen05.cu"
    if t.get("launches_per_step") != launches_per_step:
        return None, (f"capture saw {t.get('launches_per_step')} GEMM launches per step, this run {launches_per_step}")
    return t["avg_bytes_per_launch"], f"ncu --set full, {t['launches']} launches ({t.get('source', '')})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs: ONE long-running `nvidia-smi -lms 200`
    (the recipe's clocks line, B200_PROFILING.md) started before and killed after.  (Spawning nvidia-smi per sample
    re-initialises NVML every time, which stalls the driver for ~0.1 s and showed up as +5..8 ms per step in the
    host-fed e2e region.)"""

    def __init__(self, index):
        self.index = index
        self.rows = []
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
# CPU arm.  kind "reference": the VERBATIM reference classes (oracle/_ref: models_mae/MAE_ViT_MsLdCeCd.py and its
# bases, copied by tools/vendor_ref.py; timm 0.4.12's Block / PatchEmbed -- a pip dependency absent from this image --
# come from oracle/timm_shim.py) through the reference's own forward (single input, scale 2 = its in-model
# RandomResizedCrop), loss.backward() and torch.optim.AdamW, fp32, all host threads.  kind "port": oracle/restatement.py
# when oracle/_ref did not travel.  /root/reference itself does not exist on the GPU box.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_time(arch, size, batch, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    a = ARCH[arch]
    g = torch.Generator().manual_seed(1000)
    from oracle import ref_loader
    if ref_loader.reference_available():
        _, _, RefCeCd = ref_loader.reference_classes()
        with contextlib.redirect_stdout(io.StringIO()):       # the reference constructors print
            torch.manual_seed(0)
            model = RefCeCd(**a, input_size=size, patch_size=16, input_channels=3, device="cpu").train()
        decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.ndim == 1 or n.endswith(".bias"))]
        no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.ndim == 1 or n.endswith(".bias"))]
        opt = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}],
                                lr=1.5e-4, betas=(0.9, 0.95))
        x = torch.randn(batch, 3, size, size, generator=g)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            loss, _, _ = model(x, mask_ratio=MASK_RATIO)
            loss.backward()
            opt.step()
            loss.item()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return sum(times) / len(times), "reference"
    from oracle import restatement as R
    sd = R.make_state_dict(**a, input_size=size, patch_size=16, seed=0)
    leaves = {k: v.clone().requires_grad_(k not in R.FROZEN_KEYS) for k, v in sd.items()}
    opt = torch.optim.AdamW([v for v in leaves.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.95),
                            weight_decay=0.05)
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
    return sum(times) / len(times), "port"


def cpu_sample_text(kind, steps, cpu_batch):
    what = ("the verbatim reference MAE_ViT_MsLdCeCd (oracle/_ref, timm Block from oracle/timm_shim.py): "
            "zero_grad, forward(imgs, mask_ratio=0.75) with its in-model crop, backward, torch.optim.AdamW.step"
            if kind == "reference" else "oracle/restatement.py (reference algorithm): fwd+bwd+AdamW")
    return (f"{steps} timed steps of {what}, each on a {cpu_batch}-image sample of the workload's batch "
            f"(images/s is batch-insensitive on the CPU), fp32, all host threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu_batch = args.cpu_batch
    sec, kind = cpu_reference_step_time(args.arch, args.input_size, cpu_batch, args.steps, 1)
    ips = cpu_batch / sec
    cores = torch.get_num_threads()
    # `config` names the workload this arm SAMPLES (the driver matches it against the GPU arm's line); what the arm
    # actually ran -- its own batch, optimizer, precision, device -- is spelled out in `reference_arm` and in
    # `cpu_baseline.sample`
    cfg = workload_config(args, args.batch, max(1, args.gpus))
    ref_arm = {"device": "cpu", "batch_per_step": cpu_batch, "dtype": "f32", "threads": cores,
               "optimizer": "torch.optim.AdamW (betas 0.9/0.95, wd 0.05)", "processes": 1,
               "form": "single-input forward, scale 2 by the model's own RandomResizedCrop"
               if kind == "reference" else "paired inputs (oracle/restatement.py)"}
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "reference_arm": ref_arm,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": cpu_sample_text(kind, args.steps, cpu_batch)},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, batch, world, arch=None, size=None):
    arch = arch or args.arch
    size = size or args.input_size
    return {"workload": f"MAE_ViT_MsLdCeCd vit_{arch}_patch16 two-scale ({size}+{size}) "
                        f"bs={batch}/GPU mask_ratio=0.75, fwd+bwd+AdamW",
            "global_batch": batch * world, "input_size": size,
            "parallelism": (f"dp{world} ({'torch DDP' if getattr(args, 'torch_ddp', False) else 'engine-overlapped NCCL all-reduce'})"
                            if world > 1 else "single"),
            "optimizer": "torch.optim.AdamW(fused=True)" if getattr(args, "torch_adamw", False) else "csmae_b200.FusedAdamW",
            "l2_policy": "per-step working set (GBs of activations + 1.4 GB weights/grads/moments) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("NCCL_IB_DISABLE", "1")
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def timed(self, fn, k):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / k


def measure(env, args, arch, B, S, steps, warmup, full):
    """One workload: device-resident `value`, host-fed `e2e`; with full=True also the clock samples, the stock-engine
    loop, the per-kernel breakdown and the roofline of the GEMM family."""
    import csmae_b200
    from csmae_b200 import _native
    world, rank, dev = env.world, env.rank, env.dev

    torch.manual_seed(0)
    ctor = csmae_b200.mae_vit_base_patch16 if arch == "base" else csmae_b200.mae_vit_large_patch16
    model = ctor(input_size=S, device=str(dev)).to(dev).train()
    step_model = model
    if world > 1:
        # same constructor call as main_pretrain.py:417-421; csmae_b200's wrapper lets the engine all-reduce the
        # flat gradient buffer in segments overlapped with the backward (--torch-ddp: torch's own wrapper)
        wrapper = torch.nn.parallel.DistributedDataParallel if args.torch_ddp else csmae_b200.DistributedDataParallel
        step_model = wrapper(model, device_ids=[env.local_rank], find_unused_parameters=True)
    decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.ndim == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.ndim == 1 or n.endswith(".bias"))]
    groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}]
    if args.torch_adamw:
        opt = torch.optim.AdamW(groups, lr=1.5e-4, betas=(0.9, 0.95), fused=True)
    else:
        # row f1: one kernel for the whole AdamW step, which also rewrites the bf16 shadow weights
        opt = csmae_b200.FusedAdamW(groups, lr=1.5e-4, betas=(0.9, 0.95), model=model)

    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    imgs1 = torch.randn(B, 3, S, S, device=dev, generator=g)
    imgs2 = torch.randn(B, 3, S, S, device=dev, generator=g)
    host1 = imgs1.cpu().pin_memory()
    host2 = imgs2.cpu().pin_memory()
    torch.manual_seed(1 + rank)      # masking noise: per-rank stream (main_pretrain.py:368-369)

    def step(x1, x2):
        # order of engine_pretrain.py:41-75: forward, backward + optimizer step, THEN zero_grad -- the host drops the
        # 258 gradient references while the GPU runs the optimizer kernel, not before the next step's first kernel
        loss, _, _ = step_model(x1, x2, MASK_RATIO)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(max(warmup, 3)):
        loss = step(imgs1, imgs2)
    assert torch.isfinite(loss).item(), "non-finite loss"

    sampler = ClockSampler(env.local_rank)
    if rank == 0 and full:
        sampler.start()
    launches0 = _native.launch_count
    ms_step = env.timed(lambda: step(imgs1, imgs2), steps)
    launches = _native.launch_count - launches0
    if getattr(args, "quick", False):
        sampler.stop()
        return {"arch": arch, "batch": B, "size": S, "ms_step": ms_step, "launches": launches, "clocks": sampler.summary()}

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
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(steps, lag)
        e1.record()
        env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1) / steps)
    ms_e2e_sync = e2e_timed(0)       # loss.item() right after every step, as engine_pretrain.py:55 does
    out = {"arch": arch, "batch": B, "size": S, "ms_step": ms_step, "launches": launches, "ms_e2e_sync": ms_e2e_sync,
           "h2d": int(host1.numel() * 4 * 2)}
    if not full:
        del prefetcher, opt, step_model, model
        torch.cuda.empty_cache()
        return out
    out["ms_e2e_lag"] = e2e_timed(1)            # same reads, one step late

    # fwd+bwd only (the BASELINE metric's second figure)
    def fwd_bwd():
        opt.zero_grad(set_to_none=True)
        l_, _, _ = step_model(imgs1, imgs2, MASK_RATIO)
        l_.backward()
    out["ms_fwd_bwd"] = env.timed(fwd_bwd, max(3, steps // 2))
    opt.zero_grad(set_to_none=True)

    # host-side cost of enqueueing one step (no synchronisation inside): if this approaches ms_per_step the
    # path is launch-bound on the CPU
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step(imgs1, imgs2)
    out["host_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    torch.cuda.synchronize()

    # the UNMODIFIED reference loop (engine_pretrain.py:18-101, verbatim copy in oracle/_ref) driving this module:
    # loader of pinned host batches -> .to(device) -> fp16-autocast context -> model(samples, mask_ratio) (single
    # input: scale 2 is the in-model crop kernel) -> loss.item() -> loss_scaler(...) -> zero_grad -> synchronize.
    # The loop is the reference's CALLER (not a checker): only its file is loaded from oracle/_ref.
    out["stock_engine"] = stock_engine_e2e(env, args, model, step_model, opt, host1, steps)
    sampler.stop()
    out["clocks"] = sampler.summary()

    # per-kernel breakdown of one step with CUDA events on the launching stream (outside the timed region):
    # graphs off and the wgrad side stream off, so every C-ABI call is timed back to back on ONE stream and the
    # families add up to a serial step (the graphed step overlaps wgrads with the dgrad chain and is shorter)
    model._engine.use_graphs = False
    model._engine.use_side_stream = False
    step(imgs1, imgs2)
    out["breakdown"] = kernel_breakdown(_native, lambda: step(imgs1, imgs2))     # every rank: the step all-reduces
    del prefetcher, opt, step_model, model
    torch.cuda.empty_cache()
    return out


def stock_engine_e2e(env, args, model, step_model, opt, host_batch, steps):
    try:
        from oracle import ref_loader
        if not ref_loader.reference_available():
            return {"unavailable": "oracle/_ref did not travel (tools/vendor_ref.py needs /root/reference)"}
        engine = ref_loader.reference_module("engine_pretrain")
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    import csmae_b200
    scaler = csmae_b200.NativeScalerWithGradNormCount()
    if not isinstance(opt, csmae_b200.FusedAdamW):
        return {"unavailable": "stock-engine loop is measured with FusedAdamW only"}
    ns = argparse.Namespace(accum_iter=1, mask_ratio=MASK_RATIO, lr=1.5e-4, min_lr=0.0, warmup_epochs=40, epochs=800,
                            local_rank=env.local_rank, wandb_project=None)

    class Loader:
        def __init__(self, k):
            self.k = k

        def __len__(self):
            return self.k

        def __iter__(self):
            return ((host_batch, None) for _ in range(self.k))

    def epoch(k):
        with contextlib.redirect_stdout(io.StringIO()):           # the loop prints its meters; stdout carries the JSON line
            return engine.train_one_epoch(step_model, Loader(k), opt, env.dev, 0, scaler, log_writer=None, args=ns)
    epoch(4)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = epoch(steps)
    e1.record()
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
    B = host_batch.shape[0]
    return {"value": B * env.world / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "h2d_bytes_per_step": int(host_batch.numel() * 4), "d2h_bytes_per_step": 4,
            "loop": "unmodified engine_pretrain.train_one_epoch (oracle/_ref), single-input forward with the in-model "
                    "crop kernel, csmae_b200.NativeScalerWithGradNormCount + FusedAdamW, loss.item() and "
                    "torch.cuda.synchronize() every step",
            "loss": stats.get("loss")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="csmae_b200", choices=["csmae_b200", "reference"])
    ap.add_argument("--arch", default="base", choices=["base", "large"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 64 base / 32 large)")
    ap.add_argument("--input-size", type=int, default=224)
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the ViT-L/224 and ViT-L/448 extra_configs")
    ap.add_argument("--torch-ddp", action="store_true", help="N>1: wrap with torch's DistributedDataParallel")
    ap.add_argument("--torch-adamw", action="store_true", help="torch.optim.AdamW(fused=True) instead of FusedAdamW")
    ap.add_argument("--quick", action="store_true", help="development: device-resident ms/step only, one short line")
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel CUDA-event breakdown")
    args = ap.parse_args()
    default_workload = args.arch == "base" and args.batch is None and args.input_size == 224
    if args.batch is None:
        args.batch = 64 if args.arch == "base" else 32
    if args.impl == "reference":
        run_reference(args)
        return

    env = Env(args)
    world, rank = env.world, env.rank
    B, S = args.batch, args.input_size
    m = measure(env, args, args.arch, B, S, args.steps, args.warmup, full=True)
    if args.quick:
        if rank == 0:
            m["images_per_s"] = B * world / (m["ms_step"] * 1e-3)
            m["env"] = {k: v for k, v in os.environ.items() if k.startswith("CSMAE_")}
            print(json.dumps(m), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    extras = []
    if default_workload and not args.no_extra:
        # BASELINE.json configs[2] (ViT-L/16 224, bs 32/GPU) and configs[4] (ViT-L/16 448, bs 16/GPU): short runs of
        # the same step, device-resident and host-fed
        for arch, b, s in (("large", 32, 224), ("large", 16, 448)):
            k = max(5, min(args.steps, 15))
            x = measure(env, args, arch, b, s, k, 3, full=False)
            fl = flops_per_image(arch, s)
            ips = b * world / (x["ms_step"] * 1e-3)
            extras.append({"config": workload_config(args, b, world, arch, s), "steps": k, "warmup": 3,
                           "ms_per_step": x["ms_step"], "value": ips, "unit": "images/s",
                           "e2e": {"value": b * world / (x["ms_e2e_sync"] * 1e-3), "unit": "images/s",
                                   "ms_per_step": x["ms_e2e_sync"], "h2d_bytes_per_step": x["h2d"],
                                   "d2h_bytes_per_step": 4},
                           "gpu_launches": x["launches"],
                           "step_frac_of_flop_roofline": ips / world * fl["fwd_bwd"] / (measured_peaks()["tflops_sustained"] * 1e12)})
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    breakdown = m["breakdown"]
    fl = flops_per_image(args.arch, S)
    peaks = measured_peaks()
    ips = B * world / (m["ms_step"] * 1e-3)
    gemm_ms = sum(v["ms"] for k, v in breakdown.items() if k.startswith("csm_linear"))
    gemm_n = sum(v["n"] for k, v in breakdown.items() if k.startswith("csm_linear"))
    step_ms_prof = sum(v["ms"] for v in breakdown.values())
    traffic, traffic_note = gemm_traffic(gemm_n) if default_workload else (None, "captured for the default workload only")
    gemm_tflops = 3 * fl["gemm_fwd"] * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05; csm_linear_fwd/dgrad/wgrad, all launches of one step)",
                "achieved": gemm_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": gemm_tflops / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_note,
                "launches_per_step": gemm_n, "avg_launch_us": gemm_ms * 1e3 / gemm_n if gemm_n else None,
                "flops_per_launch_avg": 3 * fl["gemm_fwd"] * B / gemm_n if gemm_n else None,
                "timing": "CUDA events around every launch of one un-graphed step on one stream (launches back to back, "
                          "caches as in the step)",
                "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
                "share_of_step": gemm_ms / step_ms_prof if step_ms_prof else None,
                "step_frac_of_flop_roofline": ips / world * fl["fwd_bwd"] / (peaks["tflops_sustained"] * 1e12)}
    line = {"metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B, world),
            "fwd_bwd_ms": m["ms_fwd_bwd"], "host_enqueue_ms_per_step": m["host_ms"], "gpu_launches": m["launches"],
            "e2e": {"value": B * world / (m["ms_e2e_sync"] * 1e-3), "unit": "images/s", "ms_per_step": m["ms_e2e_sync"],
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4,
                    "loss_read": "loss.item() right after every step (engine_pretrain.py:55)",
                    "lagged_read_ms_per_step": m["ms_e2e_lag"],
                    "lagged_read_value": B * world / (m["ms_e2e_lag"] * 1e-3),
                    "stock_engine": m["stock_engine"]},
            "clocks": m["clocks"], "roofline": roofline,
            "flops_per_image": fl, "kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(
                breakdown.items(), key=lambda kv: -kv[1]["ms"])},
            "kernel_ms_note": "graphs off, single stream: families add up to a serial step; the graphed step overlaps "
                              "the wgrad side stream with the dgrad chain"}
    if extras:
        line["extra_configs"] = extras
    if world == 1 and not args.no_cpu_baseline:
        k = 4
        sec, kind = cpu_reference_step_time(args.arch, S, args.cpu_batch, k, 1)
        line["cpu_baseline"] = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": torch.get_num_threads(),
                                "kind": kind, "sample": cpu_sample_text(kind, k, args.cpu_batch)}
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
