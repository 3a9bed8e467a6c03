#!/usr/bin/env python
"""What is each kernel family worth INSIDE the step?  For every family the training step is timed with that
family's C-ABI calls skipped (results are garbage -- timing only): the drop in ms/step is what a free version of
that family would buy, which is not its stand-alone duration when side-stream work overlaps it.

    python tools/whatif.py [--arch base] [--steps 20]

Development tool: it patches csmae_b200.engine.call in this process only.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "cross-scale-mae_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import csmae_b200  # noqa: E402
import csmae_b200.engine as E  # noqa: E402

FAMILIES = {
    "none": (),
    "colsum": ("csm_colsum_bf16",),
    "wgrad": ("csm_linear_wgrad",),
    "wgrad+colsum": ("csm_linear_wgrad", "csm_colsum_bf16"),
    "dgrad": ("csm_linear_dgrad",),
    "attention_bwd": ("csm_attention_bwd",),
    "attention_fwd": ("csm_attention_fwd",),
    "layernorm_bwd": ("csm_layernorm_bwd",),
    "layernorm_fwd": ("csm_layernorm_fwd",),
    "linear_fwd": ("csm_linear_fwd",),
    "losses": ("csm_recon_loss_fwd", "csm_recon_loss_bwd", "csm_cross_mse_fwd", "csm_cross_mse_bwd",
               "csm_bn_patch_fwd", "csm_bn_patch_bwd", "csm_ntxent_fwd", "csm_ntxent_bwd"),
}


def run(arch, steps, skip):
    real_call = E.call

    def patched(name, *a):
        if name in skip:
            return 0
        return real_call(name, *a)
    E.call = patched
    try:
        dev = torch.device("cuda", 0)
        torch.manual_seed(0)
        ctor = csmae_b200.mae_vit_base_patch16 if arch == "base" else csmae_b200.mae_vit_large_patch16
        bs = 64 if arch == "base" else 32
        model = ctor(input_size=224, device=str(dev)).to(dev).train()
        opt = csmae_b200.FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1.5e-4, betas=(0.9, 0.95),
                                    model=model)
        x1 = torch.randn(bs, 3, 224, 224, device=dev)
        x2 = torch.randn(bs, 3, 224, 224, device=dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss, _, _ = model(x1, x2, 0.75)
            loss.backward()
            opt.step()
        for _ in range(6):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    finally:
        E.call = real_call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="base")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    base = None
    out = []
    for name, skip in FAMILIES.items():
        ms = run(args.arch, args.steps, set(skip))
        if base is None:
            base = ms
        rec = {"skipped": name, "ms_per_step": round(ms, 3), "saved_ms": round(base - ms, 3)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    again = run(args.arch, args.steps, set())
    print(json.dumps({"skipped": "none (again)", "ms_per_step": round(again, 3)}), flush=True)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
