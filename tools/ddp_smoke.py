#!/usr/bin/env python
"""Tiny multi-GPU smoke of the DDP path (torchrun): a few steps of a small MAE_ViT_MsLdCeCd under
DistributedDataParallel exactly as main_pretrain.py:417-421 wraps it; prints per-step losses and checks that
the replicas stay identical.  python -m torch.distributed.run --nproc-per-node 2 tools/ddp_smoke.py [--big]"""
import faulthandler
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "cross-scale-mae_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(int(os.environ.get("SMOKE_TIMEOUT", "60")), exit=True)


def main():
    import csmae_b200
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    if "--big" in sys.argv:
        model = csmae_b200.mae_vit_base_patch16(input_size=224, device=str(dev)).to(dev).train()
        B, S = 16, 224
    else:
        cfg = dict(dim_model=128, encoder_num_layers=2, encoder_num_heads=2, decoder_embed_dim=64, decoder_num_layers=2,
                   decoder_num_heads=2, input_size=96, patch_size=16, predictor_hidden_size=128)
        model = csmae_b200.MAE_ViT_MsLdCeCd(**cfg, device=str(dev)).to(dev).train()
        B, S = 8, 96
    wrapper = csmae_b200.DistributedDataParallel if "--native" in sys.argv else torch.nn.parallel.DistributedDataParallel
    ddp = wrapper(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.95))
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    torch.manual_seed(1 + rank)
    for step in range(6):
        x1 = torch.randn(B, 3, S, S, device=dev, generator=g)
        x2 = torch.randn(B, 3, S, S, device=dev, generator=g)
        opt.zero_grad(set_to_none=True)
        loss, _, _ = ddp(x1, x2, 0.75)
        loss.backward()
        if "--check-grads" in sys.argv and step in (0, 3):
            bad = []
            for n, p_ in model.named_parameters():
                if p_.grad is None:
                    continue
                gs = [torch.empty_like(p_.grad) for _ in range(world)]
                dist.all_gather(gs, p_.grad.contiguous())
                d = (gs[0] - gs[-1]).abs().max().item()
                if d != 0.0:
                    bad.append((n, d, gs[0].abs().max().item()))
            if rank == 0:
                print(f"step {step}: {len(bad)} gradients differ across ranks", bad[:12], flush=True)
        opt.step()
        print(f"[rank {rank}] step {step} loss {loss.item():.5f} graphs={len(model._engine._graphs)}", flush=True)
    # replicas must hold identical weights after identical all-reduced updates
    w = model.decoder[0].attn.qkv.weight.detach().float()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert all(torch.equal(ws[0], x) for x in ws), "replicas diverged"
    if rank == 0:
        print("ddp smoke ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
