"""Generate the committed golden fixtures from the REAL reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/models_mae/*.py verbatim through oracle/ref_loader.py
(timm Block/PatchEmbed restated in oracle/timm_shim.py) and records, for small
seeded configurations, everything a parity test needs: weights, inputs, the
masking noise the reference itself drew, outputs, loss terms and parameter
gradients.  The reference has no golden vectors of its own, so
these files are the pin for oracle/restatement.py.

Fixtures written next to this script:
  tiny_cecd.npz      MAE_ViT_MsLdCeCd, paired oracle (crop replaced by a fixed scale-2 batch)
  tiny_baseline.npz  MAE_ViT_Baseline with mask_seed
  misc.npz           sincos pos-embeds, NT-Xent known answers, lr schedule samples
  anchors.json       cfg-1 loss value (ViT-B/16 Baseline, 1x224x224, CPU fp32)
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader as rl  # noqa: E402

TINY = dict(dim_model=64, encoder_num_layers=2, encoder_num_heads=1,
            decoder_embed_dim=64, decoder_num_layers=2, decoder_num_heads=2,
            input_size=64, patch_size=16, predictor_hidden_size=128)


@contextlib.contextmanager
def record_rand(store):
    real = torch.rand

    def wrapped(*a, **k):
        out = real(*a, **k)
        store.append(out.detach().clone())
        return out
    torch.rand = wrapped
    try:
        yield
    finally:
        torch.rand = real


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def perturb(model, seed):
    """Make zero-initialised biases / unit LayerNorm weights non-trivial so every parameter
    influences the output (the reference's init leaves all biases at 0)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if not p.requires_grad:
                continue
            if n.endswith(".bias") or "norm" in n or n.startswith("predictor.1"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def main():
    Base, MsLd, CeCd = rl.reference_classes()
    out = {}

    # ---- tiny MsLdCeCd, paired ------------------------------------------------------------
    torch.manual_seed(7)
    model = quiet(CeCd, **TINY, device="cpu")
    perturb(model, 11)
    model.train()
    g = torch.Generator().manual_seed(3)
    imgs1 = torch.randn(4, 3, 64, 64, generator=g)
    imgs2 = torch.randn(4, 3, 64, 64, generator=g)
    model.crop = rl.FixedScale2(imgs2)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    noises = []
    torch.manual_seed(123)
    with record_rand(noises):
        loss, pred, mask, (e1, e2), (d1, d2) = model(imgs1, mask_ratio=0.75, return_embeds=True)
    assert len(noises) == 2
    loss.backward()
    grads = {n: (p.grad if p.grad is not None else torch.zeros(0)) for n, p in model.named_parameters()}
    fx = {"imgs1": imgs1, "imgs2": imgs2, "noise1": noises[0], "noise2": noises[1],
          "loss": loss, "pred": pred, "mask": mask, "enc1": e1, "enc2": e2, "dec1": d1, "dec2": d2,
          "bn_running_mean": model.predictor[1].running_mean,
          "bn_running_var": model.predictor[1].running_var}
    fx = to_np(fx)
    fx.update({"sd/" + k: v for k, v in to_np(sd0).items()})
    fx.update({"grad/" + k: v for k, v in to_np(grads).items()})
    np.savez_compressed(os.path.join(HERE, "tiny_cecd.npz"), **fx)
    out["tiny_cecd_loss"] = float(loss)
    out["tiny_cecd_unused_grads"] = sorted(n for n, p in model.named_parameters()
                                           if p.requires_grad and p.grad is None)

    # ---- tiny Baseline with mask_seed -----------------------------------------------------
    torch.manual_seed(5)
    kw = {k: v for k, v in TINY.items() if k != "predictor_hidden_size"}
    bmodel = quiet(Base, **kw, device="cpu")
    perturb(bmodel, 13)
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(17))
    sdb = {k: v.clone() for k, v in bmodel.state_dict().items()}
    noises = []
    with record_rand(noises):
        loss, pred, mask, enc, dec = bmodel(x, mask_ratio=0.75, mask_seed=99, return_embeds=True)
    loss.backward()
    grads = {n: (p.grad if p.grad is not None else torch.zeros(0)) for n, p in bmodel.named_parameters()}
    fx = to_np({"imgs": x, "noise": noises[0], "loss": loss, "pred": pred, "mask": mask,
                "enc": enc, "dec": dec})
    fx.update({"sd/" + k: v for k, v in to_np(sdb).items()})
    fx.update({"grad/" + k: v for k, v in to_np(grads).items()})
    np.savez_compressed(os.path.join(HERE, "tiny_baseline.npz"), **fx)
    out["tiny_baseline_loss"] = float(loss)

    # ---- misc: pos-embed, NT-Xent, lr schedule -------------------------------------------
    pe = rl.reference_module("util.pos_embed")
    cl = rl.reference_module("util.contrast_loss")
    ls = rl.reference_module("util.lr_sched")
    misc = {"pos_768_14": pe.get_2d_sincos_pos_embed(768, 14, cls_token=True),
            "pos_512_14": pe.get_2d_sincos_pos_embed(512, 14, cls_token=True),
            "pos_64_4": pe.get_2d_sincos_pos_embed(64, 4, cls_token=True)}
    g = torch.Generator().manual_seed(21)
    for bs, dim in ((4, 64), (64, 768)):
        f1 = torch.randn(bs, dim, generator=g)
        f2 = f1 + 0.5 * torch.randn(bs, dim, generator=g)
        val = cl.NTXentLoss(bs, 0.5, cos_sim=True, device=None)(f1, f2)
        misc[f"ntx_f1_{bs}"] = f1.numpy()
        misc[f"ntx_f2_{bs}"] = f2.numpy()
        misc[f"ntx_loss_{bs}"] = val.numpy()

    class A:
        lr, min_lr, warmup_epochs, epochs = 1e-3, 1e-5, 40, 400

    class Opt:
        param_groups = [{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}]
    eps = np.array([0.0, 0.5, 39.99, 40.0, 123.456, 399.0])
    misc["lr_epochs"] = eps
    misc["lr_values"] = np.array([ls.adjust_learning_rate(Opt, float(e), A) for e in eps])
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **misc)

    # ---- cfg-1 anchor (BASELINE.json configs[0]) -----------------------------------------
    torch.manual_seed(0)
    big = quiet(Base, **rl.ARGS_VIT_BASE, input_size=224, patch_size=16)
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        loss, pred, mask = big(x, mask_ratio=0.75, mask_seed=1234)
    out["cfg1_vitb_baseline_loss"] = float(loss)
    out["cfg1_mask_sum"] = float(mask.sum())
    out["cfg1_pred_abs_mean"] = float(pred.abs().mean())
    out["torch_version"] = torch.__version__
    out["tiny_config"] = TINY
    with open(os.path.join(HERE, "anchors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
