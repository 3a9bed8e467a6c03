"""Functional PyTorch restatement of the Cross-Scale MAE pretraining hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This file travels to the
GPU box (the reference does not) and is the checker the `-m gpu` parity tests,
`smoke()` and `bench.py`'s cpu_baseline leg compare against.  It is pinned
against the real reference by tests/test_oracle_golden.py (fixtures produced
by tests/golden/make_golden.py from the verbatim reference classes).

It is written as pure functions over a reference-layout `state_dict`
(parameter names of SURVEY.md section 8b) rather than as an nn.Module tree,
and takes the masking noise explicitly so the CUDA path and the oracle can be
fed identical randomness.  It uses the same torch ops as the reference at the
same places (conv2d, linear, layer_norm, einsum, batch_norm, ...), so running
it under `torch.autocast("cuda", dtype=torch.bfloat16)` reproduces the
reference's rounding points (SURVEY.md section 8a').

Each function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6        # models_mae/MAE_ViT_Baseline.py:43-45 (partial(nn.LayerNorm, eps=1e-6))
BN_EPS = 1e-5        # nn.BatchNorm1d default, models_mae/MLP.py:7
BN_MOMENTUM = 0.1
NTXENT_TAU = 0.5     # models_mae/MAE_ViT_MsLdCeCd.py:62
NTXENT_EPS = 1e-8    # util/contrast_loss.py:51


# ----------------------------------------------------------------------------
# init-time pieces
# ----------------------------------------------------------------------------
def sincos_pos_embed_2d(embed_dim, grid_size, cls_token=True):
    """util/pos_embed.py:16-63.  Row h*G+w = [sin(w*om) | cos(w*om) | sin(h*om) | cos(h*om)],
    om_k = 10000^(-k/(D/4)), float64; row 0 zeros for the cls token."""
    assert embed_dim % 4 == 0
    quarter = embed_dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    hh, ww = np.meshgrid(np.arange(grid_size, dtype=np.float64),
                         np.arange(grid_size, dtype=np.float64), indexing="ij")
    aw = ww.reshape(-1)[:, None] * omega[None, :]
    ah = hh.reshape(-1)[:, None] * omega[None, :]
    emb = np.concatenate([np.sin(aw), np.cos(aw), np.sin(ah), np.cos(ah)], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


def lr_at(epoch_frac, lr, min_lr, warmup_epochs, epochs):
    """util/lr_sched.py:9-22 -- linear warm-up then half-cosine."""
    if epoch_frac < warmup_epochs:
        return lr * epoch_frac / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (
        1.0 + math.cos(math.pi * (epoch_frac - warmup_epochs) / (epochs - warmup_epochs)))


# ----------------------------------------------------------------------------
# masking (models_mae/MAE_ViT_Shared.py:57-84)
# ----------------------------------------------------------------------------
def masking_from_noise(noise, mask_ratio):
    """noise [N, L] -> (ids_keep [N, keep] i64, mask [N, L] f32, ids_restore [N, L] i64).
    Ties are broken by index (stable), which is what the CUDA kernel defines."""
    N, L = noise.shape
    len_keep = int(L * (1 - mask_ratio))
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :len_keep]
    mask = torch.ones([N, L], device=noise.device)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, dim=1, index=ids_restore)
    return ids_keep, mask, ids_restore


# ----------------------------------------------------------------------------
# transformer block (timm 0.4.12 Block; see oracle/timm_shim.py)
# ----------------------------------------------------------------------------
def _block(x, sd, prefix, num_heads):
    B, S, C = x.shape
    d = C // num_heads
    h = F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], LN_EPS)
    qkv = F.linear(h, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"])
    qkv = qkv.reshape(B, S, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, C)
    x = x + F.linear(o, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], LN_EPS)
    h = F.linear(h, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"])
    h = F.gelu(h)
    x = x + F.linear(h, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])
    return x


def _count_layers(sd, stem):
    n = 0
    while f"{stem}.{n}.norm1.weight" in sd:
        n += 1
    return n


def patchify(imgs, p, c):
    """models_mae/MAE_ViT_Shared.py:24-39: target[n, h*w, (p, q, c)]."""
    n, _, hh, ww = imgs.shape
    assert hh == ww and hh % p == 0
    g = hh // p
    x = imgs.reshape(n, c, g, p, g, p)
    x = torch.einsum("nchpwq->nhwpqc", x)
    return x.reshape(n, g * g, p * p * c)


def baseline_pass(sd, imgs, noise, mask_ratio, enc_heads, dec_heads, norm_pix_loss=False):
    """One scale: models_mae/MAE_ViT_Baseline.py:243-320 + MAE_ViT_Shared.py:97-120,269-290.
    Returns dict(loss, pred, mask, ids_restore, enc_emb, dec_emb)."""
    w = sd["patch_embed.proj.weight"]
    p = w.shape[-1]
    c = w.shape[1]
    # forward_encoder (Baseline.py:243-266)
    x = F.conv2d(imgs, w, sd["patch_embed.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    x = x + sd["encoder_pos_embed"][:, 1:, :]
    ids_keep, mask, ids_restore = masking_from_noise(noise, mask_ratio)
    D = x.shape[-1]
    x = torch.gather(x, dim=1, index=ids_keep.unsqueeze(-1).repeat(1, 1, D))
    cls = (sd["cls_token"] + sd["encoder_pos_embed"][:, :1, :]).expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    for i in range(_count_layers(sd, "encoder")):
        x = _block(x, sd, f"encoder.{i}.", enc_heads)
    enc_emb = x  # encoder_norm(x) is computed and DISCARDED in the reference (Baseline.py:264)

    # forward_decoder (Baseline.py:268-297)
    y = F.linear(enc_emb, sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    L = ids_restore.shape[1]
    mask_tokens = sd["mask_token"].repeat(y.shape[0], L + 1 - y.shape[1], 1)
    y_ = torch.cat([y[:, 1:, :], mask_tokens], dim=1)
    y_ = torch.gather(y_, dim=1, index=ids_restore.unsqueeze(-1).repeat(1, 1, y.shape[2]))
    y = torch.cat([y[:, :1, :], y_], dim=1)
    y = y + sd["decoder_pos_embed"]
    for i in range(_count_layers(sd, "decoder")):
        y = _block(y, sd, f"decoder.{i}.", dec_heads)
    Dd = y.shape[-1]
    dec_emb = F.layer_norm(y, (Dd,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], LN_EPS)
    pred = F.linear(dec_emb, sd["decoder_pred.weight"], sd["decoder_pred.bias"])[:, 1:, :]

    # forward_loss (Shared.py:269-290 -> 97-111 -> 113-120)
    target = patchify(imgs, p, c)
    if norm_pix_loss:
        mean = target.mean(dim=-1, keepdim=True)
        var = target.var(dim=-1, keepdim=True)
        target = (target - mean) / (var + 1.0e-6) ** 0.5
    per_patch = ((pred - target) ** 2).mean(dim=-1)
    loss = (per_patch * mask).sum() / mask.sum()
    return dict(loss=loss, pred=pred, mask=mask, ids_restore=ids_restore,
                enc_emb=enc_emb, dec_emb=dec_emb)


def predictor_mlp(sd, x, running=None):
    """models_mae/MLP.py:4-10 in train mode: Linear -> BatchNorm1d(L) over [N, L, H] (channel =
    patch index) -> ReLU -> Linear.  `running` = optional (mean, var) tensors updated in place."""
    h = F.linear(x, sd["predictor.0.weight"], sd["predictor.0.bias"])
    rm, rv = (running if running is not None else (None, None))
    h = F.batch_norm(h, rm, rv, sd["predictor.1.weight"], sd["predictor.1.bias"],
                     training=True, momentum=BN_MOMENTUM, eps=BN_EPS)
    h = F.relu(h)
    return F.linear(h, sd["predictor.3.weight"], sd["predictor.3.bias"])


def ntxent(f1, f2, tau=NTXENT_TAU, eps=NTXENT_EPS):
    """util/contrast_loss.py:17-41,71-101 with cos_sim=True: positives are (i, i+-B); the
    denominator holds only the 2B-2 negatives (self and positive excluded) plus eps."""
    zi, zj = F.normalize(f1, dim=1), F.normalize(f2, dim=1)
    B = zi.shape[0]
    z = torch.cat([zi, zj], dim=0)
    sim = torch.exp(F.cosine_similarity(z.unsqueeze(1), z.unsqueeze(0), dim=-1) / tau)
    idx = torch.arange(2 * B, device=z.device)
    partner = (idx + B) % (2 * B)
    pos = sim[idx, partner]
    neg_mask = torch.ones(2 * B, 2 * B, dtype=torch.bool, device=z.device)
    neg_mask[idx, idx] = False
    neg_mask[idx, partner] = False
    neg = (sim * neg_mask).sum(dim=-1)
    return (-torch.log(pos / (neg + eps))).mean()


def cross_scale_forward(sd, imgs1, imgs2, noise1, noise2, mask_ratio, enc_heads, dec_heads,
                        reduction="sum", running=None, norm_pix_loss=False):
    """MAE_ViT_MsLd.py:37-77 + MAE_ViT_MsLdCeCd.py:27-84 with scale-2 supplied (paired form,
    SURVEY.md section 0.6).  Returns dict with the total loss, its four terms and the tensors the
    reference returns."""
    a = baseline_pass(sd, imgs1, noise1, mask_ratio, enc_heads, dec_heads, norm_pix_loss)
    b = baseline_pass(sd, imgs2, noise2, mask_ratio, enc_heads, dec_heads, norm_pix_loss)
    loss_d = a["loss"] + b["loss"]
    if reduction == "mean":
        loss_d = loss_d / 2
    cross_pred = predictor_mlp(sd, b["dec_emb"][:, 1:, :], running)
    cross_target = a["dec_emb"][:, 1:, :]          # NOT detached (MsLdCeCd.py:57-59)
    loss_cd = ((cross_pred - cross_target) ** 2).mean(dim=-1).mean()   # Shared.py:113-119, mask=None
    f1 = torch.flatten(a["enc_emb"][:, 1:, :].mean(dim=1), 1)
    f2 = torch.flatten(b["enc_emb"][:, 1:, :].mean(dim=1), 1)
    loss_ce = ntxent(f1, f2)
    return dict(loss=loss_d + loss_cd + loss_ce, loss_orig=a["loss"], loss_crop=b["loss"],
                loss_cd=loss_cd, loss_ce=loss_ce, pred=a["pred"], mask=a["mask"],
                ids_restore=(a["ids_restore"], b["ids_restore"]), mask_crop=b["mask"],
                enc_emb=(a["enc_emb"], b["enc_emb"]), dec_emb=(a["dec_emb"], b["dec_emb"]),
                pred_crop=b["pred"])


# ----------------------------------------------------------------------------
# state-dict construction with the reference's init (Baseline.py:201-241, MsLdCeCd.py:23-25)
# ----------------------------------------------------------------------------
def make_state_dict(dim_model, encoder_num_layers, encoder_num_heads, decoder_embed_dim,
                    decoder_num_layers, decoder_num_heads, input_size, patch_size=16,
                    input_channels=3, predictor_hidden_size=2048, with_predictor=True,
                    seed=0, device="cpu", simple_init=False):
    """Random-init state dict in the reference's key layout.  With simple_init=False the
    distributions match the reference's init (xavier-uniform Linears, N(0, .02) tokens, sincos
    pos-embeds) but the RNG draw ORDER does not -- use the product module (which mirrors the
    construction order) or a golden fixture when identical weights are needed."""
    g = torch.Generator().manual_seed(seed)
    D, Dd, p, c = dim_model, decoder_embed_dim, patch_size, input_channels
    L = (input_size // patch_size) ** 2

    def xavier(out_f, in_f, shape=None):
        bound = math.sqrt(6.0 / (in_f + out_f))
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        return w.reshape(shape) if shape else w

    def kaiming_linear(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound,
                (torch.rand(out_f, generator=g) * 2 - 1) * bound)

    sd = {}
    sd["cls_token"] = torch.randn(1, 1, D, generator=g) * 0.02
    sd["encoder_pos_embed"] = torch.from_numpy(
        sincos_pos_embed_2d(D, int(L ** 0.5))).float().unsqueeze(0)
    sd["mask_token"] = torch.randn(1, 1, Dd, generator=g) * 0.02
    sd["decoder_pos_embed"] = torch.from_numpy(
        sincos_pos_embed_2d(Dd, int(L ** 0.5))).float().unsqueeze(0)
    sd["patch_embed.proj.weight"] = xavier(D, c * p * p, (D, c, p, p))
    sd["patch_embed.proj.bias"] = torch.zeros(D)
    sd["decoder_embed.weight"] = xavier(Dd, D)
    sd["decoder_embed.bias"] = torch.zeros(Dd)

    def block(prefix, dim):
        sd[prefix + "norm1.weight"] = torch.ones(dim)
        sd[prefix + "norm1.bias"] = torch.zeros(dim)
        sd[prefix + "attn.qkv.weight"] = xavier(3 * dim, dim)
        sd[prefix + "attn.qkv.bias"] = torch.zeros(3 * dim)
        sd[prefix + "attn.proj.weight"] = xavier(dim, dim)
        sd[prefix + "attn.proj.bias"] = torch.zeros(dim)
        sd[prefix + "norm2.weight"] = torch.ones(dim)
        sd[prefix + "norm2.bias"] = torch.zeros(dim)
        sd[prefix + "mlp.fc1.weight"] = xavier(4 * dim, dim)
        sd[prefix + "mlp.fc1.bias"] = torch.zeros(4 * dim)
        sd[prefix + "mlp.fc2.weight"] = xavier(dim, 4 * dim)
        sd[prefix + "mlp.fc2.bias"] = torch.zeros(dim)

    for i in range(encoder_num_layers):
        block(f"encoder.{i}.", D)
    for i in range(decoder_num_layers):
        block(f"decoder.{i}.", Dd)
    sd["decoder_pred.weight"] = xavier(p * p * c, Dd)
    sd["decoder_pred.bias"] = torch.zeros(p * p * c)
    sd["decoder_norm.weight"] = torch.ones(Dd)
    sd["decoder_norm.bias"] = torch.zeros(Dd)
    sd["encoder_norm.weight"] = torch.ones(D)
    sd["encoder_norm.bias"] = torch.zeros(D)
    if with_predictor:
        w0, b0 = kaiming_linear(predictor_hidden_size, Dd)
        sd["predictor.0.weight"], sd["predictor.0.bias"] = w0, b0
        sd["predictor.1.weight"] = torch.ones(L)
        sd["predictor.1.bias"] = torch.zeros(L)
        w3, b3 = kaiming_linear(Dd, predictor_hidden_size)
        sd["predictor.3.weight"], sd["predictor.3.bias"] = w3, b3
    if simple_init:
        # perturb biases / norm affine so parity tests exercise every parameter
        for k in sd:
            if k.endswith(".bias") or "norm" in k or k.startswith("predictor.1"):
                sd[k] = sd[k] + 0.05 * torch.randn(sd[k].shape, generator=g)
    return {k: v.to(device) for k, v in sd.items()}


FROZEN_KEYS = ("encoder_pos_embed", "decoder_pos_embed")


def loss_and_grads(sd, imgs1, imgs2, noise1, noise2, mask_ratio, enc_heads, dec_heads,
                   autocast_dtype=None, paired=True, loss_scale=1.0, **forward_kwargs):
    """Runs the restatement with autograd; returns (outputs dict, grads dict keyed like sd)."""
    leaves = {k: v.detach().clone().requires_grad_(k not in FROZEN_KEYS) for k, v in sd.items()}
    dev = imgs1.device.type
    ctx = (torch.autocast(dev, dtype=autocast_dtype) if autocast_dtype is not None
           else torch.autocast(dev, enabled=False))
    with ctx:
        if paired:
            out = cross_scale_forward(leaves, imgs1, imgs2, noise1, noise2, mask_ratio,
                                      enc_heads, dec_heads, **forward_kwargs)
        else:
            out = baseline_pass(leaves, imgs1, noise1, mask_ratio, enc_heads, dec_heads)
    (out["loss"] * loss_scale).backward()
    grads = {k: (v.grad if v.grad is not None else None) for k, v in leaves.items()}
    return out, grads
