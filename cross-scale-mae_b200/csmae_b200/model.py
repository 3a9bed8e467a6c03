"""nn.Module surface of the reference's MAE model family, backed by the B200 hot-path engine.

The module tree, parameter names/shapes, constructor kwargs, init RNG order and forward signatures
mirror the reference so that it drops into main_pretrain.py / engine_pretrain.py unchanged and
checkpoints are interchangeable:
  models_mae/MAE_ViT_Shared.py    (patchify / unpatchify / random_masking / loss selection)
  models_mae/MAE_ViT_Baseline.py  (encoder / decoder, initialize_weights)
  models_mae/MAE_ViT_MsLd.py      (two scales, in-model RandomResizedCrop)
  models_mae/MAE_ViT_MsLdCeCd.py  (+ predictor cross-decoder loss + NT-Xent)
The sub-modules here (PatchEmbed, Block, ...) are parameter containers with timm 0.4.12's names;
all arithmetic runs in csmae_b200.engine through the C-ABI kernels.  There is no PyTorch fallback:
calling forward on a non-sm_100 device raises.
"""
from functools import partial

import torch
import torch.nn as nn

from .engine import CrossScaleStep, HotPathEngine
from .pos_embed import get_2d_sincos_pos_embed, sincos_pos_embed_  # noqa: F401


# ------------------------------------------------------------------------------------------------
# parameter containers with timm 0.4.12 attribute names (checkpoint keys: norm1, attn.qkv, attn.proj,
# norm2, mlp.fc1, mlp.fc2, patch_embed.proj -- cf. main_finetune.py:557-566 of the reference)
# ------------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class BatchRandomResizedCrop(nn.Module):
    """`self.crop` of the two-scale models: one random crop box per BATCH, resized back to the input size with
    bilinear + antialias filtering (reference: nn.Sequential(T.RandomResizedCrop(size, scale=ms_range,
    antialias=True)), MAE_ViT_MsLd.py:29-35).  The box is drawn by torchvision's own get_params (same CPU-generator
    draws); the resize is one csm_resized_crop kernel instead of a slice copy + ATen's
    _upsample_bilinear2d_aa."""

    def __init__(self, size, scale, ratio=(3.0 / 4.0, 4.0 / 3.0)):
        super().__init__()
        self.size = int(size)
        self.scale = tuple(scale)
        self.ratio = tuple(ratio)

    def forward(self, imgs):
        from torchvision.transforms import RandomResizedCrop
        from ._native import call
        assert imgs.dim() == 4, "expected a [N, C, H, W] batch"
        top, left, h, w = RandomResizedCrop.get_params(imgs, list(self.scale), list(self.ratio))
        x = imgs.contiguous().float()
        n, c, hh, ww = x.shape
        out = torch.empty(n, c, self.size, self.size, dtype=torch.float32, device=x.device)
        call("csm_resized_crop", x, out, n * c, hh, ww, top, left, h, w, self.size)
        return out

    def extra_repr(self):
        return f"size={self.size}, scale={self.scale}, ratio={self.ratio}, interpolation=bilinear, antialias=True"


class _FixedScale2(nn.Module):
    """Stands in for `self.crop` when the caller supplies the scale-2 batch (paired form)."""

    def __init__(self, imgs2):
        super().__init__()
        self.imgs2 = imgs2

    def forward(self, imgs):
        return self.imgs2


# ------------------------------------------------------------------------------------------------
class MAE_ViT_Baseline(nn.Module):
    """Single-scale MAE (reference: models_mae/MAE_ViT_Baseline.py:14-320)."""

    _scales = 1
    _use_cd = False
    _use_ce = False

    def __init__(self, input_size=128, input_channels=3, patch_size=16, mask_ratio=0.75, dim_model=1024,
                 encoder_num_layers=24, encoder_num_heads=16, decoder_embed_dim=512, decoder_num_layers=8,
                 decoder_num_heads=16, residual_norm_style="post", residual_dropout=0.0, ffn_name="MLP",
                 ffn_activation="gelu", ffn_ratio=4, ffn_dropout=0.0, attn_name="scaled_dot_product",
                 attn_dropout=0.0, norm_layer=None, use_xformers=False, device=None,
                 norm_pix_loss=False, loss="mse", **kwargs):
        # every other key of vars(args) is silently ignored, as upstream (MAE_ViT_Shared.py:9-15)
        super().__init__()
        self.loss = loss.lower()
        self.norm_pix_loss = norm_pix_loss
        if self.loss != "mse":
            raise NotImplementedError(f"loss={loss!r}: only the default 'mse' reconstruction loss is on the hot path")
        self.input_size = input_size
        self.input_channels = input_channels
        self.patch_size = int(patch_size)            # --patch_size arrives as str (main_pretrain.py:79-86)
        self.dim_model = dim_model
        self.decoder_embed_dim = decoder_embed_dim
        self.mask_ratio = mask_ratio
        self.use_xformers = use_xformers
        self.device = device
        self.encoder_num_heads = encoder_num_heads
        self.decoder_num_heads = decoder_num_heads
        assert input_size % self.patch_size == 0
        if use_xformers:
            raise NotImplementedError("use_xformers=True is a different (post-norm) architecture: not on the hot path")
        assert attn_name == "scaled_dot_product", f"Attention {attn_name} not supported"
        assert ffn_name == "MLP", f"Feedforward {ffn_name} not supported"
        assert ffn_activation == "gelu", f"Feedforward activation {ffn_activation} not supported"
        assert residual_dropout == 0.0 and ffn_dropout == 0.0 and attn_dropout == 0.0, "dropout is 0 on the hot path"
        assert dim_model % encoder_num_heads == 0 and decoder_embed_dim % decoder_num_heads == 0
        if norm_layer is None:
            norm_layer = partial(nn.LayerNorm, eps=1e-6)

        self.patch_embed = PatchEmbed(input_size, self.patch_size, input_channels, dim_model)
        self.num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim_model))
        self.encoder_pos_embed = nn.Parameter(torch.zeros(1, self.num_patches + 1, dim_model), requires_grad=False)
        self.decoder_embed = nn.Linear(dim_model, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, self.num_patches + 1, decoder_embed_dim),
                                              requires_grad=False)
        self.encoder = nn.ModuleList(
            [Block(dim_model, encoder_num_heads, ffn_ratio, norm_layer) for _ in range(encoder_num_layers)])
        self.decoder = nn.ModuleList(
            [Block(decoder_embed_dim, decoder_num_heads, ffn_ratio, norm_layer) for _ in range(decoder_num_layers)])
        self.decoder_pred = nn.Linear(decoder_embed_dim, self.patch_size ** 2 * input_channels, bias=True)
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.encoder_norm = norm_layer(dim_model)     # registered, never used (MAE_ViT_Baseline.py:264)
        self.initialize_weights()
        self._engine_obj = None

    # ---- init (MAE_ViT_Baseline.py:201-241) -------------------------------------------------------
    def initialize_weights(self):
        grid = int(self.patch_embed.num_patches ** 0.5)
        sincos_pos_embed_(self.encoder_pos_embed.data, grid, cls_token=True)
        sincos_pos_embed_(self.decoder_pos_embed.data, grid, cls_token=True)
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.normal_(self.cls_token, std=0.02)
        torch.nn.init.normal_(self.mask_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- helpers kept from MAE_ViT_Shared.py ------------------------------------------------------
    def patchify(self, imgs, p, c):
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        h = w = imgs.shape[2] // p
        x = imgs.reshape(shape=(imgs.shape[0], c, h, p, w, p))
        x = torch.einsum("nchpwq->nhwpqc", x)
        return x.reshape(shape=(imgs.shape[0], h * w, p ** 2 * c))

    def unpatchify(self, x, p, c):
        h = w = int(x.shape[1] ** 0.5)
        assert h * w == x.shape[1]
        x = x.reshape(shape=(x.shape[0], h, w, p, p, c))
        x = torch.einsum("nhwpqc->nchpwq", x)
        return x.reshape(shape=(x.shape[0], c, h * p, h * p))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {}

    # ---- engine plumbing ------------------------------------------------------------------------
    @property
    def _engine(self):
        if self._engine_obj is None:
            object.__setattr__(self, "_engine_obj", HotPathEngine(self, use_cd=self._use_cd, use_ce=self._use_ce))
        return self._engine_obj

    def _draw_noise(self, n, device, mask_seed):
        # same generator stream as MAE_ViT_Shared.py:66 (torch.rand(N, L, device=x.device)); with a
        # mask_seed the global generator is re-seeded before every pass (MAE_ViT_Baseline.py:301-302)
        if mask_seed is not None:
            torch.manual_seed(mask_seed)
        return torch.rand(n, self.num_patches, device=device)

    def _run(self, imgs_list, mask_ratio, mask_seed, noises=None):
        eng = self._engine
        if noises is None:
            noises = [self._draw_noise(im.shape[0], im.device, mask_seed) for im in imgs_list]
        plist = eng.snapshot()["plist"]
        # kernels first, autograd bookkeeping second (it then overlaps the GPU)
        out = eng.forward(imgs_list, noises, mask_ratio, self.training)
        if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
            loss = CrossScaleStep.apply(eng, out["loss"], eng.generation, *plist)
        else:
            loss = out["loss"].clone()
        return loss, out

    def forward(self, imgs, mask_ratio=0.75, mask_seed=None, return_embeds=False, noise=None):
        """-> (loss, pred [N, L, p*p*c], mask [N, L]) (+ enc_emb, dec_emb); MAE_ViT_Baseline.py:299-320."""
        loss, out = self._run([imgs], mask_ratio, mask_seed, None if noise is None else [noise])
        pred, mask = out["pred"][0].clone(), out["mask"][0].clone()
        if not return_embeds:
            return loss, pred, mask
        return loss, pred, mask, out["enc_emb"][0].clone(), out["dec_emb"][0].clone()


class MAE_ViT_MsLd(MAE_ViT_Baseline):
    """Two scales, reconstruction losses summed (reference: models_mae/MAE_ViT_MsLd.py:8-77)."""

    _scales = 2

    def __init__(self, ms_range=(0.25, 0.75), ms_decoder_loss_reduction="sum", **kwargs):
        super().__init__(**kwargs)
        self.ms_decoder_loss_reduction = ms_decoder_loss_reduction.lower()
        self.allowed_reductions = ["mean", "sum"]
        assert self.ms_decoder_loss_reduction in self.allowed_reductions, \
            f"ms_decoder_loss_reduction must be one of: {self.allowed_reductions}"
        # one crop box per batch, bilinear + antialias, box drawn from the CPU generator exactly as upstream
        # (MAE_ViT_MsLd.py:29-35); not used by the paired form forward(imgs1, imgs2, ...)
        self.crop = nn.Sequential(BatchRandomResizedCrop(self.input_size, ms_range))

    def _forward_two_scale(self, imgs, imgs2, mask_ratio, mask_seed, consistent_mask, noise):
        if mask_seed is not None:
            torch.manual_seed(mask_seed)
        elif consistent_mask:
            mask_seed = torch.randint(0, 2 ** 32 - 1, (1,)).item()
        imgs_crop = self.crop(imgs) if imgs2 is None else imgs2
        assert imgs_crop.shape == imgs.shape, "both scales must have the model's input size"
        return self._run([imgs, imgs_crop], mask_ratio, mask_seed, noise)

    # positional order after the image tensor(s), as in the reference class of the same name
    # (MAE_ViT_MsLd.py:37-44 / MAE_ViT_MsLdCd.py:26-33; MAE_ViT_MsLdCeCd.py:27-35 inserts contr_bs)
    _positional = ("mask_ratio", "mask_seed", "return_embeds", "consistent_mask")

    def forward(self, imgs, *args, noise=None, **kwargs):
        """forward(imgs, mask_ratio=0.75, ...)            single input, scale 2 = in-model random crop
        forward(imgs_scale1, imgs_scale2, mask_ratio)  paired form (BASELINE.json north_star)
        -> (loss, pred_orig, mask_orig) or the 5-tuple with ((enc1, enc2), (dec1, dec2))."""
        imgs2 = None
        args = list(args)
        if args and isinstance(args[0], torch.Tensor) and args[0].dim() == 4:
            imgs2 = args.pop(0)
        if len(args) > len(self._positional):
            raise TypeError(f"forward() takes at most {len(self._positional)} positional arguments after the images")
        opts = dict(mask_ratio=0.75, contr_bs=None, mask_seed=None, return_embeds=False, consistent_mask=False)
        for name, value in zip(self._positional, args):
            if name in kwargs:
                raise TypeError(f"forward() got multiple values for argument {name!r}")
            opts[name] = value
        for name in list(kwargs):
            if name in opts:
                opts[name] = kwargs.pop(name)          # anything else is swallowed, as the reference's **kwargs does
        mask_ratio = 0.75 if opts["mask_ratio"] is None else opts["mask_ratio"]
        if opts["contr_bs"]:
            assert opts["contr_bs"] == imgs.shape[0], "contr_bs must equal the batch size (NT-Xent masks are per batch)"
        loss, out = self._forward_two_scale(imgs, imgs2, mask_ratio, opts["mask_seed"], opts["consistent_mask"], noise)
        pred, mask = out["pred"][0].clone(), out["mask"][0].clone()
        if not opts["return_embeds"]:
            return loss, pred, mask
        return (loss, pred, mask, tuple(e.clone() for e in out["enc_emb"]), tuple(e.clone() for e in out["dec_emb"]))


class MAE_ViT_MsLdCd(MAE_ViT_MsLd):
    """Two scales + the cross-scale decoder loss through the predictor MLP only, no contrastive term
    (reference: models_mae/MAE_ViT_MsLdCd.py:6-65) -- the same kernels with the NT-Xent branch switched off."""

    _use_cd = True
    _use_ce = False

    def __init__(self, loss_cd=None, predictor_hidden_size=2048, **kwargs):
        super().__init__(**kwargs)
        self.loss_cd = loss_cd.lower() if loss_cd is not None else self.loss
        if self.loss_cd != "mse":
            raise NotImplementedError(f"loss_cd={loss_cd!r}: only 'mse' is on the hot path")
        self.predictor = nn.Sequential(
            nn.Linear(self.decoder_embed_dim, predictor_hidden_size),
            nn.BatchNorm1d(self.num_patches),
            nn.ReLU(inplace=True),
            nn.Linear(predictor_hidden_size, self.decoder_embed_dim),
        )


class MAE_ViT_MsLdCeCd(MAE_ViT_MsLd):
    """+ cross-scale decoder loss through the predictor MLP and NT-Xent on the encoder features
    (reference: models_mae/MAE_ViT_MsLdCeCd.py:7-84, models_mae/MLP.py, util/contrast_loss.py)."""

    _use_cd = True
    _use_ce = True
    _positional = ("mask_ratio", "contr_bs", "mask_seed", "return_embeds", "consistent_mask")

    def __init__(self, loss_cd=None, predictor_hidden_size=2048, **kwargs):
        super().__init__(**kwargs)
        self.loss_cd = loss_cd.lower() if loss_cd is not None else self.loss
        if self.loss_cd != "mse":
            raise NotImplementedError(f"loss_cd={loss_cd!r}: only 'mse' is on the hot path")
        # created after initialize_weights(): keeps PyTorch's default init (MAE_ViT_MsLdCeCd.py:16,23-25)
        self.predictor = nn.Sequential(
            nn.Linear(self.decoder_embed_dim, predictor_hidden_size),
            nn.BatchNorm1d(self.num_patches),
            nn.ReLU(inplace=True),
            nn.Linear(predictor_hidden_size, self.decoder_embed_dim),
        )


# ------------------------------------------------------------------------------------------------
# registry (reference: models_mae/__init__.py:22-162; north-star aliases per SURVEY.md 0.5)
# ------------------------------------------------------------------------------------------------
args_mae_vit_base = dict(dim_model=768, encoder_num_layers=12, encoder_num_heads=12, decoder_embed_dim=512,
                         decoder_num_layers=8, decoder_num_heads=16)
args_mae_vit_large = dict(dim_model=1024, encoder_num_layers=24, encoder_num_heads=16, decoder_embed_dim=512,
                          decoder_num_layers=8, decoder_num_heads=16)


def mae_vit_base(**kwargs):
    return MAE_ViT_Baseline(**args_mae_vit_base, **kwargs)


def mae_vit_large(**kwargs):
    return MAE_ViT_Baseline(dim_model=1024, **kwargs)


def mae_vit_base_MsLd(**kwargs):
    return MAE_ViT_MsLd(**args_mae_vit_base, **kwargs)


def mae_vit_base_MsLdCd(**kwargs):
    return MAE_ViT_MsLdCd(**args_mae_vit_base, **kwargs)


def mae_vit_base_MsLdCeCd(**kwargs):
    return MAE_ViT_MsLdCeCd(**args_mae_vit_base, **kwargs)


def mae_vit_large_MsLdCeCd(**kwargs):
    return MAE_ViT_MsLdCeCd(**args_mae_vit_large, **kwargs)


def mae_vit_base_patch16(**kwargs):
    kwargs.setdefault("patch_size", 16)
    return MAE_ViT_MsLdCeCd(**args_mae_vit_base, **kwargs)


def mae_vit_large_patch16(**kwargs):
    kwargs.setdefault("patch_size", 16)
    return MAE_ViT_MsLdCeCd(**args_mae_vit_large, **kwargs)
