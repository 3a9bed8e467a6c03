"""Restatement of the timm==0.4.12 modules the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference builds its encoder/decoder from
`timm.models.vision_transformer.Block` and `PatchEmbed`
(/root/reference/models_mae/MAE_ViT_Baseline.py:7,75-77,160-188) with timm
pinned to 0.4.12 (/root/reference/env.yml:132).  timm is not vendored in the
reference and cannot be installed offline, so its published 0.4.12 semantics
are restated here.  "parity unpinned" for this file: no timm source or test
vector is available; the restatement is corroborated only by the sub-module
names the reference itself hard-codes when it remaps checkpoints
(/root/reference/main_finetune.py:557-566: norm1, attn.qkv, attn.proj, norm2,
mlp.fc1, mlp.fc2) and by models_vit.py:39-60 (patch_embed, blocks).

Only the kwargs the reference passes are supported; dropout / drop_path are 0
on the hot path (identity) and anything else is rejected loudly.
"""
import torch
import torch.nn as nn


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (timm 0.4.12 layers/mlp.py semantics)."""

    def __init__(self, in_features, hidden_features, drop=0.0):
        super().__init__()
        assert drop == 0.0, "hot path uses drop=0"
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    """Materialised multi-head self-attention (timm 0.4.12 vision_transformer.Attention)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        assert attn_drop == 0.0 and proj_drop == 0.0, "hot path uses dropout=0"
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads)
        qkv = qkv.permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Block(nn.Module):
    """Pre-norm transformer block: x += attn(norm1(x)); x += mlp(norm2(x))."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        assert drop == 0.0 and attn_drop == 0.0 and drop_path == 0.0
        assert act_layer is nn.GELU
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class PatchEmbed(nn.Module):
    """Conv2d(k=s=patch) -> flatten(2).transpose(1, 2) (timm 0.4.12 layers/patch_embed.py)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], "input size mismatch"
        return self.proj(x).flatten(2).transpose(1, 2)
