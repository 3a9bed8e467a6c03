"""Fixed 2-D sin-cos position embedding (init-time, float64 numpy -> fp32 parameter).

Follows util/pos_embed.py:16-63 of the reference: row h*G + w is
[sin(w*omega) | cos(w*omega) | sin(h*omega) | cos(h*omega)] with omega_k = 10000^(-k / (D/4)); the
w-coordinate comes first because the reference builds its meshgrid w-first (pos_embed.py:24).
Row 0 (cls token) is all zeros.

`sincos_pos_embed_` fills a parameter in place: on a CUDA tensor through the C-ABI op `csm_sincos_pos_embed`
(SURVEY.md 8b's operator list), on a CPU tensor (a module built on the host and moved later) through the numpy
restatement below -- the two agree to the last fp32 bit except where fp64 libm differences straddle a rounding
boundary (tests/test_kernels_gpu.py).
"""
import numpy as np
import torch


def sincos_pos_embed_(param, grid_size, cls_token=True):
    """param: [1, cls + G*G, D] fp32 tensor, filled in place (MAE_ViT_Baseline.py:203-218)."""
    D = param.shape[-1]
    if param.is_cuda:
        from . import _native as nat
        if not param.is_contiguous() or param.dtype != torch.float32:
            raise nat.NativeError("sincos_pos_embed_: needs a contiguous fp32 tensor")
        with torch.cuda.device(param.device):
            nat.call("csm_sincos_pos_embed", param, D, grid_size, 1 if cls_token else 0)
    else:
        param.copy_(torch.from_numpy(get_2d_sincos_pos_embed(D, grid_size, cls_token)).float().unsqueeze(0))
    return param


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    if embed_dim % 4 != 0:
        raise AssertionError("embed_dim must be divisible by 4")
    quarter = embed_dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    coords = np.arange(grid_size, dtype=np.float64)
    w_pos = np.tile(coords, grid_size)        # varies fastest
    h_pos = np.repeat(coords, grid_size)
    out_w = w_pos[:, None] * omega[None, :]
    out_h = h_pos[:, None] * omega[None, :]
    emb = np.concatenate([np.sin(out_w), np.cos(out_w), np.sin(out_h), np.cos(out_h)], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb
