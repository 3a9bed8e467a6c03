"""csmae_b200 -- B200-native (sm_100a) implementation of the Cross-Scale MAE pretraining hot path.

Drop-in for the reference's `models_mae` namespace on that path: the constructors below build
nn.Modules with the reference's parameter names and forward signatures; all arithmetic runs in
hand-written CUDA kernels behind the C-ABI declared in include/csmae_b200.h.
"""
from .model import (MAE_ViT_Baseline, MAE_ViT_MsLd, MAE_ViT_MsLdCd, MAE_ViT_MsLdCeCd, args_mae_vit_base, args_mae_vit_large,
                    mae_vit_base, mae_vit_base_MsLd, mae_vit_base_MsLdCd, mae_vit_base_MsLdCeCd, mae_vit_base_patch16, mae_vit_large,
                    mae_vit_large_MsLdCeCd, mae_vit_large_patch16)

from .optim import FusedAdamW, NativeScalerWithGradNormCount
from .parallel import DistributedDataParallel
from .prefetch import DevicePrefetcher

__all__ = ["DevicePrefetcher", "DistributedDataParallel", "FusedAdamW", "NativeScalerWithGradNormCount", "MAE_ViT_Baseline", "MAE_ViT_MsLd", "MAE_ViT_MsLdCd", "MAE_ViT_MsLdCeCd", "mae_vit_base_MsLdCd", "args_mae_vit_base", "args_mae_vit_large",
           "mae_vit_base", "mae_vit_large", "mae_vit_base_MsLd", "mae_vit_base_MsLdCeCd", "mae_vit_large_MsLdCeCd",
           "mae_vit_base_patch16", "mae_vit_large_patch16"]
