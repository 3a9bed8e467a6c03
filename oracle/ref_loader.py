"""Import the REAL reference classes: from /root/reference in the build container, from the verbatim copy in
oracle/_ref (tools/vendor_ref.py; git-ignored, travels to the GPU box) elsewhere.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by tests/golden/make_golden.py to produce the committed
fixtures, by tests that check the restatement against the live reference, by the GPU test that runs the unmodified
engine_pretrain.train_one_epoch, and by bench.py's `--impl reference` arm.  Nothing may read /root/reference at run
time on the GPU box.

Why shims are needed: `import models_mae` fails as
shipped -- models_mae/__init__.py:16-19 star-imports four modules that are not
in the repo, and MAE_ViT_Baseline.py:7-8 / MAE_ViT_Shared.py:4-5 import timm,
xformers and pytorch_msssim, none installable offline.  We register
  * timm.models.vision_transformer.{Block,PatchEmbed} -> oracle/timm_shim.py
  * timm.loss.SoftTargetCrossEntropy, xformers.factory.{xFormer,xFormerConfig},
    pytorch_msssim.{ssim,ms_ssim} -> inert placeholders (never reached on the
    default path: use_xformers=False, loss="mse")
  * an empty `models_mae` package whose __path__ is the reference directory, so
    `from models_mae.MAE_ViT_MsLdCeCd import ...` executes the class files
    verbatim while skipping the broken __init__.
"""
import importlib
import os
import sys
import types

# /root/reference in the build container; on the GPU box the verbatim copy tools/vendor_ref.py placed in oracle/_ref
# (git-ignored, travels with the snapshot)
_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("CSM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models_mae")) and os.path.isdir(os.path.join(_VENDORED, "models_mae")):
    REFERENCE_ROOT = _VENDORED

# size dicts restated from /root/reference/models_mae/__init__.py:42-58
ARGS_VIT_BASE = dict(dim_model=768, encoder_num_layers=12, encoder_num_heads=12,
                     decoder_embed_dim=512, decoder_num_layers=8, decoder_num_heads=16)
ARGS_VIT_LARGE = dict(dim_model=1024, encoder_num_layers=24, encoder_num_heads=16,
                      decoder_embed_dim=512, decoder_num_layers=8, decoder_num_heads=16)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models_mae"))


def _placeholder(name):
    def _raise(*a, **k):
        raise RuntimeError(f"{name} is a placeholder: not on the hot path")
    return _raise


def install_shims():
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if "models_mae" in sys.modules and getattr(sys.modules["models_mae"], "_csm_shim", False):
        return
    from . import timm_shim

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    timm = mod("timm")
    timm.models = mod("timm.models")
    timm.models.vision_transformer = mod(
        "timm.models.vision_transformer", Block=timm_shim.Block, PatchEmbed=timm_shim.PatchEmbed)
    timm.loss = mod("timm.loss", SoftTargetCrossEntropy=_placeholder("SoftTargetCrossEntropy"))
    xf = mod("xformers")
    xf.factory = mod("xformers.factory", xFormer=_placeholder("xFormer"),
                     xFormerConfig=_placeholder("xFormerConfig"))
    mod("pytorch_msssim", ssim=_placeholder("ssim"), ms_ssim=_placeholder("ms_ssim"))

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    pkg = types.ModuleType("models_mae")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models_mae")]
    pkg._csm_shim = True
    sys.modules["models_mae"] = pkg


def reference_classes():
    """Returns (MAE_ViT_Baseline, MAE_ViT_MsLd, MAE_ViT_MsLdCeCd) -- the verbatim reference classes."""
    install_shims()
    base = importlib.import_module("models_mae.MAE_ViT_Baseline").MAE_ViT_Baseline
    msld = importlib.import_module("models_mae.MAE_ViT_MsLd").MAE_ViT_MsLd
    cecd = importlib.import_module("models_mae.MAE_ViT_MsLdCeCd").MAE_ViT_MsLdCeCd
    return base, msld, cecd


def reference_module(name):
    """Import a reference module that needs no shim (util.pos_embed, util.contrast_loss, engine_pretrain...)."""
    install_shims()
    return importlib.import_module(name)


def FixedScale2(imgs2):
    """nn.Module standing in for `model.crop` so the stock single-input forward
    (MAE_ViT_MsLd.py:52) sees a caller-supplied scale-2 batch: the paired
    oracle of SURVEY.md section 0.6 / 8c."""
    import torch.nn as nn

    class _Fixed(nn.Module):
        def forward(self, imgs):
            return imgs2

    return _Fixed()
