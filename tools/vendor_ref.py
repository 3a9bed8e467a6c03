"""Copies the reference files of the hot path VERBATIM from /root/reference into oracle/_ref/ (git-ignored, not
gpurun-ignored: it travels to the GPU box, where /root/reference does not exist, but never enters the history).

    python tools/vendor_ref.py            # also run by __graft_entry__.build() when /root/reference is present

What travels and why:
  engine_pretrain.py                       the unmodified step loop the module drops into (row a1)
  util/{misc,lr_sched,pos_embed,contrast_loss}.py
  models_mae/{MAE_ViT_Shared,MAE_ViT_Baseline,MAE_ViT_MsLd,MAE_ViT_MsLdCeCd,MAE_ViT_MsLdCd,MLP}.py
                                           the verbatim model classes: bench.py --impl reference and the engine test
Nothing here is product code: only tests/, smoke() and bench.py's reference arm read oracle/_ref.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CSM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "engine_pretrain.py",
    "util/misc.py", "util/lr_sched.py", "util/pos_embed.py", "util/contrast_loss.py",
    "models_mae/MAE_ViT_Shared.py", "models_mae/MAE_ViT_Baseline.py", "models_mae/MAE_ViT_MsLd.py",
    "models_mae/MAE_ViT_MsLdCeCd.py", "models_mae/MAE_ViT_MsLdCd.py", "models_mae/MLP.py",
]


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models_mae")):
        if verbose:
            print(f"vendor_ref: {SRC} not present; keeping {DST} as it is", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "models_mae"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"vendor_ref: {len(FILES)} files -> {DST}", file=sys.stderr)
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
