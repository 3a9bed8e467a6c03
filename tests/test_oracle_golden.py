"""CPU: pin oracle/restatement.py against fixtures produced by the REAL reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the live classes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import ref_loader as rl


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    sd = {k[3:]: v for k, v in t.items() if k.startswith("sd/")}
    gr = {k[5:]: v for k, v in t.items() if k.startswith("grad/")}
    return t, sd, gr


def anchors(golden_dir):
    with open(os.path.join(golden_dir, "anchors.json")) as f:
        return json.load(f)


def test_restatement_matches_reference_cecd(golden_dir):
    t, sd, gr = load(golden_dir, "tiny_cecd.npz")
    cfg = anchors(golden_dir)["tiny_config"]
    rm = torch.zeros(16)
    rv = torch.ones(16)
    sd_run = {k: v for k, v in sd.items() if "running" not in k and "num_batches" not in k}
    leaves = {k: v.clone().requires_grad_(k not in R.FROZEN_KEYS) for k, v in sd_run.items()}
    out = R.cross_scale_forward(leaves, t["imgs1"], t["imgs2"], t["noise1"], t["noise2"], 0.75,
                                cfg["encoder_num_heads"], cfg["decoder_num_heads"], running=(rm, rv))
    # CPU fp32, same op sequence -> essentially bit-exact
    assert abs(out["loss"].item() - t["loss"].item()) <= 2e-6 * abs(t["loss"].item())
    torch.testing.assert_close(out["pred"], t["pred"], rtol=1e-5, atol=1e-6)
    assert torch.equal(out["mask"], t["mask"])                      # bit-exact masking
    torch.testing.assert_close(out["enc_emb"][0], t["enc1"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["enc_emb"][1], t["enc2"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["dec_emb"][0], t["dec1"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["dec_emb"][1], t["dec2"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rm, t["bn_running_mean"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(rv, t["bn_running_var"], rtol=1e-5, atol=1e-7)
    out["loss"].backward()
    for k, g in gr.items():
        if g.numel() == 0:     # frozen pos-embeds and the dead encoder_norm (Baseline.py:264)
            assert leaves[k].grad is None, k
            continue
        torch.testing.assert_close(leaves[k].grad, g, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")
    assert sorted(k for k, v in leaves.items() if v.requires_grad and v.grad is None) == \
        anchors(golden_dir)["tiny_cecd_unused_grads"]


def test_restatement_matches_reference_baseline(golden_dir):
    t, sd, gr = load(golden_dir, "tiny_baseline.npz")
    cfg = anchors(golden_dir)["tiny_config"]
    leaves = {k: v.clone().requires_grad_(k not in R.FROZEN_KEYS) for k, v in sd.items()}
    # mask_seed=99 re-seeds the global generator (Baseline.py:301-302); the noise is then the
    # first torch.rand draw -- reproduce it instead of reading the recorded tensor.
    torch.manual_seed(99)
    noise = torch.rand(3, 16)
    assert torch.equal(noise, t["noise"])
    out = R.baseline_pass(leaves, t["imgs"], noise, 0.75, cfg["encoder_num_heads"], cfg["decoder_num_heads"])
    assert abs(out["loss"].item() - t["loss"].item()) <= 2e-6 * abs(t["loss"].item())
    torch.testing.assert_close(out["pred"], t["pred"], rtol=1e-5, atol=1e-6)
    assert torch.equal(out["mask"], t["mask"])
    out["loss"].backward()
    for k, g in gr.items():
        if g.numel():
            torch.testing.assert_close(leaves[k].grad, g, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")


def test_misc_known_answers(golden_dir):
    z = np.load(os.path.join(golden_dir, "misc.npz"))
    for dim, grid in ((768, 14), (512, 14), (64, 4)):
        np.testing.assert_allclose(R.sincos_pos_embed_2d(dim, grid), z[f"pos_{dim}_{grid}"], rtol=0, atol=1e-12)
    for bs in (4, 64):
        val = R.ntxent(torch.from_numpy(z[f"ntx_f1_{bs}"]), torch.from_numpy(z[f"ntx_f2_{bs}"]))
        np.testing.assert_allclose(val.item(), z[f"ntx_loss_{bs}"].item(), rtol=1e-6)
    for e, v in zip(z["lr_epochs"], z["lr_values"]):
        assert abs(R.lr_at(float(e), 1e-3, 1e-5, 40, 400) - v) <= 1e-15


def test_masking_edge_cases():
    # ties broken by index; len_keep = int(L*(1-r)) truncation (Shared.py:64)
    noise = torch.tensor([[0.5, 0.5, 0.1, 0.5, 0.9, 0.1, 0.3]])
    keep, mask, restore = R.masking_from_noise(noise, 0.75)
    assert keep.tolist() == [[2]] and int(mask.sum()) == 6
    assert restore.tolist() == [[3, 4, 0, 5, 6, 1, 2]]
    for L, r, k in ((196, 0.75, 49), (784, 0.75, 196), (64, 0.9, 6), (16, 0.0, 16)):
        keep, mask, restore = R.masking_from_noise(torch.rand(2, L), r)
        assert keep.shape[1] == k and int(mask[0].sum()) == L - k
        assert torch.equal(torch.sort(restore, dim=1).values, torch.arange(L).expand(2, L))


@pytest.mark.skipif(not rl.reference_available(), reason="live reference only exists in the build container")
def test_restatement_matches_live_reference_random_config():
    """A second, different configuration run against the live reference (not a committed fixture)."""
    import contextlib
    import io
    _, _, CeCd = rl.reference_classes()
    cfg = dict(dim_model=128, encoder_num_layers=1, encoder_num_heads=2, decoder_embed_dim=64,
               decoder_num_layers=1, decoder_num_heads=2, input_size=96, patch_size=16,
               predictor_hidden_size=64)
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        m = CeCd(**cfg, device="cpu")
    imgs1, imgs2 = torch.randn(3, 3, 96, 96), torch.randn(3, 3, 96, 96)
    m.crop = rl.FixedScale2(imgs2)
    sd = {k: v.clone() for k, v in m.state_dict().items() if "running" not in k and "num_batches" not in k}
    torch.manual_seed(42)
    loss, pred, mask = m(imgs1, mask_ratio=0.6)
    torch.manual_seed(42)
    n1, n2 = torch.rand(3, 36), torch.rand(3, 36)
    out = R.cross_scale_forward(sd, imgs1, imgs2, n1, n2, 0.6, 2, 2)
    assert abs(out["loss"].item() - loss.item()) <= 2e-6 * abs(loss.item())
    assert torch.equal(out["mask"], mask)
    torch.testing.assert_close(out["pred"], pred, rtol=1e-5, atol=1e-6)
