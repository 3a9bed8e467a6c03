"""GPU parity of every C-ABI kernel against the plain torch expression of the reference op it
replaces (fp32 math on the same bf16-rounded inputs).  Index outputs are bit-exact; floating-point
tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def nat():
    from csmae_b200 import _native
    _native.load()
    assert _native.sm_count(0) > 0
    return _native


def rnd(*shape, scale=1.0, dtype=f32, seed=None):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed if seed is not None else (hash(shape) & 0xFFFF))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def close(a, b, rtol, atol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {err.max().item():.3e} " \
                          f"(ref max {b.abs().max().item():.3e})"


# ------------------------------------------------------------------------------------------------ GEMM
# (6400, 768) / (6400, 2304) / (2560, 672): the tile-width chooser takes 160- / 224- / 96-column cluster tiles
GEMM_SHAPES = [(128, 128, 64), (256, 384, 128), (200, 136, 72), (3200, 768, 768), (1000, 2304, 768),
               (394, 64, 64), (64, 192, 64), (12608, 512, 2048), (6400, 768, 256), (6400, 2304, 128), (2560, 672, 128)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_linear_fwd_bf16(nat, M, N, K):
    x, w, b = rnd(M, K, dtype=bf16), rnd(N, K, scale=K ** -0.5, dtype=bf16), rnd(N)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
    nat.call("csm_linear_fwd", x, w, b, out, None, M, N, K, nat.EPI_BF16)
    ref = x.float() @ w.float().t() + b
    close(out, ref, 1e-2, 1e-2, "linear_fwd bf16")      # bf16 output rounding (2^-8 relative)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1000, 3072, 768), (200, 136, 72), (6400, 768, 128),
                                   (6400, 2304, 64), (2560, 672, 128)])
def test_linear_fwd_gelu_resid_f32(nat, M, N, K):
    x, w, b = rnd(M, K, dtype=bf16), rnd(N, K, scale=K ** -0.5, dtype=bf16), rnd(N)
    ref = x.float() @ w.float().t() + b
    gp = torch.empty(M, N, device="cuda", dtype=bf16)
    act = torch.empty(M, N, device="cuda", dtype=bf16)
    nat.call("csm_linear_fwd", x, w, b, gp, act, M, N, K, nat.EPI_GELU)
    # GELU (erf form, nn.GELU()) and its derivative of the bf16-rounded pre-activation; the pre-activation is
    # recomputed here with a different fp32 summation order, so an element that sits on a bf16 rounding
    # boundary may land one bf16 step (2^-8 relative) away: rtol 1e-2 covers output rounding + that step
    hq = ref.to(bf16).float().requires_grad_(True)
    gref = F.gelu(hq)
    gref.sum().backward()
    close(act, gref.detach(), 1e-2, 4e-3, "gelu epilogue: activation")
    close(gp, hq.grad, 1e-2, 4e-3, "gelu epilogue: derivative")
    resid = rnd(M, N, seed=5)
    out = torch.empty(M, N, device="cuda")
    nat.call("csm_linear_fwd", x, w, b, out, resid, M, N, K, nat.EPI_RESID)
    close(out - resid, ref, 1e-2, 1e-2, "residual epilogue")   # out = resid + bf16(acc + bias)
    o32 = torch.empty(M, N, device="cuda")
    nat.call("csm_linear_fwd", x, w, b, o32, None, M, N, K, nat.EPI_F32)
    close(o32, ref, 1e-4, 1e-4, "f32 epilogue")          # only fp32 accumulation-order differences


def test_gelu_epilogue_accuracy(nat):
    """gelu / gelu' of the fc1 epilogue on exactly known pre-activations (K = 8, one-hot weights) against
    float64 erf GELU: error must stay inside the bf16 output rounding (2^-8 relative) plus 1e-6 absolute
    (the Abramowitz-Stegun 26.2.17 tail error of 7.5e-8 on Phi, times |x| <= 8)."""
    M, N, K = 512, 64, 8
    vals = torch.linspace(-8, 8, M * N, device="cuda").to(bf16).view(M, N)
    # h[m, n] = x[m, :] . w[n, :]: put the wanted value in x[m, n % 8] ... needs N <= K, so use bias instead
    x = torch.zeros(M, K, device="cuda", dtype=bf16)
    w = torch.zeros(N, K, device="cuda", dtype=bf16)
    x[:, 0] = torch.linspace(-8, 8, M, device="cuda").to(bf16)
    w[:, 0] = 1.0
    bias = torch.linspace(-0.5, 0.5, N, device="cuda")
    gp = torch.empty(M, N, device="cuda", dtype=bf16)
    act = torch.empty(M, N, device="cuda", dtype=bf16)
    nat.call("csm_linear_fwd", x, w, bias, gp, act, M, N, K, nat.EPI_GELU)
    h = (x[:, :1].float() + bias[None, :]).to(bf16).double().requires_grad_(True)
    g = 0.5 * h * (1 + torch.erf(h / math.sqrt(2)))
    g.sum().backward()
    close(act, g.detach().float(), 2 ** -8, 1e-6, "gelu accuracy")
    close(gp, h.grad.float(), 2 ** -8, 1e-6, "gelu' accuracy")
    del vals


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 384, 128), (200, 136, 72), (3200, 2304, 768),
                                   (1000, 768, 3072), (394, 64, 64)])
def test_linear_dgrad(nat, M, N, K):
    dy, w = rnd(M, N, dtype=bf16), rnd(N, K, scale=N ** -0.5, dtype=bf16)
    ref = dy.float() @ w.float()
    dx = torch.empty(M, K, device="cuda")
    nat.call("csm_linear_dgrad", dy, w, dx, None, M, N, K, nat.EPI_F32)
    close(dx, ref, 1e-4, 1e-4, "dgrad f32")
    dxb = torch.empty(M, K, device="cuda", dtype=bf16)
    nat.call("csm_linear_dgrad", dy, w, dxb, None, M, N, K, nat.EPI_BF16)
    close(dxb, ref, 1e-2, 1e-2, "dgrad bf16")
    gp = rnd(M, K, dtype=bf16, seed=9)          # the stored gelu'(h)
    dh = torch.empty(M, K, device="cuda", dtype=bf16)
    nat.call("csm_linear_dgrad", dy, w, dh, gp, M, N, K, nat.EPI_DGELU)
    close(dh, ref.to(bf16).float() * gp.float(), 1e-2, 1e-2, "dgrad x gelu'")


@pytest.mark.parametrize("rows,N,K", [(128, 128, 128), (1000, 256, 192), (6400, 2304, 768), (12608, 512, 2048),
                                      (777, 64, 136), (50, 192, 64)])
def test_linear_wgrad(nat, rows, N, K):
    dy, x = rnd(rows, N, dtype=bf16), rnd(rows, K, dtype=bf16)
    dw = torch.zeros(N, K, device="cuda")
    nat.call("csm_linear_wgrad", dy, x, dw, rows, N, K, 148)
    ref = dy.float().t() @ x.float()
    close(dw, ref, 2e-4, 2e-3 * math.sqrt(rows / 128), "wgrad")   # fp32 split-K accumulation order
    nat.call("csm_linear_wgrad", dy, x, dw, rows, N, K, 148)       # accumulates
    close(dw, 2 * ref, 2e-4, 4e-3 * math.sqrt(rows / 128), "wgrad accumulate")


@pytest.fixture()
def stream_k(nat):
    nat.enable_gemm_stream_k("cuda:0", True)
    yield
    nat.enable_gemm_stream_k("cuda:0", False)


# shapes whose 256 x BN tiles do not fill whole rounds of the 74 SM pairs (the encoder's M = 6400 ones, the decoder's
# M = 25216 ones, ViT-L's M = 3200 with a half-empty last row tile, a ragged one)
@pytest.mark.parametrize("M,N,K", [(6400, 768, 768), (6400, 2304, 768), (6400, 768, 3072), (6400, 3072, 768),
                                   (25216, 512, 512), (25216, 1536, 512), (3200, 1024, 4096), (19000, 520, 328)])
def test_linear_stream_k_all_epilogues(nat, stream_k, M, N, K):
    """With the stream-K workspace registered the forward / dgrad GEMMs cut the k-block stream evenly between the
    clusters; a unit cut in two is finished through an fp32 partial tile.  Same tolerances as the whole-tile path,
    twice in a row (the arrival counters must be re-armed) and against the whole-tile result."""
    x, w, b = rnd(M, K, dtype=bf16), rnd(N, K, scale=K ** -0.5, dtype=bf16), rnd(N)
    ref = x.float() @ w.float().t() + b
    resid = rnd(M, N, seed=5)
    for rep in range(2):
        o32 = torch.full((M, N), float("nan"), device="cuda")
        nat.call("csm_linear_fwd", x, w, b, o32, None, M, N, K, nat.EPI_F32)
        close(o32, ref, 1e-4, 1e-4, f"stream-K f32 (rep {rep})")
        out = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
        nat.call("csm_linear_fwd", x, w, b, out, None, M, N, K, nat.EPI_BF16)
        close(out, ref, 1e-2, 1e-2, f"stream-K bf16 (rep {rep})")
        o = torch.full((M, N), float("nan"), device="cuda")
        nat.call("csm_linear_fwd", x, w, b, o, resid, M, N, K, nat.EPI_RESID)
        close(o - resid, ref, 1e-2, 1e-2, f"stream-K residual (rep {rep})")
        gp = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
        act = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
        nat.call("csm_linear_fwd", x, w, b, gp, act, M, N, K, nat.EPI_GELU)
        close(act, F.gelu(ref.to(bf16).float()), 1e-2, 4e-3, f"stream-K gelu (rep {rep})")
    # dgrad: dX[M,K] = dY[M,N] . W[N,K]
    dy = rnd(M, N, dtype=bf16, seed=3)
    wd = rnd(N, K, scale=N ** -0.5, dtype=bf16, seed=4)
    dref = dy.float() @ wd.float()
    d32 = torch.full((M, K), float("nan"), device="cuda")
    nat.call("csm_linear_dgrad", dy, wd, d32, None, M, N, K, nat.EPI_F32)
    close(d32, dref, 1e-4, 1e-4, "stream-K dgrad f32")
    gpk = rnd(M, K, dtype=bf16, seed=9)
    dh = torch.full((M, K), float("nan"), device="cuda", dtype=bf16)
    nat.call("csm_linear_dgrad", dy, wd, dh, gpk, M, N, K, nat.EPI_DGELU)
    close(dh, dref.to(bf16).float() * gpk.float(), 1e-2, 1e-2, "stream-K dgrad x gelu'")
    # whole-tile scheduling gives the same f32 result up to summation order
    nat.enable_gemm_stream_k("cuda:0", False)
    o_dp = torch.empty(M, N, device="cuda")
    nat.call("csm_linear_fwd", x, w, b, o_dp, None, M, N, K, nat.EPI_F32)
    close(o32, o_dp, 1e-5, 1e-5, "stream-K vs whole-tile")


def test_colsum(nat):
    dy = rnd(1234, 768, dtype=bf16)
    db = torch.zeros(768, device="cuda")
    nat.call("csm_colsum_bf16", dy, db, 1234, 768, 0, 148)
    close(db, dy.float().sum(0), 1e-4, 1e-3, "colsum")
    db.zero_()
    nat.call("csm_colsum_bf16", dy, db, 1234, 768, 50, 148)
    keep = torch.arange(1234, device="cuda") % 50 != 0
    close(db, dy.float()[keep].sum(0), 1e-4, 1e-3, "colsum skip")


# ------------------------------------------------------------------------------------------------ masking
@pytest.mark.parametrize("nimg,L,ratio", [(64, 196, 0.75), (16, 784, 0.75), (3, 16, 0.75), (5, 64, 0.9), (2, 36, 0.6)])
def test_random_masking_bit_exact(nat, nimg, L, ratio):
    torch.manual_seed(L)
    noise = torch.rand(nimg, L, device="cuda")
    noise[0, 3] = noise[0, 1]                 # force ties: stable-by-index order is the contract
    noise[-1, :4] = 0.5
    keep = int(L * (1 - ratio))
    ids_restore = torch.empty(nimg, L, dtype=torch.int64, device="cuda")
    ids_shuffle = torch.empty(nimg, L, dtype=torch.int32, device="cuda")
    mask = torch.empty(nimg, L, device="cuda")
    nat.call("csm_random_masking", noise, nimg, L, keep, ids_restore, ids_shuffle, mask)
    sh = torch.argsort(noise, dim=1, stable=True)
    rs = torch.argsort(sh, dim=1, stable=True)
    m = torch.ones(nimg, L, device="cuda")
    m[:, :keep] = 0
    m = torch.gather(m, 1, rs)
    assert torch.equal(ids_restore, rs)
    assert torch.equal(ids_shuffle.long(), sh)
    assert torch.equal(mask, m)


@pytest.mark.parametrize("L,ratio", [(196, 0.75), (784, 0.75)])
def test_random_masking_matches_default_argsort_over_many_seeds(nat, L, ratio):
    """What the reference actually calls is torch.argsort WITHOUT stable=True (MAE_ViT_Shared.py:69-72) on
    torch.rand noise: over 1024 seeds x 64 images the kernel's stable-by-index contract must reproduce its
    ids_restore / mask exactly wherever the row has no tied noise values (fp32 ties hit ~1e-3..2e-2 of the rows; there
    the unstable order is implementation-defined and the kept SET may legitimately differ only if a tie straddles the
    keep boundary -- such rows are counted and must be rare)."""
    nimg = 64
    keep = int(L * (1 - ratio))
    ids_restore = torch.empty(nimg, L, dtype=torch.int64, device="cuda")
    ids_shuffle = torch.empty(nimg, L, dtype=torch.int32, device="cuda")
    mask = torch.empty(nimg, L, device="cuda")
    rows = tie_rows = mismatch_tie_rows = 0
    for seed in range(1024):
        torch.manual_seed(seed)
        noise = torch.rand(nimg, L, device="cuda")
        nat.call("csm_random_masking", noise, nimg, L, keep, ids_restore, ids_shuffle, mask)
        sh = torch.argsort(noise, dim=1)                       # default (unstable) sort, as upstream
        rs = torch.argsort(sh, dim=1)
        m = torch.ones(nimg, L, device="cuda")
        m[:, :keep] = 0
        m = torch.gather(m, 1, rs)
        srt = noise.sort(dim=1).values
        tied = (srt[:, 1:] == srt[:, :-1]).any(dim=1)
        same = (ids_restore == rs).all(dim=1) & (mask == m).all(dim=1)
        assert bool(same[~tied].all()), f"seed {seed}: a tie-free row differs from torch.argsort"
        rows += nimg
        tie_rows += int(tied.sum())
        mismatch_tie_rows += int((~same & tied).sum())
    print(f"L={L}: {rows} rows, {tie_rows} with tied noise, {mismatch_tie_rows} of those ordered differently by the "
          f"unstable sort")
    assert tie_rows < 0.05 * rows


def test_patch_gather_and_assemble(nat):
    nimg, C, H, p, D, Dd = 5, 3, 64, 16, 64, 32
    L, keep = 16, 4
    Se, Sd = keep + 1, L + 1
    imgs = rnd(nimg, C, H, H)
    noise = torch.rand(nimg, L, device="cuda")
    sh = torch.argsort(noise, dim=1, stable=True)
    rs = torch.argsort(sh, dim=1, stable=True)
    sh32 = sh.int().contiguous()
    out = torch.full((nimg * Se, C * p * p), float("nan"), device="cuda", dtype=bf16)
    nat.call("csm_patch_gather", imgs, sh32, out, nimg, C, H, p, L, keep)
    # conv-order patches: [n, l, (c, py, px)]
    pat = imgs.reshape(nimg, C, H // p, p, H // p, p).permute(0, 2, 4, 1, 3, 5).reshape(nimg, L, C * p * p)
    ref = torch.gather(pat, 1, sh[:, :keep].unsqueeze(-1).expand(-1, -1, C * p * p))
    o = out.view(nimg, Se, -1)
    assert torch.equal(o[:, 0], torch.zeros_like(o[:, 0]))
    assert torch.equal(o[:, 1:], ref.to(bf16))

    emb = rnd(nimg * Se, D, dtype=bf16)
    pos, cls = rnd(L + 1, D), rnd(D)
    x = torch.empty(nimg * Se, D, device="cuda")
    nat.call("csm_encoder_assemble", emb, sh32, pos, cls, x, nimg, L, keep, D)
    xr = torch.empty(nimg, Se, D, device="cuda")
    xr[:, 0] = cls + pos[0]
    xr[:, 1:] = emb.view(nimg, Se, D)[:, 1:].float() + pos[1:][sh[:, :keep]]
    assert torch.equal(x.view(nimg, Se, D), xr)

    demb = rnd(nimg * Se, Dd, dtype=bf16)
    mtok, dpos = rnd(Dd), rnd(Sd, Dd)
    y = torch.empty(nimg * Sd, Dd, device="cuda")
    nat.call("csm_decoder_assemble", demb, rs.contiguous(), mtok, dpos, y, nimg, L, keep, Dd)
    de = demb.view(nimg, Se, Dd).float()
    x_ = torch.cat([de[:, 1:], mtok.expand(nimg, L - keep, Dd)], 1)
    x_ = torch.gather(x_, 1, rs.unsqueeze(-1).expand(-1, -1, Dd))
    yr = torch.cat([de[:, :1], x_], 1) + dpos
    assert torch.equal(y.view(nimg, Sd, Dd), yr)

    dy = rnd(nimg * Sd, Dd)
    d_demb = torch.empty(nimg * Se, Dd, device="cuda", dtype=bf16)
    d_mtok = torch.zeros(Dd, device="cuda")
    nat.call("csm_decoder_assemble_bwd", dy, sh32, d_demb, d_mtok, nimg, L, keep, Dd)
    de_l = de.clone().requires_grad_(True)
    mt_l = mtok.clone().requires_grad_(True)
    x_ = torch.cat([de_l[:, 1:], mt_l.expand(nimg, L - keep, Dd)], 1)
    x_ = torch.gather(x_, 1, rs.unsqueeze(-1).expand(-1, -1, Dd))
    (torch.cat([de_l[:, :1], x_], 1) + dpos).backward(dy.view(nimg, Sd, Dd))
    assert torch.equal(d_demb.view(nimg, Se, Dd), de_l.grad.to(bf16))
    close(d_mtok, mt_l.grad, 1e-5, 1e-5, "mask_token grad")


def test_encoder_out_grad_and_cls(nat):
    nimg, Se, D = 6, 5, 64
    d_enc, d_feat = rnd(nimg * Se, D, dtype=bf16), rnd(nimg, D)
    dx = torch.empty(nimg * Se, D, device="cuda")
    dx16 = torch.empty(nimg * Se, D, device="cuda", dtype=bf16)
    nat.call("csm_encoder_out_grad", d_enc, d_feat, dx, dx16, nimg, Se, D)
    ref = d_enc.float().view(nimg, Se, D).clone()
    ref[:, 1:] += (d_feat / (Se - 1)).unsqueeze(1)
    close(dx.view(nimg, Se, D), ref, 1e-6, 1e-6, "encoder_out_grad")
    assert torch.equal(dx16, dx.to(bf16))
    d_cls = torch.empty(D, device="cuda")
    nat.call("csm_cls_grad", dx, d_cls, nimg, Se, D)
    close(d_cls, dx.view(nimg, Se, D)[:, 0].sum(0), 1e-5, 1e-5, "cls grad")


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,D", [(3200, 768), (1000, 512), (300, 1024), (77, 64), (25216, 512), (131, 128), (3, 384)])
def test_layernorm_fwd_bwd(nat, rows, D):
    x, g, b = rnd(rows, D, scale=2.0), 1 + 0.1 * rnd(D), 0.1 * rnd(D, seed=3)
    o16 = torch.empty(rows, D, device="cuda", dtype=bf16)
    o32 = torch.empty(rows, D, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    nat.call("csm_layernorm_fwd", x, g, b, o16, o32, mean, rstd, rows, D, 1e-6)
    ref = F.layer_norm(x, (D,), g, b, 1e-6)
    close(o32, ref, 1e-5, 1e-5, "LN fwd f32")
    assert torch.equal(o16, o32.to(bf16))
    dy16, dy2, dres_in = rnd(rows, D, dtype=bf16, seed=1), rnd(rows, D, seed=2), rnd(rows, D, seed=4)
    dres = torch.empty(rows, D, device="cuda")
    dres16 = torch.empty(rows, D, device="cuda", dtype=bf16)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dcs = torch.zeros(D, device="cuda")
    nat.call("csm_layernorm_bwd", dy16, dy2, x, mean, rstd, g, dres_in, dres, dres16, dg, db, dcs, rows, D, 148)
    xl, gl, bl = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xl, (D,), gl, bl, 1e-6).backward(dy16.float() + dy2)
    close(dres, dres_in + xl.grad, 1e-4, 1e-4, "LN bwd dx")
    assert torch.equal(dres16, dres.to(bf16))
    close(dg, gl.grad, 1e-4, 1e-3 * math.sqrt(rows / 100), "LN dgamma")
    close(db, bl.grad, 1e-4, 1e-3 * math.sqrt(rows / 100), "LN dbeta")
    # fused bias gradient of the Linear whose dY is dres16: column sums of the bf16 copy
    close(dcs, dres16.float().sum(0), 1e-4, 1e-3 * math.sqrt(rows / 100), "LN bwd fused column sum")
    # in-place residual-gradient update + bf16-only incoming gradient
    d2 = dres_in.clone()
    nat.call("csm_layernorm_bwd", dy16, None, x, mean, rstd, g, d2, d2, dres16, dg, db, None, rows, D, 148)
    xl.grad = None
    F.layer_norm(xl, (D,), g, b, 1e-6).backward(dy16.float())
    close(d2, dres_in + xl.grad, 1e-4, 1e-4, "LN bwd in place")


def test_casts(nat):
    a = rnd(1000003)
    o = torch.empty(1000003, device="cuda", dtype=bf16)
    nat.call("csm_cast_f32_bf16", a, o, a.numel())
    assert torch.equal(o, a.to(bf16))
    srcs = [rnd(n, seed=n) for n in (8, 768 * 768, 2304 * 768, 12)]
    dsts = [torch.empty(s.numel(), device="cuda", dtype=bf16) for s in srcs]
    table = []
    for s, d in zip(srcs, dsts):
        table += [s.data_ptr(), d.data_ptr(), s.numel()]
    table = torch.tensor(table, dtype=torch.int64).cuda()
    nat.call("csm_cast_multi", table, len(srcs), 32)
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s.to(bf16))


# ------------------------------------------------------------------------------------------------ attention
def attn_ref(qkv, B, S, H, d):
    """timm 0.4.12 Attention core in fp32 on the bf16-rounded qkv (oracle/timm_shim.py)."""
    q, k, v = qkv.float().view(B, S, 3, H, d).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1)
    return (att @ v).transpose(1, 2).reshape(B * S, H * d)


@pytest.mark.parametrize("B,S,H,d", [(8, 50, 12, 64), (4, 197, 16, 32), (2, 785, 4, 32), (3, 197, 2, 64),
                                     (5, 5, 1, 64), (3, 17, 2, 32), (2, 64, 2, 32), (2, 65, 1, 64), (2, 100, 3, 32),
                                     (1, 257, 2, 32)])
def test_attention_fwd_bwd(nat, B, S, H, d):
    Dm = H * d
    qkv = rnd(B * S, 3 * Dm, dtype=bf16)
    out = torch.full((B * S, Dm), float("nan"), device="cuda", dtype=bf16)
    lse = torch.empty(B * H * S, device="cuda")
    nat.call("csm_attention_fwd", qkv, out, lse, B, S, H, d)
    ql = qkv.float().requires_grad_(True)
    ref = attn_ref(ql, B, S, H, d)
    # bf16 rounding of scores / probabilities / output: ~2^-8 relative per element
    close(out, ref, 2e-2, 2e-2, "attention fwd")
    d_out = rnd(B * S, Dm, dtype=bf16, seed=11)
    ref.backward(d_out.float())
    dqkv = torch.full((B * S, 3 * Dm), float("nan"), device="cuda", dtype=bf16)
    delta = torch.empty(B * H * S, device="cuda")
    dbias = torch.zeros(3 * Dm, device="cuda")
    nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dqkv, dbias, B, S, H, d)
    g = ql.grad
    err = (dqkv.float() - g).abs().max().item()
    assert err <= 3e-2 * g.abs().max().item() + 1e-3, f"attention bwd max err {err:.3e} vs max {g.abs().max():.3e}"
    rel = ((dqkv.float() - g).norm() / g.norm()).item()
    assert rel < 2e-2, f"attention bwd relative L2 error {rel:.3e}"
    # without the fused bias sums the launcher may pick another kernel (two-pass at d = 32): same contract
    dqkv2 = torch.full((B * S, 3 * Dm), float("nan"), device="cuda", dtype=bf16)
    nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dqkv2, None, B, S, H, d)
    err2 = (dqkv2.float() - g).abs().max().item()
    assert err2 <= 3e-2 * g.abs().max().item() + 1e-3, f"attention bwd (no bias sums) max err {err2:.3e}"
    assert ((dqkv2.float() - g).norm() / g.norm()).item() < 2e-2
    # fused qkv.bias gradient: column sums of dqkv over the tokens
    bref = g.sum(0)
    berr = (dbias - bref).abs().max().item()
    assert berr <= 2e-2 * bref.abs().max().item() + 2e-3 * math.sqrt(B * S), f"fused qkv bias grad err {berr:.3e}"


# ------------------------------------------------------------------------------------------------ losses
def patchify(imgs, p, c):
    n, _, hh, _ = imgs.shape
    g = hh // p
    x = imgs.reshape(n, c, g, p, g, p)
    return torch.einsum("nchpwq->nhwpqc", x).reshape(n, g * g, p * p * c)


@pytest.mark.parametrize("norm_pix", [0, 1])
def test_recon_loss(nat, norm_pix):
    nimg, C, H, p = 6, 3, 64, 16
    L, P = 16, 768
    imgs = rnd(nimg, C, H, H)
    pred_full = rnd(nimg * (L + 1), P, dtype=bf16)
    mask = (torch.rand(nimg, L, device="cuda") < 0.75).float()
    acc = torch.zeros(1, device="cuda")
    nat.call("csm_recon_loss_fwd", pred_full, imgs, mask, acc, nimg, C, H, p, L, norm_pix)
    pl = pred_full.float().view(nimg, L + 1, P)[:, 1:].clone().requires_grad_(True)
    tgt = patchify(imgs, p, C)
    if norm_pix:
        tgt = (tgt - tgt.mean(-1, keepdim=True)) / (tgt.var(-1, keepdim=True) + 1e-6) ** 0.5
        diff = pl - tgt
    else:
        diff = pl - tgt.to(bf16).float()          # autocast: einsum rounds the target to bf16
    loss = ((diff ** 2).mean(-1) * mask).sum()
    close(acc, loss.detach().reshape(1), 5e-3, 1e-3, "recon loss sum")       # bf16 rounding of (pred - target)
    g = torch.tensor([3.0], device="cuda")
    coef = 1.0 / (P * mask.sum().item())
    dpred = torch.full((nimg * (L + 1), P), float("nan"), device="cuda", dtype=bf16)
    nat.call("csm_recon_loss_bwd", pred_full, imgs, mask, dpred, g, coef, nimg, C, H, p, L, norm_pix)
    (loss * 3.0 / mask.sum()).backward()
    dp = dpred.float().view(nimg, L + 1, P)
    assert torch.equal(dp[:, 0], torch.zeros_like(dp[:, 0]))
    close(dp[:, 1:], pl.grad, 2e-2, 1e-6, "recon dpred")


def test_cross_mse(nat):
    N, Sd, Dd = 4, 17, 64
    cp, tgt = rnd(N * Sd, Dd, dtype=bf16), rnd(N * Sd, Dd, seed=2)
    acc = torch.zeros(1, device="cuda")
    nat.call("csm_cross_mse_fwd", cp, tgt, acc, N * Sd, Sd, Dd)
    d = (cp.float() - tgt).view(N, Sd, Dd)[:, 1:]
    close(acc, (d ** 2).sum().reshape(1), 1e-4, 1e-3, "cross mse")
    g = torch.tensor([2.0], device="cuda")
    d_cp = torch.empty(N * Sd, Dd, device="cuda", dtype=bf16)
    d_t = torch.empty(N * Sd, Dd, device="cuda")
    coef = 1.0 / (N * (Sd - 1) * Dd)
    nat.call("csm_cross_mse_bwd", cp, tgt, d_cp, d_t, g, coef, N * Sd, Sd, Dd)
    ref = torch.zeros(N, Sd, Dd, device="cuda")
    ref[:, 1:] = 2.0 * coef * 2 * d
    close(d_t.view(N, Sd, Dd), -ref, 1e-5, 1e-7, "cross mse d_tgt")
    close(d_cp.view(N, Sd, Dd), ref, 1e-2, 1e-7, "cross mse d_cp")


def test_loss_finalize_and_zero(nat):
    """csm_zero_async + csm_loss_finalize: the accumulator array every loss kernel adds into is cleared by a memset
    node and dotted with the coefficients of MAE_ViT_MsLdCeCd.py:62-69 / MAE_ViT_MsLd.py:61-66 by one warp."""
    acc = rnd(8, seed=3).abs() * 100
    coefs = torch.tensor([1 / 3.0, 0.25, 1e-3, 1.0, 0.0, 0.0, 0.0, 0.0], device="cuda")
    loss = torch.full((1,), float("nan"), device="cuda")
    nat.call("csm_loss_finalize", acc, coefs, loss, 8)
    close(loss, (acc.double() * coefs.double()).sum().float().reshape(1), 1e-6, 1e-6, "loss finalize")
    big = torch.ones(1 << 20, device="cuda")
    nat.call("csm_zero_async", big[16:], (big.numel() - 32) * 4)
    torch.cuda.synchronize()
    assert big[:16].eq(1).all() and big[-16:].eq(1).all() and big[16:-16].eq(0).all()
    with pytest.raises(nat.NativeError):
        nat.call("csm_loss_finalize", acc, coefs, loss, 33)


# N >= 8 with N/8 rows x Hp x 2 B <= 96 KB (48 KB for the backward): the channel is spread over a cluster of 8 CTAs
# (rows staged in shared memory, sums through DSMEM); (11, ...) leaves two CTAs of the cluster without rows;
# N = 4 and the eval-mode call take the one-CTA-per-channel kernels
@pytest.mark.parametrize("N,L,Hp", [(8, 16, 128), (11, 9, 256), (4, 16, 128), (64, 196, 2048)])
def test_bn_patch(nat, N, L, Hp):
    Sd = L + 1
    h = rnd(N * Sd, Hp, dtype=bf16, scale=2.0)
    gamma, beta = 1 + 0.1 * rnd(L), 0.1 * rnd(L, seed=3)
    rm, rv = torch.zeros(L, device="cuda"), torch.ones(L, device="cuda")
    out = torch.full((N * Sd, Hp), float("nan"), device="cuda", dtype=bf16)
    mean, rstd = torch.empty(L, device="cuda"), torch.empty(L, device="cuda")
    nat.call("csm_bn_patch_fwd", h, gamma, beta, out, mean, rstd, rm, rv, N, L, Hp, 1e-5, 0.1, 1)
    hl = h.float().view(N, Sd, Hp)[:, 1:].clone().requires_grad_(True)
    gl, bl = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(L, device="cuda"), torch.ones(L, device="cuda")
    y = F.relu(F.batch_norm(hl, rm_r, rv_r, gl, bl, training=True, momentum=0.1, eps=1e-5))
    o = out.float().view(N, Sd, Hp)
    assert torch.equal(o[:, 0], torch.zeros_like(o[:, 0]))
    close(o[:, 1:], y, 1e-2, 1e-2, "bn+relu fwd")
    close(rm, rm_r, 1e-4, 1e-5, "running mean")
    close(rv, rv_r, 1e-4, 1e-5, "running var")
    d_out = rnd(N * Sd, Hp, dtype=bf16, seed=7)
    dh = torch.full((N * Sd, Hp), float("nan"), device="cuda", dtype=bf16)
    dg, db = torch.empty(L, device="cuda"), torch.empty(L, device="cuda")
    nat.call("csm_bn_patch_bwd", h, out, d_out, gamma, mean, rstd, dh, dg, db, N, L, Hp, 1)
    y.backward(d_out.float().view(N, Sd, Hp)[:, 1:])
    d = dh.float().view(N, Sd, Hp)
    assert torch.equal(d[:, 0], torch.zeros_like(d[:, 0]))
    close(d[:, 1:], hl.grad, 2e-2, 2e-3, "bn bwd dh")
    close(dg, gl.grad, 1e-2, 5e-2, "bn dgamma")      # relu mask taken from the bf16-rounded output
    close(db, bl.grad, 1e-2, 5e-2, "bn dbeta")
    # eval mode uses the running statistics
    nat.call("csm_bn_patch_fwd", h, gamma, beta, out, mean, rstd, rm, rv, N, L, Hp, 1e-5, 0.1, 0)
    he = hl.detach().clone().requires_grad_(True)
    ge, be = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ye = F.relu(F.batch_norm(he, rm, rv, ge, be, training=False, eps=1e-5))
    close(out.float().view(N, Sd, Hp)[:, 1:], ye, 1e-2, 1e-2, "bn eval")
    # eval-mode backward: the statistics are constants, dh = gamma * rstd * dy (nn.BatchNorm1d.eval())
    close(mean, rm, 0, 0, "bn eval mean kept for the backward")
    nat.call("csm_bn_patch_bwd", h, out, d_out, gamma, mean, rstd, dh, dg, db, N, L, Hp, 0)
    ye.backward(d_out.float().view(N, Sd, Hp)[:, 1:])
    close(dh.float().view(N, Sd, Hp)[:, 1:], he.grad, 2e-2, 2e-3, "bn eval bwd dh")
    close(dg, ge.grad, 1e-2, 5e-2, "bn eval dgamma")
    close(db, be.grad, 1e-2, 5e-2, "bn eval dbeta")


@pytest.mark.parametrize("B,Se,D", [(64, 50, 768), (4, 5, 64), (32, 50, 1024)])
def test_ntxent(nat, B, Se, D):
    from oracle import restatement as R
    x = rnd(2 * B * Se, D)
    zhat, fnorm = torch.empty(2 * B, D, device="cuda"), torch.empty(2 * B, device="cuda")
    neg, acc = torch.empty(2 * B, device="cuda"), torch.zeros(1, device="cuda")
    nat.call("csm_ntxent_fwd", x, zhat, fnorm, neg, acc, B, Se, D, 0.5, 1e-8)
    xl = x.clone().requires_grad_(True)
    f = xl.view(2 * B, Se, D)[:, 1:].mean(1)
    loss = R.ntxent(f[:B], f[B:])
    close(acc, loss.detach().reshape(1), 1e-5, 1e-5, "ntxent loss")
    g = torch.tensor([1.5], device="cuda")
    d_feat = torch.empty(2 * B, D, device="cuda")
    nat.call("csm_ntxent_bwd", zhat, fnorm, neg, g, d_feat, B, D, 0.5, 1e-8)
    f2 = f.detach().clone().requires_grad_(True)
    (R.ntxent(f2[:B], f2[B:]) * 1.5).backward()
    close(d_feat, f2.grad, 1e-3, 1e-6 * f2.grad.abs().max().item() + 1e-9, "ntxent d_feat")


# ------------------------------------------------------------------------------------------------ in-model crop
@pytest.mark.parametrize("size,box", [(224, (17, 30, 150, 190)), (224, (0, 0, 112, 112)), (96, (3, 5, 90, 71)),
                                      (128, (10, 0, 118, 128)), (64, (0, 0, 64, 64)), (96, (0, 0, 96, 48)),
                                      (48, (1, 2, 90, 100))])
def test_resized_crop_matches_torchvision(nat, size, box):
    """csm_resized_crop against torchvision.transforms.functional.resized_crop(bilinear, antialias=True) -- what
    the reference's in-model RandomResizedCrop applies (MAE_ViT_MsLd.py:29-35,52); the last case down-scales."""
    import torchvision.transforms.functional as TF
    top, left, h, w = box
    H, W = max(size, top + h), max(size, left + w)
    imgs = rnd(3, 3, H, W, seed=size + top)
    out = torch.full((3, 3, size, size), float("nan"), device="cuda")
    nat.call("csm_resized_crop", imgs, out, 9, H, W, top, left, h, w, size)
    # ATen's CPU kernel is the yardstick (same taps, weights and fp32 summation order: agreement to 1 ulp);
    # ATen's own CUDA kernel evaluates the filter weights slightly differently and sits ~4e-5 away from both
    ref_cpu = TF.resized_crop(imgs.cpu(), top, left, h, w, [size, size], TF.InterpolationMode.BILINEAR, antialias=True)
    close(out.cpu(), ref_cpu, 1e-6, 1e-6, "resized crop vs ATen CPU")
    ref = TF.resized_crop(imgs, top, left, h, w, [size, size], TF.InterpolationMode.BILINEAR, antialias=True)
    close(out, ref, 1e-4, 1e-4, "resized crop vs ATen CUDA")


def test_batch_crop_module_draws_like_torchvision():
    """BatchRandomResizedCrop consumes the CPU generator exactly like T.RandomResizedCrop and produces the same
    batch (one box for the whole batch)."""
    import csmae_b200
    from torchvision import transforms as T
    x = rnd(4, 3, 96, 96, seed=3)
    ours = csmae_b200.model.BatchRandomResizedCrop(96, (0.25, 0.75))
    ref = T.RandomResizedCrop(size=(96, 96), scale=(0.25, 0.75), antialias=True)
    for seed in (0, 1, 2, 3):
        torch.manual_seed(seed)
        a = ours(x)
        sa = torch.get_rng_state()
        torch.manual_seed(seed)
        b = ref(x)
        sb = torch.get_rng_state()
        assert torch.equal(sa, sb), "RNG consumption differs"
        close(a, b, 1e-4, 1e-4, "batch crop")


# ------------------------------------------------------------------------------------------------ optimizer (row f1)
def test_fused_adamw_matches_torch():
    """csmae_b200.FusedAdamW against torch.optim.AdamW (the reference's optimizer, main_pretrain.py:426-427): two
    parameter groups (weight decay 0 / 0.05), ragged sizes, a learning-rate change between steps, a parameter that
    only receives a gradient from the second step on."""
    import csmae_b200
    torch.manual_seed(0)
    shapes = [(768, 768), (768,), (1, 1, 512), (37,), (5, 3, 16, 16), (2304, 768), (3,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    groups = lambda ps: [{"params": [p for p in ps if p.ndim == 1], "weight_decay": 0.0},
                         {"params": [p for p in ps if p.ndim != 1], "weight_decay": 0.05}]
    ref = torch.optim.AdamW(groups(ref_p), lr=1.5e-3, betas=(0.9, 0.95))
    ours = csmae_b200.FusedAdamW(groups(our_p), lr=1.5e-3, betas=(0.9, 0.95))
    for step in range(4):
        for a, b in zip(ref_p, our_p):
            if step == 0 and a.shape == (37,):
                a.grad = b.grad = None
                continue
            g = torch.randn_like(a) * (10.0 if step == 2 else 1.0)
            a.grad, b.grad = g.clone(), g.clone()
        if step == 2:
            for opt in (ref, ours):
                for gr in opt.param_groups:
                    gr["lr"] = 7e-4
        ref.step()
        ours.step()
        for a, b in zip(ref_p, our_p):
            close(b.detach(), a.detach(), 2e-6, 2e-7, f"FusedAdamW step {step} shape {tuple(a.shape)}")
    sa, sb = ref.state_dict()["state"], ours.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert set(sb[k].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa[k]["step"]) == float(sb[k]["step"])
        close(sb[k]["exp_avg_sq"], sa[k]["exp_avg_sq"], 2e-6, 1e-9, "exp_avg_sq")
    # global gradient norm (util/misc.py:338-355)
    want = torch.norm(torch.stack([torch.norm(p.grad, 2.0) for p in ref_p]), 2.0)
    close(ours.grad_norm().reshape(1), want.reshape(1), 1e-5, 0, "grad norm")
    # resume (util/misc.py:393-411: optimizer.load_state_dict(checkpoint["optimizer"])): a FusedAdamW that already
    # stepped takes over torch.optim.AdamW's checkpoint -- the moment tensors are replaced, the cached device table
    # must follow them -- and keeps tracking the reference optimizer afterwards
    import copy
    import io
    buf = io.BytesIO()
    torch.save(ref.state_dict(), buf)
    buf.seek(0)
    with torch.no_grad():
        for a, b in zip(ref_p, our_p):
            b.copy_(a)
    ours.load_state_dict(torch.load(buf, map_location="cpu"))
    del buf
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(8)]   # reuse freed moment storage
    for step in range(2):
        for a, b in zip(ref_p, our_p):
            g = torch.randn_like(a)
            a.grad, b.grad = g.clone(), g.clone()
        ref.step()
        ours.step()
        for a, b in zip(ref_p, our_p):
            close(b.detach(), a.detach(), 2e-6, 2e-7, f"FusedAdamW after resume, step {step}")
    del junk
    sa, sb = ref.state_dict()["state"], copy.deepcopy(ours.state_dict()["state"])
    for k in sa:
        assert float(sa[k]["step"]) == float(sb[k]["step"])
        # (torch updates exp_avg with lerp, the kernel with beta1 * m + (1 - beta1) * g: same value, different rounding)
        close(sb[k]["exp_avg"], sa[k]["exp_avg"], 1e-5, 1e-6, "exp_avg after resume")


def test_grad_stats_and_amp_update(nat):
    """csm_grad_stats_f32 / csm_amp_update (include/csmae_b200.h): norm of the unscaled gradients + found_inf in one
    pass, then the scalar GradScaler arithmetic (util/misc.py:314-326; torch._amp_update_scale_) on the device."""
    n = 1_000_003
    x = rnd(n + 1, seed=3)[:n] * 700.0
    inv = torch.tensor([1.0 / 1024.0], device="cuda")
    stats = torch.zeros(2, device="cuda")
    nat.call("csm_grad_stats_f32", x, n, inv, stats, nat.sm_count())
    want = (x.double() / 1024.0).pow(2).sum().item()
    assert abs(stats[0].item() - want) <= 1e-5 * want and stats[1].item() == 0.0
    state = torch.tensor([1024.0, 1.0, 1.0 / 1024.0, 0.0], device="cuda")
    ctl, norm = torch.zeros(2, device="cuda"), torch.zeros(1, device="cuda")
    nat.call("csm_amp_update", stats, state, ctl, norm, 0.0, 2.0, 0.5, 2)        # clean step, tracker reaches interval
    assert abs(norm.item() - want ** 0.5) <= 1e-5 * want ** 0.5
    assert ctl.tolist() == [1.0 / 1024.0, 0.0] and state.tolist()[:3] == [2048.0, 0.0, 1.0 / 2048.0]
    nat.call("csm_amp_update", stats, state, ctl, norm, 1.0, 2.0, 0.5, 2)        # clipped to norm 1
    coef = min(1.0, 1.0 / (want ** 0.5 + 1e-6))
    assert abs(ctl[0].item() - coef / 2048.0) <= 1e-6 * coef / 2048.0 and state.tolist()[:2] == [2048.0, 1.0]
    for bad in (float("inf"), float("nan")):
        y = x.clone()
        y[n - 2] = bad                                                           # in the non-vectorised tail
        stats.zero_()
        nat.call("csm_grad_stats_f32", y, n, inv, stats, nat.sm_count())
        assert stats[1].item() == 1.0
    nat.call("csm_amp_update", stats, state, ctl, norm, 0.0, 2.0, 0.5, 2)        # found_inf: backoff, tracker reset
    assert ctl[1].item() == 1.0 and state.tolist()[:3] == [1024.0, 0.0, 1.0 / 1024.0]
    # csm_adamw_multi leaves everything untouched when found_inf is set (covered end to end in
    # test_native_scaler_matches_reference_scaler)


@pytest.mark.parametrize("D,G,cls", [(768, 14, 1), (512, 14, 1), (1024, 28, 1), (64, 6, 0)])
def test_sincos_pos_embed(nat, D, G, cls):
    """csm_sincos_pos_embed against the numpy restatement of util/pos_embed.py:16-63 (itself pinned bit-for-bit to the
    live reference by tests/test_host_cpu.py): both evaluate in fp64 and round once, so the fp32 tables agree except
    where the two libms' last-bit differences straddle an fp32 rounding boundary (at most 1 ulp, a handful of
    entries)."""
    from csmae_b200.pos_embed import get_2d_sincos_pos_embed, sincos_pos_embed_
    want = torch.from_numpy(get_2d_sincos_pos_embed(D, G, cls_token=bool(cls))).float()
    out = torch.full((1, cls + G * G, D), float("nan"), device="cuda")
    sincos_pos_embed_(out, G, cls_token=bool(cls))
    got = out[0].cpu()
    diff = (got - want).abs()
    exact = (got == want).float().mean().item()
    print(f"D={D} G={G}: {exact * 100:.4f} % of entries bit-identical, max diff {diff.max().item():.3e}")
    assert diff.max().item() <= 1.2e-7 and exact >= 0.999
    if cls:
        assert torch.equal(got[0], torch.zeros(D))
