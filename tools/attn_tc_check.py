"""Development check of the tcgen05 attention kernels against a torch fp32 reference, one subprocess per case so that
a trapped kernel (watchdog) cannot take the other cases down.  Usage: python tools/attn_tc_check.py [--time]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-scale-mae_b200"))

SHAPES = [(8, 50, 12, 64), (4, 197, 16, 32), (2, 785, 4, 32), (3, 197, 2, 64), (5, 5, 1, 64), (3, 17, 2, 32),
          (2, 64, 2, 32), (2, 65, 1, 64), (2, 100, 3, 32), (1, 257, 2, 32), (128, 197, 16, 32), (128, 50, 12, 64),
          (32, 785, 16, 32), (32, 197, 16, 64)]


def one(B, S, H, d, variant, timed):
    import torch
    from csmae_b200 import _native as nat
    torch.manual_seed(0)
    Dm = H * d
    qkv = torch.randn(B * S, 3 * Dm, device="cuda").to(torch.bfloat16)
    out = torch.full((B * S, Dm), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((B * H * S,), float("nan"), device="cuda")
    nat.call("csm_attention_fwd_tc", qkv, out, lse, B, S, H, d, variant)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, S, 3, H, d).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-2, -1)) * d ** -0.5
    ref = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B * S, Dm)
    lref = (torch.logsumexp(sc, -1) * 1.4426950408889634).reshape(-1)
    err = (out.float() - ref).abs().max().item()
    lerr = (lse - lref).abs().max().item()
    msg = f"B={B} S={S} H={H} d={d} var={variant}: out max err {err:.3e} (ref max {ref.abs().max():.2f}) lse err {lerr:.3e}"
    ok = err < 2e-2 and lerr < 1e-3
    do_bwd = variant == 0
    if do_bwd:
        d_out = torch.randn(B * S, Dm, device="cuda", generator=torch.Generator(device="cuda").manual_seed(11)).to(torch.bfloat16)
        if B * S * S * H < 3e8:
            ql = qkv.float().requires_grad_(True)
            q, k, v = ql.view(B, S, 3, H, d).permute(2, 0, 3, 1, 4)
            r2 = (((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B * S, Dm)
            r2.backward(d_out.float())
            gref = ql.grad
        else:   # large case: the mma.sync kernel (already pinned against torch) is the reference
            gref = torch.empty_like(qkv)
            delta = torch.empty(B * H * S, device="cuda")
            nat.call("csm_attention_bwd_legacy", qkv, out, d_out, lse, delta, gref, None, B, S, H, d)
            gref = gref.float()
        dqkv = torch.full((B * S, 3 * Dm), float("nan"), device="cuda", dtype=torch.bfloat16)
        delta_tc = torch.empty(B * H * S, device="cuda")
        nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta_tc, dqkv, None, B, S, H, d)
        torch.cuda.synchronize()
        gerr = (dqkv.float() - gref).abs().max().item()
        rel = ((dqkv.float() - gref).norm() / gref.norm()).item()
        per = [((dqkv.float()[:, i * Dm:(i + 1) * Dm] - gref[:, i * Dm:(i + 1) * Dm]).norm() / gref[:, i * Dm:(i + 1) * Dm].norm()).item() for i in range(3)]
        msg += f" | bwd max err {gerr:.3e} (max {gref.abs().max():.2f}) rel {rel:.3e} [dq {per[0]:.1e} dk {per[1]:.1e} dv {per[2]:.1e}]"
        ok = ok and gerr <= 3e-2 * gref.abs().max().item() + 1e-3 and rel < 2e-2
    if timed and ok:
        out2 = torch.empty_like(out)
        def t(fn, n=20):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1000
        t_new = t(lambda: nat.call("csm_attention_fwd_tc", qkv, out2, lse, B, S, H, d, variant))
        t_old = t(lambda: nat.call("csm_attention_fwd_legacy", qkv, out2, lse, B, S, H, d))
        msg += f" | tc {t_new:.1f} us, mma.sync {t_old:.1f} us"
        if do_bwd:
            delta = torch.empty(B * H * S, device="cuda")
            dq2 = torch.empty_like(qkv)
            tb_new = t(lambda: nat.call("csm_attention_bwd", qkv, out, d_out, lse, delta, dq2, None, B, S, H, d))
            tb_old = t(lambda: nat.call("csm_attention_bwd_legacy", qkv, out, d_out, lse, delta, dq2, None, B, S, H, d))
            msg += f" | bwd tc {tb_new:.1f} us, mma.sync {tb_old:.1f} us"
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        B, S, H, d, variant, timed = map(int, sys.argv[2:8])
        sys.exit(one(B, S, H, d, variant, bool(timed)))
    timed = int("--time" in sys.argv)
    bad = 0
    for (B, S, H, d) in SHAPES:
        for variant in (0, 1):
            try:
                r = subprocess.run([sys.executable, __file__, "--case", *map(str, (B, S, H, d, variant, timed))],
                                   timeout=120, capture_output=True, text=True)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    bad += 1
                    if not r.stdout.strip():
                        print(f"FAIL B={B} S={S} H={H} d={d} var={variant}: rc={r.returncode} {r.stderr.strip()[-400:]}", flush=True)
            except subprocess.TimeoutExpired:
                bad += 1
                print(f"HANG B={B} S={S} H={H} d={d} var={variant}", flush=True)
    print("failures:", bad)
