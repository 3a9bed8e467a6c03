#!/usr/bin/env python
"""Per-kernel micro-benchmark of the C-ABI ops at the shapes of BASELINE.json's configs.

    python tools/opbench.py [--only gemm|attn|ln|misc] [--arch base|large] [--iters 10] [--json out.json]

Every op is timed alone with CUDA events on the launching stream; a 256 MB buffer is rewritten between
iterations so nothing is served from L2 (burst peak applies: kernel timed in isolation).  Prints one
line per (op, shape): microseconds, TFLOP/s or GB/s, fraction of MEASURED_PEAKS.json.  This is a
development tool (also the command the ncu --set full captures under profiles/ are taken from), not the
benchmark of record -- that is bench.py.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "cross-scale-mae_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

from csmae_b200 import _native as nat  # noqa: E402

bf16, f32 = torch.bfloat16, torch.float32


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["hbm_gbs"]
    return 1590.0, 6650.0


SINGLE = False


def timeit(fn, iters, flush):
    if SINGLE:                      # one launch per op: the mode the ncu captures use
        flush.add_(1.0)
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--arch", default="base")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default="")
    ap.add_argument("--single", action="store_true", help="exactly one launch per op (for ncu)")
    ap.add_argument("--stream-k", action="store_true", help="register the stream-K workspace of the GEMMs")
    ap.add_argument("--cublas", action="store_true",
                    help="also time torch.matmul (cuBLAS, no epilogue) on every GEMM shape: the library yardstick")
    args = ap.parse_args()
    global SINGLE
    SINGLE = args.single
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    nsm = nat.sm_count(dev)
    if args.stream_k:
        nat.enable_gemm_stream_k(dev, True)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    tf_peak, bw_peak = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s, dt=bf16, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(dt)
    out = []

    def report(kind, name, shape, us, flops=None, nbytes=None):
        rec = {"op": name, "shape": shape, "us": round(us, 2)}
        if SINGLE:
            flops = nbytes = None
        if flops is not None:
            rec["tflops"] = round(flops / us / 1e6, 1)
            rec["frac_of_measured_burst"] = round(rec["tflops"] / tf_peak, 3)
        if nbytes is not None:
            rec["gbs"] = round(nbytes / us / 1e3, 1)
            rec["frac_of_measured_hbm"] = round(rec["gbs"] / bw_peak, 3)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    if args.arch == "base":
        NB, D, He = 128, 768, 12
    else:
        NB, D, He = 64, 1024, 16
    Dd, Hd, Se, Sd, P = 512, 16, 50, 197, 768
    Me, Md = NB * Se, NB * Sd

    if not args.only or "gemm" in args.only:
        fwd = [("enc.qkv", Me, 3 * D, D, nat.EPI_BF16), ("enc.proj", Me, D, D, nat.EPI_RESID),
               ("enc.fc1", Me, 4 * D, D, nat.EPI_GELU), ("enc.fc2", Me, D, 4 * D, nat.EPI_RESID),
               ("dec.qkv", Md, 3 * Dd, Dd, nat.EPI_BF16), ("dec.proj", Md, Dd, Dd, nat.EPI_RESID),
               ("dec.fc1", Md, 4 * Dd, Dd, nat.EPI_GELU), ("dec.fc2", Md, Dd, 4 * Dd, nat.EPI_RESID),
               ("dec.pred", Md, P, Dd, nat.EPI_BF16), ("patch_embed", Me, D, P, nat.EPI_BF16)]
        for name, M, N, K, epi in fwd:
            x, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N, dt=f32)
            if epi == nat.EPI_RESID:
                o, aux = torch.empty(M, N, device=dev), rnd(M, N, dt=f32)
            elif epi == nat.EPI_GELU:
                o, aux = torch.empty(M, N, device=dev, dtype=bf16), torch.empty(M, N, device=dev, dtype=bf16)
            else:
                o, aux = torch.empty(M, N, device=dev, dtype=bf16), None
            us = timeit(lambda: nat.call("csm_linear_fwd", x, w, b, o, aux, M, N, K, epi), args.iters, flush)
            report("gemm", f"linear_fwd[{epi}] {name}", [M, N, K], us, flops=2.0 * M * N * K)
            if args.cublas:
                y = torch.empty(M, N, device=dev, dtype=bf16)
                us = timeit(lambda: torch.matmul(x, w.t(), out=y), args.iters, flush)
                report("gemm", f"cublas fwd {name}", [M, N, K], us, flops=2.0 * M * N * K)
        dg = [("enc.qkv", Me, 3 * D, D, nat.EPI_BF16), ("enc.proj", Me, D, D, nat.EPI_BF16),
              ("enc.fc1", Me, 4 * D, D, nat.EPI_BF16), ("enc.fc2", Me, D, 4 * D, nat.EPI_DGELU),
              ("dec.qkv", Md, 3 * Dd, Dd, nat.EPI_BF16), ("dec.proj", Md, Dd, Dd, nat.EPI_BF16),
              ("dec.fc1", Md, 4 * Dd, Dd, nat.EPI_BF16), ("dec.fc2", Md, Dd, 4 * Dd, nat.EPI_DGELU),
              ("dec.pred", Md, P, Dd, nat.EPI_BF16)]
        for name, M, N, K, epi in dg:
            dy, w = rnd(M, N), rnd(N, K, sc=N ** -0.5)
            dx = torch.empty(M, K, device=dev, dtype=bf16)
            aux = rnd(M, K) if epi == nat.EPI_DGELU else None
            us = timeit(lambda: nat.call("csm_linear_dgrad", dy, w, dx, aux, M, N, K, epi), args.iters, flush)
            report("gemm", f"linear_dgrad[{epi}] {name}", [M, N, K], us, flops=2.0 * M * N * K)
            if args.cublas:
                us = timeit(lambda: torch.matmul(dy, w, out=dx), args.iters, flush)
                report("gemm", f"cublas dgrad {name}", [M, N, K], us, flops=2.0 * M * N * K)
        for name, M, N, K, _ in dg:
            dy, x = rnd(M, N), rnd(M, K)
            dw = torch.zeros(N, K, device=dev)
            us = timeit(lambda: nat.call("csm_linear_wgrad", dy, x, dw, M, N, K, nsm), args.iters, flush)
            report("gemm", f"linear_wgrad {name}", [M, N, K], us, flops=2.0 * M * N * K)
            if args.cublas:
                dw16 = torch.empty(N, K, device=dev, dtype=bf16)
                us = timeit(lambda: torch.matmul(dy.t(), x, out=dw16), args.iters, flush)
                report("gemm", f"cublas wgrad(bf16 out) {name}", [M, N, K], us, flops=2.0 * M * N * K)

    if not args.only or "attn" in args.only:
        for name, S, H, d in [("enc", Se, He, D // He), ("dec", Sd, Hd, Dd // Hd)]:
            Dm = H * d
            qkv = rnd(NB * S, 3 * Dm)
            o = torch.empty(NB * S, Dm, device=dev, dtype=bf16)
            lse = torch.empty(NB * H * S, device=dev)
            us = timeit(lambda: nat.call("csm_attention_fwd", qkv, o, lse, NB, S, H, d), args.iters, flush)
            fl = 4.0 * NB * H * S * S * d
            report("attn", f"attention_fwd {name}", [NB, S, H, d], us, flops=fl,
                   nbytes=NB * S * Dm * 2 * 4 + NB * H * S * 4)
            do = rnd(NB * S, Dm)
            dqkv = torch.empty_like(qkv)
            delta = torch.empty_like(lse)
            dbias = torch.zeros(3 * Dm, device=dev)
            us = timeit(lambda: nat.call("csm_attention_bwd", qkv, o, do, lse, delta, dqkv, None, NB, S, H, d),
                        args.iters, flush)
            report("attn", f"attention_bwd {name}", [NB, S, H, d], us, flops=2.5 * fl,
                   nbytes=NB * S * Dm * 2 * 8 + NB * H * S * 8)

    if not args.only or "ln" in args.only:
        for name, rows, Dm in [("enc", Me, D), ("dec", Md, Dd)]:
            x = rnd(rows, Dm, dt=f32)
            gam, bet = rnd(Dm, dt=f32), rnd(Dm, dt=f32)
            o16 = torch.empty(rows, Dm, device=dev, dtype=bf16)
            mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
            us = timeit(lambda: nat.call("csm_layernorm_fwd", x, gam, bet, o16, None, mean, rstd, rows, Dm, 1e-6),
                        args.iters, flush)
            report("ln", f"layernorm_fwd {name}", [rows, Dm], us, nbytes=rows * Dm * 6 + rows * 8)
            dy = rnd(rows, Dm)
            dres = rnd(rows, Dm, dt=f32)
            dres16 = torch.empty(rows, Dm, device=dev, dtype=bf16)
            dg_, db_ = torch.zeros(Dm, device=dev), torch.zeros(Dm, device=dev)
            us = timeit(lambda: nat.call("csm_layernorm_bwd", dy, None, x, mean, rstd, gam, dres, dres, dres16, dg_, db_, dg_,
                                         rows, Dm, nsm), args.iters, flush)
            report("ln", f"layernorm_bwd {name}", [rows, Dm], us, nbytes=rows * Dm * (2 + 4 + 4 + 4 + 2) + rows * 8)
            for N in (Dm, 3 * Dm, 4 * Dm):
                dyw = rnd(rows, N)
                db2 = torch.zeros(N, device=dev)
                us = timeit(lambda: nat.call("csm_colsum_bf16", dyw, db2, rows, N, 0, nsm), args.iters, flush)
                report("ln", f"colsum {name}", [rows, N], us, nbytes=rows * N * 2)

    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
