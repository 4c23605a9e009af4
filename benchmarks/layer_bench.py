#!/usr/bin/env python
"""Per-layer timings of the contraction kernels on the shapes of SURVEY.md Appendix B (batch 256), CUDA-event timed in
a loop of back-to-back launches, printed next to the two rooflines of each layer: algorithmic HBM bytes / measured copy
bandwidth and algorithmic FLOPs / measured bf16 peak.  ``python benchmarks/layer_bench.py [--only conv] [--batch 256]``"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from vipcup_b200 import nn

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    try:
        pk = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        hbm, tf = pk["hbm_gbs"], pk["bf16_tflops"]
    except Exception:
        hbm, tf = 6650.0, 1590.0

    def rnd(*shape):
        return (torch.randn(*shape, device=dev) * 0.1).to(torch.bfloat16)

    cases = []
    # (name, kind, params)  conv: (H, C, Cout, k, stride)   gemm: (M, N, K, extras)
    for name, h, c, co, k, s in [("rs.stem2 3x3 32>32 @100", 100, 32, 32, 3, 1), ("rs.stem3 3x3 32>64 @100", 100, 32, 64, 3, 1),
                                 ("rs.stem4 3x3s2 64>64 @100", 100, 64, 64, 3, 2), ("rs.c2.conv2 3x3 64 @50", 50, 64, 64, 3, 1),
                                 ("rs.c3.conv2 3x3 128 @25", 25, 128, 128, 3, 1), ("rs.c3.b0 3x3s2 128 @50", 50, 128, 128, 3, 2),
                                 ("rs.c4.conv2 3x3 256 @13", 13, 256, 256, 3, 1), ("rs.c5.conv2 3x3 512 @7", 7, 512, 512, 3, 1),
                                 ("gc.down0 3x3s2 64>128 @56", 56, 64, 128, 3, 2), ("gc.down1 3x3s2 128>256 @28", 28, 128, 256, 3, 2)]:
        cases.append((name, "conv", (h, c, co, k, s)))
    for name, m, n, k, ex in [("rs.c2.conv1 256>64", B * 2500, 64, 256, ""), ("rs.c2.conv3 64>256 +gap", B * 2500, 256, 64, "gap2500"),
                              ("rs.c3.conv1 512>128", B * 625, 128, 512, ""), ("rs.c3.conv3 128>512 +gap", B * 625, 512, 128, "gap625"),
                              ("rs.c4.conv1 1024>256", B * 169, 256, 1024, ""), ("rs.c4.conv3 256>1024 +gap", B * 169, 1024, 256, "gap169"),
                              ("rs.c5.conv1 2048>512", B * 49, 512, 2048, ""), ("rs.c5.conv3 512>2048 +gap", B * 49, 2048, 512, "gap49"),
                              ("rs.c5.se1 2048>512", B, 512, 2048, "relu"), ("rs.c5.se2 512>2048", B, 2048, 512, "sigf32"),
                              ("gc.L0.qkv 64>192 ln", B * 3136, 192, 64, "ln"), ("gc.L0.proj 64>64 res", B * 3136, 64, 64, "res"),
                              ("gc.L0.fc1 64>192 ln gelu", B * 3136, 192, 64, "ln gelu"), ("gc.L0.fc2 192>64 res", B * 3136, 64, 192, "res"),
                              ("gc.L1.qkv 128>384 ln", B * 784, 384, 128, "ln"), ("gc.L1.fc1 128>384 ln gelu", B * 784, 384, 128, "ln gelu"),
                              ("gc.L1.fc2 384>128 res", B * 784, 128, 384, "res"),
                              ("gc.L2.qkv 256>768 ln", B * 196, 768, 256, "ln"), ("gc.L2.proj 256>256 res", B * 196, 256, 256, "res"),
                              ("gc.L2.fc1 256>768 ln gelu", B * 196, 768, 256, "ln gelu"), ("gc.L2.fc2 768>256 res", B * 196, 256, 768, "res"),
                              ("gc.L3.fc1 512>1536 ln gelu", B * 49, 1536, 512, "ln gelu"), ("gc.L3.fc2 1536>512 res", B * 49, 512, 1536, "res"),
                              ("ref 8192^3", 8192, 8192, 8192, "")]:
        cases.append((name, "gemm", (m, n, k, ex)))

    for name, h, ws, heads, glob in [("gc.L0.attn ws7 h2 @56", 56, 7, 2, False), ("gc.L1.attn ws7 h4 @28", 28, 7, 4, True),
                                     ("gc.L2.attn ws14 h8 @14", 14, 14, 8, False), ("gc.L2.attn ws14 h8 glob", 14, 14, 8, True),
                                     ("gc.L3.attn ws7 h16 @7", 7, 7, 16, False)]:
        cases.append((name, "attn", (h, ws, heads, glob)))

    print(f"{'layer':32s} {'us':>9s} {'GB/s':>8s} {'%hbm':>6s} {'TF/s':>8s} {'%tc':>6s}   bound-by-roofline us")
    for name, kind, prm in cases:
        if args.only and args.only not in name and args.only != kind:
            continue
        if kind == "attn":
            h, ws, heads, glob = prm
            c = heads * 32
            qkv = rnd(B * h * h, (2 if glob else 3) * c) * 5
            qg = rnd(B, ws * ws, c) * 5 if glob else None
            table = torch.randn(heads, (2 * ws - 1) ** 2, device=dev)
            fn = lambda: nn.window_attention(qkv, qg, table, B, h, h, c, ws, heads)
            flops = 4.0 * B * h * h * ws * ws * c
            byts = 2.0 * (qkv.numel() + B * h * h * c)
        elif kind == "conv":
            h, c, co, k, s = prm
            x = rnd(B, h, h, c)
            w = rnd(co, k * k * c)
            bias = torch.randn(co, device=dev)
            ho = (h + 2 - k) // s + 1
            fn = lambda: nn.conv2d(x, w, bias, ksize=k, stride=s, pad=1, act="relu")
            flops = 2.0 * B * ho * ho * co * k * k * c
            byts = 2.0 * (B * h * h * c + B * ho * ho * co + co * k * k * c)
        else:
            m, n, k, ex = prm
            a, w = rnd(m, k), rnd(n, k)
            bias = torch.randn(n, device=dev)
            kw = {}
            byts = 2.0 * (m * k + m * n + n * k)
            if "ln" in ex.split():
                kw.update(ln_stats=torch.tensor([0.1, 1.3], device=dev).repeat(m, 1), ln_colsum=torch.randn(n, device=dev))
            if "gelu" in ex.split():
                kw.update(act="gelu")
            if "relu" in ex.split():
                kw.update(act="relu")
            if "res" in ex.split():
                kw.update(residual=rnd(m, n), row_stats=torch.zeros(m, 3, dtype=torch.int64, device=dev))
                byts += 2.0 * m * n
            if ex.startswith("gap"):
                hw = int(ex[3:])
                kw.update(gap=torch.zeros(m // hw, n, dtype=torch.int64, device=dev), gap_rows=hw)
            out = torch.empty((m, n), dtype=torch.float32 if ex == "sigf32" else torch.bfloat16, device=dev)
            if ex == "sigf32":
                kw.update(act="sigmoid")
            fn = lambda: nn.gemm(a, w, bias=bias, out=out, **kw)
            flops = 2.0 * m * n * k
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        gbs, tfs = byts / us / 1e3, flops / us / 1e6
        print(f"{name:32s} {us:9.1f} {gbs:8.0f} {100 * gbs / hbm:6.1f} {tfs:8.1f} {100 * tfs / tf:6.1f}   "
              f"hbm {byts / hbm / 1e3:7.1f}  tc {flops / tf / 1e6:7.1f}")


if __name__ == "__main__":
    main()
