#!/usr/bin/env python
"""Fused MLP against the two contractions it replaces: python benchmarks/one_mlp.py M C HIDDEN [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from vipcup_b200 import nn

    M, C, HD = (int(v) for v in sys.argv[1:4])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    dev = torch.device("cuda:0")
    rnd = lambda *s: (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
    x, w1, w2 = rnd(M, C) * 10, rnd(HD, C), rnd(C, HD)
    xf = x.float()
    stats = torch.stack([xf.mean(1), 1.0 / torch.sqrt(xf.var(1, unbiased=False) + 1e-5)], 1).float().contiguous()
    del xf
    cs, b1, b2 = w1.float().sum(1), torch.randn(HD, device=dev), torch.randn(C, device=dev)
    rs = torch.zeros(M, 3, dtype=torch.int64, device=dev)
    ln_next = torch.empty(M, 2, device=dev)

    def fused():
        return nn.mlp_fused(x, stats, w1, cs, b1, w2, b2, ln_next=ln_next)

    def two():
        h = nn.gemm(x, w1, bias=b1, act="gelu", ln_stats=stats, ln_colsum=cs)
        return nn.gemm(h, w2, bias=b2, residual=x, row_stats=rs)

    for name, fn in (("fused", fused), ("two gemms", two)):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:10s} M={M} C={C} hidden={HD}: {us:8.1f} us   {4.0 * M * C * HD / us / 1e6:6.1f} TFLOP/s   "
              f"x in + y out {4.0 * M * C / us / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
