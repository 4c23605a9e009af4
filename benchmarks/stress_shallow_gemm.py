#!/usr/bin/env python
"""Repeats the shallow-configuration contractions (K <= 256: two CTAs per SM, several tiles per CTA) into NaN-prefilled
outputs and checks every element against fp32 torch: a guard for ring / accumulator-stage hand-over races."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from vipcup_b200 import nn

    dev = torch.device("cuda:0")
    bad_total = 0
    for (m, n, k) in [(4096, 768, 256), (8192, 768, 256), (173056, 1024, 256), (40000, 512, 128), (50000, 256, 192), (30000, 384, 256),
                      (60000, 96, 96), (100000, 288, 96)]:
        g = torch.Generator(device="cpu").manual_seed(m + n + k)
        a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        b = (torch.randn(n, k, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        res = torch.randn(m, n, generator=g).to(torch.bfloat16).to(dev)
        ref = a.float() @ b.float().t()
        tol = 2e-2 * max(1.0, ref.abs().max().item())
        for trial in range(6):
            out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=dev)
            if trial % 2 == 0:
                nn.gemm(a, b, out=out)
                r = ref
            else:
                nn.gemm(a, b, residual=res, act="relu", out=out)
                r = torch.relu(ref) + res.float()
            torch.cuda.synchronize()
            nb = int((~((out.float() - r).abs() <= tol + 2e-2 * r.abs())).sum())
            bad_total += nb
            if nb:
                print(f"{m}x{n}x{k} trial {trial}: {nb} bad elements")
        print(f"{m}x{n}x{k}: done")
    print("TOTAL BAD", bad_total)
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
