#!/usr/bin/env python
"""LayerNorm launch loop: python benchmarks/one_layernorm.py M C [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from vipcup_b200 import nn

    M, C = int(sys.argv[1]), int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    x = torch.randn(M, C, device=dev).to(torch.bfloat16)
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    for _ in range(2):
        nn.layernorm(x, g, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        nn.layernorm(x, g, b)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"layernorm M={M} C={C}: {us:.1f} us  {4.0 * M * C / us / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
