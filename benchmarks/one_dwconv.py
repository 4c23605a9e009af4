#!/usr/bin/env python
"""One depthwise 3x3 (+GELU, + SE squeeze) launch loop for profiling: python benchmarks/one_dwconv.py B H C [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from vipcup_b200 import nn

    B, H, C = (int(v) for v in sys.argv[1:4])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    dev = torch.device("cuda:0")
    x = (torch.randn(B, H, H, C, device=dev) * 0.5).to(torch.bfloat16)
    w = torch.randn(3, 3, C, device=dev) * 0.2
    gap = torch.zeros(B, C, dtype=torch.int64, device=dev)
    for _ in range(2):
        nn.dwconv3x3(x, w, gelu=True, gap=gap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        nn.dwconv3x3(x, w, gelu=True, gap=gap)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    byts = 4.0 * B * H * H * C
    print(f"dwconv3x3 B={B} H={H} C={C}: {us:.1f} us  {byts / us / 1e3:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
