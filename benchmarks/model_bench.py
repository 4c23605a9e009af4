#!/usr/bin/env python
"""Backbone forward throughput on one B200 (BASELINE.json configs[2]/[3] shapes): random-init weights, synthetic input
resident in HBM, CUDA-event timing.  ``python benchmarks/model_bench.py [--models GCViTTiny-224x224 ...] [--batch 256]``
Prints one JSON line per model: images/s, ms per batch, achieved TFLOP/s against the algorithmic FLOPs of BASELINE.md."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GFLOP_PER_IMAGE = {"ResNetRS50-200x200": 7.58, "ResNetRS101-200x200": 14.00, "GCViTTiny-224x224": 9.52,
                   "GCViTSmall-224x224": 17.06}


def main():
    import torch

    from vipcup_b200 import _lib, registry

    ap = argparse.ArgumentParser()
    ap.add_argument("--models", nargs="*", default=["ResNetRS50-200x200", "GCViTTiny-224x224"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--graph", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0
    for name in args.models:
        hw = [int(v) for v in name.rsplit("-", 1)[1].split("x")]
        model = registry.create_model(name, hw, num_classes=2, device=dev).init_random(0)
        x = torch.rand((args.batch, hw[0], hw[1], 3), device=dev).to(torch.bfloat16)
        for _ in range(3):
            out = model(x)
        torch.cuda.synchronize()
        _lib.launch_count_reset()
        model(x)
        launches = _lib.launch_count()
        run = lambda: model(x)
        if args.graph:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                model(x)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    out = model(x)
            torch.cuda.current_stream().wait_stream(s)
            run = g.replay
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        tf = GFLOP_PER_IMAGE.get(name, 0) * args.batch / ms  # GFLOP/ms = TFLOP/s
        print(json.dumps({"model": name, "batch": args.batch, "ms_per_batch": ms, "images_per_s": args.batch / ms * 1e3,
                          "tflops": tf, "frac_of_bf16_sustained": tf / peak, "kernel_launches_per_forward": int(launches),
                          "cuda_graph": bool(args.graph), "finite": bool(torch.isfinite(out).all())}))


if __name__ == "__main__":
    main()
