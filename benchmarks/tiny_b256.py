#!/usr/bin/env python
"""BASELINE.json configs[2] alone (GCViT-tiny 224x224 bf16 forward, batch 256, one CUDA graph): python benchmarks/tiny_b256.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
r = bench.bench_gcvit_tiny_b256(torch.device("cuda:0"), 20, 1418.2)
print("VIP_PDL", os.environ.get("VIP_PDL", "0"), "ms %.3f frac %.4f" % (r["ms_per_step"], r["roofline"]["frac"]))
