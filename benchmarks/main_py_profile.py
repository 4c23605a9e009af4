#!/usr/bin/env python
"""Wall-clock split of the warm configs[4] run of predict_soln on one GPU: VIP_PROFILE=1 python benchmarks/main_py_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VIP_PROFILE", "1")
import torch
import bench
t0 = time.perf_counter()
r = bench.bench_main_py("prof", 5000, 500, ["ResNetRS101-200x200", "GCViTSmall-224x224", "ResNetRS50-200x200"], 2, 0, 1, 0, None, device_batch=int(os.environ.get("BATCH", "512")))
print({k: r[k] for k in ("wall_s_cold", "wall_s_warm", "images_per_s_warm")})
