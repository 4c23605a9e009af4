#!/usr/bin/env python
"""Device JPEG decode of 1024 synthetic 200x200 files for ncu / timing: python benchmarks/one_jpeg.py [n]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vipcup_b200 import jpeg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as d:
    bench._write_synth_jpegs(d, 128, 128)
    uniq = [open(os.path.join(d, f"{i:05d}.jpg"), "rb").read() for i in range(128)]
files = [uniq[i % 128] for i in range(n)]
descs = [jpeg.parse(f) for f in files]
for _ in range(3):
    b = jpeg.decode_batch(files, descs, device=dev)
torch.cuda.synchronize()
b.check()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b = jpeg.decode_batch(files, descs, device=dev); e1.record(); torch.cuda.synchronize()
print("decode_batch ms %.3f for %d files, largest %d bytes" % (e0.elapsed_time(e1), n, max(len(f) for f in files)))
