import json, torch, sys
sys.path.insert(0, '.')
import bench
r = bench.bench_jpeg_decode(torch.device("cuda:0"), 10, 6465.2)
print("kernels_ms %.3f decode_batch_ms %.3f exact %s" % (r["kernels_ms"], r["decode_batch_ms"], r["bit_exact_vs_libjpeg_turbo"]))
