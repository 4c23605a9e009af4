#!/usr/bin/env python
"""One contraction launch for ncu / timing: python benchmarks/one_gemm.py M N K [extras: ln gelu relu res lo se] | conv H C Cout stride
(res = residual + row statistics; lo = two-plane residual stream on top of res)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipcup_b200 import nn
dev = torch.device("cuda:0")
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
a = sys.argv[1:]
if a[0] == "conv":
    h, c, co, s = map(int, a[1:5]); B = 256
    x, w, bias = rnd(B, h, h, c), rnd(co, 9 * c), torch.randn(co, device=dev)
    fn = lambda: nn.conv2d(x, w, bias, ksize=3, stride=s, pad=1, act="relu")
else:
    m, n, k = map(int, a[:3]); ex = a[3:]
    A, w, bias = rnd(m, k), rnd(n, k), torch.randn(n, device=dev)
    kw = {}
    if "ln" in ex: kw.update(ln_stats=torch.tensor([0.1, 1.3], device=dev).repeat(m, 1), ln_colsum=torch.randn(n, device=dev))
    if "gelu" in ex: kw.update(act="gelu")
    if "relu" in ex: kw.update(act="relu")
    if "res" in ex: kw.update(residual=rnd(m, n), row_stats=torch.zeros(m, 3, dtype=torch.int64, device=dev))
    if "se" in ex: kw.update(residual=rnd(m, n), act="relu", row_gate=torch.rand((m + 168) // 169, n, device=dev), gate_rows=169)
    if "lo" in ex: kw.update(residual_lo=nn.lo_plane(m, n, dev).zero_(), out_lo=nn.lo_plane(m, n, dev))
    out = torch.empty((m, n), dtype=torch.bfloat16, device=dev)
    fn = lambda: nn.gemm(A, w, bias=bias, out=out, **kw)
for _ in range(3): fn()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("us", " ".join(a), "dbg", os.environ.get("VIP_GEMM_DEBUG", "0"), "min %.1f med %.1f" % (min(ts), sorted(ts)[2]))
