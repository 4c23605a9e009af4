import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from vipcup_b200 import ops
dev = torch.device('cuda:0'); N = 4096
rng = np.random.default_rng(0)
src = torch.from_numpy(rng.integers(0, 256, (N, 200, 200, 3), dtype=np.uint8)).to(dev)
crops = torch.tensor([[10, 12, 170, 176]] * N, dtype=torch.int32, device=dev)
q = torch.from_numpy(rng.integers(65, 100, N).astype(np.int32)).to(dev)
fl = torch.from_numpy(rng.integers(0, 4, N).astype(np.uint8)).to(dev)
out = torch.empty((N, 224, 224, 3), dtype=torch.float32, device=dev)
f = lambda: ops.preprocess(src, (224, 224), crops, q, fl, out=out)
for _ in range(3): f()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print('VIP_PRE_SHFL', os.environ.get('VIP_PRE_SHFL', '1'), 'ms min %.3f med %.3f' % (min(ts), sorted(ts)[3]), 'sum', float(out.double().sum()))
