#!/usr/bin/env python
"""Error decomposition aid (GPU box): per-stage and logit error of the B200 backbones against the fp32 oracle for several
weight seeds.  python tests/tools/diag_models.py [variants...]   (VIP_TWO_PLANE=0 shows the single-plane residual stream)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gcvit as G, preprocess as P, resnet_rs as R  # noqa: E402
from vipcup_b200.models import GCViT, ResNetRS  # noqa: E402

dev = torch.device("cuda:0")


def report(name, ref_taps, taps, Wk, Wb, verbose):
    if verbose:
        for k in ref_taps:
            a, b = taps[k].float().cpu().numpy(), ref_taps[k]
            print(f"  {k:8s} scale {np.abs(b).mean():.3f} rms err {np.sqrt(((a - b) ** 2).mean()):.4e} max {np.abs(a - b).max():.3e}")
    lg, lr = taps["feat"].cpu().numpy() @ Wk + Wb, ref_taps["feat"] @ Wk + Wb
    d = lg - lr
    print(f"  {name}: logit std over images {lr.std(0).round(3)}  max |logit err| {np.abs(d).max():.3e}  "
          f"common-mode part {np.abs(d.mean(0)).max():.3e}  image-dependent part {np.abs(d - d.mean(0)).max():.3e}", flush=True)


n = 8
which = sys.argv[1:] or ["rs50", "tiny", "small"]
x200 = np.stack([P.decode_to_float(P.synth_image(i), 200, 200) for i in range(n)])
x224 = np.stack([P.decode_to_float(P.synth_image(i), 224, 224) for i in range(n)])
print("two-plane residual stream:", os.environ.get("VIP_TWO_PLANE", "1"))
if "rs50" in which:
    for seed in (3, 4):
        W = R.random_weights(50, 2, seed=seed); rt = {}; R.forward(x200, W, 50, taps=rt)
        m = ResNetRS(50, classes=2, device=dev).load_weights(W); t = {}; m(torch.from_numpy(x200).to(dev), taps=t)
        report(f"rs50 seed {seed}", rt, t, W["predictions/kernel"], W["predictions/bias"], seed == 3)
for v in ("xxtiny", "tiny", "small"):
    if v not in which:
        continue
    for seed in (1, 2, 3, 5, 7):
        W = G.random_weights(v, 2, seed=seed); rt = {}; G.forward(x224, W, v, taps=rt)
        m = GCViT(v, num_classes=2, device=dev).load_weights(W); t = {}; m(torch.from_numpy(x224).to(dev), taps=t)
        report(f"gcvit-{v} seed {seed}", rt, t, W["head/kernel"], W["head/bias"], seed == 5)
