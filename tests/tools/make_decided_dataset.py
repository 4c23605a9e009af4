#!/usr/bin/env python
"""Synthetic dataset on which "identical predicted labels" is a meaningful, assertable statement.

Random-init ensembles put many images within the bf16 error of the 0.487 threshold, where a label
comparison is a coin toss.  This tool writes ``pool`` candidate JPEGs, runs the ORACLE pipeline (Pillow decode -> oracle
preprocess -> oracle backbones -> main.py:110-145 epilogue) over them and lists in input.csv the first ``n`` whose
ensemble probability is further than ``margin`` from the threshold -- ``margin`` being chosen above the measured error
of the B200 path (tests/test_main_gpu.py prints both).  Returns the oracle's per-model probabilities of the kept images."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def oracle_model_probs(models_dir, name, imgs, tta_flags=None):
    """float32 [len(imgs), k] (or [passes * len, k] with ``tta_flags`` = list of per-pass uint8 flag arrays)."""
    from oracle import convnext as C
    from oracle import efficientnet as E
    from oracle import gcvit as G
    from oracle import nfnet as NF
    from oracle import preprocess as P
    from oracle import resnest as RN
    from oracle import resnet_rs as R
    from vipcup_b200 import registry

    arch, hw = name.rsplit("-", 1)
    dim = int(hw.split("x")[0])
    W, meta = registry.load_checkpoint(os.path.join(models_dir, name, "ckpt", "fold0.npz"))
    passes = tta_flags if tta_flags is not None else [None]
    out = []
    for fl in passes:
        xs = []
        for i, im in enumerate(imgs):
            x = P.decode_to_float(im, dim, dim)
            if fl is not None:
                x = P.apply_flags(x, int(fl[i]))
            xs.append(x)
        x = np.stack(xs)
        act = meta["head_act"] or "softmax"
        if arch.startswith("ResNetRS"):
            out.append(R.forward(x, W, int(arch[len("ResNetRS"):]), head_act=act))
        elif arch.startswith("GCViT"):
            out.append(G.forward(x, W, arch[len("GCViT"):].lower(), head_act=act))
        elif arch.startswith("convnext_"):
            out.append(C.forward(x, W, arch.split("_")[1], head_act=act))
        elif arch.startswith("EfficientNet"):
            out.append(E.forward(x, W, {"EfficientNetV2T": "v2t", "EfficientNetV1B4": "v1b4"}[arch], head_act=act))
        elif arch == "ECA_NFNetL0":
            out.append(NF.forward(x, W, head_act=act))
        elif arch == "ResNest50":
            out.append(RN.forward(x, W, head_act=act))
        else:
            raise SystemExit(f"no oracle for {arch}")
    return np.concatenate(out, 0).astype(np.float32)


def main(data_dir, models_dir, names, n, margin, pool=None, thr=0.487):
    import make_synth_dataset
    from PIL import Image

    pool = pool or 2 * n
    make_synth_dataset.main(data_dir, pool)
    cand = pd.read_csv(os.path.join(data_dir, "input.csv"))
    imgs = [np.asarray(Image.open(os.path.join(data_dir, f)).convert("RGB")) for f in cand.filename]
    probs = {name: oracle_model_probs(models_dir, name, imgs) for name in names}
    p_syn = np.mean([(1 - p[:, 0]) if p.shape[1] > 1 else p[:, 0] for p in probs.values()], axis=0, dtype=np.float64)
    keep = np.flatnonzero(np.abs(p_syn - thr) > margin)[:n]
    if len(keep) < n:
        raise SystemExit(f"only {len(keep)} of {pool} candidates are further than {margin} from the threshold")
    cand.iloc[keep].to_csv(os.path.join(data_dir, "input.csv"), index=False)
    return cand.filename.values[keep], {k: v[keep] for k, v in probs.items()}, p_syn[keep], len(keep) / (keep[-1] + 1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[5:], int(sys.argv[3]), float(sys.argv[4]))
