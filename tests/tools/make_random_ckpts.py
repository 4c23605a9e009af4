#!/usr/bin/env python
"""Writes seeded random-init checkpoints (Keras names/layouts, .npz) + ckpts.json for the synthetic runs:
``python tests/tools/make_random_ckpts.py <model_dir> ResNetRS50-200x200 GCViTTiny-224x224 ...``
(checkpoints of the reference are unpublished, README.md:13; both the oracle and the CUDA path load these files)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def weights_for(model_name, num_classes, seed):
    from oracle import convnext as C
    from oracle import efficientnet as E
    from oracle import gcvit as G
    from oracle import nfnet as NF
    from oracle import resnest as RN
    from oracle import resnet_rs as R

    arch = model_name.rsplit("-", 1)[0]
    if arch.startswith("ResNetRS"):
        return R.random_weights(int(arch[len("ResNetRS"):]), num_classes, seed)
    if arch.startswith("GCViT"):
        return G.random_weights(arch[len("GCViT"):].lower(), num_classes, seed)
    if arch.startswith("convnext_"):
        return C.random_weights(arch.split("_")[1], num_classes, seed)
    if arch in ("EfficientNetV2T", "EfficientNetV1B4"):
        return E.random_weights({"EfficientNetV2T": "v2t", "EfficientNetV1B4": "v1b4"}[arch], num_classes, seed)
    if arch == "ECA_NFNetL0":
        return NF.random_weights(num_classes, seed)
    if arch == "ResNest50":
        return RN.random_weights(num_classes, seed)
    raise SystemExit(f"no random-init generator for {arch}")


def centred(model_name, W, n_cal=16):
    """Random-init backbones are almost image-independent: re-draw the head along the first
    principal component of the calibration features so that the logit margin has std 2 over the synthetic images and its
    median sits on the 0.487 threshold.  The read-out amplifies the backbone's bf16 error by the same factor as the
    signal, so tests select images whose oracle margin exceeds the measured error (make_decided_dataset.py)."""
    from oracle import gcvit as G
    from oracle import preprocess as P
    from oracle import resnet_rs as R

    arch, hw = model_name.rsplit("-", 1)
    dim = int(hw.split("x")[0])
    x = np.stack([P.decode_to_float(P.synth_image(i), dim, dim) for i in range(n_cal)])
    taps = {}
    if arch.startswith("ResNetRS"):
        R.forward(x, W, int(arch[len("ResNetRS"):]), taps=taps)
        W = R.calibrate_head(W, taps["feat"], seed=1, target_std=2.0, direction="pca")
        return R.center_head(W, taps["feat"])
    G.forward(x, W, arch[len("GCViT"):].lower(), taps=taps)
    W = R.calibrate_head(W, taps["feat"], "head/kernel", "head/bias", seed=1, target_std=2.0, direction="pca")
    return R.center_head(W, taps["feat"], head_kernel="head/kernel", head_bias="head/bias")


def main(model_dir, names, num_classes=2, folds=1, calibrate=True):
    entries = []
    for mi, name in enumerate(names):
        hw = [int(v) for v in name.rsplit("-", 1)[1].split("x")]
        d = os.path.join(model_dir, name, "ckpt")
        os.makedirs(d, exist_ok=True)
        for f in range(folds):
            W = weights_for(name, num_classes, seed=100 * mi + f)
            if calibrate:
                W = centred(name, W)
            np.savez(os.path.join(d, f"fold{f}.npz"), __num_classes__=np.int64(num_classes),
                     __head_act__=np.array("softmax" if num_classes > 1 else "sigmoid"), **W)
        entries.append([name, hw, 0])
    with open(os.path.join(model_dir, "ckpts.json"), "w") as f:
        json.dump(entries, f, indent=1)
    print("wrote", [e[0] for e in entries], "to", model_dir)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
