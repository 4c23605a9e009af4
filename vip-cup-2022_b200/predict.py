"""``predict_soln`` -- the per-model / per-fold / TTA / ensemble loop of the reference's ``main.py:58-149`` on the B200
path.  Differences that do not change results: ``model.predict(ds.repeat(), steps=float)`` and its wrap-around padding
(main.py:109-110) become exact passes over the N images; under WORLD_SIZE > 1 each rank predicts a contiguous shard and
one all_gather collects the [models, N] probabilities before rank 0 runs the pandas epilogue (main.py:121-145)
unchanged."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from . import registry
from .dataset import build_dataset


def model_name_of(model_paths):
    """main.py:70-71"""
    return os.path.basename(os.path.dirname(os.path.dirname(model_paths[0])))


def aggregate_model(pred_passes, tta, n, agg="mean"):
    """main.py:110-114 on a float32 [tta*n, k] array: TTA mean, multi-class -> binary P(synthetic)."""
    pred = pred_passes[: tta * n, :]
    pred = getattr(np, agg)(pred.reshape((tta, n, -1)), axis=0)
    if pred.shape[1] > 1:
        pred = 1 - pred[:, 0:1]
    return pred


def ensemble_frame(test_csv, test_names, per_model_preds, thr, agg="mean"):
    """main.py:121-145: per-model DataFrames -> concat -> groupby(filename).mean() -> threshold.  ``per_model_preds`` is a
    list over models of lists over folds of float32 [N,1] arrays."""
    pred_dfs = []
    for fold_preds in per_model_preds:
        preds = getattr(np, agg)(fold_preds, axis=0)
        names = np.array(test_names)
        pred_df = pd.DataFrame(np.concatenate([names[:, None], preds], axis=1), columns=["filename", "logit"])
        pred_df = test_csv.merge(pred_df, on=["filename"], how="right").reset_index(drop=True)
        pred_dfs.append(pred_df)
    dfs = pd.concat(pred_dfs)
    dfs["logit"] = dfs["logit"].astype(np.float64)     # object column of python floats -> float64 (pandas >= 2 needs it)
    out = dfs.groupby("filename")[["logit"]].mean().reset_index()
    out["logit"] = (out.logit > thr) * 1.0
    return out


def predict_model(model, dtest, tta):
    """All TTA passes of one fold: returns float32 [tta * n_local, k] on the host (one D2H copy)."""
    outs = []
    for _ in range(max(int(tta), 1)):
        for batch in dtest:
            outs.append(model(batch))
    if not outs:
        return np.zeros((0, 1), np.float32)
    return torch.cat(outs, dim=0).float().cpu().numpy()


def predict_soln(CFG, ensemble=False, strategy=None, predict_fn=None):
    """``predict_fn(model_name, model_path, dim, local_paths) -> float32 [tta*n_local, k]`` can replace the device path
    (used by the CPU tests of the host logic); by default checkpoints are loaded and run on the current GPU."""
    from .device import ShardStrategy

    strategy = strategy or ShardStrategy()
    verbose = getattr(CFG, "verbose", 1) and strategy.rank == 0
    if verbose:
        print("=" * 35 + "\n### INFERENCE ###\n" + "=" * 35)
    test_csv = pd.read_csv(CFG.test_csv)
    test_names = test_csv.filename.values
    test_paths = [os.path.join(CFG.infer_path, name) for name in test_names]
    if getattr(CFG, "debug", 0):
        test_paths, test_names = test_paths[:100], test_names[:100]
        test_csv = test_csv.iloc[:100]
    n = len(test_paths)
    lo, hi, _ = strategy.shard_bounds(n)
    local_paths = test_paths[lo:hi]
    n_local = hi - lo
    tta = max(int(CFG.tta), 1)

    local_rows, fold_counts = [], []
    for model_idx, (model_paths, dim, idx) in enumerate(CFG.ckpt_cfg):
        model_name = model_name_of(model_paths)
        if verbose:
            print(f"> MODEL({model_idx + 1}/{len(CFG.ckpt_cfg)}): {model_name} | DIM: {dim}")
        CFG.batch_size = 8 * registry.NAME2BS.get(model_name, 16)   # main.py:85
        if verbose:
            print("> BATCH SIZE : ", CFG.batch_size)
        CFG.img_size = dim
        dtest = None
        for model_path in sorted(model_paths):
            if predict_fn is not None:
                pred = predict_fn(model_name, model_path, dim, local_paths)
            else:
                if dtest is None:
                    dtest = build_dataset(local_paths, labels=None, augment=CFG.tta > 1, repeat=True, cache=False,
                                          shuffle=False, batch_size=CFG.batch_size, drop_remainder=False, CFG=CFG)
                W, meta = registry.load_checkpoint(model_path)
                head_k = W["predictions/kernel" if "predictions/kernel" in W else "head/kernel"].shape[1]
                model = registry.create_model(model_name, dim, num_classes=head_k,
                                              head_act=meta["head_act"] or ("sigmoid" if head_k == 1 else "softmax"))
                model.load_weights(W)
                pred = predict_model(model, dtest, tta)
                del model
            local_rows.append(aggregate_model(np.asarray(pred, np.float32), tta, n_local, CFG.agg)[:, 0])
        fold_counts.append(len(model_paths))

    # one exchange step: [sum(folds), n_local] float32 per rank -> [sum(folds), N] everywhere
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and strategy.backend != "gloo" \
        else torch.device("cpu")
    local = torch.from_numpy(np.stack(local_rows).astype(np.float32) if local_rows else np.zeros((0, n_local), np.float32))
    full = strategy.gather_rows(local.to(dev), n).cpu().numpy()

    pred_df = None
    if strategy.rank == 0:
        per_model, r = [], 0
        for fc in fold_counts:
            per_model.append([full[r + f][:, None] for f in range(fc)])
            r += fc
        if ensemble:
            pred_df = ensemble_frame(test_csv, test_names, per_model, CFG.thr, CFG.agg)
            pred_df.to_csv(CFG.output_csv_path, index=False)
            if verbose:
                print("\n> FINAL PREDICTION SAVED TO ", CFG.output_csv_path)
                print(pred_df.head(2))
    return pred_df
