"""CPU ORACLE (test infrastructure) -- the aggregation epilogue of the reference's ``main.py:110-145`` restated with the
same numpy / pandas calls: TTA mean -> ``1 - p[:, 0:1]`` for multi-class heads -> fold mean -> per-model DataFrames ->
``groupby('filename').mean()`` (float64) -> ``> thr`` -> rows sorted by filename.  Parity status: this IS the reference's
host code path minus TensorFlow (pandas/numpy run here), checked on the pandas version of this image (3.0)."""
from __future__ import annotations

import numpy as np
import pandas as pd


def epilogue(test_csv: pd.DataFrame, preds_per_model, tta: int, thr: float = 0.487, agg: str = "mean") -> pd.DataFrame:
    """preds_per_model: list over models of lists over folds of float32 arrays [>= tta*N, k] as ``model.predict`` returns
    them (pass-major, possibly with wrap-around padding rows at the end, main.py:109-110)."""
    test_names = np.array(test_csv.filename.values)
    n = len(test_names)
    pred_dfs = []
    for folds in preds_per_model:
        preds = []
        for pred in folds:
            pred = pred[: tta * n, :]                                        # main.py:110
            pred = getattr(np, agg)(pred.reshape((tta, n, -1)), axis=0)      # main.py:111
            if pred.shape[1] > 1:
                pred = 1 - pred[:, 0:1]                                      # main.py:113-114
            preds.append(pred)
        preds = getattr(np, agg)(preds, axis=0)                              # main.py:121
        pred_df = pd.DataFrame(np.concatenate([test_names[:, None], preds], axis=1), columns=["filename", "logit"])
        pred_df = test_csv.merge(pred_df, on=["filename"], how="right").reset_index(drop=True)
        pred_dfs.append(pred_df)
    dfs = pd.concat(pred_dfs)
    dfs["logit"] = dfs["logit"].astype(np.float64)
    out = dfs.groupby("filename")[["logit"]].mean().reset_index()           # main.py:142
    out["logit"] = (out.logit > thr) * 1.0                                   # main.py:143
    return out
