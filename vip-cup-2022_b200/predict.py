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
    if n == 0:                       # a rank whose shard is empty (more ranks than images)
        return np.zeros((0, 1), np.float32)
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
    """All TTA passes of one fold, eagerly, batch by batch (``model.predict(dtest, steps)``, main.py:109): returns float32
    [tta * n_local, k] on the host (one D2H copy).  ``predict_soln`` uses the graph-captured :class:`GraphedModel` path
    instead; this stays as the plain operator seam."""
    outs = []
    for _ in range(max(int(tta), 1)):
        for batch in dtest:
            outs.append(model(batch))
    if not outs:
        return np.zeros((0, 1), np.float32)
    return torch.cat(outs, dim=0).float().cpu().numpy()


class SharedInput:
    """Static device buffers one batch of decoded images lands in; every GraphedModel of the same batch size reads them."""

    def __init__(self, batch, src_hw, device):
        self.batch, self.src_hw = int(batch), tuple(src_hw)
        self.src = torch.zeros((self.batch, *self.src_hw, 3), dtype=torch.uint8, device=device)
        self.flags = torch.zeros((self.batch,), dtype=torch.uint8, device=device)


class GraphedModel:
    """One fold of one model at a fixed batch size: decoded uint8 batch -> fused preprocessing to the model's resolution
    (dataset/dataset.py:31-37 + the flip / gray flags of dataset/augment.py:115-120,142-146) -> forward -> probabilities,
    captured ONCE as a CUDA graph (~250 kernel launches per replay instead of ~250 ctypes calls per batch).  Ragged tail
    batches and batches of another source size run the same code eagerly."""

    def __init__(self, model, dim, shared: SharedInput, use_graph=True):
        from . import ops

        self.model, self.dim, self.shared, self._ops = model, tuple(int(v) for v in dim), shared, ops
        self.graph, self.probs = None, None
        if use_graph:
            dev = shared.src.device
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                self.forward(shared.src, shared.flags)          # warm-up: lazy kernel attributes, allocator pools
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=s):
                    self.probs = self.forward(shared.src, shared.flags)
            torch.cuda.current_stream(dev).wait_stream(s)

    def forward(self, src, flags):
        return self.model(self._ops.preprocess(src, self.dim, None, None, flags, out_dtype=torch.bfloat16))

    def __call__(self, src, flags):
        """src uint8 [b,Hs,Ws,3] on the device, flags uint8 [b] or None -> float32 [b,k] (valid until the next call)."""
        sh = self.shared
        n = src.shape[0]
        if (self.graph is not None and n < sh.batch and 4 * n >= sh.batch and tuple(src.shape[1:]) == tuple(sh.src.shape[1:])):
            # a tail that is at least a quarter of the graph's batch: run the graph on a padded batch (the unused slots keep
            # the previous batch's images; per-image results do not depend on their neighbours) and keep the first n rows
            sh.src[:n].copy_(src, non_blocking=True)
            if flags is None:
                sh.flags.zero_()
            else:
                sh.flags[:n].copy_(flags, non_blocking=True)
            self.graph.replay()
            return self.probs[:n].clone()
        if self.graph is not None and tuple(src.shape) == tuple(sh.src.shape):
            if src.data_ptr() != sh.src.data_ptr():
                sh.src.copy_(src, non_blocking=True)
            if flags is None:
                sh.flags.zero_()
            elif flags.data_ptr() != sh.flags.data_ptr():
                sh.flags.copy_(flags, non_blocking=True)
            self.graph.replay()
            return self.probs
        return self.forward(src, flags)


def _load_model(model_name, model_path, dim, device):
    W, meta = registry.load_checkpoint(model_path)
    head_key = next((k for k in W if k in ("predictions/kernel", "head/kernel", "head/fc/kernel")
                     or k.endswith(("/predictions/kernel", "/head/kernel", "/head/fc/kernel"))), None)
    if head_key is None:
        raise ValueError(f"{model_path}: no classifier kernel (predictions/kernel, head/kernel, head/fc/kernel) in the checkpoint")
    head_k = W[head_key].shape[1]
    model = registry.create_model(model_name, dim, num_classes=head_k,
                                  head_act=meta["head_act"] or ("sigmoid" if head_k == 1 else "softmax"), device=device)
    return model.load_weights(registry.resolve_weight_names(W, list(model.weight_shapes())))


def predict_device(CFG, local_paths, tta, verbose=False, runner_cache=None, index_offset=0):
    """Device half of ``predict_soln`` for this rank's shard: every fold of every registry entry over ``local_paths``.
    Returns (list over (model, fold) of float32 [tta * n_local, k] arrays, pass-major like ``model.predict`` on the
    repeated dataset (main.py:109-111), fold counts per model).

    Loop order: the reference runs model-major (decode the whole set once per model); here the decoded batch is the outer
    loop and every fold of every model consumes it while the thread pool decodes the next ones -- each JPEG is decoded once
    per TTA pass whatever the ensemble size.  Per-image results do not depend on the order or on the batch an image is in
    (integer-atomic statistics, see csrc/stats.cuh), so the outputs are the same.

    ``runner_cache`` (a dict) keeps the loaded, graph-captured models across calls (a long-lived service / bench.py's
    steady-state leg); by default every call loads its checkpoints like main.py:106-107 does."""
    dev = torch.device("cuda", torch.cuda.current_device())
    n_local = len(local_paths)
    entries, fold_counts = [], []
    for model_idx, (model_paths, dim, idx) in enumerate(CFG.ckpt_cfg):
        model_name = model_name_of(model_paths)
        bs = 8 * registry.NAME2BS.get(model_name, 16)                 # main.py:85
        bs = int(os.environ.get("VIP_DEVICE_BATCH", bs))              # throughput knob only: results are batch-independent
        if verbose:
            print(f"> MODEL({model_idx + 1}/{len(CFG.ckpt_cfg)}): {model_name} | DIM: {dim}")
            print("> BATCH SIZE : ", bs)
        for model_path in sorted(model_paths):
            entries.append(dict(name=model_name, dim=tuple(dim), bs=bs, path=model_path))
        fold_counts.append(len(model_paths))
    if n_local == 0:
        return [np.zeros((0, 1), np.float32) for _ in entries], fold_counts
    if len({e["bs"] for e in entries}) == 1:
        # equal batches instead of full ones plus a ragged tail (results do not depend on the batch an image is in): the
        # tail would run eagerly, ~450 library calls per model instead of one graph replay
        bs = entries[0]["bs"]
        bs = -(-n_local // -(-n_local // bs))
        for e in entries:
            e["bs"] = bs
    chunk = max(e["bs"] for e in entries)
    CFG.batch_size, CFG.img_size = chunk, entries[0]["dim"]
    ds = build_dataset(local_paths, labels=None, augment=tta > 1, repeat=True, cache=False, shuffle=False,
                       batch_size=chunk, drop_remainder=False, CFG=CFG, device=dev, index_offset=index_offset)
    shared, outs = {}, []
    use_graph = os.environ.get("VIP_GRAPH", "1") != "0"
    prof = os.environ.get("VIP_PROFILE", "0") == "1"     # wall-clock breakdown of the host loop on stderr
    import time as _time
    tm = {"wait_files": 0.0, "stage": 0.0, "enqueue": 0.0, "drain": 0.0}
    for pass_idx in range(tta):
        t_prev = _time.perf_counter()
        for i0, imgs in ds.host_batches():
            t0 = _time.perf_counter()
            tm["wait_files"] += t0 - t_prev
            flags_h = ds.flags_for(i0, len(imgs), pass_idx) if tta > 1 else None
            src = ds.stage(imgs)
            tm["stage"] += _time.perf_counter() - t0
            t0 = _time.perf_counter()
            flags = None if flags_h is None else torch.from_numpy(np.ascontiguousarray(flags_h)).to(dev, non_blocking=True)
            for e in entries:
                if "runner" not in e:           # first batch: weights to the device, graph capture at this source size
                    graphed = src is not None and n_local >= e["bs"]
                    ckey = (e["path"], e["bs"], tuple(src.shape[1:3]) if graphed else None, use_graph)
                    if runner_cache is not None and ckey in runner_cache:
                        e["runner"] = runner_cache[ckey]
                    else:
                        model = _load_model(e["name"], e["path"], e["dim"], dev)
                        if graphed:
                            key = (e["bs"], tuple(src.shape[1:3]))
                            shared_map = runner_cache.setdefault("__shared__", {}) if runner_cache is not None else shared
                            if key not in shared_map:
                                shared_map[key] = SharedInput(e["bs"], key[1], dev)
                            e["runner"] = GraphedModel(model, e["dim"], shared_map[key], use_graph=use_graph)
                        else:
                            e["runner"] = GraphedModel(model, e["dim"], SharedInput(1, (8, 8), dev), use_graph=False)
                        if runner_cache is not None:
                            runner_cache[ckey] = e["runner"]
                    e["out"] = None
                r = e["runner"]
                if src is None:                  # mixed source sizes in this batch: per-size preprocessing, eager forward
                    probs = [r.model(ds.to_device(imgs, flags_h, e["dim"]))]
                else:
                    probs = []
                    for j0 in range(0, len(imgs), e["bs"]):
                        pj = r(src[j0: j0 + e["bs"]], None if flags is None else flags[j0: j0 + e["bs"]])
                        probs.append(pj.clone() if pj is r.probs else pj)
                pr = probs[0] if len(probs) == 1 else torch.cat(probs, 0)
                if e["out"] is None:
                    e["out"] = torch.empty((tta * n_local, pr.shape[1]), dtype=torch.float32, device=dev)
                    e["pos"] = 0
                e["out"][e["pos"]: e["pos"] + pr.shape[0]] = pr
                e["pos"] += pr.shape[0]
            t_prev = _time.perf_counter()
            tm["enqueue"] += t_prev - t0
    t0 = _time.perf_counter()
    for e in entries:
        outs.append(e["out"].cpu().numpy())
    tm["drain"] = _time.perf_counter() - t0
    ds.check_decode_errors()
    if prof:
        import sys as _sys

        print("predict_device wall split (s): " + ", ".join(f"{k} {v:.3f}" for k, v in tm.items()), file=_sys.stderr)
    return outs, fold_counts


def predict_soln(CFG, ensemble=False, strategy=None, predict_fn=None, runner_cache=None):
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

    if predict_fn is not None:
        raw, fold_counts = [], []
        for model_idx, (model_paths, dim, idx) in enumerate(CFG.ckpt_cfg):
            model_name = model_name_of(model_paths)
            CFG.batch_size = 8 * registry.NAME2BS.get(model_name, 16)   # main.py:85
            CFG.img_size = dim
            for model_path in sorted(model_paths):
                raw.append(predict_fn(model_name, model_path, dim, local_paths))
            fold_counts.append(len(model_paths))
    else:
        import time as _time

        _t0 = _time.perf_counter()
        raw, fold_counts = predict_device(CFG, local_paths, tta, verbose, runner_cache, index_offset=lo)
        if os.environ.get("VIP_PROFILE", "0") == "1":
            import sys as _sys

            print(f"predict_soln: setup + predict_device {_time.perf_counter() - _t0:.3f} s", file=_sys.stderr)
    local_rows = [aggregate_model(np.asarray(pred, np.float32), tta, n_local, CFG.agg)[:, 0] for pred in raw]

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
        if os.environ.get("VIP_SAVE_PROBS", "0") == "1" and getattr(CFG, "temp_save_dir", None):
            # the per-model prediction files the reference has commented out (main.py:124,129): P(synthetic), fold mean
            for (model_paths, _, _), folds in zip(CFG.ckpt_cfg, per_model):
                pd.DataFrame({"filename": test_names, "logit": getattr(np, CFG.agg)(folds, axis=0)[:, 0]}).to_csv(
                    os.path.join(CFG.temp_save_dir, model_name_of(model_paths) + "_pred.csv"), index=False)
        if ensemble:
            pred_df = ensemble_frame(test_csv, test_names, per_model, CFG.thr, CFG.agg)
            pred_df.to_csv(CFG.output_csv_path, index=False)
            if verbose:
                print("\n> FINAL PREDICTION SAVED TO ", CFG.output_csv_path)
                print(pred_df.head(2))
    return pred_df


class EnsemblePredictor:
    """Device-side half of the hot path for a fixed batch size: decoded uint8 images -> fused preprocessing (one kernel per
    distinct model resolution, dataset/dataset.py:31-37) -> every backbone forward -> head with the ensemble mean of
    P(synthetic) accumulated in float64 by the head kernel (main.py:113-114,142).  The whole sequence is captured in ONE
    CUDA graph, so a step is: H2D copy of the batch, graph replay, D2H copy of [B] probabilities.

    models: list of (model, (H, W)) already holding their weights.  ``flags`` (uint8 [B] VIP_FLAG_* bits) select the
    flip / gray test-time augmentation of dataset/augment.py:115-120,142-146."""

    def __init__(self, models, batch, src_hw=(200, 200), device=None, use_graph=True):
        from . import nn, ops

        self.models, self.batch, self.src_hw = list(models), int(batch), tuple(src_hw)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._nn, self._ops = nn, ops
        b, (hs, ws) = self.batch, self.src_hw
        self.src = torch.zeros((b, hs, ws, 3), dtype=torch.uint8, device=self.device)
        self.flags = torch.zeros((b,), dtype=torch.uint8, device=self.device)
        self.acc = torch.zeros((b,), dtype=torch.float64, device=self.device)      # ensemble mean of P(synthetic)
        self.probs = [None] * len(self.models)
        self.graph = None
        self.launches_per_step = None
        if use_graph:
            self._capture()

    def _step(self):
        nn, ops = self._nn, self._ops
        nn.zero_(self.acc)
        pre = {}
        w = 1.0 / len(self.models)
        for i, (model, dim) in enumerate(self.models):
            dim = tuple(int(v) for v in dim)
            if dim not in pre:
                pre[dim] = ops.preprocess(self.src, dim, None, None, self.flags, out_dtype=torch.bfloat16)
            self.probs[i] = model(pre[dim], acc=self.acc, acc_weight=w)

    def _capture(self):
        from . import _lib

        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._step()                      # warm-up: lazy kernel attributes, neutral vectors, allocator pools
            torch.cuda.synchronize(self.device)
            _lib.launch_count_reset()
            self._step()
            self.launches_per_step = _lib.launch_count()
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=s):
                self._step()
        torch.cuda.current_stream(self.device).wait_stream(s)

    def run(self):
        """One pass over the batch currently in ``self.src`` / ``self.flags``; results in ``self.acc`` / ``self.probs``."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.acc

    def predict_host(self, src_u8_host, out_host=None, flags_host=None):
        """Pinned host uint8 [B,Hs,Ws,3] -> pinned host float64 [B] ensemble P(synthetic) (copies on the current stream)."""
        self.src.copy_(src_u8_host, non_blocking=True)
        if flags_host is not None:
            self.flags.copy_(flags_host, non_blocking=True)
        self.run()
        if out_host is None:
            out_host = torch.empty((self.batch,), dtype=torch.float64).pin_memory()
        out_host.copy_(self.acc, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host

    def predict_host_many(self, batches, outs, flags=None):
        """A sequence of pinned host uint8 batches -> pinned host float64 [B] each, pipelined: the host-to-device copy of
        batch i + 1 runs on a copy stream into one of two staging buffers while the graph of batch i executes (what the
        reference gets from tf.data prefetch, dataset/dataset.py:100).  Returns after the last result has landed."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [torch.empty_like(self.src), torch.empty_like(self.src)]
            self._ev_h2d = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        cs, stage, ev_h2d, ev_free = self._copy_stream, self._stage, self._ev_h2d, self._ev_free
        n = len(batches)
        if n == 0:
            return outs
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            stage[0].copy_(batches[0], non_blocking=True)
            ev_h2d[0].record(cs)
        for i in range(n):
            if i + 1 < n:
                with torch.cuda.stream(cs):
                    if i >= 1:
                        cs.wait_event(ev_free[(i + 1) & 1])   # staging buffer consumed by step i - 1
                    stage[(i + 1) & 1].copy_(batches[i + 1], non_blocking=True)
                    ev_h2d[(i + 1) & 1].record(cs)
            main.wait_event(ev_h2d[i & 1])
            self.src.copy_(stage[i & 1], non_blocking=True)
            ev_free[i & 1].record(main)
            if flags is not None:
                self.flags.copy_(flags[i], non_blocking=True)
            self.run()
            outs[i].copy_(self.acc, non_blocking=True)
        main.synchronize()
        return outs
