"""CPU tests of the host-side mirror of the reference interface: aggregation epilogue vs the oracle restatement of
main.py:110-145, contiguous sharding + the single gather under a 2-rank gloo group, registry / config behaviour."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_preds(rng, n_models, folds, tta, n, k, pad=0):
    out = []
    for _ in range(n_models):
        fl = []
        for _ in range(folds):
            p = rng.random((tta * n + pad, k)).astype(np.float32)
            if k > 1:
                p = p / p.sum(1, keepdims=True)
            fl.append(p)
        out.append(fl)
    return out


@pytest.mark.parametrize("k,tta,folds", [(1, 1, 1), (2, 1, 1), (2, 2, 3), (1, 4, 2)])
def test_epilogue_matches_oracle_restatement(k, tta, folds):
    from oracle.predict import epilogue
    from vipcup_b200.predict import aggregate_model, ensemble_frame

    rng = np.random.default_rng(7 + k + tta)
    n = 41
    names = [f"{i:05d}.jpg" for i in rng.permutation(n)]          # unsorted on purpose
    test_csv = pd.DataFrame({"filename": names})
    preds = _fake_preds(rng, 3, folds, tta, n, k, pad=5)           # trailing wrap-around rows (main.py:109-110)
    ref = epilogue(test_csv, preds, tta, thr=0.487)
    per_model = [[aggregate_model(p, tta, n) for p in fl] for fl in preds]
    got = ensemble_frame(test_csv, test_csv.filename.values, per_model, 0.487)
    pd.testing.assert_frame_equal(got, ref)
    assert list(got.columns) == ["filename", "logit"] and got.filename.is_monotonic_increasing
    assert set(np.unique(got.logit)) <= {0.0, 1.0}


def test_config_and_registry():
    from vipcup_b200 import registry
    from vipcup_b200.config import Config, cfg2dict, dict2cfg

    c = Config({"a": 1})
    c.b = 2
    assert cfg2dict(c) == {"a": 1, "b": 2}
    assert dict2cfg({"class_labels": [0, 1], "class_names": ["real", "syn"]}).label2name == {0: "real", 1: "syn"}
    assert registry.arch_of("ResNetRS50-200x200") == "ResNetRS50"
    assert 8 * registry.NAME2BS.get("ResNetRS50-200x200", 16) == 128      # main.py:85
    assert {"ResNetRS50", "ResNetRS101", "GCViTTiny", "GCViTSmall", "convnext_tiny_in22k"} <= set(registry.supported_archs())
    with pytest.raises(ValueError):
        registry.create_model("HorNetBase-200x200", [200, 200], device="cpu")   # named in NAME2BS, not in ckpts.json


def test_scan_checkpoints_errors(tmp_path):
    import json

    from vipcup_b200 import registry

    (tmp_path / "ckpts.json").write_text(json.dumps([["ResNetRS50-200x200", [200, 200], 0]]))
    with pytest.raises(ValueError, match="no model found"):
        registry.scan_checkpoints(str(tmp_path), str(tmp_path / "ckpts.json"))
    d = tmp_path / "ResNetRS50-200x200" / "ckpt"
    d.mkdir(parents=True)
    np.savez(d / "fold0.npz", x=np.zeros(1))
    out = registry.scan_checkpoints(str(tmp_path), str(tmp_path / "ckpts.json"))
    assert out[0][1] == [200, 200] and out[0][0][0].endswith("fold0.npz")


def test_philox_known_answers():
    """Philox-4x32-10 against the known-answer vectors of the Random123 distribution (kat_vectors): the generator family
    of tf.random.uniform (dataset/augment.py:11-19), used here as a pure function of (image index, TTA pass)."""
    from vipcup_b200.dataset import philox4x32_10, uniform_from_bits

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in philox4x32_10([ctr], key)[0]) == out
    u = uniform_from_bits(np.array([0, 0xFFFFFFFF, 0x80000000], np.uint32))
    assert u[0] == 0.0 and u[1] == np.float32(1.0) - np.float32(2.0 ** -23) and u[2] == 0.5


def test_augment_flags_statistics_and_order_independence():
    from vipcup_b200.dataset import draw_augment_flags

    f = draw_augment_flags(np.arange(200000), 0, 42)
    assert abs((f & 1).mean() - 0.8 * 0.5) < 0.01 and abs(((f >> 1) & 1).mean() - 0.4) < 0.01
    assert abs(((f >> 2) & 1).mean() - 0.8 * 0.3) < 0.01 and abs((f == 0).mean() - (0.2 + 0.8 * 0.25 * 0.7)) < 0.01
    # counter-based: the decision of image i in pass p does not depend on which images are asked for together with it
    idx = np.array([7, 123456, 3, 99999])
    assert np.array_equal(draw_augment_flags(idx, 0, 42), f[idx])
    assert np.array_equal(draw_augment_flags(np.arange(1000, 2000), 0, 42), f[1000:2000])
    g = draw_augment_flags(np.arange(200000), 1, 42)
    assert 0.1 < (f == g).mean() < 0.3                      # another pass: independent decisions (sum p_k^2 = 0.19)
    assert not np.array_equal(draw_augment_flags(np.arange(1000), 0, 43), f[:1000])


def _worker(rank, world, port, tmp, n, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from vipcup_b200.config import Config
    from vipcup_b200.device import ShardStrategy
    from vipcup_b200.predict import predict_soln

    dist.init_process_group("gloo", rank=rank, world_size=world)
    CFG = _cfg(tmp, n)
    CFG.output_csv_path = os.path.join(tmp, f"out_w{world}.csv")
    predict_soln(CFG, ensemble=True, strategy=ShardStrategy(rank, world, rank, "gloo"), predict_fn=_fake_predict)
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def _fake_predict(model_name, model_path, dim, paths):
    """Deterministic per-image 'probabilities' from the file name: sharding must not change them."""
    seed = sum(map(ord, model_name + os.path.basename(model_path)))
    out = []
    for p in paths:
        r = np.random.default_rng(seed + int(os.path.basename(p)[:5]))
        v = r.random(2).astype(np.float32)
        out.append(v / v.sum())
    return np.stack(out) if out else np.zeros((0, 2), np.float32)


def _cfg(tmp, n):
    from vipcup_b200.config import Config

    csv = os.path.join(tmp, "input.csv")
    if not os.path.exists(csv):
        pd.DataFrame({"filename": [f"{i:05d}.jpg" for i in range(n)]}).to_csv(csv, index=False)
    c = Config({})
    c.test_csv, c.infer_path, c.verbose, c.debug, c.tta = csv, tmp, 0, 0, 1
    c.agg, c.thr, c.seed = "mean", 0.487, 42
    c.ckpt_cfg = [[[os.path.join(tmp, "ResNetRS50-200x200", "ckpt", "a.npz"),
                    os.path.join(tmp, "ResNetRS50-200x200", "ckpt", "b.npz")], [200, 200], 0],
                  [[os.path.join(tmp, "GCViTTiny-224x224", "ckpt", "a.npz")], [224, 224], 0]]
    return c


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from vipcup_b200.device import ShardStrategy
    from vipcup_b200.predict import predict_soln

    tmp, n = str(tmp_path), 53     # odd N: the last rank's shard is shorter and padded in the gather
    CFG = _cfg(tmp, n)
    CFG.output_csv_path = os.path.join(tmp, "out_w1.csv")
    single = predict_soln(CFG, ensemble=True, strategy=ShardStrategy(), predict_fn=_fake_predict)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, n, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(240) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    two = pd.read_csv(os.path.join(tmp, "out_w2.csv"))
    pd.testing.assert_frame_equal(two, pd.read_csv(os.path.join(tmp, "out_w1.csv")))
    assert len(two) == n and single is not None
    s = ShardStrategy(1, 2)
    assert s.shard_bounds(53) == (27, 53, 27) and ShardStrategy(0, 2).shard_bounds(53) == (0, 27, 27)
