#!/usr/bin/env python3
"""Drop-in for the reference's ``main.py <input.csv> <output.csv>`` (main.py:151-235) on the B200 path.

Same positional arguments, same ``ckpts/ckpts.json`` registry next to this file, same constants (tta=1, agg='mean',
resize_method='bicubic', seed=42, thr=0.487), same output CSV (columns filename,logit; rows sorted by filename).
Run on one GPU with ``python main.py in.csv out.csv`` or on all GPUs of a node with
``python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 main.py in.csv out.csv``."""
import os
import sys
import time

cwd = os.path.dirname(os.path.abspath(sys.argv[0])) or "."
sys.path.insert(0, cwd)

from vipcup_b200 import registry  # noqa: E402
from vipcup_b200.config import Config  # noqa: E402
from vipcup_b200.dataset import seeding  # noqa: E402
from vipcup_b200.device import get_device  # noqa: E402
from vipcup_b200.predict import predict_soln  # noqa: E402

if __name__ == "__main__":
    input_csv_path = sys.argv[1]   # input csv
    output_csv_path = sys.argv[2]  # output csv
    model_dir = os.environ.get("VIP_MODEL_DIR", os.path.join(cwd, "ckpts"))
    ckpt_cfg = os.path.join(model_dir, "ckpts.json")

    debug = 0
    verbose = 1
    output = os.path.dirname(output_csv_path)
    infer_path = os.path.dirname(input_csv_path)  # directory of testset
    temp_save_dir = os.path.join(output, "temp")
    os.makedirs(temp_save_dir, exist_ok=True)
    tta = int(os.environ.get("VIP_TTA", "1"))  # number of tta (reference hard-wires 1, main.py:167)

    CFG = Config({})
    CFG.test_csv = input_csv_path
    CFG.output_csv_path = output_csv_path
    CFG.verbose = verbose
    CFG.model_dir = model_dir
    CFG.temp_save_dir = temp_save_dir
    CFG.ckpt_cfg = registry.scan_checkpoints(model_dir, ckpt_cfg)
    CFG.infer_path = infer_path
    CFG.debug = debug
    CFG.tta = tta
    CFG.output_dir = output

    strategy, device = get_device()
    CFG.device = device
    CFG.replicas = strategy.num_replicas_in_sync
    if strategy.rank == 0:
        print("\n> CHECKPOINTS: ")
        [print(i) for i in CFG.ckpt_cfg]
        print("> DEBUG MODE:", bool(CFG.debug))
        print(f"> REPLICAS: {CFG.replicas}")

    CFG.agg = "mean"
    CFG.resize_method = "bicubic"
    CFG.num_classes = 1
    CFG.seed = 42
    CFG.thr = 0.487
    seeding(CFG)

    start = time.time()
    predict_soln(CFG, ensemble=True, strategy=strategy)
    eta = (time.time() - start) / 60
    if strategy.rank == 0:
        print(f"\n> TIME TO INFER: {eta:0.2f} min")
    if strategy.world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
