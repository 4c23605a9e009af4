#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on the B200 path, one JSON line on stdout (rank 0).

Workload at every N (weak scaling, no data-path collective: images are independent):
BASELINE.json configs[1] -- the augment pipeline on a 4096-image batch per GPU:
u8 [4096,200,200,3] -> random crop -> bicubic 224x224 -> /255 -> JPEG q in [65,100) -> flips -> f32 [4096,224,224,3].
A "step" is one pass of that path over one batch.

  value      images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the host-buffer C-ABI call (pinned host in/out, H2D + D2H inside the timed region)
  roofline   dominant kernel (preprocess_kernel): algorithmic bytes (722 112 B/image) / CUDA-event time / measured HBM peak
  cpu_baseline  the numpy oracle (oracle/preprocess.py) timed on this box's host cores (rank 0, N=1 only)

``--impl reference`` times the CPU restatement of the reference path on all host cores (the TensorFlow reference
itself cannot be installed offline) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMAGES = 4096
HS = WS = 200
HO = WO = 224
BYTES_PER_IMAGE = HS * WS * 3 + HO * WO * 3 * 4  # 722 112 algorithmic bytes
METRIC = "images/sec at 200x200 (preprocess+ensemble fwd)"
WORKLOAD = "configs[1]: augment pipeline (crop->bicubic 200->224->/255->JPEG q[65,100)->flip), 4096-image batch per GPU"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("preprocess_kernel_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock / throttle reasons of this rank's GPU during the timed region (NVML)."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _cpu_one(args):
    from oracle import preprocess as P

    i, crop, q, flag, backend = args
    src = P.synth_image(i % 64, HS, WS)
    return float(P.preprocess_one(src, HO, WO, crop, q, flag, jpeg_backend=backend).sum())


def cpu_baseline(n_sample: int, cores: int, backend: str):
    """Times the oracle on ``n_sample`` images of the same workload. Returns images/s."""
    from oracle import preprocess as P

    crops, q, flags = P.synth_decisions(N_IMAGES)
    jobs = [(i, crops[i], int(q[i]), int(flags[i]), backend) for i in range(n_sample)]
    for j in jobs[: min(8, n_sample)]:
        _cpu_one(j)  # warm caches / imports
    t0 = time.perf_counter()
    if cores == 1:
        for j in jobs:
            _cpu_one(j)
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=max(1, n_sample // (cores * 4)))
    dt = time.perf_counter() - t0
    return n_sample / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = 64 * max(1, min(cores, 32) // 4)  # bounded sample per step
    vals = []
    for s in range(args.warmup + args.steps):
        v = cpu_baseline(sample, cores, "pillow")
        if s >= args.warmup:
            vals.append(v)
    value = len(vals) / sum(1.0 / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (TensorFlow is not installable "
                   "offline): numpy bicubic + libjpeg-turbo (Pillow) JPEG round trip, fork pool over all host cores"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images of the 4096-image batch per step"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import numpy as np
    import torch

    from vipcup_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # synthetic inputs (generated with numpy only; the oracle module is not touched on the product path)
    rng = np.random.default_rng(1234 + rank)
    base = rng.integers(0, 256, (64, HS, WS, 3), dtype=np.uint8)
    # low-pass in x so that JPEG sees natural-ish statistics
    base = ((base.astype(np.uint16) + np.roll(base, 1, 2) + np.roll(base, 2, 2) + np.roll(base, 1, 1)) // 4).astype(np.uint8)
    src_h = torch.from_numpy(np.tile(base, (N_IMAGES // 64, 1, 1, 1))).pin_memory()
    side = rng.integers(160, 201, N_IMAGES)
    y0 = (rng.random(N_IMAGES) * (HS - side + 1)).astype(np.int64)
    x0 = (rng.random(N_IMAGES) * (WS - side + 1)).astype(np.int64)
    crops_h = torch.from_numpy(np.stack([y0, x0, side, side], 1).astype(np.int32)).pin_memory()
    q_h = torch.from_numpy(rng.integers(65, 100, N_IMAGES).astype(np.int32)).pin_memory()
    flags_h = torch.from_numpy((rng.integers(0, 2, N_IMAGES) + 2 * rng.integers(0, 2, N_IMAGES)).astype(np.uint8)).pin_memory()

    src, crops, q, flags = (t.to(dev) for t in (src_h, crops_h, q_h, flags_h))
    out = torch.empty((N_IMAGES, HO, WO, 3), dtype=torch.float32, device=dev)

    def step():
        ops.preprocess(src, (HO, WO), crops, q, flags, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    _lib.launch_count_reset()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sampler.start()
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    kernel_ms = float(np.mean(per_kernel_ms))

    # e2e through the host-buffer entry point (pinned host in, pinned host out)
    out_h = torch.empty((N_IMAGES, HO, WO, 3), dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    ops.preprocess_host(src_h, (HO, WO), crops_h, q_h, flags_h, out=out_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.preprocess_host(src_h, (HO, WO), crops_h, q_h, flags_h, out=out_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kernel_ms = (float(v) for v in t.tolist())

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = BYTES_PER_IMAGE * N_IMAGES / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * N_IMAGES * args.steps / (total_ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu_per_step": N_IMAGES, "src": [HS, WS], "out": [HO, WO],
                       "out_dtype": "f32", "l2": "inputs+outputs (2.96 GB/step) larger than the 126 MB L2",
                       "sharding": "images sharded across ranks, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": world * N_IMAGES * e2e_steps / e2e_s, "unit": "images/s",
                    "h2d_bytes_per_step": int(src_h.nbytes + crops_h.nbytes + q_h.nbytes + flags_h.nbytes),
                    "d2h_bytes_per_step": int(out_h.nbytes), "steps": e2e_steps,
                    "api": "vip_preprocess_host (pinned host buffers, chunked H2D/kernel/D2H pipeline)"},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "kernel": "preprocess_kernel<f32>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_IMAGE * N_IMAGES,
                         "avg_launch_ms": kernel_ms,
                         "note": "JPEG emulation makes the kernel ALU-issue-bound, see DESIGN.md"},
        }
        if world == 1 and not args.no_cpu_baseline:
            n_sample = 256
            v = cpu_baseline(n_sample, 1, "integer")
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": 1, "kind": "port",
                                    "sample": f"first {n_sample} images of the batch, numpy oracle, 1 thread"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
