"""Replaces ``utils/device.py:3-13`` (tf.distribute.MirroredStrategy over all GPUs of one host).

B200-native equivalent: one process per GPU (torchrun / torch.distributed, NCCL over NVLink), every model replicated,
the image list sharded contiguously across ranks, a single gather of the per-image probabilities at the end
(SURVEY.md 8e).  ``get_device()`` returns ``(strategy, 'GPU')`` where ``strategy`` carries rank / world information."""
from __future__ import annotations

import os

import torch


class ShardStrategy:
    """Stands where ``tf.distribute.Strategy`` stood: ``num_replicas_in_sync``, ``scope()`` and the shard arithmetic."""

    def __init__(self, rank=0, world=1, local_rank=0, backend=None):
        self.rank, self.world, self.local_rank, self.backend = rank, world, local_rank, backend

    @property
    def num_replicas_in_sync(self):
        return self.world

    def scope(self):
        import contextlib

        return contextlib.nullcontext()

    def shard_bounds(self, n):
        """Rank r owns the contiguous slice [r*ceil(n/R), min(n, (r+1)*ceil(n/R)))."""
        per = -(-n // self.world)
        lo = min(n, self.rank * per)
        return lo, min(n, lo + per), per

    def gather_rows(self, local, n_total):
        """local: f32/f64 tensor [M, n_local] on this rank's device -> [M, n_total] on every rank (one all_gather)."""
        if self.world == 1:
            return local
        import torch.distributed as dist

        lo, hi, per = self.shard_bounds(n_total)
        padded = torch.zeros((local.shape[0], per), dtype=local.dtype, device=local.device)
        padded[:, : hi - lo] = local
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded)
        return torch.cat(parts, dim=1)[:, :n_total]


def get_device(require_gpu=True):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and initialises NCCL when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        if require_gpu:
            raise RuntimeError("vipcup_b200 needs a CUDA device: there is no CPU path (utils/device.py's CPU branch is "
                               "intentionally not reproduced)")
        backend, device = "gloo", "CPU"
    else:
        torch.cuda.set_device(local_rank)
        backend, device = "nccl", "GPU"
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
            else:
                dist.init_process_group(backend)
    if rank == 0:
        print(f"\n> DEVICE: {device}")
    return ShardStrategy(rank, world, local_rank, backend if world > 1 else None), device
