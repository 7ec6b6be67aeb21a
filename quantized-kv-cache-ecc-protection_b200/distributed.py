"""
Multi-GPU plumbing.  The path shards with NO cross-GPU data traffic: codec and cache work by
batch / sequence / KV head, decode attention by batch.  One process per GPU (torchrun); the only collective is a
SUM all-reduce of the small int64 counter vector when statistics are read -- NCCL over NVLink on GPUs, gloo in the
CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* when launched by torchrun.
    Returns (rank, world_size, local_rank); a plain `python` launch is (0, 1, 0) and initialises nothing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of n_items for `rank` (batch rows, KV heads or sequences); remainders go to the
    first ranks so sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_counters(vec: torch.Tensor) -> torch.Tensor:
    """SUM the counter vector over the default process group (no-op when not distributed)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        vec = vec.clone()
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank timing over the default group (device-timed numbers are reported as the max over ranks)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value
