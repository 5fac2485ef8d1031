"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The path shards naturally: training = every rank processes its share of the global
coordinate batch against identical replicas of [tables | MLP] and the flat gradient arena is
summed with ONE all-reduce per optimiser step; inference = contiguous slabs of the query volume,
no communication.  Nothing here is on the reference (it is single-device); gloo is used for the
CPU tests of this host logic.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process -> 0, 0, 1)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat: torch.Tensor, group=None) -> float:
    """In-place sum of the flat gradient arena over the ranks; returns 1/world (the averaging scale the
    optimiser applies, so the update equals the single-GPU update on the concatenated global batch)."""
    w = world_size(group)
    if w == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


def split_batch(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous share [first, first+count) of a global batch of n_global coordinates."""
    first = (n_global * rank) // world
    return first, (n_global * (rank + 1)) // world - first
