"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The EBEN step shards over batch items only; the single exchange
is a sum all-reduce of each network's flat gradient bucket."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank_seed(base: int, rank: int) -> int:
    """Each rank draws its own synthetic batch (seed 42 + rank); weights use `base`."""
    return base + rank


def allreduce_sum_(bucket: torch.Tensor) -> float:
    """In-place sum all-reduce of a flat gradient bucket; returns the factor (1/world) the fused Adam
    kernel applies so that the update uses the mean gradient (DDP semantics)."""
    w = world_size()
    if w > 1:
        dist.all_reduce(bucket)
    return 1.0 / w


def max_over_ranks(value: float, device=None) -> float:
    if world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if world_size() > 1:
        dist.barrier()
