"""Multi-GPU evaluation: whole videos shard across ranks (one process per GPU); the only exchange is an
all-gather of per-video fused scores and predictions.  The reference has no distributed
code (its nn.DataParallel wrapper is never invoked, spatialModel.py:133), so this layer is new design.

Rank r of R owns the contiguous block [r*ceil(V/R), min(V, (r+1)*ceil(V/R))).  Both streams of a video run on
the owning rank, so consensus and fusion are local; the fusion kernel writes each video's row directly into the
rank's slice of the gather buffer, and one in-place NCCL all-gather over NVLink publishes all rows to all ranks.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_videos: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(first, last_exclusive, rows_per_rank) of rank's contiguous block; rows_per_rank = ceil(V / R)."""
    per = (n_videos + world - 1) // world
    lo = min(n_videos, rank * per)
    hi = min(n_videos, (rank + 1) * per)
    return lo, hi, per


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # Two-stream training with deferred updates (training.py) runs a stream's gradient all-reduce BESIDE the other
            # stream's persistent one-CTA-per-SM layer kernels, on SMs left free by va_reserve_sms: the caller that enables it
            # caps NCCL's CTAs to that reservation BEFORE this call (reserve_nccl_ctas below); nothing is capped otherwise.
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def reserve_nccl_ctas(n: int = None) -> None:
    """Cap NCCL's CTAs to the SMs the deferred-update training schedule reserves for it (VA_ALLREDUCE_SMS, default 8).  Must
    run before the process group is created.  Measured (ms per step, batch 256 per GPU and stream): N = 2: 8 SMs x 5 layer
    launches 49.41, 16 x 3 49.84, 32 x 3 49.67, 24 x 5 50.11, collective after the backward pass 50.47
    (profiles/r02_train_defer_n2.json); N = 8: 51.31 / 51.39 (16 x 6) / 52.20 (32 x 8) / 53.74 (r02_train_defer_n8.json)."""
    os.environ.setdefault("NCCL_MAX_CTAS", str(n) if n is not None else os.environ.get("VA_ALLREDUCE_SMS", "8"))


def gather_video_rows(buffers: Dict[str, torch.Tensor], rank: int, world: int, per: int) -> Dict[str, torch.Tensor]:
    """In-place all-gather: each tensor in `buffers` is [world*per, ...] with this rank's rows already written at
    [rank*per, (rank+1)*per).  After the call every rank holds every row."""
    if world == 1:
        return buffers
    for name, full in buffers.items():
        assert full.shape[0] == world * per, (name, full.shape, world, per)
        mine = full[rank * per:(rank + 1) * per]
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(full, mine)
        else:   # gloo (CPU tests of the host logic)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine.clone())
            for r, part in enumerate(parts):
                full[r * per:(r + 1) * per].copy_(part)
    return buffers


def trim_rows(buffers: Dict[str, torch.Tensor], n_videos: int) -> Dict[str, torch.Tensor]:
    """Drop the padding rows of the last rank's block."""
    return {k: v[:n_videos] for k, v in buffers.items()}
