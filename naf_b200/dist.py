"""Multi-GPU plumbing: one process per GPU, batch sharding, one weight broadcast.

The reference is single-process/single-GPU (no collective anywhere).  The
upsampling of different images is independent (GroupNorm statistics are per sample, attention is
per image), so the path shards over the batch with NO data-path collective: each rank upsamples
its contiguous slice of the batch and keeps its outputs.  The only communication is one
broadcast of the encoder weights + the RoPE `periods` buffer at start-up (662,528 + 16 fp32 for
the default model), flattened into a single buffer so it is one NCCL launch over NVLink.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def shard_range(total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) slice of `total` items owned by `rank`; sizes differ by <= 1."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank).
    A plain `python` launch (no RANK in the env) is world_size 1 and initialises nothing."""
    if "RANK" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


@torch.no_grad()
def broadcast_module_(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """Make every rank's parameters and persistent buffers equal to rank `src`'s with ONE
    broadcast of a flat fp32 buffer.  Returns the number of elements sent."""
    tensors = [t for _, t in sorted(module.state_dict().items())]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(t.numel() for t in tensors)
    dev = tensors[0].device
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n
    assert flat.device == dev
    return flat.numel()


def max_over_ranks(value: float, device) -> float:
    """MAX-reduce a scalar (device-timed milliseconds) over all ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
