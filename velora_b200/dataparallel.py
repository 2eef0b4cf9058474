"""Batch data-parallel training: one process per GPU, parameters replicated, ONE all-reduce per network backward.

The reference has no distributed code at all.  Every op on the hot path is independent per
batch row; the only cross-row reduction is the parameter-gradient sum.  Each rank runs the fused backward on its
batch shard, which leaves the whole gradient of a network in one flat pre-masked block (918 floats for the 20-neuron
actor, 4.6 k for the 128-neuron critic); that block is all-reduced over NCCL (NVLink 5 / NVSwitch) before it is
scattered to the seven `.grad` tensors.  Ranks must build their networks from the same seed (identical masks and
initial weights by construction: velora/wiring.py consumes only the global RNG streams).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

_state = {"enabled": False, "group": None, "reduce": "mean", "calls": 0}


def configure(enabled: bool = True, *, process_group=None, reduce: str = "mean") -> None:
    """Turn the gradient all-reduce on/off.  reduce='mean' matches batch-mean losses (nn.MSELoss, .mean())."""
    if reduce not in ("mean", "sum"):
        raise ValueError("reduce must be 'mean' or 'sum'")
    if enabled and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    _state.update(enabled=enabled, group=process_group, reduce=reduce)


def is_enabled() -> bool:
    return bool(_state["enabled"])


def world_size() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(_state["group"])
    return 1


def rank() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(_state["group"])
    return 0


def maybe_allreduce(flat: torch.Tensor) -> None:
    """In-place all-reduce of a flat gradient block when data-parallel mode is on (no-op otherwise)."""
    if not _state["enabled"]:
        return
    ws = world_size()
    if ws == 1:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_state["group"])
    if _state["reduce"] == "mean":
        flat.mul_(1.0 / ws)
    _state["calls"] += 1


def shard_bounds(n: int, rank: Optional[int] = None, world: Optional[int] = None):
    """Contiguous batch slice [lo, hi) of rank r."""
    if world is None:
        world = world_size()
    if rank is None:
        rank = dist.get_rank(_state["group"]) if world > 1 else 0
    if n % world != 0:
        # the noise slicing (functional.standard_normal_like), the shared index vector and reduce='mean' all assume equal shards:
        # an uneven split would hand the last rank the wrong noise rows and bias the mean of shard means
        raise ValueError(f"data-parallel batch of {n} rows does not divide evenly over {world} ranks")
    per = n // world
    return rank * per, (rank + 1) * per
