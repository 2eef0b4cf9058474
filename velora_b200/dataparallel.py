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

_state = {"enabled": False, "group": None, "reduce": "mean", "calls": 0, "oneshot": None, "oneshot_calls": 0}

# Blocks up to this many floats take the one-shot NVLink kernel (vb_allreduce_oneshot: latency-bound, one CTA); larger ones
# (the 526 k-float block of the 512-neuron class) are bandwidth problems and stay with NCCL.
ONESHOT_MAX_FLOATS = 16384


class _OneShot:
    """Symmetric buffers for vb_allreduce_oneshot: allocated and exchanged with torch.distributed._symmetric_memory (plumbing),
    read and written by this library's own kernel."""

    def __init__(self, group, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        L = _lib.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("one-shot all-reduce supports up to 16 ranks")
        self.cap = ONESHOT_MAX_FLOATS
        nbytes = int(L.vb_allreduce_oneshot_buffer_bytes(self.cap))
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        grp = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.buf, grp)
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every rank's flags are zero before anybody signals
        self.ptrs = (_lib.C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.device = device


def configure(enabled: bool = True, *, process_group=None, reduce: str = "mean", oneshot: bool | None = None) -> None:
    """Turn the gradient all-reduce on/off.  reduce='mean' matches batch-mean losses (nn.MSELoss, .mean()).

    oneshot: exchange small gradient blocks with this library's one-shot NVLink kernel instead of NCCL (None = when the symmetric
    memory set-up succeeds; False = NCCL only; True = required)."""
    if reduce not in ("mean", "sum"):
        raise ValueError("reduce must be 'mean' or 'sum'")
    if enabled and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    _state.update(enabled=enabled, group=process_group, reduce=reduce)
    if not enabled or oneshot is False or dist.get_world_size(process_group) == 1 or dist.get_backend(process_group) != "nccl":
        if oneshot is True and enabled and dist.get_world_size(process_group) > 1:
            raise RuntimeError("one-shot all-reduce needs the nccl backend (CUDA devices with peer access)")
        if oneshot is False:
            _state["oneshot"] = None      # (a plain configure(False) keeps the exchanged buffers for the next configure(True))
        return
    if _state["oneshot"] is None:
        try:
            _state["oneshot"] = _OneShot(process_group, torch.device("cuda", torch.cuda.current_device()))
        except Exception as e:      # no peer access / symmetric memory unavailable: NCCL stays the path
            if oneshot is True:
                raise
            import warnings

            warnings.warn(f"velora_b200.dataparallel: one-shot NVLink all-reduce unavailable ({e!r}); using NCCL")
            ok = torch.zeros(1, device="cuda")
        else:
            ok = torch.ones(1, device="cuda")
        # all ranks must take the same path
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
        if float(ok.item()) < 1.0:
            _state["oneshot"] = None


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
    one = _state["oneshot"]
    if (one is not None and flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() <= one.cap
            and flat.device == one.device):
        from . import _lib

        scale = 1.0 / ws if _state["reduce"] == "mean" else 1.0
        _lib.check(
            _lib.lib().vb_allreduce_oneshot(flat.data_ptr(), flat.numel(), one.ptrs, one.rank, one.world, one.cap, scale,
                                            one.counter.data_ptr(), _lib.stream_ptr(flat.device)),
            "vb_allreduce_oneshot",
        )
        _state["oneshot_calls"] += 1
    else:
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
