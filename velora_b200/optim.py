"""Fused optimizer pieces of the train step.

FusedAdam      torch.optim.Adam whose update of ALL parameters is one kernel launch (vb_adam_multi) instead of ~15 foreach
               launches; state layout is exactly Adam(capturable=True)'s, so `state_dict()` round-trips with torch's Adam
               (reference: the four `optim.Adam` instances of velora/models/nf/modules.py:116-118, 409-415, 715-717).
soft_update    velora/utils/torch.py:53-64 for all parameters of a network pair in one launch (vb_polyak_multi).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn

from . import _lib


def _i64_array(vals: List[int]):
    return (C.c_int64 * len(vals))(*vals)


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, **kw):
        if weight_decay != 0.0 or kw.get("amsgrad") or kw.get("maximize"):
            raise ValueError("FusedAdam implements the plain Adam the reference constructs (no weight decay, amsgrad or maximize)")
        kw.pop("capturable", None)
        kw.pop("foreach", None)
        kw.pop("fused", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0.0, capturable=True, **kw)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            ps, gs, ms, vs, steps = [], [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, "FusedAdam.step")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 tensors")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.to(torch.float32).contiguous()
                ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"]); steps.append(st["step"])
            if not ps:
                continue
            torch._foreach_add_(steps, 1)
            dev = ps[0].device
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                _lib.check(
                    L.vb_adam_multi(_lib.ptr_array(ps), _lib.ptr_array(gs), _lib.ptr_array(ms), _lib.ptr_array(vs), _lib.ptr_array(steps),
                                    _i64_array([p.numel() for p in ps]), len(ps), float(group["lr"]), float(b1), float(b2),
                                    float(group["eps"]), _lib.stream_ptr(dev)),
                    "vb_adam_multi",
                )
        return loss


def soft_update(source: nn.Module, target: nn.Module, *, tau: float = 0.005) -> None:
    """target <- tau * source + (1 - tau) * target for every parameter pair, one launch."""
    tps, sps = [], []
    for tp, sp in zip(target.parameters(), source.parameters()):
        if tp.dtype != torch.float32 or sp.dtype != torch.float32 or not tp.is_contiguous() or not sp.is_contiguous():
            raise RuntimeError("soft_update: parameters must be contiguous fp32 tensors")
        tps.append(tp.data)
        sps.append(sp.data)
    if not tps:
        return
    _lib.require_cuda(tps[0], "velora_b200.optim.soft_update")
    dev = tps[0].device
    with torch.cuda.device(dev):
        _lib.check(
            _lib.lib().vb_polyak_multi(_lib.ptr_array(tps), _lib.ptr_array(sps), _i64_array([t.numel() for t in tps]), len(tps), float(tau),
                                       _lib.stream_ptr(dev)),
            "vb_polyak_multi",
        )
