"""Parameter bookkeeping helpers used by the networks (mirror of velora/utils/torch.py:53-105)."""
from __future__ import annotations

import torch
import torch.nn as nn


def soft_update(source: nn.Module, target: nn.Module, *, tau: float = 0.005) -> None:
    """Polyak averaging target <- tau * source + (1 - tau) * target, parameter by parameter."""
    with torch.no_grad():
        for tp, sp in zip(target.parameters(), source.parameters()):
            tp.data.copy_(tau * sp.data + (1.0 - tau) * tp.data)


def hard_update(source: nn.Module, target: nn.Module) -> None:
    with torch.no_grad():
        for tp, sp in zip(target.parameters(), source.parameters()):
            tp.data.copy_(sp.data)


@torch.jit.ignore
def total_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


@torch.jit.ignore
def active_parameters(model: nn.Module) -> int:
    """Trainable parameters that are not exactly zero (masked weights are zeroed at construction)."""
    return sum(int((p != 0).sum().item()) for p in model.parameters() if p.requires_grad)
