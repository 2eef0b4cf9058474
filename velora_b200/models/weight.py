"""Weight initialisation schemes for SparseLinear (mirror of velora/models/weight.py:6-98).

The default, "kaiming_uniform", draws weight then bias from the global torch stream exactly like torch.nn.Linear;
the draw order is part of the bit-exactness contract for initial parameters.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Literal, Optional

import torch.nn as nn

WeightInitType = Literal["xavier", "sonar", "zero", "trunc_normal", "kaiming_uniform", "orthogonal"]


def _zero_bias(layer: nn.Module) -> None:
    if layer.bias is not None:
        nn.init.zeros_(layer.bias)


def init_linear_xavier(layer: nn.Module) -> None:
    nn.init.xavier_normal_(layer.weight)
    _zero_bias(layer)


def init_linear_sonar(layer: nn.Module, sonar_std: float = 0.006) -> None:
    bound = sonar_std * (3 / layer.in_features) ** 0.5
    nn.init.uniform_(layer.weight, a=-bound, b=bound)
    _zero_bias(layer)


def init_linear_zero(layer: nn.Module) -> None:
    nn.init.zeros_(layer.weight)
    _zero_bias(layer)


def init_linear_trunc_normal(layer: nn.Module) -> None:
    nn.init.trunc_normal_(layer.weight, std=1e-3)
    _zero_bias(layer)


def init_linear_kaiming_uniform(layer: nn.Module) -> None:
    nn.init.kaiming_uniform_(layer.weight, a=math.sqrt(5))
    if layer.bias is not None:
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(layer.weight)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(layer.bias, -bound, bound)


def init_linear_orthogonal(layer: nn.Module, gain: float = 1.0) -> None:
    nn.init.orthogonal_(layer.weight, gain=gain)
    _zero_bias(layer)


_SCHEMES: Dict[str, Callable] = {
    "xavier": init_linear_xavier,
    "sonar": init_linear_sonar,
    "zero": init_linear_zero,
    "trunc_normal": init_linear_trunc_normal,
    "kaiming_uniform": init_linear_kaiming_uniform,
    "orthogonal": init_linear_orthogonal,
}


def get_init_fn(style: Optional[str]) -> Optional[Callable]:
    return None if style is None else _SCHEMES[style]
