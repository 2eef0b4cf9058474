"""Seeding / device helpers with the behaviour of velora/utils/core.py:7-46: one integer seeds the three generators the wiring
sampler and the weight initialisers draw from (Python `random`, torch, numpy), in that order."""
from __future__ import annotations

import random
from typing import Optional

import numpy as np
import torch

_SEED_SPACE = 2**32 - 1


def set_seed(value: Optional[int] = None) -> int:
    """Seed every generator with `value` (a fresh random seed when None) and report the seed that was used."""
    seed = random.randint(0, _SEED_SPACE) if value is None else value
    for seeder in (random.seed, torch.manual_seed, np.random.seed):
        seeder(seed)
    return seed


def set_device(device: str = "auto") -> torch.device:
    """'auto' resolves to the first CUDA device when there is one, else the CPU; anything else is passed to torch.device."""
    if device != "auto":
        return torch.device(device)
    return torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
