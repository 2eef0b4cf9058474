"""Seeding / device helpers (mirror of velora/utils/core.py:7-46)."""
from __future__ import annotations

import random

import numpy as np
import torch


def set_seed(value: int | None = None) -> int:
    """Seeds Python, torch and numpy (the three streams the wiring and the initialisers draw from)."""
    if value is None:
        value = random.randint(0, 2**32 - 1)
    random.seed(value)
    torch.manual_seed(value)
    np.random.seed(value)
    return value


def set_device(device: str = "auto") -> torch.device:
    if device == "auto":
        device = "cuda:0" if torch.cuda.is_available() else "cpu"
    return torch.device(device)
