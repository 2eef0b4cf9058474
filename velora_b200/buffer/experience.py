"""BatchExperience: the struct-of-arrays minibatch `ReplayBuffer.sample` returns (velora/buffer/experience.py:6-27)."""
from dataclasses import dataclass, fields

import torch


@dataclass
class BatchExperience:
    """One [batch, width] float32 tensor per stored quantity; `hiddens` is the actor's recurrent state at collection time."""

    states: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    next_states: torch.Tensor
    dones: torch.Tensor
    hiddens: torch.Tensor


FIELD_NAMES = tuple(f.name for f in fields(BatchExperience))
