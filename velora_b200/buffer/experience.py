"""BatchExperience: the struct-of-arrays minibatch returned by ReplayBuffer.sample (velora/buffer/experience.py)."""
from dataclasses import dataclass

import torch


@dataclass
class BatchExperience:
    """Six [batch, features] tensors; `hiddens` holds the actor's recurrent state at collection time."""

    states: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    next_states: torch.Tensor
    dones: torch.Tensor
    hiddens: torch.Tensor
