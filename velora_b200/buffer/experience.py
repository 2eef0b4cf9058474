"""BatchExperience: the struct-of-arrays minibatch `ReplayBuffer.sample` returns (role of velora/buffer/experience.py:6-27).

One [batch, width] float32 tensor per stored quantity, in the order the buffer keeps them; `hiddens` is the actor's recurrent
state at collection time.  The field list is the single source of truth shared with the buffer (`FIELD_NAMES`)."""
from dataclasses import make_dataclass

import torch

FIELD_NAMES = ("states", "actions", "rewards", "next_states", "dones", "hiddens")

BatchExperience = make_dataclass(
    "BatchExperience",
    [(name, torch.Tensor) for name in FIELD_NAMES],
    namespace={"__doc__": "Six [batch, features] tensors: " + ", ".join(FIELD_NAMES) + "."},
)
BatchExperience.__module__ = __name__
