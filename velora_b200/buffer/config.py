"""BufferConfig: the metadata model ReplayBuffer.config() returns (velora/models/config.py BufferConfig)."""
from typing import Literal

from pydantic import BaseModel


class BufferConfig(BaseModel):
    type: Literal["ReplayBuffer", "RolloutBuffer"]
    capacity: int
    state_dim: int
    action_dim: int
    hidden_dim: int
