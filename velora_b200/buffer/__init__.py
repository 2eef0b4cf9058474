from velora_b200.buffer.base import BufferBase
from velora_b200.buffer.experience import BatchExperience
from velora_b200.buffer.replay import ReplayBuffer

__all__ = ["BufferBase", "BatchExperience", "ReplayBuffer"]
