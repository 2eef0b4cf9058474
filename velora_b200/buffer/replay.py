"""ReplayBuffer: uniform sampling with replacement; the six gathers are one fused kernel (vb_gather_rows).

Mirror of velora/buffer/replay.py:19-125.  The index vector is still drawn by `torch.randint` on the buffer's
device generator, so for the same seed the sampled indices - and therefore the returned rows - are bit-identical
to the reference's `tensor[indices]` results.
"""
from __future__ import annotations

from typing import TYPE_CHECKING, Optional

import torch

from velora_b200 import dataparallel as _dp
from velora_b200 import functional as VF
from velora_b200.buffer.base import FIELDS, BufferBase
from velora_b200.buffer.experience import BatchExperience

if TYPE_CHECKING:  # pragma: no cover
    from typing import Any


class ReplayBuffer(BufferBase):
    def __init__(self, capacity: int, state_dim: int, action_dim: int, hidden_dim: int, *, device: torch.device | None = None) -> None:
        super().__init__(capacity, state_dim, action_dim, hidden_dim, device=device)
        # sample from the packed shadow table (velora_b200/buffer/base.py): ~1.3x fewer DRAM sectors per transition.  Set to False to
        # read the six arrays directly (what a captured CUDA graph does: it must not depend on host-side freshness tracking).
        self.use_shadow = True

    def config(self):
        from velora_b200.buffer.config import BufferConfig

        return BufferConfig(type="ReplayBuffer", capacity=self.capacity, state_dim=self.state_dim,
                            action_dim=self.action_dim, hidden_dim=self.hidden_dim)

    def sample(self, batch_size: int) -> BatchExperience:
        """Random minibatch as six [batch_size, d] tensors."""
        if len(self) < batch_size:
            raise ValueError(f"Buffer does not contain enough experiences. Available: {len(self)}, Requested: {batch_size}")
        indices = torch.randint(0, self.size, (batch_size,), device=self.device)
        if _dp.is_enabled() and _dp.world_size() > 1:
            # data-parallel: `batch_size` is the global batch; every rank draws the same index vector (replicated buffers, identically
            # seeded generators) and gathers its contiguous slice - the union is bit-identical to a single-process sample(batch_size)
            lo, hi = _dp.shard_bounds(batch_size)
            indices = indices[lo:hi]
        return self.gather(indices)

    def gather(self, indices: torch.Tensor, n_rows: int | None = None) -> BatchExperience:
        """The gather half of `sample` for a caller-supplied int64 index vector (used by the data-parallel path,
        where every rank draws the same global vector and gathers its own slice).  `n_rows` bounds the valid indices (default: the
        rows holding data; a graph capture that must stay valid while the buffer fills passes the capacity)."""
        n_rows = self.size if n_rows is None else int(n_rows)
        arrays = [getattr(self, name) for name in FIELDS]
        if not arrays[0].is_cuda:
            raise RuntimeError(
                "velora_b200.ReplayBuffer.sample has no CPU path: build the buffer with device=torch.device('cuda')"
            )
        if self.use_shadow and not torch.cuda.is_current_stream_capturing():
            table, offsets, stride = self._shadow_table()
            out = VF.gather_packed(table, offsets, stride, [a.shape[1] for a in arrays], indices, n_rows)
        else:
            out = VF.gather_rows(arrays, indices, n_rows)
        return BatchExperience(*out)

    def warm(self, agent: "Any", n_samples: int, num_envs: int = 8) -> None:
        """Fills the buffer with `n_samples` transitions from a vectorised copy of the agent's environment."""
        if num_envs < 2:
            raise ValueError(f"'{num_envs=}' cannot be smaller than 2.")
        import gymnasium as gym   # optional dependency, only needed here

        envs = gym.make_vec(agent.env.spec.id, num_envs=num_envs, vectorization_mode="sync")
        envs = gym.wrappers.vector.NumpyToTorch(envs, agent.device)
        hidden: Optional[torch.Tensor] = None
        states, _ = envs.reset()
        while not len(self) >= n_samples:
            actions, hidden = agent.predict(states, hidden, train_mode=True)
            next_states, rewards, terminated, truncated, _ = envs.step(actions)
            self.add_multi(states, actions, rewards, next_states, terminated | truncated, hidden)
            states = next_states
        envs.close()
