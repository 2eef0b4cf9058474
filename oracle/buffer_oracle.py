"""ORACLE (test infrastructure, never shipped or timed as the product): numpy restatement of the replay buffer.

Restates the reference's
  storage        six pre-allocated float32 arrays + ring position/size     velora/buffer/base.py:61-71
  add            one row, position advances modulo capacity                velora/buffer/base.py:73-102
  add_multi      arange(position, position+n) % capacity scatter           velora/buffer/base.py:104-143
  sample         idx = torch.randint(0, size, (B,)); six row gathers       velora/buffer/replay.py:61-88

Index parity: the reference draws indices with the global torch generator of the buffer's device; the oracle
takes the index vector as an argument (tests draw it with the same torch call and seed), so the gather itself is
a pure integer/byte operation that must match bit for bit.

Parity pinned by tests/golden/buffer.npz (generated from the unmodified reference by oracle/make_golden.py) and
by the reference's ring-semantics tests restated in tests/test_buffer.py (reference tests/test_buffer.py:280-440).
"""
from __future__ import annotations

import numpy as np

FIELDS = ("states", "actions", "rewards", "next_states", "dones", "hiddens")


class RingBufferOracle:
    def __init__(self, capacity: int, state_dim: int, action_dim: int, hidden_dim: int):
        self.capacity = capacity
        self.position = 0
        self.size = 0
        self.states = np.zeros((capacity, state_dim), np.float32)
        self.actions = np.zeros((capacity, action_dim), np.float32)
        self.rewards = np.zeros((capacity, 1), np.float32)
        self.next_states = np.zeros((capacity, state_dim), np.float32)
        self.dones = np.zeros((capacity, 1), np.float32)
        self.hiddens = np.zeros((capacity, hidden_dim), np.float32)

    def add(self, state, action, reward, next_state, done, hidden) -> None:
        i = self.position
        self.states[i] = np.asarray(state, np.float32)
        self.actions[i] = np.asarray(action, np.float32)
        self.rewards[i] = np.float32(reward)
        self.next_states[i] = np.asarray(next_state, np.float32)
        self.dones[i] = np.float32(done)
        self.hiddens[i] = np.asarray(hidden, np.float32)
        self.position = (self.position + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def add_multi(self, states, actions, rewards, next_states, dones, hiddens) -> None:
        n = states.shape[0]
        idx = np.arange(self.position, self.position + n) % self.capacity
        col = lambda a: np.asarray(a, np.float32).reshape(n, -1)   # noqa: E731  (1-D -> [n,1], base.py:130-132)
        for name, src in zip(FIELDS, (states, actions, rewards, next_states, dones, hiddens)):
            dst = getattr(self, name)
            v = col(src)
            for k in range(n):          # sequential: later rows win when n > capacity
                dst[idx[k]] = v[k]
        self.position = (self.position + n) % self.capacity
        self.size = min(self.size + n, self.capacity)

    def gather(self, indices: np.ndarray):
        """replay.py:81-88 with the index vector given."""
        if self.size < len(indices):
            raise ValueError("Buffer does not contain enough experiences.")
        idx = np.asarray(indices, np.int64)
        return {name: getattr(self, name)[idx] for name in FIELDS}


def gather6(arrays, indices):
    """The bare gather: dst_k[b, :] = src_k[idx[b], :] for the six arrays."""
    idx = np.asarray(indices, np.int64)
    return [np.asarray(a)[idx] for a in arrays]
