"""Sparse NCP wiring masks (host side; cold path, bit-exact with the reference).

Mirror of the reference's `velora.wiring` public interface (`Wiring`, `NeuronCounts`, `SynapseCounts`,
`LayerMasks`; velora/wiring.py:8-338).  The masks are a pure function of the GLOBAL numpy RNG stream and, only when
a column ends up without a connection, the GLOBAL torch CPU RNG stream; this implementation consumes both in the
reference's order (inter -> command -> motor -> recurrent; per layer: column picks, polarities, then the optional
repair draw) so that `set_seed(s); Wiring(...)` yields identical int32 masks.

One documented difference: the reference scatters polarities with a torch advanced-index assignment whose winner
among duplicate (row, col) picks depends on the intra-op thread count, so the SIGN of a few entries of large
masks is not reproducible in the reference itself.  Here duplicates resolve sequentially (last write wins), which
is what the reference produces with one thread; `abs(mask)` - the only thing the networks use - is identical in
all cases.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch


@dataclass
class NeuronCounts:
    """Number of neurons per NCP category."""

    sensory: int
    inter: int
    command: int
    motor: int


@dataclass
class SynapseCounts:
    """Outgoing connections drawn per neuron, per NCP category."""

    sensory: int
    inter: int
    command: int
    motor: int


@dataclass
class LayerMasks:
    """int32 {-1, 0, 1} masks: inter [in, Ni], command [Ni, Nc], motor [Nc, out], recurrent [Nc, Nc]."""

    inter: torch.Tensor
    command: torch.Tensor
    motor: torch.Tensor
    recurrent: torch.Tensor


def _signs(shape) -> np.ndarray:
    return np.random.choice([-1, 1], shape)


class Wiring:
    """Builds the three layer masks (+ a visualisation-only recurrent mask) of an NCP network."""

    def __init__(self, in_features: int, n_neurons: int, out_features: int, *, sparsity_level: float = 0.5) -> None:
        if sparsity_level < 0.1 or sparsity_level > 0.9:
            raise ValueError(f"'{sparsity_level=}' must be between '[0.1, 0.9]'.")
        self.density_level = 1.0 - sparsity_level
        # 40 % of the decision neurons are command neurons (at least one), the rest inter neurons
        self.n_command = max(int(0.4 * n_neurons), 1)
        self.n_inter = n_neurons - self.n_command
        self.counts, self._n_connections = self._set_counts(in_features, out_features)
        self.masks = self._init_masks(in_features)
        self.build()

    # ------------------------------------------------------------------ bookkeeping
    @property
    def n_connections(self) -> SynapseCounts:
        return self._n_connections

    def _synapse_count(self, count: int, scale: int = 1) -> int:
        return max(int(count * self.density_level * scale), 1)

    def _set_counts(self, in_features: int, out_features: int) -> Tuple[NeuronCounts, SynapseCounts]:
        neurons = NeuronCounts(sensory=in_features, inter=self.n_inter, command=self.n_command, motor=out_features)
        synapses = SynapseCounts(
            sensory=self._synapse_count(self.n_inter),
            inter=self._synapse_count(self.n_command),
            command=self._synapse_count(self.n_command, scale=2),
            motor=self._synapse_count(self.n_command),
        )
        return neurons, synapses

    def _init_masks(self, n_inputs: int) -> LayerMasks:
        z = lambda r, c: torch.zeros((r, c), dtype=torch.int32)   # noqa: E731
        c = self.counts
        return LayerMasks(inter=z(n_inputs, c.inter), command=z(c.inter, c.command), motor=z(c.command, c.motor),
                          recurrent=z(c.command, c.command))

    # ------------------------------------------------------------------ sampling
    @staticmethod
    def polarity(shape: Tuple[int, ...] = (1,)) -> torch.IntTensor:
        """Random -1 / +1 matrix drawn from the global numpy stream."""
        return torch.IntTensor(_signs(shape))

    def _build_connections(self, mask: torch.Tensor, count: int) -> torch.Tensor:
        """`count` picks (with replacement) per row, then one repair connection for every empty column."""
        n_rows, n_cols = mask.shape
        m = mask.numpy()                                   # shares storage with `mask`
        picks = np.random.choice(n_cols, (n_rows, count))
        signs = _signs(picks.shape)
        m[np.arange(n_rows)[:, None], picks] = signs       # numpy: sequential, the last duplicate wins
        empty = np.flatnonzero(~m.any(axis=0))
        if empty.size:
            rows = torch.randint(0, n_rows, (int(empty.size),)).numpy()
            m[rows, empty] = _signs((int(empty.size),))
        return mask

    def _build_recurrent_connections(self, array: torch.Tensor, count: int) -> torch.Tensor:
        n = array.shape[0]
        src = np.random.choice(n, count)
        dst = np.random.choice(n, count)
        array.numpy()[src, dst] = _signs((count,))
        return array

    def build(self) -> None:
        nc = self._n_connections
        self.masks.inter = self._build_connections(self.masks.inter, nc.sensory)
        self.masks.command = self._build_connections(self.masks.command, nc.inter)
        # the motor layer is wired with the command synapse count (as in the reference; SynapseCounts.motor is unused)
        self.masks.motor = self._build_connections(self.masks.motor, nc.command)
        self.masks.recurrent = self._build_recurrent_connections(self.masks.recurrent, nc.command)

    def data(self) -> Tuple[LayerMasks, NeuronCounts]:
        return self.masks, self.counts
