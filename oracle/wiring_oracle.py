"""ORACLE (test infrastructure, never shipped or timed as the product): numpy restatement of the NCP wiring.

Restates `velora/wiring.py` of the reference:
  * neuron split and synapse counts        wiring.py:95-101, 151-162, 185-190
  * per-layer connection sampling          wiring.py:208-267 (`_build_connections`)
  * recurrent (visualisation-only) mask    wiring.py:269-292
  * layer order inter -> command -> motor -> recurrent, motor built with the *command* synapse count
                                           wiring.py:294-328

Bit-exactness contract: the masks are a pure function of the GLOBAL numpy RNG stream and (only when a column
ends up empty) the GLOBAL torch CPU RNG stream, consumed in exactly the order the reference consumes them.
Duplicate (row, col) draws are resolved sequentially, last write wins (what the reference does with one
intra-op thread; see tests/golden/PROVENANCE.txt).

Parity pinned by tests/golden/wiring.npz (generated from the unmodified reference) and by the reference's own
structural tests restated in tests/test_wiring.py (tests/models/test_wiring.py:52-54, 81-117 of the reference).
"""
from __future__ import annotations

import numpy as np
import torch


def neuron_counts(in_features: int, n_neurons: int, out_features: int):
    n_command = max(int(0.4 * n_neurons), 1)            # wiring.py:100
    n_inter = n_neurons - n_command                      # wiring.py:101
    return in_features, n_inter, n_command, out_features


def synapse_counts(n_inter: int, n_command: int, sparsity_level: float):
    density = 1.0 - sparsity_level                       # wiring.py:98

    def syn(count, scale=1):
        return max(int(count * density * scale), 1)      # wiring.py:162

    # (sensory, inter, command, motor)                     wiring.py:185-190
    return syn(n_inter), syn(n_command), syn(n_command, 2), syn(n_command)


def _polarity(shape) -> np.ndarray:
    return np.random.choice([-1, 1], shape).astype(np.int32)   # wiring.py:206


def build_connections(n_rows: int, n_cols: int, count: int) -> np.ndarray:
    mask = np.zeros((n_rows, n_cols), dtype=np.int32)
    cols = np.random.choice(n_cols, (n_rows, count))      # wiring.py:245-247 (with replacement)
    pol = _polarity(cols.shape)                            # wiring.py:248
    for r in range(n_rows):                                # wiring.py:251, sequential => last write wins
        for k in range(count):
            mask[r, cols[r, k]] = pol[r, k]
    empty = np.nonzero((mask == 0).all(axis=0))[0]         # wiring.py:256-258
    if empty.size > 0:
        rows = torch.randint(0, n_rows, (int(empty.size),)).numpy()   # wiring.py:263 (global torch RNG)
        pol2 = _polarity((int(empty.size),))               # wiring.py:264
        for r, c, p in zip(rows, empty, pol2):             # wiring.py:265
            mask[r, c] = p
    return mask


def build_recurrent(n: int, count: int) -> np.ndarray:
    arr = np.zeros((n, n), dtype=np.int32)
    src = np.random.choice(n, count)                       # wiring.py:287
    dst = np.random.choice(n, count)                       # wiring.py:288
    pol = _polarity((count,))                              # wiring.py:289
    for s, d, p in zip(src, dst, pol):
        arr[s, d] = p
    return arr


def wiring(in_features: int, n_neurons: int, out_features: int, sparsity_level: float = 0.5):
    """Returns (masks dict of int32 arrays, neuron counts tuple, synapse counts tuple)."""
    if sparsity_level < 0.1 or sparsity_level > 0.9:       # wiring.py:95-96
        raise ValueError(f"'{sparsity_level=}' must be between '[0.1, 0.9]'.")
    n_in, n_inter, n_command, n_motor = neuron_counts(in_features, n_neurons, out_features)
    syn = synapse_counts(n_inter, n_command, sparsity_level)
    masks = {
        "inter": build_connections(n_in, n_inter, syn[0]),          # wiring.py:309-312
        "command": build_connections(n_inter, n_command, syn[1]),   # wiring.py:314-317
        "motor": build_connections(n_command, n_motor, syn[2]),     # wiring.py:319-322 (command count)
        "recurrent": build_recurrent(n_command, syn[2]),            # wiring.py:325-328
    }
    return masks, (n_in, n_inter, n_command, n_motor), syn
