"""Pre-allocated ring buffer of transitions (mirror of velora/buffer/base.py:30-279).

Storage layout in HBM: six row-major float32 arrays [capacity, d] (d = S, A, 1, S, 1, H) - 76 B per transition for
the (4, 1, 8) shapes of BASELINE configs 1/2/5 - exactly the reference's state_dict, so checkpoints interchange.
"""
from __future__ import annotations

import json
from abc import abstractmethod
from pathlib import Path
from typing import Any, Dict, Literal

import torch

from velora_b200.buffer.experience import FIELD_NAMES, BatchExperience

MetaDataKeys = Literal["capacity", "state_dim", "action_dim", "hidden_dim", "position", "size", "device"]
BufferKeys = Literal["states", "actions", "rewards", "next_states", "dones", "hiddens"]
FIELDS = FIELD_NAMES


class BufferBase:
    fused_add = True      # `add` as one launch (vb_buffer_add_row); False = the reference's six indexing assignments
    def __init__(self, capacity: int, state_dim: int, action_dim: int, hidden_dim: int, *, device: torch.device | None = None) -> None:
        self.capacity = capacity
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.hidden_dim = hidden_dim
        self.device = device

        self.position = 0   # next row to write
        self.size = 0       # rows holding data

        widths = dict(states=state_dim, actions=action_dim, rewards=1, next_states=state_dim, dones=1, hiddens=hidden_dim)
        for name in FIELDS:
            setattr(self, name, torch.zeros((capacity, widths[name]), device=device))
        # packed shadow table read by the fused gather (one run of sectors per transition); built lazily on CUDA, refreshed row-wise
        # by add / add_multi, rebuilt when the six arrays (the truth: state_dict, save / load) were written from outside
        self._shadow = None
        self._shadow_key = None

    # ------------------------------------------------------------------ packed shadow table
    def _arrays_key(self):
        """(identity, data pointer, version counter) of the six arrays.  The counter is bumped by every in-place torch op on the tensor,
        NOT by writes through `.data` / detached aliases or by raw-pointer kernels: after such a write call `invalidate_shadow()`
        (or set `use_shadow = False` on the buffer).  Tensors without a version counter (inference mode) never match: the table is
        then rebuilt for every sample."""
        def version(t):
            try:
                return t._version
            except Exception:
                return object()
        return tuple((id(getattr(self, name)), getattr(self, name).data_ptr(), version(getattr(self, name))) for name in FIELDS)

    def invalidate_shadow(self) -> None:
        """Forget the packed table: the next `sample` rebuilds it from the six arrays (the truth)."""
        self._shadow_key = None

    def _shadow_current(self) -> bool:
        return self._shadow is not None and self._shadow_key == self._arrays_key()

    def _refresh_shadow(self, rows=None, row0: int = 0, n: int | None = None) -> None:
        from velora_b200 import functional as VF

        arrays = [getattr(self, name) for name in FIELDS]
        if self._shadow is None:
            self._shadow_offsets, self._shadow_stride = VF.packed_row_layout([a.shape[1] for a in arrays])
            self._shadow = torch.zeros((self.capacity, self._shadow_stride), device=arrays[0].device)
        VF.pack_rows(arrays, self._shadow, self._shadow_offsets, self._shadow_stride, rows=rows, row0=row0, n=n)
        self._shadow_key = self._arrays_key()

    def _shadow_table(self):
        """The packed table, current for rows [0, size).  Rebuilt in full when the arrays changed behind its back."""
        if not self._shadow_current():
            self._refresh_shadow(row0=0, n=max(self.size, 1))
        return self._shadow, self._shadow_offsets, self._shadow_stride

    # ------------------------------------------------------------------ writes
    def add(self, state: torch.Tensor, action: torch.Tensor, reward: float, next_state: torch.Tensor, done: bool, hidden: torch.Tensor) -> None:
        """Appends one transition (overwriting the oldest once full)."""
        i = self.position
        tracked = self._shadow_current()
        if self.fused_add and self.states.is_cuda and self._add_fused(i, (state, action, reward, next_state, done, hidden), tracked):
            self.position = (i + 1) % self.capacity
            self.size = min(self.size + 1, self.capacity)
            return
        self.states[i] = state.to(torch.float32)
        self.actions[i] = action
        self.rewards[i] = reward
        self.next_states[i] = next_state.to(torch.float32)
        self.dones[i] = done
        self.hiddens[i] = hidden
        if tracked:
            self._refresh_shadow(row0=i, n=1)
        self.position = (i + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def _add_fused(self, i: int, values, tracked: bool) -> bool:
        """The six assignments of `add` (and the shadow row) in one launch (vb_buffer_add_row) when every tensor value already lives on
        the buffer's device with the row's element count; anything else takes the reference's indexing writes.  The kernel does not bump
        the arrays' version counters, and it writes the shadow row itself, so a current shadow table stays current."""
        from velora_b200 import functional as VF

        arrays = [getattr(self, name) for name in FIELDS]
        vals = []
        f32 = torch.float32
        for a, v in zip(arrays, values):
            if isinstance(v, torch.Tensor):
                if v.device != a.device:
                    if v.numel() == 1 and v.device.type == "cpu":
                        vals.append(float(v))          # a host scalar tensor: no device round trip
                        continue
                    return False
                w = a.shape[1]
                if v.numel() != w:
                    if v.numel() != 1:
                        return False
                    v = v.detach().to(f32).reshape(1).expand(w).contiguous()
                elif v.dtype is not f32 or not v.is_contiguous() or v.requires_grad:
                    v = v.detach().to(f32).contiguous()
                vals.append(v)
            elif isinstance(v, (bool, int, float)):
                vals.append(float(v))
            else:
                return False
        # the ctypes view of the six arrays (pointers, widths, shadow layout) is rebuilt only when an array was replaced
        key = tuple(a.data_ptr() for a in arrays) + (self._shadow.data_ptr() if tracked else 0,)
        cache = getattr(self, "_add_args", None)
        if cache is None or cache[0] != key:
            cache = (key, VF.buffer_add_args(arrays, self._shadow if tracked else None,
                                            self._shadow_offsets if tracked else None, self._shadow_stride if tracked else 0))
            self._add_args = cache
        VF.buffer_add_row_cached(cache[1], vals, i)
        return True

    def add_multi(self, states: torch.Tensor, actions: torch.Tensor, rewards: torch.Tensor, next_states: torch.Tensor,
                  dones: torch.Tensor, hiddens: torch.Tensor) -> None:
        """Appends a batch of transitions with wrap-around."""
        n = states.shape[0]
        end = self.position + n
        rows = torch.arange(self.position, end) % self.capacity

        def col(t: torch.Tensor) -> torch.Tensor:
            return t.unsqueeze(-1) if t.dim() == 1 else t

        f32 = torch.float32
        tracked = self._shadow_current()
        self.states[rows] = states.to(f32)
        self.actions[rows] = col(actions).to(f32)
        self.rewards[rows] = col(rewards).to(f32)
        self.next_states[rows] = next_states.to(f32)
        self.dones[rows] = col(dones).to(f32)
        self.hiddens[rows] = hiddens
        if tracked:
            self._refresh_shadow(rows=rows)
        self.position = end % self.capacity
        self.size = min(self.size + n, self.capacity)

    @abstractmethod
    def sample(self) -> BatchExperience:
        pass  # pragma: no cover

    def __len__(self) -> int:
        return self.size

    # ------------------------------------------------------------------ persistence
    def metadata(self) -> Dict[MetaDataKeys, Any]:
        return {
            "capacity": self.capacity,
            "state_dim": self.state_dim,
            "action_dim": self.action_dim,
            "hidden_dim": self.hidden_dim,
            "position": self.position,
            "size": self.size,
            "device": str(self.device) if self.device else None,
        }

    def state_dict(self) -> Dict[BufferKeys, torch.Tensor]:
        return {name: getattr(self, name) for name in FIELDS}

    def save(self, dirpath: str | Path, prefix: str = "buffer_") -> None:
        """Writes `<prefix>metadata.json` and `<prefix>state.safetensors` (the reference's on-disk format)."""
        from safetensors.torch import save_file

        root = Path(dirpath)
        root.mkdir(parents=True, exist_ok=True)
        save_file(self.state_dict(), Path(root, f"{prefix}state").with_suffix(".safetensors"))
        with Path(root, f"{prefix}metadata").with_suffix(".json").open("w") as f:
            f.write(json.dumps(self.metadata(), indent=2))

    @classmethod
    def load(cls, state_path: str | Path, metadata: Dict[MetaDataKeys, Any]):
        from safetensors.torch import load_file

        device = metadata["device"] or "cpu"
        buffer = cls(
            capacity=metadata["capacity"],
            state_dim=metadata["state_dim"],
            action_dim=metadata["action_dim"],
            hidden_dim=metadata["hidden_dim"],
            device=torch.device(device) if device else None,
        )
        buffer.position = metadata["position"]
        buffer.size = metadata["size"]
        data = load_file(Path(state_path).with_suffix(".safetensors"), device)
        for name in FIELDS:
            setattr(buffer, name, data[name])
        return buffer
