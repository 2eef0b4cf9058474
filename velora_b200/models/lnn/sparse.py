"""SparseLinear: a linear layer whose weight is gated by a fixed 0/1 wiring mask.

Behaviour of velora/models/lnn/sparse.py:8-85 (same constructor, attributes, parameter / buffer names and RNG draw order).
Inside LiquidNCPNetwork / NCPNetwork this module is a PARAMETER CONTAINER: the fused CUDA kernels read the pre-masked weights of
all layers at once (velora_b200/csrc).  Called on its own it evaluates `x @ (weight * mask).T + bias` with framework ops on
whatever device it lives on.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from velora_b200.models.weight import WeightInitType, get_init_fn


class SparseLinear(nn.Module):
    bias: torch.Tensor
    mask: torch.Tensor

    def __init__(self, in_features: int, out_features: int, mask: torch.Tensor, *, init_type: str | WeightInitType = "kaiming_uniform",
                 bias: bool = True, device: torch.device | None = None) -> None:
        super().__init__()
        self.in_features, self.out_features, self.device = in_features, out_features, device
        self._set_mask(mask)                                    # [out_features, in_features]; a buffer: follows .to(), is in the state_dict
        self.weight = nn.Parameter(torch.empty((out_features, in_features), device=device))
        self.register_parameter("bias", nn.Parameter(torch.empty(out_features, device=device)) if bias else None)
        self.reset_parameters(init_type)
        self._zero_masked_synapses()

    def _set_mask(self, mask: torch.Tensor) -> None:
        fresh = mask.to(self.device).detach()
        if "mask" in self._buffers:
            self.mask = fresh
        else:
            self.register_buffer("mask", fresh)

    @torch.no_grad()
    def _zero_masked_synapses(self) -> None:
        """Synapses the wiring removes start at exactly zero and stay there: their gradient is masked as well."""
        self.weight.data.mul_(self.mask)

    def reset_parameters(self, style: str | WeightInitType) -> None:
        get_init_fn(style)(self)

    def update_mask(self, mask: torch.Tensor) -> None:
        self._set_mask(mask)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight * self.mask, self.bias)

    def extra_repr(self) -> str:
        fields = dict(in_features=self.in_features, out_features=self.out_features, bias=self.bias is not None)
        return ", ".join(f"{k}={v}" for k, v in fields.items())
