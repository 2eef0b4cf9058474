"""SparseLinear: a linear layer whose weight is gated by a fixed 0/1 wiring mask.

Mirror of velora/models/lnn/sparse.py:8-85 (same constructor, attributes, parameter/buffer names and draw order).
Inside LiquidNCPNetwork / NCPNetwork this module is a PARAMETER CONTAINER: the fused CUDA kernels read the
pre-masked weights of all layers at once (velora_b200/csrc).  Called on its own it evaluates the reference formula
`x @ (weight * mask).T + bias` with framework ops on whatever device it lives on.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from velora_b200.models.weight import WeightInitType, get_init_fn


class SparseLinear(nn.Module):
    bias: torch.Tensor
    mask: torch.Tensor

    def __init__(
        self,
        in_features: int,
        out_features: int,
        mask: torch.Tensor,
        *,
        init_type: str | WeightInitType = "kaiming_uniform",
        bias: bool = True,
        device: torch.device | None = None,
    ) -> None:
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.device = device

        # mask [out_features, in_features]; a buffer, so it follows .to() and is part of the state_dict
        self.register_buffer("mask", mask.to(device).detach())
        self.weight = nn.Parameter(torch.empty((out_features, in_features), device=device))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features, device=device))
        else:
            self.register_parameter("bias", None)

        self.reset_parameters(init_type)
        # masked synapses start (and, since their gradient is masked, stay) at exactly zero
        with torch.no_grad():
            self.weight.data.mul_(self.mask)

    def reset_parameters(self, style: str | WeightInitType) -> None:
        get_init_fn(style)(self)

    def update_mask(self, mask: torch.Tensor) -> None:
        self.mask = mask.to(self.device).detach()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight * self.mask, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"
