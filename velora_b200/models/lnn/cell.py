"""NCPLiquidCell: closed-form continuous-time (CfC) liquid cell over a sparse NCP wiring.

Mirror of velora/models/lnn/cell.py:10-169: five SparseLinear heads (g, h, f->g, f->h, proj) over cat([x, h]);
    gate = sigmoid(f_g + f_h);  h' = tanh(g) (1 - gate) + gate tanh(h);  y = proj + h'.
Inside LiquidNCPNetwork the heads are parameter containers for the fused network kernels; called on its own the cell runs
one fused kernel forward and one backward (vb_cell_fwd / vb_cell_bwd through functional.CellFn).  There is no CPU path.
"""

from typing import List, Tuple

import torch
import torch.nn as nn

from velora_b200 import functional as VF
from velora_b200.models.lnn.sparse import SparseLinear
from velora_b200.models.weight import WeightInitType


class NCPLiquidCell(nn.Module):
    sparsity_mask: torch.Tensor

    def __init__(
        self,
        in_features: int,
        n_hidden: int,
        mask: torch.Tensor,
        *,
        init_type: str | WeightInitType = "kaiming_uniform",
        device: torch.device | None = None,
    ) -> None:
        super().__init__()
        self.in_features = in_features
        self.n_hidden = n_hidden
        self.head_size = n_hidden + in_features
        self.init_type = init_type
        self.device = device

        self.register_buffer("sparsity_mask", self._prep_mask(mask.to(device)))

        self.tanh = nn.Tanh()
        self.sigmoid = nn.Sigmoid()

        # construction order = RNG draw order = parameter order of the checkpoint format
        self.g_head = self._make_layer()
        self.h_head = self._make_layer()
        self.f_head_to_g = self._make_layer()
        self.f_head_to_h = self._make_layer()
        self.proj = self._make_layer()

    def _make_layer(self) -> SparseLinear:
        return SparseLinear(self.head_size, self.n_hidden, self.sparsity_mask, init_type=self.init_type, device=self.device)

    def _prep_mask(self, mask: torch.Tensor) -> torch.Tensor:
        """[in, hidden] wiring mask -> [hidden, in + hidden] 0/1 head mask: the recurrent columns are dense."""
        n_hidden = mask.shape[1]
        recurrent = torch.ones((n_hidden, n_hidden), device=self.device)
        return torch.abs(torch.concatenate([mask.detach(), recurrent]).T).to(self.device)

    def update_mask(self, mask: torch.Tensor) -> None:
        self.sparsity_mask = self._prep_mask(mask.to(self.device))

    @torch.jit.ignore
    def _fused(self, x: torch.Tensor, hidden: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        heads = [self.g_head, self.h_head, self.f_head_to_g, self.f_head_to_h, self.proj]
        dev = self.g_head.weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "velora_b200.NCPLiquidCell has no CPU path: build it with device=torch.device('cuda') "
                f"(parameters are on '{dev}')"
            )
        x = x.to(dtype=torch.float32, device=dev)
        hidden = hidden.to(dtype=torch.float32, device=dev)
        params: List[torch.Tensor] = []
        for h in heads:
            params += [h.weight, h.bias]
        # masks come from each head's own buffer, as in LiquidNCPNetwork._fused_seq
        return VF.CellFn.apply(x, hidden, (self.in_features, self.n_hidden), tuple(h.mask for h in heads), *params)

    def forward(self, x: torch.Tensor, hidden: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(x [B, in], h [B, n_hidden]) -> (y [B, n_hidden], h' [B, n_hidden])   (cell.py:147-169)"""
        return self._fused(x, hidden)
