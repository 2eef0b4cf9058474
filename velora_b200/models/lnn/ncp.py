"""LiquidNCPNetwork and NCPNetwork: the public networks of the hot path, backed by the fused sm_100a kernels.

Mirror of velora/models/lnn/ncp.py:13-328 - same constructors, attributes, sub-module names, state_dict keys,
parameter order, error behaviour and return conventions.  `forward` does not run the layer chain
(26 ATen launches per step in the reference): it hands the seven (three) masked layers to one kernel
through a `@torch.jit.ignore` method, so the module still survives `copy.deepcopy` + `torch.jit.script`
(velora/models/nf/modules.py:370-392).  There is no CPU path: tensors are moved to the module's CUDA device as the
reference does (ncp.py:158), and a module built without a CUDA device raises.
"""

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from velora_b200 import functional as VF
from velora_b200.models.lnn.cell import NCPLiquidCell
from velora_b200.models.lnn.sparse import SparseLinear
from velora_b200.models.weight import WeightInitType
from velora_b200.utils.torch import active_parameters, total_parameters
from velora_b200.wiring import Wiring


class LiquidNCPNetwork(nn.Module):
    """inter (SparseLinear) -> Mish -> command (NCPLiquidCell) -> Mish -> motor (SparseLinear)."""

    def __init__(
        self,
        in_features: int,
        n_neurons: int,
        out_features: int,
        *,
        sparsity_level: float = 0.5,
        init_type: str | WeightInitType = "kaiming_uniform",
        device: torch.device | None = None,
    ) -> None:
        super().__init__()
        self.in_features = in_features
        self.n_neurons = n_neurons
        self.out_features = out_features
        self.device = device
        self.n_units = n_neurons + out_features

        self.wiring = Wiring(in_features, n_neurons, out_features, sparsity_level=sparsity_level)
        self.masks, self.counts = self.wiring.data()

        self.inter = SparseLinear(in_features, self.counts.inter, torch.abs(self.masks.inter.T), init_type=init_type, device=device).to(device)
        self.command = NCPLiquidCell(self.counts.inter, self.counts.command, self.masks.command, init_type=init_type, device=device).to(device)
        self.hidden_size = self.counts.command
        self.motor = SparseLinear(self.counts.command, self.counts.motor, torch.abs(self.masks.motor.T), init_type=init_type, device=device).to(device)
        self.act = nn.Mish()

        self.kernel_flags = 0   # VB_FLAG_* bits; 1 forces the generic kernel (parity tests)
        self._total_params = total_parameters(self)
        self._active_params = active_parameters(self)

    @property
    def total_params(self) -> int:
        return self._total_params

    @property
    def active_params(self) -> int:
        return self._active_params

    # ------------------------------------------------------------------ fused path
    @torch.jit.ignore
    def _fused_seq(self, x: torch.Tensor, h0: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """x [B,T,I] -> (y [B,T,O], h_traj [B,T,Nc], h_last [B,Nc]) through vb_liquid_fwd / vb_liquid_bwd."""
        cmd = self.command
        heads = [cmd.g_head, cmd.h_head, cmd.f_head_to_g, cmd.f_head_to_h, cmd.proj]
        # masks are read from each layer's own buffer (update_mask on the cell does not touch the heads' buffers,
        # and load_state_dict restores them independently: SURVEY.md section 7)
        masks = (self.inter.mask, *[h.mask for h in heads], self.motor.mask)
        params: List[torch.Tensor] = [self.inter.weight, self.inter.bias]
        for h in heads:
            params += [h.weight, h.bias]
        params += [self.motor.weight, self.motor.bias]
        if VF.parameters_frozen():
            params = [p.detach() for p in params]
        dims = (self.in_features, int(self.inter.out_features), int(cmd.n_hidden), int(self.motor.out_features))
        return VF.LiquidSeqFn.apply(x, h0, dims, int(self.kernel_flags), masks, *params)

    @torch.jit.ignore
    def _target_device(self) -> torch.device:
        dev = self.inter.weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "velora_b200.LiquidNCPNetwork has no CPU path: build it with device=torch.device('cuda') "
                f"(parameters are on '{dev}')"
            )
        return dev

    def forward(self, x: torch.Tensor, h_state: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(x [B, in], h [B, Nc] | None) -> (y [B, out] or [out] when B == 1, h' [B, Nc])."""
        if x.dim() != 2:
            raise ValueError(f"Unsupported dimensionality: '{x.shape}'. Should be 2 dimensional with: '(batch_size, features)'.")
        dev = self._target_device()
        x = x.to(dtype=torch.float32, device=dev)
        if h_state is not None:
            h_state = h_state.to(dev)
        y_seq, _, h_new = self._fused_seq(x.unsqueeze(1), h_state)
        y_pred = y_seq[:, 0]
        if y_pred.shape[0] == 1:
            y_pred = y_pred.squeeze(0)
        return y_pred, h_new

    @torch.jit.ignore
    def forward_seq(self, x: torch.Tensor, h0: Optional[torch.Tensor] = None, return_last: bool = False):
        """The T-fold loop of `forward` carrying h, with the time loop on chip.

        x [B, T, in] -> (y [B, T, out], h_traj [B, T, Nc]) (+ h_T [B, Nc] when return_last; using h_T instead of
        h_traj[:, -1] in a loss keeps the backward from materialising a dense dL/dh_traj).
        """
        if x.dim() != 3:
            raise ValueError(f"Unsupported dimensionality: '{x.shape}'. Should be 3 dimensional with: '(batch_size, steps, features)'.")
        dev = self._target_device()
        x = x.to(dtype=torch.float32, device=dev)
        if h0 is not None:
            h0 = h0.to(dev)
        y, h_traj, h_last = self._fused_seq(x, h0)
        return (y, h_traj, h_last) if return_last else (y, h_traj)


    @torch.jit.ignore
    def forward_seq_bf16(self, x: torch.Tensor, h0: Optional[torch.Tensor] = None, return_traj: bool = False):
        """`forward_seq` with the cell's heads on the tensor cores in bf16 (bf16 operands, fp32 accumulation and hidden state; agrees
        with `forward_seq` to ~1e-2 relative, BASELINE's bf16 contract is 2e-2).  Available for the 512-neuron class (BASELINE config 4).
        Returns (y [B,T,out], h_T [B,Nc]), differentiable w.r.t. x, h0 and the fourteen parameters (backward on tcgen05 as well:
        vb_liquid_bf16_bwd); with `return_traj` the hidden trajectory is returned too and the call is inference-only."""
        if x.dim() != 3:
            raise ValueError(f"Unsupported dimensionality: '{x.shape}'. Should be 3 dimensional with: '(batch_size, steps, features)'.")
        dev = self._target_device()
        x = x.to(dtype=torch.float32, device=dev)
        if h0 is not None:
            h0 = h0.to(dev)
        cmd = self.command
        heads = [cmd.g_head, cmd.h_head, cmd.f_head_to_g, cmd.f_head_to_h, cmd.proj]
        masks = (self.inter.mask, *[h.mask for h in heads], self.motor.mask)
        params: List[torch.Tensor] = [self.inter.weight, self.inter.bias]
        for h in heads:
            params += [h.weight, h.bias]
        params += [self.motor.weight, self.motor.bias]
        dims = (self.in_features, int(self.inter.out_features), int(cmd.n_hidden), int(self.motor.out_features))
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or (h0 is not None and h0.requires_grad) or any(p.requires_grad for p in params))
        if needs_grad and not return_traj:
            return VF.LiquidBf16SeqFn.apply(x, h0, dims, masks, *params)
        with torch.no_grad():
            return VF.liquid_bf16_forward(x, h0, dims, masks, params, want_traj=return_traj)


class NCPNetwork(nn.Module):
    """SparseLinear -> Mish -> SparseLinear -> Mish -> SparseLinear (the SAC critics' network)."""

    def __init__(
        self,
        in_features: int,
        n_neurons: int,
        out_features: int,
        *,
        sparsity_level: float = 0.5,
        init_type: str | WeightInitType = "kaiming_uniform",
        device: torch.device | None = None,
    ) -> None:
        super().__init__()
        self.in_features = in_features
        self.n_neurons = n_neurons
        self.out_features = out_features
        self.device = device
        self.n_units = n_neurons + out_features

        self.wiring = Wiring(in_features, n_neurons, out_features, sparsity_level=sparsity_level)
        self.masks, self.counts = self.wiring.data()

        self.ncp = nn.Sequential(
            SparseLinear(in_features, self.counts.inter, torch.abs(self.masks.inter.T), init_type=init_type, device=device),
            nn.Mish(),
            SparseLinear(self.counts.inter, self.counts.command, torch.abs(self.masks.command.T), init_type=init_type, device=device),
            nn.Mish(),
            SparseLinear(self.counts.command, self.counts.motor, torch.abs(self.masks.motor.T), init_type=init_type, device=device),
        ).to(device)

        self.kernel_flags = 0
        self._total_params = total_parameters(self)
        self._active_params = active_parameters(self)

    @property
    def total_params(self) -> int:
        return self._total_params

    @property
    def active_params(self) -> int:
        return self._active_params

    @torch.jit.ignore
    def _fused(self, x: torch.Tensor) -> torch.Tensor:
        # attribute access (not indexing): under torch.jit.script `self` is the RecursiveScriptModule
        l1, l2, l3 = getattr(self.ncp, "0"), getattr(self.ncp, "2"), getattr(self.ncp, "4")
        dev = l1.weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "velora_b200.NCPNetwork has no CPU path: build it with device=torch.device('cuda') "
                f"(parameters are on '{dev}')"
            )
        x = x.to(dtype=torch.float32, device=dev)
        dims = (self.in_features, int(l1.out_features), int(l2.out_features), int(l3.out_features))
        masks = (l1.mask, l2.mask, l3.mask)
        params = (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)
        if VF.parameters_frozen():
            params = tuple(p.detach() for p in params)
        return VF.NcpFn.apply(x, dims, int(self.kernel_flags), masks, *params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, in] -> y [B, out] (or [out] when B == 1)."""
        if x.dim() != 2:
            raise ValueError(f"Unsupported dimensionality: '{x.shape}'. Should be 2 dimensional with: '(batch_size, features)'.")
        y_pred = self._fused(x)
        if y_pred.shape[0] == 1:
            y_pred = y_pred.squeeze(0)
        return y_pred
