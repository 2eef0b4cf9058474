"""SAC heads over the fused networks (mirror of velora/models/sac/continuous.py:9-232).

These are the direct CALLERS of the hot path: the policy/Q heads themselves stay framework
ops (chunk, clamp, exp, reparameterised sample, tanh squash) so that sampling keeps torch's RNG semantics; `self.ncp`
is the fused LiquidNCPNetwork / NCPNetwork.  Like the reference, the modules are deep-copied and torch.jit.script-ed
by the agent modules, so the sampling helpers are `@torch.jit.ignore` methods.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch.distributions import Normal

from velora_b200.models.lnn.ncp import LiquidNCPNetwork, NCPNetwork


class SACActor(nn.Module):
    """Liquid NCP actor: Gaussian policy with tanh squashing; `ncp` outputs [mean | log_std]."""

    action_scale: torch.Tensor
    action_bias: torch.Tensor

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, action_scale: torch.Tensor, action_bias: torch.Tensor,
                 *, log_std_min: float = -5, log_std_max: float = 2, device: torch.device | None = None,
                 fused_head: bool = False) -> None:
        super().__init__()
        self.device = device
        # True: the policy-head math after the network runs as one kernel forward / one backward (vb_sac_head_*) instead of ~50
        # elementwise launches; same noise draw, same formula
        self.fused_head = fused_head
        self.ncp = LiquidNCPNetwork(num_obs, n_neurons, num_actions * 2, device=device).to(device)
        self.log_std_min = log_std_min
        self.log_std_max = log_std_max
        self.register_buffer("action_scale", action_scale)
        self.register_buffer("action_bias", action_bias)

    @torch.jit.ignore
    def get_sample(self, mean: torch.Tensor, std: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reparameterised sample and its log-density (continuous.py:57-77)."""
        # argument validation reads a device flag back to the host, which is illegal while a CUDA graph is being captured
        validate = not (mean.is_cuda and torch.cuda.is_current_stream_capturing())
        dist = Normal(mean, std, validate_args=validate)
        from velora_b200 import functional as VF

        x_t = mean + VF.standard_normal_like(mean) * std      # == dist.rsample(): same generator call, shard-aware under data parallelism
        return x_t, dist.log_prob(x_t)

    @torch.jit.ignore
    def _fused_head(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        from velora_b200 import functional as VF

        mean = x[:, : x.shape[1] // 2]
        eps = VF.standard_normal_like(mean)      # what Normal(mean, std).rsample() draws
        actions, log_prob = VF.SacHeadFn.apply(x, eps, self.action_scale, self.action_bias, self.log_std_min, self.log_std_max)
        return actions, log_prob

    @torch.jit.ignore
    def predict(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Deterministic action (continuous.py:79-104)."""
        x, new_hidden = self.ncp(obs, hidden)
        mean, _ = torch.chunk(x, 2, dim=-1)
        return torch.tanh(mean) * self.action_scale + self.action_bias, new_hidden

    def forward(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(obs, hidden) -> (actions, log_prob [B,1], new hidden)   (continuous.py:106-143)"""
        x, new_hidden = self.ncp(obs, hidden)
        if self.fused_head and x.dim() == 2:
            actions_f, log_prob_f = self._fused_head(x)
            return actions_f, log_prob_f, new_hidden
        mean, log_std = torch.chunk(x, 2, dim=-1)
        log_std = torch.clamp(log_std, self.log_std_min, self.log_std_max)
        std = log_std.exp()
        x_t, dist_log_probs = self.get_sample(mean, std)
        actions_normalized = torch.tanh(x_t)
        actions = actions_normalized * self.action_scale + self.action_bias
        log_prob = dist_log_probs - torch.log(self.action_scale * (1 - actions_normalized.pow(2)) + 1e-6)
        return actions, log_prob.sum(dim=-1, keepdim=True), new_hidden


class SACCriticNCP(nn.Module):
    """NCP critic: Q(obs, action) through the fused three-layer NCPNetwork (continuous.py:194-232)."""

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, *, device: torch.device | None = None) -> None:
        super().__init__()
        self.device = device
        self.ncp = NCPNetwork(num_obs + num_actions, n_neurons, 1, device=device).to(device)

    def forward(self, obs: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        return self.ncp(torch.cat([obs, actions], dim=-1))


class SACCritic(nn.Module):
    """Liquid NCP critic: Q(obs, action) through the fused LiquidNCPNetwork with one output (continuous.py:146-191).  Not used by the
    NeuroFlow agents (they take the NCP critics); kept for users who build their own."""

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, *, device: torch.device | None = None) -> None:
        super().__init__()
        self.device = device
        self.ncp = LiquidNCPNetwork(num_obs + num_actions, n_neurons, 1, device=device).to(device)

    def forward(self, obs: torch.Tensor, actions: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        q_values, new_hidden = self.ncp(torch.cat([obs, actions], dim=-1), hidden)
        return q_values, new_hidden
