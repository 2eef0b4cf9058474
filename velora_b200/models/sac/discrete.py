"""Discrete-action SAC heads over the fused networks (mirror of velora/models/sac/discrete.py:10-184)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch.distributions import Categorical

from velora_b200.models.lnn.ncp import LiquidNCPNetwork, NCPNetwork

_TORCH_CATEGORICAL_SAMPLE = Categorical.sample      # the fused head steps aside when somebody (a test replaying recorded draws) patched it


class SACActorDiscrete(nn.Module):
    """Liquid actor emitting a categorical distribution over actions (discrete.py:10-96)."""

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, *, device: torch.device | None = None,
                 fused_head: bool = False) -> None:
        super().__init__()
        self.device = device
        self.ncp = LiquidNCPNetwork(num_obs, n_neurons, num_actions, device=device).to(device)
        self.num_actions = num_actions
        self.softmax = nn.Softmax(dim=-1)
        self.fused_head = fused_head

    @torch.jit.ignore
    def _head(self, logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """logits -> (actions, probs, log_prob of the actions): softmax + `get_sample`, or with `fused_head` one kernel
        (functional.CategoricalHeadFn) fed with the Exp(1) noise torch.multinomial would draw - same generator consumption, same draw."""
        from velora_b200 import dataparallel as _dp

        if (self.fused_head and logits.is_cuda and logits.dim() == 2 and logits.shape[1] <= 64
                and Categorical.sample is _TORCH_CATEGORICAL_SAMPLE and not (_dp.is_enabled() and _dp.world_size() > 1)):
            from velora_b200 import functional as VF

            noise = torch.empty_like(logits, requires_grad=False).exponential_(1.0)
            actions, probs, log_prob = VF.CategoricalHeadFn.apply(logits, noise)
            return actions, probs, log_prob
        probs = torch.softmax(logits, dim=-1)        # (self.softmax is a scripted submodule here, not callable from an ignored method)
        # (forward no longer calls get_sample, so the scripted module does not bind it: call it through the class)
        actions, log_prob = SACActorDiscrete.get_sample(self, probs)
        return actions, probs, log_prob

    @torch.jit.ignore
    def get_sample(self, probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Sampled actions and their log-probabilities (discrete.py:41-57)."""
        # argument validation reads a device flag back to the host, which is illegal while a CUDA graph is being captured
        validate = not (probs.is_cuda and torch.cuda.is_current_stream_capturing())
        dist = Categorical(probs=probs, validate_args=validate)
        from velora_b200 import dataparallel as _dp

        if _dp.is_enabled() and _dp.world_size() > 1 and probs.dim() == 2:
            # data-parallel: every rank samples the GLOBAL batch from the gathered probabilities (identical generators => identical draws,
            # and the generator advances exactly like the single-process step) and keeps its own rows
            import torch.distributed as tdist

            n = probs.shape[0]
            parts = [torch.empty_like(probs) for _ in range(_dp.world_size())]
            tdist.all_gather(parts, probs.detach().contiguous(), group=_dp._state["group"])
            full = Categorical(probs=torch.cat(parts), validate_args=validate).sample()
            actions = full[_dp.rank() * n:(_dp.rank() + 1) * n]
        else:
            actions = dist.sample()
        return actions, dist.log_prob(actions).unsqueeze(-1)

    @torch.jit.ignore
    def predict(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Greedy action (discrete.py:59-78)."""
        logits, new_hidden = self.ncp(obs, hidden)
        return torch.argmax(self.softmax(logits), dim=-1), new_hidden

    def forward(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """(obs, hidden) -> (actions, probs, log_prob of the actions, new hidden)   (discrete.py:80-101)"""
        logits, new_hidden = self.ncp(obs, hidden)
        actions, probs, log_prob = self._head(logits)
        return actions, probs, log_prob, new_hidden


class SACCriticNCPDiscrete(nn.Module):
    """NCP critic: Q(obs, .) for every action through the fused three-layer NCPNetwork (discrete.py:147-184)."""

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, *, device: torch.device | None = None) -> None:
        super().__init__()
        self.device = device
        self.ncp = NCPNetwork(num_obs, n_neurons, num_actions, device=device).to(device)

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.ncp(obs)


class SACCriticDiscrete(nn.Module):
    """Liquid NCP critic: Q(obs, .) for every action through the fused LiquidNCPNetwork (discrete.py:103-145)."""

    def __init__(self, num_obs: int, n_neurons: int, num_actions: int, *, device: torch.device | None = None) -> None:
        super().__init__()
        self.device = device
        self.ncp = LiquidNCPNetwork(num_obs, n_neurons, num_actions, device=device).to(device)

    def forward(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        q_values, new_hidden = self.ncp(obs, hidden)
        return q_values, new_hidden
