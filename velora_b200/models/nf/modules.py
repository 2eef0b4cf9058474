"""Agent modules of NeuroFlowCT (mirror of velora/models/nf/modules.py: ActorModule :41, CriticModule :320,
EntropyModule :646): network + optimizer bundles.  Networks are torch.jit.script-ed after construction exactly like
the reference (modules.py:107, 388-392) - the fused forward survives because it sits behind @torch.jit.ignore."""
from copy import deepcopy
from typing import Tuple, Type

import torch
import torch.nn as nn
from torch import optim

from velora_b200.models.sac.continuous import SACActor, SACCriticNCP
from velora_b200.models.sac.discrete import SACActorDiscrete, SACCriticNCPDiscrete
from velora_b200 import dataparallel as _dp
from velora_b200.utils.torch import soft_update


class ActorModule:
    def __init__(self, state_dim: int, n_neurons: int, action_dim: int, action_scale: torch.Tensor, action_bias: torch.Tensor,
                 *, log_std_min: float = -5, log_std_max: float = 2, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, device: torch.device | None = None, optim_kwargs: dict | None = None, fused_head: bool = False):
        self.device = device
        self.network = SACActor(state_dim, n_neurons, action_dim, action_scale, action_bias, log_std_min=log_std_min,
                                log_std_max=log_std_max, device=device, fused_head=fused_head).to(device)
        self.hidden_size = self.network.ncp.hidden_size
        self.optim = optim(self.network.parameters(), lr=lr, **(optim_kwargs or {}))
        self.active_params = self.network.ncp.active_params
        self.total_params = self.network.ncp.total_params
        self.network = torch.jit.script(self.network)

    def gradient_step(self, loss: torch.Tensor) -> None:
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()

    def predict(self, obs: torch.Tensor, hidden: torch.Tensor | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.network.predict(obs, hidden)

    def forward(self, obs: torch.Tensor, hidden: torch.Tensor | None = None):
        return self.network(obs, hidden)

    def state_dict(self):
        """modules.py:165-169"""
        return {"actor": self.network.state_dict(), "actor_optim": self.optim.state_dict()}

    def load_state_dict(self, state_dict) -> None:
        self.network.load_state_dict(state_dict["actor"])
        self.optim.load_state_dict(state_dict["actor_optim"])

    def eval_mode(self) -> None:
        """modules.py:157-159"""
        self.network.eval()

    def train_mode(self) -> None:
        """modules.py:161-163"""
        self.network.train()


def _side_stream(fused: bool, device) -> "torch.cuda.Stream | None":
    dev = torch.device(device) if device is not None else None
    return torch.cuda.Stream(dev) if fused and dev is not None and dev.type == "cuda" else None


def _concurrently(side, f1, f2):
    """f1() on the current stream, f2() on `side`: the twin critics share no state, and at batch 64 each of their kernels fills a few SMs
    only, so the pair (and, since autograd replays a node on its forward stream, the pair of backward passes) overlaps instead of
    queueing.  Works inside a CUDA-graph capture: the two waits become graph dependencies."""
    from velora_b200 import dataparallel as _dp

    # data-parallel: every rank must issue its gradient exchanges in one order on one stream (the one-shot kernel numbers its calls on
    # the device), so the twins stay on the current stream there
    if side is None or _dp.is_enabled():
        return f1(), f2()
    cur = torch.cuda.current_stream(side.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        b = f2()
    a = f1()
    cur.wait_stream(side)
    for t in (b if isinstance(b, tuple) else (b,)):
        if isinstance(t, torch.Tensor):
            t.record_stream(cur)
    return a, b


class CriticModule:
    def __init__(self, state_dim: int, n_neurons: int, action_dim: int, *, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, tau: float = 0.005, device: torch.device | None = None, optim_kwargs: dict | None = None,
                 fused_update: bool = False):
        self.tau = tau
        self.fused_update = fused_update
        self.device = device
        self._side = _side_stream(fused_update, device)
        self.network1 = SACCriticNCP(state_dim, n_neurons, action_dim, device=device).to(device)
        self.network2 = SACCriticNCP(state_dim, n_neurons, action_dim, device=device).to(device)
        self.target1 = deepcopy(self.network1)
        self.target2 = deepcopy(self.network2)
        self.optim1 = optim(self.network1.parameters(), lr=lr, **(optim_kwargs or {}))
        self.optim2 = optim(self.network2.parameters(), lr=lr, **(optim_kwargs or {}))
        self.active_params = self.network1.ncp.active_params + self.network2.ncp.active_params
        self.total_params = self.network1.ncp.total_params + self.network2.ncp.total_params
        self.network1 = torch.jit.script(self.network1)
        self.network2 = torch.jit.script(self.network2)
        self.target1 = torch.jit.script(self.target1)
        self.target2 = torch.jit.script(self.target2)

    def update_targets(self) -> None:
        if self.fused_update:      # one launch per network pair (velora_b200/optim.py)
            from velora_b200.optim import soft_update as fused_soft_update

            fused_soft_update(self.network1, self.target1, tau=self.tau)
            fused_soft_update(self.network2, self.target2, tau=self.tau)
            return
        soft_update(self.network1, self.target1, tau=self.tau)
        soft_update(self.network2, self.target2, tau=self.tau)

    def gradient_step(self, c1_loss: torch.Tensor, c2_loss: torch.Tensor) -> None:
        if c1_loss.grad_fn is not None and c1_loss.grad_fn is c2_loss.grad_fn:
            # both losses come out of ONE fused node (functional.SacCriticLossFn): one backward pass; the two critics share no
            # parameters, so this equals the reference's two separate passes (modules.py:401-415)
            self.optim1.zero_grad()
            self.optim2.zero_grad()
            torch.autograd.backward([c1_loss, c2_loss])
            self.optim1.step()
            self.optim2.step()
            return
        self.optim1.zero_grad()
        c1_loss.backward()
        self.optim1.step()
        self.optim2.zero_grad()
        c2_loss.backward()
        self.optim2.step()

    def state_dict(self):
        """modules.py:452-460"""
        return {"critic": self.network1.state_dict(), "critic2": self.network2.state_dict(), "critic_target": self.target1.state_dict(),
                "critic2_target": self.target2.state_dict(), "critic_optim": self.optim1.state_dict(), "critic2_optim": self.optim2.state_dict()}

    def load_state_dict(self, state_dict) -> None:
        self.network1.load_state_dict(state_dict["critic"])
        self.network2.load_state_dict(state_dict["critic2"])
        self.target1.load_state_dict(state_dict["critic_target"])
        self.target2.load_state_dict(state_dict["critic2_target"])
        self.optim1.load_state_dict(state_dict["critic_optim"])
        self.optim2.load_state_dict(state_dict["critic2_optim"])

    def predict(self, obs: torch.Tensor, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _concurrently(self._side, lambda: self.network1(obs, actions), lambda: self.network2(obs, actions))

    def target_predict(self, obs: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        return torch.min(*_concurrently(self._side, lambda: self.target1(obs, actions), lambda: self.target2(obs, actions)))


class EntropyModule:
    """Automatic entropy tuning: one scalar log_alpha (modules.py:646-717)."""

    def __init__(self, action_dim: int, *, initial_alpha: float = 1.0, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, device: torch.device | None = None, optim_kwargs: dict | None = None):
        self.target = -action_dim
        self.log_alpha = nn.Parameter(torch.tensor(initial_alpha, device=device).log())
        self.optim = optim([self.log_alpha], lr=lr, **(optim_kwargs or {}))

    @property
    def alpha(self) -> torch.Tensor:
        return self.log_alpha.exp()

    def compute_loss(self, log_probs: torch.Tensor) -> torch.Tensor:
        entropy = (log_probs + self.target).detach()
        return -(self.log_alpha * entropy).mean()

    def state_dict(self):
        """modules.py:732-739: the reference stores the entropy optimizer only (log_alpha itself is not in the checkpoint)."""
        return {"entropy_optim": self.optim.state_dict()}

    def load_state_dict(self, state_dict) -> None:
        self.optim.load_state_dict(state_dict["entropy_optim"])

    def gradient_step(self, loss: torch.Tensor) -> None:
        self.optim.zero_grad()
        loss.backward()
        _dp.maybe_allreduce(self.log_alpha.grad)      # data-parallel: the loss is a mean over this rank's shard
        self.optim.step()


class ActorModuleDiscrete:
    """modules.py:190-318"""

    def __init__(self, state_dim: int, n_neurons: int, action_dim: int, *, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, device: torch.device | None = None, optim_kwargs: dict | None = None, fused_head: bool = False):
        self.device = device
        self.network = SACActorDiscrete(state_dim, n_neurons, action_dim, device=device, fused_head=fused_head).to(device)
        self.hidden_size = self.network.ncp.hidden_size
        self.optim = optim(self.network.parameters(), lr=lr, **(optim_kwargs or {}))
        self.active_params = self.network.ncp.active_params
        self.total_params = self.network.ncp.total_params
        self.network = torch.jit.script(self.network)

    def gradient_step(self, loss: torch.Tensor) -> None:
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()

    def predict(self, obs: torch.Tensor, hidden: torch.Tensor | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.network.predict(obs, hidden)

    def forward(self, obs: torch.Tensor, hidden: torch.Tensor | None = None):
        return self.network(obs, hidden)

    def state_dict(self):
        """modules.py:165-169"""
        return {"actor": self.network.state_dict(), "actor_optim": self.optim.state_dict()}

    def load_state_dict(self, state_dict) -> None:
        self.network.load_state_dict(state_dict["actor"])
        self.optim.load_state_dict(state_dict["actor_optim"])

    def eval_mode(self) -> None:
        self.network.eval()

    def train_mode(self) -> None:
        self.network.train()


class CriticModuleDiscrete:
    """modules.py:485-644"""

    def __init__(self, state_dim: int, n_neurons: int, action_dim: int, *, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, tau: float = 0.005, device: torch.device | None = None, optim_kwargs: dict | None = None,
                 fused_update: bool = False):
        self.tau = tau
        self.fused_update = fused_update
        self.device = device
        self._side = _side_stream(fused_update, device)
        self.network1 = SACCriticNCPDiscrete(state_dim, n_neurons, action_dim, device=device).to(device)
        self.network2 = SACCriticNCPDiscrete(state_dim, n_neurons, action_dim, device=device).to(device)
        self.target1 = deepcopy(self.network1)
        self.target2 = deepcopy(self.network2)
        self.optim1 = optim(self.network1.parameters(), lr=lr, **(optim_kwargs or {}))
        self.optim2 = optim(self.network2.parameters(), lr=lr, **(optim_kwargs or {}))
        self.active_params = self.network1.ncp.active_params + self.network2.ncp.active_params
        self.total_params = self.network1.ncp.total_params + self.network2.ncp.total_params
        self.network1 = torch.jit.script(self.network1)
        self.network2 = torch.jit.script(self.network2)
        self.target1 = torch.jit.script(self.target1)
        self.target2 = torch.jit.script(self.target2)

    update_targets = CriticModule.update_targets
    gradient_step = CriticModule.gradient_step
    state_dict = CriticModule.state_dict
    load_state_dict = CriticModule.load_state_dict

    def predict(self, obs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _concurrently(self._side, lambda: self.network1(obs), lambda: self.network2(obs))

    def target_predict(self, obs: torch.Tensor) -> torch.Tensor:
        return torch.min(*self.target_pair(obs))

    def target_pair(self, obs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Both target critics' Q(s, .) (the fused critic loss takes their minimum itself)."""
        return _concurrently(self._side, lambda: self.target1(obs), lambda: self.target2(obs))


class EntropyModuleDiscrete:
    """modules.py:751-855 (target entropy 0.98 log(1 / |A|))"""

    def __init__(self, action_dim: int, *, initial_alpha: float = 1.0, optim: Type[optim.Optimizer] = optim.Adam,
                 lr: float = 3e-4, device: torch.device | None = None, optim_kwargs: dict | None = None):
        self.target = 0.98 * torch.tensor(1 / action_dim, device=device).log()
        self.target_value = float(self.target)       # host copy for the fused loss kernel (no device read inside a graph capture)
        self.log_alpha = nn.Parameter(torch.tensor(initial_alpha, device=device).log())
        self.optim = optim([self.log_alpha], lr=lr, **(optim_kwargs or {}))

    @property
    def alpha(self) -> torch.Tensor:
        return self.log_alpha.exp()

    def compute_loss(self, probs: torch.Tensor, log_probs: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            entropy_mean = -torch.sum(probs * log_probs, dim=1).mean()
        return self.log_alpha * (entropy_mean - self.target)

    def state_dict(self):
        """modules.py:732-739: the reference stores the entropy optimizer only (log_alpha itself is not in the checkpoint)."""
        return {"entropy_optim": self.optim.state_dict()}

    def load_state_dict(self, state_dict) -> None:
        self.optim.load_state_dict(state_dict["entropy_optim"])

    def gradient_step(self, loss: torch.Tensor) -> None:
        self.optim.zero_grad()
        loss.backward()
        _dp.maybe_allreduce(self.log_alpha.grad)      # data-parallel: entropy_mean is a mean over this rank's shard
        self.optim.step()
