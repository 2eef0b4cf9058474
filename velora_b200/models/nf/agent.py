"""NeuroFlowCTCore: the environment-free core of the reference's NeuroFlowCT agent (velora/models/nf/agent.py:38-255)
- networks, replay buffer, `_update_critics` and `_train_step` - so that BASELINE configs 1/2/5 can be driven on a
synthetic buffer.  The environment loop, callbacks, metrics database and checkpoint I/O of the reference are out of
scope; gymnasium is not required.

Construction order (set_seed -> actor -> critics -> entropy -> buffer) is the reference's (agent.py:98-152), so a given
seed yields identical masks and initial parameters.
"""
from typing import Dict, Type

import torch
import torch.nn as nn
from torch import optim

from velora_b200.buffer import BatchExperience, ReplayBuffer
from velora_b200.models.nf.modules import ActorModule, CriticModule, EntropyModule
from velora_b200.utils.core import set_seed


class NeuroFlowCTCore:
    def __init__(self, state_dim: int, action_dim: int, actor_neurons: int, critic_neurons: int, action_scale: torch.Tensor,
                 action_bias: torch.Tensor, *, optim: Type[optim.Optimizer] = optim.Adam, buffer_size: int = 1_000_000,
                 actor_lr: float = 3e-4, critic_lr: float = 3e-4, alpha_lr: float = 3e-4, initial_alpha: float = 1.0,
                 log_std: tuple = (-5, 2), tau: float = 0.005, gamma: float = 0.99, device: torch.device | None = None,
                 seed: int | None = None, capturable: bool = False) -> None:
        self.device = device
        self.seed = set_seed(seed)
        self.state_dim, self.action_dim = state_dim, action_dim
        self.gamma, self.tau = gamma, tau
        # capturable=True keeps the optimizer state on the device so that a whole train step can be captured in a CUDA graph
        okw = {"capturable": True} if capturable else None
        self.actor = ActorModule(state_dim, actor_neurons, action_dim, action_scale.to(device), action_bias.to(device),
                                 log_std_min=log_std[0], log_std_max=log_std[1], optim=optim, lr=actor_lr, device=device, optim_kwargs=okw)
        self.critic = CriticModule(state_dim, critic_neurons, action_dim, optim=optim, lr=critic_lr, tau=tau, device=device, optim_kwargs=okw)
        self.hidden_dim = self.actor.hidden_size
        self.entropy = EntropyModule(action_dim, initial_alpha=initial_alpha, optim=optim, lr=alpha_lr, device=device, optim_kwargs=okw)
        self.loss = nn.MSELoss()
        self.buffer = ReplayBuffer(buffer_size, state_dim, action_dim, self.actor.hidden_size, device=device)

    def _update_critics(self, batch: BatchExperience) -> torch.Tensor:
        """agent.py:182-214"""
        with torch.no_grad():
            next_actions, next_log_probs, _ = self.actor.forward(batch.next_states, batch.hiddens)
            next_q = self.critic.target_predict(batch.next_states, next_actions)
            next_q = next_q - self.entropy.alpha * next_log_probs
            target_q = batch.rewards + (1 - batch.dones) * self.gamma * next_q
        current_q1, current_q2 = self.critic.predict(batch.states, batch.actions)
        c1_loss = self.loss(current_q1, target_q)
        c2_loss = self.loss(current_q2, target_q)
        critic_loss = c1_loss + c2_loss
        self.critic.gradient_step(c1_loss, c2_loss)
        return critic_loss

    def _train_step(self, batch_size: int) -> Dict[str, torch.Tensor]:
        """agent.py:216-255"""
        return self._train_on(self.buffer.sample(batch_size))

    def _train_on(self, batch: BatchExperience) -> Dict[str, torch.Tensor]:
        critic_loss = self._update_critics(batch)
        actions, log_probs, _ = self.actor.forward(batch.states, batch.hiddens)
        q1, q2 = self.critic.predict(batch.states, actions)
        next_q = torch.min(q1, q2)
        actor_loss = (self.entropy.alpha * log_probs - next_q).mean()
        entropy_loss = self.entropy.compute_loss(log_probs)
        self.actor.gradient_step(actor_loss)
        self.entropy.gradient_step(entropy_loss)
        self.critic.update_targets()
        return {"critic": critic_loss.detach(), "actor": actor_loss.detach(), "entropy": entropy_loss.detach()}

    def capture_train_step(self, batch_size: int, warmup: int = 3):
        """Captures one `_train_step(batch_size)` - index draw, fused gather, every network forward/backward, the SAC head
        math, the four optimizer steps and the target update - into ONE CUDA graph and returns a callable that replays it.

        At batch 64 a train step is ~250 tiny launches, i.e. pure launch/dispatch latency (the reference spends 5.9 ms per
        step on the CPU for the same reason); replaying a graph removes the per-launch host cost.
        Requires `capturable=True` at construction.  The returned losses are the graph's static output tensors.
        """
        if self.device is None or torch.device(self.device).type != "cuda":
            raise RuntimeError("capture_train_step needs a CUDA device")
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._train_step(batch_size)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._train_step(batch_size)

        def replay():
            graph.replay()
            return out

        replay.graph = graph
        return replay
