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
from velora_b200.models.nf.modules import (ActorModule, ActorModuleDiscrete, CriticModule, CriticModuleDiscrete, EntropyModule,
                                           EntropyModuleDiscrete)
from velora_b200.utils.core import set_seed


class NeuroFlowCTCore:
    def __init__(self, state_dim: int, action_dim: int, actor_neurons: int, critic_neurons: int, action_scale: torch.Tensor,
                 action_bias: torch.Tensor, *, optim: Type[optim.Optimizer] = optim.Adam, buffer_size: int = 1_000_000,
                 actor_lr: float = 3e-4, critic_lr: float = 3e-4, alpha_lr: float = 3e-4, initial_alpha: float = 1.0,
                 log_std: tuple = (-5, 2), tau: float = 0.005, gamma: float = 0.99, device: torch.device | None = None,
                 seed: int | None = None, capturable: bool = False, fused: bool = False) -> None:
        self.device = device
        self.seed = set_seed(seed)
        self.state_dim, self.action_dim = state_dim, action_dim
        self.gamma, self.tau = gamma, tau
        # fused=True also takes the loss arithmetic around the networks (critic target + both MSE losses, actor loss, entropy loss) as
        # one launch forward / one backward each (velora_b200/csrc/sac_loss.cu) instead of ~40 framework kernels per step
        self.fused_losses = bool(fused)
        self.graph_capable = bool(fused or capturable)      # optimizer state on the device: a whole train step can be captured
        # capturable=True keeps the optimizer state on the device so that a whole train step can be captured in a CUDA graph
        okw = {"capturable": True} if capturable else None
        if fused:   # one-launch Adam and Polyak updates (velora_b200/optim.py); same optimizer state layout as Adam(capturable=True)
            from velora_b200.optim import FusedAdam

            if optim is not torch.optim.Adam:
                raise ValueError("fused=True replaces torch.optim.Adam only")
            optim, okw = FusedAdam, None
        self.actor = ActorModule(state_dim, actor_neurons, action_dim, action_scale.to(device), action_bias.to(device),
                                 log_std_min=log_std[0], log_std_max=log_std[1], optim=optim, lr=actor_lr, device=device, optim_kwargs=okw,
                                 fused_head=fused)
        self.critic = CriticModule(state_dim, critic_neurons, action_dim, optim=optim, lr=critic_lr, tau=tau, device=device, optim_kwargs=okw,
                                   fused_update=fused)
        self.hidden_dim = self.actor.hidden_size
        self.entropy = EntropyModule(action_dim, initial_alpha=initial_alpha, optim=optim, lr=alpha_lr, device=device, optim_kwargs=okw)
        self.loss = nn.MSELoss()
        self.buffer = ReplayBuffer(buffer_size, state_dim, action_dim, self.actor.hidden_size, device=device)

    def _update_critics(self, batch: BatchExperience) -> torch.Tensor:
        """agent.py:182-214"""
        if getattr(self, "fused_losses", False) and batch.states.is_cuda:
            from velora_b200 import functional as VF

            with torch.no_grad():
                next_actions, next_log_probs, _ = self.actor.forward(batch.next_states, batch.hiddens)
                q1t, q2t = self.critic.target1(batch.next_states, next_actions), self.critic.target2(batch.next_states, next_actions)
            current_q1, current_q2 = self.critic.predict(batch.states, batch.actions)
            c1_loss, c2_loss = VF.SacCriticLossFn.apply(current_q1, current_q2, q1t, q2t, next_log_probs, batch.rewards, batch.dones,
                                                        self.entropy.log_alpha, self.gamma)
            critic_loss = c1_loss + c2_loss
            self.critic.gradient_step(c1_loss, c2_loss)
            return critic_loss
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
        fused = getattr(self, "fused_losses", False) and actions.is_cuda
        if fused:
            from velora_b200 import functional as VF

            # the critics' own gradients of the ACTOR loss are never used (the next critic update starts with zero_grad): dL/d(action) only
            with VF.frozen_parameters():
                q1, q2 = self.critic.predict(batch.states, actions)
        else:
            q1, q2 = self.critic.predict(batch.states, actions)
        if fused:
            actor_loss = VF.SacActorLossFn.apply(log_probs, q1, q2, self.entropy.log_alpha)
            entropy_loss = VF.SacEntropyLossFn.apply(self.entropy.log_alpha, log_probs.detach(), float(self.entropy.target))
        else:
            next_q = torch.min(q1, q2)
            actor_loss = (self.entropy.alpha * log_probs - next_q).mean()
            entropy_loss = self.entropy.compute_loss(log_probs)
        self.actor.gradient_step(actor_loss)
        self.entropy.gradient_step(entropy_loss)
        self.critic.update_targets()
        return {"critic": critic_loss.detach(), "actor": actor_loss.detach(), "entropy": entropy_loss.detach()}

    def state_dict(self):
        """{"modules": ..., "optimizers": ...} (velora/models/base.py:296-312)."""
        from velora_b200.utils.restore import _agent_state

        return _agent_state(self)

    def save(self, dirpath, *, buffer: bool = False, force: bool = False) -> None:
        """Writes model / optimizer (/ buffer) state in the reference's checkpoint format (velora/utils/restore.py:107-196)."""
        from velora_b200.utils.restore import save_model

        save_model(self, dirpath, buffer=buffer, force=force)

    def load_state(self, dirpath, *, buffer: bool = False):
        """Restores a checkpoint written by `save` or by the reference's `save_model` into this agent."""
        from velora_b200.utils.restore import load_state

        return load_state(self, dirpath, buffer=buffer)

    def predict(self, state: torch.Tensor, hidden: torch.Tensor | None = None, *, train_mode: bool = False):
        """Action for the given state (agent.py:356-388): deterministic `tanh(mean)` when `train_mode` is False, a sampled
        action otherwise; returns (action, new hidden)."""
        self.actor.eval_mode()
        with torch.no_grad():
            state = state.unsqueeze(0) if state.dim() < 2 else state
            if not train_mode:
                action, hidden = self.actor.predict(state, hidden)
            else:
                action, _, hidden = self.actor.forward(state, hidden)
        self.actor.train_mode()
        return action, hidden

    def capture_predict(self, num_envs: int = 1, *, train_mode: bool = False, warmup: int = 3):
        """The per-environment-step `predict` as ONE CUDA graph (pack + forward kernel + the policy head) with static buffers.

        Returns `act(state [num_envs, S] or [S], hidden [num_envs, H] | None) -> (action, hidden)`; the outputs are the
        graph's static tensors (overwritten by the next call).  The graph reads the live parameters, so it stays valid
        across `_train_step` / graph replays of the train step.  An eager batch-1 `predict` is ~100 us of Python dispatch
        and ~270 us on the device timeline; the replay is two tiny copies and one graph launch.  `warmup` eager calls run first
        (with `train_mode` they consume the generator: a caller that has already acted eagerly a few times passes 0).
        """
        if self.device is None or torch.device(self.device).type != "cuda":
            raise RuntimeError("capture_predict needs a CUDA device")
        dev = self.device
        s_state = torch.zeros(num_envs, self.state_dim, device=dev)
        s_hidden = torch.zeros(num_envs, self.hidden_dim, device=dev)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.predict(s_state, s_hidden, train_mode=train_mode)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.predict(s_state, s_hidden, train_mode=train_mode)

        def act(state: torch.Tensor, hidden: torch.Tensor | None = None):
            s_state.copy_(state.reshape(num_envs, self.state_dim), non_blocking=True)
            if hidden is None:
                s_hidden.zero_()
            else:
                s_hidden.copy_(hidden.reshape(num_envs, self.hidden_dim), non_blocking=True)
            graph.replay()
            return out

        act.graph = graph
        return act

    def capture_train_step(self, batch_size: int, warmup: int = 3):
        """Captures one `_train_step(batch_size)` - index draw, fused gather, every network forward/backward, the SAC head
        math, the four optimizer steps and the target update - into ONE CUDA graph and returns a callable that replays it.

        At batch 64 a train step is ~250 tiny launches, i.e. pure launch/dispatch latency (the reference spends 5.9 ms per
        step on the CPU for the same reason); replaying a graph removes the per-launch host cost.
        Requires `capturable=True` at construction and a FULL buffer (the sampling range is a host integer that the capture
        freezes).  The returned losses are the graph's static output tensors.
        """
        if self.device is None or torch.device(self.device).type != "cuda":
            raise RuntimeError("capture_train_step needs a CUDA device")
        if self.buffer.size != self.buffer.capacity:
            # `buffer.sample` draws torch.randint(0, size) and passes n_rows = size to the gather; both are HOST integers that a
            # capture freezes, so a graph captured on a partly filled buffer would never see transitions added later
            raise RuntimeError(
                f"capture_train_step: the replay buffer holds {self.buffer.size} of {self.buffer.capacity} transitions; a captured "
                "graph freezes the sampling range, so capture only once the buffer is full (or train eagerly until then)")
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


    def capture_train_on_indices(self, batch_size: int):
        """`_train_step` minus the index draw as one CUDA graph: returns (indices, replay) - write the minibatch's int64 row indices into
        the static `indices` tensor, then `replay()`.  Unlike `capture_train_step` this graph stays valid while the buffer FILLS: the
        sampling range lives outside it (`indices.copy_(torch.randint(0, len(buffer), (batch_size,), device=...))` per step is exactly
        what `ReplayBuffer.sample` draws, so a replay loop consumes the generator like the eager loop).  As for every capture of a
        backward pass, a few train steps must have run on a SIDE stream first (`warm_train_step`): autograd accumulates a parameter's
        gradient on the stream that created it, and a capture cannot synchronise with the legacy default stream."""
        if self.device is None or torch.device(self.device).type != "cuda":
            raise RuntimeError("capture_train_on_indices needs a CUDA device")
        indices = torch.zeros(batch_size, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._train_on(self.buffer.gather(indices, n_rows=self.buffer.capacity))

        def replay():
            graph.replay()
            return out

        replay.graph = graph
        return indices, replay

    def warm_train_step(self, batch_size: int):
        """One eager `_train_step` on a side stream (the warm-up a later graph capture needs, see `capture_train_on_indices`)."""
        side = getattr(self, "_warm_stream", None)
        if side is None:
            side = self._warm_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            out = self._train_step(batch_size)
        cur.wait_stream(side)
        return out


class NeuroFlowCore(NeuroFlowCTCore):
    """The environment-free core of the reference's discrete agent NeuroFlow (velora/models/nf/agent.py:411-650): liquid
    actor with a softmax head, two NCP critics emitting Q for every action, discrete-SAC entropy tuning.  BASELINE config 1
    (`NeuroFlow("CartPole-v1", 20, 128)`) is `NeuroFlowCore(4, 2, 20, 128)`.  Shares `predict`, `capture_train_step` and
    `capture_predict` with the continuous core."""

    def __init__(self, state_dim: int, action_dim: int, actor_neurons: int, critic_neurons: int, *,
                 optim: Type[optim.Optimizer] = optim.Adam, buffer_size: int = 1_000_000, actor_lr: float = 3e-4,
                 critic_lr: float = 3e-4, alpha_lr: float = 3e-4, initial_alpha: float = 1.0, tau: float = 0.005,
                 gamma: float = 0.99, device: torch.device | None = None, seed: int | None = None, capturable: bool = False,
                 fused: bool = False) -> None:
        self.device = device
        self.seed = set_seed(seed)
        self.state_dim, self.action_dim = state_dim, action_dim
        self.gamma, self.tau = gamma, tau
        okw = {"capturable": True} if capturable else None
        if fused:
            from velora_b200.optim import FusedAdam

            if optim is not torch.optim.Adam:
                raise ValueError("fused=True replaces torch.optim.Adam only")
            optim, okw = FusedAdam, None
        self.actor = ActorModuleDiscrete(state_dim, actor_neurons, action_dim, optim=optim, lr=actor_lr, device=device, optim_kwargs=okw,
                                         fused_head=fused)
        self.critic = CriticModuleDiscrete(state_dim, critic_neurons, action_dim, optim=optim, lr=critic_lr, tau=tau, device=device,
                                           optim_kwargs=okw, fused_update=fused)
        self.hidden_dim = self.actor.hidden_size
        self.entropy = EntropyModuleDiscrete(action_dim, initial_alpha=initial_alpha, optim=optim, lr=alpha_lr, device=device, optim_kwargs=okw)
        self.loss = nn.MSELoss()
        self.fused_losses = bool(fused)
        self.graph_capable = bool(fused or capturable)
        self.buffer = ReplayBuffer(buffer_size, state_dim, 1, self.actor.hidden_size, device=device)   # action index stored as one float

    def _update_critics(self, batch: BatchExperience) -> torch.Tensor:
        """agent.py:549-590"""
        if getattr(self, "fused_losses", False) and batch.states.is_cuda:
            from velora_b200 import functional as VF

            with torch.no_grad():
                _, next_probs, _, _ = self.actor.forward(batch.next_states, batch.hiddens)
                t1, t2 = self.critic.target_pair(batch.next_states)
            q1, q2 = self.critic.predict(batch.states)
            c1_loss, c2_loss = VF.SacdCriticLossFn.apply(q1, q2, batch.actions, next_probs, t1, t2, batch.rewards, batch.dones,
                                                         self.entropy.log_alpha, self.gamma)
            critic_loss = c1_loss.detach() + c2_loss.detach()
            self.critic.gradient_step(c1_loss, c2_loss)
            return critic_loss
        with torch.no_grad():
            _, next_probs, _, _ = self.actor.forward(batch.next_states, batch.hiddens)
            next_log_probs = torch.log(next_probs + 1e-8)
            min_next_q = self.critic.target_predict(batch.next_states)
            next_q = next_probs * (min_next_q - self.entropy.alpha * next_log_probs)
            next_q = torch.sum(next_q, dim=-1, keepdim=True)
            target_q = batch.rewards + (1 - batch.dones) * self.gamma * next_q
        current_q1, current_q2 = self.critic.predict(batch.states)
        current_q1 = current_q1.gather(1, batch.actions.long())
        current_q2 = current_q2.gather(1, batch.actions.long())
        c1_loss = self.loss(current_q1, target_q)
        c2_loss = self.loss(current_q2, target_q)
        critic_loss = c1_loss + c2_loss
        self.critic.gradient_step(c1_loss, c2_loss)
        return critic_loss

    def _train_on(self, batch: BatchExperience) -> Dict[str, torch.Tensor]:
        """agent.py:592-636"""
        critic_loss = self._update_critics(batch)
        _, probs, log_probs, _ = self.actor.forward(batch.states, batch.hiddens)
        if getattr(self, "fused_losses", False) and probs.is_cuda:
            from velora_b200 import functional as VF

            # Q(s, .) does not depend on the actor: with the critics' parameters frozen there is no critic backward at all in the actor
            # update (the reference computes their gradients here and discards them at the next zero_grad)
            with VF.frozen_parameters():
                q1, q2 = self.critic.predict(batch.states)
        else:
            q1, q2 = self.critic.predict(batch.states)
        if getattr(self, "fused_losses", False) and probs.is_cuda:
            actor_loss = VF.SacdActorLossFn.apply(probs, log_probs, q1, q2, self.entropy.log_alpha)
            entropy_loss = VF.SacdEntropyLossFn.apply(self.entropy.log_alpha, probs.detach(), self.entropy.target_value)
        else:
            next_q = torch.min(q1, q2)
            actor_loss = probs * (self.entropy.alpha * log_probs - next_q)
            actor_loss = torch.sum(actor_loss, dim=-1, keepdim=False).mean()
            entropy_loss = self.entropy.compute_loss(probs, torch.log(probs + 1e-8))
        self.actor.gradient_step(actor_loss)
        self.entropy.gradient_step(entropy_loss)
        self.critic.update_targets()
        return {"critic": critic_loss.detach(), "actor": actor_loss.detach(), "entropy": entropy_loss.detach()}

    def predict(self, state: torch.Tensor, hidden: torch.Tensor | None = None, *, train_mode: bool = False):
        """agent.py:700-732: greedy action when `train_mode` is False, a sampled one otherwise."""
        self.actor.eval_mode()
        with torch.no_grad():
            state = state.unsqueeze(0) if state.dim() < 2 else state
            if not train_mode:
                action, hidden = self.actor.predict(state, hidden)
            else:
                action, _, _, hidden = self.actor.forward(state, hidden)
        self.actor.train_mode()
        return action, hidden
