"""NeuroFlowCT / NeuroFlow: the exported agent names of velora.models (velora/models/__init__.py:1-9), constructed from an
environment like the reference's (velora/models/nf/agent.py:49-180, 433-545; velora/models/base.py:143-198) and driving the
environment-free cores of agent.py on the B200 kernels.

What is kept: the constructor signature, how state / action dimensions and the action scale / bias are read from the environment,
the construction order under `set_seed`, `predict`, `_train_step`, and the episode loop of `train` (act -> env.step -> buffer.add ->
one train step per environment step, warm-up through `buffer.warm`-style random acting).  What is out of scope (SURVEY.md section 2,
rows 13-18): TrainHandler, callbacks, the metrics database, console display and checkpoint I/O - `train` returns the list of episode
returns instead of writing them to SQLite.

The environment is either a Gymnasium id (needs `gymnasium`, which this image does not have) or any object with the Gymnasium
calling convention: `reset(seed=None) -> (obs, info)`, `step(action) -> (obs, reward, terminated, truncated, info)`,
`observation_space.shape`, and `action_space` with `.n` (discrete) or `.low` / `.high` (box).  Observations may be numpy arrays
or tensors.
"""
from __future__ import annotations

from typing import Any, List, Tuple, Type

import numpy as np
import torch
from torch import optim as _optim

from velora_b200.models.nf.agent import NeuroFlowCore, NeuroFlowCTCore


def _make_env(env_id: Any):
    if not isinstance(env_id, str):
        return env_id
    try:
        import gymnasium as gym
    except ImportError as e:  # pragma: no cover - gymnasium is absent in the build image
        raise ImportError(f"constructing an agent from the environment id '{env_id}' needs gymnasium; pass an environment object "
                          "with the Gymnasium calling convention instead") from e
    return gym.make(env_id)


def _is_discrete(space: Any) -> bool:
    return hasattr(space, "n")


class _EnvLoop:
    """Environment plumbing shared by both agents: tensors in, tensors out, one train step per environment step."""

    env: Any
    device: Any

    def _obs(self, o: Any) -> torch.Tensor:
        return torch.as_tensor(np.asarray(o) if not torch.is_tensor(o) else o, dtype=torch.float32, device=self.device)

    def _env_action(self, action: torch.Tensor):
        a = action.detach().reshape(-1)
        if _is_discrete(self.env.action_space):
            return int(a[0].item())
        return a.cpu().numpy()

    def _act(self, state: torch.Tensor, hidden):
        """The stochastic action of one environment step.  In graph mode (see `train`) the first calls are eager, then `predict` is one
        graph replay (`capture_predict`); the capture draws nothing from the generator, so the loop's random stream is the eager one."""
        st = getattr(self, "_act_state", None)
        if st is None:
            return self.predict(state, hidden, train_mode=True)
        if st["act"] is None:
            if st["eager_left"] > 0:
                st["eager_left"] -= 1
                return self.predict(state, hidden, train_mode=True)
            st["act"] = self.capture_predict(1, train_mode=True, warmup=0)
        return st["act"](state, hidden)

    def _collect(self, n_samples: int) -> None:
        """Role of `ReplayBuffer.warm` (velora/buffer/replay.py:90-125) on the agent's own single environment: act stochastically
        with the current policy until the buffer holds `n_samples` transitions."""
        state, _ = self.env.reset(seed=self.seed)
        state, hidden = self._obs(state), None
        while len(self.buffer) < n_samples:
            action, new_hidden = self._act(state, hidden)
            nxt, reward, term, trunc, _ = self.env.step(self._env_action(action))
            nxt = self._obs(nxt)
            done = bool(term or trunc)
            self.buffer.add(state, action.reshape(-1).float(), float(reward), nxt, done, new_hidden.reshape(-1))
            state, hidden = nxt, new_hidden
            if done:
                state, _ = self.env.reset()
                state, hidden = self._obs(state), None

    def _graph_train_step(self, batch_size: int):
        """One train step of the episode loop on a graph-capable agent (`fused=True` or `capturable=True`, CUDA, no data parallelism): the
        first three steps run eagerly on a side stream (the warm-up a capture of a backward pass needs), then the step minus its index
        draw is ONE graph replay (`capture_train_on_indices`); the indices are drawn per step with the same `torch.randint` call as
        `ReplayBuffer.sample`, so the loop consumes the generator exactly like the eager loop.  A batch-64 eager step is ~4 ms of Python
        dispatch, a replay ~0.2 ms."""
        st = self._graph_state
        if st is None or st["batch"] != batch_size:
            st = self._graph_state = dict(batch=batch_size, eager_left=3, indices=None, replay=None)
        if st["replay"] is None:
            if st["eager_left"] > 0:
                st["eager_left"] -= 1
                return self.warm_train_step(batch_size)
            st["indices"], st["replay"] = self.capture_train_on_indices(batch_size)
        if len(self.buffer) < batch_size:
            raise ValueError(f"Buffer does not contain enough experiences. Available: {len(self.buffer)}, Requested: {batch_size}")
        st["indices"].copy_(torch.randint(0, self.buffer.size, (batch_size,), device=self.device))
        return st["replay"]()

    def train(self, batch_size: int, *, n_episodes: int = 10_000, max_steps: int = 1000, warmup_steps: int | None = None,
              on_episode=None, use_graph: bool | None = None) -> List[float]:
        """The episode loop of the reference's `train` (agent.py:258-355 / 638-697).  Returns the episode returns.  `use_graph`
        (default: whenever the agent can be captured) replays the train step as a CUDA graph, see `_graph_train_step`."""
        from velora_b200 import dataparallel as _dp

        capable = (self.device is not None and torch.device(self.device).type == "cuda" and getattr(self, "graph_capable", False)
                   and not _dp.is_enabled())
        if use_graph and not capable:
            raise RuntimeError("train(use_graph=True) needs a CUDA agent built with fused=True or capturable=True, without data parallelism")
        use_graph = capable if use_graph is None else use_graph
        self._graph_state = None
        self._act_state = dict(eager_left=3, act=None) if use_graph else None
        warmup_steps = batch_size * 2 if warmup_steps is None else warmup_steps
        if warmup_steps > 0:
            self._collect(max(warmup_steps, batch_size))
        step = (lambda: self._graph_train_step(batch_size)) if use_graph else (lambda: self._train_step(batch_size))
        returns: List[float] = []
        for ep in range(1, n_episodes + 1):
            state, _ = self.env.reset()
            state, hidden, ep_return = self._obs(state), None, 0.0
            for _ in range(max_steps):
                action, hidden = self._act(state, hidden)
                nxt, reward, term, trunc, _ = self.env.step(self._env_action(action))
                nxt = self._obs(nxt)
                done = bool(term or trunc)
                self.buffer.add(state, action.reshape(-1).float(), float(reward), nxt, done, hidden.reshape(-1))
                self.last_losses = step()
                ep_return += float(reward)
                state = nxt
                if done:
                    break
            returns.append(ep_return)
            if on_episode is not None:
                on_episode(ep, ep_return)
        return returns


class NeuroFlowCT(_EnvLoop, NeuroFlowCTCore):
    """Continuous-action agent: `NeuroFlowCT(env_id, actor_neurons, critic_neurons, ...)` (agent.py:49-66)."""

    def __init__(self, env_id: Any, actor_neurons: int, critic_neurons: int, *, optim: Type[_optim.Optimizer] = _optim.Adam,
                 buffer_size: int = 1_000_000, actor_lr: float = 3e-4, critic_lr: float = 3e-4, alpha_lr: float = 3e-4,
                 initial_alpha: float = 1.0, log_std: Tuple[float, float] = (-5, 2), tau: float = 0.005, gamma: float = 0.99,
                 device: torch.device | None = None, seed: int | None = None, capturable: bool = False, fused: bool = False) -> None:
        env = _make_env(env_id)
        if _is_discrete(env.action_space):
            raise EnvironmentError(f"Invalid '{env.action_space=}'. Must be a continuous (Box) space.")
        self.env = env
        high = torch.as_tensor(np.asarray(env.action_space.high), dtype=torch.float32)
        low = torch.as_tensor(np.asarray(env.action_space.low), dtype=torch.float32)
        NeuroFlowCTCore.__init__(
            self, int(env.observation_space.shape[-1]), int(high.shape[-1]), actor_neurons, critic_neurons,
            (high - low) / 2.0, (high + low) / 2.0, optim=optim, buffer_size=buffer_size, actor_lr=actor_lr, critic_lr=critic_lr,
            alpha_lr=alpha_lr, initial_alpha=initial_alpha, log_std=log_std, tau=tau, gamma=gamma, device=device, seed=seed,
            capturable=capturable, fused=fused)
        self.actor_neurons, self.critic_neurons, self.buffer_size = actor_neurons, critic_neurons, buffer_size
        self.active_params = self.actor.active_params + self.critic.active_params
        self.total_params = self.actor.total_params + self.critic.total_params


class NeuroFlow(_EnvLoop, NeuroFlowCore):
    """Discrete-action agent: `NeuroFlow(env_id, actor_neurons, critic_neurons, ...)` (agent.py:433-449)."""

    def __init__(self, env_id: Any, actor_neurons: int, critic_neurons: int, *, optim: Type[_optim.Optimizer] = _optim.Adam,
                 buffer_size: int = 1_000_000, actor_lr: float = 3e-4, critic_lr: float = 3e-4, alpha_lr: float = 3e-4,
                 initial_alpha: float = 1.0, tau: float = 0.005, gamma: float = 0.99, device: torch.device | None = None,
                 seed: int | None = None, capturable: bool = False, fused: bool = False) -> None:
        env = _make_env(env_id)
        if not _is_discrete(env.action_space):
            raise EnvironmentError(f"Invalid '{env.action_space=}'. Must be a discrete space.")
        self.env = env
        NeuroFlowCore.__init__(
            self, int(env.observation_space.shape[-1]), int(env.action_space.n), actor_neurons, critic_neurons, optim=optim,
            buffer_size=buffer_size, actor_lr=actor_lr, critic_lr=critic_lr, alpha_lr=alpha_lr, initial_alpha=initial_alpha, tau=tau,
            gamma=gamma, device=device, seed=seed, capturable=capturable, fused=fused)
        self.actor_neurons, self.critic_neurons, self.buffer_size = actor_neurons, critic_neurons, buffer_size
        self.active_params = self.actor.active_params + self.critic.active_params
        self.total_params = self.actor.total_params + self.critic.total_params
