"""Environment shims for the agent-level BASELINE configs (gymnasium / mujoco are not installed and cannot be fetched).

NOT part of the product and NOT a reimplementation of Gymnasium: two ~40-line physics restatements with the Gymnasium calling
convention, so that `NeuroFlow("CartPole-v1", 20, 128).train(64, n_episodes=50)` (BASELINE config 1) can be driven end to end
on both arms (ours and the unmodified reference agent code).  Dynamics follow the published CartPole-v1 equations (Barto, Sutton &
Anderson 1983; Euler integration, tau = 0.02, force 10 N, termination at |x| > 2.4 or |theta| > 12 degrees, 500-step limit).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np


class CartPoleShim:
    def __init__(self, seed: int | None = None):
        self.observation_space = SimpleNamespace(shape=(4,))
        self.action_space = SimpleNamespace(n=np.int64(2))
        self.spec = SimpleNamespace(id="CartPole-v1")
        self.rng = np.random.default_rng(seed)
        self.state = np.zeros(4)
        self.t = 0

    def reset(self, seed: int | None = None, options=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.state = self.rng.uniform(-0.05, 0.05, size=4)
        self.t = 0
        return self.state.astype(np.float32), {}

    def step(self, action):
        x, x_dot, th, th_dot = self.state
        force = 10.0 if int(action) == 1 else -10.0
        g, mc, mp, l = 9.8, 1.0, 0.1, 0.5
        total, pml = mc + mp, mp * l
        ct, st = math.cos(th), math.sin(th)
        temp = (force + pml * th_dot * th_dot * st) / total
        th_acc = (g * st - ct * temp) / (l * (4.0 / 3.0 - mp * ct * ct / total))
        x_acc = temp - pml * th_acc * ct / total
        tau = 0.02
        self.state = np.array([x + tau * x_dot, x_dot + tau * x_acc, th + tau * th_dot, th_dot + tau * th_acc])
        self.t += 1
        terminated = bool(abs(self.state[0]) > 2.4 or abs(self.state[2]) > 12 * 2 * math.pi / 360)
        truncated = self.t >= 500
        return self.state.astype(np.float32), 1.0, terminated, truncated, {}


class PendulumShim:
    """A 1-D continuous-action stand-in with InvertedPendulum-v5's interface shapes (4 observations, one action in [-3, 3]): the
    CartPole dynamics driven by a continuous force 10 * a / 3.  (MuJoCo itself is not available.)"""

    def __init__(self, seed: int | None = None):
        self._cp = CartPoleShim(seed)
        self.observation_space = SimpleNamespace(shape=(4,))
        self.action_space = SimpleNamespace(low=np.array([-3.0], dtype=np.float32), high=np.array([3.0], dtype=np.float32), shape=(1,))
        self.spec = SimpleNamespace(id="InvertedPendulum-shim")

    def reset(self, seed: int | None = None, options=None):
        return self._cp.reset(seed)

    def step(self, action):
        a = float(np.asarray(action).reshape(-1)[0])
        cp = self._cp
        x, x_dot, th, th_dot = cp.state
        force = 10.0 * max(-3.0, min(3.0, a)) / 3.0
        g, mc, mp, l = 9.8, 1.0, 0.1, 0.5
        total, pml = mc + mp, mp * l
        ct, st = math.cos(th), math.sin(th)
        temp = (force + pml * th_dot * th_dot * st) / total
        th_acc = (g * st - ct * temp) / (l * (4.0 / 3.0 - mp * ct * ct / total))
        x_acc = temp - pml * th_acc * ct / total
        cp.state = np.array([x + 0.02 * x_dot, x_dot + 0.02 * x_acc, th + 0.02 * th_dot, th_dot + 0.02 * th_acc])
        cp.t += 1
        terminated = bool(abs(cp.state[0]) > 2.4 or abs(cp.state[2]) > 0.2)
        return cp.state.astype(np.float32), 1.0, terminated, cp.t >= 1000, {}
