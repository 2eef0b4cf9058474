"""ORACLE (test infrastructure / CPU baseline only): torch-CPU restatement of the NeuroFlowCT train step.

Restates velora/models/nf/agent.py:182-255 (`_update_critics`, `_train_step`), velora/models/sac/continuous.py:106-143
(actor head), :219-232 (critic head), velora/models/nf/modules.py:109-118, 394-415, 691-717 (gradient steps, target
update, entropy loss) over the op-for-op network ports of oracle/liquid_oracle.py.  Used by bench_configs.py as the
CPU baseline of configs 1/2/5 (kind "port": the Python reference cannot travel to the GPU box).  Pinned by
tests/test_oracle.py::test_train_oracle_vs_reference against tests/golden/train.npz.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F
from torch.distributions import Normal

from oracle import liquid_oracle as lo


class TrainOracle:
    def __init__(self, actor_sd: Dict, critic1_sd: Dict, critic2_sd: Dict, *, action_scale=3.0, action_bias=0.0, lr=3e-4, tau=0.005,
                 gamma=0.99, log_std=(-5.0, 2.0), action_dim=1):
        strip = lambda sd, pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}   # noqa: E731
        self.actor = lo.torch_params(strip(actor_sd, "ncp."), requires_grad=True)
        self.c1 = lo.torch_params(strip(critic1_sd, "ncp."), requires_grad=True)
        self.c2 = lo.torch_params(strip(critic2_sd, "ncp."), requires_grad=True)
        self.t1 = {k: v.detach().clone() for k, v in self.c1.items()}
        self.t2 = {k: v.detach().clone() for k, v in self.c2.items()}
        self.log_alpha = torch.zeros((), requires_grad=True)
        self.scale, self.bias, self.tau, self.gamma, self.log_std = action_scale, action_bias, tau, gamma, log_std
        self.target_entropy = -action_dim
        leaves = lambda p: [t for t in p.values() if t.requires_grad]   # noqa: E731
        self.opt_a = torch.optim.Adam(leaves(self.actor), lr=lr)
        self.opt_c1 = torch.optim.Adam(leaves(self.c1), lr=lr)
        self.opt_c2 = torch.optim.Adam(leaves(self.c2), lr=lr)
        self.opt_e = torch.optim.Adam([self.log_alpha], lr=lr)

    def actor_forward(self, obs, hidden, eps=None):
        x, h = lo.liquid_forward_torch(self.actor, obs, hidden)
        mean, log_std = torch.chunk(x, 2, dim=-1)
        std = torch.clamp(log_std, *self.log_std).exp()
        x_t = mean + eps * std if eps is not None else Normal(mean, std).rsample()
        lp = Normal(mean, std).log_prob(x_t)
        an = torch.tanh(x_t)
        actions = an * self.scale + self.bias
        lp = lp - torch.log(self.scale * (1 - an.pow(2)) + 1e-6)
        return actions, lp.sum(dim=-1, keepdim=True), h

    @staticmethod
    def q(p, obs, act):
        return lo.ncp_forward_torch(p, torch.cat([obs, act], dim=-1))

    def train_step(self, batch: Dict[str, torch.Tensor], eps: List[torch.Tensor] | None = None):
        e0, e1 = (eps if eps is not None else (None, None))
        alpha = self.log_alpha.exp()
        with torch.no_grad():
            na, nlp, _ = self.actor_forward(batch["next_states"], batch["hiddens"], e0)
            nq = torch.min(self.q(self.t1, batch["next_states"], na), self.q(self.t2, batch["next_states"], na)) - alpha * nlp
            tq = batch["rewards"] + (1 - batch["dones"]) * self.gamma * nq
        c1 = F.mse_loss(self.q(self.c1, batch["states"], batch["actions"]), tq)
        c2 = F.mse_loss(self.q(self.c2, batch["states"], batch["actions"]), tq)
        for opt, loss in ((self.opt_c1, c1), (self.opt_c2, c2)):
            opt.zero_grad()
            loss.backward()
            opt.step()
        a, lp, _ = self.actor_forward(batch["states"], batch["hiddens"], e1)
        nq = torch.min(self.q(self.c1, batch["states"], a), self.q(self.c2, batch["states"], a))
        actor_loss = (self.log_alpha.exp() * lp - nq).mean()
        ent_loss = -(self.log_alpha * (lp + self.target_entropy).detach()).mean()
        self.opt_a.zero_grad()
        actor_loss.backward()
        self.opt_a.step()
        self.opt_e.zero_grad()
        ent_loss.backward()
        self.opt_e.step()
        with torch.no_grad():
            for src, dst in ((self.c1, self.t1), (self.c2, self.t2)):
                for k, v in src.items():
                    if v.requires_grad:
                        dst[k].copy_(self.tau * v + (1.0 - self.tau) * dst[k])
        return float((c1 + c2).detach()), float(actor_loss.detach()), float(ent_loss.detach())
