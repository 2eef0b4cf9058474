"""Generate tests/golden/*.npz by running the UNMODIFIED reference (test infrastructure, build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

What is pinned (the reference's own tests hold no numeric vectors for this path, so the
goldens are outputs of the reference itself, run here with torch 2.11.0 CPU / numpy 2.3.5):

  wiring.npz   Wiring masks/counts after set_seed(s)                     velora/wiring.py:75-338
  liquid.npz   LiquidNCPNetwork parameters (init parity), y, h', every gradient, for T=1 and for the T-fold
               loop of forward() that defines forward_seq                velora/models/lnn/ncp.py:129-178
  cell.npz     NCPLiquidCell on a hand-written mask                      velora/models/lnn/cell.py:147-169
  ncp.npz      NCPNetwork parameters, y, gradients incl. dx              velora/models/lnn/ncp.py:299-328
  buffer.npz   ReplayBuffer ring writes and seeded sample()              velora/buffer/base.py:73-143, replay.py:61-88
  train.npz    three NeuroFlowCT._train_step calls (noise/indices given)  velora/models/nf/agent.py:182-255

fp64 arbiters are produced by deep-copying the reference module to float64 and composing its own sub-modules
(the reference forward() casts x to float32 at ncp.py:158, so the composition is spelled out here).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def t2n(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().copy()


# ----------------------------------------------------------------------------------------------- wiring
WIRING_CASES = [
    # (in, neurons, out, sparsity, seed)
    (10, 20, 5, 0.5, 64),
    (4, 20, 2, 0.5, 64),
    (5, 128, 1, 0.5, 64),
    (4, 128, 2, 0.5, 64),
    (8, 512, 1, 0.5, 64),
    (4, 10, 2, 0.5, 0),
    (3, 5, 1, 0.5, 7),
    (6, 40, 3, 0.1, 11),
    (6, 40, 3, 0.9, 11),
    (2, 2, 1, 0.9, 3),
    (17, 33, 9, 0.3, 123),
]


def gen_wiring(ref) -> dict:
    out = {}
    cases = []
    for idx, (i, n, o, sp, seed) in enumerate(WIRING_CASES):
        ref.set_seed(seed)
        w = ref.Wiring(i, n, o, sparsity_level=sp)
        masks, counts = w.data()
        out[f"c{idx}_inter"] = t2n(masks.inter)
        out[f"c{idx}_command"] = t2n(masks.command)
        out[f"c{idx}_motor"] = t2n(masks.motor)
        out[f"c{idx}_recurrent"] = t2n(masks.recurrent)
        out[f"c{idx}_counts"] = np.array([counts.sensory, counts.inter, counts.command, counts.motor], dtype=np.int64)
        nc = w.n_connections
        out[f"c{idx}_synapses"] = np.array([nc.sensory, nc.inter, nc.command, nc.motor], dtype=np.int64)
        # RNG state after construction pins how much of each global stream was consumed
        out[f"c{idx}_next_np"] = np.array([np.random.randint(0, 2**31 - 1)], dtype=np.int64)
        out[f"c{idx}_next_torch"] = t2n(torch.randint(0, 2**31 - 1, (1,)))
        cases.append([i, n, o, sp, seed])
    out["cases"] = np.array(cases, dtype=np.float64)
    return out


# ----------------------------------------------------------------------------------------------- liquid
def liquid_compose(net, x, h):
    """The composition of ncp.py:168-171 without the float32 cast, for any dtype."""
    a = net.act(net.inter(x))
    c, h_new = net.command(a, h)
    y = net.motor(net.act(c))
    return y, h_new


def liquid_seq_ref(net, x_seq, h0, compose=False):
    """forward_seq := the T-fold loop of forward() carrying h."""
    ys, hs = [], []
    h = h0
    for t in range(x_seq.shape[1]):
        if compose:
            y, h = liquid_compose(net, x_seq[:, t], h)
        else:
            y, h = net(x_seq[:, t], h)
            if y.dim() == 1:
                y = y.unsqueeze(0)
        ys.append(y)
        hs.append(h)
    return torch.stack(ys, 1), torch.stack(hs, 1)


LIQUID_CASES = [
    # (name, in, neurons, out, sparsity, seed, B, T, weight_on_all_h)
    ("fixture", 10, 20, 5, 0.5, 64, 7, 1, False),
    ("actor_t1", 4, 20, 2, 0.5, 64, 64, 1, False),
    ("actor_seq", 4, 20, 2, 0.5, 64, 33, 8, False),
    ("actor_seq_allh", 4, 20, 2, 0.5, 64, 5, 6, True),
    ("fixture_seq", 10, 20, 5, 0.5, 64, 9, 5, True),
    ("odd", 3, 7, 1, 0.3, 5, 6, 4, True),
    ("mid", 8, 64, 3, 0.5, 9, 5, 3, False),
    ("wide", 6, 150, 4, 0.7, 21, 3, 2, True),
]


def gen_liquid(ref) -> dict:
    out = {}
    names = []
    for (name, i, n, o, sp, seed, B, T, allh) in LIQUID_CASES:
        ref.set_seed(seed)
        net = ref.LiquidNCPNetwork(i, n, o, sparsity_level=sp)
        names.append(name)
        out[f"{name}_meta"] = np.array([i, n, o, sp, seed, B, T, int(allh)], dtype=np.float64)
        for k, v in net.state_dict().items():
            out[f"{name}_sd_{k}"] = t2n(v)
        nc = net.hidden_size
        g = torch.Generator().manual_seed(123)
        x = torch.randn(B, T, i, generator=g)
        h0 = 0.5 * torch.randn(B, nc, generator=g)
        ry = torch.randn(B, T, o, generator=g)
        rh = torch.randn(B, T, nc, generator=g)
        if not allh:
            rh[:, :-1] = 0.0
        out[f"{name}_x"] = t2n(x)
        out[f"{name}_h0"] = t2n(h0)
        out[f"{name}_ry"] = t2n(ry)
        out[f"{name}_rh"] = t2n(rh)
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            m = copy.deepcopy(net).to(dtype)
            xx = x.detach().clone().to(dtype).requires_grad_(True)
            hh = h0.detach().clone().to(dtype).requires_grad_(True)
            y, htraj = liquid_seq_ref(m, xx, hh, compose=(dtype == torch.float64))
            loss = (y * ry.to(dtype)).sum() + (htraj * rh.to(dtype)).sum()
            loss.backward()
            out[f"{name}_{tag}_y"] = t2n(y)
            out[f"{name}_{tag}_htraj"] = t2n(htraj)
            out[f"{name}_{tag}_dx"] = t2n(xx.grad)
            out[f"{name}_{tag}_dh0"] = t2n(hh.grad)
            for k, p in m.named_parameters():
                out[f"{name}_{tag}_grad_{k}"] = t2n(p.grad)
        # h_state=None path (zeros), forward only, T=1 semantics of ncp.py:162-166
        y0, h1 = net(x[:, 0], None)
        out[f"{name}_f32_y_h0none"] = t2n(y0 if y0.dim() == 2 else y0.unsqueeze(0))
        out[f"{name}_f32_h_h0none"] = t2n(h1)
    out["names"] = np.array(names)
    # The anchor recorded in SURVEY.md 8c for the (10,20,5) fixture
    ref.set_seed(64)
    net = ref.LiquidNCPNetwork(10, 20, 5)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(7, 10, generator=g)
    h = 0.5 * torch.randn(7, 8, generator=g)
    y, h2 = net(x, h)
    (y.sum() + h2.sum()).backward()
    out["anchor_sum_y"] = np.array([y.sum().item()])
    out["anchor_sum_h"] = np.array([h2.sum().item()])
    out["anchor_y0"] = t2n(y[0])
    out["anchor_sum_ghead_grad"] = np.array([net.command.g_head.weight.grad.sum().item()])
    out["anchor_total_active"] = np.array([net.total_params, net.active_params], dtype=np.int64)
    return out


def gen_liquid_big(ref) -> dict:
    """BASELINE config 4's network, LiquidNCPNetwork(8,512,1) (Ni=308, Nc=204), at B=8, T=4: parameters, fp32 outputs / data
    gradients, and the fp64 deep copy's outputs and all 14 parameter gradients (the arbiter; the reference's own fp32 distance
    from it is stored per tensor as `e32_*` instead of a second 2 MB gradient set)."""
    out = {}
    name, i, n, o, sp, seed, B, T = "big", 8, 512, 1, 0.5, 64, 8, 4
    ref.set_seed(seed)
    net = ref.LiquidNCPNetwork(i, n, o, sparsity_level=sp)
    out[f"{name}_meta"] = np.array([i, n, o, sp, seed, B, T, 1], dtype=np.float64)
    for k, v in net.state_dict().items():
        out[f"{name}_sd_{k}"] = t2n(v)
    nc = net.hidden_size
    g = torch.Generator().manual_seed(123)
    x = torch.randn(B, T, i, generator=g)
    h0 = 0.5 * torch.randn(B, nc, generator=g)
    ry = torch.randn(B, T, o, generator=g)
    rh = torch.randn(B, T, nc, generator=g) / 8
    for k, v in (("x", x), ("h0", h0), ("ry", ry), ("rh", rh)):
        out[f"{name}_{k}"] = t2n(v)
    res = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = copy.deepcopy(net).to(dtype)
        xx = x.detach().clone().to(dtype).requires_grad_(True)
        hh = h0.detach().clone().to(dtype).requires_grad_(True)
        y, htraj = liquid_seq_ref(m, xx, hh, compose=(dtype == torch.float64))
        ((y * ry.to(dtype)).sum() + (htraj * rh.to(dtype)).sum()).backward()
        res[tag] = {"y": y, "htraj": htraj, "dx": xx.grad, "dh0": hh.grad, **{f"grad_{k}": p.grad for k, p in m.named_parameters()}}
    for k, v in res["f64"].items():
        out[f"{name}_f64_{k}"] = t2n(v)
        a, b = res["f32"][k].double(), v.double()
        out[f"{name}_e32_{k}"] = np.array([float((a - b).norm() / b.norm())])
    for k in ("y", "htraj", "dx", "dh0"):
        out[f"{name}_f32_{k}"] = t2n(res["f32"][k])
    return out


# ----------------------------------------------------------------------------------------------- cell
def gen_cell(ref) -> dict:
    """NCPLiquidCell on the style of hand-written mask the reference tests use (an in x hidden {-1,0,1} matrix)."""
    out = {}
    torch.manual_seed(3)
    mask = torch.tensor(
        [
            [1, 0, -1, 0, 1, 0],
            [0, 1, 0, -1, 0, 1],
            [-1, 0, 1, 0, 0, -1],
            [0, -1, 0, 1, 1, 0],
            [1, 1, 0, 0, -1, 0],
            [0, 0, -1, 1, 0, 1],
            [1, 0, 0, -1, 0, 0],
            [0, -1, 1, 0, 0, 1],
        ],
        dtype=torch.int32,
    )  # in_features=8, n_hidden=6
    cell = ref.NCPLiquidCell(8, 6, mask)
    out["mask"] = t2n(mask)
    for k, v in cell.state_dict().items():
        out[f"sd_{k}"] = t2n(v)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 8, generator=g, requires_grad=True)
    h = torch.randn(4, 6, generator=g, requires_grad=True)
    y, hn = cell(x, h)
    (y.sum() + 2.0 * hn.sum()).backward()
    out["x"], out["h"], out["y"], out["hn"] = t2n(x), t2n(h), t2n(y), t2n(hn)
    out["dx"], out["dh"] = t2n(x.grad), t2n(h.grad)
    for k, p in cell.named_parameters():
        out[f"grad_{k}"] = t2n(p.grad)
    return out


# ----------------------------------------------------------------------------------------------- ncp (critic)
NCP_CASES = [
    ("critic_ct", 5, 128, 1, 0.5, 64, 64),
    ("critic_disc", 4, 128, 2, 0.5, 64, 17),
    ("small", 3, 10, 2, 0.5, 2, 1),
    ("odd", 7, 23, 3, 0.2, 4, 9),
]


def gen_ncp(ref) -> dict:
    out = {}
    names = []
    for (name, i, n, o, sp, seed, B) in NCP_CASES:
        ref.set_seed(seed)
        net = ref.NCPNetwork(i, n, o, sparsity_level=sp)
        names.append(name)
        out[f"{name}_meta"] = np.array([i, n, o, sp, seed, B], dtype=np.float64)
        for k, v in net.state_dict().items():
            out[f"{name}_sd_{k}"] = t2n(v)
        g = torch.Generator().manual_seed(77)
        x = torch.randn(B, i, generator=g)
        r = torch.randn(B, o, generator=g)
        out[f"{name}_x"], out[f"{name}_r"] = t2n(x), t2n(r)
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            m = copy.deepcopy(net).to(dtype)
            xx = x.detach().clone().to(dtype).requires_grad_(True)
            y = m.ncp(xx)  # the Sequential itself: no float32 cast, no squeeze (ncp.py:319-326)
            (y * r.to(dtype)).sum().backward()
            out[f"{name}_{tag}_y"] = t2n(y)
            out[f"{name}_{tag}_dx"] = t2n(xx.grad)
            for k, p in m.named_parameters():
                out[f"{name}_{tag}_grad_{k}"] = t2n(p.grad)
        y_api = net(x)
        out[f"{name}_api_shape"] = np.array(list(y_api.shape), dtype=np.int64)
    out["names"] = np.array(names)
    return out


# ----------------------------------------------------------------------------------------------- buffer
def gen_buffer(ref) -> dict:
    out = {}
    cap, S, A, H = 50, 4, 1, 8
    buf = ref.ReplayBuffer(cap, S, A, H)
    g = torch.Generator().manual_seed(11)

    def chunk(n):
        return (
            torch.randn(n, S, generator=g),
            torch.rand(n, A, generator=g) * 6 - 3,
            torch.randn(n, generator=g),
            torch.randn(n, S, generator=g),
            (torch.rand(n, generator=g) < 0.1),
            0.3 * torch.randn(n, H, generator=g),
        )

    feeds = []
    # 30 rows by add_multi, then 7 single add()s, then 25 rows by add_multi (wraps), ring position checked
    for n in (30,):
        c = chunk(n)
        buf.add_multi(*c)
        feeds.append(("multi", c))
    c = chunk(7)
    for r in range(7):
        buf.add(c[0][r], c[1][r], float(c[2][r]), c[3][r], bool(c[4][r]), c[5][r])
    feeds.append(("single", c))
    c = chunk(25)
    buf.add_multi(*c)
    feeds.append(("multi", c))
    for fi, (kind, c) in enumerate(feeds):
        out[f"feed{fi}_kind"] = np.array([kind])
        for ti, t in enumerate(c):
            out[f"feed{fi}_{ti}"] = t2n(t.to(torch.float32) if t.dtype != torch.bool else t)
    out["position_size"] = np.array([buf.position, buf.size], dtype=np.int64)
    for k, v in buf.state_dict().items():
        out[f"state_{k}"] = t2n(v)
    # seeded sample: indices are torch.randint(0, size, (B,)) on the CPU generator (replay.py:79)
    for B, seed in ((16, 1), (50, 2)):
        torch.manual_seed(seed)
        batch = buf.sample(B)
        torch.manual_seed(seed)
        idx = torch.randint(0, buf.size, (B,))
        out[f"sample{B}_idx"] = t2n(idx)
        for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens"):
            out[f"sample{B}_{k}"] = t2n(getattr(batch, k))
    out["dims"] = np.array([cap, S, A, H], dtype=np.int64)
    return out


# ----------------------------------------------------------------------------------------------- train step (cfg 1/2)
def _forget_scripted_classes() -> None:
    """torch.jit caches one concrete type per module class, ignored methods included: a class that was scripted earlier in the same
    process (another test) would keep its old `get_sample` after the patch below.  Start from an empty store."""
    import torch.jit._recursive as _rec

    _rec.concrete_type_store = _rec.ConcreteTypeStore()


def gen_train(ref) -> dict:
    """Three NeuroFlowCT._train_step calls (velora/models/nf/agent.py:182-255) on a synthetic buffer, InvertedPendulum
    shapes (S=4, A=1, scale 3, bias 0; actor 20 / critic 128 neurons), batch 32, with the policy noise and the sampled
    indices supplied from outside so that any device can replay the same step."""
    import torch.nn as nn
    from torch.distributions import Normal
    import velora.models.sac.continuous as cont
    from velora.buffer.experience import BatchExperience

    out = {}
    EPS, IDX = [], []

    def get_sample(self, mean, std):
        eps = EPS.pop(0).to(mean)
        x_t = mean + eps * std                     # == Normal(mean, std).rsample() with this eps
        return x_t, Normal(mean, std).log_prob(x_t)

    orig = cont.SACActor.get_sample
    _forget_scripted_classes()
    cont.SACActor.get_sample = torch.jit.ignore(get_sample)
    try:
        S, A, B, steps = 4, 1, 32, 3
        ref.set_seed(64)
        agent = object.__new__(ref.NeuroFlowCT)
        agent.device = None
        agent.gamma, agent.tau = 0.99, 0.005
        agent.actor = ref.ActorModule(S, 20, A, torch.tensor([3.0]), torch.tensor([0.0]))
        agent.critic = ref.CriticModule(S, 128, A)
        agent.entropy = ref.EntropyModule(A)
        agent.loss = nn.MSELoss()
        agent.buffer = ref.ReplayBuffer(512, S, A, agent.actor.hidden_size)
        g = torch.Generator().manual_seed(2)
        n = 300
        feed = (torch.randn(n, S, generator=g), torch.rand(n, A, generator=g) * 6 - 3, torch.ones(n),
                torch.randn(n, S, generator=g), (torch.rand(n, generator=g) < 0.02), 0.3 * torch.randn(n, 8, generator=g))
        agent.buffer.add_multi(*feed)
        for k, v in agent.buffer.state_dict().items():
            out[f"buf_{k}"] = t2n(v)
        for k, v in agent.actor.network.state_dict().items():
            out[f"init_actor_{k}"] = t2n(v)
        for k, v in agent.critic.network1.state_dict().items():
            out[f"init_critic1_{k}"] = t2n(v)
        for k, v in agent.critic.network2.state_dict().items():
            out[f"init_critic2_{k}"] = t2n(v)

        def sample(batch_size):
            idx = IDX.pop(0)
            b = agent.buffer
            return BatchExperience(b.states[idx], b.actions[idx], b.rewards[idx], b.next_states[idx], b.dones[idx], b.hiddens[idx])

        agent.buffer.sample = sample
        eps_all = torch.randn(steps * 2, B, A, generator=g)
        idx_all = torch.randint(0, n, (steps, B), generator=g)
        out["eps"], out["idx"] = t2n(eps_all), t2n(idx_all)
        losses = []
        for s in range(steps):
            EPS.extend([eps_all[2 * s], eps_all[2 * s + 1]])
            IDX.append(idx_all[s])
            r = agent._train_step(B)
            assert not EPS and not IDX, "the recorded noise / indices were not consumed: the patched hooks are not in use"
            losses.append([float(r["critic"]), float(r["actor"]), float(r["entropy"])])
        out["losses"] = np.array(losses, dtype=np.float64)
        for k, v in agent.actor.network.state_dict().items():
            out[f"final_actor_{k}"] = t2n(v)
        for k, v in agent.critic.network1.state_dict().items():
            out[f"final_critic1_{k}"] = t2n(v)
        for k, v in agent.critic.target1.state_dict().items():
            out[f"final_target1_{k}"] = t2n(v)
        out["final_log_alpha"] = t2n(agent.entropy.log_alpha)
        out["dims"] = np.array([S, A, B, steps, n], dtype=np.int64)
    finally:
        cont.SACActor.get_sample = orig
        _forget_scripted_classes()
    return out


def categorical_from_uniform(probs: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Inverse-CDF draw from `probs` [B, A] with recorded uniforms u [B]: the stand-in for Categorical.sample() that both the
    golden generator and the GPU test use, so that the same noise replays on any device."""
    cdf = torch.cumsum(probs, dim=-1)
    return torch.clamp((cdf < u.to(probs).unsqueeze(-1)).sum(dim=-1), max=probs.shape[-1] - 1)


def gen_train_discrete(ref) -> dict:
    """Three NeuroFlow._train_step calls (discrete agent, velora/models/nf/agent.py:549-636) on a synthetic buffer with
    CartPole shapes (S=4, 2 actions; actor 20 / critic 128 neurons = BASELINE config 1), batch 32; the categorical draws and
    the sampled indices are supplied from outside."""
    import torch.nn as nn
    import velora.models.sac.discrete as disc
    from velora.buffer.experience import BatchExperience

    out = {}
    UNI, IDX = [], []

    def get_sample(self, probs):
        actions = categorical_from_uniform(probs, UNI.pop(0))
        return actions, torch.log(torch.gather(probs, -1, actions.unsqueeze(-1)))   # == Categorical(probs).log_prob(actions)[:, None]

    orig = disc.SACActorDiscrete.get_sample
    _forget_scripted_classes()
    disc.SACActorDiscrete.get_sample = torch.jit.ignore(get_sample)
    try:
        S, A, B, steps = 4, 2, 32, 3
        ref.set_seed(64)
        agent = object.__new__(ref.NeuroFlow)
        agent.device = None
        agent.gamma, agent.tau = 0.99, 0.005
        agent.actor = ref.ActorModuleDiscrete(S, 20, A)
        agent.critic = ref.CriticModuleDiscrete(S, 128, A)
        agent.entropy = ref.EntropyModuleDiscrete(A)
        agent.loss = nn.MSELoss()
        agent.buffer = ref.ReplayBuffer(512, S, 1, agent.actor.hidden_size)
        g = torch.Generator().manual_seed(3)
        n = 300
        feed = (torch.randn(n, S, generator=g), torch.randint(0, A, (n, 1), generator=g).float(), torch.ones(n),
                torch.randn(n, S, generator=g), (torch.rand(n, generator=g) < 0.02), 0.3 * torch.randn(n, 8, generator=g))
        agent.buffer.add_multi(*feed)
        for k, v in agent.buffer.state_dict().items():
            out[f"buf_{k}"] = t2n(v)
        for k, v in agent.actor.network.state_dict().items():
            out[f"init_actor_{k}"] = t2n(v)
        for k, v in agent.critic.network1.state_dict().items():
            out[f"init_critic1_{k}"] = t2n(v)
        for k, v in agent.critic.network2.state_dict().items():
            out[f"init_critic2_{k}"] = t2n(v)

        def sample(batch_size):
            idx = IDX.pop(0)
            b = agent.buffer
            return BatchExperience(b.states[idx], b.actions[idx], b.rewards[idx], b.next_states[idx], b.dones[idx], b.hiddens[idx])

        agent.buffer.sample = sample
        uni_all = torch.rand(steps * 2, B, generator=g)
        idx_all = torch.randint(0, n, (steps, B), generator=g)
        out["uni"], out["idx"] = t2n(uni_all), t2n(idx_all)
        losses = []
        for s in range(steps):
            UNI.extend([uni_all[2 * s], uni_all[2 * s + 1]])
            IDX.append(idx_all[s])
            r = agent._train_step(B)
            assert not UNI and not IDX, "the recorded noise / indices were not consumed: the patched hooks are not in use"
            losses.append([float(r["critic"]), float(r["actor"]), float(r["entropy"])])
        out["losses"] = np.array(losses, dtype=np.float64)
        for k, v in agent.actor.network.state_dict().items():
            out[f"final_actor_{k}"] = t2n(v)
        for k, v in agent.critic.network1.state_dict().items():
            out[f"final_critic1_{k}"] = t2n(v)
        for k, v in agent.critic.target1.state_dict().items():
            out[f"final_target1_{k}"] = t2n(v)
        out["final_log_alpha"] = t2n(agent.entropy.log_alpha)
        out["dims"] = np.array([S, A, B, steps, n], dtype=np.int64)
    finally:
        disc.SACActorDiscrete.get_sample = orig
        _forget_scripted_classes()
    return out


def main() -> None:
    # The reference writes polarities with `mask[rows, cols] = pol` where (row, col) pairs repeat
    # (np.random.choice draws with replacement, wiring.py:245-251).  With several intra-op threads torch's
    # index_put_ picks an arbitrary winner among duplicates, so the SIGN of a few entries of large masks changes
    # from run to run (abs(mask), which is all the forward uses, does not).  One thread = sequential,
    # last-write-wins, which is what the goldens pin.
    torch.set_num_threads(1)
    ref = ref_import.import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    for name, fn in (("wiring", gen_wiring), ("liquid", gen_liquid), ("liquid_big", gen_liquid_big), ("cell", gen_cell), ("ncp", gen_ncp),
                     ("buffer", gen_buffer), ("train", gen_train), ("train_discrete", gen_train_discrete)):
        if only is not None and name not in only:
            continue
        data = fn(ref)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {len(data)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w" if only is None else "a") as f:
        f.write(
            "Generated by oracle/make_golden.py from the unmodified reference at /root/reference (velora v0.2.0)\n"
            f"torch {torch.__version__}, numpy {np.__version__}, CPU.\n"
        )


if __name__ == "__main__":
    main()
