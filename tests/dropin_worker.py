"""Drop-in proof (run as a subprocess by tests/test_dropin_gpu.py): the REFERENCE's own agent code - ActorModule, CriticModule,
EntropyModule (velora/models/nf/modules.py: deepcopy + torch.jit.script of the networks) and NeuroFlowCT._train_step /
_update_critics (velora/models/nf/agent.py:182-255) - runs unmodified from baseline/_ref (or /root/reference), with only the three
hot-path classes swapped for the B200 ones:

    velora.models.base.LiquidNCPNetwork / NCPNetwork  ->  velora_b200.models.lnn.LiquidNCPNetwork / NCPNetwork
    velora.models.nf.agent.ReplayBuffer               ->  velora_b200.buffer.ReplayBuffer

and is compared with tests/golden/train.npz (three steps of the all-reference agent on the CPU, recorded noise and indices).
A subprocess, because the recorded noise is injected by replacing SACActor.get_sample and torch.jit caches ignored methods per class.
Prints one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    construct_only = "--construct-only" in sys.argv
    from oracle import ref_import

    ref = ref_import.import_reference()
    import velora.models.base as rbase
    import velora.models.nf.agent as ragent
    import velora.models.sac.continuous as rcont
    from torch.distributions import Normal

    import velora_b200 as vb
    import velora_b200.buffer
    import velora_b200.models.lnn
    from velora_b200 import functional as VF

    # the swap a velora maintainer makes: three names
    rbase.LiquidNCPNetwork = vb.models.lnn.LiquidNCPNetwork
    rbase.NCPNetwork = vb.models.lnn.NCPNetwork
    ragent.ReplayBuffer = vb.buffer.ReplayBuffer

    EPS = []

    def get_sample(self, mean, std):
        eps = EPS.pop(0).to(mean)
        x_t = mean + eps * std
        return x_t, Normal(mean, std).log_prob(x_t)

    rcont.SACActor.get_sample = torch.jit.ignore(get_sample)

    d = np.load(os.path.join(ROOT, "tests", "golden", "train.npz"))
    pre = lambda p: {k[len(p):]: torch.tensor(d[k]) for k in d.files if k.startswith(p)}   # noqa: E731
    S, A, B, steps, n = (int(v) for v in d["dims"])
    dev = None if construct_only else torch.device("cuda")

    ref.set_seed(64)
    agent = object.__new__(ref.NeuroFlowCT)          # the environment is out of scope: everything below the env is the reference's
    agent.device = dev
    agent.gamma, agent.tau = 0.99, 0.005
    agent.actor = ref.ActorModule(S, 20, A, torch.tensor([3.0], device=dev), torch.tensor([0.0], device=dev), device=dev)
    agent.critic = ref.CriticModule(S, 128, A, device=dev)
    agent.entropy = ref.EntropyModule(A, device=dev)
    agent.loss = nn.MSELoss()
    agent.buffer = ragent.ReplayBuffer(512, S, A, agent.actor.hidden_size, device=dev)

    out = {
        # the scripted wrapper keeps the ignored Python methods of the class it was built from: `_fused_seq` / `_fused` exist only
        # on the B200 classes
        "actor_ncp": "velora_b200." if hasattr(agent.actor.network.ncp, "_fused_seq") and hasattr(agent.critic.network1.ncp, "_fused") else "reference",
        "scripted": isinstance(agent.actor.network, torch.jit.RecursiveScriptModule)
        and isinstance(agent.critic.network1, torch.jit.RecursiveScriptModule),
        "buffer": type(agent.buffer).__module__,
        "keys_match": list(agent.actor.network.state_dict().keys()) == list(pre("init_actor_").keys())
        and list(agent.critic.network1.state_dict().keys()) == list(pre("init_critic1_").keys()),
    }
    if construct_only:
        # CPU construction draws the RNG streams like the all-reference agent: bit-identical initial state
        out["init_equal"] = all(torch.equal(v, pre("init_actor_")[k]) for k, v in agent.actor.network.state_dict().items()) and all(
            torch.equal(v, pre("init_critic2_")[k]) for k, v in agent.critic.network2.state_dict().items())
        print(json.dumps(out))
        return

    for p, nets in (("init_actor_", [agent.actor.network]), ("init_critic1_", [agent.critic.network1, agent.critic.target1]),
                    ("init_critic2_", [agent.critic.network2, agent.critic.target2])):
        for net in nets:
            net.load_state_dict(pre(p))
    for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens"):
        getattr(agent.buffer, k).copy_(torch.tensor(d[f"buf_{k}"]))
    agent.buffer.size = n
    eps, idx = torch.tensor(d["eps"]), torch.tensor(d["idx"])
    IDX = []
    draw = torch.randint

    def recorded_randint(*a, **k):                   # the reference's sample() draws with torch.randint: feed the recorded indices
        return IDX.pop(0).to(k.get("device")) if IDX else draw(*a, **k)

    launches0 = VF.LAUNCHES
    losses = []
    for s in range(steps):
        EPS.extend([eps[2 * s], eps[2 * s + 1]])
        IDX.append(idx[s])
        torch.randint = recorded_randint
        try:
            r = agent._train_step(B)                  # reference code: sample -> _update_critics -> actor / entropy steps -> soft update
        finally:
            torch.randint = draw
        assert not EPS and not IDX, "recorded noise / indices not consumed"
        losses.append([float(r["critic"]), float(r["actor"]), float(r["entropy"])])
    torch.cuda.synchronize()
    out["launches"] = VF.LAUNCHES - launches0
    out["loss_ok"] = bool(np.allclose(np.array(losses), d["losses"], rtol=2e-5, atol=1e-6))
    out["losses"] = losses
    worst = 0.0
    for p, net in (("final_actor_", agent.actor.network), ("final_critic1_", agent.critic.network1), ("final_target1_", agent.critic.target1)):
        want = pre(p)
        for k, v in net.state_dict().items():
            worst = max(worst, float((v.detach().cpu().double() - want[k].double()).abs().max()))
    out["param_max_abs_diff"] = worst
    out["log_alpha_diff"] = abs(float(agent.entropy.log_alpha) - float(d["final_log_alpha"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
