"""Checkpoint compatibility: a checkpoint written by this package's `save` restores into the REFERENCE's modules through
the reference's own loading code (velora/utils/restore.py:198-272), and a checkpoint written by the reference's `save_model` restores
into ours - networks (weights, biases, mask buffers), targets and Adam state, key for key and bit for bit.  CPU only; skipped where no
reference tree is present."""
import json

import pytest
import torch

from oracle import ref_import


def _poke(modules):
    """One Adam step on hand-made gradients so that the optimizers hold non-trivial state (no forward needed: CPU)."""
    g = torch.Generator().manual_seed(9)
    for opt in modules:
        for group in opt.param_groups:
            for p in group["params"]:
                p.grad = torch.randn(p.shape, generator=g) * (p != 0).float() if p.dim() == 2 else torch.randn(p.shape, generator=g)
        opt.step()


def _optims(agent):
    return [agent.actor.optim, agent.critic.optim1, agent.critic.optim2, agent.entropy.optim]


def _assert_same(ours, theirs):
    pairs = [(ours.actor.network, theirs.actor.network), (ours.critic.network1, theirs.critic.network1), (ours.critic.network2, theirs.critic.network2),
             (ours.critic.target1, theirs.critic.target1), (ours.critic.target2, theirs.critic.target2)]
    for a, b in pairs:
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    for oa, ob in zip(_optims(ours), _optims(theirs)):
        sa, sb = oa.state_dict(), ob.state_dict()
        assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"] and sorted(sa["state"].keys()) == sorted(sb["state"].keys())
        for pid in sa["state"]:
            for k in sa["state"][pid]:
                assert torch.equal(torch.as_tensor(sa["state"][pid][k]).float().cpu(), torch.as_tensor(sb["state"][pid][k]).float().cpu()), (pid, k)


@pytest.mark.parametrize("kind", ["continuous", "discrete"])
def test_checkpoints_round_trip_with_the_reference(tmp_path, kind):
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    ref = ref_import.import_reference()
    import velora.utils.restore as rr
    from typing import get_args

    from velora_b200.models.nf import NeuroFlowCore, NeuroFlowCTCore

    one, zero = torch.tensor([3.0]), torch.tensor([0.0])

    def make_ours(seed):
        if kind == "continuous":
            return NeuroFlowCTCore(4, 1, 20, 128, one, zero, buffer_size=64, seed=seed)
        return NeuroFlowCore(4, 2, 20, 128, buffer_size=64, seed=seed)

    def make_ref(seed):
        ref.set_seed(seed)
        cls = ref.NeuroFlowCT if kind == "continuous" else ref.NeuroFlow
        agent = object.__new__(cls)          # everything below the environment is the reference's own code
        agent.device = None
        if kind == "continuous":
            agent.actor = ref.ActorModule(4, 20, 1, one, zero)
            agent.critic = ref.CriticModule(4, 128, 1)
            agent.entropy = ref.EntropyModule(1)
            agent.buffer = ref.ReplayBuffer(64, 4, 1, agent.actor.hidden_size)
        else:
            agent.actor = ref.ActorModuleDiscrete(4, 20, 2)
            agent.critic = ref.CriticModuleDiscrete(4, 128, 2)
            agent.entropy = ref.EntropyModuleDiscrete(2)
            agent.buffer = ref.ReplayBuffer(64, 4, 1, agent.actor.hidden_size)
        agent.metadata = {"env_id": "shim", "device": "cpu", "optim": "torch.optim.Adam", "seed": seed}
        return agent

    # ---- ours -> reference: our save(), the reference's loading steps
    ours = make_ours(3)
    _poke(_optims(ours))
    ours.save(tmp_path / "ours")
    theirs = make_ref(4)
    meta = json.load(open(tmp_path / "ours" / "metadata.json"))
    state = rr.model_from_tensor(rr.load_file(tmp_path / "ours" / "model_state.safetensors", "cpu"))
    optim_state = rr.optim_from_tensor(rr.load_file(tmp_path / "ours" / "optim_state.safetensors", "cpu"))
    meta.pop("model")
    for key in meta:                                     # restore.py:253-260
        state[key] = {"state": optim_state.get(key, {}), "param_groups": meta[key]}
    for name in get_args(rr.ModuleNames):                # restore.py:262-265
        getattr(theirs, name).load_state_dict(state)
    _assert_same(ours, theirs)

    # ---- reference -> ours: the reference's save_model(), our load_state()
    theirs = make_ref(5)
    _poke(_optims(theirs))
    rr.save_model(theirs, tmp_path / "ref", buffer=True)
    mine = make_ours(6)
    mine.load_state(tmp_path / "ref", buffer=True)
    _assert_same(mine, theirs)
    for k, v in theirs.buffer.state_dict().items():
        assert torch.equal(v, mine.buffer.state_dict()[k]), k
    with pytest.raises(FileExistsError):
        mine.save(tmp_path / "ref")
    mine.save(tmp_path / "ref", force=True, buffer=True)


def test_liquid_critics_mirror_the_reference_classes():
    """SACCritic / SACCriticDiscrete (the liquid critics the agents do not use): same parameters as the reference's for the same seed."""
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    ref = ref_import.import_reference()
    import velora.models.sac.continuous as rc
    import velora.models.sac.discrete as rd

    import velora_b200 as vb
    from velora_b200.models.sac import SACCritic, SACCriticDiscrete

    for ours_cls, ref_cls, args in ((SACCritic, rc.SACCritic, (4, 20, 1)), (SACCriticDiscrete, rd.SACCriticDiscrete, (4, 20, 2))):
        vb.set_seed(11)
        a = ours_cls(*args)
        ref.set_seed(11)
        b = ref_cls(*args)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
