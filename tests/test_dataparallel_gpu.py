"""Data-parallel correctness ON GPUs (NCCL, two ranks): the sharded fused backward + the one all-reduce of the flat gradient
block gives every rank the single-process gradients of the whole batch; replicas stay bit-identical after a train step.
Skipped on a box with fewer than two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dataparallel_gpu.py -m gpu`)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import velora_b200 as vb
    from velora_b200 import dataparallel as dp
    from velora_b200.models.lnn import LiquidNCPNetwork, NCPNetwork
    from velora_b200.models.nf import NeuroFlowCTCore, NeuroFlowCore

    res = {}
    # ---- liquid network: global batch 512 x 6, contiguous shards
    B, T = 512, 6
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, 4, generator=g)
    h0 = 0.5 * torch.randn(B, 8, generator=g)
    ry = torch.randn(B, T, 2, generator=g)
    rh = torch.randn(B, T, 8, generator=g)

    def grads(net_cls, args, lo_, hi_, inp, up):
        vb.set_seed(64)
        net = net_cls(*args, device=dev)
        if net_cls is LiquidNCPNetwork:
            y, ht = net.forward_seq(inp[0][lo_:hi_].to(dev), inp[1][lo_:hi_].to(dev))
            ((y * up[0][lo_:hi_].to(dev)).sum() + (ht * up[1][lo_:hi_].to(dev)).sum()).backward()
        else:
            q = net(inp[0][lo_:hi_].to(dev))
            (q * up[0][lo_:hi_].to(dev)).sum().backward()
        return torch.cat([p.grad.flatten() for p in net.parameters()])

    dp.configure(True, reduce="sum", oneshot=False)                   # NCCL
    lo_, hi_ = dp.shard_bounds(B)
    g_nccl = grads(LiquidNCPNetwork, (4, 20, 2), lo_, hi_, (x, h0), (ry, rh))
    dp.configure(True, reduce="sum")                                  # this library's one-shot NVLink kernel when peer memory is available
    res["oneshot_available"] = dp._state["oneshot"] is not None
    calls0, one0 = dp._state["calls"], dp._state["oneshot_calls"]
    g_dp = grads(LiquidNCPNetwork, (4, 20, 2), lo_, hi_, (x, h0), (ry, rh))
    res["liquid_allreduces"] = dp._state["calls"] - calls0            # ONE collective per backward
    res["oneshot_used"] = dp._state["oneshot_calls"] - one0
    res["oneshot_vs_nccl"] = float((g_dp - g_nccl).abs().max())
    if res["oneshot_available"]:                                      # many calls back to back: the two-slot protocol holds
        t = torch.full((918,), float(rank + 1), device=dev)
        acc = torch.zeros(918, device=dev)
        for k in range(200):
            u = t * (k + 1)
            dp.maybe_allreduce(u)
            acc += u
        res["oneshot_stress"] = float((acc - 3.0 * sum(range(1, 201))).abs().max())
    xq, rq = torch.randn(B, 5, generator=g), torch.randn(B, 1, generator=g)
    q_dp = grads(NCPNetwork, (5, 128, 1), lo_, hi_, (xq,), (rq,))
    dp.configure(False)
    g_one = grads(LiquidNCPNetwork, (4, 20, 2), 0, B, (x, h0), (ry, rh))
    q_one = grads(NCPNetwork, (5, 128, 1), 0, B, (xq,), (rq,))
    res["liquid_rel"] = float((g_dp - g_one).norm() / g_one.norm())
    res["critic_rel"] = float((q_dp - q_one).norm() / q_one.norm())
    both = [torch.empty_like(g_dp) for _ in range(world)]
    dist.all_gather(both, g_dp)
    res["replicas_equal"] = bool(torch.equal(both[0], both[1]))

    # ---- whole agents: two train steps data-parallel == single process on the whole batch (same indices, same noise), with NCCL and
    #      with the one-shot NVLink kernel as the gradient exchange
    def agent_params(kind, parallel, oneshot, fused=False):
        dp.configure(parallel, reduce="mean", oneshot=oneshot)
        if kind == "ct":
            a = NeuroFlowCTCore(4, 1, 20, 128, torch.tensor([3.0]), torch.tensor([0.0]), buffer_size=2048, seed=64, device=dev, fused=fused)
        else:
            a = NeuroFlowCore(4, 2, 20, 128, buffer_size=2048, seed=64, device=dev, fused=fused)
        gg = torch.Generator().manual_seed(3)
        b = a.buffer
        b.states.copy_(torch.randn(b.states.shape, generator=gg))
        b.next_states.copy_(torch.randn(b.states.shape, generator=gg))
        if kind == "ct":
            b.actions.copy_(torch.rand(b.actions.shape, generator=gg) * 6 - 3)
        else:
            b.actions.copy_(torch.randint(0, 2, b.actions.shape, generator=gg).float())
        b.rewards.fill_(1.0)
        b.hiddens.copy_(0.3 * torch.randn(b.hiddens.shape, generator=gg))
        b.size = 2048
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        for _ in range(2):
            a._train_step(256)
        dp.configure(False)
        flat = torch.cat([p.detach().flatten() for p in list(a.actor.network.parameters()) + list(a.critic.network1.parameters())])
        return torch.cat([flat, a.entropy.log_alpha.detach().reshape(1)])

    for kind in ("ct", "discrete"):
        p_one = agent_params(kind, False, None)
        for name, oneshot in (("nccl", False), ("oneshot", None)):
            p_dp = agent_params(kind, True, oneshot)
            res[f"{kind}_{name}_max_abs"] = float((p_dp - p_one).abs().max())
            pair = [torch.empty_like(p_dp) for _ in range(world)]
            dist.all_gather(pair, p_dp)
            res[f"{kind}_{name}_replicas_equal"] = bool(torch.equal(pair[0], pair[1]))      # includes log_alpha (discrete: all-reduced too)
        # the fused agents (one-launch optimizer / policy-head / loss kernels, frozen critics in the actor loss): data-parallel == single process
        f_one = agent_params(kind, False, None, fused=True)
        f_dp = agent_params(kind, True, None, fused=True)
        res[f"{kind}_fused_max_abs"] = float((f_dp - f_one).abs().max())
        res[f"{kind}_fused_vs_unfused"] = float((f_one - p_one).abs().max())
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gradients_equal_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "dp_gpu.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    print("two-rank results:", r)
    assert r["liquid_allreduces"] == 1
    if r["oneshot_available"]:
        assert r["oneshot_used"] == 1 and r["oneshot_stress"] == 0.0, r
    assert r["oneshot_vs_nccl"] < 1e-6 * 50, r
    assert r["liquid_rel"] < 2e-6 and r["critic_rel"] < 2e-6, r       # fp32 summation order differs between 1 and 2 shards
    assert r["replicas_equal"]
    for kind in ("ct", "discrete"):
        for name in ("nccl", "oneshot"):
            assert r[f"{kind}_{name}_max_abs"] < 1e-6, r               # two Adam steps of 3e-4
            assert r[f"{kind}_{name}_replicas_equal"], r
        assert r[f"{kind}_fused_max_abs"] < 1e-6, r
        assert r[f"{kind}_fused_vs_unfused"] < 2e-6, r                 # the fused kernels follow the framework ops to fp32 round-off
