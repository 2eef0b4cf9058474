"""The N>1 data-parallel path on CPU: world_size 2, gloo.  Checks the sharding helper and the single flat
all-reduce (the same `maybe_allreduce` call the CUDA backward makes on the packed gradient block)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np

    from oracle import liquid_oracle as lo
    from velora_b200 import dataparallel as dp
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    # identical seeds on every rank => identical masks and parameters, no broadcast needed
    vb.set_seed(64)
    net = LiquidNCPNetwork(4, 20, 2)
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same_params = all(torch.equal(g, flat) for g in gathered)

    # global batch of 10 rows sharded contiguously; per-rank gradient blocks from the oracle; one all-reduce
    g = torch.Generator().manual_seed(5)
    B, T = 10, 3
    x = torch.randn(B, T, 4, generator=g).numpy()
    ry = torch.randn(B, T, 2, generator=g).numpy()
    p = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    dp.configure(True, reduce="sum")
    lo_, hi_ = dp.shard_bounds(B)
    y, ht, caches = lo.liquid_seq_forward(p, x[lo_:hi_], None, np.float64)
    _, _, grads = lo.liquid_seq_backward(p, caches, ry[lo_:hi_], np.zeros_like(ht), np.float64)
    block = torch.tensor(np.concatenate([grads[k].ravel() for k in sorted(grads)]))
    dp.maybe_allreduce(block)
    # reference: the whole batch on one rank
    y, ht, caches = lo.liquid_seq_forward(p, x, None, np.float64)
    _, _, full = lo.liquid_seq_backward(p, caches, ry, np.zeros_like(ht), np.float64)
    whole = np.concatenate([full[k].ravel() for k in sorted(full)])
    err = float(np.abs(block.numpy() - whole).max())
    # mean reduction
    dp.configure(True, reduce="mean")
    t = torch.full((3,), float(rank + 1))
    dp.maybe_allreduce(t)
    # policy noise: every rank draws the GLOBAL batch's noise and keeps its shard (functional.standard_normal_like)
    from velora_b200 import functional as VF

    torch.manual_seed(123)
    local = VF.standard_normal_like(torch.zeros(5, 3))
    shards = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(shards, local)
    torch.manual_seed(123)
    dp.configure(False)
    whole_noise = VF.standard_normal_like(torch.zeros(5 * world, 3))
    dp.configure(True, reduce="mean")
    noise_ok = bool(torch.equal(torch.cat(shards), whole_noise)) and dp.rank() == rank
    # discrete policy: every rank samples the GLOBAL batch from the gathered probabilities and keeps its rows (sac/discrete.py)
    from velora_b200.models.sac import SACActorDiscrete

    actor = SACActorDiscrete(4, 20, 3)
    gp = torch.Generator().manual_seed(77)
    probs_all = torch.softmax(torch.randn(6 * world, 3, generator=gp), dim=-1)
    torch.manual_seed(321)
    mine, _ = actor.get_sample(probs_all[rank * 6:(rank + 1) * 6])
    after_dp = torch.rand(1)                     # the generator has advanced exactly as in the single-process draw
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    dp.configure(False)
    torch.manual_seed(321)
    whole, _ = actor.get_sample(probs_all)
    after_single = torch.rand(1)
    dp.configure(True, reduce="mean")
    sample_ok = bool(torch.equal(torch.cat(got), whole)) and bool(torch.equal(after_dp, after_single))
    if rank == 0:
        torch.save({"same_params": same_params, "err": err, "bounds": (lo_, hi_), "mean": t.tolist(), "world": dp.world_size(),
                    "noise_ok": noise_ok, "sample_ok": sample_ok}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["world"] == 2 and r["same_params"]
    assert r["bounds"] == (0, 5)
    assert r["err"] < 1e-12          # sum of shard gradients == gradient of the whole batch
    assert r["mean"] == [1.5, 1.5, 1.5]
    assert r["noise_ok"]             # the ranks' noise shards concatenate to the single-process draw of the global batch
    assert r["sample_ok"]            # and so do the discrete policy's sampled actions; the generator advances identically


def test_shard_bounds_cover_batch():
    from velora_b200 import dataparallel as dp

    import pytest

    for n in (8, 64, 1_048_576):
        for w in (1, 2, 4, 8):
            spans = [dp.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert len({b - a for a, b in spans}) == 1          # equal shards: the noise slicing and reduce='mean' rely on it
    # uneven splits are refused instead of silently giving the last rank the wrong noise rows (B = 10 over 4 ranks)
    for n, w in ((10, 4), (7, 2), (1, 8)):
        with pytest.raises(ValueError, match="does not divide evenly"):
            dp.shard_bounds(n, 0, w)
