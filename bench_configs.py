"""Secondary measurements for the other BASELINE.json configs (one GPU).  `bench.py` is the headline (config 3); this
script prints one JSON line per measurement so that profiles/ can quote configs 1/2/4/5 next to it.

    python bench_configs.py [--quick] [--no-cpu]

  cfg1/2  NeuroFlowCT train step, actor 20 / critic 128 neurons, batch 64 (InvertedPendulum shapes, synthetic buffer)
  cfg2b   batch-1 actor forward (the per-environment-step `predict`)
  cfg4    LiquidNCPNetwork(8, 512, 1) forward+backward (currently the generic fp32 tile kernels; tcgen05 path: next)
  cfg5    1M-transition replay buffer: fused gather of a 1M (and 131072) batch, and a full train step at those batches
CPU baselines use the oracle's torch-CPU ports (the Python reference cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p))
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def gpu_time_ms(fn, warmup=5, iters=20):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graph_time_ms(fn, reps=20):
    """Device time per call with `reps` calls captured in ONE CUDA graph: no host gaps between the launches.  Eager timing of a
    microsecond-scale kernel measures the Python / allocation cost per call instead (round 1 quoted 34 us for a 6.9 us gather)."""
    import torch

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time_ms(fn, warmup=1, iters=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def make_agent(dev, buffer_rows, seed=64, capturable=False, fused=False):
    import torch

    from velora_b200.models.nf import NeuroFlowCTCore

    agent = NeuroFlowCTCore(4, 1, 20, 128, torch.tensor([3.0]), torch.tensor([0.0]), buffer_size=buffer_rows, seed=seed, device=dev,
                            capturable=capturable, fused=fused)
    g = torch.Generator(device=dev).manual_seed(0)
    b = agent.buffer
    b.states.copy_(torch.randn(b.states.shape, device=dev, generator=g))
    b.next_states.copy_(torch.randn(b.states.shape, device=dev, generator=g))
    b.actions.copy_(torch.rand(b.actions.shape, device=dev, generator=g) * 6 - 3)
    b.rewards.fill_(1.0)
    b.dones.copy_((torch.rand(b.dones.shape, device=dev, generator=g) < 0.02).float())
    b.hiddens.copy_(0.3 * torch.randn(b.hiddens.shape, device=dev, generator=g))
    b.size = buffer_rows
    return agent


def cpu_train_oracle(agent):
    from oracle.train_oracle import TrainOracle

    sd = lambda net: {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}   # noqa: E731
    return TrainOracle(sd(agent.actor.network), sd(agent.critic.network1), sd(agent.critic.network2))


def main_dp(args):
    """cfg 5 at N GPUs: replicated 1M buffer, every rank draws the same global index vector and the same
    policy noise and keeps its shard; one all-reduce per network backward (2 critics + actor) and one for log_alpha; the whole step
    is one CUDA graph per rank.  Launched as `torchrun --nproc-per-node N bench_configs.py --dp`.  Before timing, rank-wise
    equivalence with the single-process step on the whole batch is checked."""
    import torch
    import torch.distributed as dist

    import velora_b200 as vb

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()

    # ---- equivalence: N-rank data-parallel step == single-process step on the global batch (same seed, noise and indices)
    gb = 4096 * world
    vb.dataparallel.configure(True, reduce="mean")
    a_dp = make_agent(dev, 100_000, seed=5, fused=True)
    torch.manual_seed(11)
    for _ in range(2):
        a_dp._train_step(gb)
    vb.dataparallel.configure(False)
    a_one = make_agent(dev, 100_000, seed=5, fused=True)
    torch.manual_seed(11)
    for _ in range(2):
        a_one._train_step(gb)
    worst = 0.0
    for pa, pb in zip(list(a_dp.actor.network.parameters()) + list(a_dp.critic.network1.parameters()) + [a_dp.entropy.log_alpha],
                      list(a_one.actor.network.parameters()) + list(a_one.critic.network1.parameters()) + [a_one.entropy.log_alpha]):
        worst = max(worst, float((pa.detach() - pb.detach()).abs().max()))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    equiv = float(t.item())
    del a_dp, a_one

    # ---- timing
    vb.dataparallel.configure(True, reduce="mean")
    per_gpu = 131072
    B = per_gpu * world
    rows = max(1_000_000, B)      # `sample` needs at least `batch` rows (replay.py:74-77): 2^20 rows at 8 GPUs
    big = make_agent(dev, rows, capturable=True, fused=True)
    replay = big.capture_train_step(B)
    for _ in range(30):
        replay()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for _ in range(iters):
        replay()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        emit(config=f"cfg5 sample + NeuroFlowCT update, data-parallel, global batch {B} = {per_gpu} x {world} GPUs, one CUDA graph per rank",
             metric="transitions_per_s", value=B / (float(ms.item()) * 1e-3), ms=float(ms.item()), n_gpus=world, buffer_rows=rows,
             max_abs_param_diff_vs_single_process_after_2_steps=equiv)
    replay = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dp", action="store_true", help="cfg5 data-parallel update under torchrun: global batch 131072 x world, one rank per GPU")
    args = ap.parse_args()
    if args.dp:
        return main_dp(args)
    import torch

    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda:0")
    pk = peaks()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    # ------------------------------------------------------------------ cfg 1/2: train step at batch 64
    agent = make_agent(dev, 100_000)
    ms = gpu_time_ms(lambda: agent._train_step(64), warmup=10, iters=50)
    cpu_ms = None
    if not args.no_cpu:
        o = cpu_train_oracle(agent)
        idx = torch.randint(0, 100_000, (64,))
        batch = {k: getattr(agent.buffer, k)[idx.to(dev)].cpu() for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens")}
        cpu_ms = cpu_time_ms(lambda: o.train_step(batch), warmup=3, iters=20)
    gagent = make_agent(dev, 100_000, capturable=True)
    replay = gagent.capture_train_step(64)
    gms = gpu_time_ms(replay, warmup=10, iters=200)
    emit(config="cfg1/2 NeuroFlowCT._train_step captured in one CUDA graph, batch 64", metric="us_per_train_step", value=1e3 * gms,
         cpu_port_us=None if cpu_ms is None else 1e3 * cpu_ms, cpu_cores=cores)
    fagent = make_agent(dev, 100_000, capturable=True, fused=True)
    freplay = fagent.capture_train_step(64)
    fms = gpu_time_ms(freplay, warmup=10, iters=200)
    emit(config="cfg1/2 NeuroFlowCT._train_step, one CUDA graph + fused Adam / Polyak kernels, batch 64", metric="us_per_train_step",
         value=1e3 * fms, graph_nodes=None)
    del fagent, freplay
    emit(config="cfg1/2 NeuroFlowCT._train_step, actor 20 / critic 128, batch 64", metric="us_per_train_step", value=1e3 * ms,
         cpu_port_us=None if cpu_ms is None else 1e3 * cpu_ms, cpu_cores=cores,
         note="sample (fused gather) + 2 liquid fwd + 1 liquid bwd + 6 NCP fwd + 4 NCP bwd through the C ABI; SAC head math, Adam, "
              "soft update are framework ops (next rows)")

    # ------------------------------------------------------------------ cfg 1: the discrete agent (CartPole shapes: 4 obs, 2 actions)
    from velora_b200.models.nf import NeuroFlowCore

    dagent = NeuroFlowCore(4, 2, 20, 128, buffer_size=100_000, seed=64, device=dev, capturable=True, fused=True)
    gd = torch.Generator(device=dev).manual_seed(0)
    db = dagent.buffer
    db.states.copy_(torch.randn(db.states.shape, device=dev, generator=gd))
    db.next_states.copy_(torch.randn(db.states.shape, device=dev, generator=gd))
    db.actions.copy_(torch.randint(0, 2, db.actions.shape, device=dev, generator=gd).float())
    db.rewards.fill_(1.0)
    db.dones.copy_((torch.rand(db.dones.shape, device=dev, generator=gd) < 0.02).float())
    db.hiddens.copy_(0.3 * torch.randn(db.hiddens.shape, device=dev, generator=gd))
    db.size = 100_000
    dms = gpu_time_ms(lambda: dagent._train_step(64), warmup=10, iters=50)
    dreplay = dagent.capture_train_step(64)
    dgms = gpu_time_ms(dreplay, warmup=10, iters=200)
    emit(config="cfg1 NeuroFlow (discrete) _train_step, actor 20 / critic 128, batch 64: one CUDA graph + fused Adam / Polyak",
         metric="us_per_train_step", value=1e3 * dgms, eager_us=1e3 * dms)
    del dagent, dreplay

    # ------------------------------------------------------------------ cfg 2b: batch-1 actor forward
    x1 = torch.randn(1, 4, device=dev)
    h1 = torch.zeros(1, 8, device=dev)
    with torch.no_grad():
        ms1 = gpu_time_ms(lambda: agent.actor.network.ncp(x1, h1), warmup=20, iters=200)
    emit(config="cfg2b LiquidNCPNetwork(4,20,2) forward, batch 1 (predict)", metric="us_per_call", value=1e3 * ms1)
    act = gagent.capture_predict(1)
    s1 = torch.randn(4, device=dev)
    hcur = [None]

    def act_once():
        _, hcur[0] = act(s1, hcur[0])

    ms1g = gpu_time_ms(act_once, warmup=20, iters=500)
    t0 = time.perf_counter()
    for _ in range(500):
        act_once()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 500
    emit(config="cfg2b NeuroFlowCT.predict captured in one CUDA graph, batch 1", metric="us_per_call", value=1e3 * ms1g, wall_us=1e6 * wall)

    # ---- the other per-environment-step call: ReplayBuffer.add
    from velora_b200.buffer import ReplayBuffer

    rb = ReplayBuffer(100_000, 4, 1, 8, device=dev)
    rb.add_multi(torch.randn(64, 4, device=dev), torch.randn(64, 1, device=dev), torch.ones(64, device=dev), torch.randn(64, 4, device=dev),
                 torch.zeros(64, device=dev), torch.randn(64, 8, device=dev))
    rb.gather(torch.arange(64, device=dev))          # shadow table built: every add keeps it current
    st, ac, ns, hd = torch.randn(4, device=dev), torch.randn(1, device=dev), torch.randn(4, device=dev), torch.randn(8, device=dev)
    add_us = {}
    for fused in (True, False):
        rb.fused_add = fused
        for _ in range(50):
            rb.add(st, ac, 1.0, ns, False, hd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(500):
            rb.add(st, ac, 1.0, ns, False, hd)
        torch.cuda.synchronize()
        add_us[fused] = 1e6 * (time.perf_counter() - t0) / 500
    rb.fused_add = True
    emit(config="cfg1/2 ReplayBuffer.add, one transition per call (with the packed shadow row)", metric="wall_us_per_call", value=add_us[True],
         six_indexing_assignments_plus_shadow_refresh_us=add_us[False], launches={"fused": 1, "indexing": 7})

    # ------------------------------------------------------------------ cfg 1 end to end: NeuroFlow("CartPole-v1", 20, 128).train(64, n_episodes=50)
    # gymnasium is not installed: the environment is the CartPole-v1 physics restatement of baseline/envs_shim.py (labelled as such)
    if not args.quick:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        from envs_shim import CartPoleShim

        from velora_b200.models import NeuroFlow

        eagent = NeuroFlow(CartPoleShim(64), 20, 128, buffer_size=100_000, device=dev, seed=64, fused=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        returns = eagent.train(64, n_episodes=50)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        n_env_steps = int(sum(returns))
        emit(config="cfg1 NeuroFlow(CartPole shim, 20, 128).train(64, n_episodes=50): act + env.step + buffer.add + one fused train step per env step (after three eager warm-up steps the act and the train step are graph replays)",
             metric="wall_s_for_50_episodes", value=wall, env_steps=n_env_steps, ms_per_env_step=1e3 * wall / max(n_env_steps, 1),
             mean_return_first10=sum(returns[:10]) / 10, mean_return_last10=sum(returns[-10:]) / 10,
             note="environment = baseline/envs_shim.py (gymnasium is not installed); includes the 128-transition warm-up")

    # ------------------------------------------------------------------ cfg 5: gather + train step at large batch
    big = make_agent(dev, 1_000_000)
    for B in ([131072] if args.quick else [131072, 1_000_000]):
        torch.manual_seed(1)
        idx = torch.randint(0, big.buffer.size, (B,), device=dev)
        from velora_b200 import functional as VF

        arrays = [getattr(big.buffer, k) for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens")]
        table, offsets, stride = big.buffer._shadow_table()
        pms = graph_time_ms(lambda: VF.gather_rows(arrays, idx, big.buffer.size))      # six separate arrays
        gms = graph_time_ms(lambda: VF.gather_packed(table, offsets, stride, [a.shape[1] for a in arrays], idx, big.buffer.size))   # packed table
        eager_ms = gpu_time_ms(lambda: big.buffer.gather(idx), warmup=5, iters=20)
        gbs = 160.0 * B / (gms * 1e-3) / 1e9
        ref_ms = gpu_time_ms(lambda: [getattr(big.buffer, k)[idx] for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens")],
                             warmup=3, iters=10)
        emit(config=f"cfg5 ReplayBuffer(1M,4,1,8) gather, batch {B}", metric="transitions_per_s", value=B / (gms * 1e-3), ms=gms,
             roofline={"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                       "algorithmic_bytes_per_transition": 160},
             six_array_gather_ms=pms, torch_six_index_kernels_ms=ref_ms, eager_call_ms=eager_ms,
             timing="20 gathers captured in one CUDA graph (kernel time); eager_call_ms includes the Python / allocation cost per call")
        tms = gpu_time_ms(lambda: big._train_step(B), warmup=3, iters=10)
        emit(config=f"cfg5 sample + NeuroFlowCT update, batch {B}, 1 GPU", metric="transitions_per_s", value=B / (tms * 1e-3), ms=tms)
        bigg = make_agent(dev, 1_000_000, capturable=True, fused=True)
        replay_big = bigg.capture_train_step(B)
        gms2 = gpu_time_ms(replay_big, warmup=3, iters=20)
        emit(config=f"cfg5 sample + NeuroFlowCT update, one CUDA graph + fused Adam / Polyak, batch {B}, 1 GPU", metric="transitions_per_s",
             value=B / (gms2 * 1e-3), ms=gms2)
        del bigg, replay_big
    if not args.no_cpu:
        B = 131072
        o = cpu_train_oracle(big)
        idxc = torch.randint(0, 1_000_000, (B,))
        cb = {k: getattr(big.buffer, k).cpu()[idxc] for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens")}
        cms = cpu_time_ms(lambda: o.train_step(cb), warmup=1, iters=2)
        emit(config=f"cfg5 CPU port of the NeuroFlowCT update, batch {B}", metric="transitions_per_s", value=B / (cms * 1e-3), ms=cms, cpu_cores=cores)
    del big

    # ------------------------------------------------------------------ cfg 3 incumbent: the reference's op sequence, eager on this GPU
    if not args.no_cpu:
        from oracle import liquid_oracle as lo      # measurement baseline only (the Python reference cannot travel to the GPU box)

        vb.set_seed(64)
        rnet = LiquidNCPNetwork(4, 20, 2, device=dev)
        rp = lo.torch_params({k: v.detach().clone() for k, v in rnet.state_dict().items()}, requires_grad=True)
        leaves = [t for t in rp.values() if t.requires_grad]
        B3, T3 = 65536, 32
        g3 = torch.Generator(device=dev).manual_seed(1234)
        x3 = torch.randn(B3, T3, 4, device=dev, generator=g3)
        ry3 = torch.randn(B3, T3, 2, device=dev, generator=g3)
        rh3 = torch.randn(B3, 8, device=dev, generator=g3)
        h03 = torch.zeros(B3, 8, device=dev)

        def ref_step():
            for t_ in leaves:
                t_.grad = None
            y_, ht_ = lo.liquid_seq_forward_torch(rp, x3, h03)
            ((y_ * ry3).sum() + (ht_[:, -1] * rh3).sum()).backward()

        rms = gpu_time_ms(ref_step, warmup=2, iters=5)
        emit(config="cfg3 the reference's ATen op sequence launched eagerly on this B200 (oracle torch port on cuda), 65536 x 32, fwd+bwd",
             metric="sample_steps_per_s", value=B3 * T3 / (rms * 1e-3), ms=rms,
             note="the incumbent a velora user has today on a GPU: ~26 launches per forward step, ~60 per backward step")
        del rnet, rp, leaves, x3, ry3, rh3, h03
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ cfg 4: 512-neuron liquid network
    vb.set_seed(64)
    net = LiquidNCPNetwork(8, 512, 1, device=dev)
    B, T = (1024, 4) if args.quick else (4096, 8)
    x = torch.randn(B, T, 8, device=dev)
    ry = torch.randn(B, T, 1, device=dev)

    def step4():
        for p in net.parameters():
            p.grad = None
        y, _, hl = net.forward_seq(x, None, return_last=True)
        torch.autograd.backward([y, hl], [ry, torch.ones_like(hl)])

    # forward only, bf16 operands on tcgen05 (the backward of this class is not built): per-GPU share of config 4 at 8 GPUs
    B4, T4 = (4096, 8) if args.quick else (32768, 64)
    x4 = torch.randn(T4, B4, 8, device=dev).permute(1, 0, 2)
    with torch.no_grad():
        msf = gpu_time_ms(lambda: net.forward_seq_bf16(x4), warmup=2, iters=5)
        xs4 = torch.randn(B, T, 8, device=dev)
        msf32 = gpu_time_ms(lambda: net.forward_seq(xs4), warmup=1, iters=3)
    fwd_flops = 2.0 * (8 * 308 + 517 * 816 + 204) * B4 * T4
    emit(config=f"cfg4 LiquidNCPNetwork(8,512,1) FORWARD, batch {B4} x seq {T4}, bf16 operands on tcgen05 (vb_liquid_bf16_fwd)",
         metric="sample_steps_per_s", value=B4 * T4 / (msf * 1e-3), ms=msf,
         roofline={"bound": "tensor", "achieved": fwd_flops / (msf * 1e-3) / 1e12, "peak": pk.get("bf16_tflops_sustained", 1400.0), "unit": "TFLOP/s",
                   "frac": fwd_flops / (msf * 1e-3) / 1e12 / pk.get("bf16_tflops_sustained", 1400.0),
                   "algorithmic_flops_per_sample_step": 2.0 * (8 * 308 + 517 * 816 + 204)},
         fp32_tile_kernels_forward_sample_steps_per_s=B * T / (msf32 * 1e-3),
         note="includes the per-call weight pack (fp32 -> bf16 plane order); forward / inference only")
    del x4, xs4
    ms4 = gpu_time_ms(step4, warmup=2, iters=5)
    flops = 3.149e6 * B * T
    emit(config=f"cfg4 LiquidNCPNetwork(8,512,1) fwd+bwd, batch {B} x seq {T}, fp32 generic tile kernels", metric="sample_steps_per_s",
         value=B * T / (ms4 * 1e-3), ms=ms4,
         roofline={"bound": "tensor", "achieved": flops / (ms4 * 1e-3) / 1e12, "peak": pk.get("bf16_tflops_sustained", 1400.0), "unit": "TFLOP/s",
                   "frac": flops / (ms4 * 1e-3) / 1e12 / pk.get("bf16_tflops_sustained", 1400.0)},
         note="the fp32 contract of the 512-neuron class (tile kernels); its bf16 tcgen05 forward + backward is timed by bench_cfg4.py")


if __name__ == "__main__":
    main()
