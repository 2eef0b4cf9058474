"""BASELINE config 4: LiquidNCPNetwork(8, 512, 1) forward + backward with bf16 tensor-core operands, batch 262144 x seq 64 over
8 GPUs = 32768 x 64 per GPU (weak scaling: every rank owns 32768 batch rows; N = 1 measures one GPU's share).

    python bench_cfg4.py [--steps K] [--warmup W] [--batch 32768] [--seq 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_cfg4.py --gpus N

One step = zero the gradients, `forward_seq_bf16` (vb_liquid_bf16_fwd, training variant), `autograd.backward` of
sum(y r_y) + sum(h_T r_h) (vb_liquid_bf16_bwd: tcgen05 recurrence + three cuBLAS GEMMs + the small kernels, then - for N > 1 - ONE NCCL
all-reduce of the 526 k-float gradient block, then vb_liquid_unpack_grads).  Timed with CUDA events, barrier + synchronize on both
sides, max over ranks.  One JSON line (rank 0): sample-steps/s and the fraction of the measured sustained bf16 tensor peak at SURVEY.md
8(d)'s algorithmic 3.149 MFLOP per sample-step (forward + dgrad + wgrad, dense).  The working set (7.8 GB of saved activations per step)
is far larger than L2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_SAMPLE_STEP = 6.0 * (308 * 8 + 5 * 204 * 512 + 1 * 204)       # 3.149 MFLOP


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32768, help="batch rows per GPU")
    ap.add_argument("--seq", type=int, default=64)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import velora_b200 as vb
    from velora_b200 import functional as VF
    from velora_b200.models.lnn import LiquidNCPNetwork

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        vb.dataparallel.configure(True, reduce="sum")
    B, T, K, W = args.batch, args.seq, args.steps, max(args.warmup, 3)
    vb.set_seed(64)
    net = LiquidNCPNetwork(8, 512, 1, device=dev)
    params = list(net.parameters())
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(T, B, 8, device=dev, generator=g).permute(1, 0, 2)          # time-major storage
    ry = torch.randn(T, B, 1, device=dev, generator=g).permute(1, 0, 2)
    rh = torch.randn(B, 204, device=dev, generator=g)

    def step():
        for p in params:
            p.grad = None
        y, hl = net.forward_seq_bf16(x, None)
        torch.autograd.backward([y, hl], [ry, rh])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    VF.PROFILE = {}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    prof, VF.PROFILE = VF.PROFILE, None
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    def avg(name):
        ev = prof.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)

    if rank == 0:
        peak_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        pk = json.load(open(peak_path)) if os.path.isfile(peak_path) else {"bf16_tflops_sustained": 1400.0}
        peak = float(pk.get("bf16_tflops_sustained", 1400.0))
        tflops = FLOPS_PER_SAMPLE_STEP * B * T / (ms * 1e-3) / 1e12           # per GPU
        line = {
            "metric": "liquid_512_fwd_bwd_sample_steps_per_s", "value": world * B * T / (ms * 1e-3), "unit": "sample-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"LiquidNCPNetwork(8,512,1) fwd+bwd, batch {B} x seq {T} per GPU, bf16 tcgen05 (BASELINE config 4)",
                       "batch_per_gpu": B, "seq_len": T},
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
                         "peak_kind": "measured sustained bf16 (MEASURED_PEAKS.json)" if os.path.isfile(peak_path) else "fallback",
                         "algorithmic_flops_per_sample_step": FLOPS_PER_SAMPLE_STEP, "per": "GPU"},
            "fwd_ms": avg("liquid_bf16_fwd"), "bwd_ms": avg("liquid_bf16_bwd"),
            "saved_activation_bytes_per_step": int(__import__("velora_b200")._lib.lib().vb_liquid_bf16_saved_bytes(8, 308, 204, 1, B, T)),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
