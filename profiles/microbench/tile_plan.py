"""Forward / backward kernel times of LiquidNCPNetwork(4,20,2), time-major, T = 32, for batch sizes around the 148-SM x 4-CTA capacity,
with full 128-sample tiles (VB_TILE_PLAN=0) or the balanced plan (default).  20 back-to-back launches between CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import velora_b200 as vb
from velora_b200 import functional as VF
from velora_b200.models.lnn import LiquidNCPNetwork

dev = torch.device("cuda")
T = 32
vb.set_seed(64)
net = LiquidNCPNetwork(4, 20, 2, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


BATCH_MAJOR = os.environ.get("BATCH_MAJOR", "0") == "1"
for B in [int(a) for a in sys.argv[1:]] or [148 * 3 * 128, 65536, 148 * 4 * 128, 148 * 2 * 128, 32768, 8192, 1024]:
    x = torch.randn(T, B, 4, device=dev).permute(1, 0, 2)
    ry = torch.randn(T, B, 2, device=dev).permute(1, 0, 2)
    if BATCH_MAJOR:
        x, ry = x.contiguous(), ry.contiguous()
    rh = torch.randn(B, 8, device=dev)

    def step():
        net.zero_grad()
        y, ht, hl = net.forward_seq(x, None, return_last=True)
        torch.autograd.backward([y, hl], [ry, rh])

    VF.PROFILE = None
    ms = timed(step)
    VF.PROFILE = {}
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    part = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in VF.PROFILE.items()}
    print(f"B={B:7d} tiles128={B / 128:7.1f} per_sm={B / 128 / 148:5.2f}  fwd {part['liquid_fwd']:.4f}  bwd {part['liquid_bwd']:.4f}  "
          f"fwd+bwd eager {ms:.4f} ms  ({B * T / ms / 1e6:.2f} G/s)", flush=True)
