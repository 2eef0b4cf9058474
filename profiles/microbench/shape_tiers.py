"""fwd+bwd time of a liquid shape on the default (tensor-core) tier against the FFMA tier: python shape_tiers.py I N O [B T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork
I, N, O = (int(v) for v in sys.argv[1:4])
B, T = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (65536, 32)
dev = torch.device("cuda")
vb.set_seed(64)
net = LiquidNCPNetwork(I, N, O, device=dev)
x = torch.randn(B, T, I, device=dev); ry = torch.randn(B, T, O, device=dev); rh = torch.randn(B, net.hidden_size, device=dev)
xt = x.permute(1, 0, 2).contiguous().permute(1, 0, 2); ryt = ry.permute(1, 0, 2).contiguous().permute(1, 0, 2)
def run(flags, xx, rr):
    net.kernel_flags = flags
    def step():
        net.zero_grad()
        y, _, hl = net.forward_seq(xx, None, return_last=True)
        torch.autograd.backward([y, hl], [rr, rh])
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
print(f"({I},{N},{O}) {B}x{T}: default tier batch-major {run(0, x, ry):.3f} ms, time-major {run(0, xt, ryt):.3f} ms, FFMA tier {run(2, x, ry):.3f} ms (eager launches)")
