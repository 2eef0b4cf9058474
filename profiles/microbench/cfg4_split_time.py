import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import velora_b200 as vb
from velora_b200 import functional as VF
from velora_b200.models.lnn import LiquidNCPNetwork
dev = torch.device("cuda")
vb.set_seed(64)
net = LiquidNCPNetwork(8, 512, 1, device=dev)
B, T = 32768, 64
x = torch.randn(T, B, 8, device=dev).permute(1, 0, 2)
ry = torch.randn(T, B, 1, device=dev).permute(1, 0, 2)
rh = torch.randn(B, 204, device=dev)
def step():
    for p in net.parameters(): p.grad = None
    y, hl = net.forward_seq_bf16(x, None)
    torch.autograd.backward([y, hl], [ry, rh])
for _ in range(2): step()
VF.PROFILE = {}
for _ in range(3): step()
torch.cuda.synchronize()
print(os.environ.get("VELORA_B200_LIB", "x/default/x").split("/")[-2], {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in VF.PROFILE.items()})
