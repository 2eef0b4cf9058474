"""One forward + backward of BASELINE config 3 (time-major, 65536 x 32) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:liquid_bwd_tc -s 2 -c 1 -o out python profiles/microbench/liquid_prof.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
vb.set_seed(64)
net = LiquidNCPNetwork(4, 20, 2, device=dev)
x = torch.randn(T, B, 4, device=dev).permute(1, 0, 2)
ry = torch.randn(T, B, 2, device=dev).permute(1, 0, 2)
rh = torch.randn(B, 8, device=dev)
for i in range(reps):
    net.zero_grad()
    y, ht, hl = net.forward_seq(x, None, return_last=True)
    torch.autograd.backward([y, hl], [ry, rh])
torch.cuda.synchronize()
print("done")
