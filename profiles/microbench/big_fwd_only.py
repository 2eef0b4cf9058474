"""One bf16 tensor-core forward of LiquidNCPNetwork(8,512,1) (for `ncu -k regex:liquid_big_fwd_kernel`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork

dev = torch.device("cuda:0")
vb.set_seed(64)
net = LiquidNCPNetwork(8, 512, 1, device=dev)
x = torch.randn(8, 37888, 8, device=dev).permute(1, 0, 2)
for _ in range(2):
    net.forward_seq_bf16(x)
torch.cuda.synchronize()
