"""clock64 phase trace of liquid_fwd_tc_kernel / liquid_bwd_tc_kernel (library built with VB_TRACE_BUILD=1 into velora_b200/lib_trace):
    VELORA_B200_LIB=velora_b200/lib_trace/libvelora_b200.so python profiles/microbench/trace_bwd.py B"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork

B = int(sys.argv[1])
T = 32
dev = torch.device("cuda")
vb.set_seed(64)
net = LiquidNCPNetwork(4, 20, 2, device=dev)
x = torch.randn(T, B, 4, device=dev).permute(1, 0, 2)
ry = torch.randn(T, B, 2, device=dev).permute(1, 0, 2)
rh = torch.randn(B, 8, device=dev)
y, ht, hl = net.forward_seq(x, None, return_last=True)
torch.autograd.backward([y, hl], [ry, rh])
torch.cuda.synchronize()
