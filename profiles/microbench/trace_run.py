import torch, velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork
import sys
B = int(sys.argv[1])
vb.set_seed(64); net = LiquidNCPNetwork(4,20,2, device=torch.device('cuda'))
x = torch.randn(B, 32, 4, device='cuda'); ry = torch.randn(B, 32, 2, device='cuda'); rh = torch.randn(B, 8, device='cuda')
for i in range(3):
    y, ht, hl = net.forward_seq(x, None, return_last=True)
    torch.autograd.backward([y, hl], [ry, rh])
torch.cuda.synchronize()
