import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_configs import make_agent
dev = torch.device("cuda:0")
a = make_agent(dev, 1_000_000)
idx = torch.randint(0, 1_000_000, (1_000_000,), device=dev)
for _ in range(3):
    a.buffer.gather(idx)
torch.cuda.synchronize()
