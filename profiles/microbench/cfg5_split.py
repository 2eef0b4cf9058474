"""Where a batch-131072 NeuroFlowCT update spends its GPU time (CUDA events around every C-ABI call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_configs import make_agent
from velora_b200 import functional as VF

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
agent = make_agent(dev, 1_000_000)
for _ in range(3):
    agent._train_step(B)
torch.cuda.synchronize()
VF.PROFILE = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 5
for _ in range(N):
    agent._train_step(B)
e1.record()
torch.cuda.synchronize()
prof, VF.PROFILE = VF.PROFILE, None
print("total ms/step", e0.elapsed_time(e1) / N)
for k, ev in prof.items():
    ms = [a.elapsed_time(b) for a, b in ev]
    print(f"{k:12s} calls/step {len(ev)/N:4.1f}  avg {sum(ms)/len(ms):7.3f} ms  per step {sum(ms)/N:7.3f} ms")
