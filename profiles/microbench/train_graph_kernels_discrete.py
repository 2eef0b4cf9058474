"""Kernel list of one graph-captured NeuroFlow (discrete) train step at batch 64 (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from velora_b200.models.nf.agent import NeuroFlowCore

dev = torch.device("cuda:0")
agent = NeuroFlowCore(4, 2, 20, 128, buffer_size=100_000, seed=64, device=dev, capturable=True, fused=True)
g = torch.Generator(device=dev).manual_seed(0)
n = 100_000
agent.buffer.add_multi(torch.randn(n, 4, device=dev, generator=g), torch.randint(0, 2, (n, 1), device=dev, generator=g).float(),
                       torch.ones(n, device=dev), torch.randn(n, 4, device=dev, generator=g), (torch.rand(n, device=dev, generator=g) < 0.02).float(),
                       0.3 * torch.randn(n, 8, device=dev, generator=g))
replay = agent.capture_train_step(64)
torch.cuda.synchronize()
torch.cuda.profiler.start()
replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
