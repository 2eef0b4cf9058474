"""Kernel list of one graph-captured NeuroFlowCT train step at batch 64 (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_configs import make_agent

dev = torch.device("cuda:0")
agent = make_agent(dev, 100_000, capturable=True, fused=True)
replay = agent.capture_train_step(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
torch.cuda.synchronize()
torch.cuda.profiler.start()
replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
