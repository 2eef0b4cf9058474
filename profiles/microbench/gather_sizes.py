"""Kernel time of the fused gather at several batch sizes, measured two ways: CUDA events around a CUDA-graph replay of 20 gathers
(no host gaps) and around 20 eager calls (what round 1 quoted: at 131072 rows that figure is the Python / allocation cost per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_configs import make_agent
dev = torch.device("cuda:0")
a = make_agent(dev, 1_000_000)
buf = a.buffer
for B in (64, 4096, 131072, 1_000_000):
    idx = torch.randint(0, 1_000_000, (B,), device=dev)
    for shadow in (True, False):
        buf.use_shadow = shadow
        for _ in range(3):
            buf.gather(idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            buf.gather(idx)
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 20 * 1e3
        # graph: 20 gathers of the six arrays (a captured gather reads the arrays directly; use the packed entry point explicitly)
        from velora_b200 import functional as VF
        if shadow:
            table, offsets, stride = buf._shadow_table()
            fn = lambda: VF.gather_packed(table, offsets, stride, [4, 1, 1, 4, 1, 8], idx, buf.size)
        else:
            arrays = [getattr(buf, n) for n in ("states", "actions", "rewards", "next_states", "dones", "hiddens")]
            fn = lambda: VF.gather_rows(arrays, idx, buf.size)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"B={B:8d} {'packed table' if shadow else 'six arrays  '}: {us:7.2f} us per gather in a graph ({160.0 * B / us / 1e3:7.1f} GB/s algorithmic), eager call {eager:7.2f} us")
