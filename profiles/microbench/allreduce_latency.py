"""Latency of the one collective on the path: all-reduce of the flat 918-float gradient block, eager and inside a CUDA graph.
torchrun --nproc-per-node N profiles/microbench/allreduce_latency.py   (NCCL_ALGO / NCCL_PROTO from the environment)"""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(int(os.environ.get("N_FLOATS", "918")), device=dev)
for _ in range(20):
    dist.all_reduce(x)
torch.cuda.synchronize()


def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


eager = timed(lambda: dist.all_reduce(x), 200)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        dist.all_reduce(x)
graph = timed(g.replay, 20) / 20
if dist.get_rank() == 0:
    print(f"ALGO={os.environ.get('NCCL_ALGO','default')} PROTO={os.environ.get('NCCL_PROTO','default')} world={dist.get_world_size()} "
          f"floats={x.numel()}: eager {eager:.1f} us/all-reduce, in-graph {graph:.1f} us/all-reduce", flush=True)
del g
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
