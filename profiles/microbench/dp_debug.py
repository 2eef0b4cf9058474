import os, sys
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import velora_b200 as vb
from velora_b200 import dataparallel as dp
from velora_b200.models.nf import NeuroFlowCTCore

def run(parallel, oneshot):
    dp.configure(parallel, reduce="mean", oneshot=oneshot)
    a = NeuroFlowCTCore(4, 1, 20, 128, torch.tensor([3.0]), torch.tensor([0.0]), buffer_size=2048, seed=64, device=dev)
    gg = torch.Generator().manual_seed(3)
    b = a.buffer
    b.states.copy_(torch.randn(b.states.shape, generator=gg)); b.next_states.copy_(torch.randn(b.states.shape, generator=gg))
    b.actions.copy_(torch.rand(b.actions.shape, generator=gg) * 6 - 3); b.rewards.fill_(1.0)
    b.hiddens.copy_(0.3 * torch.randn(b.hiddens.shape, generator=gg)); b.size = 2048
    torch.manual_seed(5); torch.cuda.manual_seed(5)
    out = []
    for _ in range(2):
        r = a._train_step(256)
        out.append([float(v) for v in r.values()])
    dp.configure(False)
    flat = torch.cat([p.detach().flatten() for p in list(a.actor.network.parameters()) + list(a.critic.network1.parameters())])
    return out, flat, float(a.entropy.log_alpha)

for name, par, one in (("single", False, None), ("nccl", True, False), ("oneshot", True, None)):
    o, f, la = run(par, one)
    if name == "single": ref = f
    print(rank, name, o, la, float((f - ref).abs().max()), flush=True)
dist.barrier(); dist.destroy_process_group()
