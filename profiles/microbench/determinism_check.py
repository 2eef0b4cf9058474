"""Run-to-run determinism of the batch-64 train step (eager and graph-replayed) on this box: the same seeded 40-step sequence, with the
per-environment-step calls (predict, add) interleaved, repeated N times; prints the largest parameter difference to the first run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from velora_b200.models.nf import NeuroFlowCore, NeuroFlowCTCore

dev = torch.device("cuda")


def run(kind, use_graph, steps=40):
    if kind == "ct":
        a = NeuroFlowCTCore(4, 1, 20, 128, torch.tensor([3.0]), torch.tensor([0.0]), buffer_size=2048, seed=3, device=dev, fused=True)
    else:
        a = NeuroFlowCore(4, 2, 20, 128, buffer_size=2048, seed=3, device=dev, fused=True)
    g = torch.Generator(device=dev).manual_seed(1)
    b = a.buffer
    for k in ("states", "next_states", "hiddens"):
        getattr(b, k).copy_(torch.randn(getattr(b, k).shape, device=dev, generator=g))
    if kind == "ct":
        b.actions.copy_(torch.rand(b.actions.shape, device=dev, generator=g) * 6 - 3)
    else:
        b.actions.copy_(torch.randint(0, 2, b.actions.shape, device=dev, generator=g).float())
    b.rewards.fill_(1.0)
    b.size, b.position = 2048, 0      # capture_train_step needs a full buffer; the eager runs start from the same state
    torch.manual_seed(99)
    state, hidden = torch.randn(4, device=dev), None
    # three eager steps first in both modes (capture_train_step's warm-up: the optimizer state must exist before the capture)
    if use_graph:
        replay = a.capture_train_step(64, warmup=3)
    else:
        replay = None
        for _ in range(3):
            a._train_step(64)
    for s in range(steps):
        action, hidden = a.predict(state, hidden, train_mode=True)
        b.add(torch.randn(4, device=dev), action.reshape(-1).float(), 1.0, torch.randn(4, device=dev), False, hidden.reshape(-1))
        replay() if use_graph else a._train_step(64)
    torch.cuda.synchronize()
    return torch.cat([p.detach().flatten() for net in (a.actor.network, a.critic.network1, a.critic.network2) for p in net.parameters()])


n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for kind in ("ct", "disc"):
    bases = {}
    for use_graph in (False, True):
        base = bases[use_graph] = run(kind, use_graph)
        diffs = [float((run(kind, use_graph) - base).abs().max()) for _ in range(n)]
        print(kind, "graph" if use_graph else "eager", "max |param diff| to the first run over", n, "repeats:", max(diffs), flush=True)
    print(kind, "graph replays vs eager steps from the same state and seed: max |param diff|", float((bases[True] - bases[False]).abs().max()), flush=True)
