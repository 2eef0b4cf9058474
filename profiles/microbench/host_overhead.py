import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import velora_b200 as vb
from velora_b200.models.lnn import LiquidNCPNetwork
dev = torch.device('cuda')
vb.set_seed(64); net = LiquidNCPNetwork(4, 20, 2, device=dev)
x = torch.randn(1, 4, device=dev); h = torch.zeros(1, 8, device=dev)
with torch.no_grad():
    for _ in range(200): net(x, h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): net(x, h)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("host us/call %.1f, incl. drain %.1f" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000): net(x, h)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('tottime').print_stats(18)
