import sys, torch
sys.path.insert(0, '/root/repo')
import velora_b200 as vb
from bench_configs import make_agent
dev = torch.device('cuda:0')
a = make_agent(dev, 4096, capturable=True)
def trycap(name, fn, warm=2):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, 'OK', flush=True)
    except Exception as e:
        print(name, 'FAIL', str(e).splitlines()[0], vb._lib.lib().vb_last_error(), flush=True)
        torch.cuda.synchronize()
which = sys.argv[1]
idx = torch.randint(0, 4096, (64,), device=dev)
b = a.buffer.gather(idx)
net = a.actor.network.ncp
if which in ('0', '1', '2'):
    net.kernel_flags = int(which)
    def f():
        with torch.no_grad(): return net(b.next_states, b.hiddens)
    trycap('liquid fwd flags ' + which, f)
if which == 'rs':
    trycap('rsample', lambda: torch.distributions.Normal(b.states, b.states.abs() + 1).rsample())
if which == 'sac':
    def f():
        with torch.no_grad(): return a.actor.network(b.next_states, b.hiddens)
    net.kernel_flags = 1
    trycap('sac actor generic', f)
