"""Parity of the CUDA liquid path (through the public modules -> autograd.Function -> C ABI) with the reference
goldens and the oracle.  Run on the B200 box: pytest -m gpu."""
import copy

import numpy as np
import pytest
import torch

from oracle import liquid_oracle as lo
from vb_testutil import FP32_RTOL, assert_close_to_reference, liquid_case_params, rel_err

pytestmark = pytest.mark.gpu

CASES = ["fixture", "actor_t1", "actor_seq", "actor_seq_allh", "fixture_seq", "odd", "mid", "wide"]


def _build(d, name, flags=0):
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    i, n, o, sp, seed, B, T, allh = d[f"{name}_meta"]
    vb.set_seed(int(seed))
    net = LiquidNCPNetwork(int(i), int(n), int(o), sparsity_level=float(sp), device=torch.device("cuda"))
    # constructed from the same seed => identical parameters and masks (checked on CPU in test_host.py);
    # load the golden state anyway so that this test only measures the kernels
    net.load_state_dict({k: torch.tensor(v) for k, v in liquid_case_params(d, name).items()})
    net.kernel_flags = flags
    return net


@pytest.mark.parametrize("flags", [0, 2, 1], ids=["default", "ffma", "generic"])
@pytest.mark.parametrize("name", CASES)
def test_forward_backward_vs_reference(golden, name, flags):
    d = golden("liquid")
    net = _build(d, name, flags)
    dev = torch.device("cuda")
    x = torch.tensor(d[f"{name}_x"], device=dev, requires_grad=True)
    h0 = torch.tensor(d[f"{name}_h0"], device=dev, requires_grad=True)
    ry = torch.tensor(d[f"{name}_ry"], device=dev)
    rh = torch.tensor(d[f"{name}_rh"], device=dev)
    y, ht = net.forward_seq(x, h0)
    assert_close_to_reference(y.detach().cpu().numpy(), d[f"{name}_f32_y"], d[f"{name}_f64_y"], "y")
    assert_close_to_reference(ht.detach().cpu().numpy(), d[f"{name}_f32_htraj"], d[f"{name}_f64_htraj"], "h_traj")
    ((y * ry).sum() + (ht * rh).sum()).backward()
    assert_close_to_reference(x.grad.cpu().numpy(), d[f"{name}_f32_dx"], d[f"{name}_f64_dx"], "dx")
    assert_close_to_reference(h0.grad.cpu().numpy(), d[f"{name}_f32_dh0"], d[f"{name}_f64_dh0"], "dh0")
    for k, p in net.named_parameters():
        g = p.grad.cpu().numpy()
        assert_close_to_reference(g, d[f"{name}_f32_grad_{k}"], d[f"{name}_f64_grad_{k}"], f"grad {k}")
        if k.endswith("weight"):
            mask = net.state_dict()[k.replace("weight", "mask")].cpu().numpy()
            assert np.all(g[mask == 0] == 0.0), f"{k}: masked gradient entries must be exactly zero"


@pytest.mark.parametrize("name", ["actor_seq", "fixture_seq"])
def test_h_last_output_matches_dense_dh(golden, name):
    """Loss on h_T through the separate h_last output == loss on h_traj[:, -1]."""
    d = golden("liquid")
    dev = torch.device("cuda")
    x = torch.tensor(d[f"{name}_x"], device=dev)
    h0 = torch.tensor(d[f"{name}_h0"], device=dev)
    r = torch.tensor(d[f"{name}_rh"], device=dev)[:, -1]
    grads = []
    for use_last in (True, False):
        net = _build(d, name)
        y, ht, hl = net.forward_seq(x, h0, return_last=True)
        assert torch.equal(hl, ht[:, -1])
        ((hl if use_last else ht[:, -1]) * r).sum().backward()
        grads.append(torch.cat([p.grad.flatten() for p in net.parameters()]).cpu().numpy())
    assert rel_err(grads[0], grads[1]) < 1e-6


def test_forward_api_conventions(golden):
    """T=1 API of ncp.py:129-178: h=None => zeros, B==1 squeeze, CPU input moved to the device, ValueError on 3-D."""
    d = golden("liquid")
    net = _build(d, "fixture")
    x = torch.tensor(d["fixture_x"][:, 0])
    y, h = net(x, None)          # CPU input, like reference tests/models/test_lnn.py:367-381
    assert y.is_cuda and h.is_cuda and y.shape == (7, 5) and h.shape == (7, 8)
    assert rel_err(y.detach().cpu().numpy(), d["fixture_f32_y_h0none"]) < FP32_RTOL
    assert rel_err(h.detach().cpu().numpy(), d["fixture_f32_h_h0none"]) < FP32_RTOL
    y1, h1 = net(x[:1])
    assert y1.shape == (5,) and h1.shape == (1, 8)     # reference tests/models/test_lnn.py:331-340
    with pytest.raises(ValueError):
        net(torch.zeros(2, 3, 10))
    # anchors of SURVEY.md 8c
    xa = torch.tensor(d["fixture_x"][:, 0], device="cuda")
    ha = torch.tensor(d["fixture_h0"], device="cuda")
    ya, _ = net(xa, ha)
    assert abs(float(ya.sum()) - float(d["anchor_sum_y"][0])) < 2e-5
    assert np.allclose(ya[0].detach().cpu().numpy(), d["anchor_y0"], atol=2e-6)


def test_deepcopy_and_jit_script_run_on_gpu(golden):
    """velora/models/nf/modules.py:370-392 deep-copies then jit.scripts the networks."""
    d = golden("liquid")
    net = _build(d, "actor_t1")
    x = torch.tensor(d["actor_t1_x"][:, 0], device="cuda")
    h = torch.tensor(d["actor_t1_h0"], device="cuda")
    y0, h0 = net(x, h)
    twin = torch.jit.script(copy.deepcopy(net))
    assert isinstance(twin, torch.jit.RecursiveScriptModule)
    y1, h1 = twin(x, h)
    assert torch.equal(y0, y1) and torch.equal(h0, h1)
    (y1.sum() + h1.sum()).backward()
    assert all(p.grad is not None for p in twin.parameters())
    assert list(twin.state_dict().keys()) == list(net.state_dict().keys())


def test_seq_equals_looped_single_steps(golden):
    d = golden("liquid")
    net = _build(d, "actor_seq")
    x = torch.tensor(d["actor_seq_x"], device="cuda")
    h = torch.tensor(d["actor_seq_h0"], device="cuda")
    y, ht = net.forward_seq(x, h)
    for t in range(x.shape[1]):
        yt, h = net(x[:, t], h)
        assert torch.equal(yt, y[:, t]) and torch.equal(h, ht[:, t])


@pytest.mark.parametrize("B,T", [(1, 1), (31, 3), (33, 9), (257, 5), (4096, 32)])
def test_ragged_batches_vs_oracle(golden, B, T):
    """Batch sizes around the 32-row warp tiles, time lengths around the staging chunks; oracle = numpy fp64."""
    d = golden("liquid")
    net = _build(d, "actor_seq")
    p = liquid_case_params(d, "actor_seq")
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, 4, generator=g)
    h0 = 0.5 * torch.randn(B, 8, generator=g)
    ry = torch.randn(B, T, 2, generator=g)
    rh = torch.randn(B, 8, generator=g)
    yo, hto, caches = lo.liquid_seq_forward(p, x.numpy(), h0.numpy(), np.float64)
    dh = np.zeros_like(hto)
    dh[:, -1] = rh.numpy()
    dxo, dh0o, go = lo.liquid_seq_backward(p, caches, ry.numpy(), dh, np.float64)
    for flags in (0, 2, 1):
        net.kernel_flags = flags
        net.zero_grad()
        xc = x.cuda().requires_grad_(True)
        hc = h0.cuda().requires_grad_(True)
        y, ht, hl = net.forward_seq(xc, hc, return_last=True)
        ((y * ry.cuda()).sum() + (hl * rh.cuda()).sum()).backward()
        assert rel_err(y.detach().cpu().numpy(), yo) < FP32_RTOL
        assert rel_err(ht.detach().cpu().numpy(), hto) < FP32_RTOL
        assert rel_err(xc.grad.cpu().numpy(), dxo) < FP32_RTOL
        assert rel_err(hc.grad.cpu().numpy(), dh0o) < FP32_RTOL
        for k, pr in net.named_parameters():
            assert rel_err(pr.grad.cpu().numpy(), go[k]) < FP32_RTOL, (k, flags)


def test_full_size_properties():
    """BASELINE config 3 at full size (B=65536, T=32): three independent kernel families agree (tensor-core, FFMA,
    generic), gradients are linear in the upstream gradient, and the run is bit-reproducible (no atomics)."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    vb.set_seed(64)
    net = LiquidNCPNetwork(4, 20, 2, device=torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(1234)
    B, T = 65536, 32
    x = torch.randn(B, T, 4, device="cuda", generator=g)
    ry = torch.randn(B, T, 2, device="cuda", generator=g)

    def run(flags, scale):
        net.kernel_flags = flags
        net.zero_grad()
        y, ht, hl = net.forward_seq(x, None, return_last=True)
        ((y * ry).sum() * scale).backward()
        return y.detach(), ht.detach(), torch.cat([p.grad.flatten() for p in net.parameters()])

    y0, h0, g0 = run(0, 1.0)
    for other in (1, 2):          # generic tile kernels, specialised FFMA kernels
        y1, h1, g1 = run(other, 1.0)
        assert rel_err(y0.cpu().numpy(), y1.cpu().numpy()) < 2e-6
        assert rel_err(h0.cpu().numpy(), h1.cpu().numpy()) < 2e-6
        assert rel_err(g0.cpu().numpy(), g1.cpu().numpy()) < FP32_RTOL
    _, _, g2 = run(0, 2.0)
    assert rel_err(g2.cpu().numpy(), 2.0 * g0.cpu().numpy()) < 1e-6
    y3, h3, g3 = run(0, 1.0)
    assert torch.equal(y0, y3) and torch.equal(h0, h3) and torch.equal(g0, g3)


@pytest.mark.parametrize("name", ["actor_seq", "actor_seq_allh", "fixture_seq"])
def test_time_major_storage_matches_reference(golden, name):
    """x and the upstream gradients stored time-major ([T,B,F], the stack of per-step batches) are read in place:
    same results as the reference, dx comes back in x's storage order, y / h_traj are [B,T,F] views."""
    d = golden("liquid")
    net = _build(d, name)
    dev = torch.device("cuda")

    def tm(a):   # [B,T,F] values in time-major storage
        return torch.tensor(a, device=dev).permute(1, 0, 2).contiguous().permute(1, 0, 2)

    x = tm(d[f"{name}_x"]).requires_grad_(True)
    assert not x.is_contiguous()
    h0 = torch.tensor(d[f"{name}_h0"], device=dev, requires_grad=True)
    ry, rh = tm(d[f"{name}_ry"]), tm(d[f"{name}_rh"])
    y, ht = net.forward_seq(x, h0)
    assert y.shape == tuple(d[f"{name}_f32_y"].shape) and ht.shape == tuple(d[f"{name}_f32_htraj"].shape)
    assert_close_to_reference(y.detach().cpu().numpy(), d[f"{name}_f32_y"], d[f"{name}_f64_y"], "y")
    assert_close_to_reference(ht.detach().cpu().numpy(), d[f"{name}_f32_htraj"], d[f"{name}_f64_htraj"], "h_traj")
    torch.autograd.backward([y, ht], [ry, rh])
    assert x.grad.stride() == x.stride()
    assert_close_to_reference(x.grad.cpu().numpy(), d[f"{name}_f32_dx"], d[f"{name}_f64_dx"], "dx")
    assert_close_to_reference(h0.grad.cpu().numpy(), d[f"{name}_f32_dh0"], d[f"{name}_f64_dh0"], "dh0")
    for k, p in net.named_parameters():
        assert_close_to_reference(p.grad.cpu().numpy(), d[f"{name}_f32_grad_{k}"], d[f"{name}_f64_grad_{k}"], f"grad {k}")


def test_time_major_flag_rejected_off_the_tensor_core_path():
    """The layout flags are a contract of the tcgen05 kernels; the C ABI refuses them elsewhere instead of misreading rows."""
    from velora_b200 import _lib

    L = _lib.lib()
    assert L.vb_liquid_time_major_ok(4, 12, 8, 2, 0) == 1
    assert L.vb_liquid_time_major_ok(4, 12, 8, 2, 2) == 0      # FORCE_FFMA
    assert L.vb_liquid_time_major_ok(8, 300, 212, 1, 0) == 0   # no tensor-core kernel for this shape
    n = L.vb_liquid_packed_floats(4, 12, 8, 2)
    packed = torch.zeros(n, device="cuda")
    x = torch.zeros(8, 3, 4, device="cuda")
    y = torch.zeros(8, 3, 2, device="cuda")
    h = torch.zeros(8, 3, 8, device="cuda")
    rc = L.vb_liquid_fwd(_lib.ptr(packed), _lib.ptr(x), None, _lib.ptr(y), _lib.ptr(h), 8, 3, 4, 12, 8, 2, None, 0, 2 | 16, None)
    assert rc == -6 and b"time-major" in L.vb_last_error()


def test_empty_batch(golden):
    """Zero rows: outputs are empty, every parameter gradient is exactly zero (what the reference's ops give for B = 0)."""
    d = golden("liquid")
    net = _build(d, "actor_seq")
    dev = torch.device("cuda")
    y, h = net(torch.zeros(0, 4, device=dev), None)
    assert y.shape == (0, 2) and h.shape == (0, 8)
    x = torch.zeros(0, 5, 4, device=dev, requires_grad=True)
    ys, ht = net.forward_seq(x, None)
    assert ys.shape == (0, 5, 2) and ht.shape == (0, 5, 8)
    (ys.sum() + ht.sum()).backward()
    assert x.grad.shape == (0, 5, 4)
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.count_nonzero(p.grad) == 0, k


@pytest.mark.parametrize("B,T,with_h0", [(300, 5, True), (128, 1, False), (1000, 12, False)])
def test_bf16_tensor_core_forward_of_the_512_neuron_class(B, T, with_h0):
    """vb_liquid_bf16_fwd (tcgen05, bf16 operands / fp32 accumulation and state) against the fp32 path of the same network:
    2e-2 relative (BASELINE's bf16 contract) for the outputs, the last hidden state and the whole hidden trajectory."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda")
    vb.set_seed(64)
    net = LiquidNCPNetwork(8, 512, 1, device=dev)
    g = torch.Generator(device=dev).manual_seed(B + T)
    x = torch.randn(B, T, 8, device=dev, generator=g)
    h0 = 0.5 * torch.randn(B, 204, device=dev, generator=g) if with_h0 else None
    with torch.no_grad():
        y_ref, ht_ref = net.forward_seq(x, h0)
    y, h_last, ht = net.forward_seq_bf16(x, h0, return_traj=True)
    assert y.shape == (B, T, 1) and h_last.shape == (B, 204) and ht.shape == (B, T, 204)
    assert rel_err(y.cpu().numpy(), y_ref.cpu().numpy()) < 2e-2
    assert rel_err(ht.cpu().numpy(), ht_ref.cpu().numpy()) < 2e-2
    assert rel_err(h_last.cpu().numpy(), ht_ref[:, -1].cpu().numpy()) < 2e-2
    # and against the oracle that applies exactly the kernel's operand rounding (bf16 z~ and head weights, exact accumulation):
    # what is left is accumulation order and an occasional one-ulp difference in the rounding of z~
    rows = slice(0, min(B, 160))
    p = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    yo, ho = lo.liquid_seq_forward_bf16(p, x[rows].cpu().numpy(), None if h0 is None else h0[rows].cpu().numpy())
    assert rel_err(y[rows].cpu().numpy(), yo) < 2e-3
    assert rel_err(ht[rows].cpu().numpy(), ho) < 2e-3
    with torch.no_grad():      # (with autograd the training variant of the kernel runs: same contract, not the same bits)
        y2, h2 = net.forward_seq_bf16(x.permute(1, 0, 2).contiguous().permute(1, 0, 2), h0)      # time-major input, no trajectory
    assert torch.equal(y2, y) and torch.equal(h2, h_last)
    with pytest.raises(RuntimeError):
        LiquidNCPNetwork(4, 20, 2, device=dev).forward_seq_bf16(torch.zeros(2, 3, 4, device=dev))


# ------------------------------------------------------------------------------------------------ BASELINE config 4's network vs the reference
def _big(golden):
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    d = golden("liquid_big")
    vb.set_seed(64)
    net = LiquidNCPNetwork(8, 512, 1, device=torch.device("cuda"))
    net.load_state_dict({k: torch.tensor(v) for k, v in liquid_case_params(d, "big").items()})
    return d, net


def _close_to_f64(ours, d, key, what, rtol=FP32_RTOL):
    """The big golden keeps the fp64 arbiter and the reference's own fp32 distance from it (e32_*): within rtol of the arbiter and
    no further from it than 4x the fp32 reference."""
    e = rel_err(ours, d[f"big_f64_{key}"])
    assert e <= max(4.0 * float(d[f"big_e32_{key}"][0]), rtol / 4) and e <= rtol, f"{what}: {e:.3e} (reference fp32: {float(d[f'big_e32_{key}'][0]):.3e})"


def test_512_neuron_network_fp32_path_vs_reference(golden):
    """LiquidNCPNetwork(8,512,1) - the network BASELINE config 4 trains - on the fp32 tile kernels (fp32 atomics for the 526 k-float
    gradient block) against the reference's golden at that shape: outputs, data gradients and all 14 parameter gradients."""
    d, net = _big(golden)
    dev = torch.device("cuda")
    x = torch.tensor(d["big_x"], device=dev, requires_grad=True)
    h0 = torch.tensor(d["big_h0"], device=dev, requires_grad=True)
    y, ht = net.forward_seq(x, h0)
    assert_close_to_reference(y.detach().cpu().numpy(), d["big_f32_y"], d["big_f64_y"], "y")
    assert_close_to_reference(ht.detach().cpu().numpy(), d["big_f32_htraj"], d["big_f64_htraj"], "h_traj")
    ((y * torch.tensor(d["big_ry"], device=dev)).sum() + (ht * torch.tensor(d["big_rh"], device=dev)).sum()).backward()
    assert_close_to_reference(x.grad.cpu().numpy(), d["big_f32_dx"], d["big_f64_dx"], "dx")
    assert_close_to_reference(h0.grad.cpu().numpy(), d["big_f32_dh0"], d["big_f64_dh0"], "dh0")
    for k, p in net.named_parameters():
        g = p.grad.cpu().numpy()
        _close_to_f64(g, d, f"grad_{k}", f"grad {k}")
        if k.endswith("weight"):
            mask = net.state_dict()[k.replace("weight", "mask")].cpu().numpy()
            assert np.all(g[mask == 0] == 0.0), k


def test_512_neuron_bf16_forward_vs_fp64_reference(golden):
    """forward_seq_bf16 (tcgen05, bf16 operands) against the fp64 deep copy of the REFERENCE at (8,512,1): BASELINE's 2e-2."""
    d, net = _big(golden)
    dev = torch.device("cuda")
    x, h0 = torch.tensor(d["big_x"], device=dev), torch.tensor(d["big_h0"], device=dev)
    y, h_last, ht = net.forward_seq_bf16(x, h0, return_traj=True)
    assert rel_err(y.cpu().numpy(), d["big_f64_y"]) < 2e-2
    assert rel_err(ht.cpu().numpy(), d["big_f64_htraj"]) < 2e-2
    assert rel_err(h_last.cpu().numpy(), d["big_f64_htraj"][:, -1]) < 2e-2


# ------------------------------------------------------------------------------------------------ the benchmarked configuration itself
@pytest.mark.parametrize("layout", ["time_major", "batch_major"])
def test_benchmarked_configuration_vs_fp64_oracle(layout):
    """bench.py's workload as bench.py runs it - LiquidNCPNetwork(4,20,2), 65536 x 32, h0 = None, loss sum(y r_y) + sum(h_T r_h),
    sequence tensors time-major - compared with the fp64 oracle on the first 4096 rows: outputs and dx
    row for row; the parameter gradients of the full batch against a second full-size run whose upstream gradient is zero outside
    those rows (so the kernels run at the benchmarked size and the oracle only has to cover 4096 rows)."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda")
    vb.set_seed(64)
    net = LiquidNCPNetwork(4, 20, 2, device=dev)
    B, T, R = 65536, 32, 4096
    g = torch.Generator(device=dev).manual_seed(1234)

    def seq(F):
        if layout == "time_major":
            return torch.randn(T, B, F, device=dev, generator=g).permute(1, 0, 2)
        return torch.randn(B, T, F, device=dev, generator=g)

    x = seq(4).requires_grad_(True)
    ry = seq(2)
    rh = torch.randn(B, 8, device=dev, generator=g)
    keep = torch.zeros(B, 1, 1, device=dev)
    keep[:R] = 1.0
    ry_m = ry * keep                                   # same storage order as ry
    rh_m = rh * keep[:, 0]
    y, ht, hl = net.forward_seq(x, None, return_last=True)
    torch.autograd.backward([y, hl], [ry_m, rh_m])
    p = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    xo = x[:R].detach().cpu().numpy()
    yo, hto, caches = lo.liquid_seq_forward(p, xo, None, np.float64)
    dh = np.zeros_like(hto)
    dh[:, -1] = rh[:R].cpu().numpy()
    dxo, _, go = lo.liquid_seq_backward(p, caches, ry[:R].cpu().numpy(), dh, np.float64)
    assert rel_err(y[:R].detach().cpu().numpy(), yo) < FP32_RTOL
    assert rel_err(ht[:R].detach().cpu().numpy(), hto) < FP32_RTOL
    assert rel_err(x.grad[:R].cpu().numpy(), dxo) < FP32_RTOL
    assert torch.count_nonzero(x.grad[R:]) == 0
    for k, pr in net.named_parameters():
        assert rel_err(pr.grad.cpu().numpy(), go[k]) < FP32_RTOL, k
    # the unmasked full-batch run (what the bench times) is the masked run plus the rest: additivity over disjoint row sets
    g_first = torch.cat([pr.grad.flatten() for pr in net.parameters()]).clone()
    net.zero_grad()
    y2, _, hl2 = net.forward_seq(x.detach(), None, return_last=True)
    torch.autograd.backward([y2, hl2], [ry - ry_m, rh - rh_m])
    g_rest = torch.cat([pr.grad.flatten() for pr in net.parameters()]).clone()
    net.zero_grad()
    y3, _, hl3 = net.forward_seq(x.detach(), None, return_last=True)
    torch.autograd.backward([y3, hl3], [ry, rh])
    g_all = torch.cat([pr.grad.flatten() for pr in net.parameters()])
    assert torch.equal(y3, y) and rel_err(g_all.cpu().numpy(), (g_first + g_rest).cpu().numpy()) < 5e-6


# ------------------------------------------------------------------------------------------------ re-entrancy and the saved block
def test_two_networks_on_two_streams_and_two_graphs_do_not_interfere(golden):
    """The tensor-core kernels keep no process-global device state (round 1 staged the inter / motor weights in a __constant__
    bank): two DIFFERENT networks run concurrently on two streams - eagerly and as two CUDA graphs replayed at the same time -
    and each reproduces what it computes alone."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda")
    nets, xs, want = [], [], []
    for seed in (64, 65):
        vb.set_seed(seed)
        net = LiquidNCPNetwork(4, 20, 2, device=dev)
        g = torch.Generator(device=dev).manual_seed(seed)
        x = torch.randn(4096, 16, 4, device=dev, generator=g)
        nets.append(net)
        xs.append(x)

    def run(k):
        net, x = nets[k], xs[k]
        net.zero_grad()
        y, ht, hl = net.forward_seq(x, None, return_last=True)
        (y.sum() + hl.sum()).backward()
        return y.detach().clone(), torch.cat([p.grad.flatten() for p in net.parameters()]).clone()

    for k in range(2):
        want.append(run(k))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    for rep in range(5):                                   # eager, interleaved launches on two streams
        got = [None, None]
        for k in range(2):
            streams[k].wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(streams[k]):
                got[k] = run(k)
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(got[k][0], want[k][0]) and torch.equal(got[k][1], want[k][1]), (rep, k)
    graphs, outs = [], []
    for k in range(2):                                     # one graph per network, replayed concurrently
        with torch.cuda.stream(streams[k]):
            run(k)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=streams[k]):
            outs.append(run(k))
        graphs.append(gr)
    for rep in range(5):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                graphs[k].replay()
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(outs[k][0], want[k][0]) and torch.equal(outs[k][1], want[k][1]), ("graph", rep, k)


def test_backward_without_a_saved_block_reruns_the_forward(golden):
    """C ABI: vb_liquid_bwd(saved = NULL) re-runs the forward into its workspace and gives bit-identical gradients to the call
    that is handed the block the forward wrote."""
    from velora_b200 import _lib

    d = golden("liquid")
    net = _build(d, "actor_seq")
    dev = torch.device("cuda")
    x = torch.tensor(d["actor_seq_x"], device=dev, requires_grad=True)
    h0 = torch.tensor(d["actor_seq_h0"], device=dev)
    y, ht = net.forward_seq(x, h0)
    x_s, h0_s, h_traj, packed, gates = y.grad_fn.saved_tensors
    assert gates.numel() > 0
    L = _lib.lib()
    B, T = x.shape[0], x.shape[1]
    dy = torch.tensor(d["actor_seq_ry"], device=dev).permute(1, 0, 2).contiguous().permute(1, 0, 2)
    lay = 32 | 64        # dy and h_traj time-major (how LiquidSeqFn allocates them); x batch-major
    outs = []
    for have in (1, 0):
        ws = torch.empty(L.vb_liquid_bwd_workspace_bytes(4, 12, 8, 2, B, T, have), dtype=torch.uint8, device=dev)
        dx = torch.empty_like(x_s)
        dp = torch.empty(L.vb_liquid_grad_floats(4, 12, 8, 2), device=dev)
        _lib.check(L.vb_liquid_bwd(_lib.ptr(packed), _lib.ptr(x_s), _lib.ptr(h0_s), _lib.ptr(h_traj), _lib.ptr(dy), None, None,
                                   _lib.ptr(gates) if have else None, gates.numel() if have else 0, _lib.ptr(dx), None, _lib.ptr(dp),
                                   B, T, 4, 12, 8, 2, _lib.ptr(ws), ws.numel(), lay, None), "vb_liquid_bwd")
        torch.cuda.synchronize()
        outs.append((dx, dp))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    small = torch.empty(1024, dtype=torch.uint8, device=dev)
    rc = L.vb_liquid_bwd(_lib.ptr(packed), _lib.ptr(x_s), _lib.ptr(h0_s), _lib.ptr(h_traj), _lib.ptr(dy), None, None, None, 0,
                         None, None, _lib.ptr(outs[0][1]), B, T, 4, 12, 8, 2, _lib.ptr(small), small.numel(), lay, None)
    assert rc == -3 and b"workspace too small" in L.vb_last_error()


# ------------------------------------------------------------------------------------------------ config 4: trainable bf16 path
def test_512_neuron_bf16_training_step_vs_fp64_reference(golden):
    """forward_seq_bf16 with autograd (vb_liquid_bf16_fwd saving gate factors + vb_liquid_bf16_bwd: tcgen05 recurrence, batched GEMMs)
    against the fp64 deep copy of the REFERENCE at (8,512,1): outputs, dx, dh0 and all 14 parameter gradients within BASELINE's
    bf16 contract of 2e-2 relative; masked weight gradients exactly zero."""
    d, net = _big(golden)
    dev = torch.device("cuda")
    x = torch.tensor(d["big_x"], device=dev, requires_grad=True)
    h0 = torch.tensor(d["big_h0"], device=dev, requires_grad=True)
    ry = torch.tensor(d["big_ry"], device=dev)
    rh = torch.tensor(d["big_rh"], device=dev)
    # the golden's loss weights every h_t; the bf16 path exposes h_T only: check it on the loss sum(y ry) + sum(h_T rh_T) against the
    # fp64 ORACLE (pinned to the reference to 1e-11 at this shape in tests/test_oracle.py)
    p = liquid_case_params(d, "big")
    yo, hto, caches = lo.liquid_seq_forward(p, d["big_x"], d["big_h0"], np.float64)
    dh = np.zeros_like(hto)
    dh[:, -1] = d["big_rh"][:, -1]
    dxo, dh0o, go = lo.liquid_seq_backward(p, caches, d["big_ry"], dh, np.float64)
    y, h_last = net.forward_seq_bf16(x, h0)
    assert rel_err(y.detach().cpu().numpy(), d["big_f64_y"]) < 2e-2
    assert rel_err(h_last.detach().cpu().numpy(), d["big_f64_htraj"][:, -1]) < 2e-2
    ((y * ry).sum() + (h_last * rh[:, -1]).sum()).backward()
    assert rel_err(x.grad.cpu().numpy(), dxo) < 2e-2
    assert rel_err(h0.grad.cpu().numpy(), dh0o) < 2e-2
    for k, pr in net.named_parameters():
        g = pr.grad.cpu().numpy()
        assert rel_err(g, go[k]) < 2e-2, (k, rel_err(g, go[k]))
        if k.endswith("weight"):
            mask = net.state_dict()[k.replace("weight", "mask")].cpu().numpy()
            assert np.all(g[mask == 0] == 0.0), k


@pytest.mark.parametrize("B,T", [(300, 5), (128, 1), (1000, 12)])
def test_512_neuron_bf16_training_vs_fp32_path(B, T):
    """Ragged batches and several tiles: the bf16 tensor-core training step against the fp32 tile kernels of the same network."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda")
    vb.set_seed(64)
    net = LiquidNCPNetwork(8, 512, 1, device=dev)
    g = torch.Generator(device=dev).manual_seed(B + T)
    x = torch.randn(T, B, 8, device=dev, generator=g).permute(1, 0, 2)       # time-major storage
    h0 = 0.5 * torch.randn(B, 204, device=dev, generator=g)
    ry = torch.randn(T, B, 1, device=dev, generator=g).permute(1, 0, 2)
    rh = torch.randn(B, 204, device=dev, generator=g) / 8

    def run(bf16):
        net.zero_grad()
        xx = x.detach().clone().requires_grad_(True)
        hh = h0.detach().clone().requires_grad_(True)
        if bf16:
            y, hl = net.forward_seq_bf16(xx, hh)
        else:
            y, _, hl = net.forward_seq(xx, hh, return_last=True)
        torch.autograd.backward([y, hl], [ry, rh])
        return [y.detach(), hl.detach(), xx.grad, hh.grad] + [p.grad.clone() for p in net.parameters()]

    ref, got = run(False), run(True)
    names = ["y", "h_last", "dx", "dh0"] + [k for k, _ in net.named_parameters()]
    for n, a, b_ in zip(names, got, ref):
        e = rel_err(a.cpu().numpy(), b_.cpu().numpy())
        assert e < 2e-2, (n, e)
    # parameters only (x without gradient): same parameter gradients, dx is not computed
    net.zero_grad()
    y, hl = net.forward_seq_bf16(x, None)
    torch.autograd.backward([y, hl], [ry, rh])
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("B,T,F", [(1, 1, 1), (33, 5, 4), (1000, 37, 2), (257, 64, 8), (70, 33, 10), (40, 9, 20), (65, 3, 64)])
def test_row_transposition_kernel(B, T, F):
    """vb_transpose_rows: [B,T,F] <-> [T,B,F], ragged tiles on both axes."""
    from velora_b200 import _lib

    dev = torch.device("cuda")
    L = _lib.lib()
    src = torch.randn(B, T, F, device=dev)
    out = torch.full((T, B, F), float("nan"), device=dev)
    _lib.check(L.vb_transpose_rows(_lib.ptr(src), _lib.ptr(out), B, T, F, 1, _lib.stream_ptr(dev)), "vb_transpose_rows")
    assert torch.equal(out, src.permute(1, 0, 2))
    back = torch.full((B, T, F), float("nan"), device=dev)
    _lib.check(L.vb_transpose_rows(_lib.ptr(out), _lib.ptr(back), B, T, F, 0, _lib.stream_ptr(dev)), "vb_transpose_rows")
    assert torch.equal(back, src)


def test_large_batch_major_call_is_relaid_and_equals_time_major():
    """A batch-major call above functional.RELAYOUT_MIN_ROWS is transposed once and runs on the time-major fast path: results bit-equal to
    the same call on time-major storage, gradients returned for the caller's tensors."""
    import velora_b200 as vb
    from velora_b200 import functional as VF
    from velora_b200.models.lnn import LiquidNCPNetwork

    dev = torch.device("cuda")
    B, T = 4096, 8
    assert B * T >= VF.RELAYOUT_MIN_ROWS
    vb.set_seed(11)
    net = LiquidNCPNetwork(4, 20, 2, device=dev)
    xt = torch.randn(T, B, 4, device=dev)
    rt = torch.randn(T, B, 2, device=dev)
    rh = torch.randn(B, 8, device=dev)
    res = []
    for batch_major in (False, True):
        x = xt.permute(1, 0, 2)
        ry = rt.permute(1, 0, 2)
        if batch_major:
            x, ry = x.contiguous(), ry.contiguous()
        x = x.detach().requires_grad_(True)
        net.zero_grad()
        y, ht, hl = net.forward_seq(x, None, return_last=True)
        torch.autograd.backward([y, hl], [ry, rh])
        res.append((y.detach().clone(), ht.detach().clone(), x.grad.clone(), [p.grad.clone() for p in net.parameters()]))
    (y0, h0, dx0, g0), (y1, h1, dx1, g1) = res
    assert torch.equal(y0, y1) and torch.equal(h0, h1) and torch.equal(dx0, dx1)
    assert all(torch.equal(a, b) for a, b in zip(g0, g1))
