"""Parity of the CUDA NCPNetwork (critic) path with the reference goldens."""
import copy

import numpy as np
import pytest
import torch

from oracle import liquid_oracle as lo
from vb_testutil import FP32_RTOL, assert_close_to_reference, liquid_case_params, rel_err

pytestmark = pytest.mark.gpu


def _build(d, name):
    import velora_b200 as vb
    from velora_b200.models.lnn import NCPNetwork

    i, n, o, sp, seed, B = d[f"{name}_meta"]
    vb.set_seed(int(seed))
    net = NCPNetwork(int(i), int(n), int(o), sparsity_level=float(sp), device=torch.device("cuda"))
    net.load_state_dict({k: torch.tensor(v) for k, v in liquid_case_params(d, name).items()})
    return net


@pytest.mark.parametrize("name", ["critic_ct", "critic_disc", "small", "odd"])
def test_ncp_vs_reference(golden, name):
    d = golden("ncp")
    net = _build(d, name)
    x = torch.tensor(d[f"{name}_x"], device="cuda", requires_grad=True)
    r = torch.tensor(d[f"{name}_r"], device="cuda")
    y = net(x)
    assert list(y.shape) == list(d[f"{name}_api_shape"])          # B == 1 squeeze (ncp.py:325-326)
    y2 = y.reshape(r.shape)
    assert_close_to_reference(y2.detach().cpu().numpy(), d[f"{name}_f32_y"], d[f"{name}_f64_y"], "y")
    (y2 * r).sum().backward()
    assert_close_to_reference(x.grad.cpu().numpy(), d[f"{name}_f32_dx"], d[f"{name}_f64_dx"], "dx")
    for k, p in net.named_parameters():
        g = p.grad.cpu().numpy()
        assert_close_to_reference(g, d[f"{name}_f32_grad_{k}"], d[f"{name}_f64_grad_{k}"], f"grad {k}")
        if k.endswith("weight"):
            mask = net.state_dict()[k.replace("weight", "mask")].cpu().numpy()
            assert np.all(g[mask == 0] == 0.0)
    with pytest.raises(ValueError):
        net(torch.zeros(2, 2, 2))


@pytest.mark.parametrize("B", [1, 33, 1000, 10001, 20003, 65536])   # 10001 / 20003: the 64- and 128-row register-blocked tiles, ragged
def test_ncp_ragged_vs_oracle(golden, B):
    d = golden("ncp")
    net = _build(d, "critic_ct")
    p = liquid_case_params(d, "critic_ct")
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 5, generator=g)
    r = torch.randn(B, 1, generator=g)
    yo, cache = lo.ncp_forward(p, x.numpy(), np.float64)
    dxo, go = lo.ncp_backward(p, cache, r.numpy(), np.float64)
    xc = x.cuda().requires_grad_(True)
    y = net(xc).reshape(B, 1)
    (y * r.cuda()).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo) < FP32_RTOL
    assert rel_err(xc.grad.cpu().numpy(), dxo) < FP32_RTOL
    for k, pr in net.named_parameters():
        assert rel_err(pr.grad.cpu().numpy(), go[k]) < FP32_RTOL, k


def test_ncp_script_and_determinism(golden):
    d = golden("ncp")
    net = _build(d, "critic_ct")
    x = torch.tensor(d["critic_ct_x"], device="cuda")
    twin = torch.jit.script(copy.deepcopy(net))
    assert torch.equal(twin(x), net(x))
    outs = []
    for _ in range(2):
        net.zero_grad()
        net(x).sum().backward()
        outs.append(torch.cat([p.grad.flatten() for p in net.parameters()]))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("name,B", [("critic_disc", 4097), ("critic_disc", 30011), ("critic_ct", 200000)])
def test_ncp_tensor_core_tier_vs_oracle_and_tile_kernels(golden, name, B):
    """Large batches run the tcgen05 critic kernels (bf16x3): same results as the fp64 oracle to 1e-5, agreement with the FFMA
    tile kernels (kernel_flags = FORCE_GENERIC) to the same bound (db3 is a cancelling sum of B random terms: its fp32
    summation order alone moves it by a few 1e-6), masked gradients exactly zero, bit-reproducible."""
    d = golden("ncp")
    net = _build(d, name)
    p = liquid_case_params(d, name)
    I, O = net.in_features, net.out_features
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, I, generator=g)
    r = torch.randn(B, O, generator=g)
    yo, cache = lo.ncp_forward(p, x.numpy(), np.float64)
    dxo, go = lo.ncp_backward(p, cache, r.numpy(), np.float64)
    res = {}
    for flags in (0, 1, 0):
        net.kernel_flags = flags
        net.zero_grad()
        xc = x.cuda().requires_grad_(True)
        y = net(xc).reshape(B, O)
        (y * r.cuda()).sum().backward()
        out = {"y": y.detach().cpu().numpy(), "dx": xc.grad.cpu().numpy()}
        out.update({k: pr.grad.cpu().numpy().copy() for k, pr in net.named_parameters()})
        res.setdefault(flags, []).append(out)
    tc, tile, tc2 = res[0][0], res[1][0], res[0][1]
    assert rel_err(tc["y"], yo) < FP32_RTOL and rel_err(tc["dx"], dxo) < FP32_RTOL
    for k in go:
        assert rel_err(tc[k], go[k]) < FP32_RTOL, k
        assert rel_err(tc[k], tile[k]) < FP32_RTOL, k
        assert np.array_equal(tc[k], tc2[k]), f"{k}: not reproducible"
        if k.endswith("weight"):
            mask = net.state_dict()[k.replace("weight", "mask")].cpu().numpy()
            assert np.all(tc[k][mask == 0] == 0.0), k
    assert rel_err(tc["y"], tile["y"]) < 2e-6 and rel_err(tc["dx"], tile["dx"]) < 2e-6


def test_ncp_empty_batch(golden):
    d = golden("ncp")
    net = _build(d, "critic_ct")
    x = torch.zeros(0, 5, device="cuda", requires_grad=True)
    y = net(x)
    assert y.shape == (0, 1)
    y.sum().backward()
    assert x.grad.shape == (0, 5)
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.count_nonzero(p.grad) == 0, k


def test_frozen_parameters_backward_returns_dx_only(golden):
    """Inside functional.frozen_parameters() a critic's parameters are constants: same output, same dL/dx, no parameter gradients
    (vb_ncp_bwd with dpacked = NULL) - the critics inside the SAC actor loss (agent.py:236-245)."""
    import velora_b200 as vb
    from velora_b200 import functional as VF
    from velora_b200.models.lnn import NCPNetwork

    dev = torch.device("cuda")
    for B in (64, 5000):                 # tile kernels / tensor-core kernels
        vb.set_seed(5)
        net = NCPNetwork(5, 128, 1, device=dev)
        x1 = torch.randn(B, 5, device=dev, requires_grad=True)
        x2 = x1.detach().clone().requires_grad_(True)
        r = torch.randn(B, device=dev)
        y1 = net(x1)
        (y1.reshape(-1) * r).sum().backward()
        ref_grads = [p.grad.clone() for p in net.parameters()]
        net.zero_grad()
        with VF.frozen_parameters():
            y2 = net(x2)
        (y2.reshape(-1) * r).sum().backward()
        assert torch.equal(y1, y2)
        assert torch.equal(x1.grad, x2.grad)
        assert all(p.grad is None or not p.grad.any() for p in net.parameters())
        assert all(g.abs().sum() > 0 for g in ref_grads)
        assert not VF.parameters_frozen()
