"""The standalone NCPLiquidCell (reference velora/models/lnn/cell.py:147-169) on the fused kernels: vb_cell_fwd / vb_cell_bwd through
functional.CellFn, against the golden written by the unmodified reference (tests/golden/cell.npz) and an fp64 restatement of the formula."""
import copy

import numpy as np
import pytest
import torch

from vb_testutil import FP32_RTOL, rel_err

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda")


def _golden_cell(d):
    from velora_b200.models.lnn.cell import NCPLiquidCell

    cell = NCPLiquidCell(8, 6, torch.tensor(d["mask"]), device=DEV).to(DEV)
    cell.load_state_dict({k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("sd_")})
    return cell


def test_cell_vs_reference_golden(golden):
    """y, h', dx, dh and the ten parameter gradients of the reference's own cell on the same inputs (loss = sum(y) + 2 sum(h'))."""
    d = golden("cell")
    cell = _golden_cell(d)
    x = torch.tensor(d["x"], device=DEV, requires_grad=True)
    h = torch.tensor(d["h"], device=DEV, requires_grad=True)
    y, hn = cell(x, h)
    assert np.allclose(y.detach().cpu().numpy(), d["y"], rtol=0, atol=2e-6)
    assert np.allclose(hn.detach().cpu().numpy(), d["hn"], rtol=0, atol=2e-6)
    (y.sum() + 2.0 * hn.sum()).backward()
    assert np.allclose(x.grad.cpu().numpy(), d["dx"], rtol=0, atol=5e-6)
    assert np.allclose(h.grad.cpu().numpy(), d["dh"], rtol=0, atol=5e-6)
    mask = cell.sparsity_mask.cpu().numpy()
    for k, p in cell.named_parameters():
        g = p.grad.cpu().numpy()
        assert np.allclose(g, d[f"grad_{k}"], rtol=0, atol=5e-6), k
        if k.endswith("weight"):
            assert np.all(g[mask == 0] == 0.0), k                                    # reference tests/models/test_lnn.py:570-576


def _f64_cell(cell, x, h, ry, rh):
    """fp64 restatement of cell.py:113-169 on the module's own parameters."""
    P = {k: v.detach().double().cpu() for k, v in cell.state_dict().items()}
    params = {k: P[k].clone().requires_grad_(True) for k in P if k.endswith("weight") or k.endswith("bias")}
    x = x.detach().double().cpu().requires_grad_(True)
    h = h.detach().double().cpu().requires_grad_(True)
    z = torch.cat([x, h], dim=1)

    def lin(name):
        return z @ (params[f"{name}.weight"] * P[f"{name}.mask"]).T + params[f"{name}.bias"]

    gate = torch.sigmoid(lin("f_head_to_g") + lin("f_head_to_h"))
    hn = torch.tanh(lin("g_head")) * (1.0 - gate) + gate * torch.tanh(lin("h_head"))
    y = lin("proj") + hn
    ((y * ry.double().cpu()).sum() + (hn * rh.double().cpu()).sum()).backward()
    return y.detach(), hn.detach(), x.grad, h.grad, {k: v.grad for k, v in params.items()}


@pytest.mark.parametrize("In,Nc,B", [(8, 6, 1), (12, 8, 37), (19, 13, 1000), (3, 40, 257), (64, 96, 129), (12, 8, 65536)])
def test_cell_shapes_vs_fp64(In, Nc, B):
    import velora_b200 as vb
    from velora_b200.models.lnn.cell import NCPLiquidCell

    vb.set_seed(In * 1000 + Nc)
    g = torch.Generator().manual_seed(7)
    mask = (torch.rand(In, Nc, generator=g) < 0.6).to(torch.int32) * (torch.randint(0, 2, (In, Nc), generator=g) * 2 - 1).to(torch.int32)
    cell = NCPLiquidCell(In, Nc, mask, device=DEV).to(DEV)
    x = torch.randn(B, In, generator=g).to(DEV).requires_grad_(True)
    h = (0.5 * torch.randn(B, Nc, generator=g)).to(DEV).requires_grad_(True)
    ry, rh = torch.randn(B, Nc, generator=g).to(DEV), torch.randn(B, Nc, generator=g).to(DEV)
    y, hn = cell(x, h)
    ((y * ry).sum() + (hn * rh).sum()).backward()
    ry64, rhn64, dx64, dh64, grads64 = _f64_cell(cell, x, h, ry, rh)
    tol = FP32_RTOL * (4 if B > 10000 else 1)          # parameter gradients are sums of B terms accumulated in fp32
    assert rel_err(y.detach().cpu().numpy(), ry64.numpy()) < FP32_RTOL
    assert rel_err(hn.detach().cpu().numpy(), rhn64.numpy()) < FP32_RTOL
    assert rel_err(x.grad.cpu().numpy(), dx64.numpy()) < FP32_RTOL
    assert rel_err(h.grad.cpu().numpy(), dh64.numpy()) < FP32_RTOL
    m = cell.sparsity_mask.cpu().numpy()
    for k, p in cell.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), grads64[k].numpy()) < tol, k
        if k.endswith("weight"):
            assert np.all(p.grad.cpu().numpy()[m == 0] == 0.0), k


def test_cell_equals_the_command_layer_of_the_network(golden):
    """LiquidNCPNetwork = motor(mish(cell(mish(inter(x)), h)))   (ncp.py:158-169): the standalone cell on the network's own command layer
    reproduces the fused network kernel's hidden state."""
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork
    import torch.nn.functional as F

    vb.set_seed(3)
    net = LiquidNCPNetwork(4, 20, 2, device=DEV)
    x = torch.randn(64, 4, device=DEV)
    h = 0.3 * torch.randn(64, int(net.command.n_hidden), device=DEV)
    with torch.no_grad():
        y_net, h_net = net(x, h)
        a = F.mish(net.inter(x))
        c, h_cell = net.command(a, h)
        y_cell = net.motor(F.mish(c))
    assert torch.allclose(h_cell, h_net, atol=2e-5, rtol=0)
    assert torch.allclose(y_cell, y_net, atol=2e-5, rtol=0)


def test_cell_conventions(golden):
    d = golden("cell")
    cell = _golden_cell(d)
    # CPU tensors are moved to the module's device, as the reference does (cell.py:160)
    y, hn = cell(torch.tensor(d["x"]), torch.tensor(d["h"]))
    assert y.is_cuda and hn.is_cuda and y.shape == (4, 6) and hn.shape == (4, 6)
    # empty batch
    y0, h0 = cell(torch.zeros(0, 8, device=DEV), torch.zeros(0, 6, device=DEV))
    assert y0.shape == (0, 6) and h0.shape == (0, 6)
    # deepcopy + torch.jit.script keep the fused path (velora/models/nf/modules.py:370-392 does this to whole networks)
    twin = torch.jit.script(copy.deepcopy(cell))
    y2, hn2 = twin(torch.tensor(d["x"], device=DEV), torch.tensor(d["h"], device=DEV))
    assert torch.equal(y2, y) and torch.equal(hn2, hn)
    # update_mask changes the cell's buffer only; the heads keep their own masks (reference behaviour)
    before = cell.g_head.mask.clone()
    cell.update_mask(torch.ones(8, 6))
    assert torch.equal(cell.g_head.mask, before)
    # gradient w.r.t. only one of the outputs
    x = torch.tensor(d["x"], device=DEV, requires_grad=True)
    _, hn3 = cell(x, torch.tensor(d["h"], device=DEV))
    hn3.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
