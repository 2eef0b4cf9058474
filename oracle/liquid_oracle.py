"""ORACLE (test infrastructure, never shipped or timed as the product): CPU restatement of the Liquid-NN hot path.

Two restatements of the same algorithm:

  * numpy, any float dtype, with a HAND-DERIVED backward (BPTT through the hidden state).  This is what the CUDA
    kernels are checked against: it does not depend on torch autograd, so it also checks the derivation in
    SURVEY.md section 8a ("Backward formulas the fused kernel must reproduce").
  * torch (CPU, fp32), op for op the ATen call sequence the reference issues (mask-multiply, addmm, cat, tanh,
    sigmoid, mish; autograd for the backward).  It is bit-identical to the reference modules on the same
    parameters, and is the timed `cpu_baseline` / `--impl reference` arm of bench.py (kind "port"): the reference
    itself is Python and cannot travel to the GPU box.

Reference lines restated:
  masked linear   y = x (W*M)^T + b                 velora/models/lnn/sparse.py:71-81
  cell mask       abs(cat([mask, ones]).T)          velora/models/lnn/cell.py:91-111
  cell            z = cat([x, h]); five heads; gate velora/models/lnn/cell.py:113-136, 147-169
  liquid network  inter -> Mish -> cell -> Mish -> motor; h=None => zeros; B==1 squeeze
                                                    velora/models/lnn/ncp.py:129-178
  NCP network     three masked linears, Mish between velora/models/lnn/ncp.py:250-274, 299-328
  Mish            x * tanh(softplus(x)) and its derivative (torch.nn.Mish; ATen, not vendored in the reference)

Parity pinned by tests/golden/{liquid,cell,ncp}.npz, generated from the unmodified reference by
oracle/make_golden.py (the reference's own tests hold no numeric vectors for this path: SURVEY.md 8c).

Parameter dictionaries use the reference's state_dict keys ("inter.weight", "command.g_head.mask", ...).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

HEADS = ("g_head", "h_head", "f_head_to_g", "f_head_to_h", "proj")


# ------------------------------------------------------------------------------------------------ elementwise
def mish(x: np.ndarray) -> np.ndarray:
    return x * np.tanh(np.logaddexp(0.0, x).astype(x.dtype))


def mish_grad(x: np.ndarray) -> np.ndarray:
    """d/dx [x tanh(softplus x)] = t + x * sigmoid(x) * (1 - t^2), t = tanh(softplus x)."""
    t = np.tanh(np.logaddexp(0.0, x).astype(x.dtype))
    sig = sigmoid(x)
    return t + x * sig * (1.0 - t * t)


def sigmoid(x: np.ndarray) -> np.ndarray:
    # stable in both tails
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e)).astype(x.dtype)


def masked_linear(x, W, M, b):
    """sparse.py:81  F.linear(x, weight * mask, bias)"""
    return x @ (W * M.astype(W.dtype)).T + b


# ------------------------------------------------------------------------------------------------ liquid network
def _cast(params: Dict[str, np.ndarray], dtype) -> Dict[str, np.ndarray]:
    return {k: np.asarray(v).astype(dtype) for k, v in params.items()}


def liquid_step(p: Dict[str, np.ndarray], x: np.ndarray, h: np.ndarray):
    """One call of LiquidNCPNetwork.forward (ncp.py:168-171) with its intermediates."""
    u = masked_linear(x, p["inter.weight"], p["inter.mask"], p["inter.bias"])
    a = mish(u)
    z = np.concatenate([a, h], axis=1)                                   # cell.py:162
    s = {k: masked_linear(z, p[f"command.{k}.weight"], p[f"command.{k}.mask"], p[f"command.{k}.bias"]) for k in HEADS}
    tg = np.tanh(s["g_head"])                                            # cell.py:127
    th = np.tanh(s["h_head"])                                            # cell.py:128
    gate = sigmoid(s["f_head_to_g"] + s["f_head_to_h"])                  # cell.py:130-133
    h_new = tg * (1.0 - gate) + gate * th                                # cell.py:134-136
    c = s["proj"] + h_new                                                # cell.py:168
    m = mish(c)
    y = masked_linear(m, p["motor.weight"], p["motor.mask"], p["motor.bias"])
    cache = dict(x=x, u=u, z=z, tg=tg, th=th, gate=gate, c=c, m=m)
    return y, h_new, cache


def liquid_seq_forward(params, x_seq: np.ndarray, h0: Optional[np.ndarray], dtype=np.float64):
    """forward_seq := the T-fold loop of forward() carrying h.  Returns y[B,T,O], h_traj[B,T,Nc], caches."""
    p = _cast(params, dtype)
    x_seq = np.asarray(x_seq).astype(dtype)
    B, T, _ = x_seq.shape
    nc = p["command.g_head.weight"].shape[0]
    h = np.zeros((B, nc), dtype=dtype) if h0 is None else np.asarray(h0).astype(dtype)   # ncp.py:162-166
    ys, hs, caches = [], [], []
    for t in range(T):
        y, h, cache = liquid_step(p, x_seq[:, t], h)
        ys.append(y)
        hs.append(h)
        caches.append(cache)
    return np.stack(ys, 1), np.stack(hs, 1), caches


def bf16_round(a: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round to nearest even) -> float32, the rounding `cvt.rn.bf16x2.f32` applies."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def liquid_seq_forward_bf16(params, x_seq: np.ndarray, h0: Optional[np.ndarray]):
    """The arithmetic contract of vb_liquid_bf16_fwd (csrc/liquid_big.cu): the cell's five heads see bf16-rounded operands
    (z = [mish(inter x) | h], the masked head weights with the two f heads summed in fp32 first, and their biases) and
    accumulate without further rounding; everything else - inter layer, activations, the hidden state, the motor layer - is
    float32 (evaluated here in float64).  Returns y[B,T,O], h_traj[B,T,Nc]."""
    p = _cast(params, np.float64)
    x_seq = np.asarray(x_seq, dtype=np.float64)
    B, T, _ = x_seq.shape
    nc = p["command.g_head.weight"].shape[0]
    Wm = lambda k: (p[f"command.{k}.weight"] * p[f"command.{k}.mask"]).astype(np.float32)   # noqa: E731
    wf = (Wm("f_head_to_g").astype(np.float32) + Wm("f_head_to_h").astype(np.float32)).astype(np.float32)
    bf = (p["command.f_head_to_g.bias"].astype(np.float32) + p["command.f_head_to_h.bias"].astype(np.float32)).astype(np.float32)
    W = {"g": bf16_round(Wm("g_head")), "h": bf16_round(Wm("h_head")), "f": bf16_round(wf), "p": bf16_round(Wm("proj"))}
    b = {"g": bf16_round(p["command.g_head.bias"]), "h": bf16_round(p["command.h_head.bias"]), "f": bf16_round(bf),
         "p": bf16_round(p["command.proj.bias"])}
    h = np.zeros((B, nc)) if h0 is None else np.asarray(h0, dtype=np.float64)
    ys, hs = [], []
    for t in range(T):
        a = mish(masked_linear(x_seq[:, t], p["inter.weight"], p["inter.mask"], p["inter.bias"]))
        z = bf16_round(np.concatenate([a, h], axis=1)).astype(np.float64)
        s = {k: z @ W[k].astype(np.float64).T + b[k].astype(np.float64) for k in W}
        tg, th, gate = np.tanh(s["g"]), np.tanh(s["h"]), sigmoid(s["f"])
        h = tg * (1.0 - gate) + gate * th
        ys.append(masked_linear(mish(s["p"] + h), p["motor.weight"], p["motor.mask"], p["motor.bias"]))
        hs.append(h)
    return np.stack(ys, 1), np.stack(hs, 1)


def liquid_seq_backward(params, caches, dy: np.ndarray, dh_traj: np.ndarray, dtype=np.float64):
    """Hand-derived BPTT.  dy[B,T,O] and dh_traj[B,T,Nc] are the loss gradients w.r.t. the two outputs.

    Returns dx[B,T,I], dh0[B,Nc], grads{state_dict key -> array} (masked entries exactly 0).
    """
    p = _cast(params, dtype)
    dy = np.asarray(dy).astype(dtype)
    dh_traj = np.asarray(dh_traj).astype(dtype)
    B, T, _ = dy.shape
    ni = p["inter.weight"].shape[0]
    grads = {k: np.zeros_like(v) for k, v in p.items() if k.endswith("weight") or k.endswith("bias")}
    Wm = lambda name: p[f"{name}.weight"] * p[f"{name}.mask"]   # noqa: E731
    dxs = [None] * T
    dh_carry = np.zeros_like(dh_traj[:, 0])
    for t in reversed(range(T)):
        c = caches[t]
        dyt = dy[:, t]
        dm = dyt @ Wm("motor")
        grads["motor.weight"] += (dyt.T @ c["m"]) * p["motor.mask"]
        grads["motor.bias"] += dyt.sum(0)
        dc = dm * mish_grad(c["c"])
        dhn = dc + dh_traj[:, t] + dh_carry
        tg, th, gate = c["tg"], c["th"], c["gate"]
        ds = {
            "g_head": dhn * (1.0 - gate) * (1.0 - tg * tg),
            "h_head": dhn * gate * (1.0 - th * th),
            "proj": dc,
        }
        dsf = dhn * (th - tg) * gate * (1.0 - gate)
        ds["f_head_to_g"] = dsf
        ds["f_head_to_h"] = dsf
        dz = np.zeros_like(c["z"])
        for k in HEADS:
            name = f"command.{k}"
            dz += ds[k] @ Wm(name)
            grads[f"{name}.weight"] += (ds[k].T @ c["z"]) * p[f"{name}.mask"]
            grads[f"{name}.bias"] += ds[k].sum(0)
        da, dh_carry = dz[:, :ni], dz[:, ni:]
        du = da * mish_grad(c["u"])
        dxs[t] = du @ Wm("inter")
        grads["inter.weight"] += (du.T @ c["x"]) * p["inter.mask"]
        grads["inter.bias"] += du.sum(0)
    return np.stack(dxs, 1), dh_carry, grads


# ------------------------------------------------------------------------------------------------ NCP network
def ncp_forward(params, x: np.ndarray, dtype=np.float64):
    """NCPNetwork.ncp: SparseLinear -> Mish -> SparseLinear -> Mish -> SparseLinear (ncp.py:250-274)."""
    p = _cast(params, dtype)
    x = np.asarray(x).astype(dtype)
    u1 = masked_linear(x, p["ncp.0.weight"], p["ncp.0.mask"], p["ncp.0.bias"])
    a1 = mish(u1)
    u2 = masked_linear(a1, p["ncp.2.weight"], p["ncp.2.mask"], p["ncp.2.bias"])
    a2 = mish(u2)
    y = masked_linear(a2, p["ncp.4.weight"], p["ncp.4.mask"], p["ncp.4.bias"])
    return y, dict(x=x, u1=u1, a1=a1, u2=u2, a2=a2)


def ncp_backward(params, cache, dy: np.ndarray, dtype=np.float64):
    p = _cast(params, dtype)
    dy = np.asarray(dy).astype(dtype)
    Wm = lambda i: p[f"ncp.{i}.weight"] * p[f"ncp.{i}.mask"]   # noqa: E731
    g = {}
    g["ncp.4.weight"] = (dy.T @ cache["a2"]) * p["ncp.4.mask"]
    g["ncp.4.bias"] = dy.sum(0)
    du2 = (dy @ Wm(4)) * mish_grad(cache["u2"])
    g["ncp.2.weight"] = (du2.T @ cache["a1"]) * p["ncp.2.mask"]
    g["ncp.2.bias"] = du2.sum(0)
    du1 = (du2 @ Wm(2)) * mish_grad(cache["u1"])
    g["ncp.0.weight"] = (du1.T @ cache["x"]) * p["ncp.0.mask"]
    g["ncp.0.bias"] = du1.sum(0)
    dx = du1 @ Wm(0)
    return dx, g


# ------------------------------------------------------------------------------------------------ torch CPU port
def torch_params(params, dtype=None, requires_grad: bool = False):
    import torch

    out = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v.detach().clone()
        if k.endswith("weight") or k.endswith("bias"):
            if dtype is not None:
                t = t.to(dtype)
            t = t.clone().requires_grad_(requires_grad)
        out[k] = t
    return out


def liquid_forward_torch(p, x, h):
    """Op-for-op the reference's ATen sequence for one LiquidNCPNetwork.forward call (fp32 CPU)."""
    import torch
    import torch.nn.functional as F

    def lin(name, v):
        return F.linear(v, p[f"{name}.weight"] * p[f"{name}.mask"], p[f"{name}.bias"])   # sparse.py:81

    a = F.mish(lin("inter", x))
    z = torch.cat([a, h], dim=1)
    g_out = lin("command.g_head", z)
    h_out = lin("command.h_head", z)
    g_head = torch.tanh(g_out)
    h_head = torch.tanh(h_out)
    fh_g = lin("command.f_head_to_g", z)
    fh_h = lin("command.f_head_to_h", z)
    gate_out = torch.sigmoid(fh_g + fh_h)
    f_head = 1.0 - gate_out
    new_hidden = g_head * f_head + gate_out * h_head
    c = lin("command.proj", z) + new_hidden
    y = lin("motor", F.mish(c))
    return y, new_hidden


def liquid_seq_forward_torch(p, x_seq, h0):
    import torch

    B, T, _ = x_seq.shape
    nc = p["command.g_head.weight"].shape[0]
    h = torch.zeros(B, nc, dtype=x_seq.dtype) if h0 is None else h0
    ys, hs = [], []
    for t in range(T):
        y, h = liquid_forward_torch(p, x_seq[:, t], h)
        ys.append(y)
        hs.append(h)
    return torch.stack(ys, 1), torch.stack(hs, 1)


def ncp_forward_torch(p, x):
    import torch.nn.functional as F

    def lin(i, v):
        return F.linear(v, p[f"ncp.{i}.weight"] * p[f"ncp.{i}.mask"], p[f"ncp.{i}.bias"])

    return lin(4, F.mish(lin(2, F.mish(lin(0, x)))))


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """||a - b|| / ||b|| (the tolerance measure of SURVEY.md 8c); 0/0 := 0."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    na = np.linalg.norm(a - b)
    if nb == 0.0:
        return 0.0 if na == 0.0 else float("inf")
    return float(na / nb)
