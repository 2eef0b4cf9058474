"""torch.autograd.Function wrappers over the C ABI (the "thin C-ABI torch.autograd.Function extension").

LiquidSeqFn   T-fold LiquidNCPNetwork.forward + its backward    (reference: velora/models/lnn/ncp.py:129-178)
NcpFn         NCPNetwork.forward + its backward                 (reference: velora/models/lnn/ncp.py:299-328)
gather_rows   ReplayBuffer.sample's six gathers                 (reference: velora/buffer/replay.py:81-88)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import threading

import torch

from . import _lib
from . import dataparallel as _dp

HEAD_NAMES = ("g_head", "h_head", "f_head_to_g", "f_head_to_h", "proj")

# Optional instrumentation used by bench.py: when PROFILE is a dict, every C-ABI compute call is bracketed by CUDA
# events on the launching stream (PROFILE[name] -> list of (start, end)); LAUNCHES counts the kernels this library
# launches (pack 1, fwd 1, bwd 2 = kernel + partial-sum reduction, unpack 1, gather 1).
PROFILE = None
LAUNCHES = 0


_WORKSPACES = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch buffer for the backward kernels, reused per (device, stream): work on one stream is ordered, so the
    next backward on that stream may overwrite it."""
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the buffer must come from (and live with) the graph's own memory pool: caching it under the
        # capture stream's handle could hand it to eager work on a later stream that happens to reuse the handle
        return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


class _timed:
    def __init__(self, name: str, launches: int):
        self.name, self.launches = name, launches

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.launches
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t if t.is_contiguous() else t.contiguous()


# vb_liquid_fwd / vb_liquid_bwd layout flags (include/velora_b200.h)
X_TM, Y_TM, H_TM, DX_TM, DH_TM = 16, 32, 64, 128, 256


def _seq(t: Optional[torch.Tensor], tm_ok: bool) -> Tuple[Optional[torch.Tensor], bool]:
    """fp32 [B,T,F] tensor -> (tensor the kernels can read in place, stored time-major?).  Batch-major contiguous and (where
    the kernel path accepts it) time-major storage - a permuted view of a contiguous [T,B,F] tensor - pass through."""
    if t is None:
        return None, False
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.is_contiguous():
        return t, False
    if tm_ok and t.permute(1, 0, 2).is_contiguous():
        return t, True
    return t.contiguous(), False


RELAYOUT_MIN_ROWS = 16384     # B * T from which a batch-major sequence tensor is re-laid time-major for the tensor-core kernels


def _time_major_copy(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 [B,T,F] -> the same values in time-major storage ([T,B,F] viewed as [B,T,F]): one tiled transposition
    (vb_transpose_rows).  The liquid tensor-core kernels read one row per thread; time-major rows of a warp are contiguous."""
    B, T, F = t.shape
    out = torch.empty((T, B, F), dtype=torch.float32, device=t.device)
    L = _lib.lib()
    with torch.cuda.device(t.device):
        _count(1)
        _lib.check(L.vb_transpose_rows(_lib.ptr(t), _lib.ptr(out), B, T, F, 1, _lib.stream_ptr(t.device)), "vb_transpose_rows")
    return out.permute(1, 0, 2)


def _relayout(t: Optional[torch.Tensor], is_tm: bool, tm_ok: bool) -> Tuple[Optional[torch.Tensor], bool]:
    if t is None or is_tm or not tm_ok or not t.is_contiguous() or t.shape[0] * t.shape[1] < RELAYOUT_MIN_ROWS or t.shape[2] > 64:
        return t, is_tm
    return _time_major_copy(t), True


class LiquidSeqFn(torch.autograd.Function):
    """y[B,T,O], h_traj[B,T,Nc], h_last[B,Nc] = liquid(x[B,T,I], h0[B,Nc] | None; 7 masked layers)."""

    @staticmethod
    def forward(ctx, x, h0, dims, flags, masks, *params):
        # params: w_in, b_in, (w,b) x5 heads, w_out, b_out ; masks: m_in, m_head x5, m_out (not differentiable)
        I, Ni, Nc, O = dims
        L = _lib.lib()
        _lib.require_cuda(x, "LiquidNCPNetwork.forward")
        dev = x.device
        B, T, _ = x.shape
        # time-major storage ([T,B,F], the stack of per-step batches) gives the one-thread-per-sample kernels contiguous
        # warp accesses; the outputs are allocated that way and returned as [B,T,F] views
        tm_ok = T > 1 and bool(L.vb_liquid_time_major_ok(I, Ni, Nc, O, flags))
        x, x_tm = _seq(x, tm_ok)
        x, x_tm = _relayout(x, x_tm, tm_ok)          # batch-major input of a large call: one transposition, then the fast layout throughout
        h0c = _f32c(h0)
        params = [_f32c(p.detach()) for p in params]
        w_in, b_in = params[0], params[1]
        w_h = [params[2 + 2 * k] for k in range(5)]
        b_h = [params[3 + 2 * k] for k in range(5)]
        w_out, b_out = params[12], params[13]
        # wiring masks are stored as transposed views (abs(mask.T) keeps the strides): passed as they are, with a type code
        masks = tuple(_lib.mask_arg(m, dev) for m in masks)
        heads_type = _lib.common_mask_type(masks[1:6])
        if heads_type is None:
            masks = (masks[0], *[m.contiguous() for m in masks[1:6]], masks[6])
            heads_type = _lib.mask_type(masks[1])
        m_in, m_h, m_out = masks[0], list(masks[1:6]), masks[6]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            packed = torch.empty(L.vb_liquid_packed_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev)
            _count(1)
            _lib.check(
                L.vb_liquid_pack(
                    _lib.ptr(w_in), _lib.ptr(b_in), _lib.ptr(m_in), _lib.mask_type(m_in),
                    _lib.ptr_array(w_h), _lib.ptr_array(b_h), _lib.ptr_array(m_h), heads_type,
                    _lib.ptr(w_out), _lib.ptr(b_out), _lib.ptr(m_out), _lib.mask_type(m_out),
                    I, Ni, Nc, O, _lib.ptr(packed), st,
                ),
                "vb_liquid_pack",
            )
            if tm_ok:
                y = torch.empty((T, B, O), dtype=torch.float32, device=dev).permute(1, 0, 2)
                h_traj = torch.empty((T, B, Nc), dtype=torch.float32, device=dev).permute(1, 0, 2)
            else:
                y = torch.empty((B, T, O), dtype=torch.float32, device=dev)
                h_traj = torch.empty((B, T, Nc), dtype=torch.float32, device=dev)
            lay = (X_TM if x_tm else 0) | ((Y_TM | H_TM) if tm_ok else 0)
            # the gate factors the backward would otherwise recompute (tensor-core path: 160 B per sample-step); only when a
            # gradient can be asked for
            saved = None
            if any(ctx.needs_input_grad):
                n_saved = L.vb_liquid_saved_bytes(I, Ni, Nc, O, B, T, flags)
                if n_saved > 0:
                    saved = torch.empty(n_saved, dtype=torch.uint8, device=dev)
            with _timed("liquid_fwd", 1):
                _lib.check(
                    L.vb_liquid_fwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(h0c), _lib.ptr(y), _lib.ptr(h_traj),
                                    B, T, I, Ni, Nc, O, _lib.ptr(saved), 0 if saved is None else saved.numel(), flags | lay, st),
                    "vb_liquid_fwd",
                )
        h_last = h_traj[:, T - 1].clone()
        ctx.dims, ctx.flags, ctx.masks = dims, flags, masks
        ctx.heads_type = heads_type
        ctx.has_h0 = h0 is not None
        ctx.layout = (tm_ok, x_tm)
        ctx.has_saved = saved is not None
        ctx.save_for_backward(x, h0c if h0c is not None else x.new_empty(0), h_traj, packed,
                              saved if saved is not None else x.new_empty(0))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable()
        return y, h_traj, h_last

    @staticmethod
    def backward(ctx, dy, dh_traj, dh_last):
        x, h0, h_traj, packed, saved = ctx.saved_tensors
        I, Ni, Nc, O = ctx.dims
        masks = ctx.masks
        L = _lib.lib()
        dev = x.device
        B, T, _ = x.shape
        h0p = h0 if ctx.has_h0 else None
        saved = saved if ctx.has_saved else None
        tm_ok, x_tm = ctx.layout
        dy, dy_tm = _seq(dy, tm_ok) if dy is not None else (torch.zeros((B, T, O), dtype=torch.float32, device=dev), False)
        dh_traj, dh_tm = _seq(dh_traj, tm_ok)
        dy, dy_tm = _relayout(dy, dy_tm, tm_ok)
        dh_traj, dh_tm = _relayout(dh_traj, dh_tm, tm_ok)
        dh_last = _f32c(dh_last)
        lay = (X_TM | DX_TM if x_tm else 0) | (H_TM if tm_ok else 0) | (Y_TM if dy_tm else 0) | (DH_TM if dh_tm else 0)
        need_x = ctx.needs_input_grad[0]
        need_h0 = ctx.has_h0 and ctx.needs_input_grad[1]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            dx = torch.empty_like(x) if need_x else None
            dh0 = torch.empty((B, Nc), dtype=torch.float32, device=dev) if need_h0 else None
            n_grad = L.vb_liquid_grad_floats(I, Ni, Nc, O)
            dpacked = torch.empty(n_grad, dtype=torch.float32, device=dev)
            ws = _workspace(dev, L.vb_liquid_bwd_workspace_bytes(I, Ni, Nc, O, B, T, 1 if saved is not None or ctx.flags & 3 else 0))
            with _timed("liquid_bwd", 2):
                _lib.check(
                    L.vb_liquid_bwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(h0p), _lib.ptr(h_traj), _lib.ptr(dy),
                                    _lib.ptr(dh_traj), _lib.ptr(dh_last), _lib.ptr(saved), 0 if saved is None else saved.numel(),
                                    _lib.ptr(dx), _lib.ptr(dh0), _lib.ptr(dpacked),
                                    B, T, I, Ni, Nc, O, _lib.ptr(ws), ws.numel(), ctx.flags | lay, st),
                    "vb_liquid_bwd",
                )
            # data-parallel: ONE collective over the flat pre-masked gradient block
            _dp.maybe_allreduce(dpacked)
            need = ctx.needs_input_grad[5:]
            shapes = [(Ni, I), (Ni,)] + [(Nc, Ni + Nc), (Nc,)] * 5 + [(O, Nc), (O,)]
            grads: List[Optional[torch.Tensor]] = [
                torch.empty(s, dtype=torch.float32, device=dev) if need[k] else None for k, s in enumerate(shapes)
            ]
            m_in, m_h, m_out = masks[0], list(masks[1:6]), masks[6]
            _count(1)
            _lib.check(
                L.vb_liquid_unpack_grads(
                    _lib.ptr(dpacked), _lib.ptr(m_in), _lib.mask_type(m_in), _lib.ptr_array(m_h), ctx.heads_type,
                    _lib.ptr(m_out), _lib.mask_type(m_out),
                    _lib.ptr(grads[0]), _lib.ptr(grads[1]),
                    _lib.ptr_array([grads[2 + 2 * k] for k in range(5)]), _lib.ptr_array([grads[3 + 2 * k] for k in range(5)]),
                    _lib.ptr(grads[12]), _lib.ptr(grads[13]), I, Ni, Nc, O, 0, st,
                ),
                "vb_liquid_unpack_grads",
            )
        return (dx, dh0, None, None, None, *grads)


class CellFn(torch.autograd.Function):
    """y[B,Nc], h_new[B,Nc] = NCPLiquidCell(x[B,In], h[B,Nc]; five masked heads) - one kernel forward, one backward (cell.py:147-169)."""

    @staticmethod
    def forward(ctx, x, h, dims, masks, *params):
        # params: (w, b) x5 in head order g, h, f->g, f->h, proj; masks: the five head masks (the cell's sparsity mask, not differentiable)
        In, Nc = dims
        L = _lib.lib()
        _lib.require_cuda(x, "NCPLiquidCell.forward")
        dev = x.device
        x, h = _f32c(x), _f32c(h)
        B = x.shape[0]
        params = [_f32c(p.detach()) for p in params]
        w_h = [params[2 * k] for k in range(5)]
        b_h = [params[2 * k + 1] for k in range(5)]
        masks = tuple(_lib.mask_arg(m, dev) for m in masks)
        heads_type = _lib.common_mask_type(masks)
        if heads_type is None:
            masks = tuple(m.contiguous() for m in masks)
            heads_type = _lib.mask_type(masks[0])
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            packed = torch.empty(L.vb_liquid_packed_floats(In, In, Nc, Nc), dtype=torch.float32, device=dev)
            _count(1)
            _lib.check(
                L.vb_liquid_pack(None, None, None, 0, _lib.ptr_array(w_h), _lib.ptr_array(b_h), _lib.ptr_array(list(masks)), heads_type,
                                 None, None, None, 0, In, In, Nc, Nc, _lib.ptr(packed), st),
                "vb_liquid_pack",
            )
            y = torch.empty((B, Nc), dtype=torch.float32, device=dev)
            h_new = torch.empty((B, Nc), dtype=torch.float32, device=dev)
            with _timed("cell_fwd", 1):
                _lib.check(L.vb_cell_fwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(h), _lib.ptr(y), _lib.ptr(h_new), B, In, Nc, st),
                           "vb_cell_fwd")
        ctx.dims, ctx.masks, ctx.heads_type = dims, masks, heads_type
        ctx.save_for_backward(x, h, packed)
        ctx.set_materialize_grads(False)
        return y, h_new

    @staticmethod
    def backward(ctx, dy, dh_new):
        x, h, packed = ctx.saved_tensors
        In, Nc = ctx.dims
        L = _lib.lib()
        dev = x.device
        B = x.shape[0]
        dy = _f32c(dy) if dy is not None else torch.zeros((B, Nc), dtype=torch.float32, device=dev)
        dh_new = _f32c(dh_new)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            dh = torch.empty_like(h) if ctx.needs_input_grad[1] else None
            dpacked = torch.empty(L.vb_liquid_grad_floats(In, In, Nc, Nc), dtype=torch.float32, device=dev)
            ws = _workspace(dev, L.vb_cell_bwd_workspace_bytes(In, Nc))
            with _timed("cell_bwd", 2):
                _lib.check(
                    L.vb_cell_bwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(h), _lib.ptr(dy), _lib.ptr(dh_new), _lib.ptr(dx), _lib.ptr(dh),
                                  _lib.ptr(dpacked), B, In, Nc, _lib.ptr(ws), ws.numel(), st),
                    "vb_cell_bwd",
                )
            _dp.maybe_allreduce(dpacked)
            need = ctx.needs_input_grad[4:]
            shapes = [(Nc, In + Nc), (Nc,)] * 5
            grads: List[Optional[torch.Tensor]] = [
                torch.empty(s, dtype=torch.float32, device=dev) if need[k] else None for k, s in enumerate(shapes)
            ]
            _count(1)
            _lib.check(
                L.vb_liquid_unpack_grads(
                    _lib.ptr(dpacked), None, 0, _lib.ptr_array(list(ctx.masks)), ctx.heads_type, None, 0,
                    None, None, _lib.ptr_array([grads[2 * k] for k in range(5)]), _lib.ptr_array([grads[2 * k + 1] for k in range(5)]),
                    None, None, In, In, Nc, Nc, 0, st,
                ),
                "vb_liquid_unpack_grads",
            )
        return (dx, dh, None, None, *grads)


# Networks called inside `frozen_parameters()` treat their parameters as constants: the backward returns dL/dx only.  The SAC actor
# update differentiates the actor loss through both critics (agent.py:236-245); the reference lets autograd fill the critics' .grad there
# and throws it away at the next critic zero_grad (modules.py:409-415) - here that work (two weight-gradient passes, two unpacks, twelve
# accumulations at batch 64) is not done at all.
_FROZEN = threading.local()


class frozen_parameters:
    def __enter__(self):
        self.prev = getattr(_FROZEN, "on", False)
        _FROZEN.on = True
        return self

    def __exit__(self, *exc):
        _FROZEN.on = self.prev
        return False


def parameters_frozen() -> bool:
    return getattr(_FROZEN, "on", False)


class NcpFn(torch.autograd.Function):
    """y[B,O] = ncp(x[B,I]; three masked layers)."""

    @staticmethod
    def forward(ctx, x, dims, flags, masks, *params):
        I, Ni, Nc, O = dims
        L = _lib.lib()
        _lib.require_cuda(x, "NCPNetwork.forward")
        dev = x.device
        x = _f32c(x)
        B = x.shape[0]
        params = [_f32c(p.detach()) for p in params]
        w = [params[0], params[2], params[4]]
        b = [params[1], params[3], params[5]]
        masks = tuple(_lib.mask_arg(m, dev) for m in masks)
        m_type = _lib.common_mask_type(masks)
        if m_type is None:
            masks = tuple(m.contiguous() for m in masks)
            m_type = _lib.mask_type(masks[0])
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            packed = torch.empty(L.vb_ncp_packed_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev)
            _count(1)
            _lib.check(
                L.vb_ncp_pack(_lib.ptr_array(w), _lib.ptr_array(b), _lib.ptr_array(list(masks)), m_type,
                              I, Ni, Nc, O, _lib.ptr(packed), st),
                "vb_ncp_pack",
            )
            y = torch.empty((B, O), dtype=torch.float32, device=dev)
            with _timed("ncp_fwd", 1):
                _lib.check(L.vb_ncp_fwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(y), B, I, Ni, Nc, O, flags, st), "vb_ncp_fwd")
        ctx.dims, ctx.flags, ctx.masks = dims, flags, masks
        ctx.m_type = m_type
        ctx.save_for_backward(x, packed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, packed = ctx.saved_tensors
        I, Ni, Nc, O = ctx.dims
        masks = ctx.masks
        L = _lib.lib()
        dev = x.device
        B = x.shape[0]
        dy = _f32c(dy)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            need = ctx.needs_input_grad[4:]
            want_params = any(need)
            if not want_params and dx is None:
                return (None,) * 10
            dpacked = torch.empty(L.vb_ncp_grad_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev) if want_params else None
            ws = _workspace(dev, L.vb_ncp_bwd_workspace_bytes(I, Ni, Nc, O, B))
            with _timed("ncp_bwd", 2):
                _lib.check(
                    L.vb_ncp_bwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(dpacked),
                                 B, I, Ni, Nc, O, _lib.ptr(ws), ws.numel(), ctx.flags, st),
                    "vb_ncp_bwd",
                )
            if not want_params:       # parameters frozen for this call (frozen_parameters()): dL/dx only
                return (dx, None, None, None) + (None,) * 6
            _dp.maybe_allreduce(dpacked)
            shapes = [(Ni, I), (Ni,), (Nc, Ni), (Nc,), (O, Nc), (O,)]
            grads = [torch.empty(s, dtype=torch.float32, device=dev) if need[k] else None for k, s in enumerate(shapes)]
            _count(1)
            _lib.check(
                L.vb_ncp_unpack_grads(_lib.ptr(dpacked), _lib.ptr_array(list(masks)), ctx.m_type,
                                      _lib.ptr_array([grads[0], grads[2], grads[4]]), _lib.ptr_array([grads[1], grads[3], grads[5]]),
                                      I, Ni, Nc, O, 0, st),
                "vb_ncp_unpack_grads",
            )
        return (dx, None, None, None, *grads)


def gather_rows(arrays: Sequence[torch.Tensor], indices: torch.Tensor, n_rows: int) -> Tuple[torch.Tensor, ...]:
    """out[k] = arrays[k][indices] for 2-D float32 CUDA arrays sharing dim 0, in one launch."""
    L = _lib.lib()
    dev = arrays[0].device
    _lib.require_cuda(arrays[0], "ReplayBuffer.sample")
    if indices.dtype != torch.int64 or indices.device != dev or not indices.is_contiguous():
        indices = indices.to(device=dev, dtype=torch.int64).contiguous()
    B = indices.numel()
    outs, widths = [], []
    for a in arrays:
        if a.dtype != torch.float32 or not a.is_contiguous() or a.dim() != 2 or a.device != dev:
            raise ValueError("gather_rows: arrays must be contiguous 2-D float32 tensors on one CUDA device")
        outs.append(torch.empty((B, a.shape[1]), dtype=torch.float32, device=dev))
        widths.append(a.shape[1])
    with torch.cuda.device(dev):
        rf = (C.c_int * len(widths))(*widths)
        _count(1)
        _lib.check(
            L.vb_gather_rows(_lib.ptr_array(list(arrays)), _lib.ptr_array(outs), rf, len(widths), _lib.ptr(indices), B,
                             int(n_rows), _lib.stream_ptr(dev)),
            "vb_gather_rows",
        )
    return tuple(outs)


def packed_row_layout(widths: Sequence[int]) -> Tuple[List[int], int]:
    """Offsets (floats) of each array's row inside one packed shadow row and the row stride: arrays whose width is a multiple
    of 4 first (they keep 16-byte alignment and move as 128-bit vectors), then even widths, then the rest; stride padded to 4."""
    order = sorted(range(len(widths)), key=lambda k: (0 if widths[k] % 4 == 0 else 1 if widths[k] % 2 == 0 else 2, k))
    offsets, off = [0] * len(widths), 0
    for k in order:
        offsets[k] = off
        off += widths[k]
    return offsets, (off + 3) // 4 * 4


def pack_rows(arrays: Sequence[torch.Tensor], packed: torch.Tensor, offsets: Sequence[int], stride: int,
              rows: Optional[torch.Tensor] = None, row0: int = 0, n: Optional[int] = None) -> None:
    """Refresh rows of the packed shadow table from the arrays: the int64 CUDA vector `rows`, or row0 .. row0 + n - 1."""
    L = _lib.lib()
    dev = packed.device
    widths = [a.shape[1] for a in arrays]
    if rows is not None:
        rows = rows.to(device=dev, dtype=torch.int64).contiguous()
        n = rows.numel()
    with torch.cuda.device(dev):
        _count(1)
        _lib.check(
            L.vb_pack_rows(_lib.ptr_array(list(arrays)), (C.c_int * len(widths))(*widths), (C.c_int * len(widths))(*offsets), len(widths),
                           _lib.ptr(packed), int(stride), _lib.ptr(rows), int(row0), int(n), int(packed.shape[0]), _lib.stream_ptr(dev)),
            "vb_pack_rows",
        )


def buffer_add_args(arrays: Sequence[torch.Tensor], shadow: Optional[torch.Tensor], offsets: Optional[Sequence[int]], stride: int):
    """The call-invariant ctypes arguments of vb_buffer_add_row for one buffer: arrays[k][row] = values[k] for all k in ONE launch (and the
    packed shadow row with it); see buffer_add_row_cached."""
    n = len(arrays)
    widths = [a.shape[1] for a in arrays]
    dev = arrays[0].device
    return dict(n=n, dev=dev, dst=_lib.ptr_array(list(arrays)), widths=(C.c_int * n)(*widths), capacity=int(arrays[0].shape[0]),
                shadow=_lib.ptr(shadow), stride=int(stride), offsets=(C.c_int * n)(*offsets) if offsets is not None else None,
                src=(C.c_void_p * n)(), scal=(C.c_float * n)(), keep=(arrays, shadow))


def buffer_add_row_cached(args, values: Sequence, row: int) -> None:
    """One transition into the buffer (vb_buffer_add_row) with the per-buffer arguments prepared once: values[k] is a float32 CUDA tensor of
    arrays[k].shape[1] elements or a Python number broadcast over the row; the per-call work is six pointer / scalar stores and one launch."""
    src, scal = args["src"], args["scal"]
    for k, v in enumerate(values):
        if isinstance(v, torch.Tensor):
            src[k] = v.data_ptr()
            scal[k] = 0.0
        else:
            src[k] = None
            scal[k] = v
    dev = args["dev"]
    global LAUNCHES
    LAUNCHES += 1
    with torch.cuda.device(dev):
        _lib.check(
            _lib.lib().vb_buffer_add_row(args["dst"], args["widths"], src, scal, args["n"], int(row), args["capacity"], args["shadow"],
                                         args["stride"], args["offsets"], _lib.stream_ptr(dev)),
            "vb_buffer_add_row",
        )


def gather_packed(packed: torch.Tensor, offsets: Sequence[int], stride: int, widths: Sequence[int], indices: torch.Tensor,
                  n_rows: int) -> Tuple[torch.Tensor, ...]:
    """gather_rows reading the packed shadow table (one run of sectors per transition)."""
    L = _lib.lib()
    dev = packed.device
    if indices.dtype != torch.int64 or indices.device != dev or not indices.is_contiguous():
        indices = indices.to(device=dev, dtype=torch.int64).contiguous()
    B = indices.numel()
    outs = [torch.empty((B, w), dtype=torch.float32, device=dev) for w in widths]
    with torch.cuda.device(dev):
        _count(1)
        _lib.check(
            L.vb_gather_packed(_lib.ptr(packed), int(stride), (C.c_int * len(widths))(*offsets), _lib.ptr_array(outs),
                               (C.c_int * len(widths))(*widths), len(widths), _lib.ptr(indices), B, int(n_rows), _lib.stream_ptr(dev)),
            "vb_gather_packed",
        )
    return tuple(outs)


def standard_normal_like(t: torch.Tensor) -> torch.Tensor:
    """The noise `Normal(mean, std).rsample()` draws (torch/distributions/normal.py: `_standard_normal(shape, dtype, device)`):
    the same generator call, so a seeded run consumes the RNG stream exactly like the reference.

    Data-parallel mode: every rank draws the noise of the GLOBAL batch (identically seeded generators give identical draws) and
    keeps the rows of its shard, so an N-rank step uses the same noise as the single-process step on the whole batch."""
    from torch.distributions.utils import _standard_normal

    ws = _dp.world_size() if _dp.is_enabled() else 1
    if ws == 1:
        return _standard_normal(t.shape, dtype=t.dtype, device=t.device)
    n = t.shape[0]
    full = _standard_normal((n * ws, *t.shape[1:]), dtype=t.dtype, device=t.device)
    r = _dp.rank()
    return full[r * n:(r + 1) * n]


class SacHeadFn(torch.autograd.Function):
    """actions [B,A], log_prob [B,1] = tail of SACActor.forward (velora/models/sac/continuous.py:121-143) for x [B,2A], eps [B,A]."""

    @staticmethod
    def forward(ctx, x, eps, scale, bias, lo, hi):
        L = _lib.lib()
        _lib.require_cuda(x, "SACActor.forward")
        dev = x.device
        x, eps = _f32c(x), _f32c(eps)
        scale, bias = _f32c(scale.reshape(-1)), _f32c(bias.reshape(-1))
        B, A2 = x.shape
        A = A2 // 2
        if scale.numel() != A or bias.numel() != A:
            scale, bias = scale.expand(A).contiguous(), bias.expand(A).contiguous()
        actions = torch.empty((B, A), dtype=torch.float32, device=dev)
        log_prob = torch.empty((B, 1), dtype=torch.float32, device=dev)
        a_norm = torch.empty((B, A), dtype=torch.float32, device=dev)
        std = torch.empty((B, A), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(
                L.vb_sac_head_fwd(_lib.ptr(x), _lib.ptr(eps), _lib.ptr(scale), _lib.ptr(bias), float(lo), float(hi), B, A,
                                  _lib.ptr(actions), _lib.ptr(log_prob), _lib.ptr(a_norm), _lib.ptr(std), _lib.stream_ptr(dev)),
                "vb_sac_head_fwd",
            )
        ctx.save_for_backward(x, eps, scale, a_norm, std)
        ctx.lims = (float(lo), float(hi))
        ctx.set_materialize_grads(False)
        return actions, log_prob

    @staticmethod
    def backward(ctx, d_actions, d_log_prob):
        x, eps, scale, a_norm, std = ctx.saved_tensors
        L = _lib.lib()
        dev = x.device
        B, A2 = x.shape
        dx = torch.empty_like(x)
        if d_actions is None and d_log_prob is None:
            return dx.zero_(), None, None, None, None, None
        da = _f32c(d_actions) if d_actions is not None else None
        dl = _f32c(d_log_prob.reshape(-1)) if d_log_prob is not None else None
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(
                L.vb_sac_head_bwd(_lib.ptr(x), _lib.ptr(eps), _lib.ptr(scale), ctx.lims[0], ctx.lims[1], B, A2 // 2, _lib.ptr(a_norm),
                                  _lib.ptr(std), _lib.ptr(da), _lib.ptr(dl), _lib.ptr(dx), _lib.stream_ptr(dev)),
                "vb_sac_head_bwd",
            )
        return dx, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------- fused SAC losses
_LOSS_WS = {}


def _loss_workspace(dev: torch.device) -> torch.Tensor:
    """Zero-initialised ticket + partial-sum area of the loss kernels, one per (device, stream); inside a graph capture a fresh one
    from the graph's pool (zeroed by a captured memset)."""
    n = int(_lib.lib().vb_sac_loss_workspace_bytes())
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros(n, dtype=torch.uint8, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _LOSS_WS.get(key)
    if ws is None:
        ws = torch.zeros(n, dtype=torch.uint8, device=dev)
        _LOSS_WS[key] = ws
    return ws


class SacCriticLossFn(torch.autograd.Function):
    """(c1_loss, c2_loss) of NeuroFlowCT._update_critics (agent.py:194-209): the target from the target critics, the next log-prob,
    reward and done flag (no gradient, as in the reference's `torch.no_grad()` block) and both MSE losses in one launch."""

    @staticmethod
    def forward(ctx, q1, q2, q1t, q2t, logp_next, rewards, dones, log_alpha, gamma):
        L = _lib.lib()
        dev = q1.device
        B = q1.numel()
        q1, q2, q1t, q2t, logp_next, rewards, dones = (_f32c(t.detach()).reshape(-1) for t in (q1, q2, q1t, q2t, logp_next, rewards, dones))
        la = _f32c(log_alpha.detach()).reshape(1)
        target = torch.empty(B, dtype=torch.float32, device=dev)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sac_critic_loss_fwd(_lib.ptr(q1), _lib.ptr(q2), _lib.ptr(q1t), _lib.ptr(q2t), _lib.ptr(logp_next), _lib.ptr(rewards),
                                                _lib.ptr(dones), _lib.ptr(la), float(gamma), B, _lib.ptr(target), _lib.ptr(losses), _lib.ptr(ws),
                                                _lib.stream_ptr(dev)), "vb_sac_critic_loss_fwd")
        ctx.save_for_backward(q1, q2, target)
        ctx.set_materialize_grads(False)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g1, g2):
        q1, q2, target = ctx.saved_tensors
        L = _lib.lib()
        dev = q1.device
        B = q1.numel()
        dq1 = torch.empty((B, 1), dtype=torch.float32, device=dev) if g1 is not None else None
        dq2 = torch.empty((B, 1), dtype=torch.float32, device=dev) if g2 is not None else None
        g1c = _f32c(g1).reshape(1) if g1 is not None else None
        g2c = _f32c(g2).reshape(1) if g2 is not None else None
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_sac_critic_loss_bwd(_lib.ptr(q1), _lib.ptr(q2), _lib.ptr(target), _lib.ptr(g1c), _lib.ptr(g2c), B, _lib.ptr(dq1),
                                                _lib.ptr(dq2), _lib.stream_ptr(dev)), "vb_sac_critic_loss_bwd")
        return dq1, dq2, None, None, None, None, None, None, None


class SacActorLossFn(torch.autograd.Function):
    """mean(exp(log_alpha) * logp - min(q1, q2)) (agent.py:238-241), gradients to logp, q1, q2 and log_alpha."""

    @staticmethod
    def forward(ctx, logp, q1, q2, log_alpha):
        L = _lib.lib()
        dev = logp.device
        B = logp.numel()
        lp, a, b = (_f32c(t.detach()).reshape(-1) for t in (logp, q1, q2))
        la = _f32c(log_alpha.detach()).reshape(1)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sac_actor_loss_fwd(_lib.ptr(lp), _lib.ptr(a), _lib.ptr(b), _lib.ptr(la), B, _lib.ptr(out), _lib.ptr(ws),
                                               _lib.stream_ptr(dev)), "vb_sac_actor_loss_fwd")
        ctx.save_for_backward(a, b, la, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a, b, la, out = ctx.saved_tensors
        L = _lib.lib()
        dev = a.device
        B = a.numel()
        gc = _f32c(g).reshape(1)
        dlogp = torch.empty((B, 1), dtype=torch.float32, device=dev)
        dq1 = torch.empty((B, 1), dtype=torch.float32, device=dev)
        dq2 = torch.empty((B, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_sac_actor_loss_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(la), _lib.ptr(gc), B, _lib.ptr(dlogp), _lib.ptr(dq1), _lib.ptr(dq2),
                                               _lib.stream_ptr(dev)), "vb_sac_actor_loss_bwd")
        # d/d log_alpha = g * alpha * mean(logp); the reference's backward populates it too (and the entropy step zeroes it again)
        dla = (gc * la.exp() * out[1]).reshape(()) if ctx.needs_input_grad[3] else None
        return dlogp, dq1, dq2, dla


class SacEntropyLossFn(torch.autograd.Function):
    """-mean(log_alpha * (logp + target_entropy).detach()) (modules.py:700-706)."""

    @staticmethod
    def forward(ctx, log_alpha, logp, target_entropy):
        L = _lib.lib()
        dev = logp.device
        B = logp.numel()
        lp = _f32c(logp.detach()).reshape(-1)
        la = _f32c(log_alpha.detach()).reshape(1)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sac_entropy_loss_fwd(_lib.ptr(lp), _lib.ptr(la), float(target_entropy), B, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr(dev)),
                       "vb_sac_entropy_loss_fwd")
        ctx.save_for_backward(out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return (-(g * out[1])).reshape(()), None, None


def _pack_liquid(L, dims, masks, params, dev, st):
    """vb_liquid_pack of the seven layers -> (packed block, masks as passed, heads mask type code)."""
    I, Ni, Nc, O = dims
    params = [_f32c(p.detach()) for p in params]
    masks = tuple(_lib.mask_arg(m, dev) for m in masks)
    heads_type = _lib.common_mask_type(masks[1:6])
    if heads_type is None:
        masks = (masks[0], *[m.contiguous() for m in masks[1:6]], masks[6])
        heads_type = _lib.mask_type(masks[1])
    packed = torch.empty(L.vb_liquid_packed_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev)
    _count(1)
    _lib.check(
        L.vb_liquid_pack(
            _lib.ptr(params[0]), _lib.ptr(params[1]), _lib.ptr(masks[0]), _lib.mask_type(masks[0]),
            _lib.ptr_array([params[2 + 2 * k] for k in range(5)]), _lib.ptr_array([params[3 + 2 * k] for k in range(5)]),
            _lib.ptr_array(list(masks[1:6])), heads_type,
            _lib.ptr(params[12]), _lib.ptr(params[13]), _lib.ptr(masks[6]), _lib.mask_type(masks[6]),
            I, Ni, Nc, O, _lib.ptr(packed), st,
        ),
        "vb_liquid_pack",
    )
    return packed, masks, heads_type


def liquid_bf16_forward(x: torch.Tensor, h0: Optional[torch.Tensor], dims, masks, params, want_traj: bool = False, train: bool = False):
    """Forward of the 512-neuron class on tcgen05 in bf16 (vb_liquid_bf16_fwd): x [B,T,I] -> y [B,T,O], h_T [B,Nc] (+ h_traj [B,T,Nc] when
    asked for).  `train`: also returns (x as the kernels read it, x time-major?, packed, masks, heads type, saved block) for the backward."""
    I, Ni, Nc, O = dims
    L = _lib.lib()
    _lib.require_cuda(x, "LiquidNCPNetwork.forward_seq_bf16")
    if not L.vb_liquid_bf16_ok(I, Ni, Nc, O):
        raise RuntimeError(f"no bf16 tensor-core forward for a liquid network of shape (in={I}, inter={Ni}, command={Nc}, out={O})")
    dev = x.device
    B, T, _ = x.shape
    x, x_tm = _seq(x, T > 1)
    h0c = _f32c(h0)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        packed, masks, heads_type = _pack_liquid(L, dims, masks, params, dev, st)
        y = torch.empty((T, B, O), dtype=torch.float32, device=dev).permute(1, 0, 2)
        h_last = torch.empty((B, Nc), dtype=torch.float32, device=dev)
        h_traj = torch.empty((T, B, Nc), dtype=torch.float32, device=dev).permute(1, 0, 2) if want_traj else None
        saved = torch.empty(L.vb_liquid_bf16_saved_bytes(I, Ni, Nc, O, B, T), dtype=torch.uint8, device=dev) if train and B > 0 else None
        ws = _workspace(dev, L.vb_liquid_bf16_fwd_workspace_bytes(I, Ni, Nc, O, B, T))
        flags = (X_TM if x_tm else 0) | Y_TM | H_TM
        _count(2)
        with _timed("liquid_bf16_fwd", 0):
            _lib.check(
                L.vb_liquid_bf16_fwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(h0c), _lib.ptr(y), _lib.ptr(h_last), _lib.ptr(h_traj),
                                     _lib.ptr(saved), 0 if saved is None else saved.numel(),
                                     B, T, I, Ni, Nc, O, _lib.ptr(ws), ws.numel(), flags, st),
                "vb_liquid_bf16_fwd",
            )
    if train:
        return y, h_last, (x, x_tm, packed, masks, heads_type, saved)
    return (y, h_last, h_traj) if want_traj else (y, h_last)


class LiquidBf16SeqFn(torch.autograd.Function):
    """y [B,T,O], h_T [B,Nc] = liquid(x [B,T,I], h0 | None) for the 512-neuron class with bf16 tensor-core operands, TRAINABLE:
    forward vb_liquid_bf16_fwd (saving the gate factors, mish(c) and z~), backward vb_liquid_bf16_bwd.  2e-2 relative contract."""

    @staticmethod
    def forward(ctx, x, h0, dims, masks, *params):
        y, h_last, (xk, x_tm, packed, masks, heads_type, saved) = liquid_bf16_forward(x, h0, dims, masks, params, train=True)
        ctx.dims, ctx.masks, ctx.heads_type, ctx.x_tm = dims, masks, heads_type, x_tm
        ctx.has_h0 = h0 is not None
        ctx.has_saved = saved is not None
        ctx.save_for_backward(xk, packed, saved if saved is not None else xk.new_empty(0))
        ctx.set_materialize_grads(False)
        return y, h_last

    @staticmethod
    def backward(ctx, dy, dh_last):
        x, packed, saved = ctx.saved_tensors
        I, Ni, Nc, O = ctx.dims
        masks = ctx.masks
        L = _lib.lib()
        dev = x.device
        B, T, _ = x.shape
        dy_tm = False
        if dy is None:
            dy = torch.zeros((B, T, O), dtype=torch.float32, device=dev)
        else:
            dy, dy_tm = _seq(dy, T > 1)
        dh_last = _f32c(dh_last)
        need_x = ctx.needs_input_grad[0]
        need_h0 = ctx.has_h0 and ctx.needs_input_grad[1]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            dx = torch.empty_like(x) if need_x else None
            dh0 = torch.empty((B, Nc), dtype=torch.float32, device=dev) if need_h0 else None
            dpacked = torch.empty(L.vb_liquid_grad_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev)
            ws = _workspace(dev, L.vb_liquid_bf16_bwd_workspace_bytes(I, Ni, Nc, O, B, T)) if B > 0 else None
            flags = (X_TM | DX_TM if ctx.x_tm else 0) | (Y_TM if dy_tm else 0)
            with _timed("liquid_bf16_bwd", 6):
                _lib.check(
                    L.vb_liquid_bf16_bwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dh_last),
                                         _lib.ptr(saved) if ctx.has_saved else None, saved.numel() if ctx.has_saved else 0,
                                         _lib.ptr(dx), _lib.ptr(dh0), _lib.ptr(dpacked), B, T, I, Ni, Nc, O,
                                         _lib.ptr(ws), 0 if ws is None else ws.numel(), flags, st),
                    "vb_liquid_bf16_bwd",
                )
            _dp.maybe_allreduce(dpacked)       # data-parallel: ONE collective over the flat block (526 k floats for this class)
            need = ctx.needs_input_grad[4:]
            shapes = [(Ni, I), (Ni,)] + [(Nc, Ni + Nc), (Nc,)] * 5 + [(O, Nc), (O,)]
            grads: List[Optional[torch.Tensor]] = [
                torch.empty(s, dtype=torch.float32, device=dev) if need[k] else None for k, s in enumerate(shapes)
            ]
            m_in, m_h, m_out = masks[0], list(masks[1:6]), masks[6]
            _count(1)
            _lib.check(
                L.vb_liquid_unpack_grads(
                    _lib.ptr(dpacked), _lib.ptr(m_in), _lib.mask_type(m_in), _lib.ptr_array(m_h), ctx.heads_type,
                    _lib.ptr(m_out), _lib.mask_type(m_out),
                    _lib.ptr(grads[0]), _lib.ptr(grads[1]),
                    _lib.ptr_array([grads[2 + 2 * k] for k in range(5)]), _lib.ptr_array([grads[3 + 2 * k] for k in range(5)]),
                    _lib.ptr(grads[12]), _lib.ptr(grads[13]), I, Ni, Nc, O, 0, st,
                ),
                "vb_liquid_unpack_grads",
            )
        return (dx, dh0, None, None, *grads)


# ------------------------------------------------------------------------------------------------ discrete agent (NeuroFlow) losses
class SacdCriticLossFn(torch.autograd.Function):
    """(c1_loss, c2_loss) of NeuroFlow._update_critics (agent.py:549-590): the soft state value of the next state from the actor's
    probabilities and the target critics (no gradient), the Q-values of the taken actions, both MSE losses - one launch."""

    @staticmethod
    def forward(ctx, q1, q2, actions, next_probs, q1t, q2t, rewards, dones, log_alpha, gamma):
        L = _lib.lib()
        dev = q1.device
        B, A = q1.shape
        q1, q2, next_probs, q1t, q2t = (_f32c(t.detach()) for t in (q1, q2, next_probs, q1t, q2t))
        actions, rewards, dones = (_f32c(t.detach()).reshape(-1) for t in (actions, rewards, dones))
        la = _f32c(log_alpha.detach()).reshape(1)
        target = torch.empty(B, dtype=torch.float32, device=dev)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sacd_critic_loss_fwd(_lib.ptr(q1), _lib.ptr(q2), _lib.ptr(actions), _lib.ptr(next_probs), _lib.ptr(q1t), _lib.ptr(q2t),
                                                 _lib.ptr(rewards), _lib.ptr(dones), _lib.ptr(la), float(gamma), B, A, _lib.ptr(target),
                                                 _lib.ptr(losses), _lib.ptr(ws), _lib.stream_ptr(dev)), "vb_sacd_critic_loss_fwd")
        ctx.save_for_backward(q1, q2, actions, target)
        ctx.set_materialize_grads(False)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g1, g2):
        q1, q2, actions, target = ctx.saved_tensors
        L = _lib.lib()
        dev = q1.device
        B, A = q1.shape
        dq1 = torch.empty((B, A), dtype=torch.float32, device=dev) if g1 is not None else None
        dq2 = torch.empty((B, A), dtype=torch.float32, device=dev) if g2 is not None else None
        g1c = _f32c(g1).reshape(1) if g1 is not None else None
        g2c = _f32c(g2).reshape(1) if g2 is not None else None
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_sacd_critic_loss_bwd(_lib.ptr(q1), _lib.ptr(q2), _lib.ptr(actions), _lib.ptr(target), _lib.ptr(g1c), _lib.ptr(g2c), B, A,
                                                 _lib.ptr(dq1), _lib.ptr(dq2), _lib.stream_ptr(dev)), "vb_sacd_critic_loss_bwd")
        return (dq1, dq2) + (None,) * 8


class SacdActorLossFn(torch.autograd.Function):
    """mean_b sum_a p(a) (exp(log_alpha) logp_b - min(q1, q2)(a)) (agent.py:605-612); gradients to p, logp and log_alpha - the critics are
    constants here (their gradient of this loss is never used)."""

    @staticmethod
    def forward(ctx, probs, logp, q1, q2, log_alpha):
        L = _lib.lib()
        dev = probs.device
        B, A = probs.shape
        p, a, b = (_f32c(t.detach()) for t in (probs, q1, q2))
        lp = _f32c(logp.detach()).reshape(-1)
        la = _f32c(log_alpha.detach()).reshape(1)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sacd_actor_loss_fwd(_lib.ptr(p), _lib.ptr(lp), _lib.ptr(a), _lib.ptr(b), _lib.ptr(la), B, A, _lib.ptr(out), _lib.ptr(ws),
                                                _lib.stream_ptr(dev)), "vb_sacd_actor_loss_fwd")
        ctx.save_for_backward(p, lp, a, b, la, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        p, lp, a, b, la, out = ctx.saved_tensors
        L = _lib.lib()
        dev = p.device
        B, A = p.shape
        gc = _f32c(g).reshape(1)
        dprobs = torch.empty((B, A), dtype=torch.float32, device=dev)
        dlogp = torch.empty((B, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_sacd_actor_loss_bwd(_lib.ptr(p), _lib.ptr(lp), _lib.ptr(a), _lib.ptr(b), _lib.ptr(la), _lib.ptr(gc), B, A,
                                                _lib.ptr(dprobs), _lib.ptr(dlogp), _lib.stream_ptr(dev)), "vb_sacd_actor_loss_bwd")
        dla = (gc * la.exp() * out[1]).reshape(()) if ctx.needs_input_grad[4] else None
        return dprobs, dlogp, None, None, dla


class SacdEntropyLossFn(torch.autograd.Function):
    """log_alpha (H - target), H = -mean_b sum_a p(a) log(p(a) + 1e-8) with no gradient into p (modules.py:820-838)."""

    @staticmethod
    def forward(ctx, log_alpha, probs, target_entropy):
        L = _lib.lib()
        dev = probs.device
        B, A = probs.shape
        p = _f32c(probs.detach())
        la = _f32c(log_alpha.detach()).reshape(1)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _loss_workspace(dev)
            _count(1)
            _lib.check(L.vb_sacd_entropy_loss_fwd(_lib.ptr(p), _lib.ptr(la), float(target_entropy), B, A, _lib.ptr(out), _lib.ptr(ws),
                                                  _lib.stream_ptr(dev)), "vb_sacd_entropy_loss_fwd")
        ctx.save_for_backward(out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return (g * out[1]).reshape(()), None, None


class CategoricalHeadFn(torch.autograd.Function):
    """(actions [B] int64, probs [B,A], log_prob [B,1]) = the tail of SACActorDiscrete.forward (discrete.py:41-57, 80-101) for logits [B,A] and
    Exp(1) noise [B,A] drawn by the caller with torch's generator - one launch forward, one backward."""

    @staticmethod
    def forward(ctx, logits, noise):
        L = _lib.lib()
        dev = logits.device
        B, A = logits.shape
        x, q = _f32c(logits.detach()), _f32c(noise)
        probs = torch.empty((B, A), dtype=torch.float32, device=dev)
        actions = torch.empty((B,), dtype=torch.int64, device=dev)
        logp = torch.empty((B, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_categorical_head_fwd(_lib.ptr(x), _lib.ptr(q), B, A, _lib.ptr(probs), _lib.ptr(actions), _lib.ptr(logp),
                                                 _lib.stream_ptr(dev)), "vb_categorical_head_fwd")
        ctx.save_for_backward(probs, actions)
        ctx.mark_non_differentiable(actions)
        ctx.set_materialize_grads(False)
        return actions, probs, logp

    @staticmethod
    def backward(ctx, _ga, d_probs, d_logp):
        probs, actions = ctx.saved_tensors
        if d_probs is None and d_logp is None:
            return None, None
        L = _lib.lib()
        dev = probs.device
        B, A = probs.shape
        dp = _f32c(d_probs) if d_probs is not None else None
        dl = _f32c(d_logp).reshape(-1) if d_logp is not None else None
        dx = torch.empty((B, A), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _count(1)
            _lib.check(L.vb_categorical_head_bwd(_lib.ptr(probs), _lib.ptr(actions), _lib.ptr(dp), _lib.ptr(dl), B, A, _lib.ptr(dx),
                                                 _lib.stream_ptr(dev)), "vb_categorical_head_bwd")
        return dx, None
