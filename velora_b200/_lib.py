"""ctypes binding of libvelora_b200.so (the C ABI declared in include/velora_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VELORA_B200_LIB selects another build of the same ABI (the clock64-instrumented copy used by profiles/microbench/trace_run.py)
LIB_PATH = os.environ.get("VELORA_B200_LIB") or os.path.join(_HERE, "lib", "libvelora_b200.so")

VB_MASK_I32, VB_MASK_F32 = 0, 1
VB_FLAG_FORCE_GENERIC = 1

_lib: Optional[C.CDLL] = None

_i, _i64, _p = C.c_int, C.c_int64, C.c_void_p
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/velora_b200.h one to one
SIGNATURES = {
    "vb_query": (_i, [_i]),
    "vb_last_error": (C.c_char_p, []),
    "vb_liquid_packed_floats": (_i64, [_i, _i, _i, _i]),
    "vb_liquid_grad_floats": (_i64, [_i, _i, _i, _i]),
    "vb_liquid_has_fast_path": (_i, [_i, _i, _i, _i]),
    "vb_liquid_time_major_ok": (_i, [_i, _i, _i, _i, _i]),
    "vb_liquid_pack": (_i, [_p, _p, _p, _i, _pp, _pp, _pp, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "vb_liquid_saved_bytes": (_i64, [_i, _i, _i, _i, _i64, _i, _i]),
    "vb_liquid_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i64, _i, _i]),
    "vb_liquid_fwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "vb_liquid_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "vb_liquid_unpack_grads": (_i, [_p, _p, _i, _pp, _i, _p, _i, _p, _p, _pp, _pp, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vb_cell_bwd_workspace_bytes": (_i64, [_i, _i]),
    "vb_cell_fwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _p]),
    "vb_cell_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _i64, _p]),
    "vb_liquid_bf16_ok": (_i, [_i, _i, _i, _i]),
    "vb_liquid_bf16_fwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i64, _i]),
    "vb_liquid_bf16_saved_bytes": (_i64, [_i, _i, _i, _i, _i64, _i]),
    "vb_liquid_bf16_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "vb_liquid_bf16_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i64, _i]),
    "vb_liquid_bf16_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "vb_ncp_packed_floats": (_i64, [_i, _i, _i, _i]),
    "vb_ncp_grad_floats": (_i64, [_i, _i, _i, _i]),
    "vb_ncp_pack": (_i, [_pp, _pp, _pp, _i, _i, _i, _i, _i, _p, _p]),
    "vb_ncp_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i64]),
    "vb_ncp_fwd": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, _p]),
    "vb_ncp_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "vb_ncp_unpack_grads": (_i, [_p, _pp, _i, _pp, _pp, _i, _i, _i, _i, _i, _p]),
    "vb_gather_rows": (_i, [_pp, _pp, C.POINTER(C.c_int), _i, _p, _i64, _i64, _p]),
    "vb_pack_rows": (_i, [_pp, C.POINTER(C.c_int), C.POINTER(C.c_int), _i, _p, _i, _p, _i64, _i64, _i64, _p]),
    "vb_transpose_rows": (_i, [_p, _p, _i64, _i, _i, _i, _p]),
    "vb_buffer_add_row": (_i, [_pp, _p, _pp, _p, _i, _i64, _i64, _p, _i, _p, _p]),
    "vb_gather_packed": (_i, [_p, _i, C.POINTER(C.c_int), _pp, C.POINTER(C.c_int), _i, _p, _i64, _i64, _p]),
    "vb_adam_multi": (_i, [_pp, _pp, _pp, _pp, _pp, C.POINTER(C.c_int64), _i, C.c_float, C.c_float, C.c_float, C.c_float, _p]),
    "vb_sac_head_fwd": (_i, [_p, _p, _p, _p, C.c_float, C.c_float, _i64, _i, _p, _p, _p, _p, _p]),
    "vb_sac_head_bwd": (_i, [_p, _p, _p, C.c_float, C.c_float, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "vb_sac_loss_workspace_bytes": (_i64, []),
    "vb_sac_critic_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, C.c_float, _i64, _p, _p, _p, _p]),
    "vb_sac_critic_loss_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "vb_sac_actor_loss_fwd": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p]),
    "vb_sac_actor_loss_bwd": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "vb_sac_entropy_loss_fwd": (_i, [_p, _p, C.c_float, _i64, _p, _p, _p]),
    "vb_categorical_head_fwd": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p]),
    "vb_categorical_head_bwd": (_i, [_p, _p, _p, _p, _i64, _i, _p, _p]),
    "vb_sacd_critic_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, C.c_float, _i64, _i, _p, _p, _p, _p]),
    "vb_sacd_critic_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _p, _p, _p]),
    "vb_sacd_actor_loss_fwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _p, _p, _p]),
    "vb_sacd_actor_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _p, _p, _p]),
    "vb_sacd_entropy_loss_fwd": (_i, [_p, _p, C.c_float, _i64, _i, _p, _p, _p]),
    "vb_allreduce_oneshot_buffer_bytes": (_i64, [_i64]),
    "vb_allreduce_oneshot": (_i, [_p, _i64, _pp, _i, _i, _i64, C.c_float, _p, _p]),
    "vb_polyak_multi": (_i, [_pp, _pp, C.POINTER(C.c_int64), _i, C.c_float, _p]),
}


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into velora_b200/lib/ (nvcc cross-compiles without a GPU)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    r = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libvelora_b200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(velora_b200 has no CPU or eager fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(ts))()
    for k, t in enumerate(ts):
        arr[k] = None if t is None else t.data_ptr()
    return arr


VB_MASK_TRANSPOSED = 2


def mask_type(t: torch.Tensor) -> int:
    """Type code of a 2-D mask that is either row-major or a transposed view of a row-major tensor (see mask_arg)."""
    if t.dtype == torch.int32:
        code = VB_MASK_I32
    elif t.dtype == torch.float32:
        code = VB_MASK_F32
    else:
        raise TypeError(f"mask dtype {t.dtype} not supported (int32 or float32)")
    return code if t.is_contiguous() else code | VB_MASK_TRANSPOSED


def common_mask_type(masks) -> "int | None":
    """One type code for a group of masks that the C ABI takes with a single code, or None when they really differ.  A mask with
    a dimension of size 1 (the [1, Nc] motor mask of a one-output critic) is row-major AND column-major: it adopts the others'."""
    codes = set()
    dtype_codes = set()
    for m in masks:
        dtype_codes.add(mask_type(m) & ~VB_MASK_TRANSPOSED)
        ambiguous = m.dim() == 2 and m.is_contiguous() and m.t().is_contiguous()
        if not ambiguous:
            codes.add(mask_type(m))
    if len(dtype_codes) != 1 or len(codes) > 1:
        return None
    return codes.pop() if codes else mask_type(masks[0])


def mask_arg(m: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """The reference stores `abs(mask.T)`: a transposed VIEW.  Row-major and transposed-row-major masks are passed
    as they are (the kernels index them through the type code); anything else is copied once."""
    if m.device != dev:
        m = m.to(dev)
    if m.is_contiguous() or (m.dim() == 2 and m.t().is_contiguous()):
        return m
    return m.contiguous()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: velora_b200 runs on CUDA (sm_100a) only and has no CPU fallback; got a tensor on '{t.device}'. "
            "Construct the network/buffer with device=torch.device('cuda')."
        )
