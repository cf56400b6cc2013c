"""ctypes binding of the C-ABI library (``include/b200_ops.h``).

The product path has no CPU fallback: if the library is missing, fails to load, or the device is not sm_100, every op
raises.  Pointers cross the boundary as raw addresses (``tensor.data_ptr()``) plus sizes and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libb200.so")

_p, _i, _f, _l = c_void_p, c_int, c_float, c_int64

# name -> argtypes (return type is always int unless listed in _RESTYPES)
SIGNATURES = {
    "b200_abi_version": [],
    "b200_device_check": [],
    "b200_last_error": [],
    "b200_gemm_bf16": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_gemv_bf16": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "b200_gemm_bf16_2sm": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_gemm_bf16_grouped": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_gemm_glu_bf16": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_gemm_bf16_1sm": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_embedding_fwd": [_p, _p, _p, _i, _i, _i, _f, _i, _p, _p],
    "b200_embedding_bwd": [_p, _p, _p, _i, _i, _i, _l, _f, _i, _p],
    "b200_rmsnorm_fwd": [_p, _p, _p, _p, _p, _p, _i, _i, _f, _i, _p],
    "b200_rmsnorm_bwd_workspace_rows": [],
    "b200_rmsnorm_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "b200_rope": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_rope_table": [_p, _p, _p, _p, _i, _i, _f, _p],
    "b200_glu_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b200_glu_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "b200_add_bf16": [_p, _p, _p, _l, _p],
    "b200_kv_append": [_p, _p, _p, _p, _i, _i, _i, _i] + [_l] * 9 + [_i, _i, _p],
    "b200_moe_route": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "b200_moe_gather": [_p, _p, _p, _i, _i, _p],
    "b200_moe_combine": [_p, _p, _p, _p, _i, _i, _i, _p],
    "b200_ce_fwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _l, _f, _p],
    "b200_attn_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i] + [_l] * 12 + [_f, _f, _i, _i, _p, _p, _p],
    "b200_attn_decode_splits": [_i, _i, _i],
    "b200_attn_decode": [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i] + [_l] * 10 + [_f, _f, _i, _p, _p, _p],
    "b200_attn_bwd": [_p] * 10 + [_i] * 7 + [_p, _f, _f, _i, _i, _p, _p, _p],
    "b200_ce_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _l, _p],
    "b200_ce_bwd_sharded": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "b200_optim_chunk_elems": [],
    "b200_adamw_step": [_p, _p, _i, _i, _f, _f, _f, _f, _f, _f, _f, _p, _p],
    "b200_grad_norm": [_p, _p, _i, _p, _f, _p, _p],
    "b200_grad_scale": [_p, _p, _i, _p, _p],
    "b200_gemm_bf16_scatter": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_pull_reduce_bf16": [_p, _i, _l, _l, _p, _p, _p],
}
_RESTYPES = {"b200_last_error": c_char_p}

_lib = None


class B200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libb200.so (building is the job of ``__graft_entry__.build()`` / ``transformers_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m transformers_b200.build). "
            "There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header/library drift -> fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200Error(f"{what} failed (code {rc}): {last_error()}")


_device_ok = False


def require_device() -> None:
    """Raise unless the current CUDA device is a B200 (sm_100)."""
    global _device_ok
    if _device_ok:
        return
    import torch

    if not torch.cuda.is_available():
        raise B200Error("transformers_b200 needs a CUDA sm_100 (B200) device; no CPU fallback exists")
    check(load().b200_device_check(), "device check")
    _device_ok = True
