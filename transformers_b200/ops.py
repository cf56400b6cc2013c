"""Tensor-level wrappers over the C-ABI (no autograd here; see ``functional.py``).

Every function takes CUDA bf16 tensors, allocates its outputs with torch (device memory + streams are the only things
torch is used for), passes raw pointers / sizes / the current stream to ``libb200.so`` and raises on any non-zero
return code.  There is deliberately no CPU or library fallback.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import B200Error
from ._lib import check as _check

BF16 = torch.bfloat16

# kernels launched per C-ABI entry point
_KERNELS_PER_CALL = {"b200_attn_decode": 2, "b200_rmsnorm_bwd": 2, "b200_ce_fwd": 2, "b200_attn_bwd": 3, "b200_moe_route": 3, "b200_grad_norm": 2}


def check(rc: int, what: str) -> None:
    _check(rc, what)
    _count(_KERNELS_PER_CALL.get(what, 1))


_LAUNCHES = 0  # kernels launched through the C-ABI by this process (bench.py reports it as gpu_launches)


def launch_count() -> int:
    return _LAUNCHES


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _lib_ready():
    _lib.require_device()
    return _lib.load()


def _chk_bf16(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != BF16:
            raise B200Error(f"expected a CUDA bfloat16 tensor, got {t.device} {t.dtype}")


# ----------------------------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out: torch.Tensor | None = None,
         accumulate: bool = False) -> torch.Tensor:
    """D[M,N] (+)= sum_k A(m,k) B(n,k).  ``a`` is [M,K] (or [K,M] when a_mn), ``b`` is [N,K] (or [K,N] when b_mn);
    both 2-D with unit inner stride."""
    lib = _lib_ready()
    _chk_bf16(a, b, out)
    if a.dim() != 2 or b.dim() != 2 or (a.stride(1) != 1 and a.shape[1] != 1) or (b.stride(1) != 1 and b.shape[1] != 1):
        raise B200Error("gemm operands must be 2-D with unit inner stride")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise B200Error(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        if accumulate:
            raise B200Error("gemm: accumulate needs an output tensor")
        out = torch.empty(M, N, device=a.device, dtype=BF16)
    if out.shape != (M, N) or out.stride(1) != 1:
        raise B200Error("gemm: bad output tensor")
    if M == 0 or N == 0:
        return out
    check(lib.b200_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                             int(a_mn), int(b_mn), int(accumulate), _stream()), "b200_gemm_bf16")
    return out


GLU_BLOCK = 128  # column block of the interleaved gate|up layout (half of the GEMM's 256-column tile)


def interleave_gate_up(wg: torch.Tensor, wu: torch.Tensor) -> torch.Tensor:
    """[I, K] gate and up weights -> [2I, K] with 128-row blocks alternating gate / up (what b200_gemm_glu_bf16 expects)."""
    I, K = wg.shape
    if wu.shape != wg.shape or I % GLU_BLOCK:
        raise B200Error(f"interleave_gate_up: shapes {tuple(wg.shape)} / {tuple(wu.shape)}, I must be a multiple of {GLU_BLOCK}")
    return torch.stack([wg.reshape(I // GLU_BLOCK, GLU_BLOCK, K), wu.reshape(I // GLU_BLOCK, GLU_BLOCK, K)], dim=1).reshape(2 * I, K).contiguous()


def deinterleave_gate_up(t: torch.Tensor):
    """Inverse on the leading dimension: [2I, ...] interleaved rows -> (gate [I, ...], up [I, ...])."""
    n = t.shape[0] // (2 * GLU_BLOCK)
    v = t.reshape(n, 2, GLU_BLOCK, *t.shape[1:])
    return v[:, 0].reshape(n * GLU_BLOCK, *t.shape[1:]), v[:, 1].reshape(n * GLU_BLOCK, *t.shape[1:])


def gemm_glu(a: torch.Tensor, w_ilv: torch.Tensor, gelu: bool = False, gu_out: torch.Tensor | None = None,
             h_out: torch.Tensor | None = None):
    """a [M, K] @ block-interleaved gate|up weight [2I, K] -> (gu [M, 2I] interleaved columns, h [M, I] = act(gate) * up);
    one kernel: the activation runs in the GEMM's epilogue (csrc/gemm2.cu GLU mode)."""
    lib = _lib_ready()
    _chk_bf16(a, w_ilv, gu_out, h_out)
    M, K = a.shape
    I = w_ilv.shape[0] // 2
    if a.stride(1) != 1 or w_ilv.stride(1) != 1 or w_ilv.shape[1] != K:
        raise B200Error("gemm_glu: operands must be 2-D with unit inner stride and matching contraction")
    gu = gu_out if gu_out is not None else torch.empty(M, 2 * I, device=a.device, dtype=BF16)
    h = h_out if h_out is not None else torch.empty(M, I, device=a.device, dtype=BF16)
    if gu.shape != (M, 2 * I) or h.shape != (M, I) or gu.stride(1) != 1 or h.stride(1) != 1:
        raise B200Error("gemm_glu: bad output tensors")
    check(lib.b200_gemm_glu_bf16(a.data_ptr(), w_ilv.data_ptr(), gu.data_ptr(), h.data_ptr(), M, I, K, a.stride(0), w_ilv.stride(0),
                                 gu.stride(0), h.stride(0), int(gelu), _stream()), "b200_gemm_glu_bf16")
    return gu, h


def gemm_grouped(a: torch.Tensor, b: torch.Tensor, offsets: torch.Tensor, *, b_mn: bool = False,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """Rows [offsets[g], offsets[g+1]) of a [M, K] times expert g's matrix b[g] ([N, K], or [K, N] when ``b_mn``) -> the same
    rows of out [M, N]; ``offsets`` int32 [G + 1] on the device (no host sync, one launch for all experts)."""
    lib = _lib_ready()
    _chk_bf16(a, b, out)
    G = b.shape[0]
    M, K = a.shape
    N = b.shape[2] if b_mn else b.shape[1]
    if b.dim() != 3 or not b.is_contiguous() or a.stride(1) != 1 or (b.shape[1] if b_mn else b.shape[2]) != K:
        raise B200Error("gemm_grouped: b must be a contiguous [G, N, K] (or [G, K, N]) tensor matching a's contraction")
    if offsets.dtype != torch.int32 or offsets.numel() != G + 1 or not offsets.is_cuda:
        raise B200Error("gemm_grouped: offsets must be a device int32 tensor with G + 1 entries")
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=BF16)
    if out.shape != (M, N) or out.stride(1) != 1:
        raise B200Error("gemm_grouped: bad output tensor")
    check(lib.b200_gemm_bf16_grouped(a.data_ptr(), b.data_ptr(), out.data_ptr(), offsets.data_ptr(), G, M, N, K, a.stride(0),
                                     b.stride(1), out.stride(0), int(b_mn), _stream()), "b200_gemm_bf16_grouped")
    return out


def grouped_ok(gate_up: torch.Tensor, down: torch.Tensor) -> bool:
    """Expert stacks the grouped kernel takes (<= 64 experts, widths in whole 64-column chunks)."""
    E, I2, H = gate_up.shape
    return E <= 64 and I2 % 64 == 0 and H % 64 == 0 and down.shape[2] % 64 == 0 and gate_up.is_contiguous() and down.is_contiguous()


def glu_fusable(M: int, I: int) -> bool:
    """Shapes the GLU-epilogue GEMM takes (CTA-pair kernel: more than one 128-row tile; whole 128-column blocks)."""
    return M > 128 and I % GLU_BLOCK == 0


# ------------------------------------------------------------------------------------------------------ embedding
class _DeferredFlag:
    """Out-of-range token ids: the gather kernel raises a device flag instead of asserting (the reference's F.embedding
    device-asserts).  Reading the flag right away would put a host sync on every forward, so it is read at the NEXT call --
    by then the kernel that wrote it has long finished (``event.query()``; if it has not, the check moves on to the call
    after) -- and raises there.  The flag travels to PINNED host memory with an asynchronous copy right behind the kernel, so
    reading it never synchronises the stream (a ``.item()`` on the device flag would wait for everything queued behind it).
    ``poll(force=True)`` waits for the copies (tests, end of a step)."""

    def __init__(self):
        self.pending = []  # (flag tensor, event or None)

    def push(self, flag, event=None):
        self.pending.append((flag, event))

    def poll(self, force: bool = False):
        keep = []
        for flag, event in self.pending:
            if force and event is not None:
                event.synchronize()
            if force or event is None or event.query():
                if int(flag[0]) != 0:
                    self.pending = []
                    raise B200Error("embedding: input_ids contain values outside [0, num_embeddings) "
                                    "(detected at the call after the offending forward)")
            else:
                keep.append((flag, event))
        self.pending = keep


_EMBED_FLAGS = _DeferredFlag()


def embedding_check_now() -> None:
    """Synchronously raise if any earlier embedding gather saw an out-of-range token id."""
    _EMBED_FLAGS.poll(force=True)


def embedding_fwd(ids: torch.Tensor, weight: torch.Tensor, scale: float | None = None) -> torch.Tensor:
    lib = _lib_ready()
    _chk_bf16(weight)
    _EMBED_FLAGS.poll()
    ids_c = ids.contiguous().view(-1)
    if ids_c.dtype != torch.int64:
        ids_c = ids_c.to(torch.int64)
    T, (V, H) = ids_c.numel(), weight.shape
    out = torch.empty(*ids.shape, H, device=weight.device, dtype=BF16)
    err = torch.zeros(1, device=weight.device, dtype=torch.int32)
    w = weight if weight.is_contiguous() else weight.contiguous()
    check(lib.b200_embedding_fwd(ids_c.data_ptr(), w.data_ptr(), out.data_ptr(), T, H, V,
                                 float(scale) if scale is not None else 1.0, int(scale is not None), err.data_ptr(), _stream()),
          "b200_embedding_fwd")
    host = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host.copy_(err, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _EMBED_FLAGS.push(host, ev)
    return out


def embedding_bwd(ids: torch.Tensor, dout: torch.Tensor, num_embeddings: int, padding_idx: int | None,
                  scale: float | None = None) -> torch.Tensor:
    lib = _lib_ready()
    ids_c = ids.contiguous().view(-1).to(torch.int64)
    dout = dout.contiguous()
    H = dout.shape[-1]
    dw = torch.zeros(num_embeddings, H, device=dout.device, dtype=BF16)
    check(lib.b200_embedding_bwd(ids_c.data_ptr(), dout.data_ptr(), dw.data_ptr(), ids_c.numel(), H, num_embeddings,
                                 -1 if padding_idx is None else int(padding_idx),
                                 float(scale) if scale is not None else 1.0, int(scale is not None), _stream()),
          "b200_embedding_bwd")
    return dw


# -------------------------------------------------------------------------------------------------------- RMSNorm
def rmsnorm_fwd(x: torch.Tensor, weight: torch.Tensor, eps: float, gemma: bool = False, residual: torch.Tensor | None = None):
    """Returns (y, rstd, residual_out).  With ``residual`` the kernel first forms r = bf16(x + residual)."""
    lib = _lib_ready()
    _chk_bf16(x, weight, residual)
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    T = x2.shape[0]
    y = torch.empty_like(x2)
    rstd = torch.empty(T, device=x.device, dtype=torch.float32)
    res_out = None
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, H).contiguous()
        res_out = torch.empty_like(x2)
    check(lib.b200_rmsnorm_fwd(x2.data_ptr(), r2.data_ptr() if r2 is not None else None, weight.data_ptr(),
                               res_out.data_ptr() if res_out is not None else None, y.data_ptr(), rstd.data_ptr(), T, H,
                               float(eps), int(gemma), _stream()), "b200_rmsnorm_fwd")
    return y.view(x.shape), rstd, (res_out.view(x.shape) if res_out is not None else None)


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, rstd: torch.Tensor, gemma: bool = False):
    lib = _lib_ready()
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    dy2 = dy.reshape(-1, H)
    if not dy2.is_contiguous():
        dy2 = dy2.contiguous()
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    T = x2.shape[0]
    dx = torch.empty_like(x2)
    dw = torch.empty(H, device=x.device, dtype=BF16)
    ws = torch.empty(lib.b200_rmsnorm_bwd_workspace_rows() * H, device=x.device, dtype=torch.float32)
    check(lib.b200_rmsnorm_bwd(dy2.data_ptr(), x2.data_ptr(), weight.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                               dw.data_ptr(), ws.data_ptr(), T, H, int(gemma), 0, _stream()), "b200_rmsnorm_bwd")
    return dx.view(x.shape), dw


# ----------------------------------------------------------------------------------------------------------- RoPE
def rope_table(inv_freq: torch.Tensor, position_ids: torch.Tensor, attention_scaling: float = 1.0, dtype: torch.dtype = BF16):
    """inv_freq fp32 [D/2], position_ids [B, S] (any integer dtype) -> (cos, sin) bf16 [B, S, D]
    (LlamaRotaryEmbedding.forward, models/llama/modeling_llama.py:113-127)."""
    lib = _lib_ready()
    if dtype != BF16:
        raise B200Error(f"rope_table: only bfloat16 tables are produced, got {dtype}")
    if not inv_freq.is_cuda or position_ids.device != inv_freq.device:
        raise B200Error(f"rope_table: expected CUDA tensors on one device, got {inv_freq.device} / {position_ids.device}")
    f = inv_freq.detach().to(torch.float32).contiguous()
    pos = position_ids.to(torch.int64).contiguous()
    B, S = pos.shape
    D = 2 * f.numel()
    cos = torch.empty(B, S, D, device=f.device, dtype=BF16)
    sin = torch.empty(B, S, D, device=f.device, dtype=BF16)
    check(lib.b200_rope_table(f.data_ptr(), pos.data_ptr(), cos.data_ptr(), sin.data_ptr(), B * S, D, float(attention_scaling),
                              _stream()), "b200_rope_table")
    return cos, sin


def rope_(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_rot_heads: int, head_dim: int, backward: bool = False):
    """In place on qkv [B, S, W] (heads packed along W, the first ``n_rot_heads`` are rotated)."""
    lib = _lib_ready()
    _chk_bf16(qkv, cos, sin)
    B, S, W = qkv.shape
    if not qkv.is_contiguous():
        raise B200Error("rope_: qkv must be contiguous")
    cos = cos.contiguous()
    sin = sin.contiguous()
    cb = cos.shape[0] if cos.dim() == 3 else 1
    if cos.shape[-2] != S or cos.shape[-1] != head_dim:
        raise B200Error(f"rope_: cos/sin shape {tuple(cos.shape)} does not match S={S}, D={head_dim}")
    check(lib.b200_rope(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, S, n_rot_heads, head_dim, W, cb, int(backward),
                        _stream()), "b200_rope")
    return qkv


# ------------------------------------------------------------------------------------------------------------ GLU
def glu_fwd(gu: torch.Tensor, gelu: bool = False, interleaved: bool = False) -> torch.Tensor:
    """gu [..., 2I] = [gate | up] (``interleaved``: 128-column blocks alternating gate / up)  ->  act(gate) * up  [..., I]"""
    lib = _lib_ready()
    _chk_bf16(gu)
    I = gu.shape[-1] // 2
    g2 = gu.reshape(-1, 2 * I)
    T = g2.shape[0]
    out = torch.empty(T, I, device=gu.device, dtype=BF16)
    up_off = 2 * (GLU_BLOCK if interleaved else I)  # bytes from the gate pointer to the up pointer
    check(lib.b200_glu_fwd(g2.data_ptr(), g2.data_ptr() + up_off, out.data_ptr(), T, I, g2.stride(0), I,
                           int(gelu) | (2 if interleaved else 0), _stream()), "b200_glu_fwd")
    return out.view(*gu.shape[:-1], I)


def glu_bwd(dh: torch.Tensor, gu: torch.Tensor, gelu: bool = False, interleaved: bool = False) -> torch.Tensor:
    """d(gate|up) in the layout of ``gu`` (plain halves or interleaved 128-column blocks)."""
    lib = _lib_ready()
    I = gu.shape[-1] // 2
    g2 = gu.reshape(-1, 2 * I)
    d2 = dh.reshape(-1, I)
    if not d2.is_contiguous():
        d2 = d2.contiguous()
    T = g2.shape[0]
    dgu = torch.empty(T, 2 * I, device=gu.device, dtype=BF16)
    up_off = 2 * (GLU_BLOCK if interleaved else I)
    check(lib.b200_glu_bwd(d2.data_ptr(), g2.data_ptr(), g2.data_ptr() + up_off, dgu.data_ptr(), dgu.data_ptr() + up_off, T, I,
                           I, g2.stride(0), 2 * I, int(gelu) | (2 if interleaved else 0), _stream()), "b200_glu_bwd")
    return dgu.view(gu.shape)


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    lib = _lib_ready()
    _chk_bf16(a, b)
    a = a.contiguous()
    b = b.contiguous()
    out = torch.empty_like(a)
    check(lib.b200_add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "b200_add_bf16")
    return out


# ------------------------------------------------------------------------------------------------------ attention
def _bsh_strides(t: torch.Tensor):
    """t is a [B, S, h, D] view with unit last stride -> (batch, row, head) strides in elements."""
    if t.stride(3) != 1:
        raise B200Error("attention operands need a unit stride on head_dim")
    return t.stride(0), t.stride(1), t.stride(2)


def _lse_stride(sq: int) -> int:
    return (sq + 127) // 128 * 128


def attn_fwd(q, k, v, *, scale: float, causal: bool, window: int = 0, softcap: float = 0.0, kv_start=None, kv_end=None,
             out: torch.Tensor | None = None, decode_kernel: bool | None = None):
    """q [B,Sq,Hq,D], k/v [B,Skv,Hkv,D] strided views -> (out [B,Sq,Hq,D], lse [B,Hq,lse_stride] fp32).
    q_len == 1 goes to the split-context decode kernels (``decode_kernel=False`` forces the tensor-core kernel: tests)."""
    lib = _lib_ready()
    _chk_bf16(q, k, v)
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    if out is None:
        out = torch.empty(B, Sq, Hq, D, device=q.device, dtype=BF16)
    ls = _lse_stride(Sq)
    lse = torch.empty(B, Hq, ls, device=q.device, dtype=torch.float32)
    if Sq == 1 and Hq // Hkv in (1, 2, 4, 8) and decode_kernel is not False:
        # decode step: split-context CUDA-core kernel (attention_decode.cu) instead of a 1/128-full tensor-core q tile
        nsplit = lib.b200_attn_decode_splits(B, Hkv, Skv)
        ws = torch.empty(B * Hq * nsplit * (D + 2), device=q.device, dtype=torch.float32)
        kb, kr, kh = _bsh_strides(k)
        vb, vr, vh = _bsh_strides(v)
        check(lib.b200_attn_decode(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), ls, ws.data_ptr(),
                                   B, Skv, Hq, Hkv, D, q.stride(0), q.stride(2), kb, kr, kh, vb, vr, vh, out.stride(0),
                                   out.stride(2), float(scale), float(softcap or 0.0), int(window or 0),
                                   kv_start.data_ptr() if kv_start is not None else None,
                                   kv_end.data_ptr() if kv_end is not None else None, _stream()), "b200_attn_decode")
        return out, lse
    check(lib.b200_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), ls, B, Sq, Skv, Hq, Hkv, D,
                            *_bsh_strides(q), *_bsh_strides(k), *_bsh_strides(v), *_bsh_strides(out), float(scale),
                            float(softcap or 0.0), int(causal), int(window or 0),
                            kv_start.data_ptr() if kv_start is not None else None,
                            kv_end.data_ptr() if kv_end is not None else None, _stream()), "b200_attn_fwd")
    return out, lse


def attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, *, scale: float, causal: bool, window: int = 0, softcap: float = 0.0,
             kv_start=None, kv_end=None):
    """Writes dq/dk/dv (strided [B,S,h,D] views, e.g. slices of one packed buffer)."""
    lib = _lib_ready()
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    ls = lse.shape[-1]
    ws = torch.empty(2 * B * Hq * ls, device=q.device, dtype=torch.float32)
    strides = []
    for t in (q, k, v, out, dout, dq, dk, dv):
        strides += list(_bsh_strides(t))
    sarr = (ctypes.c_int64 * 24)(*strides)
    check(lib.b200_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                            dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), B, Sq, Skv, Hq, Hkv, D, ls,
                            ctypes.cast(sarr, ctypes.c_void_p), float(scale), float(softcap or 0.0), int(causal),
                            int(window or 0), kv_start.data_ptr() if kv_start is not None else None,
                            kv_end.data_ptr() if kv_end is not None else None, _stream()), "b200_attn_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------------------------------------------- KV cache
def kv_append(k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, offset: int) -> None:
    """Write k_new / v_new [B,H,q,D] (strided) into rows [offset, offset+q) of the [B,H,capacity,D] caches, in place."""
    lib = _lib_ready()
    _chk_bf16(k_new, v_new, k_cache, v_cache)
    B, H, q, D = k_new.shape
    if k_new.stride(3) != 1 or v_new.stride(3) != 1 or k_cache.stride(3) != 1 or k_cache.stride() != v_cache.stride():
        raise B200Error("kv_append: unit inner stride and identical k/v cache layouts required")
    check(lib.b200_kv_append(k_new.data_ptr(), v_new.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), B, H, q, D,
                             k_new.stride(0), k_new.stride(1), k_new.stride(2), v_new.stride(0), v_new.stride(1), v_new.stride(2),
                             k_cache.stride(0), k_cache.stride(1), k_cache.stride(2), int(offset), k_cache.shape[2], _stream()),
          "b200_kv_append")


# ------------------------------------------------------------------------------------------------------------ MoE
def moe_experts_forward(x: torch.Tensor, top_k_index: torch.Tensor, top_k_weights: torch.Tensor, gate_up: torch.Tensor,
                        down: torch.Tensor, gelu: bool = False) -> torch.Tensor:
    """x [T,H] bf16, top_k_index [T,k] int64, top_k_weights [T,k], gate_up [E,2I,H], down [E,H,I] -> [T,H] (inference)."""
    lib = _lib_ready()
    _chk_bf16(x, gate_up, down)
    T, H = x.shape
    k = top_k_index.shape[1]
    E, I2, _ = gate_up.shape
    I = I2 // 2
    dev = x.device
    n = T * k
    idx = top_k_index.contiguous().to(torch.int64)
    w = top_k_weights.contiguous().to(torch.float32)
    counts = torch.zeros(E, device=dev, dtype=torch.int32)
    offsets = torch.empty(E + 1, device=dev, dtype=torch.int32)
    cursor = torch.empty(E, device=dev, dtype=torch.int32)
    slot = torch.empty(n, device=dev, dtype=torch.int32)
    tok = torch.empty(n, device=dev, dtype=torch.int32)
    check(lib.b200_moe_route(idx.data_ptr(), counts.data_ptr(), offsets.data_ptr(), cursor.data_ptr(), slot.data_ptr(),
                             tok.data_ptr(), T, k, E, _stream()), "b200_moe_route")
    xs = torch.empty(n, H, device=dev, dtype=BF16)
    x = x.contiguous()
    check(lib.b200_moe_gather(x.data_ptr(), tok.data_ptr(), xs.data_ptr(), n, H, _stream()), "b200_moe_gather")
    if grouped_ok(gate_up, down):
        # ONE grouped GEMM per projection: the expert row ranges stay on the device (no host sync, no per-expert launches)
        gu = gemm_grouped(xs, gate_up, offsets)
        act = glu_fwd(gu, gelu)
        ys = gemm_grouped(act, down, offsets)
    else:
        off = offsets.tolist()  # odd widths: per-expert launches, their row ranges are host-side launch parameters
        gu = torch.empty(n, 2 * I, device=dev, dtype=BF16)
        ys = torch.empty(n, H, device=dev, dtype=BF16)
        for e in range(E):
            lo, hi = off[e], off[e + 1]
            if hi > lo:
                gemm(xs[lo:hi], gate_up[e], out=gu[lo:hi])
        act = glu_fwd(gu, gelu)
        for e in range(E):
            lo, hi = off[e], off[e + 1]
            if hi > lo:
                gemm(act[lo:hi], down[e], out=ys[lo:hi])
    out = torch.empty(T, H, device=dev, dtype=BF16)
    check(lib.b200_moe_combine(ys.data_ptr(), slot.data_ptr(), w.data_ptr(), out.data_ptr(), T, k, H, _stream()), "b200_moe_combine")
    return out


def moe_route(top_k_index: torch.Tensor, E: int):
    """top_k_index [T,k] int64 -> (offsets int32 [E+1], slot int32 [T*k] position of pair (t,j) in expert-sorted order,
    tok int32 [T*k] token of every sorted slot)."""
    lib = _lib_ready()
    T, k = top_k_index.shape
    dev, n = top_k_index.device, T * k
    idx = top_k_index.contiguous().to(torch.int64)
    counts = torch.zeros(E, device=dev, dtype=torch.int32)
    offsets = torch.empty(E + 1, device=dev, dtype=torch.int32)
    cursor = torch.empty(E, device=dev, dtype=torch.int32)
    slot = torch.empty(n, device=dev, dtype=torch.int32)
    tok = torch.empty(n, device=dev, dtype=torch.int32)
    check(lib.b200_moe_route(idx.data_ptr(), counts.data_ptr(), offsets.data_ptr(), cursor.data_ptr(), slot.data_ptr(),
                             tok.data_ptr(), T, k, E, _stream()), "b200_moe_route")
    return offsets, slot, tok


def moe_gather(x: torch.Tensor, tok: torch.Tensor) -> torch.Tensor:
    """rows x[tok[s]] -> [n, H]"""
    lib = _lib_ready()
    _chk_bf16(x)
    x = x.contiguous()
    n, H = tok.numel(), x.shape[1]
    xs = torch.empty(n, H, device=x.device, dtype=BF16)
    check(lib.b200_moe_gather(x.data_ptr(), tok.data_ptr(), xs.data_ptr(), n, H, _stream()), "b200_moe_gather")
    return xs


def moe_combine(ys: torch.Tensor, slot: torch.Tensor, weights: torch.Tensor, T: int, k: int) -> torch.Tensor:
    """out[t] = sum_j weights[t,j] * ys[slot[t*k+j]]  ([T, H] bf16; weights fp32 [T,k])"""
    lib = _lib_ready()
    _chk_bf16(ys)
    H = ys.shape[1]
    w = weights.contiguous().to(torch.float32)
    out = torch.empty(T, H, device=ys.device, dtype=BF16)
    check(lib.b200_moe_combine(ys.data_ptr(), slot.data_ptr(), w.data_ptr(), out.data_ptr(), T, k, H, _stream()), "b200_moe_combine")
    return out


# ----------------------------------------------------------------------------------------------------------- loss
def ce_fwd(logits: torch.Tensor, labels: torch.Tensor, shift: bool = True, ignore_index: int = -100,
           num_items: float | None = None):
    """logits [B,S,V] bf16, labels [B,S] int64 -> (loss fp32 scalar tensor, lse [B*S], denom [1])."""
    lib = _lib_ready()
    _chk_bf16(logits)
    B, S, V = logits.shape
    lg = logits.reshape(B * S, V)
    if not lg.is_contiguous():
        lg = lg.contiguous()
    labels = labels.contiguous().to(torch.int64)
    dev = logits.device
    lse = torch.empty(B * S, device=dev, dtype=torch.float32)
    rows = torch.empty(B * S, device=dev, dtype=torch.float32)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    denom = torch.empty(1, device=dev, dtype=torch.float32)
    check(lib.b200_ce_fwd(lg.data_ptr(), labels.data_ptr(), lse.data_ptr(), rows.data_ptr(), loss.data_ptr(), denom.data_ptr(),
                          B, S, V, lg.stride(0), int(shift), int(ignore_index), float(num_items) if num_items else 0.0, _stream()),
          "b200_ce_fwd")
    return loss, lse, denom


def ce_bwd(logits: torch.Tensor, labels: torch.Tensor, lse: torch.Tensor, dloss: torch.Tensor, denom: torch.Tensor,
           shift: bool = True, ignore_index: int = -100, out: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib_ready()
    B, S, V = logits.shape
    lg = logits.reshape(B * S, V)
    if not lg.is_contiguous():
        lg = lg.contiguous()
    labels = labels.contiguous().to(torch.int64)
    if out is not None:
        _chk_bf16(out)
        if out.shape != (B * S, V) or not out.is_contiguous():
            raise B200Error("ce_bwd: `out` must be a contiguous [B*S, V] bf16 tensor")
    dl = out if out is not None else torch.empty(B * S, V, device=logits.device, dtype=BF16)
    dloss = dloss.reshape(1).to(torch.float32).contiguous()
    check(lib.b200_ce_bwd(lg.data_ptr(), labels.data_ptr(), lse.data_ptr(), dloss.data_ptr(), denom.data_ptr(), dl.data_ptr(), B, S,
                          V, lg.stride(0), V, int(shift), int(ignore_index), _stream()), "b200_ce_bwd")
    return dl.view(B, S, V)


def ce_row_lse(logits: torch.Tensor) -> torch.Tensor:
    """logits [T, V] bf16 -> fp32 log-sum-exp of every row over these V columns (b200_ce_fwd with no targets)."""
    lib = _lib_ready()
    _chk_bf16(logits)
    T, V = logits.shape
    lg = logits if logits.is_contiguous() else logits.contiguous()
    dev = logits.device
    labels = torch.full((T,), -100, device=dev, dtype=torch.int64)
    lse = torch.empty(T, device=dev, dtype=torch.float32)
    rows = torch.empty(T, device=dev, dtype=torch.float32)
    scratch = torch.empty(2, device=dev, dtype=torch.float32)
    check(lib.b200_ce_fwd(lg.data_ptr(), labels.data_ptr(), lse.data_ptr(), rows.data_ptr(), scratch.data_ptr(),
                          scratch.data_ptr() + 4, 1, T, V, lg.stride(0), 0, -100, 1.0, _stream()), "b200_ce_fwd")
    return lse


def ce_bwd_sharded(logits: torch.Tensor, target_local: torch.Tensor, lse_global: torch.Tensor,
                   row_scale: torch.Tensor) -> torch.Tensor:
    """Vocabulary-sharded CE gradient: logits [T, V_local] bf16, target_local [T] int64 (outside [0, V_local) when another
    rank owns the target), lse_global [T] fp32, row_scale [T] fp32 -> dlogits [T, V_local]."""
    lib = _lib_ready()
    _chk_bf16(logits)
    T, V = logits.shape
    lg = logits if logits.is_contiguous() else logits.contiguous()
    dl = torch.empty(T, V, device=logits.device, dtype=BF16)
    tl = target_local.contiguous().to(torch.int64)
    ls = lse_global.contiguous().to(torch.float32)
    rs = row_scale.contiguous().to(torch.float32)
    check(lib.b200_ce_bwd_sharded(lg.data_ptr(), tl.data_ptr(), ls.data_ptr(), rs.data_ptr(), dl.data_ptr(), T, V, lg.stride(0),
                                  V, _stream()), "b200_ce_bwd_sharded")
    return dl


# ------------------------------------------------------------------------------------------------------ optimizer
def optim_chunk_elems() -> int:
    return int(_lib.load().b200_optim_chunk_elems())


def adamw_step(table: torch.Tensor, chunk_map: torch.Tensor, *, state_fp32: bool, master: bool = False, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float, bias_correction1: float, bias_correction2_sqrt: float,
               grad_scale: torch.Tensor | None = None) -> None:
    """One AdamW update of every tensor listed in ``table`` (device int64 [n,6], see include/b200_ops.h), in place."""
    lib = _lib_ready()
    check(lib.b200_adamw_step(table.data_ptr(), chunk_map.data_ptr(), chunk_map.shape[0], int(state_fp32) | (2 if master else 0), float(lr), float(beta1),
                              float(beta2), float(eps), float(weight_decay), float(bias_correction1),
                              float(bias_correction2_sqrt), grad_scale.data_ptr() if grad_scale is not None else None,
                              _stream()), "b200_adamw_step")


def grad_norm(table: torch.Tensor, chunk_map: torch.Tensor, max_norm: float = 0.0) -> torch.Tensor:
    """-> fp32 [2] on the device: (global L2 norm of the listed gradients, clip coefficient for ``max_norm``)."""
    lib = _lib_ready()
    n = chunk_map.shape[0]
    ws = torch.empty(max(n, 1), device=table.device, dtype=torch.float32)
    out = torch.empty(2, device=table.device, dtype=torch.float32)
    check(lib.b200_grad_norm(table.data_ptr(), chunk_map.data_ptr(), n, ws.data_ptr(), float(max_norm), out.data_ptr(), _stream()),
          "b200_grad_norm")
    return out


def grad_scale_(table: torch.Tensor, chunk_map: torch.Tensor, coef: torch.Tensor) -> None:
    lib = _lib_ready()
    check(lib.b200_grad_scale(table.data_ptr(), chunk_map.data_ptr(), chunk_map.shape[0], coef.data_ptr(), _stream()),
          "b200_grad_scale")


# ---------------------------------------------------------------------------------------------------- peer memory
def gemm_scatter(a: torch.Tensor, b: torch.Tensor, dest_ptrs: list[int], rank: int, *, a_mn: bool = False,
                 b_mn: bool = False) -> None:
    """GEMM whose epilogue stores row block r of the [M, N] result into ``dest_ptrs[r]`` ([M / world, N] bf16, contiguous:
    slots of peer-mapped buffers) -- the first half of a reduce-scatter fused into the GEMM (csrc/gemm2.cu SCATTER mode)."""
    lib = _lib_ready()
    _chk_bf16(a, b)
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb or a.stride(1) != 1 or b.stride(1) != 1:
        raise B200Error("gemm_scatter: operands must be 2-D with unit inner stride and matching contraction")
    arr = (ctypes.c_void_p * len(dest_ptrs))(*dest_ptrs)
    check(lib.b200_gemm_bf16_scatter(a.data_ptr(), b.data_ptr(), ctypes.cast(arr, ctypes.c_void_p), len(dest_ptrs), int(rank),
                                     M, N, K, a.stride(0), b.stride(0), N, int(a_mn), int(b_mn), _stream()),
          "b200_gemm_bf16_scatter")


def pull_reduce(peer_ptrs: list[int], offset_elems: int, n_elems: int, out: torch.Tensor,
                residual: torch.Tensor | None = None) -> torch.Tensor:
    """out[i] = bf16(residual[i] + sum_s peer_s[offset + i]) where peer_ptrs are the base addresses of every rank's
    peer-mapped partial buffer (rank order, own buffer included).  The loads are the NVLink transfer."""
    lib = _lib_ready()
    _chk_bf16(out, residual)
    if not out.is_contiguous() or out.numel() != n_elems or (residual is not None and (not residual.is_contiguous() or residual.numel() != n_elems)):
        raise B200Error("pull_reduce: out / residual must be contiguous with n_elems elements")
    arr = (ctypes.c_void_p * len(peer_ptrs))(*peer_ptrs)
    check(lib.b200_pull_reduce_bf16(ctypes.cast(arr, ctypes.c_void_p), len(peer_ptrs), int(offset_elems), int(n_elems),
                                    residual.data_ptr() if residual is not None else None, out.data_ptr(), _stream()),
          "b200_pull_reduce_bf16")
    return out
