"""The CUDA-core kernels that were written without a GPU (optim.cu, peer.cu, gemv.cu, attention_decode.cu, ce_sharded.cu), executed on the
HOST by tests/emu (one std::thread per CUDA thread, real warp-shuffle and barrier semantics): the kernels' own source is
compiled with g++ and checked against torch.  This pins indexing, vector/tail paths, shuffle reductions and shared-memory
merges before the first device run; device-only aspects (coalescing, latency, the launchers) are left to
tests/test_experimental_gpu.py."""
import ctypes
import math
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CUDA_INC = "/usr/local/cuda/include"
p, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "libemu.so")
    r = subprocess.run([gxx, "-std=c++20", "-O1", "-fPIC", "-shared", "-Wno-attributes", "-I", EMU, "-I",
                        os.path.join(ROOT, "transformers_b200", "csrc"), "-I", CUDA_INC, os.path.join(EMU, "emu_kernels.cpp"),
                        "-o", out, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = ctypes.CDLL(out)
    lib.emu_adamw.argtypes = [p, p, i32, i32, f32, f32, f32, f32, f32, f32, f32, p]
    lib.emu_grad_norm.argtypes = [p, p, i32, p, f32, p]
    lib.emu_grad_scale.argtypes = [p, p, i32, p]
    lib.emu_pull_reduce.argtypes = [p, i32, i64, i64, p, p, i32]
    lib.emu_gemv.argtypes = [p, p, p, i32, i32, i32, i32, i32, i32]
    lib.emu_ce_bwd_sharded.argtypes = [p, p, p, p, p, i32, i32, i32, i32]
    lib.emu_attn_decode.argtypes = [p, p, p, p, p, i32, p, i32, i32, i32, i32, i32] + [i64] * 10 + [f32, f32, i32, p, p, i32]
    return lib


BF = torch.bfloat16


def _tables(entries, chunk=32768):
    rows, cmap = [], []
    for i, (a, b, c, d, n) in enumerate(entries):
        rows.append((a, b, c, d, n, 0))
        cmap.extend((i, k) for k in range((n + chunk - 1) // chunk))
    return torch.tensor(rows, dtype=torch.int64), torch.tensor(cmap, dtype=torch.int32)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("state_dtype", [BF, torch.float32])
def test_adamw_norm_scale_kernels_emulated(emu, state_dtype):
    from oracle import adamw_oracle as O

    torch.manual_seed(0)
    shapes = [7, 64 * 33, 40003, 8]  # scalar tail only, vector body, multi-chunk with a ragged tail, exactly one vector
    ps = [(torch.randn(n) * 0.5).to(BF) for n in shapes]
    gs = [(torch.randn(n) * 2).to(BF) for n in shapes]
    ms = [(torch.randn(n) * 0.1).to(state_dtype) for n in shapes]
    vs = [(torch.rand(n) * 0.1).to(state_dtype) for n in shapes]
    want = [O.adamw_step(a, b, c, d, 3, 1e-2, 0.9, 0.95, 1e-8, 0.1, grad_scale=0.5) for a, b, c, d in zip(ps, gs, ms, vs)]
    table, cmap = _tables([(a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), a.numel()) for a, b, c, d in zip(ps, gs, ms, vs)])
    # global norm + clip coefficient, then in-place scale of a copy
    partial = torch.empty(cmap.shape[0], dtype=torch.float32)
    out2 = torch.empty(2, dtype=torch.float32)
    emu.emu_grad_norm(table.data_ptr(), cmap.data_ptr(), cmap.shape[0], partial.data_ptr(), 1.0, out2.data_ptr())
    total, coef = O.grad_norm_and_coef(gs, 1.0)
    assert abs(out2[0].item() - total) < 1e-4 * total and abs(out2[1].item() - coef) < 1e-6
    g2 = [g.clone() for g in gs]
    t2, c2 = _tables([(0, g.data_ptr(), 0, 0, g.numel()) for g in g2])
    coef_t = torch.tensor([0.5], dtype=torch.float32)
    emu.emu_grad_scale(t2.data_ptr(), c2.data_ptr(), c2.shape[0], coef_t.data_ptr())
    for a, b in zip(g2, gs):
        assert torch.equal(a, (b.float() * 0.5).to(BF))
    # AdamW step 3 with fused grad scale 0.5
    gsc = torch.tensor([0.5], dtype=torch.float32)
    emu.emu_adamw(table.data_ptr(), cmap.data_ptr(), cmap.shape[0], int(state_dtype == torch.float32), 1e-2, 0.9, 0.95, 1e-8, 0.1,
                  1 - 0.9 ** 3, math.sqrt(1 - 0.95 ** 3), gsc.data_ptr())
    for (pw, mw, vw), a, c, d in zip(want, ps, ms, vs):
        torch.testing.assert_close(a.float(), pw.float(), atol=1e-6, rtol=8e-3)  # <= 1 bf16 ulp
        tol = dict(atol=1e-7, rtol=1e-5) if state_dtype == torch.float32 else dict(atol=1e-6, rtol=8e-3)
        torch.testing.assert_close(c.float(), mw.float(), **tol)
        torch.testing.assert_close(d.float(), vw.float(), **tol)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_pull_reduce_kernel_emulated(emu, world):
    torch.manual_seed(world)
    rows, cols = 24, 88  # 2112 elements per rank slice = 264 vectors: grid-stride loop with a partial last wave
    bufs = [torch.randn(world * rows, cols).to(BF) for _ in range(world)]
    res = torch.randn(rows, cols).to(BF)
    ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    for rank in (0, world - 1):
        out = torch.empty(rows, cols, dtype=BF)
        emu.emu_pull_reduce(ctypes.cast(ptrs, p), world, rank * rows * cols, rows * cols, res.data_ptr(), out.data_ptr(), 2)
        want = res.float()
        for b in bufs:
            want = want + b[rank * rows:(rank + 1) * rows].float()
        assert torch.equal(out, want.to(BF))
        emu.emu_pull_reduce(ctypes.cast(ptrs, p), world, rank * rows * cols, rows * cols, None, out.data_ptr(), 1)
        want = sum(b[rank * rows:(rank + 1) * rows].float() for b in bufs)
        assert torch.equal(out, want.to(BF))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,N,K", [(1, 19, 264), (4, 8, 1024), (3, 9, 8), (2, 16, 520)])
def test_gemv_kernel_emulated(emu, M, N, K):
    torch.manual_seed(N + K)
    x = torch.randn(M, K).to(BF)
    w = (torch.randn(N, K) * 0.2).to(BF)
    y = torch.full((M, N), float("nan"), dtype=BF)
    assert emu.emu_gemv(x.data_ptr(), w.data_ptr(), y.data_ptr(), M, N, K, K, K, N) == 0
    torch.testing.assert_close(y.float(), x.float() @ w.float().t(), atol=2e-2, rtol=8e-3)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("B,Hq,Hkv,D,ctx,window,softcap,nsplit", [
    (2, 4, 2, 64, 37, 0, 0.0, 3), (1, 8, 1, 128, 70, 16, 0.0, 2), (1, 2, 2, 256, 9, 0, 30.0, 4), (2, 4, 4, 64, 1, 0, 0.0, 2),
    (1, 4, 1, 128, 21, 0, 0.0, 1),
])
def test_decode_attention_kernels_emulated(emu, B, Hq, Hkv, D, ctx, window, softcap, nsplit):
    torch.manual_seed(ctx + D)
    cap = ctx + 3
    kc = torch.randn(B, Hkv, cap, D).to(BF)  # KV-cache layout: [B, Hkv, capacity, D]
    vc = torch.randn(B, Hkv, cap, D).to(BF)
    q = torch.randn(B, 1, Hq, D).to(BF)
    out = torch.full((B, 1, Hq, D), float("nan"), dtype=BF)
    lse = torch.full((B, Hq, 128), float("nan"), dtype=torch.float32)
    ws = torch.full((B * Hq * nsplit * (D + 2),), float("nan"), dtype=torch.float32)
    kv_start = torch.tensor([0] + [3] * (B - 1), dtype=torch.int32) if ctx > 8 else None
    kv_end = torch.tensor([ctx] + [ctx - 2] * (B - 1), dtype=torch.int32) if ctx > 8 else None
    scale = D ** -0.5
    rc = emu.emu_attn_decode(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), lse.data_ptr(), 128, ws.data_ptr(), B, ctx,
                             Hq, Hkv, D, q.stride(0), q.stride(2), kc.stride(0), kc.stride(2), kc.stride(1), vc.stride(0),
                             vc.stride(2), vc.stride(1), out.stride(0), out.stride(2), scale, softcap, window,
                             kv_start.data_ptr() if kv_start is not None else None,
                             kv_end.data_ptr() if kv_end is not None else None, nsplit)
    assert rc == 0
    G = Hq // Hkv
    k = kc[:, :, :ctx].transpose(1, 2).float().repeat_interleave(G, dim=2)
    v = vc[:, :, :ctx].transpose(1, 2).float().repeat_interleave(G, dim=2)
    s = torch.einsum("bhd,bkhd->bhk", q[:, 0].float(), k) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    idx = torch.arange(ctx)
    valid = torch.ones(B, ctx, dtype=torch.bool)
    if window:
        valid &= idx[None] >= ctx - window
    if kv_start is not None:
        valid &= (idx[None] >= kv_start[:, None]) & (idx[None] < kv_end[:, None])
    s = s.masked_fill(~valid[:, None], float("-inf"))
    want = torch.einsum("bhk,bkhd->bhd", torch.softmax(s, -1), v)
    torch.testing.assert_close(out[:, 0].float(), want, atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(lse[..., 0], torch.logsumexp(s, -1), atol=1e-3, rtol=1e-3)


@pytest.mark.timeout(300)
def test_ce_bwd_sharded_kernel_emulated(emu):
    """Two vocabulary shards: the sharded gradient kernel with the GLOBAL lse must reproduce the full softmax - one-hot."""
    torch.manual_seed(5)
    T, V, N = 3, 8 * 1030 + 4, 2  # per shard: > 1024 vectors (second loop trip for some threads) plus a scalar tail
    Vl = V // N
    logits = (torch.randn(T, V) * 2).to(BF)
    tgt = torch.tensor([5, V - 2, Vl + 1])
    row_scale = torch.tensor([0.25, 0.0, 0.5])
    lse = torch.logsumexp(logits.float(), -1)
    want = torch.softmax(logits.float(), -1)
    want[torch.arange(T), tgt] -= 1.0
    want = want * row_scale[:, None]
    for r in range(N):
        shard = logits[:, r * Vl:(r + 1) * Vl].contiguous()
        local = tgt - r * Vl
        local = torch.where((local >= 0) & (local < Vl), local, torch.full_like(local, -1))
        out = torch.full((T, Vl), float("nan"), dtype=BF)
        emu.emu_ce_bwd_sharded(shard.data_ptr(), local.data_ptr(), lse.data_ptr(), row_scale.data_ptr(), out.data_ptr(), T, Vl, Vl, Vl)
        torch.testing.assert_close(out.float(), want[:, r * Vl:(r + 1) * Vl], atol=2e-6, rtol=8e-3)
