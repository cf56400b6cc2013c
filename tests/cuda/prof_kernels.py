"""Launch every hot kernel at the Llama-3-8B shapes (twice each: warm-up + measured) -- the target of the ncu captures
committed under profiles/.  Usage on the GPU box:
  ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tests/cuda/prof_kernels.py [names...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transformers_b200 import ops

which = set(sys.argv[1:])
want = lambda n: not which or n in which
dev = "cuda"
BF = torch.bfloat16
T, H, I, Hq, Hkv, D, B, S = 16384, 4096, 14336, 32, 8, 128, 4, 4096
rn = lambda *s: torch.randn(*s, device=dev).to(BF)
reps = 2

if want("gemm"):
    x, w = rn(T, H), rn(2 * I, H) * 0.02
    for _ in range(reps): y = ops.gemm(x, w)                       # fwd gate|up   [NT]
    dy = rn(T, 2 * I)
    for _ in range(reps): dx = ops.gemm(dy, w, b_mn=True)          # dgrad         [NN]
    for _ in range(reps): dw = ops.gemm(dy, x, a_mn=True, b_mn=True)  # wgrad      [TT]
    for _ in range(reps): ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)  # wgrad, TMA reduce-add epilogue
    w_ilv = ops.interleave_gate_up(w[:I], w[I:])
    for _ in range(reps): gu, h = ops.gemm_glu(x, w_ilv)          # fwd gate|up with the GLU epilogue
    del x, w, y, dy, dx, dw, w_ilv, gu, h
if want("attn"):
    qkv = rn(B, S, (Hq + 2 * Hkv) * D)
    q = qkv[..., : Hq * D].view(B, S, Hq, D); k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D); v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
    for _ in range(reps): out, lse = ops.attn_fwd(q, k, v, scale=D ** -0.5, causal=True)
    do = rn(B, S, Hq, D); dqkv = torch.empty_like(qkv)
    dq = dqkv[..., : Hq * D].view(B, S, Hq, D); dk = dqkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D); dv = dqkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
    for _ in range(reps): ops.attn_bwd(q, k, v, out, do, lse, dq, dk, dv, scale=D ** -0.5, causal=True)
    del qkv, out, do, dqkv
if want("elementwise"):
    x, w = rn(T, H), rn(H)
    for _ in range(reps): y, rstd, _ = ops.rmsnorm_fwd(x, w, 1e-5)
    for _ in range(reps): ops.rmsnorm_bwd(y, x, w, rstd)
    for _ in range(reps): ops.rmsnorm_fwd(x, w, 1e-5, residual=y)
    gu = rn(T, 2 * I)
    for _ in range(reps): h = ops.glu_fwd(gu)
    for _ in range(reps): ops.glu_bwd(h, gu)
    qkv = rn(B, S, (Hq + 2 * Hkv) * D); cos = rn(1, S, D); sin = rn(1, S, D)
    for _ in range(reps): ops.rope_(qkv, cos, sin, Hq + Hkv, D)
    ids = torch.randint(0, 128256, (B, S), device=dev); emb = rn(128256, H)
    for _ in range(reps): e = ops.embedding_fwd(ids, emb)
    for _ in range(reps): ops.add(x, y)
    inv = torch.rand(D // 2, device=dev); pos = torch.arange(S, device=dev)[None]
    for _ in range(reps): ops.rope_table(inv, pos)
    gi = rn(T, 2 * I)
    for _ in range(reps): ops.glu_bwd(h, gi, interleaved=True)
    del gu, h, qkv, emb, e, gi
if want("ce"):
    logits = rn(B, S, 128256); labels = torch.randint(0, 128256, (B, S), device=dev)
    for _ in range(reps): loss, lse, denom = ops.ce_fwd(logits, labels)
    for _ in range(reps): ops.ce_bwd(logits, labels, lse, torch.ones((), device=dev), denom)
torch.cuda.synchronize()
print("done")
