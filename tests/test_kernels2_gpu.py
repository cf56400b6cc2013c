"""GPU parity of the kernels behind decode (GEMV, split-context attention), the optimizer step, the vocabulary-sharded
loss, the peer-memory reduction, the GEMM scatter epilogue, MoE backward and the Gemma (v1) class map -- written at the end
of round 1 without a GPU, first run on a B200 in round 2 (profiles/r02_call1_validation.md) and part of the regular `-m gpu`
suite since."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ce_sharded_matches_full_ce():
    """Two vocabulary shards combined by hand must reproduce b200_ce_fwd / b200_ce_bwd on the full logits bit for bit in the
    statistics (same fp32 reductions per shard) and to bf16 rounding in the gradient."""
    from transformers_b200 import ops

    torch.manual_seed(0)
    B, S, V, N = 2, 64, 4096, 2
    logits = (torch.randn(B, S, V, device="cuda") * 3).to(torch.bfloat16)
    labels = torch.randint(0, V, (B, S), device="cuda")
    labels[0, :5] = -100
    loss, lse, denom = ops.ce_fwd(logits, labels, shift=True)
    dl = ops.ce_bwd(logits, labels, lse, torch.ones((), device="cuda"), denom, shift=True)
    tgt = torch.full_like(labels, -100)
    tgt[:, :-1] = labels[:, 1:]
    tgt = tgt.reshape(-1)
    valid = tgt != -100
    shards = logits.reshape(B * S, V).chunk(N, dim=-1)
    lses = torch.stack([ops.ce_row_lse(s.contiguous()) for s in shards])
    lse_g = torch.logsumexp(lses, 0)
    torch.testing.assert_close(lse_g, lse, atol=1e-5, rtol=1e-6)
    scale = valid.float() / denom
    for r, s in enumerate(shards):
        local = tgt - r * (V // N)
        local = torch.where(valid & (local >= 0) & (local < V // N), local, torch.full_like(local, -1))
        d = ops.ce_bwd_sharded(s.contiguous(), local, lse, scale)
        torch.testing.assert_close(d.float(), dl.reshape(B * S, V).chunk(N, dim=-1)[r].float(), atol=1e-6, rtol=1e-2)


@pytest.mark.parametrize("state_dtype", [None, torch.float32])
def test_adamw_and_clip_kernels_match_oracle(state_dtype):
    """b200_adamw_step / b200_grad_norm / b200_grad_scale vs oracle/adamw_oracle.py on odd-sized tensors (vector body,
    scalar tail, a tensor spanning several chunks), 4 steps, fused clipping."""
    from oracle import adamw_oracle as O
    from transformers_b200.optim import B200AdamW, clip_grad_norm_

    shapes = [(7,), (64, 33), (3, 5, 8), (100003,), (4096, 64)]
    g = torch.Generator().manual_seed(0)
    host = [(torch.randn(s, generator=g) * 0.5).to(torch.bfloat16) for s in shapes]
    ps = [torch.nn.Parameter(h.clone().cuda()) for h in host]
    sdt = state_dtype or torch.bfloat16
    mine = [(h.clone(), torch.zeros(h.shape, dtype=sdt), torch.zeros(h.shape, dtype=sdt)) for h in host]
    opt = B200AdamW(ps, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, max_grad_norm=1.0, state_dtype=state_dtype)
    for step in range(1, 5):
        grads = [(torch.randn(s, generator=g) * 2).to(torch.bfloat16) for s in shapes]
        for p, gr in zip(ps, grads):
            p.grad = gr.cuda()
        total, coef = O.grad_norm_and_coef(grads, 1.0)
        opt.step()
        assert abs(float(opt.grad_norm) - total) < 1e-3 * total
        mine = [O.adamw_step(p, gr, m, v, step, 1e-2, 0.9, 0.95, 1e-8, 0.1, grad_scale=coef) for (p, m, v), gr in zip(mine, grads)]
    for (p, m, v), q in zip(mine, ps):
        torch.testing.assert_close(q.detach().cpu().float(), p.float(), atol=2e-3, rtol=1.6e-2)  # <= 2 bf16 ulps
        torch.testing.assert_close(opt.state[q]["exp_avg"].cpu().float(), m.float(), atol=2e-3, rtol=1.6e-2)
        torch.testing.assert_close(opt.state[q]["exp_avg_sq"].cpu().float(), v.float(), atol=2e-3, rtol=1.6e-2)
    grads = [(torch.randn(s, generator=g) * 2).to(torch.bfloat16) for s in shapes]
    for p, gr in zip(ps, grads):
        p.grad = gr.cuda()
    total, coef = O.grad_norm_and_coef(grads, 0.25)
    n = clip_grad_norm_(ps, 0.25)
    assert abs(float(n) - total) < 1e-3 * total
    for p, gr in zip(ps, grads):
        torch.testing.assert_close(p.grad.cpu().float(), (gr.float() * coef).to(torch.bfloat16).float(), atol=1e-6, rtol=8e-3)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_pull_reduce_kernel_on_local_buffers(world):
    """b200_pull_reduce_bf16 with all `world` "peer" buffers on this GPU (the addressing / summation logic; the NVLink side
    is covered by tests/cuda/tp_check.py with B200_TP_PEER=1): fp32 sum in rank order + residual, rounded once."""
    from transformers_b200 import ops

    torch.manual_seed(world)
    rows, cols = 96, 264  # 25344 elements per rank slice: several CTAs, not a multiple of the block size
    bufs = [torch.randn(world * rows, cols, device="cuda").to(torch.bfloat16) for _ in range(world)]
    res = torch.randn(rows, cols, device="cuda").to(torch.bfloat16)
    for rank in range(world):
        sl = slice(rank * rows, (rank + 1) * rows)
        want = res.float()
        for b in bufs:
            want = want + b[sl].float()
        out = torch.empty(rows, cols, device="cuda", dtype=torch.bfloat16)
        ops.pull_reduce([b.data_ptr() for b in bufs], rank * rows * cols, rows * cols, out, residual=res)
        assert torch.equal(out, want.to(torch.bfloat16))
        want0 = sum(b[sl].float() for b in bufs)
        ops.pull_reduce([b.data_ptr() for b in bufs], rank * rows * cols, rows * cols, out)
        assert torch.equal(out, want0.to(torch.bfloat16))


def test_mixtral_moe_backward_matches_oracle():
    """functional.MoEExpertsFn (forward + backward on the routing / gather / combine kernels and the expert GEMMs) vs autograd
    through the oracle's restatement of MixtralExperts.forward (models/mixtral/modeling_mixtral.py:69-93) in fp32, on the
    SAME routing decisions -- a whole-model comparison is ill-posed here: a router logit that differs in the last bf16 bit
    flips a top-k choice and with it that token's entire gradient."""
    from oracle import decoder_oracle as O
    from transformers_b200 import functional as Fn

    T, H, I, E, k = 200, 256, 512, 4, 2
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16)
    x, gate_up, down, w_gate = rnd(T, H), rnd(E, 2 * I, H, sc=0.05), rnd(E, H, I, sc=0.05), rnd(E, H, sc=0.2)
    tw, ti = O.moe_router(x, w_gate, k)
    ti = torch.where(ti == 3, torch.full_like(ti, 2), ti)  # one expert receives no token
    dout = rnd(T, H)
    ref_in = [t.detach().clone().float().requires_grad_(True) for t in (x, tw, gate_up, down)]
    ref = O.moe_experts(ref_in[0], ti, ref_in[1], ref_in[2], ref_in[3])
    ref.backward(dout.float())
    ours_in = [t.detach().to(torch.bfloat16).cuda().requires_grad_(True) for t in (x, tw, gate_up, down)]
    out = Fn.MoEExpertsFn.apply(ours_in[0], ti.cuda(), ours_in[1], ours_in[2], ours_in[3], False)
    out.backward(dout.cuda())
    rel = lambda a, b: ((a.float().cpu() - b).abs().max() / (b.abs().max() + 1e-8)).item()
    assert rel(out, ref.detach()) < 2e-2
    for name, a, b in zip(("x", "top_k_weights", "gate_up_proj", "down_proj"), ours_in, ref_in):
        assert a.grad is not None, name
        assert rel(a.grad, b.grad) < 3e-2, f"{name}: {rel(a.grad, b.grad)}"
    assert torch.count_nonzero(ours_in[2].grad[3]) == 0 and torch.count_nonzero(ours_in[3].grad[3]) == 0


@pytest.mark.parametrize("M,N,K", [(1, 4096, 4096), (1, 6144, 3584), (4, 28672, 4096), (3, 1000, 14336), (2, 128256, 4096), (1, 7, 264)])
def test_gemv_decode_rows_match_matmul(M, N, K):
    """b200_gemv_bf16 (direct C-ABI call) vs fp32 matmul on the same bf16 operands; then timing against the bytes of W."""
    import ctypes

    from transformers_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(M * 7 + N)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.b200_gemv_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), M, N, K, K, K, N, ctypes.c_void_p(st))
    assert rc == 0, _lib.last_error()
    ref = x.float() @ w.float().t()
    torch.testing.assert_close(y.float(), ref, atol=2e-2, rtol=1.6e-2)
    if N * K >= 4096 * 4096:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            lib.b200_gemv_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), M, N, K, K, K, N, ctypes.c_void_p(st))
        e0.record()
        for _ in range(10):
            lib.b200_gemv_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), M, N, K, K, K, N, ctypes.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        gbs = N * K * 2 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9
        print(f"gemv M={M} N={N} K={K}: {gbs:.0f} GB/s of weight stream")


@pytest.mark.parametrize("B,Hq,Hkv,D,ctx,window,softcap", [
    (1, 16, 8, 256, 8704, 0, 50.0),     # Gemma-2-9B full layer at the config-5 context
    (1, 16, 8, 256, 4095, 4096, 50.0),  # its sliding layers (cropped cache)
    (4, 32, 8, 128, 1000, 0, 0.0),      # Llama-3-8B heads
    (2, 8, 8, 64, 257, 0, 0.0), (3, 8, 1, 128, 300, 128, 0.0), (1, 4, 2, 128, 1, 0, 0.0), (2, 4, 4, 64, 255, 0, 30.0),
])
def test_decode_attention_matches_fp32_reference(B, Hq, Hkv, D, ctx, window, softcap):
    """b200_attn_decode (split-context kernel pair) on a KV-cache-shaped buffer [B, Hkv, capacity, D] vs an fp32 softmax
    reference and vs the prefill kernel (q_len == 1 through b200_attn_fwd), including left padding."""
    from transformers_b200 import ops

    torch.manual_seed(ctx + D)
    cap = ctx + 37
    kc = torch.randn(B, Hkv, cap, D, device="cuda").to(torch.bfloat16)
    vc = torch.randn(B, Hkv, cap, D, device="cuda").to(torch.bfloat16)
    q = torch.randn(B, 1, Hq, D, device="cuda").to(torch.bfloat16)
    k = kc[:, :, :ctx].transpose(1, 2)  # [B, ctx, Hkv, D] strided view, as the attention entry point passes it
    v = vc[:, :, :ctx].transpose(1, 2)
    kv_start = torch.tensor([0] + [5] * (B - 1), device="cuda", dtype=torch.int32) if ctx > 8 else None
    scale = D ** -0.5
    out_ref_kernel, lse_ref_kernel = ops.attn_fwd(q, k, v, scale=scale, causal=False, window=window, softcap=softcap,
                                                  kv_start=kv_start, decode_kernel=False)
    out, lse = ops.attn_fwd(q, k, v, scale=scale, causal=False, window=window, softcap=softcap, kv_start=kv_start)
    G = Hq // Hkv
    kf = k.float().repeat_interleave(G, dim=2)
    vf = v.float().repeat_interleave(G, dim=2)
    s = torch.einsum("bhd,bkhd->bhk", q[:, 0].float(), kf) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    idx = torch.arange(ctx, device="cuda")
    valid = torch.ones(B, ctx, dtype=torch.bool, device="cuda")
    if window:
        valid &= idx[None] >= ctx - window
    if kv_start is not None:
        valid &= idx[None] >= kv_start[:, None]
    s = s.masked_fill(~valid[:, None], float("-inf"))
    want = torch.einsum("bhk,bkhd->bhd", torch.softmax(s, -1), vf)
    torch.testing.assert_close(out[:, 0].float(), want, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(lse[..., 0], torch.logsumexp(s, -1), atol=2e-3, rtol=1e-3)
    torch.testing.assert_close(out.float(), out_ref_kernel.float(), atol=2e-2, rtol=2e-2)


def test_gemma_v1_forward_backward_matches_oracle():
    """Gemma (v1) through the plugin ((1+w) norm, GeGLU, scaled embeddings, tied head) vs the oracle, like the regular
    tests/test_model_gpu.py cases."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _hf import import_transformers
    from oracle import decoder_oracle as O

    tf = import_transformers()
    import transformers_b200

    transformers_b200.enable()
    cfg = tf.GemmaConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=1, head_dim=64, max_position_embeddings=512,
                         rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
    tf.set_seed(42)
    model = tf.GemmaForCausalLM._from_config(cfg, attn_implementation="b200", dtype=torch.bfloat16)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ocfg = O.config_from_hf(cfg)
    ids = torch.randint(1, cfg.vocab_size, (2, 200))
    p32 = {k: v.float().requires_grad_(True) for k, v in sd.items()}
    lo32, loss32, _ = O.model_forward(ids, p32, ocfg, labels=ids)
    loss32.backward()
    with torch.no_grad():
        lobf, _, _ = O.model_forward(ids, {k: v.clone() for k, v in sd.items()}, ocfg, labels=ids)
    model = model.cuda().train()
    transformers_b200.accelerate(model, fused_head_loss=False)
    assert type(model.model.layers[0].mlp).__name__ == "B200GemmaMLP"
    out = model(input_ids=ids.cuda(), labels=ids.cuda())
    out.loss.backward()
    got = out.logits.float().cpu()
    torch.testing.assert_close(got, lobf.float(), atol=3e-2, rtol=3e-2)
    assert (got - lo32.detach()).abs().max() <= 2 * (lobf.float() - lo32.detach()).abs().max() + 1e-2
    assert abs(out.loss.item() - loss32.item()) < 3e-2
    for n, p in model.named_parameters():
        g32 = p32[n].grad
        assert (p.grad.float().cpu() - g32).abs().max() / (g32.abs().max() + 1e-6) < 6e-2, n


@pytest.mark.parametrize("world,layout", [(2, "nt"), (4, "nt"), (4, "nn"), (8, "nt")])
def test_gemm_scatter_epilogue_on_local_buffers(world, layout):
    """b200_gemm_bf16_scatter with all destination slots on this GPU: every row block must land in its slot bit-identical to
    the plain GEMM's rows (same kernel, only the TMA-store target differs)."""
    from transformers_b200 import ops

    torch.manual_seed(world)
    rows, N, K = 256, 384, 512
    M = rows * world
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    b_mn = layout == "nn"
    bb = b.t().contiguous() if b_mn else b  # [K, N] storage for the dgrad layout
    want = ops.gemm(a, bb, b_mn=b_mn)
    for rank in range(world):
        slots = torch.full((world, rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm_scatter(a, bb, [slots[r].data_ptr() for r in range(world)], rank, b_mn=b_mn)
        torch.cuda.synchronize()
        assert torch.equal(slots.view(M, N), want), (world, rank)
