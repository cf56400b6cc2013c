"""Parity at BASELINE.json configs[1] SHAPES (Llama-3-8B width, S = 4096), not just through size-independent properties:
the CUDA path through the C-ABI / the plugin against the oracle's formulas evaluated in fp32 ON THE GPU (the oracle is
plain torch, so it runs unchanged under `torch.device("cuda")`; torch's fp32 matmul is the checker here, never the thing
measured).  Bars are the reference's own: 3e-2 vs flash-style kernels (tests/causal_lm_tester.py:398-446).

VERDICT r1 "parity is green only on toy shapes": every GEMM wave / TMEM-buffer parity, multi-tile attention at S = 4096 and
one full-width decoder layer + head + loss, forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from _hf import import_transformers  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402

BF = torch.bfloat16
T = 16384  # B * S of configs[1]


def _ops():
    from transformers_b200 import ops

    return ops


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


# name, (M, N, K), a_mn, b_mn  -- the GEMM launches of one Llama-3-8B step (bench.py kernels_ms: NT fwd, NN dgrad, TT wgrad)
GEMM_SHAPES = [
    ("qkv_fwd", (T, 6144, 4096), False, False),
    ("o_fwd", (T, 4096, 4096), False, False),
    ("gate_up_fwd", (T, 28672, 4096), False, False),
    ("down_fwd", (T, 4096, 14336), False, False),
    ("gate_up_dgrad", (T, 4096, 28672), False, True),
    ("down_dgrad", (T, 14336, 4096), False, True),
    ("gate_up_wgrad", (28672, 4096, T), True, True),
    ("down_wgrad", (4096, 14336, T), True, True),
    ("lm_head_fwd", (T, 128256, 4096), False, False),
]


@pytest.mark.parametrize("name,shape,a_mn,b_mn", GEMM_SHAPES, ids=[s[0] for s in GEMM_SHAPES])
def test_gemm_bench_shapes_every_tile(name, shape, a_mn, b_mn):
    """Rows sampled from EVERY 128-row block of the output (so from every 256 x 256 cluster tile, every persistent wave
    and both TMEM accumulator parities), all N columns, against fp32."""
    ops = _ops()
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(sum(map(ord, name)))
    A = (torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g) * 0.5).to(BF)
    Bm = (torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g) * 0.05).to(BF)
    out = ops.gemm(A, Bm, a_mn=a_mn, b_mn=b_mn)
    rows = (torch.arange(0, M, 128, device="cuda")[:, None] + torch.tensor([3, 77], device="cuda")[None, :]).reshape(-1)
    rows = rows[rows < M]
    a_rows = (A[:, rows].t() if a_mn else A[rows]).float()
    ref = a_rows @ (Bm.float() if b_mn else Bm.float().t())
    got = out[rows].float()
    torch.testing.assert_close(got, ref, atol=2e-2 * ref.abs().max().item(), rtol=1e-2)
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("B,window", [(1, 0), (2, 1024)])
def test_attention_s4096_fwd_bwd_vs_fp32(B, window):
    """Llama-3-8B attention geometry (32 / 8 heads, D = 128, S = 4096: 32 q tiles x up to 32 kv tiles per head) against
    eager_attention_forward's formulas in fp32 (models/llama/modeling_llama.py:191-213)."""
    ops = _ops()
    S, Hq, Hkv, D = 4096, 32, 8, 128
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, S, (Hq + 2 * Hkv) * D, device="cuda", generator=g).to(BF)
    q = qkv[..., : Hq * D].view(B, S, Hq, D)
    k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
    v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
    dout = torch.randn(B, S, Hq, D, device="cuda", generator=g).to(BF)
    scale = D**-0.5
    out, lse = ops.attn_fwd(q, k, v, scale=scale, causal=True, window=window)
    dqkv = torch.empty_like(qkv)
    dq = dqkv[..., : Hq * D].view(B, S, Hq, D)
    dk = dqkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
    dv = dqkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
    ops.attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, scale=scale, causal=True, window=window)
    with torch.device("cuda"):
        mask = O.eager_mask(1, S, S, torch.float32, sliding_window=window or None)
        for b in range(B):  # one batch row at a time: 2.1 GB of fp32 scores each
            qr, kr, vr = (t[b:b + 1].transpose(1, 2).float().detach().requires_grad_(True) for t in (q, k, v))
            ref, _ = O.eager_attention(qr, kr, vr, mask, scale, None)
            ref.backward(dout[b:b + 1].float())
            torch.testing.assert_close(out[b:b + 1].float(), ref.detach(), atol=3e-2, rtol=3e-2)
            assert _rel(dq[b:b + 1].transpose(1, 2), qr.grad) < 3e-2
            assert _rel(dk[b:b + 1].transpose(1, 2), kr.grad) < 3e-2
            assert _rel(dv[b:b + 1].transpose(1, 2), vr.grad) < 3e-2
            del qr, kr, vr, ref


def test_llama3_8b_full_width_layer_fwd_bwd_vs_fp32_oracle():
    """One full-width Llama-3-8B decoder layer + embedding + final norm + lm_head + loss (hidden 4096, 32 / 8 heads,
    intermediate 14336, vocab 128256), B = 1, S = 4096, built by the reference's own `_from_config` with the plugin on,
    forward + backward, against the oracle evaluated in fp32 on the same bf16 weights (3e-2: tests/causal_lm_tester.py:441)."""
    tf = import_transformers()
    import transformers_b200

    transformers_b200.enable()
    cfg = tf.LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=1,
                         num_attention_heads=32, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-5,
                         max_position_embeddings=8192, attention_bias=False, mlp_bias=False, tie_word_embeddings=False,
                         hidden_act="silu", rope_parameters={"rope_type": "default", "rope_theta": 500000.0}, use_cache=False)
    tf.set_seed(42)
    with torch.device("cuda"):
        model = tf.LlamaForCausalLM._from_config(cfg, attn_implementation="b200", dtype=BF)
    transformers_b200.accelerate(model, fused_head_loss=False)
    model.train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_((torch.randn_like(p.float()) * 0.1).to(p.dtype))
    torch.manual_seed(0)
    ids = torch.randint(0, cfg.vocab_size, (1, 4096), device="cuda")
    out = model(input_ids=ids, labels=ids)
    out.loss.backward()
    ocfg = O.config_from_hf(cfg)
    p32 = {k: v.detach().float().requires_grad_(True) for k, v in model.state_dict().items()}
    with torch.device("cuda"):
        logits_ref, loss_ref, _ = O.model_forward(ids, p32, ocfg, labels=ids)
    loss_ref.backward()
    with torch.device("cuda"), torch.no_grad():  # the reference's own arithmetic: the same eager formulas in bf16
        logits_bf, loss_bf, _ = O.model_forward(ids, {k: v.detach() for k, v in model.state_dict().items()}, ocfg, labels=ids)
    assert abs(out.loss.item() - loss_ref.item()) < 2e-2, (out.loss.item(), loss_ref.item())
    assert abs(out.loss.item() - loss_bf.item()) < 2e-2, (out.loss.item(), loss_bf.item())
    # at this width (K = 4096 / 14336 dot products of bf16-rounded activations) the reference's own bf16 eager logits sit
    # ~0.1 away from the fp32 result, so an absolute 3e-2 bar is not meaningful: require (1) the 3e-2 bar relative to the
    # logit range, (2) ours at least as close to fp32 as 2x the reference's bf16 path is (tests/test_model_gpu.py does the same)
    lo32 = logits_ref.detach()
    err_ours = (out.logits.float() - lo32).abs().max().item()
    err_ref = (logits_bf.float() - lo32).abs().max().item()
    assert err_ours / lo32.abs().max().item() < 3e-2, (err_ours, lo32.abs().max().item())
    assert err_ours <= 2 * err_ref + 1e-2, (err_ours, err_ref)
    del logits_bf
    for n, p in model.named_parameters():
        ref = p32[n].grad
        assert p.grad is not None, n
        rel = _rel(p.grad, ref)
        assert rel < 3e-2, f"grad {n}: rel err {rel}"
    # the same step through the chunked fused lm_head + loss (what accelerate() installs by default for training forwards):
    # no [T, V] logits, same loss, same gradients
    from transformers_b200.integration import _install_fused_head_loss

    loss_unfused = out.loss.item()
    del out, logits_ref, lo32
    model.zero_grad(set_to_none=True)
    _install_fused_head_loss(model)
    out2 = model(input_ids=ids, labels=ids)
    assert out2.logits.shape[-1] == 0
    out2.loss.backward()
    assert abs(out2.loss.item() - loss_unfused) < 2e-3, (out2.loss.item(), loss_unfused)
    for n, p in model.named_parameters():
        rel = _rel(p.grad, p32[n].grad)
        assert rel < 3e-2, f"fused head+loss, grad {n}: rel err {rel}"
