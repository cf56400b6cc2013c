"""Parity of every CUDA kernel, called through the C-ABI (transformers_b200.ops -> ctypes -> libb200.so), against the
oracle on the same seeded inputs; plus size-independent properties at BASELINE.json's full sizes.
Tolerances: bit-exact for integer indexing / pure copies; the reference's own bf16 bars (atol=rtol=1e-2 eager<->sdpa,
3e-2 vs flash-style kernels: tests/test_modeling_common.py:203-232, tests/causal_lm_tester.py:441) otherwise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import decoder_oracle as O  # noqa: E402

BF = torch.bfloat16


def _ops():
    from transformers_b200 import ops

    return ops


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max() + 1e-6)).item()


def _randn(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (384, 512, 192), (200, 264, 72), (1, 8, 8), (130, 16, 4096)])
def test_gemm_layouts_vs_oracle(M, N, K):
    ops = _ops()
    a, b = _randn(M, K, seed=1), _randn(N, K, seed=2)
    ref = torch.nn.functional.linear(a.float(), b.float())  # F.linear: the op behind nn.Linear (modeling_llama.py:169-171)
    for a_mn in (False, True):
        for b_mn in (False, True):
            if (a_mn and M % 8) or (b_mn and N % 8):
                continue  # an MN-major operand needs a 16-byte aligned row pitch
            A = (a.t().contiguous() if a_mn else a).cuda()
            B = (b.t().contiguous() if b_mn else b).cuda()
            out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn)
            assert _rel(out, ref) < 1e-2, (a_mn, b_mn)
    c0 = _randn(M, N, seed=3)
    out = ops.gemm(a.cuda(), b.cuda(), out=c0.cuda().clone(), accumulate=True)
    assert _rel(out, ref + c0.float()) < 1e-2


def test_gemm_linearity_full_size():
    """Llama-3-8B o_proj shape: (A1 + A2) W^T == A1 W^T + A2 W^T up to bf16 rounding; zero input -> exact zeros."""
    ops = _ops()
    T, N, K = 16384, 4096, 4096
    w = torch.randn(N, K, device="cuda").to(BF) * 0.02
    a1 = torch.randn(T, K, device="cuda").to(BF)
    a2 = torch.randn(T, K, device="cuda").to(BF)
    y12 = ops.gemm((a1.float() + a2.float()).to(BF), w)
    y1, y2 = ops.gemm(a1, w), ops.gemm(a2, w)
    assert _rel(y12, y1.float() + y2.float()) < 3e-2
    assert torch.count_nonzero(ops.gemm(torch.zeros_like(a1), w)) == 0
    ref = a1[:256].float() @ w.float().t()  # spot-check rows against fp32
    assert _rel(y1[:256], ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(384, 512, 192), (2048, 4096, 1024), (300, 264, 520)])
def test_gemm_accumulate_through_tma_reduce_add(M, N, K):
    """accumulate=True on the CTA-pair kernel: C += A B^T leaves the SM as a TMA reduce-add of the bf16 tile (no read-modify-
    write of C in the epilogue); repeated accumulation, M / N tails, and the MN-major (wgrad) operand layout."""
    ops = _ops()
    a, b = _randn(M, K, seed=11).cuda(), _randn(N, K, seed=12, scale=0.1).cuda()
    ref = a.float() @ b.float().t()
    c0 = _randn(M, N, seed=13).cuda()
    c = c0.clone()
    ops.gemm(a, b, out=c, accumulate=True)
    assert _rel(c, ref + c0.float()) < 1e-2
    ops.gemm(a, b, out=c, accumulate=True)
    assert _rel(c, 2 * ref + c0.float()) < 1.5e-2
    if M % 8 == 0 and N % 8 == 0:
        c = c0.clone()
        ops.gemm(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True, out=c, accumulate=True)
        assert _rel(c, ref + c0.float()) < 1e-2


@pytest.mark.parametrize("M,I,K,gelu", [(384, 256, 192, False), (1000, 384, 520, True), (2048, 1792, 4096, False)])
def test_gemm_glu_epilogue_is_bit_identical_to_gemm_then_glu(M, I, K, gelu):
    """b200_gemm_glu_bf16 (gate|up projection on the block-interleaved weight + gated activation in the epilogue) vs the two
    kernels it replaces, b200_gemm_bf16 on the concatenated weight + b200_glu_fwd: same accumulation order, same rounding
    points -> bit-identical projections and activations; then the interleaved-layout GLU backward vs the plain one."""
    ops = _ops()
    x = _randn(M, K, seed=14).cuda()
    wg, wu = _randn(I, K, seed=15, scale=0.05).cuda(), _randn(I, K, seed=16, scale=0.05).cuda()
    gu_plain = ops.gemm(x, torch.cat([wg, wu]))
    h_plain = ops.glu_fwd(gu_plain, gelu)
    gu_ilv, h = ops.gemm_glu(x, ops.interleave_gate_up(wg, wu), gelu)
    blocks = gu_ilv.view(M, I // 128, 2, 128)
    assert torch.equal(blocks[:, :, 0].reshape(M, I), gu_plain[:, :I]) and torch.equal(blocks[:, :, 1].reshape(M, I), gu_plain[:, I:])
    assert torch.equal(h, h_plain)
    assert torch.equal(ops.glu_fwd(gu_ilv, gelu, interleaved=True), h_plain)
    dh = _randn(M, I, seed=17).cuda()
    d_plain = ops.glu_bwd(dh, gu_plain, gelu)
    d_ilv = ops.glu_bwd(dh, gu_ilv, gelu, interleaved=True).view(M, I // 128, 2, 128)
    assert torch.equal(d_ilv[:, :, 0].reshape(M, I), d_plain[:, :I]) and torch.equal(d_ilv[:, :, 1].reshape(M, I), d_plain[:, I:])
    g, u = ops.deinterleave_gate_up(ops.interleave_gate_up(wg, wu))
    assert torch.equal(g, wg) and torch.equal(u, wu)


@pytest.mark.parametrize("counts", [[300, 0, 5, 256, 1027, 64, 9, 130], [512, 512, 512, 512], [0, 0, 77], [4096]])
@pytest.mark.parametrize("b_mn", [False, True])
def test_gemm_grouped_matches_per_expert_gemms(counts, b_mn):
    """b200_gemm_bf16_grouped (row ranges read from device memory, one launch for all experts) vs one b200_gemm_bf16 per
    expert on the same rows: same tile arithmetic -> bit-identical rows; ragged ranges (not multiples of 8, empty experts,
    ranges ending inside a 32-row strip), rows beyond the last range untouched.  Forward ([N, K]) and dgrad ([K, N]) layouts."""
    ops = _ops()
    E, N, K = len(counts), 320, 192
    M = sum(counts)
    off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device="cuda")
    a = _randn(M + 40, K, seed=21).cuda()  # 40 extra rows that belong to no expert
    b = _randn(E, K, N, seed=22, scale=0.1).cuda() if b_mn else _randn(E, N, K, seed=22, scale=0.1).cuda()
    out = torch.full((M + 40, N), float("nan"), device="cuda", dtype=BF)
    ops.gemm_grouped(a, b, off, b_mn=b_mn, out=out)
    lo = 0
    for e, c in enumerate(counts):
        if c:
            want = ops.gemm(a[lo:lo + c], b[e], b_mn=b_mn)
            if c > 128:  # same CTA-pair kernel family: identical tile arithmetic
                assert torch.equal(out[lo:lo + c], want), (e, c)
            else:        # the dispatcher takes the 1-CTA kernel for a single row tile
                torch.testing.assert_close(out[lo:lo + c].float(), want.float(), atol=1e-2, rtol=1e-2)
        lo += c
    assert torch.isnan(out[M:]).all()  # nothing written past the last range


def test_embedding_out_of_range_ids_are_reported():
    """Token ids outside the table raise (the reference's F.embedding device-asserts): not at the offending call -- that would
    cost a host sync per forward -- but at the next embedding call / embedding_check_now()."""
    ops = _ops()
    from transformers_b200 import B200Error

    ops.embedding_check_now()
    V, H = 100, 64
    w = _randn(V, H, seed=8).cuda()
    ok = ops.embedding_fwd(torch.tensor([[1, 2, 99]]).cuda(), w)
    assert torch.equal(ok[0, 2], w[99])
    ops.embedding_check_now()
    ops.embedding_fwd(torch.tensor([[3, V + 5, 7]]).cuda(), w)  # bad id: flagged on the device, reported later
    torch.cuda.synchronize()
    with pytest.raises(B200Error):
        ops.embedding_fwd(torch.tensor([[1]]).cuda(), w)
    ops.embedding_check_now()  # the report cleared the pending flags


def test_rope_table_bit_exact_vs_reference_ops():
    """b200_rope_table vs LlamaRotaryEmbedding.forward's six torch ops (models/llama/modeling_llama.py:113-127) run on the same
    device: fp32 outer product, cat, cos / sin, scale, cast -- bit for bit; and within one bf16 ulp of the CPU oracle."""
    ops = _ops()
    cfg = O.DecoderConfig(vocab_size=8, hidden_size=64, intermediate_size=64, num_hidden_layers=1, num_attention_heads=1,
                          num_key_value_heads=1, head_dim=128, rope_theta=500000.0)
    inv_freq = O.rope_inv_freq(cfg)
    pos = torch.stack([torch.arange(4096), torch.arange(4096) + 3000])  # two rows, large angles included
    for scaling in (1.0, 0.8333):
        cos, sin = ops.rope_table(inv_freq.cuda(), pos.cuda(), scaling)
        want_cos, want_sin = O.rope_tables(inv_freq.cuda(), pos.cuda(), BF, attention_scaling=scaling)
        # same arithmetic in the same order (cosf / sinf are what torch's CUDA cos / sin call); allow a handful of elements whose
        # fp32 value sits on a bf16 rounding boundary in case the two toolkits' libdevice differ in the last fp32 bit
        for got, want in ((cos, want_cos), (sin, want_sin)):
            bad = (got != want).sum().item()
            assert bad <= 1e-4 * got.numel(), f"{bad} of {got.numel()} table entries differ from the reference's ops"
            assert (got.float() - want.float()).abs().max() <= 2 ** -7
        cpu_cos, cpu_sin = O.rope_tables(inv_freq, pos, BF, attention_scaling=scaling)
        assert (cos.cpu().float() - cpu_cos.float()).abs().max() <= 2 ** -7 and (sin.cpu().float() - cpu_sin.float()).abs().max() <= 2 ** -7


def test_embedding_bit_exact_and_scatter():
    ops = _ops()
    V, H = 1000, 256
    w = _randn(V, H, seed=4)
    ids = torch.randint(0, V, (3, 50), generator=torch.Generator().manual_seed(5))
    out = ops.embedding_fwd(ids.cuda(), w.cuda())
    assert torch.equal(out.cpu(), O.embedding(ids, w))  # integer indexing: bit-exact
    out_s = ops.embedding_fwd(ids.cuda(), w.cuda(), scale=float(torch.tensor(H**0.5).to(BF)))
    assert torch.equal(out_s.cpu(), O.embedding(ids, w, scale=H**0.5))
    dout = _randn(3, 50, H, seed=6)
    wr = w.float().requires_grad_(True)
    O.embedding(ids, wr, padding_idx=7).backward(dout.float())
    dw = ops.embedding_bwd(ids.cuda(), dout.cuda(), V, 7)
    assert _rel(dw, wr.grad) < 2e-2
    with pytest.raises(Exception):
        ops.embedding_fwd(ids.cuda(), w.cuda()[:, :7])  # H % 8 != 0


@pytest.mark.parametrize("gemma", [False, True])
@pytest.mark.parametrize("T,H", [(37, 64), (300, 4096), (16, 3584), (8, 8192)])
def test_rmsnorm_fwd_bwd_vs_oracle(T, H, gemma):
    ops = _ops()
    x = _randn(T, H, seed=7)
    w = (_randn(H, seed=8).float() * 0.1 + (0.0 if gemma else 1.0)).to(BF)
    eps = 1e-6 if gemma else 1e-5
    ref = O.rms_norm(x, w, eps, gemma)
    y, rstd, _ = ops.rmsnorm_fwd(x.cuda(), w.cuda(), eps, gemma)
    torch.testing.assert_close(y.cpu().float(), ref.float(), atol=1e-2, rtol=1e-2)
    assert (y.cpu() != ref).float().mean() < 1e-3  # only rare 1-ulp flips from the reduction order
    # fused residual variant
    r = _randn(T, H, seed=9)
    y2, _, res = ops.rmsnorm_fwd(x.cuda(), w.cuda(), eps, gemma, residual=r.cuda())
    assert torch.equal(res.cpu(), x + r)
    torch.testing.assert_close(y2.cpu().float(), O.rms_norm(x + r, w, eps, gemma).float(), atol=1e-2, rtol=1e-2)
    # backward vs fp32 autograd of the oracle
    dy = _randn(T, H, seed=10)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    O.rms_norm(xr, wr, eps, gemma).backward(dy.float())
    dx, dw = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd, gemma)
    assert _rel(dx, xr.grad) < 2e-2 and _rel(dw, wr.grad) < 2e-2


def test_rope_bit_exact_and_inverse():
    ops = _ops()
    B, S, Hq, Hkv, D = 2, 40, 4, 2, 64
    cfg = O.DecoderConfig(vocab_size=8, hidden_size=8, intermediate_size=8, num_hidden_layers=1, num_attention_heads=Hq,
                          num_key_value_heads=Hkv, head_dim=D)
    cos, sin = O.rope_tables(O.rope_inv_freq(cfg), torch.arange(S)[None], BF)
    qkv = _randn(B, S, (Hq + 2 * Hkv) * D, seed=11)
    q = qkv[..., : Hq * D].view(B, S, Hq, D).transpose(1, 2)
    k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D).transpose(1, 2)
    qe, ke = O.apply_rope(q, k, cos, sin)
    buf = qkv.cuda().clone()
    ops.rope_(buf, cos.cuda(), sin.cuda(), Hq + Hkv, D)
    got = buf.cpu()
    assert torch.equal(got[..., : Hq * D].view(B, S, Hq, D).transpose(1, 2), qe)  # elementwise bf16: bit-exact
    assert torch.equal(got[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D).transpose(1, 2), ke)
    assert torch.equal(got[..., (Hq + Hkv) * D:], qkv[..., (Hq + Hkv) * D:])  # v untouched
    # backward is the transpose: <rope(x), y> == <x, rope^T(y)>
    y = _randn(B, S, (Hq + 2 * Hkv) * D, seed=12)
    yb = y.cuda().clone()
    ops.rope_(yb, cos.cuda(), sin.cuda(), Hq + Hkv, D, backward=True)
    lhs = (got.float() * y.float()).sum()
    rhs = (qkv.float() * yb.cpu().float()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 2e-2


@pytest.mark.parametrize("act", ["silu", "gelu_pytorch_tanh"])
def test_glu_vs_oracle(act):
    ops = _ops()
    T, I = 50, 176
    gu = _randn(T, 2 * I, seed=13, scale=2.0)
    g, u = gu[:, :I], gu[:, I:]
    ref = O.act_fn(g, act) * u
    out = ops.glu_fwd(gu.cuda(), act != "silu")
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=1e-2, rtol=1e-2)
    dh = _randn(T, I, seed=14)
    gr, ur = g.float().requires_grad_(True), u.float().requires_grad_(True)
    (O.act_fn(gr, act) * ur).backward(dh.float())
    dgu = ops.glu_bwd(dh.cuda(), gu.cuda(), act != "silu").cpu()
    assert _rel(dgu[:, :I], gr.grad) < 2e-2 and _rel(dgu[:, I:], ur.grad) < 2e-2


ATTN_CASES = [
    dict(B=2, Sq=128, Skv=128, Hq=2, Hkv=1, D=128),
    dict(B=2, Sq=200, Skv=200, Hq=4, Hkv=2, D=64),
    dict(B=1, Sq=320, Skv=320, Hq=4, Hkv=2, D=128, window=100),
    dict(B=2, Sq=256, Skv=256, Hq=2, Hkv=2, D=128, softcap=20.0),
    dict(B=2, Sq=1, Skv=77, Hq=4, Hkv=2, D=128),            # decode: q_len 1 over the cache
    dict(B=2, Sq=60, Skv=260, Hq=4, Hkv=1, D=64),           # chunked prefill (bottom-right aligned causal)
    dict(B=2, Sq=150, Skv=150, Hq=2, Hkv=1, D=128, pad=True),
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_attention_fwd_bwd_vs_oracle(case):
    ops = _ops()
    B, Sq, Skv, Hq, Hkv, D = (case[k] for k in ("B", "Sq", "Skv", "Hq", "Hkv", "D"))
    window, softcap, pad = case.get("window", 0), case.get("softcap", 0.0), case.get("pad", False)
    q, k, v = _randn(B, Hq, Sq, D, seed=20), _randn(B, Hkv, Skv, D, seed=21), _randn(B, Hkv, Skv, D, seed=22)
    scale = D**-0.5
    pm = None
    if pad:
        pm = torch.ones(B, Skv, dtype=torch.bool)
        pm[0, -9:] = False
        pm[1, :4] = False
    causal = Sq > 1
    mask = O.eager_mask(B, Sq, Skv, torch.float32, q_offset=Skv - Sq, sliding_window=window or None, padding_mask=pm)
    if not causal:
        mask = None if pm is None else torch.where(pm[:, None, None, :], 0.0, torch.finfo(torch.float32).min)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, _ = O.eager_attention(qr, kr, vr, mask, scale, softcap or None)  # fp32 oracle: [B,Sq,Hq,D]
    dout = _randn(B, Sq, Hq, D, seed=23)
    ref.backward(dout.float())
    ks = ke = None
    if pad:
        from transformers_b200.modules import mask_to_kv_ranges

        ks, ke = mask_to_kv_ranges(pm.cuda())
    qc, kc, vc = (t.cuda().transpose(1, 2) for t in (q, k, v))  # strided [B,S,h,D] views
    out, lse = ops.attn_fwd(qc, kc, vc, scale=scale, causal=causal, window=window, softcap=softcap, kv_start=ks, kv_end=ke)
    valid = torch.ones(B, Sq, dtype=torch.bool)
    if pad and Sq == Skv:
        valid = pm  # fully padded query rows are garbage in the reference too
    torch.testing.assert_close(out.cpu().float()[valid], ref.detach()[valid], atol=3e-2, rtol=3e-2)
    dq, dk, dv = (torch.empty(t.shape, device="cuda", dtype=BF) for t in (qc, kc, vc))
    doc = dout.cuda()
    if pad and Sq == Skv:
        doc = doc * pm.cuda()[:, :, None, None]
        for t in (qr, kr, vr):
            t.grad = None
        ref2, _ = O.eager_attention(qr, kr, vr, mask, scale, softcap or None)
        ref2.backward(dout.float() * pm[:, :, None, None])
    ops.attn_bwd(qc, kc, vc, out, doc, lse, dq, dk, dv, scale=scale, causal=causal, window=window, softcap=softcap,
                 kv_start=ks, kv_end=ke)
    assert _rel(dq.transpose(1, 2), qr.grad) < 3e-2
    assert _rel(dk.transpose(1, 2), kr.grad) < 3e-2
    assert _rel(dv.transpose(1, 2), vr.grad) < 3e-2


def test_attention_properties_full_size():
    """Llama-3-8B attention shape (B=4, S=4096, 32/8 heads, D=128), checked through size-independent properties:
    v == 1 -> output exactly 1 (softmax rows sum to one); causality: changing future keys leaves earlier rows bit-identical;
    lse of a constant-score row == log(row length)."""
    ops = _ops()
    B, S, Hq, Hkv, D = 4, 4096, 32, 8, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, S, (Hq + 2 * Hkv) * D, device="cuda", generator=g).to(BF)
    q = qkv[..., : Hq * D].view(B, S, Hq, D)
    k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
    v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
    v.fill_(1.0)
    out, lse = ops.attn_fwd(q, k, v, scale=D**-0.5, causal=True)
    assert (out.float() - 1.0).abs().max() < 1e-2
    out_a = out.clone()
    k[:, 3000:].normal_(generator=g)  # perturb the "future"
    out_b, lse_b = ops.attn_fwd(q, k, v, scale=D**-0.5, causal=True)
    assert torch.equal(lse[..., :3000], lse_b[..., :3000])
    assert torch.equal(out_a[:, :3000], out_b[:, :3000])
    q.zero_()
    _, lse0 = ops.attn_fwd(q, k, v, scale=D**-0.5, causal=True)
    expect = torch.log(torch.arange(1, S + 1, device="cuda", dtype=torch.float32))
    torch.testing.assert_close(lse0[0, 0, :S], expect, atol=1e-3, rtol=1e-4)


def test_causal_lm_loss_vs_oracle():
    ops = _ops()
    B, S, V = 2, 33, 1000
    logits = _randn(B, S, V, seed=30, scale=3.0)
    labels = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(31))
    labels[0, 5:9] = -100
    lr = logits.float().requires_grad_(True)
    ref = O.causal_lm_loss(lr.to(BF), labels)
    ref.backward()
    loss, lse, denom = ops.ce_fwd(logits.cuda(), labels.cuda())
    torch.testing.assert_close(loss.cpu(), ref.detach(), atol=1e-4, rtol=1e-4)
    dl = ops.ce_bwd(logits.cuda(), labels.cuda(), lse, torch.ones((), device="cuda"), denom)
    assert _rel(dl, lr.grad) < 2e-2
    # sum / num_items_in_batch variant (loss/loss_utils.py:40-44)
    loss2, _, _ = ops.ce_fwd(logits.cuda(), labels.cuda(), num_items=17.0)
    torch.testing.assert_close(loss2.cpu(), O.causal_lm_loss(logits, labels, num_items_in_batch=17.0), atol=1e-3, rtol=1e-4)


def test_kv_append_bit_exact_vs_reference_cat():
    """DynamicLayer.update (cache_utils.py:127-146) semantics: returned K/V == torch.cat of everything appended so far
    (byte copy -> bit-exact), through prefill + decode steps and one capacity growth."""
    from transformers_b200.cache import layer_class

    B, H, D = 2, 2, 64
    layer = layer_class()()
    layer.min_capacity = 8
    ref_k = ref_v = None
    g = torch.Generator().manual_seed(40)
    for q in (5, 1, 1, 7, 1, 30):
        # new states arrive as transposed views of [B, S, H, D] storage, exactly as the attention module produces them
        k = torch.randn(B, q, H, D, generator=g).to(BF).cuda().transpose(1, 2)
        v = torch.randn(B, q, H, D, generator=g).to(BF).cuda().transpose(1, 2)
        ks, vs = layer.update(k, v)
        ref_k = k if ref_k is None else torch.cat([ref_k, k], dim=-2)
        ref_v = v if ref_v is None else torch.cat([ref_v, v], dim=-2)
        assert ks.shape == ref_k.shape and torch.equal(ks, ref_k) and torch.equal(vs, ref_v)
        assert layer.get_seq_length() == ref_k.shape[-2]
    layer.crop(-10)
    assert layer.get_seq_length() == ref_k.shape[-2] - 10 and torch.equal(layer.keys, ref_k[:, :, :-10])
    layer.reset()
    assert layer.get_seq_length() == 0


def test_attention_head_dim_256_forward_vs_oracle():
    """Gemma-2-9B geometry (head_dim 256, softcap, sliding window): forward only (configs[4] is generate())."""
    ops = _ops()
    B, Sq, Skv, Hq, Hkv, D = 1, 300, 300, 4, 2, 256
    q, k, v = _randn(B, Hq, Sq, D, seed=50), _randn(B, Hkv, Skv, D, seed=51), _randn(B, Hkv, Skv, D, seed=52)
    for window, softcap in ((0, 0.0), (128, 50.0)):
        mask = O.eager_mask(B, Sq, Skv, torch.float32, sliding_window=window or None)
        ref, _ = O.eager_attention(q.float(), k.float(), v.float(), mask, 256**-0.5, softcap or None)
        out, lse = ops.attn_fwd(*(t.cuda().transpose(1, 2) for t in (q, k, v)), scale=256**-0.5, causal=True, window=window,
                                softcap=softcap)
        torch.testing.assert_close(out.cpu().float(), ref, atol=3e-2, rtol=3e-2)
    # decode step over a long cache
    q1 = _randn(B, Hq, 1, D, seed=53)
    ref, _ = O.eager_attention(q1.float(), k.float(), v.float(), None, 256**-0.5, None)
    out, _ = ops.attn_fwd(q1.cuda().transpose(1, 2), k.cuda().transpose(1, 2), v.cuda().transpose(1, 2), scale=256**-0.5, causal=False)
    torch.testing.assert_close(out.cpu().float(), ref, atol=3e-2, rtol=3e-2)


def test_moe_experts_vs_oracle():
    """Mixtral experts path (route -> gather -> per-expert GEMMs -> GLU -> GEMMs -> weighted combine) vs the oracle's
    restatement of MixtralExperts.forward; also an expert that receives no token and top-1 routing."""
    ops = _ops()
    T, H, I, E = 200, 128, 256, 8
    x = _randn(T, H, seed=60)
    gate_up, down = _randn(E, 2 * I, H, seed=61, scale=0.05), _randn(E, H, I, seed=62, scale=0.05)
    w_gate = _randn(E, H, seed=63, scale=0.2)
    for topk in (2, 1):
        tw, ti = O.moe_router(x, w_gate, topk)
        ti = torch.where(ti == 5, torch.full_like(ti, 4), ti)  # expert 5 gets nothing
        ref = O.moe_experts(x.float(), ti, tw, gate_up.float(), down.float())
        ref_bf = O.moe_experts(x, ti, tw, gate_up, down)
        out = ops.moe_experts_forward(x.cuda(), ti.cuda(), tw.cuda(), gate_up.cuda(), down.cuda())
        torch.testing.assert_close(out.cpu().float(), ref_bf.float(), atol=2e-2, rtol=3e-2)
        assert _rel(out, ref) < 2e-2
