"""The kernel path's HOST logic on CPU: the real ``functional.py`` / ``modules.py`` / ``integration.py`` code runs, with the
C-ABI calls replaced by the torch stand-ins of ``tests/_fake_ops.py``.  What is under test is everything between the
reference's module boundary and the kernel launches: fused-weight packing, the strided q/k/v views handed to attention, RoPE
on the packed buffer, gradient routing to the individual q/k/v / gate/up parameters, mask -> kv range conversion, the loss
hook.  (The kernels themselves are tested against the oracle in tests/test_kernels_gpu.py / test_model_gpu.py.)

Reference for the numbers: the stock eager forward/backward of the same model in the same process (fp32)."""
import copy

import pytest
import torch

import _fake_ops
from _hf import import_transformers

transformers = import_transformers()
import transformers_b200  # noqa: E402


@pytest.fixture(autouse=True)
def _fakes(monkeypatch):
    transformers_b200.enable()
    _fake_ops.install(monkeypatch.setattr)
    yield


def _llama_cfg(**kw):
    base = dict(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, rms_norm_eps=1e-5, max_position_embeddings=128,
                rope_parameters={"rope_type": "default", "rope_theta": 500000.0})
    base.update(kw)
    return transformers.LlamaConfig(**base)


def _pair(cls, cfg, fused_head_loss=False):
    """(stock eager model, accelerated copy on the faked kernel path) with identical weights.  ``fused_head_loss`` is off
    unless a test is about it: most tests compare the logits of training-mode forwards, which that path does not produce."""
    from transformers.monkey_patching import clear_patch_mapping

    import transformers_b200.integration as integ

    clear_patch_mapping()
    try:
        transformers.set_seed(0)
        ref = cls._from_config(cfg, attn_implementation="eager", dtype=torch.float32)
    finally:
        integ._enabled = False
        transformers_b200.enable()
    ours = copy.deepcopy(ref)
    transformers_b200.accelerate(ours, fused_head_loss=fused_head_loss)
    return ref, ours


def _compare(ref, ours, ids, attention_mask=None, atol=2e-5):
    kw = dict(input_ids=ids, labels=ids, use_cache=False)
    if attention_mask is not None:
        kw["attention_mask"] = attention_mask
        keep = attention_mask.bool()
        keep[:, 1:] &= attention_mask.bool()[:, :-1]  # a target predicted FROM a pad position is garbage in both paths
        kw["labels"] = ids.masked_fill(~keep, -100)
    a = ref(**kw)
    a.loss.backward()
    _fake_ops.CALLS.clear()
    b = ours(**kw)
    b.loss.backward()
    valid = slice(None) if attention_mask is None else attention_mask.bool()
    torch.testing.assert_close(b.logits[valid], a.logits[valid], atol=atol, rtol=1e-4)
    torch.testing.assert_close(b.loss, a.loss, atol=1e-5, rtol=1e-5)
    ga = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        assert p.grad is not None, n
        torch.testing.assert_close(p.grad, ga[n].grad, atol=atol, rtol=1e-3, msg=lambda m, n=n: f"{n}: {m}")


def test_llama_fwd_bwd_through_kernel_path_host_logic():
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    torch.manual_seed(1)
    ids = torch.randint(0, 160, (2, 24))
    _compare(ref, ours, ids)
    names = [c[0] for c in _fake_ops.CALLS]
    # per layer: qkv GEMM + o GEMM + gate|up GEMM + down GEMM forward (4), one attention, one GLU; + lm_head
    assert names.count("attn_fwd") == 2 and names.count("attn_bwd") == 2
    assert names.count("glu_fwd") == 2 and names.count("rope_") == 4
    assert names.count("gemm") == 3 * (4 * 2 + 1)  # fwd + dgrad + wgrad for each fused linear
    assert names.count("ce_fwd") == 1 and names.count("embedding_bwd") == 1


def test_llama_left_and_right_padding():
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    torch.manual_seed(2)
    ids = torch.randint(0, 160, (3, 20))
    am = torch.ones(3, 20, dtype=torch.long)
    am[0, :5] = 0   # left padded
    am[1, -4:] = 0  # right padded
    _compare(ref, ours, ids, am)


def test_mistral_sliding_window():
    cfg = transformers.MistralConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                     num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=8,
                                     max_position_embeddings=128)
    ref, ours = _pair(transformers.MistralForCausalLM, cfg)
    torch.manual_seed(3)
    _compare(ref, ours, torch.randint(0, 160, (2, 32)))


def test_gemma2_softcap_scaled_embedding_and_alternating_window():
    cfg = transformers.Gemma2Config(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                    num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=8,
                                    query_pre_attn_scalar=16, max_position_embeddings=128, pad_token_id=0)
    ref, ours = _pair(transformers.Gemma2ForCausalLM, cfg)
    torch.manual_seed(4)
    ids = torch.randint(1, 160, (2, 24))
    _compare(ref, ours, ids, atol=5e-5)
    assert ours.loss_function is not transformers_b200.integration.b200_causal_lm_loss  # final softcapping: stock loss stays


def test_kv_cache_path_matches_full_forward():
    """generate()-style prefill + one-token steps through QKVRopeFn + Cache.update + the registry entry."""
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    ours.eval()
    ref.eval()
    torch.manual_seed(5)
    ids = torch.randint(0, 160, (2, 12))
    with torch.no_grad():
        full = ref(input_ids=ids).logits
        cache = transformers.DynamicCache(config=ours.config)
        out = ours(input_ids=ids[:, :8], past_key_values=cache, use_cache=True)
        steps = [out.logits]
        for t in range(8, 12):
            out = ours(input_ids=ids[:, t:t + 1], past_key_values=cache, use_cache=True)
            steps.append(out.logits)
    torch.testing.assert_close(torch.cat(steps, 1), full, atol=2e-5, rtol=1e-4)


def test_gradient_checkpointing_replays_functions():
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    ours.train()
    ref.train()
    ours.gradient_checkpointing_enable()
    torch.manual_seed(6)
    _compare(ref, ours, torch.randint(0, 160, (2, 16)))


def test_packed_weights_are_parameter_views_and_survive_optimizer_steps():
    from transformers_b200.modules import fused_weight

    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    names_before = {k: v.shape for k, v in ours.state_dict().items()}
    assert transformers_b200.pack_weights(ours) == 2 * 2  # (qkv, gate|up) x 2 layers
    assert {k: v.shape for k, v in ours.state_dict().items()} == names_before
    att = ours.model.layers[0].self_attn
    ws = [att.q_proj.weight, att.k_proj.weight, att.v_proj.weight]
    buf = fused_weight(att, "qkv", ws)
    assert buf.data_ptr() == ws[0].data_ptr() and buf.shape[0] == sum(w.shape[0] for w in ws)
    assert ws[1].data_ptr() == buf.data_ptr() + ws[0].numel() * 4  # adjacent row views, no copy
    torch.manual_seed(7)
    ids = torch.randint(0, 160, (2, 16))
    _compare(ref, ours, ids)
    # an optimizer step updates the parameters in place -> the packed operand follows without re-concatenation
    opt_a = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt_b = torch.optim.SGD(ours.parameters(), lr=0.1)
    opt_a.step()
    opt_b.step()
    assert fused_weight(att, "qkv", ws) is buf and torch.equal(buf[: ws[0].shape[0]], ws[0])
    ref.zero_grad(set_to_none=True)
    ours.zero_grad(set_to_none=True)
    _compare(ref, ours, ids)
    # re-allocating the parameters (dtype change) breaks the views: silently back to the cached concatenation
    ours.to(torch.float64)
    ws = [att.q_proj.weight, att.k_proj.weight, att.v_proj.weight]
    again = fused_weight(att, "qkv", ws)
    assert again.dtype == torch.float64 and torch.equal(again, torch.cat([w.detach() for w in ws]))
    assert "qkv" not in att.__dict__["_b200_packed"]
    transformers_b200.unpack_weights(ours)
    assert "_b200_packed" not in att.__dict__


def test_inplace_kv_cache_grows_crops_and_reorders_like_the_reference(monkeypatch):
    """B200DynamicLayer on the kernel path (append kernel faked): prefill + decode through make_cache() must reproduce the
    full forward, across a buffer re-allocation; crop / batch re-ordering keep the reference layer's semantics."""
    from transformers_b200.cache import layer_class, make_cache

    monkeypatch.setattr(layer_class(), "min_capacity", 4)  # force a geometric re-allocation inside the test
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    ours.eval()
    ref.eval()
    torch.manual_seed(8)
    ids = torch.randint(0, 160, (2, 14))
    with torch.no_grad():
        full = ref(input_ids=ids).logits
        cache = make_cache(ours.config)
        out = ours(input_ids=ids[:, :6], past_key_values=cache, use_cache=True)
        steps = [out.logits]
        for t in range(6, 14):
            steps.append(ours(input_ids=ids[:, t:t + 1], past_key_values=cache, use_cache=True).logits)
    torch.testing.assert_close(torch.cat(steps, 1), full, atol=2e-5, rtol=1e-4)
    layer = cache.layers[0]
    assert type(layer) is layer_class() and layer.get_seq_length() == 14 and layer._buf_k.shape[2] == 16
    assert [c[0] for c in _fake_ops.CALLS].count("kv_append") > 2 * 9  # appends + the copies of two re-allocations
    keys_before = layer.keys.clone()
    layer.crop(10)
    assert layer.get_seq_length() == 10 and torch.equal(layer.keys, keys_before[:, :, :10])
    layer.reorder_cache(torch.tensor([1, 0]))
    assert torch.equal(layer.keys, keys_before[[1, 0]][:, :, :10])
    layer.reset()
    assert layer.get_seq_length() == 0


def test_loss_hook_honours_num_items_in_batch_and_ignore_index():
    """Trainer passes num_items_in_batch under gradient accumulation (loss/loss_utils.py:48-70): sum / num_items."""
    from transformers.loss.loss_utils import ForCausalLMLoss

    from transformers_b200.integration import b200_causal_lm_loss

    torch.manual_seed(9)
    logits = torch.randn(2, 10, 50, requires_grad=True)
    labels = torch.randint(0, 50, (2, 10))
    labels[0, 3:6] = -100
    mine = logits.detach().clone().requires_grad_(True)
    a = ForCausalLMLoss(logits, labels, vocab_size=50, num_items_in_batch=torch.tensor(40))
    b = b200_causal_lm_loss(mine, labels, vocab_size=50, num_items_in_batch=torch.tensor(40))
    torch.testing.assert_close(b, a, atol=1e-6, rtol=1e-6)
    a.backward()
    b.backward()
    torch.testing.assert_close(mine.grad, logits.grad, atol=1e-7, rtol=1e-5)


def test_inplace_sliding_window_layer_matches_the_reference_layer():
    """B200SlidingWindowLayer vs DynamicSlidingWindowLayer (cache_utils.py:203-262) on random update sequences: prefill longer
    and shorter than the window, single-token decode across several buffer wrap-arounds, record_past + crop, beam re-order."""
    from transformers.cache_utils import DynamicSlidingWindowLayer

    from transformers_b200.cache import sliding_layer_class

    g = torch.Generator().manual_seed(11)
    for window, first in ((5, 3), (5, 12), (8, 8), (2, 1)):
        ref, ours = DynamicSlidingWindowLayer(sliding_window=window), sliding_layer_class()(sliding_window=window)
        for step, q in enumerate([first] + [1] * 23 + [3, 1, 1, 4, 4, 2, 4, 1]):
            k, v = torch.randn(2, 2, q, 4, generator=g), torch.randn(2, 2, q, 4, generator=g)
            rk, rv = ref.update(k, v)
            ok, ov = ours.update(k, v)
            assert torch.equal(ok, rk) and torch.equal(ov, rv), (window, step)
            assert torch.equal(ours.keys, ref.keys) and ours.cumulative_length == ref.cumulative_length
            assert ours.get_mask_sizes(1) == ref.get_mask_sizes(1) and ours.get_seq_length() == ref.get_seq_length()
            if step == 10:  # the base class re-orders the batch (beam search): detected and rebuilt on the next update
                idx = torch.tensor([1, 0])
                ref.reorder_cache(idx)
                ours.reorder_cache(idx)
        assert ours._buf_k.shape[2] <= 2 * (window - 1) + max(first, 4) + 1  # bounded by window + longest update, not by the context
    appended = [c for c in _fake_ops.CALLS if c[0] == "kv_append"]
    assert len(appended) > 4 * 32  # every update is one append (+ occasional window moves), never a torch.cat of the window


def test_make_cache_inplace_sliding_generates_like_the_reference_cache():
    cfg = transformers.Gemma2Config(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                    num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=6,
                                    query_pre_attn_scalar=16, max_position_embeddings=128, pad_token_id=0)
    from transformers_b200.cache import layer_class, make_cache, sliding_layer_class

    ref, ours = _pair(transformers.Gemma2ForCausalLM, cfg)
    ref.eval()
    ours.eval()
    torch.manual_seed(12)
    ids = torch.randint(1, 160, (2, 20))
    with torch.no_grad():
        full = ref(input_ids=ids).logits
        cache = make_cache(ours.config, inplace_sliding=True)
        assert type(cache.layers[0]) is sliding_layer_class() and type(cache.layers[1]) is layer_class()
        outs = [ours(input_ids=ids[:, :9], past_key_values=cache, use_cache=True).logits]
        for t in range(9, 20):
            outs.append(ours(input_ids=ids[:, t:t + 1], past_key_values=cache, use_cache=True).logits)
    torch.testing.assert_close(torch.cat(outs, 1), full, atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("inter", [96, 128])
def test_mixtral_moe_forward_and_backward_through_the_experts_registry(inter):
    """config 4 family: the b200 experts entry (ExpertsInterface) under autograd = functional.MoEExpertsFn; gradients of the
    hidden states, the router (through top_k_weights) and the stacked expert weights vs the stock eager experts.  Widths in
    whole 64-column chunks (inter = 128) take the grouped GEMM (expert row ranges stay on the device: one launch per
    projection, forward and dgrad); odd widths (inter = 96) the per-expert launches."""
    cfg = transformers.MixtralConfig(vocab_size=160, hidden_size=64, intermediate_size=inter, num_hidden_layers=2,
                                     num_attention_heads=4, num_key_value_heads=2, head_dim=16, num_local_experts=4,
                                     num_experts_per_tok=2, max_position_embeddings=128, sliding_window=None,
                                     router_jitter_noise=0.0, output_router_logits=False)
    ref, ours = _pair(transformers.MixtralForCausalLM, cfg)
    assert ours.config._experts_implementation == "b200"
    torch.manual_seed(13)
    ids = torch.randint(0, 160, (2, 24))
    _compare(ref, ours, ids, atol=5e-5)
    names = [c[0] for c in _fake_ops.CALLS]
    assert names.count("moe_route") == 2 and names.count("moe_combine") == 2 * 2  # fwd un-permute + bwd dX per layer
    assert names.count("gemm_grouped") == (2 * 4 if inter == 128 else 0)  # per layer: gate|up, down forward + their dgrads
    with torch.no_grad():  # inference path (no autograd bookkeeping) gives the same logits
        torch.testing.assert_close(ours(input_ids=ids).logits, ref(input_ids=ids).logits, atol=5e-5, rtol=1e-4)


def test_fused_residual_decoder_layer_matches_the_stock_layer():
    """accelerate(model, fuse_residual=True): residual add + post-attention RMSNorm in one kernel, second add on ours."""
    for cls, cfg in ((transformers.LlamaForCausalLM, _llama_cfg()),
                     (transformers.MistralForCausalLM,
                      transformers.MistralConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                                 num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=8,
                                                 max_position_embeddings=128))):
        ref, ours = _pair(cls, cfg)
        transformers_b200.accelerate(ours, fuse_residual=True, fused_head_loss=False)
        assert type(ours.model.layers[0]).__name__.startswith("B200")
        torch.manual_seed(14)
        _compare(ref, ours, torch.randint(0, 160, (2, 20)))
        names = [c[0] for c in _fake_ops.CALLS]
        # per layer: forward 1 add (the other is inside rmsnorm_fwd), backward 1 add (dr + dx of the fused norm)
        assert names.count("add") == 2 * 2 and names.count("rmsnorm_fwd") == 2 * 2 + 1
    ours.gradient_checkpointing_enable()
    ours.train()
    ref.train()
    _compare(ref, ours, torch.randint(0, 160, (2, 20)))


def test_gemma_v1_blocks_run_on_the_kernel_path():
    """modeling_gemma (north star: "modeling_llama/mistral/gemma"): Llama structure + (1 + w) RMSNorm + GeGLU + scaled
    embeddings (the scaling is reference code in GemmaModel.forward)."""
    cfg = transformers.GemmaConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=1, head_dim=16, max_position_embeddings=128,
                                   pad_token_id=0)
    ref, ours = _pair(transformers.GemmaForCausalLM, cfg)
    layer = ours.model.layers[0]
    assert type(layer.self_attn).__name__ == "B200GemmaAttention" and type(layer.mlp).__name__ == "B200GemmaMLP"
    assert type(layer.input_layernorm).__name__ == "B200GemmaRMSNorm" and layer.input_layernorm._b200_gemma
    torch.manual_seed(15)
    _compare(ref, ours, torch.randint(1, 160, (2, 20)), atol=5e-5)
    names = [c[0] for c in _fake_ops.CALLS]
    assert names.count("attn_fwd") == 2 and names.count("glu_fwd") == 2


# ------------------------------------------------------------------------------------------- packed (varlen) batches
def _packed_position_ids(lengths):
    return torch.cat([torch.arange(n) for n in lengths])[None]


@pytest.mark.parametrize("cls_cfg", ["llama", "mistral_window"])
def test_packed_batch_from_position_ids_matches_stock_eager(cls_cfg):
    """Padding-free packed batch: position_ids restart at every sequence boundary (what DataCollatorWithFlattening emits).
    The reference composes the sequence indices into the mask function (masking_utils.py:728-757,973-974); the b200 mask
    entry must hand them on and attention must stay inside each sequence (round-1 bug: it attended across boundaries).
    Mirrors the flash path's packed handling, modeling_flash_attention_utils.py:536,796-822."""
    if cls_cfg == "llama":
        ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    else:
        cfg = transformers.MistralConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                                         num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=5,
                                         max_position_embeddings=128)
        ref, ours = _pair(transformers.MistralForCausalLM, cfg)
    lengths = [7, 1, 11, 5]
    torch.manual_seed(2)
    ids = torch.randint(0, 160, (1, sum(lengths)))
    pos = _packed_position_ids(lengths)
    kw = dict(input_ids=ids, position_ids=pos, labels=ids, use_cache=False)
    a = ref(**kw)
    a.loss.backward()
    _fake_ops.CALLS.clear()
    b = ours(**kw)
    b.loss.backward()
    torch.testing.assert_close(b.logits, a.logits, atol=2e-5, rtol=1e-4)
    ga = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        torch.testing.assert_close(p.grad, ga[n].grad, atol=2e-5, rtol=1e-3, msg=lambda m, n=n: f"{n}: {m}")
    names = [c[0] for c in _fake_ops.CALLS]
    assert names.count("attn_fwd") == 2 * len(lengths) and names.count("attn_bwd") == 2 * len(lengths)  # one launch per sequence
    # and the un-packed answer is different (the test would not notice a regression otherwise)
    c = ref(input_ids=ids, labels=ids, use_cache=False)
    assert (c.logits - a.logits).abs().max() > 1e-3


def test_packed_batch_two_rows_and_cu_seq_lens_kwargs():
    """(1) batch size 2, each row packed differently (find_packed_sequence_indices handles any batch size);
    (2) explicit cu_seq_lens_* kwargs of a flattened batch (modeling_flash_attention_utils.py:570-590) without the mask
    entry seeing a position reset (attention function called directly)."""
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg())
    torch.manual_seed(3)
    ids = torch.randint(0, 160, (2, 16))
    pos = torch.stack([_packed_position_ids([4, 12])[0], _packed_position_ids([9, 3, 4])[0]])
    a = ref(input_ids=ids, position_ids=pos, use_cache=False).logits
    b = ours(input_ids=ids, position_ids=pos, use_cache=False).logits
    torch.testing.assert_close(b, a, atol=2e-5, rtol=1e-4)

    from transformers_b200.integration import b200_attention_forward

    B, H, S, D = 1, 2, 12, 16
    q, k, v = (torch.randn(B, H, S, D) for _ in range(3))
    cu = torch.tensor([0, 5, 6, 12], dtype=torch.int32)
    mod = torch.nn.Module()
    mod.is_causal = True
    out, _ = b200_attention_forward(mod, q, k, v, None, cu_seq_lens_q=cu, cu_seq_lens_k=cu, max_length_q=6, max_length_k=6)
    want = torch.cat([torch.nn.functional.scaled_dot_product_attention(q[:, :, s:e], k[:, :, s:e], v[:, :, s:e], is_causal=True)
                      for s, e in ((0, 5), (5, 6), (6, 12))], dim=2).transpose(1, 2)
    torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
    with pytest.raises(transformers_b200.B200Error):
        b200_attention_forward(mod, q, k, v, None, cu_seq_lens_q=torch.tensor([0, 5, 11]), cu_seq_lens_k=torch.tensor([0, 5, 11]))


def test_masks_the_kernels_cannot_express_raise():
    """Interior zeros in a padding mask and mask overlays other than causal / window / packed are refused loudly."""
    from transformers.masking_utils import and_masks, causal_mask_function, chunked_overlay, or_masks

    from transformers_b200.integration import b200_attention_mask
    from transformers_b200.modules import mask_to_kv_ranges

    holes = torch.tensor([[1, 1, 0, 1, 1, 0]], dtype=torch.bool)
    with pytest.raises(transformers_b200.B200Error):
        mask_to_kv_ranges(holes)
    ok = torch.tensor([[0, 0, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0]], dtype=torch.bool)
    s, e = mask_to_kv_ranges(ok)
    assert s.tolist() == [2, 0] and e.tolist() == [5, 0]
    assert mask_to_kv_ranges(ok)[0] is s  # cached on the tensor object: one validation per forward, not per layer
    assert b200_attention_mask(1, 6, 6, mask_function=causal_mask_function) is None
    for fn in (or_masks(causal_mask_function, causal_mask_function),
               and_masks(chunked_overlay(4, torch.zeros(1, dtype=torch.int64)), causal_mask_function)):
        with pytest.raises(transformers_b200.B200Error):
            b200_attention_mask(1, 6, 6, mask_function=fn)


# ------------------------------------------------------------------------------------- chunked fused lm_head + loss
@pytest.mark.parametrize("variant", ["mean", "num_items", "padded_labels"])
def test_fused_head_loss_matches_stock_head_and_loss(variant, monkeypatch):
    """Training forward with labels: lm_head + ForCausalLMLoss run chunk by chunk without the [T, V] logits
    (functional.FusedHeadLossFn: gradients produced inside the forward, wgrad accumulated across chunks).  Loss and every
    gradient must equal the stock model's; ``out.logits`` is the documented empty placeholder; eval forwards, forwards
    without labels and ``fused_head_loss=False`` still return logits."""
    from transformers_b200 import functional as Fn

    monkeypatch.setattr(Fn.FusedHeadLossFn, "CHUNK_ROWS", 16)  # T = 48 -> three chunks
    ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg(), fused_head_loss=True)
    ref.train()
    ours.train()
    torch.manual_seed(5)
    ids = torch.randint(0, 160, (2, 24))
    labels = ids.clone()
    kw = {}
    if variant == "padded_labels":
        labels[0, 5:11] = -100
        labels[1, -4:] = -100
    if variant == "num_items":
        kw["num_items_in_batch"] = torch.tensor(37)
    a = ref(input_ids=ids, labels=labels, use_cache=False, **kw)
    a.loss.backward()
    _fake_ops.CALLS.clear()
    b = ours(input_ids=ids, labels=labels, use_cache=False, **kw)
    assert b.logits.shape == (2, 24, 0)
    (b.loss * 0.5).backward()  # a non-unit upstream gradient (Trainer divides by the accumulation steps)
    torch.testing.assert_close(b.loss, a.loss, atol=1e-5, rtol=1e-5)
    ga = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        torch.testing.assert_close(p.grad, 0.5 * ga[n].grad, atol=2e-5, rtol=1e-3, msg=lambda m, n=n: f"{n}: {m}")
    names = [c[0] for c in _fake_ops.CALLS]
    assert names.count("ce_fwd") == 3 and names.count("ce_bwd") == 3
    with torch.no_grad():
        assert ours(input_ids=ids, labels=labels, use_cache=False).logits.shape == (2, 24, 160)  # no grad: logits as usual
    assert ours(input_ids=ids, use_cache=False).logits.shape == (2, 24, 160)
    ours.eval()
    torch.testing.assert_close(ours(input_ids=ids, labels=labels, use_cache=False).logits,
                               ref.eval()(input_ids=ids, use_cache=False).logits, atol=2e-5, rtol=1e-4)


# ------------------------------------------------------------------------------------- GLU in the gate|up GEMM epilogue
@pytest.mark.parametrize("kind", ["llama_silu", "gemma2_gelu"])
def test_gate_up_glu_fused_path_matches_stock_mlp(kind):
    """MLP widths that are whole 128-column blocks and more than one 128-row tile of tokens take the fused path: ONE launch for
    gate|up projection + activation on the block-interleaved weight (functional.GateUpGluFn), GLU backward on the interleaved
    layout, weight gradient de-interleaved into gate_proj / up_proj.  Everything must equal the stock model."""
    if kind == "llama_silu":
        ref, ours = _pair(transformers.LlamaForCausalLM, _llama_cfg(intermediate_size=384))
    else:
        cfg = transformers.Gemma2Config(vocab_size=160, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                                        num_attention_heads=4, num_key_value_heads=2, head_dim=16, sliding_window=32,
                                        query_pre_attn_scalar=16, max_position_embeddings=256,
                                        layer_types=["sliding_attention", "full_attention"])
        ref, ours = _pair(transformers.Gemma2ForCausalLM, cfg)
    torch.manual_seed(7)
    ids = torch.randint(0, 160, (2, 80))  # T = 160 > 128
    _compare(ref, ours, ids)
    names = [c[0] for c in _fake_ops.CALLS]
    assert names.count("gemm_glu") == 2 and "glu_fwd" not in names and names.count("glu_bwd") == 2
    # the interleaved copy follows parameter updates (same cache discipline as fused_weight)
    mlp = ours.model.layers[0].mlp
    with torch.no_grad():
        mlp.gate_proj.weight.add_(0.01)
        ref.model.layers[0].mlp.gate_proj.weight.add_(0.01)
    ref.zero_grad()
    ours.zero_grad()
    _compare(ref, ours, ids)
