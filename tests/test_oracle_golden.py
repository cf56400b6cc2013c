"""Pin the oracle: every oracle function against fixtures produced by the real reference (tests/golden/make_golden.py)
and against the reference's own known-answer vectors.  CPU only."""
import pytest
import torch

from conftest import load_golden
from oracle import decoder_oracle as O

# fp32 CPU matmuls may block differently across hosts -> tight but not bitwise; bf16 per the reference's own bars
# (tests/test_modeling_common.py:203-232)
TOL = {"fp32": dict(atol=2e-5, rtol=1e-4), "bf16": dict(atol=1e-2, rtol=1e-2)}

# tests/utils/test_modeling_rope_utils.py:239-252 (default rope, theta 1e4, dim 128)
EXPECTED_DEFAULT_INV_FREQ = [
    1.0000e+00, 8.6596e-01, 7.4989e-01, 6.4938e-01, 5.6234e-01, 4.8697e-01, 4.2170e-01, 3.6517e-01, 3.1623e-01, 2.7384e-01,
    2.3714e-01, 2.0535e-01, 1.7783e-01, 1.5399e-01, 1.3335e-01, 1.1548e-01, 1.0000e-01, 8.6596e-02, 7.4989e-02, 6.4938e-02,
    5.6234e-02, 4.8697e-02, 4.2170e-02, 3.6517e-02, 3.1623e-02, 2.7384e-02, 2.3714e-02, 2.0535e-02, 1.7783e-02, 1.5399e-02,
    1.3335e-02, 1.1548e-02, 1.0000e-02, 8.6596e-03, 7.4989e-03, 6.4938e-03, 5.6234e-03, 4.8697e-03, 4.2170e-03, 3.6517e-03,
    3.1623e-03, 2.7384e-03, 2.3714e-03, 2.0535e-03, 1.7783e-03, 1.5399e-03, 1.3335e-03, 1.1548e-03, 1.0000e-03, 8.6596e-04,
    7.4989e-04, 6.4938e-04, 5.6234e-04, 4.8697e-04, 4.2170e-04, 3.6517e-04, 3.1623e-04, 2.7384e-04, 2.3714e-04, 2.0535e-04,
    1.7783e-04, 1.5399e-04, 1.3335e-04, 1.1548e-04,
]
# tests/utils/test_modeling_rope_utils.py:1026-1041 (llama3, factor 10? no: factor 8 is below) -> values for
# rope_theta 1e4, factor 10... the reference test builds them with factor=10, low=1, high=4, original_max=2048
EXPECTED_LLAMA3_INV_FREQ = [
    1.0000e+00, 8.6596e-01, 7.4989e-01, 6.4938e-01, 5.6234e-01, 4.8697e-01, 4.2170e-01, 3.6517e-01, 3.1623e-01, 2.7384e-01,
    2.3714e-01, 2.0535e-01, 1.7783e-01, 1.5399e-01, 1.3335e-01, 1.1548e-01, 1.0000e-01, 8.6596e-02, 7.4989e-02, 6.4938e-02,
    5.6234e-02, 4.8697e-02, 4.2170e-02, 3.6517e-02, 3.1623e-02, 2.7384e-02, 2.3714e-02, 2.0535e-02, 1.7783e-02, 1.5399e-02,
    1.3335e-02, 1.0730e-02, 7.7785e-03, 5.6009e-03, 3.9991e-03, 2.8248e-03, 1.9675e-03, 1.3449e-03, 8.9549e-04, 5.7363e-04,
    3.4539e-04, 2.7384e-04, 2.3714e-04, 2.0535e-04, 1.7783e-04, 1.5399e-04, 1.3335e-04, 1.1548e-04, 1.0000e-04, 8.6596e-05,
    7.4989e-05, 6.4938e-05, 5.6234e-05, 4.8697e-05, 4.2170e-05, 3.6517e-05, 3.1623e-05, 2.7384e-05, 2.3714e-05, 2.0535e-05,
    1.7783e-05, 1.5399e-05, 1.3335e-05, 1.1548e-05,
]


def _cfg(**kw):
    base = dict(vocab_size=8, hidden_size=4096, intermediate_size=8, num_hidden_layers=1, num_attention_heads=32,
                num_key_value_heads=32, head_dim=128, rope_theta=10000.0)
    base.update(kw)
    return O.DecoderConfig(**base)


def test_rope_inv_freq_reference_known_answers():
    torch.testing.assert_close(O.rope_inv_freq(_cfg()), torch.tensor(EXPECTED_DEFAULT_INV_FREQ), rtol=1e-4, atol=1e-8)
    c = _cfg(rope_type="llama3", rope_extra=dict(factor=10.0, low_freq_factor=1, high_freq_factor=4,
                                                 original_max_position_embeddings=2048))
    torch.testing.assert_close(O.rope_inv_freq(c), torch.tensor(EXPECTED_LLAMA3_INV_FREQ), rtol=1e-4, atol=1e-8)


@pytest.fixture(scope="module")
def ops():
    return load_golden("ops")


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_ops_match_reference(ops, tag):
    tol = TOL[tag]
    f = ops[f"rmsnorm_llama_{tag}"]
    torch.testing.assert_close(O.rms_norm(f["x"], f["w"], f["eps"]), f["y"], **tol)
    f = ops[f"rmsnorm_gemma_{tag}"]
    torch.testing.assert_close(O.rms_norm(f["x"], f["w"], f["eps"], gemma=True), f["y"], **tol)
    for rope in ("default", "llama3"):
        f = ops[f"rope_{rope}_{tag}"]
        cfg = O.config_from_hf(f["config"])
        inv = O.rope_inv_freq(cfg)
        torch.testing.assert_close(inv, f["inv_freq"], rtol=1e-6, atol=0)
        cos, sin = O.rope_tables(inv, f["pos"], f["q"].dtype)
        torch.testing.assert_close(cos, f["cos"], **tol)
        torch.testing.assert_close(sin, f["sin"], **tol)
        q, k = O.apply_rope(f["q"], f["k"], f["cos"], f["sin"])
        assert torch.equal(q, f["q_out"]) and torch.equal(k, f["k_out"])  # pure elementwise: bit-exact
    f = ops[f"attn_llama_{tag}"]
    o, w = O.eager_attention(f["q"], f["k"], f["v"], f["mask"], f["scaling"])
    torch.testing.assert_close(o, f["out"], **tol)
    torch.testing.assert_close(w, f["weights"], **tol)
    f = ops[f"attn_softcap_{tag}"]
    o, w = O.eager_attention(f["q"], f["k"], f["v"], f["mask"], f["scaling"], softcap=f["softcap"])
    torch.testing.assert_close(o, f["out"], **tol)
    f = ops[f"mlp_silu_{tag}"]
    torch.testing.assert_close(O.mlp(f["x"], f["wg"], f["wu"], f["wd"]), f["y"], **tol)
    f = ops[f"loss_{tag}"]
    torch.testing.assert_close(O.causal_lm_loss(f["logits"], f["labels"]), f["loss"], atol=1e-5, rtol=1e-5)


def test_masks_match_reference(ops):
    m = ops["mask_causal"]
    assert torch.equal(O.eager_mask(2, 10, 10, m.dtype), m.expand(2, 1, 10, 10))
    pad = torch.ones(2, 10, dtype=torch.long)
    pad[1, -3:] = 0
    assert torch.equal(O.eager_mask(2, 10, 10, m.dtype, padding_mask=pad), ops["mask_causal_padded"])
    assert torch.equal(O.eager_mask(2, 10, 10, m.dtype, sliding_window=4), ops["mask_sliding4"].expand(2, 1, 10, 10))


MODELS = ["llama_tiny", "llama_tiny_padded", "llama_tiny_packed", "llama3rope_tiny", "mistral_tiny", "gemma1_tiny", "gemma2_tiny",
          "mixtral_tiny"]


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
@pytest.mark.parametrize("name", MODELS)
def test_model_forward_backward_matches_reference(name, tag):
    fx = load_golden(f"{name}_{tag}")
    tol = TOL[tag]
    cfg = O.config_from_hf(fx["config"])
    params = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    logits, loss, last = O.model_forward(fx["input_ids"], params, cfg, labels=fx["labels"], padding_mask=fx["attention_mask"],
                                         position_ids=fx.get("position_ids"))
    # embedding gather is integer indexing: bit-exact rows
    emb = O.embedding(fx["input_ids"], fx["state_dict"]["model.embed_tokens.weight"])
    assert torch.equal(emb[0, 0], fx["state_dict"]["model.embed_tokens.weight"][fx["input_ids"][0, 0]])
    valid = slice(None)
    if fx["attention_mask"] is not None:  # padded query rows are garbage-in/garbage-out in the reference too
        keep = fx["attention_mask"].bool()
        torch.testing.assert_close(logits[keep], fx["logits"][keep], **tol)
    else:
        torch.testing.assert_close(logits, fx["logits"], **tol)
        torch.testing.assert_close(last, fx["last_hidden"], **tol)
    torch.testing.assert_close(loss, fx["loss"], atol=1e-4 if tag == "fp32" else 2e-2, rtol=1e-4 if tag == "fp32" else 1e-2)
    loss.backward()
    gtol = dict(atol=1e-5, rtol=1e-3) if tag == "fp32" else dict(atol=1e-2, rtol=5e-2)
    for n, g in fx["grads"].items():
        key = n
        if key not in params:  # tied head
            continue
        torch.testing.assert_close(params[key].grad, g, **gtol, msg=lambda m, n=n: f"grad {n}: {m}")
