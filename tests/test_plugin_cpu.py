"""Boundary behaviour on CPU (no kernels run): registry entries, patch mapping, class swap, switchability, loud failure.
Mirrors tests/test_modeling_common.py:4621-4668 (test_can_set_attention_dynamically) and the GPT-2 plumbing config
(BASELINE.json configs[0])."""
import pytest
import torch

from _hf import import_transformers
from conftest import load_golden

transformers = import_transformers()
import transformers_b200  # noqa: E402
from transformers_b200 import B200Error  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _enabled():
    transformers_b200.enable()
    yield


def _tiny_cfg(**kw):
    base = dict(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, rms_norm_eps=1e-5, max_position_embeddings=128,
                rope_parameters={"rope_type": "default", "rope_theta": 500000.0})
    base.update(kw)
    return transformers.LlamaConfig(**base)


def test_registries_hold_b200():
    from transformers import AttentionInterface, AttentionMaskInterface
    from transformers.monkey_patching import get_patch_mapping

    assert "b200" in AttentionInterface().valid_keys()
    assert "b200" in AttentionMaskInterface().valid_keys()
    mapping = get_patch_mapping()
    for k in ("LlamaAttention", "LlamaMLP", "LlamaRMSNorm", "MistralAttention", "Gemma2RMSNorm", "Gemma2Attention"):
        assert k in mapping and issubclass(mapping[k], torch.nn.Module)


def test_patched_model_keeps_names_and_matches_reference_on_cpu():
    fx = load_golden("llama_tiny_fp32")
    cfg = transformers.LlamaConfig(**{k: v for k, v in fx["config"].items() if k not in ("model_type", "transformers_version", "architectures")})
    model = transformers.LlamaForCausalLM._from_config(cfg, attn_implementation="eager", dtype=torch.float32)
    layer = model.model.layers[0]
    assert type(layer.self_attn).__name__ == "B200LlamaAttention"
    assert type(layer.mlp).__name__ == "B200LlamaMLP"
    assert type(layer.input_layernorm).__name__ == "B200LlamaRMSNorm"
    assert isinstance(layer.self_attn, transformers.models.llama.modeling_llama.LlamaAttention)
    assert set(model.state_dict()) == set(fx["state_dict"])  # parameter names / shapes unchanged
    model.load_state_dict(fx["state_dict"])
    out = model(input_ids=fx["input_ids"], labels=fx["labels"])
    # CPU tensors + eager: our modules defer to the stock forward -> the reference's own numbers
    torch.testing.assert_close(out.logits, fx["logits"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(out.loss, fx["loss"], atol=1e-5, rtol=1e-5)


def test_switchable_and_fails_loudly_on_cpu():
    model = transformers.LlamaForCausalLM._from_config(_tiny_cfg(), attn_implementation="b200", dtype=torch.bfloat16)
    assert model.config._attn_implementation == "b200"
    ids = torch.randint(0, 160, (1, 8))
    with pytest.raises(B200Error):  # no CPU fallback: selecting b200 without a B200 must raise, not silently run eager
        model(ids)
    model.set_attn_implementation("eager")
    assert torch.isfinite(model(ids).logits.float()).all()
    model.set_attn_implementation("b200")
    assert model.config._attn_implementation == "b200"


def test_accelerate_swaps_classes_in_place():
    from transformers.monkey_patching import clear_patch_mapping

    clear_patch_mapping()
    try:
        model = transformers.LlamaForCausalLM._from_config(_tiny_cfg(), attn_implementation="eager", dtype=torch.float32)
        assert type(model.model.layers[0].mlp).__name__ == "LlamaMLP"
        before = {k: v.clone() for k, v in model.state_dict().items()}
        transformers_b200.accelerate(model)
        assert type(model.model.layers[0].mlp).__name__ == "B200LlamaMLP"
        assert type(model.model.embed_tokens).__name__ == "B200Embedding"
        assert type(model.lm_head).__name__ == "B200Linear"
        assert model.config._attn_implementation == "b200"
        after = model.state_dict()
        assert set(before) == set(after) and all(torch.equal(before[k], after[k]) for k in before)
    finally:
        import transformers_b200.integration as integ

        integ._enabled = False
        transformers_b200.enable()


def test_mask_entry_returns_2d_padding_mask_or_none():
    from transformers_b200.integration import b200_attention_mask
    from transformers_b200.modules import mask_to_kv_ranges

    am = torch.ones(2, 10, dtype=torch.bool)
    assert b200_attention_mask(2, 10, 10, attention_mask=am) is None
    am[1, :3] = False
    am[0, -2:] = False
    m = b200_attention_mask(2, 10, 10, attention_mask=am)
    assert m.shape == (2, 10)
    s, e = mask_to_kv_ranges(m)
    assert s.tolist() == [0, 3] and e.tolist() == [8, 10]


def test_gpt2_plumbing_config_unchanged():
    """configs[0]: GPT-2-small eager CPU forward, B=1 S=128 -- with our backend registered but eager selected the
    reference's own path runs (same torch ops)."""
    cfg = transformers.GPT2Config()
    torch.manual_seed(0)
    model = transformers.GPT2LMHeadModel._from_config(cfg, attn_implementation="eager", dtype=torch.float32).eval()
    ids = torch.randint(0, 50257, (1, 128))
    with torch.no_grad():
        a = model(ids).logits
        b = model(ids).logits
    assert a.shape == (1, 128, 50257) and torch.equal(a, b)
    assert type(model.transformer.h[0].attn).__name__ == "GPT2Attention"
