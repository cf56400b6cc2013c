"""Generate golden fixtures from the REAL reference (huggingface/transformers @ /root/reference, v5.16.0.dev0).

Run in the authoring container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/*.pt (small: tiny random-init models, CPU eager path).  Each fixture holds the config dict, the
state_dict, the inputs, and the reference's outputs: logits, loss, last hidden state and every parameter gradient.
Per-op fixtures hold inputs/outputs of the reference's own module classes / functions.
"""
import os
import sys
import types

REF = "/root/reference/src"
sys.path.insert(0, REF)
stub = types.ModuleType("transformers.dependency_versions_check")  # tokenizers version gate
stub.dep_version_check = lambda *a, **k: None
sys.modules["transformers.dependency_versions_check"] = stub

import torch  # noqa: E402
import transformers  # noqa: E402

assert transformers.__file__.startswith(REF), transformers.__file__
from transformers import Gemma2Config, Gemma2ForCausalLM, GemmaConfig, GemmaForCausalLM, LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM, MixtralConfig, MixtralForCausalLM, set_seed  # noqa: E402
from transformers.models.llama import modeling_llama as ml  # noqa: E402
from transformers.models.gemma2 import modeling_gemma2 as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def model_fixture(name, cls, cfg, dtype, B=2, S=24, pad=False, packed=None):
    set_seed(42)
    extra = {"experts_implementation": "eager"} if hasattr(cfg, "num_local_experts") else {}  # per-expert loop = the eager path
    model = cls._from_config(cfg, attn_implementation="eager", dtype=dtype, **extra)
    model.train()
    # non-trivial norm weights so their gradients are exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(torch.randn_like(p.float()).to(p.dtype) * 0.1)
    torch.manual_seed(0)
    if packed is not None:  # padding-free packed batch: several sequences along one row, position_ids restart at each of them
        B, S = 1, sum(packed)
    ids = torch.randint(0, cfg.vocab_size, (B, S))
    labels = ids.clone()
    kw = {}
    if packed is not None:
        kw["position_ids"] = torch.cat([torch.arange(n) for n in packed])[None]
        kw["use_cache"] = False  # the reference only looks for packed sequences when no cache is in play (masking_utils.py:852-860)
    if pad:
        am = torch.ones(B, S, dtype=torch.long)
        am[1, -5:] = 0  # right padding on the second sequence
        labels[am == 0] = -100
        kw["attention_mask"] = am
    out = model(input_ids=ids, labels=labels, output_hidden_states=True, **kw)
    out.loss.backward()
    fx = {
        "reference_version": transformers.__version__,
        "torch_version": torch.__version__,
        "config": cfg.to_dict(),
        "dtype": str(dtype),
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "input_ids": ids,
        "labels": labels,
        "attention_mask": kw.get("attention_mask"),
        "position_ids": kw.get("position_ids"),
        "logits": out.logits.detach().clone(),
        "loss": out.loss.detach().clone(),
        "last_hidden": out.hidden_states[-1].detach().clone(),
        "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None},
    }
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "loss", float(out.loss), "logits", tuple(out.logits.shape))


def llama_cfg(**kw):
    base = dict(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, rms_norm_eps=1e-5, max_position_embeddings=128,
                rope_parameters={"rope_type": "default", "rope_theta": 500000.0}, attention_bias=False, mlp_bias=False,
                tie_word_embeddings=False, hidden_act="silu", attention_dropout=0.0)
    base.update(kw)
    return LlamaConfig(**base)


def op_fixtures():
    torch.manual_seed(1)
    fx = {}
    for dtype in (torch.float32, torch.bfloat16):
        tag = "bf16" if dtype == torch.bfloat16 else "fp32"
        # RMSNorm (Llama + Gemma2)
        x = torch.randn(3, 5, 64).to(dtype)
        n = ml.LlamaRMSNorm(64, eps=1e-5).to(dtype)
        with torch.no_grad():
            n.weight.copy_((1 + 0.1 * torch.randn(64)).to(dtype))
        fx[f"rmsnorm_llama_{tag}"] = {"x": x, "w": n.weight.detach().clone(), "eps": 1e-5, "y": n(x).detach()}
        g = mg.Gemma2RMSNorm(64, eps=1e-6).to(dtype)
        with torch.no_grad():
            g.weight.copy_((0.1 * torch.randn(64)).to(dtype))
        fx[f"rmsnorm_gemma_{tag}"] = {"x": x, "w": g.weight.detach().clone(), "eps": 1e-6, "y": g(x).detach()}
        # RoPE tables + apply (llama3-scaled and default)
        for rope in ("default", "llama3"):
            rp = {"rope_type": "default", "rope_theta": 500000.0} if rope == "default" else {
                "rope_type": "llama3", "rope_theta": 500000.0, "factor": 8.0, "low_freq_factor": 1.0,
                "high_freq_factor": 4.0, "original_max_position_embeddings": 64}
            cfg = llama_cfg(rope_parameters=rp, head_dim=32, hidden_size=128, max_position_embeddings=256)
            rot = ml.LlamaRotaryEmbedding(cfg)
            pos = torch.arange(40)[None, :]
            q = torch.randn(2, 4, 40, 32).to(dtype)
            k = torch.randn(2, 2, 40, 32).to(dtype)
            cos, sin = rot(q, pos)
            qe, ke = ml.apply_rotary_pos_emb(q, k, cos, sin)
            fx[f"rope_{rope}_{tag}"] = {"config": cfg.to_dict(), "inv_freq": rot.inv_freq.clone(), "pos": pos, "q": q, "k": k,
                                        "cos": cos, "sin": sin, "q_out": qe, "k_out": ke}
        # eager attention (GQA, causal additive mask; with and without softcap)
        q = torch.randn(2, 4, 12, 16).to(dtype)
        k = torch.randn(2, 2, 12, 16).to(dtype)
        v = torch.randn(2, 2, 12, 16).to(dtype)
        mod = types.SimpleNamespace(num_key_value_groups=2, training=False, head_dim=16)
        mn = torch.finfo(dtype).min
        mask = torch.triu(torch.full((12, 12), mn, dtype=dtype), diagonal=1)[None, None].expand(2, 1, 12, 12)
        o, w = ml.eager_attention_forward(mod, q, k, v, mask, scaling=0.25)
        fx[f"attn_llama_{tag}"] = {"q": q, "k": k, "v": v, "mask": mask, "scaling": 0.25, "out": o, "weights": w}
        o, w = mg.eager_attention_forward(mod, q, k, v, mask, scaling=0.25, softcap=5.0)
        fx[f"attn_softcap_{tag}"] = {"q": q, "k": k, "v": v, "mask": mask, "scaling": 0.25, "softcap": 5.0, "out": o, "weights": w}
        # MLP
        cfg = llama_cfg()
        m = ml.LlamaMLP(cfg).to(dtype)
        x = torch.randn(2, 7, 64).to(dtype)
        fx[f"mlp_silu_{tag}"] = {"x": x, "wg": m.gate_proj.weight.detach().clone(), "wu": m.up_proj.weight.detach().clone(),
                                 "wd": m.down_proj.weight.detach().clone(), "y": m(x).detach()}
        # loss
        logits = torch.randn(2, 9, 50).to(dtype)
        labels = torch.randint(0, 50, (2, 9))
        labels[0, 3] = -100
        from transformers.loss.loss_utils import ForCausalLMLoss
        fx[f"loss_{tag}"] = {"logits": logits, "labels": labels, "loss": ForCausalLMLoss(logits, labels, 50)}
    # masks: the 4-D additive mask the eager backend receives (causal, sliding, padded)
    from transformers.masking_utils import create_causal_mask, create_sliding_window_causal_mask
    cfg = llama_cfg()
    cfg._attn_implementation = "eager"
    emb = torch.zeros(2, 10, 64)
    pad = torch.ones(2, 10, dtype=torch.long)
    pad[1, -3:] = 0
    fx["mask_causal"] = create_causal_mask(config=cfg, inputs_embeds=emb, attention_mask=None, past_key_values=None, position_ids=torch.arange(10)[None])
    fx["mask_causal_padded"] = create_causal_mask(config=cfg, inputs_embeds=emb, attention_mask=pad, past_key_values=None, position_ids=torch.arange(10)[None])
    mcfg = MistralConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=16, sliding_window=4)
    mcfg._attn_implementation = "eager"
    fx["mask_sliding4"] = create_sliding_window_causal_mask(config=mcfg, inputs_embeds=emb, attention_mask=None, past_key_values=None, position_ids=torch.arange(10)[None])
    torch.save(fx, os.path.join(OUT, "ops.pt"))
    print("ops", len(fx))


if __name__ == "__main__":
    if not os.environ.get("GOLDEN_ONLY"):
        op_fixtures()
    only = os.environ.get("GOLDEN_ONLY")  # e.g. GOLDEN_ONLY=mixtral regenerates just that family
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        jobs = {
            "llama_tiny": lambda: model_fixture(f"llama_tiny_{tag}", LlamaForCausalLM, llama_cfg(), dtype),
            "llama_tiny_padded": lambda: model_fixture(f"llama_tiny_padded_{tag}", LlamaForCausalLM, llama_cfg(), dtype, pad=True),
            "llama_tiny_packed": lambda: model_fixture(f"llama_tiny_packed_{tag}", LlamaForCausalLM, llama_cfg(), dtype, packed=[7, 1, 11, 5]),
            "llama3rope_tiny": lambda: model_fixture(f"llama3rope_tiny_{tag}", LlamaForCausalLM, llama_cfg(rope_parameters={
                "rope_type": "llama3", "rope_theta": 500000.0, "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                "original_max_position_embeddings": 16}), dtype),
            "mistral_tiny": lambda: model_fixture(f"mistral_tiny_{tag}", MistralForCausalLM, MistralConfig(
                vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, sliding_window=8, rms_norm_eps=1e-5,
                rope_parameters={"rope_type": "default", "rope_theta": 10000.0}), dtype),
            "mixtral_tiny": lambda: model_fixture(f"mixtral_tiny_{tag}", MixtralForCausalLM, MixtralConfig(
                vocab_size=160, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, num_local_experts=4, num_experts_per_tok=2, sliding_window=None,
                rms_norm_eps=1e-5, rope_parameters={"rope_type": "default", "rope_theta": 10000.0}), dtype),
            "gemma1_tiny": lambda: model_fixture(f"gemma1_tiny_{tag}", GemmaForCausalLM, GemmaConfig(
                vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=1, head_dim=32, rope_parameters={"rope_type": "default", "rope_theta": 10000.0}), dtype),
            "gemma2_tiny": lambda: model_fixture(f"gemma2_tiny_{tag}", Gemma2ForCausalLM, Gemma2Config(
                vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, head_dim=32, sliding_window=8, query_pre_attn_scalar=32,
                attn_logit_softcapping=50.0, final_logit_softcapping=30.0, layer_types=["sliding_attention", "full_attention"],
                rope_parameters={"rope_type": "default", "rope_theta": 10000.0}), dtype),
        }
        for name, job in jobs.items():
            if only is None or only in name:
                job()
