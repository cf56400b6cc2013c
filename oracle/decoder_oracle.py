"""CPU oracle: a plain restatement of the reference's *eager* decoder hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module; the product path (``transformers_b200``) never does and has no CPU fallback.

The arithmetic of the reference lives in PyTorch/ATen (``torch==2.11``, third party, not under /root/reference), so the
restatement calls the same primitive torch ops (matmul, softmax, rsqrt, silu, cross_entropy ...) in the same order and
with the same rounding points as the reference's model files; the backward oracle is torch autograd over this
restatement.  Every function cites the reference lines it follows (paths relative to /root/reference/src/transformers).

Pinning: ``tests/test_oracle_golden.py`` checks every function here against fixtures produced by importing the real
reference in the authoring container (``tests/golden/make_golden.py``), and against the reference's own known-answer
vectors for RoPE inverse frequencies (tests/utils/test_modeling_rope_utils.py:239-273, :1024-1045).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------- config
@dataclass
class DecoderConfig:
    """The subset of LlamaConfig / MistralConfig / Gemma2Config the hot path reads
    (models/llama/configuration_llama.py:64-84, models/gemma2/configuration_gemma2.py:49-101)."""

    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_type: str = "default"
    rope_extra: dict = field(default_factory=dict)  # llama3: factor, low/high_freq_factor, original_max_position_embeddings
    hidden_act: str = "silu"  # or "gelu_pytorch_tanh"
    model_type: str = "llama"  # llama | mistral | mixtral | gemma | gemma2
    sliding_window: int | None = None
    layer_types: list | None = None  # per layer "full_attention" | "sliding_attention"
    attn_logit_softcapping: float | None = None
    final_logit_softcapping: float | None = None
    query_pre_attn_scalar: float | None = None
    tie_word_embeddings: bool = False
    pad_token_id: int | None = None  # nn.Embedding(padding_idx=...) models/llama/modeling_llama.py:350-353
    num_local_experts: int = 0  # > 0: Mixtral sparse MoE block instead of the dense MLP (models/mixtral/modeling_mixtral.py:114-130)
    num_experts_per_tok: int = 2

    @property
    def gemma(self) -> bool:
        """Gemma-2 block structure: four norms per layer, softcaps, query_pre_attn_scalar (models/gemma2/modeling_gemma2.py)."""
        return self.model_type == "gemma2"

    @property
    def gemma_norm(self) -> bool:
        """(1 + w) RMSNorm in fp32 and sqrt(hidden)-scaled embeddings: Gemma (models/gemma/modeling_gemma.py:64-79, :374)
        and Gemma-2 (models/gemma2/modeling_gemma2.py:49-63, :386-389)."""
        return self.model_type in ("gemma", "gemma2")

    @property
    def scaling(self) -> float:
        # LlamaAttention.__init__ models/llama/modeling_llama.py:226; Gemma2Attention models/gemma2/modeling_gemma2.py:229
        if self.gemma and self.query_pre_attn_scalar is not None:
            return self.query_pre_attn_scalar**-0.5
        return self.head_dim**-0.5

    def layer_window(self, layer_idx: int) -> int | None:
        if self.layer_types is not None:
            return self.sliding_window if self.layer_types[layer_idx] == "sliding_attention" else None
        return self.sliding_window


# ------------------------------------------------------------------------------------------------------------- ops
def embedding(ids: torch.Tensor, weight: torch.Tensor, scale: float | None = None, padding_idx: int | None = None) -> torch.Tensor:
    """nn.Embedding row gather, models/llama/modeling_llama.py:353,381 (bit-exact copy of weight rows).
    Gemma2TextScaledWordEmbedding models/gemma2/modeling_gemma2.py:338-349: * bf16(sqrt(hidden)) in weight dtype."""
    out = F.embedding(ids, weight, padding_idx=padding_idx)  # padding_idx only zeroes that row's gradient
    if scale is not None:
        out = out * torch.tensor(scale).to(weight.dtype)
    return out


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, gemma: bool = False) -> torch.Tensor:
    """LlamaRMSNorm.forward models/llama/modeling_llama.py:62-67; Gemma2RMSNorm models/gemma2/modeling_gemma2.py:55-63."""
    if gemma:
        xf = x.float()
        out = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        out = out * (1.0 + weight.float())
        return out.type_as(x)
    input_dtype = x.dtype
    h = x.to(torch.float32)
    variance = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(variance + eps)
    return weight * h.to(input_dtype)


def rope_inv_freq(cfg: DecoderConfig) -> torch.Tensor:
    """compute_default_rope_parameters models/llama/modeling_llama.py:88-111 and _compute_llama3_parameters
    modeling_rope_utils.py:580-665."""
    dim = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).to(dtype=torch.float) / dim))
    if cfg.rope_type == "default":
        return inv_freq
    if cfg.rope_type != "llama3":
        raise ValueError(f"oracle: unsupported rope_type {cfg.rope_type}")
    factor = cfg.rope_extra["factor"]
    low_freq_factor = cfg.rope_extra["low_freq_factor"]
    high_freq_factor = cfg.rope_extra["high_freq_factor"]
    old_context_len = cfg.rope_extra["original_max_position_embeddings"]
    low_freq_wavelen = old_context_len / low_freq_factor
    high_freq_wavelen = old_context_len / high_freq_factor
    wavelen = 2 * math.pi / inv_freq
    inv_freq_llama = torch.where(wavelen > low_freq_wavelen, inv_freq / factor, inv_freq)
    smooth_factor = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
    smoothed_inv_freq = (1 - smooth_factor) * inv_freq_llama / factor + smooth_factor * inv_freq_llama
    is_medium_freq = ~(wavelen < high_freq_wavelen) * ~(wavelen > low_freq_wavelen)
    return torch.where(is_medium_freq, smoothed_inv_freq, inv_freq_llama)


def rope_tables(inv_freq: torch.Tensor, position_ids: torch.Tensor, dtype: torch.dtype, attention_scaling: float = 1.0):
    """LlamaRotaryEmbedding.forward models/llama/modeling_llama.py:113-127 -> (cos, sin) of shape [B, S, D] in `dtype`."""
    inv_freq_expanded = inv_freq[None, :, None].expand(position_ids.shape[0], -1, 1).to(dtype=torch.float)
    position_ids_expanded = position_ids[:, None, :].float()
    freqs = (inv_freq_expanded @ position_ids_expanded).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos = emb.cos() * attention_scaling
    sin = emb.sin() * attention_scaling
    return cos.to(dtype=dtype), sin.to(dtype=dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """models/llama/modeling_llama.py:130-134."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, unsqueeze_dim: int = 1):
    """apply_rotary_pos_emb models/llama/modeling_llama.py:138-160 (q, k are [B, h, S, D])."""
    cos = cos.unsqueeze(unsqueeze_dim)
    sin = sin.unsqueeze(unsqueeze_dim)
    q_embed = (q * cos) + (rotate_half(q) * sin)
    k_embed = (k * cos) + (rotate_half(k) * sin)
    return q_embed, k_embed


def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    """models/llama/modeling_llama.py:179-188: kv head j serves q heads j*n_rep .. (j+1)*n_rep-1."""
    batch, num_key_value_heads, slen, head_dim = hidden_states.shape
    if n_rep == 1:
        return hidden_states
    hidden_states = hidden_states[:, :, None, :, :].expand(batch, num_key_value_heads, n_rep, slen, head_dim)
    return hidden_states.reshape(batch, num_key_value_heads * n_rep, slen, head_dim)


def packed_sequence_ids(position_ids: torch.Tensor) -> torch.Tensor | None:
    """find_packed_sequence_indices masking_utils.py:728-757: the index of the sequence every token of a padding-free packed
    batch belongs to (a new sequence starts wherever consecutive position ids do not differ by one); None when no row holds
    more than one sequence."""
    first = position_ids[:, :1] - 1
    ids = (torch.diff(position_ids, prepend=first, dim=-1) != 1).cumsum(-1)
    return None if bool((ids[:, -1] == 0).all()) else ids


def eager_mask(
    batch: int,
    q_len: int,
    kv_len: int,
    dtype: torch.dtype,
    q_offset: int = 0,
    sliding_window: int | None = None,
    padding_mask: torch.Tensor | None = None,
    sequence_ids: torch.Tensor | None = None,
) -> torch.Tensor:
    """The additive 4-D mask the eager backend receives: 0 where attended, finfo(dtype).min elsewhere.
    causal_mask_function masking_utils.py:76-80 (kv_idx <= q_idx), sliding_window_overlay :92-101
    (kv_idx > q_idx - sliding_window), padding :104-115, packed_sequence_mask_function :182-190 (q and kv in the same sequence,
    and-ed in by create_causal_mask :973-974), eager_mask :538-604 (min-value fill :599-603)."""
    q_idx = torch.arange(q_len)[:, None] + q_offset
    kv_idx = torch.arange(kv_len)[None, :]
    allowed = kv_idx <= q_idx
    if sliding_window is not None:
        allowed = allowed & (kv_idx > q_idx - sliding_window)
    allowed = allowed[None, None].expand(batch, 1, q_len, kv_len)
    if padding_mask is not None:
        allowed = allowed & padding_mask[:, None, None, :kv_len].bool()
    if sequence_ids is not None:
        allowed = allowed & (sequence_ids[:, None, :, None] == sequence_ids[:, None, None, :])
    min_dtype = torch.finfo(dtype).min
    return torch.where(allowed, torch.tensor(0.0, dtype=dtype), torch.tensor(min_dtype, dtype=dtype))


def eager_attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    attention_mask: torch.Tensor | None,
    scaling: float,
    softcap: float | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """eager_attention_forward models/llama/modeling_llama.py:191-213 (softcap branch: models/gemma2/modeling_gemma2.py
    :201-208).  q [B,Hq,Sq,D], k/v [B,Hkv,Skv,D] -> (out [B,Sq,Hq,D], weights [B,Hq,Sq,Skv])."""
    n_rep = q.shape[1] // k.shape[1]
    key_states = repeat_kv(k, n_rep)
    value_states = repeat_kv(v, n_rep)
    attn_weights = torch.matmul(q, key_states.transpose(2, 3)) * scaling
    if softcap is not None:
        attn_weights = attn_weights / softcap
        attn_weights = torch.tanh(attn_weights)
        attn_weights = attn_weights * softcap
    if attention_mask is not None:
        attn_weights = attn_weights + attention_mask
    attn_weights = F.softmax(attn_weights, dim=-1, dtype=torch.float32).to(q.dtype)
    attn_output = torch.matmul(attn_weights, value_states)
    attn_output = attn_output.transpose(1, 2).contiguous()
    return attn_output, attn_weights


def act_fn(x: torch.Tensor, name: str) -> torch.Tensor:
    """ACT2FN activations.py: "silu" -> F.silu (:92-103), "gelu_pytorch_tanh" -> gelu(approximate="tanh") (:30-49)."""
    if name == "silu":
        return F.silu(x)
    if name == "gelu_pytorch_tanh":
        return F.gelu(x, approximate="tanh")
    raise ValueError(name)


def mlp(x: torch.Tensor, w_gate: torch.Tensor, w_up: torch.Tensor, w_down: torch.Tensor, act: str = "silu") -> torch.Tensor:
    """LlamaMLP.forward models/llama/modeling_llama.py:174-176 (no biases for Llama-3: configuration_llama.py:81-83)."""
    return F.linear(act_fn(F.linear(x, w_gate), act) * F.linear(x, w_up), w_down)


def moe_router(x: torch.Tensor, w_gate: torch.Tensor, top_k: int):
    """MixtralTopKRouter.forward models/mixtral/modeling_mixtral.py:104-111: linear -> fp32 softmax -> top-k -> renormalise."""
    router_logits = F.linear(x, w_gate)
    router_probs = F.softmax(router_logits.float(), dim=-1)
    top_value, top_index = torch.topk(router_probs, top_k, dim=-1)
    top_value = top_value / top_value.sum(dim=-1, keepdim=True)
    return top_value, top_index


def moe_experts(x: torch.Tensor, top_k_index: torch.Tensor, top_k_weights: torch.Tensor, gate_up: torch.Tensor,
                down: torch.Tensor, act: str = "silu") -> torch.Tensor:
    """MixtralExperts.forward models/mixtral/modeling_mixtral.py:69-93 (per-expert loop, index_add un-permute)."""
    E = gate_up.shape[0]
    final = torch.zeros_like(x)
    expert_mask = F.one_hot(top_k_index, num_classes=E).permute(2, 1, 0)
    for e in range(E):
        top_k_pos, token_idx = torch.where(expert_mask[e])
        if token_idx.numel() == 0:
            continue
        cur = x[token_idx]
        gate, up = F.linear(cur, gate_up[e]).chunk(2, dim=-1)
        h = act_fn(gate, act) * up
        h = F.linear(h, down[e])
        h = h * top_k_weights[token_idx, top_k_pos, None]
        final.index_add_(0, token_idx, h.to(final.dtype))
    return final


def moe_block(x: torch.Tensor, p: dict, prefix: str, cfg: "DecoderConfig") -> torch.Tensor:
    """MixtralSparseMoeBlock.forward models/mixtral/modeling_mixtral.py:121-130 (no jitter in eval)."""
    B, S, H = x.shape
    flat = x.view(-1, H)
    w, idx = moe_router(flat, p[prefix + "gate.weight"], cfg.num_experts_per_tok)
    out = moe_experts(flat, idx, w.to(flat.dtype) if False else w, p[prefix + "experts.gate_up_proj"], p[prefix + "experts.down_proj"], cfg.hidden_act)
    return out.reshape(B, S, H)


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100, num_items_in_batch=None):
    """ForCausalLMLoss loss/loss_utils.py:48-70 + fixed_cross_entropy :32-45."""
    vocab_size = logits.shape[-1]
    logits = logits.float()
    labels = F.pad(labels, (0, 1), value=ignore_index)
    shift_labels = labels[..., 1:].contiguous()
    logits = logits.view(-1, vocab_size)
    shift_labels = shift_labels.view(-1)
    reduction = "sum" if num_items_in_batch is not None else "mean"
    loss = F.cross_entropy(logits, shift_labels, ignore_index=ignore_index, reduction=reduction)
    if reduction == "sum":
        loss = loss / num_items_in_batch
    return loss


# ---------------------------------------------------------------------------------------------------------- modules
def attention_block(
    x: torch.Tensor,
    p: dict,
    prefix: str,
    cfg: DecoderConfig,
    cos: torch.Tensor,
    sin: torch.Tensor,
    mask: torch.Tensor | None,
) -> torch.Tensor:
    """LlamaAttention.forward models/llama/modeling_llama.py:243-281 (Mistral :141-178, Gemma2 :248-288), eager backend."""
    B, S, _ = x.shape
    hidden_shape = (B, S, -1, cfg.head_dim)
    q = F.linear(x, p[prefix + "q_proj.weight"]).view(hidden_shape).transpose(1, 2)
    k = F.linear(x, p[prefix + "k_proj.weight"]).view(hidden_shape).transpose(1, 2)
    v = F.linear(x, p[prefix + "v_proj.weight"]).view(hidden_shape).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    out, _ = eager_attention(q, k, v, mask, cfg.scaling, cfg.attn_logit_softcapping if cfg.gemma else None)
    out = out.reshape(B, S, -1).contiguous()
    return F.linear(out, p[prefix + "o_proj.weight"])


def decoder_layer(x: torch.Tensor, p: dict, layer_idx: int, cfg: DecoderConfig, cos, sin, mask) -> torch.Tensor:
    """LlamaDecoderLayer.forward models/llama/modeling_llama.py:295-324; Gemma2DecoderLayer.forward
    models/gemma2/modeling_gemma2.py:304-335 (post-norms applied before each residual add)."""
    pre = f"model.layers.{layer_idx}."
    eps, g, gn = cfg.rms_norm_eps, cfg.gemma, cfg.gemma_norm
    residual = x
    h = rms_norm(x, p[pre + "input_layernorm.weight"], eps, gn)
    h = attention_block(h, p, pre + "self_attn.", cfg, cos, sin, mask)
    if g:
        h = rms_norm(h, p[pre + "post_attention_layernorm.weight"], eps, gn)
    h = residual + h
    residual = h
    if g:
        h = rms_norm(h, p[pre + "pre_feedforward_layernorm.weight"], eps, gn)
    else:
        h = rms_norm(h, p[pre + "post_attention_layernorm.weight"], eps, gn)
    if cfg.num_local_experts:
        h = moe_block(h, p, pre + "mlp.", cfg)
    else:
        h = mlp(h, p[pre + "mlp.gate_proj.weight"], p[pre + "mlp.up_proj.weight"], p[pre + "mlp.down_proj.weight"], cfg.hidden_act)
    if g:
        h = rms_norm(h, p[pre + "post_feedforward_layernorm.weight"], eps, gn)
    return residual + h


def model_forward(ids: torch.Tensor, p: dict, cfg: DecoderConfig, labels: torch.Tensor | None = None,
                  padding_mask: torch.Tensor | None = None, position_ids: torch.Tensor | None = None):
    """LlamaModel.forward models/llama/modeling_llama.py:367-418 + LlamaForCausalLM.forward :438-490
    (Gemma2: scaled embedding :386-389, final logit softcap :527-530).  Returns (logits, loss, last_hidden)."""
    B, S = ids.shape
    w_emb = p["model.embed_tokens.weight"]
    dtype = w_emb.dtype
    scale = cfg.hidden_size**0.5 if cfg.gemma_norm else None
    h = embedding(ids, w_emb, scale, cfg.pad_token_id)
    seq_ids = None
    if position_ids is None:
        position_ids = torch.arange(S)[None, :]
    elif padding_mask is None:  # packed batches are only recognised without a padding mask (masking_utils.py:852-860)
        seq_ids = packed_sequence_ids(position_ids.expand(B, -1))
    cos, sin = rope_tables(rope_inv_freq(cfg), position_ids, dtype)
    masks = {}
    for li in range(cfg.num_hidden_layers):
        win = cfg.layer_window(li)
        if win not in masks:
            masks[win] = eager_mask(B, S, S, dtype, sliding_window=win, padding_mask=padding_mask, sequence_ids=seq_ids)
        h = decoder_layer(h, p, li, cfg, cos, sin, masks[win])
    h = rms_norm(h, p["model.norm.weight"], cfg.rms_norm_eps, cfg.gemma_norm)
    w_head = w_emb if cfg.tie_word_embeddings else p["lm_head.weight"]
    logits = F.linear(h, w_head)
    if cfg.final_logit_softcapping is not None:
        logits = logits / cfg.final_logit_softcapping
        logits = torch.tanh(logits)
        logits = logits * cfg.final_logit_softcapping
    loss = causal_lm_loss(logits, labels) if labels is not None else None
    return logits, loss, h


def config_from_hf(hf_cfg) -> DecoderConfig:
    """Build a DecoderConfig from a transformers config object or dict (used by tests and the bench)."""
    d = hf_cfg if isinstance(hf_cfg, dict) else hf_cfg.to_dict()
    rp = d.get("rope_parameters") or {}
    if "rope_type" not in rp and rp and all(isinstance(v, dict) for v in rp.values()):
        rp = rp.get("full_attention", next(iter(rp.values())))
    head_dim = d.get("head_dim") or d["hidden_size"] // d["num_attention_heads"]
    extra = {k: rp[k] for k in ("factor", "low_freq_factor", "high_freq_factor", "original_max_position_embeddings") if k in rp}
    return DecoderConfig(
        vocab_size=d["vocab_size"], hidden_size=d["hidden_size"], intermediate_size=d["intermediate_size"],
        num_hidden_layers=d["num_hidden_layers"], num_attention_heads=d["num_attention_heads"],
        num_key_value_heads=d.get("num_key_value_heads") or d["num_attention_heads"], head_dim=head_dim,
        rms_norm_eps=d.get("rms_norm_eps", 1e-6), rope_theta=rp.get("rope_theta", d.get("rope_theta", 10000.0)),
        rope_type=rp.get("rope_type", "default"), rope_extra=extra,
        hidden_act=d.get("hidden_act") or d.get("hidden_activation") or "silu", model_type=d.get("model_type", "llama"),
        sliding_window=d.get("sliding_window"), layer_types=d.get("layer_types"),
        attn_logit_softcapping=d.get("attn_logit_softcapping"), final_logit_softcapping=d.get("final_logit_softcapping"),
        query_pre_attn_scalar=d.get("query_pre_attn_scalar"), tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
        pad_token_id=d.get("pad_token_id"),
        num_local_experts=(d.get("num_local_experts") or 0) if d.get("model_type") == "mixtral" else 0,
        num_experts_per_tok=d.get("num_experts_per_tok") or 2,
    )
