"""Registration behind the reference's own extension points:

1. ``AttentionInterface.register("b200", fn)``            (utils/generic.py:1130-1132, modeling_utils.py:5092-5130)
2. ``AttentionMaskInterface.register("b200", mask_fn)``   (masking_utils.py:711-725)
3. ``register_patch_mapping({...})``                      (monkey_patching.py:82-154) for the Attention / MLP / RMSNorm
   classes of Llama, Mistral and Gemma2, applied by the reference inside ``_from_config`` / ``from_pretrained``
4. ``accelerate(model)``: the same swaps on an already-built model (``module.__class__`` assignment), plus the pieces the
   class map cannot reach: ``nn.Embedding`` gather, ``lm_head`` GEMM and the causal-LM loss.

``Trainer`` and ``generate()`` then call the model unchanged.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as Fn
from . import modules as M
from ._lib import B200Error

ATTN_NAME = M.ATTN_NAME
_enabled = False


# ------------------------------------------------------------------------------------------- attention registry entry
def b200_attention_forward(module, query, key, value, attention_mask, dropout: float = 0.0, scaling: float | None = None,
                           sliding_window: int | None = None, softcap: float | None = None, is_causal: bool | None = None,
                           **kwargs):
    """AttentionInterface entry (signature: docs/source/en/attention_interface.md:164-175).

    query [B,Hq,Sq,D], key/value [B,Hkv,Skv,D] with arbitrary batch/head/seq strides (the reference hands us transposed
    views of [B,S,h,D] storage, or the cat'ed KV cache) -> (attn_output [B,Sq,Hq,D], None).  GQA is resolved inside the
    kernel (kv_head = q_head // n_rep).  The mask is never materialised: causal / sliding-window come from indices, padding
    from the 2-D mask our AttentionMaskInterface entry forwards."""
    if dropout:
        raise B200Error("b200 attention: dropout is not supported")
    if kwargs.get("s_aux") is not None:
        raise B200Error("b200 attention: attention sinks (s_aux) are not supported")
    if query.stride(-1) != 1:
        query = query.contiguous()
    if key.stride(-1) != 1:
        key = key.contiguous()
    if value.stride(-1) != 1:
        value = value.contiguous()
    Sq = query.shape[2]
    if scaling is None:
        scaling = query.shape[-1] ** -0.5
    if is_causal is None:
        is_causal = getattr(module, "is_causal", True)
    # decode (q_len 1 over the whole cache) needs no causal mask (integrations/sdpa_attention.py:124)
    causal = bool(is_causal) and Sq > 1
    kv_start, kv_end, segments = M.mask_info(attention_mask, kwargs, query.shape[0] * Sq)
    if segments is not None and (len(segments) != query.shape[0] or key.shape[2] != Sq):
        raise B200Error("b200 attention: packed batches need one range list per row and q_len == kv_len (no KV cache)")
    if torch.is_tensor(attention_mask) and attention_mask.shape[1] != key.shape[2]:
        raise B200Error("b200 attention: padding mask length does not match the kv length")
    q = query.transpose(1, 2)  # [B, S, h, D] views; no copies
    k = key.transpose(1, 2)
    v = value.transpose(1, 2)
    out = Fn.FlashAttentionFn.apply(q, k, v, float(scaling), causal, int(sliding_window or 0), float(softcap or 0.0),
                                    kv_start, kv_end, segments)
    return out, None


def _packed_ids_of(mask_function):
    """Walk the mask function the reference composed (masking_utils.py:864-996) and return the packed-sequence index
    tensor it carries, if any.  Patterns the kernels compute from indices pass (causal, sliding-window overlay -- the
    window itself comes from the attention module); patterns they cannot express raise instead of being ignored."""
    if mask_function is None:
        return None
    name = getattr(mask_function, "__qualname__", type(mask_function).__name__)
    cells = dict(zip(getattr(getattr(mask_function, "__code__", None), "co_freevars", ()),
                     (c.cell_contents for c in (getattr(mask_function, "__closure__", None) or ()))))
    if name == "causal_mask_function" or name.startswith("sliding_window_overlay.<locals>"):
        return None
    if name.startswith("packed_sequence_mask_function.<locals>"):
        return cells["packed_sequence_mask"]
    if name.startswith("and_masks.<locals>"):
        found = [t for t in (_packed_ids_of(f) for f in cells["mask_functions"]) if t is not None]
        if len(found) > 1:
            raise B200Error("b200 attention mask: more than one packed-sequence overlay")
        return found[0] if found else None
    raise B200Error(f"b200 attention mask: the mask pattern `{name}` cannot be expressed by the b200 kernels (supported: "
                    "causal, sliding window, left / right padding, packed sequences)")


def b200_attention_mask(batch_size, q_length, kv_length, q_offset=0, kv_offset=0, mask_function=None, attention_mask=None,
                        **kwargs):
    """AttentionMaskInterface entry: like the flash backends (masking_utils.py:607-647) return the 2-D padding mask, or
    None when nothing is padded; causal / sliding patterns are computed inside the kernel.  When the reference detected a
    padding-free packed batch (position_ids that restart, masking_utils.py:728-757) it composes the sequence indices into
    ``mask_function`` (:973-974); they are handed on as ``modules.SegmentIds`` so attention stays inside each sequence."""
    packed = _packed_ids_of(mask_function)
    if attention_mask is not None:
        attention_mask = attention_mask[:, -kv_length:]
        if attention_mask.shape[1] == kv_length and attention_mask.all():
            attention_mask = None
    if packed is not None:
        if attention_mask is not None or q_length != kv_length:
            raise B200Error("b200 attention mask: packed sequences together with padding or a KV cache are not supported")
        return M.SegmentIds(packed)
    return attention_mask


# ---------------------------------------------------------------------------------------------- experts registry entry
def b200_experts_forward(self, hidden_states, top_k_index, top_k_weights):
    """ExpertsInterface entry (integrations/moe.py:481-506, call site :568-570): fn(module, hidden_states [T,H],
    top_k_index [T,k] int64, top_k_weights [T,k]) -> [T,H].  The module supplies gate_up_proj [E,2I,H], down_proj [E,H,I]
    and act_fn (MixtralExperts models/mixtral/modeling_mixtral.py:56-93).  Under autograd the call goes through
    functional.MoEExpertsFn (same kernels, with backward); without it through the allocation-lean inference path."""
    from . import ops

    if getattr(self, "is_transposed", False) or getattr(self, "has_bias", False) or not getattr(self, "has_gate", True):
        raise B200Error("b200 experts: only concatenated, untransposed, bias-free gate_up_proj experts are supported")
    act = getattr(self.config, "hidden_act", "silu")
    if act not in ("silu", "gelu_pytorch_tanh"):
        raise B200Error(f"b200 experts: activation {act} not supported")
    shape = hidden_states.shape
    x = hidden_states.reshape(-1, shape[-1])
    gelu = act == "gelu_pytorch_tanh"
    if torch.is_grad_enabled() and (x.requires_grad or top_k_weights.requires_grad or self.gate_up_proj.requires_grad
                                    or self.down_proj.requires_grad):
        out = Fn.MoEExpertsFn.apply(x, top_k_index, top_k_weights, self.gate_up_proj, self.down_proj, gelu)
    else:
        out = ops.moe_experts_forward(x, top_k_index, top_k_weights, self.gate_up_proj, self.down_proj, gelu)
    return out.view(shape)


# -------------------------------------------------------------------------------------------------------- class maps
_CLASS_MAP = None


def _class_map() -> dict:
    global _CLASS_MAP
    if _CLASS_MAP is None:
        _CLASS_MAP = _build_class_map()
    return _CLASS_MAP


def _build_class_map() -> dict:
    from transformers.models.gemma2 import modeling_gemma2 as g2
    from transformers.models.llama import modeling_llama as ll
    from transformers.models.mistral import modeling_mistral as mi
    from transformers.models.mixtral import modeling_mixtral as mx

    mk = M.make_class
    extra = {}
    try:  # Gemma (v1): Llama block structure with the (1 + w) norm and a GeGLU MLP
        from transformers.models.gemma import modeling_gemma as g1

        extra = {
            "GemmaRMSNorm": mk(g1.GemmaRMSNorm, M.B200RMSNormMixin, _b200_gemma=True),
            "GemmaMLP": mk(g1.GemmaMLP, M.B200MLPMixin),
            "GemmaAttention": mk(g1.GemmaAttention, M.B200AttentionMixin),
            "GemmaRotaryEmbedding": mk(g1.GemmaRotaryEmbedding, M.B200RotaryEmbeddingMixin),
        }
    except ImportError:  # pragma: no cover
        pass
    return {
        **extra,
        "LlamaRMSNorm": mk(ll.LlamaRMSNorm, M.B200RMSNormMixin),
        "LlamaMLP": mk(ll.LlamaMLP, M.B200MLPMixin),
        "LlamaAttention": mk(ll.LlamaAttention, M.B200AttentionMixin),
        "LlamaRotaryEmbedding": mk(ll.LlamaRotaryEmbedding, M.B200RotaryEmbeddingMixin),
        "MistralRotaryEmbedding": mk(mi.MistralRotaryEmbedding, M.B200RotaryEmbeddingMixin),
        "MixtralRotaryEmbedding": mk(mx.MixtralRotaryEmbedding, M.B200RotaryEmbeddingMixin),
        "Gemma2RotaryEmbedding": mk(g2.Gemma2RotaryEmbedding, M.B200RotaryEmbeddingMixin),
        "MistralRMSNorm": mk(mi.MistralRMSNorm, M.B200RMSNormMixin),
        "MistralMLP": mk(mi.MistralMLP, M.B200MLPMixin),
        "MistralAttention": mk(mi.MistralAttention, M.B200AttentionMixin),
        "MixtralRMSNorm": mk(mx.MixtralRMSNorm, M.B200RMSNormMixin),
        "MixtralAttention": mk(mx.MixtralAttention, M.B200AttentionMixin),
        "Gemma2RMSNorm": mk(g2.Gemma2RMSNorm, M.B200RMSNormMixin, _b200_gemma=True),
        "Gemma2MLP": mk(g2.Gemma2MLP, M.B200MLPMixin),
        "Gemma2Attention": mk(g2.Gemma2Attention, M.B200AttentionMixin),
    }


def enable(patch_modules: bool = True) -> None:
    """Register the b200 backend with the reference's registries (idempotent)."""
    global _enabled
    from transformers import AttentionInterface, AttentionMaskInterface

    AttentionInterface.register(ATTN_NAME, b200_attention_forward)
    AttentionMaskInterface.register(ATTN_NAME, b200_attention_mask)
    try:  # Mixtral-style experts (config 4): selected with experts_implementation="b200"
        from transformers.integrations.moe import ExpertsInterface

        ExpertsInterface.register(ATTN_NAME, b200_experts_forward)
    except ImportError:  # pragma: no cover - very old transformers
        pass
    if patch_modules and not _enabled:
        from transformers.monkey_patching import register_patch_mapping

        register_patch_mapping(_class_map(), overwrite=True)
    from .tp_styles import register_tp_styles

    register_tp_styles()  # "b200_colwise" / "b200_rowwise" / "b200_colwise_gather_output" in the reference's ParallelInterface
    _enabled = True


def b200_causal_lm_loss(logits, labels, vocab_size=None, num_items_in_batch=None, ignore_index: int = -100,
                        shift_labels=None, **kwargs):
    """Drop-in for ForCausalLMLoss (loss/loss_utils.py:48-70) on the fused CE kernels."""
    if shift_labels is not None:
        labels, shift = shift_labels, False
    else:
        shift = True
    if not M._on_b200(logits) or logits.dtype not in M.KERNEL_DTYPES:
        raise B200Error("b200 loss: expects CUDA bf16 logits")
    if torch.is_tensor(num_items_in_batch):
        num_items_in_batch = float(num_items_in_batch)
    lazy = getattr(logits, "_b200_lazy_head", None)
    if lazy is not None:  # lm_head skipped the logits (accelerate(fused_head_loss=True), training forward with labels)
        h, w_fused, w_param = lazy
        return Fn.FusedHeadLossFn.apply(h, w_fused, labels, ignore_index, num_items_in_batch, shift, w_param)
    shard_group = getattr(logits, "_b200_vocab_shard", None)
    if shard_group is not None:  # lm_head left its output vocabulary-sharded (parallel.tensor_parallelize(vocab_parallel_loss=True))
        return Fn.VocabParallelLossFn.apply(logits, labels, ignore_index, num_items_in_batch, shift, shard_group)
    if logits.dim() == 2:
        logits = logits.unsqueeze(0)
        labels = labels.reshape(1, -1)
    return Fn.CausalLMLossFn.apply(logits, labels, ignore_index, num_items_in_batch, shift)


def _install_fused_head_loss(model: nn.Module) -> None:
    """Training forwards that carry ``labels`` (and use our loss) do not materialise the logits: lm_head hands hidden states
    and weight to the loss, which walks the tokens in chunks (functional.FusedHeadLossFn).  ``out.logits`` of such a forward
    is an empty [B, S, 0] placeholder -- like the fused linear-cross-entropy kernels the reference points at
    (integrations/hub_kernels.py:509-515); eval-mode forwards and forwards without labels return logits as usual."""
    head = model.lm_head
    if model.__dict__.get("_b200_fused_head_hooks"):
        return

    def before(module, args, kwargs):
        ours = getattr(module, "loss_function", None) is b200_causal_lm_loss
        keep = kwargs.get("logits_to_keep", 0)
        head.__dict__["_b200_lazy_logits"] = bool(ours and module.training and kwargs.get("labels") is not None
                                                   and isinstance(keep, int) and keep == 0
                                                   and not head.__dict__.get("_b200_keep_vocab_shard", False))

    def after(module, args, kwargs, output):
        head.__dict__["_b200_lazy_logits"] = False

    model.register_forward_pre_hook(before, with_kwargs=True)
    model.register_forward_hook(after, with_kwargs=True, always_call=True)
    model.__dict__["_b200_fused_head_hooks"] = True


def accelerate(model: nn.Module, attn: bool = True, head_and_loss: bool = True, pack_weights: bool = False,
               fuse_residual: bool = True, fused_head_loss: bool = True, fuse_glu: bool = True) -> nn.Module:
    """Convert an already constructed reference model in place (class swap, parameters untouched).
    ``pack_weights``: additionally make q/k/v and gate/up weights row views of one buffer (modules.pack_weights).
    ``fuse_residual`` (default): Llama / Mistral decoder layers run their residual adds on our kernels, the first fused
    with the post-attention RMSNorm (modules.B200DecoderLayerMixin); False keeps the reference's ``torch.add``.
    ``fused_head_loss`` (default): training forwards with labels run lm_head + loss chunk by chunk without materialising the
    [T, V] logits (``_install_fused_head_loss``); False always returns logits.
    ``fuse_glu`` (default): the MLP's gate|up projection runs the gated activation in the GEMM epilogue (one kernel,
    functional.GateUpGluFn) whenever the shapes allow; False keeps GEMM + GLU kernel (A/B measurements)."""
    enable()
    cmap = _class_map()
    by_base = {cls.__mro__[2]: cls for cls in cmap.values()}  # (B200X, mixin, base, ...)
    for mod in model.modules():
        target = by_base.get(type(mod))
        if target is not None:
            mod.__class__ = target
        elif type(mod) is nn.Embedding or type(mod).__name__ == "Gemma2TextScaledWordEmbedding":
            mod.__class__ = M.make_class(type(mod), M.B200EmbeddingMixin)
    if head_and_loss and hasattr(model, "lm_head") and type(model.lm_head) is nn.Linear and model.lm_head.bias is None:
        model.lm_head.__class__ = M.make_class(nn.Linear, M.B200LinearMixin)
        if getattr(model.config, "final_logit_softcapping", None) is None:
            model.loss_function = b200_causal_lm_loss
            if fused_head_loss:
                _install_fused_head_loss(model)
    if attn and hasattr(model, "set_attn_implementation"):
        model.set_attn_implementation(ATTN_NAME)
    if attn and any(hasattr(m, "gate_up_proj") and hasattr(m, "num_experts") for m in model.modules()):
        # MoE experts modules dispatch on config._experts_implementation at call time (integrations/moe.py:568-570); older
        # transformers only validate the built-in names at construction, so the switch happens here
        cfgs = [model.config] + ([model.config.get_text_config()] if hasattr(model.config, "get_text_config") else [])
        for c in cfgs:
            try:
                c._experts_implementation = ATTN_NAME
            except Exception:
                c._experts_implementation_internal = ATTN_NAME
    if not fuse_glu:
        for mod in model.modules():
            if isinstance(mod, M.B200MLPMixin):
                mod.__dict__["_b200_fuse_glu"] = False
    if pack_weights:
        M.pack_weights(model)
    if fuse_residual:
        for mod in model.modules():
            if type(mod).__name__ in ("LlamaDecoderLayer", "MistralDecoderLayer"):
                mod.__class__ = M.make_class(type(mod), M.B200DecoderLayerMixin)
    return model
