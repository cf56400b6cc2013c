"""Drop-in replacements for the reference's decoder sub-modules (Llama / Mistral / Gemma2).

Each class subclasses the reference class it replaces, keeps its constructor signature, child-module names and
parameter shapes (``q_proj/k_proj/v_proj/o_proj``, ``gate_proj/up_proj/down_proj``, ``weight``) so state dicts,
``tp_plan`` keys and ``_can_record_outputs`` keep working, and overrides only ``forward``.
They hold no extra constructor state, so an existing model can also be converted by swapping ``module.__class__``
(``transformers_b200.accelerate(model)``).

When ``config._attn_implementation`` is not ``"b200"`` the attention module defers to the stock reference forward, so
``model.set_attn_implementation("eager" | "sdpa" | "b200")`` keeps switching backends (modeling_utils.py:2041-2141).
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as Fn
from ._lib import B200Error

ATTN_NAME = "b200"
KERNEL_DTYPES = (torch.bfloat16,)


def _on_b200(t: torch.Tensor) -> bool:
    """Whether ``t`` takes the kernel path.  Off-GPU tensors defer to the stock reference forward (plumbing / CPU tests);
    the kernels themselves never fall back (ops.py raises)."""
    return t.is_cuda


# ----------------------------------------------------------------------------------------------------- weight fusion
def fused_weight(module: nn.Module, key: str, params: list[torch.Tensor]) -> torch.Tensor:
    """Row-wise concatenation of several [N_i, K] weights as ONE contiguous bf16 buffer, cached on the module and
    rebuilt when any parameter is replaced or updated in place (optimizer step bumps ``_version``)."""
    if len(params) == 1:
        w = params[0]
        return w if w.is_contiguous() else w.contiguous()
    packed = module.__dict__.get("_b200_packed")
    if packed is not None and key in packed:
        if _is_packed(packed[key], params):
            return packed[key]  # the parameters ARE row views of this buffer: nothing to concatenate, ever
        del packed[key]  # the parameters were re-allocated (.to(), load with assign=True, TP re-shard): fall back to the copy
    sig = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params)
    cache = module.__dict__.setdefault("_b200_fused", {})
    hit = cache.get(key)
    if hit is not None and hit[0] == sig:
        return hit[1]
    with torch.no_grad():
        buf = torch.cat([p.detach() for p in params], dim=0).contiguous()
    cache[key] = (sig, buf)
    return buf


def interleaved_weight(module: nn.Module, wg: torch.Tensor, wu: torch.Tensor) -> torch.Tensor:
    """The [2I, K] block-interleaved gate|up weight of the GLU-epilogue GEMM (128 gate rows, the matching 128 up rows, ...),
    cached on the module and rebuilt when either parameter is replaced or updated (same key as ``fused_weight``)."""
    from . import ops

    sig = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in (wg, wu))
    cache = module.__dict__.setdefault("_b200_fused", {})
    hit = cache.get("gate_up_ilv")
    if hit is not None and hit[0] == sig:
        return hit[1]
    with torch.no_grad():
        buf = ops.interleave_gate_up(wg.detach(), wu.detach())
    cache["gate_up_ilv"] = (sig, buf)
    return buf


def _is_packed(buf: torch.Tensor, params: list[torch.Tensor]) -> bool:
    if buf.dim() != 2 or not buf.is_contiguous() or sum(p.shape[0] for p in params) != buf.shape[0]:
        return False
    K, ptr, es = buf.shape[1], buf.data_ptr(), buf.element_size()
    for p in params:
        if (p.dim() != 2 or p.shape[1] != K or p.dtype != buf.dtype or p.device != buf.device or p.stride() != (K, 1)
                or p.data_ptr() != ptr):
            return False
        ptr += p.shape[0] * K * es
    return True


PACK_GROUPS = {"qkv": ("q_proj", "k_proj", "v_proj"), "gate_up": ("gate_proj", "up_proj")}


def pack_weights(model: nn.Module) -> int:
    """Checkpoint-compatible fused weight layout: re-home the q/k/v and the gate/up projection weights
    of every block as adjacent row views of ONE contiguous buffer.  Names, shapes and ``state_dict`` are unchanged (each
    ``*_proj.weight`` is still its own Parameter, optimizers update it in place), but the packed operand of the fused GEMM
    now IS the parameter storage: no second copy in HBM (-9 GB for Llama-3-8B) and no re-concatenation after every
    optimizer step.  Returns the number of groups packed.  Call after ``.cuda()`` / ``tensor_parallelize``; a later
    re-allocation of the parameters silently falls back to the cached concatenation."""
    n = 0
    for mod in model.modules():
        if not isinstance(mod, (B200AttentionMixin, B200MLPMixin)):
            continue
        for key, names in PACK_GROUPS.items():
            lins = [getattr(mod, nm, None) for nm in names]
            if any(not isinstance(lin, nn.Linear) or lin.bias is not None or hasattr(lin.weight, "to_local") for lin in lins):
                continue
            ws = [lin.weight for lin in lins]
            if len({(w.dtype, w.device, w.shape[1]) for w in ws}) != 1:
                continue
            with torch.no_grad():
                buf = torch.cat([w.detach() for w in ws], dim=0).contiguous()
                off = 0
                for w in ws:
                    w.data = buf[off:off + w.shape[0]]
                    off += w.shape[0]
            mod.__dict__.setdefault("_b200_packed", {})[key] = buf
            mod.__dict__.get("_b200_fused", {}).pop(key, None)
            n += 1
    return n


def unpack_weights(model: nn.Module) -> None:
    """Undo ``pack_weights`` (every parameter gets its own storage again), e.g. before a safetensors export that rejects
    tensors sharing storage."""
    for mod in model.modules():
        packed = mod.__dict__.pop("_b200_packed", None)
        if not packed:
            continue
        for key in packed:
            for nm in PACK_GROUPS[key]:
                w = getattr(mod, nm).weight
                w.data = w.data.clone()


def _local(p: torch.Tensor) -> torch.Tensor:
    """Tensor-parallel parameters are DTensors whose local shard is the operand ([out/tp, in] colwise, [out, in/tp]
    rowwise: distributed/tensor_parallel.py:161-334)."""
    return p.to_local() if hasattr(p, "to_local") else p


def _check_no_bias(*linears):
    for lin in linears:
        if getattr(lin, "bias", None) is not None:
            raise B200Error("transformers_b200: projection biases are not supported on the fused path")


def mask_to_kv_ranges(attention_mask: torch.Tensor | None):
    """2-D padding mask [B, kv_len] (what our AttentionMaskInterface entry returns) -> int32 (kv_start, kv_end).
    Padding must be contiguous on the left and/or right of each row (what tokenizers produce); a mask with interior zeros
    (concatenated samples, custom masks) cannot be expressed as one range per row and raises instead of silently
    attending across the holes.  The result is cached on the mask tensor object: the reference hands the same object to
    every decoder layer of a forward, so the validation (one host sync) runs once per forward, not once per layer."""
    if attention_mask is None:
        return None, None
    if attention_mask.dim() != 2:
        raise B200Error(f"b200 attention expects a 2-D padding mask or None, got shape {tuple(attention_mask.shape)}")
    hit = attention_mask.__dict__.get("_b200_ranges")
    if hit is not None and hit[0] == attention_mask._version:
        return hit[1], hit[2]
    m = attention_mask.to(torch.bool)
    L = m.shape[1]
    idx = torch.arange(L, device=m.device)
    start = torch.where(m, idx, L).amin(dim=1)
    end = torch.where(m, idx, -1).amax(dim=1) + 1
    count = m.sum(dim=1)
    if bool((count != (end - start).clamp(min=0)).any()):
        raise B200Error("b200 attention: the padding mask has zeros between valid tokens; only left / right padding is "
                        "supported (packed batches are recognised through position_ids or cu_seq_lens_*)")
    empty = count == 0  # fully masked row: an empty range (the row's output is zero, as for the un-padding flash path)
    start = torch.where(empty, torch.zeros_like(start), start).to(torch.int32).contiguous()
    end = torch.where(empty, torch.zeros_like(end), end).to(torch.int32).contiguous()
    attention_mask.__dict__["_b200_ranges"] = (attention_mask._version, start, end)
    return start, end


class SegmentIds:
    """What our AttentionMaskInterface entry returns for a padding-free packed batch (several sequences concatenated along
    the sequence axis, recognised by the reference from position_ids that restart: masking_utils.py:728-757,973-974):
    the [B, S] tensor of sequence indices plus, lazily, the per-row (start, end) token ranges on the host.  It travels
    through the decoder layers in the ``attention_mask`` slot, like the 2-D mask of the flash backends."""

    def __init__(self, ids: torch.Tensor | None = None, ranges: list | None = None):
        self.ids, self._ranges = ids, ranges

    @classmethod
    def from_cu_seqlens(cls, cu, total: int):
        """cu_seq_lens_q of a flattened batch (modeling_flash_attention_utils.py:570-590): [n + 1] cumulative lengths."""
        c = [int(v) for v in (cu.tolist() if torch.is_tensor(cu) else cu)]
        if len(c) < 2 or c[0] != 0 or c[-1] != total or any(b < a for a, b in zip(c, c[1:])):
            raise B200Error(f"b200 attention: cu_seq_lens {c} do not partition the {total} tokens of the batch")
        return cls(None, [[(a, b) for a, b in zip(c, c[1:]) if b > a]])

    def ranges(self) -> list:
        """[[(start, end), ...] per batch row]; one host sync, once per forward (cached)."""
        if self._ranges is None:
            ids = self.ids
            B, S = ids.shape
            cuts = (ids[:, 1:] != ids[:, :-1]).nonzero().tolist()
            rows = [[0] for _ in range(B)]
            for b, i in cuts:
                rows[b].append(i + 1)
            self._ranges = [[(a, e) for a, e in zip(r, r[1:] + [S])] for r in rows]
        return self._ranges


def mask_info(attention_mask, kwargs=None, total_tokens=None):
    """What the attention kernels need from the ``attention_mask`` slot and the flash-style kwargs:
    (kv_start, kv_end, segments) -- per-row padding ranges (2-D bool mask), or the packed-sequence ranges."""
    if isinstance(attention_mask, SegmentIds):
        return None, None, attention_mask.ranges()
    cu_q = (kwargs or {}).get("cu_seq_lens_q")
    if cu_q is not None:
        cu_k = kwargs.get("cu_seq_lens_k")
        if attention_mask is not None:
            raise B200Error("b200 attention: cu_seq_lens_* together with a padding mask is not supported")
        seg = cu_q.__dict__.get("_b200_segments") if torch.is_tensor(cu_q) else None
        if seg is None:
            if cu_k is not None and cu_k is not cu_q and not torch.equal(torch.as_tensor(cu_k), torch.as_tensor(cu_q)):
                raise B200Error("b200 attention: cu_seq_lens_q and cu_seq_lens_k differ (cross-length packing is not supported)")
            seg = SegmentIds.from_cu_seqlens(cu_q, total_tokens)
            if torch.is_tensor(cu_q):
                cu_q.__dict__["_b200_segments"] = seg  # the same tensor reaches every layer: one sync per forward
        return None, None, seg.ranges()
    kv_start, kv_end = mask_to_kv_ranges(attention_mask)
    return kv_start, kv_end, None


# ------------------------------------------------------------------------------------------------------------ mixins
class B200RMSNormMixin:
    _b200_gemma = False

    def forward(self, hidden_states):  # LlamaRMSNorm.forward models/llama/modeling_llama.py:62-67
        if not _on_b200(hidden_states):
            return super().forward(hidden_states)
        eps = getattr(self, "variance_epsilon", None)
        if eps is None:
            eps = self.eps
        return Fn.RMSNormFn.apply(hidden_states, _local(self.weight), eps, self._b200_gemma)


_ROPE_FWD = None


def _rope_forward():
    """The kernel-path body of the rotary module's forward, wrapped like the reference's own (no_grad + dynamic_rope_update,
    which re-derives ``inv_freq`` for the dynamic / longrope types before the tables are computed:
    modeling_rope_utils.py)."""
    global _ROPE_FWD
    if _ROPE_FWD is None:
        from transformers.modeling_rope_utils import dynamic_rope_update

        from . import ops

        @torch.no_grad()
        @dynamic_rope_update
        def fwd(self, x, position_ids):
            return ops.rope_table(self.inv_freq, position_ids, float(self.attention_scaling), x.dtype)

        _ROPE_FWD = fwd
    return _ROPE_FWD


class B200RotaryEmbeddingMixin:
    def forward(self, x, position_ids, *args, **kwargs):  # LlamaRotaryEmbedding.forward models/llama/modeling_llama.py:113-127
        if (args or kwargs or not _on_b200(x) or x.dtype not in KERNEL_DTYPES or position_ids.dim() != 2
                or not _on_b200(self.inv_freq)):
            return super().forward(x, position_ids, *args, **kwargs)
        return _rope_forward()(self, x, position_ids)


class B200MLPMixin:
    def forward(self, x):  # LlamaMLP.forward models/llama/modeling_llama.py:174-176
        if not _on_b200(x):
            return _tp_allreduce(self, super().forward(_tp_copy(self, x)))
        col, row, _ = _tp_modes(self)
        _check_no_bias(self.gate_proj, self.up_proj, self.down_proj)
        wg, wu, wd = _local(self.gate_proj.weight), _local(self.up_proj.weight), _local(self.down_proj.weight)
        act = getattr(self.config, "hidden_act", None) or getattr(self.config, "hidden_activation", "silu")
        if act not in ("silu", "gelu_pytorch_tanh"):
            raise B200Error(f"transformers_b200: activation {act} not supported")
        gelu = act == "gelu_pytorch_tanh"
        sp = _sp_state(self)
        tokens = (sp.full_shape[0] * sp.full_shape[1]) if sp is not None else x.numel() // x.shape[-1]
        from . import ops

        if self.__dict__.get("_b200_fuse_glu", True) and ops.glu_fusable(tokens, wg.shape[0]) and x.dtype in KERNEL_DTYPES:
            # gate|up GEMM with the gated activation in its epilogue: one kernel, the projections are never read back
            h = Fn.GateUpGluFn.apply(x, interleaved_weight(self, wg, wu), gelu, col, wg, wu)
        else:
            gu = Fn.FusedLinearFn.apply(x, fused_weight(self, "gate_up", [wg, wu]), col, wg, wu)
            h = Fn.GluFn.apply(gu, gelu)
        return Fn.FusedLinearFn.apply(h, fused_weight(self, "down", [wd]), row, wd)  # row mode: all-reduce inside, overlapped


class B200AttentionMixin:
    """LlamaAttention.forward models/llama/modeling_llama.py:243-281 (Mistral :141-178, Gemma2 :248-288)."""

    def _b200_window(self):
        if hasattr(self, "sliding_window"):  # Gemma2: per-layer attribute (None on full-attention layers)
            return self.sliding_window
        return getattr(self.config, "sliding_window", None)  # Mistral (modeling_mistral.py:172); Llama: None

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        if self.config._attn_implementation != ATTN_NAME or not _on_b200(hidden_states):
            out, w = super().forward(_tp_copy(self, hidden_states), position_embeddings=position_embeddings,
                                     attention_mask=attention_mask, past_key_values=past_key_values, **kwargs)
            return _tp_allreduce(self, out), w
        col, row, sp = _tp_modes(self)
        _check_no_bias(self.q_proj, self.k_proj, self.v_proj, self.o_proj)
        if self.training and getattr(self, "attention_dropout", 0.0):
            raise B200Error("transformers_b200: attention dropout is not supported")
        wq, wk, wv, wo = (_local(self.q_proj.weight), _local(self.k_proj.weight), _local(self.v_proj.weight),
                          _local(self.o_proj.weight))
        D = self.head_dim
        Hq, Hkv = wq.shape[0] // D, wk.shape[0] // D  # local head counts (module is head-count agnostic under TP)
        # sequence parallel: hidden_states is this rank's token shard [1, T/N, H]; the all-gather happens inside the fused
        # Function and (B, S) are those of the unsharded batch
        B, S = sp.full_shape[:2] if sp is not None else hidden_states.shape[:2]
        cos, sin = position_embeddings
        window = self._b200_window() or 0
        softcap = getattr(self, "attn_logit_softcapping", None) or 0.0
        wqkv = fused_weight(self, "qkv", [wq, wk, wv])
        if kwargs.get("s_aux") is not None:
            raise B200Error("b200 attention: attention sinks (s_aux) are not supported")
        if past_key_values is None:
            # padding ranges, or the sequence ranges of a padding-free packed batch (position_ids that restart reach us as
            # SegmentIds through our mask entry; collators that flatten pass cu_seq_lens_* instead)
            kv_start, kv_end, segments = mask_info(attention_mask, kwargs, B * S)
            if segments is not None and len(segments) != B:
                raise B200Error("b200 attention: cu_seq_lens_* describe a flattened batch (batch size 1)")
            cfg = (Hq, Hkv, D, float(self.scaling), True, int(window), float(softcap))
            attn = Fn.QKVRopeAttentionFn.apply(hidden_states, wqkv, cos, sin, cfg, kv_start, kv_end, col, segments, wq, wk, wv)
        else:
            from .integration import b200_attention_forward

            qkv = Fn.QKVRopeFn.apply(hidden_states, wqkv, cos, sin, Hq + Hkv, D, col, wq, wk, wv)
            q = qkv[..., : Hq * D].view(B, S, Hq, D).transpose(1, 2)
            k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D).transpose(1, 2)
            v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D).transpose(1, 2)
            k, v = past_key_values.update(k, v, self.layer_idx)
            attn, _ = b200_attention_forward(self, q, k, v, attention_mask, dropout=0.0, scaling=self.scaling,
                                             sliding_window=window or None, softcap=softcap or None, **kwargs)
            attn = attn.reshape(B, S, Hq * D)
        out = Fn.FusedLinearFn.apply(attn, fused_weight(self, "o", [wo]), row, wo)  # row mode: all-reduce inside, overlapped
        return out, None


class B200DecoderLayerMixin:
    """LlamaDecoderLayer / MistralDecoderLayer.forward (models/llama/modeling_llama.py:295-324) with the two residual adds on
    our kernels: the first fused with the post-attention RMSNorm (one pass over the row instead of two), the second a plain
    vectorised add.  Only installed by ``accelerate(model, fuse_residual=True)``; anything unexpected defers to the stock
    forward."""

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_values=None, use_cache=False,
                position_embeddings=None, **kwargs):
        norm2 = self.post_attention_layernorm
        if not _on_b200(hidden_states) or not isinstance(norm2, B200RMSNormMixin) or hidden_states.dtype not in KERNEL_DTYPES:
            return super().forward(hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                                   past_key_values=past_key_values, use_cache=use_cache,
                                   position_embeddings=position_embeddings, **kwargs)
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                                          past_key_values=past_key_values, use_cache=use_cache,
                                          position_embeddings=position_embeddings, **kwargs)
        eps = getattr(norm2, "variance_epsilon", None)
        if eps is None:
            eps = norm2.eps
        residual, hidden_states = Fn.AddRMSNormFn.apply(hidden_states, residual, _local(norm2.weight), eps, norm2._b200_gemma)
        hidden_states = self.mlp(hidden_states)
        return Fn.AddFn.apply(residual, hidden_states)


class B200EmbeddingMixin:
    def forward(self, input_ids):  # nn.Embedding.forward; Gemma2TextScaledWordEmbedding models/gemma2/modeling_gemma2.py:348
        w = _local(self.weight)
        if not _on_b200(w):
            return super().forward(input_ids)
        scale = None
        if hasattr(self, "embed_scale"):
            scale = float(self.embed_scale.to(w.dtype))  # bf16(sqrt(hidden)) as the reference rounds it
        return Fn.EmbeddingFn.apply(input_ids, w, self.padding_idx, scale)


class B200LinearMixin:
    """nn.Linear without bias on our GEMM (used for lm_head, models/llama/modeling_llama.py:480)."""

    def forward(self, x):
        w = _local(self.weight)
        gather = self.__dict__.get("_b200_tp_gather", False)
        if (self.__dict__.get("_b200_lazy_logits", False) and not gather and _on_b200(w) and self.bias is None
                and x.dtype in KERNEL_DTYPES and torch.is_grad_enabled()):
            # training forward with labels and our loss (integration._install_fused_head_loss): skip the [T, V] logits; the
            # loss function finds hidden states and weight on the placeholder and runs functional.FusedHeadLossFn
            y = x.new_empty(*x.shape[:-1], 0)
            y._b200_lazy_head = (x, fused_weight(self, "w", [w]), self.weight)
            return y
        if not _on_b200(w) or self.bias is not None:
            y = super().forward(_tp_copy(self, x) if gather else x)
        else:
            col = (self.__dict__["_b200_tp_group"], "col") if gather else None
            y = Fn.FusedLinearFn.apply(x, fused_weight(self, "w", [w]), col, w)
        if gather and self.__dict__.get("_b200_keep_vocab_shard", False):
            y._b200_vocab_shard = self.__dict__["_b200_tp_group"]  # consumed by integration.b200_causal_lm_loss
            return y
        if gather:
            from .parallel import gather_last_dim

            y = gather_last_dim(y, self.__dict__["_b200_tp_group"])
        return y


def _sp_state(module):
    st = module.__dict__.get("_b200_sp")
    return st if st is not None and st.active else None


def _tp_group(module):
    """Process group this block reduces over: set by parallel.tensor_parallelize, or -- when the model was sharded through the
    reference's own dispatch with the styles of tp_styles.py -- found on the block's rowwise Linear."""
    group = module.__dict__.get("_b200_tp_group")
    if group is None:
        for child in ("o_proj", "down_proj"):
            lin = module.__dict__.get("_modules", {}).get(child)
            if lin is not None and "_b200_tp_group" in lin.__dict__:
                group = module.__dict__["_b200_tp_group"] = lin.__dict__["_b200_tp_group"]
                break
    return group


def _tp_modes(module):
    """(col, row, sp) descriptors for FusedLinearFn / QKVRopeAttentionFn: ``(group, mode[, sp_state])`` or None."""
    group = _tp_group(module)
    if group is None:
        return None, None, None
    sp = _sp_state(module)
    if sp is not None:
        return (group, "col_sp", sp), (group, "row_sp", sp), sp
    return (group, "col"), (group, "row"), None


def _tp_copy(module, x):
    group = _tp_group(module)
    if group is None:
        return x
    from . import parallel

    sp = _sp_state(module)
    return parallel.gather_tokens(x, sp) if sp is not None else parallel.copy_to_group(x, group)


def _tp_allreduce(module, out):
    group = _tp_group(module)
    if group is None:
        return out
    from . import parallel

    sp = _sp_state(module)
    return parallel.reduce_scatter_tokens(out, sp) if sp is not None else parallel.all_reduce_sum(out, group)


# --------------------------------------------------------------------------------------------------- class factories
_CLASS_CACHE: dict = {}


def make_class(base: type, mixin: type, **attrs) -> type:
    """``class B200<Base>(mixin, base)`` -- created once per reference class, picklable by name lookup."""
    if issubclass(base, mixin):  # already one of ours (e.g. a module attribute left patched by apply_patches)
        return base
    key = (base, mixin, tuple(sorted(attrs.items())))
    if key not in _CLASS_CACHE:
        name = "B200" + base.__name__
        cls = type(name, (mixin, base), {"__module__": __name__, "__doc__": f"{base.__name__} on B200 kernels", **attrs})
        globals()[name] = cls
        _CLASS_CACHE[key] = cls
    return _CLASS_CACHE[key]
