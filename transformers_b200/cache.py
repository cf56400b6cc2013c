"""KV cache with in-place append behind the reference's cache protocol.

``B200DynamicLayer`` subclasses the reference's ``DynamicLayer`` (cache_utils.py:113-200): same ``update`` contract
(returns the *full* K/V to attend over, [B, Hkv, ctx, D]), same bookkeeping methods, but K/V live in a preallocated
[B, Hkv, capacity, D] buffer that grows geometrically, and ``update`` launches one ``b200_kv_append`` kernel that writes
only the new rows -- no O(context) ``torch.cat`` per token.  The returned tensors are strided views of the buffer; the
b200 attention kernel reads them through strided TMA descriptors, so nothing is copied.

    cache = transformers_b200.cache.make_cache(model.config)          # sliding layers keep the reference layer type
    model.generate(ids, past_key_values=cache, ...)                   # generation/utils.py:1945-1957 accepts user caches
"""
from __future__ import annotations

import torch

from . import ops


def _layer_base():
    from transformers.cache_utils import DynamicLayer

    return DynamicLayer


def _make_layer_class():
    DynamicLayer = _layer_base()

    class B200DynamicLayer(DynamicLayer):
        """DynamicLayer (cache_utils.py:113) with a preallocated buffer and an in-place append kernel."""

        min_capacity = 256

        def lazy_initialization(self, key_states, value_states):
            self.dtype, self.device = key_states.dtype, key_states.device
            self._len = 0
            self._buf_k = self._buf_v = None
            self.keys = torch.tensor([], dtype=self.dtype, device=self.device)
            self.values = torch.tensor([], dtype=self.dtype, device=self.device)
            self.is_initialized = True

        def _reserve(self, B, H, D, need):
            cap = 0 if self._buf_k is None else self._buf_k.shape[2]
            if need <= cap and self._buf_k.shape[0] == B:
                return
            new_cap = max(self.min_capacity, cap)
            while new_cap < need:
                new_cap *= 2
            nk = torch.empty(B, H, new_cap, D, dtype=self.dtype, device=self.device)
            nv = torch.empty(B, H, new_cap, D, dtype=self.dtype, device=self.device)
            if self._len > 0:  # amortised O(1): geometric growth
                ops.kv_append(self._buf_k[:, :, : self._len], self._buf_v[:, :, : self._len], nk, nv, 0)
            self._buf_k, self._buf_v = nk, nv

        def update(self, key_states, value_states, *args, **kwargs):
            if not self.is_initialized:
                self.lazy_initialization(key_states, value_states)
            from . import modules as M

            if not M._on_b200(key_states) or key_states.dtype not in M.KERNEL_DTYPES:
                return super().update(key_states, value_states, *args, **kwargs)
            B, H, q, D = key_states.shape
            self._reserve(B, H, D, self._len + q)
            ops.kv_append(key_states, value_states, self._buf_k, self._buf_v, self._len)
            self._len += q
            self.keys = self._buf_k[:, :, : self._len]
            self.values = self._buf_v[:, :, : self._len]
            return self.keys, self.values

        def _stock(self) -> bool:
            """True while this layer is running on the reference's own ``torch.cat`` path (``update`` deferred to the base
            class: CPU tensors, or a dtype the kernels do not take): every bookkeeping method then defers too, so that
            lengths, mask sizes, crop and beam re-ordering keep following ``self.keys`` / ``self.values``."""
            return getattr(self, "_buf_k", None) is None

        def get_seq_length(self) -> int:
            if self._stock():
                return super().get_seq_length()
            return self._len if self.is_initialized else 0

        def crop(self, tokens_to_remove: int = 0, **kw) -> None:
            """cache_utils.py:166-189: negative = remove that many tokens, positive (deprecated) = absolute final length,
            0 = nothing to do."""
            if not self.is_initialized:
                return
            if self._stock():
                return super().crop(kw.get("max_length", tokens_to_remove))
            n = kw.get("max_length", tokens_to_remove)
            new_len = self._len - abs(n) if n < 0 else min(self._len, n if n > 0 else self._len)
            self._len = max(0, new_len)
            self.keys = self._buf_k[:, :, : self._len]
            self.values = self._buf_v[:, :, : self._len]

        def reset(self) -> None:
            if self._stock():
                return super().reset()
            if self.is_initialized:
                self._len = 0
                if self._buf_k is not None:
                    self.keys = self._buf_k[:, :, :0]
                    self.values = self._buf_v[:, :, :0]

        def _rebuffer(self, k, v):
            self._buf_k, self._buf_v = k.contiguous(), v.contiguous()
            self.keys = self._buf_k[:, :, : self._len]
            self.values = self._buf_v[:, :, : self._len]

        def reorder_cache(self, beam_idx) -> None:
            if self._stock():
                return super().reorder_cache(beam_idx)
            if self.get_seq_length() > 0:
                self._rebuffer(self._buf_k.index_select(0, beam_idx.to(self.device)), self._buf_v.index_select(0, beam_idx.to(self.device)))

        def batch_repeat_interleave(self, repeats: int) -> None:
            if self._stock():
                return super().batch_repeat_interleave(repeats)
            if self.get_seq_length() > 0:
                self._rebuffer(self._buf_k.repeat_interleave(repeats, dim=0), self._buf_v.repeat_interleave(repeats, dim=0))

        def batch_select_indices(self, indices) -> None:
            if self._stock():
                return super().batch_select_indices(indices)
            if self.get_seq_length() > 0:
                self._rebuffer(self._buf_k[indices, ...], self._buf_v[indices, ...])

    return B200DynamicLayer


def _make_sliding_class():
    from transformers.cache_utils import DynamicSlidingWindowLayer

    class B200SlidingWindowLayer(DynamicSlidingWindowLayer):
        """DynamicSlidingWindowLayer (cache_utils.py:203-262) without the per-token ``torch.cat`` of the whole window: K/V
        live in a buffer of at least 2 x (window - 1) rows; ``update`` appends the new rows in place (``b200_kv_append``)
        and returns views; the kept window just advances its start offset, and the last window - 1 rows are moved back to
        the front (one copy per ~window tokens, regions never overlap) when the end of the buffer is reached.
        Any manipulation the base class does on ``self.keys`` / ``self.values`` (crop, beam re-ordering, batch selection)
        is detected on the next ``update`` and the buffer is rebuilt from them."""

        def _views_are_ours(self):
            return (self._buf_k is not None and self.keys.dim() == 4 and self.keys.shape[2] == self._n
                    and self.keys.shape[0] == self._buf_k.shape[0]
                    and (self._n == 0 or self.keys.data_ptr() == self._buf_k[:, :, self._s:].data_ptr()))

        def update(self, key_states, value_states, *args, **kwargs):
            from . import modules as M

            if not self.is_initialized:
                self.lazy_initialization(key_states, value_states)
                self._buf_k = self._buf_v = None
                self._s = self._n = 0
            if not M._on_b200(key_states) or key_states.dtype not in M.KERNEL_DTYPES:
                return super().update(key_states, value_states, *args, **kwargs)
            B, H, q, D = key_states.shape
            keep_max = self.sliding_window - 1
            if not self._views_are_ours():  # first call, or the base class re-sliced / re-ordered keys and values
                old_k, old_v = self.keys, self.values
                self._n = old_k.shape[2] if old_k.dim() == 4 else 0
                self._buf_k = self._buf_v = None
                self._s = 0
            else:
                old_k = old_v = None
            need = self._n + q
            cap = 0 if self._buf_k is None else self._buf_k.shape[2]
            wrap = self._buf_k is not None and need <= cap and self._s + need > cap
            if wrap and self._s >= self._n:  # end of the buffer: move the kept rows to the front, regions do not overlap
                if self._n > 0:
                    ops.kv_append(self._buf_k[:, :, self._s:self._s + self._n], self._buf_v[:, :, self._s:self._s + self._n],
                                  self._buf_k, self._buf_v, 0)
                self._s = 0
            elif self._buf_k is None or need > cap or wrap:  # (re)allocate: room for the window twice, or for a long prefill
                new_cap = max(2 * keep_max + 1, need + keep_max, cap)
                nk = torch.empty(B, H, new_cap, D, dtype=self.dtype, device=self.device)
                nv = torch.empty(B, H, new_cap, D, dtype=self.dtype, device=self.device)
                src_k = old_k if old_k is not None else (self._buf_k[:, :, self._s:self._s + self._n] if self._buf_k is not None else None)
                src_v = old_v if old_v is not None else (self._buf_v[:, :, self._s:self._s + self._n] if self._buf_v is not None else None)
                if self._n > 0:
                    ops.kv_append(src_k, src_v, nk, nv, 0)
                self._buf_k, self._buf_v, self._s = nk, nv, 0
            self.cumulative_length += q
            ops.kv_append(key_states, value_states, self._buf_k, self._buf_v, self._s + self._n)
            full_k = self._buf_k[:, :, self._s:self._s + need]
            full_v = self._buf_v[:, :, self._s:self._s + need]
            keep = need if getattr(self, "record_past", False) else min(need, keep_max)  # record_past: newer transformers only
            self._s += need - keep
            self._n = keep
            self.keys = self._buf_k[:, :, self._s:self._s + keep]
            self.values = self._buf_v[:, :, self._s:self._s + keep]
            return full_k, full_v

    return B200SlidingWindowLayer


_LAYER_CLS = None
_SLIDING_CLS = None


def sliding_layer_class():
    global _SLIDING_CLS
    if _SLIDING_CLS is None:
        _SLIDING_CLS = _make_sliding_class()
    return _SLIDING_CLS


def layer_class():
    global _LAYER_CLS
    if _LAYER_CLS is None:
        _LAYER_CLS = _make_layer_class()
    return _LAYER_CLS


def make_cache(config, inplace_sliding: bool = True):
    """DynamicCache(config=...) (cache_utils.py:1773-1815) with every full-attention layer replaced by B200DynamicLayer and
    every sliding-window layer by B200SlidingWindowLayer (same semantics as DynamicSlidingWindowLayer without the per-token
    copy of the window; ``inplace_sliding=False`` keeps the reference layer type for those)."""
    from transformers.cache_utils import DynamicCache, DynamicLayer, DynamicSlidingWindowLayer

    cache = DynamicCache(config=config)
    cls = layer_class()
    n = config.get_text_config(decoder=True).num_hidden_layers
    if len(cache.layers) == 0:
        cache.layers = [cls() for _ in range(n)]
    else:
        cache.layers = [cls() if type(l) is DynamicLayer else l for l in cache.layers]
    if inplace_sliding:
        scls = sliding_layer_class()
        cache.layers = [scls(sliding_window=l.sliding_window) if type(l) is DynamicSlidingWindowLayer else l for l in cache.layers]
    return cache
