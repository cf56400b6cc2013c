"""Host logic of the in-place KV cache layer on CPU (where it must defer to the reference's DynamicLayer.update) and of
make_cache(); the CUDA append itself is covered by tests/test_kernels_gpu.py::test_kv_append_bit_exact_vs_reference_cat."""
import torch

from _hf import import_transformers

tf = import_transformers()


def test_layer_defers_to_reference_update_on_cpu():
    from transformers.cache_utils import DynamicLayer

    from transformers_b200.cache import layer_class

    cls = layer_class()
    assert issubclass(cls, DynamicLayer)
    a, b = cls(), DynamicLayer()
    g = torch.Generator().manual_seed(0)
    for q in (3, 1, 1):
        k, v = torch.randn(2, 2, q, 8, generator=g), torch.randn(2, 2, q, 8, generator=g)
        ka, va = a.update(k, v)
        kb, vb = b.update(k, v)
        assert torch.equal(ka, kb) and torch.equal(va, vb)
        assert a.get_seq_length() == b.get_seq_length() and a.get_mask_sizes(1) == b.get_mask_sizes(1)
    # bookkeeping follows the reference on the deferred path too (lengths stayed 0, crop / reorder were skipped)
    assert a.get_seq_length() == 5
    idx = torch.tensor([1, 0])
    for op in (lambda l: l.crop(-2), lambda l: l.reorder_cache(idx), lambda l: l.batch_repeat_interleave(2),
               lambda l: l.batch_select_indices(torch.tensor([0, 3])), lambda l: l.crop(0)):
        op(a)
        op(b)
        assert a.get_seq_length() == b.get_seq_length() and torch.equal(a.keys, b.keys) and torch.equal(a.values, b.values)
    a.reset()
    b.reset()
    assert a.get_seq_length() == b.get_seq_length()


def test_make_cache_replaces_only_full_attention_layers():
    from transformers.cache_utils import DynamicLayer, DynamicSlidingWindowLayer

    from transformers_b200.cache import layer_class, make_cache

    cfg = tf.LlamaConfig(vocab_size=32, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2)
    c = make_cache(cfg)
    assert len(c.layers) == 3 and all(type(l) is layer_class() for l in c.layers)
    gcfg = tf.Gemma2Config(vocab_size=32, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                           num_key_value_heads=1, head_dim=16, sliding_window=8,
                           layer_types=["sliding_attention", "full_attention"])
    g = make_cache(gcfg)
    assert isinstance(g.layers[0], DynamicSlidingWindowLayer) and type(g.layers[1]) is layer_class()
    assert isinstance(g.layers[1], DynamicLayer)
