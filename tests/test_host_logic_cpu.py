"""Host-side logic that needs no kernels: padding-mask -> kv range conversion, fused-weight cache invalidation,
tp_plan resolution / sharding shapes, C-ABI argument validation (errno-style codes before any launch)."""
import pytest
import torch

from _hf import import_transformers

tf = import_transformers()


def test_mask_to_kv_ranges_left_right_and_full():
    from transformers_b200.modules import mask_to_kv_ranges

    assert mask_to_kv_ranges(None) == (None, None)
    m = torch.tensor([[1, 1, 1, 1, 0, 0], [0, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [0, 1, 1, 1, 0, 0]])
    s, e = mask_to_kv_ranges(m)
    assert s.tolist() == [0, 2, 0, 1] and e.tolist() == [4, 6, 6, 4]
    assert s.dtype == torch.int32 and e.dtype == torch.int32
    from transformers_b200 import B200Error

    with pytest.raises(B200Error):
        mask_to_kv_ranges(torch.ones(2, 1, 4, 4))  # 4-D (eager-style) masks are not what the b200 mask entry produces


def test_fused_weight_cache_tracks_parameter_updates():
    from transformers_b200.modules import fused_weight

    mod = torch.nn.Module()
    a = torch.nn.Parameter(torch.randn(4, 8))
    b = torch.nn.Parameter(torch.randn(2, 8))
    f1 = fused_weight(mod, "ab", [a, b])
    assert f1.shape == (6, 8) and torch.equal(f1[:4], a) and torch.equal(f1[4:], b)
    assert fused_weight(mod, "ab", [a, b]) is f1  # cache hit
    with torch.no_grad():
        a.add_(1.0)  # optimizer-style in-place update bumps _version
    f2 = fused_weight(mod, "ab", [a, b])
    assert f2 is not f1 and torch.equal(f2[:4], a)
    c = torch.nn.Parameter(torch.randn(4, 8))
    assert torch.equal(fused_weight(mod, "ab", [c, b])[:4], c)  # replaced parameter object
    assert fused_weight(mod, "single", [a]) is a.data or fused_weight(mod, "single", [a]).data_ptr() == a.data_ptr()


def test_resolve_plan_reads_the_reference_tp_plan():
    from transformers_b200.parallel import resolve_plan

    cfg = tf.LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=8)
    model = tf.LlamaForCausalLM._from_config(cfg, attn_implementation="eager")
    plan = resolve_plan(model)
    assert plan["model.layers.*.self_attn.q_proj"] == "colwise" and plan["model.layers.*.self_attn.o_proj"] == "rowwise"
    assert plan["model.layers.*.mlp.down_proj"] == "rowwise" and plan["lm_head"] == "colwise_gather_output"


def test_c_abi_validates_arguments_without_a_gpu():
    from transformers_b200 import _lib

    lib = _lib.load()
    assert lib.b200_rmsnorm_fwd(None, None, None, None, None, None, 4, 7, 1e-5, 0, None) == -22  # H % 8
    assert "multiple of 8" in _lib.last_error()
    assert lib.b200_rope(None, None, None, 1, 4, 2, 24, 64, 1, 0, None) == -22  # head_dim % 16
    assert lib.b200_kv_append(None, None, None, None, 1, 1, 4, 64, 8, 8, 8, 8, 8, 8, 8, 8, 8, 10, 12, None) == -22  # over capacity
    assert lib.b200_attn_fwd(None, None, None, None, None, 128, 1, 8, 8, 4, 3, 128, *([8] * 12), 1.0, 0.0, 1, 0, None, None, None) == -22
    assert "multiple of Hkv" in _lib.last_error()
    assert lib.b200_moe_route(None, None, None, None, None, None, 4, 2, 0, None) == -22


def test_bench_reference_arm_times_stock_layer_even_after_enable(monkeypatch):
    """bench.py's CPU arm must run the reference library's own LlamaDecoderLayer (no B200 class in it), also in a process
    where the plugin is enabled and a patched model has been built (the cpu_baseline leg of the b200 arm)."""
    import bench
    import transformers_b200

    transformers = tf
    transformers_b200.enable()
    tiny = dict(bench.LLAMA3_8B, vocab_size=64, hidden_size=64, intermediate_size=128, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, num_hidden_layers=2)
    patched = transformers.LlamaForCausalLM._from_config(transformers.LlamaConfig(**tiny), attn_implementation="eager")
    assert type(patched.model.layers[0].mlp).__name__ == "B200LlamaMLP"
    monkeypatch.setattr(bench, "LLAMA3_8B", tiny)
    got = bench.cpu_reference_sample_stock(seq=16, repeats=0)
    assert got is not None and got[0] > 0 and got[2] == transformers.__version__
    t, threads, kind, what = bench.cpu_sample(seq=16, repeats=0)
    assert kind == "reference" and "stock transformers" in what
    monkeypatch.setenv("B200_BENCH_ORACLE_BASELINE", "1")
    assert bench.cpu_sample(seq=16, repeats=0)[2] == "port"


def test_bench_reference_arm_protocol_on_a_tiny_config(monkeypatch, capsys):
    """`bench.py --impl reference`: 1- and 2-layer stock models through the public API, per-layer time by
    difference, one JSON line with impl / cpu_baseline / e2e; the model it times contains no B200 class even though the
    plugin was enabled earlier in this process."""
    import argparse
    import json

    import bench
    import transformers_b200
    import transformers_b200.integration as integ

    transformers_b200.enable()
    tiny = dict(bench.LLAMA3_8B, vocab_size=64, hidden_size=64, intermediate_size=128, num_attention_heads=4,
                num_key_value_heads=2, head_dim=16, num_hidden_layers=3)
    monkeypatch.setattr(bench, "LLAMA3_8B", tiny)
    try:
        bench.run_reference(argparse.Namespace(seq=32, batch=4, steps=2, warmup=1, gpus=1))
    finally:
        integ._enabled = False  # the arm clears the patch mapping (it must time stock classes): re-register for later tests
        transformers_b200.enable()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference" and line["value"] > 0
    assert "1-layer model" in line["cpu_baseline"]["sample"] and line["e2e"]["value"] == line["value"]
    assert abs(line["ms_per_step"] - 4 * 32 / line["value"] * 1e3) < 1e-6 * line["ms_per_step"] + 1e-9


def test_c_abi_validates_new_entry_points_without_a_gpu():
    import ctypes

    from transformers_b200 import _lib

    lib = _lib.load()
    assert lib.b200_gemv_bf16(None, None, None, 5, 64, 64, 64, 64, 64, None) == -22 and "1..4" in _lib.last_error()
    assert lib.b200_gemv_bf16(None, None, None, 1, 64, 60, 64, 64, 64, None) == -22  # K % 8
    ptrs = (ctypes.c_void_p * 3)(16, 32, 48)
    assert lib.b200_pull_reduce_bf16(ctypes.cast(ptrs, ctypes.c_void_p), 3, 0, 64, None, ctypes.c_void_p(64), None) == -22
    assert "not instantiated" in _lib.last_error()
    assert lib.b200_pull_reduce_bf16(ctypes.cast(ptrs, ctypes.c_void_p), 2, 4, 64, None, ctypes.c_void_p(64), None) == -22  # offset % 8
    assert lib.b200_adamw_step(None, None, 1, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0, 1.0, None, None) == -22  # step 0: bc1 == 0
    assert "bias corrections" in _lib.last_error()
    assert lib.b200_adamw_step(None, None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.03, None, None) == 0  # empty group: no launch
    assert lib.b200_grad_scale(None, None, -1, None, None) == -22
    assert lib.b200_ce_bwd_sharded(None, None, None, None, None, 4, 64, 60, 64, None) == -22  # ld % 8
    assert lib.b200_optim_chunk_elems() == 32768


def test_deferred_embedding_range_flag():
    """Out-of-range token ids are reported at the next call, once the kernel that raised the flag has finished, without a
    host sync on the hot path (the flag was written but never read)."""
    from transformers_b200 import B200Error
    from transformers_b200.ops import _DeferredFlag

    class Ev:
        def __init__(self, done):
            self.done = done

        def query(self):
            return self.done

        def synchronize(self):
            self.done = True

    d = _DeferredFlag()
    busy = Ev(False)
    d.push(torch.ones(1, dtype=torch.int32), busy)   # bad ids, kernel still running: nothing to read yet
    d.poll()
    assert len(d.pending) == 1
    busy.done = True
    with pytest.raises(B200Error):
        d.poll()
    assert d.pending == []
    d.push(torch.zeros(1, dtype=torch.int32), Ev(True))
    d.poll()
    assert d.pending == []
    d.push(torch.ones(1, dtype=torch.int32), Ev(False))
    with pytest.raises(B200Error):
        d.poll(force=True)
