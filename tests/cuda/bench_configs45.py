"""Measurement script for BASELINE.json configs[3] and configs[4] (not bench.py lines: the contract's bench is configs[1]):

  config 4  Mixtral-8x7B bf16 forward, seq 2048, 1 x B200 (router + expert SwiGLU path)      -> forward tokens/s
  config 5  Gemma-2-9B generate(): prefill 8192 + decode 512, 1 x B200 (KV-cache append path) -> prefill and decode tokens/s

Random-init weights, synthetic ids, CUDA-event timing, one JSON line per config on stdout.  Usage on the GPU box:
  python tests/cuda/bench_configs45.py mixtral        # needs ~95 GB of HBM for the 46.7 B parameters
  python tests/cuda/bench_configs45.py gemma2 [--inplace-sliding]     (B200_GEMV=1 B200_DECODE_ATTN=1 for the decode kernels)
Written after round 1's GPU budget was spent (never run on a device yet)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from _hf import import_transformers  # noqa: E402

tf = import_transformers()
import transformers_b200  # noqa: E402
from transformers_b200 import ops  # noqa: E402

BF = torch.bfloat16


def timed(fn, warmup=1, iters=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def mixtral():
    cfg = tf.MixtralConfig(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                           num_attention_heads=32, num_key_value_heads=8, head_dim=128, num_local_experts=8,
                           num_experts_per_tok=2, max_position_embeddings=32768, sliding_window=None, rms_norm_eps=1e-5,
                           rope_parameters={"rope_type": "default", "rope_theta": 1000000.0}, use_cache=False)
    transformers_b200.enable()
    t0 = time.time()
    with torch.device("cuda"):
        model = tf.MixtralForCausalLM._from_config(cfg, attn_implementation="b200", experts_implementation="eager", dtype=BF)
    transformers_b200.accelerate(model)
    model.eval()
    S = 2048
    ids = torch.randint(0, cfg.vocab_size, (1, S), device="cuda")
    n0 = ops.launch_count()
    with torch.no_grad():
        ms = timed(lambda: model(input_ids=ids).logits)
    launches = (ops.launch_count() - n0) // 4
    active = 12.88e9  # parameters touched per token: attention + 2 of 8 experts + embeddings (Mixtral-8x7B)
    print(json.dumps({"config": "Mixtral-8x7B bf16 forward seq=2048 on 1xB200 (configs[3])", "metric": "tokens/s forward",
                      "value": S / (ms * 1e-3), "ms": ms, "gpu_launches": launches, "model_tflops": 2 * active * S / (ms * 1e-3) / 1e12,
                      "init_s": round(time.time() - t0, 1), "data": "synthetic", "dtype": "bf16"}), flush=True)


def gemma2(inplace_sliding):
    from transformers_b200.cache import make_cache

    cfg = tf.Gemma2Config(vocab_size=256000, hidden_size=3584, intermediate_size=14336, num_hidden_layers=42,
                          num_attention_heads=16, num_key_value_heads=8, head_dim=256, sliding_window=4096,
                          query_pre_attn_scalar=256, attn_logit_softcapping=50.0, final_logit_softcapping=30.0,
                          max_position_embeddings=16384, rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
    transformers_b200.enable()
    with torch.device("cuda"):
        model = tf.Gemma2ForCausalLM._from_config(cfg, attn_implementation="b200", dtype=BF)
    transformers_b200.accelerate(model)
    model.eval()
    P, Dn = 8192, 512
    ids = torch.randint(1, cfg.vocab_size, (1, P), device="cuda")

    def run(new_tokens):
        cache = make_cache(model.config, inplace_sliding=inplace_sliding)
        with torch.no_grad():
            return model.generate(ids, max_new_tokens=new_tokens, min_new_tokens=new_tokens, do_sample=False, pad_token_id=0,
                                  past_key_values=cache)

    run(2)  # warm-up (allocator, cuBLAS-free: all ours)
    ms_prefill = timed(lambda: run(1), warmup=0, iters=2)
    ms_total = timed(lambda: run(Dn), warmup=0, iters=1)
    ms_decode = max(ms_total - ms_prefill, 1e-3)
    kv_bytes = 21 * 2 * 8 * 256 * 2 * (P + Dn / 2) + 21 * 2 * 8 * 256 * 2 * 4095  # full + sliding layers, per decode step
    w_bytes = 9.24e9 * 2
    print(json.dumps({"config": "Gemma-2-9B generate(): prefill 8192 + decode 512 on 1xB200 (configs[4])",
                      "prefill_tokens_per_s": P / (ms_prefill * 1e-3), "decode_tokens_per_s": (Dn - 1) / (ms_decode * 1e-3),
                      "ms_prefill": ms_prefill, "ms_per_decode_step": ms_decode / (Dn - 1),
                      "decode_hbm_floor_ms": (kv_bytes + w_bytes) / 6.566e12 * 1e3,
                      "switches": {k: os.environ.get(k, "0") for k in ("B200_GEMV", "B200_DECODE_ATTN")},
                      "inplace_sliding": inplace_sliding, "data": "synthetic", "dtype": "bf16"}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "gemma2"
    if which == "mixtral":
        mixtral()
    else:
        gemma2("--inplace-sliding" in sys.argv)
