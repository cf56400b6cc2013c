"""transformers_b200: a Blackwell (sm_100a) forward/backward engine for decoder-only transformers that plugs in behind
huggingface/transformers' own extension points (AttentionInterface, AttentionMaskInterface, register_patch_mapping).

    import transformers_b200
    transformers_b200.enable()                       # register "b200" + module patches
    model = AutoModelForCausalLM.from_config(cfg, attn_implementation="b200", dtype=torch.bfloat16).cuda()
    transformers_b200.accelerate(model)              # embedding gather, lm_head GEMM, fused causal-LM loss
    from transformers_b200.optim import B200AdamW    # Trainer(optimizer_cls_and_kwargs=(B200AdamW, {...})): fused multi-tensor step

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot op is a hand-written CUDA kernel
reached through the C-ABI in ``include/b200_ops.h`` (``transformers_b200/lib/libb200.so``).  No CPU fallback exists.
"""
from ._lib import B200Error, LIB_PATH  # noqa: F401

__version__ = "0.1.0"


def enable(*args, **kwargs):
    from .integration import enable as _enable

    return _enable(*args, **kwargs)


def accelerate(model, *args, **kwargs):
    from .integration import accelerate as _accelerate

    return _accelerate(model, *args, **kwargs)


def pack_weights(model):
    from .modules import pack_weights as _pack

    return _pack(model)


def unpack_weights(model):
    from .modules import unpack_weights as _unpack

    return _unpack(model)
