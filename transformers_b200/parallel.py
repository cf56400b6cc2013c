"""Tensor parallelism by the reference's own ``tp_plan`` (models/llama/configuration_llama.py:49-57:
q/k/v/gate/up ``colwise``, o/down ``rowwise``; ``lm_head`` ``colwise_gather_output`` models/llama/modeling_llama.py:423).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the transport.
Unlike the reference (DTensor hooks around every nn.Linear, distributed/tensor_parallel.py:147-334) the shards are plain
local tensors: our fused modules consume them directly and issue the two collectives per block themselves
(Megatron-style f / g operators):

    x --copy_to_group--> [colwise GEMMs -> ... -> rowwise GEMM] --all_reduce_sum--> y
          (bwd: all-reduce dX)                                      (bwd: identity)
"""
from __future__ import annotations

import fnmatch

import torch
import torch.distributed as dist
from torch import nn


class _CopyToGroup(torch.autograd.Function):
    """Identity forward; backward all-reduces the input gradient of the colwise layers
    (C2 in SURVEY.md §2.2: distributed/tensor_parallel.py:219-226)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, group=ctx.group)
        return g, None


class _ReduceFromGroup(torch.autograd.Function):
    """All-reduce(sum) of the rowwise partial outputs forward (C1: distributed/tensor_parallel.py:320-328); identity backward."""

    @staticmethod
    def forward(ctx, x, group):
        if x.is_contiguous():
            ctx.mark_dirty(x)  # reduce in place: x is the fresh output of the rowwise GEMM
        else:
            x = x.contiguous()
        dist.all_reduce(x, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


class _GatherLastDim(torch.autograd.Function):
    """colwise_gather_output (C3: distributed/tensor_parallel.py:239-240,748): all-gather shards on the last dim;
    backward keeps this rank's slice."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.n = group, x.shape[-1]
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=-1)

    @staticmethod
    def backward(ctx, g):
        r = dist.get_rank(ctx.group)
        return g[..., r * ctx.n:(r + 1) * ctx.n].contiguous(), None


def copy_to_group(x, group):
    return _CopyToGroup.apply(x, group)


def all_reduce_sum(x, group):
    return _ReduceFromGroup.apply(x, group)


def gather_last_dim(x, group):
    return _GatherLastDim.apply(x, group)


def _shard(param: nn.Parameter, dim: int, rank: int, world: int) -> nn.Parameter:
    if param.shape[dim] % world:
        raise ValueError(f"cannot shard dim {dim} of {tuple(param.shape)} over {world} ranks")
    piece = param.detach().chunk(world, dim=dim)[rank].contiguous().clone()
    return nn.Parameter(piece, requires_grad=param.requires_grad)


def resolve_plan(model) -> dict:
    """{module-name pattern: style} from the model's own config (base_model_tp_plan) + class-level _tp_plan."""
    plan = {}
    base = getattr(model.config, "base_model_tp_plan", None) or {}
    prefix = getattr(model, "base_model_prefix", "model")
    has_prefix = hasattr(model, prefix)
    for k, v in base.items():
        plan[(prefix + "." + k) if has_prefix else k] = v
    # class-level plan (e.g. {"lm_head": "colwise_gather_output"}); some transformers versions shadow it on the instance
    # with the merged base plan, so read both
    for klass in type(model).__mro__:
        cls_plan = klass.__dict__.get("_tp_plan")
        if isinstance(cls_plan, dict):
            for k, v in cls_plan.items():
                plan.setdefault(k, v)
    inst_plan = getattr(model, "_tp_plan", None)
    if isinstance(inst_plan, dict):
        for k, v in inst_plan.items():
            plan.setdefault(k, v)
    return plan


def tensor_parallelize(model: nn.Module, group=None, plan: dict | None = None) -> nn.Module:
    """Shard an already materialised model in place (each rank keeps its slice) and tell the block modules which group
    to reduce over.  Mirrors apply_tensor_parallelism (distributed/tensor_parallel.py:773-796) for colwise / rowwise /
    colwise_gather_output; embeddings and norms stay replicated."""
    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan = plan if plan is not None else resolve_plan(model)
    kv = getattr(model.config, "num_key_value_heads", None)
    if kv is not None and kv % world:
        raise ValueError(f"num_key_value_heads={kv} is not divisible by tp size {world} (SURVEY.md §8e)")
    touched = set()
    for name, mod in model.named_modules():
        if not isinstance(mod, nn.Linear):
            continue
        style = next((s for pat, s in plan.items() if fnmatch.fnmatchcase(name, pat)), None)
        if style is None:
            continue
        if style in ("colwise", "colwise_gather_output"):
            mod.weight = _shard(mod.weight, 0, rank, world)
            if mod.bias is not None:
                mod.bias = _shard(mod.bias, 0, rank, world)
            mod.out_features = mod.weight.shape[0]
            if style == "colwise_gather_output":
                mod.__dict__["_b200_tp_group"] = group
                mod.__dict__["_b200_tp_gather"] = True
        elif style == "rowwise":
            mod.weight = _shard(mod.weight, 1, rank, world)
            mod.in_features = mod.weight.shape[1]
        else:
            raise ValueError(f"tp style {style!r} for {name} is not supported (colwise / rowwise / colwise_gather_output)")
        mod.__dict__.pop("_b200_fused", None)
        touched.add(name.rsplit(".", 1)[0])
    for name, mod in model.named_modules():
        if name in touched and hasattr(mod, "forward") and type(mod).__name__.startswith("B200"):
            mod.__dict__["_b200_tp_group"] = group
            mod.__dict__.pop("_b200_fused", None)
    model.__dict__["_b200_tp_world"] = world
    return model
