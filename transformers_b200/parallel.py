"""Tensor parallelism by the reference's own ``tp_plan`` (models/llama/configuration_llama.py:49-57:
q/k/v/gate/up ``colwise``, o/down ``rowwise``; ``lm_head`` ``colwise_gather_output`` models/llama/modeling_llama.py:423).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the transport.
Unlike the reference (DTensor hooks around every nn.Linear, distributed/tensor_parallel.py:147-334) the shards are plain
local tensors: our fused modules consume them directly and issue the two collectives per block themselves
(Megatron-style f / g operators):

    x --copy_to_group--> [colwise GEMMs -> ... -> rowwise GEMM] --all_reduce_sum--> y
          (bwd: all-reduce dX)                                      (bwd: identity)

``tensor_parallelize(..., sequence_parallel=True)`` keeps the same weight shards but splits the two all-reduces of a block
into all-gather (before the colwise GEMMs) + reduce-scatter (after the rowwise GEMM): the hidden states BETWEEN the blocks
-- residual stream, RMSNorms, residual adds, their gradients -- then live token-sharded ([1, T/N, H] per rank) instead of
replicated, which removes the replicated element-wise work that caps plain TP scaling.  Same bytes on the wire, same
results; the reference has the building block (``SequenceParallel`` style, distributed/tensor_parallel.py) but Llama's
plan does not use it.  See ``SequenceParallelState`` for the token layout.
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


# ------------------------------------------------------------------------------------------------ sequence parallel
class SequenceParallelState:
    """Shared by every module of one model.  Token layout: the T = B*S tokens (flattened) are cut into ``chunks`` equal
    row blocks; inside each block rank r owns the r-th of ``world`` equal pieces; a rank's shard is the concatenation of its
    pieces ([T/world, H]).  Chunking lets the collectives of one block pipeline against the GEMM of the other
    (functional.py): all-gather(c+1) runs under GEMM(c), reduce-scatter(c) under GEMM(c+1).

    ``active`` / ``full_shape`` describe the forward currently in flight; the first decoder layer's pre-hook sets them
    (sequence parallelism is skipped for KV-cache forwards and token counts that do not divide)."""

    def __init__(self, group, chunks: int = 2, peer=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.chunks = int(chunks)
        self.active = False
        self.full_shape = None
        # symm.PeerWorkspace: the kernel path then runs its all-gathers / reduce-scatters itself over NVLink peer memory
        # (functional._peer_*) instead of NCCL; the layout is then one row block per rank (chunks == 1)
        self.peer = peer
        if peer is not None and self.chunks != 1:
            raise ValueError("the peer-memory transport uses the plain one-block-per-rank token layout (chunks=1)")

    def usable(self, tokens: int) -> bool:
        return tokens > 0 and tokens % (self.world * self.chunks) == 0

    def chunk_rows(self, rows: int):
        """Row slices of the ``chunks`` blocks of a [rows, *] tensor (full: rows = T; shard: rows = T/world)."""
        n = rows // self.chunks
        return [slice(c * n, (c + 1) * n) for c in range(self.chunks)]


def sp_all_gather(local2: torch.Tensor, st: SequenceParallelState):
    """[T/N, H] shard -> ([T, H] full, [async work per chunk]); wait on work c before touching chunk c."""
    rows, H = local2.shape
    full = local2.new_empty(rows * st.world, H)
    works = [dist.all_gather_into_tensor(full[fr], local2[lr], group=st.group, async_op=True)
             for fr, lr in zip(st.chunk_rows(rows * st.world), st.chunk_rows(rows))]
    return full, works


def sp_reduce_scatter_chunk(full_chunk: torch.Tensor, local_out: torch.Tensor, c: int, st: SequenceParallelState):
    """Sum chunk ``c`` ([T/chunks, H], a partial sum on every rank) over the group into this rank's piece of ``local_out``."""
    return dist.reduce_scatter_tensor(local_out[st.chunk_rows(local_out.shape[0])[c]], full_chunk, group=st.group, async_op=True)


def sp_take_local(full2: torch.Tensor, st: SequenceParallelState) -> torch.Tensor:
    rows = full2.shape[0]
    n = rows // st.chunks // st.world
    return torch.cat([full2[fr][st.rank * n:(st.rank + 1) * n] for fr in st.chunk_rows(rows)], dim=0)


def _sp_gather_now(local: torch.Tensor, st: SequenceParallelState) -> torch.Tensor:
    full, works = sp_all_gather(local.reshape(-1, local.shape[-1]).contiguous(), st)
    for w in works:
        w.wait()
    return full


def _sp_scatter_now(full: torch.Tensor, st: SequenceParallelState) -> torch.Tensor:
    full2 = full.reshape(-1, full.shape[-1]).contiguous()
    out = full2.new_empty(full2.shape[0] // st.world, full2.shape[1])
    works = [sp_reduce_scatter_chunk(full2[fr], out, c, st) for c, fr in enumerate(st.chunk_rows(full2.shape[0]))]
    for w in works:
        w.wait()
    return out


class _ShardTokens(torch.autograd.Function):
    """[B, S, H] replicated -> [1, T/N, H] shard (entry of the decoder stack); backward all-gathers the shard gradients
    so the embedding gradient stays complete on every rank."""

    @staticmethod
    def forward(ctx, x, st):
        ctx.st, ctx.shape = st, x.shape
        return sp_take_local(x.reshape(-1, x.shape[-1]), st).unsqueeze(0)

    @staticmethod
    def backward(ctx, g):
        return _sp_gather_now(g, ctx.st).view(ctx.shape), None


class _GatherTokens(torch.autograd.Function):
    """[1, T/N, H] shard -> [B, S, H] (exit of the decoder stack, after the final norm).  The incoming gradient is already
    complete on every rank (lm_head's colwise input gradient is all-reduced), so backward keeps the local rows."""

    @staticmethod
    def forward(ctx, x, st):
        ctx.st = st
        return _sp_gather_now(x, st).view(st.full_shape)

    @staticmethod
    def backward(ctx, g):
        return sp_take_local(g.reshape(-1, g.shape[-1]), ctx.st).unsqueeze(0), None


class _GatherTokensSumBwd(torch.autograd.Function):
    """Block entry off the kernel path: all-gather forward, reduce-scatter of the partial input gradients backward."""

    @staticmethod
    def forward(ctx, x, st):
        ctx.st = st
        return _sp_gather_now(x, st).view(st.full_shape)

    @staticmethod
    def backward(ctx, g):
        return _sp_scatter_now(g, ctx.st).unsqueeze(0), None


class _ReduceScatterTokens(torch.autograd.Function):
    """Block exit off the kernel path: reduce-scatter of the rowwise partial sums forward, all-gather backward."""

    @staticmethod
    def forward(ctx, x, st):
        ctx.st, ctx.shape = st, x.shape
        return _sp_scatter_now(x, st).unsqueeze(0)

    @staticmethod
    def backward(ctx, g):
        return _sp_gather_now(g, ctx.st).view(ctx.shape), None


def shard_tokens(x, st):
    return _ShardTokens.apply(x, st)


def gather_tokens(x, st, summed_grad: bool = False):
    return (_GatherTokens if summed_grad else _GatherTokensSumBwd).apply(x, st)


def reduce_scatter_tokens(x, st):
    return _ReduceScatterTokens.apply(x, st)


def _install_sequence_parallel(model: nn.Module, st: SequenceParallelState, block_names: set) -> None:
    prefix = getattr(model, "base_model_prefix", "model")
    base = getattr(model, prefix, model)
    layers, norm = getattr(base, "layers", None), getattr(base, "norm", None)
    if not isinstance(layers, nn.ModuleList) or len(layers) == 0 or norm is None:
        raise ValueError("sequence_parallel needs a decoder stack with `.layers` and a final `.norm`")

    def enter(module, args, kwargs):
        hidden = args[0] if args else kwargs["hidden_states"]
        cached = kwargs.get("past_key_values", kwargs.get("past_key_value")) is not None
        st.active = hidden.dim() == 3 and not cached and st.usable(hidden.shape[0] * hidden.shape[1])
        if not st.active:
            return None
        st.full_shape = tuple(hidden.shape)
        hidden = shard_tokens(hidden, st)
        if args:
            return (hidden, *args[1:]), kwargs
        return args, {**kwargs, "hidden_states": hidden}

    def leave(module, args, output):
        return gather_tokens(output, st, summed_grad=True) if st.active else None

    layers[0].register_forward_pre_hook(enter, with_kwargs=True)
    norm.register_forward_hook(leave)

    # replicated parameters that now see only this rank's tokens (the norms): their gradients are partial sums
    def reduce_grad(g):
        if not st.active:
            return None
        g = g.contiguous().clone()
        dist.all_reduce(g, group=st.group)
        return g

    base_name = prefix + "." if base is not model else ""
    for name, p in base.named_parameters():
        full = base_name + name
        owner = full.rsplit(".", 1)[0]
        in_stack = name.startswith("layers.") or name.startswith("norm.")
        in_block = any(owner == b or owner.startswith(b + ".") for b in block_names)
        if in_stack and not in_block and p.requires_grad:
            p.register_hook(reduce_grad)


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


def _install_vocab_parallel_loss(model: nn.Module) -> None:
    """When a forward carries ``labels`` and the loss is ours, lm_head skips the logits all-gather and the loss combines
    per-row statistics across the ranks instead (functional.VocabParallelLossFn).  ``out.logits`` of such a forward is this
    rank's vocabulary shard; forwards without labels (generate) still return gathered logits."""
    from .integration import b200_causal_lm_loss

    head = getattr(model, "lm_head", None)
    if head is None or not head.__dict__.get("_b200_tp_gather", False):
        raise ValueError("vocab_parallel_loss needs an lm_head sharded as colwise_gather_output")

    def before(module, args, kwargs):
        ours = getattr(module, "loss_function", None) is b200_causal_lm_loss
        head.__dict__["_b200_keep_vocab_shard"] = bool(ours and kwargs.get("labels") is not None)

    def after(module, args, kwargs, output):
        head.__dict__["_b200_keep_vocab_shard"] = False

    model.register_forward_pre_hook(before, with_kwargs=True)
    model.register_forward_hook(after, with_kwargs=True, always_call=True)


def tensor_parallelize(model: nn.Module, group=None, plan: dict | None = None, sequence_parallel: bool = False,
                       chunks: int = 2, vocab_parallel_loss: bool = False, peer_workspace=None) -> nn.Module:
    """Shard an already materialised model in place (each rank keeps its slice) and tell the block modules which group
    to reduce over.  Mirrors apply_tensor_parallelism (distributed/tensor_parallel.py:773-796) for colwise / rowwise /
    colwise_gather_output; embeddings and norms stay replicated.  ``sequence_parallel``: see the module docstring;
    ``vocab_parallel_loss``: see ``_install_vocab_parallel_loss``; ``peer_workspace`` (a ``symm.PeerWorkspace``, needs
    ``sequence_parallel``): the kernel path's all-gathers / reduce-scatters run over NVLink peer memory with our own
    kernels instead of NCCL."""
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
    blocks = set()
    for name, mod in model.named_modules():
        if name in touched and hasattr(mod, "forward") and type(mod).__name__.startswith("B200"):
            mod.__dict__["_b200_tp_group"] = group
            mod.__dict__.pop("_b200_fused", None)
            blocks.add(name)
    model.__dict__["_b200_tp_world"] = world
    if sequence_parallel and world > 1:
        st = SequenceParallelState(group, 1 if peer_workspace is not None else chunks, peer=peer_workspace)
        for name, mod in model.named_modules():
            if name in blocks and "_b200_tp_gather" not in mod.__dict__:
                mod.__dict__["_b200_sp"] = st
        _install_sequence_parallel(model, st, blocks)
        model.__dict__["_b200_sp"] = st
    if vocab_parallel_loss and world > 1:
        _install_vocab_parallel_loss(model)
    return model
