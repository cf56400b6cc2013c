"""Tensor-parallel styles registered with the reference's own ``ParallelInterface`` (distributed/tensor_parallel.py:742-770),
so that the reference's dispatch -- ``model.tp_plan = {...}`` (distributed/mixin.py:105-134) followed by
``apply_tensor_parallelism(model, mesh)`` (:773-796; what ``from_pretrained(..., distributed_config=...)`` runs,
modeling_utils.py:4323) -- shards a model for the b200 blocks instead of our own ``parallel.tensor_parallelize``:

    transformers_b200.enable()                      # registers "b200_colwise" / "b200_rowwise" / "b200_colwise_gather_output"
    model.tp_plan = transformers_b200.tp_styles.b200_tp_plan(model)
    apply_tensor_parallelism(model, init_device_mesh("cuda", (world,)))

Parameters become DTensor placeholders exactly as with the reference's ``colwise`` / ``rowwise`` styles (the styles below
inherit ``validate_param`` / ``shard_param``), i.e. the loader reads only this rank's slice (shard-on-read,
distributed/sharding_utils.py).  What differs is the forward: the reference wraps every nn.Linear in DTensor redistributes
(one blocking collective per projection, :219-226, :320-328); here the Linear keeps plain local tensors and the enclosing
b200 block (attention / MLP / lm_head) issues its two collectives itself, overlapped with its own GEMMs
(functional.FusedLinearFn).  The styles therefore only (1) run the Linear's own forward on the local shard when something
calls it directly (CPU tensors, an un-patched parent) and (2) tell the parent block which process group to reduce over.
The sequence-parallel / peer-memory variants need model-level hooks and stay with ``parallel.tensor_parallelize``."""
from __future__ import annotations

_REGISTERED = False

STYLE_OF = {"colwise": "b200_colwise", "rowwise": "b200_rowwise", "colwise_gather_output": "b200_colwise_gather_output"}


def _group(mesh):
    return mesh.get_group() if mesh.ndim == 1 else mesh.get_group("tp")


def register_tp_styles() -> bool:
    """Idempotent; returns False when this transformers has no ParallelInterface / torch.distributed is unavailable."""
    global _REGISTERED
    if _REGISTERED:
        return True
    try:
        from torch.distributed.tensor import Replicate, Shard
        from transformers.distributed.tensor_parallel import ColwiseParallel, ParallelInterface, RowwiseParallel
    except Exception:
        return False

    class _LocalForward:
        """Forward hooks of a b200 style: local tensors in, local tensors out, no redistribute; the parent block learns the
        process group from the marker left on the Linear (modules._tp_modes)."""

        b200_gather = False

        def should_use_local_tensors(self, module):
            return True

        def transform_inputs_pre_forward(self, module, args, kwargs, mesh):
            return args, kwargs

        def transform_output_post_forward(self, module, output, mesh):
            return output

        def install_forward(self, module, mesh):
            module.__dict__["_b200_tp_group"] = _group(mesh)
            if self.b200_gather:
                module.__dict__["_b200_tp_gather"] = True
            return super().install_forward(module, mesh)

    class B200ColwiseParallel(_LocalForward, ColwiseParallel):
        """colwise (q/k/v/gate/up): weight Shard(0); the block's input gradient is all-reduced inside its Function."""

    class B200RowwiseParallel(_LocalForward, RowwiseParallel):
        """rowwise (o/down): weight Shard(1); the block all-reduces the partial output, overlapped with its GEMM."""

    class B200ColwiseGatherParallel(_LocalForward, ColwiseParallel):
        """colwise_gather_output (lm_head): the b200 Linear gathers the vocabulary shards itself."""

        b200_gather = True

    ParallelInterface.register("b200_colwise", B200ColwiseParallel(input_layouts=Replicate(), output_layouts=Shard(-1)))
    ParallelInterface.register("b200_rowwise", B200RowwiseParallel(input_layouts=Shard(-1), output_layouts=Replicate()))
    ParallelInterface.register("b200_colwise_gather_output",
                               B200ColwiseGatherParallel(input_layouts=Replicate(), output_layouts=Replicate()))
    _REGISTERED = True
    return True


def b200_tp_plan(model) -> dict:
    """The model's own tp_plan (config.base_model_tp_plan + the class-level ``_tp_plan``) with every colwise / rowwise /
    colwise_gather_output entry mapped to its b200 style; entries in other styles are kept as they are."""
    from .parallel import resolve_plan

    return {pattern: STYLE_OF.get(style, style) for pattern, style in resolve_plan(model).items()}
