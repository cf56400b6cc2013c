"""The b200 tensor-parallel styles through the REFERENCE's own dispatch (ParallelInterface registry -> model.tp_plan ->
apply_tensor_parallelism, distributed/tensor_parallel.py:742-796), 2 processes over gloo: logits, loss and the local gradient
shards must equal the single-process model -- on the stock CPU path (the styles run each Linear on its local shard, the b200
blocks issue the collectives) and on the kernel path (C-ABI calls replaced by tests/_fake_ops.py)."""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, kernel_path):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from _hf import import_transformers

        tf = import_transformers()
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.tensor import DTensor
        from transformers.distributed.tensor_parallel import ALL_PARALLEL_STYLES, apply_tensor_parallelism

        import transformers_b200
        from transformers_b200.tp_styles import b200_tp_plan

        transformers_b200.enable()
        assert all(k in ALL_PARALLEL_STYLES for k in ("b200_colwise", "b200_rowwise", "b200_colwise_gather_output"))
        cfg = tf.LlamaConfig(vocab_size=160, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                             num_key_value_heads=2, head_dim=16, max_position_embeddings=512,
                             rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
        tf.set_seed(0)
        model = tf.LlamaForCausalLM._from_config(cfg, attn_implementation="eager", dtype=torch.float32)
        transformers_b200.accelerate(model, attn=False, fused_head_loss=False)
        model.loss_function = None
        del model._loss_function
        torch.manual_seed(1)
        ids = torch.randint(0, 160, (2, 96))
        model.config.use_cache = False
        ref = model(input_ids=ids, labels=ids)
        ref.loss.backward()
        ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        ref_logits, ref_loss = ref.logits.detach().clone(), ref.loss.detach().clone()
        model.zero_grad(set_to_none=True)

        plan = b200_tp_plan(model)
        assert plan["model.layers.*.self_attn.q_proj"] == "b200_colwise" and plan["model.layers.*.mlp.down_proj"] == "b200_rowwise"
        assert plan["lm_head"] == "b200_colwise_gather_output"
        model.tp_plan = plan  # the reference's setter validates every style name against its registry
        if kernel_path:
            import _fake_ops

            _fake_ops.install()
        apply_tensor_parallelism(model, init_device_mesh("cpu", (world,)))
        if kernel_path:
            model.set_attn_implementation("b200")
            model.loss_function = transformers_b200.integration.b200_causal_lm_loss
        att = model.model.layers[0].self_attn
        assert isinstance(att.q_proj.weight, DTensor) and att.q_proj.weight.to_local().shape == (4 * 16 // world, 64)
        assert att.o_proj.weight.to_local().shape == (64, 4 * 16 // world)
        assert model.lm_head.weight.to_local().shape == (160 // world, 64)
        out = model(input_ids=ids, labels=ids)
        out.loss.backward()
        torch.testing.assert_close(out.logits, ref_logits, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(out.loss, ref_loss, atol=1e-5, rtol=1e-5)
        if kernel_path:
            names = [c[0] for c in _fake_ops.CALLS]
            assert names.count("attn_fwd") == 2 and names.count("gemm_glu") == 2  # the fused blocks ran on the local shards
        styles = {"q_proj": 0, "k_proj": 0, "v_proj": 0, "gate_proj": 0, "up_proj": 0, "lm_head": 0, "o_proj": 1, "down_proj": 1}
        for n, p in model.named_parameters():
            g = ref_grads[n]
            leaf = n.split(".")[-2]
            if leaf in styles:
                g = g.chunk(world, dim=styles[leaf])[rank]
            got = p.grad.to_local() if isinstance(p.grad, DTensor) else p.grad
            torch.testing.assert_close(got, g, atol=1e-5, rtol=1e-4, msg=lambda m, n=n: f"{n}: {m}")
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("kernel_path", [False, True])
def test_b200_styles_through_the_reference_dispatch(kernel_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kernel_path)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=280) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
