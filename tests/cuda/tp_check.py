"""Real-NCCL check of the tensor-parallel path (run under torchrun on N GPUs): a tiny Llama through the plugin, sharded
by the reference tp_plan, must reproduce the single-GPU logits / loss / gradient shards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from _hf import import_transformers
tf = import_transformers()
import transformers_b200
from transformers_b200.parallel import tensor_parallelize

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
transformers_b200.enable()
cfg = tf.LlamaConfig(vocab_size=1024, hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
                     num_key_value_heads=8, head_dim=64, max_position_embeddings=2048, use_cache=False,
                     rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
tf.set_seed(0)
model = tf.LlamaForCausalLM._from_config(cfg, attn_implementation="b200", dtype=torch.bfloat16).cuda()
transformers_b200.accelerate(model, fused_head_loss=False)  # the single-GPU reference pass compares logits
torch.manual_seed(1)
ids = torch.randint(0, 1024, (2, 256 * max(1, world // 2)), device="cuda")  # T / world stays a multiple of 256 (scatter epilogue)
ref = model(input_ids=ids, labels=ids); ref.loss.backward()
ref_logits, ref_loss = ref.logits.detach().float().clone(), ref.loss.item()
ref_g = {n: p.grad.detach().float().clone() for n, p in model.named_parameters()}
model.zero_grad(set_to_none=True)
SP = int(os.environ.get("B200_TP_SP", "0"))  # >0: sequence parallel with that many pipelined chunks
VP = bool(int(os.environ.get("B200_TP_VOCAB_LOSS", "0")))  # vocabulary-parallel loss (no logits all-gather)
PEER = int(os.environ.get("B200_TP_PEER", "0"))  # 1: collectives over NVLink peer memory, 2: + GEMM scatter epilogue (needs B200_TP_SP>0)
ws = None
if PEER:
    from transformers_b200.symm import PeerWorkspace
    ws = PeerWorkspace(dist.group.WORLD, scatter_epilogue=PEER == 2)
tensor_parallelize(model, sequence_parallel=SP > 0, chunks=max(SP, 1), vocab_parallel_loss=VP, peer_workspace=ws)
out = model(input_ids=ids, labels=ids); out.loss.backward()
if VP:
    ref_logits = ref_logits.chunk(world, dim=-1)[rank]
err = (out.logits.float() - ref_logits).abs().max().item()
styles = {"q_proj": 0, "k_proj": 0, "v_proj": 0, "gate_proj": 0, "up_proj": 0, "lm_head": 0, "o_proj": 1, "down_proj": 1}
worst = 0.0
for n, p in model.named_parameters():
    g = ref_g[n]; leaf = n.split(".")[-2]
    if leaf in styles: g = g.chunk(world, dim=styles[leaf])[rank]
    worst = max(worst, ((p.grad.float() - g).abs().max() / (g.abs().max() + 1e-8)).item())
ok = err < 5e-2 and abs(out.loss.item() - ref_loss) < 2e-2 and worst < 5e-2
print(f"rank {rank}/{world} sp={SP} vp={int(VP)} peer={int(PEER)}: logits max err {err:.4f}, loss {out.loss.item():.4f} vs {ref_loss:.4f}, worst grad rel err {worst:.4f} -> {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
