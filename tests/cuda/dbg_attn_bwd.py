"""Per-role cycle counters of CTA 0 (kv tile 0) of attn_bwd_dkdv at the Llama-3-8B shape."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transformers_b200 import _lib, ops
lib = _lib.load(); _lib.require_device()
B, S, Hq, Hkv, D = 4, 4096, 32, 8, 128
qkv = torch.randn(B, S, (Hq + 2 * Hkv) * D, device="cuda").to(torch.bfloat16)
q = qkv[..., : Hq * D].view(B, S, Hq, D); k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D); v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
out, lse = ops.attn_fwd(q, k, v, scale=D ** -0.5, causal=True)
do = torch.randn(B, S, Hq, D, device="cuda").to(torch.bfloat16); dqkv = torch.empty_like(qkv)
dq = dqkv[..., : Hq * D].view(B, S, Hq, D); dk = dqkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D); dv = dqkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
for _ in range(2): ops.attn_bwd(q, k, v, out, do, lse, dq, dk, dv, scale=D ** -0.5, causal=True)
lib.b200_debug_set_buffer(dbg.data_ptr())
ops.attn_bwd(q, k, v, out, do, lse, dq, dk, dv, scale=D ** -0.5, causal=True)
torch.cuda.synchronize()
lib.b200_debug_set_buffer(None)
d = dbg.tolist(); n = max(d[2], 1)
print(f"iters {d[2]}")
print(f"TMA : total {d[0]} ({d[0]/n:.0f}/it)  wait q_empty {d[1]} ({d[1]/n:.0f}/it)")
print(f"MMA : total {d[4]} ({d[4]/n:.0f}/it)  wait q_full {d[5]/n:.0f}/it  wait pds_full {d[6]/n:.0f}/it  issue sdp(16xSS N64) {d[7]/n:.0f}/it  issue dvdk(8xTS N128) {d[11]/n:.0f}/it")
