"""GPU bring-up probe for the tcgen05 flash-attention forward/backward (and the HBM-bound kernels) vs torch references."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from transformers_b200 import _lib

lib = _lib.load(); _lib.require_device()
dev = "cuda"; lines = []
def log(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); lines.append(s)
st = lambda: torch.cuda.current_stream().cuda_stream

def ref_attn(q, k, v, scale, causal, window, softcap, kv_start=None, kv_end=None):
    # q [B,Sq,Hq,D] k,v [B,Skv,Hkv,D] -> out [B,Sq,Hq,D], lse [B,Hq,Sq]; fp32 math
    B, Sq, Hq, D = q.shape; Skv, Hkv = k.shape[1], k.shape[2]
    qf = q.float().transpose(1, 2); kf = k.float().transpose(1, 2).repeat_interleave(Hq // Hkv, 1); vf = v.float().transpose(1, 2).repeat_interleave(Hq // Hkv, 1)
    s = qf @ kf.transpose(-1, -2) * scale
    if softcap > 0: s = torch.tanh(s / softcap) * softcap
    qi = torch.arange(Sq, device=dev)[:, None] + (Skv - Sq); ki = torch.arange(Skv, device=dev)[None, :]
    allowed = torch.ones(Sq, Skv, dtype=torch.bool, device=dev)
    if causal: allowed &= ki <= qi
    if window > 0: allowed &= ki > qi - window
    allowed = allowed[None, None].expand(B, 1, Sq, Skv).clone()
    if kv_start is not None: allowed &= (ki[None, None] >= kv_start[:, None, None, None])
    if kv_end is not None: allowed &= (ki[None, None] < kv_end[:, None, None, None])
    s = s.masked_fill(~allowed, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1); p = torch.nan_to_num(p, 0.0)
    return (p @ vf).transpose(1, 2), lse

def make_qkv(B, Sq, Skv, Hq, Hkv, D, packed, g):
    if packed and Sq == Skv:
        qkv = torch.randn(B, Sq, (Hq + 2 * Hkv) * D, device=dev, generator=g).to(torch.bfloat16)
        q = qkv[..., : Hq * D].view(B, Sq, Hq, D); k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, Skv, Hkv, D); v = qkv[..., (Hq + Hkv) * D:].view(B, Skv, Hkv, D)
    else:
        q = torch.randn(B, Sq, Hq, D, device=dev, generator=g).to(torch.bfloat16)
        k = torch.randn(B, Skv, Hkv, D, device=dev, generator=g).to(torch.bfloat16)
        v = torch.randn(B, Skv, Hkv, D, device=dev, generator=g).to(torch.bfloat16)
    return q, k, v

def call_fwd(q, k, v, out, lse, causal, window, softcap, ks=None, ke=None):
    B, Sq, Hq, D = q.shape; Skv, Hkv = k.shape[1], k.shape[2]
    return lib.b200_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), lse.shape[-1], B, Sq, Skv, Hq, Hkv, D,
                             q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0), v.stride(1), v.stride(2),
                             out.stride(0), out.stride(1), out.stride(2), D ** -0.5, softcap, causal, window,
                             ks.data_ptr() if ks is not None else None, ke.data_ptr() if ke is not None else None, st())

def run_attn(B, Sq, Skv, Hq, Hkv, D, causal=1, window=0, softcap=0.0, packed=False, pad=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    q, k, v = make_qkv(B, Sq, Skv, Hq, Hkv, D, packed, g)
    out = torch.empty(B, Sq, Hq, D, device=dev, dtype=torch.bfloat16)
    Sp = (Sq + 127) // 128 * 128
    lse = torch.empty(B, Hq, Sp, device=dev, dtype=torch.float32)
    ks = ke = None
    if pad:
        ks = torch.tensor(([0, 5] + [0] * B)[:B], device=dev, dtype=torch.int32)
        ke = torch.tensor(([Skv - 7, Skv] + [Skv] * B)[:B], device=dev, dtype=torch.int32)
    rc = call_fwd(q, k, v, out, lse, causal, window, softcap, ks, ke)
    if rc: return f"rc={rc} {_lib.last_error()}"
    torch.cuda.synchronize()
    qr, kr_, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ro, rl = ref_attn(qr, kr_, vr, D ** -0.5, causal, window, softcap, ks, ke)
    eo = (out.float() - ro).abs().max().item()
    fin = torch.isfinite(rl)
    el = (lse[..., :Sq][fin] - rl[fin]).abs().max().item() if fin.any() else 0.0
    # backward
    do = torch.randn(B, Sq, Hq, D, device=dev, generator=g).to(torch.bfloat16)
    ro.backward(do.float())
    dq = torch.full_like(q.contiguous(), float("nan")); dk = torch.full_like(k.contiguous(), float("nan")); dv = torch.full_like(v.contiguous(), float("nan"))
    ws = torch.empty(2 * B * Hq * Sp, device=dev, dtype=torch.float32)
    strides = []
    for t in (q, k, v, out, do, dq, dk, dv): strides += [t.stride(0), t.stride(1), t.stride(2)]
    import ctypes
    sarr = (ctypes.c_int64 * 24)(*strides)
    rc = lib.b200_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                           ws.data_ptr(), B, Sq, Skv, Hq, Hkv, D, Sp, ctypes.cast(sarr, ctypes.c_void_p), D ** -0.5, softcap, causal, window,
                           ks.data_ptr() if ks is not None else None, ke.data_ptr() if ke is not None else None, st())
    if rc: return f"bwd rc={rc} {_lib.last_error()}"
    torch.cuda.synchronize()
    rel = lambda a, r: ((a.float() - r).abs().max() / (r.abs().max() + 1e-6)).item()
    return eo, el, rel(dq, qr.grad), rel(dk, kr_.grad), rel(dv, vr.grad)

log(torch.cuda.get_device_name(0))
cases = [
    dict(B=1, Sq=128, Skv=128, Hq=1, Hkv=1, D=128), dict(B=1, Sq=128, Skv=128, Hq=1, Hkv=1, D=128, causal=0),
    dict(B=1, Sq=256, Skv=256, Hq=2, Hkv=1, D=128), dict(B=2, Sq=512, Skv=512, Hq=4, Hkv=2, D=128, packed=True),
    dict(B=2, Sq=200, Skv=200, Hq=4, Hkv=2, D=128), dict(B=2, Sq=384, Skv=384, Hq=4, Hkv=4, D=64),
    dict(B=2, Sq=512, Skv=512, Hq=4, Hkv=2, D=128, window=100), dict(B=2, Sq=512, Skv=512, Hq=4, Hkv=2, D=128, softcap=30.0),
    dict(B=2, Sq=1, Skv=300, Hq=4, Hkv=2, D=128), dict(B=2, Sq=64, Skv=320, Hq=4, Hkv=2, D=128),
    dict(B=2, Sq=300, Skv=300, Hq=4, Hkv=2, D=128, pad=True), dict(B=1, Sq=1024, Skv=1024, Hq=8, Hkv=2, D=128, packed=True),
]
for c in cases:
    try: r = run_attn(**c)
    except Exception as e: r = f"EXC {e}"
    ok = isinstance(r, tuple) and r[0] < 2e-2 and r[1] < 2e-2 and max(r[2:]) < 2e-2
    log("attn", c, "->", r, "OK" if ok else "FAIL")

# timing at the Llama-3-8B shape
try:
    B, S, Hq, Hkv, D = 4, 4096, 32, 8, 128
    q, k, v = make_qkv(B, S, S, Hq, Hkv, D, True, torch.Generator(device=dev).manual_seed(1))
    out = torch.empty(B, S, Hq, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, Hq, S, device=dev, dtype=torch.float32)
    f = lambda: call_fwd(q, k, v, out, lse, 1, 0, 0.0)
    import ctypes
    do = torch.randn(B, S, Hq, D, device=dev).to(torch.bfloat16)
    dq = torch.empty_like(do); dk = torch.empty(B, S, Hkv, D, device=dev, dtype=torch.bfloat16); dv = torch.empty_like(dk)
    ws = torch.empty(2 * B * Hq * S, device=dev, dtype=torch.float32)
    strides = []
    for t in (q, k, v, out, do, dq, dk, dv): strides += [t.stride(0), t.stride(1), t.stride(2)]
    sarr = (ctypes.c_int64 * 24)(*strides)
    fb = lambda: lib.b200_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                           ws.data_ptr(), B, S, S, Hq, Hkv, D, S, ctypes.cast(sarr, ctypes.c_void_p), D ** -0.5, 0.0, 1, 0, None, None, st())
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * Hq * S * S * D / 2
    log(f"attn fwd B4 S4096 H32/8 D128 causal: {ms:.3f} ms = {fl/ms/1e9:.0f} TF/s")
    for _ in range(3): fb()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fb()
    e1.record(); torch.cuda.synchronize(); msb = e0.elapsed_time(e1) / 10
    log(f"attn bwd same shape: {msb:.3f} ms = {2.5*fl/msb/1e9:.0f} TF/s (5-matmul count), {3.5*fl/msb/1e9:.0f} TF/s executed")
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    g = lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=True, enable_gqa=True)
    for _ in range(3): g()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): g()
    e1.record(); torch.cuda.synchronize(); ms2 = e0.elapsed_time(e1) / 10
    log(f"torch sdpa same shape: {ms2:.3f} ms = {fl/ms2/1e9:.0f} TF/s")
    ro = g().transpose(1, 2)
    log("max diff vs sdpa:", (out.float() - ro.float()).abs().max().item())
except Exception as e:
    log("timing EXC", e)

# ---- HBM-bound kernels vs torch (quick sanity; the real parity tests live in tests/)
try:
    T, H = 4096, 4096
    x = torch.randn(T, H, device=dev).to(torch.bfloat16); w = (1 + 0.1 * torch.randn(H, device=dev)).to(torch.bfloat16)
    y = torch.empty_like(x); rstd = torch.empty(T, device=dev)
    lib.b200_rmsnorm_fwd(x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(), rstd.data_ptr(), T, H, 1e-5, 0, st())
    xf = x.float(); ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16)
    log("rmsnorm fwd maxdiff", (y.float() - ref.float()).abs().max().item(), "mismatch frac", (y != ref).float().mean().item())
    dy = torch.randn(T, H, device=dev).to(torch.bfloat16); dx = torch.empty_like(x); dw = torch.zeros(H, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(lib.b200_rmsnorm_bwd_workspace_rows() * H, device=dev)
    lib.b200_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dw.data_ptr(), ws.data_ptr(), T, H, 0, 0, st())
    xr = x.float().requires_grad_(True); wr = w.float().requires_grad_(True)
    (wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5))).backward(dy.float())
    log("rmsnorm bwd dx maxdiff", (dx.float() - xr.grad).abs().max().item(), "dw rel", ((dw.float() - wr.grad).abs().max() / wr.grad.abs().max()).item())
    for name, fn in [("rmsnorm_fwd", lambda: lib.b200_rmsnorm_fwd(x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(), rstd.data_ptr(), T, H, 1e-5, 0, st())),
                     ("rmsnorm_bwd", lambda: lib.b200_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dw.data_ptr(), ws.data_ptr(), T, H, 0, 0, st()))]:
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
        nbytes = T * H * 2 * (2 if name == "rmsnorm_fwd" else 3)
        log(f"{name} T={T} H={H}: {ms*1e3:.1f} us = {nbytes/ms/1e6:.0f} GB/s")
except Exception as e:
    log("elementwise EXC", e)
os.makedirs("gpurun_out", exist_ok=True); open("gpurun_out/bringup_attn.txt", "w").write("\n".join(lines) + "\n")
