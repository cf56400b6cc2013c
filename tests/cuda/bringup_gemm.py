"""GPU bring-up probe for the tcgen05 GEMM: every operand layout x descriptor variant against torch.matmul.
Run on a B200: python tests/cuda/bringup_gemm.py  (writes gpurun_out/bringup_gemm.txt)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transformers_b200 import _lib

lib = _lib.load()
_lib.require_device()
dev = "cuda"
out_lines = []
def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out_lines.append(s)

def run(M, N, K, a_mn, b_mn, variant, accumulate=0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)   # logical A[m,k]
    B = torch.randn(N, K, device=dev, generator=g).to(torch.bfloat16)   # logical B[n,k]
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.t().contiguous() if b_mn else B.contiguous()
    C0 = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16) if accumulate else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    C = C0.clone()
    if variant == 2:
        rc = lib.b200_gemm_bf16_2sm(As.data_ptr(), Bs.data_ptr(), C.data_ptr(), M, N, K, As.stride(0), Bs.stride(0), C.stride(0),
                                    a_mn, b_mn, accumulate, torch.cuda.current_stream().cuda_stream)
    else:
        rc = lib.b200_gemm_bf16_1sm(As.data_ptr(), Bs.data_ptr(), C.data_ptr(), M, N, K, As.stride(0), Bs.stride(0), C.stride(0),
                                    a_mn, b_mn, accumulate, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        return f"rc={rc} {_lib.last_error()}"
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if accumulate: ref = ref + C0.float()
    err = (C.float() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    return rel

log(torch.cuda.get_device_name(0))
for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1), (1, 0)]:
    for variant in [0, 2]:
        for (M, N, K) in [(128, 256, 64), (256, 256, 64), (256, 512, 128), (200, 264, 72), (1024, 1024, 1024), (1000, 520, 264), (4096, 6144, 4096)]:
            try:
                r = run(M, N, K, a_mn, b_mn, variant)
            except Exception as e:
                r = f"EXC {e}"
            ok = isinstance(r, float) and r < 2e-2
            log(f"a_mn={a_mn} b_mn={b_mn} variant={variant} M={M} N={N} K={K} rel_err={r} {'OK' if ok else 'FAIL'}")
log("accumulate:", run(256, 512, 128, 0, 0, 0, accumulate=1), run(512, 512, 128, 0, 0, 2, accumulate=1))

# timing, Llama-3-8B shapes
def bench(M, N, K, a_mn, b_mn, iters=20):
    A = torch.randn(K, M, device=dev).to(torch.bfloat16) if a_mn else torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(K, N, device=dev).to(torch.bfloat16) if b_mn else torch.randn(N, K, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.b200_gemm_bf16_2sm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, A.stride(0), B.stride(0), C.stride(0), a_mn, b_mn, 0, st)
    f1 = lambda: lib.b200_gemm_bf16_1sm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, A.stride(0), B.stride(0), C.stride(0), a_mn, b_mn, 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): f1()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): f1()
    e1.record(); torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / iters
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # cuBLAS reference
    Ar = torch.randn(M, K, device=dev).to(torch.bfloat16); Br = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for _ in range(3): torch.matmul(Ar, Br.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): torch.matmul(Ar, Br.t())
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / 1e12
    log(f"bench M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: 2sm {ms:.3f} ms = {tf/ms*1e3:.0f} TF/s | 1sm {ms1:.3f} ms = {tf/ms1*1e3:.0f} TF/s | cuBLAS {ms_ref:.3f} ms = {tf/ms_ref*1e3:.0f} TF/s")

try:
    for (M, N, K, a, b) in [(16384, 6144, 4096, 0, 0), (16384, 4096, 4096, 0, 0), (16384, 28672, 4096, 0, 0), (16384, 4096, 14336, 0, 0),
                            (16384, 4096, 6144, 0, 1), (16384, 4096, 28672, 0, 1), (6144, 4096, 16384, 1, 1), (28672, 4096, 16384, 1, 1), (8192, 8192, 8192, 0, 0)]:
        bench(M, N, K, a, b)
except Exception as e:
    log("bench EXC", e)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/bringup_gemm.txt", "w").write("\n".join(out_lines) + "\n")
