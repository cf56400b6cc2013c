"""Grouped expert GEMM at the Mixtral-8x7B / seq 2048 shape (4096 routed rows over 8 experts, gate|up projection 28672 x 4096):
timing against the per-expert launches it replaced (CUDA events), and -- when run under ncu -- the capture target.
  python tests/cuda/prof_grouped.py            # timing lines
  ncu --set full --clock-control none -k regex:gemm_bf16 -c 4 -o gpurun_out/r02_grouped python tests/cuda/prof_grouped.py ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transformers_b200 import ops

dev, BF = "cuda", torch.bfloat16
E, N, K = 8, 28672, 4096
counts = [480, 530, 512, 600, 450, 500, 512, 512]
M = sum(counts)
off_host = [0]
for c in counts:
    off_host.append(off_host[-1] + c)
offsets = torch.tensor(off_host, dtype=torch.int32, device=dev)
xs = torch.randn(M, K, device=dev).to(BF)
w = (torch.randn(E, N, K, device=dev) * 0.02).to(BF)
out = torch.empty(M, N, device=dev, dtype=BF)
under_ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
reps = 2 if under_ncu else 10


def grouped():
    ops.gemm_grouped(xs, w, offsets, out=out)


def per_expert():
    for e in range(E):
        ops.gemm(xs[off_host[e]:off_host[e + 1]], w[e], out=out[off_host[e]:off_host[e + 1]])


for name, fn in (("grouped (1 launch, device-side offsets)", grouped), ("per expert (8 launches, host-side offsets)", per_expert)):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if not under_ncu:
        print(f"{name}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TF/s useful ({sum((c + 255) // 256 for c in counts) * 256} rows computed for {M})")
    if under_ncu:
        break
print("done")
