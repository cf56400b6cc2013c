#!/usr/bin/env bash
# Round-2 GPU call 4 (1 GPU): first device run of the two-tile attention forward, the TMA reduce-add accumulate, the GLU-epilogue
# GEMM and the chunked fused lm_head + loss.  Risky groups run in their own processes (a trapped kernel poisons its CUDA
# context); the bench falls back feature by feature so that a broken feature is named by the first line that works.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@" 2>&1 | grep -v "^\[ERROR\]"; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=200 run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "two_tile or attention"
T=200 run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "glu or accumulate or gemm"
T=200 run python -m pytest tests/test_model_gpu.py -q -m gpu -k "fused_head"
T=900 run python -m pytest tests -q -m gpu
T=100 run python -c "import __graft_entry__ as g; g.smoke()"
T=200 run python tests/cuda/bringup_attn.py
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
T=300 run $B || T=300 run $B --attn-one-tile 1 || T=300 run $B --attn-one-tile 1 --fuse-glu 0 || T=300 run $B --attn-one-tile 1 --fuse-glu 0 --fused-head-loss 0
T=300 run $B --attn-one-tile 1 --fuse-glu 0 --fused-head-loss 0
T=300 run python bench.py --config llama3-8b-trainer-step --steps 4 --warmup 3
T=300 run python bench.py --config mixtral-8x7b-forward --steps 5 --warmup 3
T=300 run python bench.py --config gemma2-9b-generate
