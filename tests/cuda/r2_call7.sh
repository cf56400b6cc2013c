#!/usr/bin/env bash
# Round-2 GPU call 7 (1 GPU): grouped GEMM with device-side expert offsets (MoE forward + dgrad): kernel parity, the MoE paths
# that now use it, the whole suite, then configs[3] through bench.py.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@" 2>&1 | grep -v "^\[ERROR\]"; echo "--- exit ${PIPESTATUS[0]}"; }
T=120 run python -m pytest tests/test_kernels_gpu.py tests/test_kernels2_gpu.py -q -m gpu -k "grouped or moe"
T=300 run python -m pytest tests -q -m gpu
T=300 run python bench.py --config mixtral-8x7b-forward --steps 5 --warmup 3
