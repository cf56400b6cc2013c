#!/usr/bin/env bash
# Round-2 GPU call 8 (1 GPU): the final tree -- whole GPU suite, smoke() (now through the GLU-epilogue GEMM and the chunked fused
# head + loss), and a short bench run after the deferred embedding range check went onto the hot path.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@" 2>&1 | grep -v "^\[ERROR\]"; echo "--- exit ${PIPESTATUS[0]}"; }
T=300 run python -m pytest tests -q -m gpu
T=100 run python -c "import __graft_entry__ as g; g.smoke()"
T=200 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
