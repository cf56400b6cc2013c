#!/usr/bin/env bash
# Round-2 GPU call 3a (1 GPU): the regular suite after the default flips (0 skipped expected), then GEMM tuning A/B.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@" 2>&1 | grep -v "^\[ERROR\]"; echo "--- exit ${PIPESTATUS[0]}"; }
T=900 run python -m pytest tests -q -m gpu -x
T=120 run python -c "import __graft_entry__ as g; g.smoke()"
for tune in "0,0,0,0" "0,0,8192,0" "0,0,0,0" "0,0,8192,0" "4,0,0,0" "16,0,0,0"; do
  T=300 run python bench.py --steps 8 --warmup 3 --layers 8 --no-cpu-baseline --gemm-tuning $tune
done
