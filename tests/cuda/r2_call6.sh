#!/usr/bin/env bash
# Round-2 GPU call 6 (1 GPU): final-tree suite, the ncu evidence (launch list of one bench step + full-set captures of the hot
# kernels), the final bench line and one run of the reference arm's new protocol.  Numbers printed under ncu are never bench values.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@" 2>&1 | grep -v "^\[ERROR\]"; echo "--- exit ${PIPESTATUS[0]}"; }
T=300 run python -m pytest tests -q -m gpu
T=300 run python bench.py --steps 10 --warmup 3
T=400 run python bench.py --impl reference --steps 1 --warmup 1
T=300 run ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline
T=400 run ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attn_|rmsnorm|glu_|rope|ce_|add_kernel" \
  -o gpurun_out/r02_prof python tests/cuda/prof_kernels.py
ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw.csv 2>/dev/null
gzip -f gpurun_out/r02_ncu_raw.csv gpurun_out/r02_ncu_launches_bench.csv
ls -la gpurun_out/ | head -20
rm -f gpurun_out/r02_prof.ncu-rep   # keep the merged scratch small; the raw page carries every metric
