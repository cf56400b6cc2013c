#!/usr/bin/env bash
# Round-2 GPU call 10 (1 GPU): grouped expert GEMM at the Mixtral shape -- timing next to the per-expert launches, one ncu capture.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== timing"; timeout 100 python tests/cuda/prof_grouped.py 2>&1 | grep -v "^\[ERROR\]"; echo "--- exit $?"
echo "=== ncu"; timeout 150 ncu --set full --clock-control none -k regex:gemm_bf16 -c 4 -o gpurun_out/r02_grouped python tests/cuda/prof_grouped.py ncu 2>&1 | tail -3
ncu -i gpurun_out/r02_grouped.ncu-rep --page raw --csv > gpurun_out/r02_ncu_grouped_raw.csv 2>/dev/null; gzip -f gpurun_out/r02_ncu_grouped_raw.csv; rm -f gpurun_out/r02_grouped.ncu-rep
echo "--- exit $?"
