#!/usr/bin/env bash
# Round-2 GPU call 1 (single GPU): validate every opt-in written without a GPU, then measure configs[3]/[4].
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout "${T:-300}" "$@"; echo "--- exit $?"; }
python -c "import sys; sys.path.insert(0,'baseline'); import ref_import; print(ref_import.where())"
T=400 run env B200_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q -m gpu -s
T=600 run python -m pytest tests -x -q -m gpu
T=600 run env B200_GEMV=1 B200_DECODE_ATTN=1 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "generate or cache or decode"
T=600 run env B200_ATTN_FWD_ILP=1 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "attn or attention or llama"
T=300 run python tests/cuda/bringup_attn.py
T=300 run env B200_ATTN_FWD_ILP=1 python tests/cuda/bringup_attn.py
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline --pack-weights 1
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline --fuse-residual 1
T=600 run env B200_GEMM2_SYNC=1 B200_GEMM2_SYNC_MIN_K=1 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k gemm
T=300 run env B200_GEMM2_SYNC=1 python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline
T=600 run python tests/cuda/bench_configs45.py gemma2
T=600 run env B200_GEMV=1 B200_DECODE_ATTN=1 python tests/cuda/bench_configs45.py gemma2 --inplace-sliding
T=600 run python tests/cuda/bench_configs45.py mixtral
