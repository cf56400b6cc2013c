#!/usr/bin/env bash
# First GPU call of the next round: validates everything that was written after round 1's GPU budget was spent
# (never run on a device), cheapest first.  Two parts, because a multi-GPU box is charged N x its time:
#   gpurun --timeout 2700 -- 'PART=single bash tests/cuda/round2_validate.sh 2>&1 | tee gpurun_out/round2_single.log'
#   gpurun --gpus 2 --timeout 900 -- 'PART=multi bash tests/cuda/round2_validate.sh 2>&1 | tee gpurun_out/round2_multi.log'
# Every step is bounded by `timeout`; a failing step is reported and the script goes on.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-2}
PART=${PART:-single}
run() { echo "=== $*"; timeout "${T:-300}" "$@"; echo "--- exit $?"; }

if [ "$PART" = single ]; then
# 1. single-GPU kernels: vocabulary-sharded CE backward, multi-tensor AdamW / grad-norm / scale
T=300 run env B200_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q -m gpu -s
# 1b. decode path with the GEMV and split-context attention dispatch on: generate() parity tests
T=600 run env B200_GEMV=1 B200_DECODE_ATTN=1 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "generate or cache or decode"
# 2. regular suite still green on this build
T=600 run python -m pytest tests -x -q -m gpu
# 2b. attention forward softmax variant (batched TMEM loads, split max / sum chains): parity, then timing next to the default
T=600 run env B200_ATTN_FWD_ILP=1 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "attn or attention or llama"
T=300 run python tests/cuda/bringup_attn.py
T=300 run env B200_ATTN_FWD_ILP=1 python tests/cuda/bringup_attn.py
fi
if [ "$PART" = multi ]; then
# 3. NCCL parity of the tensor-parallel variants on a tiny model (logits / loss / gradient shards vs single GPU)
for cfg in "0 0" "2 0" "0 1" "2 1" "4 1"; do
  set -- $cfg
  T=240 run env B200_TP_SP=$1 B200_TP_VOCAB_LOSS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29611 tests/cuda/tp_check.py
done
# 3b. peer-memory transport (symmetric allocations, our pull-reduce kernel + copy-engine gathers instead of NCCL)
T=240 run env B200_TP_SP=1 B200_TP_PEER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
  --master-addr 127.0.0.1 --master-port 29613 tests/cuda/tp_check.py
T=240 run env B200_TP_SP=1 B200_TP_PEER=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
  --master-addr 127.0.0.1 --master-port 29615 tests/cuda/tp_check.py
for tr in peer peer-scatter; do
  T=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29614 \
    bench.py --gpus "$N" --steps 3 --warmup 3 --layers 8 --no-cpu-baseline --sequence-parallel 1 --tp-transport $tr
done
# 4. what the variants buy at the real shapes (8 of 32 layers: relative numbers only, NOT a bench value)
for flags in "" "--sequence-parallel 2" "--sequence-parallel 4" "--sequence-parallel 2 --vocab-parallel-loss 1"; do
  T=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus "$N" --steps 3 --warmup 3 --layers 8 --no-cpu-baseline $flags
done
fi
if [ "$PART" = single ]; then
# 5. packed weights (parameters as views of the fused buffer) on one GPU, 8 layers
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline --pack-weights 1
# 5b. residual adds on our kernels (fused add + RMSNorm decoder layer)
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline --fuse-residual 1
# 6. GEMM DRAM re-reads (profiles/README.md: 2-8x the algorithmic bytes): rasterisation group size sweep under ncu
#    (one GPU; ncu numbers are for traffic only, never a bench value)
for gm in 2 4 8 16 32; do
  T=300 run env B200_GEMM2_GROUP_M=$gm ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_gm$gm.csv python tests/cuda/prof_kernels.py gemm
done
# 7. GEMM lock-step variant: parity (all GEMM tests with the sync forced on for every K), then traffic + time next to step 6's gm=8
T=600 run env B200_GEMM2_SYNC=1 B200_GEMM2_SYNC_MIN_K=1 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k gemm
T=300 run env B200_GEMM2_SYNC=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_sync.csv python tests/cuda/prof_kernels.py gemm
T=300 run env B200_GEMM2_SYNC=1 python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline
T=300 run python bench.py --steps 3 --warmup 3 --layers 8 --no-cpu-baseline
# 8. the other BASELINE configs, measured (JSON lines): Mixtral-8x7B forward (configs[3]), Gemma-2-9B generate (configs[4])
T=600 run python tests/cuda/bench_configs45.py mixtral
T=600 run python tests/cuda/bench_configs45.py gemma2
T=600 run env B200_GEMV=1 B200_DECODE_ATTN=1 python tests/cuda/bench_configs45.py gemma2 --inplace-sliding
fi
