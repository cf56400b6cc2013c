#!/usr/bin/env bash
# Round-2 GPU call 5 (2 GPUs): TP regression after the GLU-epilogue GEMM (col / col_sp modes), the multi-stream copy-engine
# all-gather and the fused head + loss default: NCCL-free parity on the tiny model, then the 8-layer bench next to call 2's
# numbers (plain 126.2 ms, peer-scatter 115.6 ms on that box).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-2}
for cfg in "1 1 2" "0 0 0"; do
  set -- $cfg
  echo "=== tp_check sp=$1 vp=$2 peer=$3"
  B200_TP_SP=$1 B200_TP_VOCAB_LOSS=$2 B200_TP_PEER=$3 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29611 tests/cuda/tp_check.py 2>&1 | grep -v "^\[ERROR\]" | tail -4
  echo "--- exit $?"
done
for flags in "" "--fuse-glu 0" "--tp-transport nccl --sequence-parallel 0 --vocab-parallel-loss 0"; do
  echo "=== bench $flags"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus "$N" --steps 6 --warmup 3 --layers 8 --no-cpu-baseline $flags 2>&1 | grep -v "^\[ERROR\]" | tail -2
  echo "--- exit $?"
done
