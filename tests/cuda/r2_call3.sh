#!/usr/bin/env bash
# Round-2 GPU call 3 (8 GPUs): NCCL parity at TP=8 (plain plan and the peer-memory scatter transport), then the full
# 32-layer bench line with the new multi-GPU defaults next to the plain-NCCL plan.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-8}
for cfg in "0 0 0" "1 1 2"; do
  set -- $cfg
  echo "=== tp_check sp=$1 vp=$2 peer=$3"
  B200_TP_SP=$1 B200_TP_VOCAB_LOSS=$2 B200_TP_PEER=$3 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29611 tests/cuda/tp_check.py 2>&1 | grep -v "^\[ERROR\]" | tail -10
  echo "--- exit $?"
done
for flags in "" "--tp-transport peer" "--tp-transport nccl --sequence-parallel 0 --vocab-parallel-loss 0"; do
  echo "=== bench $flags"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus "$N" --steps 10 --warmup 3 --no-cpu-baseline $flags 2>&1 | grep -v "^\[ERROR\]" | tail -3
  echo "--- exit $?"
done
