#!/usr/bin/env bash
# Round-2 GPU call 9 (2 GPUs): the final tree's default multi-GPU path (eager peer-workspace reservation in bench.py).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-2}
echo "=== tp_check sp=1 vp=1 peer=2"
B200_TP_SP=1 B200_TP_VOCAB_LOSS=1 B200_TP_PEER=2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
  --master-addr 127.0.0.1 --master-port 29611 tests/cuda/tp_check.py 2>&1 | grep -v "^\[ERROR\]" | tail -3
echo "--- exit $?"
echo "=== bench (defaults, 8 layers)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus "$N" --steps 6 --warmup 3 --layers 8 --no-cpu-baseline 2>&1 | grep -v "^\[ERROR\]" | tail -2
echo "--- exit $?"
