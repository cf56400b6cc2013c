#!/usr/bin/env bash
# Round-2 GPU call 2 (N GPUs): NCCL parity of the TP variants on a tiny model, then what each buys at real shapes (8 layers).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-2}
run() { echo "=== $*"; timeout "${T:-300}" "$@"; echo "--- exit $?"; }
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
for cfg in "0 0 0" "2 1 0" "4 1 0" "1 0 1" "1 1 2"; do
  set -- $cfg
  echo "=== tp_check sp=$1 vp=$2 peer=$3"
  B200_TP_SP=$1 B200_TP_VOCAB_LOSS=$2 B200_TP_PEER=$3 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29611 tests/cuda/tp_check.py 2>&1 | grep -v "^\[ERROR\]" | tail -6
  echo "--- exit $?"
done
for flags in "" "--sequence-parallel 2 --vocab-parallel-loss 1" "--sequence-parallel 4 --vocab-parallel-loss 1" \
             "--sequence-parallel 1 --vocab-parallel-loss 1 --tp-transport peer" "--sequence-parallel 1 --vocab-parallel-loss 1 --tp-transport peer-scatter"; do
  echo "=== bench $flags"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus "$N" --steps 5 --warmup 3 --layers 8 --no-cpu-baseline $flags 2>&1 | grep -v "^\[ERROR\]" | tail -4
  echo "--- exit $?"
done
