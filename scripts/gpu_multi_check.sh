#!/bin/bash
# Shortest multi-GPU regression check: sharded parity + one graph-replay timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N="${1:-2}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 \
  tests/multigpu/check_sharded_update.py 2>&1 | grep -E "MULTIGPU|rror|False" | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 \
  bench.py --gpus $N --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror" | cut -c1-120
