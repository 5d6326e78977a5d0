#!/bin/bash
# Multi-GPU timing ablations (gpurun --gpus N): graph-replay ms per update with pieces of the sharded path switched off.
# TIMING ONLY -- the MINPPO_PX_ABLATE runs compute wrong gradients.  Output: gpurun_out/multi_ablate_gN.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N="${1:-2}"
OUT=gpurun_out/multi_ablate_g$N.log
: > $OUT
port=29600
for cfg in "MINPPO_PDL=1" "MINPPO_PX_ABLATE=1" "MINPPO_PX_ABLATE=2" "MINPPO_SHARE_PERM=0" "MINPPO_PDL=0" "MINPPO_NCCL_ALLREDUCE=1" $EXTRA_ABLATE; do
  port=$((port + 1))
  echo "## $cfg" >> $OUT
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror" | cut -c1-160 >> $OUT
done
cat $OUT
