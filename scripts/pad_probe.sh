#!/bin/bash
# Single-GPU timing probes for the env-sharded code path (no second GPU needed):
#   MINPPO_EMULATE_SHARD_PAD=2  row lists / grids sized like a sharded rank's worst case (1.5 x mean + 256 rows)
#   MINPPO_DW_SPLITS=S          force the split-K factor of the weight-gradient GEMM
#   MINPPO_BENCH_ENABLE_P2P=1   enable peer access on the context (needs a 2-GPU box): cost per kernel boundary
#   MINPPO_NO_FORK=1            operand staging on the main stream instead of the forked graph branch
# Round-1 findings: padded shapes cost 1.5 ms / update while the split-K factor grew to 18 and
# left 4 spare CTAs for the L2 prefetch; peer mappings cost 0.34 ms / update.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "X=0" "MINPPO_EMULATE_SHARD_PAD=2" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=16" "MINPPO_NO_FORK=1" "MINPPO_PDL=0" $EXTRA; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror|trace" | cut -c1-100
done
