cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=14" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=15" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=17" "MINPPO_DW_SPLITS=17" "MINPPO_DW_SPLITS=18" "MINPPO_DW_SPLITS=14" "MINPPO_EMULATE_SHARD_PAD=3 MINPPO_DW_SPLITS=16" "MINPPO_EMULATE_SHARD_PAD=3 MINPPO_DW_SPLITS=18"; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror|trace" | cut -c1-60
done
