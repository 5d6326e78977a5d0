cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "MINPPO_EMULATE_SHARD_PAD=2" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=16" "MINPPO_EMULATE_SHARD_PAD=3" "MINPPO_EMULATE_SHARD_PAD=3 MINPPO_DW_SPLITS=16" "MINPPO_DW_SPLITS=16" "MINPPO_DW_SPLITS=15"; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror|trace" | cut -c1-100
done
MINPPO_EMULATE_SHARD_PAD=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_pad.csv python bench.py --quick --steps 2 --warmup 3 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/launches_pad.csv | head -8
