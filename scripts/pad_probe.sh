cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "X=0" "MINPPO_EMULATE_SHARD_PAD=2" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_PDL=0" "MINPPO_PDL=0" $EXTRA; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror|trace" | cut -c1-100
done
