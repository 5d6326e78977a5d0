cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "X=0" "MINPPO_EMULATE_SHARD_PAD=2" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=16" "MINPPO_EMULATE_SHARD_PAD=2 MINPPO_DW_SPLITS=18" "MINPPO_EMULATE_SHARD_PAD=3"; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror|trace" | cut -c1-60
done
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_gemm.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -3
