cd /root/repo; mkdir -p gpurun_out
for cfg in "X=0" "MINPPO_BENCH_ENABLE_P2P=1" "X=1"; do
  echo "## $cfg"; env $cfg timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | grep -E "quick|rror" | cut -c1-90
done
