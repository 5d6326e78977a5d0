cd "$(dirname "$0")/.."; mkdir -p gpurun_out
MINPPO_EMULATE_SHARD_PAD=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwopt_kernel" -s 20 -c 1 -f -o gpurun_out/prof_dwopt_pad \
    python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_dwopt_pad.log 2>&1
tail -n 2 gpurun_out/ncu_dwopt_pad.log
