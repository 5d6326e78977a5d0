#!/bin/bash
# ncu --set full of the two step kernels of the final build (warm caches: --cache-control none), 4 launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:"fused_step_kernel|dwopt_kernel" -s 40 -c 4 -f -o gpurun_out/prof_r02_final \
  python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "exit $?" >> gpurun_out/ncu_full.log
tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r02_final.ncu-rep
