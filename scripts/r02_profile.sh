#!/bin/bash
# round 2 profiling session (one GPU): ncu launch list, ncu --set full of the two step kernels with the caches left warm
# (--cache-control none: steady-state L2 residency, which is what the graph replay sees), compute-sanitizer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv \
  --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
cat gpurun_out/launches_summary.txt
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:"fused_step_kernel|dwopt_kernel" -s 40 -c 4 -f -o gpurun_out/prof_r02 \
  python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "exit $?" >> gpurun_out/ncu_full.log
tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
bash scripts/sanitize.sh memcheck racecheck initcheck synccheck
