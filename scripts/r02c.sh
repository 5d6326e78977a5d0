#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
echo "exit $?" >> gpurun_out/trace_fused.log
tail -n 75 gpurun_out/trace_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv \
  --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
cat gpurun_out/launches_summary.txt
