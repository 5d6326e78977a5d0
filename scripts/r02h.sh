#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
MINPPO_PERSISTENT=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
MINPPO_PERSISTENT=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
MINPPO_PERSISTENT=1 timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused_persistent.log 2>&1
grep -A6 "persistent loop" gpurun_out/trace_fused_persistent.log; tail -n 20 gpurun_out/trace_fused_persistent.log | grep -v "phase 1 of"
