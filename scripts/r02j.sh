#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "## default build (MMA thread busy-polls)"
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
echo "## lib_alt (MMA thread uses the suspending try_wait)"
MINPPO_B200_LIB=$PWD/minppo_b200/lib_alt/libminppo_b200.so timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
MINPPO_B200_LIB=$PWD/minppo_b200/lib_alt/libminppo_b200.so timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
MINPPO_B200_LIB=$PWD/minppo_b200/lib_alt/libminppo_b200.so timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused_alt.log 2>&1
head -n 27 gpurun_out/trace_fused_alt.log
timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
head -n 27 gpurun_out/trace_fused.log
