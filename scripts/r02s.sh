#!/bin/bash
# A/B of development builds of the library (build/dev2 = variant under test; build/dev = the same with all-warp epilogue-2 stamps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
  echo "## product"; timeout 300 python bench.py --quick --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c1-90
  echo "## variant"; MINPPO_B200_LIB=$PWD/build/dev2/libminppo_b200.so timeout 300 python bench.py --quick --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c1-90
done
TRACE_EPI2_ALL=1 MINPPO_B200_LIB=$PWD/build/dev/libminppo_b200.so timeout 300 python scripts/trace_fused.py > gpurun_out/cycle_trace_epi2.txt 2>&1
grep -n "straggler\|Error" gpurun_out/cycle_trace_epi2.txt | cut -c1-300
echo "## product D=415/A=20"; timeout 300 python bench.py --quick --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 2>&1 | tail -n 1 | cut -c1-90
echo "## variant D=415/A=20"; MINPPO_B200_LIB=$PWD/build/dev2/libminppo_b200.so timeout 300 python bench.py --quick --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 2>&1 | tail -n 1 | cut -c1-90
