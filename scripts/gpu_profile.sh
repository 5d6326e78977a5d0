#!/bin/bash
# ncu --set full capture of the learner step kernels (one GPU).  Output: gpurun_out/prof_<tag>.ncu-rep
cd "$(dirname "$0")/.."
TAG="${1:-r1}"
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:"umma_gemm_kernel|opt_kernel|head_loss_kernel|fused_step_kernel" -s 36 -c 6 -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "exit $?" >> gpurun_out/ncu_full_$TAG.log
tail -n 3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/*.ncu-rep
