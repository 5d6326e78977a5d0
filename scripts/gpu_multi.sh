#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded-update parity, then bench at 1..N GPUs.  Output: gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N="${1:-2}"
nvidia-smi -L > gpurun_out/nvidia_smi_multi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/multigpu/check_sharded_update.py > gpurun_out/multigpu_parity.log 2>&1
echo "exit $?" >> gpurun_out/multigpu_parity.log
grep -E "^\[|MULTIGPU|exit|Error|error" gpurun_out/multigpu_parity.log | tail -n 12
timeout 600 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_g1.log 2>&1
tail -n 1 gpurun_out/bench_g1.log
for g in $(seq 2 $N); do
  if [ $g -eq 2 ] || [ $g -eq 4 ] || [ $g -eq 8 ]; then
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 \
      --master-port 29512 bench.py --gpus $g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.log 2>&1
    echo "exit $?" >> gpurun_out/bench_g$g.log
    tail -n 2 gpurun_out/bench_g$g.log | cut -c1-900
  fi
done
