#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded-update parity (peer-memory exchange and NCCL fallback), then bench at N GPUs
# with both exchange paths.  Output: gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N="${1:-2}"
nvidia-smi -L > gpurun_out/nvidia_smi_multi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/nvidia_smi_multi.txt 2>&1
for mode in 0 1; do
  MINPPO_NCCL_ALLREDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$mode tests/multigpu/check_sharded_update.py > gpurun_out/multigpu_parity_nccl$mode.log 2>&1
  echo "exit $?" >> gpurun_out/multigpu_parity_nccl$mode.log
  echo "## parity, MINPPO_NCCL_ALLREDUCE=$mode"
  grep -E "^\[(small|wide)|MULTIGPU|^exit|MinppoError" gpurun_out/multigpu_parity_nccl$mode.log | tail -n 6
done
MINPPO_TRACE=1 MINPPO_PDL=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29530 bench.py --gpus $N --quick --steps 3 --warmup 3 > gpurun_out/bench_trace_g$N.log 2>&1
grep "dwopt trace" gpurun_out/bench_trace_g$N.log
timeout 600 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_g1.log 2>&1
tail -n 1 gpurun_out/bench_g1.log | cut -c1-200
for g in 2 4 8; do
  if [ $g -le $N ]; then
    for mode in 0 1; do
      MINPPO_NCCL_ALLREDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 \
        --master-port 2952$mode bench.py --gpus $g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g${g}_nccl$mode.log 2>&1
      echo "exit $?" >> gpurun_out/bench_g${g}_nccl$mode.log
      echo "## bench $g GPUs, MINPPO_NCCL_ALLREDUCE=$mode"
      python - <<PY
import json
for l in open('gpurun_out/bench_g${g}_nccl$mode.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms_per_step', round(d['ms_per_step'], 3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
        print({k: round(v['ms_per_update'], 3) for k, v in d['kernel_classes'].items() if v['ms_per_update'] > 0})
    elif 'rror' in l:
        print(l.rstrip()[:300])
PY
    done
  fi
done
