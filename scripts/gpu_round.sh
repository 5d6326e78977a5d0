#!/bin/bash
# One GPU session: probes, parity tests (one process per file), bench, ncu launch list.
# Everything lands in gpurun_out/.  Usage: bash scripts/gpu_round.sh [quick]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
for m in 0 1 2 3; do
  timeout 120 python scripts/gemm_probe.py $m > gpurun_out/gemm_probe_$m.log 2>&1
  echo "exit $?" >> gpurun_out/gemm_probe_$m.log
done
for f in gae perm gemm update policy; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_$f.log 2>&1
  echo "exit $?" >> gpurun_out/pytest_$f.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "exit $?" >> gpurun_out/bench.log
if [ "$1" != "quick" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "exit $?" >> gpurun_out/ncu_bench.log
fi
tail -n 5 gpurun_out/gemm_probe_*.log gpurun_out/pytest_*.log gpurun_out/smoke.log
tail -c 3000 gpurun_out/bench.log
