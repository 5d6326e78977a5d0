#!/bin/bash
# one tuning iteration on a B200 box: parity subset (update, capture, policy tests), quick benches at D = 225 / A = 10 and D = 415 / A = 20, cycle trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_capture.py tests/test_gpu_policy.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "not configs3" > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 4 gpurun_out/pytest_update.log
for i in 1 2; do echo "## D=225/A=10"; timeout 300 python bench.py --quick --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c1-90; done
echo "## D=415/A=20"; timeout 300 python bench.py --quick --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 2>&1 | tail -n 1 | cut -c1-90
timeout 300 python scripts/trace_fused.py > gpurun_out/cycle_trace.txt 2>&1; sed -n 1,30p gpurun_out/cycle_trace.txt | cut -c1-150; grep "total cycles" gpurun_out/cycle_trace.txt
