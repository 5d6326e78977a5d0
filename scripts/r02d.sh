#!/bin/bash
# round 2, GPU call D: persistent all-steps kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not configs3" --durations=5 > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 30 gpurun_out/pytest_update.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 2
MINPPO_PERSISTENT=0 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1
MINPPO_STEPS_PER_LAUNCH=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1
timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
tail -n 22 gpurun_out/trace_fused.log
