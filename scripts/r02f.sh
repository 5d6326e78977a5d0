#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not configs3" --durations=5 > gpurun_out/pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/pytest_all.log
tail -n 15 gpurun_out/pytest_all.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1
MINPPO_PERSISTENT=0 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1
timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
grep -A6 "persistent loop" gpurun_out/trace_fused.log; tail -n 22 gpurun_out/trace_fused.log
