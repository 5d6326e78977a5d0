#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_gemm.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not configs3" > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 4 gpurun_out/pytest_update.log
for i in 1 2; do
echo "## N-halved dW tiles (default)"
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
echo "## whole-N dW tiles (MINPPO_DW_NSPLIT=1)"
MINPPO_DW_NSPLIT=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
done
timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
grep -A6 "dwopt kernel, cycles" gpurun_out/trace_fused.log; tail -n 10 gpurun_out/trace_fused.log
