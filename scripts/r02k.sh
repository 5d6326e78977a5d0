#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_ffi_shim_mock.py tests/test_gpu_policy.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not configs3" > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 6 gpurun_out/pytest_update.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused.log 2>&1
head -n 27 gpurun_out/trace_fused.log | tail -n 16; grep "total cycles" gpurun_out/trace_fused.log
