#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not configs3 and not configs1" > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 4 gpurun_out/pytest_update.log
echo "## D=225/A=10"; timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 1 | cut -c1-80
echo "## D=415/A=20"; timeout 300 python bench.py --quick --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 2>&1 | tail -n 1 | cut -c1-80
echo "## D=415/A=20, whole-N dW tiles"; MINPPO_DW_NSPLIT=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 2>&1 | tail -n 1 | cut -c1-80
