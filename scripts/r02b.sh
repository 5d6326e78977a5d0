#!/bin/bash
# round 2, GPU call B: new fused step kernel (generic D / A, 16 worker warps) + multi-unit dwopt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "not configs3" --durations=5 > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
tail -n 40 gpurun_out/pytest_update.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -n 3
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-proxy > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.log
tail -c 1800 gpurun_out/bench.log; tail -n 3 gpurun_out/bench.err
