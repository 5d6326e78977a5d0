#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_capture.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_capture.log 2>&1
echo "exit $?" >> gpurun_out/pytest_capture.log
tail -n 8 gpurun_out/pytest_capture.log
TRACE_ENVS=512 timeout 300 python scripts/trace_fused.py > gpurun_out/trace_fused_512.log 2>&1
echo "## 512 envs (32 CTAs):"; head -n 28 gpurun_out/trace_fused_512.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.log
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'blocking', d['e2e_blocking']['ms_per_step'])
        print('policy', d['policy_step']); print('proxy', {k:v for k,v in d['gpu_proxy'].items() if 'ms' in k or 'speedup' in k}); print('parity', d.get('parity_at_bench_shape'), d.get('max_rel_err_losses'))
PY
tail -n 3 gpurun_out/bench.err
