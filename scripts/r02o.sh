#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "overflow or medium" > gpurun_out/pytest_overflow.log 2>&1
echo "exit $?" >> gpurun_out/pytest_overflow.log
tail -n 6 gpurun_out/pytest_overflow.log
timeout 900 python bench.py --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 > gpurun_out/bench_d415.log 2>gpurun_out/bench_d415.err
echo "exit $?" >> gpurun_out/bench_d415.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_d415.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('D=415/A=20: ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_at_bench_shape'), d.get('max_rel_err_losses'))
        print('roofline', d['roofline']['kernel'][:40], round(d['roofline']['frac'],4), {k:round(v['ms_per_update'],3) for k,v in d['kernel_classes'].items() if v['scopes']})
        print('proxy', {k:round(v,1) for k,v in d['gpu_proxy'].items() if isinstance(v,float)})
PY
tail -n 3 gpurun_out/bench_d415.err
