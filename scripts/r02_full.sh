#!/bin/bash
# the driver's round-end sequence, run by the builder: full `pytest -m gpu`, smoke(), bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu_full.log
tail -n 14 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.log
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_at_bench_shape'), 'launches', d['gpu_launches'])
        print('roofline', d['roofline']['kernel'][:30], round(d['roofline']['frac'],4), 'gae', round(d['gae_roofline']['frac'],3), 'policy', d['policy_step'])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -n 1 | cut -c1-400
