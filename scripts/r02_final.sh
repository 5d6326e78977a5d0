#!/bin/bash
# final check of round 2: full `pytest -m gpu`, smoke(), bench (both arms), bench at D = 415 / A = 20, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=4 > gpurun_out/pytest_gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu_full.log
tail -n 10 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --obs-dim 415 --act-dim 20 > gpurun_out/bench_d415.log 2>gpurun_out/bench_d415.err
echo "exit $?" >> gpurun_out/bench_d415.log
python - <<'PY'
import json
for f in ('gpurun_out/bench.log', 'gpurun_out/bench_d415.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, 'ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_at_bench_shape'), 'roofline', d['roofline']['kernel'][:24], round(d['roofline']['frac'],4), 'gae', round(d['gae_roofline']['frac'],3))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv \
  --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
head -n 4 gpurun_out/launches_summary.txt
