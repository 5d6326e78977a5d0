#!/bin/bash
# quick GPU check: GEMM + update parity tests + bench (no ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gemm.log
timeout 900 python -m pytest tests/test_gpu_update.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_update.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "exit $?" >> gpurun_out/bench.log
tail -n 4 gpurun_out/pytest_gemm.log
tail -n 6 gpurun_out/pytest_update.log
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms_per_step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])
        print({k:round(v['ms_per_update'],3) for k,v in d['kernel_classes'].items()})
        print('roofline',d['roofline']['kernel'],d['roofline']['frac'],'gae',d['gae_roofline']['frac'])
    elif not l.startswith('exit 0'): print(l.rstrip()[:300])
PY
