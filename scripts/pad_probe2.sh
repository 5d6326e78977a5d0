cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for cfg in "X=0" "MINPPO_EMULATE_SHARD_PAD=2"; do
  echo "## $cfg"
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms', round(d['ms_per_step'], 3), {k: round(v['ms_per_update'], 3) for k, v in d['kernel_classes'].items() if v['ms_per_update'] > 0})
    elif 'rror' in l: print(l.rstrip()[:200])
"
  env $cfg MINPPO_TRACE=1 MINPPO_PDL=0 timeout 200 python bench.py --quick --steps 3 --warmup 3 2>&1 | grep -E "trace" | cut -c1-100
  env $cfg timeout 200 python scripts/trace_fused.py 2>&1 | tail -n 75 | grep -vE "^\s*$" | head -80
done
