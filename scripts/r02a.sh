#!/bin/bash
# round 2, GPU call A: new parity tests at the bench shapes, bench (whole-update CPU arm, GPU proxy, parity flag),
# sanitizer, full GAE sweep + ncu dram counters.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1; nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests/test_gpu_update.py -m gpu -q --timeout=900 -p no:cacheprovider -k "bench_shape or negative_lr or test_update_parity" --durations=8 > gpurun_out/pytest_update_new.log 2>&1
echo "exit $?" >> gpurun_out/pytest_update_new.log
tail -n 25 gpurun_out/pytest_update_new.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.log
tail -c 2500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
bash scripts/sanitize.sh memcheck racecheck initcheck synccheck
timeout 900 python scripts/gae_sweep.py > gpurun_out/gae_sweep.jsonl 2>&1
tail -n 12 gpurun_out/gae_sweep.jsonl
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
  --clock-control none -k regex:gae --csv --log-file gpurun_out/gae_ncu.csv python scripts/gae_ncu_target.py > gpurun_out/gae_ncu.log 2>&1
echo "exit $?" >> gpurun_out/gae_ncu.log
tail -n 14 gpurun_out/gae_ncu.csv
