#!/bin/bash
# Build libminppo_b200.so for sm_100a (B200). Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${MINPPO_OUT:-$HERE/../lib}"      # MINPPO_OUT: build a development variant beside the product library
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function"
pids=()
for f in gae prng_sort minibatch head_loss policy adam learner; do
  rm -f "$OUT/$f.o"                      # a failed compile must not link a stale object
  $NVCC $FLAGS "$@" -c "$HERE/$f.cu" -o "$OUT/$f.o" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libminppo_b200.so" "$OUT"/{gae,prng_sort,minibatch,head_loss,policy,adam,learner}.o -lcudart -ldl
echo "built $OUT/libminppo_b200.so"
