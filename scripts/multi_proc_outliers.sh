#!/bin/bash
# N independent processes (one per GPU) running the e2e loop with allocation tracing: which events coincide with slow steps?
N=${1:-4}; STEPS=${2:-400}; THREADS=${3:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i VGB_ALLOC_TRACE=1 B200SDF_TRACE=1 python scripts/e2e_sweep.py noto $STEPS $THREADS > gpurun_out/mp_$i.log 2>&1 &
done
wait
for i in $(seq 0 $((N-1))); do
  echo "== proc $i: $(tail -1 gpurun_out/mp_$i.log | cut -c1-120)"
  grep -n -E "step .* took|vgb alloc\]|b200sdf trace" gpurun_out/mp_$i.log | grep -v "top-up" | awk '/took/ {n=split($0,a," "); if (a[n-1]+0 > 3.0) print; next} {print}' | awk -F: '$1 > 400' | tail -12
done
