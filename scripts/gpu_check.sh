#!/bin/bash
# Runs on the GPU box under gpurun: GPU tests, smoke, a bench line, then (only if all of that exited 0)
# the ncu launch list and one full capture of the SDF kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; rc=$?
tail -25 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && { echo "GPU TESTS FAILED rc=$rc"; exit $rc; }
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || { cat gpurun_out/smoke.log; echo "SMOKE FAILED"; exit 1; }
cat gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err || { tail -20 gpurun_out/bench.err; echo "BENCH FAILED"; exit 1; }
cat gpurun_out/bench.json
if [ "${1:-}" = "ncu" ]; then
  python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sdf_tiles -s 3 -c 2 -f -o gpurun_out/prof_sdf \
      python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
