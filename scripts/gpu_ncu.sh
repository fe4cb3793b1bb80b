#!/bin/bash
# ncu only (after a plain run of the same command): launch list + one full capture of the SDF kernel.
set -u
mkdir -p gpurun_out
WL=${1:-noto}
python bench.py --kernel-only --steps 3 --warmup 3 --workload $WL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --kernel-only --steps 3 --warmup 3 --workload $WL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py --kernel-only --steps 3 --warmup 3 --workload $WL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sdf_tiles -s 3 -c 2 -f -o gpurun_out/prof_sdf \
    python bench.py --kernel-only --steps 3 --warmup 3 --workload $WL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/plain.log | cut -c1-300
