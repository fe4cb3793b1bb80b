#!/bin/bash
# ncu only (after a plain run of the same command): launch list + one full capture of each kernel.
set -u
mkdir -p gpurun_out
WL=${1:-noto}
CMD="python bench.py --kernel-only --diag --steps 3 --warmup 3 --workload $WL"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
for k in sdf_tiles_strided_kernel glyf_decode_kernel sdf_tiles_persistent_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
cut -c1-400 gpurun_out/plain.log | tail -2
