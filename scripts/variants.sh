#!/bin/bash
# Build kernel variants (different -D flags) into build/variants/<name>/ and time each on the GPU box.
# usage: scripts/variants.sh build   (here, no GPU)      scripts/variants.sh run [workload]   (under gpurun)
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
declare -A V=(
  [base]=""
  [g5]="-DB200SDF_GLYF_MIN_CTAS=5 -DB200SDF_GLYF_MAX_POINTS=1792"
  [g6]="-DB200SDF_GLYF_MIN_CTAS=6 -DB200SDF_GLYF_MAX_POINTS=1536"
  [g8]="-DB200SDF_GLYF_MIN_CTAS=8 -DB200SDF_GLYF_MAX_POINTS=1024"
)
if [ "${1:-}" = "build" ]; then
  for n in "${!V[@]}"; do
    d=$ROOT/build/variants/$n; mkdir -p $d
    make -s -C $ROOT/versatiles_glyphs_rs_b200/csrc OUT=$d EXTRA="${V[$n]}" 2>&1 | grep -E "error" 
    echo "$n: $(grep -E 'Used [0-9]+ registers' $d/libb200sdf.ptxas.log | tail -1)"
  done
else
  WL=${2:-noto}
  mkdir -p $ROOT/gpurun_out
  for n in $(ls $ROOT/build/variants); do
    VGB200_LIBDIR=$ROOT/build/variants/$n python $ROOT/bench.py --kernel-only --diag --steps 30 --warmup 5 --workload $WL 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$n', round(d['ms_per_step'],4), 'ms  decode', round(d['detail']['decode_kernel_ms'],4), 'sdf', round(d['detail']['sdf_kernel_ms'],4), 'host-planned', round(d['detail']['host_planned_sdf_kernel_ms'],4), 'devplan-1cta', round(d['detail']['device_plan_in_one_cta_per_job_kernel_ms'],4))
    elif 'Error' in l or 'error' in l: print('$n', l.strip())
" | tee -a $ROOT/gpurun_out/variants_$WL.txt
  done
fi
