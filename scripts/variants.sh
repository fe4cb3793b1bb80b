#!/bin/bash
# Build kernel variants (different -D flags) into build/variants/<name>/ and time each on the GPU box.
# usage: scripts/variants.sh build   (here, no GPU)      scripts/variants.sh run [workload]   (under gpurun)
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
declare -A V=(
  [base]=""
  [pack0]="-DB200SDF_PACK=0"
  [pack1]="-DB200SDF_PACK=1"
  [unroll2]="-DB200SDF_UNROLL=2"
  [mini32]="-DB200SDF_MINI=32"
  [mini128]="-DB200SDF_MINI=128"
  [mini128p1]="-DB200SDF_MINI=128 -DB200SDF_PACK=1"
  [mini96]="-DB200SDF_MINI=96"
  [t4x2]="-DB200SDF_TILE_W=4 -DB200SDF_TILE_H=2 -DB200SDF_MAX_ITEMS=128"
  [t4x2p1]="-DB200SDF_TILE_W=4 -DB200SDF_TILE_H=2 -DB200SDF_MAX_ITEMS=128 -DB200SDF_PACK=1"
  [t8x2]="-DB200SDF_TILE_W=8 -DB200SDF_TILE_H=2"
  [csm128]="-DB200SDF_CURVE_SMEM=128"
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
    VGB200_LIBDIR=$ROOT/build/variants/$n python $ROOT/bench.py --kernel-only --steps 30 --warmup 5 --workload $WL 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$n', round(d['ms_per_step'],4), 'ms  frac', round(d['roofline']['frac'],4), 'ctas', d['config']['ctas'])
    elif 'Error' in l or 'error' in l: print('$n', l.strip())
" | tee -a $ROOT/gpurun_out/variants_$WL.txt
  done
fi
