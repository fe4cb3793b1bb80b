#!/bin/bash
# SDF kernel form A/B on the GPU box: persistent claim loop against one CTA per tile job over a guessed grid.
cd "$(dirname "$0")/.."
run() {
  echo "$@ WL=${WL:-noto}: $(env "$@" python bench.py --kernel-only --steps 30 --warmup 5 --workload ${WL:-noto} 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'],4), 'ms  decode', round(d['detail']['decode_kernel_ms'],4), 'sdf', round(d['detail']['sdf_kernel_ms'],4), 'checksum', d['detail']['bitmap_checksum'])
")"
}
for wl in ${WLS:-noto c4 dense}; do
  export WL=$wl
  run B200SDF_SDF_KERNEL=persistent
  run B200SDF_SDF_KERNEL=strided
  run B200SDF_SDF_KERNEL=strided B200SDF_STRIDED_PER_GLYPH=1
  run B200SDF_SDF_KERNEL=strided B200SDF_STRIDED_PER_GLYPH=3
  run B200SDF_SDF_KERNEL=strided B200SDF_STRIDED_PER_GLYPH=6
done
