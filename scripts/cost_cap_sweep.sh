#!/bin/bash
# device-resident step against the tile-job cost cap (heavier glyphs are cut into several rectangles)
cd "$(dirname "$0")/.."
for wl in ${WLS:-noto dense c4}; do
for cap in 0 16384 32768 65536 131072 262144 1048576; do
  echo "$wl cap=$cap: $(B200SDF_GLYPH_COST_CAP=$cap python bench.py --kernel-only --steps 30 --warmup 5 --workload $wl 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'],4), 'decode', round(d['detail']['decode_kernel_ms'],4), 'sdf', round(d['detail']['sdf_kernel_ms'],4), 'tiles', d['detail']['tile_jobs'])
")"
done; done
