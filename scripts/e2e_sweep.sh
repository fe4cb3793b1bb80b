#!/bin/bash
# e2e tuning sweep on the GPU box: part size x batches per worker (median ms of FontManager.render_glyphs, C2)
for pg in 16 32 64; do for bw in 2 3 4 6; do
  echo -n "part_glyphs=$pg batches_per_worker=$bw: "
  VGB_PART_GLYPHS=$pg VGB_BATCHES_PER_WORKER=$bw python scripts/e2e_trace.py 2>&1 | grep median
done; done
