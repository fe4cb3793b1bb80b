#!/bin/bash
# e2e knob sweep on the GPU box (one process per setting)
cd "$(dirname "$0")/.."
run() { echo "$@ T=${THREADS:-8}: $(env "$@" python scripts/e2e_sweep.py ${WL:-noto} 40 ${THREADS:-8} 2>/dev/null | tail -1)"; }
for t in 4 8 16; do
THREADS=$t run VGB_GROUP_MAX=1
THREADS=$t run VGB_GROUP_MAX=4
THREADS=$t run VGB_GROUP_MAX=16
THREADS=$t run VGB_GROUP_MAX=16 VGB_BATCHES_PER_WORKER=3
THREADS=$t run VGB_GROUP_MAX=16 VGB_BATCHES_PER_WORKER=4
done
