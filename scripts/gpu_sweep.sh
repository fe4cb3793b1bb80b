#!/bin/bash
# e2e knob sweep on the GPU box (one process per setting)
cd "$(dirname "$0")/.."
run() { echo "$@ T=${THREADS:-8}: $(env "$@" python scripts/e2e_sweep.py ${WL:-noto} 60 ${THREADS:-8} 2>/dev/null | tail -1 | sed 's/.*: min/min/')"; }
for t in 8 16; do
THREADS=$t run A=1
THREADS=$t run VGB_PART_GLYPHS=64 VGB_OPEN_GLYPHS=64
THREADS=$t run VGB_PART_GLYPHS=64 VGB_OPEN_GLYPHS=128
THREADS=$t run VGB_PART_GLYPHS=128 VGB_OPEN_GLYPHS=128
THREADS=$t run VGB_PART_GLYPHS=32 VGB_OPEN_GLYPHS=32
THREADS=$t run VGB_PART_GLYPHS=128 VGB_OPEN_GLYPHS=192
done
