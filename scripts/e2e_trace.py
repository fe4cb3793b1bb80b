#!/usr/bin/env python
"""Per-worker timeline of one FontManager.render_glyphs call on the GPU box (VGB_TRACE=1), after warm-up."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (fixture paths only)
import versatiles_glyphs_rs_b200 as V  # noqa: E402

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = V.FontManager(parallel=True)
m.add_font_with_name("Noto Sans Regular", O.noto_paths())
r = V.Renderer.new_precise(device=int(os.environ.get("VGB_DEVICE", "0")))
ts = []
for i in range(30):
    w = V.Writer.new_memory()
    if os.environ.get("B200SDF_TRACE"):
        print(f"--- step {i}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    st = m.render_glyphs(w, r, threads=threads)
    ts.append((time.perf_counter() - t) * 1e3)
print("steps ms:", " ".join(f"{x:.3f}" for x in ts[5:]), file=sys.stderr)
print(f"median {sorted(ts[5:])[len(ts[5:]) // 2]:.3f} ms  workers {st.workers} submits {st.submits}", file=sys.stderr)
os.environ["VGB_TRACE"] = "1"
for k in range(2):
    print(f"--- traced call {k}", file=sys.stderr, flush=True)
    m.render_glyphs(V.Writer.new_memory(), r, threads=threads)
