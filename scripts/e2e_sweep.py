"""e2e timing of FontManager.render_glyphs under the environment's tuning knobs (run one process per setting:
the knobs are read once).  usage: python scripts/e2e_sweep.py [noto|c4|dense|fira] [steps] [threads] [copies]"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (workload helpers only)
import versatiles_glyphs_rs_b200 as V  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "noto"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
copies = int(sys.argv[4]) if len(sys.argv) > 4 else 1
_, fonts = bench.workload_fonts(wl, copies)
m = V.FontManager(parallel=True)
for name, blobs in fonts:
    for b in blobs:
        m.add_font_bytes_with_name(name, b)
r = V.Renderer.new_precise(device=0)
for _ in range(8):
    st = m.render_glyphs(V.Writer.new_memory(), r, threads=threads)
import gc

gc.collect()
gc.disable()
ms = []
for k in range(steps):
    if os.environ.get("VGB_ALLOC_TRACE"):
        print(f"--- step {k}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    st = m.render_glyphs(V.Writer.new_memory(), r, threads=threads)
    ms.append(1e3 * (time.perf_counter() - t))
    if os.environ.get("VGB_ALLOC_TRACE"):
        print(f"--- step {k} took {ms[-1]:.2f} ms", file=sys.stderr, flush=True)
knobs = {k: v for k, v in os.environ.items() if k.startswith(("VGB_", "B200SDF_"))}
print(f"{wl} x{copies} threads={threads or 'all'} {knobs}: min/median/max {min(ms):.3f}/{statistics.median(ms):.3f}/{max(ms):.3f} ms  "
      f"workers {st.workers} submits {st.submits} req/w {st.outline_ns / 1e6 / st.workers:.3f} enc/w {st.encode_ns / 1e6 / st.workers:.3f} "
      f"wait/w {st.wait_ns / 1e6 / st.workers:.3f} sub/w {st.submit_ns / 1e6 / st.workers:.3f}")
if os.environ.get("VGB_TRACE"):
    pass
