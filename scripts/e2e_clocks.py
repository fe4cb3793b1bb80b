#!/usr/bin/env python
"""SM clock while FontManager.render_glyphs runs back to back (is the GPU clocked down by the bursty e2e load?)."""
import os, subprocess, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import versatiles_glyphs_rs_b200 as V
m = V.FontManager(parallel=True)
m.add_font_with_name("Noto Sans Regular", O.noto_paths())
r = V.Renderer.new_precise(device=0)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t_end = time.time() + 3.0
ts = []
while time.time() < t_end:
    t = time.perf_counter(); m.render_glyphs(V.Writer.new_memory(), r); ts.append((time.perf_counter() - t) * 1e3)
p.terminate()
out = p.stdout.read().strip().splitlines()
print("steps", len(ts), "median ms", sorted(ts)[len(ts)//2])
print("clock samples:", out[:3], "...", out[-8:])
