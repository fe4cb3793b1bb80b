#!/usr/bin/env python
"""One traced FontManager.render_glyphs call on C2 (after 30 warm-up calls): where the step goes and how busy the GPU is.

    python scripts/e2e_timeline.py [threads] > profiles/rNN_e2e_timeline.txt      (on the GPU box)

Runs scripts/e2e_trace.py with B200SDF_GPU_TRACE=1 (CUDA events around every submission's kernels, reported when the
submission is reaped) and VGB_TRACE=1 (host threads), keeps the last call and summarises it.  Tracing itself costs
time (two extra event records per submission and the log lines): compare shares, not the absolute step time."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
threads = sys.argv[1] if len(sys.argv) > 1 else "8"
env = dict(os.environ, B200SDF_GPU_TRACE="1")
p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "e2e_trace.py"), threads], env=env, capture_output=True, text=True)
lines = p.stderr.splitlines()
last = max(i for i, l in enumerate(lines) if l.startswith("--- traced call"))
untraced = [l for l in lines if l.startswith("median")]
call = lines[last + 1:]
gpu = []
for l in call:
    m = re.match(r"\[b200sdf gpu\] submit ([\d.]+) us\s+kernel ([\d.]+) \.\. ([\d.]+) us \(([\d.]+) us, (\d+) CTAs\)\s+seen ([\d.]+) us", l)
    if m:
        gpu.append(tuple(float(x) for x in m.groups()))
host = [l for l in call if l.startswith("[vgb trace]")]
print(f"== C2 (Noto Sans merge), FontManager.render_glyphs, {threads} host threads: one traced call after 30 warm-up calls ==")
if untraced:
    print("untraced calls of the same process:", untraced[-1])
if gpu:
    t0 = min(g[0] for g in gpu)
    ivs = sorted((g[1], g[2]) for g in gpu)
    busy, cur_a, cur_b = 0.0, ivs[0][0], ivs[0][1]
    for a, b in ivs[1:]:
        if a > cur_b:
            busy += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    busy += cur_b - cur_a
    span = max(g[5] for g in gpu) - t0
    total = None
    for l in host:
        m = re.search(r"total ([\d.]+) us", l)
        if m:
            total = float(m.group(1))
    print(f"submissions {len(gpu)} (queued batches share a submission); first submit .. last completion seen {span:.0f} us"
          + (f"; call wall time (traced) {total:.0f} us" if total else ""))
    print(f"GPU busy (union of the submissions' [decode + SDF] kernel intervals, CUDA events): {busy:.0f} us = "
          f"{100 * busy / span:.0f} % of first submit .. last completion" + (f", {100 * busy / total:.0f} % of the call" if total else ""))
    print(f"sum of the submissions' kernel durations {sum(g[3] for g in gpu):.0f} us (they overlap on the device)")
    print("per submission: submit (host clock, us since the call's first submit) | kernels start .. end | duration | glyph requests | seen by host")
    for g in sorted(gpu):
        print(f"  submit {g[0] - t0:7.1f}  kernels {g[1] - t0:7.1f} .. {g[2] - t0:7.1f}  ({g[3]:6.1f} us, {int(g[4]):5d} requests)  seen {g[5] - t0:7.1f}")
print("host threads (us since the call began; B begin, o batch opened, s handed to the CUDA thread, e encode, F files written, E end;")
print("              last line = the pumping thread: s/S submit begin/end, d completion seen):")
for l in host:
    print("  " + l)
