#!/usr/bin/env python
"""Kernel-only probes (run on the GPU box): where does the SDF kernel lose time?

  uniform   N copies of ONE glyph (no load imbalance, no tail) -> the per-CTA code's own efficiency
  c2        the Noto merge batch (what bench.py times)
  c2x4      the same batch four times in one launch (amortises ramp/tail)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (fixture paths only)
import versatiles_glyphs_rs_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
ctx = V.SdfContext(0, 1)
peak, _ = ctx.measure_fp32_peak(3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)


def to_dev(a):
    return torch.from_numpy(np.frombuffer(a.tobytes() or b"\0" * 16, dtype=np.uint8).copy()).to(dev)


def time_batch(name, curves, segs, jobs, steps=20):
    out_bytes = int((jobs["out_off"] + jobs["width"].astype(np.uint64) * jobs["height"]).max())
    tiles, n_tiles, pairs = ctx.plan_outline_tiles(jobs, len(curves), len(segs), out_bytes)
    d = [to_dev(x) for x in (curves, segs, jobs)] + [torch.from_numpy(tiles).to(dev)]
    d_out = torch.zeros(out_bytes + 16, dtype=torch.uint8, device=dev)
    ms = []
    with torch.cuda.stream(stream):
        for i in range(steps + 3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.render_outlines_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n_tiles,
                                       d_out.data_ptr(), stream.cuda_stream)
            b.record(stream)
            ms.append((a, b))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ms[3:]])
    tf = 11 * pairs / (t.mean() * 1e-3) / 1e12
    px = int((jobs["width"].astype(np.int64) * jobs["height"]).sum())
    print(f"{name:28s} {t.mean():8.4f} ms (min {t.min():.4f})  ctas {n_tiles:6d}  pairs {pairs:.3e}  {tf:6.2f} TFLOP/s  "
          f"frac {tf / peak:.4f}  glyphs/s {len(jobs) / (t.mean() * 1e-3):.3e}  px {px}")
    return t.mean()


def noto_batch():
    fonts = [V.FontFileEntry(path=p) for p in O.noto_paths()]
    owner = {}
    for f in fonts:
        for cp in f.codepoints().tolist():
            if cp <= 0xFFFF and cp not in owner:
                owner[cp] = f
    r = V.Renderer.new_dummy()
    b = r.new_batch()
    for cp in sorted(owner):
        b.add_glyph(owner[cp], cp)
    return b.curves(), b.segments().copy(), b.jobs()


curves, segs, jobs = noto_batch()
print("fp32 peak", round(peak, 2), "TFLOP/s; lib", os.environ.get("VGB200_LIBDIR", "in-tree"))
time_batch("c2", curves, segs, jobs)

# the same batch 4x in one launch
j4 = np.concatenate([jobs] * 4)
stride = int((jobs["out_off"] + jobs["width"].astype(np.uint64) * jobs["height"]).max())
j4["out_off"] = np.concatenate([jobs["out_off"] + k * stride for k in range(4)])
time_batch("c2 x4 in one launch", curves, segs, j4)

# uniform: N copies of one glyph near the median (S ~ 528, 20x24) and of a heavy one
order = np.argsort(jobs["seg_cnt"])
for label, idx, n in (("uniform median glyph", order[len(order) // 2], 16384), ("uniform p99 glyph", order[int(len(order) * 0.99)], 4096),
                      ("uniform small glyph", order[len(order) // 10], 32768)):
    j = jobs[idx]
    ju = np.repeat(j[None], n)
    sz = int(j["width"]) * int(j["height"])
    ju["out_off"] = np.arange(n, dtype=np.uint64) * sz
    print(f"  {label}: S={int(j['seg_cnt'])} {int(j['width'])}x{int(j['height'])}")
    time_batch(label, curves, segs, ju)
