#!/usr/bin/env python
"""GPU-side cost of batching (run on the GPU box): the C2 job rendered as ONE launch vs split into chunks of G glyphs
launched on n streams (device-resident arrays, no host pipeline).  Tells how much of the e2e tail is the GPU running
small kernels latency-bound."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (fixture paths only)
import versatiles_glyphs_rs_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
ctx = V.SdfContext(0, 1)
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
curves, segs, jobs = b.curves(), b.segments().copy(), b.jobs()
out_bytes = int((jobs["out_off"] + jobs["width"].astype(np.uint64) * jobs["height"]).max())


def to_dev(a):
    return torch.from_numpy(np.frombuffer(a.tobytes() or b"\0" * 16, dtype=np.uint8).copy()).to(dev)


d_curves, d_segs, d_jobs = to_dev(curves), to_dev(segs), to_dev(jobs)
d_out = torch.zeros(out_bytes + 16, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(1)


def run(G, n_streams, shuffle):
    order = rng.permutation(len(jobs)) if shuffle else np.arange(len(jobs))
    chunks = []
    for i in range(0, len(jobs), G):
        idx = np.sort(order[i:i + G])
        # tiles refer to outline jobs by index into the device job array: plan over the full array, keep my glyphs' tiles
        chunks.append(idx)
    tile_dt = np.dtype([("seg_off", "<u4"), ("seg_cnt", "<u4"), ("out_off", "<u8"), ("width", "<u2"), ("height", "<u2"),
                        ("tx0", "<u2"), ("ty0", "<u2"), ("ntx", "<u2"), ("nty", "<u2"), ("job", "<u4")])
    plans = []
    for idx in chunks:
        # plan the chunk on its own (the planner's small-batch rule sees the chunk's cost), then point the tiles at
        # the glyphs' indices in the full device job array
        t, n, _ = ctx.plan_outline_tiles(jobs[idx], len(curves), len(segs), out_bytes)
        t = t.view(tile_dt)[:n].copy()
        t["job"] = idx[t["job"]]
        plans.append((torch.from_numpy(t.view(np.uint8).copy()).to(dev), n))
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in streams:
            s.wait_event(a)
        for k, (dt, n) in enumerate(plans):
            s = streams[k % n_streams]
            ctx.render_outlines_device(d_curves.data_ptr(), d_segs.data_ptr(), d_jobs.data_ptr(), dt.data_ptr(), n,
                                       d_out.data_ptr(), s.cuda_stream)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            torch.cuda.current_stream().wait_event(ev)
        e.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, a.elapsed_time(e))
    return best, len(plans)


print("one launch:", "%.3f ms" % run(len(jobs), 1, False)[0])
for G in (64, 128, 256):
    for ns in (1, 32):
        ms, n = run(G, ns, True)
        print(f"G={G:4d} ({n:3d} launches) streams={ns:2d}: {ms:.3f} ms")
