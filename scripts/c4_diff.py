"""Render C4 several ways and report where outputs differ (diagnostics)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import versatiles_glyphs_rs_b200 as V

_, fonts = bench.workload_fonts("c4")
m = V.FontManager(parallel=True)
for name, blobs in fonts:
    for b in blobs:
        m.add_font_bytes_with_name(name, b)
r = V.Renderer.new_precise(device=0)

def run(**kw):
    shards = kw.pop("shards", 1)
    got = {}
    for s in range(shards):
        w = V.Writer.new_memory()
        m.render_glyphs(w, r, shard=s, n_shards=shards, **kw)
        got.update({n: d for n, is_dir, d in w.entries() if not is_dir})
    return got

base = run()
for label, kw in (("again", {}), ("threads3", {"threads": 3}), ("shards4", {"shards": 4}), ("shards4b", {"shards": 4}), ("shards3", {"shards": 3}),
                  ("threads1", {"threads": 1})):
    other = run(**kw)
    bad = [n for n in base if base[n] != other.get(n)]
    print(label, "differing files:", bad)
    for n in bad[:3]:
        ga = {g.id: g for g in V.decode_pbf(base[n])[2]}
        gb = {g.id: g for g in V.decode_pbf(other[n])[2]}
        for i in ga:
            a, b = ga[i], gb[i]
            if (a.width, a.height, a.left, a.top, a.advance) != (b.width, b.height, b.left, b.top, b.advance):
                print("   metrics differ", hex(i), (a.width, a.height, a.left, a.top), (b.width, b.height, b.left, b.top))
            elif a.bitmap is not None and a.bitmap != b.bitmap:
                x = np.frombuffer(a.bitmap, dtype=np.uint8).astype(np.int16).reshape(a.height + 6, a.width + 6)
                y = np.frombuffer(b.bitmap, dtype=np.uint8).astype(np.int16).reshape(b.height + 6, b.width + 6)
                d = np.argwhere(x != y)
                print("   bitmap differs", hex(i), "size", x.shape, "n", len(d), "at", d[:8].tolist(), "vals", [(int(x[p[0], p[1]]), int(y[p[0], p[1]])) for p in d[:8]])
