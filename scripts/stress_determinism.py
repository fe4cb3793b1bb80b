"""Render a workload repeatedly and count runs whose output differs from the majority (race hunting)."""
import hashlib, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import versatiles_glyphs_rs_b200 as V
wl = sys.argv[1] if len(sys.argv) > 1 else "noto"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
_, fonts = bench.workload_fonts(wl)
m = V.FontManager(parallel=True)
for name, blobs in fonts:
    for b in blobs:
        m.add_font_bytes_with_name(name, b)
r = V.Renderer.new_precise(device=0)
import numpy as np
runs = []
raw = []
for i in range(reps):
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, r, threads=threads)
    ent = {n: d for n, is_dir, d in w.entries() if not is_dir}
    raw.append(ent)
    runs.append({n: hashlib.sha1(d).hexdigest() for n, d in ent.items()})
files = sorted(runs[0])
bad_runs = 0
for n in files:
    c = collections.Counter(run[n] for run in runs)
    if len(c) > 1:
        major = c.most_common(1)[0][0]
        odd = [i for i, run in enumerate(runs) if run[n] != major]
        print("  ", n, "differs in runs", odd)
        bad_runs += 1
        good = next(i for i, run in enumerate(runs) if run[n] == major)
        ga = {g.id: g for g in V.decode_pbf(raw[good][n])[2]}
        gb = {g.id: g for g in V.decode_pbf(raw[odd[0]][n])[2]}
        # index of every good bitmap of the job, to find out where a wrong one came from
        origin = {}
        for fn, blob in raw[good].items():
            for gg in V.decode_pbf(blob)[2]:
                if gg.bitmap is not None:
                    origin.setdefault(bytes(gg.bitmap), []).append(gg.id)
        for gid in ga:
            a, b = ga[gid], gb[gid]
            if (a.width, a.height, a.left, a.top, a.advance) != (b.width, b.height, b.left, b.top, b.advance):
                print("      metrics differ", hex(gid))
            elif a.bitmap != b.bitmap:
                x = np.frombuffer(a.bitmap, dtype=np.uint8).astype(np.int16).reshape(a.height + 6, a.width + 6)
                y = np.frombuffer(b.bitmap, dtype=np.uint8).astype(np.int16).reshape(b.height + 6, b.width + 6)
                d = np.argwhere(x != y)
                print("      wrong bitmap equals good bitmap of", [hex(v) for v in origin.get(bytes(b.bitmap), [])][:4], "zeros" if not any(b.bitmap) else "")
                print("      bitmap differs", hex(gid), "size", x.shape, "n", len(d), "first", d[:6].tolist(), "rows", sorted(set(d[:, 0].tolist()))[:12],
                      "vals", [(int(x[p[0], p[1]]), int(y[p[0], p[1]])) for p in d[:6]])
print(f"{wl} flatten={os.environ.get('VGB_FLATTEN','glyf')} threads={threads}: {reps} runs, {bad_runs} files with differing content, handed_back {st.handed_back}")
