"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json): glyph metrics and PBF structure bit-exact; SDF bitmaps within +-1 u8 per pixel
with >= 99.9 % of pixels identical.  Floating point tolerance is therefore stated in u8 steps:
MAX_DIFF = 1, MIN_IDENTICAL = 0.999.
"""
import os

import numpy as np
import pytest

import oracle_lib as O
import synth_font
import versatiles_glyphs_rs_b200 as V

pytestmark = pytest.mark.gpu

MAX_DIFF = 1
MIN_IDENTICAL = 0.999


@pytest.fixture(scope="module")
def renderer():
    return V.Renderer.new_precise(device=0)


@pytest.fixture(scope="module")
def ctx():
    return V.SdfContext(device=0, n_slots=2)


JOB_DT = np.dtype([("seg_off", "<u4"), ("seg_cnt", "<u4"), ("width", "<u4"), ("height", "<u4"), ("out_off", "<u8")])


def compare_bitmaps(got, want, what=""):
    d = np.abs(got.astype(np.int16).reshape(-1) - want.astype(np.int16).reshape(-1))
    assert d.max(initial=0) <= MAX_DIFF, (what, int(d.max()))
    return d.size, int((d == 0).sum())


def check_pbf_block(got_bytes, want_bytes, what):
    """PBF structure + metrics bit-exact, bitmaps within tolerance; returns (pixels, identical)."""
    gname, grange, gg = O.decode_pbf(got_bytes)
    wname, wrange, wg = O.decode_pbf(want_bytes)
    assert (gname, grange) == (wname, wrange), what
    assert len(got_bytes) == len(want_bytes), what
    assert [g["id"] for g in gg] == [g["id"] for g in wg], what
    px = same = 0
    for a, b in zip(gg, wg):
        for k in ("width", "height", "left", "top", "advance"):
            assert a[k] == b[k], (what, a["id"], k)
        assert (a["bitmap"] is None) == (b["bitmap"] is None)
        if a["bitmap"] is not None:
            assert len(a["bitmap"]) == (a["width"] + 6) * (a["height"] + 6)  # renderer.rs:164-166
            p, s = compare_bitmaps(a["bitmap"], b["bitmap"], (what, a["id"]))
            px += p
            same += s
    return px, same


# ---- reference golden tests, now through the CUDA path ---------------------------------------------------
def test_square_golden(ctx):
    """reference src/render/renderer_precise.rs:91-135: axis-aligned square in a 10x10 bitmap."""
    ring = np.array([(1, 2), (5, 2), (5, 6), (1, 6), (1, 2)], dtype=np.float64)
    x0, y0, W, H = -2, -1, 10, 10
    segs = np.concatenate([ring[:-1], ring[1:]], axis=1) - np.array([x0, y0, x0, y0])
    jobs = np.array([(0, 4, W, H, 0)], dtype=JOB_DT)
    bm = ctx.render(segs.astype(np.float32), jobs, W * H)
    want = O.renderer_precise(x0, y0, W, H, [ring.tolist()])
    assert np.array_equal(bm, want)
    assert O.bitmap_as_digit_art(bm, W) == [
        "30 38 42 43 43 43 43 42 38 30",
        "38 48 54 55 55 55 55 54 48 38",
        "42 54 65 68 68 68 68 65 54 42",
        "43 55 68 80 80 80 80 68 55 43",
        "43 55 68 80 93 93 80 68 55 43",
        "43 55 68 80 93 93 80 68 55 43",
        "43 55 68 80 80 80 80 68 55 43",
        "42 54 65 68 68 68 68 65 54 42",
        "38 48 54 55 55 55 55 54 48 38",
        "30 38 42 43 43 43 43 42 38 30",
    ]


def test_render_glyph_goldens(renderer):
    """reference src/render/renderer.rs:176-287 — metrics exact, art bands as in the reference."""
    font, ofont = V.FontFileEntry(path=O.FIRA), O.Font(O.FIRA)
    g = renderer.render_glyph(font, 0x20)
    assert (g.width, g.height, g.left, g.top, g.advance, g.bitmap) == (0, 0, 0, 0, 6, None)
    for cp, metrics in ((0x41, (14, 17, 0, -7, 13)), (0xE6, (19, 14, 0, -11, 19)), (0x60, (7, 5, 0, -4, 7))):
        g = renderer.render_glyph(font, cp)
        assert (g.width, g.height, g.left, g.top, g.advance) == metrics
        assert len(g.bitmap) == (g.width + 6) * (g.height + 6)
        want = ofont.render_glyph(cp)
        compare_bitmaps(np.frombuffer(g.bitmap, dtype=np.uint8), want["bitmap"], hex(cp))
    assert renderer.render_glyph(font, 0xD800) is None
    g = renderer.render_glyph(font, 0x41)
    art = O.bitmap_as_ascii_art(np.frombuffer(g.bitmap, dtype=np.uint8), g.width + 6)
    want = O.bitmap_as_ascii_art(ofont.render_glyph(0x41)["bitmap"], g.width + 6)
    assert art == want


# ---- C1 / C2: whole fonts through FontManager::render_glyphs ---------------------------------------------
def _render_all(name, paths, renderer, **kw):
    m = V.FontManager(parallel=True)
    m.add_font_with_name(name, paths)
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, renderer, **kw)
    return {n: d for n, is_dir, d in w.entries() if not is_dir}, st, m


def test_c1_fira_recurse_parity(renderer):
    files, st, _ = _render_all("Fira Sans - Regular", [O.FIRA], renderer)
    oset = O.FontSet("Fira Sans - Regular", [O.FIRA])
    assert (st.glyphs, st.bitmaps, st.segments, st.pixels) == (1686, 1679, 600952, 758736)
    px = same = 0
    for b in range(256):
        p, s = check_pbf_block(files[f"fira_sans_regular/{b * 256}-{b * 256 + 255}.pbf"], oset.render_block(b), b)
        px += p
        same += s
    assert px == 758736
    assert same / px >= MIN_IDENTICAL, same / px
    print(f"C1 Fira: {px} px, {px - same} differ by 1 ({100 * same / px:.4f}% identical)")


def test_c2_noto_merge_parity(renderer):
    paths = O.noto_paths()
    files, st, _ = _render_all("Noto Sans Regular", paths, renderer)
    oset = O.FontSet("Noto Sans Regular", paths)
    assert (st.glyphs, st.bitmaps, st.segments, st.pixels) == (6480, 6445, 3956999, 3295280)
    pop = oset.block_population()
    px = same = 0
    for b in range(256):
        if pop[b] == 0:
            assert files[f"noto_sans_regular/{b * 256}-{b * 256 + 255}.pbf"] == oset.render_block(b)
            continue
        p, s = check_pbf_block(files[f"noto_sans_regular/{b * 256}-{b * 256 + 255}.pbf"], oset.render_block(b), b)
        px += p
        same += s
    assert px == 3295280
    assert same / px >= MIN_IDENTICAL, same / px
    print(f"C2 Noto merge: {px} px, {px - same} differ by 1 ({100 * same / px:.4f}% identical)")


def test_c5_every_fixture_font_on_its_own_parity(renderer):
    """C5-like: the 21 fixture fonts as 21 separate fonts of ONE FontManager (recurse over a directory without
    fonts.json, recurse.rs:127-133), so that every glyph of every file is rendered — the merge of C2 hides the code
    points a later file shares with an earlier one (14 k glyphs here against 6.5 k there).  Compared block by block
    with the oracle rendering each file alone."""
    paths = [O.FIRA] + O.noto_paths()
    m = V.FontManager(parallel=True)
    names = []
    for p in paths:
        name = os.path.basename(p)[:-4]
        names.append(name)
        m.add_font_with_name(name, [p])
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, renderer)
    files = {n: d for n, is_dir, d in w.entries() if not is_dir}
    assert len(files) == 256 * len(paths) and st.blocks == 256 * len(paths)
    px = same = glyphs = 0
    for name, p in zip(names, paths):
        fid = V.name_to_id(name)
        oset = O.FontSet(name, [p])
        pop = oset.block_population()
        glyphs += sum(pop)
        for b in range(256):
            got = files[f"{fid}/{b * 256}-{b * 256 + 255}.pbf"]
            if pop[b] == 0:
                assert got == oset.render_block(b)
                continue
            a, c = check_pbf_block(got, oset.render_block(b), (name, b))
            px += a
            same += c
    assert st.glyphs == glyphs
    assert same / px >= MIN_IDENTICAL, same / px
    print(f"C5 every fixture font alone: {glyphs} glyphs, {px} px, {px - same} differ by 1 ({100 * same / px:.4f}% identical)")


def test_single_thread_and_sharded_runs_are_identical(renderer):
    """--single-thread (recurse.rs:52-53) and the multi-GPU sharding give byte-identical PBFs: the
    kernel is deterministic (min and integer adds only)."""
    full, _, m = _render_all("Fira Sans - Regular", [O.FIRA], renderer)
    w = V.Writer.new_memory()
    m.render_glyphs(w, renderer, threads=1)
    assert {n: d for n, is_dir, d in w.entries() if not is_dir} == full
    got = {}
    for s in range(4):
        w = V.Writer.new_memory()
        m.render_glyphs(w, renderer, shard=s, n_shards=4)
        got.update({n: d for n, is_dir, d in w.entries() if not is_dir})
    assert got == full


# ---- C3 / C4: synthetic fonts ---------------------------------------------------------------------------
def _font_parity(data, renderer, name, blocks):
    m = V.FontManager(parallel=True)
    m.add_font_bytes_with_name(name, data)
    fid = V.name_to_id(name)
    path = f"/tmp/_synth_{fid}.ttf"
    open(path, "wb").write(data)
    oset = O.FontSet(name, [path])
    assert m.block_population(fid).tolist() == oset.block_population()
    px = same = 0
    for b in blocks:
        p, s = check_pbf_block(m.render_block(fid, b, renderer), oset.render_block(b), (name, b))
        px += p
        same += s
    os.unlink(path)
    assert px > 0
    assert same / px >= MIN_IDENTICAL, (name, same / px)
    return px, same


def test_first_calls_on_a_fresh_renderer_are_deterministic():
    """The first render_glyphs calls of a renderer — slots used for the first time, pooled buffers still growing,
    fonts being uploaded — give the same bytes as every later call.  (A device-side counter block that was zeroed on
    the legacy default stream instead of the slot's stream once lost the last tile jobs of a slot's first batch.)"""
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Noto Sans Regular", O.noto_paths())
    for _ in range(3):
        r = V.Renderer.new_precise(device=0)
        runs = []
        for _ in range(6):
            w = V.Writer.new_memory()
            m.render_glyphs(w, r)
            runs.append({n: d for n, is_dir, d in w.entries() if not is_dir})
        for k in range(1, len(runs)):
            bad = [n for n in runs[0] if runs[0][n] != runs[k][n]]
            assert not bad, (k, bad[:4])
        del r


def test_c3_dense_outlines_parity(renderer):
    """Synthetic dense outlines (K = 8..64 strokes, up to ~10 k segments per glyph, overlapping and
    counter-wound rings): segment staging over many chunks + winding numbers beyond 0/1."""
    data = synth_font.dense_font(n_glyphs=96, first_cp=0x4E00)
    px, same = _font_parity(data, renderer, "Synth Dense", [0x4E])
    print(f"C3 dense: {px} px, {100 * same / px:.4f}% identical")


def test_c3_full_dense_font_strided_sample(renderer):
    """The bench's C3 itself — all 4096 dense glyphs (0.5 k .. 9 k segments each) through FontManager.render_glyphs — with
    every 37th glyph checked against the oracle (metrics exact, bitmap within 1, >= 99.9 % identical)."""
    data = synth_font.dense_font(4096)
    name, fid = "Synth Dense", "synth_dense"
    m = V.FontManager(parallel=True)
    m.add_font_bytes_with_name(name, data)
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, renderer)
    assert st.glyphs == 4096 and st.bitmaps == 4096 and st.handed_back == 0
    glyphs = {}
    for n, is_dir, d in w.entries():
        if not is_dir and n.endswith(".pbf"):
            for g in V.decode_pbf(d)[2]:
                glyphs[g.id] = g
    assert len(glyphs) == 4096
    ofont = O.Font(data)
    px = same = 0
    for cp in range(0x4E00, 0x4E00 + 4096, 37):
        want, got = ofont.render_glyph(cp), glyphs[cp]
        assert (got.width, got.height, got.left, got.top, got.advance) == (
            want["width"], want["height"], want["left"], want["top"], want["advance"]), hex(cp)
        diff = np.abs(np.frombuffer(got.bitmap, np.uint8).astype(np.int16) - want["bitmap"].astype(np.int16))
        assert diff.max() <= 1, hex(cp)
        px += diff.size
        same += int((diff == 0).sum())
    assert px > 80000 and same / px >= MIN_IDENTICAL, same / px
    print(f"C3 full, strided sample: {px} px, {100 * same / px:.4f}% identical")


def test_c5_recurse_directory_with_fonts_json(renderer, tmp_path):
    """BASELINE.json configs[4]: a `recurse` input directory — fira/<file> + noto/fonts.json naming the 20 Noto files as
    ONE font (commands/recurse.rs:104-133: a directory with a fonts.json contributes exactly what it lists) — through
    FontManager.scan, rendered sharded 8 ways by font x GlyphBlock into a directory sink; the union of the shards is the
    unsharded run, index.json / font_families.json are written, and every block matches the oracle."""
    import json
    import shutil

    src = tmp_path / "fonts"
    (src / "fira").mkdir(parents=True)
    (src / "noto").mkdir()
    shutil.copy(O.FIRA, src / "fira" / os.path.basename(O.FIRA))
    names = []
    for p in O.noto_paths():
        shutil.copy(p, src / "noto" / os.path.basename(p))
        names.append(os.path.basename(p))
    (src / "noto" / "fonts.json").write_text(json.dumps([{"name": "Noto Sans Regular", "sources": names}]))
    (src / "noto" / "README.txt").write_text("not a font")
    m = V.FontManager(parallel=True)
    m.scan(str(src))
    assert m.font_ids() == ["fira_sans_regular", "noto_sans_regular"]
    whole = V.Writer.new_memory()
    st = m.render_glyphs(whole, renderer)
    want = {n: d for n, is_dir, d in whole.entries() if not is_dir}
    assert st.glyphs == 1686 + 6480 and len(want) == 512
    out = tmp_path / "out"
    out.mkdir()
    costs = []
    for shard in range(8):
        w = V.Writer.new_file(str(out))
        s8 = m.render_glyphs(w, renderer, shard=shard, n_shards=8)
        costs.append(s8.cost_shard)
        if shard == 0:
            m.write_index_json(w)
            m.write_families_json(w)
        w.finish()
    assert max(costs) <= 1.1 * sum(costs) / 8, costs  # LPT balance of the estimated costs
    got = {}
    for fid in m.font_ids():
        for f in os.listdir(out / fid):
            got[f"{fid}/{f}"] = (out / fid / f).read_bytes()
    assert got == want, "union of the 8 shards differs from the unsharded run"
    assert json.loads((out / "index.json").read_text()) == ["fira_sans_regular", "noto_sans_regular"]
    fam = json.loads((out / "font_families.json").read_text())
    assert {f["name"] for f in fam} == {"Fira Sans", "Noto Sans"}
    px = same = 0
    for fid, oset in (("fira_sans_regular", O.FontSet("Fira Sans - Regular", [O.FIRA])),
                      ("noto_sans_regular", O.FontSet("Noto Sans Regular", O.noto_paths()))):
        pop = oset.block_population()
        for b in range(256):
            blob = got[f"{fid}/{b * 256}-{b * 256 + 255}.pbf"]
            if pop[b] == 0:
                assert blob == oset.render_block(b)
                continue
            a, c = check_pbf_block(blob, oset.render_block(b), (fid, b))
            px += a
            same += c
    assert px == 758736 + 3295280 and same / px >= MIN_IDENTICAL
    print(f"C5 recurse dir: {px} px, {100 * same / px:.4f}% identical, shard cost max/mean {max(costs) * 8 / sum(costs):.3f}")


def test_c4_full_bmp_subset_parity(renderer):
    """Every 97th BMP code point of the C4 font: all block ranges, surrogate gap, format-4 cmap."""
    data = synth_font.full_bmp_font(stride=97)
    px, same = _font_parity(data, renderer, "Synth Full", list(range(0, 256, 5)) + [0xD7, 0xD8, 0xDF, 0xE0, 0xFF])
    print(f"C4 subset: {px} px, {100 * same / px:.4f}% identical")


@pytest.mark.parametrize("cid", [False, True])
def test_cff_font_parity(renderer, cid):
    """CFF (.otf) outlines, name-keyed and CID-keyed: Type 2 charstrings -> cubic flattening on the host
    (ring.rs:159-187, adaptive, so these glyphs take the segment-level jobs) -> SDF kernel, against the oracle's
    own CFF interpreter and f64 renderer."""
    data, cps, _ = synth_font.cff_test_font(n_glyphs=60, cid=cid)
    px, same = _font_parity(data, renderer, "Synth CFF CID" if cid else "Synth CFF", [0])
    print(f"CFF cid={cid}: {px} px, {100 * same / px:.4f}% identical")


def test_cff2_font_parity(renderer):
    """CFF 2 (variable .otf at its default instance): charstrings with blend / vsindex on the host -> cubic records ->
    subdivision in the decode kernel (kind PATH) -> SDF kernel, against the oracle's own CFF 2 interpreter and f64
    renderer."""
    data, cps, _ = synth_font.cff2_test_font(n_glyphs=40)
    px, same = _font_parity(data, renderer, "Synth CFF2", [0])
    print(f"CFF2: {px} px, {100 * same / px:.4f}% identical")


# ---- edge cases through the raw C ABI ---------------------------------------------------------------------
def test_empty_and_degenerate_batches(ctx):
    # empty batch
    out = ctx.render(np.zeros((0, 4), np.float32), np.zeros(0, JOB_DT), 0)
    assert out.size == 0
    # a glyph with zero segments renders as "infinitely far outside" = 0 everywhere
    out = ctx.render(np.zeros((0, 4), np.float32), np.array([(0, 0, 7, 9, 0)], JOB_DT), 63)
    assert not out.any()
    # zero-length segment = distance to a point (segment.rs:58-61); one-segment glyph
    segs = np.array([[5.5, 5.5, 5.5, 5.5]], np.float32)
    out = ctx.render(segs, np.array([(0, 1, 11, 11, 0)], JOB_DT), 121).reshape(11, 11)
    want = O.renderer_precise(0, 0, 11, 11, [[(5.5, 5.5), (5.5, 5.5)]]).reshape(11, 11)
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1
    assert out[5, 5] == 191  # on the point: 255 - 64


def test_bad_arguments_are_rejected(ctx):
    segs = np.zeros((4, 4), np.float32)
    for job in [(0, 5, 8, 8, 0), (0, 4, 0, 8, 0), (0, 4, 8, 8, 1), (0, 4, 70000, 8, 0)]:
        with pytest.raises(V.B200Error):
            ctx.render(segs, np.array([job], JOB_DT), 64)


def test_unaligned_and_sparse_output_offsets(ctx):
    """out_off may have any alignment (vectorised stores must handle head/tail bytes)."""
    rng = np.random.default_rng(5)
    jobs, segs, rings, off = [], [], [], 3
    for i in range(40):
        W, H = int(rng.integers(7, 40)), int(rng.integers(7, 40))
        cx, cy = W / 2 + rng.normal() * 0.3, H / 2 + rng.normal() * 0.3
        r = min(W, H) / 2 - 3
        n = int(rng.integers(3, 40))
        ang = np.linspace(0, 2 * np.pi, n + 1)
        ring = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], axis=1)
        ring[-1] = ring[0]
        s = np.concatenate([ring[:-1], ring[1:]], axis=1)
        jobs.append((len(np.concatenate(segs)) if segs else 0, n, W, H, off))
        segs.append(s)
        rings.append(ring)
        off += W * H + int(rng.integers(0, 7))
    allsegs = np.concatenate(segs).astype(np.float32)
    jobs = np.array(jobs, JOB_DT)
    out = ctx.render(allsegs, jobs, off)
    px = same = 0
    for j, ring in zip(jobs, rings):
        W, H = int(j["width"]), int(j["height"])
        got = out[int(j["out_off"]) : int(j["out_off"]) + W * H]
        ring32 = ring.astype(np.float32).astype(np.float64)  # the oracle sees what the kernel sees
        want = O.renderer_precise(0, 0, W, H, [ring32.tolist()])
        p, s = compare_bitmaps(got, want, (W, H))
        px += p
        same += s
    assert same / px >= MIN_IDENTICAL
    # bytes between bitmaps are never written
    mask = np.ones(off, bool)
    for j in jobs:
        mask[int(j["out_off"]) : int(j["out_off"]) + int(j["width"]) * int(j["height"])] = False
    assert not out[mask].any()


def test_vertex_band_long_decomposition_edge_cases(ctx):
    """The kernel takes min over segments as vertices + band interiors of short segments + clamped projection of
    long ones.  Shapes chosen to sit on its seams: segment lengths at and around the 0.5 px
    threshold, tiny segments (0.01..0.3 px) in every direction incl. exactly axis-aligned and 45 degrees, end points and
    whole segments exactly on pixel centres, repeated points (zero length), open polylines (raw segments need both
    end points as vertices), bands crossing tile-rectangle boundaries of a glyph that is cut into several CTAs."""
    rng = np.random.default_rng(20261018)
    jobs, segs, rings_per_job, off = [], [], [], 0

    def add(W, H, rings):
        nonlocal off
        s = [np.concatenate([np.asarray(r)[:-1], np.asarray(r)[1:]], axis=1) for r in rings]
        s = np.concatenate(s).astype(np.float32)
        jobs.append((sum(len(x) for x in segs), len(s), W, H, off))
        segs.append(s)
        rings_per_job.append([np.asarray(r, np.float32).astype(np.float64).tolist() for r in rings])
        off += W * H

    # 1. polygons whose edges are chains of equal steps of a given length (incl. exactly 0.5 and its neighbours)
    for step in (0.01, 0.05, 0.3, 0.4999, 0.5, 0.5001, 0.75, 3.0):
        W, H = 31, 27
        corners = np.array([[5.5, 5.5], [25.5, 6.0], [24.0, 21.5], [12.25, 19.0], [6.0, 22.0]])
        pts = []
        for a, b in zip(corners, np.roll(corners, -1, axis=0)):
            n = max(1, int(np.ceil(np.linalg.norm(b - a) / step)))
            d = (b - a) / np.linalg.norm(b - a)
            for k in range(n):
                pts.append(a + d * min(k * step, np.linalg.norm(b - a)))
        pts.append(pts[0])
        add(W, H, [pts])
    # 2. axis-aligned and diagonal staircases through pixel centres, with repeated points
    stair = [(4.5, 4.5), (4.5, 4.5), (14.5, 4.5), (14.5, 9.5), (14.5, 9.5), (19.5, 14.5), (19.5, 20.5), (9.5, 20.5),
             (4.5, 15.5), (4.5, 4.5)]
    add(25, 25, [stair])
    fine = []
    for a, b in zip(stair[:-1], stair[1:]):  # the same outline in 0.125 px steps (exact in binary)
        a, b = np.array(a), np.array(b)
        n = int(round(np.abs(b - a).max() / 0.125))
        for k in range(max(n, 1)):
            fine.append(a + (b - a) * (k / max(n, 1)))
    fine.append(fine[0])
    add(25, 25, [fine])
    # 3. random smooth blobs: thousands of tiny segments in every direction
    for _ in range(12):
        W, H = int(rng.integers(12, 60)), int(rng.integers(12, 60))
        n = int(rng.integers(200, 3000))
        t = np.linspace(0, 2 * np.pi, n + 1)
        r = (min(W, H) / 2 - 4) * (1 + 0.25 * np.sin(3 * t + rng.uniform(0, 6)) + 0.1 * np.sin(7 * t + rng.uniform(0, 6)))
        ring = np.stack([W / 2 + r * np.cos(t), H / 2 + r * np.sin(t)], axis=1)
        ring[-1] = ring[0]
        hole = np.stack([W / 2 + 2.2 * np.cos(-t[::8]), H / 2 + 2.2 * np.sin(-t[::8])], axis=1)
        hole[-1] = hole[0]
        add(W, H, [ring, hole])
    # 4. a glyph large enough to be cut into many rectangles, outline in 0.07 px steps
    W, H = 150, 110
    t = np.linspace(0, 2 * np.pi, 5001)
    ring = np.stack([75 + 60 * np.cos(t) + 6 * np.cos(9 * t), 55 + 42 * np.sin(t) + 5 * np.sin(11 * t)], axis=1)
    ring[-1] = ring[0]
    add(W, H, [ring])
    allsegs = np.concatenate(segs)
    jb = np.array(jobs, JOB_DT)
    out = ctx.render(allsegs, jb, off)
    px = same = 0
    for j, rings in zip(jb, rings_per_job):
        Wj, Hj = int(j["width"]), int(j["height"])
        got = out[int(j["out_off"]) : int(j["out_off"]) + Wj * Hj]
        want = O.renderer_precise(0, 0, Wj, Hj, rings)
        p, sm = compare_bitmaps(got, want, (Wj, Hj))
        px += p
        same += sm
    print(f"decomposition edge cases: {px} px, {100.0 * same / px:.4f}% identical")
    assert same / px >= MIN_IDENTICAL
    # 5. an OPEN polyline through the raw-segment seam: distances only (no interior), both end points count
    open_pts = np.array([[3.2, 3.7], [9.9, 4.1], [10.0, 4.1], [10.05, 4.15], [16.5, 12.25]], np.float32)
    s = np.concatenate([open_pts[:-1], open_pts[1:]], axis=1)
    got = ctx.render(s, np.array([(0, len(s), 20, 16, 0)], JOB_DT), 320).reshape(16, 20).astype(int)
    yy, xx = np.mgrid[0:16, 0:20]
    pxc, pyc = xx + 0.5, (15 - yy) + 0.5  # row 0 = top
    d2 = np.full((16, 20), np.inf)
    for x0, y0, x1, y1 in s.astype(np.float64):
        dx, dy = x1 - x0, y1 - y0
        tt = np.clip(((pxc - x0) * dx + (pyc - y0) * dy) / (dx * dx + dy * dy), 0, 1)
        d2 = np.minimum(d2, (pxc - x0 - tt * dx) ** 2 + (pyc - y0 - tt * dy) ** 2)
    want = np.rint(np.clip(191 - 32 * np.sqrt(d2), 0, 255)).astype(int)  # winding 0 everywhere: outside
    # (an open polyline has no inside; crossings of its segments still toggle the reference's winding sum on some
    # rows, so compare only where no segment is crossed to the left: rows above and below the polyline)
    rows_clear = [r for r in range(16) if (15 - r) + 0.5 > 12.25 or (15 - r) + 0.5 < 3.7]
    assert np.abs(got[rows_clear] - want[rows_clear]).max() <= 1


def test_large_glyph_is_tiled_over_several_ctas(ctx):
    """A 300x200 px glyph (one CTA covers at most 128 4x4 tiles) incl. a counter-wound hole and a
    self-overlapping second ring: winding prefix across tile-rectangle boundaries, column strips."""
    W, H = 300, 200
    def circle(cx, cy, r, n, rev=False):
        a = np.linspace(0, 2 * np.pi, n + 1)
        if rev:
            a = a[::-1]
        p = np.stack([cx + r * np.cos(a), cy + r * np.sin(a)], axis=1)
        p[-1] = p[0]
        return p.astype(np.float32).astype(np.float64)
    rings = [circle(150.3, 100.2, 90.0, 700), circle(150.3, 100.2, 60.0, 500, rev=True), circle(200.1, 120.4, 70.5, 300)]
    segs = np.concatenate([np.concatenate([r[:-1], r[1:]], axis=1) for r in rings]).astype(np.float32)
    jobs = np.array([(0, len(segs), W, H, 5)], JOB_DT)
    tiles, n_tiles, pairs = ctx.plan_tiles(jobs, len(segs), W * H + 5)
    assert n_tiles > 20 and pairs == W * H * len(segs)
    out = ctx.render(segs, jobs, W * H + 5)[5:]
    want = O.renderer_precise(0, 0, W, H, [r.tolist() for r in rings])
    px, same = compare_bitmaps(out, want, "large")
    assert same / px >= MIN_IDENTICAL
    assert {0, 255} <= set(np.unique(out).tolist())


def _glyph_bitmaps(renderer, batch):
    """Per-glyph bitmaps of a rendered batch (a glyph-level batch keeps its bitmaps in slots sized on the host: the
    bytes between them are undefined, so whole buffers are not comparable)."""
    renderer.finalize_batch(batch)
    return [None if (bm := batch.bitmap_of(i)) is None else bm.copy() for i in range(len(batch))]


def _same_bitmaps(a, b):
    return len(a) == len(b) and all((x is None and y is None) or (x is not None and y is not None and np.array_equal(x, y))
                                    for x, y in zip(a, b))


def test_many_batches_in_flight_from_threads(renderer):
    """submit/wait from several host threads on one context (reference: rayon workers, manager.rs:117-118)."""
    import threading

    font = V.FontFileEntry(path=O.FIRA)
    cps = font.codepoints().tolist()
    ref_batch = renderer.new_batch()
    for cp in cps[:300]:
        ref_batch.add_glyph(font, cp)
    renderer.render_batch(ref_batch)
    want = _glyph_bitmaps(renderer, ref_batch)
    errors = []

    def work():
        try:
            for _ in range(5):
                b = renderer.new_batch()
                for cp in cps[:300]:
                    b.add_glyph(font, cp)
                t = renderer.submit_batch(b)
                renderer.wait_batch(t)
                if not _same_bitmaps(_glyph_bitmaps(renderer, b), want):
                    errors.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work) for _ in range(6)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


def test_prepared_submission_and_polling(renderer):
    """b200sdf_submit_planned (tiles planned beforehand) + b200sdf_poll give the bitmaps of submit + wait."""
    import time

    font = V.FontFileEntry(path=O.FIRA)
    cps = font.codepoints().tolist()[:400]
    a, b = renderer.new_batch(), renderer.new_batch()
    for cp in cps:
        a.add_glyph(font, cp)
        b.add_glyph(font, cp)
    renderer.render_batch(a)
    renderer.prepare_batch(b)
    t = renderer.submit_batch(b)
    deadline = time.time() + 30
    while not renderer.poll_batch(t):
        assert time.time() < deadline
    assert _same_bitmaps(_glyph_bitmaps(renderer, a), _glyph_bitmaps(renderer, b))
    with pytest.raises(V.B200Error):
        renderer.poll_batch(t)  # the ticket was consumed
    with pytest.raises(V.B200Error):
        renderer.wait_batch(t)


def test_device_resident_path_matches_host_path(ctx):
    """b200sdf_render_device over torch-owned HBM buffers on a torch stream = b200sdf_render."""
    import torch

    font = V.FontFileEntry(path=O.FIRA)
    r = V.Renderer.new_dummy()  # only used to build the batch (host side)
    r.set_flatten("host")
    batch = r.new_batch()
    for cp in font.codepoints().tolist()[:500]:
        batch.add_glyph(font, cp)
    segs, jobs = batch.segments().copy(), batch.glyph_jobs()
    out_bytes = int(jobs["out_off"][-1] + jobs["width"][-1] * jobs["height"][-1])
    want = ctx.render(segs, jobs, out_bytes)
    tiles, n_tiles, pairs = ctx.plan_tiles(jobs, len(segs), out_bytes)
    assert pairs == batch.pairs
    dev = torch.device("cuda:0")
    d_segs = torch.from_numpy(segs).to(dev)
    d_tiles = torch.from_numpy(tiles).to(dev)
    d_out = torch.zeros(out_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx.render_device(d_segs.data_ptr(), d_tiles.data_ptr(), n_tiles, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)


# ---- outline-level path: flattening on the device ---------------------------------------------------------
def _both_batches(font, cps):
    r = V.Renderer.new_dummy()  # host side only: builds the batches
    dev_batch = r.new_batch()
    r.set_flatten("host")
    host_batch = r.new_batch()
    for cp in cps:
        assert dev_batch.add_glyph(font, cp) == host_batch.add_glyph(font, cp)
    return dev_batch, host_batch


@pytest.mark.parametrize("path", [O.FIRA] + O.noto_paths())
def test_device_flattening_is_bit_identical_to_host_flattening(ctx, path):
    """The device regenerates Ring::add_quadratic_bezier's points (ring.rs:119-144) from curve records:
    every f32 segment equals the host's literal flattening bit for bit, for every glyph of every fixture."""
    font = V.FontFileEntry(path=path)
    cps = [cp for cp in font.codepoints().tolist() if cp <= 0xFFFF]
    dev_batch, host_batch = _both_batches(font, cps)
    jobs = dev_batch.jobs()
    curve_jobs = jobs[jobs["kind"] == 0]
    got = ctx.flatten_outlines(dev_batch.curves(), curve_jobs)
    hsegs = host_batch.segments()
    hjobs = host_batch.jobs()
    assert len(jobs) == len(hjobs)
    pos = 0
    for dj, hj in zip(jobs, hjobs):
        assert (dj["width"], dj["height"], dj["x0"], dj["y0"], dj["seg_cnt"], dj["out_off"]) == (
            hj["width"], hj["height"], hj["x0"], hj["y0"], hj["seg_cnt"], hj["out_off"])
        if dj["kind"] != 0:
            continue
        n = int(dj["seg_cnt"])
        want = hsegs[int(hj["src_off"]) : int(hj["src_off"]) + n]
        assert np.array_equal(got[pos : pos + n], want), (path, int(dj["src_off"]))
        pos += n
    assert pos == len(got)


def test_outline_path_bitmaps_equal_segment_path_bitmaps(ctx):
    """Same segments in, same bytes out: b200sdf_submit_outlines == b200sdf_submit, byte for byte."""
    font = V.FontFileEntry(path=os.path.join(O.NOTO_DIR, "Noto Sans - Regular.ttf"))
    cps = [cp for cp in font.codepoints().tolist() if cp <= 0xFFFF][:1500]
    dev_batch, host_batch = _both_batches(font, cps)
    jobs = dev_batch.jobs()
    out_bytes = int(jobs["out_off"][-1] + jobs["width"][-1].astype(np.uint64) * jobs["height"][-1])
    a = ctx.render_outlines(dev_batch.curves(), dev_batch.segments(), jobs, out_bytes)
    b = ctx.render(host_batch.segments(), host_batch.glyph_jobs(), out_bytes)
    assert np.array_equal(a, b)
    assert (jobs["kind"] == 0).sum() > 1400


def test_outline_jobs_are_validated(ctx):
    curves = np.zeros(2, dtype=V.api.CURVE_DT)
    curves["depth"] = [1, 2]
    curves["seg_off"] = [0, 2]
    good = np.zeros(1, dtype=V.api.OUTLINE_JOB_DT)
    good[0] = (0, 0, 2, 6, 8, 8, 0, 0, 1.0, 0.0, 0)
    ctx.render_outlines(curves, np.zeros((0, 4), np.float32), good, 64)
    for field, value in (("seg_cnt", 5), ("src_cnt", 3), ("kind", 7), ("width", 0)):
        bad = good.copy()
        bad[field] = value
        with pytest.raises(V.B200Error):
            ctx.render_outlines(curves, np.zeros((0, 4), np.float32), bad, 64)
    curves["seg_off"] = [0, 3]
    with pytest.raises(V.B200Error):
        ctx.render_outlines(curves, np.zeros((0, 4), np.float32), good, 64)


# ---- full-size C4 (63 487 glyphs, ~90 M segments): size-independent properties + sampled oracle parity ----
def test_c4_full_font_properties(renderer):
    data = synth_font.full_bmp_font()
    name, fid = "Synth Full", "synth_full"
    m = V.FontManager(parallel=True)
    m.add_font_bytes_with_name(name, data)
    runs = []
    for kw in ({}, {"threads": 3}):
        w = V.Writer.new_memory()
        st = m.render_glyphs(w, renderer, **kw)
        runs.append({n: d for n, is_dir, d in w.entries() if not is_dir})
    assert st.glyphs == 63487 and st.blocks == 256 and st.bitmaps == 63487
    # (1) deterministic: identical bytes whatever the worker count / batching
    assert runs[0] == runs[1]
    # (2) sharded over 4 "GPUs": union of the shards == the whole job, byte for byte
    got = {}
    for s in range(4):
        w = V.Writer.new_memory()
        m.render_glyphs(w, renderer, shard=s, n_shards=4)
        got.update({n: d for n, is_dir, d in w.entries() if not is_dir})
    assert got == runs[0]
    # (3) structure: every block has exactly the byte length of the dummy renderer's block (metrics + PBF framing)
    wd = V.Writer.new_memory()
    m.render_glyphs(wd, V.Renderer.new_dummy())
    dummy = {n: d for n, is_dir, d in wd.entries() if not is_dir}
    assert {n: len(d) for n, d in runs[0].items()} == {n: len(d) for n, d in dummy.items()}
    # (4) every bitmap: outermost ring of the 3 px buffer is at least 2 px outside the outline -> value <= 191 - 64 + 1
    rng = np.random.default_rng(11)
    blocks = sorted(rng.choice(256, size=24, replace=False).tolist())
    path = "/tmp/_synth_full.ttf"
    open(path, "wb").write(data)
    oset = O.FontSet(name, [path])
    px = same = 0
    for b in blocks:
        blob = runs[0][f"{fid}/{b * 256}-{b * 256 + 255}.pbf"]
        _, _, glyphs = O.decode_pbf(blob)
        for g in glyphs:
            bm = g["bitmap"].reshape(g["height"] + 6, g["width"] + 6)
            edge = np.concatenate([bm[0], bm[-1], bm[:, 0], bm[:, -1]])
            assert edge.max() <= 128, (b, g["id"], int(edge.max()))
        # (5) oracle parity on a sample of whole blocks
        if b % 3 == 0 and not 0xD8 <= b <= 0xDF:
            p, s = check_pbf_block(blob, oset.render_block(b), ("c4", b))
            px += p
            same += s
    os.unlink(path)
    assert px > 100000 and same / px >= MIN_IDENTICAL
    print(f"C4 full: {len(runs[0])} blocks, oracle sample {px} px, {100 * same / px:.4f}% identical")


def test_submit_planned_rejects_tile_lists_that_do_not_fit(ctx):
    """b200sdf_submit_planned takes the caller's tile list: every entry is checked against the arrays it indexes before
    anything is enqueued (a wrong list would otherwise be an out-of-bounds device write)."""
    import ctypes as C

    from versatiles_glyphs_rs_b200 import _native as N
    from versatiles_glyphs_rs_b200.api import OUTLINE_JOB_DT, TILE_JOB_DT

    segs = np.array([[2, 2, 9, 2], [9, 2, 9, 9], [9, 9, 2, 9], [2, 9, 2, 2]], np.float32)
    jobs = np.zeros(1, OUTLINE_JOB_DT)
    jobs[0]["kind"], jobs[0]["src_off"], jobs[0]["src_cnt"], jobs[0]["seg_cnt"] = N.KIND_SEGMENTS, 0, 4, 4
    jobs[0]["width"], jobs[0]["height"], jobs[0]["out_off"] = 12, 12, 0
    good = np.zeros(1, TILE_JOB_DT)
    good[0]["seg_off"], good[0]["seg_cnt"], good[0]["out_off"] = 0, 4, 0
    good[0]["width"], good[0]["height"], good[0]["tx0"], good[0]["ty0"], good[0]["ntx"], good[0]["nty"] = 12, 12, 0, 0, 3, 3
    good[0]["job"] = 0xFFFFFFFF
    out = np.zeros(144, np.uint8)

    def submit(tiles):
        t = C.c_uint64()
        rc = N.sdf.b200sdf_submit_planned(ctx._h, None, 0, segs.ctypes.data, len(segs), jobs.ctypes.data, 1, tiles.ctypes.data,
                                          len(tiles), out.ctypes.data, out.size, C.byref(t))
        if rc == 0:
            assert N.sdf.b200sdf_wait(ctx._h, t.value) == 0
        return rc

    assert submit(good) == 0
    want = O.renderer_precise(0, 0, 12, 12, [[(2, 2), (9, 2), (9, 9), (2, 9), (2, 2)]])
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1
    for field, value in [("seg_cnt", 5), ("seg_off", 1), ("out_off", 1), ("width", 13), ("ntx", 4), ("ty0", 1), ("nty", 0),
                         ("ntx", 65), ("job", 0), ("job", 7), ("width", 0)]:
        bad = good.copy()
        bad[0][field] = value
        assert submit(bad) != 0, field
        assert b"tile job 0" in N.sdf.b200sdf_last_error(ctx._h)
