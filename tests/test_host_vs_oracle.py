"""CPU tests (no GPU): the C++ host side of the product against the oracle.

Everything that decides a metric or feeds the kernel is produced on the host, so it can be
checked bit-for-bit here: cmap / hmtx / outline flattening, glyph frames, the f32 segment buffer,
and — with the reference's own `Renderer::new_dummy()` fake (src/render/renderer.rs:39-43) — the
complete PBF byte stream of every block, incl. the reference's golden sizes.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import versatiles_glyphs_rs_b200 as V
from versatiles_glyphs_rs_b200 import _native as N


@pytest.fixture(scope="module")
def fira():
    return V.FontFileEntry(path=O.FIRA), O.Font(O.FIRA)


def test_library_exports_every_declared_symbol():
    """Every function declared in include/*.h must be exported (and nothing is called here)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for header, lib, table in (("b200sdf.h", N.sdf, N.SDF_SYMBOLS), ("vgb200_host.h", N.host, N.HOST_SYMBOLS)):
        text = open(os.path.join(root, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared = set(re.findall(r"\b((?:b200sdf|vgb)_[a-z0-9_]+)\s*\(", text))
        assert declared, header
        for name in declared:
            assert hasattr(lib, name), f"{header}: {name} not exported"
        assert declared == set(table), (header, declared ^ set(table))
    assert N.sdf.b200sdf_abi_version() == 2


def test_no_cuda_device_fails_loudly():
    """No CPU fallback: without a GPU the precise renderer refuses to exist (the dummy one works)."""
    if V.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(V.B200Error):
        V.Renderer(dummy=False)
    with pytest.raises(V.B200Error):
        V.SdfContext()
    assert V.Renderer(dummy=True).dummy


def test_face_basics_match_reference_goldens(fira):
    f, _ = fira
    # reference src/font/metadata.rs:136-153, src/font/file_entry.rs:66-71
    assert len(f.codepoints()) == 1686
    assert f.number_of_glyphs == 2677
    assert f.units_per_em == 1000
    noto = V.FontFileEntry(path=os.path.join(O.NOTO_DIR, "Noto Sans - Regular.ttf"))
    assert len(noto.codepoints()) == 3094
    assert f.glyph_index(0xD800) is None


@pytest.mark.parametrize("path", [O.FIRA] + O.noto_paths())
def test_cmap_hmtx_outline_match_oracle(path):
    """cmap enumeration, glyph ids, advances and flattened rings (font units, f64) are identical."""
    f, o = V.FontFileEntry(path=path), O.Font(path)
    cps = f.codepoints()
    assert cps.tolist() == o.codepoints()
    step = max(1, len(cps) // 400)  # every glyph of small fonts, a stride through the large ones
    for cp in cps[::step].tolist():
        gid = f.glyph_index(cp)
        assert gid == o.glyph_index(cp)
        assert f.glyph_hor_advance(gid) == o.hor_advance(gid)
        pts, starts = f.outline_rings(gid)
        rings = o.outline_rings(gid)
        assert len(rings) == len(starts) - 1
        for i, r in enumerate(rings):
            assert np.array_equal(pts[starts[i] : starts[i + 1]], r), (path, cp)


def test_flatten_point_counts():
    """reference src/render/ring_builder.rs:197-229: quad and cubic at precision 0.01 give 17 points."""
    arr = lambda p: (C.c_double * 2)(*p)
    out = (C.c_double * 4096)()
    n = N.host.vgb_flatten_quad(arr((0, 0)), arr((10, 10)), arr((20, 0)), 0.01, out, 2048)
    assert n + 1 == 17
    got = np.array(out[: 2 * n]).reshape(-1, 2)
    assert np.array_equal(got, O.flatten_quad((0, 0), (10, 10), (20, 0)))
    n = N.host.vgb_flatten_cubic(arr((0, 0)), arr((10, 10)), arr((20, 10)), arr((30, 0)), 0.01, out, 2048)
    assert n + 1 == 17
    got = np.array(out[: 2 * n]).reshape(-1, 2)
    assert np.array_equal(got, O.flatten_cubic((0, 0), (10, 10), (20, 10), (30, 0)))
    rng = np.random.default_rng(7)
    for _ in range(200):
        p = rng.integers(-2000, 2000, size=(4, 2)).astype(float)
        n = N.host.vgb_flatten_cubic(arr(p[0]), arr(p[1]), arr(p[2]), arr(p[3]), 0.01, out, 2048)
        assert np.array_equal(np.array(out[: 2 * n]).reshape(-1, 2), O.flatten_cubic(p[0], p[1], p[2], p[3]))


def test_segment_distance_cases():
    """reference src/geometry/segment.rs:117-198 (projection / clamp cases)."""
    d = N.host.vgb_segment_sqdist
    assert d(0, 0, 10, 0, 5, 3) == 9.0
    assert d(0, 0, 10, 0, -3, 4) == 25.0  # clamps to start
    assert d(0, 0, 10, 0, 13, 4) == 25.0  # clamps to end
    assert d(2, 2, 2, 2, 5, 6) == 25.0  # zero-length segment -> distance to the point
    rng = np.random.default_rng(3)
    for _ in range(500):
        a = rng.normal(size=6) * 10
        assert d(*a) == O.segment_sqdist(a[0:2], a[2:4], a[4:6])


def test_name_to_id():
    """reference src/font/manager.rs:141-147 and its tests (:224-232)."""
    assert V.name_to_id("Fira Sans - Regular") == "fira_sans_regular"
    assert V.name_to_id("  Noto--Sans__Bold  Italic ") == "noto_sans_bold_italic"
    assert V.name_to_id("Noto Sans Regular") == "noto_sans_regular"


def test_frames_and_segments_match_oracle_fira(fira):
    """Host flattening: glyph frames bit-exact; the f32 segment buffer is the oracle's f64 segments, origin-relative."""
    f, o = fira
    r = V.Renderer(dummy=True)
    r.set_flatten("host")
    batch = r.new_batch()
    cps = f.codepoints().tolist()
    for cp in cps:
        assert batch.add_glyph(f, cp)
    assert not batch.add_glyph(f, 0xD800) and not batch.add_glyph(f, 0x110000) and not batch.add_glyph(f, 0x0378)
    assert len(batch) == len(cps)
    segs = batch.segments()
    n_bitmaps = 0
    for i, cp in enumerate(cps):
        g = batch.glyph_info(i)
        osegs, (x0, y0, W, H) = o.glyph_segments(cp)
        assert g.id == cp
        if len(osegs) == 0:
            assert not g.has_bitmap
            continue
        n_bitmaps += 1
        assert (g.x0, g.y0, g.bm_width, g.bm_height) == (x0, y0, W, H), cp
        assert g.kind == N.KIND_SEGMENTS and g.seg_cnt == g.src_cnt == len(osegs)
        want = (osegs - np.array([x0, y0, x0, y0], dtype=np.float64)).astype(np.float32)
        assert np.array_equal(segs[g.src_off : g.src_off + g.seg_cnt], want), cp
    assert n_bitmaps == 1679
    assert batch.pairs == sum(
        int(batch.glyph_info(i).bm_width) * batch.glyph_info(i).bm_height * batch.glyph_info(i).seg_cnt for i in range(len(cps))
    )
    assert len(batch.glyph_jobs()) == n_bitmaps and len(batch.curves()) == 0


@pytest.mark.parametrize("path", [O.FIRA] + O.noto_paths())
def test_outline_records_give_exact_frames_and_counts(path):
    """Device-flatten mode: the frame comes from the analytic bounding box of the curve records and the
    segment count from the root flatness test — both must equal the literal flattening for EVERY glyph."""
    f, o = V.FontFileEntry(path=path), O.Font(path)
    r = V.Renderer(dummy=True)
    batch = r.new_batch()  # default: flatten on the device
    cps = [cp for cp in f.codepoints().tolist() if cp <= 0xFFFF]
    for cp in cps:
        assert batch.add_glyph(f, cp)
    curves = batch.curves()
    n_curve_glyphs = 0
    for i, cp in enumerate(cps):
        g = batch.glyph_info(i)
        want = o.render_glyph(cp, O.MODE_DUMMY)
        assert (g.id, g.advance) == (cp, want["advance"])
        if want["bitmap"] is None:
            assert not g.has_bitmap
            continue
        assert (g.width, g.height, g.left, g.top) == (want["width"], want["height"], want["left"], want["top"]), hex(cp)
        assert g.seg_cnt == want["n_segments"], hex(cp)
        if g.kind == N.KIND_CURVES:
            n_curve_glyphs += 1
            rec = curves[g.src_off : g.src_off + g.src_cnt]
            assert rec["seg_off"][0] == 0 and np.array_equal(np.cumsum(1 << rec["depth"])[:-1], rec["seg_off"][1:])
            assert int((1 << rec["depth"].astype(np.int64)).sum()) == g.seg_cnt
    # only glyphs with scaled composite components cannot be represented exactly (6 in all fixtures)
    assert batch.fallback_glyphs <= 6
    assert n_curve_glyphs > 0


def test_render_glyph_metrics_goldens_dummy(fira):
    """reference src/render/renderer.rs:176-287 metrics (bitmaps need the GPU: tests/test_gpu_parity.py)."""
    f, _ = fira
    r = V.Renderer.new_dummy()
    g = r.render_glyph(f, 0x20)
    assert (g.width, g.height, g.left, g.top, g.advance, g.bitmap) == (0, 0, 0, 0, 6, None)
    g = r.render_glyph(f, 0x41)
    assert (g.width, g.height, g.left, g.top, g.advance) == (14, 17, 0, -7, 13)
    assert len(g.bitmap) == (14 + 6) * (17 + 6) and not any(g.bitmap)
    g = r.render_glyph(f, 0xE6)
    assert (g.width, g.height, g.left, g.top, g.advance) == (19, 14, 0, -11, 19)
    g = r.render_glyph(f, 0x60)
    assert (g.width, g.height, g.left, g.top, g.advance) == (7, 5, 0, -4, 7)
    assert r.render_glyph(f, 0xD800) is None and r.render_glyph(f, 0x0378) is None


# reference src/commands/recurse.rs:341-367 — golden per-block PBF sizes of Fira Sans (dummy == precise sizes)
FIRA_PBF_SIZES = {
    0: 80022, 1024: 118037, 11264: 3579, 1280: 26296, 256: 130750, 3584: 592, 42752: 5761, 43776: 487,
    512: 92634, 64256: 1032, 65024: 50, 7424: 7260, 768: 63760, 7680: 87078, 7936: 124520, 8192: 20301,
    8448: 17395, 8704: 6511, 8960: 4375, 9472: 853,
}


def test_fira_dummy_pipeline_matches_oracle_bytes_and_golden_sizes():
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    assert m.font_ids() == ["fira_sans_regular"]
    pop = m.block_population("fira_sans_regular")
    oset = O.FontSet("Fira Sans - Regular", [O.FIRA])
    assert pop.tolist() == oset.block_population()
    # reference src/font/wrapper.rs:197-220 (first entries)
    assert pop[:3].tolist() == [192, 256, 219]
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, V.Renderer.new_dummy())
    entries = w.entries()
    assert entries[0] == ("fira_sans_regular/", True, b"")
    files = {n: d for n, is_dir, d in entries if not is_dir}
    assert len(files) == 256 and st.blocks == 256 and st.glyphs == 1686 and st.bitmaps == 1679
    assert st.segments == 600952 and st.pixels == 758736
    for start, size in FIRA_PBF_SIZES.items():
        assert len(files[f"fira_sans_regular/{start}-{start + 255}.pbf"]) == size
    nonempty = [n for n, d in files.items() if len(d) > 40]
    assert len(nonempty) == len(FIRA_PBF_SIZES)
    for b in range(256):
        assert files[f"fira_sans_regular/{b * 256}-{b * 256 + 255}.pbf"] == oset.render_block(b, O.MODE_DUMMY), b
    assert st.pbf_bytes == sum(len(d) for d in files.values())


def test_noto_merge_dummy_pipeline_matches_oracle():
    """C2: 20 Noto Sans files merged into one glyph set, first file wins (glyph_block.rs:34-36)."""
    paths = O.noto_paths()
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Noto Sans Regular", paths)
    oset = O.FontSet("Noto Sans Regular", paths)
    assert m.block_population("noto_sans_regular").tolist() == oset.block_population()
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, V.Renderer.new_dummy(), threads=4)
    assert (st.glyphs, st.bitmaps, st.segments, st.pixels) == (6480, 6445, 3956999, 3295280)
    files = {n: d for n, is_dir, d in w.entries() if not is_dir}
    for b in range(256):
        assert files[f"noto_sans_regular/{b * 256}-{b * 256 + 255}.pbf"] == oset.render_block(b, O.MODE_DUMMY), b
    # single block API = GlyphBlock::render
    assert m.render_block("noto_sans_regular", 9, V.Renderer.new_dummy()) == oset.render_block(9, O.MODE_DUMMY)


def test_sharded_render_covers_every_task_once():
    """font x block sharding (multi-GPU): shards are disjoint and their union is the full job."""
    m = V.FontManager(parallel=False)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    m.add_font_with_name("Noto Sans Regular", O.noto_paths()[:2])
    full = V.Writer.new_memory()
    r = V.Renderer.new_dummy()
    m.render_glyphs(full, r)
    want = {n: d for n, is_dir, d in full.entries() if not is_dir}
    assert len(want) == 512
    got = {}
    for s in range(3):
        w = V.Writer.new_memory()
        m.render_glyphs(w, r, shard=s, n_shards=3)
        part = {n: d for n, is_dir, d in w.entries() if not is_dir}
        assert not set(part) & set(got)
        got.update(part)
    assert got == want


def test_pbf_roundtrip_and_index_json():
    m = V.FontManager(parallel=False)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    data = m.render_block("fira_sans_regular", 0, V.Renderer.new_dummy())
    name, rng, glyphs = V.decode_pbf(data)
    oname, orng, oglyphs = O.decode_pbf(data)
    assert (name, rng) == (oname, orng) == ("fira_sans_regular", "0-255")
    assert len(glyphs) == len(oglyphs) == 192
    for a, b in zip(glyphs, oglyphs):
        assert (a.id, a.width, a.height, a.left, a.top, a.advance) == (
            b["id"], b["width"], b["height"], b["left"], b["top"], b["advance"])
        assert (a.bitmap is None) == (b["bitmap"] is None)
    w = V.Writer.new_memory()
    m.write_index_json(w)
    assert w.entries() == [("index.json", False, b'[\n  "fira_sans_regular"\n]')]


def test_file_writer(tmp_path):
    m = V.FontManager(parallel=False)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    out = tmp_path / "glyphs"
    out.mkdir()
    m.render_glyphs(V.Writer.new_file(str(out)), V.Renderer.new_dummy())
    names = sorted(os.listdir(out / "fira_sans_regular"))
    assert len(names) == 256
    assert os.path.getsize(out / "fira_sans_regular" / "0-255.pbf") == 80022  # recurse.rs:345


def test_bad_font_errors():
    with pytest.raises(V.B200Error, match="Could not parse font data"):
        V.FontFileEntry(data=b"not a font")
    m = V.FontManager()
    with pytest.raises(V.B200Error):
        m.add_font_with_name("x", ["/nonexistent/font.ttf"])


def test_pipeline_with_asynchronous_completion_is_byte_identical():
    """The submitter / worker queues of FontManager::render_glyphs under out-of-order, delayed completion
    (VGB_FAKE_LATENCY: a dummy batch finishes only 400 us after its submission): split blocks are
    reassembled byte-identically, for many worker counts and for the inline single-thread pump."""
    import subprocess
    import sys

    code = r"""
import hashlib, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_lib as O, versatiles_glyphs_rs_b200 as V
m = V.FontManager(parallel=True)
m.add_font_with_name("Noto Sans Regular", O.noto_paths())
m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
r = V.Renderer.new_dummy()
import time
for threads in (1, 2, 3, 8, 0, 5, 1) * 3:
    w = V.Writer.new_memory()
    t0 = time.perf_counter()
    st = m.render_glyphs(w, r, threads=threads)
    # the workers' condition waits have a safety-net timeout (VGB_WAIT_MS, 3 s here): a lost wake-up would show
    # up as a call that takes that long instead of milliseconds
    assert time.perf_counter() - t0 < 2.0, (threads, time.perf_counter() - t0)
    files = sorted((n, d) for n, is_dir, d in w.entries() if not is_dir)
    h = hashlib.sha256()
    for n, d in files:
        h.update(n.encode()); h.update(d)
    print(threads, len(files), st.glyphs, h.hexdigest())
""" % (O.ROOT, os.path.join(O.ROOT, "tests"))
    outs = []
    for fake in (False, True):
        env = dict(os.environ)
        env.pop("VGB_FAKE_LATENCY", None)
        env["VGB_WAIT_MS"] = "3000"
        if fake:
            env["VGB_FAKE_LATENCY"] = "400"  # microseconds per batch: long enough for idle workers to go to sleep
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append([ln.split() for ln in res.stdout.strip().splitlines()])
    lines = outs[0] + outs[1]
    assert len(lines) == 42
    assert {ln[1] for ln in lines} == {"512"} and {ln[2] for ln in lines} == {str(6480 + 1686)}
    assert len({ln[3] for ln in lines}) == 1, lines  # every run produced exactly the same files


@pytest.mark.parametrize("fmt", [4, 6, 10, 12, 13])
def test_cmap_formats_known_answers_and_oracle(fmt):
    """cmap subtable formats beyond the fixtures' 4 / 12: a synthetic font per format, checked
    against the mapping it was built with (known answer) and host against oracle (code point set, glyph ids,
    full dummy pipeline bytes).  Formats 6 / 10 return the stored id as is (0 included), 13 maps whole
    ranges to one glyph, as ttf-parser 0.25.1 does."""
    import synth_font

    cps = [0x41, 0x42, 0x43, 0x50, 0x51, 0x60] + list(range(0x4E00, 0x4E10))
    m2o = [(0x41, 0x5A, 2), (0x100, 0x10F, 5), (0x4E00, 0x4E7F, 7)]
    data = synth_font.build_font(cps, lambda cp: 3, seed=7, family="Cmap Test", cmap_format=fmt, many_to_one=m2o)
    f, o = V.FontFileEntry(data=data), O.Font(data)
    if fmt == 13:
        expect = {cp: g for s0, e0, g in m2o for cp in range(s0, e0 + 1)}
    elif fmt in (6, 10):
        expect = {cp: 0 for cp in range(cps[0], cps[-1] + 1)}  # holes: glyph 0 (.notdef), still "mapped"
        expect.update({cp: i + 1 for i, cp in enumerate(cps)})
    else:
        expect = {cp: i + 1 for i, cp in enumerate(cps)}
    for cp in list(expect) + [0x20, 0x7F, 0x5B, 0x4DFF, 0x4E80, 0x1F600]:
        want = expect.get(cp)
        assert f.glyph_index(cp) == want, (fmt, hex(cp))
        assert o.glyph_index(cp) == want, (fmt, hex(cp))
    assert f.codepoints().tolist() == sorted(expect) == list(o.codepoints())
    # the whole host path on this font equals the oracle's
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name("Cmap Test", data)
    path = f"/tmp/_cmap_{fmt}.ttf"
    open(path, "wb").write(data)
    try:
        oset = O.FontSet("Cmap Test", [path])
        assert m.block_population("cmap_test").tolist() == oset.block_population()
        for b in (0, 1, 0x4E):
            assert m.render_block("cmap_test", b, V.Renderer.new_dummy()) == oset.render_block(b, O.MODE_DUMMY), (fmt, b)
    finally:
        os.unlink(path)


def _expected_rings(contours):
    """RingBuilder (ring_builder.rs:33-117) over absolute path commands: cubics flattened by the pinned
    add_cubic_bezier restatement, rings closed, rings with < 3 points (< 4 closed) dropped."""
    rings = []
    for ct in contours:
        pts = []
        for c in ct:
            if c[0] == "M":
                pts = [(float(c[1]), float(c[2]))]
            elif c[0] == "L":
                pts.append((float(c[1]), float(c[2])))
            else:
                pts += [tuple(p) for p in O.flatten_cubic(pts[-1], c[1:3], c[3:5], c[5:7]).tolist()]
        if len(pts) < 3:
            continue
        if pts[0] != pts[-1]:
            pts.append(pts[0])
        if len(pts) >= 4:
            rings.append(np.array(pts))
    return rings


@pytest.mark.parametrize("cid,fdsel,charset", [(False, 3, None), (True, 3, None), (True, 0, None), (False, 3, "iso"),
                                               (False, 3, 0), (False, 3, 1), (False, 3, 2)])
def test_cff_charstrings_known_answers_and_oracle(cid, fdsel, charset):
    """CFF 1 outlines: a synthetic .otf whose Type 2 charstrings are generated from absolute path
    commands (every path operator, every number encoding, hint operators, width prefixes, local / global / nested
    subroutines, `seac` composites; name-keyed and CID-keyed with FDSelect formats 0 and 3, charset formats 0 / 1 / 2
    and the predefined one).  Host and oracle interpreters must both
    reproduce the commands they were generated from, and the whole dummy pipeline must agree byte for byte."""
    import synth_font

    if charset is None:
        data, cps, expected = synth_font.cff_test_font(cid=cid, fd_select_format=fdsel)
    else:  # `seac` composites through the predefined ISOAdobe charset or a charset table of format 0 / 1 / 2
        data, cps, expected = synth_font.cff_test_font(n_glyphs=70, seac=True, charset_format=None if charset == "iso" else charset)
    f, o = V.FontFileEntry(data=data), O.Font(data)
    assert f.codepoints().tolist() == cps == list(o.codepoints())
    n_rings = 0
    for cp, contours in zip(cps, expected):
        gid = f.glyph_index(cp)
        assert gid == o.glyph_index(cp) == cp - cps[0] + 1
        want = _expected_rings(contours)
        pts, starts = f.outline_rings(gid)
        got_o = o.outline_rings(gid)
        assert len(want) == len(starts) - 1 == len(got_o), (hex(cp), len(want), len(starts) - 1, len(got_o))
        for i, r in enumerate(want):
            assert np.array_equal(pts[starts[i] : starts[i + 1]], r), (hex(cp), i)
            assert np.array_equal(got_o[i], r), (hex(cp), i)
        n_rings += len(want)
    assert n_rings > len(cps)
    # .notdef is `endchar` only: no outline, no error
    assert len(f.outline_rings(0)[1]) == 1 and len(o.outline_rings(0)) == 0
    # frames, metrics and PBF bytes of the whole block through the dummy renderer
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name("Synth CFF", data)
    path = f"/tmp/_cff_{int(cid)}_{fdsel}_{charset}.otf"
    open(path, "wb").write(data)
    try:
        oset = O.FontSet("Synth CFF", [path])
        assert m.block_population("synth_cff").tolist() == oset.block_population()
        assert m.render_block("synth_cff", 0, V.Renderer.new_dummy()) == oset.render_block(0, O.MODE_DUMMY)
    finally:
        os.unlink(path)


def test_cmap_format_2_known_answers_and_oracle():
    """cmap format 2 (high-byte mapping through table) under a Unicode platform id: one-byte codes through subheader 0,
    two-byte codes through the lead byte's subheader, idDelta (negative too), stored glyph 0 = unmapped; the code point
    set is what ttf-parser's `codepoints` enumerates, filtered by `glyph_index` (metadata.rs:104-118)."""
    import synth_font

    cps = list(range(0x4E00, 0x4E10))  # 16 glyphs, ids 1..16, whatever the base font's own cmap says
    base = synth_font.build_font(cps, lambda cp: 3, seed=5, family="Cmap2 Test")
    data = synth_font.with_cmap_format2(base, (0x41, [1, 2, 0, 4, 5]),
                                        [(0x4E, 0x00, 5, [1, 2, 3, 4, 0, 6, 7, 8]), (0x81, 0x40, -2, [12, 13, 14, 2])])
    expect = {0x41: 1, 0x42: 2, 0x44: 4, 0x45: 5}
    expect.update({0x4E00 + k: g + 5 for k, g in enumerate([1, 2, 3, 4, 0, 6, 7, 8]) if g})
    expect.update({0x8140 + k: g - 2 for k, g in enumerate([12, 13, 14, 2]) if g})
    f, o = V.FontFileEntry(data=data), O.Font(data)
    for cp in list(expect) + [0x20, 0x40, 0x43, 0x46, 0x4E04, 0x4E08, 0x813F, 0x8144, 0x9000, 0x1F600]:
        assert f.glyph_index(cp) == expect.get(cp), hex(cp)
        assert o.glyph_index(cp) == expect.get(cp), hex(cp)
    assert f.codepoints().tolist() == sorted(expect) == list(o.codepoints())
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name("Cmap2 Test", data)
    path = "/tmp/_cmap_2.ttf"
    open(path, "wb").write(data)
    try:
        oset = O.FontSet("Cmap2 Test", [path])
        assert m.block_population("cmap2_test").tolist() == oset.block_population()
        for b in (0, 0x4E, 0x81):
            assert m.render_block("cmap2_test", b, V.Renderer.new_dummy()) == oset.render_block(b, O.MODE_DUMMY), b
    finally:
        os.unlink(path)
    # truncations: both parsers answer nothing rather than reading past the table
    for cut in (100, 517, 530, len(data)):
        n = int.from_bytes(data[4:6], "big")
        rec = next(12 + 16 * i for i in range(n) if data[12 + 16 * i : 16 + 16 * i] == b"cmap")
        b = bytearray(data)
        b[rec + 12 : rec + 16] = min(cut, int.from_bytes(data[rec + 12 : rec + 16], "big")).to_bytes(4, "big")
        ft, ot = V.FontFileEntry(data=bytes(b)), O.Font(bytes(b))
        assert ft.codepoints().tolist() == list(ot.codepoints())
        for cp in list(expect)[:6]:
            assert ft.glyph_index(cp) == ot.glyph_index(cp)


def test_cmap_format_14_subtable_is_skipped_like_ttf_parser_does():
    """A cmap with a format 14 (Unicode variation sequences) subtable in front of the usual one: ttf-parser's
    `Face::glyph_index` asks every Unicode subtable in turn and format 14 answers None for a plain code point; its
    `codepoints` callback enumerates nothing for format 14.  So lookups and the code point set are those of the font
    without the subtable — host and oracle — and the variation selectors themselves are not code points of the font."""
    import synth_font

    cps = [0x41, 0x42, 0x43, 0x50] + list(range(0x4E00, 0x4E08))
    plain = synth_font.build_font(cps, lambda cp: 3, seed=9, family="UVS Test", cmap_format=12)
    data = synth_font.with_variation_selectors(plain, [(0xFE00, [(0x41, 2), (0x4E00, 3)]), (0xE0100, [(0x4E01, 4)])])
    f0, f, o = V.FontFileEntry(data=plain), V.FontFileEntry(data=data), O.Font(data)
    assert f.codepoints().tolist() == f0.codepoints().tolist() == sorted(cps) == list(o.codepoints())
    for cp in cps + [0x20, 0xFE00, 0xE0100, 0x4E09]:
        assert f.glyph_index(cp) == f0.glyph_index(cp) == o.glyph_index(cp), hex(cp)
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name("UVS Test", data)
    path = "/tmp/_cmap_14.ttf"
    open(path, "wb").write(data)
    try:
        oset = O.FontSet("UVS Test", [path])
        assert m.block_population("uvs_test").tolist() == oset.block_population()
        assert m.render_block("uvs_test", 0, V.Renderer.new_dummy()) == oset.render_block(0, O.MODE_DUMMY)
    finally:
        os.unlink(path)


def test_cff2_charstrings_known_answers_and_oracle():
    """CFF 2 outlines (variable .otf; ttf-parser cff2.rs at the face's default coordinates — the reference never sets
    any): 32-bit INDEX counts, no width / endchar / return, `blend` with one and several operands, `vsindex`, a region
    whose scalar is 1 at the default instance, hint operators, local / global / nested subroutines.  Host and oracle
    must both reproduce the commands the charstrings were generated from, and the dummy pipeline must agree byte for
    byte."""
    import synth_font

    data, cps, expected = synth_font.cff2_test_font()
    f, o = V.FontFileEntry(data=data), O.Font(data)
    assert f.codepoints().tolist() == cps == list(o.codepoints())
    n_rings = 0
    for cp, contours in zip(cps, expected):
        gid = f.glyph_index(cp)
        assert gid == o.glyph_index(cp) == cp - cps[0] + 1
        want = _expected_rings(contours)
        pts, starts = f.outline_rings(gid)
        got_o = o.outline_rings(gid)
        assert len(want) == len(starts) - 1 == len(got_o), (hex(cp), len(want), len(starts) - 1, len(got_o))
        for i, r in enumerate(want):
            assert np.array_equal(pts[starts[i] : starts[i + 1]], r), (hex(cp), i)
            assert np.array_equal(got_o[i], r), (hex(cp), i)
        n_rings += len(want)
    assert n_rings > len(cps)
    assert len(f.outline_rings(0)[1]) == 1 and len(o.outline_rings(0)) == 0  # .notdef: an empty charstring
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name("Synth CFF2", data)
    path = "/tmp/_cff2_synth.otf"
    open(path, "wb").write(data)
    try:
        oset = O.FontSet("Synth CFF2", [path])
        assert m.block_population("synth_cff2").tolist() == oset.block_population()
        assert m.render_block("synth_cff2", 0, V.Renderer.new_dummy()) == oset.render_block(0, O.MODE_DUMMY)
    finally:
        os.unlink(path)


def test_cff2_malformed_inputs_do_not_crash():
    """Truncations and byte flips of the CFF2 and fvar tables: both parsers stay in bounds and agree on what comes out."""
    import synth_font

    data, cps, _ = synth_font.cff2_test_font(n_glyphs=10)
    n_tables = int.from_bytes(data[4:6], "big")
    rng = np.random.default_rng(4)
    variants = []
    for tag in (b"CFF2", b"fvar"):
        rec = next(12 + 16 * i for i in range(n_tables) if data[12 + 16 * i : 16 + 16 * i] == tag)
        off, length = int.from_bytes(data[rec + 8 : rec + 12], "big"), int.from_bytes(data[rec + 12 : rec + 16], "big")
        for cut in (3, 9, 40, length // 2, length - 7):
            b = bytearray(data)
            b[rec + 12 : rec + 16] = max(0, min(cut, length)).to_bytes(4, "big")
            variants.append(bytes(b))
        for _ in range(120 if tag == b"CFF2" else 20):
            b = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                b[off + int(rng.integers(0, length))] = int(rng.integers(0, 256))
            variants.append(bytes(b))
    agreed = 0
    for v in variants:
        f, o = V.FontFileEntry(data=v), O.Font(v)
        for gid in range(0, len(cps) + 1):
            pts, starts = f.outline_rings(gid)
            got_o = o.outline_rings(gid)
            assert len(starts) - 1 == len(got_o)
            for i, r in enumerate(got_o):
                assert np.array_equal(pts[starts[i] : starts[i + 1]], r, equal_nan=True)
            agreed += 1
    assert agreed > 1000


def test_cff_malformed_inputs_do_not_crash():
    """Truncations and byte flips of the CFF table: both parsers must stay in bounds and agree on what comes out
    (a rejected table = a face without outlines; a charstring error keeps the callbacks made before it)."""
    import synth_font

    data, cps, _ = synth_font.cff_test_font(n_glyphs=12, cid=True)
    n_tables = int.from_bytes(data[4:6], "big")
    rec = next(12 + 16 * i for i in range(n_tables) if data[12 + 16 * i : 16 + 16 * i] == b"CFF ")
    off, length = int.from_bytes(data[rec + 8 : rec + 12], "big"), int.from_bytes(data[rec + 12 : rec + 16], "big")
    rng = np.random.default_rng(3)
    variants = []
    for cut in (3, 10, 40, length // 2, length - 7):
        b = bytearray(data)
        b[rec + 12 : rec + 16] = cut.to_bytes(4, "big")
        variants.append(bytes(b))
    for _ in range(150):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 4))):
            b[off + int(rng.integers(0, length))] = int(rng.integers(0, 256))
        variants.append(bytes(b))
    for v in variants:
        f, o = V.FontFileEntry(data=v), O.Font(v)
        for gid in range(0, len(cps) + 1):
            pts, starts = f.outline_rings(gid)
            rings = o.outline_rings(gid)
            assert len(rings) == len(starts) - 1
            for i, r in enumerate(rings):
                assert np.array_equal(pts[starts[i] : starts[i + 1]], r, equal_nan=True)


def test_pipeline_buffers_are_allocated_in_the_first_call_only():
    """Pooled batch buffers (pinned memory on the GPU, where an allocation stalls the CUDA context): after the first
    render_glyphs call the pool holds every batch a call needs at the capacity marks, so later calls allocate nothing.
    Single-threaded and dummy renderer: batch composition is deterministic, the pool logic is the product's."""
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import versatiles_glyphs_rs_b200 as V, oracle_lib as O\n"
        "m = V.FontManager(parallel=True); m.add_font_with_name('Noto Sans Regular', O.noto_paths())\n"
        "r = V.Renderer.new_dummy()\n"
        "for i in range(6):\n"
        "    sys.stderr.write('== call %%d\\n' %% i); sys.stderr.flush()\n"
        "    m.render_glyphs(V.Writer.new_memory(), r, threads=1)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VGB_ALLOC_TRACE="1", VGB_FAKE_LATENCY="100")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    call, per_call = -1, {}
    for line in res.stderr.splitlines():
        if line.startswith("== call"):
            call = int(line.split()[2])
        elif line.startswith("[vgb alloc]") and "->" in line:
            per_call[call] = per_call.get(call, 0) + 1
    assert per_call.get(0, 0) > 0, res.stderr[-2000:]
    assert set(per_call) == {0}, per_call
