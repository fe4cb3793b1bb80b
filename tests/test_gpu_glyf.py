"""GPU parity of the glyph-level path: glyf decoding, outline recording, metrics and tile planning on the device
(csrc/glyf_kernel.cuh) against

  * the host's OutlineRecorder (csrc/host/render.cc) — curve records, segment counts and frames must be
    bit-identical for every glyph of every fixture font, and the bitmaps of the two GPU paths byte-identical;
  * the CPU oracle (oracle/vg_oracle.c, the restatement of the reference) — metrics exact, bitmaps within 1 with
    >= 99.9 % of the pixels identical (BASELINE.json north_star).

Everything goes through the C ABI (ctypes).  Reference behaviour being reproduced: ttf-parser 0.25.1's glyf walker
 -> RingBuilder (reference src/render/ring_builder.rs:67-117) -> prepare_glyph
(src/render/renderer.rs:64-91)."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O  # noqa: E402
import synth_font  # noqa: E402
import versatiles_glyphs_rs_b200 as V  # noqa: E402
from versatiles_glyphs_rs_b200 import _native as N  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    return V.Renderer.new_precise(device=0, n_slots=4)


def _batches(renderer, font, cps):
    """The same glyphs as a glyph-level batch (device decoding) and as a curve-record batch (host recorder)."""
    renderer.set_flatten("glyf")
    g = renderer.new_batch()
    renderer.set_flatten("device")
    d = renderer.new_batch()
    renderer.set_flatten("glyf")
    for cp in cps:
        a, b = g.add_glyph(font, cp), d.add_glyph(font, cp)
        assert a == b
    return g, d


def _compare_glyf_with_recorder(renderer, font, cps):
    g, d = _batches(renderer, font, cps)
    ctx = V.SdfContext.of_renderer(renderer)
    reqs, parts = g.requests(), g.parts()
    frames, jobs, curves, tiles = ctx.decode_glyphs(reqs, parts, g.curve_slots, curves=g.curves(), n_seg=len(g.segments()),
                                                    est_cost=g.est_cost)
    assert not (frames["status"] == N.GLYPH_BAD_REQUEST).any()
    handed = int((frames["status"] == N.GLYPH_NEEDS_HOST).sum())
    ok = np.flatnonzero((frames["status"] == N.GLYPH_OK) & (reqs["kind"] == N.KIND_GLYF))
    # the claim order of the tile jobs: every glyph planned, heaviest cost class first (a factor of two per class, the
    # lightest class open-ended)
    assert len(tiles) >= int((frames["status"] == N.GLYPH_OK).sum())
    cost = tiles["ntx"].astype(np.int64) * tiles["nty"] * (tiles["seg_cnt"].astype(np.int64) + 8)
    cap = max(g.est_cost // (148 * 6), 65536)  # b200sdf.cu glyph_cost_cap
    cls = np.zeros(len(cost), dtype=np.int64)
    lim = cap >> 1
    for c in range(1, 8):  # glyf_kernel.cuh tile_class
        cls[cost <= lim] = c
        lim >>= 1
    assert (np.diff(cls) >= 0).all(), "tile jobs are not claimed in descending cost-class order"
    # end to end: both GPU paths give the same glyph metrics and the same bytes
    renderer.render_batch(g)
    renderer.render_batch(d)
    assert len(g) == len(d)
    req_of = {int(o): k for k, o in enumerate(reqs["out_off"])}  # (requests are not in glyph order: heavy glyphs lead)
    dc = d.curves()
    n_px = n_rec = 0
    for i in range(len(g)):
        a, b = g.glyph_info(i), d.glyph_info(i)
        assert (a.id, a.advance, a.has_bitmap, a.x0, a.y0, a.bm_width, a.bm_height, a.width, a.height, a.left, a.top, a.seg_cnt) == (
            b.id, b.advance, b.has_bitmap, b.x0, b.y0, b.bm_width, b.bm_height, b.width, b.height, b.left, b.top, b.seg_cnt), hex(a.id)
        if not a.has_bitmap:
            continue
        assert np.array_equal(g.bitmap_of(i), d.bitmap_of(i)), hex(a.id)
        n_px += a.bm_width * a.bm_height
        # record level: what the device decoded for this glyph's request == what the host recorder wrote
        k = req_of[int(a.out_off)]
        if frames["status"][k] != N.GLYPH_OK or b.kind != N.KIND_CURVES:
            continue
        j = jobs[k]
        assert (j["width"], j["height"], j["x0"], j["y0"], j["seg_cnt"], j["src_cnt"]) == (
            b.bm_width, b.bm_height, b.x0, b.y0, b.seg_cnt, b.src_cnt), hex(a.id)
        ra = curves[j["src_off"] : j["src_off"] + j["src_cnt"]]
        rb = dc[b.src_off : b.src_off + b.src_cnt]
        assert ra.tobytes() == rb.tobytes(), f"curve records of U+{a.id:04X} differ"
        n_rec += int(j["src_cnt"])
    return {"glyphs": len(g), "requests": len(reqs), "ok": len(ok), "handed_back": handed, "records": n_rec, "pixels": n_px,
            "host_recorded": int((reqs["kind"] != N.KIND_GLYF).sum()), "batch_handed_back": g.handed_back}


@pytest.mark.parametrize("path", [O.FIRA] + O.noto_paths())
def test_device_glyf_decoding_is_bit_identical_to_host_recorder(renderer, path):
    """Every BMP glyph of every fixture font: device-decoded curve records, segment counts and frames equal the host
    recorder's bit for bit, nothing is handed back, and the bitmaps of the two paths are byte-identical."""
    font = V.FontFileEntry(path=path)
    cps = [int(c) for c in font.codepoints() if c <= 0xFFFF]
    res = _compare_glyf_with_recorder(renderer, font, cps)
    assert res["handed_back"] == 0 and res["batch_handed_back"] == 0, res
    assert res["ok"] > 0 and res["records"] > 0
    # scaled composites (4 in Noto Sans Arabic, 2 in Myanmar — SURVEY.md 8c) are recorded on the host
    assert res["host_recorded"] <= 8, res


def _contour(points):
    return [(int(x), int(y), int(on)) for x, y, on in points]


def _edge_case_font():
    """Glyphs that walk every branch of the contour rules and of the glyf encodings."""
    sq = [(100, 0, 1), (700, 0, 1), (700, 600, 1), (100, 600, 1)]
    # first point off-curve, second on
    off_first = [(400, -50, 0), (700, 300, 1), (400, 650, 0), (100, 300, 1)]
    # first two points off-curve (ring starts at their midpoint), last point off-curve too
    off_off = [(100, 0, 0), (700, 0, 0), (700, 600, 0), (100, 600, 0)]
    # all on-curve with a repeated point (zero-length segment) and a long run of equal flags (REPEAT)
    stairs = []
    for k in range(12):
        stairs += [(50 * k, 40 * k, 1), (50 * k + 50, 40 * k, 1)]
    stairs += [(650, 700, 1), (650, 700, 1), (0, 700, 1)]
    # odd coordinates: implied midpoints are half-integers
    halves = [(101, 3, 0), (703, 7, 0), (699, 611, 0), (97, 605, 0), (301, 301, 1)]
    # mixed: on, off, off, on, off, on
    mixed = [(0, 0, 1), (300, -200, 0), (600, -100, 0), (800, 200, 1), (900, 700, 0), (400, 800, 1), (0, 500, 1)]
    # large deltas (two-byte coordinates) and negative positions
    big = [(-3000, -2000, 1), (4000, -2000, 1), (4000, 3000, 0), (500, 5000, 1), (-3000, 3000, 0)]
    recs = {
        "square": [sq],
        "off_first": [off_first],
        "off_off": [off_off],
        "stairs": [stairs],
        "halves": [halves],
        "mixed": [mixed, [(p[0] + 100, p[1] + 100, p[2]) for p in sq[::-1]]],
        "big": [big],
        # degenerate contours: one point (on / off), two on-curve points (3-point ring, dropped), then a real one
        "degenerate": [[(10, 10, 1)], [(20, 20, 0)], [(0, 0, 1), (500, 500, 1)], sq],
        # a lone off/on pair: QUAD from the on point through the off point back (1 + 2^k points)
        "pair": [[(300, 900, 0), (300, 0, 1)], sq],
        "only_dropped": [[(0, 0, 1), (500, 500, 1)]],   # no ring survives -> empty glyph
        "flat": [[(0, 100, 1), (300, 100, 1), (600, 100, 1), (300, 100, 1)]],  # no extent in y: still a frame (bbox.rs:56-58)
        "point": [[(5, 5, 1), (5, 5, 1), (5, 5, 1), (5, 5, 1)]],  # no extent at all -> empty glyph
    }
    names = list(recs)
    records = [synth_font._glyph_bytes_compact([_contour(c) for c in recs[n]]) for n in names]
    # instructions in front of the flags must be skipped
    records[0] = synth_font._glyph_bytes_compact([_contour(sq)], instructions=b"\x08\x08\x88\xff\x08")
    n_simple = len(records)
    first_extra = 1 + n_simple + 6  # glyph ids of extra records start after the mapped glyphs (6 composites below)
    comp = [
        synth_font.composite_bytes([(1, 0, 0, None)]),                                   # plain reference
        synth_font.composite_bytes([(1, 100, -50, None), (2, -300, 400, None)]),         # two translated components
        synth_font.composite_bytes([(1 + n_simple + 1, 30, 40, None), (5, 2000, 0, None)]),  # nested (component is a composite)
        synth_font.composite_bytes([(1, 0, 0, 0.5), (2, 300, 0, None)]),                 # scaled component -> host recorder
        synth_font.composite_bytes([(1, 10, 10, (1.0, 0.0, 0.0, 1.0))]),                 # explicit identity matrix: still a translation
        synth_font.composite_bytes([(3, 0, 0, (0.0, 1.0, -1.0, 0.0))]),                  # rotated -> host recorder
    ]
    assert first_extra == 1 + n_simple + len(comp)
    cps = list(range(0x100, 0x100 + n_simple + len(comp)))
    return synth_font.build_font(cps, None, records=records + comp, family="Synth Edge"), cps, names


def test_glyf_decoder_edge_cases(renderer):
    """Contours starting off-curve / with two off-curve points, implied half-integer midpoints, degenerate contours,
    REPEAT flags, one- and two-byte deltas, instructions, translated / nested / scaled / rotated composites: device
    decoding == host recorder bit for bit, and both match the oracle."""
    blob, cps, names = _edge_case_font()
    font, ofont = V.FontFileEntry(data=blob), O.Font(blob)
    res = _compare_glyf_with_recorder(renderer, font, cps)
    assert res["handed_back"] == 0, res
    assert res["host_recorded"] == 2, res  # the scaled and the rotated composite
    renderer.set_flatten("glyf")
    batch = renderer.new_batch()
    for cp in cps:
        assert batch.add_glyph(font, cp)
    renderer.render_batch(batch)
    total = same = 0
    seen_empty = 0
    for i, cp in enumerate(cps):
        want = ofont.render_glyph(cp)
        got = batch.glyph_info(i)
        assert (got.width, got.height, got.left, got.top, got.advance) == (
            want["width"], want["height"], want["left"], want["top"], want["advance"]), hex(cp)
        if want["bitmap"] is None:
            assert not got.has_bitmap, hex(cp)
            seen_empty += 1
            continue
        diff = np.abs(batch.bitmap_of(i).reshape(-1).astype(np.int16) - want["bitmap"].astype(np.int16))
        assert diff.max() <= 1, (hex(cp), int(diff.max()))
        total += diff.size
        same += int((diff == 0).sum())
    assert seen_empty == 2  # "only_dropped" and "point"
    assert same / total >= 0.999, same / total


def test_glyf_decoder_hands_back_what_it_cannot_take(renderer):
    """Truncated records and coordinates beyond the exactly representable range come back as NEEDS_HOST and are then
    recorded by the host: the result still equals the oracle's."""
    sq = [(100, 0, 1), (700, 0, 1), (700, 600, 1), (100, 600, 1)]
    good = synth_font._glyph_bytes_compact([_contour(sq)])
    far = synth_font.composite_bytes([(1, 32700, 0, None)])  # 32700 + 100 > 2^15: outside the recorder's exact range
    # header bounding box far too small: the real frame does not fit the slot reserved from it
    liar = bytearray(synth_font._glyph_bytes([_contour([(0, 0, 1), (3000, 0, 1), (3000, 3000, 1), (0, 3000, 1)])]))
    liar[2:10] = struct.pack(">hhhh", 0, 0, 10, 10)
    cps = [0x41, 0x42, 0x43]
    blob = synth_font.build_font(cps, None, records=[good, far, bytes(liar)], family="Synth Back")
    font, ofont = V.FontFileEntry(data=blob), O.Font(blob)
    renderer.set_flatten("glyf")
    batch = renderer.new_batch()
    for cp in cps:
        assert batch.add_glyph(font, cp)
    renderer.render_batch(batch)
    assert batch.handed_back == 2
    for i, cp in enumerate(cps):
        want = ofont.render_glyph(cp)
        got = batch.glyph_info(i)
        assert (got.width, got.height, got.left, got.top, got.advance) == (
            want["width"], want["height"], want["left"], want["top"], want["advance"]), hex(cp)
        diff = np.abs(batch.bitmap_of(i).reshape(-1).astype(np.int16) - want["bitmap"].astype(np.int16))
        assert diff.max() <= 1


def test_glyph_requests_are_validated(renderer):
    """Out-of-range requests are answered with BAD_REQUEST, never with an out-of-bounds access."""
    font = V.FontFileEntry(path=O.FIRA)
    renderer.set_flatten("glyf")
    batch = renderer.new_batch()
    assert batch.add_glyph(font, 0x41)
    ctx = V.SdfContext.of_renderer(renderer)
    reqs, parts = batch.requests(), batch.parts()
    for field, value in (("src_off", 7), ("curve_off", 1 << 30)):
        bad = reqs.copy()
        bad[field][0] = value
        frames, _, _, _ = ctx.decode_glyphs(bad, parts, batch.curve_slots)
        assert frames["status"][0] == N.GLYPH_BAD_REQUEST
    badp = parts.copy()
    badp["glyf_off"][0] = 0x7FFFFFF0
    frames, _, _, _ = ctx.decode_glyphs(reqs, badp, batch.curve_slots)
    assert frames["status"][0] == N.GLYPH_BAD_REQUEST
    badp = parts.copy()
    badp["font"][0] = 4000
    frames, _, _, _ = ctx.decode_glyphs(reqs, badp, batch.curve_slots)
    assert frames["status"][0] == N.GLYPH_BAD_REQUEST
    # a slot too small for the frame: handed back, not overrun
    small = reqs.copy()
    small["out_cap"][0] = 16
    frames, _, _, _ = ctx.decode_glyphs(small, parts, batch.curve_slots)
    assert frames["status"][0] == N.GLYPH_NEEDS_HOST
    # too short a tile list fails the batch at wait
    with pytest.raises(V.B200Error):
        big = renderer.new_batch()
        for cp in range(0x21, 0x7F):
            big.add_glyph(font, cp)
        r = big.requests()
        ctx.render_glyphs(r, big.parts(), big.curve_slots, 2, int((r["out_off"] + r["out_cap"]).max()))


def test_glyph_level_submission_from_pageable_memory(renderer):
    """b200sdf_submit_glyphs over plain (not pinned) host arrays takes the staged-copy path: same answers."""
    font = V.FontFileEntry(path=O.FIRA)
    cps = [int(c) for c in font.codepoints() if c <= 0xFFFF][:300]
    renderer.set_flatten("glyf")
    batch = renderer.new_batch()
    for cp in cps:
        batch.add_glyph(font, cp)
    reqs, parts = batch.requests(), batch.parts()
    out_bytes = int((reqs["out_off"] + reqs["out_cap"]).max())
    ctx = V.SdfContext.of_renderer(renderer)
    frames, out = ctx.render_glyphs(reqs, parts, batch.curve_slots, batch.tile_cap, out_bytes, est_cost=batch.est_cost)
    renderer.render_batch(batch)
    req_of = {int(o): k for k, o in enumerate(reqs["out_off"])}
    seen = 0
    for i in range(len(batch)):
        g = batch.glyph_info(i)
        if not g.has_bitmap:
            continue
        k = req_of[int(g.out_off)]
        f = frames[k]
        assert f["status"] == N.GLYPH_OK
        assert (f["x0"], f["y0"], f["width"], f["height"]) == (g.x0, g.y0, g.bm_width, g.bm_height)
        n = g.bm_width * g.bm_height
        assert np.array_equal(out[int(reqs["out_off"][k]) : int(reqs["out_off"][k]) + n], batch.bitmap_of(i).reshape(-1))
        seen += 1
    assert seen == int((frames["status"] == N.GLYPH_OK).sum()) > 250


def test_several_batches_in_one_submission(renderer):
    """b200sdf_submit_glyph_batches: 5 batches of different fonts and sizes (one of them empty) share one decode launch
    and one SDF launch; every batch gets exactly the frames and bitmaps it gets when it is submitted alone."""
    renderer.set_flatten("glyf")
    ctx = V.SdfContext.of_renderer(renderer)
    fonts = [V.FontFileEntry(path=p) for p in (O.FIRA, O.noto_paths()[0])]
    group, alone = [], []
    est = 0
    for k, n in enumerate((300, 7, 0, 512, 129)):
        font = fonts[k % 2]
        cps = [int(c) for c in font.codepoints() if c <= 0xFFFF][40 * k : 40 * k + n]
        batch = renderer.new_batch()
        for cp in cps:
            batch.add_glyph(font, cp)
        reqs, parts = batch.requests(), batch.parts()
        out_bytes = int((reqs["out_off"] + reqs["out_cap"]).max()) if len(reqs) else 0
        group.append((reqs, parts, batch.curve_slots, batch.tile_cap, out_bytes))
        est = max(est, batch.est_cost)
        alone.append(ctx.render_glyphs(reqs, parts, batch.curve_slots, batch.tile_cap, out_bytes, est_cost=batch.est_cost))
    together = ctx.render_glyph_batches(group, est_cost=est)
    checked = 0
    for (reqs, *_), (fa, oa), (fb, ob) in zip(group, alone, together):
        assert np.array_equal(fa, fb)
        for k in range(len(reqs)):
            if fa[k]["status"] != N.GLYPH_OK:
                continue
            o, n = int(reqs["out_off"][k]), int(fa[k]["width"]) * int(fa[k]["height"])
            assert np.array_equal(oa[o : o + n], ob[o : o + n])
            checked += 1
    assert checked > 800
    with pytest.raises(V.B200Error):  # 17 batches: over B200SDF_MAX_BATCHES
        ctx.render_glyph_batches([group[1]] * 17)


@pytest.mark.parametrize("cid", [False, True])
def test_cubic_outlines_are_flattened_on_the_device(renderer, cid, tmp_path):
    """CFF glyphs (cubic curves, Ring::add_cubic_bezier ring.rs:159-187) travel as kind PATH: the host counts the leaves
    of the adaptive subdivision, the decode kernel repeats it literally and writes the segments.  Frames, metrics and
    every bitmap byte equal the host-flattened rendering of the same glyphs (same SDF kernel, uploaded segments)."""
    data, cps, _ = synth_font.cff_test_font(n_glyphs=60, cid=cid)
    path = tmp_path / "synth.otf"
    path.write_bytes(data)
    font = V.FontFileEntry(path=str(path))
    got = {}
    for mode in ("glyf", "host"):
        renderer.set_flatten(mode)
        batch = renderer.new_batch()
        for cp in cps:
            batch.add_glyph(font, cp)
        if mode == "glyf":
            reqs = batch.requests()
            assert batch.path_glyphs > 40 and int((reqs["kind"] == N.KIND_PATH).sum()) == batch.path_glyphs
            assert not (reqs["kind"] == N.KIND_SEGMENTS).any()  # nothing was flattened on the host
        else:
            assert batch.path_glyphs == 0
        renderer.render_batch(batch)
        got[mode] = [(batch.glyph_info(i), batch.bitmap_of(i).copy() if batch.glyph_info(i).has_bitmap else None)
                     for i in range(len(batch))]
    renderer.set_flatten("glyf")
    assert len(got["glyf"]) == len(got["host"]) == len(cps)
    px = 0
    for (ga, ba), (gb, bb) in zip(got["glyf"], got["host"]):
        assert (ga.id, ga.advance, ga.has_bitmap) == (gb.id, gb.advance, gb.has_bitmap)
        if not ga.has_bitmap:
            continue
        assert (ga.x0, ga.y0, ga.bm_width, ga.bm_height) == (gb.x0, gb.y0, gb.bm_width, gb.bm_height)
        assert np.array_equal(ba, bb)
        px += ba.size
    assert px > 20000


def test_path_requests_equal_uploaded_segments_for_random_outlines(renderer):
    """Kind PATH at the C ABI, without a font: random closed outlines of lines, quadratics and cubics (integer font
    units; some cubics with control points far outside, i.e. ten levels of subdivision; large coordinates at
    upm 16384).  The decode kernel's flattening must give exactly the segments the reference's flattening gives
    (oracle, f64, literal stack): the bitmap of every PATH request equals the bitmap of the same glyph sent as
    uploaded segments, byte for byte, and the reference's leaf counts are accepted."""
    import ctypes as C

    from versatiles_glyphs_rs_b200.api import CURVE_DT, GLYPH_REQ_DT

    rng = np.random.default_rng(17)
    ctx = V.SdfContext.of_renderer(renderer)
    recs, segs, reqs = [], [], []
    out_off = gen_off = tile_cap = 0
    n_cubic_leaves = 0
    for g in range(120):
        upm, span = (1000, 1000) if g % 3 else (16384, 30000)
        scale, dx = 24.0 / upm, float(rng.uniform(-0.25, 0.25))
        k = int(rng.integers(3, 7))
        anchors = rng.integers(-span // 8, span, size=(k, 2)).astype(np.float64)
        first_rec, seg_off, pts = len(recs), 0, [anchors[0]]
        for i in range(k):
            s, e = anchors[i], anchors[(i + 1) % k]
            kind = int(rng.integers(0, 3))
            if kind == 0:
                recs.append((s[0], s[1], s[0], s[1], e[0], e[1], seg_off, 0))
                pts.append(e)
                seg_off += 1
            elif kind == 1:
                c = rng.integers(-span // 8, span, size=2).astype(np.float64)
                p = O.flatten_quad(s, c, e)
                depth = int(np.log2(len(p)))
                assert 1 << depth == len(p) and depth <= 12
                recs.append((s[0], s[1], c[0], c[1], e[0], e[1], seg_off, depth))
                pts.extend(p)
                seg_off += len(p)
            else:
                far = span * (4 if g % 5 == 0 else 1)
                c1, c2 = (rng.integers(-far, far, size=2).astype(np.float64) for _ in range(2))
                p = O.flatten_cubic(s, c1, c2, e)
                assert 1 <= len(p) < 4096
                recs.append((s[0], s[1], c1[0], c1[1], c2[0], c2[1], seg_off, N.CURVE_CUBIC | len(p)))
                recs.append((e[0], e[1], 0, 0, 0, 0, 0, N.CURVE_TAIL))
                pts.extend(p)
                seg_off += len(p)
                n_cubic_leaves += len(p)
        pts = np.array(pts)
        px, py = pts[:, 0] * scale + dx, pts[:, 1] * scale + 0.0
        x0, y0 = int(np.floor(px.min())) - 3, int(np.floor(py.min())) - 3
        w, h = int(np.ceil(px.max())) + 3 - x0, int(np.ceil(py.max())) + 3 - y0
        fx, fy = (px - x0).astype(np.float32), (py - y0).astype(np.float32)
        first_seg = sum(len(sg) for sg in segs)
        segs.append(np.stack([fx[:-1], fy[:-1], fx[1:], fy[1:]], axis=1))
        assert len(segs[-1]) == seg_off
        for kind, src_off, src_cnt in ((N.KIND_PATH, first_rec, len(recs) - first_rec), (N.KIND_SEGMENTS, first_seg, seg_off)):
            r = np.zeros(1, GLYPH_REQ_DT)[0]
            r["kind"], r["src_off"], r["src_cnt"], r["seg_cnt"] = kind, src_off, src_cnt, seg_off
            r["width"], r["height"], r["x0"], r["y0"], r["scale"], r["dx"] = w, h, x0, y0, scale, dx
            r["out_off"], r["out_cap"] = out_off, w * h
            if kind == N.KIND_PATH:
                r["curve_off"], r["curve_cap"] = gen_off, seg_off
                gen_off += seg_off
            out_off += (w * h + 15) & ~15
            tile_cap += N.sdf.b200sdf_glyph_tile_bound(w, h)
            reqs.append(r)
    reqs = np.array(reqs, GLYPH_REQ_DT)
    curves = np.array(recs, dtype=CURVE_DT)
    frames, out = ctx.render_glyphs(reqs, np.zeros(0, V.api.GLYPH_PART_DT), 1, tile_cap, out_off, curves=curves,
                                    segs=np.concatenate(segs), est_cost=0)
    assert (frames["status"] == N.GLYPH_OK).all(), frames["status"]
    assert n_cubic_leaves > 2000
    px = 0
    for a, b in zip(reqs[0::2], reqs[1::2]):
        n = int(a["width"]) * int(a["height"])
        assert np.array_equal(out[int(a["out_off"]) : int(a["out_off"]) + n], out[int(b["out_off"]) : int(b["out_off"]) + n])
        px += n
    assert px > 100000
    # a leaf count that is not the subdivision's: the device notices and hands the glyph back
    bad = curves.copy()
    k = int(np.nonzero(bad["depth"] & N.CURVE_CUBIC)[0][0])
    bad["depth"][k] += 1
    frames, _ = ctx.render_glyphs(reqs, np.zeros(0, V.api.GLYPH_PART_DT), 1, tile_cap, out_off, curves=bad, segs=np.concatenate(segs))
    hit = next(i for i in range(0, len(reqs), 2) if reqs["src_off"][i] <= k < reqs["src_off"][i] + reqs["src_cnt"][i])
    assert frames["status"][hit] == N.GLYPH_NEEDS_HOST
    assert (np.delete(frames["status"], hit) == N.GLYPH_OK).all()
    assert N.sdf.b200sdf_reserve_glyphs(ctx._h, 4096, 1 << 16, 1 << 16, 1 << 14) == 0
