"""Independent cross-check of the outline decoding against FreeType (CPU test).

The reference gets its outlines from ttf-parser 0.25.1, which is not vendored; the oracle and the C++ host both
restate it.  What no golden of the reference pins — composite glyphs (offset-only, nested,
scaled), the raw flag / delta decoding of every glyph, CFF charstrings — is checked here against a third, unrelated
implementation: FreeType 2.14 as bundled with Pillow (pillow.libs), driven through ctypes with FT_LOAD_NO_SCALE
(font units, no hinting).

  * TrueType: FreeType's raw outline (points, on/off tags, contour ends — composites already resolved) is pushed
    through the contour rules of Appendix C and must reproduce the host's move/line/quad callback stream EXACTLY for
    every cmap-reachable BMP glyph of all 21 fixture fonts; advances must equal FreeType's.  Components that carry a
    scale are rounded to integers by FreeType and kept in f32 by ttf-parser: those few glyphs are compared within one
    font unit.
  * CFF (synthetic .otf fonts of tests/synth_font.py): FreeType's decomposed path (move/line/cubic) must equal the
    host's callback stream, i.e. the charstring interpreter agrees with FreeType's on every operator the fonts use.
"""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O  # noqa: E402
import synth_font  # noqa: E402
import versatiles_glyphs_rs_b200 as V  # noqa: E402

FT_LOAD_NO_SCALE = 1


class FreeType:
    """The few FreeType calls needed, with the 64-bit struct offsets of FT_FaceRec / FT_GlyphSlotRec (checked against
    units_per_EM and the glyph count of a known font when the library is opened)."""

    def __init__(self):
        import PIL
        import PIL._imagingft  # noqa: F401  (loads libfreetype and its siblings into the process)

        libs = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libfreetype-*.so*"))
        if not libs:
            raise OSError("no bundled libfreetype")
        self.ft = C.CDLL(libs[0])
        self.lib = C.c_void_p()
        assert self.ft.FT_Init_FreeType(C.byref(self.lib)) == 0
        self.ft.FT_New_Memory_Face.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_long, C.POINTER(C.c_void_p)]
        self.ft.FT_Get_Char_Index.argtypes = [C.c_void_p, C.c_ulong]
        self.ft.FT_Get_Char_Index.restype = C.c_uint
        self.ft.FT_Load_Glyph.argtypes = [C.c_void_p, C.c_uint, C.c_int]
        self.ft.FT_Done_Face.argtypes = [C.c_void_p]
        v = [C.c_int(), C.c_int(), C.c_int()]
        self.ft.FT_Library_Version(self.lib, *[C.byref(x) for x in v])
        self.version = tuple(x.value for x in v)
        self._keep = []

    def face(self, data: bytes):
        f = C.c_void_p()
        assert self.ft.FT_New_Memory_Face(self.lib, data, len(data), 0, C.byref(f)) == 0
        self._keep.append(data)
        return f

    @staticmethod
    def units_per_em(face):
        return C.c_ushort.from_address(face.value + 136).value

    @staticmethod
    def num_glyphs(face):
        return C.c_long.from_address(face.value + 32).value

    def load(self, face, gid):
        """-> (advance, points int64 [n, 2], tags uint8 [n], contour ends uint16 [c]) in font units, or None."""
        if self.ft.FT_Load_Glyph(face, gid, FT_LOAD_NO_SCALE) != 0:
            return None
        slot = C.c_void_p.from_address(face.value + 152).value
        adv = C.c_long.from_address(slot + 48 + 32).value  # metrics.horiAdvance
        nc = C.c_ushort.from_address(slot + 200).value
        npnt = C.c_ushort.from_address(slot + 202).value
        if npnt == 0:
            return adv, np.zeros((0, 2), np.int64), np.zeros(0, np.uint8), np.zeros(0, np.uint16)
        pts = np.ctypeslib.as_array((C.c_long * (2 * npnt)).from_address(C.c_void_p.from_address(slot + 208).value)).reshape(-1, 2).copy()
        tags = np.ctypeslib.as_array((C.c_ubyte * npnt).from_address(C.c_void_p.from_address(slot + 216).value)).copy()
        ends = np.ctypeslib.as_array((C.c_ushort * nc).from_address(C.c_void_p.from_address(slot + 224).value)).copy()
        return adv, pts.astype(np.int64), tags, ends


@pytest.fixture(scope="module")
def ft():
    try:
        lib = FreeType()
    except (ImportError, OSError) as e:
        pytest.skip(f"FreeType (pillow.libs) not available: {e}")
    f = lib.face(open(O.FIRA, "rb").read())
    assert lib.units_per_em(f) == 1000 and lib.num_glyphs(f) == 2677  # the struct offsets are right
    return lib


def ttf_parser_commands(pts, tags, ends):
    """SURVEY.md Appendix C (ttf-parser's glyf walker) over raw points: the expected callback stream, f32."""
    f32 = np.float32
    out = []
    if len(pts) == 1:
        return out  # a glyph with a single point yields nothing
    start = 0

    def mid(a, b):
        return (f32(a[0] + f32(0.5) * (b[0] - a[0])), f32(a[1] + f32(0.5) * (b[1] - a[1])))

    for end in ends:
        first_on = first_off = last_off = None
        for i in range(start, int(end) + 1):
            p = (f32(pts[i][0]), f32(pts[i][1]))
            on = bool(tags[i] & 1)
            if first_on is None:
                if on:
                    first_on = p
                    out.append((0, p))
                elif first_off is not None:
                    m = mid(first_off, p)
                    first_on = m
                    last_off = p
                    out.append((0, m))
                else:
                    first_off = p
            elif last_off is not None:
                c = last_off
                if on:
                    last_off = None
                    out.append((2, c, p))
                else:
                    last_off = p
                    out.append((2, c, mid(c, p)))
            elif on:
                out.append((1, p))
            else:
                last_off = p
        if first_off is not None and last_off is not None:
            c = last_off
            last_off = None
            out.append((2, c, mid(c, first_off)))
        if first_on is not None and first_off is not None:
            out.append((2, first_off, first_on))
        elif first_on is not None and last_off is not None:
            out.append((2, last_off, first_on))
        elif first_on is not None:
            out.append((1, first_on))
        out.append((4,))
        start = int(end) + 1
    return out


def as_rows(cmds):
    rows = np.zeros((len(cmds), 7), np.float32)
    for k, c in enumerate(cmds):
        rows[k, 0] = c[0]
        if c[0] in (0, 1):
            rows[k, 5:7] = c[1]
        elif c[0] == 2:
            rows[k, 1:3] = c[1]
            rows[k, 5:7] = c[2]
    return rows


@pytest.mark.parametrize("path", [O.FIRA] + O.noto_paths())
def test_truetype_outlines_and_advances_match_freetype(ft, path):
    data = open(path, "rb").read()
    font = V.FontFileEntry(data=data)
    face = ft.face(data)
    cps = [int(c) for c in font.codepoints() if c <= 0xFFFF]
    gids = sorted({font.glyph_index(cp) for cp in cps})
    exact = scaled = points = 0
    for cp in cps[:: max(1, len(cps) // 200)]:  # cmap: our lookup against FreeType's, on a sample
        assert ft.ft.FT_Get_Char_Index(face, cp) == font.glyph_index(cp), hex(cp)
    for gid in gids:
        got = ft.load(face, gid)
        assert got is not None, gid
        adv, pts, tags, ends = got
        assert adv == (font.glyph_hor_advance(gid) or 0), gid
        assert not (tags & 2).any(), "cubic points in a glyf font"
        want = as_rows(ttf_parser_commands(pts, tags, ends))
        have = font.outline_commands(gid)
        assert have.shape == want.shape, (gid, have.shape, want.shape)
        points += len(pts)
        if np.array_equal(have, want):
            exact += 1
            continue
        # a component with a scale: FreeType rounds the scaled points to integers, ttf-parser keeps them in f32
        assert np.array_equal(have[:, 0], want[:, 0]), gid
        assert np.abs(have - want).max() <= 1.0, (gid, float(np.abs(have - want).max()))
        scaled += 1
    assert scaled <= 6, (path, scaled)  # 4 in Noto Sans Arabic, 2 in Myanmar, none elsewhere
    assert exact + scaled == len(gids) and points > 0
    print(f"{os.path.basename(path)}: {len(gids)} glyphs, {points} points, {exact} exact, {scaled} within 1 unit (scaled components)")


def test_edge_case_font_matches_freetype(ft):
    """The synthetic glyphs that walk every branch of the contour rules and of the composite resolution."""
    from test_gpu_glyf import _edge_case_font

    blob, cps, names = _edge_case_font()
    font = V.FontFileEntry(data=blob)
    face = ft.face(blob)
    n_scaled = 0
    for cp in cps:
        gid = font.glyph_index(cp)
        adv, pts, tags, ends = ft.load(face, gid)
        assert adv == font.glyph_hor_advance(gid)
        want = as_rows(ttf_parser_commands(pts, tags, ends))
        have = font.outline_commands(gid)
        assert have.shape == want.shape, hex(cp)
        if not np.array_equal(have, want):
            assert np.abs(have - want).max() <= 1.0, hex(cp)
            n_scaled += 1
    assert n_scaled <= 2  # the scaled and the rotated composite


def ft_decompose(ft, face, gid):
    """FT_Outline_Decompose of the loaded glyph -> rows like FontFileEntry.outline_commands (no close rows)."""
    assert ft.ft.FT_Load_Glyph(face, gid, FT_LOAD_NO_SCALE) == 0
    slot = C.c_void_p.from_address(face.value + 152).value
    rows = []

    class Vec(C.Structure):
        _fields_ = [("x", C.c_long), ("y", C.c_long)]

    MOVE = C.CFUNCTYPE(C.c_int, C.POINTER(Vec), C.c_void_p)
    CONIC = C.CFUNCTYPE(C.c_int, C.POINTER(Vec), C.POINTER(Vec), C.c_void_p)
    CUBIC = C.CFUNCTYPE(C.c_int, C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec), C.c_void_p)

    class Funcs(C.Structure):
        _fields_ = [("move_to", MOVE), ("line_to", MOVE), ("conic_to", CONIC), ("cubic_to", CUBIC), ("shift", C.c_int), ("delta", C.c_long)]

    def mv(p, _):
        rows.append([0, 0, 0, 0, 0, p[0].x, p[0].y])
        return 0

    def ln(p, _):
        rows.append([1, 0, 0, 0, 0, p[0].x, p[0].y])
        return 0

    def cn(c, p, _):
        rows.append([2, c[0].x, c[0].y, 0, 0, p[0].x, p[0].y])
        return 0

    def cb(a, b, p, _):
        rows.append([3, a[0].x, a[0].y, b[0].x, b[0].y, p[0].x, p[0].y])
        return 0

    funcs = Funcs(MOVE(mv), MOVE(ln), CONIC(cn), CUBIC(cb), 0, 0)
    ft.ft.FT_Outline_Decompose.argtypes = [C.c_void_p, C.POINTER(Funcs), C.c_void_p]
    assert ft.ft.FT_Outline_Decompose(slot + 200, C.byref(funcs), None) == 0
    return np.array(rows, np.float32).reshape(-1, 7)


@pytest.mark.parametrize("cid", [False, True])
def test_cff_charstrings_match_freetype(ft, cid):
    """CFF 1 outlines (host/cff.cc): same path as FreeType's interpreter, operator for operator."""
    blob, cps, _ = synth_font.cff_test_font(n_glyphs=40, cid=cid)
    font = V.FontFileEntry(data=blob)
    face = ft.face(blob)
    checked = 0
    for cp in cps:
        gid = font.glyph_index(cp)
        have = font.outline_commands(gid)
        want = ft_decompose(ft, face, gid)
        # ttf-parser closes every contour explicitly (close rows) and, like FreeType, returns to the start point with a
        # line when the charstring did not: compare the drawing commands after dropping close rows and closing lines that
        # end on the contour's start
        def drawing(rows):
            out, start = [], None
            for r in rows:
                if r[0] == 4:
                    continue
                if r[0] == 0:
                    start = (r[5], r[6])
                out.append(r)
            return np.array(out, np.float32).reshape(-1, 7), start

        h, _ = drawing(have)
        w, _ = drawing(want)

        def strip_closing_lines(rows):
            keep, start, last = [], None, None
            for k, r in enumerate(rows):
                if r[0] == 0:
                    start = (r[5], r[6])
                nxt_is_new = k + 1 == len(rows) or rows[k + 1][0] == 0
                if r[0] == 1 and nxt_is_new and (r[5], r[6]) == start:
                    continue
                keep.append(r)
            return np.array(keep, np.float32).reshape(-1, 7)

        h, w = strip_closing_lines(h), strip_closing_lines(w)
        assert h.shape == w.shape, (hex(cp), h.shape, w.shape)
        assert np.array_equal(h, w), hex(cp)
        checked += 1
    assert checked >= 30


def test_cff2_charstrings_match_freetype(ft):
    """CFF 2 outlines at the default instance (host/cff.cc parse2 / blend / vsindex): same path as FreeType's
    interpreter, operator for operator — including the glyph whose variation region has scalar 1 at the default."""
    blob, cps, _ = synth_font.cff2_test_font()
    font = V.FontFileEntry(data=blob)
    face = ft.face(blob)
    checked = 0
    for cp in cps:
        gid = font.glyph_index(cp)
        have = font.outline_commands(gid)
        want = ft_decompose(ft, face, gid)

        def drawing(rows):
            return np.array([r for r in rows if r[0] != 4], np.float32).reshape(-1, 7)

        def strip_closing_lines(rows):
            keep, start = [], None
            for k, r in enumerate(rows):
                if r[0] == 0:
                    start = (r[5], r[6])
                nxt_is_new = k + 1 == len(rows) or rows[k + 1][0] == 0
                if r[0] == 1 and nxt_is_new and (r[5], r[6]) == start:
                    continue
                keep.append(r)
            return np.array(keep, np.float32).reshape(-1, 7)

        h, w = strip_closing_lines(drawing(have)), strip_closing_lines(drawing(want))
        assert h.shape == w.shape, (hex(cp), h.shape, w.shape)
        assert np.array_equal(h, w), hex(cp)
        checked += 1
    assert checked >= 30


def test_cmap_with_variation_selector_subtable_matches_freetype(ft):
    """cmap format 14 in front of the Unicode subtable: plain lookups are FreeType's (which also ignores it for them)."""
    cps = [0x41, 0x42, 0x43, 0x50] + list(range(0x4E00, 0x4E08))
    plain = synth_font.build_font(cps, lambda cp: 3, seed=9, family="UVS Test", cmap_format=12)
    blob = synth_font.with_variation_selectors(plain, [(0xFE00, [(0x41, 2), (0x4E00, 3)])])
    font = V.FontFileEntry(data=blob)
    face = ft.face(blob)
    for cp in cps + [0x20, 0xFE00, 0x4E09]:
        assert ft.ft.FT_Get_Char_Index(face, cp) == (font.glyph_index(cp) or 0), hex(cp)


def test_cmap_format_2_matches_freetype(ft):
    """cmap format 2: the mapped one-byte and two-byte codes (and plainly unmapped ones) resolve as in FreeType."""
    cps = list(range(0x4E00, 0x4E10))
    base = synth_font.build_font(cps, lambda cp: 3, seed=5, family="Cmap2 Test")
    blob = synth_font.with_cmap_format2(base, (0x41, [1, 2, 0, 4, 5]),
                                        [(0x4E, 0x00, 5, [1, 2, 3, 4, 0, 6, 7, 8]), (0x81, 0x40, -2, [12, 13, 14, 2])])
    font = V.FontFileEntry(data=blob)
    face = ft.face(blob)
    for cp in [0x41, 0x42, 0x43, 0x44, 0x45, 0x46, 0x20] + list(range(0x4E00, 0x4E09)) + list(range(0x8140, 0x8145)) + [0x9000]:
        assert ft.ft.FT_Get_Char_Index(face, cp) == (font.glyph_index(cp) or 0), hex(cp)
