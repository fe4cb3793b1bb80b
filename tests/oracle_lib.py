"""ctypes binding of the CPU oracle (oracle/libvg_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libvg_oracle.so")
TESTDATA = os.path.join(ROOT, "testdata")
FIRA = os.path.join(TESTDATA, "Fira Sans - Regular.ttf")
NOTO_DIR = os.path.join(TESTDATA, "Noto Sans")

MODE_PRECISE, MODE_DUMMY = 0, 1


def noto_paths():
    """C2 input order: byte-wise lexicographic path order."""
    names = sorted(os.listdir(NOTO_DIR), key=lambda s: s.encode())
    return [os.path.join(NOTO_DIR, n) for n in names if n.endswith(".ttf")]


class Rings(C.Structure):
    _fields_ = [("xy", C.POINTER(C.c_double)), ("ring_start", C.POINTER(C.c_uint32)),
                ("n_rings", C.c_uint32), ("n_points", C.c_uint32)]


class Glyph(C.Structure):
    _fields_ = [("id", C.c_uint32), ("has_bitmap", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("left", C.c_int32), ("top", C.c_int32), ("advance", C.c_uint32),
                ("bitmap", C.POINTER(C.c_uint8)), ("bitmap_len", C.c_uint64), ("n_segments", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("glyphs", "bitmaps", "pixels", "segments", "pairs", "pbf_bytes", "pbf_checksum")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64, dbl = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_double
    pd = C.POINTER(C.c_double)
    sig = {
        "vgo_font_parse": (vp, [C.c_char_p, C.c_size_t]),
        "vgo_font_free": (None, [vp]),
        "vgo_font_units_per_em": (u32, [vp]),
        "vgo_font_num_glyphs": (u32, [vp]),
        "vgo_font_glyph_index": (i32, [vp, u32]),
        "vgo_font_hor_advance": (i32, [vp, u32]),
        "vgo_font_codepoints": (C.c_size_t, [vp, C.POINTER(u32), C.c_size_t]),
        "vgo_outline_rings": (C.c_int, [vp, u32, C.POINTER(Rings)]),
        "vgo_rings_free": (None, [C.POINTER(Rings)]),
        "vgo_flatten_quad": (C.c_size_t, [pd, pd, pd, dbl, pd, C.c_size_t]),
        "vgo_flatten_cubic": (C.c_size_t, [pd, pd, pd, pd, dbl, pd, C.c_size_t]),
        "vgo_segment_sqdist": (dbl, [dbl] * 6),
        "vgo_min_distance": (dbl, [pd, u32, dbl, dbl, dbl]),
        "vgo_renderer_precise": (C.c_int, [i32, i32, u32, u32, pd, C.POINTER(u32), u32, C.POINTER(C.c_uint8)]),
        "vgo_render_glyph": (C.c_int, [vp, u32, C.c_int, C.POINTER(Glyph)]),
        "vgo_glyph_free": (None, [C.POINTER(Glyph)]),
        "vgo_glyph_segments": (u32, [vp, u32, C.POINTER(pd), C.POINTER(i32 * 4)]),
        "vgo_free": (None, [vp]),
        "vgo_fontset_new": (vp, [C.c_char_p]),
        "vgo_fontset_free": (None, [vp]),
        "vgo_fontset_add": (None, [vp, vp]),
        "vgo_fontset_block_population": (None, [vp, C.POINTER(u32 * 256)]),
        "vgo_fontset_render_block": (C.c_int, [vp, u32, C.c_int, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(u64)]),
        "vgo_fontset_render_all": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, C.POINTER(Stats)]),
        "vgo_fontset_render_strided": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, u32, C.POINTER(Stats)]),
        "vgo_name_to_id": (C.c_char_p, [C.c_char_p, C.c_char_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class Font:
    """Restated ttf_parser::Face (the subset the path uses)."""

    def __init__(self, path_or_bytes):
        data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
        self._h = lib().vgo_font_parse(bytes(data), len(data))
        if not self._h:
            raise ValueError("Could not parse font data")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vgo_font_free(self._h)
            self._h = None

    @property
    def units_per_em(self):
        return lib().vgo_font_units_per_em(self._h)

    @property
    def number_of_glyphs(self):
        return lib().vgo_font_num_glyphs(self._h)

    def glyph_index(self, cp):
        g = lib().vgo_font_glyph_index(self._h, cp)
        return None if g < 0 else g

    def hor_advance(self, gid):
        a = lib().vgo_font_hor_advance(self._h, gid)
        return None if a < 0 else a

    def codepoints(self):
        n = lib().vgo_font_codepoints(self._h, None, 0)
        buf = (C.c_uint32 * max(n, 1))()
        lib().vgo_font_codepoints(self._h, buf, n)
        return list(buf[:n])

    def outline_rings(self, gid):
        """Flattened rings in font units: list of (n,2) float64 arrays."""
        r = Rings()
        lib().vgo_outline_rings(self._h, gid, C.byref(r))
        out = []
        for i in range(r.n_rings):
            a, b = r.ring_start[i], r.ring_start[i + 1]
            out.append(np.ctypeslib.as_array(r.xy, shape=(r.n_points * 2,))[2 * a:2 * b].reshape(-1, 2).copy())
        lib().vgo_rings_free(C.byref(r))
        return out

    def render_glyph(self, cp, mode=MODE_PRECISE):
        """Renderer::render_glyph → dict or None."""
        g = Glyph()
        if not lib().vgo_render_glyph(self._h, cp, mode, C.byref(g)):
            return None
        d = dict(id=g.id, width=g.width, height=g.height, left=g.left, top=g.top, advance=g.advance,
                 bitmap=None, n_segments=g.n_segments)
        if g.has_bitmap:
            d["bitmap"] = np.ctypeslib.as_array(g.bitmap, shape=(g.bitmap_len,)).copy()
        lib().vgo_glyph_free(C.byref(g))
        return d

    def glyph_segments(self, cp):
        """(segments float64 (n,4) in pixel space, (x0,y0,W,H)) exactly as renderer_precise sees them."""
        p = C.POINTER(C.c_double)()
        frame = (C.c_int32 * 4)()
        n = lib().vgo_glyph_segments(self._h, cp, C.byref(p), C.byref(frame))
        if n == 0:
            return np.zeros((0, 4)), tuple(frame)
        segs = np.ctypeslib.as_array(p, shape=(n * 4,)).reshape(n, 4).copy()
        lib().vgo_free(p)
        return segs, tuple(frame)


class FontSet:
    """FontWrapper (first file wins) + GlyphBlock::render + FontManager::render_glyphs."""

    def __init__(self, name, paths):
        buf = C.create_string_buffer(256)
        self.id = lib().vgo_name_to_id(name.encode(), buf, 256).decode()
        self.fonts = [Font(p) for p in paths]
        self._h = lib().vgo_fontset_new(self.id.encode())
        for f in self.fonts:
            lib().vgo_fontset_add(self._h, f._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vgo_fontset_free(self._h)
            self._h = None

    def block_population(self):
        out = (C.c_uint32 * 256)()
        lib().vgo_fontset_block_population(self._h, C.byref(out))
        return list(out)

    def render_block(self, block, mode=MODE_PRECISE):
        p = C.POINTER(C.c_uint8)()
        n = C.c_uint64()
        rc = lib().vgo_fontset_render_block(self._h, block, mode, C.byref(p), C.byref(n))
        assert rc == 0
        data = bytes(np.ctypeslib.as_array(p, shape=(n.value,))) if n.value else b""
        lib().vgo_free(p)
        return data

    def render_all(self, mode=MODE_PRECISE, threads=1, block_lo=0, block_hi=256, stride=1):
        st = Stats()
        lib().vgo_fontset_render_strided(self._h, mode, threads, block_lo, block_hi, stride, C.byref(st))
        return st.as_dict()


def renderer_precise(x0, y0, W, H, rings):
    """renderer_precise on explicit pixel-space rings (list of point lists)."""
    pts, starts = [], [0]
    for r in rings:
        pts.extend(r)
        starts.append(len(pts))
    xy = np.asarray(pts, dtype=np.float64).reshape(-1)
    rs = np.asarray(starts, dtype=np.uint32)
    bm = np.zeros(W * H, dtype=np.uint8)
    lib().vgo_renderer_precise(x0, y0, W, H, xy.ctypes.data_as(C.POINTER(C.c_double)),
                               rs.ctypes.data_as(C.POINTER(C.c_uint32)), len(rings),
                               bm.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bm


def flatten_quad(s, c, e, tol=0.01):
    arr = lambda p: (C.c_double * 2)(*p)
    out = (C.c_double * 4096)()
    n = lib().vgo_flatten_quad(arr(s), arr(c), arr(e), tol, out, 2048)
    return np.array(out[:2 * n]).reshape(-1, 2)


def flatten_cubic(s, c1, c2, e, tol=0.01):
    arr = lambda p: (C.c_double * 2)(*p)
    out = (C.c_double * 8192)()
    n = lib().vgo_flatten_cubic(arr(s), arr(c1), arr(c2), arr(e), tol, out, 4096)
    return np.array(out[:2 * n]).reshape(-1, 2)


def segment_sqdist(v, w, p):
    return lib().vgo_segment_sqdist(v[0], v[1], w[0], w[1], p[0], p[1])


def min_distance(segs, p, radius):
    a = np.asarray(segs, dtype=np.float64).reshape(-1)
    return lib().vgo_min_distance(a.ctypes.data_as(C.POINTER(C.c_double)), len(a) // 4, p[0], p[1], radius)


# ---- art decoders of the reference's golden bitmaps (src/utils/decode_bitmap.rs:15-28, 60-78) ----
def bitmap_as_digit_art(bitmap, width):
    rows = np.asarray(bitmap).reshape(-1, width)
    return [" ".join("%02d" % min(int(x) * 100 // 256, 99) for x in row) for row in rows]


def bitmap_as_ascii_art(bitmap, width):
    def band(x):
        return "  " if x <= 60 else "░░" if x <= 120 else "▒▒" if x <= 180 else "▓▓" if x <= 240 else "█"
    rows = np.asarray(bitmap).reshape(-1, width)
    return ["".join(band(int(x)) for x in row) for row in rows]


# ---- minimal glyphs-PBF decoder (mirror of prost decode in src/commands/debug.rs:60-79) ----
def _varint(buf, i):
    v = s = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return v, i


def _fields(buf):
    i = 0
    while i < len(buf):
        key, i = _varint(buf, i)
        tag, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 2:
            n, i = _varint(buf, i)
            v = bytes(buf[i:i + n])
            i += n
        else:
            raise ValueError("unexpected wire type %d" % wt)
        yield tag, wt, v


def _unzig(v):
    return (v >> 1) ^ -(v & 1)


def decode_pbf(data):
    """→ (name, range, [glyph dicts sorted by id])."""
    stacks = [v for t, _, v in _fields(data) if t == 1]
    assert len(stacks) == 1
    name = rng = None
    glyphs = []
    for t, _, v in _fields(stacks[0]):
        if t == 1:
            name = v.decode()
        elif t == 2:
            rng = v.decode()
        elif t == 3:
            g = dict(bitmap=None)
            for gt, _, gv in _fields(v):
                if gt == 1:
                    g["id"] = gv
                elif gt == 2:
                    g["bitmap"] = np.frombuffer(gv, dtype=np.uint8)
                elif gt == 3:
                    g["width"] = gv
                elif gt == 4:
                    g["height"] = gv
                elif gt == 5:
                    g["left"] = _unzig(gv)
                elif gt == 6:
                    g["top"] = _unzig(gv)
                elif gt == 7:
                    g["advance"] = gv
            glyphs.append(g)
    glyphs.sort(key=lambda g: g["id"])
    return name, rng, glyphs
