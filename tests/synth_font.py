"""Synthetic TrueType fonts for the parity tests and the bench (BASELINE.json configs C3 / C4).

A real .ttf is written (head, hhea, maxp, hmtx, cmap format 4, loca, glyf, name), so the product's
host (C++ Face) and the oracle (its own C parser) both go through their whole path: cmap lookup,
glyf decoding, flattening, metrics, SDF.  Outlines follow SURVEY.md §8(d): every glyph is K
"strokes", a stroke = a rounded rectangle (4 lines + 4 quadratic corners), alternate strokes are
wound the other way so winding numbers other than 0/1 occur.
"""
import struct

import numpy as np


def _table_checksum(data: bytes) -> int:
    pad = (-len(data)) % 4
    arr = np.frombuffer(data + b"\0" * pad, dtype=">u4")
    return int(arr.sum(dtype=np.uint64) & 0xFFFFFFFF)


def _rounded_rect(x, y, w, h, r, reverse):
    """12 points (on, off, on per corner) of a rounded rectangle; returns (xs, ys, on_curve)."""
    r = min(r, w // 2 - 1, h // 2 - 1)
    r = max(r, 1)
    pts = [
        (x + r, y, 1), (x + w - r, y, 1), (x + w, y, 0),
        (x + w, y + r, 1), (x + w, y + h - r, 1), (x + w, y + h, 0),
        (x + w - r, y + h, 1), (x + r, y + h, 1), (x, y + h, 0),
        (x, y + h - r, 1), (x, y + r, 1), (x, y, 0),
    ]
    if reverse:
        pts = pts[::-1]
    return pts


def _glyph_bytes(contours):
    """Simple-glyph record from a list of contours (lists of (x, y, on))."""
    if not contours:
        return b""
    xs = np.array([p[0] for c in contours for p in c], dtype=np.int32)
    ys = np.array([p[1] for c in contours for p in c], dtype=np.int32)
    on = np.array([p[2] for c in contours for p in c], dtype=np.uint8)
    ends = np.cumsum([len(c) for c in contours]) - 1
    dx = np.diff(xs, prepend=0).astype(">i2")
    dy = np.diff(ys, prepend=0).astype(">i2")
    head = struct.pack(">hhhhh", len(contours), int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max()))
    body = ends.astype(">u2").tobytes() + struct.pack(">H", 0) + on.tobytes() + dx.tobytes() + dy.tobytes()
    data = head + body
    return data + b"\0" * ((-len(data)) % 4)


def make_outline(rng, k_strokes, hole_every=2):
    contours = []
    for s in range(k_strokes):
        w = int(rng.integers(60, 701))
        h = int(rng.integers(30, 121))
        if rng.integers(0, 2):
            w, h = h, w
        x = int(rng.integers(0, max(1, 1000 - w)))
        y = int(rng.integers(-200, max(-199, 800 - h)))
        r = int(rng.integers(20, 61))
        rect = _rounded_rect(x, y, w, h, r, reverse=False)
        if s % 3 == 2:
            rect = rect[2:] + rect[:2]  # contour whose first point is off-curve
        contours.append(rect)
        if hole_every and s % hole_every == 1 and w > 40 and h > 40:
            contours.append(_rounded_rect(x + 12, y + 12, w - 24, h - 24, max(2, r - 12), reverse=True))
    return contours


def build_font(codepoints, strokes_for, seed=0xB200, family="Synth B200", cmap_format=4, many_to_one=None):
    """codepoints: ascending BMP code points (no surrogates, no 0xFFFF); strokes_for(cp) -> K.
    Glyph id i+1 belongs to codepoints[i]; glyph 0 is an empty .notdef."""
    cps = [int(c) for c in codepoints]
    assert cps == sorted(set(cps)) and all(0 <= c < 0xFFFF and not 0xD800 <= c <= 0xDFFF for c in cps)
    n_glyphs = len(cps) + 1
    assert n_glyphs <= 0xFFFF
    rng = np.random.default_rng(seed)
    glyf = [b""]
    advances = [500]
    for cp in cps:
        glyf.append(_glyph_bytes(make_outline(rng, strokes_for(cp))))
        advances.append(int(rng.integers(400, 1101)))
    offsets = np.zeros(n_glyphs + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(g) for g in glyf])
    glyf_table = b"".join(glyf)
    loca_table = offsets.astype(">u4").tobytes()

    # cmap format 4: one segment per run of consecutive code points, idDelta maps to glyph ids
    runs = []
    start = prev = cps[0]
    gid0 = 1
    for i, c in enumerate(cps[1:], start=1):
        if c != prev + 1:
            runs.append((start, prev, gid0))
            start, gid0 = c, i + 1
        prev = c
    runs.append((start, prev, gid0))
    runs.append((0xFFFF, 0xFFFF, 0))
    segx2 = 2 * len(runs)
    ends = b"".join(struct.pack(">H", e) for _, e, _ in runs)
    starts = b"".join(struct.pack(">H", s) for s, _, _ in runs)
    deltas = b"".join(struct.pack(">H", (g - s) & 0xFFFF if s != 0xFFFF else 1) for s, _, g in runs)
    ranges = b"\0\0" * len(runs)
    sub = struct.pack(">HHHHHHH", 4, 16 + 4 * segx2, 0, segx2, 0, 0, 0) + ends + b"\0\0" + starts + deltas + ranges
    cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 1, 12) + sub
    real_runs = runs[:-1]
    if cmap_format == 12:  # segmented coverage, (platform 0, encoding 4)
        groups = b"".join(struct.pack(">III", s0, e0, g0) for s0, e0, g0 in real_runs)
        sub = struct.pack(">HHIII", 12, 0, 16 + len(groups), 0, len(real_runs)) + groups
        cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 0, 4, 12) + sub
    elif cmap_format == 13:  # many-to-one: many_to_one = [(first cp, last cp, glyph id)], ascending
        groups = b"".join(struct.pack(">III", s0, e0, g0) for s0, e0, g0 in many_to_one)
        sub = struct.pack(">HHIII", 13, 0, 16 + len(groups), 0, len(many_to_one)) + groups
        cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 10, 12) + sub
    elif cmap_format in (6, 10):  # trimmed arrays over [cps[0], cps[-1]]; holes map to glyph 0
        first, count = cps[0], cps[-1] - cps[0] + 1
        gids = np.zeros(count, dtype=">u2")
        for i, c in enumerate(cps):
            gids[c - first] = i + 1
        if cmap_format == 6:
            sub = struct.pack(">HHHHH", 6, 10 + 2 * count, 0, first, count) + gids.tobytes()
            cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 1, 12) + sub
        else:
            sub = struct.pack(">HHIIII", 10, 0, 20 + 2 * count, 0, first, count) + gids.tobytes()
            cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 0, 4, 12) + sub
    else:
        assert cmap_format == 4

    head = struct.pack(">IIIIHHqqhhhhHHhhh", 0x00010000, 0x00010000, 0, 0x5F0F3CF5, 0, 1000, 0, 0, 0, -200, 1000, 800,
                       0, 8, 2, 1, 0)
    hhea = struct.pack(">IhhhHhhhhhhhhhhhH", 0x00010000, 800, -200, 0, 1100, 0, 0, 1000, 1, 0, 0, 0, 0, 0, 0, 0, n_glyphs)
    maxp = struct.pack(">IH", 0x00005000, n_glyphs)
    hm = np.zeros((n_glyphs, 2), dtype=">u2")
    hm[:, 0] = advances
    hmtx = hm.tobytes()
    fam = family.encode("utf-16-be")
    name = struct.pack(">HHH", 0, 1, 6 + 12) + struct.pack(">HHHHHH", 3, 1, 0x409, 1, len(fam), 0) + fam

    tables = {b"cmap": cmap_table, b"glyf": glyf_table, b"head": head, b"hhea": hhea, b"hmtx": hmtx, b"loca": loca_table,
              b"maxp": maxp, b"name": name}
    tags = sorted(tables)
    n = len(tags)
    out = struct.pack(">IHHHH", 0x00010000, n, 128, 3, n * 16 - 128)
    offset = 12 + 16 * n
    records, blobs = b"", b""
    for tag in tags:
        data = tables[tag]
        records += tag + struct.pack(">III", _table_checksum(data), offset, len(data))
        padded = data + b"\0" * ((-len(data)) % 4)
        blobs += padded
        offset += len(padded)
    return out + records + blobs


def dense_font(n_glyphs=4096, first_cp=0x4E00, seed=0xB200):
    """C3: dense outlines, K in {8,16,32,64} in equal shares."""
    cps = list(range(first_cp, first_cp + n_glyphs))
    ks = (8, 16, 32, 64)
    return build_font(cps, lambda cp: ks[cp % 4], seed=seed, family="Synth Dense")


def full_bmp_font(seed=0xB200, stride=1):
    """C4: every BMP code point except surrogates and U+FFFF (format 4's terminator); K = 2 + hash(cp) % 14.
    stride > 1 keeps every stride-th code point (smaller fixture for tests)."""
    cps = [c for c in range(0, 0xFFFF, stride) if not 0xD800 <= c <= 0xDFFF]
    return build_font(cps, lambda cp: 2 + ((cp * 2654435761) >> 7) % 14, seed=seed, family="Synth Full")
