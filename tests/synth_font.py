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


def _glyph_bytes_compact(contours, instructions=b""):
    """The same record with the compact encodings real fonts use: one-byte deltas (flags 0x02 / 0x04 with the sign
    bits), "same as previous" coordinates (0x10 / 0x20 without a byte), and runs of equal flags folded with REPEAT (0x08)."""
    if not contours:
        return b""
    pts = [p for c in contours for p in c]
    ends = np.cumsum([len(c) for c in contours]) - 1
    flags, xb, yb = [], b"", b""
    px = py = 0
    for x, y, on in pts:
        f = 1 if on else 0
        dx, dy = x - px, y - py
        px, py = x, y
        if dx == 0:
            f |= 0x10
        elif -255 <= dx <= 255:
            f |= 0x02 | (0x10 if dx > 0 else 0)
            xb += bytes([abs(dx)])
        else:
            xb += struct.pack(">h", dx)
        if dy == 0:
            f |= 0x20
        elif -255 <= dy <= 255:
            f |= 0x04 | (0x20 if dy > 0 else 0)
            yb += bytes([abs(dy)])
        else:
            yb += struct.pack(">h", dy)
        flags.append(f)
    fb, i = b"", 0
    while i < len(flags):
        j = i
        while j + 1 < len(flags) and flags[j + 1] == flags[i] and j - i < 255:
            j += 1
        if j > i:
            fb += bytes([flags[i] | 0x08, j - i])
        else:
            fb += bytes([flags[i]])
        i = j + 1
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]
    head = struct.pack(">hhhhh", len(contours), min(xs), min(ys), max(xs), max(ys))
    data = head + ends.astype(">u2").tobytes() + struct.pack(">H", len(instructions)) + instructions + fb + xb + yb
    return data + b"\0" * ((-len(data)) % 4)


def composite_bytes(components, bbox=(0, 0, 1000, 1000)):
    """Composite glyph record: components = [(glyph id, dx, dy, transform)], transform = None (translation only),
    a float (uniform scale), (sx, sy) or (a, b, c, d)."""
    out = struct.pack(">hhhhh", -1, *bbox)
    for k, (gid, dx, dy, tr) in enumerate(components):
        fl = 0x0002  # ARGS_ARE_XY
        words = not (-128 <= dx <= 127 and -128 <= dy <= 127)
        if words:
            fl |= 0x0001
        if k + 1 < len(components):
            fl |= 0x0020
        f2 = lambda v: struct.pack(">h", int(round(v * 16384)))
        tail = b""
        if tr is not None:
            if isinstance(tr, (int, float)):
                fl |= 0x0008
                tail = f2(tr)
            elif len(tr) == 2:
                fl |= 0x0040
                tail = f2(tr[0]) + f2(tr[1])
            else:
                fl |= 0x0080
                tail = b"".join(f2(v) for v in tr)
        out += struct.pack(">HH", fl, gid) + (struct.pack(">hh", dx, dy) if words else struct.pack(">bb", dx, dy)) + tail
    return out + b"\0" * ((-len(out)) % 4)


def build_font(codepoints, strokes_for, seed=0xB200, family="Synth B200", cmap_format=4, many_to_one=None, records=None,
               extra_records=(), compact=False):
    """codepoints: ascending BMP code points (no surrogates, no 0xFFFF); strokes_for(cp) -> K.
    Glyph id i+1 belongs to codepoints[i]; glyph 0 is an empty .notdef.
    records: explicit glyf records for the code points instead of generated strokes; extra_records: further records
    (components of composites) that get the glyph ids after the mapped ones; compact: generated strokes use the
    compact glyf encodings."""
    cps = [int(c) for c in codepoints]
    assert cps == sorted(set(cps)) and all(0 <= c < 0xFFFF and not 0xD800 <= c <= 0xDFFF for c in cps)
    n_glyphs = len(cps) + 1 + len(extra_records)
    assert n_glyphs <= 0xFFFF
    rng = np.random.default_rng(seed)
    glyf = [b""]
    advances = [500]
    for i, cp in enumerate(cps):
        if records is not None:
            glyf.append(records[i])
        else:
            contours = make_outline(rng, strokes_for(cp))
            glyf.append(_glyph_bytes_compact(contours) if compact else _glyph_bytes(contours))
        advances.append(int(rng.integers(400, 1101)))
    for rec in extra_records:
        glyf.append(rec)
        advances.append(600)
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
    # left side bearing = the record's xMin, as in any well-formed font (FreeType shifts an outline whose xMin differs
    # from its lsb; ttf-parser — the reference — does not look at it)
    hm[:, 1] = [struct.unpack(">h", g[2:4])[0] & 0xFFFF if len(g) >= 10 else 0 for g in glyf]
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


# ---------------------------------------------------------------------------------------------------------------
# Synthetic CFF (.otf) fonts — SURVEY.md §8(f) rank 3.  Glyphs are given as absolute path commands
# ('M', x, y) / ('L', x, y) / ('C', x1, y1, x2, y2, x, y) with integer coordinates; the encoder below turns them
# into Type 2 charstrings and deliberately spreads them over every operator and number encoding, so the
# absolute commands are the known answer for both CFF interpreters (host C++ and oracle C).
# ---------------------------------------------------------------------------------------------------------------
def _t2_num(v, style=0):
    """Type 2 charstring operand.  style 1 forces the 3-byte form (28), style 2 the 16.16 form (255)."""
    v = int(v)
    if style == 2:
        return b"\xff" + struct.pack(">i", v << 16)
    if style == 1 or not -1131 <= v <= 1131:
        return b"\x1c" + struct.pack(">h", v)
    if -107 <= v <= 107:
        return bytes([v + 139])
    if v > 0:
        v -= 108
        return bytes([247 + (v >> 8), v & 255])
    v = -v - 108
    return bytes([251 + (v >> 8), v & 255])


def _dict_int5(v):
    return b"\x1d" + struct.pack(">i", int(v))


def _dict_real(text):
    nib = {".": 0xA, "E": 0xB, "-": 0xE}
    out = []
    i = 0
    while i < len(text):
        if text[i : i + 2] == "E-":
            out.append(0xC)
            i += 2
        else:
            out.append(nib.get(text[i], int(text[i]) if text[i].isdigit() else None))
            i += 1
    out.append(0xF)
    if len(out) & 1:
        out.append(0xF)
    return b"\x1e" + bytes((out[k] << 4) | out[k + 1] for k in range(0, len(out), 2))


def _cff_index(items):
    if not items:
        return b"\0\0"
    offs = [1]
    for it in items:
        offs.append(offs[-1] + len(it))
    osz = 1 if offs[-1] < 256 else 2 if offs[-1] < 65536 else 4
    enc = b"".join(o.to_bytes(osz, "big") for o in offs)
    return struct.pack(">HB", len(items), osz) + enc + b"".join(items)


def _subr_bias(n):
    return 107 if n < 1240 else 1131 if n < 33900 else 32768


def encode_charstring(cmds, variant=0, width=None, hints=False, call=None, endchar=True, blend_regions=0):
    """Absolute commands -> Type 2 charstring.  `variant` rotates which of the equivalent operators is used;
    `call` = (operator byte 10 / 29, biased index) is issued right after the first moveto (the subroutine
    draws a closed square relative to the current point and returns to it, see cff_test_font)."""
    out = b""
    x = y = 0
    first = True
    stack_w = [] if width is None else [_t2_num(width)]
    if hints:  # hstemhm (+ width), then hintmask with two implied vstems: 3 stems -> one mask byte
        out += b"".join(stack_w) + _t2_num(10) + _t2_num(20) + b"\x12"
        stack_w = []
        out += _t2_num(30) + _t2_num(40) + _t2_num(50) + _t2_num(60) + b"\x13\xe0"
    i = 0
    k = variant
    blend_ctr = [variant]
    while i < len(cmds):
        c = cmds[i]
        k += 1
        style = (0, 0, 1, 0, 2)[k % 5]
        num = lambda v: _t2_num(v, style)
        if blend_regions:  # CFF 2: every other operand carries variation deltas (default d1 .. dk 1 blend)
            def num(v, _style=style, _ctr=blend_ctr):
                _ctr[0] += 1
                if _ctr[0] % 2:
                    return _t2_num(v, _style)
                deltas = b"".join(_t2_num(((_ctr[0] * 7 + r * 13) % 41) - 20) for r in range(blend_regions))
                return _t2_num(v, _style) + deltas + _t2_num(1) + b"\x10"
        if c[0] == "M":
            dx, dy = c[1] - x, c[2] - y
            w = b"".join(stack_w)
            stack_w = []
            if dy == 0 and k % 2:
                out += w + num(dx) + b"\x16"
            elif dx == 0 and k % 2:
                out += w + num(dy) + b"\x04"
            else:
                out += w + num(dx) + num(dy) + b"\x15"
            x, y = c[1], c[2]
            if first and call is not None:
                out += _t2_num(call[1]) + bytes([call[0]])
            first = False
            i += 1
        elif c[0] == "L":
            run = []
            while i < len(cmds) and cmds[i][0] == "L":
                run.append(cmds[i])
                i += 1
            # alternating horizontal / vertical runs use hlineto / vlineto, the rest rlineto; a run that is
            # followed by a curve may become rlinecurve
            j = 0
            while j < len(run):
                dx, dy = run[j][1] - x, run[j][2] - y
                if (dx == 0) != (dy == 0):
                    horiz = dy == 0
                    args, h, xx, yy, jj = [], horiz, x, y, j
                    while jj < len(run):
                        ddx, ddy = run[jj][1] - xx, run[jj][2] - yy
                        if h and ddy == 0 and ddx != 0:
                            args.append(ddx)
                        elif not h and ddx == 0 and ddy != 0:
                            args.append(ddy)
                        else:
                            break
                        xx, yy = run[jj][1], run[jj][2]
                        h = not h
                        jj += 1
                    out += b"".join(num(a) for a in args) + (b"\x06" if horiz else b"\x07")
                    x, y, j = xx, yy, jj
                else:
                    args = []
                    while j < len(run):
                        dx, dy = run[j][1] - x, run[j][2] - y
                        if (dx == 0) != (dy == 0) and args:
                            break
                        args += [dx, dy]
                        x, y = run[j][1], run[j][2]
                        j += 1
                    if j == len(run) and i < len(cmds) and cmds[i][0] == "C" and k % 3 == 0 and len(args) <= 30:
                        cc = cmds[i]
                        args += [cc[1] - x, cc[2] - y, cc[3] - cc[1], cc[4] - cc[2], cc[5] - cc[3], cc[6] - cc[4]]
                        x, y = cc[5], cc[6]
                        i += 1
                        out += b"".join(num(a) for a in args) + b"\x19"  # rlinecurve
                    else:
                        out += b"".join(num(a) for a in args) + b"\x05"
        else:
            d = [c[1] - x, c[2] - y, c[3] - c[1], c[4] - c[2], c[5] - c[3], c[6] - c[4]]
            nxt = cmds[i + 1] if i + 1 < len(cmds) else None
            if nxt is not None and nxt[0] == "C" and k % 4 == 0:  # two curves as flex (depth argument unused)
                d2 = [nxt[1] - c[5], nxt[2] - c[6], nxt[3] - nxt[1], nxt[4] - nxt[2], nxt[5] - nxt[3], nxt[6] - nxt[4]]
                out += b"".join(num(a) for a in d + d2) + num(50) + b"\x0c\x23"
                x, y = nxt[5], nxt[6]
                i += 2
                continue
            delta = lambda c0, px, py: [c0[1] - px, c0[2] - py, c0[3] - c0[1], c0[4] - c0[2], c0[5] - c0[3], c0[6] - c0[4]]
            hv = lambda q: q[1] == 0 and q[4] == 0  # horizontal start, vertical end
            vh = lambda q: q[0] == 0 and q[5] == 0  # vertical start, horizontal end
            if hv(d) or vh(d):
                # hvcurveto / vhcurveto: following curves join the same operator while they keep alternating
                args, horiz, op = [], hv(d), (b"\x1f" if hv(d) else b"\x1e")
                while True:
                    args += [d[0], d[2], d[3], d[5]] if horiz else [d[1], d[2], d[3], d[4]]
                    x, y = c[5], c[6]
                    i += 1
                    horiz = not horiz
                    if i >= len(cmds) or cmds[i][0] != "C" or len(args) > 36:
                        break
                    c = cmds[i]
                    d = delta(c, x, y)
                    if not (hv(d) if horiz else vh(d)):
                        break
                out += b"".join(num(a) for a in args) + op
                continue
            elif d[1] == 0 and d[5] == 0:
                out += b"".join(num(a) for a in (d[0], d[2], d[3], d[4])) + b"\x1b"  # hhcurveto
            elif d[0] == 0 and d[4] == 0:
                out += b"".join(num(a) for a in (d[1], d[2], d[3], d[5])) + b"\x1a"  # vvcurveto
            elif d[5] == 0 and k % 2:
                out += b"".join(num(a) for a in (d[1], d[0], d[2], d[3], d[4])) + b"\x1b"  # hhcurveto with dy1
            elif d[4] == 0 and k % 2:
                out += b"".join(num(a) for a in (d[0], d[1], d[2], d[3], d[5])) + b"\x1a"  # vvcurveto with dx1
            elif d[1] == 0 and k % 2:
                out += b"".join(num(a) for a in (d[0], d[2], d[3], d[5], d[4])) + b"\x1f"  # hvcurveto, 5 arguments
            elif d[0] == 0 and k % 2:
                out += b"".join(num(a) for a in (d[1], d[2], d[3], d[4], d[5])) + b"\x1e"  # vhcurveto, 5 arguments
            elif nxt is not None and nxt[0] == "L" and k % 3 == 1:
                out += b"".join(num(a) for a in d + [nxt[1] - c[5], nxt[2] - c[6]]) + b"\x18"  # rcurveline
                x, y = nxt[1], nxt[2]
                i += 2
                continue
            elif nxt is not None and nxt[0] == "C" and k % 2 == 0:  # two general curves in one rrcurveto
                out += b"".join(num(a) for a in d + delta(nxt, c[5], c[6])) + b"\x08"
                x, y = nxt[5], nxt[6]
                i += 2
                continue
            else:
                out += b"".join(num(a) for a in d) + b"\x08"
            x, y = c[5], c[6]
            i += 1
    return out + b"".join(stack_w) + (b"\x0e" if endchar else b"")


def _blob(x, y, w, h, r):
    """Closed contour with cubic corners and axis-parallel edges, absolute integer coordinates."""
    kx = (r * 11) // 20
    return [
        ("M", x + r, y), ("L", x + w - r, y), ("C", x + w - r + kx, y, x + w, y + r - kx, x + w, y + r),
        ("L", x + w, y + h - r), ("C", x + w, y + h - r + kx, x + w - r + kx, y + h, x + w - r, y + h),
        ("L", x + r, y + h), ("C", x + r - kx, y + h, x, y + h - r + kx, x, y + h - r),
        ("L", x, y + r), ("C", x, y + r - kx, x + r - kx, y, x + r, y),
    ]


def _wave(x, y, w, h, rng):
    """Closed contour of general cubics and slanted lines (exercises rrcurveto / rlineto / flex / rcurveline)."""
    p = lambda fx, fy: (x + int(fx * w) + int(rng.integers(-9, 10)), y + int(fy * h) + int(rng.integers(-9, 10)))
    a, b, c, d, e, f, g = p(0, 0), p(0.3, 0.25), p(0.6, -0.2), p(1, 0.1), p(0.9, 0.6), p(0.5, 1.0), p(0.1, 0.8)
    m1, m2, m3, m4 = p(0.95, 0.3), p(1.05, 0.45), p(0.75, 0.9), p(0.6, 1.1)
    return [("M",) + a, ("C",) + b + c + d, ("C",) + m1 + m2 + e, ("L",) + m3, ("C",) + m4 + p(0.55, 1.05) + f,
            ("L",) + g, ("L",) + p(0.05, 0.5), ("L",) + a]


def _raw(*items):
    return b"".join(it if isinstance(it, bytes) else _t2_num(it) for it in items)


# hand-assembled charstrings for the three flex forms the encoder does not produce: (charstring, expected contours)
_FLEX_CASES = [
    (_raw(100, 100, b"\x15", 10, 20, 30, 40, 50, 60, 70, b"\x0c\x22", b"\x0e"),  # hflex
     [[("M", 100, 100), ("C", 110, 100, 130, 130, 170, 130), ("C", 220, 130, 280, 100, 350, 100)]]),
    (_raw(100, 100, b"\x15", 10, 5, 20, 30, 40, 50, 60, -30, 70, b"\x0c\x24", b"\x0e"),  # hflex1
     [[("M", 100, 100), ("C", 110, 105, 130, 135, 170, 135), ("C", 220, 135, 280, 105, 350, 100)]]),
    (_raw(100, 100, b"\x15", 10, 5, 20, 30, 40, 0, 50, 0, 60, -30, 70, b"\x0c\x25", b"\x0e"),  # flex1, horizontal
     [[("M", 100, 100), ("C", 110, 105, 130, 135, 170, 135), ("C", 220, 135, 280, 105, 350, 100)]]),
    (_raw(100, 100, b"\x15", 5, 10, 30, 20, 0, 40, 0, 50, -30, 60, 70, b"\x0c\x25", b"\x0e"),  # flex1, vertical
     [[("M", 100, 100), ("C", 105, 110, 135, 130, 135, 170), ("C", 135, 220, 105, 280, 100, 350)]]),
]


def _specials(x, y):
    """Contours built to reach the operators random shapes rarely hit: multi-argument hlineto / vlineto, hmoveto /
    vmoveto (consecutive contours share a coordinate), hhcurveto / vvcurveto with and without the leading argument,
    the 5-argument hvcurveto, chained hv / vh curves, and two general curves in a row (flex or 12-argument rrcurveto)."""
    stair_h = [("M", x, y), ("L", x + 60, y), ("L", x + 60, y + 50), ("L", x + 130, y + 50), ("L", x + 130, y + 120),
               ("L", x, y + 120), ("L", x, y)]
    bx = x + 300
    stair_v = [("M", bx, y), ("L", bx, y + 70), ("L", bx + 40, y + 70), ("L", bx + 40, y + 150), ("L", bx + 110, y + 150),
               ("L", bx + 110, y), ("L", bx, y)]
    cy = y + 200
    curvy = [("M", bx, cy), ("C", bx + 40, cy, bx + 60, cy + 60, bx + 100, cy + 60),
             ("C", bx + 100, cy + 100, bx + 130, cy + 120, bx + 130, cy + 160),
             ("C", bx + 80, cy + 190, bx + 20, cy + 210, bx - 20, cy + 210),
             ("C", bx - 40, cy + 170, bx - 70, cy + 120, bx - 70, cy + 60),
             ("C", bx - 50, cy + 60, bx - 30, cy + 30, bx, cy)]
    ox, oy = bx, cy + 300  # quarter arcs: hv, vh, hv, vh in one operator
    ring = [("M", ox + 50, oy), ("C", ox + 78, oy, ox + 100, oy + 22, ox + 100, oy + 50),
            ("C", ox + 100, oy + 78, ox + 78, oy + 100, ox + 50, oy + 100),
            ("C", ox + 22, oy + 100, ox, oy + 78, ox, oy + 50), ("C", ox, oy + 22, ox + 22, oy, ox + 50, oy)]
    lx, ly = x, y + 400
    lens = [("M", lx, ly), ("C", lx + 30, ly + 60, lx + 90, ly + 80, lx + 150, ly + 20),
            ("C", lx + 110, ly - 50, lx + 40, ly - 45, lx, ly)]
    return [stair_h, stair_v, curvy, ring, lens]


_SUBR_RING = [(20, 0), (0, 20), (-20, 0), (0, -20)]  # relative square drawn by the shared subroutines (returns to its start)


def cff_test_font(n_glyphs=40, first_cp=0x41, seed=5, cid=False, fd_select_format=3, charset_format=None, seac=False):
    """Returns (font bytes, code points, expected) where expected[i] = list of contours (absolute commands) of
    the glyph of cps[i], subroutine-drawn parts included."""
    rng = np.random.default_rng(seed)
    cps = list(range(first_cp, first_cp + n_glyphs + len(_FLEX_CASES) + (2 if seac else 0)))
    n_all = len(cps) + 1
    # subroutines: draw the small square relative to the current point, come back to it, open a new contour
    # there.  Global subr 0 nests local subr 0 (or, in a CID font, the glyph's own FD's local subr 0).
    def tri(scale):
        s = b""
        for dx, dy in _SUBR_RING:
            s += _t2_num(dx * scale) + _t2_num(dy * scale) + b"\x05"
        return s
    n_fd = 2 if cid else 1
    fd_of = lambda g: g * n_fd // n_all
    local_subrs = [[tri(1 + fd) + b"\x0b", tri(3 + fd) + _t2_num(0) + _t2_num(0) + b"\x15" + b"\x0b"] for fd in range(n_fd)]
    global_subrs = [_t2_num(0 - 107) + b"\x0a" + b"\x0b", tri(2) + b"\x0b"]

    def tri_contour(x, y, scale):
        pts, out = (x, y), [("M", x, y)]
        for dx, dy in _SUBR_RING:
            pts = (pts[0] + dx * scale, pts[1] + dy * scale)
            out.append(("L",) + pts)
        return out

    charstrings = [b"\x0e"]  # .notdef: endchar only -> no outline
    expected, advances = [], [500]
    for gi, cp in enumerate(cps[:n_glyphs]):
        contours = []
        for s in range(1 + gi % 3):
            w, h = int(rng.integers(120, 600)), int(rng.integers(100, 500))
            x, y = int(rng.integers(0, 1000 - w)), int(rng.integers(-200, 800 - h))
            contours.append(_blob(x, y, w, h, int(rng.integers(20, 50))) if (gi + s) % 2 == 0 else _wave(x, y, w, h, rng))
        if gi % 3 == 1:
            contours += _specials(int(rng.integers(0, 500)), int(rng.integers(-150, 100)))
        fd = fd_of(gi + 1)
        call = None
        kind = gi % 5
        first = contours[0][0]
        cmds = [c for ct in contours for c in ct]
        exp = [list(ct) for ct in contours]
        if kind == 1:  # local subr 0: triangle appended to the first contour's start, before its own segments
            call = (10, 0 - 107)
            exp[0] = tri_contour(first[1], first[2], 1 + fd) + contours[0][1:]
        elif kind == 2:  # global subr 0 -> local subr 0
            call = (29, 0 - 107)
            exp[0] = tri_contour(first[1], first[2], 1 + fd) + contours[0][1:]
        elif kind == 3:  # global subr 1
            call = (29, 1 - 107)
            exp[0] = tri_contour(first[1], first[2], 2) + contours[0][1:]
        elif kind == 4:  # local subr 1: a triangle contour of its own, then a fresh moveto at the same point
            call = (10, 1 - 107)
            exp = [tri_contour(first[1], first[2], 3 + fd), [("M", first[1] + 0, first[2] + 0)] + contours[0][1:]] + exp[1:]
        charstrings.append(encode_charstring(cmds, variant=gi, width=(300 + gi) if gi % 2 else None, hints=gi % 4 == 2,
                                             call=call))
        expected.append(exp)
        advances.append(int(rng.integers(400, 1101)))
    for code, exp in _FLEX_CASES:
        charstrings.append(code)
        expected.append(exp)
        advances.append(600)
    if seac:
        # `endchar` with four arguments (five with a width): accent glyph drawn at (adx, ady) over the base glyph, both
        # named by StandardEncoding codes that go code -> SID -> glyph through the charset.  Predefined ISOAdobe
        # charset: glyph id = SID = code - 31.  Custom charset (below): glyphs 1..10 have SIDs 30..39, the rest 108 + gid.
        assert not cid
        base_gid, accent_gid, base_code, accent_code = (34, 65, 65, 96) if charset_format is None else (5, 16, 65, 193)
        shift = lambda ct, dx, dy: [(c[0],) + tuple(v + (dx if k % 2 == 0 else dy) for k, v in enumerate(c[1:])) for c in ct]
        for args in ((450, 30, 40), (-25, 60)):
            charstrings.append(_raw(*args, base_code, accent_code, b"\x0e"))
            adx, ady = args[-2:]
            expected.append([list(ct) for ct in expected[base_gid - 1]] + [shift(ct, adx, ady) for ct in expected[accent_gid - 1]])
            advances.append(700)
    assert n_all == len(charstrings)

    # ---- CFF table: header, Name, Top DICT, String, Global Subrs | charset | FDSelect | CharStrings | FDArray |
    # Private DICTs + Local Subrs.  Offsets in DICTs use the fixed 5-byte form so sizes do not depend on them.
    header = bytes([1, 0, 4, 4])
    name_index = _cff_index([b"SynthCFF"])
    string_index = _cff_index([b"Adobe", b"Identity"]) if cid else _cff_index([])
    gsubr_index = _cff_index(global_subrs)
    font_matrix = b"".join(_dict_real(t) for t in ("0.001", "0", "0", "1E-3", "0", "0")) + b"\x0c\x07"

    def top_dict(o):
        d = b""
        if cid:
            d += _dict_int5(391) + _dict_int5(392) + _dict_int5(0) + b"\x0c\x1e"
        d += font_matrix
        if cid:
            d += _dict_int5(o["charset"]) + b"\x0f" + _dict_int5(o["fdarray"]) + b"\x0c\x24" + _dict_int5(o["fdselect"]) + b"\x0c\x25"
        else:
            if charset_format is not None:
                d += _dict_int5(o["charset"]) + b"\x0f"
            d += _dict_int5(o["priv_size"][0]) + _dict_int5(o["priv"][0]) + b"\x12"
        return d + _dict_int5(o["charstrings"]) + b"\x11"

    def private_dict(subrs_off):
        return _dict_real("0.039625") + b"\x0c\x09" + _dict_int5(500) + b"\x14" + _dict_int5(subrs_off) + b"\x13"

    priv_len = len(private_dict(0))
    zero = {"charset": 0, "fdarray": 0, "fdselect": 0, "charstrings": 0, "priv": [0] * n_fd, "priv_size": [priv_len] * n_fd}
    top_len = len(_cff_index([top_dict(zero)]))
    pos = len(header) + len(name_index) + top_len + len(string_index) + len(gsubr_index)
    o = dict(zero)
    charset = struct.pack(">BHH", 2, 1, n_all - 2) if cid else b""
    if not cid and charset_format is not None:
        sids = [29 + g if g <= 10 else 108 + g for g in range(1, n_all)]
        if charset_format == 0:
            charset = b"\0" + b"".join(struct.pack(">H", v) for v in sids)
        elif charset_format == 1:
            charset = b"\1" + struct.pack(">HB", 30, 9) + struct.pack(">HB", 119, n_all - 12)
        else:
            charset = b"\2" + struct.pack(">HH", 30, 9) + struct.pack(">HH", 119, n_all - 12)
    o["charset"] = pos
    pos += len(charset)
    if fd_select_format == 3:
        bounds = [g for g in range(n_all) if g == 0 or fd_of(g) != fd_of(g - 1)]
        fdselect = struct.pack(">BH", 3, len(bounds)) + b"".join(struct.pack(">HB", g, fd_of(g)) for g in bounds)
        fdselect += struct.pack(">H", n_all)
    else:
        fdselect = b"\0" + bytes(fd_of(g) for g in range(n_all))
    if not cid:
        fdselect = b""
    o["fdselect"] = pos
    pos += len(fdselect)
    cs_index = _cff_index(charstrings)
    o["charstrings"] = pos
    pos += len(cs_index)
    font_dict = lambda off: _dict_int5(priv_len) + _dict_int5(off) + b"\x12"
    fdarray = _cff_index([font_dict(0)] * n_fd) if cid else b""
    o["fdarray"] = pos
    pos += len(fdarray)
    privs = b""
    o["priv"] = []
    for fd in range(n_fd):
        o["priv"].append(pos)
        blob = private_dict(priv_len) + _cff_index(local_subrs[fd])  # Subrs right behind the Private DICT
        privs += blob
        pos += len(blob)
    if cid:
        fdarray = _cff_index([font_dict(off) for off in o["priv"]])
    cff = header + name_index + _cff_index([top_dict(o)]) + string_index + gsubr_index + charset + fdselect + cs_index + fdarray + privs
    assert len(_cff_index([top_dict(o)])) == top_len and len(cff) == pos

    # ---- sfnt wrapper ('OTTO'): cmap format 4 with one segment, hmtx, head, hhea, maxp 0.5, name
    sub = struct.pack(">HHHHHHH", 4, 16 + 16, 0, 4, 0, 0, 0) + struct.pack(">HH", cps[-1], 0xFFFF) + b"\0\0" + \
        struct.pack(">HH", cps[0], 0xFFFF) + struct.pack(">HH", (1 - cps[0]) & 0xFFFF, 1) + b"\0\0\0\0"
    cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 1, 12) + sub
    head = struct.pack(">IIIIHHqqhhhhHHhhh", 0x00010000, 0x00010000, 0, 0x5F0F3CF5, 0, 1000, 0, 0, 0, -200, 1000, 800,
                       0, 8, 2, 0, 0)
    hhea = struct.pack(">IhhhHhhhhhhhhhhhH", 0x00010000, 800, -200, 0, 1100, 0, 0, 1000, 1, 0, 0, 0, 0, 0, 0, 0, n_all)
    maxp = struct.pack(">IH", 0x00005000, n_all)
    hm = np.zeros((n_all, 2), dtype=">u2")
    hm[:, 0] = advances
    fam = ("Synth CFF CID" if cid else "Synth CFF").encode("utf-16-be")
    name = struct.pack(">HHH", 0, 1, 6 + 12) + struct.pack(">HHHHHH", 3, 1, 0x409, 1, len(fam), 0) + fam
    tables = {b"CFF ": cff, b"cmap": cmap_table, b"head": head, b"hhea": hhea, b"hmtx": hm.tobytes(), b"maxp": maxp,
              b"name": name}
    tags = sorted(tables)
    out = struct.pack(">4sHHHH", b"OTTO", len(tags), 64, 2, len(tags) * 16 - 64)
    offset = 12 + 16 * len(tags)
    records, blobs = b"", b""
    for tag in tags:
        data = tables[tag]
        records += tag + struct.pack(">III", _table_checksum(data), offset, len(data))
        padded = data + b"\0" * ((-len(data)) % 4)
        blobs += padded
        offset += len(padded)
    return out + records + blobs, cps, expected


def _cff2_index(items):
    """CFF 2 INDEX: a 32-bit count; the empty INDEX is the count alone."""
    if not items:
        return b"\0\0\0\0"
    offs = [1]
    for it in items:
        offs.append(offs[-1] + len(it))
    osz = 1 if offs[-1] < 256 else 2 if offs[-1] < 65536 else 4
    return struct.pack(">IB", len(items), osz) + b"".join(o.to_bytes(osz, "big") for o in offs) + b"".join(items)


def cff2_test_font(n_glyphs=30, first_cp=0x41, seed=11):
    """A variable OpenType font with CFF 2 outlines (two axes).  Returns (font bytes, code points, expected) like
    cff_test_font: expected[i] = contours of the DEFAULT instance — what a renderer that never sets variation coordinates
    draws.  Covers: 32-bit INDEX counts, charstrings without width / endchar, `blend` (one and several operands, one and
    two regions), `vsindex`, a region whose peaks are all 0 (its scalar is 1 at every coordinate, the default included),
    local and global subroutines without `return`, hint operators without a width."""
    rng = np.random.default_rng(seed)
    specials = [
        # (charstring, expected contours)
        (_raw(100, 5, 1, b"\x10", 200, -7, 1, b"\x10", b"\x15", 300, 9, 1, b"\x10", b"\x06", 0, 150, -40, 60, 3, 4, 5, 6, 4, b"\x10", b"\x05"),
         [[("M", 100, 200), ("L", 400, 200), ("L", 400, 350), ("L", 360, 410)]]),
        # vsindex 1: two regions, both 0 at the default
        (_raw(1, b"\x0f", 10, 20, 1, 2, 3, 4, 2, b"\x10", b"\x15", 250, 30, -30, 1, b"\x10", b"\x06", 250, b"\x07", -250, b"\x06"),
         [[("M", 10, 20), ("L", 260, 20), ("L", 260, 270), ("L", 10, 270)]]),
        # vsindex 2: a region with every peak 0 has scalar 1 -> its deltas apply at the default instance too
        (_raw(2, b"\x0f", 100, 11, 1, b"\x10", 200, 22, 1, b"\x10", b"\x15", 80, 5, 1, b"\x10", 0, 0, 90, -80, 0, b"\x05"),
         [[("M", 111, 222), ("L", 196, 222), ("L", 196, 312), ("L", 116, 312)]]),
        # hints without width: hstemhm (odd count: the last value is dropped), hintmask with implied vstems
        (_raw(10, 20, 30, b"\x12", 30, 40, 50, 60, b"\x13\xe0", 50, 60, b"\x15", 100, b"\x06", 100, b"\x07", -100, b"\x06"),
         [[("M", 50, 60), ("L", 150, 60), ("L", 150, 160), ("L", 50, 160)]]),
    ]
    cps = list(range(first_cp, first_cp + n_glyphs + len(specials)))
    n_all = len(cps) + 1

    def tri(scale):
        return b"".join(_t2_num(dx * scale) + _t2_num(dy * scale) + b"\x05" for dx, dy in _SUBR_RING)

    local_subrs = [tri(1), tri(3)]  # (no `return`: a CFF 2 subroutine ends with its data)
    global_subrs = [_t2_num(0 - 107) + b"\x0a", tri(2)]

    def tri_contour(x, y, scale):
        pts, out = (x, y), [("M", x, y)]
        for dx, dy in _SUBR_RING:
            pts = (pts[0] + dx * scale, pts[1] + dy * scale)
            out.append(("L",) + pts)
        return out

    charstrings = [b""]  # .notdef: nothing
    expected, advances = [], [500]
    for gi in range(n_glyphs):
        contours = []
        for k in range(1 + gi % 3):
            w, h = int(rng.integers(120, 600)), int(rng.integers(100, 500))
            x, y = int(rng.integers(0, 1000 - w)), int(rng.integers(-200, 800 - h))
            contours.append(_blob(x, y, w, h, int(rng.integers(20, 50))) if (gi + k) % 2 == 0 else _wave(x, y, w, h, rng))
        if gi % 3 == 1:
            contours += _specials(int(rng.integers(0, 500)), int(rng.integers(-150, 100)))
        first = contours[0][0]
        cmds = [c for ct in contours for c in ct]
        exp = [list(ct) for ct in contours]
        call = None
        kind = gi % 4
        if kind == 1:
            call = (10, 0 - 107)
            exp[0] = tri_contour(first[1], first[2], 1) + contours[0][1:]
        elif kind == 2:  # global subr 0 -> local subr 0
            call = (29, 0 - 107)
            exp[0] = tri_contour(first[1], first[2], 1) + contours[0][1:]
        elif kind == 3:
            call = (29, 1 - 107)
            exp[0] = tri_contour(first[1], first[2], 2) + contours[0][1:]
        charstrings.append(encode_charstring(cmds, variant=gi, call=call, endchar=False, blend_regions=1 if gi % 2 else 0))
        expected.append(exp)
        advances.append(int(rng.integers(400, 1101)))
    for code, exp in specials:
        charstrings.append(code)
        expected.append(exp)
        advances.append(600)
    assert n_all == len(charstrings)

    # ItemVariationStore: 3 regions over 2 axes; data 0 -> [region 0], data 1 -> [0, 1], data 2 -> [2]
    f2 = lambda v: int(round(v * 16384)) & 0xFFFF
    region = lambda *axes: b"".join(struct.pack(">HHH", f2(a), f2(b), f2(c)) for a, b, c in axes)
    regions = struct.pack(">HH", 2, 3) + region((0, 1, 1), (0, 0, 0)) + region((0, 0, 0), (0, 1, 1)) + region((0, 0, 0), (0, 0, 0))
    vdata = [struct.pack(">HHH", 0, 0, len(r)) + b"".join(struct.pack(">H", x) for x in r) for r in ([0], [0, 1], [2])]
    head_len = 8 + 4 * len(vdata)
    offs, pos = [], head_len + len(regions)
    for d in vdata:
        offs.append(pos)
        pos += len(d)
    store = struct.pack(">HIH", 1, head_len, len(vdata)) + b"".join(struct.pack(">I", o) for o in offs) + regions + b"".join(vdata)
    vstore = struct.pack(">H", len(store)) + store

    gsubr_index = _cff2_index(global_subrs)
    cs_index = _cff2_index(charstrings)
    private_dict = lambda subrs_off: _dict_int5(500) + b"\x14" + _dict_int5(subrs_off) + b"\x13"
    priv_len = len(private_dict(0))
    font_dict = lambda off: _dict_int5(priv_len) + _dict_int5(off) + b"\x12"
    top_dict = lambda o: _dict_int5(o["charstrings"]) + b"\x11" + _dict_int5(o["vstore"]) + b"\x18" + _dict_int5(o["fdarray"]) + b"\x0c\x24"
    zero = {"charstrings": 0, "vstore": 0, "fdarray": 0}
    top_len = len(top_dict(zero))
    pos = 5 + top_len + len(gsubr_index)
    o = {}
    o["vstore"] = pos
    pos += len(vstore)
    o["charstrings"] = pos
    pos += len(cs_index)
    o["fdarray"] = pos
    fdarray_len = len(_cff2_index([font_dict(0)]))
    pos += fdarray_len
    priv_off = pos
    fdarray = _cff2_index([font_dict(priv_off)])
    assert len(fdarray) == fdarray_len
    privs = private_dict(priv_len) + _cff2_index(local_subrs)
    cff2 = struct.pack(">BBBH", 2, 0, 5, top_len) + top_dict(o) + gsubr_index + vstore + cs_index + fdarray + privs

    fixed = lambda v: struct.pack(">i", int(round(v * 65536)))
    axis = lambda tag, lo, de, hi, name_id: tag + fixed(lo) + fixed(de) + fixed(hi) + struct.pack(">HH", 0, name_id)
    fvar = struct.pack(">HHHHHHHH", 1, 0, 16, 2, 2, 20, 0, 4 + 2 * 4) + axis(b"wght", 100, 400, 900, 256) + axis(b"wdth", 50, 100, 200, 257)

    sub = struct.pack(">HHHHHHH", 4, 16 + 16, 0, 4, 0, 0, 0) + struct.pack(">HH", cps[-1], 0xFFFF) + b"\0\0" + \
        struct.pack(">HH", cps[0], 0xFFFF) + struct.pack(">HH", (1 - cps[0]) & 0xFFFF, 1) + b"\0\0\0\0"
    cmap_table = struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 1, 12) + sub
    head = struct.pack(">IIIIHHqqhhhhHHhhh", 0x00010000, 0x00010000, 0, 0x5F0F3CF5, 0, 1000, 0, 0, 0, -200, 1000, 800,
                       0, 8, 2, 0, 0)
    hhea = struct.pack(">IhhhHhhhhhhhhhhhH", 0x00010000, 800, -200, 0, 1100, 0, 0, 1000, 1, 0, 0, 0, 0, 0, 0, 0, n_all)
    maxp = struct.pack(">IH", 0x00005000, n_all)
    hm = np.zeros((n_all, 2), dtype=">u2")
    hm[:, 0] = advances
    names = [(1, "Synth CFF2"), (256, "Weight"), (257, "Width")]
    strings = [t.encode("utf-16-be") for _, t in names]
    recs, off = b"", 0
    for (nid, _), st in zip(names, strings):
        recs += struct.pack(">HHHHHH", 3, 1, 0x409, nid, len(st), off)
        off += len(st)
    name = struct.pack(">HHH", 0, len(names), 6 + 12 * len(names)) + recs + b"".join(strings)
    tables = {b"CFF2": cff2, b"cmap": cmap_table, b"fvar": fvar, b"head": head, b"hhea": hhea, b"hmtx": hm.tobytes(),
              b"maxp": maxp, b"name": name}
    tags = sorted(tables)
    out = struct.pack(">4sHHHH", b"OTTO", len(tags), 128, 3, len(tags) * 16 - 128)
    offset = 12 + 16 * len(tags)
    records, blobs = b"", b""
    for tag in tags:
        data = tables[tag]
        records += tag + struct.pack(">III", _table_checksum(data), offset, len(data))
        padded = data + b"\0" * ((-len(data)) % 4)
        blobs += padded
        offset += len(padded)
    return out + records + blobs, cps, expected


def replace_table(blob: bytes, tag: bytes, new: bytes) -> bytes:
    """The sfnt `blob` with table `tag` replaced by `new` (directory and offsets rebuilt, checksums recomputed)."""
    n = int.from_bytes(blob[4:6], "big")
    tables = {}
    for i in range(n):
        rec = 12 + 16 * i
        t = blob[rec : rec + 4]
        off, length = int.from_bytes(blob[rec + 8 : rec + 12], "big"), int.from_bytes(blob[rec + 12 : rec + 16], "big")
        tables[t] = blob[off : off + length]
    assert tag in tables
    tables[tag] = new
    tags = sorted(tables)
    out = blob[:12]
    offset = 12 + 16 * len(tags)
    records, blobs = b"", b""
    for t in tags:
        data = tables[t]
        records += t + struct.pack(">III", _table_checksum(data), offset, len(data))
        padded = data + b"\0" * ((-len(data)) % 4)
        blobs += padded
        offset += len(padded)
    return out + records + blobs


def with_variation_selectors(blob: bytes, sequences) -> bytes:
    """The font with a cmap format 14 subtable (platform 0, encoding 5) put IN FRONT of its existing subtable:
    sequences = [(variation selector, [(base code point, glyph id)])] as non-default UVS mappings."""
    n = int.from_bytes(blob[4:6], "big")
    rec = next(12 + 16 * i for i in range(n) if blob[12 + 16 * i : 16 + 16 * i] == b"cmap")
    off, length = int.from_bytes(blob[rec + 8 : rec + 12], "big"), int.from_bytes(blob[rec + 12 : rec + 16], "big")
    cmap = blob[off : off + length]
    assert int.from_bytes(cmap[2:4], "big") == 1
    platform, encoding, sub_off = struct.unpack(">HHI", cmap[4:12])
    old_sub = cmap[sub_off:]
    header_len = 10 + 11 * len(sequences)
    tables, pos = b"", header_len
    records = b""
    for vs, maps in sequences:
        t = struct.pack(">I", len(maps)) + b"".join(cp.to_bytes(3, "big") + struct.pack(">H", g) for cp, g in maps)
        records += vs.to_bytes(3, "big") + struct.pack(">II", 0, pos)
        tables += t
        pos += len(t)
    f14 = struct.pack(">HII", 14, header_len + len(tables), len(sequences)) + records + tables
    new = struct.pack(">HH", 0, 2) + struct.pack(">HHI", 0, 5, 20) + struct.pack(">HHI", platform, encoding, 20 + len(f14)) + f14 + old_sub
    return replace_table(blob, b"cmap", new)


def with_cmap_format2(blob: bytes, single, double) -> bytes:
    """The font with its cmap replaced by ONE format 2 subtable (platform 0, encoding 3).  single = (first code, [glyph
    ids]) for one-byte codes; double = [(lead byte, first low byte, id delta, [stored glyph ids])]."""
    subs = [(single[0], 0, list(single[1]))] + [(lo, d, list(g)) for _, lo, d, g in double]
    keys = [0] * 256
    for k, (lead, _, _, _) in enumerate(double):
        keys[lead] = 8 * (k + 1)
    n_sub = len(subs)
    array_start = 518 + 8 * n_sub
    headers, array = b"", []
    for i, (first, delta, glyphs) in enumerate(subs):
        field = 518 + 8 * i + 6
        headers += struct.pack(">HHhH", first, len(glyphs), delta, array_start + 2 * len(array) - field)
        array += glyphs
    body = b"".join(struct.pack(">H", k) for k in keys) + headers + b"".join(struct.pack(">H", g) for g in array)
    sub = struct.pack(">HHH", 2, 6 + len(body), 0) + body
    return replace_table(blob, b"cmap", struct.pack(">HH", 0, 1) + struct.pack(">HHI", 0, 3, 12) + sub)
