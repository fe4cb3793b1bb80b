"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §4 / §8c).  CPU-only.  Each test cites the reference test it restates."""
import math

import numpy as np
import pytest

import oracle_lib as O


@pytest.fixture(scope="module")
def fira():
    return O.Font(O.FIRA)


# ---- src/render/renderer_precise.rs:91-135 test_render_sdf_simple_square ----
def test_render_sdf_simple_square():
    bm = O.renderer_precise(-2, -1, 10, 10, [[(1, 2), (5, 2), (5, 6), (1, 6), (1, 2)]])
    assert O.bitmap_as_digit_art(bm, 10) == [
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


# ---- src/render/renderer.rs:176-185 test_render_glyph_32 ----
def test_render_glyph_32(fira):
    g = fira.render_glyph(32)
    assert (g["width"], g["height"], g["left"], g["top"], g["advance"]) == (0, 0, 0, 0, 6)
    assert g["bitmap"] is None


def _check(g, w, h, left, top, adv):
    assert (g["width"], g["height"], g["left"], g["top"], g["advance"]) == (w, h, left, top, adv)
    assert len(g["bitmap"]) == (g["width"] + 6) * (g["height"] + 6)  # renderer.rs:164-166
    return O.bitmap_as_ascii_art(g["bitmap"], g["width"] + 6)


# ---- src/render/renderer.rs:188-224 test_render_glyph_65 ----
def test_render_glyph_65(fira):
    art = _check(fira.render_glyph(65), 14, 17, 0, -7, 13)
    assert art == [
        "            ░░░░░░░░░░░░░░░░            ",
        "          ░░░░▒▒▒▒▒▒▒▒▒▒░░░░░░          ",
        "        ░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░          ",
        "        ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒░░░░        ",
        "        ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒░░░░        ",
        "      ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░        ",
        "      ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░      ",
        "      ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░      ",
        "      ░░▒▒▒▒▓▓▓▓▓▓▒▒▓▓▓▓▓▓▒▒▒▒░░░░      ",
        "    ░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░    ",
        "    ░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░    ",
        "    ░░░░▒▒▓▓▓▓▓▓▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░    ",
        "  ░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒░░░░    ",
        "  ░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░  ",
        "  ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░  ",
        "░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░  ",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░",
        "░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░",
        "░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░░░░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░",
        "░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░",
        "░░▒▒▒▒▒▒▒▒▒▒▒▒░░░░  ░░░░░░▒▒▒▒▒▒▒▒▒▒░░░░",
        "░░░░░░░░░░░░░░░░░░    ░░░░░░░░░░░░░░░░░░",
    ]


# ---- src/render/renderer.rs:227-260 test_render_glyph_230 ----
def test_render_glyph_230(fira):
    art = _check(fira.render_glyph(230), 19, 14, 0, -11, 19)
    assert art == [
        "      ░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░      ",
        "    ░░░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░▒▒▒▒▒▒▒▒▒▒▒▒░░░░░░    ",
        "  ░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░  ",
        "  ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░░░",
        "  ░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░░░",
        "  ░░░░▒▒▒▒▓▓▓▓▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▒▒▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "  ░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░",
        "  ░░░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒░░",
        "  ░░░░▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "░░░░▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒▒▒▓▓▓▓▓▓▓▓▒▒▒▒▒▒▒▒▒▒▓▓▒▒▒▒░░░░",
        "░░░░▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "░░░░▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▓▒▒▒▒▒▒░░",
        "  ░░░░▒▒▒▒▒▒▒▒▓▓▓▓▒▒▒▒▒▒▒▒▒▒▒▒▒▒▓▓▓▓▒▒▒▒▒▒▒▒▒▒░░░░",
        "    ░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░░░  ",
        "      ░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░░    ",
        "        ░░░░░░░░░░░░░░░░  ░░░░░░░░░░░░░░░░        ",
    ]


# ---- src/render/renderer.rs:263-287 test_render_glyph_96 ----
def test_render_glyph_96(fira):
    art = _check(fira.render_glyph(96), 7, 5, 0, -4, 7)
    assert art == [
        "    ░░░░░░░░░░            ",
        "  ░░░░░░░░░░░░░░░░        ",
        "  ░░░░▒▒▒▒▒▒▒▒░░░░░░░░    ",
        "░░░░▒▒▒▒▒▒▒▒▒▒▒▒▒▒░░░░░░  ",
        "░░░░▒▒▒▒▓▓▓▓▓▓▒▒▒▒▒▒░░░░░░",
        "░░░░▒▒▓▓▓▓▓▓▓▓▓▓▒▒▒▒▒▒▒▒░░",
        "░░░░▒▒▒▒▒▒▓▓▓▓▓▓▓▓▓▓▒▒▒▒░░",
        "░░░░░░▒▒▒▒▒▒▒▒▒▒▓▓▒▒▒▒▒▒░░",
        "  ░░░░░░░░▒▒▒▒▒▒▒▒▒▒▒▒░░░░",
        "      ░░░░░░░░▒▒▒▒▒▒░░░░░░",
        "          ░░░░░░░░░░░░░░  ",
    ]


# ---- src/render/ring_builder.rs:197-229: quad and cubic flatten to exactly 17 points ----
def test_flatten_point_counts():
    q = O.flatten_quad((0, 0), (10, 10), (20, 0))
    assert len(q) + 1 == 17  # + the move_to point
    assert tuple(q[-1]) == (20.0, 0.0)
    c = O.flatten_cubic((0, 0), (10, 10), (20, 10), (30, 0))
    assert len(c) + 1 == 17
    assert tuple(c[-1]) == (30.0, 0.0)
    # uniform 2^k rule for quadratics: points at t = i/16
    t = np.arange(1, 17) / 16.0
    bx = 2 * (1 - t) * t * 10 + t * t * 20
    by = 2 * (1 - t) * t * 10
    assert np.allclose(q, np.stack([bx, by], 1), atol=1e-12)


# ---- src/geometry/ring.rs tests: a straight "curve" emits just the end point ----
def test_flatten_degenerate():
    assert len(O.flatten_quad((0, 0), (5, 0), (10, 0))) == 1
    assert len(O.flatten_cubic((0, 0), (1, 0), (2, 0), (3, 0))) >= 1


# ---- src/commands/recurse.rs:341-367 (and merge.rs:158-185): PBF byte sizes of all Fira blocks ----
FIRA_PBF_SIZES = {
    0: 80022, 1024: 118037, 11264: 3579, 1280: 26296, 256: 130750, 3584: 592, 42752: 5761, 43776: 487,
    512: 92634, 64256: 1032, 65024: 50, 7424: 7260, 768: 63760, 7680: 87078, 7936: 124520, 8192: 20301,
    8448: 17395, 8704: 6511, 8960: 4375, 9472: 853,
}
# ---- src/font/wrapper.rs:197-220 test_get_blocks ----
FIRA_POPULATION = [(0, 192), (256, 256), (512, 219), (768, 177), (1024, 240), (1280, 48), (3584, 1), (7424, 20),
                   (7680, 157), (7936, 233), (8192, 67), (8448, 28), (8704, 16), (8960, 5), (9472, 2), (11264, 7),
                   (42752, 14), (43776, 1), (64256, 2), (65024, 1)]


@pytest.fixture(scope="module")
def fira_set():
    return O.FontSet("Fira Sans Regular", [O.FIRA])


def test_fira_block_population(fira_set):
    assert fira_set.id == "fira_sans_regular"
    pop = fira_set.block_population()
    assert len(pop) == 256
    assert [(256 * i, n) for i, n in enumerate(pop) if n] == FIRA_POPULATION


def test_fira_pbf_sizes_dummy(fira_set):
    """The reference's size goldens use the dummy renderer (identical sizes by construction)."""
    for start, size in FIRA_PBF_SIZES.items():
        assert len(fira_set.render_block(start // 256, O.MODE_DUMMY)) == size, start
    # empty blocks: 2 + 19 + 2 + len(range) bytes; tests filter 32..34 (recurse.rs:204-206)
    assert len(fira_set.render_block(3840 // 256, O.MODE_DUMMY)) == 2 + 19 + 2 + len("3840-4095")


def test_fira_pbf_sizes_precise_small_blocks(fira_set):
    """Same sizes from the precise renderer on the small blocks (the full set runs in test_oracle_full)."""
    for start in (3584, 43776, 64256, 65024, 9472):
        data = fira_set.render_block(start // 256, O.MODE_PRECISE)
        assert len(data) == FIRA_PBF_SIZES[start]
        name, rng, glyphs = O.decode_pbf(data)
        assert name == "fira_sans_regular" and rng == "%d-%d" % (start, start + 255)
        assert [g["id"] for g in glyphs] == sorted(g["id"] for g in glyphs)


# ---- src/font/metadata.rs:136-153, src/font/file_entry.rs:66-71 ----
def test_codepoint_counts(fira):
    assert len(fira.codepoints()) == 1686
    assert fira.number_of_glyphs == 2677
    noto = O.Font(O.noto_paths()[0])
    assert O.noto_paths()[0].endswith("Noto Sans - Regular.ttf")
    assert len(noto.codepoints()) == 3094


# ---- src/render/rtree_segments.rs:92-198 ----
def test_min_distance_helper():
    assert abs(O.min_distance([(0, 0, 4, 0)], (2, 1), 5.0) - 1.0) < 2.3e-16
    assert math.isinf(O.min_distance([(0, 0, 4, 0)], (100, 100), 5.0))
    segs = [(0, 0, 4, 0), (2, 2, 2, 6), (-1, -1, -1, -5)]
    assert abs(O.min_distance(segs, (2, 1), 5.0) - 1.0) < 2.3e-16
    assert O.min_distance(segs, (-1, -3), 5.0) == 0.0
    assert O.min_distance([(1, 1, 5, 1)], (3, 1), 2.0) == 0.0


def test_rtree_matches_linear_scan():
    """The STR tree stands in for rstar: same candidate set semantics as a linear AABB scan."""
    rng = np.random.default_rng(7)
    segs = rng.uniform(-10, 40, size=(500, 4))
    for _ in range(200):
        p = rng.uniform(-20, 50, size=2)
        lo, hi = p - 8.0, p + 8.0
        best = math.inf
        for s in segs:
            if min(s[0], s[2]) > hi[0] or max(s[0], s[2]) < lo[0] or min(s[1], s[3]) > hi[1] or max(s[1], s[3]) < lo[1]:
                continue
            best = min(best, O.segment_sqdist(s[:2], s[2:], p))
        assert O.min_distance(segs, p, 8.0) == math.sqrt(best)


# ---- src/geometry/segment.rs:117-198 ----
def test_segment_projection_cases():
    assert O.segment_sqdist((0, 0), (10, 0), (5, 3)) == 9.0      # interior
    assert O.segment_sqdist((0, 0), (10, 0), (-3, 4)) == 25.0    # clamps to start
    assert O.segment_sqdist((0, 0), (10, 0), (13, 4)) == 25.0    # clamps to end
    assert O.segment_sqdist((2, 2), (2, 2), (5, 6)) == 25.0      # zero-length segment
    assert O.segment_sqdist((0, 0), (10, 10), (5, 5)) == 0.0     # on the segment


# ---- src/font/manager.rs name_to_id tests ----
def test_name_to_id():
    import ctypes as C
    buf = C.create_string_buffer(128)
    f = lambda s: O.lib().vgo_name_to_id(s.encode(), buf, 128).decode()
    assert f("Fira Sans Regular") == "fira_sans_regular"
    assert f("  Noto--Sans__Regular \t") == "noto_sans_regular"
    assert f("Open Sans - Bold_Italic") == "open_sans_bold_italic"


# ---- src/render/renderer.rs:104: surrogates are skipped; unmapped code points are None ----
def test_render_glyph_none_cases(fira):
    assert fira.render_glyph(0xD800) is None
    assert fira.render_glyph(0x4E00) is None  # CJK not in Fira
