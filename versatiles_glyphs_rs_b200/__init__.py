"""versatiles_glyphs_rs_b200 — B200-native drop-in for the per-glyph SDF rendering path of
versatiles_glyphs (reference v0.9.1).  See DESIGN.md.

The heavy lifting is native: ``libb200sdf.so`` (hand-written sm_100a CUDA behind the C ABI of
include/b200sdf.h) and ``libvgb200host.so`` (C++ host mirror of the reference's Rust host).
Importing this package without those libraries raises: there is no CPU or Python fallback.
"""
from .api import (  # noqa: F401
    BUFFER,
    GLYPH_BLOCK_SIZE,
    GLYPH_SIZE,
    B200Error,
    FontFileEntry,
    FontManager,
    GlyphBatch,
    PbfGlyph,
    Renderer,
    RenderStats,
    SdfContext,
    Writer,
    decode_pbf,
    device_count,
    encode_codeblocks,
    name_to_id,
    parse_font_name,
)
