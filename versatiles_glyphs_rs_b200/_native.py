"""ctypes bindings of the two in-tree libraries.

libb200sdf.so     — CUDA kernels + C ABI (include/b200sdf.h): the drop-in boundary
libvgb200host.so  — C++ host mirror of the reference's Rust host (include/vgb200_host.h)

Both are built in-tree by ``__graft_entry__.build()`` (or ``make -C versatiles_glyphs_rs_b200/csrc``).
Loading fails loudly when they are missing: there is no Python or CPU fallback for the SDF path.
"""
import ctypes as C
import os

_HERE = os.environ.get("VGB200_LIBDIR") or os.path.dirname(os.path.abspath(__file__))  # override: kernel-variant experiments
SDF_LIB_PATH = os.path.join(_HERE, "libb200sdf.so")
HOST_LIB_PATH = os.path.join(_HERE, "libvgb200host.so")


class NativeLibraryMissing(ImportError):
    pass


def _load(path):
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the SDF path."
        )
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


sdf = _load(SDF_LIB_PATH)
host = _load(HOST_LIB_PATH)

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class Segment(C.Structure):
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("x1", C.c_float), ("y1", C.c_float)]


class GlyphJob(C.Structure):
    _fields_ = [
        ("seg_off", C.c_uint32),
        ("seg_cnt", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("out_off", C.c_uint64),
    ]


class TileJob(C.Structure):
    _fields_ = [
        ("seg_off", C.c_uint32),
        ("seg_cnt", C.c_uint32),
        ("out_off", C.c_uint64),
        ("width", C.c_uint16),
        ("height", C.c_uint16),
        ("tx0", C.c_uint16),
        ("ty0", C.c_uint16),
        ("ntx", C.c_uint16),
        ("nty", C.c_uint16),
        ("job", C.c_uint32),
    ]


class Curve(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("sx", "sy", "cx", "cy", "ex", "ey")] + [("seg_off", C.c_uint32), ("depth", C.c_uint32)]


class OutlineJob(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32),
        ("src_off", C.c_uint32),
        ("src_cnt", C.c_uint32),
        ("seg_cnt", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("x0", C.c_int32),
        ("y0", C.c_int32),
        ("scale", C.c_double),
        ("dx", C.c_double),
        ("out_off", C.c_uint64),
    ]


KIND_CURVES, KIND_SEGMENTS = 0, 1


class Glyph(C.Structure):
    _fields_ = [
        ("id", C.c_uint32),
        ("has_bitmap", C.c_int32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("left", C.c_int32),
        ("top", C.c_int32),
        ("advance", C.c_uint32),
        ("bitmap", u8p),
        ("bitmap_len", C.c_uint64),
        ("n_segments", C.c_uint32),
    ]


class BatchGlyph(C.Structure):
    _fields_ = [
        ("id", C.c_uint32),
        ("advance", C.c_uint32),
        ("has_bitmap", C.c_int32),
        ("x0", C.c_int32),
        ("y0", C.c_int32),
        ("bm_width", C.c_uint32),
        ("bm_height", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("left", C.c_int32),
        ("top", C.c_int32),
        ("kind", C.c_uint32),
        ("src_off", C.c_uint32),
        ("src_cnt", C.c_uint32),
        ("seg_cnt", C.c_uint32),
        ("out_off", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "glyphs", "bitmaps", "pixels", "segments", "pairs", "pbf_bytes", "blocks",
        "outline_ns", "submit_ns", "wait_ns", "encode_ns", "write_ns", "wall_ns", "submits", "workers",
        "handed_back", "h2d_bytes", "cost_total", "cost_shard")]


class GlyphPart(C.Structure):
    _fields_ = [("font", C.c_uint32), ("glyf_off", C.c_uint32), ("glyf_len", C.c_uint32), ("ox", C.c_float), ("oy", C.c_float)]


class GlyphReq(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("src_off", C.c_uint32), ("src_cnt", C.c_uint32), ("seg_cnt", C.c_uint32),
        ("width", C.c_uint32), ("height", C.c_uint32), ("x0", C.c_int32), ("y0", C.c_int32),
        ("scale", C.c_double), ("dx", C.c_double), ("out_off", C.c_uint64),
        ("out_cap", C.c_uint32), ("curve_off", C.c_uint32), ("curve_cap", C.c_uint32), ("reserved", C.c_uint32),
    ]


class GlyphBatchDesc(C.Structure):
    """b200sdf_glyph_batch: one batch of a multi-batch submission"""
    _fields_ = [("reqs", C.c_void_p), ("n_reqs", C.c_uint32), ("parts", C.c_void_p), ("n_parts", C.c_uint32),
                ("curves", C.c_void_p), ("n_curves", C.c_uint32), ("segs", C.c_void_p), ("n_seg", C.c_uint32),
                ("curve_slots", C.c_uint32), ("tile_cap", C.c_uint32), ("frames", C.c_void_p), ("out", C.c_void_p),
                ("out_bytes", C.c_uint64)]


class GlyphFrame(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("seg_cnt", C.c_uint32), ("status", C.c_uint32)]


KIND_GLYF = 2
KIND_PATH = 3  # host-recorded outline with cubic curves, flattened by the decode kernel
CURVE_CUBIC, CURVE_TAIL = 0x80000000, 0x40000000
GLYPH_OK, GLYPH_EMPTY, GLYPH_NEEDS_HOST, GLYPH_BAD_REQUEST = 0, 1, 2, 3


# name -> (restype, argtypes); the single source of truth for "every symbol the headers declare"
SDF_SYMBOLS = {
    "b200sdf_abi_version": (C.c_int, []),
    "b200sdf_device_count": (C.c_int, []),
    "b200sdf_create": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]),
    "b200sdf_destroy": (None, [C.c_void_p]),
    "b200sdf_last_error": (C.c_char_p, [C.c_void_p]),
    "b200sdf_device": (C.c_int, [C.c_void_p]),
    "b200sdf_reserve": (C.c_int, [C.c_void_p]),
    "b200sdf_reserve_glyphs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "b200sdf_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "b200sdf_free_pinned": (None, [C.c_void_p]),
    "b200sdf_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, u64p]),
    "b200sdf_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200sdf_poll": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200sdf_submit_planned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                         C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, u64p]),
    "b200sdf_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "b200sdf_plan_tiles": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, u32p, u64p]),
    "b200sdf_render_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "b200sdf_submit_outlines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, u64p]),
    "b200sdf_flatten_outlines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "b200sdf_plan_outline_tiles": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, u32p, u64p]),
    "b200sdf_plan_outline_tiles_ex": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p,
                                                C.c_uint32, u32p, u64p]),
    "b200sdf_render_outlines_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "b200sdf_font_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u32p]),
    "b200sdf_glyph_tile_bound": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "b200sdf_submit_glyphs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "b200sdf_submit_glyph_batches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, u64p]),
    "b200sdf_render_glyphs_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                               C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                               C.c_void_p, C.c_void_p]),
    "b200sdf_decode_glyphs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, u32p, C.c_void_p, C.c_uint32]),
    "b200sdf_measure_fp32_peak": (C.c_int, [C.c_void_p, C.c_int, f64p, f64p]),
    "b200sdf_launch_count": (C.c_uint64, [C.c_void_p]),
}

HOST_SYMBOLS = {
    "vgb_last_error": (C.c_char_p, []),
    "vgb_free": (None, [C.c_void_p]),
    "vgb_font_from_bytes": (C.c_void_p, [C.c_char_p, C.c_size_t]),
    "vgb_font_from_path": (C.c_void_p, [C.c_char_p]),
    "vgb_font_free": (None, [C.c_void_p]),
    "vgb_font_units_per_em": (C.c_uint32, [C.c_void_p]),
    "vgb_font_number_of_glyphs": (C.c_uint32, [C.c_void_p]),
    "vgb_font_glyph_index": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "vgb_font_hor_advance": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "vgb_font_codepoints": (C.c_size_t, [C.c_void_p, u32p, C.c_size_t]),
    "vgb_font_outline_rings": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(f64p), C.POINTER(u32p), u32p]),
    "vgb_font_outline_commands": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_float))]),
    "vgb_flatten_quad": (C.c_size_t, [f64p, f64p, f64p, C.c_double, f64p, C.c_size_t]),
    "vgb_flatten_cubic": (C.c_size_t, [f64p, f64p, f64p, f64p, C.c_double, f64p, C.c_size_t]),
    "vgb_segment_sqdist": (C.c_double, [C.c_double] * 6),
    "vgb_name_to_id": (C.c_char_p, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "vgb_parse_font_name": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint16), C.c_char_p, C.c_size_t]),
    "vgb_encode_codeblocks": (C.c_size_t, [u32p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "vgb_font_metadata": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint16), C.c_char_p, C.c_char_p, C.c_size_t]),
    "vgb_renderer_new": (C.c_void_p, [C.c_int, C.c_int, C.c_uint32]),
    "vgb_renderer_free": (None, [C.c_void_p]),
    "vgb_renderer_is_dummy": (C.c_int, [C.c_void_p]),
    "vgb_renderer_context": (C.c_void_p, [C.c_void_p]),
    "vgb_renderer_render_glyph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Glyph)]),
    "vgb_batch_new": (C.c_void_p, [C.c_void_p]),
    "vgb_batch_free": (None, [C.c_void_p]),
    "vgb_batch_clear": (None, [C.c_void_p]),
    "vgb_batch_add_glyph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "vgb_batch_add_rings": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, f64p, u32p, C.c_uint32]),
    "vgb_batch_glyph_count": (C.c_uint32, [C.c_void_p]),
    "vgb_batch_glyph_info": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(BatchGlyph)]),
    "vgb_batch_segments": (C.POINTER(Segment), [C.c_void_p, u32p]),
    "vgb_batch_jobs": (C.POINTER(OutlineJob), [C.c_void_p, u32p]),
    "vgb_batch_curves": (C.POINTER(Curve), [C.c_void_p, u32p]),
    "vgb_batch_total_segments": (C.c_uint64, [C.c_void_p]),
    "vgb_batch_fallback_glyphs": (C.c_uint32, [C.c_void_p]),
    "vgb_renderer_set_flatten": (None, [C.c_void_p, C.c_int]),
    "vgb_renderer_flatten": (C.c_int, [C.c_void_p]),
    "vgb_batch_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgb_batch_requests": (C.POINTER(GlyphReq), [C.c_void_p, u32p]),
    "vgb_batch_parts": (C.POINTER(GlyphPart), [C.c_void_p, u32p]),
    "vgb_batch_curve_slots": (C.c_uint32, [C.c_void_p]),
    "vgb_batch_tile_cap": (C.c_uint32, [C.c_void_p]),
    "vgb_batch_est_cost": (C.c_uint64, [C.c_void_p]),
    "vgb_batch_handed_back": (C.c_uint32, [C.c_void_p]),
    "vgb_batch_path_glyphs": (C.c_uint32, [C.c_void_p]),
    "vgb_batch_glyph_bitmap": (u8p, [C.c_void_p, C.c_uint32, u64p]),
    "vgb_batch_bitmaps": (u8p, [C.c_void_p, u64p]),
    "vgb_batch_pairs": (C.c_uint64, [C.c_void_p]),
    "vgb_renderer_render_batch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgb_renderer_submit_batch": (C.c_int, [C.c_void_p, C.c_void_p, u64p]),
    "vgb_renderer_wait_batch": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vgb_renderer_prepare_batch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgb_renderer_poll_batch": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vgb_writer_new_file": (C.c_void_p, [C.c_char_p]),
    "vgb_writer_new_memory": (C.c_void_p, []),
    "vgb_writer_new_tar": (C.c_void_p, [C.c_char_p]),
    "vgb_writer_new_tar_memory": (C.c_void_p, []),
    "vgb_writer_write_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64]),
    "vgb_writer_write_directory": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vgb_writer_finish": (C.c_int, [C.c_void_p]),
    "vgb_writer_tar_bytes": (u8p, [C.c_void_p, u64p]),
    "vgb_writer_free": (None, [C.c_void_p]),
    "vgb_writer_entry_count": (C.c_uint32, [C.c_void_p]),
    "vgb_writer_entry": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(u8p), u64p]),
    "vgb_manager_new": (C.c_void_p, [C.c_int]),
    "vgb_manager_free": (None, [C.c_void_p]),
    "vgb_manager_add_path": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vgb_manager_scan": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vgb_manager_font_file_names": (C.c_size_t, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "vgb_manager_add_font_with_name": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32]),
    "vgb_manager_add_font_bytes_with_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "vgb_manager_font_count": (C.c_uint32, [C.c_void_p]),
    "vgb_manager_font_id": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "vgb_manager_block_population": (C.c_int, [C.c_void_p, C.c_char_p, u32p]),
    "vgb_manager_render_block": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.POINTER(u8p), u64p]),
    "vgb_manager_render_glyphs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(Stats)]),
    "vgb_manager_shard_owners": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint16), C.c_size_t, u64p]),
    "vgb_manager_write_index_json": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgb_manager_write_families_json": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgb_pbf_decode": (C.c_int32, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(Glyph))]),
    "vgb_glyphs_free": (None, [C.POINTER(Glyph), C.c_int32]),
}

for _lib, _table in ((sdf, SDF_SYMBOLS), (host, HOST_SYMBOLS)):
    for _name, (_res, _args) in _table.items():
        _fn = getattr(_lib, _name)  # AttributeError here = header/library mismatch: fail loudly
        _fn.restype = _res
        _fn.argtypes = _args


def host_error():
    return (host.vgb_last_error() or b"").decode("utf-8", "replace")
