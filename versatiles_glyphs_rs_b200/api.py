"""Python mirror of the reference's public Rust API for the rendering path
(reference src/lib.rs:8-13: FontManager, FontWrapper, GlyphBlock, Writer, Renderer, PbfGlyph).

Thin ctypes wrappers over libvgb200host.so / libb200sdf.so — all work happens in native code.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as N

GLYPH_SIZE = 24  # reference src/render/mod.rs:52
BUFFER = 3  # reference src/render/mod.rs:58
GLYPH_BLOCK_SIZE = 256  # reference src/font/glyph_block.rs:7


GLYPH_JOB_DT = np.dtype([("seg_off", "<u4"), ("seg_cnt", "<u4"), ("width", "<u4"), ("height", "<u4"), ("out_off", "<u8")])
OUTLINE_JOB_DT = np.dtype(
    [("kind", "<u4"), ("src_off", "<u4"), ("src_cnt", "<u4"), ("seg_cnt", "<u4"), ("width", "<u4"), ("height", "<u4"),
     ("x0", "<i4"), ("y0", "<i4"), ("scale", "<f8"), ("dx", "<f8"), ("out_off", "<u8")]
)
CURVE_DT = np.dtype(
    [("sx", "<f4"), ("sy", "<f4"), ("cx", "<f4"), ("cy", "<f4"), ("ex", "<f4"), ("ey", "<f4"), ("seg_off", "<u4"), ("depth", "<u4")]
)
assert OUTLINE_JOB_DT.itemsize == C.sizeof(N.OutlineJob) and CURVE_DT.itemsize == C.sizeof(N.Curve)

TILE_JOB_DT = np.dtype([("seg_off", "<u4"), ("seg_cnt", "<u4"), ("out_off", "<u8"), ("width", "<u2"), ("height", "<u2"), ("tx0", "<u2"),
                        ("ty0", "<u2"), ("ntx", "<u2"), ("nty", "<u2"), ("job", "<u4")])
assert TILE_JOB_DT.itemsize == C.sizeof(N.TileJob) == 32
GLYPH_PART_DT = np.dtype([("font", "<u4"), ("glyf_off", "<u4"), ("glyf_len", "<u4"), ("ox", "<f4"), ("oy", "<f4")])
GLYPH_REQ_DT = np.dtype([
    ("kind", "<u4"), ("src_off", "<u4"), ("src_cnt", "<u4"), ("seg_cnt", "<u4"), ("width", "<u4"), ("height", "<u4"),
    ("x0", "<i4"), ("y0", "<i4"), ("scale", "<f8"), ("dx", "<f8"), ("out_off", "<u8"), ("out_cap", "<u4"),
    ("curve_off", "<u4"), ("curve_cap", "<u4"), ("reserved", "<u4")])
GLYPH_FRAME_DT = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("width", "<u4"), ("height", "<u4"), ("seg_cnt", "<u4"), ("status", "<u4")])
assert GLYPH_PART_DT.itemsize == C.sizeof(N.GlyphPart) and GLYPH_REQ_DT.itemsize == C.sizeof(N.GlyphReq) == 72
assert GLYPH_FRAME_DT.itemsize == C.sizeof(N.GlyphFrame) == 24



class B200Error(RuntimeError):
    pass


@dataclass
class PbfGlyph:
    """reference src/protobuf/glyph.rs:10-41"""

    id: int
    bitmap: Optional[bytes]
    width: int
    height: int
    left: int
    top: int
    advance: int
    n_segments: int = 0


def _glyph_from_c(g: N.Glyph) -> PbfGlyph:
    bm = C.string_at(g.bitmap, g.bitmap_len) if g.has_bitmap else None
    return PbfGlyph(g.id, bm, g.width, g.height, g.left, g.top, g.advance, g.n_segments)


class FontFileEntry:
    """reference src/font/file_entry.rs:13-56 (Face + code point set)."""

    def __init__(self, data: Optional[bytes] = None, path: Optional[str] = None):
        if data is not None:
            self._h = N.host.vgb_font_from_bytes(data, len(data))
        else:
            self._h = N.host.vgb_font_from_path(path.encode())
        if not self._h:
            raise B200Error(N.host_error())

    def __del__(self):
        if getattr(self, "_h", None):
            N.host.vgb_font_free(self._h)
            self._h = None

    @property
    def metadata(self) -> dict:
        """FontMetadata (reference src/font/metadata.rs:20-64,84-129) + generate_name()."""
        cap = 1024
        bufs = [C.create_string_buffer(cap) for _ in range(5)]
        weight = C.c_uint16()
        if N.host.vgb_font_metadata(self._h, bufs[0], bufs[1], bufs[2], C.byref(weight), bufs[3], bufs[4], cap) != 0:
            raise B200Error(N.host_error())
        name, family, style, width, generated = (b.value.decode() for b in bufs)
        return {"name": name, "family": family, "style": style, "weight": int(weight.value), "width": width,
                "generated_name": generated}

    @property
    def units_per_em(self) -> int:
        return N.host.vgb_font_units_per_em(self._h)

    @property
    def number_of_glyphs(self) -> int:
        return N.host.vgb_font_number_of_glyphs(self._h)

    def glyph_index(self, cp: int) -> Optional[int]:
        g = N.host.vgb_font_glyph_index(self._h, cp)
        return None if g < 0 else g

    def glyph_hor_advance(self, gid: int) -> Optional[int]:
        a = N.host.vgb_font_hor_advance(self._h, gid)
        return None if a < 0 else a

    def codepoints(self) -> np.ndarray:
        n = N.host.vgb_font_codepoints(self._h, None, 0)
        out = np.zeros(n, dtype=np.uint32)
        N.host.vgb_font_codepoints(self._h, out.ctypes.data_as(N.u32p), n)
        return out

    def outline_commands(self, gid: int) -> np.ndarray:
        """The raw outline callbacks (ttf_parser::OutlineBuilder) of a glyph: float32 [n, 7] = kind (0 move_to, 1 line_to,
        2 quad_to, 3 curve_to, 4 close), x1, y1, x2, y2, x, y."""
        p = C.POINTER(C.c_float)()
        n = N.host.vgb_font_outline_commands(self._h, gid, C.byref(p))
        if n < 0:
            raise B200Error(N.host_error())
        if n == 0:
            return np.zeros((0, 7), dtype=np.float32)
        out = np.ctypeslib.as_array(p, shape=(n, 7)).copy()
        N.host.vgb_free(p)
        return out

    def outline_rings(self, gid: int):
        """RingBuilder over outline_glyph: (xy float64 [n,2], ring_start uint32 [n_rings+1]) in font units."""
        xy = N.f64p()
        rs = N.u32p()
        npts = C.c_uint32()
        nr = N.host.vgb_font_outline_rings(self._h, gid, C.byref(xy), C.byref(rs), C.byref(npts))
        pts = np.ctypeslib.as_array(xy, shape=(max(npts.value, 1) * 2,))[: npts.value * 2].copy().reshape(-1, 2)
        starts = np.ctypeslib.as_array(rs, shape=(nr + 1,)).copy()
        N.host.vgb_free(xy)
        N.host.vgb_free(rs)
        return pts, starts


class Renderer:
    """reference src/render/renderer.rs:17-43.  ``Renderer(dummy=False)`` is the CUDA renderer
    (the reference's ``Precise`` arm replaced by the sm_100a kernel); ``Renderer(dummy=True)`` is the
    zero-bitmap fake of renderer_dummy.rs.  Raises when no B200 is usable: there is no CPU path."""

    def __init__(self, dummy: bool = False, device: int = 0, n_slots: int = 0):
        self._h = N.host.vgb_renderer_new(1 if dummy else 0, device, n_slots)
        if not self._h:
            raise B200Error(N.host_error())
        self.dummy = dummy

    @classmethod
    def new_precise(cls, device: int = 0, n_slots: int = 0):
        return cls(False, device, n_slots)

    @classmethod
    def new_dummy(cls):
        return cls(True)

    def __del__(self):
        if getattr(self, "_h", None):
            N.host.vgb_renderer_free(self._h)
            self._h = None

    @property
    def context(self):
        """b200sdf_ctx* of the CUDA renderer (None for dummy)."""
        return N.host.vgb_renderer_context(self._h)

    def set_flatten(self, mode):
        """What batches created from now on send to the device: "glyf" / 2 = glyf record references (the device decodes,
        records, measures, plans and renders; default of the CUDA renderer), True / "device" / 1 = curve records made on
        the host, flattened on the device, False / "host" / 0 = segments flattened on the host (the literal
        renderer_precise seam)."""
        mode = {"glyf": 2, "device": 1, "host": 0}.get(mode, mode)
        N.host.vgb_renderer_set_flatten(self._h, int(mode))

    @property
    def flatten(self) -> int:
        return N.host.vgb_renderer_flatten(self._h)

    def render_glyph(self, font: FontFileEntry, index: int) -> Optional[PbfGlyph]:
        """reference src/render/renderer.rs:103-149"""
        g = N.Glyph()
        rc = N.host.vgb_renderer_render_glyph(self._h, font._h, index, C.byref(g))
        if rc < 0:
            raise B200Error(N.host_error())
        if rc == 0:
            return None
        out = _glyph_from_c(g)
        N.host.vgb_free(g.bitmap)
        return out

    def new_batch(self) -> "GlyphBatch":
        return GlyphBatch(self)

    def render_batch(self, batch: "GlyphBatch"):
        if N.host.vgb_renderer_render_batch(self._h, batch._h) != 0:
            raise B200Error(N.host_error())

    def submit_batch(self, batch: "GlyphBatch") -> int:
        t = C.c_uint64()
        if N.host.vgb_renderer_submit_batch(self._h, batch._h, C.byref(t)) != 0:
            raise B200Error(N.host_error())
        return t.value

    def wait_batch(self, ticket: int):
        if N.host.vgb_renderer_wait_batch(self._h, ticket) != 0:
            raise B200Error(N.host_error())

    def prepare_batch(self, batch: "GlyphBatch"):
        """Plan the batch's tiles now (any thread); submit_batch then only enqueues (b200sdf_submit_planned)."""
        if N.host.vgb_renderer_prepare_batch(self._h, batch._h) != 0:
            raise B200Error(N.host_error())

    def poll_batch(self, ticket: int) -> bool:
        """Non-blocking wait_batch: True = finished (the ticket is consumed), False = still running."""
        rc = N.host.vgb_renderer_poll_batch(self._h, ticket)
        if rc < 0:
            raise B200Error(N.host_error())
        return rc == 1

    def finalize_batch(self, batch: "GlyphBatch"):
        """After wait_batch / poll_batch: take the frames a glyph-level batch got back from the device (no-op otherwise)."""
        if N.host.vgb_batch_finalize(self._h, batch._h) != 0:
            raise B200Error(N.host_error())


class GlyphBatch:
    """The flat segment buffer of one GlyphBlock (north star: "packed into a flat SoA segment buffer
    per GlyphBlock and uploaded once")."""

    def __init__(self, renderer: Renderer):
        self._h = N.host.vgb_batch_new(renderer._h)

    def __del__(self):
        if getattr(self, "_h", None):
            N.host.vgb_batch_free(self._h)
            self._h = None

    def clear(self):
        N.host.vgb_batch_clear(self._h)

    def add_glyph(self, font: FontFileEntry, cp: int) -> bool:
        return N.host.vgb_batch_add_glyph(self._h, font._h, cp) == 1

    def add_rings(self, gid: int, x0: int, y0: int, width: int, height: int, xy: np.ndarray, ring_start: np.ndarray):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        rs = np.ascontiguousarray(ring_start, dtype=np.uint32)
        rc = N.host.vgb_batch_add_rings(
            self._h, gid, x0, y0, width, height, xy.ctypes.data_as(N.f64p), rs.ctypes.data_as(N.u32p), len(rs) - 1
        )
        if rc != 1:
            raise B200Error(N.host_error())

    def __len__(self):
        return N.host.vgb_batch_glyph_count(self._h)

    def glyph_info(self, i: int) -> N.BatchGlyph:
        g = N.BatchGlyph()
        if N.host.vgb_batch_glyph_info(self._h, i, C.byref(g)) != 0:
            raise B200Error(N.host_error())
        return g

    def segments(self) -> np.ndarray:
        n = C.c_uint32()
        p = N.host.vgb_batch_segments(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros((0, 4), dtype=np.float32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 4))

    def jobs(self) -> np.ndarray:
        """b200sdf_outline_job records (one per bitmap)."""
        n = C.c_uint32()
        p = N.host.vgb_batch_jobs(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=OUTLINE_JOB_DT)
        buf = C.string_at(p, n.value * C.sizeof(N.OutlineJob))
        return np.frombuffer(buf, dtype=OUTLINE_JOB_DT).copy()

    def glyph_jobs(self) -> np.ndarray:
        """The same jobs as b200sdf_glyph_job records (only valid when every glyph is a SEGMENTS glyph)."""
        j = self.jobs()
        assert (j["kind"] == N.KIND_SEGMENTS).all()
        out = np.zeros(len(j), dtype=GLYPH_JOB_DT)
        out["seg_off"], out["seg_cnt"] = j["src_off"], j["seg_cnt"]
        out["width"], out["height"], out["out_off"] = j["width"], j["height"], j["out_off"]
        return out

    def curves(self) -> np.ndarray:
        n = C.c_uint32()
        p = N.host.vgb_batch_curves(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=CURVE_DT)
        return np.frombuffer(C.string_at(p, n.value * C.sizeof(N.Curve)), dtype=CURVE_DT).copy()

    @property
    def total_segments(self) -> int:
        return N.host.vgb_batch_total_segments(self._h)

    @property
    def fallback_glyphs(self) -> int:
        return N.host.vgb_batch_fallback_glyphs(self._h)

    def bitmaps(self) -> np.ndarray:
        n = C.c_uint64()
        p = N.host.vgb_batch_bitmaps(self._h, C.byref(n))
        if n.value == 0 or not p:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    @property
    def pairs(self) -> int:
        return N.host.vgb_batch_pairs(self._h)

    def bitmap_of(self, i: int) -> Optional[np.ndarray]:
        g = self.glyph_info(i)
        if not g.has_bitmap:
            return None
        n = C.c_uint64()
        p = N.host.vgb_batch_glyph_bitmap(self._h, i, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).reshape(g.bm_height, g.bm_width)

    # ---- glyph-level batches (the device decodes glyf records) ----
    def requests(self) -> np.ndarray:
        n = C.c_uint32()
        p = N.host.vgb_batch_requests(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=GLYPH_REQ_DT)
        return np.frombuffer(C.string_at(p, n.value * C.sizeof(N.GlyphReq)), dtype=GLYPH_REQ_DT).copy()

    def parts(self) -> np.ndarray:
        n = C.c_uint32()
        p = N.host.vgb_batch_parts(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=GLYPH_PART_DT)
        return np.frombuffer(C.string_at(p, n.value * C.sizeof(N.GlyphPart)), dtype=GLYPH_PART_DT).copy()

    @property
    def curve_slots(self) -> int:
        return N.host.vgb_batch_curve_slots(self._h)

    @property
    def tile_cap(self) -> int:
        return N.host.vgb_batch_tile_cap(self._h)

    @property
    def est_cost(self) -> int:
        return N.host.vgb_batch_est_cost(self._h)

    @property
    def handed_back(self) -> int:
        return N.host.vgb_batch_handed_back(self._h)

    @property
    def path_glyphs(self) -> int:
        """glyphs with cubic curves sent as kind PATH (the device flattens them)"""
        return N.host.vgb_batch_path_glyphs(self._h)


class Writer:
    """reference src/writer/mod.rs:27-96 (directory sink and in-memory recorder)."""

    def __init__(self, folder: Optional[str] = None, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            self._h = N.host.vgb_writer_new_file(folder.encode()) if folder else N.host.vgb_writer_new_memory()

    @classmethod
    def new_file(cls, folder: str):
        return cls(folder)

    @classmethod
    def new_memory(cls):
        return cls(None)

    @classmethod
    def new_tar(cls, path: str):
        """reference src/writer/mod.rs:27-33 — a ustar stream (src/writer/tar.rs) written to `path`."""
        return cls(_handle=N.host.vgb_writer_new_tar(path.encode()))

    @classmethod
    def new_tar_memory(cls):
        """TarWriter over an in-memory buffer, as in the reference's src/writer/tar.rs tests."""
        return cls(_handle=N.host.vgb_writer_new_tar_memory())

    def write_file(self, filename: str, data: bytes):
        if N.host.vgb_writer_write_file(self._h, filename.encode(), data, len(data)) != 0:
            raise B200Error(N.host_error())

    def write_directory(self, dirname: str):
        if N.host.vgb_writer_write_directory(self._h, dirname.encode()) != 0:
            raise B200Error(N.host_error())

    def finish(self):
        if N.host.vgb_writer_finish(self._h) != 0:
            raise B200Error(N.host_error())

    def tar_bytes(self) -> bytes:
        n = C.c_uint64()
        p = N.host.vgb_writer_tar_bytes(self._h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def __del__(self):
        if getattr(self, "_h", None):
            N.host.vgb_writer_free(self._h)
            self._h = None

    def entries(self):
        """[(name, is_dir, bytes)] recorded by the in-memory writer."""
        out = []
        for i in range(N.host.vgb_writer_entry_count(self._h)):
            name = C.c_char_p()
            is_dir = C.c_int32()
            data = N.u8p()
            n = C.c_uint64()
            N.host.vgb_writer_entry(self._h, i, C.byref(name), C.byref(is_dir), C.byref(data), C.byref(n))
            out.append((name.value.decode(), bool(is_dir.value), C.string_at(data, n.value) if n.value else b""))
        return out


@dataclass
class RenderStats:
    glyphs: int
    bitmaps: int
    pixels: int
    segments: int
    pairs: int
    pbf_bytes: int
    blocks: int
    # host time per phase in ns, summed over workers (wall_ns: the whole call)
    outline_ns: int = 0
    submit_ns: int = 0
    wait_ns: int = 0
    encode_ns: int = 0
    write_ns: int = 0
    wall_ns: int = 0
    submits: int = 0
    workers: int = 0
    handed_back: int = 0  # glyphs the device decoder returned to the host recorder
    h2d_bytes: int = 0    # request / record / segment bytes the device read from host memory
    cost_total: int = 0   # estimated cost of the whole job (0 when not sharded)
    cost_shard: int = 0   # ... and of this shard


class FontManager:
    """reference src/font/manager.rs:18-147."""

    def __init__(self, parallel: bool = True):
        self._h = N.host.vgb_manager_new(1 if parallel else 0)

    def __del__(self):
        if getattr(self, "_h", None):
            N.host.vgb_manager_free(self._h)
            self._h = None

    def add_path(self, path: str):
        if N.host.vgb_manager_add_path(self._h, path.encode()) != 0:
            raise B200Error(N.host_error())

    def font_file_names(self, font_id: str) -> List[str]:
        """metadata.name of the files merged into one font, in the order they were added."""
        n = N.host.vgb_manager_font_file_names(self._h, font_id.encode(), None, 0)
        buf = C.create_string_buffer(n + 1)
        N.host.vgb_manager_font_file_names(self._h, font_id.encode(), buf, n + 1)
        return buf.value.decode().split("\n") if n else []

    def scan(self, path: str):
        """reference src/commands/recurse.rs:104-133: font files, fonts.json manifests, recursion into directories."""
        if N.host.vgb_manager_scan(self._h, path.encode()) != 0:
            raise B200Error(N.host_error())

    def add_paths(self, paths: List[str]):
        for p in paths:
            self.add_path(p)

    def add_font_with_name(self, name: str, sources: List[str]):
        arr = (C.c_char_p * len(sources))(*[s.encode() for s in sources])
        if N.host.vgb_manager_add_font_with_name(self._h, name.encode(), arr, len(sources)) != 0:
            raise B200Error(N.host_error())

    def add_font_bytes_with_name(self, name: str, data: bytes):
        if N.host.vgb_manager_add_font_bytes_with_name(self._h, name.encode(), data, len(data)) != 0:
            raise B200Error(N.host_error())

    def font_ids(self) -> List[str]:
        return [N.host.vgb_manager_font_id(self._h, i).decode() for i in range(N.host.vgb_manager_font_count(self._h))]

    def block_population(self, font_id: str) -> np.ndarray:
        out = np.zeros(256, dtype=np.uint32)
        if N.host.vgb_manager_block_population(self._h, font_id.encode(), out.ctypes.data_as(N.u32p)) != 0:
            raise B200Error(N.host_error())
        return out

    def render_block(self, font_id: str, block: int, renderer: Renderer) -> bytes:
        """GlyphBlock::render (reference src/font/glyph_block.rs:69-80), glyphs in ascending id order."""
        p = N.u8p()
        n = C.c_uint64()
        if N.host.vgb_manager_render_block(self._h, font_id.encode(), block, renderer._h, C.byref(p), C.byref(n)) != 0:
            raise B200Error(N.host_error())
        out = C.string_at(p, n.value)
        N.host.vgb_free(p)
        return out

    def render_glyphs(self, writer: Writer, renderer: Renderer, shard: int = 0, n_shards: int = 1, threads: int = 0) -> RenderStats:
        """reference src/font/manager.rs:81-125 as an async batch pipeline over CUDA streams."""
        st = N.Stats()
        if N.host.vgb_manager_render_glyphs(self._h, writer._h, renderer._h, shard, n_shards, threads, C.byref(st)) != 0:
            raise B200Error(N.host_error())
        return RenderStats(*[int(getattr(st, n)) for n, _ in N.Stats._fields_])

    def shard_owners(self, n_shards: int):
        """-> (owner[font, block] uint16 array over font_ids() x 256 blocks, estimated cost per shard)."""
        n = len(self.font_ids()) * 256
        owner = np.zeros(n, dtype=np.uint16)
        loads = np.zeros(max(1, n_shards), dtype=np.uint64)
        rc = N.host.vgb_manager_shard_owners(self._h, n_shards, owner.ctypes.data_as(C.POINTER(C.c_uint16)), n,
                                             loads.ctypes.data_as(N.u64p))
        if rc < 0:
            raise B200Error(N.host_error())
        return owner.reshape(-1, 256), loads

    def write_index_json(self, writer: Writer):
        if N.host.vgb_manager_write_index_json(self._h, writer._h) != 0:
            raise B200Error(N.host_error())


    def write_families_json(self, writer: Writer):
        """reference src/font/manager.rs:134-137 (font_families.json, src/font/index_files.rs:115-139)."""
        if N.host.vgb_manager_write_families_json(self._h, writer._h) != 0:
            raise B200Error(N.host_error())


def parse_font_name(family: str, ps_name: str):
    """(family, style, weight, width) — reference src/font/parse_font_name.rs:214-293."""
    cap = 4 * (len(family) + len(ps_name)) + 64
    fam, style, width = (C.create_string_buffer(cap) for _ in range(3))
    weight = C.c_uint16()
    if N.host.vgb_parse_font_name(family.encode(), ps_name.encode(), fam, style, C.byref(weight), width, cap) != 0:
        raise B200Error(N.host_error())
    return fam.value.decode(), style.value.decode(), int(weight.value), width.value.decode()


def encode_codeblocks(codepoints) -> str:
    """reference src/font/index_files.rs:60-95."""
    cps = np.ascontiguousarray(codepoints, dtype=np.uint32)
    p = cps.ctypes.data_as(N.u32p)
    n = N.host.vgb_encode_codeblocks(p, cps.size, None, 0)
    buf = C.create_string_buffer(n + 1)
    N.host.vgb_encode_codeblocks(p, cps.size, buf, n + 1)
    return buf.value.decode()


def name_to_id(name: str) -> str:
    buf = C.create_string_buffer(4 * len(name) + 8)
    return N.host.vgb_name_to_id(name.encode(), buf, len(buf)).decode()


def decode_pbf(data: bytes):
    """(name, range, [PbfGlyph]) — mirror of the prost decode in reference src/commands/debug.rs:60-79."""
    name = C.create_string_buffer(512)
    rng = C.create_string_buffer(64)
    gl = C.POINTER(N.Glyph)()
    n = N.host.vgb_pbf_decode(data, len(data), name, len(name), rng, len(rng), C.byref(gl))
    if n < 0:
        raise B200Error(N.host_error())
    out = [_glyph_from_c(gl[i]) for i in range(n)]
    N.host.vgb_glyphs_free(gl, n)
    return name.value.decode(), rng.value.decode(), out


class SdfContext:
    """Direct view of the device ABI (include/b200sdf.h) over numpy buffers."""

    def __init__(self, device: int = 0, n_slots: int = 2, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        rc = N.sdf.b200sdf_create(device, n_slots, C.byref(h))
        if rc != 0:
            raise B200Error(f"b200sdf_create failed with code {rc}: a B200 (sm_100) GPU is required; there is no CPU fallback")
        self._h = h

    @classmethod
    def of_renderer(cls, renderer: "Renderer"):
        """The renderer's own context (not owned): the fonts it uploaded are resident there."""
        return cls(_borrowed=renderer.context)

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_owned", False):
            N.sdf.b200sdf_destroy(self._h)
        self._h = None

    # ---- glyph-level path (device glyf decoding) ----
    def font_upload(self, glyf: bytes) -> int:
        h = C.c_uint32()
        buf = np.frombuffer(glyf, dtype=np.uint8)
        rc = N.sdf.b200sdf_font_upload(self._h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(h))
        if rc != 0:
            raise B200Error(f"b200sdf_font_upload: {rc}: {self.last_error()}")
        return h.value

    def decode_glyphs(self, reqs: np.ndarray, parts: np.ndarray, curve_slots: int, curves: Optional[np.ndarray] = None, n_seg: int = 0,
                      est_cost: int = 0):
        """b200sdf_decode_glyphs -> (frames, outline jobs, curve scratch, tile jobs in claim order).  curves / n_seg:
        the host-recorded arrays CURVES / SEGMENTS requests index."""
        reqs = np.ascontiguousarray(reqs, dtype=GLYPH_REQ_DT)
        parts = np.ascontiguousarray(parts, dtype=GLYPH_PART_DT)
        hc = np.zeros(0, dtype=CURVE_DT) if curves is None else np.ascontiguousarray(curves, dtype=CURVE_DT)
        frames = np.zeros(len(reqs), dtype=GLYPH_FRAME_DT)
        jobs = np.zeros(len(reqs), dtype=OUTLINE_JOB_DT)
        curves = np.zeros(max(1, curve_slots), dtype=CURVE_DT)
        n_tiles = C.c_uint32()
        cap = max(64, 64 * len(reqs))  # decode_glyphs' own tile capacity
        tiles = np.zeros(cap, dtype=TILE_JOB_DT)
        rc = N.sdf.b200sdf_decode_glyphs(self._h, reqs.ctypes.data, len(reqs), parts.ctypes.data, len(parts),
                                         hc.ctypes.data if len(hc) else None, len(hc), n_seg, curve_slots, est_cost, frames.ctypes.data,
                                         jobs.ctypes.data, curves.ctypes.data, C.byref(n_tiles), tiles.ctypes.data, cap)
        if rc != 0:
            raise B200Error(f"b200sdf_decode_glyphs: {rc}: {self.last_error()}")
        return frames, jobs, curves[:curve_slots], tiles[: min(cap, n_tiles.value)]

    def render_glyphs(self, reqs: np.ndarray, parts: np.ndarray, curve_slots: int, tile_cap: int, out_bytes: int,
                      curves: Optional[np.ndarray] = None, segs: Optional[np.ndarray] = None, est_cost: int = 0):
        """b200sdf_submit_glyphs + b200sdf_wait over (pageable) host buffers -> (frames, bitmaps)."""
        reqs = np.ascontiguousarray(reqs, dtype=GLYPH_REQ_DT)
        parts = np.ascontiguousarray(parts, dtype=GLYPH_PART_DT)
        curves = np.zeros(0, dtype=CURVE_DT) if curves is None else np.ascontiguousarray(curves, dtype=CURVE_DT)
        segs = np.zeros((0, 4), dtype=np.float32) if segs is None else np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
        frames = np.zeros(max(1, len(reqs)), dtype=GLYPH_FRAME_DT)
        out = np.zeros(max(out_bytes, 1), dtype=np.uint8)
        t = C.c_uint64()
        rc = N.sdf.b200sdf_submit_glyphs(self._h, reqs.ctypes.data, len(reqs), parts.ctypes.data, len(parts),
                                         curves.ctypes.data if len(curves) else None, len(curves),
                                         segs.ctypes.data if len(segs) else None, len(segs), curve_slots, tile_cap, est_cost,
                                         frames.ctypes.data, out.ctypes.data, out_bytes, C.byref(t))
        if rc == 0:
            rc = N.sdf.b200sdf_wait(self._h, t.value)
        if rc != 0:
            raise B200Error(f"b200sdf_submit_glyphs: {rc}: {self.last_error()}")
        return frames[: len(reqs)], out[:out_bytes]

    def render_glyph_batches(self, batches, est_cost: int = 0):
        """b200sdf_submit_glyph_batches + b200sdf_wait: `batches` is a list of (reqs, parts, curve_slots, tile_cap, out_bytes);
        every buffer is copied into b200sdf_alloc_pinned memory first (several batches in one submission must be
        pinned).  -> [(frames, bitmaps)] per batch."""
        held, views = [], []

        def pinned(a: np.ndarray):
            nbytes = max(int(a.nbytes), 64)
            p = N.sdf.b200sdf_alloc_pinned(nbytes)
            if not p:
                raise B200Error("b200sdf_alloc_pinned failed")
            held.append(p)
            v = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))
            v[: a.nbytes] = a.view(np.uint8).reshape(-1)
            return p, v

        try:
            desc = (N.GlyphBatchDesc * len(batches))()
            for k, (reqs, parts, curve_slots, tile_cap, out_bytes) in enumerate(batches):
                reqs = np.ascontiguousarray(reqs, dtype=GLYPH_REQ_DT)
                parts = np.ascontiguousarray(parts, dtype=GLYPH_PART_DT)
                d = desc[k]
                d.reqs, _ = pinned(reqs)
                d.parts, _ = pinned(parts)
                d.n_reqs, d.n_parts = len(reqs), len(parts)
                d.curve_slots, d.tile_cap = int(curve_slots), int(tile_cap)
                d.frames, fv = pinned(np.zeros(max(1, len(reqs)), dtype=GLYPH_FRAME_DT))
                d.out, ov = pinned(np.zeros(max(1, out_bytes), dtype=np.uint8))
                d.out_bytes = int(out_bytes)
                views.append((fv, ov, len(reqs), int(out_bytes)))
            t = C.c_uint64()
            rc = N.sdf.b200sdf_submit_glyph_batches(self._h, desc, len(batches), est_cost, C.byref(t))
            if rc == 0:
                rc = N.sdf.b200sdf_wait(self._h, t.value)
            if rc != 0:
                raise B200Error(f"b200sdf_submit_glyph_batches: {rc}: {self.last_error()}")
            return [(fv[: n * GLYPH_FRAME_DT.itemsize].view(GLYPH_FRAME_DT).copy(), ov[:ob].copy()) for fv, ov, n, ob in views]
        finally:
            for p in held:
                N.sdf.b200sdf_free_pinned(p)

    def render_glyphs_device(self, d_reqs: int, n_reqs: int, d_parts: int, n_parts: int, d_curves: int, n_curves: int, d_segs: int,
                             n_seg: int, curve_slots: int, tile_cap: int, est_cost: int, d_frames: int, d_out: int, out_bytes: int,
                             stream: int = 0, mid_event: int = 0):
        rc = N.sdf.b200sdf_render_glyphs_device(self._h, d_reqs, n_reqs, d_parts, n_parts, d_curves, n_curves, d_segs, n_seg,
                                                curve_slots, tile_cap, est_cost, d_frames, d_out, out_bytes, stream, mid_event)
        if rc != 0:
            raise B200Error(f"b200sdf_render_glyphs_device: {rc}: {self.last_error()}")

    def last_error(self) -> str:
        return (N.sdf.b200sdf_last_error(self._h) or b"").decode()

    def render(self, segs: np.ndarray, jobs: np.ndarray, out_bytes: int) -> np.ndarray:
        """b200sdf_render over host buffers: segs float32 [n,4], jobs structured (see GlyphBatch.jobs)."""
        segs = np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
        jobs = np.ascontiguousarray(jobs)
        out = np.zeros(max(out_bytes, 1), dtype=np.uint8)
        rc = N.sdf.b200sdf_render(self._h, segs.ctypes.data, len(segs), jobs.ctypes.data, len(jobs), out.ctypes.data, out_bytes)
        if rc != 0:
            raise B200Error(f"b200sdf_render: {rc}: {self.last_error()}")
        return out[:out_bytes]

    def plan_tiles(self, jobs: np.ndarray, n_seg: int, out_bytes: int):
        jobs = np.ascontiguousarray(jobs)
        n = C.c_uint32()
        pairs = C.c_uint64()
        rc = N.sdf.b200sdf_plan_tiles(jobs.ctypes.data, len(jobs), n_seg, out_bytes, None, 0, C.byref(n), C.byref(pairs))
        if rc != 0:
            raise B200Error(f"b200sdf_plan_tiles: {rc}")
        tiles = np.zeros(n.value * C.sizeof(N.TileJob), dtype=np.uint8)
        rc = N.sdf.b200sdf_plan_tiles(jobs.ctypes.data, len(jobs), n_seg, out_bytes, tiles.ctypes.data, n.value, C.byref(n), C.byref(pairs))
        if rc != 0:
            raise B200Error(f"b200sdf_plan_tiles: {rc}")
        return tiles, n.value, pairs.value

    def render_device(self, d_segs: int, d_tiles: int, n_tiles: int, d_out: int, stream: int = 0):
        rc = N.sdf.b200sdf_render_device(self._h, d_segs, d_tiles, n_tiles, d_out, stream)
        if rc != 0:
            raise B200Error(f"b200sdf_render_device: {rc}: {self.last_error()}")

    # ---- outline-level path (device flattening) ----
    def render_outlines(self, curves: np.ndarray, segs: np.ndarray, jobs: np.ndarray, out_bytes: int) -> np.ndarray:
        """b200sdf_submit_outlines + b200sdf_wait over host buffers."""
        curves = np.ascontiguousarray(curves, dtype=CURVE_DT)
        segs = np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
        jobs = np.ascontiguousarray(jobs, dtype=OUTLINE_JOB_DT)
        out = np.zeros(max(out_bytes, 1), dtype=np.uint8)
        t = C.c_uint64()
        rc = N.sdf.b200sdf_submit_outlines(self._h, curves.ctypes.data, len(curves), segs.ctypes.data, len(segs),
                                           jobs.ctypes.data, len(jobs), out.ctypes.data, out_bytes, C.byref(t))
        if rc == 0:
            rc = N.sdf.b200sdf_wait(self._h, t.value)
        if rc != 0:
            raise B200Error(f"b200sdf_submit_outlines: {rc}: {self.last_error()}")
        return out[:out_bytes]

    def flatten_outlines(self, curves: np.ndarray, jobs: np.ndarray) -> np.ndarray:
        """Device flattening only: float32 [sum seg_cnt, 4] origin-relative segments, glyph after glyph."""
        curves = np.ascontiguousarray(curves, dtype=CURVE_DT)
        jobs = np.ascontiguousarray(jobs, dtype=OUTLINE_JOB_DT)
        n = int(jobs["seg_cnt"].sum())
        out = np.zeros((max(n, 1), 4), dtype=np.float32)
        rc = N.sdf.b200sdf_flatten_outlines(self._h, curves.ctypes.data, len(curves), jobs.ctypes.data, len(jobs), out.ctypes.data, n)
        if rc != 0:
            raise B200Error(f"b200sdf_flatten_outlines: {rc}: {self.last_error()}")
        return out[:n]

    def plan_outline_tiles(self, jobs: np.ndarray, n_curves: int, n_seg: int, out_bytes: int):
        jobs = np.ascontiguousarray(jobs, dtype=OUTLINE_JOB_DT)
        n = C.c_uint32()
        pairs = C.c_uint64()
        args = (jobs.ctypes.data, len(jobs), n_curves, n_seg, out_bytes)
        rc = N.sdf.b200sdf_plan_outline_tiles(*args, None, 0, C.byref(n), C.byref(pairs))
        if rc != 0:
            raise B200Error(f"b200sdf_plan_outline_tiles: {rc}")
        tiles = np.zeros(n.value * C.sizeof(N.TileJob), dtype=np.uint8)
        rc = N.sdf.b200sdf_plan_outline_tiles(*args, tiles.ctypes.data, n.value, C.byref(n), C.byref(pairs))
        if rc != 0:
            raise B200Error(f"b200sdf_plan_outline_tiles: {rc}")
        return tiles, n.value, pairs.value

    def render_outlines_device(self, d_curves: int, d_segs: int, d_jobs: int, d_tiles: int, n_tiles: int, d_out: int, stream: int = 0):
        rc = N.sdf.b200sdf_render_outlines_device(self._h, d_curves, d_segs, d_jobs, d_tiles, n_tiles, d_out, stream)
        if rc != 0:
            raise B200Error(f"b200sdf_render_outlines_device: {rc}: {self.last_error()}")

    def measure_fp32_peak(self, reps: int = 5, packed: bool = False):
        """FFMA-chain microbenchmark -> (TFLOP/s, ms); packed=True uses FFMA2 (two FMAs per instruction)."""
        t = C.c_double()
        ms = C.c_double()
        rc = N.sdf.b200sdf_measure_fp32_peak(self._h, -reps if packed else reps, C.byref(t), C.byref(ms))
        if rc != 0:
            raise B200Error(f"b200sdf_measure_fp32_peak: {rc}: {self.last_error()}")
        return t.value, ms.value

    @property
    def launch_count(self) -> int:
        return N.sdf.b200sdf_launch_count(self._h)


def device_count() -> int:
    return N.sdf.b200sdf_device_count()
