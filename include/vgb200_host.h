/*
 * vgb200_host.h — C view of the C++ host library (libvgb200host.so).
 *
 * The reference's host is Rust (FontManager / FontWrapper / GlyphBlock / Writer / Renderer,
 * reference src/lib.rs:8-13) and there is no Rust toolchain in this image, so the host side above
 * the device ABI (b200sdf.h) is written in C++ with the same names, argument meaning and error
 * behaviour; this header exposes it to C / ctypes.  Each function cites what it mirrors.
 * Conventions: handles are opaque pointers; functions returning int give 0 on success and a
 * negative value on error with a message retrievable through vgb_last_error() (thread-local).
 */
#ifndef VGB200_HOST_H
#define VGB200_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "b200sdf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgb_font vgb_font;         /* FontFileEntry  (src/font/file_entry.rs:13) */
typedef struct vgb_renderer vgb_renderer; /* Renderer       (src/render/renderer.rs:17) */
typedef struct vgb_batch vgb_batch;       /* flat segment buffer of one GlyphBlock (new) */
typedef struct vgb_manager vgb_manager;   /* FontManager    (src/font/manager.rs:18)    */
typedef struct vgb_writer vgb_writer;     /* Writer         (src/writer/mod.rs:21)      */

const char *vgb_last_error(void);
void vgb_free(void *p);

/* ---- PbfGlyph (src/protobuf/glyph.rs:10-41) ---- */
typedef struct {
	uint32_t id;
	int32_t has_bitmap;
	uint32_t width, height;
	int32_t left, top;
	uint32_t advance;
	uint8_t *bitmap;     /* malloc'd, (width+6)*(height+6) bytes when has_bitmap; free with vgb_free */
	uint64_t bitmap_len;
	uint32_t n_segments; /* segments handed to the SDF pass */
} vgb_glyph;

/* ---- FontFileEntry / Face ---- */
vgb_font *vgb_font_from_bytes(const uint8_t *data, size_t len);           /* FontFileEntry::new, file_entry.rs:32 */
vgb_font *vgb_font_from_path(const char *path);
void vgb_font_free(vgb_font *f);
uint32_t vgb_font_units_per_em(const vgb_font *f);                         /* renderer.rs:107 */
uint32_t vgb_font_number_of_glyphs(const vgb_font *f);                     /* file_entry.rs:66-71 */
int32_t vgb_font_glyph_index(const vgb_font *f, uint32_t codepoint);       /* renderer.rs:106; -1 = None */
int32_t vgb_font_hor_advance(const vgb_font *f, uint32_t glyph_id);        /* renderer.rs:115; -1 = None */
size_t vgb_font_codepoints(const vgb_font *f, uint32_t *out, size_t cap);  /* metadata.rs:104-118 */
/* RingBuilder over outline_glyph (ring_builder.rs): flattened rings in font units.
 * xy: malloc'd 2*n_points doubles; ring_start: malloc'd n_rings+1 offsets.  Returns n_rings. */
int32_t vgb_font_outline_rings(const vgb_font *f, uint32_t glyph_id, double **xy, uint32_t **ring_start,
                               uint32_t *n_points);

/* The raw outline callbacks of Face::outline_glyph (ttf_parser::OutlineBuilder, src/render/renderer.rs:109-110), before
 * any flattening: malloc'd 7 floats per command — kind (0 move_to, 1 line_to, 2 quad_to, 3 curve_to, 4 close),
 * x1, y1, x2, y2, x, y (unused fields 0).  Returns the number of commands; free with vgb_free.  Used by the
 * cross-check against FreeType (tests/test_freetype_crosscheck.py). */
int32_t vgb_font_outline_commands(const vgb_font *f, uint32_t glyph_id, float **cmds);

/* ---- geometry (src/geometry/ring.rs:119-187, segment.rs:96-99) ---- */
size_t vgb_flatten_quad(const double s[2], const double c[2], const double e[2], double tol_sq, double *out_xy, size_t cap);
size_t vgb_flatten_cubic(const double s[2], const double c1[2], const double c2[2], const double e[2], double tol_sq,
                         double *out_xy, size_t cap);
double vgb_segment_sqdist(double vx, double vy, double wx, double wy, double px, double py);
const char *vgb_name_to_id(const char *name, char *buf, size_t cap);        /* manager.rs:141-147 */

/* ---- font naming + index files (not on the accelerated path; they name the output files) ---- */
/* parse_font_name (font/parse_font_name.rs:214-293): family/style/width are written NUL-terminated
 * into caller buffers of `cap` bytes each; returns 0, or -1 when a buffer is too small. */
int vgb_parse_font_name(const char *family, const char *ps_name, char *out_family, char *out_style, uint16_t *out_weight,
                        char *out_width, size_t cap);
/* encode_codeblocks (font/index_files.rs:60-95): returns the length needed (excluding NUL) */
size_t vgb_encode_codeblocks(const uint32_t *codepoints, size_t n, char *buf, size_t cap);
/* FontMetadata (font/metadata.rs:20-64,84-129) of a parsed file; generate_name() in `generated` */
int vgb_font_metadata(const vgb_font *f, char *name, char *family, char *style, uint16_t *weight, char *width,
                      char *generated, size_t cap);

/* ---- Renderer ---- */
/* Renderer::new(dummy) (renderer.rs:25-31).  dummy=0 -> the CUDA renderer on `device` with
 * n_slots batches in flight (0 = default); NULL when no B200 is usable (no CPU fallback). */
vgb_renderer *vgb_renderer_new(int dummy, int device, uint32_t n_slots);
void vgb_renderer_free(vgb_renderer *r);
int vgb_renderer_is_dummy(const vgb_renderer *r);
/* What new batches send to the device: 2 = glyf record references — the device decodes, records, measures and plans
 * (default of the CUDA renderer), 1 = curve records made by the host's OutlineRecorder, flattened on the device,
 * 0 = segments flattened on the host (the literal renderer_precise seam, src/render/renderer_precise.rs:8) */
void vgb_renderer_set_flatten(vgb_renderer *r, int mode);
int vgb_renderer_flatten(const vgb_renderer *r);
b200sdf_ctx *vgb_renderer_context(const vgb_renderer *r); /* NULL for the dummy renderer */
/* Renderer::render_glyph (renderer.rs:103-149): 1 = Some(glyph), 0 = None, <0 = error */
int vgb_renderer_render_glyph(const vgb_renderer *r, const vgb_font *f, uint32_t codepoint, vgb_glyph *out);

/* ---- GlyphBatch ---- */
typedef struct {
	uint32_t id, advance;
	int32_t has_bitmap;
	int32_t x0, y0;          /* integer origin of the bitmap (RenderResult.x0/.y0) */
	uint32_t bm_width, bm_height; /* RenderResult.width/.height (buffer included) */
	uint32_t width, height;  /* PbfGlyph fields */
	int32_t left, top;
	uint32_t kind;            /* B200SDF_KIND_CURVES / B200SDF_KIND_SEGMENTS */
	uint32_t src_off, src_cnt; /* range in the batch's curve or segment array */
	uint32_t seg_cnt;          /* flattened segments */
	uint64_t out_off;
} vgb_batch_glyph;

vgb_batch *vgb_batch_new(const vgb_renderer *r);
void vgb_batch_free(vgb_batch *b);
void vgb_batch_clear(vgb_batch *b);
int vgb_batch_add_glyph(vgb_batch *b, const vgb_font *f, uint32_t codepoint); /* 1 added, 0 None */
/* renderer_precise's own inputs: frame + closed rings in pixel space (renderer_precise.rs:8) */
int vgb_batch_add_rings(vgb_batch *b, uint32_t id, int32_t x0, int32_t y0, uint32_t width, uint32_t height,
                        const double *xy, const uint32_t *ring_start, uint32_t n_rings);
uint32_t vgb_batch_glyph_count(const vgb_batch *b);
int vgb_batch_glyph_info(const vgb_batch *b, uint32_t i, vgb_batch_glyph *out);
const b200sdf_segment *vgb_batch_segments(const vgb_batch *b, uint32_t *n_seg);
const b200sdf_outline_job *vgb_batch_jobs(const vgb_batch *b, uint32_t *n_jobs);
const b200sdf_curve *vgb_batch_curves(const vgb_batch *b, uint32_t *n_curves);
uint64_t vgb_batch_total_segments(const vgb_batch *b);   /* flattened segments of all glyphs, either source */
uint32_t vgb_batch_fallback_glyphs(const vgb_batch *b);  /* device-flatten glyphs that had to be flattened on the host */
const uint8_t *vgb_batch_bitmaps(const vgb_batch *b, uint64_t *bytes);
uint64_t vgb_batch_pairs(const vgb_batch *b);
int vgb_renderer_render_batch(const vgb_renderer *r, vgb_batch *b);
int vgb_renderer_submit_batch(const vgb_renderer *r, vgb_batch *b, uint64_t *ticket);
int vgb_renderer_wait_batch(const vgb_renderer *r, uint64_t ticket);
/* Two-step submission (pipelines with one CUDA thread): prepare in any thread (bitmap buffer, job validation,
 * tile planning), then vgb_renderer_submit_batch only enqueues.  poll: 1 = finished (ticket consumed like
 * wait), 0 = still running, < 0 = error. */
int vgb_renderer_prepare_batch(const vgb_renderer *r, vgb_batch *b);
int vgb_renderer_poll_batch(const vgb_renderer *r, uint64_t ticket);
/* Glyph-level batches (mode 2): after wait / poll reported the batch finished, take frames and bitmap presence from
 * the device's answers; glyphs the device handed back are recorded on the host and rendered now (blocking).
 * vgb_renderer_render_batch does this itself.  No-op for the other modes. */
int vgb_batch_finalize(const vgb_renderer *r, vgb_batch *b);
/* The request arrays of a glyph-level batch (what b200sdf_submit_glyphs receives) */
const b200sdf_glyph_req *vgb_batch_requests(const vgb_batch *b, uint32_t *n);
const b200sdf_glyph_part *vgb_batch_parts(const vgb_batch *b, uint32_t *n);
uint32_t vgb_batch_curve_slots(const vgb_batch *b);
uint32_t vgb_batch_tile_cap(const vgb_batch *b);
uint64_t vgb_batch_est_cost(const vgb_batch *b); /* the est_cost argument of b200sdf_submit_glyphs */
uint32_t vgb_batch_handed_back(const vgb_batch *b);
/* glyphs with cubic curves sent as kind PATH (flattened by the device) */
uint32_t vgb_batch_path_glyphs(const vgb_batch *b);
/* bitmap of glyph i (NULL when it has none); valid after the batch was rendered */
const uint8_t *vgb_batch_glyph_bitmap(const vgb_batch *b, uint32_t i, uint64_t *len);

/* ---- Writer ---- */
vgb_writer *vgb_writer_new_file(const char *folder); /* Writer::new_file, writer/mod.rs:35 */
vgb_writer *vgb_writer_new_memory(void);             /* Writer::new_dummy + contents, writer/mod.rs:44 */
vgb_writer *vgb_writer_new_tar(const char *path);    /* Writer::new_tar over a file, writer/mod.rs:27-33, writer/tar.rs */
vgb_writer *vgb_writer_new_tar_memory(void);         /* TarWriter over a Vec<u8> (tar.rs tests) */
int vgb_writer_write_file(vgb_writer *w, const char *filename, const uint8_t *bytes, uint64_t len); /* mod.rs:51-56 */
int vgb_writer_write_directory(vgb_writer *w, const char *dirname);                                /* mod.rs:58-63 */
int vgb_writer_finish(vgb_writer *w);                                                               /* mod.rs:67-73 */
const uint8_t *vgb_writer_tar_bytes(const vgb_writer *w, uint64_t *len); /* the ustar stream of new_tar_memory */
void vgb_writer_free(vgb_writer *w);
uint32_t vgb_writer_entry_count(const vgb_writer *w);
int vgb_writer_entry(const vgb_writer *w, uint32_t i, const char **name, int32_t *is_dir, const uint8_t **bytes,
                     uint64_t *len);

/* ---- FontManager ---- */
typedef struct {
	uint64_t glyphs, bitmaps, pixels, segments, pairs, pbf_bytes, blocks;
	/* host time per phase, ns summed over workers; wall_ns = the whole call */
	uint64_t outline_ns, submit_ns, wait_ns, encode_ns, write_ns, wall_ns;
	uint64_t submits, workers;
	uint64_t handed_back;            /* glyphs the device decoder returned to the host recorder */
	uint64_t h2d_bytes;              /* request / record / segment bytes the device read from host memory */
	uint64_t cost_total, cost_shard; /* estimated cost of the whole job / of this shard (0 when not sharded) */
} vgb_stats;

vgb_manager *vgb_manager_new(int parallel);                                        /* manager.rs:28-33 */
void vgb_manager_free(vgb_manager *m);
int vgb_manager_add_path(vgb_manager *m, const char *path);                        /* manager.rs:39-53 */
int vgb_manager_add_font_with_name(vgb_manager *m, const char *name, const char *const *sources, uint32_t n); /* :66-75 */
/* `scan` of the recurse command (commands/recurse.rs:104-133): font files, fonts.json manifests, recursion */
int vgb_manager_scan(vgb_manager *m, const char *path);
int vgb_manager_add_font_bytes_with_name(vgb_manager *m, const char *name, const uint8_t *data, size_t len);
uint32_t vgb_manager_font_count(const vgb_manager *m);
const char *vgb_manager_font_id(const vgb_manager *m, uint32_t i);
/* metadata.name (name id 1) of the font's files, in the order they were added, '\n'-separated; returns the length
 * needed (excluding NUL) */
size_t vgb_manager_font_file_names(const vgb_manager *m, const char *font_id, char *buf, size_t cap);
/* FontWrapper::get_blocks populations (wrapper.rs:53-76): out[256] = glyphs per block */
int vgb_manager_block_population(const vgb_manager *m, const char *font_id, uint32_t out[256]);
/* GlyphBlock::render (glyph_block.rs:69-80): malloc'd PBF bytes of one block */
int vgb_manager_render_block(const vgb_manager *m, const char *font_id, uint32_t block, const vgb_renderer *r,
                             uint8_t **pbf, uint64_t *len);
/* FontManager::render_glyphs (manager.rs:81-125).  shard/n_shards select every n-th (font, block)
 * task — the multi-GPU sharding; threads = host threads the call may use, the calling thread included
 * (0 = one per core).  With 14 or more the CALLING thread is the pipeline's only CUDA thread and the others record
 * outlines and encode PBFs; with 2..13 every thread does that and whichever is free submits / polls; 1 = the
 * reference's --single-thread, everything on the calling thread. */
int vgb_manager_render_glyphs(const vgb_manager *m, vgb_writer *w, const vgb_renderer *r, uint32_t shard,
                              uint32_t n_shards, int threads, vgb_stats *stats);
/* The shard of every (font, block) task of an n_shards-way run: owner[font * 256 + block], fonts in id order
 * (longest-processing-time-first over per-block cost estimates; SURVEY.md 8(e), task list manager.rs:88-97).
 * loads (optional, n_shards entries): estimated cost per shard.  Returns the number of entries written or < 0. */
int vgb_manager_shard_owners(const vgb_manager *m, uint32_t n_shards, uint16_t *owner, size_t cap, uint64_t *loads);
int vgb_manager_write_index_json(const vgb_manager *m, vgb_writer *w);             /* manager.rs:128-131 */
int vgb_manager_write_families_json(const vgb_manager *m, vgb_writer *w);          /* manager.rs:134-137 */

/* ---- PBF decode (commands/debug.rs:60-79) ---- */
/* Decodes one glyphs PBF; glyphs: malloc'd array (each bitmap malloc'd); returns count or <0. */
int32_t vgb_pbf_decode(const uint8_t *data, size_t len, char *name, size_t name_cap, char *range, size_t range_cap,
                       vgb_glyph **glyphs);
void vgb_glyphs_free(vgb_glyph *glyphs, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
