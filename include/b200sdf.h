/*
 * b200sdf.h — C ABI of the B200-native SDF glyph renderer (libb200sdf.so).
 *
 * This is the drop-in boundary for the per-glyph rendering path of versatiles_glyphs v0.9.1.
 * The reference has no FFI; its only backend switch is the private enum
 *     RendererMode { Precise, Dummy }                    reference src/render/renderer.rs:11-15
 * matched at src/render/renderer.rs:140-143, where both arms have the shape
 *     fn(&mut RenderResult [x0,y0,width,height filled], Rings) -> glyph.bitmap = Some(Vec<u8; W*H>)
 * (src/render/renderer_precise.rs:8, src/render/renderer_dummy.rs:3).  The entry points below
 * are what a third arm `RendererMode::Cuda` binds (INTEGRATION.md shows the Rust side): the
 * per-glyph call becomes "append a job", the per-GlyphBlock call (src/font/glyph_block.rs:69-80)
 * becomes one b200sdf_submit(), and FontManager::render_glyphs (src/font/manager.rs:81-125)
 * keeps several submits in flight on separate CUDA streams.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * B200SDF_E_ code; nothing throws or unwinds across this boundary; the caller owns every buffer.
 * There is NO CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef B200SDF_H
#define B200SDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SDF_ABI_VERSION 1

enum {
	B200SDF_OK = 0,
	B200SDF_E_ARG = -1,     /* bad argument (null pointer, out-of-range job, zero-sized bitmap ...) */
	B200SDF_E_CUDA = -2,    /* a CUDA runtime call failed; see b200sdf_last_error() */
	B200SDF_E_NODEVICE = -3,/* no CUDA device / driver */
	B200SDF_E_NOMEM = -4,   /* host or device allocation failed */
	B200SDF_E_TICKET = -5   /* unknown or already-waited ticket */
};

typedef struct b200sdf_ctx b200sdf_ctx;

/*
 * One flattened outline segment = geometry::Segment (reference src/geometry/segment.rs:9-14) as
 * produced by Rings::get_segments (src/geometry/rings.rs:75-81): consecutive points of one ring,
 * AFTER rings.scale()/translate() (src/render/renderer.rs:122-131), narrowed to f32 AFTER
 * subtracting the glyph's integer origin (RenderResult.x0, .y0), i.e. in pixel units with the
 * bitmap's lower-left corner at (0,0) and pixel centres at (i+0.5, j+0.5).
 */
typedef struct {
	float x0, y0, x1, y1;
} b200sdf_segment;

/*
 * One glyph = one call of renderer_precise (src/render/renderer_precise.rs:8-84).
 * width/height are RenderResult.width/.height (buffer included).  The bitmap is written to
 * out[out_off + (height-1-y)*width + x] — row-major, top row first, exactly the Vec<u8> the
 * reference moves into PbfGlyph.bitmap (renderer_precise.rs:78-83).  out_off may have any
 * alignment; bitmaps must not overlap.
 */
typedef struct {
	uint32_t seg_off; /* first segment of this glyph in the segment array */
	uint32_t seg_cnt; /* number of segments (0 allowed: every pixel is "outside, infinitely far" = 0) */
	uint32_t width;
	uint32_t height;
	uint64_t out_off; /* byte offset of this glyph's bitmap in the output buffer */
} b200sdf_glyph_job;

/*
 * Outline-level input (the seam one step earlier: the RingBuilder callbacks of
 * src/render/ring_builder.rs:67-117 BEFORE flattening).  One record = one line or one quadratic
 * Bezier of a closed ring, control points in FONT UNITS exactly as the outline callbacks deliver
 * them (f32).  The device flattens it the way Ring::add_quadratic_bezier does
 * (src/geometry/ring.rs:119-144): `depth` = k means 2^k segments at t = j/2^k, evaluated in f64.
 * For inputs whose coordinates are small dyadic rationals (every TrueType outline without scaled
 * components) the reference's midpoint recursion is exact in f64 and has uniform depth, so the
 * device points equal the reference's points bit for bit; the host decides k with the reference's
 * own flatness test and falls back to uploading explicit segments otherwise.
 */
typedef struct {
	float sx, sy, cx, cy, ex, ey; /* start, control, end; for a line cx,cy = sx,sy and depth = 0 */
	uint32_t seg_off;             /* index of this record's first segment within its glyph */
	uint32_t depth;               /* k */
} b200sdf_curve;

enum { B200SDF_KIND_CURVES = 0, B200SDF_KIND_SEGMENTS = 1 };

/*
 * One glyph of a mixed batch.  kind = CURVES: src_off/src_cnt index the curve array, the device
 * applies rings.scale(scale); rings.translate((dx, 0)) (src/render/renderer.rs:122-131) in f64 and
 * subtracts the integer origin (x0, y0) before narrowing to f32.  kind = SEGMENTS: src_off/src_cnt
 * index the segment array (already origin-relative pixel units; scale/dx/x0/y0 unused) and
 * seg_cnt = src_cnt.
 */
typedef struct {
	uint32_t kind;
	uint32_t src_off, src_cnt;
	uint32_t seg_cnt; /* total flattened segments of the glyph */
	uint32_t width, height;
	int32_t x0, y0;
	double scale, dx;
	uint64_t out_off;
} b200sdf_outline_job;

/*
 * Device-side work item: a rectangle of pixel tiles of one glyph, rendered by one CTA
 * ("each CTA owns one glyph, or a pixel tile of a large glyph").  Produced by b200sdf_plan_tiles.
 */
typedef struct {
	uint32_t seg_off;       /* first segment (segment array), or first curve record when `job` is set */
	uint32_t seg_cnt;       /* flattened segments of the glyph */
	uint64_t out_off;
	uint16_t width, height; /* whole-glyph bitmap size */
	uint16_t tx0, ty0;      /* first tile column / row (tiles are B200SDF_TILE_W x B200SDF_TILE_H pixels, y upward) */
	uint16_t ntx, nty;      /* tile columns / rows in this rectangle; ntx*nty <= B200SDF_MAX_ITEMS */
	uint32_t job;           /* B200SDF_NO_JOB = raw segments; else index of the b200sdf_outline_job (curves) */
} b200sdf_tile_job;
#define B200SDF_NO_JOB 0xFFFFFFFFu

#ifndef B200SDF_TILE_W
#define B200SDF_TILE_W 4
#endif
#ifndef B200SDF_TILE_H
#define B200SDF_TILE_H 4
#endif
#ifndef B200SDF_MAX_ITEMS
#define B200SDF_MAX_ITEMS 64
#endif
#define B200SDF_MAX_DIM 16384 /* largest accepted glyph width/height in pixels */

/* ---- context --------------------------------------------------------------------------------- */
int b200sdf_abi_version(void);
int b200sdf_device_count(void);
/* One context <-> one GPU.  n_slots = number of batches that may be in flight (one CUDA stream,
 * one set of device buffers each); buffers grow on demand.  Thread-safe: submit/wait may be
 * called concurrently from several host threads (the reference calls render_glyph from rayon
 * workers, src/font/manager.rs:117-118). */
int b200sdf_create(int device, uint32_t n_slots, b200sdf_ctx **out);
void b200sdf_destroy(b200sdf_ctx *ctx);
const char *b200sdf_last_error(const b200sdf_ctx *ctx);
int b200sdf_device(const b200sdf_ctx *ctx);

/* Pinned host memory for segment / bitmap buffers (plain malloc'd memory also works, slower). */
void *b200sdf_alloc_pinned(size_t bytes);
void b200sdf_free_pinned(void *p);

/* ---- host-buffer path (what RendererMode::Cuda calls) ------------------------------------------ */
/* Enqueue one batch: H2D(segments, jobs) -> kernel -> D2H(bitmaps) on the slot's stream.  Returns
 * immediately; `out` is valid after b200sdf_wait(ticket).  Blocks only while all slots are busy. */
int b200sdf_submit(b200sdf_ctx *ctx, const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_glyph_job *jobs,
                   uint32_t n_jobs, uint8_t *out, uint64_t out_bytes, uint64_t *ticket);
int b200sdf_wait(b200sdf_ctx *ctx, uint64_t ticket);
/* Non-blocking form of b200sdf_wait: 1 = the batch has finished (ticket consumed, as by wait),
 * 0 = still running (ticket stays valid), < 0 = error.  Lets ONE thread own all CUDA traffic of a
 * pipeline (submit + completion polling) while the others only produce and consume batches. */
int b200sdf_poll(b200sdf_ctx *ctx, uint64_t ticket);
/* submit + wait */
int b200sdf_render(b200sdf_ctx *ctx, const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_glyph_job *jobs,
                   uint32_t n_jobs, uint8_t *out, uint64_t out_bytes);

/* ---- outline-level path: flattening happens on the device ------------------------------------- */
/* Like b200sdf_submit, but glyphs are given as curve records (kind CURVES) and/or explicit
 * segments (kind SEGMENTS).  H2D shrinks from 16 B per flattened segment to 32 B per curve. */
int b200sdf_submit_outlines(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves,
                            const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_outline_job *jobs,
                            uint32_t n_jobs, uint8_t *out, uint64_t out_bytes, uint64_t *ticket);
/* The same with the tile planning done by the caller beforehand (b200sdf_plan_outline_tiles over the
 * SAME job array; the tile list is trusted, not re-validated): a pipeline can then plan in its worker
 * threads and keep the thread that talks to CUDA down to one copy, one launch and one event. */
int b200sdf_submit_planned(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves,
                           const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_outline_job *jobs,
                           uint32_t n_jobs, const b200sdf_tile_job *tiles, uint32_t n_tiles, uint8_t *out,
                           uint64_t out_bytes, uint64_t *ticket);
/* Device flattening only: writes the f32 origin-relative segments of every CURVES glyph, glyph
 * after glyph in job order, into out_segs (host buffer, n_out = sum of seg_cnt).  Blocking. */
int b200sdf_flatten_outlines(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves,
                             const b200sdf_outline_job *jobs, uint32_t n_jobs, b200sdf_segment *out_segs,
                             uint64_t n_out);

/* ---- device-resident path (segments and bitmaps stay in HBM) ----------------------------------- */
/* Split glyph jobs into CTA work items, largest first.  tiles may be NULL to query the count.
 * pairs (optional) receives sum(width*height*seg_cnt), the algorithmic pixel x segment pairs. */
int b200sdf_plan_tiles(const b200sdf_glyph_job *jobs, uint32_t n_jobs, uint32_t n_seg, uint64_t out_bytes,
                       b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles, uint64_t *pairs);
/* Launch the SDF kernel on `stream` (a cudaStream_t, NULL = legacy default stream) over device
 * pointers.  Asynchronous; exactly one kernel launch. */
int b200sdf_render_device(b200sdf_ctx *ctx, const b200sdf_segment *d_segs, const b200sdf_tile_job *d_tiles,
                          uint32_t n_tiles, uint8_t *d_out, void *stream);
/* Same for a mixed batch: tiles from b200sdf_plan_outline_tiles, d_jobs = the outline jobs in HBM. */
int b200sdf_plan_outline_tiles(const b200sdf_outline_job *jobs, uint32_t n_jobs, uint32_t n_curves, uint32_t n_seg,
                               uint64_t out_bytes, b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles,
                               uint64_t *pairs);
/* The same with planning flags.  B200SDF_PLAN_LATENCY: the caller will wait for this batch alone (the last batches
 * of a pipeline) — heavy glyphs are cut as finely as allowed, which shortens the kernel at the price of repeated
 * staging; without it small batches are planned for throughput. */
#define B200SDF_PLAN_LATENCY 1u
int b200sdf_plan_outline_tiles_ex(const b200sdf_outline_job *jobs, uint32_t n_jobs, uint32_t n_curves, uint32_t n_seg,
                                  uint64_t out_bytes, uint32_t flags, b200sdf_tile_job *tiles, uint32_t cap,
                                  uint32_t *n_tiles, uint64_t *pairs);
int b200sdf_render_outlines_device(b200sdf_ctx *ctx, const b200sdf_curve *d_curves, const b200sdf_segment *d_segs,
                                   const b200sdf_outline_job *d_jobs, const b200sdf_tile_job *d_tiles,
                                   uint32_t n_tiles, uint8_t *d_out, void *stream);

/* ---- measurement helpers ----------------------------------------------------------------------- */
/* Dependent-FFMA-chain microbenchmark: measured FP32 (non-tensor) peak of this device in TFLOP/s
 * (2 flop per FFMA), best of |reps|.  Roofline denominator for the SDF kernel.  reps < 0 runs the
 * packed FFMA2 variant (two FMAs per lane per instruction) instead. */
int b200sdf_measure_fp32_peak(b200sdf_ctx *ctx, int reps, double *tflops, double *ms);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
uint64_t b200sdf_launch_count(const b200sdf_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
