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

#define B200SDF_ABI_VERSION 2

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

/*
 * Glyph-level input (the seam one step earlier still: Face::outline_glyph itself, reference
 * src/render/renderer.rs:109-111).  The font's `glyf` table lives in HBM (b200sdf_font_upload); a request names the
 * simple-glyph record(s) of one glyph and the device does what ttf-parser's glyf walker, RingBuilder, Rings::scale /
 * translate and prepare_glyph (renderer.rs:64-91) do: it decodes flags and coordinate deltas, emits the curve
 * records, decides the flattening depth, computes the integer frame, plans the tile jobs and renders.  The host keeps
 * cmap, hmtx, loca and the composite tree (components that are only translated become parts; anything else is
 * recorded on the host and sent as kind CURVES / SEGMENTS in the same batch).
 */
enum { B200SDF_KIND_GLYF = 2, B200SDF_KIND_PATH = 3 };

/*
 * kind PATH: a glyph recorded on the host whose outline has CUBIC curves (CFF fonts).  Its records live in the call's
 * curve array like those of kind CURVES, with one more record form:
 *   cubic head   {sx, sy = start; cx, cy = first control point; ex, ey = second control point; seg_off;
 *                 depth = B200SDF_CURVE_CUBIC | n} — n = segments Ring::add_cubic_bezier (src/geometry/ring.rs:159-187)
 *                 makes of this curve: the host runs the adaptive flatness test to count them (it needs the points for
 *                 the bounding box anyway), the device repeats the subdivision literally, in f64, and writes them
 *   cubic tail   {sx, sy = end point; depth = B200SDF_CURVE_TAIL}, directly after its head; holds no segments.
 * The decode kernel flattens every record of such a glyph (lines, quadratics in closed form, cubics by subdivision)
 * into origin-relative f32 segments in device memory; the SDF kernel reads them like uploaded segments, so nothing is
 * flattened on the host and 64 bytes per cubic cross PCIe instead of 16 per segment.  In the request, curve_off /
 * curve_cap name the glyph's slot in the submission's generated-segment area (curve_cap = seg_cnt).  A glyph whose
 * subdivision does not come out at the host's counts is reported B200SDF_GLYPH_NEEDS_HOST.  Host buffers only:
 * b200sdf_render_glyphs_device and b200sdf_decode_glyphs answer B200SDF_GLYPH_BAD_REQUEST.
 */
#define B200SDF_CURVE_CUBIC 0x80000000u
#define B200SDF_CURVE_TAIL 0x40000000u
#define B200SDF_CUBIC_STACK 24 /* deepest pending-halves stack of the device's subdivision (the host falls back beyond) */

typedef struct {
	uint32_t font;     /* handle from b200sdf_font_upload */
	uint32_t glyf_off; /* byte range of ONE simple glyph record inside that font's glyf table */
	uint32_t glyf_len;
	float ox, oy;      /* translation of the component in font units (0, 0 for a simple glyph) */
} b200sdf_glyph_part;

typedef struct {
	uint32_t kind;             /* B200SDF_KIND_GLYF, or _CURVES / _SEGMENTS for a glyph recorded on the host */
	uint32_t src_off, src_cnt; /* GLYF: parts; CURVES: curve records; SEGMENTS: segments (in the arrays of this call) */
	uint32_t seg_cnt;          /* CURVES / SEGMENTS: flattened segments (GLYF: computed by the device) */
	uint32_t width, height;    /* CURVES / SEGMENTS: the frame (GLYF: computed by the device) */
	int32_t x0, y0;
	double scale, dx;          /* GLYF / CURVES: rings.scale(scale), rings.translate((dx, 0)) */
	uint64_t out_off;          /* where the bitmap goes; GLYF: start of a slot of out_cap bytes */
	uint32_t out_cap;          /* GLYF: bytes reserved for the bitmap (from the glyph header's bounding box) */
	uint32_t curve_off;        /* GLYF / CURVES: first record of this glyph's slot in the device's curve scratch ... */
	uint32_t curve_cap;        /* ... and its size in records (GLYF: >= points of all parts) */
	uint32_t reserved;
} b200sdf_glyph_req;

enum {
	B200SDF_GLYPH_OK = 0,         /* rendered: frame valid, bitmap at out_off, row stride = width */
	B200SDF_GLYPH_EMPTY = 1,      /* no ring survived / empty bounding box: PbfGlyph::empty (renderer.rs:118-120,133-137) */
	B200SDF_GLYPH_NEEDS_HOST = 2, /* not representable by the device decoder (see glyf_kernel.cuh): record it on the host */
	B200SDF_GLYPH_BAD_REQUEST = 3 /* request out of range */
};

typedef struct {
	int32_t x0, y0;         /* RenderResult.x0 / .y0 */
	uint32_t width, height; /* RenderResult.width / .height (buffer included) */
	uint32_t seg_cnt;       /* flattened segments handed to the SDF pass */
	uint32_t status;
} b200sdf_glyph_frame;

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

/* Bring the device buffers of EVERY idle slot to the largest sizes any batch of this context has needed so far.  A slot
 * sizes its buffers when it is first used; a pipeline calls this between jobs so that a slot first needed at a moment of
 * peak concurrency does not allocate device memory in the middle of a job (with several processes on one box such an
 * allocation was measured at 10-70 ms). */
int b200sdf_reserve(b200sdf_ctx *ctx);

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

/* ---- glyph-level path: glyf decoding, metrics and tile planning on the device ------------------ */
/* Make a font's `glyf` table resident in HBM for the lifetime of the context (what FontFileEntry::new does for host
 * memory, reference src/font/file_entry.rs:32-56).  handle indexes b200sdf_glyph_part.font.  Blocking; load time. */
int b200sdf_font_upload(b200sdf_ctx *ctx, const uint8_t *glyf, uint64_t len, uint32_t *handle);
/* Upper bound of the tile jobs the device may plan for a glyph whose frame is at most width x height:
 * tile_cap of a submission = the sum over its requests. */
uint32_t b200sdf_glyph_tile_bound(uint32_t width, uint32_t height);
/* Enqueue one batch of glyph requests: decode + frame + tile planning (one kernel), SDF (one kernel, one CTA per tile job).
 * frames[i] (written by the device) and the bitmaps are valid after b200sdf_wait / b200sdf_poll(ticket).
 * curves / segs: the arrays CURVES / SEGMENTS requests index (may be NULL / 0).  curve_slots = size of the device's
 * curve scratch in records (>= every request's curve_off + curve_cap); tile_cap = capacity of the device's tile
 * list (the sum of b200sdf_glyph_tile_bound over the requests is always enough).  A tile list that turns out too
 * short fails the batch at wait / poll with B200SDF_E_ARG.  est_cost = the caller's estimate of the batch's cost in
 * tile x segment units (sum over glyphs of ceil(W/4) * ceil(H/4) * segments; about 8 segments per outline point):
 * decides how finely heavy glyphs are cut (0 = never below the small-batch floor).  It affects speed only. */
int b200sdf_submit_glyphs(b200sdf_ctx *ctx, const b200sdf_glyph_req *reqs, uint32_t n_reqs, const b200sdf_glyph_part *parts,
                          uint32_t n_parts, const b200sdf_curve *curves, uint32_t n_curves, const b200sdf_segment *segs,
                          uint32_t n_seg, uint32_t curve_slots, uint32_t tile_cap, uint64_t est_cost, b200sdf_glyph_frame *frames,
                          uint8_t *out, uint64_t out_bytes, uint64_t *ticket);
/* Several batches in ONE submission (one decode launch, one SDF launch, one ticket): what a pipeline does with the
 * batches that queued up while the previous submission was being made — a kernel pair over a few hundred glyphs lasts as
 * long as its slowest glyph and its heaviest tile, so many small submissions cost the GPU more than one large one.
 * Every batch keeps its own arrays, frames and bitmap area.  With more than one batch all arrays must come from
 * b200sdf_alloc_pinned (a single batch may live in pageable memory and is then staged).  At most B200SDF_MAX_BATCHES. */
#define B200SDF_MAX_BATCHES 16
typedef struct {
	const b200sdf_glyph_req *reqs;
	uint32_t n_reqs;
	const b200sdf_glyph_part *parts;
	uint32_t n_parts;
	const b200sdf_curve *curves;
	uint32_t n_curves;
	const b200sdf_segment *segs;
	uint32_t n_seg;
	uint32_t curve_slots, tile_cap;
	b200sdf_glyph_frame *frames;
	uint8_t *out;
	uint64_t out_bytes;
} b200sdf_glyph_batch;
/* Pre-size every slot for glyph-level submissions of up to these totals (requests, uploaded + generated segments, curve
 * slots, tile jobs), then b200sdf_reserve.  A pipeline that merges queued batches makes submissions whose size depends
 * on timing; it names the largest one it can make, so that none of them ever grows a device buffer in mid-call. */
int b200sdf_reserve_glyphs(b200sdf_ctx *ctx, uint32_t n_reqs, uint32_t n_seg, uint32_t curve_slots, uint32_t tile_cap);
int b200sdf_submit_glyph_batches(b200sdf_ctx *ctx, const b200sdf_glyph_batch *batches, uint32_t n_batches, uint64_t est_cost,
                                 uint64_t *ticket);
/* The same over device pointers on `stream` (asynchronous, two kernel launches, scratch owned by the context);
 * mid_event (a cudaEvent_t or NULL) is recorded between the decode kernel and the SDF kernel. */
int b200sdf_render_glyphs_device(b200sdf_ctx *ctx, const b200sdf_glyph_req *d_reqs, uint32_t n_reqs,
                                 const b200sdf_glyph_part *d_parts, uint32_t n_parts, const b200sdf_curve *d_curves,
                                 uint32_t n_curves, const b200sdf_segment *d_segs, uint32_t n_seg, uint32_t curve_slots,
                                 uint32_t tile_cap, uint64_t est_cost, b200sdf_glyph_frame *d_frames, uint8_t *d_out,
                                 uint64_t out_bytes, void *stream, void *mid_event);
/* Decode only (tests, diagnostics): frames, the outline job of every request (src_off / src_cnt locate its records
 * in curves_out, which must hold curve_slots records), the number of tile jobs planned (n_tiles_out, may be NULL) and
 * the tile jobs in the order the SDF kernel would claim them (tiles_out[tiles_cap], may be NULL).  Blocking. */
int b200sdf_decode_glyphs(b200sdf_ctx *ctx, const b200sdf_glyph_req *reqs, uint32_t n_reqs, const b200sdf_glyph_part *parts,
                          uint32_t n_parts, const b200sdf_curve *curves, uint32_t n_curves, uint32_t n_seg, uint32_t curve_slots,
                          uint64_t est_cost, b200sdf_glyph_frame *frames, b200sdf_outline_job *jobs_out, b200sdf_curve *curves_out,
                          uint32_t *n_tiles_out, b200sdf_tile_job *tiles_out, uint32_t tiles_cap);

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
