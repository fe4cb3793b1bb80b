// render.h — host side of the rendering path (mirror of reference src/render/*.rs).
//
//   RingBuilder     = src/render/ring_builder.rs      outline callbacks -> flattened closed rings (literal)
//   OutlineRecorder = the same callbacks recorded as curve records for on-device flattening,
//                     with the exact glyph bounding box computed analytically
//   RenderResult    = src/render/result.rs            integer frame of a glyph (+ bitmap)
//   Renderer        = src/render/renderer.rs          render_glyph(face, codepoint) -> Option<PbfGlyph>
//   GlyphBatch      = NEW: the flat outline buffer of one GlyphBlock, uploaded once (north star)
//
// Everything that decides a METRIC (advance, bbox, floor/ceil, width/height/left/top) is computed
// here in f64 with the reference's operation order; only the per-pixel SDF work (and, in outline
// mode, the flattening that feeds it) goes to the GPU through include/b200sdf.h.  There is no CPU
// rasteriser in this library.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include "../../../include/b200sdf.h"
#include "face.h"
#include "geometry.h"

namespace vgb {

// render/mod.rs:52-68
constexpr int32_t GLYPH_SIZE = 24;
constexpr int32_t BUFFER = 3;
constexpr double SDF_RADIUS = 8.0;
constexpr double CUTOFF = 0.25 * 256.0;
constexpr double PRECISION = 0.01; // ring_builder.rs:62 (tolerance_sq, font units squared)

// protobuf/glyph.rs:10-41
struct PbfGlyph {
	uint32_t id = 0;
	bool has_bitmap = false;
	std::vector<uint8_t> bitmap; // (width+6)*(height+6) bytes when has_bitmap
	uint32_t width = 0, height = 0;
	int32_t left = 0, top = 0;
	uint32_t advance = 0;
	// glyph.rs:60-70
	static PbfGlyph empty(uint32_t id, uint32_t advance)
	{
		PbfGlyph g;
		g.id = id;
		g.advance = advance;
		return g;
	}
};

// render/result.rs:7-29
struct RenderResult {
	int32_t x0 = 0, x1 = 0, y0 = 0, y1 = 0;
	uint32_t width = 0, height = 0;
	// result.rs:66-76 (bitmap attached by the caller)
	PbfGlyph into_pbf_glyph(uint32_t id, uint32_t advance) const
	{
		PbfGlyph g;
		g.id = id;
		g.has_bitmap = true;
		g.width = width - 2 * (uint32_t)BUFFER;
		g.height = height - 2 * (uint32_t)BUFFER;
		g.left = x0 + BUFFER;
		g.top = y1 - BUFFER;
		g.advance = advance;
		return g;
	}
};

// render/ring_builder.rs:8-117 — literal flattening on the host
class RingBuilder : public OutlineBuilder {
  public:
	explicit RingBuilder(RingSet &rings, double precision = PRECISION) : rings_(rings), precision_(precision) {}
	void move_to(float x, float y) override;
	void line_to(float x, float y) override;
	void quad_to(float x1, float y1, float x, float y) override;
	void curve_to(float x1, float y1, float x2, float y2, float x, float y) override;
	void close() override;
	void finish(); // into_rings(): saves the ring under construction
  private:
	void save_ring();
	RingSet &rings_;
	double precision_;
};

// The same callbacks, recorded instead of flattened.  For outlines whose coordinates are small
// dyadic rationals (|v| <= 2^15, v * 2^10 integral — every TrueType glyph without scaled
// components), Ring::add_quadratic_bezier's midpoint recursion (ring.rs:119-144) is exact in f64:
// the second difference quarters at every level, so the recursion depth k is uniform and the
// emitted points are exactly B(j / 2^k).  Then
//   * k follows from the reference's own flatness test evaluated at the root,
//   * the bounding box of the flattened points is the box of the end points plus the grid points
//     next to each coordinate's vertex (a quadratic is monotone either side of it),
//   * ring bookkeeping (save_ring's 3 / 4 point rules, Ring::close) needs only point counts,
// and the device can regenerate every flattened point bit-for-bit.  Anything else (non-dyadic
// coordinates, cubic curves, absurd depth) clears exact() and the caller flattens literally.
class OutlineRecorder : public OutlineBuilder {
  public:
	void begin();
	void move_to(float x, float y) override;
	void line_to(float x, float y) override;
	void quad_to(float x1, float y1, float x, float y) override;
	void curve_to(float x1, float y1, float x2, float y2, float x, float y) override;
	void close() override;
	void finish();

	// Cubic curves (CFF outlines): recorded as head + tail records for the device's literal subdivision (kind PATH,
	// include/b200sdf.h) when allowed; otherwise a cubic makes the glyph inexact (flattened on the host).
	void allow_cubics(bool on) { allow_cubics_ = on; }
	bool has_cubic() const { return has_cubic_; }
	bool exact() const { return exact_; }
	bool is_empty() const { return rings_ == 0; } // Rings::is_empty
	const std::vector<b200sdf_curve> &records() const { return recs_; }
	uint32_t segment_count() const { return n_seg_; }
	const BBox &bbox() const { return bbox_; } // font units, over every flattened point of the kept rings

  private:
	struct P {
		float x, y;
	};
	void add_line(P a, P b);
	void save_ring();
	void axis_extrema(double s, double c, double e, uint32_t k, double &lo, double &hi) const;

	std::vector<b200sdf_curve> recs_;
	size_t ring_first_rec_ = 0;
	uint32_t ring_points_ = 0;
	P ring_first_{0, 0}, ring_last_{0, 0};
	BBox ring_bbox_;
	BBox bbox_;
	uint32_t n_seg_ = 0;
	size_t rings_ = 0;
	bool exact_ = true;
	bool allow_cubics_ = false, has_cubic_ = false;
};

// Growable host buffer, pinned when a CUDA renderer owns it.
class HostBuffer {
  public:
	explicit HostBuffer(bool pinned) : pinned_(pinned) {}
	~HostBuffer();
	HostBuffer(const HostBuffer &) = delete;
	HostBuffer &operator=(const HostBuffer &) = delete;
	// keeps the first `keep` bytes
	bool reserve(size_t bytes, size_t keep, bool exact = false);
	uint8_t *data() { return p_; }
	const uint8_t *data() const { return p_; }
	size_t capacity() const { return cap_; }

  private:
	bool pinned_;
	uint8_t *p_ = nullptr;
	size_t cap_ = 0;
};

// One glyph of a batch: everything Renderer::render_glyph knows before/after the SDF pass.
struct BatchGlyph {
	uint32_t id = 0;
	uint32_t advance = 0;
	bool has_bitmap = false; // false = PbfGlyph::empty
	bool pending = false;    // Glyf mode: frame and has_bitmap come from the device (GlyphBatch::finalize)
	RenderResult frame;      // y1 already rebased (renderer.rs:146)
	uint32_t job = 0;        // index into jobs when has_bitmap (or pending)
	const Face *face = nullptr; // Glyf mode: where the glyph came from (for glyphs the device hands back)
	int64_t extra_off = -1;  // >= 0: the bitmap lives in the batch's side buffer (glyph re-rendered by finalize)
};

// Where outline decoding and flattening happen for a batch.
enum class Flatten {
	Device, // record outlines on the host, upload curve records, flatten on the GPU (glyphs that are not exactly representable fall back per glyph)
	Host,   // flatten on the host, upload b200sdf_segment (the literal renderer_precise seam)
	Glyf    // send glyf record references; the GPU decodes, records, measures, plans and renders (csrc/glyf_kernel.cuh);
	        // glyphs it cannot take (scaled components, CFF) are recorded on the host as in Device mode
};
class Renderer;

// The flat outline buffer of one GlyphBlock (or of any group of glyphs): curve records and/or
// segments of all glyphs back to back, one b200sdf_outline_job per bitmap, bitmaps packed back to back.
class GlyphBatch {
  public:
	GlyphBatch(bool pinned, Flatten mode, const Renderer *owner = nullptr);
	void clear();
	Flatten mode() const { return mode_; }
	// First half of Renderer::render_glyph (renderer.rs:103-137): cmap lookup, outline, advance,
	// scale+shift, integer frame; appends the glyph's curve records or segments.  Returns false for
	// "None" (code point not a char / not in the font), true otherwise.
	bool add_glyph(const Face &face, uint32_t codepoint);
	// From explicit rings already in pixel space (renderer_precise's own signature; used by tests
	// with synthetic outlines).  Always has a bitmap, always uploaded as segments.
	bool add_rings(uint32_t id, uint32_t advance, const RenderResult &frame, const RingSet &rings);

	const std::vector<BatchGlyph> &glyphs() const { return glyphs_; }
	const b200sdf_outline_job *jobs() const { return reinterpret_cast<const b200sdf_outline_job *>(jobs_.data()); }
	uint32_t job_count() const { return n_jobs_; }
	const b200sdf_segment *segments() const { return reinterpret_cast<const b200sdf_segment *>(segs_.data()); }
	uint32_t segment_count() const { return n_seg_; }       // host-flattened segments uploaded as such
	const b200sdf_curve *curves() const { return reinterpret_cast<const b200sdf_curve *>(curves_.data()); }
	uint32_t curve_count() const { return n_curves_; }
	uint64_t total_segments() const { return total_seg_; }  // flattened segments of all glyphs (either source)
	uint32_t fallback_glyphs() const { return n_fallback_; } // Device-mode glyphs that had to be flattened on the host
	uint8_t *bitmaps() { return out_.data(); }
	const uint8_t *bitmaps() const { return out_.data(); }
	uint64_t bitmap_bytes() const { return out_bytes_; } // extent of the bitmap area (Glyf mode: slots, not pixels)
	uint64_t pixel_count() const { return pixels_; }     // bitmap pixels (Glyf mode: known after finalize)
	uint64_t pairs() const { return pairs_; }
	// where glyph b's bitmap is (valid after the batch was rendered and, in Glyf mode, finalized)
	const uint8_t *bitmap_of(const BatchGlyph &b) const
	{
		return b.extra_off >= 0 ? extra_.data() + b.extra_off : out_.data() + jobs()[b.job].out_off;
	}
	// ---- Glyf mode ----
	const b200sdf_glyph_req *reqs() const { return reinterpret_cast<const b200sdf_glyph_req *>(reqs_.data()); }
	const b200sdf_glyph_part *parts() const { return reinterpret_cast<const b200sdf_glyph_part *>(parts_.data()); }
	b200sdf_glyph_frame *frames() { return reinterpret_cast<b200sdf_glyph_frame *>(frames_.data()); }
	const b200sdf_glyph_frame *frames() const { return reinterpret_cast<const b200sdf_glyph_frame *>(frames_.data()); }
	uint32_t part_count() const { return n_parts_; }
	uint32_t curve_slots() const { return curve_slots_; }
	uint32_t generated_segment_slots() const { return gen_seg_slots_; }
	uint32_t path_glyphs() const { return n_path_; } // glyphs with cubic curves flattened by the device (kind PATH)
	uint32_t tile_cap() const { return tile_cap_; }
	// estimated tile x segment units of the work the device renders together with this batch (b200sdf_submit_glyphs
	// est_cost): the batch itself, or — when a pipeline keeps many batches of one job in flight — that job
	uint64_t est_cost() const { return est_cost_ > cost_context_ ? est_cost_ : cost_context_; }
	void set_cost_context(uint64_t units) { cost_context_ = units; }
	bool ensure_frames(); // allocate the frame array (after the last add)
	// After the batch came back: take frames and bitmap presence from the device's answers; glyphs it handed back
	// (B200SDF_GLYPH_NEEDS_HOST) are recorded on the host and rendered through `renderer` now.  false + *err on failure.
	bool finalize(const Renderer &renderer, std::string *err);
	bool finalized() const { return finalized_; }
	// add_glyph answers false both for "None" (the code point has no glyph: skipped, glyph_block.rs:74-76) and when the
	// batch could not take the glyph (allocation failure): the second case is sticky and reported here, so callers that
	// treat false as "skip" still notice that the batch is incomplete.
	bool failed() const { return failed_; }
	const char *failure() const { return failure_; }
	uint32_t handed_back() const { return n_handed_back_; }
	// bytes of this batch the device reads from host memory (requests + parts, or job records + tile list; curve
	// records and segments of glyphs recorded on the host)
	uint64_t upload_bytes() const
	{
		const uint64_t recorded = (uint64_t)n_curves_ * sizeof(b200sdf_curve) + (uint64_t)n_seg_ * sizeof(b200sdf_segment);
		if (mode_ == Flatten::Glyf)
			return recorded + (uint64_t)n_jobs_ * sizeof(b200sdf_glyph_req) + (uint64_t)n_parts_ * sizeof(b200sdf_glyph_part);
		return recorded + (uint64_t)n_jobs_ * sizeof(b200sdf_outline_job) + (uint64_t)n_tiles_ * sizeof(b200sdf_tile_job);
	}
	bool ensure_output(); // allocate the bitmap area (after the last add)
	// plan the CTA work items of the recorded jobs into the batch's own (pinned) tile buffer; false + *why
	// when a job is invalid.  After this, tiles()/tile_count() feed b200sdf_submit_planned.
	bool plan_tiles(const char **why, bool latency = false);
	bool prepared() const { return prepared_; }
	const b200sdf_tile_job *tiles() const { return reinterpret_cast<const b200sdf_tile_job *>(tiles_.data()); }
	uint32_t tile_count() const { return n_tiles_; }
	// buffer capacities in bytes (jobs, segments, curves, bitmaps) — the pool sizes new leases from them
	static constexpr int kBuffers = 8; // jobs, segments, curves, bitmaps, tile list, requests, parts, frames
	void capacities(size_t caps[kBuffers]) const;
	void reserve_capacity(const size_t caps[kBuffers]);
	// PbfGlyph i with its bitmap copied out of the batch (valid after the batch was rendered)
	PbfGlyph take_glyph(size_t i) const;

  private:
	bool append_segments(const RingSet &rings, double ox, double oy);
	bool push_job(const b200sdf_outline_job &j);
	bool add_flattened(uint32_t index, uint32_t advance, double advance_float, double scale);
	bool add_glyf_request(const Face &face, uint32_t index, uint32_t advance, double advance_float, double scale);
	bool push_req(const b200sdf_glyph_req &r);
	Flatten mode_;
	const Renderer *owner_ = nullptr;
	HostBuffer reqs_, parts_, frames_;
	std::vector<Face::GlyfPart> parts_tmp_;
	std::vector<uint8_t> extra_;
	uint32_t n_parts_ = 0, curve_slots_ = 0, tile_cap_ = 0, n_handed_back_ = 0;
	uint32_t gen_seg_slots_ = 0, n_path_ = 0; // kind PATH: room for the segments the device generates / such glyphs
	uint64_t pixels_ = 0, est_cost_ = 0, cost_context_ = 0;
	// Glyf mode: requests of glyphs with many outline points are kept at the front of the request array — the decode
	// kernel takes requests in order, one warp each, and a glyph of several hundred points is that kernel's critical path
	std::vector<uint32_t> job_glyph_; // request index -> glyph index
	uint32_t n_heavy_ = 0;
	void move_to_front(uint32_t job);
	bool finalized_ = false;
	bool failed_ = false;
	const char *failure_ = "";
	bool fail_alloc()
	{
		failed_ = true;
		failure_ = "out of (pinned) host memory while recording a glyph";
		return false;
	}
	RingSet scratch_;
	OutlineRecorder recorder_;
	std::vector<BatchGlyph> glyphs_;
	HostBuffer jobs_, segs_, curves_, out_;
	HostBuffer tiles_;       // CTA work items planned by prepare() (pipeline: planned in the worker thread)
	uint32_t n_tiles_ = 0;
	bool prepared_ = false;
	uint32_t n_jobs_ = 0, n_seg_ = 0, n_curves_ = 0, n_fallback_ = 0;
	uint64_t total_seg_ = 0, out_bytes_ = 0, pairs_ = 0;
};

// render/renderer.rs:17-43.  `Precise` in the reference is the CPU loop; here the precise
// renderer IS the CUDA kernel (RendererMode::Cuda), and `Dummy` stays the zero-bitmap fake the
// reference uses in its pipeline tests (renderer_dummy.rs:3-5).
class Renderer {
  public:
	enum class Mode { Cuda, Dummy };
	// n_slots = batches in flight on the GPU (0 = 2 x host cores, capped at 64)
	static std::unique_ptr<Renderer> create(bool dummy, int device = 0, uint32_t n_slots = 0, std::string *err = nullptr);
	static std::unique_ptr<Renderer> new_precise(int device = 0, uint32_t n_slots = 0, std::string *err = nullptr);
	static std::unique_ptr<Renderer> new_dummy();
	~Renderer();

	Mode mode() const { return mode_; }
	uint32_t slots() const { return slots_; }
	b200sdf_ctx *context() const { return ctx_; }
	Flatten flatten() const { return flatten_; }
	void set_flatten(Flatten f) { flatten_ = f; }
	std::unique_ptr<GlyphBatch> new_batch() const { return std::make_unique<GlyphBatch>(mode_ == Mode::Cuda, flatten_, this); }
	// handle of the face's glyf table on this renderer's device (uploaded on first use: b200sdf_font_upload); false on error
	bool font_handle(const Face &face, uint32_t *handle) const;
	// Batch pool: pinned buffers are expensive to allocate, so the pipeline recycles batches.
	// in_pipeline: the batch counts towards the per-call total that sizes the pool (top_up_pool)
	std::unique_ptr<GlyphBatch> acquire_batch(bool in_pipeline = false) const;
	void release_batch(std::unique_ptr<GlyphBatch> b) const;
	// Called when a render_glyphs call is over: brings the pool to twice the largest number of batches that were
	// ever out at once — or, while that is affordable, to 5/4 of the batches one call uses in total, which makes
	// a later shortage impossible whatever the timing — every one sized to the marks, so that later calls neither
	// create a batch nor grow one in the middle of their pipeline (a pinned allocation stalls the whole CUDA context, and far longer when eight
	// ranks allocate on one host).
	void top_up_pool() const;

	// renderer.rs:103-149 — a batch of one.  nullopt = None.
	std::optional<PbfGlyph> render_glyph(const Face &face, uint32_t index, std::string *err = nullptr) const;
	// Render every bitmap of the batch (blocking).  false on error (message in *err).
	bool render_batch(GlyphBatch &batch, std::string *err = nullptr) const;
	// Asynchronous form: several batches in flight on the context's streams.
	bool submit_batch(GlyphBatch &batch, uint64_t *ticket, std::string *err = nullptr) const;
	bool wait_batch(uint64_t ticket, std::string *err = nullptr) const;
	// Several prepared glyph-level batches in ONE submission (b200sdf_submit_glyph_batches: one decode launch, one SDF
	// launch, one ticket for all of them); other kinds of batches, or a single one, go through submit_batch.
	static constexpr size_t kMaxGroup = B200SDF_MAX_BATCHES;
	bool submit_batches(GlyphBatch *const *batches, size_t n, uint64_t *ticket, std::string *err = nullptr) const;
	// Two-step submission for pipelines with one CUDA thread: prepare_batch (any thread: sizes the bitmap
	// buffer, validates the jobs and plans the tiles) then submit_batch (enqueue only).
	bool prepare_batch(GlyphBatch &batch, std::string *err = nullptr, bool latency = false) const;
	// non-blocking: *done = the batch has finished (the ticket is consumed, as by wait_batch)
	bool poll_batch(uint64_t ticket, bool *done, std::string *err = nullptr) const;

  private:
	Renderer() = default;
	Mode mode_ = Mode::Dummy;
	b200sdf_ctx *ctx_ = nullptr;
	uint32_t slots_ = 0;
	Flatten flatten_ = Flatten::Device;
	mutable std::mutex pool_mu_;
	mutable std::vector<std::unique_ptr<GlyphBatch>> pool_;
	mutable size_t hwm_[GlyphBatch::kBuffers] = {0, 0, 0, 0, 0, 0, 0, 0}; // largest buffer capacities any batch of this renderer reached
	uint64_t id_ = 0; // distinguishes renderers in Face::device_tag
	mutable std::mutex fonts_mu_;
	mutable std::vector<std::pair<uint64_t, uint32_t>> fonts_; // Face::uid -> handle of its glyf table on ctx_
	mutable size_t out_now_ = 0, out_max_ = 0; // batches handed out and not yet returned; the most that ever were
	mutable size_t acq_call_ = 0, acq_max_ = 0; // batches handed out since the last top-up; the most per call
	// glyph-level batches submitted so far: the largest device needs of ONE batch and the fewest glyphs a batch had — from
	// these top_up_pool bounds the largest merged submission the pipeline can make (b200sdf_reserve_glyphs)
	struct GlyfMarks {
		uint64_t reqs = 0;                                // most glyph requests in one batch
		double segs = 0.0, curve_slots = 0.0, tile_cap = 0.0; // most per request of any batch
	};
	mutable GlyfMarks glyf_marks_;
	mutable size_t pool_target_ = 0;
	mutable uint64_t glyf_group_bound_ = 0; // most glyph requests one merged submission of the pipeline can hold (0: not told)
	void note_glyf_batch(const GlyphBatch &b) const;

  public:
	// The pipeline's own, timing-independent figures for the same purpose: the heaviest TASK (a fixed range of a block's
	// glyphs — the same in every call, unlike the batches the tasks happen to be grouped into) per request, and the
	// most requests a merged submission can hold.
	void note_glyf_density(double segs_per_req, double curve_slots_per_req, double tile_cap_per_req) const;
	void set_glyf_group_bound(uint64_t requests) const;
	// ... and for the pooled (pinned) batch buffers: capacities no batch of this job can exceed (bytes per buffer, in
	// GlyphBatch::capacities order).  Batches differ from call to call (workers claim tasks dynamically); without this
	// a batch a little larger than any before re-sized the whole pool in the middle of a long run (50 pinned
	// allocations, 40 ms).
	void raise_batch_marks(const size_t caps[GlyphBatch::kBuffers]) const;
	// The most batches the pipeline can hold at one time (its back-pressure limit + one per worker): the pool is kept at
	// that size, instead of at a multiple of what past calls happened to use.
	void set_pool_target(size_t batches) const;

  private:
};

} // namespace vgb
