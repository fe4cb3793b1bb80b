// render.h — host side of the rendering path (mirror of reference src/render/*.rs).
//
//   RingBuilder   = src/render/ring_builder.rs        outline callbacks -> flattened closed rings
//   RenderResult  = src/render/result.rs              integer frame of a glyph (+ bitmap)
//   Renderer      = src/render/renderer.rs            render_glyph(face, codepoint) -> Option<PbfGlyph>
//   GlyphBatch    = NEW: the flat segment buffer of one GlyphBlock, uploaded once (north star)
//
// Everything that decides a METRIC (advance, bbox, floor/ceil, width/height/left/top) is computed
// here in f64 with the reference's operation order; only the per-pixel SDF work goes to the GPU
// through include/b200sdf.h.  There is no CPU rasteriser in this library.
#pragma once

#include <cstdint>
#include <memory>
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

// render/ring_builder.rs:8-117
class RingBuilder : public OutlineBuilder {
  public:
	explicit RingBuilder(RingSet &rings, double precision = 0.01) : rings_(rings), precision_(precision) {}
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

// Growable host buffer, pinned when a CUDA renderer owns it.
class HostBuffer {
  public:
	explicit HostBuffer(bool pinned) : pinned_(pinned) {}
	~HostBuffer();
	HostBuffer(const HostBuffer &) = delete;
	HostBuffer &operator=(const HostBuffer &) = delete;
	// keeps the first `keep` bytes
	bool reserve(size_t bytes, size_t keep);
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
	RenderResult frame;      // y1 already rebased (renderer.rs:146)
	uint32_t job = 0;        // index into jobs when has_bitmap
};

// The flat segment buffer of one GlyphBlock (or of any group of glyphs): segments of all glyphs
// back to back as b200sdf_segment, one b200sdf_glyph_job per bitmap, bitmaps packed back to back.
class GlyphBatch {
  public:
	explicit GlyphBatch(bool pinned);
	void clear();
	// First half of Renderer::render_glyph (renderer.rs:103-137): cmap lookup, outline, flatten,
	// advance, scale+shift, integer frame; appends the glyph's segments.  Returns false for
	// "None" (code point not a char / not in the font), true otherwise.
	bool add_glyph(const Face &face, uint32_t codepoint);
	// Same from explicit rings already in pixel space (renderer_precise's own signature; used by
	// tests with synthetic outlines).  Always has a bitmap.
	bool add_rings(uint32_t id, uint32_t advance, const RenderResult &frame, const RingSet &rings);

	const std::vector<BatchGlyph> &glyphs() const { return glyphs_; }
	const std::vector<b200sdf_glyph_job> &jobs() const { return jobs_; }
	const b200sdf_segment *segments() const { return reinterpret_cast<const b200sdf_segment *>(segs_.data()); }
	uint32_t segment_count() const { return n_seg_; }
	uint8_t *bitmaps() { return out_.data(); }
	const uint8_t *bitmaps() const { return out_.data(); }
	uint64_t bitmap_bytes() const { return out_bytes_; }
	uint64_t pairs() const { return pairs_; }
	bool ensure_output(); // allocate the bitmap area (after the last add)
	// PbfGlyph i with its bitmap copied out of the batch (valid after the batch was rendered)
	PbfGlyph take_glyph(size_t i) const;

  private:
	bool append_segments(const RingSet &rings, double ox, double oy);
	RingSet scratch_;
	std::vector<BatchGlyph> glyphs_;
	std::vector<b200sdf_glyph_job> jobs_;
	HostBuffer segs_;
	HostBuffer out_;
	uint32_t n_seg_ = 0;
	uint64_t out_bytes_ = 0;
	uint64_t pairs_ = 0;
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
	std::unique_ptr<GlyphBatch> new_batch() const { return std::make_unique<GlyphBatch>(mode_ == Mode::Cuda); }

	// renderer.rs:103-149 — a batch of one.  nullopt = None.
	std::optional<PbfGlyph> render_glyph(const Face &face, uint32_t index, std::string *err = nullptr) const;
	// Render every bitmap of the batch (blocking).  false on error (message in *err).
	bool render_batch(GlyphBatch &batch, std::string *err = nullptr) const;
	// Asynchronous form: several batches in flight on the context's streams.
	bool submit_batch(GlyphBatch &batch, uint64_t *ticket, std::string *err = nullptr) const;
	bool wait_batch(uint64_t ticket, std::string *err = nullptr) const;

  private:
	Renderer() = default;
	Mode mode_ = Mode::Dummy;
	b200sdf_ctx *ctx_ = nullptr;
	uint32_t slots_ = 0;
};

} // namespace vgb
