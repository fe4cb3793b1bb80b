// render.cc — RingBuilder, GlyphBatch, Renderer (see render.h for the reference map).
#include "render.h"

#include <algorithm>
#include <cmath>
#include <thread>
#include <cstdlib>
#include <cstring>

namespace vgb {

// ---- RingBuilder (ring_builder.rs) -------------------------------------------------------------------
void RingBuilder::save_ring()
{
	// ring_builder.rs:33-54: < 3 points is not a polygon; close; < 4 after closing is dropped
	if (rings_.open_len() < 3) {
		rings_.open_clear();
		return;
	}
	rings_.open_close();
	if (rings_.open_len() < 4) {
		rings_.open_clear();
		return;
	}
	rings_.open_commit();
}

void RingBuilder::move_to(float x, float y)
{
	save_ring();
	rings_.open_add(Point::from_f32(x, y));
}

void RingBuilder::line_to(float x, float y) { rings_.open_add(Point::from_f32(x, y)); }

void RingBuilder::quad_to(float x1, float y1, float x, float y)
{
	if (rings_.open_len() == 0)
		return; // ring_builder.rs:83
	const Point start = rings_.open_last();
	rings_.open_add_quadratic_bezier(start, Point::from_f32(x1, y1), Point::from_f32(x, y), precision_);
}

void RingBuilder::curve_to(float x1, float y1, float x2, float y2, float x, float y)
{
	if (rings_.open_len() == 0)
		return; // ring_builder.rs:99
	const Point start = rings_.open_last();
	rings_.open_add_cubic_bezier(start, Point::from_f32(x1, y1), Point::from_f32(x2, y2), Point::from_f32(x, y),
	                             precision_);
}

void RingBuilder::close() { save_ring(); }
void RingBuilder::finish() { save_ring(); }

// ---- HostBuffer ------------------------------------------------------------------------------------
HostBuffer::~HostBuffer()
{
	if (!p_)
		return;
	if (pinned_)
		b200sdf_free_pinned(p_);
	else
		std::free(p_);
}

bool HostBuffer::reserve(size_t bytes, size_t keep)
{
	if (bytes <= cap_)
		return true;
	size_t n = cap_ ? cap_ : 4096;
	while (n < bytes)
		n += n / 2 + 4096;
	n = (n + 4095) & ~size_t(4095);
	uint8_t *q = pinned_ ? (uint8_t *)b200sdf_alloc_pinned(n) : (uint8_t *)std::malloc(n);
	if (!q)
		return false;
	if (p_ && keep)
		std::memcpy(q, p_, keep);
	if (p_) {
		if (pinned_)
			b200sdf_free_pinned(p_);
		else
			std::free(p_);
	}
	p_ = q;
	cap_ = n;
	return true;
}

// ---- GlyphBatch ------------------------------------------------------------------------------------
GlyphBatch::GlyphBatch(bool pinned) : segs_(pinned), out_(pinned) {}

void GlyphBatch::clear()
{
	glyphs_.clear();
	jobs_.clear();
	n_seg_ = 0;
	out_bytes_ = 0;
	pairs_ = 0;
}

// Rings::get_segments (rings.rs:75-81) narrowed to f32 relative to the integer origin (ox, oy).
bool GlyphBatch::append_segments(const RingSet &rings, double ox, double oy)
{
	const size_t n = rings.segment_count();
	if (!segs_.reserve(((size_t)n_seg_ + n) * sizeof(b200sdf_segment), (size_t)n_seg_ * sizeof(b200sdf_segment)))
		return false;
	b200sdf_segment *dst = reinterpret_cast<b200sdf_segment *>(segs_.data()) + n_seg_;
	const Point *pts = rings.points();
	for (size_t r = 0; r < rings.ring_count(); ++r) {
		const size_t b = rings.ring_begin(r), e = rings.ring_end(r);
		float px = (float)(pts[b].x - ox), py = (float)(pts[b].y - oy);
		for (size_t i = b + 1; i < e; ++i) {
			const float qx = (float)(pts[i].x - ox), qy = (float)(pts[i].y - oy);
			*dst++ = b200sdf_segment{px, py, qx, qy};
			px = qx;
			py = qy;
		}
	}
	n_seg_ += (uint32_t)n;
	return true;
}

bool GlyphBatch::add_rings(uint32_t id, uint32_t advance, const RenderResult &frame, const RingSet &rings)
{
	b200sdf_glyph_job job;
	job.seg_off = n_seg_;
	job.seg_cnt = (uint32_t)rings.segment_count();
	job.width = frame.width;
	job.height = frame.height;
	job.out_off = out_bytes_;
	if (!append_segments(rings, (double)frame.x0, (double)frame.y0))
		return false;
	BatchGlyph g;
	g.id = id;
	g.advance = advance;
	g.has_bitmap = true;
	g.frame = frame;
	g.job = (uint32_t)jobs_.size();
	jobs_.push_back(job);
	glyphs_.push_back(g);
	out_bytes_ += (uint64_t)frame.width * frame.height;
	pairs_ += (uint64_t)frame.width * frame.height * job.seg_cnt;
	return true;
}

bool GlyphBatch::add_glyph(const Face &face, uint32_t index)
{
	// char::from_u32(index)? — renderer.rs:104
	if ((index >= 0xD800 && index <= 0xDFFF) || index > 0x10FFFF)
		return false;
	const auto glyph_id = face.glyph_index(index); // :106
	if (!glyph_id)
		return false;
	const double scale = (double)GLYPH_SIZE / (double)face.units_per_em(); // :107

	scratch_.clear();
	RingBuilder builder(scratch_);
	face.outline_glyph(*glyph_id, builder); // :109-110
	builder.finish();                       // into_rings, :111

	// :115-116 — (adv * scale) * 0.95, round half away from zero, saturating cast
	const double advance_float = (double)face.glyph_hor_advance(*glyph_id).value_or(0) * scale * 0.95;
	const double rounded = std::round(advance_float);
	const uint32_t advance = rounded <= 0.0 ? 0u : (rounded >= 4294967295.0 ? 4294967295u : (uint32_t)rounded);

	BatchGlyph g;
	g.id = index;
	g.advance = advance;
	if (scratch_.is_empty()) { // :118-120
		glyphs_.push_back(g);
		return true;
	}
	// :122-131 — scale, then shift by half the advance rounding error
	const double dx = ((double)advance - advance_float) / 2.0;
	scratch_.scale_translate(scale, dx, 0.0);

	// prepare_glyph — :64-91
	const BBox bbox = scratch_.get_bbox();
	if (bbox.is_empty()) { // :133-137
		glyphs_.push_back(g);
		return true;
	}
	RenderResult fr;
	fr.x0 = (int32_t)std::floor(bbox.min.x) - BUFFER;
	fr.y0 = (int32_t)std::floor(bbox.min.y) - BUFFER;
	fr.x1 = (int32_t)std::ceil(bbox.max.x) + BUFFER;
	fr.y1 = (int32_t)std::ceil(bbox.max.y) + BUFFER;
	fr.width = (uint32_t)(fr.x1 - fr.x0);
	fr.height = (uint32_t)(fr.y1 - fr.y0);
	if (!add_rings(index, advance, fr, scratch_))
		return false;
	glyphs_.back().frame.y1 -= GLYPH_SIZE; // :146
	return true;
}

bool GlyphBatch::ensure_output() { return out_.reserve((size_t)out_bytes_ + 16, 0); }

PbfGlyph GlyphBatch::take_glyph(size_t i) const
{
	const BatchGlyph &b = glyphs_[i];
	if (!b.has_bitmap)
		return PbfGlyph::empty(b.id, b.advance);
	PbfGlyph g = b.frame.into_pbf_glyph(b.id, b.advance);
	const b200sdf_glyph_job &j = jobs_[b.job];
	const size_t n = (size_t)j.width * j.height;
	g.bitmap.assign(out_.data() + j.out_off, out_.data() + j.out_off + n);
	return g;
}

// ---- Renderer --------------------------------------------------------------------------------------
std::unique_ptr<Renderer> Renderer::create(bool dummy, int device, uint32_t n_slots, std::string *err)
{
	return dummy ? new_dummy() : new_precise(device, n_slots, err);
}

std::unique_ptr<Renderer> Renderer::new_dummy()
{
	std::unique_ptr<Renderer> r(new Renderer());
	r->mode_ = Mode::Dummy;
	return r;
}

std::unique_ptr<Renderer> Renderer::new_precise(int device, uint32_t n_slots, std::string *err)
{
	b200sdf_ctx *ctx = nullptr;
	if (n_slots == 0)
		n_slots = 2 * std::max(1u, std::thread::hardware_concurrency());
	n_slots = std::min(n_slots, 64u);
	const int rc = b200sdf_create(device, n_slots, &ctx);
	if (rc != 0) {
		// No CPU fallback by design: the precise renderer is the CUDA kernel.
		if (err)
			*err = "b200sdf_create failed (code " + std::to_string(rc) + "): a B200 (sm_100) GPU is required";
		return nullptr;
	}
	std::unique_ptr<Renderer> r(new Renderer());
	r->mode_ = Mode::Cuda;
	r->ctx_ = ctx;
	r->slots_ = n_slots;
	return r;
}

Renderer::~Renderer()
{
	if (ctx_)
		b200sdf_destroy(ctx_);
}

bool Renderer::submit_batch(GlyphBatch &batch, uint64_t *ticket, std::string *err) const
{
	if (!batch.ensure_output()) {
		if (err)
			*err = "out of host memory for the bitmap buffer";
		return false;
	}
	if (mode_ == Mode::Dummy) {
		// renderer_dummy.rs:3-5 — zero-filled bitmaps of the right size
		std::memset(batch.bitmaps(), 0, (size_t)batch.bitmap_bytes());
		*ticket = ~0ull;
		return true;
	}
	const int rc = b200sdf_submit(ctx_, batch.segments(), batch.segment_count(), batch.jobs().data(),
	                              (uint32_t)batch.jobs().size(), batch.bitmaps(), batch.bitmap_bytes(), ticket);
	if (rc != 0) {
		if (err)
			*err = std::string("b200sdf_submit: ") + b200sdf_last_error(ctx_);
		return false;
	}
	return true;
}

bool Renderer::wait_batch(uint64_t ticket, std::string *err) const
{
	if (mode_ == Mode::Dummy)
		return true;
	const int rc = b200sdf_wait(ctx_, ticket);
	if (rc != 0) {
		if (err)
			*err = std::string("b200sdf_wait: ") + b200sdf_last_error(ctx_);
		return false;
	}
	return true;
}

bool Renderer::render_batch(GlyphBatch &batch, std::string *err) const
{
	uint64_t t = 0;
	return submit_batch(batch, &t, err) && wait_batch(t, err);
}

std::optional<PbfGlyph> Renderer::render_glyph(const Face &face, uint32_t index, std::string *err) const
{
	GlyphBatch batch(mode_ == Mode::Cuda);
	if (!batch.add_glyph(face, index))
		return std::nullopt;
	if (!render_batch(batch, err))
		return std::nullopt;
	return batch.take_glyph(0);
}

} // namespace vgb
