// render.cc — RingBuilder, OutlineRecorder, GlyphBatch, Renderer (see render.h for the reference map).
#include "render.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>

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

// ---- OutlineRecorder -----------------------------------------------------------------------------------
namespace {

constexpr uint32_t MAX_DEPTH = 12;

// |v| <= 2^15 and v is a multiple of 2^-10: with depth <= 12 every intermediate of the midpoint
// recursion needs at most 17 + 10 + 24 = 51 significant bits, i.e. f64 arithmetic is exact.
inline bool dyadic_ok(float v)
{
	if (!(v >= -32768.0f && v <= 32768.0f))
		return false;                 // also rejects NaN
	const float w = v * 1024.0f;      // exact (power of two), |w| <= 2^25 fits int32
	return (float)(int32_t)w == w;
}

inline double lerp_exact(double a, double b, double t) { return a + t * (b - a); }

// B(t) of one coordinate, t = i / 2^k exactly — the same expression the device evaluates (sdf_kernel.cuh curve_point)
inline double curve_coord(double s, double c, double e, double t)
{
	return lerp_exact(lerp_exact(s, c, t), lerp_exact(c, e, t), t);
}

// 2^-k, exact
inline double pow2_neg(uint32_t k)
{
	const uint64_t bits = (uint64_t)(1023 - k) << 52;
	double v;
	std::memcpy(&v, &bits, sizeof(v));
	return v;
}

// PRECISION * 16^j: scaling by a power of two is exact, so `v * 16^-j > PRECISION` <=> `v > kDepthThreshold[j]`
struct DepthThresholds {
	double t[MAX_DEPTH + 1];
	DepthThresholds()
	{
		double v = PRECISION;
		for (uint32_t j = 0; j <= MAX_DEPTH; ++j, v *= 16.0)
			t[j] = v;
	}
};
const DepthThresholds kDepthThreshold;

} // namespace

void OutlineRecorder::begin()
{
	recs_.clear();
	ring_first_rec_ = 0;
	ring_points_ = 0;
	ring_bbox_ = BBox();
	bbox_ = BBox();
	n_seg_ = 0;
	rings_ = 0;
	exact_ = true;
	has_cubic_ = false;
}

void OutlineRecorder::add_line(P a, P b)
{
	b200sdf_curve r;
	r.sx = a.x, r.sy = a.y, r.cx = a.x, r.cy = a.y, r.ex = b.x, r.ey = b.y;
	r.seg_off = 0;
	r.depth = 0;
	recs_.push_back(r);
	ring_bbox_.include_point(Point::from_f32(b.x, b.y));
	ring_points_ += 1;
	ring_last_ = b;
}

void OutlineRecorder::move_to(float x, float y)
{
	save_ring();
	if (!dyadic_ok(x) || !dyadic_ok(y))
		exact_ = false;
	ring_first_ = ring_last_ = P{x, y};
	ring_points_ = 1;
	ring_bbox_.include_point(Point::from_f32(x, y));
}

void OutlineRecorder::line_to(float x, float y)
{
	if (!dyadic_ok(x) || !dyadic_ok(y))
		exact_ = false;
	if (ring_points_ == 0) { // Ring::add_point on an empty ring: the ring starts here (ring_builder.rs:75-77)
		ring_first_ = ring_last_ = P{x, y};
		ring_points_ = 1;
		ring_bbox_.include_point(Point::from_f32(x, y));
		return;
	}
	add_line(ring_last_, P{x, y});
}

// Extremes over the grid points i/2^k, 0 < i < 2^k, of one coordinate of a quadratic: a quadratic is
// monotone either side of its vertex t* = (s-c)/(s-2c+e), so only the grid points next to t* matter.
void OutlineRecorder::axis_extrema(double s, double c, double e, uint32_t k, double &lo, double &hi) const
{
	// control point between the end points: the coordinate is monotone on [0,1], the end points bound it
	// (coordinates passed dyadic_ok: no NaN)
	const double mn = s < e ? s : e, mx = s < e ? e : s;
	if (k == 0 || (c >= mn && c <= mx))
		return;
	const double a = s - c * 2.0 + e;
	if (a == 0.0)
		return;
	const double ts = (s - c) / a;
	if (!(ts > 0.0 && ts < 1.0))
		return;
	const int64_t n = (int64_t)1 << k;
	const double step = pow2_neg(k);
	const int64_t i0 = (int64_t)(ts * (double)n); // ts > 0: truncation is floor
	// one extra point each side absorbs the rounding of t*; candidates outside 1..n-1 are clamped onto grid points
	// of the polyline, which belong into the box anyway
	double l = lo, h = hi;
	for (int64_t d = -1; d <= 2; ++d) {
		int64_t i = i0 + d;
		i = i < 1 ? 1 : (i > n - 1 ? n - 1 : i);
		const double v = curve_coord(s, c, e, (double)i * step);
		l = v < l ? v : l;
		h = v > h ? v : h;
	}
	lo = l, hi = h;
}

void OutlineRecorder::quad_to(float x1, float y1, float x, float y)
{
	if (ring_points_ == 0)
		return; // ring_builder.rs:83
	if (!dyadic_ok(x1) || !dyadic_ok(y1) || !dyadic_ok(x) || !dyadic_ok(y)) {
		exact_ = false;
		return;
	}
	const double sx = ring_last_.x, sy = ring_last_.y, cx = x1, cy = y1, ex = x, ey = y;
	// Ring::add_quadratic_bezier's test at the root (ring.rs:128-131); at depth j both differences are
	// exactly 4^-j times the root's, so the tested value is exactly 16^-j times this one.
	const double dx = sx + ex - cx * 2.0;
	const double dy = sy + ey - cy * 2.0;
	const double v = dx * dx + dy * dy;
	// depth = how often the tested value has to be divided by 16 to pass: counted against the scaled thresholds
	// without a data-dependent loop exit (the depth varies from curve to curve and mispredicts)
	uint32_t k = 0;
	for (uint32_t j = 0; j <= MAX_DEPTH; ++j)
		k += v > kDepthThreshold.t[j] ? 1u : 0u;
	if (k > MAX_DEPTH) {
		exact_ = false;
		return;
	}
	b200sdf_curve r;
	r.sx = ring_last_.x, r.sy = ring_last_.y, r.cx = x1, r.cy = y1, r.ex = x, r.ey = y;
	r.seg_off = 0;
	r.depth = k;
	recs_.push_back(r);
	ring_bbox_.include_point(Point::from_f32(x, y));
	axis_extrema(sx, cx, ex, k, ring_bbox_.min.x, ring_bbox_.max.x);
	axis_extrema(sy, cy, ey, k, ring_bbox_.min.y, ring_bbox_.max.y);
	ring_points_ += 1u << k;
	ring_last_ = P{x, y};
}

void OutlineRecorder::curve_to(float x1, float y1, float x2, float y2, float x, float y)
{
	// Cubic flattening (ring.rs:159-187) is adaptive, not uniform.  The device repeats the subdivision literally (kind
	// PATH); here the same walk only COUNTS the leaves and collects their end points into the bounding box — nothing is
	// stored, nothing is scaled or narrowed.  The arithmetic (f64, unfused) is the device's, operation for operation.
	if (!allow_cubics_) {
		exact_ = false;
		return;
	}
	if (ring_points_ == 0)
		return; // ring_builder.rs:99
	if (!dyadic_ok(x1) || !dyadic_ok(y1) || !dyadic_ok(x2) || !dyadic_ok(y2) || !dyadic_ok(x) || !dyadic_ok(y)) {
		exact_ = false;
		return;
	}
	struct Node {
		double sx, sy, ax, ay, bx, by, ex, ey;
	};
	Node stack[B200SDF_CUBIC_STACK];
	int top = 0;
	stack[top++] = Node{(double)ring_last_.x, (double)ring_last_.y, (double)x1, (double)y1, (double)x2, (double)y2, (double)x, (double)y};
	uint32_t leaves = 0;
	while (top > 0) {
		const Node q = stack[--top];
		const double dx = (q.bx + q.ax) - (q.sx + q.ex);
		const double dy = (q.by + q.ay) - (q.sy + q.ey);
		if (dx * dx + dy * dy <= PRECISION) {
			ring_bbox_.include_point(Point(q.ex, q.ey));
			if (++leaves > 0x00ffffffu) {
				exact_ = false;
				return;
			}
			continue;
		}
		if (top + 2 > B200SDF_CUBIC_STACK) { // deeper than the device's stack: let the host flatten this glyph
			exact_ = false;
			return;
		}
		const double p01x = (q.sx + q.ax) / 2.0, p01y = (q.sy + q.ay) / 2.0;
		const double p12x = (q.ax + q.bx) / 2.0, p12y = (q.ay + q.by) / 2.0;
		const double p23x = (q.bx + q.ex) / 2.0, p23y = (q.by + q.ey) / 2.0;
		const double p012x = (p01x + p12x) / 2.0, p012y = (p01y + p12y) / 2.0;
		const double p123x = (p12x + p23x) / 2.0, p123y = (p12y + p23y) / 2.0;
		const double mx = (p012x + p123x) / 2.0, my = (p012y + p123y) / 2.0;
		stack[top++] = Node{mx, my, p123x, p123y, p23x, p23y, q.ex, q.ey};
		stack[top++] = Node{q.sx, q.sy, p01x, p01y, p012x, p012y, mx, my};
	}
	b200sdf_curve h, t;
	h.sx = ring_last_.x, h.sy = ring_last_.y, h.cx = x1, h.cy = y1, h.ex = x2, h.ey = y2;
	h.seg_off = 0;
	h.depth = B200SDF_CURVE_CUBIC | leaves;
	std::memset(&t, 0, sizeof(t));
	t.sx = x, t.sy = y;
	t.depth = B200SDF_CURVE_TAIL;
	recs_.push_back(h);
	recs_.push_back(t);
	has_cubic_ = true;
	ring_points_ += leaves;
	ring_last_ = P{x, y};
}

void OutlineRecorder::close() { save_ring(); }
void OutlineRecorder::finish() { save_ring(); }

void OutlineRecorder::save_ring()
{
	// ring_builder.rs:33-54 on point counts
	if (ring_points_ >= 3) {
		// Ring::close — ring.rs:53-63
		const double eps = std::numeric_limits<double>::epsilon();
		if (std::fabs((double)ring_first_.x - (double)ring_last_.x) > eps ||
		    std::fabs((double)ring_first_.y - (double)ring_last_.y) > eps)
			add_line(ring_last_, ring_first_);
	}
	if (ring_points_ >= 4) {
		uint32_t off = n_seg_;
		for (size_t i = ring_first_rec_; i < recs_.size(); ++i) {
			recs_[i].seg_off = off;
			const uint32_t d = recs_[i].depth;
			off += (d & B200SDF_CURVE_TAIL) ? 0u : (d & B200SDF_CURVE_CUBIC) ? (d & 0x00ffffffu) : (1u << d);
		}
		n_seg_ = off;
		bbox_.include_point(ring_bbox_.min);
		bbox_.include_point(ring_bbox_.max);
		rings_++;
	} else {
		recs_.resize(ring_first_rec_);
	}
	ring_first_rec_ = recs_.size();
	ring_points_ = 0;
	ring_bbox_ = BBox();
}

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

bool HostBuffer::reserve(size_t bytes, size_t keep, bool exact)
{
	if (bytes <= cap_)
		return true;
	// Pinned allocations are slow and stall the whole CUDA context, so grow them in big steps: a pooled
	// batch reaches its steady-state size after one or two uses.  (exact: sizing to a known capacity.)
	const size_t step = pinned_ ? (size_t)256 << 10 : (size_t)16 << 10;
	// ... and with half as much again as is needed now, so that batches a little bigger than any seen so far
	// (their composition varies from call to call) do not grow the buffer in the middle of a later call.
	const size_t target = bytes + bytes / 2;
	size_t n = cap_ ? cap_ : step;
	while (n < target)
		n += n + step;
	if (exact)
		n = bytes;
	n = (n + 4095) & ~size_t(4095);
	static const bool trace = std::getenv("VGB_ALLOC_TRACE") != nullptr;
	if (trace)
		std::fprintf(stderr, "[vgb alloc] %zu -> %zu bytes (%s%s)\n", cap_, n, pinned_ ? "pinned" : "heap", exact ? ", to the pool's mark" : "");
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
GlyphBatch::GlyphBatch(bool pinned, Flatten mode, const Renderer *owner)
    : mode_(mode), owner_(owner), reqs_(pinned), parts_(pinned), frames_(pinned), jobs_(pinned), segs_(pinned),
      curves_(pinned), out_(pinned), tiles_(pinned)
{
}

void GlyphBatch::clear()
{
	glyphs_.clear();
	n_jobs_ = n_seg_ = n_curves_ = n_fallback_ = 0;
	total_seg_ = out_bytes_ = pairs_ = 0;
	n_tiles_ = 0;
	prepared_ = false;
	n_parts_ = curve_slots_ = tile_cap_ = n_handed_back_ = 0;
	gen_seg_slots_ = n_path_ = 0;
	pixels_ = est_cost_ = cost_context_ = 0;
	job_glyph_.clear();
	n_heavy_ = 0;
	finalized_ = false;
	failed_ = false;
	failure_ = "";
	extra_.clear();
}

bool GlyphBatch::plan_tiles(const char **why, bool latency)
{
	*why = "";
	// most glyphs are one tile job; heavy ones are cut into several: start generous, retry once if short
	uint32_t cap = std::max<uint32_t>(64u, n_jobs_ * 2u + 64u);
	for (int attempt = 0; attempt < 2; ++attempt) {
		if (!tiles_.reserve((size_t)cap * sizeof(b200sdf_tile_job), 0)) {
			*why = "out of host memory for the tile list";
			return false;
		}
		uint32_t n = 0;
		const int rc = b200sdf_plan_outline_tiles_ex(jobs(), n_jobs_, n_curves_, n_seg_, out_bytes_,
		                                             latency ? B200SDF_PLAN_LATENCY : 0u,
		                                             reinterpret_cast<b200sdf_tile_job *>(tiles_.data()), cap, &n, nullptr);
		if (rc != 0 && n <= cap) { // (a short buffer also answers non-zero, with the needed count in n)
			*why = "invalid outline job (b200sdf_plan_outline_tiles)";
			return false;
		}
		if (rc == 0 && n <= cap) {
			n_tiles_ = n;
			prepared_ = true;
			return true;
		}
		cap = n;
	}
	*why = "tile planning did not converge";
	return false;
}

bool GlyphBatch::push_req(const b200sdf_glyph_req &r)
{
	// (one request per job, same index)
	if (!reqs_.reserve(((size_t)n_jobs_ + 1) * sizeof(r), (size_t)n_jobs_ * sizeof(r)))
		return fail_alloc();
	reinterpret_cast<b200sdf_glyph_req *>(reqs_.data())[n_jobs_] = r;
	return true;
}

bool GlyphBatch::push_job(const b200sdf_outline_job &j)
{
	if (!jobs_.reserve(((size_t)n_jobs_ + 1) * sizeof(j), (size_t)n_jobs_ * sizeof(j)))
		return fail_alloc();
	if (mode_ == Flatten::Glyf) {
		// a glyph recorded on the host travels in a glyph-level batch as a CURVES / SEGMENTS request
		b200sdf_glyph_req r;
		std::memset(&r, 0, sizeof(r));
		r.kind = j.kind;
		r.src_off = j.src_off, r.src_cnt = j.src_cnt, r.seg_cnt = j.seg_cnt;
		r.width = j.width, r.height = j.height, r.x0 = j.x0, r.y0 = j.y0;
		r.scale = j.scale, r.dx = j.dx;
		r.out_off = j.out_off;
		r.out_cap = j.width * j.height;
		if (j.kind == B200SDF_KIND_CURVES) {
			r.curve_off = curve_slots_;
			r.curve_cap = j.src_cnt;
			curve_slots_ += j.src_cnt;
		} else if (j.kind == B200SDF_KIND_PATH) { // its slot in the generated-segment area
			r.curve_off = gen_seg_slots_;
			r.curve_cap = j.seg_cnt;
			gen_seg_slots_ += j.seg_cnt;
			n_path_++;
		}
		tile_cap_ += b200sdf_glyph_tile_bound(j.width, j.height);
		est_cost_ += (uint64_t)((j.width + 3) / 4) * ((j.height + 3) / 4) * ((uint64_t)j.seg_cnt + 8);
		if (!push_req(r))
			return false;
		job_glyph_.push_back((uint32_t)glyphs_.size()); // (the glyph is appended right after its job)
	}
	reinterpret_cast<b200sdf_outline_job *>(jobs_.data())[n_jobs_++] = j;
	out_bytes_ += (uint64_t)j.width * j.height;
	pixels_ += (uint64_t)j.width * j.height;
	pairs_ += (uint64_t)j.width * j.height * j.seg_cnt;
	total_seg_ += j.seg_cnt;
	return true;
}

// Rings::get_segments (rings.rs:75-81) narrowed to f32 relative to the integer origin (ox, oy).
bool GlyphBatch::append_segments(const RingSet &rings, double ox, double oy)
{
	const size_t n = rings.segment_count();
	if (!segs_.reserve(((size_t)n_seg_ + n) * sizeof(b200sdf_segment), (size_t)n_seg_ * sizeof(b200sdf_segment)))
		return fail_alloc();
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
	b200sdf_outline_job job;
	std::memset(&job, 0, sizeof(job));
	job.kind = B200SDF_KIND_SEGMENTS;
	job.src_off = n_seg_;
	job.src_cnt = job.seg_cnt = (uint32_t)rings.segment_count();
	job.width = frame.width;
	job.height = frame.height;
	job.x0 = frame.x0;
	job.y0 = frame.y0;
	job.out_off = out_bytes_;
	if (!append_segments(rings, (double)frame.x0, (double)frame.y0))
		return false;
	BatchGlyph g;
	g.id = id;
	g.advance = advance;
	g.has_bitmap = true;
	g.frame = frame;
	g.job = n_jobs_;
	if (!push_job(job))
		return false;
	glyphs_.push_back(g);
	return true;
}

namespace {

// prepare_glyph — renderer.rs:64-91
inline RenderResult frame_of(const BBox &bbox)
{
	RenderResult fr;
	fr.x0 = (int32_t)std::floor(bbox.min.x) - BUFFER;
	fr.y0 = (int32_t)std::floor(bbox.min.y) - BUFFER;
	fr.x1 = (int32_t)std::ceil(bbox.max.x) + BUFFER;
	fr.y1 = (int32_t)std::ceil(bbox.max.y) + BUFFER;
	fr.width = (uint32_t)(fr.x1 - fr.x0);
	fr.height = (uint32_t)(fr.y1 - fr.y0);
	return fr;
}

} // namespace

// Literal host flattening of the glyph whose rings (font units) are in scratch_: renderer.rs:122-146
bool GlyphBatch::add_flattened(uint32_t index, uint32_t advance, double advance_float, double scale)
{
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
	const BBox bbox = scratch_.get_bbox();
	if (bbox.is_empty()) { // :133-137
		glyphs_.push_back(g);
		return true;
	}
	if (!add_rings(index, advance, frame_of(bbox), scratch_))
		return false;
	glyphs_.back().frame.y1 -= GLYPH_SIZE; // :146
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
	// :115-116 — (adv * scale) * 0.95, round half away from zero, saturating cast
	const double advance_float = (double)face.glyph_hor_advance(*glyph_id).value_or(0) * scale * 0.95;
	const double rounded = std::round(advance_float);
	const uint32_t advance = rounded <= 0.0 ? 0u : (rounded >= 4294967295.0 ? 4294967295u : (uint32_t)rounded);

	if (mode_ == Flatten::Glyf) {
		parts_tmp_.clear();
		const Face::GlyfPlan plan = face.glyf_parts(*glyph_id, parts_tmp_);
		if (plan == Face::GlyfPlan::None) { // no outline: rings.is_empty(), :118-120
			BatchGlyph g;
			g.id = index;
			g.advance = advance;
			glyphs_.push_back(g);
			return true;
		}
		if (plan == Face::GlyfPlan::Parts)
			return add_glyf_request(face, index, advance, advance_float, scale);
		// Host: recorded below like in Device mode
	}
	if (mode_ == Flatten::Device || mode_ == Flatten::Glyf) {
		static const bool device_cubics = [] { // VGB_DEVICE_CUBICS=0: cubic outlines are flattened on the host (experiments)
			const char *e = std::getenv("VGB_DEVICE_CUBICS");
			return !(e && e[0] == '0');
		}();
		recorder_.allow_cubics(mode_ == Flatten::Glyf && device_cubics); // kind PATH exists at the glyph-level seam only
		recorder_.begin();
		face.outline_glyph(*glyph_id, recorder_); // :109-110
		recorder_.finish();                       // into_rings, :111
		if (recorder_.exact()) {
			BatchGlyph g;
			g.id = index;
			g.advance = advance;
			if (recorder_.is_empty()) { // :118-120
				glyphs_.push_back(g);
				return true;
			}
			// :122-131 applied to the font-unit box: x*scale then +dx is monotone, so the box of the
			// transformed points is the transformed box
			const double dx = ((double)advance - advance_float) / 2.0;
			BBox bbox;
			Point lo = recorder_.bbox().min, hi = recorder_.bbox().max;
			lo.x *= scale, lo.y *= scale, hi.x *= scale, hi.y *= scale;
			lo.x += dx, lo.y += 0.0, hi.x += dx, hi.y += 0.0;
			bbox.include_point(lo);
			bbox.include_point(hi);
			if (bbox.is_empty()) { // :133-137
				glyphs_.push_back(g);
				return true;
			}
			const RenderResult fr = frame_of(bbox);
			const std::vector<b200sdf_curve> &recs = recorder_.records();
			if (!curves_.reserve(((size_t)n_curves_ + recs.size()) * sizeof(b200sdf_curve), (size_t)n_curves_ * sizeof(b200sdf_curve)))
				return fail_alloc();
			std::memcpy(curves_.data() + (size_t)n_curves_ * sizeof(b200sdf_curve), recs.data(), recs.size() * sizeof(b200sdf_curve));
			b200sdf_outline_job job;
			std::memset(&job, 0, sizeof(job));
			job.kind = recorder_.has_cubic() ? (uint32_t)B200SDF_KIND_PATH : (uint32_t)B200SDF_KIND_CURVES;
			job.src_off = n_curves_;
			job.src_cnt = (uint32_t)recs.size();
			job.seg_cnt = recorder_.segment_count();
			job.width = fr.width;
			job.height = fr.height;
			job.x0 = fr.x0;
			job.y0 = fr.y0;
			job.scale = scale;
			job.dx = dx;
			job.out_off = out_bytes_;
			n_curves_ += (uint32_t)recs.size();
			g.has_bitmap = true;
			g.frame = fr;
			g.frame.y1 -= GLYPH_SIZE; // :146
			g.job = n_jobs_;
			if (!push_job(job))
				return false;
			glyphs_.push_back(g);
			return true;
		}
		n_fallback_++;
	}
	scratch_.clear();
	RingBuilder builder(scratch_);
	face.outline_glyph(*glyph_id, builder); // :109-110
	builder.finish();                       // into_rings, :111
	return add_flattened(index, advance, advance_float, scale);
}

// A glyph the device decodes itself: one request naming its glyf records, a bitmap slot sized from the records'
// header boxes (the device checks that the real frame fits and hands the glyph back otherwise) and a curve slot of
// one record per point.
bool GlyphBatch::add_glyf_request(const Face &face, uint32_t index, uint32_t advance, double advance_float, double scale)
{
	uint32_t font = 0;
	if (!owner_ || !owner_->font_handle(face, &font)) {
		failed_ = true;
		failure_ = "could not make the font's glyf table resident on the device";
		return false;
	}
	const double dx = ((double)advance - advance_float) / 2.0;
	BBox box;
	uint32_t points = 0;
	if (!parts_.reserve(((size_t)n_parts_ + parts_tmp_.size()) * sizeof(b200sdf_glyph_part), (size_t)n_parts_ * sizeof(b200sdf_glyph_part)))
		return fail_alloc();
	b200sdf_glyph_part *dst = reinterpret_cast<b200sdf_glyph_part *>(parts_.data()) + n_parts_;
	for (const Face::GlyfPart &p : parts_tmp_) {
		dst->font = font, dst->glyf_off = p.off, dst->glyf_len = p.len, dst->ox = p.ox, dst->oy = p.oy;
		++dst;
		points += p.points;
		box.include_point(Point((double)p.xmin + (double)p.ox, (double)p.ymin + (double)p.oy));
		box.include_point(Point((double)p.xmax + (double)p.ox, (double)p.ymax + (double)p.oy));
	}
	// the slot: frame_of the (untrusted) header box, one pixel of slack each side, 16-byte aligned start
	double w = std::ceil(box.max.x * scale + dx) - std::floor(box.min.x * scale + dx) + 2.0 * BUFFER + 2.0;
	double h = std::ceil(box.max.y * scale) - std::floor(box.min.y * scale) + 2.0 * BUFFER + 2.0;
	if (!(w >= 8.0))
		w = 8.0;
	if (!(h >= 8.0))
		h = 8.0;
	if (w > 4096.0)
		w = 4096.0; // absurd header boxes: the device hands the glyph back if the real frame is larger
	if (h > 4096.0)
		h = 4096.0;
	const uint32_t wi = (uint32_t)w, hi = (uint32_t)h;
	b200sdf_glyph_req r;
	std::memset(&r, 0, sizeof(r));
	r.kind = B200SDF_KIND_GLYF;
	r.src_off = n_parts_, r.src_cnt = (uint32_t)parts_tmp_.size();
	r.scale = scale, r.dx = dx;
	r.out_off = (out_bytes_ + 15u) & ~(uint64_t)15u;
	r.out_cap = wi * hi;
	r.curve_off = curve_slots_;
	r.curve_cap = points;
	b200sdf_outline_job job; // host-side mirror, completed by finalize()
	std::memset(&job, 0, sizeof(job));
	job.kind = B200SDF_KIND_CURVES;
	job.scale = scale, job.dx = dx;
	job.out_off = r.out_off;
	if (!jobs_.reserve(((size_t)n_jobs_ + 1) * sizeof(job), (size_t)n_jobs_ * sizeof(job)) || !push_req(r))
		return fail_alloc();
	BatchGlyph g;
	g.id = index;
	g.advance = advance;
	g.pending = true;
	g.job = n_jobs_;
	g.face = &face;
	reinterpret_cast<b200sdf_outline_job *>(jobs_.data())[n_jobs_++] = job;
	n_parts_ += (uint32_t)parts_tmp_.size();
	curve_slots_ += points;
	tile_cap_ += b200sdf_glyph_tile_bound(wi, hi);
	est_cost_ += (uint64_t)((wi + 3) / 4) * ((hi + 3) / 4) * ((uint64_t)points * 8 + 8); // ~8 flattened segments per outline point
	out_bytes_ = r.out_off + r.out_cap;
	job_glyph_.push_back((uint32_t)glyphs_.size());
	glyphs_.push_back(g);
	constexpr uint32_t kHeavyPoints = 160;
	if (points >= kHeavyPoints)
		move_to_front(g.job);
	return true;
}

void GlyphBatch::move_to_front(uint32_t job)
{
	const uint32_t to = n_heavy_++;
	if (job == to)
		return;
	b200sdf_glyph_req *rv = reinterpret_cast<b200sdf_glyph_req *>(reqs_.data());
	b200sdf_outline_job *jv = reinterpret_cast<b200sdf_outline_job *>(jobs_.data());
	std::swap(rv[job], rv[to]);
	std::swap(jv[job], jv[to]);
	std::swap(job_glyph_[job], job_glyph_[to]);
	glyphs_[job_glyph_[job]].job = job;
	glyphs_[job_glyph_[to]].job = to;
}

bool GlyphBatch::ensure_output() { return out_.reserve((size_t)out_bytes_ + 16, 0); }
bool GlyphBatch::ensure_frames() { return frames_.reserve(((size_t)n_jobs_ + 1) * sizeof(b200sdf_glyph_frame), 0); }

bool GlyphBatch::finalize(const Renderer &renderer, std::string *err)
{
	if (mode_ != Flatten::Glyf || finalized_)
		return true;
	finalized_ = true;
	b200sdf_outline_job *jv = reinterpret_cast<b200sdf_outline_job *>(jobs_.data());
	const b200sdf_glyph_frame *fv = frames();
	for (BatchGlyph &g : glyphs_) {
		if (!g.pending) {
			// recorded on the host (kinds CURVES / SEGMENTS / PATH): the frame is known, the device must have taken it
			if (g.has_bitmap && g.extra_off < 0 && fv[g.job].status != B200SDF_GLYPH_OK) {
				if (err)
					*err = "the device rejected a host-recorded glyph (U+" + std::to_string(g.id) + ", status " +
					       std::to_string(fv[g.job].status) + ")";
				return false;
			}
			continue;
		}
		g.pending = false;
		const b200sdf_glyph_frame &f = fv[g.job];
		b200sdf_outline_job &j = jv[g.job];
		if (f.status == B200SDF_GLYPH_OK) {
			j.width = f.width, j.height = f.height, j.x0 = f.x0, j.y0 = f.y0, j.seg_cnt = f.seg_cnt;
			g.has_bitmap = true;
			g.frame.x0 = f.x0, g.frame.y0 = f.y0;
			g.frame.x1 = f.x0 + (int32_t)f.width, g.frame.y1 = f.y0 + (int32_t)f.height;
			g.frame.width = f.width, g.frame.height = f.height;
			g.frame.y1 -= GLYPH_SIZE; // renderer.rs:146
			pixels_ += (uint64_t)f.width * f.height;
			pairs_ += (uint64_t)f.width * f.height * f.seg_cnt;
			total_seg_ += f.seg_cnt;
		} else if (f.status == B200SDF_GLYPH_EMPTY) {
			g.has_bitmap = false; // renderer.rs:118-120 / :133-137
		} else if (f.status == B200SDF_GLYPH_NEEDS_HOST) {
			// the closed-form device decoder does not cover this glyph: record it on the host (exact recorder or literal
			// flattening) and render it on its own; the bitmap goes to the side buffer
			n_handed_back_++;
			std::unique_ptr<GlyphBatch> one(new GlyphBatch(false, Flatten::Device, owner_));
			if (!one->add_glyph(*g.face, g.id) || one->glyphs().size() != 1) {
				if (err)
					*err = "could not record a glyph handed back by the device";
				return false;
			}
			const BatchGlyph &o = one->glyphs()[0];
			if (o.has_bitmap) {
				if (!renderer.render_batch(*one, err))
					return false;
				const b200sdf_outline_job &oj = one->jobs()[o.job];
				const size_t n = (size_t)oj.width * oj.height;
				g.extra_off = (int64_t)extra_.size();
				extra_.insert(extra_.end(), one->bitmaps() + oj.out_off, one->bitmaps() + oj.out_off + n);
				j.width = oj.width, j.height = oj.height, j.x0 = oj.x0, j.y0 = oj.y0, j.seg_cnt = oj.seg_cnt;
				pixels_ += n;
				pairs_ += (uint64_t)n * oj.seg_cnt;
				total_seg_ += oj.seg_cnt;
			}
			g.has_bitmap = o.has_bitmap;
			g.frame = o.frame;
		} else {
			if (err)
				*err = "the device rejected a glyph request (status " + std::to_string(f.status) + ")";
			return false;
		}
	}
	return true;
}

void GlyphBatch::capacities(size_t caps[kBuffers]) const
{
	caps[0] = jobs_.capacity(), caps[1] = segs_.capacity(), caps[2] = curves_.capacity(), caps[3] = out_.capacity();
	caps[4] = tiles_.capacity();
	caps[5] = reqs_.capacity(), caps[6] = parts_.capacity(), caps[7] = frames_.capacity();
}

void GlyphBatch::reserve_capacity(const size_t caps[kBuffers])
{
	// only called on an empty batch: nothing to keep
	jobs_.reserve(caps[0], 0, true), segs_.reserve(caps[1], 0, true), curves_.reserve(caps[2], 0, true),
	    out_.reserve(caps[3], 0, true), tiles_.reserve(caps[4], 0, true);
	reqs_.reserve(caps[5], 0, true), parts_.reserve(caps[6], 0, true), frames_.reserve(caps[7], 0, true);
}

PbfGlyph GlyphBatch::take_glyph(size_t i) const
{
	const BatchGlyph &b = glyphs_[i];
	if (!b.has_bitmap)
		return PbfGlyph::empty(b.id, b.advance);
	PbfGlyph g = b.frame.into_pbf_glyph(b.id, b.advance);
	const b200sdf_outline_job &j = jobs()[b.job];
	const size_t n = (size_t)j.width * j.height;
	const uint8_t *src = bitmap_of(b);
	g.bitmap.assign(src, src + n);
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
	static std::atomic<uint64_t> next_id{1};
	r->id_ = next_id.fetch_add(1);
	// default: the device decodes glyf records itself (VGB_FLATTEN=device / host select the older seams: curve records
	// recorded on the host / segments flattened on the host)
	r->flatten_ = Flatten::Glyf;
	if (const char *e = std::getenv("VGB_FLATTEN")) {
		if (!std::strcmp(e, "device"))
			r->flatten_ = Flatten::Device;
		else if (!std::strcmp(e, "host"))
			r->flatten_ = Flatten::Host;
	}
	return r;
}

bool Renderer::font_handle(const Face &face, uint32_t *handle) const
{
	if (mode_ != Mode::Cuda)
		return false;
	const uint64_t tag = face.device_tag();
	if ((tag >> 32) == id_ && (tag & 0xffffffffu)) {
		*handle = (uint32_t)(tag & 0xffffffffu) - 1u;
		return true;
	}
	std::lock_guard<std::mutex> g(fonts_mu_);
	for (const auto &kv : fonts_)
		if (kv.first == face.uid()) {
			*handle = kv.second;
			face.set_device_tag((id_ << 32) | ((uint64_t)kv.second + 1u));
			return true;
		}
	uint32_t h = 0;
	if (b200sdf_font_upload(ctx_, face.glyf_data(), face.glyf_size(), &h) != 0)
		return false;
	fonts_.emplace_back(face.uid(), h);
	face.set_device_tag((id_ << 32) | ((uint64_t)h + 1u));
	*handle = h;
	return true;
}

Renderer::~Renderer()
{
	pool_.clear(); // pinned buffers go before the context
	if (ctx_)
		b200sdf_destroy(ctx_);
}

namespace {
constexpr size_t kPoolMax = 128;  // batches kept when they come back
constexpr size_t kPoolTopUp = 96; // batches allocated ahead of need
}

std::unique_ptr<GlyphBatch> Renderer::acquire_batch(bool in_pipeline) const
{
	std::unique_ptr<GlyphBatch> b;
	size_t caps[GlyphBatch::kBuffers];
	{
		std::lock_guard<std::mutex> g(pool_mu_);
		while (!pool_.empty() && !b) {
			b = std::move(pool_.back());
			pool_.pop_back();
			if (b->mode() != flatten_)
				b.reset();
		}
		for (int i = 0; i < GlyphBatch::kBuffers; ++i)
			caps[i] = hwm_[i];
		out_max_ = std::max(out_max_, ++out_now_);
		acq_call_ += in_pipeline ? 1 : 0;
	}
	if (!b)
		b = new_batch();
	b->clear();
	// Size every buffer to the largest any batch of this renderer ever needed: (pinned) allocations then
	// happen once per pooled batch, up front, instead of whenever a batch first meets a big block.
	b->reserve_capacity(caps);
	return b;
}

void Renderer::release_batch(std::unique_ptr<GlyphBatch> b) const
{
	size_t caps[GlyphBatch::kBuffers];
	b->capacities(caps);
	std::lock_guard<std::mutex> g(pool_mu_);
	for (int i = 0; i < GlyphBatch::kBuffers; ++i)
		hwm_[i] = std::max(hwm_[i], caps[i]);
	if (out_now_ > 0)
		out_now_--;
	if (pool_.size() < kPoolMax)
		pool_.push_back(std::move(b));
}

void Renderer::top_up_pool() const
{
	size_t caps[GlyphBatch::kBuffers], want, have;
	std::vector<std::unique_ptr<GlyphBatch>> mine;
	static const bool trace = std::getenv("VGB_ALLOC_TRACE") != nullptr;
	if (trace)
		std::fprintf(stderr, "[vgb alloc] end of call: pool top-up\n");
	{
		std::lock_guard<std::mutex> g(pool_mu_);
		for (int i = 0; i < GlyphBatch::kBuffers; ++i)
			caps[i] = hwm_[i];
		acq_max_ = std::max(acq_max_, acq_call_);
		acq_call_ = 0;
		have = pool_.size();
		// A call can never run short while the pool holds as many batches as the call hands out in total.  Top up
		// only when that is about to stop being true, and then with room to spare, so that calls which use a batch
		// or two more than any before do not each end with an allocation.
		const size_t floor = std::max(out_max_ + out_max_ / 2, acq_max_ + 2);
		want = have < floor ? std::max(2 * out_max_, acq_max_ + acq_max_ / 2) : have;
		if (pool_target_) // the pipeline's own bound on the batches it can hold at a time: nothing else is ever needed
			want = std::max(have, pool_target_);
		want = std::min(kPoolTopUp, want);
		mine.swap(pool_); // size the pooled batches outside the lock
	}
	if (ctx_) {
		// every slot's device buffers at the marks too: no device allocation in the middle of a later call.  The pipeline
		// merges whatever batches are queued into one submission (at most kMaxGroup of them, and none once it holds 4096
		// glyphs): how large those get depends on timing, so the slots are sized for the largest one possible — as many
		// glyphs as a submission can hold, each as heavy as the heaviest batch's average glyph.
		GlyfMarks m;
		uint64_t bound;
		{
			std::lock_guard<std::mutex> g(pool_mu_);
			m = glyf_marks_;
			bound = glyf_group_bound_;
		}
		if (m.reqs) {
			auto u32 = [](double v) { return (uint32_t)std::min(v + 1.0, 4294967295.0); };
			// (a submission stops growing at 4096 glyphs: it holds less than 4096 + one batch)
			const double reqs = (double)std::max<uint64_t>(bound, std::min<uint64_t>((uint64_t)kMaxGroup * m.reqs, 4096 + m.reqs));
			b200sdf_reserve_glyphs(ctx_, u32(reqs), u32(m.segs * reqs), u32(m.curve_slots * reqs), u32(m.tile_cap * reqs));
		} else {
			b200sdf_reserve(ctx_);
		}
	}
	for (auto &b : mine)
		if (b->mode() == flatten_)
			b->reserve_capacity(caps);
	while (have < want) {
		std::unique_ptr<GlyphBatch> b = new_batch();
		b->reserve_capacity(caps);
		mine.push_back(std::move(b));
		have++;
	}
	std::lock_guard<std::mutex> g(pool_mu_);
	for (auto &b : mine)
		if (pool_.size() < kPoolMax)
			pool_.push_back(std::move(b));
}

namespace {
uint64_t fake_now_ns()
{
	return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
uint64_t fake_latency_ns()
{
	static const uint64_t ns = [] {
		const char *e = std::getenv("VGB_FAKE_LATENCY");
		const long us = e ? std::atol(e) : 0;
		return (uint64_t)(us > 0 ? us : 0) * 1000ull;
	}();
	return ns;
}
} // namespace

bool Renderer::prepare_batch(GlyphBatch &batch, std::string *err, bool latency) const
{
	if (!batch.ensure_output()) {
		if (err)
			*err = "out of host memory for the bitmap buffer";
		return false;
	}
	if (mode_ == Mode::Dummy)
		return true;
	if (batch.mode() == Flatten::Glyf) { // frames come back from the device, which also plans the tiles
		if (!batch.ensure_frames()) {
			if (err)
				*err = "out of host memory for the frame buffer";
			return false;
		}
		return true;
	}
	const char *why = "";
	if (!batch.plan_tiles(&why, latency)) {
		if (err)
			*err = why;
		return false;
	}
	return true;
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
		*ticket = fake_latency_ns() ? fake_now_ns() + fake_latency_ns() : ~0ull; // (test hook, see poll_batch)
		return true;
	}
	if (batch.mode() == Flatten::Glyf) {
		if (!batch.ensure_frames()) {
			if (err)
				*err = "out of host memory for the frame buffer";
			return false;
		}
		note_glyf_batch(batch);
		const int grc = b200sdf_submit_glyphs(ctx_, batch.reqs(), batch.job_count(), batch.parts(), batch.part_count(), batch.curves(),
		                                      batch.curve_count(), batch.segments(), batch.segment_count(), batch.curve_slots(),
		                                      batch.tile_cap(), batch.est_cost(), batch.frames(), batch.bitmaps(), batch.bitmap_bytes(),
		                                      ticket);
		if (grc != 0) {
			if (err)
				*err = std::string("b200sdf_submit_glyphs: ") + b200sdf_last_error(ctx_);
			return false;
		}
		return true;
	}
	const int rc =
	    batch.prepared()
	        ? b200sdf_submit_planned(ctx_, batch.curves(), batch.curve_count(), batch.segments(), batch.segment_count(),
	                                 batch.jobs(), batch.job_count(), batch.tiles(), batch.tile_count(), batch.bitmaps(),
	                                 batch.bitmap_bytes(), ticket)
	        : b200sdf_submit_outlines(ctx_, batch.curves(), batch.curve_count(), batch.segments(), batch.segment_count(),
	                                  batch.jobs(), batch.job_count(), batch.bitmaps(), batch.bitmap_bytes(), ticket);
	if (rc != 0) {
		if (err)
			*err = std::string("b200sdf_submit: ") + b200sdf_last_error(ctx_);
		return false;
	}
	return true;
}

void Renderer::note_glyf_batch(const GlyphBatch &b) const
{
	std::lock_guard<std::mutex> g(pool_mu_);
	const double n = (double)std::max<uint32_t>(1, b.job_count());
	glyf_marks_.reqs = std::max<uint64_t>(glyf_marks_.reqs, b.job_count());
	glyf_marks_.segs = std::max(glyf_marks_.segs, ((double)b.segment_count() + (double)b.generated_segment_slots()) / n);
	glyf_marks_.curve_slots = std::max(glyf_marks_.curve_slots, (double)b.curve_slots() / n);
	glyf_marks_.tile_cap = std::max(glyf_marks_.tile_cap, (double)b.tile_cap() / n);
}

void Renderer::note_glyf_density(double segs_per_req, double curve_slots_per_req, double tile_cap_per_req) const
{
	std::lock_guard<std::mutex> g(pool_mu_);
	glyf_marks_.segs = std::max(glyf_marks_.segs, segs_per_req);
	glyf_marks_.curve_slots = std::max(glyf_marks_.curve_slots, curve_slots_per_req);
	glyf_marks_.tile_cap = std::max(glyf_marks_.tile_cap, tile_cap_per_req);
}

void Renderer::raise_batch_marks(const size_t caps[GlyphBatch::kBuffers]) const
{
	std::lock_guard<std::mutex> g(pool_mu_);
	// (bounds, not needs: a font with absurd header boxes makes them astronomical — past 64 MiB per buffer the pool
	// keeps growing with what batches really use, as it does without this call)
	for (int i = 0; i < GlyphBatch::kBuffers; ++i)
		if (caps[i] <= ((size_t)64 << 20))
			hwm_[i] = std::max(hwm_[i], caps[i]);
}

void Renderer::set_pool_target(size_t batches) const
{
	std::lock_guard<std::mutex> g(pool_mu_);
	pool_target_ = std::max(pool_target_, batches);
}

void Renderer::set_glyf_group_bound(uint64_t requests) const
{
	std::lock_guard<std::mutex> g(pool_mu_);
	glyf_group_bound_ = std::max(glyf_group_bound_, requests);
}

bool Renderer::submit_batches(GlyphBatch *const *batches, size_t n, uint64_t *ticket, std::string *err) const
{
	if (n == 1 || mode_ != Mode::Cuda)
		return n == 1 && submit_batch(*batches[0], ticket, err);
	if (n == 0 || n > kMaxGroup) {
		if (err)
			*err = "submit_batches: between 1 and 16 batches";
		return false;
	}
	b200sdf_glyph_batch desc[kMaxGroup];
	uint64_t est = 0;
	for (size_t k = 0; k < n; ++k) {
		GlyphBatch &b = *batches[k];
		if (b.mode() != Flatten::Glyf || !b.ensure_output() || !b.ensure_frames()) {
			if (err)
				*err = b.mode() != Flatten::Glyf ? "submit_batches: only glyph-level batches can share a submission"
				                                 : "out of host memory for the bitmap / frame buffer";
			return false;
		}
		b200sdf_glyph_batch &d = desc[k];
		d.reqs = b.reqs(), d.n_reqs = b.job_count(), d.parts = b.parts(), d.n_parts = b.part_count();
		d.curves = b.curves(), d.n_curves = b.curve_count(), d.segs = b.segments(), d.n_seg = b.segment_count();
		d.curve_slots = b.curve_slots(), d.tile_cap = b.tile_cap(), d.frames = b.frames(), d.out = b.bitmaps();
		d.out_bytes = b.bitmap_bytes();
		est = std::max(est, b.est_cost());
		note_glyf_batch(b);
	}
	const int rc = b200sdf_submit_glyph_batches(ctx_, desc, (uint32_t)n, est, ticket);
	if (rc != 0) {
		if (err)
			*err = std::string("b200sdf_submit_glyph_batches: ") + b200sdf_last_error(ctx_);
		return false;
	}
	return true;
}

bool Renderer::wait_batch(uint64_t ticket, std::string *err) const
{
	if (mode_ == Mode::Dummy) {
		while (ticket != ~0ull && fake_now_ns() < ticket)
			std::this_thread::yield();
		return true;
	}
	const int rc = b200sdf_wait(ctx_, ticket);
	if (rc != 0) {
		if (err)
			*err = std::string("b200sdf_wait: ") + b200sdf_last_error(ctx_);
		return false;
	}
	return true;
}

bool Renderer::poll_batch(uint64_t ticket, bool *done, std::string *err) const
{
	*done = true;
	if (mode_ == Mode::Dummy) {
		// test hook: VGB_FAKE_LATENCY=<microseconds> makes a dummy batch "finish" only that long after its
		// submission (the ticket carries the submission time), so the CPU tests exercise the asynchronous paths of
		// the pipeline: polling, sleeping workers, out-of-order hand-back (tests/test_host_vs_oracle.py)
		if (ticket != ~0ull)
			*done = fake_now_ns() >= ticket;
		return true;
	}
	const int rc = b200sdf_poll(ctx_, ticket);
	if (rc < 0) {
		if (err)
			*err = std::string("b200sdf_poll: ") + b200sdf_last_error(ctx_);
		return false;
	}
	*done = rc == 1;
	return true;
}

bool Renderer::render_batch(GlyphBatch &batch, std::string *err) const
{
	uint64_t t = 0;
	return submit_batch(batch, &t, err) && wait_batch(t, err) && batch.finalize(*this, err);
}

std::optional<PbfGlyph> Renderer::render_glyph(const Face &face, uint32_t index, std::string *err) const
{
	std::unique_ptr<GlyphBatch> batch = acquire_batch();
	std::optional<PbfGlyph> out;
	if (batch->add_glyph(face, index) && render_batch(*batch, err))
		out = batch->take_glyph(0);
	release_batch(std::move(batch));
	return out;
}

} // namespace vgb
