// geometry.h — f64 2-D primitives of the host side (mirror of reference src/geometry/*.rs).
//
// The reference keeps one heap Vec<Point> per Ring and a Vec<Ring> per glyph; here all rings of
// a glyph live in ONE flat point array with ring offsets (RingSet), because the next step packs
// them into the flat segment buffer that is uploaded once per GlyphBlock.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace vgb {

// geometry/point.rs:9-14
struct Point {
	double x = 0.0, y = 0.0;
	Point() = default;
	Point(double x_, double y_) : x(x_), y(y_) {}
	// point.rs:108-112 — ttf-parser hands out f32, widened losslessly
	static Point from_f32(float x, float y) { return Point((double)x, (double)y); }
	// point.rs:29-31
	Point midpoint(const Point &o) const { return Point((x + o.x) / 2.0, (y + o.y) / 2.0); }
	// point.rs:38-42
	double squared_distance_to(const Point &o) const
	{
		const double dx = o.x - x, dy = o.y - y;
		return dx * dx + dy * dy;
	}
};

// geometry/bbox.rs:26-69
struct BBox {
	Point min{std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()};
	Point max{-std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity()};
	void include_point(const Point &p)
	{
		// f64::min / f64::max for non-NaN inputs (coordinates are never NaN); plain compares compile to
		// minsd/maxsd instead of a libm call
		min.x = p.x < min.x ? p.x : min.x;
		min.y = p.y < min.y ? p.y : min.y;
		max.x = p.x > max.x ? p.x : max.x;
		max.y = p.y > max.y ? p.y : max.y;
	}
	// bbox.rs:56-58: empty only when there is no extent in BOTH axes
	bool is_empty() const { return max.x <= min.x && max.y <= min.y; }
};

// geometry/segment.rs:54-99
struct Segment {
	Point start, end;
	Point project_point_on(const Point &p) const
	{
		const double l2 = start.squared_distance_to(end);
		if (l2 == 0.0)
			return start;
		const double t = ((p.x - start.x) * (end.x - start.x) + (p.y - start.y) * (end.y - start.y)) / l2;
		if (t < 0.0)
			return start;
		if (t > 1.0)
			return end;
		return Point(start.x + t * (end.x - start.x), start.y + t * (end.y - start.y));
	}
	double squared_distance_to_point(const Point &p) const { return p.squared_distance_to(project_point_on(p)); }
};

// All rings of one glyph: geometry/rings.rs + ring.rs over flat storage.
class RingSet {
  public:
	void clear()
	{
		pts_.clear();
		starts_.clear();
		open_ = 0;
	}
	bool is_empty() const { return starts_.empty(); }   // rings.rs: no ring saved
	size_t ring_count() const { return starts_.size(); }
	size_t point_count() const { return open_; }         // points of saved rings only
	size_t ring_begin(size_t r) const { return starts_[r]; }
	size_t ring_end(size_t r) const { return r + 1 < starts_.size() ? starts_[r + 1] : open_; }
	const Point *points() const { return pts_.data(); }
	Point *points() { return pts_.data(); }

	// ---- the ring under construction occupies pts_[open_ ..) ----
	size_t open_len() const { return pts_.size() - open_; }
	void open_clear() { pts_.resize(open_); }
	void open_add(const Point &p) { pts_.push_back(p); }
	const Point &open_last() const { return pts_.back(); }
	const Point &open_first() const { return pts_[open_]; }
	// Ring::close — ring.rs:53-63
	void open_close()
	{
		if (open_len() == 0)
			return;
		const Point first = open_first();
		const Point &last = open_last();
		const double eps = std::numeric_limits<double>::epsilon();
		if (std::fabs(first.x - last.x) > eps || std::fabs(first.y - last.y) > eps)
			pts_.push_back(first);
	}
	// Rings::add_ring of the ring under construction
	void open_commit()
	{
		starts_.push_back((uint32_t)open_);
		open_ = pts_.size();
	}

	// Ring::add_quadratic_bezier — ring.rs:119-144 (explicit stack, right half pushed first)
	void open_add_quadratic_bezier(const Point &start, const Point &ctrl, const Point &end, double tolerance_sq);
	// Ring::add_cubic_bezier — ring.rs:159-187
	void open_add_cubic_bezier(const Point &start, const Point &c1, const Point &c2, const Point &end, double tolerance_sq);

	// Rings::scale then Rings::translate — rings.rs:50-63, point.rs:83-99 (two separate roundings)
	void scale_translate(double scale, double dx, double dy)
	{
		for (size_t i = 0; i < open_; ++i) {
			Point &p = pts_[i];
			p.x *= scale;
			p.y *= scale;
			p.x += dx;
			p.y += dy;
		}
	}
	// Rings::get_bbox — rings.rs:65-73
	BBox get_bbox() const
	{
		BBox b;
		for (size_t i = 0; i < open_; ++i)
			b.include_point(pts_[i]);
		return b;
	}
	// Rings::get_segments count — rings.rs:75-81: consecutive pairs within each ring
	size_t segment_count() const { return open_ - starts_.size(); }

  private:
	std::vector<Point> pts_;
	std::vector<uint32_t> starts_;
	size_t open_ = 0; // first point of the ring under construction
};

} // namespace vgb
