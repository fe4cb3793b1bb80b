// geometry.cc — Bezier flattening (reference src/geometry/ring.rs:119-187).
#include "geometry.h"

namespace vgb {

namespace {
struct Quad {
	Point s, c, e;
};
struct Cubic {
	Point s, c1, c2, e;
};
} // namespace

void RingSet::open_add_quadratic_bezier(const Point &start, const Point &ctrl, const Point &end, double tolerance_sq)
{
	// Depth-first De Casteljau with an explicit stack; the right half is pushed first so the left
	// half is processed next and points come out in start -> end order (ring.rs:136-142).
	Quad stack[96];
	int top = 0;
	stack[top++] = Quad{start, ctrl, end};
	while (top > 0) {
		const Quad q = stack[--top];
		const double dx = q.s.x + q.e.x - q.c.x * 2.0;
		const double dy = q.s.y + q.e.y - q.c.y * 2.0;
		if (dx * dx + dy * dy <= tolerance_sq || top + 2 > 96) {
			open_add(q.e);
			continue;
		}
		const Point mid_1 = q.s.midpoint(q.c);
		const Point mid_2 = q.c.midpoint(q.e);
		const Point mid = mid_1.midpoint(mid_2);
		stack[top++] = Quad{mid, mid_2, q.e};
		stack[top++] = Quad{q.s, mid_1, mid};
	}
}

void RingSet::open_add_cubic_bezier(const Point &start, const Point &c1, const Point &c2, const Point &end,
                                    double tolerance_sq)
{
	Cubic stack[96];
	int top = 0;
	stack[top++] = Cubic{start, c1, c2, end};
	while (top > 0) {
		const Cubic q = stack[--top];
		const double dx = (q.c2.x + q.c1.x) - (q.s.x + q.e.x);
		const double dy = (q.c2.y + q.c1.y) - (q.s.y + q.e.y);
		if (dx * dx + dy * dy <= tolerance_sq || top + 2 > 96) {
			open_add(q.e);
			continue;
		}
		const Point p01 = q.s.midpoint(q.c1);
		const Point p12 = q.c1.midpoint(q.c2);
		const Point p23 = q.c2.midpoint(q.e);
		const Point p012 = p01.midpoint(p12);
		const Point p123 = p12.midpoint(p23);
		const Point mid = p012.midpoint(p123);
		stack[top++] = Cubic{mid, p123, p23, q.e};
		stack[top++] = Cubic{q.s, p01, p012, mid};
	}
}

} // namespace vgb
