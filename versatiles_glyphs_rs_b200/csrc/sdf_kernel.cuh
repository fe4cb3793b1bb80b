// sdf_kernel.cuh — the batched SDF kernel for sm_100a (B200).
//
// Replaces the hot loops of renderer_precise (reference src/render/renderer_precise.rs:33-81):
//   loop 2  (row x segment crossing collection, :41-51)      -> crossing scatter in stage_segment()
//   loop 3  (winding sweep per pixel, :61-67)                 -> prefix sum of the scattered deltas
//   loop 4  (min_distance_to_line_segment, rtree_segments.rs:40-68 over
//            Segment::squared_distance_to_point, geometry/segment.rs:54-99) -> the FP32 pair loop
//   quantisation (:75-79)                                     -> epilogue
// and, for outline-level jobs, Ring::add_quadratic_bezier (src/geometry/ring.rs:119-144) plus
// rings.scale/translate (src/render/renderer.rs:122-131): flattening happens while staging, in
// f64, so flattened segments never touch HBM.
//
// One CTA (128 threads) renders one tile job = a rectangle of TW x TH pixel tiles of one glyph.
// Every thread owns one tile (TW x TH pixels, "item"); a warp owns a group of up to 32 items.
// When the rectangle has fewer than 4 groups the warps split the glyph's segments between them
// (warp slices), and lanes left over inside a group split a warp's segments further (lane
// slices); slices are merged at the end with shared-memory atomicMin on the float bit patterns.
//
// Two staging structures, by the source of the glyph's segments:
//   source B (curve records — every glyph decoded or recorded as curves, i.e. the default path, B200SDF_SHARED_STAGE):
//            the CTA works in ROUNDS of 4 x 64 segments.  The glyph's curve list is bulk-copied into shared memory once
//            per job (TMA unit, one mbarrier); in a round warp w evaluates the end points of the w-th 64 segments in
//            f64 — every segment is staged exactly once per CTA — and writes its own list (vertices, compacted long
//            segments), scatters row crossings and rasterises the bands of its short segments; then a CTA-wide barrier,
//            then EVERY warp evaluates its pixels against its share of all four lists, then a second barrier before the
//            lists are overwritten.  Two __syncthreads per round: 20 % of the warp samples wait there (profiles/), the
//            price of staging each segment once instead of once per item group.
//   source A (raw segments uploaded by the caller): each warp streams its own contiguous range of segments through a
//            private double buffer (cp.async.bulk + per-warp mbarrier) and consumes only its own list — no CTA-wide
//            barrier in that loop.
// A staged segment contributes its start vertex to the vertex list, its row crossings to the winding deltas, a band of
// pixels (short segments: interior nearer than both ends) or a record of the clamped-projection loop (long segments);
// the warp then evaluates its pixels against the vertices with packed FP32 (FADD2 / FMUL2 / FFMA2: two horizontally
// adjacent pixels per instruction, FMNMX3 for two vertices per minimum).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200sdf.h"
#include "glyf_kernel.cuh"

#ifndef B200SDF_MINI
#define B200SDF_MINI 64 // segments a warp stages at a time
#endif
#ifndef B200SDF_PACK
#define B200SDF_PACK 2 // 0 = scalar FP32 pair loop, 1 = FFMA2 for the projection only, 2 = FFMA2/FMUL2 throughout
#endif
#ifndef B200SDF_UNROLL
#define B200SDF_UNROLL 1 // pair-loop unroll over segments (measured: 1 beats 2 and 4 — register pressure)
#endif
#ifndef B200SDF_MIN_CTAS
#define B200SDF_MIN_CTAS 8 // __launch_bounds__ minimum CTAs per SM: 64 registers (measured 6 / 7 / 8 with the shared-staging loop: 0.475 / 0.459 / 0.442 ms)
#endif
#define B200SDF_BOUNDS __launch_bounds__(128, B200SDF_MIN_CTAS)
#ifndef B200SDF_SPLIT_BARRIER
// shared staging: consume the warp's own list between arrive and wait of the round barrier.  Measured and left off:
// C2 / C3 / C4 0.475 / 2.52 / 10.38 ms with it against 0.447 / 2.36 / 9.67 ms without (the duplicated inner loops cost
// more than the overlap gains)
#define B200SDF_SPLIT_BARRIER 0
#endif
#ifndef B200SDF_PERSISTENT_MIN_CTAS
#define B200SDF_PERSISTENT_MIN_CTAS 6 // resident CTAs per SM of the persistent kernel (80 registers)
#endif
#ifndef B200SDF_ALGO
// 1 = every pixel x every segment through the clamped projection (11 flop per pair);
// 2 = the same minimum split into   min over vertices  (one FFMA + half an FMNMX3 per pair)
//                                 + interiors of short segments, rasterised as bands
//                                 + long segments (> 0.5 px) through the clamped projection.
#define B200SDF_ALGO 2
#endif
#ifndef B200SDF_SHARED_STAGE
#define B200SDF_SHARED_STAGE 1 // curve glyphs: stage each segment once per CTA, all warps consume all staged lists
#endif
#ifndef B200SDF_VUNROLL
#define B200SDF_VUNROLL 1 // vertex-loop unroll (pairs of vertices per trip); 1 beats 2 and 4 at 64 registers
#endif
#ifndef B200SDF_EVEN_ROUNDS
// shared staging: a round's segments split evenly between the four warps.  Measured and left off: C2 / C4 0.4515 / 9.566 ms
// with it against 0.4446 / 9.506 ms without
#define B200SDF_EVEN_ROUNDS 0
#endif
#ifndef B200SDF_BAND_MASK
#define B200SDF_BAND_MASK 1 // band rasterisation: cheap row scan into a bit mask, exact test only for the rows it keeps
#endif
#ifndef B200SDF_ROW_SCAN
// epilogue: rectangles wider than this many pixels get their winding numbers from a shuffle scan per row (linear in the
// width) instead of a loop over the columns left of each pixel (quadratic per row, but faster for narrow rectangles)
#define B200SDF_ROW_SCAN 48
#endif
#ifndef B200SDF_VPACK
// vertex loop: squared distances with FFMA2 (two pixels per instruction; 49 instead of 65 instructions per vertex pair).
// Measured C2 / C4: 0.4446 / 9.506 ms against 0.4468 / 9.671 ms (persistent form), 0.408 against 0.426 ms (one CTA per job)
#define B200SDF_VPACK 1
#endif
#ifndef B200SDF_CURVE_SMEM
#define B200SDF_CURVE_SMEM 256 // curve records (32 B) kept in shared memory per CTA
#endif

// -DB200SDF_DEBUG_BOUNDS: trap on any shared-memory index outside its array (compute-sanitizer stand-in)
#ifdef B200SDF_DEBUG_BOUNDS
#include <assert.h>
#define B200SDF_CHECK(cond) assert(cond)
#else
#define B200SDF_CHECK(cond) ((void)0)
#endif

namespace b200sdf {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kTileW = B200SDF_TILE_W;
constexpr int kTileH = B200SDF_TILE_H;
constexpr int kMaxItems = B200SDF_MAX_ITEMS;
constexpr int kMaxPix = kMaxItems * kTileW * kTileH;
constexpr int kMini = B200SDF_MINI;
constexpr int kCurveSmem = B200SDF_CURVE_SMEM;
constexpr int kUnroll = B200SDF_UNROLL;
constexpr int kVUnroll = B200SDF_VUNROLL;
static_assert(kTileW % 2 == 0 && kTileH % 2 == 0, "the packed-FP32 loops pair pixels along x and rows along y");
static_assert(kMaxItems <= 32 * kWarps, "one item per thread");

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA unit; SASS: UBLKCP / SYNCS) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "WAIT_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE_%=;\n\t"
	    "bra WAIT_%=;\n\t"
	    "DONE_%=:\n\t"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
// named barriers (barrier 0 is __syncthreads): arrive without waiting / arrive and wait; `count` threads per phase
__device__ __forceinline__ void named_bar_arrive(int id, int count)
{
	asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// three-input minimum (sm_100+: one FMNMX3 on the ALU pipe instead of two FMNMX)
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
	float r;
	asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
	return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Per-segment record of the pair loop (24 bytes).  Origin and direction are stored negated so the
// loop needs only adds and fused multiply-adds.
struct __align__(16) SegA {
	float nvx, nvy, ndx, ndy; // -start, -(end - start)
};
struct __align__(8) SegN {
	float dxn, dyn; // direction / |direction|^2  (0,0 for a zero-length segment: segment.rs:58-61)
};

constexpr int kVtx = 2 * kMini + 2; // source A stages both end points of a segment
struct WarpStage {
	SegA recA[kMini]; // ALGO 2: only the long segments, compacted
	SegN recN[kMini];
#if B200SDF_ALGO == 2
	float2 vtx[kVtx]; // negated vertices, read two at a time (LDS.128)
#endif
};
constexpr float kLongL2 = 0.25f; // squared length above which a segment takes the clamped-projection loop

// A CTA reads either raw segments or curve records, never both: the two staging areas share storage.
union SourceStage {
	float4 raw[kWarps][2][kMini];     // source A: per-warp TMA destinations, double buffered
	b200sdf_curve curves[kCurveSmem]; // source B: the glyph's curve list
};

struct SharedStorage {
	WarpStage warp[kWarps];
	SourceStage src;
	int delta[kMaxPix];               // signed crossing deltas per pixel of the rectangle (winding sweep)
	unsigned d2[kMaxPix];             // min squared distance per pixel, float bits
	uint8_t obuf[kMaxPix + 32];
	uint64_t bar[kWarps][2];          // per-warp mbarriers of the raw double buffer
	uint64_t curve_bar;
	int st_nv[kWarps], st_nl[kWarps]; // shared staging: vertices / long records each warp staged this round
	// mbarrier phases of a CTA that renders several tile jobs one after the other (the barriers are initialised once):
	// completed uses of curve_bar, and of each warp's two raw-segment barriers.  Kept here, not in registers: the
	// persistent kernel's loop must not carry state through the hot loops.
	uint32_t ph_curve;
	uint32_t ph_raw[kWarps][2];
	// persistent kernel: the tile job claimed for the CTA's next round (claimed while the current job's epilogue runs)
	uint32_t next_t, total;
	uint32_t class_end[kTileClasses]; // claim index space: class c holds claims [class_end[c - 1], class_end[c])
	b200sdf_tile_job next_job;
};

struct Rect {
	int rx0, ry0, rw, rh;
};

// Turn one segment (origin-relative pixel units) into its record and, if `scatter`, add its row
// crossings to the winding deltas.
// Crossing rule = renderer_precise.rs:44-50: upward  s.y <= py <  e.y -> sign +1
//                                             downward s.y >  py >= e.y -> sign -1
// and the sweep (:63-66) subtracts the sign of every crossing with x_c <= px, so a crossing adds
// -sign to the first pixel column whose centre is >= x_c; columns left of the rectangle clamp to
// its first column, columns right of it are dropped.
__device__ __forceinline__ void scatter_crossings(const float4 s, const float dx, const float dy, int *delta, const Rect &R)
{
	const float lo = fminf(s.y, s.w), hi = fmaxf(s.y, s.w);
	// rows r (glyph space) whose centre r+0.5 lies in [lo, hi)
	int r0 = (int)ceilf(lo - 0.5f), r1 = (int)ceilf(hi - 0.5f);
	r0 = max(r0, R.ry0);
	r1 = min(r1, R.ry0 + R.rh);
#pragma unroll 1
	for (int r = r0; r < r1; ++r) {
		const float py = (float)r + 0.5f;
		const bool up = (s.y <= py) && (s.w > py);
		const bool down = (s.y > py) && (s.w <= py);
		if (!(up || down))
			continue;
		const float t = (py - s.y) / dy;
		const float xc = s.x + t * dx;
		int c = (int)ceilf(xc - 0.5f) - R.rx0; // first column with centre >= x_c
		c = max(c, 0);
		if (c < R.rw) {
			B200SDF_CHECK(r >= R.ry0 && r < R.ry0 + R.rh && c >= 0 && (r - R.ry0) * R.rw + c < kMaxPix);
			atomicAdd(&delta[(r - R.ry0) * R.rw + c], up ? -1 : 1);
		}
	}
}

__device__ __forceinline__ void stage_segment(const float4 s, SegA &ra, SegN &rn, int *delta, const Rect &R, bool scatter)
{
	const float dx = s.z - s.x, dy = s.w - s.y;
	const float l2 = dx * dx + dy * dy;
	const float inv = l2 > 0.0f ? __frcp_rn(l2) : 0.0f;
	ra = SegA{-s.x, -s.y, -dx, -dy};
	rn = SegN{dx * inv, dy * inv};
	if (scatter)
		scatter_crossings(s, dx, dy, delta, R);
}

#if B200SDF_ALGO == 2
// Interior of a SHORT segment (0 < |d|^2 <= kLongL2).  A pixel p is nearer to the interior of the
// segment than to its end points iff 0 < u < |d|^2 with u = (p - s).d  (segment.rs:62-70, 0 < t < 1);
// its squared distance is then ((p - s) x d)^2 / |d|^2.  Those pixels form a band of width |d|
// perpendicular to the segment: walk it along its major axis (rows for a flat segment, columns for a
// steep one).  |d| <= 0.5 px makes the band at most 0.7072 px wide along the minor axis, so at most
// ONE pixel centre per row (column) can lie in it.  A band pixel at perpendicular distance D sits
// D |d_c| / |d| <= D beyond its foot point along the major axis, so rows (columns) whose centre is more
// than 6 px past the segment hold only distances > 6 px, which saturate the output (191 - 32 d < 0.5).
__device__ __forceinline__ void band_scatter(const float4 s, float dx, float dy, const float l2, const float inv, unsigned *d2,
                                             const Rect &R)
{
	// minor axis "c" (x for a flat segment), major axis "m" (the band runs along it)
	const bool steep = fabsf(dy) > fabsf(dx);
	const float sc = steep ? s.y : s.x, sm_ = steep ? s.x : s.y;
	const float ec_m = steep ? s.z : s.w; // end point, major coordinate
	const float dc = steep ? dy : dx, dm = steep ? dx : dy;
	const int c0 = steep ? R.ry0 : R.rx0, cn = steep ? R.rh : R.rw;
	const int m0 = steep ? R.rx0 : R.ry0, mn = steep ? R.rw : R.rh;
	const int stride_c = steep ? R.rw : 1, stride_m = steep ? 1 : R.rw;
	const float slope = dm / dc;  // |slope| <= 1
	const float w = l2 / dc;      // signed band width along c, |w| <= sqrt(2) |d|
	int ma = (int)ceilf(fminf(sm_, ec_m) - 6.5f), mb = (int)floorf(fmaxf(sm_, ec_m) + 5.5f); // centres within 6 px
	ma = max(ma, m0);
	mb = min(mb, m0 + mn - 1);
	// u = pac * dc + pam * dm in (0, l2)  <=>  pac between -pam * slope and -pam * slope + w;
	// the first centre at or after the lower end (with slack) is  ceil(k0 - pam * slope)
	const float k0 = (fminf(w, 0.0f) + (sc - 0.5f)) - 1e-4f;
#if B200SDF_BAND_MASK
	// Two passes.  The scan only asks, per row, whether the first pixel centre past the band's lower end can lie inside
	// the band: with a = k0 - pam * slope and frac = ceil(a) - a (exact: both lie within one unit of each other) that is
	// 1e-4 < frac < |w| + 1e-4 in exact arithmetic (u / dc = frac + min(w, 0) - 1e-4); the test here is that interval
	// widened by more than the rounding of a, k0 and u can amount to, so it lets through a superset of the rows the exact
	// test below accepts.  One row in ten passes it, but nearly every scan step has SOME lane that does: collecting the
	// rows in a bit mask keeps the exact test, the cross product and the atomic (two thirds of the old loop body) out of
	// the scan — they run for a warp's two or three busiest rows instead of for all thirteen.
	unsigned rows = 0;
	{
		const float tol = 5e-5f + 2e-6f * (fabsf(sc) + fabsf(sm_) + 8.0f);
		const float f_lo = 1e-4f - tol, f_hi = fabsf(w) + 1e-4f + tol;
		float mf = (float)ma + 0.5f;
		unsigned bit = 1u;
#pragma unroll 1
		for (int m = ma; m <= mb; ++m, mf += 1.0f, bit <<= 1) {
			const float a = fmaf(-(mf - sm_), slope, k0);
			const float frac = ceilf(a) - a;
			if (frac > f_lo && frac < f_hi)
				rows |= bit;
		}
	}
#pragma unroll 1
	while (rows) {
		const int m = ma + (__ffs((int)rows) - 1);
		rows &= rows - 1u;
		const float mf = (float)m + 0.5f;
		unsigned *cell = d2 + (m - m0) * stride_m - c0 * stride_c;
#else
	float mf = (float)ma + 0.5f;
	unsigned *cell = d2 + (ma - m0) * stride_m - c0 * stride_c;
#pragma unroll 1
	for (int m = ma; m <= mb; ++m, mf += 1.0f, cell += stride_m) {
#endif
		const float pam = mf - sm_;
		const float cf = ceilf(fmaf(-pam, slope, k0));
		const float pac = (cf + 0.5f) - sc;
		const float u = fmaf(pac, dc, pam * dm);
		const int c = (int)cf;
		if (u > 0.0f && u < l2 && (unsigned)(c - c0) < (unsigned)cn) {
			const float cr = fmaf(pac, dm, -(pam * dc));
			B200SDF_CHECK((cell + c * stride_c) >= d2 && (cell + c * stride_c) < d2 + R.rw * R.rh && R.rw * R.rh <= kMaxPix);
			atomicMin(cell + c * stride_c, __float_as_uint((cr * cr) * inv));
		}
	}
}
#endif

// ---- device flattening (ring.rs:119-144 in closed form, exact for dyadic inputs) --------------------
__device__ __forceinline__ double lerp_rn(double a, double b, double t)
{
	return __dadd_rn(a, __dmul_rn(t, __dsub_rn(b, a))); // never contracted into an FMA
}

// Point j (0..2^k) of a curve record, in font units.
__device__ __forceinline__ void curve_point(const b200sdf_curve &c, uint32_t j, double &x, double &y)
{
	// t = j / 2^k, exact
	const double t = (double)j * __longlong_as_double((long long)(1023 - (int)c.depth) << 52);
	const double sx = c.sx, sy = c.sy, cx = c.cx, cy = c.cy, ex = c.ex, ey = c.ey;
	x = lerp_rn(lerp_rn(sx, cx, t), lerp_rn(cx, ex, t), t);
	y = lerp_rn(lerp_rn(sy, cy, t), lerp_rn(cy, ey, t), t);
}

// Segment g of a CURVES glyph as origin-relative f32 pixel coordinates — the same arithmetic, in
// the same order, as the host path: p*scale, +dx (renderer.rs:122-131), -origin, narrow to f32.
// binary search: last record with seg_off <= g
__device__ __forceinline__ uint32_t find_curve(const b200sdf_curve *__restrict__ curves, uint32_t n_curves, uint32_t g)
{
	uint32_t lo = 0, hi = n_curves;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (curves[mid].seg_off <= g)
			lo = mid;
		else
			hi = mid;
	}
	return lo;
}

__device__ __forceinline__ float4 flatten_segment_of(const b200sdf_curve c, uint32_t g, double scale, double dx, double ox, double oy)
{
	const uint32_t j = g - c.seg_off;
	double x0, y0, x1, y1;
	curve_point(c, j, x0, y0);
	curve_point(c, j + 1, x1, y1);
	float4 s;
	s.x = (float)__dsub_rn(__dadd_rn(__dmul_rn(x0, scale), dx), ox);
	s.y = (float)__dsub_rn(__dadd_rn(__dmul_rn(y0, scale), 0.0), oy);
	s.z = (float)__dsub_rn(__dadd_rn(__dmul_rn(x1, scale), dx), ox);
	s.w = (float)__dsub_rn(__dadd_rn(__dmul_rn(y1, scale), 0.0), oy);
	return s;
}

__device__ __forceinline__ float4 flatten_segment(const b200sdf_curve *__restrict__ curves, uint32_t n_curves, uint32_t g,
                                                  double scale, double dx, double ox, double oy)
{
	return flatten_segment_of(curves[find_curve(curves, n_curves, g)], g, scale, dx, ox, oy);
}

// The curve records of 32 CONSECUTIVE segments g0 + lane, without a search per lane: every record holds at least
// one segment, so they lie in [c_lo, c_lo + 32] where c_lo is the record of g0.  Lane k looks at record c_lo + 1 + k;
// if it starts at g0 + d (1 <= d <= 31) it sets bit d; the OR of all lanes' bits, counted up to a lane's own
// position, is how many records that lane is past c_lo.  Returns this lane's record index; c_lo_next receives the
// record of segment g0 + 32 (the next pass continues from there).
__device__ __forceinline__ uint32_t curves_of_32(const b200sdf_curve *__restrict__ curves, uint32_t n_curves, uint32_t c_lo,
                                                 uint32_t g0, int lane, uint32_t &c_lo_next)
{
	const uint32_t k = c_lo + 1u + (uint32_t)lane;
	const uint32_t start = k < n_curves ? curves[k].seg_off : 0xffffffffu;
	const uint32_t d = start - g0; // >= 1
	const uint32_t bit = d <= 31u ? (1u << d) : 0u;
	const uint32_t mask = __reduce_or_sync(0xffffffffu, bit);
	// records starting exactly at g0 + 32 belong to the next pass
	const uint32_t starts32 = __ballot_sync(0xffffffffu, d <= 32u); // records starting at or before g0 + 32
	c_lo_next = c_lo + (uint32_t)__popc(starts32);
	return c_lo + (uint32_t)__popc(mask & (0xffffffffu >> (31 - lane)));
}

static_assert(kThreads == 128, "B200SDF_BOUNDS");

// One tile job.  `rot` rotates the warps' roles (see below); the caller has initialised the mbarriers and made sure
// (a CTA-wide barrier) that nobody still reads the shared storage of a previous job.
// where claim number t lives: class after class, heaviest first
__device__ __forceinline__ const b200sdf_tile_job *claimed_tile(const SharedStorage &sm, const b200sdf_tile_job *tiles,
                                                               uint32_t tile_cap, uint32_t t)
{
	int c = 0;
	uint32_t first = 0;
#pragma unroll
	for (int k = 0; k < kTileClasses - 1; ++k)
		if (t >= sm.class_end[k]) { // non-decreasing: the last hit is the class boundary below t
			c = k + 1;
			first = sm.class_end[k];
		}
	return tiles + (size_t)c * tile_cap + (t - first);
}

// claim: persistent kernel only — the batch's cursor and its tile lists; null otherwise.
// One copy of the tile code, called (not inlined) by both kernels: measured on C2, the call costs nothing and the
// separately allocated function is faster than either inlined copy (0.452 -> 0.433 ms one CTA per job, 0.524 -> 0.485 ms
// persistent); the persistent kernel additionally prefers 6 fatter CTAs per SM to 8 (0.485 -> 0.453 ms).
#ifndef B200SDF_TILE_INLINE
#define B200SDF_TILE_INLINE __noinline__
#endif
__device__ B200SDF_TILE_INLINE void render_tile(SharedStorage &sm, const b200sdf_tile_job job, const uint32_t rot,
                                            BatchCounters *__restrict__ claim_ctr,
                                            const b200sdf_tile_job *__restrict__ claim_tiles, const uint32_t claim_cap,
                                            const float4 *__restrict__ segs,
                                            const b200sdf_curve *__restrict__ curves,
                                            const b200sdf_outline_job *__restrict__ ojobs, uint8_t *__restrict__ out)
{
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;

	const int W = job.width, H = job.height;
	Rect R;
	R.rx0 = job.tx0 * kTileW;
	R.ry0 = job.ty0 * kTileH; // rectangle origin (pixels, y upward)
	R.rw = min((int)job.ntx * kTileW, W - R.rx0);
	R.rh = min((int)job.nty * kTileH, H - R.ry0);
	const int rpix = R.rw * R.rh;
	const int n_items = (int)job.ntx * (int)job.nty;
	const uint32_t S = job.seg_cnt;
	const bool from_curves = job.job != B200SDF_NO_JOB;

	// source A: raw segments; source B: curve records of this glyph
	const float4 *gsegs = segs + job.seg_off;
	const b200sdf_curve *gcurves = curves + job.seg_off;
	uint32_t n_curves = 0;
	double g_scale = 0.0, g_dx = 0.0, g_ox = 0.0, g_oy = 0.0;
	if (from_curves) {
		const b200sdf_outline_job oj = ojobs[job.job];
		n_curves = oj.src_cnt;
		g_scale = oj.scale;
		g_dx = oj.dx;
		g_ox = (double)oj.x0;
		g_oy = (double)oj.y0;
	}
	const bool curves_in_smem = from_curves && n_curves != 0 && n_curves <= (uint32_t)kCurveSmem;

	for (int i = tid; i < rpix; i += kThreads) {
		sm.delta[i] = 0;
		sm.d2[i] = 0x7f800000u; // +inf
	}
	__syncthreads();
	if (tid == 0 && curves_in_smem) {
		mbar_expect_tx(&sm.curve_bar, n_curves * 32u);
		bulk_g2s(sm.src.curves, gcurves, n_curves * 32u, &sm.curve_bar);
	}

	// ---- work split: item group per warp, then warp slices x lane slices over the segments ----
	// A warp's scheduler (SM sub-partition) is fixed by its index, and so would be the role it plays if
	// roles followed the index: rotate roles by the CTA index so that heavy and light roles of the CTAs
	// resident on one SM spread over all four sub-partitions.
	const int vw = (warp + (int)rot) & 3;
	const int n_groups = (n_items + 31) >> 5; // 1..4
	int wslices, group, wslice;
	if (n_groups == 1) {
		wslices = 4, group = 0, wslice = vw;
	} else if (n_groups == 2) {
		// the second group is a remainder; if it is small enough to run >= 3 lane slices, one warp
		// takes it with all segments and the full group gets three warps (work 1/3 : <= 1/3 each)
		const int rem = n_items - 32;
		if (32 / rem >= 3) {
			group = vw == 3 ? 1 : 0;
			wslices = vw == 3 ? 1 : 3;
			wslice = vw == 3 ? 0 : vw;
		} else {
			wslices = 2, group = vw & 1, wslice = vw >> 1;
		}
	} else {
		wslices = 1, group = vw, wslice = 0;
	}
	const bool warp_active = group < n_groups;
	const int g_items = warp_active ? min(32, n_items - group * 32) : 1; // items in my group
	const int lslices = 32 / g_items;
	const int item = group * 32 + lane % g_items;
	const int lslice = (lane / g_items) % lslices;
	// this warp's contiguous share of the glyph's segments
	const uint32_t s_begin = (uint32_t)(((uint64_t)S * (uint32_t)wslice) / (uint32_t)wslices);
	const uint32_t s_end = (uint32_t)(((uint64_t)S * (uint32_t)(wslice + 1)) / (uint32_t)wslices);
	// warps of different groups that share a slice stage the same segments; only group 0 scatters crossings
	const bool scatter = group == 0;

	const int tx = item % (int)job.ntx, ty = item / (int)job.ntx;
	// pixel block origin relative to the glyph origin; pixel centres at +0.5
	const float px0 = (float)(R.rx0 + tx * kTileW) + 0.5f;
	const float py0 = (float)(R.ry0 + ty * kTileH) + 0.5f;

	float mn[kTileH][kTileW];
#pragma unroll
	for (int r = 0; r < kTileH; ++r)
#pragma unroll
		for (int j = 0; j < kTileW; ++j)
			mn[r][j] = __int_as_float(0x7f800000);

	float2 pxp[kTileW / 2];
#pragma unroll
	for (int j = 0; j < kTileW / 2; ++j)
		pxp[j] = make_float2(px0 + (float)(2 * j), px0 + (float)(2 * j + 1));
#if B200SDF_ALGO == 2
	float2 pyp[kTileH / 2];
#pragma unroll
	for (int r = 0; r < kTileH / 2; ++r)
		pyp[r] = make_float2(py0 + (float)(2 * r), py0 + (float)(2 * r + 1));
#endif

	if (curves_in_smem) {
		mbar_wait(&sm.curve_bar, sm.ph_curve & 1u); // (incremented at the end of the job, after a CTA-wide barrier)
		gcurves = sm.src.curves;
	}

	// ---- the two halves of a staging pass, shared by both loop structures below ----
	// long-segment loop: my pixels x records [first, count) step stride of one staged list
	auto long_loop = [&](const WarpStage &ws, int count, int first, int stride) {
		const SegA *__restrict__ A = ws.recA;
		const SegN *__restrict__ Nn = ws.recN;
#pragma unroll kUnroll
		for (int i = first; i < count; i += stride) {
				const float4 a = *reinterpret_cast<const float4 *>(&A[i]);
				const SegN q = Nn[i];
#if B200SDF_PACK >= 1
				const float2 nvx = make_float2(a.x, a.x), ndx = make_float2(a.z, a.z), ndy = make_float2(a.w, a.w);
				float2 pax[kTileW / 2];
#pragma unroll
				for (int j = 0; j < kTileW / 2; ++j)
					pax[j] = __fadd2_rn(pxp[j], nvx);
#pragma unroll
				for (int r = 0; r < kTileH; ++r) {
					const float pay1 = (py0 + (float)r) + a.y;
					const float2 pay = make_float2(pay1, pay1);
					const float cr = pay1 * q.dyn;
#pragma unroll
					for (int j = 0; j < kTileW / 2; ++j) {
						float2 t;
						t.x = __saturatef(fmaf(pax[j].x, q.dxn, cr));
						t.y = __saturatef(fmaf(pax[j].y, q.dxn, cr));
						const float2 qx = __ffma2_rn(t, ndx, pax[j]);
						const float2 qy = __ffma2_rn(t, ndy, pay);
#if B200SDF_PACK >= 2
						const float2 d2 = __ffma2_rn(qx, qx, __fmul2_rn(qy, qy));
#else
						float2 d2;
						d2.x = fmaf(qx.x, qx.x, qy.x * qy.x);
						d2.y = fmaf(qx.y, qx.y, qy.y * qy.y);
#endif
						mn[r][2 * j] = fminf(mn[r][2 * j], d2.x);
						mn[r][2 * j + 1] = fminf(mn[r][2 * j + 1], d2.y);
					}
				}
#else
				float pax[kTileW];
#pragma unroll
				for (int j = 0; j < kTileW; ++j)
					pax[j] = (px0 + (float)j) + a.x;
#pragma unroll
				for (int r = 0; r < kTileH; ++r) {
					const float pay = (py0 + (float)r) + a.y;
					const float cr = pay * q.dyn;
#pragma unroll
					for (int j = 0; j < kTileW; ++j) {
						const float t = __saturatef(fmaf(pax[j], q.dxn, cr));
						const float qx = fmaf(t, a.z, pax[j]);
						const float qy = fmaf(t, a.w, pay);
						const float d2 = fmaf(qx, qx, qy * qy);
						mn[r][j] = fminf(mn[r][j], d2);
					}
				}
#endif
		}
	};
#if B200SDF_ALGO == 2
	// stage: one warp turns segments [base, base + n) into the vertex list, the long list, crossings and bands
	auto stage = [&](WarpStage &ws, const float4 *rawbuf, uint32_t base, int n, uint32_t &c_lo, bool do_scatter,
	                 int &nv_out) -> int {
			int n_long = 0;
			for (int i0 = 0; i0 < n; i0 += 32) { // uniform trip count: the long list is compacted by ballot
				const int i = i0 + lane;
				bool is_long = false;
				float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
				float dx = 0.f, dy = 0.f, inv = 0.f;
				uint32_t my_curve = 0;
				if (from_curves) {
					uint32_t next_lo;
					my_curve = curves_of_32(gcurves, n_curves, c_lo, base + (uint32_t)i0, lane, next_lo);
					c_lo = next_lo;
				}
				if (i < n) {
					s = from_curves ? flatten_segment_of(gcurves[min(my_curve, n_curves - 1u)], base + (uint32_t)i, g_scale, g_dx, g_ox, g_oy)
					                : rawbuf[i];
					dx = s.z - s.x, dy = s.w - s.y;
					const float l2 = dx * dx + dy * dy;
					inv = l2 > 0.0f ? __frcp_rn(l2) : 0.0f;
					is_long = l2 > kLongL2;
					B200SDF_CHECK(i >= 0 && 2 * i + 1 < kVtx);
					if (from_curves) {
						ws.vtx[i] = make_float2(-s.x, -s.y); // rings are closed: every end is another segment's start
					} else {
						ws.vtx[2 * i] = make_float2(-s.x, -s.y);
						ws.vtx[2 * i + 1] = make_float2(-s.z, -s.w);
					}
					if (do_scatter) {
						scatter_crossings(s, dx, dy, sm.delta, R);
						if (!is_long && l2 > 0.0f)
							band_scatter(s, dx, dy, l2, inv, sm.d2, R);
					}
				}
				const unsigned mask = __ballot_sync(0xffffffffu, is_long);
				if (is_long) {
					const int k = n_long + __popc(mask & ((1u << lane) - 1u));
					B200SDF_CHECK(k >= 0 && k < kMini);
					ws.recA[k] = SegA{-s.x, -s.y, -dx, -dy};
					ws.recN[k] = SegN{dx * inv, dy * inv};
				}
				n_long += __popc(mask);
			}
		const int nv = from_curves ? n : 2 * n;
		__syncwarp();
		if ((nv & 1) && lane == 0)
			ws.vtx[nv] = ws.vtx[nv - 1]; // pad to a whole pair
		nv_out = (nv + 1) & ~1;
		return n_long;
	};
	// vertex loop: my pixels x vertex pairs [first, nv/2) step stride of one staged list
	auto vertex_loop = [&](const WarpStage &ws, int nv, int first, int stride) {
				const float4 *__restrict__ V4 = reinterpret_cast<const float4 *>(ws.vtx);
				const int npair = nv >> 1;
#pragma unroll kVUnroll
				for (int i = first; i < npair; i += stride) {
					const float4 v = V4[i];
					// two vertices a = (v.x, v.y), b = (v.z, v.w): the per-column / per-row terms with packed
					// FP32 (half the issue slots), the 16 + 16 squared distances with scalar FFMA
					float2 paxa[kTileW / 2], paxb[kTileW / 2], sya[kTileH / 2], syb[kTileH / 2];
#pragma unroll
					for (int j = 0; j < kTileW / 2; ++j) {
						paxa[j] = __fadd2_rn(pxp[j], make_float2(v.x, v.x));
						paxb[j] = __fadd2_rn(pxp[j], make_float2(v.z, v.z));
					}
#pragma unroll
					for (int r = 0; r < kTileH / 2; ++r) {
						const float2 pa = __fadd2_rn(pyp[r], make_float2(v.y, v.y));
						const float2 pb = __fadd2_rn(pyp[r], make_float2(v.w, v.w));
						sya[r] = __fmul2_rn(pa, pa);
						syb[r] = __fmul2_rn(pb, pb);
					}
#pragma unroll
					for (int r = 0; r < kTileH; ++r) {
						const float ya = (r & 1) ? sya[r >> 1].y : sya[r >> 1].x;
						const float yb = (r & 1) ? syb[r >> 1].y : syb[r >> 1].x;
#if B200SDF_VPACK
						// two horizontally adjacent pixels per FFMA2 (the row term is a broadcast scalar operand): the
						// squared distances take half the issue slots, which is what this loop runs out of first
#pragma unroll
						for (int j = 0; j < kTileW / 2; ++j) {
							const float2 da = __ffma2_rn(paxa[j], paxa[j], make_float2(ya, ya));
							const float2 db = __ffma2_rn(paxb[j], paxb[j], make_float2(yb, yb));
							mn[r][2 * j] = fmin3(mn[r][2 * j], da.x, db.x);
							mn[r][2 * j + 1] = fmin3(mn[r][2 * j + 1], da.y, db.y);
						}
#else
#pragma unroll
						for (int j = 0; j < kTileW; ++j) {
							const float xa = (j & 1) ? paxa[j >> 1].y : paxa[j >> 1].x;
							const float xb = (j & 1) ? paxb[j >> 1].y : paxb[j >> 1].x;
							mn[r][j] = fmin3(mn[r][j], fmaf(xa, xa, ya), fmaf(xb, xb, yb));
						}
#endif
					}
				}
	};
#endif

#if B200SDF_ALGO == 2 && B200SDF_SHARED_STAGE
	// Curve glyphs: the CTA stages 4 x kMini segments per round — warp w the w-th quarter, every segment exactly
	// once — and then every warp consumes its share of ALL four staged lists for its own tiles.  (Before, the
	// warps of different item groups each staged the segments they needed: a glyph of more than 32 tiles, i.e.
	// most of the work, was staged twice, and in the 3 + 1 split one warp staged everything alone.)
	if (from_curves) {
		const int cstride = wslices * lslices;                 // consumers of my item group
		const int cs = wslice * lslices + lslice;              // my position among them
		const uint32_t per_round = 4u * (uint32_t)kMini;
		for (uint32_t r0 = 0; r0 < S; r0 += per_round) {
#if B200SDF_EVEN_ROUNDS
			// the round's segments in four equal shares (a short last round is staged in one pass per warp instead of two
			// passes by the first warps while the others wait at the barrier)
			const uint32_t n_round = min(per_round, S - r0);
			const uint32_t share = (n_round + 3u) >> 2;
			const uint32_t base = r0 + (uint32_t)warp * share;
			const int n = (uint32_t)warp * share < n_round ? (int)min(share, n_round - (uint32_t)warp * share) : 0;
#else
			const uint32_t base = r0 + (uint32_t)warp * (uint32_t)kMini;
			const int n = base < S ? (int)min((uint32_t)kMini, S - base) : 0;
#endif
			int nv = 0, nl = 0;
			if (n > 0) {
				uint32_t c_lo = find_curve(gcurves, n_curves, base);
				nl = stage(sm.warp[warp], nullptr, base, n, c_lo, true, nv);
			}
			if (lane == 0) {
				sm.st_nv[warp] = nv;
				sm.st_nl[warp] = nl;
			}
#if B200SDF_SPLIT_BARRIER
			// Split barrier (experiment, see B200SDF_SPLIT_BARRIER): announce "my list is staged", consume MY OWN list
			// (visible to my warp already) while the slower warps finish staging theirs, and only then wait for them.
			__syncwarp();
			named_bar_arrive(1, 2 * kThreads);
			if (warp_active) {
				const int first = (cs + warp) % cstride;
				vertex_loop(sm.warp[warp], nv, first, cstride);
				long_loop(sm.warp[warp], nl, first, cstride);
			}
			named_bar_sync(1, 2 * kThreads);
			if (warp_active) {
#pragma unroll 1
				for (int k = 1; k < kWarps; ++k) {
					const int q = (warp + k) & (kWarps - 1);
					const int first = (cs + q) % cstride; // rotate: list lengths are not multiples of the stride
					vertex_loop(sm.warp[q], sm.st_nv[q], first, cstride);
					long_loop(sm.warp[q], sm.st_nl[q], first, cstride);
				}
			}
#else
			__syncthreads();
			if (warp_active) {
#pragma unroll 1
				for (int q = 0; q < kWarps; ++q) {
					const int first = (cs + q) % cstride; // rotate: list lengths are not multiples of the stride
					vertex_loop(sm.warp[q], sm.st_nv[q], first, cstride);
					long_loop(sm.warp[q], sm.st_nl[q], first, cstride);
				}
			}
#endif
			__syncthreads(); // the lists are overwritten by the next round
		}
	} else
#endif
	if (warp_active && s_end > s_begin) {
		WarpStage &ws = sm.warp[warp];
		float4(*raw)[kMini] = sm.src.raw[warp];
		const uint32_t n_mini = (s_end - s_begin + kMini - 1) / kMini;
		if (!from_curves && lane == 0) {
			for (uint32_t m = 0; m < 2 && m < n_mini; ++m) {
				const uint32_t n = min((uint32_t)kMini, s_end - (s_begin + m * kMini));
				mbar_expect_tx(&sm.bar[warp][m], n * 16u);
				bulk_g2s(raw[m], gsegs + s_begin + m * kMini, n * 16u, &sm.bar[warp][m]);
			}
		}
#if B200SDF_ALGO == 2
		// record of the warp's first segment; later passes advance it without searching (curves_of_32)
		uint32_t c_lo = from_curves ? find_curve(gcurves, n_curves, s_begin) : 0u;
#endif
		for (uint32_t m = 0; m < n_mini; ++m) {
			const uint32_t base = s_begin + m * kMini;
			const int n = (int)min((uint32_t)kMini, s_end - base);
			const int b = (int)(m & 1);
			// ---- stage: every lane turns up to kMini/32 segments into records ----
#if B200SDF_ALGO == 2
			if (!from_curves)
				mbar_wait(&sm.bar[warp][b], (sm.ph_raw[warp][b] + (m >> 1)) & 1u);
			int nv = 0;
			const int n_long = stage(ws, raw[b], base, n, c_lo, scatter, nv);
#else
			const int n_long = n;
			if (!from_curves) {
				mbar_wait(&sm.bar[warp][b], (sm.ph_raw[warp][b] + (m >> 1)) & 1u);
				for (int i = lane; i < n; i += 32)
					stage_segment(raw[b][i], ws.recA[i], ws.recN[i], sm.delta, R, scatter);
			} else {
				for (int i = lane; i < n; i += 32) {
					const float4 s = flatten_segment(gcurves, n_curves, base + (uint32_t)i, g_scale, g_dx, g_ox, g_oy);
					stage_segment(s, ws.recA[i], ws.recN[i], sm.delta, R, scatter);
				}
			}
#endif
			__syncwarp();
			if (!from_curves && lane == 0 && m + 2 < n_mini) {
				const uint32_t n2 = min((uint32_t)kMini, s_end - (base + 2 * kMini));
				fence_proxy_async();
				mbar_expect_tx(&sm.bar[warp][b], n2 * 16u);
				bulk_g2s(raw[b], gsegs + base + 2 * kMini, n2 * 16u, &sm.bar[warp][b]);
			}
#if B200SDF_ALGO == 2
			vertex_loop(ws, nv, lslice, lslices);
#endif
			long_loop(ws, n_long, lslice, lslices);
			__syncwarp(); // records are overwritten by the next staging pass
		}
		if (!from_curves && lane == 0) { // buffer b was used for passes b, b + 2, ... (the loop ended with a __syncwarp)
			sm.ph_raw[warp][0] += (n_mini + 1) >> 1;
			sm.ph_raw[warp][1] += n_mini >> 1;
		}
		__syncwarp();
	}

	// ---- merge slices ----
	if (warp_active) {
#pragma unroll
		for (int r = 0; r < kTileH; ++r) {
			const int y = ty * kTileH + r; // row inside the rectangle
#pragma unroll
			for (int j = 0; j < kTileW; ++j) {
				const int x = tx * kTileW + j;
				if (x < R.rw && y < R.rh) {
					B200SDF_CHECK(y * R.rw + x < kMaxPix);
					atomicMin(&sm.d2[y * R.rw + x], __float_as_uint(mn[r][j]));
				}
			}
		}
	}
	__syncthreads();
	if (tid == 0 && curves_in_smem)
		sm.ph_curve++; // every thread is past its wait on curve_bar
	// Persistent kernel: claim the next job NOW — late enough that the greedy largest-first order still holds (a claim
	// made at the start of a job would hand the heaviest jobs out two at a time), early enough that the atomic and the
	// load of the job record overlap this job's epilogue.
	uint32_t claimed = 0xffffffffu;
	if (claim_ctr && tid == kThreads - 1)
		claimed = atomicAdd(&claim_ctr->next_tile, 1u);

	// ---- epilogue: winding prefix, quantise (renderer_precise.rs:67-79), stage in output order ----
	// Output rows run top (largest y) to bottom; the rectangle's rows [ry0, ry0+rh) map to output
	// rows H-1-y.  For a full-width rectangle they form one contiguous byte range.
	const bool full_width = (R.rx0 == 0 && R.rw == W);
	const size_t gbase = (size_t)job.out_off + (size_t)(H - R.ry0 - R.rh) * (size_t)W; // first byte (full-width case)
	const uint32_t mis = full_width ? (uint32_t)((uintptr_t)(out + gbase) & 15u) : 0u;
	// winding numbers = running sum of the deltas along each row, in place: one warp per row, 32 columns per shuffle scan
	// (linear in the width; a loop over the columns left of every pixel was quadratic per row)
	// (only for wide rectangles: at the usual 20-30 pixels the plain loop below is faster — C2 0.412 against 0.418 ms)
	const bool row_scan = R.rw > B200SDF_ROW_SCAN;
	if (row_scan) {
	for (int y = warp; y < R.rh; y += kWarps) {
		int *drow = &sm.delta[y * R.rw];
		int carry = 0;
		for (int x0 = 0; x0 < R.rw; x0 += 32) {
			const int x = x0 + lane;
			int v = x < R.rw ? drow[x] : 0;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int up = __shfl_up_sync(0xffffffffu, v, d);
				if (lane >= d)
					v += up;
			}
			v += carry;
			if (x < R.rw)
				drow[x] = v;
			carry = __shfl_sync(0xffffffffu, v, 31);
		}
	}
	__syncthreads();
	}
	for (int p = tid; p < rpix; p += kThreads) {
		const int y = p / R.rw, x = p - y * R.rw;
		int wn = 0;
		if (row_scan) {
			wn = sm.delta[p];
		} else {
			const int *drow = &sm.delta[y * R.rw];
			for (int k = 0; k <= x; ++k)
				wn += drow[k];
		}
		const float d = sqrtf(__uint_as_float(sm.d2[p]));
		// value = 255 - (+-d * 32 + 64), clamped, rounded half away from zero
		float v = wn != 0 ? fmaf(d, 32.0f, 191.0f) : fmaf(d, -32.0f, 191.0f);
		v = fminf(fmaxf(v, 0.0f), 255.0f);
		const uint8_t q = (uint8_t)(int)floorf(v + 0.5f);
		if (full_width)
			sm.obuf[mis + (uint32_t)((R.rh - 1 - y) * R.rw + x)] = q;
		else
			out[(size_t)job.out_off + (size_t)(H - 1 - (R.ry0 + y)) * (size_t)W + (size_t)(R.rx0 + x)] = q;
	}
	if (claim_ctr && tid == kThreads - 1) {
		sm.next_t = claimed;
		if (claimed < sm.total)
			sm.next_job = *claimed_tile(sm, claim_tiles, claim_cap, claimed);
	}
	if (!full_width)
		return;
	__syncthreads();
	// coalesced 16-byte stores; smem offset is congruent to the global address mod 16
	uint8_t *gdst = out + gbase - mis; // 16-byte aligned
	const uint32_t nbytes = (uint32_t)rpix;
	const uint32_t nvec = (mis + nbytes + 15u) >> 4;
	for (uint32_t v = tid; v < nvec; v += kThreads) {
		const uint32_t lo = v << 4, hi = lo + 16u;
		if (lo >= mis && hi <= mis + nbytes) {
			*reinterpret_cast<uint4 *>(gdst + lo) = *reinterpret_cast<const uint4 *>(&sm.obuf[lo]);
		} else {
			for (uint32_t k = max(lo, mis); k < min(hi, mis + nbytes); ++k)
				gdst[k] = sm.obuf[k];
		}
	}
}

__device__ __forceinline__ void init_barriers(SharedStorage &sm)
{
	const int tid = threadIdx.x;
	if (tid == 0) {
		mbar_init(&sm.curve_bar, 1);
		sm.ph_curve = 0;
		fence_mbar_init();
	}
	if ((tid & 31) == 0) {
		mbar_init(&sm.bar[tid >> 5][0], 1);
		mbar_init(&sm.bar[tid >> 5][1], 1);
		sm.ph_raw[tid >> 5][0] = 0;
		sm.ph_raw[tid >> 5][1] = 0;
		fence_mbar_init();
	}
}

// One CTA per tile job, jobs planned (and sorted, largest first) by the host.
__global__ void B200SDF_BOUNDS sdf_tiles_kernel(const float4 *__restrict__ segs, const b200sdf_curve *__restrict__ curves,
                                                const b200sdf_outline_job *__restrict__ ojobs,
                                                const b200sdf_tile_job *__restrict__ jobs, uint8_t *__restrict__ out)
{
	__shared__ __align__(128) SharedStorage sm;
	init_barriers(sm); // render_tile's first CTA-wide barrier orders this before any use
	render_tile(sm, jobs[blockIdx.x], blockIdx.x, nullptr, nullptr, 0, segs, curves, ojobs, out);
}

// Persistent form for batches planned on the device (glyf_decode_kernel): the grid is sized for the machine, not for
// the batch — whose tile count the host never learns — and every CTA claims tile jobs from the batch's cursor, cost
// class after cost class (heaviest first: what the host's largest-first sort does for the other kernel), until none
// is left.  The next job is claimed while the current one's epilogue runs (render_tile).  The
// last CTA to finish publishes the batch's overflow flag to *status_out and zeroes the counters for the slot's next
// batch.
__global__ void __launch_bounds__(128, B200SDF_PERSISTENT_MIN_CTAS) sdf_tiles_persistent_kernel(
    const float4 *__restrict__ segs, const b200sdf_curve *__restrict__ curves, const b200sdf_outline_job *__restrict__ ojobs,
    const b200sdf_tile_job *__restrict__ tiles, const uint32_t tile_cap, BatchCounters *__restrict__ ctr,
    uint32_t *__restrict__ status_out, uint8_t *__restrict__ out)
{
	__shared__ __align__(128) SharedStorage sm;
	init_barriers(sm);
	if (threadIdx.x == 0) {
		uint32_t run = 0;
#pragma unroll
		for (int c = 0; c < kTileClasses; ++c) {
			// (read at L2, where the decode kernel's atomics left them)
			run += min(__ldcg(&ctr->class_count[c]), tile_cap);
			sm.class_end[c] = run;
		}
		const uint32_t t = atomicAdd(&ctr->next_tile, 1u);
		sm.total = run;
		sm.next_t = t;
		if (t < run)
			sm.next_job = *claimed_tile(sm, tiles, tile_cap, t);
	}
	for (;;) {
		__syncthreads(); // the claim is visible; everybody is done with the previous job's shared storage
		const uint32_t t = sm.next_t;
		if (t >= sm.total)
			break;
		const b200sdf_tile_job job = sm.next_job;
		__syncthreads(); // everybody holds the job before the claiming thread may overwrite it
		render_tile(sm, job, t, ctr, tiles, tile_cap, segs, curves, ojobs, out);
	}
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(&ctr->done_ctas, 1u) == gridDim.x - 1) {
			// every other CTA has left its loop: nobody reads the counters any more
			if (status_out)
				*status_out = ctr->overflow;
#pragma unroll
			for (int c = 0; c < kTileClasses; ++c)
				ctr->class_count[c] = 0;
			ctr->next_tile = 0;
			ctr->overflow = 0;
			ctr->done_ctas = 0;
			__threadfence_system();
		}
	}
}

// The same tile lists, one CTA per tile job where the grid allows it: the host does not know the batch's tile count when
// it makes the launch, so the grid is a guess (a small multiple of the glyph count).  CTA b renders claim number b; if the
// batch has more tile jobs than the grid has CTAs the first CTAs go on with b + gridDim.x, ...; CTAs past the end leave
// at once.  Against the persistent form: the hardware's CTA scheduler does the dynamic assignment (no cursor, no claim,
// no job hand-over through shared memory) and the kernel fits 64 registers, 8 CTAs per SM.
#ifndef B200SDF_STRIDED_MIN_CTAS
#define B200SDF_STRIDED_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(128, B200SDF_STRIDED_MIN_CTAS) sdf_tiles_strided_kernel(
    const float4 *__restrict__ segs, const b200sdf_curve *__restrict__ curves, const b200sdf_outline_job *__restrict__ ojobs,
    const b200sdf_tile_job *__restrict__ tiles, const uint32_t tile_cap, BatchCounters *__restrict__ ctr,
    uint32_t *__restrict__ status_out, uint8_t *__restrict__ out)
{
	__shared__ __align__(128) SharedStorage sm;
	init_barriers(sm);
	if (threadIdx.x < 32) {
		const int lane = threadIdx.x;
		// (read at L2, where the decode kernel's atomics left them)
		uint32_t run = lane < kTileClasses ? min(__ldcg(&ctr->class_count[lane]), tile_cap) : 0u;
#pragma unroll
		for (int d = 1; d < kTileClasses; d <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, run, d);
			if (lane >= d)
				run += up;
		}
		if (lane < kTileClasses)
			sm.class_end[lane] = run;
		if (lane == kTileClasses - 1)
			sm.total = run;
	}
	__syncthreads();
	const uint32_t total = sm.total;
	for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
		const b200sdf_tile_job job = *claimed_tile(sm, tiles, tile_cap, t);
		render_tile(sm, job, t, nullptr, nullptr, 0, segs, curves, ojobs, out);
		__syncthreads(); // everybody is done with this job's shared storage
	}
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(&ctr->done_ctas, 1u) == gridDim.x - 1) {
			// every other CTA has read the counts and finished: nobody reads the counters any more
			if (status_out)
				*status_out = ctr->overflow;
#pragma unroll
			for (int c = 0; c < kTileClasses; ++c)
				ctr->class_count[c] = 0;
			ctr->next_tile = 0;
			ctr->overflow = 0;
			ctr->done_ctas = 0;
			__threadfence_system();
		}
	}
}

// Device flattening only (b200sdf_flatten_outlines): one CTA per glyph, one thread per segment.
__global__ void __launch_bounds__(256) flatten_kernel(const b200sdf_curve *__restrict__ curves,
                                                      const b200sdf_outline_job *__restrict__ ojobs,
                                                      const uint64_t *__restrict__ seg_base, float4 *__restrict__ out)
{
	const b200sdf_outline_job oj = ojobs[blockIdx.x];
	if (oj.kind != B200SDF_KIND_CURVES)
		return;
	for (uint32_t g = threadIdx.x; g < oj.seg_cnt; g += blockDim.x)
		out[seg_base[blockIdx.x] + g] =
		    flatten_segment(curves + oj.src_off, oj.src_cnt, g, oj.scale, oj.dx, (double)oj.x0, (double)oj.y0);
}

// ---- FP32 peak microbenchmark: 8 independent dependent-FFMA chains per thread ----------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b)
{
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
	float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			x0 = fmaf(x0, a, b);
			x1 = fmaf(x1, a, b);
			x2 = fmaf(x2, a, b);
			x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b);
			x5 = fmaf(x5, a, b);
			x6 = fmaf(x6, a, b);
			x7 = fmaf(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Same with packed FFMA2 (2 FMAs per lane per instruction): tells whether packed FP32 raises the
// FLOP ceiling or only halves the issue slots.
__global__ void __launch_bounds__(256) fp32x2_peak_kernel(float *out, int iters, float a, float b)
{
	float2 x0 = make_float2(threadIdx.x * 1e-3f, 0.5f), x1 = x0, x2 = x0, x3 = x0, x4 = x0, x5 = x0, x6 = x0, x7 = x0;
	x1.x += 1.f, x2.x += 2.f, x3.x += 3.f, x4.x += 4.f, x5.x += 5.f, x6.x += 6.f, x7.x += 7.f;
	const float2 aa = make_float2(a, a), bb = make_float2(b, b);
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			x0 = __ffma2_rn(x0, aa, bb);
			x1 = __ffma2_rn(x1, aa, bb);
			x2 = __ffma2_rn(x2, aa, bb);
			x3 = __ffma2_rn(x3, aa, bb);
			x4 = __ffma2_rn(x4, aa, bb);
			x5 = __ffma2_rn(x5, aa, bb);
			x6 = __ffma2_rn(x6, aa, bb);
			x7 = __ffma2_rn(x7, aa, bb);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] =
	    ((x0.x + x1.x) + (x2.x + x3.x)) + ((x4.x + x5.x) + (x6.x + x7.x)) + ((x0.y + x1.y) + (x2.y + x3.y)) +
	    ((x4.y + x5.y) + (x6.y + x7.y));
}

} // namespace b200sdf
