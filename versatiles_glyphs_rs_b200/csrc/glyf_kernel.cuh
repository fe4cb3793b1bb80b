// glyf_kernel.cuh — TrueType `glyf` decoding, outline recording, metrics and tile planning on the device.
//
// Replaces, for glyphs whose outline is a simple `glyf` record (or a composite of simple records that are only
// translated), everything the host did per glyph between the cmap lookup and the SDF kernel:
//   ttf-parser 0.25.1 glyf::outline (flags / coordinate deltas / implied on-curve points)
//       -> the RingBuilder callbacks of reference src/render/ring_builder.rs:67-117
//       -> OutlineRecorder (host/render.cc): one b200sdf_curve per line / quadratic, flattening depth from the
//          reference's own flatness test (src/geometry/ring.rs:128-131), save_ring's 3 / 4 point rules
//          (ring_builder.rs:33-54), the exact bounding box of the flattened points
//       -> rings.scale / translate and prepare_glyph (src/render/renderer.rs:64-91,122-131): integer frame
//       -> tile planning (b200sdf.cu plan_glyph)
// One WARP owns one glyph request.  The font's glyf table is resident in HBM (b200sdf_font_upload); the host sends
// 72 bytes per glyph (cmap / hmtx / loca lookups stay there) instead of recording ~1.4 KB of curve records.
//
// The arithmetic is the host recorder's, operation for operation (f32 midpoints, f64 flatness test and extrema with
// explicit round-to-nearest operations, never contracted), so curve records, segment counts and frames are
// bit-identical to the host path — tests/test_gpu_parity.py compares them for every glyph of every fixture font.
// Anything the closed-form recorder does not cover (coordinates beyond 2^15, depth > 12, truncated or inconsistent
// glyph data, more points than fit in shared memory, a frame larger than the slot the host reserved) sets a status
// code instead; the host then records that glyph literally (host/render.cc) — still rendered by the SDF kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200sdf.h"

#ifndef B200SDF_GLYF_MAX_POINTS
#define B200SDF_GLYF_MAX_POINTS 2048 // points of one simple glyph record held in shared memory per warp
#endif
#ifndef B200SDF_GLYF_STAGE_BYTES
#define B200SDF_GLYF_STAGE_BYTES 1536 // glyph records up to this size are copied to shared memory before parsing
#endif

namespace b200sdf {

constexpr int kGlyfWarps = 4;
constexpr int kGlyfThreads = 32 * kGlyfWarps;
constexpr int kGlyfMaxPts = B200SDF_GLYF_MAX_POINTS;
constexpr uint32_t kGlyfMaxDepth = 12;

// Device-side bookkeeping of one batch.  glyf_decode_kernel appends every tile job to the list of its cost class
// (class_count); the SDF kernel takes them class after class, heaviest first — CTA b the b-th one (strided form), or by
// claiming from next_tile (persistent form) — and its last CTA reports overflow and zeroes everything for the slot's
// next batch.
constexpr int kTileClasses = 8; // class c: cost in (cap / 2^(c+1), cap / 2^c], the last one open-ended
struct BatchCounters {
	uint32_t class_count[kTileClasses];
	uint32_t next_tile;
	uint32_t overflow;
	uint32_t done_ctas; // CTAs of the SDF kernel that have finished
	uint32_t pad;
};

constexpr int kGlyfStageBytes = B200SDF_GLYF_STAGE_BYTES;
struct GlyfWarpScratch {
	int16_t px[kGlyfMaxPts];
	int16_t py[kGlyfMaxPts];
	uint8_t flags[kGlyfMaxPts];
	// The record itself: parsing is a chain of dependent byte reads (header -> end points -> instruction length ->
	// flags -> x -> y), a microsecond each from cold HBM; one coalesced copy puts the whole chain into shared memory.
	// Larger records are parsed in place.
	__align__(16) uint8_t bytes[kGlyfStageBytes + 32];
};

// One caller-side batch of a submission (several queued batches of a pipeline are decoded and rendered by ONE pair of
// launches: a kernel pair over a few hundred glyphs runs as long as its slowest glyph / heaviest tile, whatever its size)
constexpr int kMaxSubBatches = 16;
struct SubBatch {
	const b200sdf_glyph_req *reqs;
	const b200sdf_glyph_part *parts;
	const b200sdf_curve *host_curves; // records of host-recorded CURVES requests
	b200sdf_glyph_frame *frames;      // result per request (pinned host memory or device mirror)
	uint64_t out_addr;                // device-side address of the batch's bitmap area
	uint64_t out_bytes;
	uint32_t n_reqs, req_base;        // requests [req_base, req_base + n_reqs) of the submission
	uint32_t n_parts, n_host_curves, n_host_segs;
	uint32_t seg_base;                // where the batch's segments start in the submission's segment array
	uint32_t curve_base, curve_slots; // the batch's part of the curve scratch
	uint32_t gen_base, gen_slots;     // the batch's part of the generated-segment area (kind PATH), in segments
};

struct DecodeParams {
	SubBatch sub[kMaxSubBatches];
	uint32_t n_sub;
	uint32_t n_reqs;                  // all batches together
	const uint8_t *const *font_base; // device table: glyf bytes of every uploaded font
	const uint64_t *font_len;
	uint32_t n_fonts;
	float4 *segs;                     // the submission's segment array: uploaded segments, then the generated ones
	b200sdf_curve *curves;            // device scratch: every glyph's records at sub.curve_base + req.curve_off
	b200sdf_outline_job *ojobs;       // device scratch: one per request (index = request number in the submission)
	b200sdf_tile_job *tiles;          // kTileClasses lists of tile_cap entries each
	uint32_t tile_cap;
	BatchCounters *counters;
	uint32_t cost_cap;                // largest tile job (item x segment units) before a glyph is cut
	uint32_t min_items;
};

__device__ __forceinline__ uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | (uint32_t)p[1]; }
__device__ __forceinline__ int32_t be16s(const uint8_t *p) { return (int32_t)(int16_t)be16(p); }

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int o = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= d)
			v += o;
	}
	return v;
}

// host/render.cc dyadic_ok
__device__ __forceinline__ bool dyadic_ok_dev(float v)
{
	if (!(v >= -32768.0f && v <= 32768.0f))
		return false;
	const float w = v * 1024.0f;
	return (float)(int32_t)w == w;
}

__device__ __forceinline__ double lerp_rn_g(double a, double b, double t)
{
	return __dadd_rn(a, __dmul_rn(t, __dsub_rn(b, a)));
}
__device__ __forceinline__ double curve_coord_dev(double s, double c, double e, double t)
{
	return lerp_rn_g(lerp_rn_g(s, c, t), lerp_rn_g(c, e, t), t);
}

// host/render.cc OutlineRecorder::axis_extrema: the extremes of one coordinate over the grid points i / 2^k,
// 0 < i < 2^k.  The coordinate is monotone either side of its vertex t* = (s - c) / (s - 2c + e), so the extreme grid
// point is floor(t* 2^k) or ceil(t* 2^k); every candidate evaluated is a point of the polyline, so ANY candidate set
// that contains those two yields the host's minimum and maximum bit for bit.  The host locates t* with an f64
// division and looks at four neighbours; an f32 division is off by far less than one grid step (2^k <= 4096), so
// the same four-neighbour window around the f32 estimate contains them too — without the f64 divide.
__device__ __forceinline__ void axis_extrema_dev(double s, double c, double e, uint32_t k, double &lo, double &hi)
{
	const double mn = s < e ? s : e, mx = s < e ? e : s;
	if (k == 0 || (c >= mn && c <= mx))
		return;
	const double a = __dadd_rn(__dsub_rn(s, __dmul_rn(c, 2.0)), e);
	if (a == 0.0)
		return;
	const int n = 1 << k;
	const float x = __fdividef((float)__dsub_rn(s, c), (float)a) * (float)n; // ~ t* 2^k, |error| << 1
	if (!(x > -1.0f && x < (float)n + 1.0f))
		return; // vertex outside (0, 1) — cannot happen when c lies outside [s, e], kept as a guard
	const double step = __longlong_as_double((long long)(1023 - (int)k) << 52);
	const int i0 = (int)floorf(x);
	double l = lo, h = hi;
#pragma unroll
	for (int d = -1; d <= 2; ++d) {
		int i = i0 + d;
		i = i < 1 ? 1 : (i > n - 1 ? n - 1 : i);
		const double v = curve_coord_dev(s, c, e, __dmul_rn((double)i, step));
		l = v < l ? v : l;
		h = v > h ? v : h;
	}
	lo = l, hi = h;
}

// 0.01 * 16^j, j = 0..12 (host/render.cc kDepthThreshold: PRECISION scaled by exact powers of two)
__device__ __forceinline__ double depth_threshold(uint32_t j)
{
	// 0.01 = 0x3F847AE147AE147B; multiplying by 16^j adds 4j to the exponent field
	return __longlong_as_double(0x3F847AE147AE147Bll + ((long long)(4 * j) << 52));
}

__device__ __forceinline__ double warp_min_d(double v)
{
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		const double o = __shfl_xor_sync(0xffffffffu, v, d);
		v = o < v ? o : v;
	}
	return v;
}
__device__ __forceinline__ double warp_max_d(double v)
{
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		const double o = __shfl_xor_sync(0xffffffffu, v, d);
		v = o > v ? o : v;
	}
	return v;
}

// Cost class of a tile job: the largest-first order the host planner sorts into matters for the heavy jobs (one of
// them is a large share of what a CTA renders in the whole kernel) and not at all for the many light ones, so a
// factor of two per class is as good as a sort — and costs the decode kernel one atomic per job instead of a pass.
__device__ __forceinline__ int tile_class(uint64_t cost, uint32_t cost_cap)
{
	int c = 0;
	uint64_t lim = (uint64_t)cost_cap >> 1;
	while (c < kTileClasses - 1 && cost <= lim) {
		++c;
		lim >>= 1;
	}
	return c;
}

struct GlyphAcc {
	uint32_t n_rec;  // records written so far (kept rings only)
	uint32_t n_seg;  // flattened segments so far
	uint32_t rings;  // kept rings
	double bx0, by0, bx1, by1;
	uint32_t status; // 0 while everything is representable
};

// One simple glyph record (g, len) translated by (ox, oy): appends the records of its kept rings at
// curves[acc.n_rec ..] (at most `room` more) and updates acc.  Warp-collective; returns false on anomaly (acc.status set).
__device__ __forceinline__ bool decode_simple_glyph(const uint8_t *g, uint32_t len, float ox, float oy,
                                                    GlyfWarpScratch &ws, b200sdf_curve *__restrict__ curves, uint32_t room,
                                                    GlyphAcc &acc, int lane)
{
	// host/face.cc outline_impl, simple-glyph arm
	if (len < 10) {
		return true; // `g.len < 10 -> return`: no callbacks
	}
	if (len <= (uint32_t)kGlyfStageBytes) {
		__syncwarp(); // the previous part's parse is over
		const uint32_t mis = (uint32_t)((uintptr_t)g & 15u);
		const uint4 *src = reinterpret_cast<const uint4 *>(g - mis); // font blobs are 256-byte aligned and padded by 16 bytes
		uint4 *dst = reinterpret_cast<uint4 *>(ws.bytes);
		const uint32_t nvec = (mis + len + 15u) >> 4;
		for (uint32_t i = (uint32_t)lane; i < nvec; i += 32)
			dst[i] = __ldg(src + i);
		__syncwarp();
		g = ws.bytes + mis;
	}
	const int32_t n_contours = be16s(g);
	if (n_contours == 0)
		return true;
	if (n_contours < 0) {
		acc.status = B200SDF_GLYPH_NEEDS_HOST; // a composite record is the host's business (it resolves the components)
		return false;
	}
	const uint32_t nc = (uint32_t)n_contours;
	uint32_t pos = 10;
	if (pos + 2 * nc + 2 > len)
		return true; // truncated header: no callbacks
	const uint8_t *end_pts = g + pos;
	const uint32_t n_points = be16(end_pts + 2 * (nc - 1)) + 1;
	if (n_points == 1)
		return true; // a single point is not an outline
	if (n_points > (uint32_t)kGlyfMaxPts) {
		acc.status = B200SDF_GLYPH_NEEDS_HOST;
		return false;
	}
	pos += 2 * nc;
	const uint32_t instr_len = be16(g + pos);
	pos += 2 + instr_len;
	if (pos > len)
		return true;
	// end points must be strictly increasing for the per-contour walk below (anything else: host)
	{
		bool bad = false;
		for (uint32_t c = lane; c < nc; c += 32) {
			const uint32_t e = be16(end_pts + 2 * c);
			const uint32_t p = c ? be16(end_pts + 2 * (c - 1)) + 1 : 0;
			bad |= e < p || e >= n_points;
		}
		if (__any_sync(0xffffffffu, bad)) {
			acc.status = B200SDF_GLYPH_NEEDS_HOST;
			return false;
		}
	}

	// ---- flags (REPEAT 0x08), 32 bytes per pass ------------------------------------------------------
	// Whether a byte is a flag or a repeat count depends on its predecessor: count[i] = repeat_bit[i-1] & !count[i-1],
	// i.e. a byte is a count iff the run of repeat-bit bytes ending just before it has odd length.
	uint32_t flags_end = 0;
	uint32_t x_bytes = 0, y_bytes = 0; // sizes of the coordinate arrays, summed while the flags are expanded
	{
		uint32_t k = 0, p = pos;
		bool carry_rep = false;
		while (k < n_points) {
			if (p >= len) { // ran out of glyph data before every point had a flag (host: emits a partial outline)
				acc.status = B200SDF_GLYPH_NEEDS_HOST;
				return false;
			}
			const uint32_t off = p + (uint32_t)lane;
			const bool valid = off < len;
			const uint32_t byte = valid ? g[off] : 0u;
			const bool has_next = off + 1 < len;
			const uint32_t nxt = has_next ? g[off + 1] : 0u;
			const bool b = valid && (byte & 0x08u);
			const uint32_t m = __ballot_sync(0xffffffffu, b);
			const uint32_t below = (1u << lane) - 1u;
			const uint32_t inv = ~m & below;
			bool isrep;
			if (inv == 0u)
				isrep = carry_rep != ((lane & 1) != 0);
			else
				isrep = (((uint32_t)lane - 1u - (31u - (uint32_t)__clz((int)inv))) & 1u) != 0u;
			const bool isflag = valid && !isrep;
			const int cnt = isflag ? 1 + (b ? (int)nxt : 0) : 0;
			const int incl = warp_incl_scan(cnt, lane);
			const uint32_t start = k + (uint32_t)(incl - cnt);
			const bool consumed = isflag && start < n_points;
			if (__any_sync(0xffffffffu, consumed && b && !has_next)) { // repeat count beyond the record
				acc.status = B200SDF_GLYPH_NEEDS_HOST;
				return false;
			}
			uint32_t sizes = 0; // x bytes | y bytes << 16 of the points this flag byte stands for
			if (consumed) {
				const uint32_t stop = min(start + (uint32_t)cnt, n_points);
				for (uint32_t q = start; q < stop; ++q)
					ws.flags[q] = (uint8_t)byte;
				const uint32_t xs = (byte & 0x02u) ? 1u : ((byte & 0x10u) ? 0u : 2u);
				const uint32_t ys = (byte & 0x04u) ? 1u : ((byte & 0x20u) ? 0u : 2u);
				sizes = (xs | (ys << 16)) * (stop - start); // <= 2 * 2048 per half
			}
			sizes = __reduce_add_sync(0xffffffffu, sizes);
			x_bytes += sizes & 0xffffu;
			y_bytes += sizes >> 16;
			const uint32_t total = (uint32_t)__shfl_sync(0xffffffffu, incl, 31);
			const uint32_t cm = __ballot_sync(0xffffffffu, consumed);
			if (k + total >= n_points) {
				const int last = 31 - __clz((int)cm);
				const uint32_t last_b = (uint32_t)__shfl_sync(0xffffffffu, (int)(b ? 1 : 0), last);
				flags_end = p + (uint32_t)last + 1u + last_b;
				k = n_points;
			} else {
				// every valid byte of this pass was consumed; an invalid lane means the data ended
				k += total;
				carry_rep = __shfl_sync(0xffffffffu, (int)(b && !isrep), 31) != 0;
				p += 32;
			}
		}
	}
	__syncwarp();

	// ---- coordinates: byte offsets and values by prefix sums ----------------------------------------------
	if (flags_end + x_bytes + y_bytes > len) { // truncated coordinate arrays (host: partial outline)
		acc.status = B200SDF_GLYPH_NEEDS_HOST;
		return false;
	}
	{
		uint32_t xpos = flags_end, ypos = flags_end + x_bytes;
		int32_t xacc = 0, yacc = 0;
		for (uint32_t i0 = 0; i0 < n_points; i0 += 32) {
			const uint32_t i = i0 + (uint32_t)lane;
			const uint32_t f = i < n_points ? ws.flags[i] : 0x30u;
			const int xs = (f & 0x02u) ? 1 : ((f & 0x10u) ? 0 : 2);
			const int ys = (f & 0x04u) ? 1 : ((f & 0x20u) ? 0 : 2);
			const int both = warp_incl_scan(xs | (ys << 16), lane); // both prefix sums at once (<= 64 each)
			const int xi = both & 0xffff, yi = both >> 16;
			const uint8_t *xp = g + xpos + (uint32_t)(xi - xs);
			const uint8_t *yp = g + ypos + (uint32_t)(yi - ys);
			int32_t dx = 0, dy = 0;
			if (xs == 1)
				dx = (f & 0x10u) ? (int32_t)xp[0] : -(int32_t)xp[0];
			else if (xs == 2)
				dx = be16s(xp);
			if (ys == 1)
				dy = (f & 0x20u) ? (int32_t)yp[0] : -(int32_t)yp[0];
			else if (ys == 2)
				dy = be16s(yp);
			const int32_t sx = warp_incl_scan(dx, lane), sy = warp_incl_scan(dy, lane);
			if (i < n_points) {
				ws.px[i] = (int16_t)(xacc + sx); // i16 accumulation wraps (ttf-parser: wrapping_add)
				ws.py[i] = (int16_t)(yacc + sy);
			}
			xacc += __shfl_sync(0xffffffffu, sx, 31);
			yacc += __shfl_sync(0xffffffffu, sy, 31);
			xpos += (uint32_t)__shfl_sync(0xffffffffu, xi, 31);
			ypos += (uint32_t)__shfl_sync(0xffffffffu, yi, 31);
		}
	}
	__syncwarp();

	// ---- contours -> records ----------------------------------------------------------------------------
	// ttf-parser's contour walk (host/face.cc ContourEmitter) is a cyclic rule: arriving at point i from its
	// predecessor p (and p's predecessor q),
	//     i on,  p on   -> LINE(p, i)
	//     i on,  p off  -> QUAD(start, ctrl = p, end = i)
	//     i off, p off  -> QUAD(start, ctrl = p, end = mid(p, i))
	//     i off, p on   -> nothing
	// with start = q if q is on-curve else mid(q, p).  The ring starts at the first on-curve point (or at the midpoint
	// of the first two points when both are off-curve): arrivals run s+1 (s+2), ..., e, s (, s+1).
	//
	// "Slot" t = the t-th arrival of the record in emission order (contour after contour, each in its rotated order);
	// the warp takes 32 slots at a time whatever contours they belong to — a glyph of a hundred small contours is a
	// dozen passes, not a hundred.  A ring is dropped when it ends with fewer than 4 points (ring_builder.rs:33-54):
	// only contours of at most 4 points can be (five arrivals give at least three segments), and a window never splits
	// such a contour, so its segment total is a few shuffles away.
	auto contour_of = [&](uint32_t t) { // smallest c with end_pts[c] >= t
		uint32_t lo = 0, hi = nc - 1;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (be16(end_pts + 2 * mid) < t)
				lo = mid + 1;
			else
				hi = mid;
		}
		return lo;
	};
	double rx0 = __longlong_as_double(0x7ff0000000000000ll), ry0 = rx0;
	double rx1 = __longlong_as_double(0xfff0000000000000ll), ry1 = rx1;
	bool bad = false, over = false;
	uint32_t rec_base = acc.n_rec, seg_base = acc.n_seg, rings_kept = 0;
	for (uint32_t t0 = 0; t0 < n_points;) {
		uint32_t t1 = min(t0 + 32u, n_points);
		if (t1 < n_points) { // do not split a contour that may be dropped
			const uint32_t cl = contour_of(t1 - 1);
			const uint32_t sl = cl ? be16(end_pts + 2 * (cl - 1)) + 1 : 0, el = be16(end_pts + 2 * cl);
			if (el >= t1 && el - sl + 1 <= 4)
				t1 = sl; // > t0: the contour has at most 4 slots, the window 32
		}
		const uint32_t t = t0 + (uint32_t)lane;
		const bool valid = t < t1;
		bool has = false;
		uint32_t s = 0, e = 0, n = 1;
		b200sdf_curve r;
		r.sx = r.sy = r.cx = r.cy = r.ex = r.ey = 0.f;
		r.seg_off = 0, r.depth = 0;
		if (valid) {
			const uint32_t c = contour_of(t);
			s = c ? be16(end_pts + 2 * (c - 1)) + 1 : 0;
			e = be16(end_pts + 2 * c);
			n = e - s + 1;
		}
		if (valid && n >= 2) { // (one on-curve point: a 2-point ring, dropped; one off-curve point: no callbacks)
			const uint32_t a0 = (ws.flags[s] & 1u) ? 1u : 2u; // first arrival, relative to s
			uint32_t ri = a0 + (t - s);
			ri = ri >= n ? ri - n : ri; // a0 + j < n + 2 <= 2n
			const uint32_t rp = ri ? ri - 1 : n - 1, rq = rp ? rp - 1 : n - 1;
			const uint32_t i = s + ri, p = s + rp, q = s + rq;
			const bool on_i = ws.flags[i] & 1u, on_p = ws.flags[p] & 1u, on_q = ws.flags[q] & 1u;
			has = on_i || !on_p;
			if (has) {
				const float ix = (float)ws.px[i], iy = (float)ws.py[i];
				const float pxf = (float)ws.px[p], pyf = (float)ws.py[p];
				// (the host transforms AFTER taking midpoints of the untransformed points: same order here)
				if (on_i && on_p) {
					r.sx = __fadd_rn(pxf, ox), r.sy = __fadd_rn(pyf, oy);
					r.cx = r.sx, r.cy = r.sy;
					r.ex = __fadd_rn(ix, ox), r.ey = __fadd_rn(iy, oy);
				} else {
					const float qx = (float)ws.px[q], qy = (float)ws.py[q];
					float sx = qx, sy = qy;
					if (!on_q) { // lerp_half(q, p) = q + 0.5 (p - q)
						sx = __fadd_rn(qx, __fmul_rn(0.5f, __fsub_rn(pxf, qx)));
						sy = __fadd_rn(qy, __fmul_rn(0.5f, __fsub_rn(pyf, qy)));
					}
					float ex = ix, ey = iy;
					if (!on_i) { // lerp_half(p, i)
						ex = __fadd_rn(pxf, __fmul_rn(0.5f, __fsub_rn(ix, pxf)));
						ey = __fadd_rn(pyf, __fmul_rn(0.5f, __fsub_rn(iy, pyf)));
					}
					r.sx = __fadd_rn(sx, ox), r.sy = __fadd_rn(sy, oy);
					r.cx = __fadd_rn(pxf, ox), r.cy = __fadd_rn(pyf, oy);
					r.ex = __fadd_rn(ex, ox), r.ey = __fadd_rn(ey, oy);
					// Ring::add_quadratic_bezier's test at the root (ring.rs:128-131), depth counted against 0.01 * 16^j
					const double ddx = __dsub_rn(__dadd_rn((double)r.sx, (double)r.ex), __dmul_rn((double)r.cx, 2.0));
					const double ddy = __dsub_rn(__dadd_rn((double)r.sy, (double)r.ey), __dmul_rn((double)r.cy, 2.0));
					const double v = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
					uint32_t k = 0;
#pragma unroll
					for (uint32_t d = 0; d <= kGlyfMaxDepth; ++d)
						k += v > depth_threshold(d) ? 1u : 0u;
					bad |= k > kGlyfMaxDepth;
					r.depth = k > kGlyfMaxDepth ? 0u : k;
				}
				bad |= !dyadic_ok_dev(r.sx) || !dyadic_ok_dev(r.sy) || !dyadic_ok_dev(r.cx) || !dyadic_ok_dev(r.cy) ||
				       !dyadic_ok_dev(r.ex) || !dyadic_ok_dev(r.ey);
			}
		}
		int nseg = has ? (1 << r.depth) : 0;
		// segments of my contour when it is small enough to be dropped (it lies inside this window)
		bool keep = n >= 2;
		{
			int tot = 0;
			const int first = (int)(s - t0); // lane of the contour's first slot (only used when n <= 4)
#pragma unroll
			for (int d = 0; d < 4; ++d) {
				const int v = __shfl_sync(0xffffffffu, nseg, (first + d) & 31);
				tot += (s + (uint32_t)d <= e) ? v : 0;
			}
			if (n <= 4)
				keep = n >= 2 && 1 + tot >= 4;
		}
		const bool emit = has && keep;
		nseg = emit ? nseg : 0;
		const uint32_t em = __ballot_sync(0xffffffffu, emit);
		const int sincl = warp_incl_scan(nseg, lane);
		if (emit) {
			const uint32_t idx = rec_base + (uint32_t)__popc(em & ((1u << lane) - 1u));
			r.seg_off = seg_base + (uint32_t)(sincl - nseg);
			if (idx < room)
				curves[idx] = r;
			else
				over = true;
			// bounding box of the flattened points: end point + the grid points next to each coordinate's vertex
			const double exd = (double)r.ex, eyd = (double)r.ey;
			rx0 = exd < rx0 ? exd : rx0, rx1 = exd > rx1 ? exd : rx1;
			ry0 = eyd < ry0 ? eyd : ry0, ry1 = eyd > ry1 ? eyd : ry1;
			if (r.depth) {
				axis_extrema_dev((double)r.sx, (double)r.cx, exd, r.depth, rx0, rx1);
				axis_extrema_dev((double)r.sy, (double)r.cy, eyd, r.depth, ry0, ry1);
			}
		}
		rec_base += (uint32_t)__popc(em);
		seg_base += (uint32_t)__shfl_sync(0xffffffffu, sincl, 31);
		rings_kept += (uint32_t)__popc(__ballot_sync(0xffffffffu, valid && t == e && keep));
		t0 = t1;
	}
	if (__any_sync(0xffffffffu, bad || over)) {
		acc.status = B200SDF_GLYPH_NEEDS_HOST;
		return false;
	}
	if (rings_kept) {
		acc.n_rec = rec_base;
		acc.n_seg = seg_base;
		acc.rings += rings_kept;
		rx0 = warp_min_d(rx0), ry0 = warp_min_d(ry0), rx1 = warp_max_d(rx1), ry1 = warp_max_d(ry1);
		acc.bx0 = rx0 < acc.bx0 ? rx0 : acc.bx0, acc.by0 = ry0 < acc.by0 ? ry0 : acc.by0;
		acc.bx1 = rx1 > acc.bx1 ? rx1 : acc.bx1, acc.by1 = ry1 > acc.by1 ? ry1 : acc.by1;
	}
	__syncwarp();
	return true;
}

// b200sdf.cu plan_glyph / items_cap, one lane per rectangle; tile jobs go to the cost class of their size.
__device__ __forceinline__ void plan_tiles_dev(const DecodeParams &P, uint32_t src_off, uint32_t seg_cnt, uint32_t width,
                                               uint32_t height, uint64_t out_off, uint32_t job, int lane)
{
	const uint32_t nx = (width + B200SDF_TILE_W - 1) / B200SDF_TILE_W, ny = (height + B200SDF_TILE_H - 1) / B200SDF_TILE_H;
	uint64_t items = (uint64_t)P.cost_cap / (uint64_t)(seg_cnt + 8);
	items = items < P.min_items ? P.min_items : items;
	const uint32_t max_items = (uint32_t)(items > B200SDF_MAX_ITEMS ? B200SDF_MAX_ITEMS : items);
	const uint32_t col_parts = (nx + max_items - 1) / max_items;
	const uint32_t cols_per = (nx + col_parts - 1) / col_parts;
	const uint32_t max_rows = max(1u, max_items / cols_per);
	const uint32_t row_parts = (ny + max_rows - 1) / max_rows;
	const uint32_t rows_per = (ny + row_parts - 1) / row_parts;
	const uint32_t n_cols = (nx + cols_per - 1) / cols_per, n_rows = (ny + rows_per - 1) / rows_per;
	const uint32_t n_rects = n_cols * n_rows;
	for (uint32_t r = (uint32_t)lane; r < n_rects; r += 32) {
		const uint32_t ty = (r / n_cols) * rows_per, tx = (r % n_cols) * cols_per;
		b200sdf_tile_job t;
		t.seg_off = src_off;
		t.seg_cnt = seg_cnt;
		t.out_off = out_off;
		t.width = (uint16_t)width;
		t.height = (uint16_t)height;
		t.tx0 = (uint16_t)tx;
		t.ty0 = (uint16_t)ty;
		t.ntx = (uint16_t)min(cols_per, nx - tx);
		t.nty = (uint16_t)min(rows_per, ny - ty);
		t.job = job;
		const uint64_t cost = (uint64_t)t.ntx * t.nty * (uint64_t)(seg_cnt + 8);
		const int cls = tile_class(cost, P.cost_cap);
		const uint32_t at = atomicAdd(&P.counters->class_count[cls], 1u);
		if (at < P.tile_cap)
			P.tiles[(size_t)cls * P.tile_cap + at] = t;
		else
			atomicExch(&P.counters->overflow, 1u);
	}
}

// ---- kind PATH: literal flattening of host-recorded outlines with cubic curves -------------------------------------
// origin-relative f32 pixel coordinates of a font-unit point: p * scale, + dx (renderer.rs:122-131), - origin, narrow
__device__ __forceinline__ float2 path_pixel(double x, double y, double scale, double dx, double ox, double oy)
{
	return make_float2((float)__dsub_rn(__dadd_rn(__dmul_rn(x, scale), dx), ox), (float)__dsub_rn(__dadd_rn(__dmul_rn(y, scale), 0.0), oy));
}
__device__ __forceinline__ double mid_rn(double a, double b) { return __dmul_rn(__dadd_rn(a, b), 0.5); } // (a + b) / 2.0, point.rs:29-31

struct CubicNode {
	double sx, sy, ax, ay, bx, by, ex, ey; // start, control 1, control 2, end
};

// Ring::add_cubic_bezier (src/geometry/ring.rs:159-187) as the reference runs it: explicit stack, right half pushed
// first, a leaf adds its end point.  Writes segment j of the curve to dst[j] for j < cap; returns the number of leaves.
__device__ __noinline__ uint32_t flatten_cubic_dev(const b200sdf_curve &h, const float ex, const float ey, float4 *dst, const uint32_t cap,
                                                   const double scale, const double dx, const double ox, const double oy, bool &overflow)
{
	CubicNode stack[B200SDF_CUBIC_STACK];
	int top = 0;
	stack[top++] = CubicNode{(double)h.sx, (double)h.sy, (double)h.cx, (double)h.cy, (double)h.ex, (double)h.ey, (double)ex, (double)ey};
	float2 prev = path_pixel((double)h.sx, (double)h.sy, scale, dx, ox, oy);
	uint32_t n = 0;
	while (top > 0) {
		const CubicNode q = stack[--top];
		const double ddx = __dsub_rn(__dadd_rn(q.bx, q.ax), __dadd_rn(q.sx, q.ex));
		const double ddy = __dsub_rn(__dadd_rn(q.by, q.ay), __dadd_rn(q.sy, q.ey));
		const bool flat = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)) <= 0.01; // tolerance_sq, ring_builder.rs:62
		if (flat || top + 2 > B200SDF_CUBIC_STACK) {
			overflow |= !flat; // the host never sends a curve that needs a deeper stack: its counts would not match
			const float2 p = path_pixel(q.ex, q.ey, scale, dx, ox, oy);
			if (n < cap)
				dst[n] = make_float4(prev.x, prev.y, p.x, p.y);
			prev = p;
			++n;
			continue;
		}
		const double p01x = mid_rn(q.sx, q.ax), p01y = mid_rn(q.sy, q.ay);
		const double p12x = mid_rn(q.ax, q.bx), p12y = mid_rn(q.ay, q.by);
		const double p23x = mid_rn(q.bx, q.ex), p23y = mid_rn(q.by, q.ey);
		const double p012x = mid_rn(p01x, p12x), p012y = mid_rn(p01y, p12y);
		const double p123x = mid_rn(p12x, p23x), p123y = mid_rn(p12y, p23y);
		const double mx = mid_rn(p012x, p123x), my = mid_rn(p012y, p123y);
		stack[top++] = CubicNode{mx, my, p123x, p123y, p23x, p23y, q.ex, q.ey};
		stack[top++] = CubicNode{q.sx, q.sy, p01x, p01y, p012x, p012y, mx, my};
	}
	return n;
}

// One lane per record: every record of the glyph becomes its segments in dst[seg_off ...].  Returns the number of
// segments this lane wrote (or would have written); `bad` = malformed records.
__device__ __forceinline__ uint32_t flatten_path_dev(const b200sdf_curve *recs, const uint32_t n_recs, float4 *dst, const uint32_t cap,
                                                     const double scale, const double dx, const double ox, const double oy, const int lane,
                                                     bool &bad)
{
	uint32_t made = 0;
	for (uint32_t i = (uint32_t)lane; i < n_recs; i += 32) {
		const b200sdf_curve r = recs[i];
		if ((r.depth & 0xC0000000u) == B200SDF_CURVE_TAIL)
			continue;
		if (r.depth & B200SDF_CURVE_CUBIC) {
			const uint32_t want = r.depth & 0x3fffffffu;
			if (i + 1 >= n_recs || (recs[i + 1].depth & 0xC0000000u) != B200SDF_CURVE_TAIL || r.seg_off > cap || want > cap - r.seg_off) {
				bad = true;
				continue;
			}
			bool overflow = false;
			const uint32_t n = flatten_cubic_dev(r, recs[i + 1].sx, recs[i + 1].sy, dst + r.seg_off, want, scale, dx, ox, oy, overflow);
			bad |= overflow || n != want;
			made += n;
		} else {
			const uint32_t k = r.depth;
			if (k > kGlyfMaxDepth || r.seg_off > cap || (1u << k) > cap - r.seg_off) {
				bad = true;
				continue;
			}
			const double sx = r.sx, sy = r.sy, cx = r.cx, cy = r.cy, ex = r.ex, ey = r.ey;
			const double step = __longlong_as_double((long long)(1023 - (int)k) << 52); // 2^-k
			float2 prev = path_pixel(sx, sy, scale, dx, ox, oy);
			for (uint32_t j = 1; j <= (1u << k); ++j) {
				const double t = (double)j * step;
				// (the same closed form as the SDF kernel's curve_point: exact for the dyadic inputs the recorder admits)
				const float2 p = path_pixel(curve_coord_dev(sx, cx, ex, t), curve_coord_dev(sy, cy, ey, t), scale, dx, ox, oy);
				dst[r.seg_off + j - 1] = make_float4(prev.x, prev.y, p.x, p.y);
				prev = p;
			}
			made += 1u << k;
		}
	}
	return made;
}

__device__ __forceinline__ void decode_request(const DecodeParams &P, const uint32_t gi, GlyfWarpScratch &wscratch, const int lane)
{
	int k = 0;
#pragma unroll 1
	for (int q = 1; q < (int)P.n_sub; ++q)
		k = gi >= P.sub[q].req_base ? q : k;
	const SubBatch &B = P.sub[k];
	const uint32_t li = gi - B.req_base;
	const b200sdf_glyph_req rq = B.reqs[li];
	b200sdf_glyph_frame fr;
	fr.x0 = rq.x0, fr.y0 = rq.y0, fr.width = 0, fr.height = 0, fr.seg_cnt = 0, fr.status = B200SDF_GLYPH_EMPTY;
	b200sdf_outline_job oj;
	oj.kind = B200SDF_KIND_CURVES;
	oj.src_off = B.curve_base + rq.curve_off, oj.src_cnt = 0, oj.seg_cnt = 0;
	oj.width = oj.height = 0;
	oj.x0 = oj.y0 = 0;
	oj.scale = rq.scale, oj.dx = rq.dx;
	oj.out_off = B.out_addr + rq.out_off; // absolute: the SDF kernel's bitmap base is 0

	if (rq.kind == B200SDF_KIND_GLYF) {
		GlyphAcc acc;
		acc.n_rec = acc.n_seg = acc.rings = 0;
		acc.bx0 = acc.by0 = __longlong_as_double(0x7ff0000000000000ll);
		acc.bx1 = acc.by1 = __longlong_as_double(0xfff0000000000000ll);
		acc.status = 0;
		const bool range_ok = (uint64_t)rq.src_off + rq.src_cnt <= B.n_parts && (uint64_t)rq.curve_off + rq.curve_cap <= B.curve_slots;
		if (!range_ok)
			acc.status = B200SDF_GLYPH_BAD_REQUEST;
		for (uint32_t pi = 0; pi < rq.src_cnt && acc.status == 0; ++pi) {
			const b200sdf_glyph_part pt = B.parts[rq.src_off + pi];
			if (pt.font >= P.n_fonts || (uint64_t)pt.glyf_off + pt.glyf_len > P.font_len[pt.font]) {
				acc.status = B200SDF_GLYPH_BAD_REQUEST;
				break;
			}
			if (!decode_simple_glyph(P.font_base[pt.font] + pt.glyf_off, pt.glyf_len, pt.ox, pt.oy, wscratch,
			                         P.curves + B.curve_base + rq.curve_off, rq.curve_cap, acc, lane))
				break;
		}
		if (acc.status != 0) {
			fr.status = acc.status;
		} else if (acc.rings != 0) {
			// renderer.rs:122-131 on the font-unit box (x * scale, then + dx: monotone, so the box of the transformed
			// points is the transformed box), bbox.rs:56-58, prepare_glyph renderer.rs:64-91
			const double lx = __dadd_rn(__dmul_rn(acc.bx0, rq.scale), rq.dx), ly = __dadd_rn(__dmul_rn(acc.by0, rq.scale), 0.0);
			const double hx = __dadd_rn(__dmul_rn(acc.bx1, rq.scale), rq.dx), hy = __dadd_rn(__dmul_rn(acc.by1, rq.scale), 0.0);
			if (!(hx <= lx && hy <= ly)) {
				const double fx0 = floor(lx) - 3.0, fy0 = floor(ly) - 3.0, fx1 = ceil(hx) + 3.0, fy1 = ceil(hy) + 3.0;
				const double w = fx1 - fx0, h = fy1 - fy0;
				if (!(fabs(fx0) < 1e9 && fabs(fy0) < 1e9 && w >= 1.0 && h >= 1.0 && w <= (double)B200SDF_MAX_DIM && h <= (double)B200SDF_MAX_DIM) ||
				    (uint64_t)w * (uint64_t)h > (uint64_t)rq.out_cap || rq.out_off + (uint64_t)w * (uint64_t)h > B.out_bytes) {
					fr.status = B200SDF_GLYPH_NEEDS_HOST; // frame does not fit the slot the host reserved from the header bbox
				} else {
					fr.x0 = (int32_t)fx0, fr.y0 = (int32_t)fy0;
					fr.width = (uint32_t)w, fr.height = (uint32_t)h;
					fr.seg_cnt = acc.n_seg;
					fr.status = B200SDF_GLYPH_OK;
					oj.src_cnt = acc.n_rec;
					oj.seg_cnt = acc.n_seg;
				}
			}
		}
	} else if (rq.kind == B200SDF_KIND_CURVES) {
		// host-recorded glyph (scaled composites, ...): frame and records are given; copy the records next to the others
		const bool ok = (uint64_t)rq.src_off + rq.src_cnt <= B.n_host_curves && rq.src_cnt <= rq.curve_cap &&
		                (uint64_t)rq.curve_off + rq.curve_cap <= B.curve_slots && rq.width >= 1 && rq.height >= 1 &&
		                rq.width <= B200SDF_MAX_DIM && rq.height <= B200SDF_MAX_DIM &&
		                rq.out_off + (uint64_t)rq.width * rq.height <= B.out_bytes;
		if (!ok) {
			fr.status = B200SDF_GLYPH_BAD_REQUEST;
		} else {
			const uint4 *src = reinterpret_cast<const uint4 *>(B.host_curves + rq.src_off);
			uint4 *dst = reinterpret_cast<uint4 *>(P.curves + B.curve_base + rq.curve_off);
			for (uint32_t i = (uint32_t)lane; i < rq.src_cnt * 2u; i += 32)
				dst[i] = src[i];
			fr.width = rq.width, fr.height = rq.height, fr.seg_cnt = rq.seg_cnt, fr.status = B200SDF_GLYPH_OK;
			oj.src_cnt = rq.src_cnt;
			oj.seg_cnt = rq.seg_cnt;
		}
	} else if (rq.kind == B200SDF_KIND_PATH) {
		const bool ok = (uint64_t)rq.src_off + rq.src_cnt <= B.n_host_curves && rq.seg_cnt >= 1 && rq.seg_cnt == rq.curve_cap &&
		                (uint64_t)rq.curve_off + rq.curve_cap <= B.gen_slots && rq.width >= 1 && rq.height >= 1 &&
		                rq.width <= B200SDF_MAX_DIM && rq.height <= B200SDF_MAX_DIM &&
		                rq.out_off + (uint64_t)rq.width * rq.height <= B.out_bytes;
		if (!ok) {
			fr.status = B200SDF_GLYPH_BAD_REQUEST;
		} else {
			bool bad = false;
			float4 *dst = P.segs + B.gen_base + rq.curve_off;
			uint32_t made = flatten_path_dev(B.host_curves + rq.src_off, rq.src_cnt, dst, rq.seg_cnt, rq.scale, rq.dx, (double)rq.x0,
			                                 (double)rq.y0, lane, bad);
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1)
				made += __shfl_xor_sync(0xffffffffu, made, d);
			bad = __any_sync(0xffffffffu, bad) || made != rq.seg_cnt;
			if (bad) {
				fr.status = B200SDF_GLYPH_NEEDS_HOST; // the host's counts and the device's subdivision disagree
			} else {
				fr.width = rq.width, fr.height = rq.height, fr.seg_cnt = rq.seg_cnt, fr.status = B200SDF_GLYPH_OK;
				oj.kind = B200SDF_KIND_SEGMENTS;
				oj.src_off = B.gen_base + rq.curve_off, oj.src_cnt = rq.seg_cnt, oj.seg_cnt = rq.seg_cnt;
			}
		}
	} else if (rq.kind == B200SDF_KIND_SEGMENTS) {
		const bool ok = (uint64_t)rq.src_off + rq.src_cnt <= B.n_host_segs && rq.seg_cnt == rq.src_cnt && rq.width >= 1 &&
		                rq.height >= 1 && rq.width <= B200SDF_MAX_DIM && rq.height <= B200SDF_MAX_DIM &&
		                rq.out_off + (uint64_t)rq.width * rq.height <= B.out_bytes;
		if (!ok) {
			fr.status = B200SDF_GLYPH_BAD_REQUEST;
		} else {
			fr.width = rq.width, fr.height = rq.height, fr.seg_cnt = rq.seg_cnt, fr.status = B200SDF_GLYPH_OK;
			oj.kind = B200SDF_KIND_SEGMENTS;
			oj.src_off = B.seg_base + rq.src_off, oj.src_cnt = rq.src_cnt, oj.seg_cnt = rq.seg_cnt;
		}
	} else {
		fr.status = B200SDF_GLYPH_BAD_REQUEST;
	}

	if (fr.status == B200SDF_GLYPH_OK) {
		oj.width = fr.width, oj.height = fr.height;
		oj.x0 = fr.x0, oj.y0 = fr.y0;
	}
	__syncwarp(); // this warp's curve records are written before any lane publishes the jobs
	if (lane == 0) {
		P.ojobs[gi] = oj;
		B.frames[li] = fr;
	}
	if (fr.status == B200SDF_GLYPH_OK)
		plan_tiles_dev(P, oj.src_off, oj.seg_cnt, oj.width, oj.height, oj.out_off,
		               oj.kind == B200SDF_KIND_SEGMENTS ? B200SDF_NO_JOB : gi, lane);
}

#ifndef B200SDF_GLYF_MIN_CTAS
#define B200SDF_GLYF_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(kGlyfThreads, B200SDF_GLYF_MIN_CTAS) glyf_decode_kernel(const DecodeParams P)
{
	__shared__ GlyfWarpScratch scratch[kGlyfWarps];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t gi = blockIdx.x * kGlyfWarps + (uint32_t)warp;
	if (gi < P.n_reqs)
		decode_request(P, gi, scratch[warp], lane);
}

} // namespace b200sdf
