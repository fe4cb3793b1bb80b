// sdf_kernel.cuh — the batched SDF kernel for sm_100a (B200).
//
// Replaces the hot loops of renderer_precise (reference src/render/renderer_precise.rs:33-81):
//   loop 2  (row x segment crossing collection, :41-51)      -> crossing scatter in stage_chunk()
//   loop 3  (winding sweep per pixel, :61-67)                 -> prefix sum of the scattered deltas
//   loop 4  (min_distance_to_line_segment, rtree_segments.rs:40-68 over
//            Segment::squared_distance_to_point, geometry/segment.rs:54-99) -> the FP32 pair loop
//   quantisation (:75-79)                                     -> epilogue
//
// One CTA (128 threads) renders one tile job = a rectangle of TW x TH pixel tiles of one glyph.
// The glyph's segments stream HBM -> shared memory in chunks through the TMA unit
// (cp.async.bulk + mbarrier, double buffered); each chunk is turned into per-segment records
// (origin, direction, direction / |direction|^2) once, then every thread evaluates its TW x TH
// pixels against the chunk.  Threads that would idle because the rectangle has fewer than 128
// items instead take a slice of the chunk's segments (warp slices and lane slices); slices are
// merged with shared-memory atomicMin on the non-negative float bit patterns.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200sdf.h"

namespace b200sdf {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kTileW = B200SDF_TILE_W;
constexpr int kTileH = B200SDF_TILE_H;
constexpr int kMaxItems = B200SDF_MAX_ITEMS;
constexpr int kMaxPix = kMaxItems * kTileW * kTileH;
constexpr int kChunk = 256; // segments per staged chunk

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
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

struct __align__(16) SegA {
	float vx, vy, dx, dy; // start point, direction (end - start)
};
struct __align__(8) SegB {
	float dxn, dyn; // direction / |direction|^2  (0,0 for a zero-length segment: segment.rs:58-61)
};

struct SharedStorage {
	float4 raw[2][kChunk];  // staged b200sdf_segment chunks (TMA destination)
	SegA recA[2][kChunk];
	SegB recB[2][kChunk];
	int delta[kMaxPix];     // signed crossing deltas per pixel of the rectangle (winding sweep)
	unsigned d2[kMaxPix];   // min squared distance per pixel, float bits
	uint8_t obuf[kMaxPix + 32];
	uint64_t bar[2];
};

// Turn one staged chunk into records and scatter its row crossings.
// Crossing rule = renderer_precise.rs:44-50: upward  s.y <= py <  e.y -> sign +1
//                                             downward s.y >  py >= e.y -> sign -1
// and the sweep (:63-66) subtracts the sign of every crossing with x_c <= px, so a crossing adds
// -sign to the first pixel column whose centre is >= x_c; pixels left of the rectangle clamp to
// its first column, pixels right of it are dropped.
__device__ __forceinline__ void stage_chunk(const float4 *raw, SegA *recA, SegB *recB, int n, int *delta, int rx0, int ry0,
                                            int rw, int rh, int tid)
{
	for (int i = tid; i < n; i += kThreads) {
		const float4 s = raw[i];
		const float dx = s.z - s.x, dy = s.w - s.y;
		const float l2 = dx * dx + dy * dy;
		const float inv = l2 > 0.0f ? __frcp_rn(l2) : 0.0f;
		recA[i] = SegA{s.x, s.y, dx, dy};
		recB[i] = SegB{dx * inv, dy * inv};

		const float lo = fminf(s.y, s.w), hi = fmaxf(s.y, s.w);
		// rows r (glyph space) whose centre r+0.5 lies in [lo, hi)
		int r0 = (int)ceilf(lo - 0.5f), r1 = (int)ceilf(hi - 0.5f);
		r0 = max(r0, ry0);
		r1 = min(r1, ry0 + rh);
		for (int r = r0; r < r1; ++r) {
			const float py = (float)r + 0.5f;
			const bool up = (s.y <= py) && (s.w > py);
			const bool down = (s.y > py) && (s.w <= py);
			if (!(up || down))
				continue;
			const float t = (py - s.y) / dy;
			const float xc = s.x + t * dx;
			int c = (int)ceilf(xc - 0.5f) - rx0; // first column with centre >= x_c
			c = max(c, 0);
			if (c < rw)
				atomicAdd(&delta[(r - ry0) * rw + c], up ? -1 : 1);
		}
	}
}

__global__ void __launch_bounds__(kThreads) sdf_tiles_kernel(const float4 *__restrict__ segs,
                                                             const b200sdf_tile_job *__restrict__ jobs,
                                                             uint8_t *__restrict__ out)
{
	__shared__ __align__(128) SharedStorage sm;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;

	// 32-byte job record, uniform across the CTA
	const b200sdf_tile_job job = jobs[blockIdx.x];
	const int W = job.width, H = job.height;
	const int rx0 = job.tx0 * kTileW, ry0 = job.ty0 * kTileH; // rectangle origin (pixels, y upward)
	const int rw = min((int)job.ntx * kTileW, W - rx0);
	const int rh = min((int)job.nty * kTileH, H - ry0);
	const int rpix = rw * rh;
	const int n_items = (int)job.ntx * (int)job.nty;
	const uint32_t S = job.seg_cnt;
	const float4 *gsegs = segs + job.seg_off;
	const int n_chunks = (int)((S + kChunk - 1) / kChunk);

	if (tid == 0) {
		mbar_init(&sm.bar[0], 1);
		mbar_init(&sm.bar[1], 1);
		fence_mbar_init();
	}
	for (int i = tid; i < rpix; i += kThreads) {
		sm.delta[i] = 0;
		sm.d2[i] = 0x7f800000u; // +inf
	}
	__syncthreads();
	if (tid == 0) {
		for (int c = 0; c < 2 && c < n_chunks; ++c) {
			const uint32_t n = min((uint32_t)kChunk, S - (uint32_t)c * kChunk);
			mbar_expect_tx(&sm.bar[c], n * 16u);
			bulk_g2s(sm.raw[c], gsegs + (size_t)c * kChunk, n * 16u, &sm.bar[c]);
		}
	}

	// ---- work split: item group per warp, then warp slices x lane slices over the segments ----
	const int n_groups = (n_items + 31) >> 5;              // 1..4
	const int wslices = n_groups == 1 ? 4 : (n_groups == 2 ? 2 : 1);
	const int group = n_groups == 1 ? 0 : (n_groups == 2 ? (warp & 1) : warp);
	const int wslice = n_groups == 1 ? warp : (n_groups == 2 ? (warp >> 1) : 0);
	const bool warp_active = group < n_groups;
	const int g_items = warp_active ? min(32, n_items - group * 32) : 1; // items in my group
	const int lslices = 32 / g_items;
	const int item = group * 32 + lane % g_items;
	const int lslice = (lane / g_items) % lslices;
	const int T = wslices * lslices;             // total segment slices for my item
	const int sid = wslice * lslices + lslice;   // my slice

	const int tx = item % (int)job.ntx, ty = item / (int)job.ntx;
	// pixel block origin relative to the glyph origin; pixel centres at +0.5
	const float px0 = (float)(rx0 + tx * kTileW) + 0.5f;
	const float py0 = (float)(ry0 + ty * kTileH) + 0.5f;

	float mn[kTileH][kTileW];
#pragma unroll
	for (int r = 0; r < kTileH; ++r)
#pragma unroll
		for (int j = 0; j < kTileW; ++j)
			mn[r][j] = __int_as_float(0x7f800000);

	for (int c = 0; c < n_chunks; ++c) {
		const int b = c & 1;
		const int n = (int)min((uint32_t)kChunk, S - (uint32_t)c * kChunk);
		mbar_wait(&sm.bar[b], (uint32_t)((c >> 1) & 1));
		stage_chunk(sm.raw[b], sm.recA[b], sm.recB[b], n, sm.delta, rx0, ry0, rw, rh, tid);
		__syncthreads(); // records of chunk c visible; raw[b] consumed; everyone is past chunk c-1
		if (tid == 0 && c + 2 < n_chunks) {
			const uint32_t n2 = min((uint32_t)kChunk, S - (uint32_t)(c + 2) * kChunk);
			fence_proxy_async();
			mbar_expect_tx(&sm.bar[b], n2 * 16u);
			bulk_g2s(sm.raw[b], gsegs + (size_t)(c + 2) * kChunk, n2 * 16u, &sm.bar[b]);
		}
		if (warp_active) {
			const SegA *__restrict__ A = sm.recA[b];
			const SegB *__restrict__ B = sm.recB[b];
#pragma unroll 2
			for (int i = sid; i < n; i += T) {
				const SegA a = A[i];
				const SegB q = B[i];
				float pax[kTileW];
#pragma unroll
				for (int j = 0; j < kTileW; ++j)
					pax[j] = (px0 + (float)j) - a.vx;
#pragma unroll
				for (int r = 0; r < kTileH; ++r) {
					const float pay = (py0 + (float)r) - a.vy;
					const float cr = pay * q.dyn;
#pragma unroll
					for (int j = 0; j < kTileW; ++j) {
						const float t = __saturatef(fmaf(pax[j], q.dxn, cr));
						const float qx = fmaf(-t, a.dx, pax[j]);
						const float qy = fmaf(-t, a.dy, pay);
						const float d2 = fmaf(qx, qx, qy * qy);
						mn[r][j] = fminf(mn[r][j], d2);
					}
				}
			}
		}
	}

	// ---- merge slices ----
	if (warp_active) {
#pragma unroll
		for (int r = 0; r < kTileH; ++r) {
			const int y = ty * kTileH + r; // row inside the rectangle
#pragma unroll
			for (int j = 0; j < kTileW; ++j) {
				const int x = tx * kTileW + j;
				if (x < rw && y < rh)
					atomicMin(&sm.d2[y * rw + x], __float_as_uint(mn[r][j]));
			}
		}
	}
	__syncthreads();

	// ---- epilogue: winding prefix, quantise (renderer_precise.rs:67-79), stage in output order ----
	// Output rows run top (largest y) to bottom; the rectangle's rows [ry0, ry0+rh) map to output
	// rows H-1-y.  For a full-width rectangle they form one contiguous byte range.
	const bool full_width = (rx0 == 0 && rw == W);
	const size_t gbase = (size_t)job.out_off + (size_t)(H - ry0 - rh) * (size_t)W; // first byte (full-width case)
	const uint32_t mis = full_width ? (uint32_t)((uintptr_t)(out + gbase) & 15u) : 0u;
	for (int p = tid; p < rpix; p += kThreads) {
		const int y = p / rw, x = p - y * rw;
		int wn = 0;
		const int *drow = &sm.delta[y * rw];
		for (int k = 0; k <= x; ++k)
			wn += drow[k];
		const float d = sqrtf(__uint_as_float(sm.d2[p]));
		// value = 255 - (±d * 32 + 64), clamped, rounded half away from zero
		float v = wn != 0 ? fmaf(d, 32.0f, 191.0f) : fmaf(d, -32.0f, 191.0f);
		v = fminf(fmaxf(v, 0.0f), 255.0f);
		const uint8_t q = (uint8_t)(int)floorf(v + 0.5f);
		if (full_width)
			sm.obuf[mis + (uint32_t)((rh - 1 - y) * rw + x)] = q;
		else
			out[(size_t)job.out_off + (size_t)(H - 1 - (ry0 + y)) * (size_t)W + (size_t)(rx0 + x)] = q;
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

} // namespace b200sdf
