// b200sdf.cu — C ABI around the sm_100a SDF kernel (include/b200sdf.h).
//
// Context = one GPU + a pool of slots; a slot = one CUDA stream + device buffers + pinned staging
// for the tile-job list + one completion event.  submit() plans tiles on the host, then enqueues
// H2D(curves / segments / jobs) -> H2D(tiles) -> kernel -> D2H(bitmaps) on the slot's stream and
// returns.  This is the async batch pipeline that replaces the serial per-block loop of
// FontManager::render_glyphs (reference src/font/manager.rs:104-121).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "sdf_kernel.cuh"

namespace {

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
};

// Pinned host memory handed out by b200sdf_alloc_pinned: [start, start + bytes) -> device-side address.
// A batch whose job list and bitmap buffer live in such memory is rendered without staging copies for
// them: the kernel reads the (small) job records and writes the bitmaps straight over PCIe, so the slot's
// stream carries one H2D copy and one kernel instead of four copies, a kernel and a copy back.
struct PinnedRegistry {
	std::mutex mu;
	std::map<uintptr_t, std::pair<size_t, uintptr_t>> ranges; // start -> (bytes, device address)
	void add(void *p, size_t bytes, void *dev)
	{
		std::lock_guard<std::mutex> g(mu);
		ranges[(uintptr_t)p] = std::make_pair(bytes, (uintptr_t)dev);
	}
	void remove(void *p)
	{
		std::lock_guard<std::mutex> g(mu);
		ranges.erase((uintptr_t)p);
	}
	// device address of host range [p, p + bytes) if it lies inside one pinned allocation, else nullptr
	void *device_ptr(const void *p, size_t bytes)
	{
		std::lock_guard<std::mutex> g(mu);
		auto it = ranges.upper_bound((uintptr_t)p);
		if (it == ranges.begin())
			return nullptr;
		--it;
		const uintptr_t off = (uintptr_t)p - it->first;
		if (off > it->second.first || bytes > it->second.first - off)
			return nullptr;
		return (void *)(it->second.second + off);
	}
};
PinnedRegistry &pinned_registry()
{
	static PinnedRegistry *r = new PinnedRegistry(); // leaked on purpose: outlives static destructors
	return *r;
}
#ifndef B200SDF_ZEROCOPY_CURVES
#define B200SDF_ZEROCOPY_CURVES 1
#endif
// B200SDF_ZEROCOPY: 0 = stage everything, 1 (default) = inputs and bitmaps in place, 2 = bitmaps only
// Device of the most recently created context: threads that only allocate pinned memory (a pipeline's workers never
// call cudaSetDevice themselves) must not touch — and thereby create a CUDA context on — device 0 in a process that
// renders on another GPU.
std::atomic<int> g_home_device{-1};

int zero_copy_mode()
{
	static const int mode = [] {
		const char *e = std::getenv("B200SDF_ZEROCOPY");
		return e ? std::atoi(e) : 1;
	}();
	return mode;
}

struct Slot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	DevBuf segs, curves, ojobs, tiles, out, seg_base;
	// glyph-level submissions (b200sdf_submit_glyphs): staged copies of the request arrays when the caller's memory is
	// not pinned, the device-written curve scratch, and the batch counters with their pinned mirror
	DevBuf reqs, parts, frames, gcurves;
	b200sdf::BatchCounters *counters = nullptr; // device
	uint32_t *h_status = nullptr;               // pinned + mapped: the batch's overflow flag, written by the SDF kernel's last CTA
	uint32_t *d_status = nullptr;               // its device-side address
	bool check_overflow = false;
	void *h_tiles = nullptr; // pinned staging: tile jobs
	size_t h_tiles_cap = 0;
	bool busy = false;
	uint64_t generation = 0;
	// B200SDF_GPU_TRACE=1: device-side timeline of every batch (kernel start / stop relative to the context's epoch)
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	uint64_t host_submit_ns = 0;
	uint32_t traced_tiles = 0;
};

} // namespace

struct b200sdf_ctx {
	int device = 0;
	std::mutex mu;
	std::condition_variable cv;
	std::vector<Slot> slots;
	std::string err;
	uint64_t launches = 0;
	float *d_peak = nullptr;
	cudaEvent_t epoch = nullptr; // B200SDF_GPU_TRACE
	uint64_t epoch_host_ns = 0;
	size_t hwm[5] = {0, 0, 0, 0, 0}; // largest per-batch buffer sizes seen (segments, curves, jobs, tiles, bitmaps)
	size_t ghwm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // same for glyph-level submissions
	// fonts resident in HBM (b200sdf_font_upload): glyf tables + the device-side tables the decoder indexes
	struct FontBlob {
		void *p = nullptr;
		uint64_t len = 0;
	};
	std::vector<FontBlob> fonts;
	const uint8_t **d_font_base = nullptr;
	uint64_t *d_font_len = nullptr;
	// scratch of the device-resident entry point (b200sdf_render_glyphs_device)
	DevBuf dv_gcurves, dv_ojobs, dv_tiles;
	b200sdf::BatchCounters *dv_counters = nullptr;
};
constexpr uint32_t kMaxFonts = 8192;

namespace {

int fail_cuda(b200sdf_ctx *ctx, cudaError_t e, const char *what)
{
	std::lock_guard<std::mutex> g(ctx->mu);
	ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
	return (e == cudaErrorMemoryAllocation) ? B200SDF_E_NOMEM : B200SDF_E_CUDA;
}
int fail_arg(b200sdf_ctx *ctx, const char *what)
{
	if (ctx) {
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->err = what;
	}
	return B200SDF_E_ARG;
}

#define CU_TRY(ctx, call)                       \
	do {                                        \
		cudaError_t e_ = (call);                \
		if (e_ != cudaSuccess)                  \
			return fail_cuda((ctx), e_, #call); \
	} while (0)

size_t grown(size_t need, size_t cap)
{
	size_t n = std::max(need, cap + cap / 2);
	return (n + 255) & ~size_t(255);
}

// Stream-ordered (re)allocation on the slot's own stream: unlike cudaFree / cudaMalloc it does not
// synchronise the device, so a slot that meets a bigger batch does not stall the other slots' work.
int grow_device(b200sdf_ctx *ctx, DevBuf &b, size_t need, cudaStream_t stream)
{
	if (need <= b.cap)
		return 0;
	const size_t n = grown(need, b.cap);
	if (b.p)
		cudaFreeAsync(b.p, stream);
	b.p = nullptr;
	b.cap = 0;
	if (std::getenv("B200SDF_TRACE"))
		std::fprintf(stderr, "[b200sdf trace] grow_device %zu -> %zu bytes\n", need, n);
	cudaError_t e = cudaMallocAsync(&b.p, n, stream);
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "cudaMallocAsync");
	b.cap = n;
	return 0;
}

int grow_pinned(b200sdf_ctx *ctx, void **p, size_t *cap, size_t need)
{
	if (need <= *cap)
		return 0;
	if (*p)
		b200sdf_free_pinned(*p);
	*p = nullptr;
	const size_t n = grown(need, *cap);
	*cap = 0;
	*p = b200sdf_alloc_pinned(n);
	if (!*p)
		return fail_cuda(ctx, cudaErrorMemoryAllocation, "cudaHostAlloc");
	*cap = n;
	return 0;
}

// ---- tile planning ---------------------------------------------------------------------------------
struct Planned {
	b200sdf_tile_job t;
	uint64_t cost;
};

// Work model: one item (tile) x one segment = one unit.  A CTA's latency is what bounds the kernel
// when a batch has a few very heavy glyphs (thousands of segments), so no tile job may cost more
// than its fair share of the batch: heavy glyphs are cut into smaller rectangles (never below
// kMinItems items, so that staging — which every rectangle repeats — stays a small fraction).
constexpr uint32_t kSMs = 148;          // B200
constexpr uint32_t kJobsPerSM = 8;      // target number of equal-cost jobs per SM
constexpr uint32_t kMinItems = 16;
constexpr uint64_t kMinJobCost = 16384; // item x segment units

inline uint64_t glyph_cost(uint32_t seg_cnt, uint32_t width, uint32_t height)
{
	using namespace b200sdf;
	const uint64_t nx = (width + kTileW - 1) / kTileW, ny = (height + kTileH - 1) / kTileH;
	return nx * ny * (uint64_t)(seg_cnt + 8);
}

// A batch too small to fill the GPU (a pipeline's 64..256-glyph batches) is latency-bound: its kernel lasts as
// long as its heaviest CTA running alone on an SM, so heavy glyphs are cut finer there (down to kMinItemsSmall
// items per CTA) even though every extra rectangle repeats the staging.
constexpr uint32_t kMinItemsSmall = 8;
// ... but not finer than this per CTA: a pipeline keeps many small batches in flight, so what it needs from each
// is throughput more than latency, and every extra rectangle repeats the staging of all the glyph's segments
// (measured, C2 in 128-glyph kernels on 32 streams: 0.99 ms with 16384, 0.62 ms with 65536; one launch: 0.46 ms)
constexpr uint64_t kMinJobCostSmall = 65536;
constexpr uint64_t kSmallBatchCost = (uint64_t)kSMs * kJobsPerSM * kMinJobCost; // ~19 M units, ~1000 median glyphs

inline uint32_t items_cap(uint64_t total_cost, uint32_t seg_cnt, bool latency = false)
{
	static const uint32_t min_small = [] { // B200SDF_MIN_ITEMS_SMALL: tuning knob
		const char *e = std::getenv("B200SDF_MIN_ITEMS_SMALL");
		const int v = e ? std::atoi(e) : 0;
		return (uint32_t)(v >= 1 && v <= 64 ? v : (int)kMinItemsSmall);
	}();
	static const uint64_t min_cost_small = [] { // B200SDF_MIN_COST_SMALL: tuning knob
		const char *e = std::getenv("B200SDF_MIN_COST_SMALL");
		const long v = e ? std::atol(e) : 0;
		return (uint64_t)(v >= 256 ? v : (long)kMinJobCostSmall);
	}();
	const bool small = total_cost < kSmallBatchCost;
	// latency: the caller waits for THIS batch alone (the last batches of a pipeline): cut as finely as allowed, so the
	// kernel's critical path — its heaviest CTA — is short, whatever that costs in repeated staging
	const uint64_t floor_cost = !small ? kMinJobCost : latency ? kMinJobCost : min_cost_small;
	const uint64_t cap = std::max<uint64_t>(total_cost / (kSMs * kJobsPerSM), floor_cost);
	const uint64_t items = cap / (uint64_t)(seg_cnt + 8);
	const uint64_t floor_items = small ? (latency ? std::min<uint64_t>(4, min_small) : min_small) : kMinItems;
	return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(items, floor_items), (uint64_t)b200sdf::kMaxItems);
}

// Split one glyph into rectangles of tiles with at most max_items items each.
void plan_glyph(uint32_t src_off, uint32_t seg_cnt, uint32_t width, uint32_t height, uint64_t out_off, uint32_t job,
                uint32_t max_items, std::vector<Planned> &out)
{
	using namespace b200sdf;
	const uint32_t nx = (width + kTileW - 1) / kTileW, ny = (height + kTileH - 1) / kTileH;
	// column strips only when one tile row alone exceeds the item budget
	const uint32_t col_parts = (nx + max_items - 1) / max_items;
	const uint32_t cols_per = (nx + col_parts - 1) / col_parts;
	const uint32_t max_rows = std::max<uint32_t>(1, max_items / cols_per);
	const uint32_t row_parts = (ny + max_rows - 1) / max_rows;
	const uint32_t rows_per = (ny + row_parts - 1) / row_parts;
	for (uint32_t ty = 0; ty < ny; ty += rows_per)
		for (uint32_t tx = 0; tx < nx; tx += cols_per) {
			Planned p;
			p.t.seg_off = src_off;
			p.t.seg_cnt = seg_cnt;
			p.t.out_off = out_off;
			p.t.width = (uint16_t)width;
			p.t.height = (uint16_t)height;
			p.t.tx0 = (uint16_t)tx;
			p.t.ty0 = (uint16_t)ty;
			p.t.ntx = (uint16_t)std::min(cols_per, nx - tx);
			p.t.nty = (uint16_t)std::min(rows_per, ny - ty);
			p.t.job = job;
			p.cost = (uint64_t)p.t.ntx * p.t.nty * (uint64_t)(seg_cnt + 8);
			out.push_back(p);
		}
}

bool frame_ok(uint32_t w, uint32_t h, uint64_t out_off, uint64_t out_bytes, const char **why)
{
	if (w == 0 || h == 0 || w > B200SDF_MAX_DIM || h > B200SDF_MAX_DIM) {
		*why = "glyph job with zero or oversized width/height";
		return false;
	}
	if (out_off + (uint64_t)w * h > out_bytes) {
		*why = "glyph job bitmap exceeds out_bytes";
		return false;
	}
	return true;
}

void sort_plan(std::vector<Planned> &v)
{
	// largest first: the hardware CTA scheduler then approximates longest-processing-time-first
	std::stable_sort(v.begin(), v.end(), [](const Planned &a, const Planned &b) { return a.cost > b.cost; });
}

int plan_segments(const b200sdf_glyph_job *jobs, uint32_t n_jobs, uint32_t n_seg, uint64_t out_bytes, std::vector<Planned> &v,
                  uint64_t *pairs, const char **why)
{
	uint64_t pr = 0, total = 0;
	v.clear();
	v.reserve(n_jobs + n_jobs / 8 + 4);
	for (uint32_t i = 0; i < n_jobs; ++i)
		total += glyph_cost(jobs[i].seg_cnt, jobs[i].width, jobs[i].height);
	for (uint32_t i = 0; i < n_jobs; ++i) {
		const b200sdf_glyph_job &j = jobs[i];
		if (!frame_ok(j.width, j.height, j.out_off, out_bytes, why))
			return B200SDF_E_ARG;
		if ((uint64_t)j.seg_off + j.seg_cnt > n_seg) {
			*why = "glyph job segment range exceeds n_seg";
			return B200SDF_E_ARG;
		}
		pr += (uint64_t)j.width * j.height * j.seg_cnt;
		plan_glyph(j.seg_off, j.seg_cnt, j.width, j.height, j.out_off, B200SDF_NO_JOB, items_cap(total, j.seg_cnt), v);
	}
	sort_plan(v);
	if (pairs)
		*pairs = pr;
	return 0;
}

// curves may be null (device-resident planning): then the per-record consistency check is skipped
int plan_outlines(const b200sdf_outline_job *jobs, uint32_t n_jobs, const b200sdf_curve *curves, uint32_t n_curves,
                  uint32_t n_seg, uint64_t out_bytes, std::vector<Planned> &v, uint64_t *pairs, const char **why,
                  bool latency = false)
{
	uint64_t pr = 0, total = 0;
	v.clear();
	v.reserve(n_jobs + n_jobs / 8 + 4);
	for (uint32_t i = 0; i < n_jobs; ++i)
		total += glyph_cost(jobs[i].seg_cnt, jobs[i].width, jobs[i].height);
	for (uint32_t i = 0; i < n_jobs; ++i) {
		const b200sdf_outline_job &j = jobs[i];
		if (!frame_ok(j.width, j.height, j.out_off, out_bytes, why))
			return B200SDF_E_ARG;
		if (j.kind == B200SDF_KIND_SEGMENTS) {
			if ((uint64_t)j.src_off + j.src_cnt > n_seg || j.seg_cnt != j.src_cnt) {
				*why = "outline job (segments) range exceeds n_seg or seg_cnt != src_cnt";
				return B200SDF_E_ARG;
			}
			plan_glyph(j.src_off, j.seg_cnt, j.width, j.height, j.out_off, B200SDF_NO_JOB, items_cap(total, j.seg_cnt, latency), v);
		} else if (j.kind == B200SDF_KIND_CURVES) {
			if ((uint64_t)j.src_off + j.src_cnt > n_curves || (j.src_cnt == 0 && j.seg_cnt != 0)) {
				*why = "outline job (curves) range exceeds n_curves";
				return B200SDF_E_ARG;
			}
			if (curves) {
				// the device's binary search needs seg_off = running sum of 2^depth starting at 0
				uint64_t run = 0;
				for (uint32_t c = 0; c < j.src_cnt; ++c) {
					const b200sdf_curve &cv = curves[j.src_off + c];
					if (cv.depth > 20 || cv.seg_off != run) {
						*why = "curve record with depth > 20 or inconsistent seg_off";
						return B200SDF_E_ARG;
					}
					run += 1ull << cv.depth;
				}
				if (run != j.seg_cnt) {
					*why = "outline job seg_cnt does not match its curve records";
					return B200SDF_E_ARG;
				}
			}
			plan_glyph(j.src_off, j.seg_cnt, j.width, j.height, j.out_off, i, items_cap(total, j.seg_cnt, latency), v);
		} else {
			*why = "outline job with unknown kind";
			return B200SDF_E_ARG;
		}
		pr += (uint64_t)j.width * j.height * j.seg_cnt;
	}
	sort_plan(v);
	if (pairs)
		*pairs = pr;
	return 0;
}

int copy_plan(const std::vector<Planned> &v, b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles)
{
	*n_tiles = (uint32_t)v.size();
	if (tiles) {
		if (cap < v.size())
			return B200SDF_E_ARG;
		for (size_t i = 0; i < v.size(); ++i)
			tiles[i] = v[i].t;
	}
	return 0;
}

void launch_sdf(const void *d_segs, const void *d_curves, const void *d_ojobs, const void *d_tiles, uint32_t n_tiles,
                void *d_out, cudaStream_t stream)
{
	b200sdf::sdf_tiles_kernel<<<n_tiles, b200sdf::kThreads, 0, stream>>>(
	    reinterpret_cast<const float4 *>(d_segs), reinterpret_cast<const b200sdf_curve *>(d_curves),
	    reinterpret_cast<const b200sdf_outline_job *>(d_ojobs), reinterpret_cast<const b200sdf_tile_job *>(d_tiles),
	    reinterpret_cast<uint8_t *>(d_out));
}

// ---- slots -----------------------------------------------------------------------------------------
size_t acquire_slot(b200sdf_ctx *ctx)
{
	std::unique_lock<std::mutex> lk(ctx->mu);
	size_t si;
	for (;;) {
		for (si = 0; si < ctx->slots.size(); ++si)
			if (!ctx->slots[si].busy)
				break;
		if (si < ctx->slots.size())
			break;
		ctx->cv.wait(lk);
	}
	ctx->slots[si].busy = true;
	ctx->slots[si].generation++;
	return si;
}

uint64_t now_ns_mono()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

void report_gpu_trace(b200sdf_ctx *ctx, Slot &s)
{
	if (!s.t1 || !s.traced_tiles || !ctx->epoch)
		return;
	float a = 0.f, b = 0.f;
	if (cudaEventElapsedTime(&a, ctx->epoch, s.t0) == cudaSuccess && cudaEventElapsedTime(&b, ctx->epoch, s.t1) == cudaSuccess)
		std::fprintf(stderr, "[b200sdf gpu] submit %.1f us  kernel %.1f .. %.1f us (%.1f us, %u CTAs)  seen %.1f us  abs0 %llu\n",
		             (double)(s.host_submit_ns - ctx->epoch_host_ns) * 1e-3, a * 1e3, b * 1e3, (b - a) * 1e3, s.traced_tiles,
		             (double)(now_ns_mono() - ctx->epoch_host_ns) * 1e-3, (unsigned long long)ctx->epoch_host_ns);
	s.traced_tiles = 0;
}

int release_slot(b200sdf_ctx *ctx, Slot &s, int code)
{
	std::lock_guard<std::mutex> g(ctx->mu);
	s.busy = false;
	ctx->cv.notify_one();
	return code;
}

// The common submit path: everything already validated and planned.
int submit_planned(b200sdf_ctx *ctx, const std::vector<Planned> &plan, const b200sdf_curve *curves, uint32_t n_curves,
                   const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_outline_job *ojobs, uint32_t n_ojobs,
                   uint8_t *out, uint64_t out_bytes, uint64_t *ticket, const b200sdf_tile_job *ext_tiles = nullptr,
                   size_t n_ext_tiles = 0)
{
	// B200SDF_TRACE=1: report submissions slower than 200 us with per-phase timestamps (diagnostics)
	static const bool trace = std::getenv("B200SDF_TRACE") != nullptr;
	uint64_t tt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	auto now = [] {
		timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
	};
	if (trace)
		tt[0] = now();
	const size_t si = acquire_slot(ctx);
	if (trace)
		tt[1] = now();
	Slot &s = ctx->slots[si];
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e != cudaSuccess)
		return release_slot(ctx, s, fail_cuda(ctx, e, "cudaSetDevice"));
	const size_t n_tiles = ext_tiles ? n_ext_tiles : plan.size();
	// Size a slot's buffers to the largest request any slot of this context has seen: a slot then
	// (re)allocates at most once after the workload's largest batch has shown up, instead of stalling
	// the context with cudaFree/cudaMalloc whenever it first meets a bigger batch.
	size_t need[5] = {(size_t)n_seg * sizeof(b200sdf_segment), (size_t)n_curves * sizeof(b200sdf_curve),
	                  (size_t)n_ojobs * sizeof(b200sdf_outline_job), n_tiles * sizeof(b200sdf_tile_job), (size_t)out_bytes};
	// zero-copy legs (see PinnedRegistry): job records read, bitmaps written, in place in pinned host memory
	const bool zc = zero_copy_mode() == 1;
	const void *k_ojobs = zc && n_ojobs ? pinned_registry().device_ptr(ojobs, need[2]) : nullptr;
	void *k_out = zero_copy_mode() != 0 && out_bytes ? pinned_registry().device_ptr(out, need[4]) : nullptr;
	// Curve lists are read ONCE per CTA with a bulk copy into shared memory when they fit there; only then
	// is reading them in place (over PCIe) as good as a staged copy.  A glyph with more curve records than
	// the kernel keeps in shared memory searches them in global memory: such a batch is staged.
	const void *k_curves = nullptr;
	if (zc && n_curves && B200SDF_ZEROCOPY_CURVES) {
		uint32_t most = 0;
		for (uint32_t i = 0; i < n_ojobs; ++i)
			if (ojobs[i].kind == B200SDF_KIND_CURVES)
				most = std::max(most, ojobs[i].src_cnt);
		if (most <= (uint32_t)B200SDF_CURVE_SMEM)
			k_curves = pinned_registry().device_ptr(curves, need[1]);
	}
	if (k_ojobs)
		need[2] = 0; // no device mirror needed
	if (k_out)
		need[4] = 0;
	if (k_curves)
		need[1] = 0;
	const size_t tiles_bytes = need[3];
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		for (int k = 0; k < 5; ++k) {
			// ... rounded up to a power of two (>= 64 KiB): batch composition varies from call to call
			// (dynamic scheduling), and a mark that creeps up by a few bytes would stall the device again
			if (need[k] == 0)
				continue;
			size_t r = (size_t)64 << 10;
			while (r < need[k])
				r <<= 1;
			ctx->hwm[k] = std::max(ctx->hwm[k], r);
			need[k] = ctx->hwm[k];
		}
	}
	int rc;
	if ((rc = grow_device(ctx, s.segs, need[0], s.stream)) || (rc = grow_device(ctx, s.curves, need[1], s.stream)) ||
	    (rc = grow_device(ctx, s.ojobs, need[2], s.stream)) || (rc = grow_device(ctx, s.tiles, need[3], s.stream)) ||
	    (rc = grow_device(ctx, s.out, need[4], s.stream)) || (rc = ext_tiles ? 0 : grow_pinned(ctx, &s.h_tiles, &s.h_tiles_cap, need[3])))
		return release_slot(ctx, s, rc);
	(void)tiles_bytes;
	if (trace)
		tt[2] = now();
	const b200sdf_tile_job *ht = ext_tiles;
	if (!ext_tiles) {
		b200sdf_tile_job *mine = reinterpret_cast<b200sdf_tile_job *>(s.h_tiles);
		for (size_t i = 0; i < n_tiles; ++i)
			mine[i] = plan[i].t;
		ht = mine;
	}

#define SUB_TRY(call)                                                \
	do {                                                             \
		cudaError_t e_ = (call);                                     \
		if (e_ != cudaSuccess)                                       \
			return release_slot(ctx, s, fail_cuda(ctx, e_, #call));  \
	} while (0)
	if (n_seg)
		SUB_TRY(cudaMemcpyAsync(s.segs.p, segs, (size_t)n_seg * sizeof(b200sdf_segment), cudaMemcpyHostToDevice, s.stream));
	if (n_curves && !k_curves) {
		SUB_TRY(cudaMemcpyAsync(s.curves.p, curves, (size_t)n_curves * sizeof(b200sdf_curve), cudaMemcpyHostToDevice, s.stream));
		k_curves = s.curves.p;
	}
	if (trace)
		tt[3] = now();
	const void *k_tiles = zc && n_tiles ? pinned_registry().device_ptr(ht, n_tiles * sizeof(b200sdf_tile_job)) : nullptr;
	if (n_ojobs && !k_ojobs) {
		SUB_TRY(cudaMemcpyAsync(s.ojobs.p, ojobs, (size_t)n_ojobs * sizeof(b200sdf_outline_job), cudaMemcpyHostToDevice, s.stream));
		k_ojobs = s.ojobs.p;
	}
	if (n_tiles) {
		if (!k_tiles) {
			SUB_TRY(cudaMemcpyAsync(s.tiles.p, ht, n_tiles * sizeof(b200sdf_tile_job), cudaMemcpyHostToDevice, s.stream));
			k_tiles = s.tiles.p;
		}
		// Bitmaps may be sparse in `out` (caller-chosen out_off): bytes between bitmaps are defined (zero).
		uint64_t covered = 0;
		for (size_t i = 0; i < n_tiles; ++i)
			if (ht[i].tx0 == 0 && ht[i].ty0 == 0)
				covered += (uint64_t)ht[i].width * ht[i].height;
		static const bool gpu_trace = std::getenv("B200SDF_GPU_TRACE") != nullptr;
		if (gpu_trace) {
			if (!s.t0) {
				cudaEventCreate(&s.t0);
				cudaEventCreate(&s.t1);
			}
			{
				std::lock_guard<std::mutex> g(ctx->mu);
				if (!ctx->epoch) {
					cudaEventCreate(&ctx->epoch);
					cudaEventRecord(ctx->epoch, s.stream);
					ctx->epoch_host_ns = now();
				}
			}
			s.host_submit_ns = now();
			s.traced_tiles = (uint32_t)n_tiles;
			cudaEventRecord(s.t0, s.stream);
		}
		if (k_out) {
			if (covered < out_bytes)
				std::memset(out, 0, (size_t)out_bytes); // the caller's buffer is ours until wait()
			launch_sdf(s.segs.p, k_curves, k_ojobs, k_tiles, (uint32_t)n_tiles, k_out, s.stream);
			SUB_TRY(cudaGetLastError());
		} else {
			// staged: the device buffer mirrors the layout and is copied back whole
			if (covered < out_bytes)
				SUB_TRY(cudaMemsetAsync(s.out.p, 0, (size_t)out_bytes, s.stream));
			launch_sdf(s.segs.p, k_curves, k_ojobs, k_tiles, (uint32_t)n_tiles, s.out.p, s.stream);
			SUB_TRY(cudaGetLastError());
			SUB_TRY(cudaMemcpyAsync(out, s.out.p, (size_t)out_bytes, cudaMemcpyDeviceToHost, s.stream));
		}
	}
	if (trace)
		tt[4] = now();
	if (s.t1 && s.traced_tiles)
		cudaEventRecord(s.t1, s.stream);
	SUB_TRY(cudaEventRecord(s.done, s.stream));
#undef SUB_TRY
	if (trace) {
		tt[5] = now();
		if (tt[5] - tt[0] > 200000)
			std::fprintf(stderr, "[b200sdf trace] slow submit slot %zu: acquire %.0f grow %.0f tiles+h2d %.0f launch %.0f event %.0f us (curves %u tiles %zu)\n",
			             si, (tt[1] - tt[0]) * 1e-3, (tt[2] - tt[1]) * 1e-3, (tt[3] - tt[2]) * 1e-3, (tt[4] - tt[3]) * 1e-3,
			             (tt[5] - tt[4]) * 1e-3, n_curves, n_tiles);
	}
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (n_tiles)
			ctx->launches++;
		*ticket = ((uint64_t)s.generation << 8) | (uint64_t)si;
	}
	return 0;
}

} // namespace

extern "C" {

int b200sdf_abi_version(void) { return B200SDF_ABI_VERSION; }

int b200sdf_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int b200sdf_create(int device, uint32_t n_slots, b200sdf_ctx **out)
{
	if (!out)
		return B200SDF_E_ARG;
	*out = nullptr;
	int n = b200sdf_device_count();
	if (n <= 0)
		return B200SDF_E_NODEVICE;
	if (device < 0 || device >= n)
		return B200SDF_E_ARG;
	if (n_slots == 0)
		n_slots = 1;
	if (n_slots > 64)
		n_slots = 64;
	b200sdf_ctx *ctx = new (std::nothrow) b200sdf_ctx();
	if (!ctx)
		return B200SDF_E_NOMEM;
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess) {
		delete ctx;
		return B200SDF_E_CUDA;
	}
	g_home_device.store(device, std::memory_order_relaxed);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
		// sm_100a code only: refuse anything else instead of failing at the first launch
		delete ctx;
		return B200SDF_E_NODEVICE;
	}
	// keep freed stream-ordered allocations in the pool (grow_device): no trimming at synchronisation points
	{
		cudaMemPool_t pool = nullptr;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
			uint64_t keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
	}
	ctx->slots.resize(n_slots);
	for (auto &s : ctx->slots) {
		if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
			b200sdf_destroy(ctx);
			return B200SDF_E_CUDA;
		}
	}
	*out = ctx;
	return 0;
}

void b200sdf_destroy(b200sdf_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	for (auto &s : ctx->slots) {
		if (s.stream)
			cudaStreamSynchronize(s.stream);
		for (DevBuf *b : {&s.segs, &s.curves, &s.ojobs, &s.tiles, &s.out, &s.seg_base, &s.reqs, &s.parts, &s.frames, &s.gcurves})
			if (b->p)
				cudaFree(b->p);
		if (s.counters)
			cudaFree(s.counters);
		if (s.h_status)
			cudaFreeHost(s.h_status);
		if (s.h_tiles)
			b200sdf_free_pinned(s.h_tiles);
		if (s.done)
			cudaEventDestroy(s.done);
		if (s.stream)
			cudaStreamDestroy(s.stream);
	}
	if (ctx->d_peak)
		cudaFree(ctx->d_peak);
	for (auto &f : ctx->fonts)
		if (f.p)
			cudaFree(f.p);
	for (DevBuf *b : {&ctx->dv_gcurves, &ctx->dv_ojobs, &ctx->dv_tiles})
		if (b->p)
			cudaFree(b->p);
	if (ctx->d_font_base)
		cudaFree((void *)ctx->d_font_base);
	if (ctx->d_font_len)
		cudaFree(ctx->d_font_len);
	if (ctx->dv_counters)
		cudaFree(ctx->dv_counters);
	delete ctx;
}

const char *b200sdf_last_error(const b200sdf_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int b200sdf_device(const b200sdf_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t b200sdf_launch_count(const b200sdf_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *b200sdf_alloc_pinned(size_t bytes)
{
	void *p = nullptr;
	const int home = g_home_device.load(std::memory_order_relaxed);
	if (home >= 0) {
		int cur = -1;
		if (cudaGetDevice(&cur) != cudaSuccess || cur != home)
			cudaSetDevice(home);
	}
	if (std::getenv("B200SDF_TRACE"))
		std::fprintf(stderr, "[b200sdf trace] alloc_pinned %zu bytes\n", bytes);
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	void *dev = nullptr;
	if (cudaHostGetDevicePointer(&dev, p, 0) == cudaSuccess && dev)
		pinned_registry().add(p, bytes ? bytes : 1, dev);
	else
		cudaGetLastError(); // not mapped: the staged path is used for this buffer
	return p;
}
void b200sdf_free_pinned(void *p)
{
	if (!p)
		return;
	pinned_registry().remove(p);
	cudaFreeHost(p);
}

int b200sdf_plan_tiles(const b200sdf_glyph_job *jobs, uint32_t n_jobs, uint32_t n_seg, uint64_t out_bytes,
                       b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles, uint64_t *pairs)
{
	if ((!jobs && n_jobs) || !n_tiles)
		return B200SDF_E_ARG;
	std::vector<Planned> v;
	const char *why = "";
	int rc = plan_segments(jobs, n_jobs, n_seg, out_bytes, v, pairs, &why);
	return rc ? rc : copy_plan(v, tiles, cap, n_tiles);
}

int b200sdf_plan_outline_tiles(const b200sdf_outline_job *jobs, uint32_t n_jobs, uint32_t n_curves, uint32_t n_seg,
                               uint64_t out_bytes, b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles, uint64_t *pairs)
{
	if ((!jobs && n_jobs) || !n_tiles)
		return B200SDF_E_ARG;
	std::vector<Planned> v;
	const char *why = "";
	int rc = plan_outlines(jobs, n_jobs, nullptr, n_curves, n_seg, out_bytes, v, pairs, &why);
	return rc ? rc : copy_plan(v, tiles, cap, n_tiles);
}

int b200sdf_plan_outline_tiles_ex(const b200sdf_outline_job *jobs, uint32_t n_jobs, uint32_t n_curves, uint32_t n_seg,
                                  uint64_t out_bytes, uint32_t flags, b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles,
                                  uint64_t *pairs)
{
	if ((!jobs && n_jobs) || !n_tiles)
		return B200SDF_E_ARG;
	std::vector<Planned> v;
	const char *why = "";
	int rc = plan_outlines(jobs, n_jobs, nullptr, n_curves, n_seg, out_bytes, v, pairs, &why, (flags & B200SDF_PLAN_LATENCY) != 0);
	return rc ? rc : copy_plan(v, tiles, cap, n_tiles);
}

int b200sdf_render_outlines_device(b200sdf_ctx *ctx, const b200sdf_curve *d_curves, const b200sdf_segment *d_segs,
                                   const b200sdf_outline_job *d_jobs, const b200sdf_tile_job *d_tiles, uint32_t n_tiles,
                                   uint8_t *d_out, void *stream)
{
	if (!ctx)
		return B200SDF_E_ARG;
	if (n_tiles == 0)
		return 0;
	if (!d_tiles || !d_out)
		return fail_arg(ctx, "render_device: null device pointer");
	if ((((uintptr_t)d_segs) & 15u) != 0 || (((uintptr_t)d_curves) & 15u) != 0)
		return fail_arg(ctx, "render_device: segment / curve arrays must be 16-byte aligned");
	launch_sdf(d_segs, d_curves, d_jobs, d_tiles, n_tiles, d_out, (cudaStream_t)stream);
	CU_TRY(ctx, cudaGetLastError());
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->launches++;
	}
	return 0;
}

int b200sdf_render_device(b200sdf_ctx *ctx, const b200sdf_segment *d_segs, const b200sdf_tile_job *d_tiles,
                          uint32_t n_tiles, uint8_t *d_out, void *stream)
{
	return b200sdf_render_outlines_device(ctx, nullptr, d_segs, nullptr, d_tiles, n_tiles, d_out, stream);
}

int b200sdf_submit(b200sdf_ctx *ctx, const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_glyph_job *jobs,
                   uint32_t n_jobs, uint8_t *out, uint64_t out_bytes, uint64_t *ticket)
{
	if (!ctx || !ticket)
		return B200SDF_E_ARG;
	if ((n_seg && !segs) || (n_jobs && !jobs) || (out_bytes && !out))
		return fail_arg(ctx, "submit: null buffer");
	std::vector<Planned> plan;
	const char *why = "";
	int rc = plan_segments(jobs, n_jobs, n_seg, out_bytes, plan, nullptr, &why);
	if (rc)
		return fail_arg(ctx, why);
	return submit_planned(ctx, plan, nullptr, 0, segs, n_seg, nullptr, 0, out, out_bytes, ticket);
}

int b200sdf_submit_outlines(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves, const b200sdf_segment *segs,
                            uint32_t n_seg, const b200sdf_outline_job *jobs, uint32_t n_jobs, uint8_t *out,
                            uint64_t out_bytes, uint64_t *ticket)
{
	if (!ctx || !ticket)
		return B200SDF_E_ARG;
	if ((n_seg && !segs) || (n_curves && !curves) || (n_jobs && !jobs) || (out_bytes && !out))
		return fail_arg(ctx, "submit_outlines: null buffer");
	std::vector<Planned> plan;
	const char *why = "";
	int rc = plan_outlines(jobs, n_jobs, curves, n_curves, n_seg, out_bytes, plan, nullptr, &why);
	if (rc)
		return fail_arg(ctx, why);
	return submit_planned(ctx, plan, curves, n_curves, segs, n_seg, jobs, n_jobs, out, out_bytes, ticket);
}

int b200sdf_submit_planned(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves, const b200sdf_segment *segs,
                           uint32_t n_seg, const b200sdf_outline_job *jobs, uint32_t n_jobs, const b200sdf_tile_job *tiles,
                           uint32_t n_tiles, uint8_t *out, uint64_t out_bytes, uint64_t *ticket)
{
	if (!ctx || !ticket)
		return B200SDF_E_ARG;
	if ((n_seg && !segs) || (n_curves && !curves) || (n_jobs && !jobs) || (out_bytes && !out) || (n_tiles && !tiles))
		return fail_arg(ctx, "submit_planned: null buffer");
	// The caller's tile list drives device writes: check every entry against the arrays it indexes (one pass, a few
	// comparisons per tile job) instead of trusting it.
	for (uint32_t i = 0; i < n_tiles; ++i) {
		const b200sdf_tile_job &t = tiles[i];
		const uint32_t items = (uint32_t)t.ntx * t.nty;
		const uint32_t tiles_x = (t.width + B200SDF_TILE_W - 1) / B200SDF_TILE_W, tiles_y = (t.height + B200SDF_TILE_H - 1) / B200SDF_TILE_H;
		bool ok = t.width >= 1 && t.height >= 1 && t.width <= B200SDF_MAX_DIM && t.height <= B200SDF_MAX_DIM && items >= 1 &&
		          items <= B200SDF_MAX_ITEMS && (uint32_t)t.tx0 + t.ntx <= tiles_x && (uint32_t)t.ty0 + t.nty <= tiles_y &&
		          t.out_off <= out_bytes && (uint64_t)t.width * t.height <= out_bytes - t.out_off;
		if (ok && t.job == B200SDF_NO_JOB) {
			ok = t.seg_off <= n_seg && t.seg_cnt <= n_seg - t.seg_off;
		} else if (ok) {
			ok = t.job < n_jobs;
			if (ok) {
				const b200sdf_outline_job &j = jobs[t.job];
				ok = j.kind == B200SDF_KIND_CURVES && t.seg_off == j.src_off && j.src_off <= n_curves && j.src_cnt <= n_curves - j.src_off &&
				     t.seg_cnt == j.seg_cnt && t.width == j.width && t.height == j.height && t.out_off == j.out_off;
			}
		}
		if (!ok)
			return fail_arg(ctx, ("submit_planned: tile job " + std::to_string(i) + " does not fit the arrays it refers to").c_str());
	}
	static const std::vector<Planned> none;
	return submit_planned(ctx, none, curves, n_curves, segs, n_seg, jobs, n_jobs, out, out_bytes, ticket, tiles, n_tiles);
}

namespace {
// glyph-level batches: the device reports a tile list that was too short through the counters' mirror
int finish_slot(b200sdf_ctx *ctx, Slot &s, cudaError_t e, const char *what)
{
	const bool overflow = e == cudaSuccess && s.check_overflow && s.h_status && *(volatile uint32_t *)s.h_status != 0;
	s.check_overflow = false;
	report_gpu_trace(ctx, s);
	release_slot(ctx, s, 0);
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, what);
	if (overflow)
		return fail_arg(ctx, "submit_glyphs: tile_cap too small for this batch (see b200sdf_glyph_tile_bound)");
	return 0;
}
} // namespace

int b200sdf_wait(b200sdf_ctx *ctx, uint64_t ticket)
{
	if (!ctx)
		return B200SDF_E_ARG;
	const size_t si = (size_t)(ticket & 0xff);
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (si >= ctx->slots.size() || !ctx->slots[si].busy || ctx->slots[si].generation != (ticket >> 8)) {
			ctx->err = "wait: unknown ticket";
			return B200SDF_E_TICKET;
		}
	}
	Slot &s = ctx->slots[si];
	const cudaError_t e = cudaEventSynchronize(s.done);
	return finish_slot(ctx, s, e, "cudaEventSynchronize");
}

int b200sdf_poll(b200sdf_ctx *ctx, uint64_t ticket)
{
	if (!ctx)
		return B200SDF_E_ARG;
	const size_t si = (size_t)(ticket & 0xff);
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (si >= ctx->slots.size() || !ctx->slots[si].busy || ctx->slots[si].generation != (ticket >> 8)) {
			ctx->err = "poll: unknown ticket";
			return B200SDF_E_TICKET;
		}
	}
	Slot &s = ctx->slots[si];
	const cudaError_t e = cudaEventQuery(s.done);
	if (e == cudaErrorNotReady)
		return 0;
	const int rc = finish_slot(ctx, s, e, "cudaEventQuery");
	return rc ? rc : 1;
}

int b200sdf_render(b200sdf_ctx *ctx, const b200sdf_segment *segs, uint32_t n_seg, const b200sdf_glyph_job *jobs,
                   uint32_t n_jobs, uint8_t *out, uint64_t out_bytes)
{
	uint64_t t = 0;
	int rc = b200sdf_submit(ctx, segs, n_seg, jobs, n_jobs, out, out_bytes, &t);
	if (rc)
		return rc;
	return b200sdf_wait(ctx, t);
}

int b200sdf_flatten_outlines(b200sdf_ctx *ctx, const b200sdf_curve *curves, uint32_t n_curves, const b200sdf_outline_job *jobs,
                             uint32_t n_jobs, b200sdf_segment *out_segs, uint64_t n_out)
{
	if (!ctx)
		return B200SDF_E_ARG;
	if ((n_curves && !curves) || (n_jobs && !jobs) || (n_out && !out_segs))
		return fail_arg(ctx, "flatten_outlines: null buffer");
	// validate exactly like submit (bitmap frames are irrelevant here: give each job a fake 1x1 frame check)
	std::vector<uint64_t> base(n_jobs + 1, 0);
	for (uint32_t i = 0; i < n_jobs; ++i) {
		const b200sdf_outline_job &j = jobs[i];
		if (j.kind != B200SDF_KIND_CURVES || (uint64_t)j.src_off + j.src_cnt > n_curves)
			return fail_arg(ctx, "flatten_outlines: only CURVES jobs with valid ranges");
		uint64_t run = 0;
		for (uint32_t c = 0; c < j.src_cnt; ++c) {
			const b200sdf_curve &cv = curves[j.src_off + c];
			if (cv.depth > 20 || cv.seg_off != run)
				return fail_arg(ctx, "curve record with depth > 20 or inconsistent seg_off");
			run += 1ull << cv.depth;
		}
		if (run != j.seg_cnt)
			return fail_arg(ctx, "outline job seg_cnt does not match its curve records");
		base[i + 1] = base[i] + run;
	}
	if (base[n_jobs] != n_out)
		return fail_arg(ctx, "flatten_outlines: n_out != sum of seg_cnt");
	if (n_jobs == 0 || n_out == 0)
		return 0;
	const size_t si = acquire_slot(ctx);
	Slot &s = ctx->slots[si];
	int rc;
	if ((rc = grow_device(ctx, s.curves, (size_t)n_curves * sizeof(b200sdf_curve), s.stream)) ||
	    (rc = grow_device(ctx, s.ojobs, (size_t)n_jobs * sizeof(b200sdf_outline_job), s.stream)) ||
	    (rc = grow_device(ctx, s.seg_base, (size_t)(n_jobs + 1) * sizeof(uint64_t), s.stream)) ||
	    (rc = grow_device(ctx, s.segs, (size_t)n_out * sizeof(b200sdf_segment), s.stream)))
		return release_slot(ctx, s, rc);
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(s.curves.p, curves, (size_t)n_curves * sizeof(b200sdf_curve), cudaMemcpyHostToDevice, s.stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(s.ojobs.p, jobs, (size_t)n_jobs * sizeof(b200sdf_outline_job), cudaMemcpyHostToDevice, s.stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(s.seg_base.p, base.data(), (size_t)(n_jobs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream);
	if (e == cudaSuccess) {
		b200sdf::flatten_kernel<<<n_jobs, 256, 0, s.stream>>>(
		    reinterpret_cast<const b200sdf_curve *>(s.curves.p), reinterpret_cast<const b200sdf_outline_job *>(s.ojobs.p),
		    reinterpret_cast<const uint64_t *>(s.seg_base.p), reinterpret_cast<float4 *>(s.segs.p));
		e = cudaGetLastError();
	}
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(out_segs, s.segs.p, (size_t)n_out * sizeof(b200sdf_segment), cudaMemcpyDeviceToHost, s.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s.stream);
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->launches++;
	}
	release_slot(ctx, s, 0);
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "flatten_outlines");
	return 0;
}

int b200sdf_measure_fp32_peak(b200sdf_ctx *ctx, int reps, double *tflops, double *ms_out)
{
	if (!ctx || !tflops)
		return B200SDF_E_ARG;
	CU_TRY(ctx, cudaSetDevice(ctx->device));
	int sms = 0;
	CU_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
	const int blocks = sms * 8, threads = 256, iters = 4096;
	if (!ctx->d_peak)
		CU_TRY(ctx, cudaMalloc(&ctx->d_peak, (size_t)blocks * threads * sizeof(float)));
	cudaEvent_t a, b;
	CU_TRY(ctx, cudaEventCreate(&a));
	CU_TRY(ctx, cudaEventCreate(&b));
	const bool packed = reps < 0; // negative reps: measure the packed FFMA2 variant instead
	reps = std::max(1, reps < 0 ? -reps : reps);
	double best = 1e30;
	for (int r = 0; r < reps + 2; ++r) {
		cudaEventRecord(a, 0);
		if (packed)
			b200sdf::fp32x2_peak_kernel<<<blocks, threads>>>(ctx->d_peak, iters, 0.999f, 0.001f);
		else
			b200sdf::fp32_peak_kernel<<<blocks, threads>>>(ctx->d_peak, iters, 0.999f, 0.001f);
		cudaEventRecord(b, 0);
		cudaError_t e = cudaEventSynchronize(b);
		if (e != cudaSuccess) {
			cudaEventDestroy(a);
			cudaEventDestroy(b);
			return fail_cuda(ctx, e, "fp32_peak_kernel");
		}
		float ms = 0;
		cudaEventElapsedTime(&ms, a, b);
		if (r >= 2 && ms < best)
			best = ms;
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	const double flop = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
	*tflops = flop / (best * 1e-3) / 1e12;
	if (ms_out)
		*ms_out = best;
	return 0;
}

/* ---- glyph-level path: glyf decoding, metrics and tile planning on the device ---------------------- */
namespace {

// the batch's tile jobs: one list of tile_cap entries per cost class
constexpr size_t kTileScratchPerJob = sizeof(b200sdf_tile_job) * b200sdf::kTileClasses;
void set_tile_scratch(b200sdf::DecodeParams &P, void *base, uint32_t tile_cap)
{
	P.tiles = reinterpret_cast<b200sdf_tile_job *>(base);
	P.tile_cap = tile_cap;
}

void single_sub(b200sdf::DecodeParams &P, const void *reqs, uint32_t n_reqs, const void *parts, uint32_t n_parts,
                const void *host_curves, uint32_t n_curves, uint32_t n_seg, uint32_t curve_slots, void *frames, void *out,
                uint64_t out_bytes)
{
	std::memset(&P, 0, sizeof(P));
	b200sdf::SubBatch &S = P.sub[0];
	S.reqs = reinterpret_cast<const b200sdf_glyph_req *>(reqs), S.n_reqs = n_reqs, S.req_base = 0;
	S.parts = reinterpret_cast<const b200sdf_glyph_part *>(parts), S.n_parts = n_parts;
	S.host_curves = reinterpret_cast<const b200sdf_curve *>(host_curves), S.n_host_curves = n_curves, S.n_host_segs = n_seg;
	S.frames = reinterpret_cast<b200sdf_glyph_frame *>(frames);
	S.out_addr = (uint64_t)(uintptr_t)out, S.out_bytes = out_bytes;
	S.seg_base = 0, S.curve_base = 0, S.curve_slots = curve_slots;
	P.n_sub = 1, P.n_reqs = n_reqs;
}

uint32_t persistent_grid(uint32_t n_reqs)
{
	static const uint32_t forced = [] { // B200SDF_PERSISTENT_GRID: experiments
		const char *e = std::getenv("B200SDF_PERSISTENT_GRID");
		const long v = e ? std::atol(e) : 0;
		return (uint32_t)(v > 0 ? v : 0);
	}();
	if (forced)
		return forced;
	// one wave of resident CTAs at most; small batches do not need the whole machine
	const uint64_t want = std::max<uint64_t>(kSMs, (uint64_t)n_reqs * 2u);
	return (uint32_t)std::min<uint64_t>((uint64_t)kSMs * B200SDF_PERSISTENT_MIN_CTAS, want);
}

// Largest tile job (item x segment units) before a glyph is cut into several rectangles: the batch's fair share per
// resident CTA, like the host planner's items_cap — but from the caller's ESTIMATE of the batch's cost (the device has
// not decoded anything when the launch is made).  Every rectangle repeats the staging of all the glyph's segments, so
// small batches are not cut below kMinJobCostSmall.  B200SDF_GLYPH_COST_CAP overrides (experiments).
uint32_t glyph_cost_cap(uint64_t est_cost)
{
	static const uint32_t forced = [] {
		const char *e = std::getenv("B200SDF_GLYPH_COST_CAP");
		const long x = e ? std::atol(e) : 0;
		return (uint32_t)(x >= 1024 ? x : 0);
	}();
	if (forced)
		return forced;
	const uint64_t share = est_cost / ((uint64_t)kSMs * B200SDF_PERSISTENT_MIN_CTAS);
	return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(share, kMinJobCostSmall), 1u << 30);
}
constexpr uint32_t kGlyphMinItems = 8;

// The counters are zero when this is called (zeroed when they were allocated, and again by the last CTA of every
// SDF kernel): a batch is two launches, nothing else.
void launch_glyph_pipeline(const b200sdf::DecodeParams &P, const void *d_segs, uint32_t *d_status, cudaStream_t stream, cudaEvent_t mid)
{
	using namespace b200sdf;
	glyf_decode_kernel<<<(P.n_reqs + kGlyfWarps - 1) / kGlyfWarps, kGlyfThreads, 0, stream>>>(P);
	if (mid)
		cudaEventRecord(mid, stream);
	// (the tile jobs carry absolute bitmap addresses — several batches, several bitmap areas: the kernel's base is 0)
	static const int form = [] { // B200SDF_SDF_KERNEL=persistent|strided (experiments)
		const char *e = std::getenv("B200SDF_SDF_KERNEL");
		return e && e[0] == 'p' ? 0 : 1;
	}();
	static const uint32_t per_glyph = [] { // B200SDF_STRIDED_PER_GLYPH: CTAs per glyph request in the strided form's grid
		const char *e = std::getenv("B200SDF_STRIDED_PER_GLYPH");
		const long v = e ? std::atol(e) : 0;
		return (uint32_t)(v > 0 ? v : 1);
	}();
	if (form == 0) {
		sdf_tiles_persistent_kernel<<<persistent_grid(P.n_reqs), kThreads, 0, stream>>>(
		    reinterpret_cast<const float4 *>(d_segs), P.curves, P.ojobs, P.tiles, P.tile_cap, P.counters, d_status, nullptr);
	} else {
		const uint64_t grid = std::min<uint64_t>((uint64_t)P.tile_cap, std::max<uint64_t>(kSMs, (uint64_t)P.n_reqs * per_glyph));
		sdf_tiles_strided_kernel<<<(uint32_t)grid, kThreads, 0, stream>>>(
		    reinterpret_cast<const float4 *>(d_segs), P.curves, P.ojobs, P.tiles, P.tile_cap, P.counters, d_status, nullptr);
	}
}

int ensure_font_tables(b200sdf_ctx *ctx)
{
	if (ctx->d_font_base)
		return 0;
	CU_TRY(ctx, cudaMalloc((void **)&ctx->d_font_base, kMaxFonts * sizeof(void *)));
	CU_TRY(ctx, cudaMalloc((void **)&ctx->d_font_len, kMaxFonts * sizeof(uint64_t)));
	CU_TRY(ctx, cudaMemset((void *)ctx->d_font_base, 0, kMaxFonts * sizeof(void *)));
	CU_TRY(ctx, cudaMemset(ctx->d_font_len, 0, kMaxFonts * sizeof(uint64_t)));
	CU_TRY(ctx, cudaDeviceSynchronize()); // the memsets run on the legacy default stream: done before any table entry is written or read
	return 0;
}

// plain (synchronising) growth for the context-wide scratch of the device-resident entry point
int grow_plain(b200sdf_ctx *ctx, DevBuf &b, size_t need)
{
	if (need <= b.cap)
		return 0;
	if (b.p)
		cudaFree(b.p);
	b.p = nullptr, b.cap = 0;
	const size_t n = grown(need, 0);
	cudaError_t e = cudaMalloc(&b.p, n);
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "cudaMalloc");
	b.cap = n;
	return 0;
}

} // namespace

int b200sdf_reserve(b200sdf_ctx *ctx)
{
	using namespace b200sdf;
	if (!ctx)
		return B200SDF_E_ARG;
	CU_TRY(ctx, cudaSetDevice(ctx->device));
	size_t g[9], h[5];
	{
		std::lock_guard<std::mutex> lk(ctx->mu);
		std::memcpy(g, ctx->ghwm, sizeof(g));
		std::memcpy(h, ctx->hwm, sizeof(h));
	}
	for (size_t si = 0; si < ctx->slots.size(); ++si) {
		Slot &s = ctx->slots[si];
		{
			std::lock_guard<std::mutex> lk(ctx->mu);
			if (s.busy)
				continue;
			s.busy = true;
		}
		int rc = 0;
		// glyph-level submissions: [0] segments [1] host curves [2] requests [3] parts [4] frames [5] bitmaps [6] curve
		// scratch [7] tile lists [8] outline jobs; outline-level ones: segments, curves, jobs, tiles, bitmaps
		const size_t want[10] = {std::max(g[0], h[0]), std::max(g[1], h[1]), g[2], g[3], g[4], std::max(g[5], h[4]), g[6], std::max(g[7], h[3]),
		                         std::max(g[8], h[2]), 0};
		DevBuf *bufs[9] = {&s.segs, &s.curves, &s.reqs, &s.parts, &s.frames, &s.out, &s.gcurves, &s.tiles, &s.ojobs};
		for (int k = 0; k < 9 && rc == 0; ++k)
			rc = grow_device(ctx, *bufs[k], want[k], s.stream);
		if (rc == 0 && !s.counters && g[6]) {
			cudaError_t e;
			if ((e = cudaMalloc((void **)&s.counters, sizeof(BatchCounters))) != cudaSuccess ||
			    (e = cudaMemsetAsync(s.counters, 0, sizeof(BatchCounters), s.stream)) != cudaSuccess ||
			    (e = cudaHostAlloc((void **)&s.h_status, 64, cudaHostAllocPortable | cudaHostAllocMapped)) != cudaSuccess ||
			    (e = cudaHostGetDevicePointer((void **)&s.d_status, s.h_status, 0)) != cudaSuccess)
				rc = fail_cuda(ctx, e, "cudaMalloc(counters)");
		}
		release_slot(ctx, s, 0);
		if (rc)
			return rc;
	}
	return 0;
}

int b200sdf_reserve_glyphs(b200sdf_ctx *ctx, uint32_t n_reqs, uint32_t n_seg, uint32_t curve_slots, uint32_t tile_cap)
{
	using namespace b200sdf;
	if (!ctx)
		return B200SDF_E_ARG;
	// the same sizes, rounded the same way, as b200sdf_submit_glyph_batches asks for (buffers in pinned memory are read
	// in place: only the device scratch counts)
	const size_t need[9] = {(size_t)n_seg * sizeof(b200sdf_segment), 0, 0, 0, 0, 0,
	                        (size_t)std::max(1u, curve_slots) * sizeof(b200sdf_curve), (size_t)std::max(1u, tile_cap) * kTileScratchPerJob,
	                        (size_t)n_reqs * sizeof(b200sdf_outline_job)};
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		for (int k = 0; k < 9; ++k) {
			if (need[k] == 0)
				continue;
			// (a bound, not a need: a font with absurd header boxes can make it astronomical — beyond 256 MiB per buffer
			// and slot the buffer is left to grow when a submission really needs it)
			if (need[k] > ((size_t)256 << 20))
				continue;
			size_t r = (size_t)64 << 10;
			while (r < need[k])
				r <<= 1;
			ctx->ghwm[k] = std::max(ctx->ghwm[k], r);
		}
	}
	return b200sdf_reserve(ctx);
}

int b200sdf_font_upload(b200sdf_ctx *ctx, const uint8_t *glyf, uint64_t len, uint32_t *handle)
{
	if (!ctx || !handle || (len && !glyf))
		return B200SDF_E_ARG;
	CU_TRY(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_font_tables(ctx);
	if (rc)
		return rc;
	void *d = nullptr;
	CU_TRY(ctx, cudaMalloc(&d, (size_t)len + 16)); // the decoder reads at most one byte past a checked position
	if (len)
		CU_TRY(ctx, cudaMemcpy(d, glyf, (size_t)len, cudaMemcpyHostToDevice));
	CU_TRY(ctx, cudaMemset((uint8_t *)d + len, 0, 16));
	CU_TRY(ctx, cudaStreamSynchronize(0)); // (legacy default stream: the slots' non-blocking streams do not wait for it)
	uint32_t h;
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (ctx->fonts.size() >= kMaxFonts) {
			cudaFree(d);
			ctx->err = "font_upload: too many fonts resident on this context";
			return B200SDF_E_NOMEM;
		}
		h = (uint32_t)ctx->fonts.size();
		b200sdf_ctx::FontBlob fb;
		fb.p = d, fb.len = len;
		ctx->fonts.push_back(fb);
	}
	const void *dp = d;
	CU_TRY(ctx, cudaMemcpy((void *)(ctx->d_font_base + h), &dp, sizeof(void *), cudaMemcpyHostToDevice));
	CU_TRY(ctx, cudaMemcpy(ctx->d_font_len + h, &len, sizeof(uint64_t), cudaMemcpyHostToDevice));
	CU_TRY(ctx, cudaStreamSynchronize(0));
	*handle = h;
	return 0;
}

uint32_t b200sdf_glyph_tile_bound(uint32_t width, uint32_t height)
{
	// plan_tiles_dev: at most ceil(nx / min_items) column strips x ny rows of rectangles, whatever the glyph costs
	const uint32_t nx = (width + B200SDF_TILE_W - 1) / B200SDF_TILE_W, ny = (height + B200SDF_TILE_H - 1) / B200SDF_TILE_H;
	return std::max(1u, ((nx + kGlyphMinItems - 1) / kGlyphMinItems) * ny);
}

int b200sdf_submit_glyph_batches(b200sdf_ctx *ctx, const b200sdf_glyph_batch *batches, uint32_t n_batches, uint64_t est_cost,
                                 uint64_t *ticket)
{
	using namespace b200sdf;
	if (!ctx || !ticket)
		return B200SDF_E_ARG;
	if (n_batches == 0 || n_batches > (uint32_t)kMaxSubBatches || !batches)
		return fail_arg(ctx, "submit_glyph_batches: between 1 and B200SDF_MAX_BATCHES batches per submission");
	uint64_t n_reqs = 0, n_seg = 0, curve_slots = 0, tile_cap = 0, gen_total = 0;
	uint64_t gen_slots[kMaxSubBatches];
	for (uint32_t b = 0; b < n_batches; ++b) {
		const b200sdf_glyph_batch &B = batches[b];
		if ((B.n_reqs && (!B.reqs || !B.frames)) || (B.n_parts && !B.parts) || (B.n_curves && !B.curves) || (B.n_seg && !B.segs) ||
		    (B.out_bytes && !B.out))
			return fail_arg(ctx, "submit_glyphs: null buffer");
		n_reqs += B.n_reqs, n_seg += B.n_seg, curve_slots += B.curve_slots, tile_cap += std::max(1u, B.tile_cap);
		// kind PATH (host-recorded outlines with cubics, only where there are host-recorded records at all): the segments
		// the device makes of them go behind the uploaded ones; the requests say how much room that takes
		uint64_t gen = 0;
		if (B.n_curves)
			for (uint32_t i = 0; i < B.n_reqs; ++i)
				if (B.reqs[i].kind == B200SDF_KIND_PATH)
					gen = std::max<uint64_t>(gen, (uint64_t)B.reqs[i].curve_off + B.reqs[i].curve_cap);
		gen_slots[b] = gen;
		gen_total += gen;
	}
	if (n_reqs > 0xffffffffull || n_seg + gen_total > 0xffffffffull || curve_slots > 0xffffffffull || tile_cap > 0xffffffffull)
		return fail_arg(ctx, "submit_glyphs: submission too large");
	const size_t si = acquire_slot(ctx);
	Slot &s = ctx->slots[si];
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e != cudaSuccess)
		return release_slot(ctx, s, fail_cuda(ctx, e, "cudaSetDevice"));
	int rc = ensure_font_tables(ctx);
	if (rc)
		return release_slot(ctx, s, rc);
	if (!s.counters) {
		// (zeroed ON THE SLOT'S STREAM: a plain cudaMemset runs on the legacy default stream, which a non-blocking stream
		// does not wait for — the first batch's decode kernel would count its tile jobs into counters that are wiped
		// a moment later)
		if ((e = cudaMalloc((void **)&s.counters, sizeof(BatchCounters))) != cudaSuccess ||
		    (e = cudaMemsetAsync(s.counters, 0, sizeof(BatchCounters), s.stream)) != cudaSuccess ||
		    (e = cudaHostAlloc((void **)&s.h_status, 64, cudaHostAllocPortable | cudaHostAllocMapped)) != cudaSuccess ||
		    (e = cudaHostGetDevicePointer((void **)&s.d_status, s.h_status, 0)) != cudaSuccess)
			return release_slot(ctx, s, fail_cuda(ctx, e, "cudaMalloc(counters)"));
	}
	*s.h_status = 0;
	const bool zc = zero_copy_mode() == 1;
	const bool one = n_batches == 1;
	const b200sdf_glyph_batch &B0 = batches[0];
	// What has to exist on the device: [0] segments (always staged: the raw-segment path reads them through the TMA
	// unit) [1] host curves [2] requests [3] parts [4] frames [5] bitmaps — each only for a single batch whose buffer
	// is not pinned + mapped — [6] curve scratch [7] tile lists [8] outline jobs
	size_t need[9] = {(size_t)(n_seg + gen_total) * sizeof(b200sdf_segment), 0, 0, 0, 0, 0,
	                  (size_t)std::max<uint64_t>(1, curve_slots) * sizeof(b200sdf_curve), (size_t)tile_cap * kTileScratchPerJob,
	                  (size_t)n_reqs * sizeof(b200sdf_outline_job)};
	DecodeParams P;
	std::memset(&P, 0, sizeof(P));
	bool staged_frames = false, staged_out = false;
	{
		uint32_t req_base = 0, seg_base = 0, curve_base = 0, gen_base = (uint32_t)n_seg;
		for (uint32_t b = 0; b < n_batches; ++b) {
			const b200sdf_glyph_batch &B = batches[b];
			SubBatch &S = P.sub[b];
			const size_t bytes[6] = {0, (size_t)B.n_curves * sizeof(b200sdf_curve), (size_t)B.n_reqs * sizeof(b200sdf_glyph_req),
			                         (size_t)B.n_parts * sizeof(b200sdf_glyph_part), (size_t)B.n_reqs * sizeof(b200sdf_glyph_frame),
			                         (size_t)B.out_bytes};
			const void *k_hc = zc && B.n_curves ? pinned_registry().device_ptr(B.curves, bytes[1]) : nullptr;
			const void *k_rq = zc && B.n_reqs ? pinned_registry().device_ptr(B.reqs, bytes[2]) : nullptr;
			const void *k_pt = zc && B.n_parts ? pinned_registry().device_ptr(B.parts, bytes[3]) : nullptr;
			void *k_fr = zc && B.n_reqs ? pinned_registry().device_ptr(B.frames, bytes[4]) : nullptr;
			void *k_out = zero_copy_mode() != 0 && B.out_bytes ? pinned_registry().device_ptr(B.out, bytes[5]) : nullptr;
			if (!one && ((B.n_curves && !k_hc) || (B.n_reqs && (!k_rq || !k_fr)) || (B.n_parts && !k_pt) || (B.out_bytes && !k_out)))
				return release_slot(ctx, s, fail_arg(ctx, "submit_glyph_batches: several batches in one submission need buffers from b200sdf_alloc_pinned"));
			if (one) { // a single batch may live in pageable memory: staged copies
				need[1] = k_hc ? 0 : bytes[1], need[2] = k_rq ? 0 : bytes[2], need[3] = k_pt ? 0 : bytes[3];
				need[4] = k_fr ? 0 : bytes[4], need[5] = k_out ? 0 : bytes[5];
			}
			S.host_curves = reinterpret_cast<const b200sdf_curve *>(k_hc);
			S.reqs = reinterpret_cast<const b200sdf_glyph_req *>(k_rq);
			S.parts = reinterpret_cast<const b200sdf_glyph_part *>(k_pt);
			S.frames = reinterpret_cast<b200sdf_glyph_frame *>(k_fr);
			S.out_addr = (uint64_t)(uintptr_t)k_out;
			S.out_bytes = B.out_bytes;
			S.n_reqs = B.n_reqs, S.req_base = req_base;
			S.n_parts = B.n_parts, S.n_host_curves = B.n_curves, S.n_host_segs = B.n_seg;
			S.seg_base = seg_base, S.curve_base = curve_base, S.curve_slots = B.curve_slots;
			S.gen_base = gen_base, S.gen_slots = (uint32_t)gen_slots[b];
			req_base += B.n_reqs, seg_base += B.n_seg, curve_base += B.curve_slots, gen_base += (uint32_t)gen_slots[b];
		}
	}
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		for (int k = 0; k < 9; ++k) {
			if (need[k] == 0)
				continue;
			size_t r = (size_t)64 << 10;
			while (r < need[k])
				r <<= 1;
			ctx->ghwm[k] = std::max(ctx->ghwm[k], r);
			need[k] = ctx->ghwm[k];
		}
		P.n_fonts = (uint32_t)ctx->fonts.size();
	}
	if ((rc = grow_device(ctx, s.segs, need[0], s.stream)) || (rc = grow_device(ctx, s.curves, need[1], s.stream)) ||
	    (rc = grow_device(ctx, s.reqs, need[2], s.stream)) || (rc = grow_device(ctx, s.parts, need[3], s.stream)) ||
	    (rc = grow_device(ctx, s.frames, need[4], s.stream)) || (rc = grow_device(ctx, s.out, need[5], s.stream)) ||
	    (rc = grow_device(ctx, s.gcurves, need[6], s.stream)) || (rc = grow_device(ctx, s.tiles, need[7], s.stream)) ||
	    (rc = grow_device(ctx, s.ojobs, need[8], s.stream)))
		return release_slot(ctx, s, rc);
#define SUB_TRY(call)                                                \
	do {                                                             \
		cudaError_t e_ = (call);                                     \
		if (e_ != cudaSuccess)                                       \
			return release_slot(ctx, s, fail_cuda(ctx, e_, #call));  \
	} while (0)
	for (uint32_t b = 0; b < n_batches; ++b)
		if (batches[b].n_seg)
			SUB_TRY(cudaMemcpyAsync(reinterpret_cast<b200sdf_segment *>(s.segs.p) + P.sub[b].seg_base, batches[b].segs,
			                        (size_t)batches[b].n_seg * sizeof(b200sdf_segment), cudaMemcpyHostToDevice, s.stream));
	if (one) {
		SubBatch &S = P.sub[0];
		if (B0.n_curves && !S.host_curves) {
			SUB_TRY(cudaMemcpyAsync(s.curves.p, B0.curves, (size_t)B0.n_curves * sizeof(b200sdf_curve), cudaMemcpyHostToDevice, s.stream));
			S.host_curves = reinterpret_cast<const b200sdf_curve *>(s.curves.p);
		}
		if (B0.n_reqs && !S.reqs) {
			SUB_TRY(cudaMemcpyAsync(s.reqs.p, B0.reqs, (size_t)B0.n_reqs * sizeof(b200sdf_glyph_req), cudaMemcpyHostToDevice, s.stream));
			S.reqs = reinterpret_cast<const b200sdf_glyph_req *>(s.reqs.p);
		}
		if (B0.n_parts && !S.parts) {
			SUB_TRY(cudaMemcpyAsync(s.parts.p, B0.parts, (size_t)B0.n_parts * sizeof(b200sdf_glyph_part), cudaMemcpyHostToDevice, s.stream));
			S.parts = reinterpret_cast<const b200sdf_glyph_part *>(s.parts.p);
		}
		if (B0.n_reqs && !S.frames) {
			S.frames = reinterpret_cast<b200sdf_glyph_frame *>(s.frames.p);
			staged_frames = true;
		}
		if (B0.out_bytes && !S.out_addr) {
			S.out_addr = (uint64_t)(uintptr_t)s.out.p;
			staged_out = true;
		}
	}
	static const bool gpu_trace = std::getenv("B200SDF_GPU_TRACE") != nullptr;
	if (n_reqs) {
		if (gpu_trace) {
			if (!s.t0) {
				cudaEventCreate(&s.t0);
				cudaEventCreate(&s.t1);
			}
			{
				std::lock_guard<std::mutex> g(ctx->mu);
				if (!ctx->epoch) {
					cudaEventCreate(&ctx->epoch);
					cudaEventRecord(ctx->epoch, s.stream);
					ctx->epoch_host_ns = now_ns_mono();
				}
			}
			s.host_submit_ns = now_ns_mono();
			s.traced_tiles = (uint32_t)n_reqs;
			cudaEventRecord(s.t0, s.stream);
		}
		P.n_sub = n_batches;
		P.n_reqs = (uint32_t)n_reqs;
		P.font_base = ctx->d_font_base;
		P.font_len = ctx->d_font_len;
		P.segs = reinterpret_cast<float4 *>(s.segs.p);
		P.curves = reinterpret_cast<b200sdf_curve *>(s.gcurves.p);
		P.ojobs = reinterpret_cast<b200sdf_outline_job *>(s.ojobs.p);
		set_tile_scratch(P, s.tiles.p, (uint32_t)tile_cap);
		P.counters = s.counters;
		P.cost_cap = glyph_cost_cap(est_cost);
		P.min_items = kGlyphMinItems;
		launch_glyph_pipeline(P, s.segs.p, s.d_status, s.stream, nullptr);
		SUB_TRY(cudaGetLastError());
		if (staged_frames)
			SUB_TRY(cudaMemcpyAsync(B0.frames, s.frames.p, (size_t)B0.n_reqs * sizeof(b200sdf_glyph_frame), cudaMemcpyDeviceToHost, s.stream));
		if (staged_out)
			SUB_TRY(cudaMemcpyAsync(B0.out, s.out.p, (size_t)B0.out_bytes, cudaMemcpyDeviceToHost, s.stream));
		s.check_overflow = true;
		if (s.t1 && s.traced_tiles)
			cudaEventRecord(s.t1, s.stream);
	}
	SUB_TRY(cudaEventRecord(s.done, s.stream));
#undef SUB_TRY
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (n_reqs)
			ctx->launches += 2;
		*ticket = ((uint64_t)s.generation << 8) | (uint64_t)si;
	}
	return 0;
}

int b200sdf_submit_glyphs(b200sdf_ctx *ctx, const b200sdf_glyph_req *reqs, uint32_t n_reqs, const b200sdf_glyph_part *parts,
                          uint32_t n_parts, const b200sdf_curve *curves, uint32_t n_curves, const b200sdf_segment *segs,
                          uint32_t n_seg, uint32_t curve_slots, uint32_t tile_cap, uint64_t est_cost, b200sdf_glyph_frame *frames,
                          uint8_t *out, uint64_t out_bytes, uint64_t *ticket)
{
	b200sdf_glyph_batch B;
	B.reqs = reqs, B.n_reqs = n_reqs, B.parts = parts, B.n_parts = n_parts, B.curves = curves, B.n_curves = n_curves;
	B.segs = segs, B.n_seg = n_seg, B.curve_slots = curve_slots, B.tile_cap = tile_cap, B.frames = frames, B.out = out;
	B.out_bytes = out_bytes;
	return b200sdf_submit_glyph_batches(ctx, &B, 1, est_cost, ticket);
}

int b200sdf_render_glyphs_device(b200sdf_ctx *ctx, const b200sdf_glyph_req *d_reqs, uint32_t n_reqs,
                                 const b200sdf_glyph_part *d_parts, uint32_t n_parts, const b200sdf_curve *d_curves,
                                 uint32_t n_curves, const b200sdf_segment *d_segs, uint32_t n_seg, uint32_t curve_slots,
                                 uint32_t tile_cap, uint64_t est_cost, b200sdf_glyph_frame *d_frames, uint8_t *d_out,
                                 uint64_t out_bytes, void *stream, void *mid_event)
{
	using namespace b200sdf;
	if (!ctx)
		return B200SDF_E_ARG;
	if (n_reqs == 0)
		return 0;
	if (!d_reqs || !d_frames || (n_parts && !d_parts) || (out_bytes && !d_out))
		return fail_arg(ctx, "render_glyphs_device: null device pointer");
	CU_TRY(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_font_tables(ctx);
	if (rc)
		return rc;
	if (tile_cap == 0)
		tile_cap = 1;
	// context-wide scratch: grown outside any timed loop by the first call of a given size (synchronising)
	if ((rc = grow_plain(ctx, ctx->dv_gcurves, (size_t)std::max(1u, curve_slots) * sizeof(b200sdf_curve))) ||
	    (rc = grow_plain(ctx, ctx->dv_ojobs, (size_t)n_reqs * sizeof(b200sdf_outline_job))) ||
	    (rc = grow_plain(ctx, ctx->dv_tiles, (size_t)tile_cap * kTileScratchPerJob)))
		return rc;
	if (!ctx->dv_counters) {
		CU_TRY(ctx, cudaMalloc((void **)&ctx->dv_counters, sizeof(BatchCounters)));
		CU_TRY(ctx, cudaMemsetAsync(ctx->dv_counters, 0, sizeof(BatchCounters), (cudaStream_t)stream)); // ordered before the kernels below
	}
	DecodeParams P;
	single_sub(P, d_reqs, n_reqs, d_parts, n_parts, d_curves, n_curves, n_seg, curve_slots, d_frames, d_out, out_bytes);
	P.font_base = ctx->d_font_base, P.font_len = ctx->d_font_len;
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		P.n_fonts = (uint32_t)ctx->fonts.size();
	}
	P.curves = reinterpret_cast<b200sdf_curve *>(ctx->dv_gcurves.p);
	P.ojobs = reinterpret_cast<b200sdf_outline_job *>(ctx->dv_ojobs.p);
	set_tile_scratch(P, ctx->dv_tiles.p, tile_cap);
	P.counters = ctx->dv_counters;
	P.cost_cap = glyph_cost_cap(est_cost);
	P.min_items = kGlyphMinItems;
	launch_glyph_pipeline(P, d_segs, nullptr, (cudaStream_t)stream, (cudaEvent_t)mid_event);
	CU_TRY(ctx, cudaGetLastError());
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->launches += 2;
	}
	return 0;
}

int b200sdf_decode_glyphs(b200sdf_ctx *ctx, const b200sdf_glyph_req *reqs, uint32_t n_reqs, const b200sdf_glyph_part *parts,
                          uint32_t n_parts, const b200sdf_curve *curves, uint32_t n_curves, uint32_t n_seg, uint32_t curve_slots,
                          uint64_t est_cost, b200sdf_glyph_frame *frames, b200sdf_outline_job *jobs_out, b200sdf_curve *curves_out,
                          uint32_t *n_tiles_out, b200sdf_tile_job *tiles_out, uint32_t tiles_cap)
{
	using namespace b200sdf;
	if (!ctx)
		return B200SDF_E_ARG;
	if (n_reqs == 0)
		return 0;
	if (!reqs || !frames || (n_parts && !parts))
		return fail_arg(ctx, "decode_glyphs: null buffer");
	CU_TRY(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_font_tables(ctx);
	if (rc)
		return rc;
	uint64_t tile_cap64 = 0, out_bytes = 0;
	for (uint32_t i = 0; i < n_reqs; ++i) {
		tile_cap64 += 64; // decode only: nothing is rendered, a generous fixed allowance per glyph
		out_bytes = std::max<uint64_t>(out_bytes, reqs[i].out_off + reqs[i].out_cap);
	}
	const uint32_t tile_cap = (uint32_t)std::min<uint64_t>(tile_cap64, 1u << 24);
	if (n_curves && !curves)
		return fail_arg(ctx, "decode_glyphs: null curve array");
	void *d_reqs = nullptr, *d_parts = nullptr, *d_frames = nullptr, *d_curves = nullptr, *d_ojobs = nullptr, *d_tiles = nullptr, *d_ctr = nullptr;
	void *d_hcurves = nullptr;
	cudaError_t e = cudaSuccess;
	auto alloc = [&](void **p, size_t n) {
		if (e == cudaSuccess)
			e = cudaMalloc(p, std::max<size_t>(n, 16));
	};
	alloc(&d_reqs, (size_t)n_reqs * sizeof(b200sdf_glyph_req));
	alloc(&d_parts, (size_t)n_parts * sizeof(b200sdf_glyph_part));
	alloc(&d_frames, (size_t)n_reqs * sizeof(b200sdf_glyph_frame));
	alloc(&d_curves, (size_t)std::max(1u, curve_slots) * sizeof(b200sdf_curve));
	alloc(&d_ojobs, (size_t)n_reqs * sizeof(b200sdf_outline_job));
	alloc(&d_tiles, (size_t)tile_cap * kTileScratchPerJob);
	alloc(&d_ctr, sizeof(BatchCounters));
	alloc(&d_hcurves, (size_t)n_curves * sizeof(b200sdf_curve));
	if (e == cudaSuccess && n_curves)
		e = cudaMemcpy(d_hcurves, curves, (size_t)n_curves * sizeof(b200sdf_curve), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(d_reqs, reqs, (size_t)n_reqs * sizeof(b200sdf_glyph_req), cudaMemcpyHostToDevice);
	if (e == cudaSuccess && n_parts)
		e = cudaMemcpy(d_parts, parts, (size_t)n_parts * sizeof(b200sdf_glyph_part), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemset(d_ctr, 0, sizeof(BatchCounters));
	if (e == cudaSuccess)
		e = cudaMemset(d_curves, 0, (size_t)std::max(1u, curve_slots) * sizeof(b200sdf_curve));
	if (e == cudaSuccess) {
		DecodeParams P;
		single_sub(P, d_reqs, n_reqs, d_parts, n_parts, d_hcurves, n_curves, n_seg, curve_slots, d_frames, nullptr, out_bytes);
		P.font_base = ctx->d_font_base, P.font_len = ctx->d_font_len;
		{
			std::lock_guard<std::mutex> g(ctx->mu);
			P.n_fonts = (uint32_t)ctx->fonts.size();
		}
		P.curves = reinterpret_cast<b200sdf_curve *>(d_curves);
		P.ojobs = reinterpret_cast<b200sdf_outline_job *>(d_ojobs);
		set_tile_scratch(P, d_tiles, tile_cap);
		P.counters = reinterpret_cast<BatchCounters *>(d_ctr);
		P.cost_cap = glyph_cost_cap(est_cost), P.min_items = kGlyphMinItems;
		glyf_decode_kernel<<<(n_reqs + kGlyfWarps - 1) / kGlyfWarps, kGlyfThreads>>>(P);
		e = cudaGetLastError();
		if (e == cudaSuccess)
			e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess)
		e = cudaMemcpy(frames, d_frames, (size_t)n_reqs * sizeof(b200sdf_glyph_frame), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && jobs_out)
		e = cudaMemcpy(jobs_out, d_ojobs, (size_t)n_reqs * sizeof(b200sdf_outline_job), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && curves_out && curve_slots)
		e = cudaMemcpy(curves_out, d_curves, (size_t)curve_slots * sizeof(b200sdf_curve), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && (n_tiles_out || tiles_out)) {
		// the tile jobs in the order the SDF kernel would claim them: class after class
		BatchCounters h;
		e = cudaMemcpy(&h, d_ctr, sizeof(h), cudaMemcpyDeviceToHost);
		uint32_t n = 0;
		for (int c = 0; c < kTileClasses && e == cudaSuccess; ++c) {
			const uint32_t k = std::min(h.class_count[c], tile_cap);
			if (tiles_out && n < tiles_cap && k)
				e = cudaMemcpy(tiles_out + n, reinterpret_cast<b200sdf_tile_job *>(d_tiles) + (size_t)c * tile_cap,
				               (size_t)std::min(k, tiles_cap - n) * sizeof(b200sdf_tile_job), cudaMemcpyDeviceToHost);
			n += k;
		}
		if (n_tiles_out)
			*n_tiles_out = n;
	}
	for (void *q : {d_reqs, d_parts, d_frames, d_curves, d_ojobs, d_tiles, d_ctr, d_hcurves})
		if (q)
			cudaFree(q);
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->launches++;
	}
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "decode_glyphs");
	return 0;
}

} // extern "C"
