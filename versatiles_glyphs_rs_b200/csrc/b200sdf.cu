// b200sdf.cu — C ABI around the sm_100a SDF kernel (include/b200sdf.h).
//
// Context = one GPU + a pool of slots; a slot = one CUDA stream + device buffers + pinned staging
// for the tile-job list + one completion event.  submit() plans tiles on the host, then enqueues
// H2D(segments) -> H2D(tiles) -> kernel -> D2H(bitmaps) on the slot's stream and returns.
// This is the async batch pipeline that replaces the serial per-block loop of
// FontManager::render_glyphs (reference src/font/manager.rs:104-121).
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "sdf_kernel.cuh"

namespace {

struct Slot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	void *d_segs = nullptr;
	size_t segs_cap = 0;
	void *d_tiles = nullptr;
	size_t tiles_cap = 0;
	void *d_out = nullptr;
	size_t out_cap = 0;
	b200sdf_tile_job *h_tiles = nullptr; // pinned
	size_t h_tiles_cap = 0;
	bool busy = false;
	uint64_t generation = 0;
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
};

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

#define CU_TRY(ctx, call)                                  \
	do {                                                   \
		cudaError_t e_ = (call);                           \
		if (e_ != cudaSuccess)                             \
			return fail_cuda((ctx), e_, #call);            \
	} while (0)

template <typename F> int grow(b200sdf_ctx *ctx, void **p, size_t *cap, size_t need, F alloc_free)
{
	if (need <= *cap)
		return 0;
	size_t n = std::max(need, *cap + *cap / 2);
	n = (n + 255) & ~size_t(255);
	return alloc_free(ctx, p, cap, n);
}

int grow_device(b200sdf_ctx *ctx, void **p, size_t *cap, size_t need)
{
	return grow(ctx, p, cap, need, [](b200sdf_ctx *c, void **pp, size_t *cc, size_t n) -> int {
		if (*pp)
			cudaFree(*pp);
		*pp = nullptr;
		*cc = 0;
		cudaError_t e = cudaMalloc(pp, n);
		if (e != cudaSuccess)
			return fail_cuda(c, e, "cudaMalloc");
		*cc = n;
		return 0;
	});
}

int grow_pinned(b200sdf_ctx *ctx, void **p, size_t *cap, size_t need)
{
	return grow(ctx, p, cap, need, [](b200sdf_ctx *c, void **pp, size_t *cc, size_t n) -> int {
		if (*pp)
			cudaFreeHost(*pp);
		*pp = nullptr;
		*cc = 0;
		cudaError_t e = cudaMallocHost(pp, n);
		if (e != cudaSuccess)
			return fail_cuda(c, e, "cudaMallocHost");
		*cc = n;
		return 0;
	});
}

// Split one glyph into rectangles of tiles with at most kMaxItems items each.
struct Planned {
	b200sdf_tile_job t;
	uint64_t cost;
};

void plan_glyph(const b200sdf_glyph_job &j, std::vector<Planned> &out)
{
	using namespace b200sdf;
	const uint32_t nx = (j.width + kTileW - 1) / kTileW, ny = (j.height + kTileH - 1) / kTileH;
	// column strips only when one tile row alone exceeds the item budget
	const uint32_t col_parts = (nx + kMaxItems - 1) / kMaxItems;
	const uint32_t cols_per = (nx + col_parts - 1) / col_parts;
	const uint32_t max_rows = std::max<uint32_t>(1, kMaxItems / cols_per);
	const uint32_t row_parts = (ny + max_rows - 1) / max_rows;
	const uint32_t rows_per = (ny + row_parts - 1) / row_parts;
	for (uint32_t ty = 0; ty < ny; ty += rows_per)
		for (uint32_t tx = 0; tx < nx; tx += cols_per) {
			Planned p;
			p.t.seg_off = j.seg_off;
			p.t.seg_cnt = j.seg_cnt;
			p.t.out_off = j.out_off;
			p.t.width = (uint16_t)j.width;
			p.t.height = (uint16_t)j.height;
			p.t.tx0 = (uint16_t)tx;
			p.t.ty0 = (uint16_t)ty;
			p.t.ntx = (uint16_t)std::min(cols_per, nx - tx);
			p.t.nty = (uint16_t)std::min(rows_per, ny - ty);
			p.t.reserved = 0;
			p.cost = (uint64_t)p.t.ntx * p.t.nty * (uint64_t)(j.seg_cnt + 8);
			out.push_back(p);
		}
}

int plan_impl(const b200sdf_glyph_job *jobs, uint32_t n_jobs, uint32_t n_seg, uint64_t out_bytes, std::vector<Planned> &v,
              uint64_t *pairs, const char **why)
{
	uint64_t pr = 0;
	v.clear();
	v.reserve(n_jobs + n_jobs / 8 + 4);
	for (uint32_t i = 0; i < n_jobs; ++i) {
		const b200sdf_glyph_job &j = jobs[i];
		if (j.width == 0 || j.height == 0 || j.width > B200SDF_MAX_DIM || j.height > B200SDF_MAX_DIM) {
			*why = "glyph job with zero or oversized width/height";
			return B200SDF_E_ARG;
		}
		if ((uint64_t)j.seg_off + j.seg_cnt > n_seg) {
			*why = "glyph job segment range exceeds n_seg";
			return B200SDF_E_ARG;
		}
		if (j.out_off + (uint64_t)j.width * j.height > out_bytes) {
			*why = "glyph job bitmap exceeds out_bytes";
			return B200SDF_E_ARG;
		}
		pr += (uint64_t)j.width * j.height * j.seg_cnt;
		plan_glyph(j, v);
	}
	// largest first: the hardware CTA scheduler then approximates longest-processing-time-first
	std::stable_sort(v.begin(), v.end(), [](const Planned &a, const Planned &b) { return a.cost > b.cost; });
	if (pairs)
		*pairs = pr;
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
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
		// sm_100a code only: refuse anything else instead of failing at the first launch
		delete ctx;
		return B200SDF_E_NODEVICE;
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
		if (s.d_segs)
			cudaFree(s.d_segs);
		if (s.d_tiles)
			cudaFree(s.d_tiles);
		if (s.d_out)
			cudaFree(s.d_out);
		if (s.h_tiles)
			cudaFreeHost(s.h_tiles);
		if (s.done)
			cudaEventDestroy(s.done);
		if (s.stream)
			cudaStreamDestroy(s.stream);
	}
	if (ctx->d_peak)
		cudaFree(ctx->d_peak);
	delete ctx;
}

const char *b200sdf_last_error(const b200sdf_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int b200sdf_device(const b200sdf_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t b200sdf_launch_count(const b200sdf_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *b200sdf_alloc_pinned(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
void b200sdf_free_pinned(void *p)
{
	if (p)
		cudaFreeHost(p);
}

int b200sdf_plan_tiles(const b200sdf_glyph_job *jobs, uint32_t n_jobs, uint32_t n_seg, uint64_t out_bytes,
                       b200sdf_tile_job *tiles, uint32_t cap, uint32_t *n_tiles, uint64_t *pairs)
{
	if ((!jobs && n_jobs) || !n_tiles)
		return B200SDF_E_ARG;
	std::vector<Planned> v;
	const char *why = "";
	int rc = plan_impl(jobs, n_jobs, n_seg, out_bytes, v, pairs, &why);
	if (rc)
		return rc;
	*n_tiles = (uint32_t)v.size();
	if (tiles) {
		if (cap < v.size())
			return B200SDF_E_ARG;
		for (size_t i = 0; i < v.size(); ++i)
			tiles[i] = v[i].t;
	}
	return 0;
}

int b200sdf_render_device(b200sdf_ctx *ctx, const b200sdf_segment *d_segs, const b200sdf_tile_job *d_tiles,
                          uint32_t n_tiles, uint8_t *d_out, void *stream)
{
	if (!ctx)
		return B200SDF_E_ARG;
	if (n_tiles == 0)
		return 0;
	if (!d_tiles || !d_out)
		return fail_arg(ctx, "render_device: null device pointer");
	if (((uintptr_t)d_segs & 15u) != 0)
		return fail_arg(ctx, "render_device: segment array must be 16-byte aligned");
	b200sdf::sdf_tiles_kernel<<<n_tiles, b200sdf::kThreads, 0, (cudaStream_t)stream>>>(
	    reinterpret_cast<const float4 *>(d_segs), d_tiles, d_out);
	CU_TRY(ctx, cudaGetLastError());
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->launches++;
	}
	return 0;
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
	int rc = plan_impl(jobs, n_jobs, n_seg, out_bytes, plan, nullptr, &why);
	if (rc)
		return fail_arg(ctx, why);

	size_t si;
	{
		std::unique_lock<std::mutex> lk(ctx->mu);
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
	}
	Slot &s = ctx->slots[si];
	auto release = [&](int code) {
		std::lock_guard<std::mutex> g(ctx->mu);
		s.busy = false;
		ctx->cv.notify_one();
		return code;
	};
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e != cudaSuccess)
		return release(fail_cuda(ctx, e, "cudaSetDevice"));
	const size_t n_tiles = plan.size();
	void *ht = s.h_tiles;
	if ((rc = grow_device(ctx, &s.d_segs, &s.segs_cap, (size_t)n_seg * sizeof(b200sdf_segment))) ||
	    (rc = grow_device(ctx, &s.d_tiles, &s.tiles_cap, n_tiles * sizeof(b200sdf_tile_job))) ||
	    (rc = grow_device(ctx, &s.d_out, &s.out_cap, (size_t)out_bytes)) ||
	    (rc = grow_pinned(ctx, &ht, &s.h_tiles_cap, n_tiles * sizeof(b200sdf_tile_job)))) {
		s.h_tiles = (b200sdf_tile_job *)ht;
		return release(rc);
	}
	s.h_tiles = (b200sdf_tile_job *)ht;
	for (size_t i = 0; i < n_tiles; ++i)
		s.h_tiles[i] = plan[i].t;

#define SUB_TRY(call)                                        \
	do {                                                     \
		cudaError_t e_ = (call);                             \
		if (e_ != cudaSuccess)                               \
			return release(fail_cuda(ctx, e_, #call));       \
	} while (0)
	if (n_seg)
		SUB_TRY(cudaMemcpyAsync(s.d_segs, segs, (size_t)n_seg * sizeof(b200sdf_segment), cudaMemcpyHostToDevice, s.stream));
	if (n_tiles) {
		SUB_TRY(cudaMemcpyAsync(s.d_tiles, s.h_tiles, n_tiles * sizeof(b200sdf_tile_job), cudaMemcpyHostToDevice, s.stream));
		b200sdf::sdf_tiles_kernel<<<(unsigned)n_tiles, b200sdf::kThreads, 0, s.stream>>>(
		    reinterpret_cast<const float4 *>(s.d_segs), reinterpret_cast<const b200sdf_tile_job *>(s.d_tiles),
		    reinterpret_cast<uint8_t *>(s.d_out));
		SUB_TRY(cudaGetLastError());
		// bitmaps may be sparse in `out` (caller-chosen out_off); the device buffer mirrors the layout
		SUB_TRY(cudaMemcpyAsync(out, s.d_out, (size_t)out_bytes, cudaMemcpyDeviceToHost, s.stream));
	}
	SUB_TRY(cudaEventRecord(s.done, s.stream));
#undef SUB_TRY
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		if (n_tiles)
			ctx->launches++;
		*ticket = ((uint64_t)s.generation << 8) | (uint64_t)si;
	}
	return 0;
}

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
	cudaError_t e = cudaEventSynchronize(s.done);
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		s.busy = false;
		ctx->cv.notify_one();
	}
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "cudaEventSynchronize");
	return 0;
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
	if (reps < 1)
		reps = 1;
	double best = 1e30;
	for (int r = 0; r < reps + 2; ++r) {
		cudaEventRecord(a, 0);
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

} // extern "C"
