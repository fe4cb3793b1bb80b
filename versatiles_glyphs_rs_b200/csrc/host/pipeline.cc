// pipeline.cc — FontManager::render_glyphs (reference src/font/manager.rs:81-125) as an asynchronous batch pipeline:
// workers record outlines / plan / encode, one thread talks to CUDA; see DESIGN.md 3.4.
#include "font.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <malloc.h>
#include <pthread.h>
#include <sched.h>
#include <thread>

namespace vgb {

namespace {
inline void cpu_pause()
{
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#else
	std::this_thread::yield();
#endif
}
// CPUs this process may run on (a cpuset / taskset can be narrower than the machine)
inline int usable_cpus()
{
	cpu_set_t set;
	CPU_ZERO(&set);
	if (sched_getaffinity(0, sizeof(set), &set) == 0) {
		const int n = CPU_COUNT(&set);
		if (n > 0)
			return n;
	}
	return (int)std::max(1u, std::thread::hardware_concurrency());
}
inline uint64_t now_ns()
{
	return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
	    .count();
}

// Persistent host workers (the reference keeps a rayon pool alive the same way): creating and joining
// 16 threads costs more than rendering the whole Noto merge on the GPU.  run() is not re-entrant across
// concurrent callers; they serialise on run_mu_.
class WorkerPool {
  public:
	static WorkerPool &instance()
	{
		static WorkerPool *p = new WorkerPool(); // intentionally leaked: workers must outlive static destructors
		return *p;
	}
	// run fn(0..n-1) on the pool and, meanwhile, `here` on the calling thread (which is already running on a
	// core: the pipeline's submitter must not wait for a sleeping pool thread to be scheduled)
	void run(int n, const std::function<void(int)> &fn, const std::function<void()> &here = nullptr)
	{
		std::lock_guard<std::mutex> serial(run_mu_);
		{
			std::unique_lock<std::mutex> lk(mu_);
			while ((int)threads_.size() < n) {
				const int id = (int)threads_.size();
				threads_.emplace_back([this, id] { loop(id); });
				pin(threads_.back(), id);
			}
			fn_ = &fn;
			want_ = n;
			remaining_ = n;
			++generation_;
		}
		cv_.notify_all();
		if (here)
			here();
		std::unique_lock<std::mutex> lk(mu_);
		done_cv_.wait(lk, [this] { return remaining_ == 0; });
		fn_ = nullptr;
	}

  private:
	// VGB_PIN_WORKERS=1: worker i stays on the (i+1)-th CPU this process may use (the first is left to the
	// calling thread, the pipeline's submitter) — fewer migrations, steadier step times.  Off by default.
	static void pin(std::thread &t, int id)
	{
		static const bool on = [] {
			const char *e = std::getenv("VGB_PIN_WORKERS");
			return e && e[0] == '1';
		}();
		if (!on)
			return;
		cpu_set_t allowed;
		CPU_ZERO(&allowed);
		if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0)
			return;
		std::vector<int> cpus;
		for (int c = 0; c < CPU_SETSIZE; ++c)
			if (CPU_ISSET(c, &allowed))
				cpus.push_back(c);
		if ((int)cpus.size() < 2)
			return;
		cpu_set_t one;
		CPU_ZERO(&one);
		CPU_SET(cpus[(size_t)(id + 1) % cpus.size()], &one);
		pthread_setaffinity_np(t.native_handle(), sizeof(one), &one);
	}
	void loop(int id)
	{
		uint64_t seen = 0;
		for (;;) {
			const std::function<void(int)> *fn;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&] { return generation_ != seen && id < want_; });
				seen = generation_;
				fn = fn_;
			}
			(*fn)(id);
			{
				std::lock_guard<std::mutex> lk(mu_);
				if (--remaining_ == 0)
					done_cv_.notify_all();
			}
		}
	}
	std::mutex mu_, run_mu_;
	std::condition_variable cv_, done_cv_;
	std::vector<std::thread> threads_;
	const std::function<void(int)> *fn_ = nullptr;
	int want_ = 0, remaining_ = 0;
	uint64_t generation_ = 0;
};
} // namespace

void FontManager::shard_owners(uint32_t n_shards, std::vector<uint16_t> &owner, std::vector<uint64_t> *loads) const
{
	constexpr uint32_t kBlocks = 0x10000 / GLYPH_BLOCK_SIZE;
	struct Item {
		uint64_t cost;
		uint32_t index;
	};
	if (n_shards == 0)
		n_shards = 1;
	std::vector<Item> items;
	items.reserve(fonts_.size() * kBlocks);
	uint32_t fi = 0;
	for (const auto &kv : fonts_) {
		const std::vector<uint64_t> &costs = kv.second.block_costs();
		for (uint32_t i = 0; i < kBlocks; ++i)
			items.push_back(Item{costs[i], fi * kBlocks + i});
		++fi;
	}
	// longest processing time first: heaviest task to the least loaded shard (ties: lowest task index, lowest shard)
	std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.cost > b.cost; });
	std::vector<uint64_t> load(n_shards, 0);
	owner.assign(items.size(), 0);
	for (const Item &it : items) {
		uint32_t best = 0;
		for (uint32_t k = 1; k < n_shards; ++k)
			if (load[k] < load[best])
				best = k;
		load[best] += it.cost;
		owner[it.index] = (uint16_t)best;
	}
	if (loads)
		*loads = std::move(load);
}

bool FontManager::render_glyphs(Writer &writer, const Renderer &renderer, std::string *err, RenderStats *stats,
                                uint32_t shard, uint32_t n_shards, int threads) const
{
	// glibc serves allocations >= 128 KiB (every full block's PBF) with mmap: fresh pages, i.e. ~30 page faults per
	// file, every call.  Keep such blocks on the heap, where the pages freed after one call are reused by the next.
	// (Process-wide malloc tuning, therefore only on request: VGB_TUNE_MALLOC=1.)
	static const bool malloc_tuned = [] {
		// opt-in (VGB_TUNE_MALLOC=1): a library entry point does not change process-wide allocator settings on its own
		const char *e = std::getenv("VGB_TUNE_MALLOC");
		if (!e || e[0] != '1')
			return false;
#if defined(__GLIBC__)
		mallopt(M_MMAP_THRESHOLD, 64 << 20);
		mallopt(M_TRIM_THRESHOLD, 512 << 20);
		return true;
#else
		return false;
#endif
	}();
	(void)malloc_tuned;
	// A task is a slot range of one block holding at most kPartGlyphs (64) glyphs: full blocks are split so that
	// no worker is stuck recording 256 outlines while the others (and the GPU) run dry.  The parts of a
	// block are encoded independently (Fontstack.glyphs entries) and the worker that finishes the last
	// one assembles and writes the file.
	// (When the device decodes the outlines itself a glyph costs the host a quarter of a microsecond: blocks stay whole.)
	static const size_t kPartGlyphsEnv = [] { // tuning knob (measured 16 / 32 / 64 on C2 with host-side recording)
		const char *e = std::getenv("VGB_PART_GLYPHS");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 256 ? v : 0);
	}();
	const bool glyf_mode = renderer.flatten() == Flatten::Glyf && renderer.mode() == Renderer::Mode::Cuda;
	const size_t kPartGlyphs = kPartGlyphsEnv ? kPartGlyphsEnv : (glyf_mode ? 256 : 64);
	struct BlockState {
		const std::string *name = nullptr;
		const GlyphBlock *blk = nullptr;
		std::vector<std::vector<uint8_t>> parts;
		std::atomic<uint32_t> remaining{0};
	};
	struct Todo {
		BlockState *bs;
		uint32_t part, slot0, slot1, glyphs;
	};
	const uint64_t t_begin = now_ns();
	if (n_shards == 0)
		n_shards = 1;
	constexpr uint32_t kBlocks = 0x10000 / GLYPH_BLOCK_SIZE; // wrapper.rs:53-76: always 256 BMP blocks per font
	std::vector<std::unique_ptr<BlockState[]>> fonts_blocks;
	std::vector<Todo> tasks;
	tasks.reserve(fonts_.size() * kBlocks + 64);
	size_t total_glyphs = 0;
	// Shards: longest-processing-time-first over the (font, block) tasks by estimated cost (the task list
	// is manager.rs:88-97).  Every rank computes the same assignment from the same cost tables and keeps its own part.
	std::vector<uint16_t> owner; // owner[font * 256 + block]
	uint64_t cost_total = 0, cost_mine = 0;
	if (n_shards > 1) {
		std::vector<uint64_t> load;
		shard_owners(n_shards, owner, &load);
		for (uint64_t l : load)
			cost_total += l;
		cost_mine = shard < n_shards ? load[shard] : 0;
	}
	// What the device will be rendering while any one batch of this call is in flight is (most of) the call: heavy glyphs
	// are cut into rectangles by the call's fair share per resident CTA, not by a single batch's (every rectangle
	// repeats the staging of all the glyph's segments — C3 in 13 batches: 6.5 ms cut per batch, 2.9 ms cut per call).
	uint64_t call_units = 0;
	uint32_t index = 0;
	for (const auto &kv : fonts_) {
		if (!writer.write_directory(kv.first + "/", err))
			return false;
		fonts_blocks.emplace_back(new BlockState[kBlocks]);
		BlockState *bsv = fonts_blocks.back().get();
		const std::vector<GlyphBlock> &table = kv.second.blocks(); // = get_blocks(), built once per font set
		for (uint32_t i = 0; i < kBlocks; ++i) {
			bsv[i].name = &kv.first;
			bsv[i].blk = &table[i];
		}
		for (uint32_t i = 0; i < kBlocks; ++i) {
			const uint32_t task = index++;
			if (n_shards > 1 && owner[task] != shard)
				continue;
			BlockState *bs = &bsv[i];
			total_glyphs += bs->blk->len();
			if (glyf_mode && !bs->blk->is_empty())
				call_units += kv.second.block_units()[i];
			uint32_t part = 0, slot0 = 0, count = 0;
			if (bs->blk->len() > kPartGlyphs) {
				for (uint32_t k = 0; k < GLYPH_BLOCK_SIZE; ++k) {
					if (!bs->blk->font_of((uint8_t)k))
						continue;
					if (count == kPartGlyphs) {
						tasks.push_back(Todo{bs, part++, slot0, k, count});
						slot0 = k, count = 0;
					}
					++count;
				}
			} else {
				count = (uint32_t)bs->blk->len(); // small and empty blocks: one part
			}
			tasks.push_back(Todo{bs, part++, slot0, GLYPH_BLOCK_SIZE, count});
			bs->parts.resize(part);
			bs->remaining.store(part, std::memory_order_relaxed);
		}
	}

	// Empty blocks first (they need no GPU and would otherwise all be written after the last wait), then the
	// fullest parts: dynamic scheduling ends with the cheap ones (the reference's rayon par_iter makes no
	// order promise either, manager.rs:117-118).
	std::stable_sort(tasks.begin(), tasks.end(), [](const Todo &a, const Todo &b) {
		const uint32_t ka = a.glyphs ? a.glyphs : 0xffffffffu, kb = b.glyphs ? b.glyphs : 0xffffffffu;
		return ka > kb;
	});

	// `threads` = host threads this call may use, the calling thread included (0 = one per core).  With fourteen or more,
	// the calling thread becomes the dedicated submitter (all CUDA traffic, see below) and the rest are workers; with
	// fewer — several ranks sharing a box's cores — a spinning submitter would burn a large share of them, so every
	// thread is a worker and whoever is free pumps the queues (one at a time); with one, everything runs inline.
	static const int kDedicatedMin = [] { // VGB_DEDICATED_MIN: tuning knob
		const char *e = std::getenv("VGB_DEDICATED_MIN");
		const int v = e ? std::atoi(e) : 0;
		return v >= 2 ? v : 14; // measured on C2: cooperative 2.36 / 1.53 ms at 6 / 12 threads against 2.62 / 1.66 dedicated; 16: 1.40 against 1.31
	}();
	int total_threads = 1;
	if (parallel_) {
		total_threads = threads > 0 ? threads : usable_cpus();
		// No more threads than there is host work for: a worker should have a few hundred microseconds of glyphs to
		// prepare and encode, or waking it costs more than it contributes (measured: 32 threads were slower than 16 on
		// the 6480-glyph Noto merge).  VGB_GLYPHS_PER_WORKER overrides the share; the pool itself is unbounded.
		static const size_t per_worker = [] {
			const char *e = std::getenv("VGB_GLYPHS_PER_WORKER");
			const long v = e ? std::atol(e) : 0;
			return (size_t)(v >= 1 ? v : 0);
		}();
		const size_t share = per_worker ? per_worker : (renderer.flatten() == Flatten::Glyf ? 768 : 384);
		const size_t useful = std::max<size_t>(1, (total_glyphs + share - 1) / share);
		total_threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)total_threads, useful + 1));
	}
	const bool dedicated = total_threads >= kDedicatedMin;
	const int workers = dedicated ? total_threads - 1 : total_threads;
	// One submission carries whole blocks until it holds about `target` glyphs: small jobs keep one
	// block per submission (parallelism), big jobs amortise the per-submission cost.
	static const size_t kBatchesPerWorkerEnv = [] { // tuning knob (measured 1..6 on C2)
		const char *e = std::getenv("VGB_BATCHES_PER_WORKER");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 64 ? v : 0);
	}();
	const size_t kBatchesPerWorker = kBatchesPerWorkerEnv ? kBatchesPerWorkerEnv : 4;
	constexpr int kEarlyWorkers = 4;
	static const size_t kOpenGlyphs = [] { // glyphs in a worker's opening batch, at least (tuning knob)
		const char *e = std::getenv("VGB_OPEN_GLYPHS");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 4096 ? v : 192);
	}();
	static const bool kLatencyTail = [] { // VGB_LATENCY_TAIL=1: plan each worker's last batch for latency
		const char *e = std::getenv("VGB_LATENCY_TAIL"); // (measured on C2: 1.30-1.35 ms with, 1.26-1.34 without: off)
		return e && e[0] == '1';
	}();
	const size_t target = glyf_mode ? std::min<size_t>(4096, std::max<size_t>(192, (total_glyphs + (size_t)workers * kBatchesPerWorker - 1) /
	                                                                                   ((size_t)workers * kBatchesPerWorker)))
	                                : std::min<size_t>(2048, std::max<size_t>(1, total_glyphs / ((size_t)workers * kBatchesPerWorker)));
	std::atomic<size_t> next{0};
	std::atomic<size_t> glyphs_taken{0};
	std::atomic<bool> failed{false};
	std::mutex writer_mutex, err_mutex;
	std::vector<RenderStats> per_worker((size_t)workers);

	auto fail = [&](const std::string &msg) {
		std::lock_guard<std::mutex> g(err_mutex);
		if (!failed.exchange(true) && err)
			*err = msg;
	};

	// VGB_TRACE=1: print a per-worker timeline (us since the call began) to stderr — diagnostics only
	const bool trace = std::getenv("VGB_TRACE") != nullptr;
	std::vector<std::vector<std::pair<char, uint64_t>>> events((size_t)workers + 1); // [workers] = the submitter
	const uint64_t t_setup = now_ns();

	struct Part {
		const Todo *todo;
		size_t g0, g1;
	};
	struct Flight {
		std::unique_ptr<GlyphBatch> batch;
		std::vector<Part> parts;
		uint64_t ticket = 0;
		std::vector<Flight *> mates; // the other batches of the same submission (they finish together)
		// A finished batch is encoded PART BY PART by whichever workers are free (the batches of a merged submission
		// come back together, at the end of a call all at once: one worker per batch left most of them idle for the
		// call's last 0.1 ms).  next_part is claimed under the queue lock; the first claimer finalizes the batch
		// (fin: 0 not yet, 1 in progress, 2 done, 3 failed) while the others wait for it; whoever finishes the last
		// part gives the batch back.
		size_t next_part = 0;
		std::atomic<size_t> parts_left{0};
		std::atomic<int> fin{0};
	};
	std::mutex qm;
	std::condition_variable qcv;
	std::deque<Flight *> submit_q, done_q;
	size_t outstanding = 0; // batches submitted (or queued for it) and not yet encoded
	int workers_done = 0;
	std::atomic<uint64_t> submit_ns{0};
	static const int kSpinBudget = [] { // pause iterations an idle worker spins before it sleeps (VGB_SPIN: tuning knob)
		const char *e = std::getenv("VGB_SPIN");
		const int v = e ? std::atoi(e) : 0;
		return v > 0 ? v : 4000;
	}();
	static const int kWaitMs = [] { // safety-net timeout of the workers' condition waits (diagnostics knob)
		const char *e = std::getenv("VGB_WAIT_MS");
		const int v = e ? std::atoi(e) : 0;
		return v > 0 ? v : 50;
	}();
	std::atomic<uint64_t> done_seq{0}; // batches handed back so far (workers spin on it before they sleep)
	std::atomic<int> sleepers{0};      // workers blocked in qcv.wait (the submitter only pays for a wake-up then)
	std::function<void(bool)> pump;
	const bool inline_pump = workers == 1;      // single thread: the worker pumps, and may block in the pump
	const bool coop_pump = !dedicated && !inline_pump; // few threads: whoever is free pumps (never blocking)
	std::mutex pump_mu;                         // cooperative mode: one pumping thread at a time
	// pump from a worker (inline / cooperative modes); `idle` = the caller has nothing else to do
	// (idle_spins: the caller's count of consecutive idle calls.  Idle workers must not spin forever: with every core
	// busy spinning, the one thread that holds pump_mu or qm can be pre-empted for a whole scheduler slice — a 70 ms
	// step was measured — so they yield after a while and then nap.)
	auto worker_pump = [&](bool idle, unsigned *idle_spins) {
		if (inline_pump) {
			pump(idle);
			return;
		}
		{
			std::unique_lock<std::mutex> pl(pump_mu, std::try_to_lock);
			if (pl.owns_lock())
				pump(false);
		}
		if (!idle)
			return;
		const unsigned n = idle_spins ? ++*idle_spins : 0;
		if (n > 2000)
			std::this_thread::sleep_for(std::chrono::microseconds(50));
		else if (n > 200)
			std::this_thread::yield();
		else
			for (int k = 0; k < 64; ++k)
				cpu_pause();
	};
	// never more batches on their way than the renderer has slots for (submit would block the submitter)
	const size_t max_outstanding =
	    renderer.mode() == Renderer::Mode::Cuda ? std::max<size_t>(2, renderer.slots()) : (size_t)(2 * workers + 2);

	renderer.set_pool_target(max_outstanding + (size_t)workers + 2);

	// what a finished (and finalized) batch adds to the call's statistics
	auto account = [](RenderStats &st, const GlyphBatch &b) {
		st.glyphs += b.glyphs().size();
		for (const BatchGlyph &g : b.glyphs())
			st.bitmaps += g.has_bitmap ? 1 : 0;
		st.pixels += b.pixel_count();
		st.segments += b.total_segments();
		st.pairs += b.pairs();
		st.handed_back += b.handed_back();
		st.h2d_bytes += b.upload_bytes();
	};
	std::vector<std::array<double, 3>> task_density((size_t)workers, std::array<double, 3>{0.0, 0.0, 0.0});
	// bytes of each batch buffer per glyph of the heaviest task (GlyphBatch::capacities order)
	std::vector<std::array<double, GlyphBatch::kBuffers>> task_bytes((size_t)workers);
	for (auto &a : task_bytes)
		a.fill(0.0);
	auto work = [&](int wid) {
		RenderStats &st = per_worker[(size_t)wid];
		std::array<double, 3> &dens = task_density[(size_t)wid]; // heaviest task: segments, curve slots, tile jobs per request
		std::array<double, GlyphBatch::kBuffers> &bpg = task_bytes[(size_t)wid];
		auto &ev = events[(size_t)wid];
		auto mark = [&](char what) {
			if (trace)
				ev.emplace_back(what, now_ns() - t_begin);
		};
		mark('B');
		std::vector<std::pair<std::string, std::vector<uint8_t>>> pending;
		// one finished part: encode its glyph entries; the last part of a block assembles and writes the file
		auto finish_part = [&](const Todo &todo, const GlyphBatch &batch, size_t g0, size_t g1) -> bool {
			BlockState &bs = *todo.bs;
			uint64_t t0 = now_ns();
			std::vector<uint8_t> data;
			const bool whole = bs.parts.size() == 1;
			if (whole)
				data = bs.blk->encode_range(*bs.name, batch, g0, g1);
			else
				bs.parts[todo.part] = encode_glyph_entries(batch, g0, g1);
			if (!whole && bs.remaining.fetch_sub(1, std::memory_order_acq_rel) != 1) {
				st.encode_ns += now_ns() - t0;
				return true;
			}
			if (!whole) {
				mark('a');
				data = assemble_glyphs_pbf(*bs.name, bs.blk->range(), bs.parts);
				mark('A');
			}
			st.encode_ns += now_ns() - t0;
			st.pbf_bytes += data.size();
			st.blocks++;
			pending.emplace_back(*bs.name + "/" + bs.blk->filename(), std::move(data));
			return true;
		};
		// finished files are handed to the writer in groups: one lock acquisition per retired batch instead
		// of one per file (512 files per font pair, most of them tiny, would otherwise queue on the mutex)
		auto flush = [&]() -> bool {
			if (pending.empty())
				return true;
			const uint64_t t0 = now_ns();
			std::string e;
			bool ok = true;
			if (writer.concurrent_files()) { // files in a directory: every worker writes its own, side by side
				for (auto &f : pending)
					if (ok)
						ok = writer.write_file(f.first, std::move(f.second), &e);
			} else {
				std::lock_guard<std::mutex> g(writer_mutex);
				for (auto &f : pending)
					if (ok)
						ok = writer.write_file(f.first, std::move(f.second), &e);
			}
			pending.clear();
			st.write_ns += now_ns() - t0;
			mark('F');
			if (!ok)
				fail(e);
			return ok;
		};
		// ---- the pipeline ------------------------------------------------------------------------------
		// Workers only record outlines and encode results.  Every CUDA call (submit, completion polling) is
		// made by ONE thread — the submitter: 16 threads entering the driver concurrently were measured to
		// stall each other's launches for hundreds of microseconds.  Batches travel through two queues:
		//   worker --submit_q--> submitter --(GPU)--> submitter --done_q--> any worker (encode, write)
		// With a single worker (the reference's --single-thread) the worker pumps the queues itself.
		unsigned n_batches = 0;
		unsigned idle_spins = 0;
		bool more = true;
		for (;;) {
			if (failed.load())
				break;
			// 1. finished batches first: encoding frees the batch and gets files out early
			if (coop_pump)
				worker_pump(false, nullptr);
			Flight *done = nullptr;
			size_t part = 0;
			{
				std::lock_guard<std::mutex> g(qm);
				if (!done_q.empty()) {
					done = done_q.front();
					part = done->next_part++;
					if (done->next_part >= done->parts.size()) // (also a batch without parts: an error path dropped its results)
						done_q.pop_front();
				}
			}
			if (done) {
				idle_spins = 0;
				mark('e');
				bool ok = true;
				const bool empty = done->parts.empty();
				if (!empty) {
					int expected = 0;
					if (done->fin.compare_exchange_strong(expected, 1, std::memory_order_acq_rel)) {
						std::string e;
						const bool fok = done->batch->finalize(renderer, &e);
						if (!fok)
							fail(e);
						account(st, *done->batch);
						done->fin.store(fok ? 2 : 3, std::memory_order_release);
					} else {
						while (done->fin.load(std::memory_order_acquire) < 2)
							cpu_pause();
					}
					ok = done->fin.load(std::memory_order_acquire) == 2;
					if (ok) {
						const Part &p = done->parts[part];
						ok = finish_part(*p.todo, *done->batch, p.g0, p.g1);
					}
				}
				ok = flush() && ok;
				if (empty || done->parts_left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
					// the batch's last part: give it back
					renderer.release_batch(std::move(done->batch));
					delete done;
					{
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
					}
					qcv.notify_all();
				}
				if (!ok)
					break;
				continue;
			}
			// 2. record the next batch
			if (more) {
				{
					std::unique_lock<std::mutex> lk(qm);
					if (outstanding >= max_outstanding) { // back-pressure: wait for a completion
						if (inline_pump || coop_pump)
							lk.unlock(), worker_pump(true, &idle_spins);
						else {
							sleepers.fetch_add(1, std::memory_order_acq_rel);
							qcv.wait_for(lk, std::chrono::milliseconds(kWaitMs),
							             [&] { return !done_q.empty() || outstanding < max_outstanding || failed.load(); });
							sleepers.fetch_sub(1, std::memory_order_acq_rel);
						}
						continue;
					}
					++outstanding; // reserve the place now: several workers pass this check at the same time
				}
				std::unique_ptr<Flight> cur(new Flight());
				cur->batch = renderer.acquire_batch(true);
				cur->batch->set_cost_context(call_units);
				mark('o');
				uint64_t t0 = now_ns();
				// Batch size.  One thread enqueues every batch (about 10 us each), so batches are as large as the
				// pipeline allows: a few workers open with a single part so that the GPU starts early, the rest
				// record `target` glyphs at a time, and the size tapers with the work that is left — the last
				// batches of all workers are submitted together and their latency is the tail of the call.
				const size_t left = total_glyphs - std::min(total_glyphs, glyphs_taken.load(std::memory_order_relaxed));
				// (half of each worker's share of what is left: sizes fall geometrically towards the end)
				size_t want = std::max<size_t>(kPartGlyphs, std::min(target, left / ((size_t)workers * 2)));
				if (n_batches == 0) // staggered openings: the workers do not all submit at the same moments
					want = wid < kEarlyWorkers ? kPartGlyphs : std::min(target, kPartGlyphs * (size_t)(1 + wid % 4));
				// Device-side decoding: filling a batch costs microseconds, submitting one costs the CUDA thread ~15 us and
				// a kernel pair under a few hundred glyphs runs as long as its heaviest tile: equal, large batches.
				if (glyf_mode) // (each worker opens with a quarter-size batch: the GPU has work after tens of microseconds)
					want = n_batches == 0 ? std::max<size_t>(kOpenGlyphs, target / 4) : target;
				++n_batches;
				while (cur->batch->glyphs().size() < want) {
					const size_t ti = next.fetch_add(1);
					if (ti >= tasks.size()) {
						more = false;
						break;
					}
					const Todo &todo = tasks[ti];
					const size_t g0 = cur->batch->glyphs().size();
					glyphs_taken.fetch_add(todo.glyphs, std::memory_order_relaxed);
					const GlyphBatch &bb = *cur->batch;
					const uint64_t r0 = bb.job_count(), s0 = (uint64_t)bb.segment_count() + bb.generated_segment_slots(),
					               c0 = bb.curve_slots(), k0 = bb.tile_cap();
					const uint64_t u0[GlyphBatch::kBuffers] = {(uint64_t)bb.job_count() * sizeof(b200sdf_outline_job),
					                                           (uint64_t)bb.segment_count() * sizeof(b200sdf_segment),
					                                           (uint64_t)bb.curve_count() * sizeof(b200sdf_curve), bb.bitmap_bytes(), 0,
					                                           (uint64_t)bb.job_count() * sizeof(b200sdf_glyph_req),
					                                           (uint64_t)bb.part_count() * sizeof(b200sdf_glyph_part),
					                                           (uint64_t)bb.job_count() * sizeof(b200sdf_glyph_frame)};
					todo.bs->blk->append_to_batch(*cur->batch, todo.slot0, todo.slot1);
					if (bb.glyphs().size() > g0) {
						const uint64_t u1[GlyphBatch::kBuffers] = {(uint64_t)bb.job_count() * sizeof(b200sdf_outline_job),
						                                           (uint64_t)bb.segment_count() * sizeof(b200sdf_segment),
						                                           (uint64_t)bb.curve_count() * sizeof(b200sdf_curve), bb.bitmap_bytes(), 0,
						                                           (uint64_t)bb.job_count() * sizeof(b200sdf_glyph_req),
						                                           (uint64_t)bb.part_count() * sizeof(b200sdf_glyph_part),
						                                           (uint64_t)bb.job_count() * sizeof(b200sdf_glyph_frame)};
						const double n = (double)(bb.glyphs().size() - g0);
						for (int k = 0; k < GlyphBatch::kBuffers; ++k)
							bpg[(size_t)k] = std::max(bpg[(size_t)k], (double)(u1[k] - u0[k]) / n);
					}
					if (glyf_mode && bb.job_count() > r0) { // what this task needs on the device per request: the same in every call
						const double n = (double)(bb.job_count() - r0);
						dens[0] = std::max(dens[0], (double)((uint64_t)bb.segment_count() + bb.generated_segment_slots() - s0) / n);
						dens[1] = std::max(dens[1], (double)(bb.curve_slots() - c0) / n);
						dens[2] = std::max(dens[2], (double)(bb.tile_cap() - k0) / n);
					}
					cur->parts.push_back(Part{&todo, g0, cur->batch->glyphs().size()});
				}
				st.outline_ns += now_ns() - t0;
				if (cur->batch->failed()) { // a glyph could not be recorded (allocation failure): the output would be incomplete
					fail(cur->batch->failure());
					renderer.release_batch(std::move(cur->batch));
					{
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
					}
					qcv.notify_all();
					break;
				}
				if (cur->parts.empty()) {
					renderer.release_batch(std::move(cur->batch));
					{
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
					}
					qcv.notify_all(); // a worker may be asleep waiting for exactly this reservation to go away
					continue;
				}
				if (cur->batch->job_count() == 0) {
					// nothing to rasterise (empty blocks, or only bitmap-less glyphs): no GPU round trip
					bool ok = true;
					account(st, *cur->batch);
					for (const Part &p : cur->parts)
						if (ok)
							ok = finish_part(*p.todo, *cur->batch, p.g0, p.g1);
					ok = ok && flush();
					renderer.release_batch(std::move(cur->batch));
					{
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
					}
					qcv.notify_all();
					if (!ok)
						break;
					continue;
				}
				st.submits++;
				{
					// everything that is not a CUDA call happens here, in the worker: bitmap buffer, tile planning
					std::string e;
					t0 = now_ns();
					// (a worker's last batch — the task queue ran dry while filling it — is planned for latency: the
					// call ends when it comes back)
					const bool prepared = renderer.prepare_batch(*cur->batch, &e, !more && kLatencyTail);
					st.submit_ns += now_ns() - t0;
					if (!prepared) {
						fail(e);
						renderer.release_batch(std::move(cur->batch));
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
						break;
					}
				}
				mark('s');
				{
					std::lock_guard<std::mutex> g(qm);
					cur->parts_left.store(cur->parts.size(), std::memory_order_relaxed); // (published by the queue lock)
					submit_q.push_back(cur.release());
				}
				if (inline_pump || coop_pump)
					worker_pump(false, nullptr);
				continue;
			}
			// 3. nothing left to record: help until every batch has come back
			{
				std::unique_lock<std::mutex> lk(qm);
				if (outstanding == 0)
					break;
				if (!done_q.empty())
					continue;
				const uint64_t t0 = now_ns();
				if (inline_pump || coop_pump) {
					lk.unlock(), worker_pump(true, &idle_spins);
				} else {
					// spin briefly on the hand-back counter before sleeping: at the end of a call the next batch is
					// usually tens of microseconds away, less than a sleep / wake-up round trip
					const uint64_t seen = done_seq.load(std::memory_order_acquire);
					lk.unlock();
					bool changed = false;
					for (int spin = 0; spin < kSpinBudget && !changed; ++spin) {
						cpu_pause();
						changed = done_seq.load(std::memory_order_acquire) != seen || failed.load(std::memory_order_relaxed);
					}
					lk.lock();
					if (!changed && done_q.empty() && outstanding != 0 && !failed.load()) {
						sleepers.fetch_add(1, std::memory_order_acq_rel);
						qcv.wait_for(lk, std::chrono::milliseconds(kWaitMs), [&] { return !done_q.empty() || outstanding == 0 || failed.load(); });
						sleepers.fetch_sub(1, std::memory_order_acq_rel);
					}
				}
				st.wait_ns += now_ns() - t0;
			}
		}
		flush();
		if (failed.load())
			qcv.notify_all(); // sleepers look at `failed`
		mark('E');
	};

	// The submitter: submit whatever the workers queued, poll what is in flight, hand back what finished.
	// (With one worker it is called inline: `block` = nothing else to do, wait for the oldest batch.)
	std::deque<Flight *> inflight; // touched by the pumping thread only
	std::vector<Flight *> finished_now;
	pump = [&](bool block) {
		bool progressed = false;
		for (;;) {
			Flight *f = nullptr;
			{
				// Everything the workers queued since the last look goes into ONE submission (glyph-level batches): a
				// kernel pair over a few hundred glyphs runs as long as its slowest glyph and heaviest tile, so 20 small
				// submissions cost the GPU 0.7 ms where the same glyphs in a few large ones cost 0.55
				std::lock_guard<std::mutex> g(qm);
				static const size_t group_max = [] {
					const char *e = std::getenv("VGB_GROUP_MAX");
					const long v = e ? std::atol(e) : 0;
					return v >= 1 && v <= (long)Renderer::kMaxGroup ? (size_t)v : Renderer::kMaxGroup;
				}();
				const size_t take = glyf_mode ? group_max : 1;
				size_t group_glyphs = 0; // large batches fill the GPU on their own: a submission stops growing at 4096 glyphs
				while (!submit_q.empty() && (!f || (f->mates.size() + 1 < take && group_glyphs < 4096))) {
					Flight *q = submit_q.front();
					group_glyphs += q->batch->job_count();
					submit_q.pop_front();
					if (!f)
						f = q;
					else
						f->mates.push_back(q);
				}
			}
			if (f) {
				std::string e;
				const uint64_t t0 = now_ns();
				if (trace)
					events[(size_t)workers].emplace_back('s', t0 - t_begin);
				GlyphBatch *group[Renderer::kMaxGroup];
				group[0] = f->batch.get();
				for (size_t k = 0; k < f->mates.size(); ++k)
					group[k + 1] = f->mates[k]->batch.get();
				const bool ok = renderer.submit_batches(group, f->mates.size() + 1, &f->ticket, &e);
				submit_ns.fetch_add(now_ns() - t0, std::memory_order_relaxed);
				if (trace)
					events[(size_t)workers].emplace_back('S', now_ns() - t_begin);
				if (ok) {
					inflight.push_back(f);
				} else {
					fail(e);
					f->parts.clear(); // results are dropped
					{
						std::lock_guard<std::mutex> g(qm);
						done_q.push_back(f);
						for (Flight *q : f->mates) {
							q->parts.clear();
							done_q.push_back(q);
						}
					}
					done_seq.fetch_add(1 + f->mates.size(), std::memory_order_release);
					f->mates.clear();
					qcv.notify_all();
				}
			}
			// a polling sweep after every submission: slots come back only through here.  Batches finish
			// roughly in submission order, so only the oldest few are asked (a query costs about a microsecond).
			const size_t sweep = f ? 2 : 8;
			finished_now.clear();
			for (size_t i = 0; i < inflight.size() && i < sweep;) {
				Flight *q = inflight[i];
				bool finished = false;
				std::string e;
				bool ok;
				if (block && !f && i == 0 && !progressed) {
					ok = renderer.wait_batch(q->ticket, &e);
					finished = true;
				} else {
					ok = renderer.poll_batch(q->ticket, &finished, &e);
				}
				if (!ok) {
					fail(e);
					q->parts.clear();
					for (Flight *m : q->mates)
						m->parts.clear();
					finished = true;
				}
				if (!finished) {
					++i;
					continue;
				}
				progressed = true;
				if (trace)
					events[(size_t)workers].emplace_back('d', now_ns() - t_begin);
				inflight.erase(inflight.begin() + (long)i);
				finished_now.push_back(q);
				for (Flight *m : q->mates)
					finished_now.push_back(m);
				q->mates.clear();
			}
			if (!finished_now.empty()) {
				// hand everything that finished in this sweep over at once: one lock, one wake-up (a futex wake per
				// batch cost the submitter ~10 us each and completions queued up behind it at the end of a call)
				{
					std::lock_guard<std::mutex> g(qm);
					for (Flight *q : finished_now)
						done_q.push_back(q);
				}
				done_seq.fetch_add((uint64_t)finished_now.size(), std::memory_order_release);
				if (sleepers.load(std::memory_order_acquire) > 0)
					qcv.notify_all();
			}
			if (!f)
				break;
		}
	};
	auto submitter = [&]() {
		// watchdog: a pipeline that makes no progress for this long is reported as an error instead of hanging
		static const uint64_t stall_ns = [] {
			const char *e = std::getenv("VGB_STALL_SECONDS");
			const double v = e ? std::atof(e) : 0.0;
			return (uint64_t)((v > 0.0 ? v : 60.0) * 1e9);
		}();
		uint64_t last_progress = now_ns();
		size_t last_state = ~(size_t)0;
		for (;;) {
			pump(false);
			size_t state;
			{
				std::lock_guard<std::mutex> g(qm);
				if (workers_done == workers && submit_q.empty() && inflight.empty())
					break;
				state = next.load() * 131 + outstanding * 17 + done_q.size() * 7 + inflight.size() + (size_t)workers_done * 1000003;
			}
			const uint64_t t = now_ns();
			if (state != last_state) {
				last_state = state;
				last_progress = t;
			} else if (t - last_progress > stall_ns && !failed.load()) {
				std::lock_guard<std::mutex> g(qm);
				fail("render_glyphs pipeline stalled: tasks " + std::to_string(next.load()) + "/" + std::to_string(tasks.size()) +
				     ", outstanding " + std::to_string(outstanding) + ", in flight " + std::to_string(inflight.size()) +
				     ", to submit " + std::to_string(submit_q.size()) + ", done " + std::to_string(done_q.size()) +
				     ", workers done " + std::to_string(workers_done) + "/" + std::to_string(workers));
				qcv.notify_all();
				last_progress = t;
			}
			for (int k = 0; k < 16; ++k)
				cpu_pause();
		}
	};

	if (inline_pump) {
		work(0);
	} else if (coop_pump) {
		// the calling thread is worker 0
		WorkerPool::instance().run(workers - 1, [&](int id) { work(id + 1); }, [&] { work(0); });
	} else {
		WorkerPool::instance().run(
		    workers,
		    [&](int id) {
			    work(id);
			    std::lock_guard<std::mutex> g(qm);
			    ++workers_done;
		    },
		    submitter); // the calling thread is the submitter
	}
	// error paths leave batches behind: nothing is in flight any more (the submitter drained), free them
	if (inline_pump || coop_pump)
		while (!inflight.empty()) {
			renderer.wait_batch(inflight.front()->ticket, nullptr);
			done_q.push_back(inflight.front());
			for (Flight *m : inflight.front()->mates)
				done_q.push_back(m);
			inflight.front()->mates.clear();
			inflight.pop_front();
		}
	for (std::deque<Flight *> *q : {&submit_q, &done_q})
		for (Flight *f : *q) {
			if (f->batch)
				renderer.release_batch(std::move(f->batch));
			delete f;
		}
	if (glyf_mode) {
		// what the largest merged submission of a later call can need, from figures that do not depend on timing: the
		// heaviest task's needs per request, and 4096 glyphs (where a submission stops growing) + one full batch
		std::array<double, 3> d{0.0, 0.0, 0.0};
		for (const auto &w : task_density)
			for (int k = 0; k < 3; ++k)
				d[(size_t)k] = std::max(d[(size_t)k], w[(size_t)k]);
		renderer.note_glyf_density(d[0], d[1], d[2]);
		renderer.set_glyf_group_bound(std::min<uint64_t>(total_glyphs, 4096 + target + kPartGlyphs));
		// the pooled batches: no batch holds more than target + one task's glyphs, none is heavier per glyph than the
		// heaviest task (+ 1/8, rounded to the 256 KiB steps pinned buffers grow in)
		size_t caps[GlyphBatch::kBuffers];
		const double max_glyphs = (double)std::min<uint64_t>(total_glyphs, target + kPartGlyphs);
		for (int k = 0; k < GlyphBatch::kBuffers; ++k) {
			double b = 0.0;
			for (const auto &w : task_bytes)
				b = std::max(b, w[(size_t)k]);
			const double bytes = b * max_glyphs * 1.125 + 64.0;
			caps[k] = b > 0.0 ? (((size_t)bytes + ((size_t)256 << 10) - 1) & ~(((size_t)256 << 10) - 1)) : 0;
		}
		renderer.raise_batch_marks(caps);
	}
	renderer.top_up_pool();
	if (trace) {
		std::fprintf(stderr, "[vgb trace] setup %.1f us, %d workers, %zu tasks, target %zu glyphs, total %.1f us\n",
		             (double)(t_setup - t_begin) * 1e-3, workers, tasks.size(), target, (double)(now_ns() - t_begin) * 1e-3);
		for (int w = 0; w <= workers; ++w) {
			std::fprintf(stderr, "[vgb trace] %c%02d", w == workers ? 'S' : 'w', w);
			for (const auto &e : events[(size_t)w])
				std::fprintf(stderr, " %c%.0f", e.first, (double)e.second * 1e-3);
			std::fprintf(stderr, "\n");
		}
	}
	if (stats) {
		*stats = RenderStats();
		for (const RenderStats &s : per_worker) {
			stats->glyphs += s.glyphs;
			stats->bitmaps += s.bitmaps;
			stats->pixels += s.pixels;
			stats->segments += s.segments;
			stats->pairs += s.pairs;
			stats->pbf_bytes += s.pbf_bytes;
			stats->blocks += s.blocks;
			stats->outline_ns += s.outline_ns;
			stats->submit_ns += s.submit_ns;
			stats->wait_ns += s.wait_ns;
			stats->encode_ns += s.encode_ns;
			stats->write_ns += s.write_ns;
			stats->submits += s.submits;
			stats->handed_back += s.handed_back;
			stats->h2d_bytes += s.h2d_bytes;
		}
		stats->cost_total = cost_total;
		stats->cost_shard = cost_mine;
		stats->submit_ns += submit_ns.load();
		stats->workers = (uint64_t)workers;
		stats->wall_ns = now_ns() - t_begin;
	}
	return !failed.load();
}

} // namespace vgb
