// font.cc — see font.h.
#include "font.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <fstream>
#include <functional>
#include <sys/stat.h>
#include <thread>

namespace vgb {

// ---- FontFileEntry -----------------------------------------------------------------------------------
std::unique_ptr<FontFileEntry> FontFileEntry::from_bytes(std::vector<uint8_t> data, std::string *err)
{
	std::unique_ptr<Face> face = Face::parse(std::move(data));
	if (!face) {
		if (err)
			*err = "Could not parse font data"; // file_entry.rs:48
		return nullptr;
	}
	std::unique_ptr<FontFileEntry> e(new FontFileEntry());
	e->metadata = FontMetadata::from_face(*face);
	e->codepoints = e->metadata.codepoints;
	e->family = e->metadata.name;
	e->face = std::move(face);
	return e;
}

std::unique_ptr<FontFileEntry> FontFileEntry::from_path(const std::string &path, std::string *err)
{
	std::ifstream f(path, std::ios::binary);
	if (!f) {
		if (err)
			*err = "reading font file \"" + path + "\": " + std::strerror(errno); // wrapper.rs:33
		return nullptr;
	}
	std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	return from_bytes(std::move(data), err);
}

// ---- GlyphBlock --------------------------------------------------------------------------------------
std::string GlyphBlock::range() const
{
	return std::to_string(start_index_) + "-" + std::to_string(start_index_ + GLYPH_BLOCK_SIZE - 1);
}
std::string GlyphBlock::filename() const { return range() + ".pbf"; }

void GlyphBlock::append_to_batch(GlyphBatch &batch, uint32_t slot0, uint32_t slot1) const
{
	for (uint32_t i = slot0; i < slot1 && i < GLYPH_BLOCK_SIZE; ++i) {
		const FontFileEntry *f = font_of((uint8_t)i);
		if (f)
			batch.add_glyph(*f->face, start_index_ + i); // false = None = skipped (glyph_block.rs:74-76)
	}
}

bool GlyphBlock::fill_batch(GlyphBatch &batch) const
{
	batch.clear();
	append_to_batch(batch);
	return true;
}

std::vector<uint8_t> GlyphBlock::encode_range(const std::string &font_name, const GlyphBatch &batch, size_t g0, size_t g1) const
{
	return encode_batch_range(font_name, range(), batch, g0, g1);
}

std::vector<uint8_t> GlyphBlock::encode_batch(const std::string &font_name, const GlyphBatch &batch) const
{
	PbfGlyphs glyphs(font_name, range());
	for (size_t i = 0; i < batch.glyphs().size(); ++i)
		glyphs.push(batch.take_glyph(i));
	return glyphs.into_vec();
}

bool GlyphBlock::render(const std::string &font_name, const Renderer &renderer, std::vector<uint8_t> &out,
                        std::string *err) const
{
	std::unique_ptr<GlyphBatch> batch = renderer.acquire_batch();
	fill_batch(*batch);
	const bool ok = renderer.render_batch(*batch, err);
	if (ok)
		out = encode_batch(font_name, *batch);
	renderer.release_batch(std::move(batch));
	return ok;
}

// ---- FontWrapper -------------------------------------------------------------------------------------
bool FontWrapper::add_paths(const std::vector<std::string> &sources, std::string *err)
{
	for (const std::string &p : sources) {
		std::unique_ptr<FontFileEntry> e = FontFileEntry::from_path(p, err);
		if (!e)
			return false;
		files_.push_back(std::move(e));
	}
	return true;
}

std::vector<GlyphBlock> FontWrapper::get_blocks() const
{
	constexpr uint32_t BMP_BLOCK_COUNT = 0x10000 / GLYPH_BLOCK_SIZE;
	std::vector<GlyphBlock> blocks;
	blocks.reserve(BMP_BLOCK_COUNT);
	for (uint32_t i = 0; i < BMP_BLOCK_COUNT; ++i)
		blocks.emplace_back(i * GLYPH_BLOCK_SIZE);
	std::vector<GlyphBlock *> ptrs(BMP_BLOCK_COUNT);
	for (uint32_t i = 0; i < BMP_BLOCK_COUNT; ++i)
		ptrs[i] = &blocks[i];
	assign_blocks(ptrs.data());
	return blocks;
}

void FontWrapper::assign_blocks(GlyphBlock *const *blocks) const
{
	for (const auto &file : files_)
		for (uint32_t cp : file->codepoints) {
			if (cp > 0xFFFF)
				continue;
			blocks[cp / GLYPH_BLOCK_SIZE]->set_glyph_font((uint8_t)(cp % GLYPH_BLOCK_SIZE), file.get());
		}
}

// ---- Writer ------------------------------------------------------------------------------------------
Writer Writer::new_file(const std::string &folder)
{
	Writer w;
	w.to_disk_ = true;
	w.folder_ = folder;
	return w;
}
Writer Writer::new_memory() { return Writer(); }

Writer Writer::new_tar_memory()
{
	Writer w;
	w.to_tar_ = true;
	return w;
}

Writer Writer::new_tar(const std::string &path)
{
	Writer w;
	w.to_tar_ = true;
	w.folder_ = path;
	std::FILE *f = std::fopen(path.c_str(), "wb");
	if (f)
		w.tar_file_ = std::shared_ptr<std::FILE>(f, [](std::FILE *p) { std::fclose(p); });
	return w;
}

bool Writer::tar_put(const uint8_t *p, size_t n, std::string *err)
{
	if (!folder_.empty()) {
		if (!tar_file_ || (n && std::fwrite(p, 1, n, tar_file_.get()) != n)) {
			if (err)
				*err = "writing tar \"" + folder_ + "\" failed";
			return false;
		}
		return true;
	}
	tar_.insert(tar_.end(), p, p + n);
	return true;
}

// writer/tar.rs:49-99 — one 512-byte ustar header; octal fields are zero-filled and end with a space
bool Writer::tar_header(const std::string &path, uint64_t size, uint64_t mode, char typeflag, std::string *err)
{
	uint8_t h[512];
	std::memset(h, 0, sizeof(h));
	if (path.size() > 100) { // tar.rs:160-172
		if (err)
			*err = "tar header field overflow: \"" + path + "\" is " + std::to_string(path.size()) + " bytes, max 100";
		return false;
	}
	std::memcpy(h, path.data(), path.size());
	auto octal = [&](size_t off, size_t len, uint64_t v) { // tar.rs:147-156
		h[off + len - 1] = ' ';
		for (size_t i = len - 1; i-- > 0;) {
			h[off + i] = (uint8_t)('0' + (v & 7));
			v >>= 3;
		}
	};
	octal(100, 8, mode);
	octal(108, 8, 0);
	octal(116, 8, 0);
	octal(124, 12, size);
	octal(136, 12, (uint64_t)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count());
	h[156] = (uint8_t)typeflag;
	std::memcpy(h + 257, "ustar\0", 6);
	std::memcpy(h + 263, "00", 2);
	std::memset(h + 148, ' ', 8);
	uint32_t sum = 0;
	for (uint8_t b : h)
		sum += b;
	octal(148, 8, sum);
	return tar_put(h, sizeof(h), err);
}

static bool mkdirs(const std::string &path, std::string *err)
{
	std::string cur;
	for (size_t i = 0; i <= path.size(); ++i) {
		if (i == path.size() || path[i] == '/') {
			if (!cur.empty() && ::mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) {
				if (err)
					*err = "mkdir " + cur + ": " + std::strerror(errno);
				return false;
			}
		}
		if (i < path.size())
			cur.push_back(path[i]);
	}
	return true;
}

bool Writer::write_file(const std::string &filename, const uint8_t *bytes, size_t len, std::string *err)
{
	bytes_written_ += len;
	if (to_tar_) { // writer/tar.rs:101-120
		static const uint8_t zeros[512] = {0};
		if (!tar_header(filename, len, 0644, '0', err) || !tar_put(bytes, len, err))
			return false;
		const size_t rem = len % 512;
		return rem == 0 || tar_put(zeros, 512 - rem, err);
	}
	if (!to_disk_) {
		Entry e;
		e.name = filename;
		e.bytes.assign(bytes, bytes + len);
		entries_.push_back(std::move(e));
		return true;
	}
	const std::string path = folder_ + "/" + filename;
	std::FILE *f = std::fopen(path.c_str(), "wb");
	if (!f) {
		if (err)
			*err = "open " + path + ": " + std::strerror(errno);
		return false;
	}
	const bool ok = len == 0 || std::fwrite(bytes, 1, len, f) == len;
	std::fclose(f);
	if (!ok && err)
		*err = "write " + path + " failed";
	return ok;
}

bool Writer::write_file(const std::string &filename, std::vector<uint8_t> &&bytes, std::string *err)
{
	if (to_disk_ || to_tar_)
		return write_file(filename, bytes.data(), bytes.size(), err);
	bytes_written_ += bytes.size();
	Entry e;
	e.name = filename;
	e.bytes = std::move(bytes);
	entries_.push_back(std::move(e));
	return true;
}

bool Writer::write_directory(const std::string &dirname, std::string *err)
{
	if (to_tar_) { // writer/tar.rs:122-126
		if (dirname.empty() || dirname.back() != '/') {
			if (err)
				*err = "dirname must end with a slash";
			return false;
		}
		return tar_header(dirname, 0, 0755, '5', err);
	}
	if (!to_disk_) {
		Entry e;
		e.name = dirname;
		e.is_dir = true;
		entries_.push_back(std::move(e));
		return true;
	}
	return mkdirs(folder_ + "/" + dirname, err);
}

bool Writer::finish(std::string *err)
{
	if (finished_) // writer/mod.rs:67-73
		return true;
	finished_ = true;
	if (to_tar_) { // writer/tar.rs:133-137: two zero blocks
		static const uint8_t zeros[1024] = {0};
		if (!tar_put(zeros, sizeof(zeros), err))
			return false;
		if (tar_file_)
			std::fflush(tar_file_.get());
	}
	return true;
}

// ---- FontManager -------------------------------------------------------------------------------------
std::string FontManager::name_to_id(const std::string &name)
{
	// lowercase; every run of [-_\s] becomes one separator; trim; separators -> '_'
	std::string out;
	bool pending = false;
	for (unsigned char ch : name) {
		if (ch >= 'A' && ch <= 'Z')
			ch = (unsigned char)(ch - 'A' + 'a');
		const bool sep = ch == '-' || ch == '_' || ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\f' || ch == '\v';
		if (sep) {
			pending = true;
			continue;
		}
		if (pending && !out.empty())
			out.push_back('_');
		pending = false;
		out.push_back((char)ch);
	}
	return out;
}

bool FontManager::add_font_with_name(const std::string &name, const std::vector<std::string> &sources, std::string *err)
{
	return fonts_[name_to_id(name)].add_paths(sources, err);
}

bool FontManager::add_font_bytes_with_name(const std::string &name, std::vector<uint8_t> data, std::string *err)
{
	std::unique_ptr<FontFileEntry> e = FontFileEntry::from_bytes(std::move(data), err);
	if (!e)
		return false;
	fonts_[name_to_id(name)].add_file(std::move(e));
	return true;
}

bool FontManager::add_path(const std::string &path, std::string *err)
{
	std::unique_ptr<FontFileEntry> e = FontFileEntry::from_path(path, err);
	if (!e)
		return false;
	// manager.rs:42: id = name_to_id(metadata.generate_name())
	const std::string name = e->metadata.generate_name();
	fonts_[name_to_id(name)].add_file(std::move(e));
	return true;
}

bool FontManager::add_paths(const std::vector<std::string> &paths, std::string *err)
{
	for (const std::string &p : paths)
		if (!add_path(p, err))
			return false;
	return true;
}

bool FontManager::write_index_json(Writer &writer, std::string *err) const
{
	// serde_json::to_vec_pretty of the sorted id list (index_files.rs:109-113)
	std::string s;
	if (fonts_.empty()) {
		s = "[]";
	} else {
		s = "[\n";
		size_t k = 0;
		for (const auto &kv : fonts_) {
			s += "  \"" + kv.first + "\"";
			s += (++k < fonts_.size()) ? ",\n" : "\n";
		}
		s += "]";
	}
	return writer.write_file("index.json", (const uint8_t *)s.data(), s.size(), err);
}

bool FontManager::write_families_json(Writer &writer, std::string *err) const
{
	std::string e;
	const std::string s = build_font_families_json(fonts_, &e);
	if (s.empty()) {
		if (err)
			*err = e;
		return false;
	}
	return writer.write_file("font_families.json", (const uint8_t *)s.data(), s.size(), err);
}

namespace {
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
	void run(int n, const std::function<void(int)> &fn)
	{
		std::lock_guard<std::mutex> serial(run_mu_);
		{
			std::unique_lock<std::mutex> lk(mu_);
			while ((int)threads_.size() < n) {
				const int id = (int)threads_.size();
				threads_.emplace_back([this, id] { loop(id); });
			}
			fn_ = &fn;
			want_ = n;
			remaining_ = n;
			++generation_;
		}
		cv_.notify_all();
		std::unique_lock<std::mutex> lk(mu_);
		done_cv_.wait(lk, [this] { return remaining_ == 0; });
		fn_ = nullptr;
	}

  private:
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

bool FontManager::render_glyphs(Writer &writer, const Renderer &renderer, std::string *err, RenderStats *stats,
                                uint32_t shard, uint32_t n_shards, int threads) const
{
	// A task is a slot range of one block holding at most kPartGlyphs (64) glyphs: full blocks are split so that
	// no worker is stuck recording 256 outlines while the others (and the GPU) run dry.  The parts of a
	// block are encoded independently (Fontstack.glyphs entries) and the worker that finishes the last
	// one assembles and writes the file.
	static const size_t kPartGlyphs = [] { // tuning knob (default 64; measured 16 / 32 / 64 on C2)
		const char *e = std::getenv("VGB_PART_GLYPHS");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 256 ? v : 64);
	}();
	struct BlockState {
		const std::string *name = nullptr;
		GlyphBlock block;
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
	uint32_t index = 0;
	size_t total_glyphs = 0;
	for (const auto &kv : fonts_) {
		if (!writer.write_directory(kv.first + "/", err))
			return false;
		fonts_blocks.emplace_back(new BlockState[kBlocks]);
		BlockState *bsv = fonts_blocks.back().get();
		GlyphBlock *ptrs[kBlocks];
		for (uint32_t i = 0; i < kBlocks; ++i) {
			bsv[i].name = &kv.first;
			bsv[i].block.reset(i * GLYPH_BLOCK_SIZE);
			ptrs[i] = &bsv[i].block;
		}
		kv.second.assign_blocks(ptrs); // = get_blocks(), written in place
		for (uint32_t i = 0; i < kBlocks; ++i) {
			if (index++ % n_shards != shard)
				continue;
			BlockState *bs = &bsv[i];
			total_glyphs += bs->block.len();
			uint32_t part = 0, slot0 = 0, count = 0;
			if (bs->block.len() > kPartGlyphs) {
				for (uint32_t k = 0; k < GLYPH_BLOCK_SIZE; ++k) {
					if (!bs->block.font_of((uint8_t)k))
						continue;
					if (count == kPartGlyphs) {
						tasks.push_back(Todo{bs, part++, slot0, k, count});
						slot0 = k, count = 0;
					}
					++count;
				}
			} else {
				count = (uint32_t)bs->block.len(); // small and empty blocks: one part
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

	int workers = 1;
	if (parallel_) {
		workers = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
		workers = std::max(1, std::min(workers, 32));
		// each worker keeps two batches in flight; never more than the renderer has slots for
		if (renderer.mode() == Renderer::Mode::Cuda)
			workers = std::max(1, std::min(workers, (int)renderer.slots() / 2));
	}
	// One submission carries whole blocks until it holds about `target` glyphs: small jobs keep one
	// block per submission (parallelism), big jobs amortise the per-submission cost.
	static const size_t kBatchesPerWorker = [] { // tuning knob (default 4)
		const char *e = std::getenv("VGB_BATCHES_PER_WORKER");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 64 ? v : 4);
	}();
	const size_t target = std::min<size_t>(2048, std::max<size_t>(1, total_glyphs / ((size_t)workers * kBatchesPerWorker)));
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
	std::vector<std::vector<std::pair<char, uint64_t>>> events((size_t)workers);
	const uint64_t t_setup = now_ns();

	auto work = [&](int wid) {
		RenderStats &st = per_worker[(size_t)wid];
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
				data = bs.block.encode_range(*bs.name, batch, g0, g1);
			else
				bs.parts[todo.part] = encode_glyph_entries(batch, g0, g1);
			if (!whole && bs.remaining.fetch_sub(1, std::memory_order_acq_rel) != 1) {
				st.encode_ns += now_ns() - t0;
				return true;
			}
			if (!whole) {
				mark('a');
				data = assemble_glyphs_pbf(*bs.name, bs.block.range(), bs.parts);
				mark('A');
			}
			st.encode_ns += now_ns() - t0;
			st.pbf_bytes += data.size();
			st.blocks++;
			pending.emplace_back(*bs.name + "/" + bs.block.filename(), std::move(data));
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
			{
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
		// Two batches per worker: while batch A is on the GPU, batch B is being filled.
		// (batches come from the renderer's pool: their pinned buffers survive across calls)
		struct Part {
			const Todo *todo;
			size_t g0, g1;
		};
		struct Flight {
			std::unique_ptr<GlyphBatch> batch;
			std::vector<Part> parts;
			uint64_t ticket = 0;
			bool active = false;
		} flights[2];
		flights[0].batch = renderer.acquire_batch();
		flights[1].batch = renderer.acquire_batch();
		auto retire = [&](Flight &f) -> bool {
			if (!f.active)
				return true;
			f.active = false;
			std::string e;
			uint64_t t0 = now_ns();
			mark('w');
			const bool waited = renderer.wait_batch(f.ticket, &e);
			mark('W');
			st.wait_ns += now_ns() - t0;
			if (!waited) {
				fail(e);
				return false;
			}
			for (const Part &p : f.parts)
				if (!finish_part(*p.todo, *f.batch, p.g0, p.g1))
					return false;
			return flush();
		};
		int k = 0;
		unsigned n_batches = 0;
		bool more = true;
		while (more && !failed.load()) {
			Flight &cur = flights[k];
			cur.batch->clear();
			cur.parts.clear();
			mark('o');
			uint64_t t0 = now_ns();
			// Batch size: small first batches (the GPU starts early), the steady-state target, then a taper —
			// the last batches of all workers are submitted together and their latency is the tail of the call.
			const size_t left = total_glyphs - std::min(total_glyphs, glyphs_taken.load(std::memory_order_relaxed));
			size_t want = std::min(target, std::max<size_t>(kPartGlyphs, left / ((size_t)workers * 2)));
			if (n_batches < 2)
				want = std::min(want, std::max<size_t>(kPartGlyphs, target >> (2 - n_batches)));
			++n_batches;
			while (cur.batch->glyphs().size() < want) {
				const size_t ti = next.fetch_add(1);
				if (ti >= tasks.size()) {
					more = false;
					break;
				}
				const Todo &todo = tasks[ti];
				const size_t g0 = cur.batch->glyphs().size();
				glyphs_taken.fetch_add(todo.glyphs, std::memory_order_relaxed);
				todo.bs->block.append_to_batch(*cur.batch, todo.slot0, todo.slot1);
				cur.parts.push_back(Part{&todo, g0, cur.batch->glyphs().size()});
			}
			st.outline_ns += now_ns() - t0;
			if (cur.parts.empty())
				break;
			st.glyphs += cur.batch->glyphs().size();
			st.bitmaps += cur.batch->job_count();
			st.pixels += cur.batch->bitmap_bytes();
			st.segments += cur.batch->total_segments();
			st.pairs += cur.batch->pairs();
			if (cur.batch->job_count() == 0) {
				// nothing to rasterise (empty blocks, or only bitmap-less glyphs): no GPU round trip
				cur.ticket = ~0ull;
				cur.active = false;
				for (const Part &p : cur.parts)
					if (!finish_part(*p.todo, *cur.batch, p.g0, p.g1))
						break;
				if (!flush())
					break;
				continue;
			}
			t0 = now_ns();
			std::string e;
			mark('s');
			const bool submitted = renderer.submit_batch(*cur.batch, &cur.ticket, &e);
			mark('S');
			st.submit_ns += now_ns() - t0;
			st.submits++;
			if (!submitted) {
				fail(e);
				break;
			}
			cur.active = true;
			k ^= 1;
			if (!retire(flights[k]))
				break;
		}
		for (Flight &f : flights) {
			if (f.active && failed.load()) {
				renderer.wait_batch(f.ticket, nullptr); // drain; results are dropped
				f.active = false;
			}
			retire(f);
			renderer.release_batch(std::move(f.batch));
		}
		mark('E');
	};

	if (workers == 1) {
		work(0);
	} else {
		WorkerPool::instance().run(workers, work);
	}
	if (trace) {
		std::fprintf(stderr, "[vgb trace] setup %.1f us, %d workers, %zu tasks, target %zu glyphs, total %.1f us\n",
		             (double)(t_setup - t_begin) * 1e-3, workers, tasks.size(), target, (double)(now_ns() - t_begin) * 1e-3);
		for (int w = 0; w < workers; ++w) {
			std::fprintf(stderr, "[vgb trace] w%02d", w);
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
		}
		stats->workers = (uint64_t)workers;
		stats->wall_ns = now_ns() - t_begin;
	}
	return !failed.load();
}

} // namespace vgb
