// font.cc — see font.h.
#include "font.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
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
	e->codepoints = face->codepoints();
	e->family = face->name(1);
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

bool GlyphBlock::fill_batch(GlyphBatch &batch) const
{
	batch.clear();
	for (uint32_t i = 0; i < GLYPH_BLOCK_SIZE; ++i) {
		const FontFileEntry *f = fonts_[i];
		if (f)
			batch.add_glyph(*f->face, start_index_ + i); // false = None = skipped (glyph_block.rs:74-76)
	}
	return true;
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
	for (const auto &file : files_)
		for (uint32_t cp : file->codepoints) {
			if (cp > 0xFFFF)
				continue;
			blocks[cp / GLYPH_BLOCK_SIZE].set_glyph_font((uint8_t)(cp % GLYPH_BLOCK_SIZE), file.get());
		}
	return blocks;
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

bool Writer::write_directory(const std::string &dirname, std::string *err)
{
	if (!to_disk_) {
		Entry e;
		e.name = dirname;
		e.is_dir = true;
		entries_.push_back(std::move(e));
		return true;
	}
	return mkdirs(folder_ + "/" + dirname, err);
}

bool Writer::finish(std::string *)
{
	finished_ = true;
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
	// The reference derives "<family> [<width>] <weight> [<style>]" with the heuristics of
	// font/parse_font_name.rs (out of this path's scope).  Subset: family name + "Regular" unless
	// the family already ends in a weight word — enough for the fixtures; use
	// add_font_with_name() (what `recurse` does with fonts.json) for exact control.
	std::string name = e->family.empty() ? path.substr(path.find_last_of('/') + 1) : e->family;
	static const char *weights[] = {"Thin", "ExtraLight", "Light", "Regular", "Medium", "SemiBold", "Bold", "ExtraBold", "Black"};
	bool has_weight = false;
	for (const char *w : weights) {
		const size_t n = std::strlen(w);
		if (name.size() > n && name.compare(name.size() - n, n, w) == 0 && name[name.size() - n - 1] == ' ')
			has_weight = true;
	}
	if (!has_weight)
		name += " Regular";
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

bool FontManager::render_glyphs(Writer &writer, const Renderer &renderer, std::string *err, RenderStats *stats,
                                uint32_t shard, uint32_t n_shards, int threads) const
{
	struct Todo {
		const std::string *name;
		GlyphBlock block;
	};
	if (n_shards == 0)
		n_shards = 1;
	std::vector<Todo> tasks;
	uint32_t index = 0;
	for (const auto &kv : fonts_) {
		if (!writer.write_directory(kv.first + "/", err))
			return false;
		for (GlyphBlock &b : kv.second.get_blocks()) {
			if (index++ % n_shards == shard)
				tasks.push_back(Todo{&kv.first, std::move(b)});
		}
	}

	int workers = 1;
	if (parallel_) {
		workers = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
		workers = std::max(1, std::min(workers, 32));
		// each worker keeps two batches in flight; never more than the renderer has slots for
		if (renderer.mode() == Renderer::Mode::Cuda)
			workers = std::max(1, std::min(workers, (int)renderer.slots() / 2));
	}
	std::atomic<size_t> next{0};
	std::atomic<bool> failed{false};
	std::mutex writer_mutex, err_mutex;
	std::vector<RenderStats> per_worker((size_t)workers);

	auto fail = [&](const std::string &msg) {
		std::lock_guard<std::mutex> g(err_mutex);
		if (!failed.exchange(true) && err)
			*err = msg;
	};

	auto work = [&](int wid) {
		// Two batches per worker: while batch A is on the GPU, batch B is being flattened.
		// (batches come from the renderer's pool: their pinned buffers survive across calls)
		struct Lease {
			const Renderer &r;
			std::unique_ptr<GlyphBatch> b[2];
			explicit Lease(const Renderer &rr) : r(rr) { b[0] = r.acquire_batch(), b[1] = r.acquire_batch(); }
			~Lease() { r.release_batch(std::move(b[0])), r.release_batch(std::move(b[1])); }
		} batches(renderer);
		struct InFlight {
			const Todo *todo = nullptr;
			GlyphBatch *batch = nullptr;
			uint64_t ticket = 0;
		} pending;
		RenderStats &st = per_worker[(size_t)wid];
		auto retire = [&](InFlight &p) -> bool {
			std::string e;
			if (!renderer.wait_batch(p.ticket, &e)) {
				fail(e);
				return false;
			}
			const std::vector<uint8_t> data = p.todo->block.encode_batch(*p.todo->name, *p.batch);
			st.pbf_bytes += data.size();
			st.blocks++;
			std::lock_guard<std::mutex> g(writer_mutex);
			if (!writer.write_file(*p.todo->name + "/" + p.todo->block.filename(), data.data(), data.size(), &e)) {
				fail(e);
				return false;
			}
			return true;
		};
		int k = 0;
		while (!failed.load()) {
			const size_t ti = next.fetch_add(1);
			if (ti >= tasks.size())
				break;
			const Todo &todo = tasks[ti];
			GlyphBatch *cur = batches.b[k].get();
			k ^= 1;
			todo.block.fill_batch(*cur);
			st.glyphs += cur->glyphs().size();
			st.bitmaps += cur->job_count();
			st.pixels += cur->bitmap_bytes();
			st.segments += cur->total_segments();
			st.pairs += cur->pairs();
			InFlight now;
			now.todo = &todo;
			now.batch = cur;
			std::string e;
			if (!renderer.submit_batch(*cur, &now.ticket, &e)) {
				fail(e);
				break;
			}
			if (pending.todo && !retire(pending)) {
				pending.todo = nullptr;
				renderer.wait_batch(now.ticket, nullptr);
				return;
			}
			pending = now;
		}
		if (pending.todo)
			retire(pending);
	};

	if (workers == 1) {
		work(0);
	} else {
		std::vector<std::thread> pool;
		for (int w = 0; w < workers; ++w)
			pool.emplace_back(work, w);
		for (auto &t : pool)
			t.join();
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
		}
	}
	return !failed.load();
}

} // namespace vgb
