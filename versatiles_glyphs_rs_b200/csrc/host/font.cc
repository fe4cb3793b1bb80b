// font.cc — see font.h.
#include "font.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <condition_variable>
#include <fstream>
#include <functional>
#include <malloc.h>
#include <pthread.h>
#include <sched.h>
#include <dirent.h>
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
		add_file(std::move(e));
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

const std::vector<GlyphBlock> &FontWrapper::blocks() const
{
	std::lock_guard<std::mutex> g(blocks_mu_);
	if (blocks_.empty())
		blocks_ = get_blocks();
	return blocks_;
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

namespace {
// The subset of JSON a fonts.json needs: an array of objects whose "name" is a string and whose "sources" is an
// array of strings (serde would reject anything else for Vec<FontConfig>, recurse.rs:57-63); unknown keys are skipped.
struct JsonCursor {
	const std::string &s;
	size_t i = 0;
	bool ok = true;
	void ws()
	{
		while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r'))
			++i;
	}
	bool eat(char c)
	{
		ws();
		if (i < s.size() && s[i] == c) {
			++i;
			return true;
		}
		return false;
	}
	static void utf8(std::string &o, uint32_t cp)
	{
		if (cp < 0x80)
			o.push_back((char)cp);
		else if (cp < 0x800) {
			o.push_back((char)(0xC0 | (cp >> 6)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		} else if (cp < 0x10000) {
			o.push_back((char)(0xE0 | (cp >> 12)));
			o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		} else {
			o.push_back((char)(0xF0 | (cp >> 18)));
			o.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
			o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		}
	}
	bool hex4(uint32_t &v)
	{
		if (i + 4 > s.size())
			return false;
		v = 0;
		for (int k = 0; k < 4; ++k) {
			const char c = s[i++];
			v <<= 4;
			if (c >= '0' && c <= '9')
				v |= (uint32_t)(c - '0');
			else if (c >= 'a' && c <= 'f')
				v |= (uint32_t)(c - 'a' + 10);
			else if (c >= 'A' && c <= 'F')
				v |= (uint32_t)(c - 'A' + 10);
			else
				return false;
		}
		return true;
	}
	bool string(std::string &out)
	{
		out.clear();
		if (!eat('"'))
			return ok = false;
		while (i < s.size()) {
			const char c = s[i++];
			if (c == '"')
				return true;
			if (c != '\\') {
				out.push_back(c);
				continue;
			}
			if (i >= s.size())
				break;
			const char e = s[i++];
			switch (e) {
			case '"': out.push_back('"'); break;
			case '\\': out.push_back('\\'); break;
			case '/': out.push_back('/'); break;
			case 'b': out.push_back('\b'); break;
			case 'f': out.push_back('\f'); break;
			case 'n': out.push_back('\n'); break;
			case 'r': out.push_back('\r'); break;
			case 't': out.push_back('\t'); break;
			case 'u': {
				uint32_t cp;
				if (!hex4(cp))
					return ok = false;
				if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= s.size() && s[i] == '\\' && s[i + 1] == 'u') {
					i += 2;
					uint32_t lo;
					if (!hex4(lo) || lo < 0xDC00 || lo > 0xDFFF)
						return ok = false;
					cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
				}
				utf8(out, cp);
				break;
			}
			default: return ok = false;
			}
		}
		return ok = false;
	}
	// skips any value (for keys FontConfig does not have)
	bool skip()
	{
		ws();
		if (i >= s.size())
			return ok = false;
		const char c = s[i];
		if (c == '"') {
			std::string t;
			return string(t);
		}
		if (c == '{' || c == '[') {
			const char close = c == '{' ? '}' : ']';
			++i;
			if (eat(close))
				return true;
			for (;;) {
				if (c == '{') {
					std::string k;
					if (!string(k) || !eat(':'))
						return ok = false;
				}
				if (!skip())
					return false;
				if (eat(','))
					continue;
				return eat(close) ? true : (ok = false);
			}
		}
		const size_t start = i;
		while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && s[i] != ' ' && s[i] != '\n' && s[i] != '\r' && s[i] != '\t')
			++i;
		return i > start ? true : (ok = false);
	}
};

struct FontConfig {
	std::string name;
	std::vector<std::string> sources;
};

bool parse_fonts_json(const std::string &text, std::vector<FontConfig> &out)
{
	JsonCursor c{text};
	if (!c.eat('['))
		return false;
	if (c.eat(']'))
		return true;
	for (;;) {
		if (!c.eat('{'))
			return false;
		FontConfig fc;
		bool has_name = false, has_sources = false;
		if (!c.eat('}')) {
			for (;;) {
				std::string key;
				if (!c.string(key) || !c.eat(':'))
					return false;
				if (key == "name") {
					if (!c.string(fc.name))
						return false;
					has_name = true;
				} else if (key == "sources") {
					if (!c.eat('['))
						return false;
					fc.sources.clear();
					if (!c.eat(']'))
						for (;;) {
							std::string v;
							if (!c.string(v))
								return false;
							fc.sources.push_back(v);
							if (c.eat(','))
								continue;
							if (!c.eat(']'))
								return false;
							break;
						}
					has_sources = true;
				} else if (!c.skip()) {
					return false;
				}
				if (c.eat(','))
					continue;
				if (!c.eat('}'))
					return false;
				break;
			}
		}
		if (!has_name || !has_sources)
			return false; // serde: missing field
		out.push_back(std::move(fc));
		if (c.eat(','))
			continue;
		if (!c.eat(']'))
			return false;
		break;
	}
	c.ws();
	return c.i == text.size();
}

bool has_font_extension(const std::string &path)
{
	const size_t slash = path.find_last_of('/');
	const size_t dot = path.find_last_of('.');
	if (dot == std::string::npos || (slash != std::string::npos && dot < slash))
		return false;
	const std::string ext = path.substr(dot + 1);
	return ext == "ttf" || ext == "otf"; // recurse.rs:106-108 (case-sensitive)
}
} // namespace

bool FontManager::scan(const std::string &path, std::string *err)
{
	struct stat st;
	if (stat(path.c_str(), &st) != 0)
		return true; // neither file nor directory: ignored like the reference's two is_* tests
	if (S_ISREG(st.st_mode)) {
		if (has_font_extension(path))
			return add_path(path, err);
		return true;
	}
	if (!S_ISDIR(st.st_mode))
		return true;
	const std::string manifest = path + "/fonts.json";
	struct stat ms;
	if (stat(manifest.c_str(), &ms) == 0) {
		std::ifstream in(manifest, std::ios::binary);
		std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
		if (!in.good() && !in.eof()) {
			if (err)
				*err = "Failed to read \"" + manifest + "\"";
			return false;
		}
		std::vector<FontConfig> configs;
		if (!parse_fonts_json(text, configs)) {
			if (err)
				*err = "invalid fonts.json: \"" + manifest + "\"";
			return false;
		}
		for (const FontConfig &c : configs) {
			std::vector<std::string> sources;
			for (const std::string &src : c.sources)
				sources.push_back(path + "/" + src);
			if (!add_font_with_name(c.name, sources, err))
				return false;
		}
		return true;
	}
	DIR *d = opendir(path.c_str());
	if (!d) {
		if (err)
			*err = "cannot read directory \"" + path + "\": " + std::strerror(errno);
		return false;
	}
	std::vector<std::string> names;
	while (const dirent *e = readdir(d)) {
		const std::string n = e->d_name;
		if (n != "." && n != "..")
			names.push_back(n);
	}
	closedir(d);
	std::sort(names.begin(), names.end());
	for (const std::string &n : names)
		if (!scan(path + "/" + n, err))
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
inline void cpu_pause()
{
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#else
	std::this_thread::yield();
#endif
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

bool FontManager::render_glyphs(Writer &writer, const Renderer &renderer, std::string *err, RenderStats *stats,
                                uint32_t shard, uint32_t n_shards, int threads) const
{
	// glibc serves allocations >= 128 KiB (every full block's PBF) with mmap: fresh pages, i.e. ~30 page faults per
	// file, every call.  Keep such blocks on the heap, where the pages freed after one call are reused by the next.
	// (Process-wide malloc tuning; VGB_KEEP_MALLOC_DEFAULTS=1 leaves the allocator alone.)
	static const bool malloc_tuned = [] {
		if (std::getenv("VGB_KEEP_MALLOC_DEFAULTS"))
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
	static const size_t kPartGlyphs = [] { // tuning knob (default 64; measured 16 / 32 / 64 on C2)
		const char *e = std::getenv("VGB_PART_GLYPHS");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 256 ? v : 64);
	}();
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
	uint32_t index = 0;
	size_t total_glyphs = 0;
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
			if (index++ % n_shards != shard)
				continue;
			BlockState *bs = &bsv[i];
			total_glyphs += bs->blk->len();
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
		total_threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
		total_threads = std::max(1, std::min(total_threads, 33));
	}
	const bool dedicated = total_threads >= kDedicatedMin;
	const int workers = dedicated ? total_threads - 1 : total_threads;
	// One submission carries whole blocks until it holds about `target` glyphs: small jobs keep one
	// block per submission (parallelism), big jobs amortise the per-submission cost.
	static const size_t kBatchesPerWorker = [] { // tuning knob (default 4; measured 1..6 on C2)
		const char *e = std::getenv("VGB_BATCHES_PER_WORKER");
		const long v = e ? std::atol(e) : 0;
		return (size_t)(v >= 1 && v <= 64 ? v : 4);
	}();
	constexpr int kEarlyWorkers = 4;
	static const bool kLatencyTail = [] { // VGB_LATENCY_TAIL=1: plan each worker's last batch for latency
		const char *e = std::getenv("VGB_LATENCY_TAIL"); // (measured on C2: 1.30-1.35 ms with, 1.26-1.34 without: off)
		return e && e[0] == '1';
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
	};
	std::mutex qm;
	std::condition_variable qcv;
	std::deque<Flight *> submit_q, done_q;
	size_t outstanding = 0; // batches submitted (or queued for it) and not yet encoded
	int workers_done = 0;
	std::atomic<uint64_t> submit_ns{0};
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
			{
				std::lock_guard<std::mutex> g(qm);
				if (!done_q.empty()) {
					done = done_q.front();
					done_q.pop_front();
				}
			}
			if (done) {
				idle_spins = 0;
				mark('e');
				bool ok = true;
				for (const Part &p : done->parts)
					if (ok)
						ok = finish_part(*p.todo, *done->batch, p.g0, p.g1);
				ok = ok && flush();
				renderer.release_batch(std::move(done->batch));
				delete done;
				{
					std::lock_guard<std::mutex> g(qm);
					--outstanding;
				}
				qcv.notify_all();
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
				cur->batch = renderer.acquire_batch();
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
					todo.bs->blk->append_to_batch(*cur->batch, todo.slot0, todo.slot1);
					cur->parts.push_back(Part{&todo, g0, cur->batch->glyphs().size()});
				}
				st.outline_ns += now_ns() - t0;
				if (cur->parts.empty()) {
					renderer.release_batch(std::move(cur->batch));
					{
						std::lock_guard<std::mutex> g(qm);
						--outstanding;
					}
					qcv.notify_all(); // a worker may be asleep waiting for exactly this reservation to go away
					continue;
				}
				st.glyphs += cur->batch->glyphs().size();
				st.bitmaps += cur->batch->job_count();
				st.pixels += cur->batch->bitmap_bytes();
				st.segments += cur->batch->total_segments();
				st.pairs += cur->batch->pairs();
				if (cur->batch->job_count() == 0) {
					// nothing to rasterise (empty blocks, or only bitmap-less glyphs): no GPU round trip
					bool ok = true;
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
					for (int spin = 0; spin < 4000 && !changed; ++spin) {
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
				std::lock_guard<std::mutex> g(qm);
				if (!submit_q.empty()) {
					f = submit_q.front();
					submit_q.pop_front();
				}
			}
			if (f) {
				std::string e;
				const uint64_t t0 = now_ns();
				if (trace)
					events[(size_t)workers].emplace_back('s', t0 - t_begin);
				const bool ok = renderer.submit_batch(*f->batch, &f->ticket, &e);
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
					}
					done_seq.fetch_add(1, std::memory_order_release);
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
			inflight.pop_front();
		}
	for (std::deque<Flight *> *q : {&submit_q, &done_q})
		for (Flight *f : *q) {
			if (f->batch)
				renderer.release_batch(std::move(f->batch));
			delete f;
		}
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
		}
		stats->submit_ns += submit_ns.load();
		stats->workers = (uint64_t)workers;
		stats->wall_ns = now_ns() - t_begin;
	}
	return !failed.load();
}

} // namespace vgb
