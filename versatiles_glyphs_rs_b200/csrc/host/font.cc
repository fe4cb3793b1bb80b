// font.cc — FontFileEntry, GlyphBlock, FontWrapper and the bookkeeping half of FontManager (mirror of reference
// src/font/{file_entry,glyph_block,wrapper,manager}.rs); the pipeline is in pipeline.cc, the sinks in writer.cc, the
// directory rules in scan.cc.  See font.h.
#include "font.h"

#include <algorithm>
#include <cstring>
#include <fstream>

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
	if (batch->failed()) {
		if (err)
			*err = batch->failure();
		renderer.release_batch(std::move(batch));
		return false;
	}
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

const std::vector<uint64_t> &FontWrapper::block_costs() const
{
	const std::vector<GlyphBlock> &table = blocks();
	std::lock_guard<std::mutex> g(blocks_mu_);
	if (!block_costs_.empty())
		return block_costs_;
	std::vector<uint64_t> costs(table.size(), 0);
	std::vector<uint64_t> units(table.size(), 0);
	std::vector<Face::GlyfPart> parts;
	for (size_t b = 0; b < table.size(); ++b) {
		uint64_t c = 64; // an empty block still costs a file
		uint64_t u = 0;
		for (uint32_t k = 0; k < GLYPH_BLOCK_SIZE; ++k) {
			const FontFileEntry *f = table[b].font_of((uint8_t)k);
			if (!f)
				continue;
			c += 40000; // per glyph: request, decode, encode
			const auto gid = f->face->glyph_index(table[b].start_index() + k);
			if (!gid)
				continue;
			parts.clear();
			const Face::GlyfPlan plan = f->face->glyf_parts(*gid, parts);
			if (plan == Face::GlyfPlan::Host) {
				c += 600000; // no header to go by: a typical glyph
				u += 600000 / 16;
				continue;
			}
			const double scale = (double)GLYPH_SIZE / (double)std::max<uint16_t>(1, f->face->units_per_em());
			for (const Face::GlyfPart &p : parts) {
				const double w = ((double)p.xmax - (double)p.xmin) * scale + 8.0, h = ((double)p.ymax - (double)p.ymin) * scale + 8.0;
				const double area = std::min(std::max(w, 8.0), 4096.0) * std::min(std::max(h, 8.0), 4096.0);
				c += (uint64_t)(area * (double)p.points * 8.0); // ~8 flattened segments per point
				u += (uint64_t)(area / 16.0 * ((double)p.points * 8.0 + 8.0));
			}
		}
		costs[b] = c;
		units[b] = u;
	}
	block_costs_ = std::move(costs);
	block_units_ = std::move(units);
	return block_costs_;
}

const std::vector<uint64_t> &FontWrapper::block_units() const
{
	block_costs();
	return block_units_;
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

// ---- FontManager -------------------------------------------------------------------------------------
namespace {
// Unicode simple lowercase for the scripts font names are written in (Rust's str::to_lowercase, manager.rs:144, maps
// every cased letter; this table covers ASCII, Latin-1, Latin Extended-A, Greek and Cyrillic — the context rule for a
// final capital sigma is not applied).
uint32_t lower_cp(uint32_t c)
{
	if (c >= 'A' && c <= 'Z')
		return c + 32;
	if (c < 0xC0)
		return c;
	if (c <= 0xDE)
		return c == 0xD7 ? c : c + 32;
	if (c >= 0x100 && c <= 0x137)
		return c == 0x130 ? 'i' : ((c & 1) ? c : c + 1); // (U+0130 lowers to "i" + U+0307; the dot is dropped here)
	if (c >= 0x139 && c <= 0x148)
		return (c & 1) ? c + 1 : c;
	if (c >= 0x14A && c <= 0x177)
		return (c & 1) ? c : c + 1;
	if (c == 0x178)
		return 0xFF;
	if (c >= 0x179 && c <= 0x17E)
		return (c & 1) ? c + 1 : c;
	if (c == 0x386)
		return 0x3AC;
	if (c >= 0x388 && c <= 0x38A)
		return c + 37;
	if (c == 0x38C)
		return 0x3CC;
	if (c == 0x38E || c == 0x38F)
		return c + 63;
	if (c >= 0x391 && c <= 0x3AB)
		return c == 0x3A2 ? c : c + 32;
	if (c >= 0x400 && c <= 0x40F)
		return c + 80;
	if (c >= 0x410 && c <= 0x42F)
		return c + 32;
	if (c >= 0x460 && c <= 0x481)
		return (c & 1) ? c : c + 1;
	if (c >= 0x48A && c <= 0x4BF)
		return (c & 1) ? c : c + 1;
	return c;
}
// char::is_whitespace (what str::trim removes)
bool unicode_space(uint32_t c)
{
	return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
	       c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
void put_utf8(std::string &out, uint32_t cp)
{
	if (cp < 0x80)
		out.push_back((char)cp);
	else if (cp < 0x800) {
		out.push_back((char)(0xC0 | (cp >> 6)));
		out.push_back((char)(0x80 | (cp & 0x3F)));
	} else if (cp < 0x10000) {
		out.push_back((char)(0xE0 | (cp >> 12)));
		out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
		out.push_back((char)(0x80 | (cp & 0x3F)));
	} else {
		out.push_back((char)(0xF0 | (cp >> 18)));
		out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
		out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
		out.push_back((char)(0x80 | (cp & 0x3F)));
	}
}
} // namespace

std::string FontManager::name_to_id(const std::string &name)
{
	// manager.rs:141-147: to_lowercase (Unicode); every run of [-_\s] — regex-lite's \s is ASCII-only — becomes one
	// space; trim (Unicode whitespace, both ends); spaces -> '_'
	std::vector<uint32_t> cps;
	for (size_t i = 0; i < name.size();) {
		const unsigned char b = (unsigned char)name[i];
		uint32_t cp = b;
		size_t n = 1;
		if (b >= 0xF0 && i + 3 < name.size())
			cp = ((b & 7u) << 18) | (((unsigned char)name[i + 1] & 0x3Fu) << 12) | (((unsigned char)name[i + 2] & 0x3Fu) << 6) |
			     ((unsigned char)name[i + 3] & 0x3Fu), n = 4;
		else if (b >= 0xE0 && i + 2 < name.size())
			cp = ((b & 15u) << 12) | (((unsigned char)name[i + 1] & 0x3Fu) << 6) | ((unsigned char)name[i + 2] & 0x3Fu), n = 3;
		else if (b >= 0xC0 && i + 1 < name.size())
			cp = ((b & 31u) << 6) | ((unsigned char)name[i + 1] & 0x3Fu), n = 2;
		i += n;
		cp = lower_cp(cp);
		const bool sep = cp == '-' || cp == '_' || cp == ' ' || (cp >= 9 && cp <= 13);
		if (sep) {
			if (cps.empty() || cps.back() != ' ')
				cps.push_back(' ');
		} else {
			cps.push_back(cp);
		}
	}
	size_t a = 0, b = cps.size();
	while (a < b && unicode_space(cps[a]))
		++a;
	while (b > a && unicode_space(cps[b - 1]))
		--b;
	std::string out;
	for (size_t i = a; i < b; ++i)
		put_utf8(out, cps[i] == ' ' ? (uint32_t)'_' : cps[i]);
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
			s += "  \"" + json_escape(kv.first) + "\""; // (serde_json escapes: an id may hold quotes or control bytes)
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

} // namespace vgb

