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
	std::vector<Face::GlyfPart> parts;
	for (size_t b = 0; b < table.size(); ++b) {
		uint64_t c = 64; // an empty block still costs a file
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
				continue;
			}
			const double scale = (double)GLYPH_SIZE / (double)std::max<uint16_t>(1, f->face->units_per_em());
			for (const Face::GlyfPart &p : parts) {
				const double w = ((double)p.xmax - (double)p.xmin) * scale + 8.0, h = ((double)p.ymax - (double)p.ymin) * scale + 8.0;
				const double area = std::min(std::max(w, 8.0), 4096.0) * std::min(std::max(h, 8.0), 4096.0);
				c += (uint64_t)(area * (double)p.points * 8.0); // ~8 flattened segments per point
			}
		}
		costs[b] = c;
	}
	block_costs_ = std::move(costs);
	return block_costs_;
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

} // namespace vgb

