// font.h — block / font / manager orchestration (mirror of reference src/font/*.rs) and the
// writer facade (src/writer/mod.rs).
//
//   FontFileEntry  = font/file_entry.rs:13-56     owned font bytes + parsed Face + code point set
//   GlyphBlock     = font/glyph_block.rs:9-89     <= 256 code points, first file wins
//   FontWrapper    = font/wrapper.rs:15-76        files of one font id -> 256 BMP blocks
//   FontManager    = font/manager.rs:18-147       fonts by id; render_glyphs = the batch pipeline
//   Writer         = writer/mod.rs:27-96          directory / in-memory sink
#pragma once

#include <atomic>
#include <array>
#include <cstdio>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "face.h"
#include "pbf.h"
#include "render.h"

namespace vgb {

constexpr uint32_t GLYPH_BLOCK_SIZE = 256; // glyph_block.rs:7

// font/metadata.rs:20-64
struct FontMetadata {
	std::string name;   // name id 1 as stored in the font
	std::string family; // after parse_font_name stripped width / weight / script words
	std::vector<uint32_t> codepoints;
	std::string style = "normal";
	uint16_t weight = 400;
	std::string width = "normal";
	std::string generate_name() const;              // "<family> [<width>] <Weight> [<style>]", metadata.rs:29-55
	static FontMetadata from_face(const Face &face); // metadata.rs:84-129
};
// font/parse_font_name.rs:214-293
void parse_font_name(const std::string &family, const std::string &ps_name, std::string &out_family, std::string &style,
                     uint16_t &weight, std::string &width);
// font/index_files.rs:60-95
std::string encode_codeblocks(const std::vector<uint32_t> &codepoints);
// JSON string escaping as serde_json writes it (quotes, backslash, control characters)
std::string json_escape(const std::string &s);
class FontWrapper;
// font/index_files.rs:115-139; empty string + *err when a font has no files
std::string build_font_families_json(const std::map<std::string, FontWrapper> &fonts, std::string *err);

struct FontFileEntry {
	std::unique_ptr<Face> face;
	FontMetadata metadata;
	std::vector<uint32_t> codepoints; // = metadata.codepoints (metadata.rs:104-118)
	std::string family;               // = metadata.name, name id 1 (metadata.rs:97)
	// file_entry.rs:32-56; nullptr + *err on unparsable data
	static std::unique_ptr<FontFileEntry> from_bytes(std::vector<uint8_t> data, std::string *err);
	static std::unique_ptr<FontFileEntry> from_path(const std::string &path, std::string *err);
};

class GlyphBlock {
  public:
	explicit GlyphBlock(uint32_t start_index = 0) : start_index_(start_index) { slot_.fill(0); }
	uint32_t start_index() const { return start_index_; }
	void reset(uint32_t start_index)
	{
		start_index_ = start_index;
		slot_.fill(0);
		owners_.clear();
		len_ = 0;
	}
	// glyph_block.rs:34-36 — entry().or_insert(): the first file that has the code point keeps it
	void set_glyph_font(uint8_t char_index, const FontFileEntry *font)
	{
		if (slot_[char_index])
			return;
		size_t k = 0;
		while (k < owners_.size() && owners_[k] != font)
			++k;
		if (k == owners_.size())
			owners_.push_back(font);
		slot_[char_index] = (uint16_t)(k + 1);
		++len_;
	}
	size_t len() const { return len_; }
	bool is_empty() const { return len_ == 0; }
	const FontFileEntry *font_of(uint8_t char_index) const
	{
		const uint16_t k = slot_[char_index];
		return k ? owners_[k - 1u] : nullptr;
	}
	std::string range() const;    // "{start}-{start+255}"            glyph_block.rs:52-58
	std::string filename() const; // "{range}.pbf"                     glyph_block.rs:85-87
	// glyph_block.rs:69-80.  Glyphs are emitted in ascending code point order (the reference
	// iterates a HashMap, i.e. in unspecified order).
	bool render(const std::string &font_name, const Renderer &renderer, std::vector<uint8_t> &out, std::string *err) const;
	// Split form used by the pipeline: fill a batch, later turn the rendered batch into the PBF.
	bool fill_batch(GlyphBatch &batch) const;
	std::vector<uint8_t> encode_batch(const std::string &font_name, const GlyphBatch &batch) const;
	// Several blocks may share one batch: append this block's glyphs (no clear) / encode glyphs [g0, g1)
	// straight from the batch's bitmap buffer (no intermediate PbfGlyph copies).
	void append_to_batch(GlyphBatch &batch, uint32_t slot0 = 0, uint32_t slot1 = GLYPH_BLOCK_SIZE) const;
	std::vector<uint8_t> encode_range(const std::string &font_name, const GlyphBatch &batch, size_t g0, size_t g1) const;

  private:
	uint32_t start_index_;
	// slot -> 1 + index into owners_ (0 = no glyph): a block is 512 bytes, so building the 256 blocks of a
	// font on every render_glyphs call (as the reference does) touches 128 KiB instead of 512 KiB
	std::array<uint16_t, GLYPH_BLOCK_SIZE> slot_;
	std::vector<const FontFileEntry *> owners_; // the files that own at least one slot, in first-seen order
	size_t len_ = 0;
};

class FontWrapper {
  public:
	void add_file(std::unique_ptr<FontFileEntry> file)
	{
		files_.push_back(std::move(file));
		blocks_.clear(); // the block table is rebuilt on next use
		block_costs_.clear();
		block_units_.clear();
	}
	bool add_paths(const std::vector<std::string> &sources, std::string *err);
	const std::vector<std::unique_ptr<FontFileEntry>> &files() const { return files_; }
	// wrapper.rs:53-76 — always 256 BMP blocks, code points > 0xFFFF ignored
	std::vector<GlyphBlock> get_blocks() const;
	// the same assignment written into 256 caller-owned blocks (block i must start at i * 256)
	void assign_blocks(GlyphBlock *const *blocks) const;
	// get_blocks() as a table owned by the wrapper: a pure function of the files, so it is built when first
	// needed after the last add_file (i.e. at load time, like the parsed cmap) and reused by every render_glyphs
	const std::vector<GlyphBlock> &blocks() const;
	// Estimated rendering cost of each of the 256 blocks (pixel x segment pairs plus a per-glyph constant), from the glyf
	// headers only: what FontManager::render_glyphs balances shards with.  Like blocks(), a pure function of the
	// files, built when first needed.
	const std::vector<uint64_t> &block_costs() const;
	// The GPU share of those estimates in the device's own unit (4 x 4 pixel tile x segment): what a batch tells the
	// device about the work it is part of (b200sdf_submit_glyphs est_cost)
	const std::vector<uint64_t> &block_units() const;

  private:
	std::vector<std::unique_ptr<FontFileEntry>> files_;
	mutable std::vector<GlyphBlock> blocks_;
	mutable std::vector<uint64_t> block_costs_, block_units_;
	mutable std::mutex blocks_mu_;
};

// writer/mod.rs.  kinds: directory on disk (file.rs), in-memory recorder (dummy.rs + content).
class Writer {
  public:
	static Writer new_file(const std::string &folder);
	static Writer new_memory();
	// Writer::new_tar (writer/mod.rs:27-33, writer/tar.rs): a ustar stream, to a file (path) or kept in memory
	static Writer new_tar(const std::string &path);
	static Writer new_tar_memory();
	const std::vector<uint8_t> &tar_bytes() const { return tar_; } // new_tar_memory only
	bool write_file(const std::string &filename, const uint8_t *bytes, size_t len, std::string *err);
	bool write_file(const std::string &filename, std::vector<uint8_t> &&bytes, std::string *err); // no copy for the memory writer
	bool write_directory(const std::string &dirname, std::string *err);
	bool finish(std::string *err);
	struct Entry {
		std::string name;
		bool is_dir = false;
		std::vector<uint8_t> bytes;
	};
	const std::vector<Entry> &entries() const { return entries_; } // memory writer only
	uint64_t bytes_written() const { return __atomic_load_n(&bytes_written_, __ATOMIC_RELAXED); }
	// The directory sink writes every file on its own (open, write, close): calls for different files need no lock
	// between them.  (The reference serialises every write behind one mutex, manager.rs:108-111; the tar stream and the
	// in-memory list do need it.)
	bool concurrent_files() const { return to_disk_ && !to_tar_; }

  private:
	bool tar_header(const std::string &path, uint64_t size, uint64_t mode, char typeflag, std::string *err);
	bool tar_put(const uint8_t *p, size_t n, std::string *err);
	bool to_disk_ = false, to_tar_ = false;
	std::string folder_;
	std::vector<uint8_t> tar_;
	std::shared_ptr<std::FILE> tar_file_;
	std::shared_ptr<bool> tar_closed_; // finish() closes the file itself so that close errors are reported
	std::vector<Entry> entries_;
	uint64_t bytes_written_ = 0; // (updated with atomic adds: directory-sink writes run side by side; the class stays movable)
	bool finished_ = false;
};

struct RenderStats {
	uint64_t glyphs = 0, bitmaps = 0, pixels = 0, segments = 0, pairs = 0, pbf_bytes = 0, blocks = 0;
	// where the host time went, nanoseconds summed over workers (wall_ns: the whole call)
	uint64_t outline_ns = 0, submit_ns = 0, wait_ns = 0, encode_ns = 0, write_ns = 0, wall_ns = 0;
	uint64_t submits = 0, workers = 0;
	uint64_t handed_back = 0;               // glyphs the device decoder returned to the host recorder
	uint64_t h2d_bytes = 0;                 // bytes of glyph requests / records / segments the device read from host memory
	uint64_t cost_total = 0, cost_shard = 0; // estimated cost of the whole job and of this shard (0 when not sharded)
};

class FontManager {
  public:
	explicit FontManager(bool parallel) : parallel_(parallel) {}
	// manager.rs:39-53: id = name_to_id(family name); see DESIGN.md for the naming subset
	bool add_path(const std::string &path, std::string *err);
	bool add_paths(const std::vector<std::string> &paths, std::string *err);
	// manager.rs:66-75
	bool add_font_with_name(const std::string &name, const std::vector<std::string> &sources, std::string *err);
	// commands/recurse.rs:104-133 (`scan`): a .ttf / .otf file is added by path; a directory holding a fonts.json
	// contributes exactly the merged fonts that file lists ([{"name": ..., "sources": [...]}], sources relative to the
	// directory); any other directory is searched recursively.  Directory entries are visited in byte-wise name order
	// (the reference takes fs::read_dir order, i.e. unspecified — the order decides which file owns a shared code point).
	bool scan(const std::string &path, std::string *err);
	bool add_font_bytes_with_name(const std::string &name, std::vector<uint8_t> data, std::string *err);
	const std::map<std::string, FontWrapper> &fonts() const { return fonts_; }
	// manager.rs:81-125.  Blocks are flattened by host workers, rendered on the renderer's CUDA
	// streams and encoded while later blocks are in flight; writes go through one mutex.
	// shard/n_shards: render only tasks with (task_index % n_shards == shard) — the font x block
	// sharding used across GPUs (every shard still writes its own directories).
	bool render_glyphs(Writer &writer, const Renderer &renderer, std::string *err, RenderStats *stats = nullptr,
	                   uint32_t shard = 0, uint32_t n_shards = 1, int threads = 0) const;
	// The shard every (font, block) task belongs to when the job is cut into n_shards: owner[font * 256 + block], fonts in
	// id order.  Longest-processing-time-first over FontWrapper::block_costs; deterministic, so every
	// rank of a multi-GPU run computes the same table.  loads (optional): estimated cost per shard.
	void shard_owners(uint32_t n_shards, std::vector<uint16_t> &owner, std::vector<uint64_t> *loads = nullptr) const;
	// manager.rs:128-131 (font ids as a JSON array)
	bool write_index_json(Writer &writer, std::string *err) const;
	// manager.rs:134-137 (font_families.json, index_files.rs:115-139)
	bool write_families_json(Writer &writer, std::string *err) const;
	static std::string name_to_id(const std::string &name); // manager.rs:141-147

  private:
	std::map<std::string, FontWrapper> fonts_; // ordered: deterministic task order
	bool parallel_;
};

} // namespace vgb
