// cff.h — the part of ttf-parser 0.25.1's `cff::Table` (CFF version 1) that `Face::outline_glyph` reaches
// (crate not vendored in the reference; call site src/render/renderer.rs:110, which ignores the result, so the
// callbacks made before a charstring error count).  Type 2 charstrings → move_to / line_to / curve_to / close
// in f32 font units; name-keyed (SID) and CID-keyed fonts (FDSelect formats 0 / 3), local and global
// subroutines, `seac` (accented composites in `endchar`, through StandardEncoding and the charset).  FontMatrix is
// read past but not applied (ttf-parser exposes it through `Table::matrix()` without transforming the outline).
// SURVEY.md §8(f) rank 3.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>

namespace vgb {

class OutlineBuilder;

class CffTable {
  public:
	// nullptr where cff::Table::parse returns None (the face then has no CFF outlines at all)
	static std::unique_ptr<CffTable> parse(const uint8_t *data, size_t len);
	// the same for a `CFF2` table (cff2::Table::parse); axis_count = the face's variation coordinates (fvar axes, all 0)
	static std::unique_ptr<CffTable> parse2(const uint8_t *data, size_t len, uint16_t axis_count);
	// false where Table::outline returns Err (callbacks may already have been made)
	bool outline(uint16_t glyph_id, OutlineBuilder &builder) const;
	uint32_t number_of_glyphs() const { return char_strings_.count; }

	struct Bytes {
		const uint8_t *p = nullptr;
		size_t len = 0;
	};
	// INDEX: count, offSize, (count + 1) one-based offsets, data
	struct Index {
		uint32_t count = 0;
		uint8_t off_size = 0;
		const uint8_t *offsets = nullptr;
		Bytes data;
		bool get(uint32_t i, Bytes &out) const;
	};

  private:
	struct Interp;
	bool run(Interp &in, Bytes code, int depth) const;
	bool cid_local_subrs(uint16_t glyph_id, Index &out) const;
	bool seac_glyph(float code, uint16_t &glyph_id) const;
	bool blend_scalars(uint16_t index, float *scalars, int &count) const;

	Bytes table_;
	Index global_subrs_, char_strings_, local_subrs_, fd_array_;
	bool cid_ = false;
	uint8_t fd_select_format_ = 0;
	Bytes fd_select_; // format 0: one byte per glyph; format 3: everything after the format byte
	// charset: 0 ISOAdobe / 1 Expert / 2 ExpertSubset (predefined), 3 + f for a table of format f (records in charset_)
	int charset_kind_ = 0;
	Bytes charset_;
	uint32_t charset_records_ = 0;
	// CFF 2: ItemVariationStore (regions x axes of start / peak / end, data subtables -> region indices)
	bool cff2_ = false, has_var_store_ = false;
	uint16_t axis_count_ = 0, region_axes_ = 0, n_regions_ = 0;
	Bytes var_store_, var_data_offsets_, regions_;
};

} // namespace vgb
