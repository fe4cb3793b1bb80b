// pbf.h — Mapbox/MapLibre "glyphs" protobuf writer (mirror of reference src/protobuf/*.rs, which
// derives the encoding with prost 0.14: proto2 required scalars are always written, the optional
// `bitmap` is omitted when None, fields appear in tag order, left/top are sint32 zig-zag).
//
//   PbfGlyphs { 1: repeated Fontstack }                         protobuf/glyphs.rs:11-16
//   Fontstack { 1: name, 2: range, 3: repeated PbfGlyph }       protobuf/fontstack.rs:9-25
//   PbfGlyph  { 1: id, 2: bitmap?, 3: width, 4: height,
//               5: sint32 left, 6: sint32 top, 7: advance }     protobuf/glyph.rs:10-41
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "render.h"

namespace vgb {

class PbfGlyphs {
  public:
	PbfGlyphs(std::string name, std::string range) : name_(std::move(name)), range_(std::move(range)) {}
	void push(PbfGlyph g) { glyphs_.push_back(std::move(g)); }
	size_t len() const { return glyphs_.size(); }
	const std::vector<PbfGlyph> &glyphs() const { return glyphs_; }
	// protobuf/glyphs.rs:66-70
	std::vector<uint8_t> into_vec() const;

  private:
	std::string name_, range_;
	std::vector<PbfGlyph> glyphs_;
};

// Same bytes as PbfGlyphs(name, range) + push(batch.take_glyph(i)) for i in [g0, g1) + into_vec(),
// written straight from the batch's bitmap buffer (one allocation, one copy per bitmap).
std::vector<uint8_t> encode_batch_range(const std::string &name, const std::string &range, const GlyphBatch &batch, size_t g0,
                                        size_t g1);

// A block rendered in several parts (FontManager::render_glyphs splits full blocks so that no worker is
// stuck with 256 glyphs): the Fontstack.glyphs entries of glyphs [g0, g1) alone, and the final message
// built from the parts in code point order — byte-identical to encode_batch_range over the whole block.
std::vector<uint8_t> encode_glyph_entries(const GlyphBatch &batch, size_t g0, size_t g1);
std::vector<uint8_t> assemble_glyphs_pbf(const std::string &name, const std::string &range,
                                         const std::vector<std::vector<uint8_t>> &parts);

// Decoder for tests / the debug differ (mirror of commands/debug.rs:38-98's prost decode).
bool pbf_decode(const uint8_t *data, size_t len, std::string &name, std::string &range, std::vector<PbfGlyph> &glyphs);

} // namespace vgb
