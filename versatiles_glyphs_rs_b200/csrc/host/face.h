// face.h — the subset of ttf-parser 0.25.1's `Face` that the reference's rendering path calls
// (crate not vendored in the reference; call sites: src/render/renderer.rs:106,107,110,115,
// src/font/file_entry.rs:48, src/font/metadata.rs:91-116).  TrueType `glyf` outlines, cmap
// formats 0/4/6/10/12/13, hmtx advances, name table strings; `CFF ` outlines through cff.h when the face
// has no usable glyf / loca pair (ttf-parser's order: glyf, then cff).  No CFF2, no variations (none of
// the reference's fixtures use them).
#pragma once

#include <atomic>
#include <cstdint>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace vgb {

class CffTable;

// ttf_parser::OutlineBuilder — coordinates arrive as f32 font units
class OutlineBuilder {
  public:
	virtual ~OutlineBuilder() = default;
	virtual void move_to(float x, float y) = 0;
	virtual void line_to(float x, float y) = 0;
	virtual void quad_to(float x1, float y1, float x, float y) = 0;
	virtual void curve_to(float x1, float y1, float x2, float y2, float x, float y) = 0;
	virtual void close() = 0;
};

class Face {
  public:
	// Face::parse(data, 0): nullptr when the mandatory tables are missing / truncated
	static std::unique_ptr<Face> parse(std::vector<uint8_t> data);
	~Face();

	uint16_t units_per_em() const { return upm_; }
	uint16_t number_of_glyphs() const { return num_glyphs_; }
	// first unicode cmap subtable (table order) that maps the code point
	std::optional<uint16_t> glyph_index(uint32_t codepoint) const;
	std::optional<uint16_t> glyph_hor_advance(uint16_t glyph_id) const;
	// returns false when the glyph has no outline (no callbacks were made)
	bool outline_glyph(uint16_t glyph_id, OutlineBuilder &builder) const;
	// sorted union over unicode cmap subtables of the code points they map (metadata.rs:104-118)
	std::vector<uint32_t> codepoints() const;
	// name table entry (first record with that id that decodes; UTF-16BE / Mac Roman), "" if none
	std::string name(uint16_t name_id) const;

	// ---- device-side glyf decoding (csrc/glyf_kernel.cuh) ----
	// How outline_glyph(gid) can be handed to the device: as the simple-glyph records it is made of, each only
	// translated (appended to `parts`), or not at all (scaled / rotated components, CFF outlines, more points than the
	// decoder holds: the caller records the outline on the host), or there is nothing to draw.
	enum class GlyfPlan { None, Parts, Host };
	struct GlyfPart {
		uint32_t off, len;   // byte range inside the glyf table
		float ox, oy;        // translation (font units)
		uint32_t points;     // numberOfPoints of the record
		int16_t xmin, ymin, xmax, ymax; // the record's header bounding box (not trusted: the device checks the fit)
	};
	GlyfPlan glyf_parts(uint16_t glyph_id, std::vector<GlyfPart> &parts) const;
	const uint8_t *glyf_data() const { return data_.data() + glyf_.off; }
	size_t glyf_size() const { return glyf_.len; }
	// process-wide unique number of this face (addresses are reused once a face is freed; this is not)
	uint64_t uid() const { return uid_; }
	// handle of this face's glyf table on a device context, remembered per face: tag = (owner id << 32) | (handle + 1)
	uint64_t device_tag() const { return device_tag_.load(std::memory_order_acquire); }
	void set_device_tag(uint64_t t) const { device_tag_.store(t, std::memory_order_release); }

  private:
	struct Span {
		size_t off = 0, len = 0;
	};
	struct CmapSubtable {
		uint16_t platform, encoding, format;
		Span data;
		bool is_unicode() const
		{
			return platform == 0 || (platform == 3 && encoding == 1) ||
			       (platform == 3 && encoding == 10 && (format == 12 || format == 13));
		}
	};
	struct Transform {
		float a = 1.f, b = 0.f, c = 0.f, d = 1.f, e = 0.f, f = 0.f;
		bool is_default() const { return a == 1.f && b == 0.f && c == 0.f && d == 1.f && e == 0.f && f == 0.f; }
		static Transform combine(const Transform &t1, const Transform &t2);
		void apply_to(float &x, float &y) const;
	};
	class ContourEmitter;

	bool table(const char tag[4], Span &out) const;
	uint16_t u16(size_t off) const { return (uint16_t)((data_[off] << 8) | data_[off + 1]); }
	int16_t i16(size_t off) const { return (int16_t)u16(off); }
	uint32_t u32(size_t off) const
	{
		return ((uint32_t)data_[off] << 24) | ((uint32_t)data_[off + 1] << 16) | ((uint32_t)data_[off + 2] << 8) |
		       (uint32_t)data_[off + 3];
	}
	std::optional<uint16_t> lookup(const CmapSubtable &s, uint32_t cp) const;
	template <typename F> void enumerate(const CmapSubtable &s, F &&f) const;
	bool glyph_range(uint16_t gid, Span &out) const;
	void outline_impl(Span glyph, int depth, const Transform &t, OutlineBuilder &b) const;
	bool parts_impl(Span glyph, int depth, const Transform &t, std::vector<GlyfPart> &parts) const;
	mutable std::atomic<uint64_t> device_tag_{0};
	uint64_t uid_ = 0;

	std::vector<uint8_t> data_;
	uint16_t upm_ = 0, num_glyphs_ = 0, num_hmetrics_ = 0;
	bool loca_long_ = false;
	Span hmtx_, loca_, glyf_, cmap_, name_;
	std::vector<CmapSubtable> subtables_;
	std::unique_ptr<CffTable> cff_, cff2_;
	Face();
};

} // namespace vgb
