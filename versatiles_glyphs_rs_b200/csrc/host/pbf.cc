// pbf.cc — see pbf.h.
#include "pbf.h"

#include <cstring>

namespace vgb {

namespace {

inline size_t varint_size(uint64_t v)
{
	size_t n = 1;
	while (v >= 0x80) {
		v >>= 7;
		++n;
	}
	return n;
}
inline uint8_t *put_varint(uint8_t *p, uint64_t v)
{
	while (v >= 0x80) {
		*p++ = (uint8_t)(v | 0x80);
		v >>= 7;
	}
	*p++ = (uint8_t)v;
	return p;
}
inline uint32_t zigzag32(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }

size_t glyph_body_size(const PbfGlyph &g)
{
	size_t n = 1 + varint_size(g.id);
	if (g.has_bitmap)
		n += 1 + varint_size(g.bitmap.size()) + g.bitmap.size();
	n += 1 + varint_size(g.width);
	n += 1 + varint_size(g.height);
	n += 1 + varint_size(zigzag32(g.left));
	n += 1 + varint_size(zigzag32(g.top));
	n += 1 + varint_size(g.advance);
	return n;
}

} // namespace

std::vector<uint8_t> PbfGlyphs::into_vec() const
{
	std::vector<size_t> body(glyphs_.size());
	size_t stack = 1 + varint_size(name_.size()) + name_.size() + 1 + varint_size(range_.size()) + range_.size();
	for (size_t i = 0; i < glyphs_.size(); ++i) {
		body[i] = glyph_body_size(glyphs_[i]);
		stack += 1 + varint_size(body[i]) + body[i];
	}
	std::vector<uint8_t> out(1 + varint_size(stack) + stack);
	uint8_t *p = out.data();
	*p++ = 0x0A; // PbfGlyphs.stacks
	p = put_varint(p, stack);
	*p++ = 0x0A; // Fontstack.name
	p = put_varint(p, name_.size());
	std::memcpy(p, name_.data(), name_.size());
	p += name_.size();
	*p++ = 0x12; // Fontstack.range
	p = put_varint(p, range_.size());
	std::memcpy(p, range_.data(), range_.size());
	p += range_.size();
	for (size_t i = 0; i < glyphs_.size(); ++i) {
		const PbfGlyph &g = glyphs_[i];
		*p++ = 0x1A; // Fontstack.glyphs
		p = put_varint(p, body[i]);
		*p++ = 0x08;
		p = put_varint(p, g.id);
		if (g.has_bitmap) {
			*p++ = 0x12;
			p = put_varint(p, g.bitmap.size());
			if (!g.bitmap.empty())
				std::memcpy(p, g.bitmap.data(), g.bitmap.size());
			p += g.bitmap.size();
		}
		*p++ = 0x18;
		p = put_varint(p, g.width);
		*p++ = 0x20;
		p = put_varint(p, g.height);
		*p++ = 0x28;
		p = put_varint(p, zigzag32(g.left));
		*p++ = 0x30;
		p = put_varint(p, zigzag32(g.top));
		*p++ = 0x38;
		p = put_varint(p, g.advance);
	}
	return out;
}

namespace {
struct RangeItem {
	uint32_t id, width, height, left_zz, top_zz, advance;
	const uint8_t *bitmap; // nullptr = None
	size_t bitmap_len, body;
};
// gathers glyphs [g0, g1) and returns the encoded size of their Fontstack.glyphs entries
size_t gather_items(const GlyphBatch &batch, size_t g0, size_t g1, std::vector<RangeItem> &items)
{
	items.reserve(g1 - g0);
	size_t bytes = 0;
	for (size_t i = g0; i < g1; ++i) {
		const BatchGlyph &b = batch.glyphs()[i];
		RangeItem it;
		it.id = b.id;
		it.advance = b.advance;
		if (b.has_bitmap) {
			const PbfGlyph m = b.frame.into_pbf_glyph(b.id, b.advance); // metrics only (result.rs:66-76)
			const b200sdf_outline_job &j = batch.jobs()[b.job];
			it.width = m.width, it.height = m.height;
			it.left_zz = zigzag32(m.left), it.top_zz = zigzag32(m.top);
			it.bitmap = batch.bitmap_of(b);
			it.bitmap_len = (size_t)j.width * j.height;
		} else {
			it.width = it.height = 0;
			it.left_zz = it.top_zz = 0;
			it.bitmap = nullptr;
			it.bitmap_len = 0;
		}
		it.body = 1 + varint_size(it.id) + (it.bitmap ? 1 + varint_size(it.bitmap_len) + it.bitmap_len : 0) + 1 +
		          varint_size(it.width) + 1 + varint_size(it.height) + 1 + varint_size(it.left_zz) + 1 + varint_size(it.top_zz) +
		          1 + varint_size(it.advance);
		bytes += 1 + varint_size(it.body) + it.body;
		items.push_back(it);
	}
	return bytes;
}
uint8_t *put_items(uint8_t *p, const std::vector<RangeItem> &items)
{
	for (const RangeItem &it : items) {
		*p++ = 0x1A;
		p = put_varint(p, it.body);
		*p++ = 0x08;
		p = put_varint(p, it.id);
		if (it.bitmap) {
			*p++ = 0x12;
			p = put_varint(p, it.bitmap_len);
			std::memcpy(p, it.bitmap, it.bitmap_len);
			p += it.bitmap_len;
		}
		*p++ = 0x18;
		p = put_varint(p, it.width);
		*p++ = 0x20;
		p = put_varint(p, it.height);
		*p++ = 0x28;
		p = put_varint(p, it.left_zz);
		*p++ = 0x30;
		p = put_varint(p, it.top_zz);
		*p++ = 0x38;
		p = put_varint(p, it.advance);
	}
	return p;
}
size_t stack_header_size(const std::string &name, const std::string &range)
{
	return 1 + varint_size(name.size()) + name.size() + 1 + varint_size(range.size()) + range.size();
}
uint8_t *put_headers(uint8_t *p, const std::string &name, const std::string &range, size_t stack)
{
	*p++ = 0x0A;
	p = put_varint(p, stack);
	*p++ = 0x0A;
	p = put_varint(p, name.size());
	std::memcpy(p, name.data(), name.size());
	p += name.size();
	*p++ = 0x12;
	p = put_varint(p, range.size());
	std::memcpy(p, range.data(), range.size());
	return p + range.size();
}
} // namespace

std::vector<uint8_t> encode_batch_range(const std::string &name, const std::string &range, const GlyphBatch &batch, size_t g0,
                                        size_t g1)
{
	std::vector<RangeItem> items;
	const size_t stack = stack_header_size(name, range) + gather_items(batch, g0, g1, items);
	std::vector<uint8_t> out(1 + varint_size(stack) + stack);
	put_items(put_headers(out.data(), name, range, stack), items);
	return out;
}

std::vector<uint8_t> encode_glyph_entries(const GlyphBatch &batch, size_t g0, size_t g1)
{
	std::vector<RangeItem> items;
	std::vector<uint8_t> out(gather_items(batch, g0, g1, items));
	put_items(out.data(), items);
	return out;
}

std::vector<uint8_t> assemble_glyphs_pbf(const std::string &name, const std::string &range,
                                         const std::vector<std::vector<uint8_t>> &parts)
{
	size_t stack = stack_header_size(name, range);
	for (const auto &part : parts)
		stack += part.size();
	std::vector<uint8_t> out(1 + varint_size(stack) + stack);
	uint8_t *p = put_headers(out.data(), name, range, stack);
	for (const auto &part : parts) {
		if (!part.empty())
			std::memcpy(p, part.data(), part.size());
		p += part.size();
	}
	return out;
}

// ---- decoder ----------------------------------------------------------------------------------------
namespace {

struct Cursor {
	const uint8_t *p, *end;
	bool ok = true;
	uint64_t varint()
	{
		uint64_t v = 0;
		int shift = 0;
		while (p < end && shift < 64) {
			const uint8_t b = *p++;
			v |= (uint64_t)(b & 0x7F) << shift;
			if (!(b & 0x80))
				return v;
			shift += 7;
		}
		ok = false;
		return 0;
	}
	Cursor sub()
	{
		const uint64_t n = varint();
		if (!ok || n > (uint64_t)(end - p)) {
			ok = false;
			return Cursor{p, p, false};
		}
		Cursor c{p, p + n, true};
		p += n;
		return c;
	}
	bool skip(uint32_t wire)
	{
		switch (wire) {
		case 0:
			varint();
			return ok;
		case 1:
			if (end - p < 8)
				return ok = false;
			p += 8;
			return true;
		case 2:
			sub();
			return ok;
		case 5:
			if (end - p < 4)
				return ok = false;
			p += 4;
			return true;
		default:
			return ok = false;
		}
	}
};

bool decode_glyph(Cursor c, PbfGlyph &g)
{
	while (c.ok && c.p < c.end) {
		const uint64_t key = c.varint();
		const uint32_t tag = (uint32_t)(key >> 3), wire = (uint32_t)(key & 7);
		if (tag == 2 && wire == 2) {
			Cursor b = c.sub();
			if (!c.ok)
				return false;
			g.has_bitmap = true;
			g.bitmap.assign(b.p, b.end);
		} else if (wire == 0 && tag >= 1 && tag <= 7) {
			const uint64_t v = c.varint();
			switch (tag) {
			case 1:
				g.id = (uint32_t)v;
				break;
			case 3:
				g.width = (uint32_t)v;
				break;
			case 4:
				g.height = (uint32_t)v;
				break;
			case 5:
				g.left = (int32_t)(((uint32_t)v >> 1) ^ (uint32_t)-(int32_t)(v & 1));
				break;
			case 6:
				g.top = (int32_t)(((uint32_t)v >> 1) ^ (uint32_t)-(int32_t)(v & 1));
				break;
			case 7:
				g.advance = (uint32_t)v;
				break;
			default:
				break;
			}
		} else if (!c.skip(wire)) {
			return false;
		}
	}
	return c.ok;
}

} // namespace

bool pbf_decode(const uint8_t *data, size_t len, std::string &name, std::string &range, std::vector<PbfGlyph> &glyphs)
{
	Cursor top{data, data + len, true};
	while (top.ok && top.p < top.end) {
		const uint64_t key = top.varint();
		if ((key >> 3) == 1 && (key & 7) == 2) {
			Cursor st = top.sub();
			while (top.ok && st.ok && st.p < st.end) {
				const uint64_t k = st.varint();
				const uint32_t tag = (uint32_t)(k >> 3), wire = (uint32_t)(k & 7);
				if (wire == 2 && tag >= 1 && tag <= 3) {
					Cursor f = st.sub();
					if (!st.ok)
						return false;
					if (tag == 1)
						name.assign((const char *)f.p, (size_t)(f.end - f.p));
					else if (tag == 2)
						range.assign((const char *)f.p, (size_t)(f.end - f.p));
					else {
						PbfGlyph g;
						if (!decode_glyph(f, g))
							return false;
						glyphs.push_back(std::move(g));
					}
				} else if (!st.skip(wire)) {
					return false;
				}
			}
			if (!st.ok)
				return false;
		} else if (!top.skip((uint32_t)(key & 7))) {
			return false;
		}
	}
	return top.ok;
}

} // namespace vgb
