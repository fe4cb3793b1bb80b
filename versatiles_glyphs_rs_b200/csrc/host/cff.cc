// cff.cc — CFF 1 outlines for Face::outline_glyph (see cff.h for scope and the reference call site).
#include "cff.h"

#include <cmath>
#include <cstdlib>
#include <vector>

#include "face.h"

namespace vgb {
namespace {

using Bytes = CffTable::Bytes;
using Index = CffTable::Index;

// Bounds-checked big-endian cursor; `ok` latches to false on the first short read (ttf-parser's `?`).
struct Cursor {
	const uint8_t *p;
	size_t len, pos = 0;
	bool ok = true;
	explicit Cursor(Bytes b, size_t at = 0) : p(b.p), len(b.len), pos(at)
	{
		if (at > len)
			ok = false;
	}
	bool at_end() const { return !ok || pos >= len; }
	bool need(size_t n)
	{
		if (!ok || len - pos < n)
			ok = false;
		return ok;
	}
	uint8_t u8() { return need(1) ? p[pos++] : 0; }
	uint16_t u16()
	{
		if (!need(2))
			return 0;
		const uint16_t v = (uint16_t)((p[pos] << 8) | p[pos + 1]);
		pos += 2;
		return v;
	}
	uint32_t u32()
	{
		if (!need(4))
			return 0;
		const uint32_t v = ((uint32_t)p[pos] << 24) | ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3];
		pos += 4;
		return v;
	}
	Bytes take(size_t n)
	{
		Bytes b;
		if (need(n)) {
			b.p = p + pos, b.len = n;
			pos += n;
		}
		return b;
	}
	void skip(size_t n)
	{
		if (need(n))
			pos += n;
	}
};

uint32_t read_offset(const uint8_t *p, uint8_t size)
{
	uint32_t v = 0;
	for (uint8_t i = 0; i < size; ++i)
		v = (v << 8) | p[i];
	return v;
}

// parse_index::<u16> (CFF 1) / parse_index::<u32> (CFF 2: `wide`): false = None.  count 0 is the empty INDEX.
bool parse_index(Cursor &c, Index &out, bool wide = false)
{
	out = Index();
	const uint32_t count = wide ? c.u32() : c.u16();
	if (!c.ok || count == 0xffffffffu)
		return false;
	if (count == 0)
		return true;
	const uint8_t off_size = c.u8();
	if (!c.ok || off_size < 1 || off_size > 4)
		return false;
	const Bytes offs = c.take(((size_t)count + 1) * off_size);
	if (!c.ok)
		return false;
	const uint32_t last = read_offset(offs.p + (size_t)count * off_size, off_size);
	if (last == 0) // offsets are one-based: no valid last offset, an empty INDEX and nothing consumed
		return true;
	const Bytes data = c.take(last - 1);
	if (!c.ok)
		return false;
	out.count = count, out.off_size = off_size, out.offsets = offs.p, out.data = data;
	return true;
}

// DICT operand / operator stream (dict.rs): collects the operands in front of each operator.
struct Dict {
	static constexpr int kMaxOperands = 48;
	Cursor c;
	double operands[kMaxOperands];
	int n = 0;
	bool operands_ok = true;
	explicit Dict(Bytes b) : c(b) {}

	static bool is_operator(uint8_t b) { return b <= 27 || b == 31 || b == 255; }

	static bool parse_real(Cursor &c, double &v)
	{
		char buf[66];
		int k = 0;
		bool done = false;
		while (!done) {
			if (c.at_end())
				return false;
			const uint8_t b = c.u8();
			for (int h = 0; h < 2 && !done; ++h) {
				const uint8_t nib = h == 0 ? (b >> 4) : (b & 15);
				if (nib == 0xf) {
					done = true;
					break;
				}
				if (k >= 62)
					return false;
				if (nib <= 9)
					buf[k++] = (char)('0' + nib);
				else if (nib == 0xa)
					buf[k++] = '.';
				else if (nib == 0xb)
					buf[k++] = 'E';
				else if (nib == 0xc)
					buf[k++] = 'E', buf[k++] = '-';
				else if (nib == 0xe)
					buf[k++] = '-';
				else
					return false;
			}
		}
		buf[k] = 0;
		char *end = nullptr;
		v = std::strtod(buf, &end);
		return k > 0 && end == buf + k;
	}

	static bool parse_number(uint8_t b0, Cursor &c, double &v)
	{
		if (b0 == 28) {
			v = (int16_t)c.u16();
		} else if (b0 == 29) {
			v = (int32_t)c.u32();
		} else if (b0 == 30) {
			return parse_real(c, v);
		} else if (b0 >= 32 && b0 <= 246) {
			v = (int)b0 - 139;
		} else if (b0 >= 247 && b0 <= 250) {
			v = ((int)b0 - 247) * 256 + (int)c.u8() + 108;
		} else if (b0 >= 251 && b0 <= 254) {
			v = -((int)b0 - 251) * 256 - (int)c.u8() - 108;
		} else {
			return false;
		}
		return c.ok;
	}

	// next operator (two-byte ones as 1200 + second byte), -1 at the end of the data or on a malformed number
	int next()
	{
		n = 0;
		operands_ok = true;
		while (!c.at_end()) {
			const uint8_t b = c.u8();
			if (is_operator(b)) {
				int op = b;
				if (b == 12) {
					op = 1200 + c.u8();
					if (!c.ok)
						return -1;
				}
				return op;
			}
			double v;
			if (!parse_number(b, c, v))
				return -1;
			if (n < kMaxOperands)
				operands[n++] = v;
		}
		return -1;
	}
	bool offset(size_t &out) const
	{
		if (n != 1 || (int32_t)operands[0] < 0)
			return false;
		out = (size_t)(int32_t)operands[0];
		return true;
	}
	bool range(size_t &start, size_t &end) const
	{
		if (n != 2 || (int32_t)operands[0] < 0 || (int32_t)operands[1] < 0)
			return false;
		start = (size_t)(int32_t)operands[1];
		end = start + (size_t)(int32_t)operands[0];
		return true;
	}
};

bool sub(Bytes t, size_t start, size_t end, Bytes &out)
{
	if (start > end || end > t.len)
		return false;
	out.p = t.p + start, out.len = end - start;
	return true;
}

// Subrs offset of a Private DICT, if it has one (operator 19)
bool private_subrs_offset(Bytes priv, size_t &off)
{
	Dict d(priv);
	bool have = false;
	for (int op = d.next(); op >= 0; op = d.next())
		if (op == 19)
			have = d.offset(off);
	return have;
}

// parse_charset: the records behind the format byte (format 0: SIDs of glyphs 1.., formats 1 / 2: ranges that
// together cover every glyph but .notdef); false = malformed, which makes cff::Table::parse fail
bool parse_charset(Bytes t, size_t off, uint16_t n_glyphs, uint8_t &format, Bytes &records, uint32_t &n_records)
{
	Cursor c(t, off);
	format = c.u8();
	if (!c.ok)
		return false;
	const size_t begin = c.pos;
	n_records = 0;
	if (format == 0) {
		n_records = (uint32_t)n_glyphs - 1;
		c.skip((size_t)n_records * 2);
	} else if (format == 1 || format == 2) {
		uint32_t left = (uint32_t)n_glyphs - 1;
		while (left > 0) {
			c.skip(2);
			const uint32_t n = (format == 1 ? c.u8() : c.u16()) + 1u;
			if (!c.ok || n > left)
				return false;
			left -= n;
			n_records++;
		}
	} else {
		return false;
	}
	if (!c.ok)
		return false;
	records.p = t.p + begin, records.len = c.pos - begin;
	return true;
}

// Adobe StandardEncoding: character code -> SID (0 = .notdef)
const uint8_t kStandardEncoding[256] = {
	0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
	1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
	33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64,
	65, 66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 0,
	0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
	0, 96, 97, 98, 99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 0, 111, 112, 113, 114, 0, 115, 116, 117, 118, 119, 120, 121, 122, 0, 123,
	0, 124, 125, 126, 127, 128, 129, 130, 131, 0, 132, 133, 0, 134, 135, 136, 137, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
	0, 138, 0, 139, 0, 0, 0, 0, 140, 141, 142, 143, 0, 0, 0, 0, 0, 144, 0, 0, 0, 145, 0, 0, 146, 147, 148, 149, 0, 0, 0, 0,
};

bool encoding_well_formed(Bytes t, size_t off)
{
	Cursor c(t, off);
	const uint8_t format = c.u8();
	const uint8_t count = c.u8();
	if (!c.ok)
		return false;
	if ((format & 0x7f) == 0)
		c.skip(count);
	else if ((format & 0x7f) == 1)
		c.skip((size_t)count * 2);
	else
		return false;
	if (format & 0x80)
		c.skip((size_t)c.u8() * 3);
	return c.ok;
}

constexpr int kStackLimit = 10;   // subroutine nesting
constexpr int kMaxArguments = 48; // Type 2 argument stack
constexpr int kMaxArguments2 = 513; // CFF 2 argument stack

} // namespace

bool Index::get(uint32_t i, Bytes &out) const
{
	if (i >= count)
		return false;
	const uint32_t a = read_offset(offsets + (size_t)i * off_size, off_size);
	const uint32_t b = read_offset(offsets + ((size_t)i + 1) * off_size, off_size);
	if (a == 0 || b == 0)
		return false;
	return sub(data, a - 1, b - 1, out);
}

std::unique_ptr<CffTable> CffTable::parse(const uint8_t *data, size_t len)
{
	std::unique_ptr<CffTable> t(new CffTable());
	t->table_.p = data, t->table_.len = len;
	Cursor c(t->table_);
	const uint8_t major = c.u8();
	c.skip(1); // minor
	const uint8_t header_size = c.u8();
	c.skip(1); // absolute offset size
	if (!c.ok || major != 1)
		return nullptr;
	if (header_size > 4)
		c.skip(header_size - 4u);
	Index names, top, strings;
	if (!parse_index(c, names) || !parse_index(c, top))
		return nullptr;
	Bytes top_dict;
	if (!top.get(0, top_dict))
		return nullptr;

	size_t charset_off = 0, encoding_off = 0, char_strings_off = 0, priv_start = 0, priv_end = 0, fd_array_off = 0, fd_select_off = 0;
	bool has_charset = false, has_encoding = false, has_priv = false, has_ros = false, has_fd_array = false, has_fd_select = false;
	{
		Dict d(top_dict);
		for (int op = d.next(); op >= 0; op = d.next()) {
			switch (op) {
			case 15: has_charset = d.offset(charset_off); break;
			case 16: has_encoding = d.offset(encoding_off); break;
			case 17:
				if (!d.offset(char_strings_off))
					return nullptr;
				break;
			case 18: has_priv = d.range(priv_start, priv_end); break;
			case 1230: has_ros = true; break;
			case 1236: has_fd_array = d.offset(fd_array_off); break;
			case 1237: has_fd_select = d.offset(fd_select_off); break;
			default: break; // FontMatrix (1207) and the rest: read past
			}
		}
	}
	if (char_strings_off == 0)
		return nullptr;
	if (!parse_index(c, strings) || !parse_index(c, t->global_subrs_))
		return nullptr;
	{
		Cursor cs(t->table_, char_strings_off);
		if (!parse_index(cs, t->char_strings_) || t->char_strings_.count == 0)
			return nullptr;
	}
	const uint16_t n_glyphs = (uint16_t)t->char_strings_.count;
	if (has_charset && charset_off <= 2) {
		t->charset_kind_ = (int)charset_off;
	} else if (has_charset) {
		uint8_t format;
		if (!parse_charset(t->table_, charset_off, n_glyphs, format, t->charset_, t->charset_records_))
			return nullptr;
		t->charset_kind_ = 3 + format;
	}

	if (has_ros) {
		if (!has_charset || !has_fd_array || !has_fd_select || charset_off == 0 || fd_array_off == 0 || fd_select_off == 0)
			return nullptr;
		t->cid_ = true;
		Cursor fa(t->table_, fd_array_off);
		if (!parse_index(fa, t->fd_array_))
			return nullptr;
		Cursor fs(t->table_, fd_select_off);
		t->fd_select_format_ = fs.u8();
		if (!fs.ok)
			return nullptr;
		if (t->fd_select_format_ == 0) {
			t->fd_select_ = fs.take(n_glyphs);
			if (!fs.ok)
				return nullptr;
		} else if (t->fd_select_format_ == 3) {
			t->fd_select_ = fs.take(fs.len - fs.pos);
		} else {
			return nullptr;
		}
	} else {
		if (has_encoding && encoding_off > 1 && !encoding_well_formed(t->table_, encoding_off))
			return nullptr;
		if (has_priv) {
			Bytes priv;
			if (!sub(t->table_, priv_start, priv_end, priv))
				return nullptr;
			size_t subrs_off;
			if (private_subrs_offset(priv, subrs_off)) {
				// relative to the beginning of the Private DICT data
				Cursor ls(t->table_, priv_start + subrs_off);
				if (!ls.ok || !parse_index(ls, t->local_subrs_))
					return nullptr;
			}
		}
	}
	return t;
}

// ---- CFF 2 (cff2.rs): variable-font outlines at the face's variation coordinates — which the reference never sets,
// so every coordinate is 0 and the blend scalars are those of the default instance ------------------------------------
std::unique_ptr<CffTable> CffTable::parse2(const uint8_t *data, size_t len, uint16_t axis_count)
{
	std::unique_ptr<CffTable> t(new CffTable());
	t->table_.p = data, t->table_.len = len;
	t->cff2_ = true;
	t->axis_count_ = axis_count;
	Cursor c(t->table_);
	const uint8_t major = c.u8();
	c.skip(1); // minor
	const uint8_t header_size = c.u8();
	const uint16_t top_dict_length = c.u16();
	if (!c.ok || major != 2 || header_size < 5)
		return nullptr;
	c.skip(header_size - 5u);
	const Bytes top_dict = c.take(top_dict_length);
	if (!c.ok)
		return nullptr;
	size_t char_strings_off = 0, vstore_off = 0, fd_array_off = 0;
	bool has_vstore = false, has_fd_array = false;
	{
		Dict d(top_dict);
		for (int op = d.next(); op >= 0; op = d.next()) {
			switch (op) {
			case 17:
				if (!d.offset(char_strings_off))
					return nullptr;
				break;
			case 24: has_vstore = d.offset(vstore_off); break;
			case 1236: has_fd_array = d.offset(fd_array_off); break;
			default: break;
			}
		}
	}
	if (char_strings_off == 0)
		return nullptr;
	if (!parse_index(c, t->global_subrs_, true))
		return nullptr;
	{
		Cursor cs(t->table_, char_strings_off);
		if (!cs.ok || !parse_index(cs, t->char_strings_, true))
			return nullptr;
	}
	if (has_vstore) {
		// a u16 length, then an ItemVariationStore (var_store.rs): format 1, the region list, the data subtables
		Cursor vs(t->table_, vstore_off);
		vs.skip(2);
		const size_t base = vs.pos;
		const uint16_t format = vs.u16();
		const uint32_t regions_off = vs.u32();
		const uint16_t n_data = vs.u16();
		if (!vs.ok || format != 1)
			return nullptr;
		t->var_data_offsets_ = vs.take((size_t)n_data * 4);
		if (!vs.ok)
			return nullptr;
		t->var_store_ = Bytes{t->table_.p + base, t->table_.len - base};
		Cursor rs(t->var_store_, regions_off);
		t->region_axes_ = rs.u16();
		const uint16_t n_regions = rs.u16();
		t->regions_ = rs.take((size_t)t->region_axes_ * n_regions * 6);
		if (!rs.ok)
			return nullptr;
		t->n_regions_ = n_regions;
		t->has_var_store_ = true;
	}
	if (has_fd_array) {
		// (no FDSelect in ttf-parser's CFF 2: the first Font DICT whose Private DICT has local subroutines supplies them)
		Cursor fa(t->table_, fd_array_off);
		Index fonts;
		if (!fa.ok || !parse_index(fa, fonts, true))
			return nullptr;
		for (uint32_t i = 0; i < fonts.count; ++i) {
			Bytes fd;
			if (!fonts.get(i, fd))
				return nullptr;
			size_t priv_start = 0, priv_end = 0;
			bool has_priv = false;
			Dict d(fd);
			for (int op = d.next(); op >= 0; op = d.next())
				if (op == 18)
					has_priv = d.range(priv_start, priv_end);
			if (!has_priv)
				return nullptr;
			Bytes priv;
			if (!sub(t->table_, priv_start, priv_end, priv))
				return nullptr;
			size_t subrs_off;
			if (private_subrs_offset(priv, subrs_off)) {
				Cursor ls(t->table_, priv_start + subrs_off);
				if (!ls.ok || !parse_index(ls, t->local_subrs_, true))
					return nullptr;
				break;
			}
		}
	}
	return t;
}

// CharStringParserContext::update_scalars at all-zero coordinates: one scalar per region of ItemVariationData[index]
// (RegionList::evaluate_region / evaluate_axis, var_store.rs).  false = InvalidItemVariationDataIndex / too many regions.
bool CffTable::blend_scalars(uint16_t index, float *scalars, int &count) const
{
	count = 0;
	if (!has_var_store_)
		return false; // (the default store has no data subtables: every index is invalid)
	if ((size_t)index * 4 + 4 > var_data_offsets_.len)
		return false;
	const uint32_t off = read_offset(var_data_offsets_.p + (size_t)index * 4, 4);
	Cursor d(var_store_, off);
	d.skip(4); // item count, short delta count
	const uint16_t n = d.u16();
	if (!d.ok)
		return false;
	for (uint16_t k = 0; k < n; ++k) {
		const uint16_t region = d.u16();
		if (!d.ok)
			return false;
		float v = 1.0f;
		for (uint16_t axis = 0; axis < axis_count_; ++axis) {
			// RegionList::get(index, axis): None -> the region evaluates to 0
			if (region >= n_regions_ || axis >= region_axes_) {
				v = 0.0f;
				break;
			}
			const uint8_t *r = regions_.p + ((size_t)region * region_axes_ + axis) * 6;
			const int16_t start = (int16_t)((r[0] << 8) | r[1]), peak = (int16_t)((r[2] << 8) | r[3]), end = (int16_t)((r[4] << 8) | r[5]);
			// evaluate_axis(coord = 0)
			float factor;
			if (start > peak || peak > end)
				factor = 1.0f;
			else if (start < 0 && end > 0 && peak != 0)
				factor = 1.0f;
			else if (peak == 0)
				factor = 1.0f;
			else if (0 <= start || end <= 0)
				factor = 0.0f;
			else if (0 < peak)
				factor = (float)(0 - start) / (float)(peak - start);
			else
				factor = (float)(end - 0) / (float)(end - peak);
			if (factor == 0.0f) {
				v = 0.0f;
				break;
			}
			v *= factor;
		}
		if (count == 64)
			return false; // BlendRegionsLimitReached
		scalars[count++] = v;
	}
	return true;
}

// seac_code_to_glyph_id: StandardEncoding code -> SID -> glyph through the charset
bool CffTable::seac_glyph(float code, uint16_t &glyph_id) const
{
	if (!(code > -1.0f && code < 256.0f))
		return false;
	const uint8_t ch = (uint8_t)(int)code;
	const uint16_t sid = kStandardEncoding[ch];
	if (charset_kind_ == 0) { // ISOAdobe: glyph id = SID, defined up to 228
		if (ch > 228)
			return false;
		glyph_id = sid;
		return true;
	}
	if (charset_kind_ < 3)
		return false; // Expert, ExpertSubset
	if (sid == 0) {
		glyph_id = 0;
		return true;
	}
	const uint8_t *r = charset_.p;
	if (charset_kind_ == 3) {
		for (uint32_t i = 0; i < charset_records_; ++i)
			if ((uint16_t)((r[2 * i] << 8) | r[2 * i + 1]) == sid) {
				glyph_id = (uint16_t)(i + 1);
				return true;
			}
		return false;
	}
	const size_t rec = charset_kind_ == 4 ? 3 : 4;
	uint32_t gid = 1;
	for (uint32_t i = 0; i < charset_records_; ++i, r += rec) {
		const uint32_t first = (uint32_t)((r[0] << 8) | r[1]);
		const uint32_t left = rec == 3 ? r[2] : (uint32_t)((r[2] << 8) | r[3]);
		if (first <= sid && sid <= first + left) {
			glyph_id = (uint16_t)(gid + (sid - first));
			return true;
		}
		gid += left + 1;
	}
	return false;
}

// parse_cid_local_subrs: FDSelect → Font DICT → Private DICT → Subrs
bool CffTable::cid_local_subrs(uint16_t glyph_id, Index &out) const
{
	uint8_t fd = 0;
	if (fd_select_format_ == 0) {
		if (glyph_id >= fd_select_.len)
			return false;
		fd = fd_select_.p[glyph_id];
	} else {
		Cursor c(fd_select_);
		const uint16_t n_ranges = c.u16();
		if (!c.ok || n_ranges == 0 || n_ranges == 0xffff)
			return false;
		uint16_t prev_first = c.u16();
		uint8_t prev_fd = c.u8();
		bool found = false;
		for (uint32_t i = 1; i < (uint32_t)n_ranges + 1; ++i) { // the sentinel closes the last range
			const uint16_t first = c.u16();
			if (!c.ok)
				return false;
			if (glyph_id >= prev_first && glyph_id < first) {
				found = true;
				break;
			}
			prev_fd = c.u8();
			if (!c.ok)
				return false;
			prev_first = first;
		}
		if (!found)
			return false;
		fd = prev_fd;
	}
	Bytes font_dict;
	if (!fd_array_.get(fd, font_dict))
		return false;
	size_t start = 0, end = 0;
	bool have = false;
	{
		Dict d(font_dict);
		for (int op = d.next(); op >= 0; op = d.next())
			if (op == 18) {
				have = d.range(start, end);
				break;
			}
	}
	Bytes priv;
	if (!have || !sub(table_, start, end, priv))
		return false;
	size_t subrs_off;
	if (!private_subrs_offset(priv, subrs_off))
		return false;
	Cursor ls(table_, start + subrs_off);
	return ls.ok && parse_index(ls, out);
}

// ---- Type 2 charstring interpreter (cff1.rs `_parse_char_string`, charstring.rs) ---------------------
struct CffTable::Interp {
	OutlineBuilder &b;
	uint16_t glyph_id;
	float stack[kMaxArguments2];
	int max_len = kMaxArguments;
	// CFF 2: blend scalars of the current ItemVariationData, `vsindex` / `blend` seen
	bool cff2 = false, had_vsindex = false, had_blend = false;
	float scalars[64];
	int n_scalars = 0;
	int len = 0;
	float x = 0.f, y = 0.f;
	bool has_move_to = false, is_first_move_to = true;
	bool have_width = false, has_endchar = false, has_seac = false, drew = false;
	uint32_t stems = 0;
	bool have_local = false;
	Index local;

	Interp(OutlineBuilder &bb, uint16_t g) : b(bb), glyph_id(g) {}
	bool push(float v)
	{
		if (len == max_len)
			return false;
		stack[len++] = v;
		return true;
	}
	float at(int i) const { return stack[i]; }
	float pop() { return stack[--len]; }
	void width_from_first() { have_width = true; }

	void move(float nx, float ny)
	{
		if (is_first_move_to)
			is_first_move_to = false;
		else
			b.close();
		has_move_to = true;
		x = nx, y = ny;
		b.move_to(x, y);
		drew = true;
		len = 0;
	}
	void line() { b.line_to(x, y); }
	void curve(float x1, float y1, float x2, float y2, float ex, float ey)
	{
		x = ex, y = ey;
		b.curve_to(x1, y1, x2, y2, x, y);
	}
	// {dxa dya dxb dyb dxc dyc} starting at stack[i]
	void rr_curve(int i)
	{
		const float x1 = x + at(i), y1 = y + at(i + 1);
		const float x2 = x1 + at(i + 2), y2 = y1 + at(i + 3);
		curve(x1, y1, x2, y2, x2 + at(i + 4), y2 + at(i + 5));
	}

	bool moveto(int op)
	{
		const int want = op == 21 ? 2 : 1;
		int i = 0;
		if (cff2 && len != want) // CFF 2 charstrings carry no width
			return false;
		if (len == want + 1) { // one argument too many: the first is the width (also inside a seac component)
			have_width = true;
			i = 1;
		}
		if (len != i + want)
			return false;
		if (op == 21)
			move(x + at(i), y + at(i + 1));
		else if (op == 22)
			move(x + at(i), y);
		else
			move(x, y + at(i));
		return true;
	}
	bool rlineto()
	{
		if (!has_move_to || (len & 1))
			return false;
		for (int i = 0; i < len; i += 2) {
			x += at(i), y += at(i + 1);
			line();
		}
		len = 0;
		return true;
	}
	bool hvlineto(bool horizontal)
	{
		if (!has_move_to || len == 0)
			return false;
		for (int i = 0; i < len; ++i) {
			if (horizontal)
				x += at(i);
			else
				y += at(i);
			horizontal = !horizontal;
			line();
		}
		len = 0;
		return true;
	}
	bool rrcurveto()
	{
		if (!has_move_to || len % 6 != 0)
			return false;
		for (int i = 0; i < len; i += 6)
			rr_curve(i);
		len = 0;
		return true;
	}
	bool rcurveline()
	{
		if (!has_move_to || len < 8 || (len - 2) % 6 != 0)
			return false;
		int i = 0;
		for (; i < len - 2; i += 6)
			rr_curve(i);
		x += at(i), y += at(i + 1);
		line();
		len = 0;
		return true;
	}
	bool rlinecurve()
	{
		if (!has_move_to || len < 8 || ((len - 6) & 1))
			return false;
		int i = 0;
		for (; i < len - 6; i += 2) {
			x += at(i), y += at(i + 1);
			line();
		}
		rr_curve(i);
		len = 0;
		return true;
	}
	bool hhcurveto()
	{
		if (!has_move_to)
			return false;
		int i = 0;
		if (len & 1) {
			y += at(0);
			i = 1;
		}
		if ((len - i) % 4 != 0)
			return false;
		for (; i < len; i += 4) {
			const float x1 = x + at(i), y1 = y;
			const float x2 = x1 + at(i + 1), y2 = y1 + at(i + 2);
			curve(x1, y1, x2, y2, x2 + at(i + 3), y2);
		}
		len = 0;
		return true;
	}
	bool vvcurveto()
	{
		if (!has_move_to)
			return false;
		int i = 0;
		if (len & 1) {
			x += at(0);
			i = 1;
		}
		if ((len - i) % 4 != 0)
			return false;
		for (; i < len; i += 4) {
			const float x1 = x, y1 = y + at(i);
			const float x2 = x1 + at(i + 1), y2 = y1 + at(i + 2);
			curve(x1, y1, x2, y2, x2, y2 + at(i + 3));
		}
		len = 0;
		return true;
	}
	// hvcurveto / vhcurveto: curves alternate between starting horizontal and starting vertical; a single
	// left-over argument is the last curve's free end coordinate
	bool alternating(bool horizontal)
	{
		if (!has_move_to || len < 4)
			return false;
		int i = 0;
		while (i < len) {
			if (len - i < 4)
				return false;
			const bool last = len - i == 5;
			if (horizontal) {
				const float x1 = x + at(i), y1 = y;
				const float x2 = x1 + at(i + 1), y2 = y1 + at(i + 2);
				const float ey = y2 + at(i + 3);
				float ex = x2;
				if (last)
					ex += at(i + 4);
				curve(x1, y1, x2, y2, ex, ey);
			} else {
				const float x1 = x, y1 = y + at(i);
				const float x2 = x1 + at(i + 1), y2 = y1 + at(i + 2);
				const float ex = x2 + at(i + 3);
				float ey = y2;
				if (last)
					ey += at(i + 4);
				curve(x1, y1, x2, y2, ex, ey);
			}
			i += last ? 5 : 4;
			horizontal = !horizontal;
		}
		len = 0;
		return true;
	}
	bool flex(int op2)
	{
		if (!has_move_to)
			return false;
		const float sx = x, sy = y;
		if (op2 == 35) { // flex: two rrcurves, the 13th argument (flex depth) unused
			if (len != 13)
				return false;
			rr_curve(0);
			rr_curve(6);
		} else if (op2 == 34) { // hflex
			if (len != 7)
				return false;
			const float x1 = x + at(0), y1 = y;
			const float x2 = x1 + at(1), y2 = y1 + at(2);
			curve(x1, y1, x2, y2, x2 + at(3), y2);
			const float x4 = x + at(4), x5 = x4 + at(5);
			curve(x4, y2, x5, sy, x5 + at(6), sy);
		} else if (op2 == 36) { // hflex1
			if (len != 9)
				return false;
			const float x1 = x + at(0), y1 = y + at(1);
			const float x2 = x1 + at(2), y2 = y1 + at(3);
			curve(x1, y1, x2, y2, x2 + at(4), y2);
			const float x4 = x + at(5), x5 = x4 + at(6), y5 = y2 + at(7);
			curve(x4, y2, x5, y5, x5 + at(8), sy);
		} else { // flex1: the last argument moves along the dominant axis, the other returns to the start
			if (len != 11)
				return false;
			const float x1 = x + at(0), y1 = y + at(1);
			const float x2 = x1 + at(2), y2 = y1 + at(3);
			const float x3 = x2 + at(4), y3 = y2 + at(5);
			const float x4 = x3 + at(6), y4 = y3 + at(7);
			const float x5 = x4 + at(8), y5 = y4 + at(9);
			float ex = sx, ey = sy;
			if (std::fabs(x5 - sx) > std::fabs(y5 - sy))
				ex = x5 + at(10);
			else
				ey = y5 + at(10);
			curve(x1, y1, x2, y2, x3, y3);
			curve(x4, y4, x5, y5, ex, ey);
		}
		len = 0;
		return true;
	}
};

bool CffTable::run(Interp &in, Bytes code, int depth) const
{
	Cursor s(code);
	while (!s.at_end()) {
		const uint8_t op = s.u8();
		if (cff2_ && (op == 11 || op == 14))
			return false; // `return` and `endchar` do not exist in CFF 2
		if (cff2_ && op == 15) { // vsindex: once, before the first blend
			if (in.had_blend || in.had_vsindex || in.len != 1)
				return false;
			const float v = in.pop();
			if (!(v > -2147483904.f && v < 2147483648.f) || (int32_t)v < 0 || (int32_t)v > 65535)
				return false;
			if (!blend_scalars((uint16_t)(int32_t)v, in.scalars, in.n_scalars))
				return false;
			in.had_vsindex = true;
			in.len = 0;
			continue;
		}
		if (cff2_ && op == 16) { // blend: n defaults followed by n * k deltas, k = regions of the current variation data
			in.had_blend = true;
			if (in.len == 0)
				return false;
			const float nv = in.pop();
			if (!(nv > -2147483904.f && nv < 2147483648.f) || (int32_t)nv < 0 || (int32_t)nv > 65535)
				return false;
			const int n = (int32_t)nv, k = in.n_scalars;
			const int need = n * (k + 1);
			if (in.len < need)
				return false;
			const int start = in.len - need;
			for (int i = n - 1; i >= 0; --i)
				for (int j = 0; j < k; ++j) {
					const float delta = in.pop();
					in.stack[start + i] += delta * in.scalars[k - j - 1];
				}
			continue;
		}
		switch (op) {
		case 0: case 2: case 9: case 13: case 15: case 16: case 17:
			return false; // reserved
		case 1: case 3: case 18: case 23: { // stem hints: an odd count means the first value is the width
			int n = in.len;
			if ((n & 1) && !in.have_width) {
				in.have_width = true;
				n -= 1;
			}
			in.stems += (uint32_t)n >> 1;
			in.len = 0;
			break;
		}
		case 19: case 20: { // hintmask / cntrmask: implied vstem arguments, then one mask bit per stem
			int n = in.len;
			in.len = 0;
			if (n & 1) {
				in.have_width = true;
				n -= 1;
			}
			in.stems += (uint32_t)n >> 1;
			s.skip((in.stems + 7) >> 3);
			if (!s.ok)
				return false;
			break;
		}
		case 4: case 21: case 22:
			if (!in.moveto(op))
				return false;
			break;
		case 5:
			if (!in.rlineto())
				return false;
			break;
		case 6: case 7:
			if (!in.hvlineto(op == 6))
				return false;
			break;
		case 8:
			if (!in.rrcurveto())
				return false;
			break;
		case 10: case 29: { // callsubr / callgsubr
			if (in.len == 0 || depth == kStackLimit)
				return false;
			const Index *subrs = &global_subrs_;
			if (op == 10) {
				if (!in.have_local) {
					if (!cid_) {
						in.local = local_subrs_;
						in.have_local = true;
					} else if (cid_local_subrs(in.glyph_id, in.local)) {
						in.have_local = true;
					}
				}
				if (!in.have_local)
					return false;
				subrs = &in.local;
			}
			const int bias = subrs->count < 1240 ? 107 : (subrs->count < 33900 ? 1131 : 32768);
			const float v = in.pop();
			if (!(v > -2147483904.f && v < 2147483648.f))
				return false;
			const int64_t idx = (int64_t)(int32_t)v + bias;
			Bytes body;
			if (idx < 0 || !subrs->get((uint32_t)idx, body))
				return false;
			if (!run(in, body, depth + 1))
				return false;
			if (in.has_endchar && !in.has_seac) {
				if (!s.at_end())
					return false;
				return true;
			}
			break;
		}
		case 11: return true; // return
		case 12: {
			const uint8_t op2 = s.u8();
			if (!s.ok || op2 < 34 || op2 > 37)
				return false; // arithmetic / storage operators are not supported by ttf-parser either
			if (!in.flex(op2))
				return false;
			break;
		}
		case 14: { // endchar
			if (in.len == 4 || (!in.have_width && in.len == 5)) { // seac: adx ady bchar achar
				uint16_t accent, base;
				if (!seac_glyph(in.pop(), accent) || !seac_glyph(in.pop(), base))
					return false;
				const float dy = in.pop(), dx = in.pop();
				if (!in.have_width && in.len != 0) {
					in.have_width = true;
					in.len--;
				}
				in.has_seac = true;
				if (depth == kStackLimit)
					return false;
				Bytes part;
				if (!char_strings_.get(base, part) || !run(in, part, depth + 1))
					return false;
				in.x = dx, in.y = dy;
				if (!char_strings_.get(accent, part) || !run(in, part, depth + 1))
					return false;
			} else if (in.len == 1 && !in.have_width) {
				in.have_width = true;
				in.len = 0;
			}
			if (!in.is_first_move_to) {
				in.is_first_move_to = true;
				in.b.close();
			}
			if (!s.at_end())
				return false;
			in.has_endchar = true;
			return true;
		}
		case 24:
			if (!in.rcurveline())
				return false;
			break;
		case 25:
			if (!in.rlinecurve())
				return false;
			break;
		case 26:
			if (!in.vvcurveto())
				return false;
			break;
		case 27:
			if (!in.hhcurveto())
				return false;
			break;
		case 28: {
			const int16_t v = (int16_t)s.u16();
			if (!s.ok || !in.push((float)v))
				return false;
			break;
		}
		case 30: case 31:
			if (!in.alternating(op == 31))
				return false;
			break;
		case 255: {
			const int32_t v = (int32_t)s.u32();
			if (!s.ok || !in.push((float)v / 65536.0f))
				return false;
			break;
		}
		default: {
			int v;
			if (op <= 246) {
				v = (int)op - 139;
			} else {
				const int b1 = s.u8();
				if (!s.ok)
					return false;
				v = op <= 250 ? ((int)op - 247) * 256 + b1 + 108 : -((int)op - 251) * 256 - b1 - 108;
			}
			if (!in.push((float)v))
				return false;
			break;
		}
		}
	}
	return true;
}

bool CffTable::outline(uint16_t glyph_id, OutlineBuilder &builder) const
{
	Bytes code;
	if (!char_strings_.get(glyph_id, code))
		return false;
	Interp in(builder, glyph_id);
	if (cff2_) {
		in.cff2 = true;
		in.max_len = kMaxArguments2;
		in.have_width = true; // (no width anywhere: an odd stem count just drops its last value)
		if (!blend_scalars(0, in.scalars, in.n_scalars)) // "load scalars at default index"
			return false;
		if (!run(in, code, 0))
			return false;
		return in.drew; // ZeroBBox
	}
	if (!run(in, code, 0))
		return false;
	return in.has_endchar && in.drew; // MissingEndChar; ZeroBBox (the box changes with the first move_to)
}

} // namespace vgb
