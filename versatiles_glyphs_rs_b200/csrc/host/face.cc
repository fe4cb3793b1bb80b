// face.cc — sfnt / cmap / hmtx / glyf reader (behaviour of ttf-parser 0.25.1 as documented in
// SURVEY.md Appendix C; anchored on the reference's call sites, see face.h).
#include "face.h"

#include "cff.h"

#include <algorithm>
#include <cstring>

namespace vgb {

// ---- Transform (f32, like ttf-parser's) ------------------------------------------------------------
Face::Transform Face::Transform::combine(const Transform &t1, const Transform &t2)
{
	Transform r;
	r.a = t1.a * t2.a + t1.c * t2.b;
	r.b = t1.b * t2.a + t1.d * t2.b;
	r.c = t1.a * t2.c + t1.c * t2.d;
	r.d = t1.b * t2.c + t1.d * t2.d;
	r.e = t1.a * t2.e + t1.c * t2.f + t1.e;
	r.f = t1.b * t2.e + t1.d * t2.f + t1.f;
	return r;
}

void Face::Transform::apply_to(float &x, float &y) const
{
	const float tx = x, ty = y;
	x = a * tx + c * ty + e;
	y = b * tx + d * ty + f;
}

// ---- parse -----------------------------------------------------------------------------------------
bool Face::table(const char tag[4], Span &out) const
{
	if (data_.size() < 12)
		return false;
	const uint16_t n = u16(4);
	for (uint16_t i = 0; i < n; ++i) {
		const size_t rec = 12 + (size_t)i * 16;
		if (rec + 16 > data_.size())
			return false;
		if (std::memcmp(&data_[rec], tag, 4) != 0)
			continue;
		const uint32_t off = u32(rec + 8), len = u32(rec + 12);
		if ((size_t)off + len > data_.size())
			return false;
		out.off = off;
		out.len = len;
		return true;
	}
	return false;
}

Face::Face()
{
	static std::atomic<uint64_t> next{1};
	uid_ = next.fetch_add(1);
}
Face::~Face() = default;

std::unique_ptr<Face> Face::parse(std::vector<uint8_t> data)
{
	std::unique_ptr<Face> f(new Face());
	f->data_ = std::move(data);
	Span head, maxp, hhea;
	if (!f->table("head", head) || head.len < 54)
		return nullptr;
	if (!f->table("maxp", maxp) || maxp.len < 6)
		return nullptr;
	if (!f->table("hhea", hhea) || hhea.len < 36)
		return nullptr;
	f->upm_ = f->u16(head.off + 18);
	if (f->upm_ < 16 || f->upm_ > 16384)
		return nullptr; // ttf-parser head::Table::parse rejects such a head table: "Could not parse font data"
	f->loca_long_ = f->i16(head.off + 50) != 0;
	f->num_glyphs_ = f->u16(maxp.off + 4);
	f->num_hmetrics_ = f->u16(hhea.off + 34);
	f->table("hmtx", f->hmtx_);
	f->table("loca", f->loca_);
	f->table("glyf", f->glyf_);
	f->table("name", f->name_);
	Span cff;
	if (f->table("CFF ", cff))
		f->cff_ = CffTable::parse(f->data_.data() + cff.off, cff.len);
	Span cff2, fvar;
	if (f->table("CFF2", cff2)) {
		// the face's variation coordinates: one per fvar axis (at most 64), all 0 — the reference never sets any
		uint16_t axes = 0;
		if (f->table("fvar", fvar) && fvar.len >= 10)
			axes = std::min<uint16_t>(f->u16(fvar.off + 8), 64);
		f->cff2_ = CffTable::parse2(f->data_.data() + cff2.off, cff2.len, axes);
	}
	if (f->table("cmap", f->cmap_) && f->cmap_.len >= 4) {
		const uint16_t n = f->u16(f->cmap_.off + 2);
		for (uint16_t i = 0; i < n; ++i) {
			const size_t rec = f->cmap_.off + 4 + (size_t)i * 8;
			if (rec + 8 > f->cmap_.off + f->cmap_.len)
				break;
			const uint32_t off = f->u32(rec + 4);
			if ((size_t)off + 2 > f->cmap_.len)
				continue;
			CmapSubtable s;
			s.platform = f->u16(rec);
			s.encoding = f->u16(rec + 2);
			s.format = f->u16(f->cmap_.off + off);
			s.data.off = f->cmap_.off + off;
			s.data.len = f->cmap_.len - off;
			f->subtables_.push_back(s);
		}
	}
	return f;
}

// ---- cmap ------------------------------------------------------------------------------------------
std::optional<uint16_t> Face::lookup(const CmapSubtable &s, uint32_t cp) const
{
	const size_t base = s.data.off, len = s.data.len;
	switch (s.format) {
	case 0: {
		if (cp >= 256 || len < 6 + 256)
			return std::nullopt;
		const uint8_t g = data_[base + 6 + cp];
		if (g == 0)
			return std::nullopt;
		return (uint16_t)g;
	}
	case 4: {
		if (cp > 0xFFFF || len < 16)
			return std::nullopt;
		const uint32_t segx2 = u16(base + 6);
		const uint32_t nseg = segx2 / 2;
		const size_t ends = base + 14, starts = base + 16 + segx2, deltas = starts + segx2, ranges = deltas + segx2;
		if (ranges + segx2 > base + len)
			return std::nullopt;
		// first segment whose endCode >= cp
		uint32_t lo = 0, hi = nseg;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (u16(ends + 2 * mid) < cp)
				lo = mid + 1;
			else
				hi = mid;
		}
		if (lo == nseg)
			return std::nullopt;
		const uint16_t start = u16(starts + 2 * lo);
		if (start > cp)
			return std::nullopt;
		const uint16_t delta = u16(deltas + 2 * lo);
		const uint16_t range = u16(ranges + 2 * lo);
		if (range == 0)
			return (uint16_t)(cp + delta);
		if (range == 0xFFFF)
			return std::nullopt;
		const size_t pos = ranges + 2 * (size_t)lo + range + 2 * (size_t)(cp - start);
		if (pos + 2 > base + len)
			return std::nullopt;
		const uint16_t v = u16(pos);
		if (v == 0)
			return std::nullopt;
		return (uint16_t)(v + delta);
	}
	case 2: { // high-byte mapping through table (ttf-parser cmap/format2.rs; legacy CJK encodings)
		if (cp > 0xFFFF || len < 518)
			return std::nullopt;
		uint32_t max_key = 0;
		for (uint32_t k = 0; k < 256; ++k)
			max_key = std::max<uint32_t>(max_key, u16(base + 6 + 2 * k));
		const uint32_t n_sub = max_key / 8 + 1; // the subheader array is as long as the largest key says
		if (518 + (size_t)n_sub * 8 > len)
			return std::nullopt; // Subtable2::parse fails: the subtable answers nothing
		const uint32_t high = cp >> 8, low = cp & 0xFF;
		const uint32_t i = cp < 0xFF ? 0u : u16(base + 6 + 2 * high) / 8u; // "subheader 0 is for single-byte codes"
		if (i >= n_sub)
			return std::nullopt;
		const size_t sh = base + 518 + (size_t)i * 8;
		const uint32_t first = u16(sh), count = u16(sh + 2);
		const int32_t delta = (int16_t)u16(sh + 4);
		const uint32_t range_offset = u16(sh + 6);
		if (first + count > 0xFFFF) // checked_add on u16
			return std::nullopt;
		if (low < first || low >= first + count)
			return std::nullopt;
		// idRangeOffset counts from its own position to the glyphIndexArray element of firstCode
		const size_t pos = (size_t)518 + 8 * ((size_t)i + 1) - 2 + range_offset + 2 * (size_t)(low - first);
		if (pos + 2 > len)
			return std::nullopt;
		const int32_t glyph = u16(base + pos);
		if (glyph == 0)
			return std::nullopt;
		const int32_t v = (glyph + delta) % 65536; // (negative stays negative: u16::try_from fails)
		if (v < 0)
			return std::nullopt;
		return (uint16_t)v;
	}
	case 6: {
		if (len < 10)
			return std::nullopt;
		const uint32_t first = u16(base + 6), count = u16(base + 8);
		if (cp < first || cp >= first + count)
			return std::nullopt;
		const size_t pos = base + 10 + 2 * (size_t)(cp - first);
		if (pos + 2 > base + len)
			return std::nullopt;
		return u16(pos);
	}
	case 12: {
		if (len < 16)
			return std::nullopt;
		const uint32_t n = u32(base + 12);
		if (16 + (size_t)n * 12 > len)
			return std::nullopt;
		uint32_t lo = 0, hi = n;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			const size_t g = base + 16 + (size_t)mid * 12;
			const uint32_t sc = u32(g), ec = u32(g + 4);
			if (cp < sc)
				hi = mid;
			else if (cp > ec)
				lo = mid + 1;
			else {
				const uint64_t gid = (uint64_t)u32(g + 8) + (cp - sc);
				if (gid > 0xFFFF)
					return std::nullopt;
				return (uint16_t)gid;
			}
		}
		return std::nullopt;
	}
	case 10: { // trimmed array with 32-bit code points: the stored glyph id as is (ttf-parser cmap/format10.rs)
		if (len < 20)
			return std::nullopt;
		const uint32_t first = u32(base + 12), count = u32(base + 16);
		if (cp < first || cp - first >= count)
			return std::nullopt;
		const size_t pos = base + 20 + 2 * (size_t)(cp - first);
		if (pos + 2 > base + len)
			return std::nullopt;
		return u16(pos);
	}
	case 13: { // many-to-one ranges, scanned in order (ttf-parser cmap/format13.rs)
		if (len < 16)
			return std::nullopt;
		const uint32_t n = u32(base + 12);
		if (16 + (size_t)n * 12 > len)
			return std::nullopt;
		for (uint32_t i = 0; i < n; ++i) {
			const size_t g = base + 16 + (size_t)i * 12;
			if (cp >= u32(g) && cp <= u32(g + 4)) {
				const uint32_t gid = u32(g + 8);
				if (gid > 0xFFFF)
					return std::nullopt;
				return (uint16_t)gid;
			}
		}
		return std::nullopt;
	}
	default:
		return std::nullopt;
	}
}

template <typename F> void Face::enumerate(const CmapSubtable &s, F &&f) const
{
	const size_t base = s.data.off, len = s.data.len;
	switch (s.format) {
	case 0:
		for (uint32_t cp = 0; cp < 256; ++cp)
			f(cp);
		break;
	case 4: {
		if (len < 16)
			return;
		const uint32_t segx2 = u16(base + 6);
		const size_t ends = base + 14, starts = base + 16 + segx2;
		if (starts + segx2 > base + len)
			return;
		for (uint32_t i = 0; i < segx2 / 2; ++i) {
			const uint32_t sc = u16(starts + 2 * i), ec = u16(ends + 2 * i);
			if (sc == 0xFFFF && ec == 0xFFFF)
				break; // terminator segment
			for (uint32_t cp = sc; cp <= ec; ++cp)
				f(cp);
		}
		break;
	}
	case 2: { // format2.rs codepoints_inner: every code of every subheader's range; stops at the first malformed entry
		if (len < 518)
			return;
		uint32_t max_key = 0;
		for (uint32_t k = 0; k < 256; ++k)
			max_key = std::max<uint32_t>(max_key, u16(base + 6 + 2 * k));
		const uint32_t n_sub = max_key / 8 + 1;
		if (518 + (size_t)n_sub * 8 > len)
			return;
		for (uint32_t first_byte = 0; first_byte < 256; ++first_byte) {
			const uint32_t i = u16(base + 6 + 2 * first_byte) / 8u;
			if (i >= n_sub)
				return;
			const size_t sh = base + 518 + (size_t)i * 8;
			const uint32_t first = u16(sh), count = u16(sh + 2);
			if (i == 0) {
				if (first + count > 0xFFFF)
					return;
				if (first_byte >= first && first_byte < first + count)
					f(first_byte);
			} else {
				const uint32_t b = first + (first_byte << 8);
				if (b > 0xFFFF)
					return;
				for (uint32_t k = 0; k < count; ++k) {
					if (b + k > 0xFFFF)
						return;
					f(b + k);
				}
			}
		}
		break;
	}
	case 6: {
		if (len < 10)
			return;
		const uint32_t first = u16(base + 6), count = u16(base + 8);
		for (uint32_t cp = first; cp < first + count; ++cp)
			f(cp);
		break;
	}
	case 12:
	case 13: {
		if (len < 16)
			return;
		const uint32_t n = u32(base + 12);
		if (16 + (size_t)n * 12 > len)
			return;
		for (uint32_t i = 0; i < n; ++i) {
			const size_t g = base + 16 + (size_t)i * 12;
			const uint64_t sc = u32(g), ec = u32(g + 4);
			for (uint64_t cp = sc; cp <= ec; ++cp)
				f((uint32_t)cp);
		}
		break;
	}
	case 10: {
		if (len < 20)
			return;
		const uint64_t first = u32(base + 12), count = u32(base + 16);
		for (uint64_t i = 0; i < count && first + i <= 0xFFFFFFFFull; ++i)
			f((uint32_t)(first + i));
		break;
	}
	default:
		break;
	}
}

std::optional<uint16_t> Face::glyph_index(uint32_t codepoint) const
{
	for (const CmapSubtable &s : subtables_) {
		if (!s.is_unicode())
			continue;
		if (auto g = lookup(s, codepoint))
			return g;
	}
	return std::nullopt;
}

std::vector<uint32_t> Face::codepoints() const
{
	std::vector<uint32_t> cps;
	for (const CmapSubtable &s : subtables_) {
		if (!s.is_unicode())
			continue;
		enumerate(s, [&](uint32_t cp) {
			if (lookup(s, cp))
				cps.push_back(cp);
		});
	}
	std::sort(cps.begin(), cps.end());
	cps.erase(std::unique(cps.begin(), cps.end()), cps.end());
	return cps;
}

std::optional<uint16_t> Face::glyph_hor_advance(uint16_t gid) const
{
	if (gid >= num_glyphs_ || num_hmetrics_ == 0)
		return std::nullopt;
	const size_t i = std::min<size_t>(gid, (size_t)num_hmetrics_ - 1);
	if (i * 4 + 2 > hmtx_.len)
		return std::nullopt;
	return u16(hmtx_.off + i * 4);
}

std::string Face::name(uint16_t name_id) const
{
	if (name_.len < 6)
		return std::string();
	const size_t base = name_.off;
	const uint16_t count = u16(base + 2), storage = u16(base + 4);
	for (uint16_t i = 0; i < count; ++i) {
		const size_t rec = base + 6 + (size_t)i * 12;
		if (rec + 12 > base + name_.len)
			break;
		if (u16(rec + 6) != name_id)
			continue;
		const uint16_t platform = u16(rec), encoding = u16(rec + 2);
		const size_t len = u16(rec + 8), off = (size_t)storage + u16(rec + 10);
		if (off + len > name_.len)
			continue;
		const bool utf16 = platform == 0 || (platform == 3 && (encoding == 0 || encoding == 1 || encoding == 10));
		const bool roman = platform == 1 && encoding == 0;
		if (!utf16 && !roman)
			continue;
		std::string out;
		if (roman) {
			for (size_t k = 0; k < len; ++k) {
				const uint8_t ch = data_[base + off + k];
				out.push_back(ch < 0x80 ? (char)ch : '?');
			}
		} else {
			for (size_t k = 0; k + 1 < len; k += 2) {
				uint32_t cp = u16(base + off + k);
				if (cp >= 0xD800 && cp < 0xDC00 && k + 3 < len) {
					const uint32_t lo = u16(base + off + k + 2);
					if (lo >= 0xDC00 && lo < 0xE000) {
						cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
						k += 2;
					}
				}
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
		}
		return out;
	}
	return std::string();
}

// ---- glyf ------------------------------------------------------------------------------------------
bool Face::glyph_range(uint16_t gid, Span &out) const
{
	if (gid >= num_glyphs_)
		return false;
	size_t a, b;
	if (loca_long_) {
		if (((size_t)gid + 2) * 4 > loca_.len)
			return false;
		a = u32(loca_.off + (size_t)gid * 4);
		b = u32(loca_.off + (size_t)gid * 4 + 4);
	} else {
		if (((size_t)gid + 2) * 2 > loca_.len)
			return false;
		a = 2 * (size_t)u16(loca_.off + (size_t)gid * 2);
		b = 2 * (size_t)u16(loca_.off + (size_t)gid * 2 + 2);
	}
	if (b <= a || b > glyf_.len)
		return false;
	out.off = glyf_.off + a;
	out.len = b - a;
	return true;
}

// Walks the points of one simple-glyph contour and emits move/line/quad callbacks the way
// ttf-parser's glyf builder does: the contour starts at the first on-curve point (or at the
// midpoint of the first two off-curve points), consecutive off-curve points get an implied
// on-curve midpoint, and the contour is always finished by an explicit segment back to the start
// followed by close().
class Face::ContourEmitter {
  public:
	ContourEmitter(OutlineBuilder &b, const Transform &t) : b_(b), t_(t), identity_(t.is_default()) {}

	void push(float x, float y, bool on_curve, bool last)
	{
		const Pt p{x, y};
		if (!have_first_on_) {
			if (on_curve) {
				first_on_ = p;
				have_first_on_ = true;
				move_to(p);
			} else if (have_first_off_) {
				const Pt mid = lerp_half(first_off_, p);
				first_on_ = mid;
				have_first_on_ = true;
				last_off_ = p;
				have_last_off_ = true;
				move_to(mid);
			} else {
				first_off_ = p;
				have_first_off_ = true;
			}
		} else if (have_last_off_) {
			const Pt ctrl = last_off_;
			if (on_curve) {
				have_last_off_ = false;
				quad_to(ctrl, p);
			} else {
				last_off_ = p;
				quad_to(ctrl, lerp_half(ctrl, p));
			}
		} else if (on_curve) {
			line_to(p);
		} else {
			last_off_ = p;
			have_last_off_ = true;
		}
		if (last)
			finish();
	}

  private:
	struct Pt {
		float x, y;
	};
	static Pt lerp_half(Pt a, Pt b) { return Pt{a.x + 0.5f * (b.x - a.x), a.y + 0.5f * (b.y - a.y)}; }
	void xf(Pt &p) const
	{
		if (!identity_)
			t_.apply_to(p.x, p.y);
	}
	void move_to(Pt p)
	{
		xf(p);
		b_.move_to(p.x, p.y);
	}
	void line_to(Pt p)
	{
		xf(p);
		b_.line_to(p.x, p.y);
	}
	void quad_to(Pt c, Pt p)
	{
		xf(c);
		xf(p);
		b_.quad_to(c.x, c.y, p.x, p.y);
	}
	void finish()
	{
		if (have_first_off_ && have_last_off_) {
			const Pt ctrl = last_off_;
			have_last_off_ = false;
			quad_to(ctrl, lerp_half(ctrl, first_off_));
		}
		if (have_first_on_ && have_first_off_)
			quad_to(first_off_, first_on_);
		else if (have_first_on_ && have_last_off_)
			quad_to(last_off_, first_on_);
		else if (have_first_on_)
			line_to(first_on_);
		have_first_on_ = have_first_off_ = have_last_off_ = false;
		b_.close();
	}

	OutlineBuilder &b_;
	const Transform &t_;
	bool identity_;
	bool have_first_on_ = false, have_first_off_ = false, have_last_off_ = false;
	Pt first_on_{0, 0}, first_off_{0, 0}, last_off_{0, 0};
};

void Face::outline_impl(Span g, int depth, const Transform &t, OutlineBuilder &builder) const
{
	if (depth >= 32 || g.len < 10)
		return;
	const size_t end = g.off + g.len;
	const int16_t n_contours = i16(g.off);
	size_t pos = g.off + 10;
	if (n_contours > 0) {
		const size_t nc = (size_t)n_contours;
		if (pos + 2 * nc + 2 > end)
			return;
		const size_t end_pts = pos;
		const uint32_t n_points = (uint32_t)u16(end_pts + 2 * (nc - 1)) + 1;
		if (n_points == 1)
			return; // a single point is not an outline
		pos += 2 * nc;
		const uint16_t instr_len = u16(pos);
		pos += 2 + (size_t)instr_len;
		if (pos > end)
			return;
		// expand flags (REPEAT 0x08)
		thread_local std::vector<uint8_t> flags; // scratch: no allocation per glyph (simple glyphs do not recurse)
		flags.resize(n_points);
		for (uint32_t k = 0; k < n_points;) {
			if (pos >= end)
				return;
			const uint8_t fl = data_[pos++];
			flags[k++] = fl;
			if (fl & 0x08) {
				if (pos >= end)
					return;
				uint8_t rep = data_[pos++];
				while (rep-- && k < n_points)
					flags[k++] = fl;
			}
		}
		size_t x_bytes = 0;
		for (uint32_t k = 0; k < n_points; ++k)
			x_bytes += (flags[k] & 0x02) ? 1 : ((flags[k] & 0x10) ? 0 : 2);
		size_t xpos = pos, ypos = pos + x_bytes;
		if (ypos > end)
			return;
		ContourEmitter emit(builder, t);
		int16_t x = 0, y = 0;
		size_t contour = 0;
		uint32_t contour_end = u16(end_pts);
		for (uint32_t k = 0; k < n_points; ++k) {
			const uint8_t fl = flags[k];
			if (fl & 0x02) { // x is one byte, sign in 0x10
				if (xpos >= end)
					return;
				const int16_t d = data_[xpos++];
				x = (int16_t)(x + ((fl & 0x10) ? d : -d));
			} else if (!(fl & 0x10)) {
				if (xpos + 2 > end)
					return;
				x = (int16_t)(x + i16(xpos));
				xpos += 2;
			}
			if (fl & 0x04) {
				if (ypos >= end)
					return;
				const int16_t d = data_[ypos++];
				y = (int16_t)(y + ((fl & 0x20) ? d : -d));
			} else if (!(fl & 0x20)) {
				if (ypos + 2 > end)
					return;
				y = (int16_t)(y + i16(ypos));
				ypos += 2;
			}
			const bool last = (k == contour_end);
			emit.push((float)x, (float)y, (fl & 0x01) != 0, last);
			if (last && ++contour < nc)
				contour_end = u16(end_pts + 2 * contour);
		}
	} else if (n_contours < 0) {
		enum : uint16_t {
			ARG_WORDS = 0x0001,
			ARGS_ARE_XY = 0x0002,
			HAVE_SCALE = 0x0008,
			MORE_COMPONENTS = 0x0020,
			HAVE_XY_SCALE = 0x0040,
			HAVE_2X2 = 0x0080
		};
		for (;;) {
			if (pos + 4 > end)
				return;
			const uint16_t fl = u16(pos), child = u16(pos + 2);
			pos += 4;
			Transform ct;
			if (fl & ARG_WORDS) {
				if (pos + 4 > end)
					return;
				if (fl & ARGS_ARE_XY) {
					ct.e = (float)i16(pos);
					ct.f = (float)i16(pos + 2);
				}
				pos += 4;
			} else {
				if (pos + 2 > end)
					return;
				if (fl & ARGS_ARE_XY) {
					ct.e = (float)(int8_t)data_[pos];
					ct.f = (float)(int8_t)data_[pos + 1];
				}
				pos += 2;
			}
			auto f2dot14 = [&](size_t p) { return (float)i16(p) / 16384.0f; };
			if (fl & HAVE_2X2) {
				if (pos + 8 > end)
					return;
				ct.a = f2dot14(pos);
				ct.b = f2dot14(pos + 2);
				ct.c = f2dot14(pos + 4);
				ct.d = f2dot14(pos + 6);
				pos += 8;
			} else if (fl & HAVE_XY_SCALE) {
				if (pos + 4 > end)
					return;
				ct.a = f2dot14(pos);
				ct.d = f2dot14(pos + 2);
				pos += 4;
			} else if (fl & HAVE_SCALE) {
				if (pos + 2 > end)
					return;
				ct.a = f2dot14(pos);
				ct.d = ct.a;
				pos += 2;
			}
			Span cg;
			if (glyph_range(child, cg))
				outline_impl(cg, depth + 1, Transform::combine(t, ct), builder);
			if (!(fl & MORE_COMPONENTS))
				break;
		}
	}
}

// outline_impl's walk with the simple-glyph arm replaced by "remember the record": returns false as soon as a
// component carries anything but a translation (the caller then records the whole glyph on the host).
bool Face::parts_impl(Span g, int depth, const Transform &t, std::vector<GlyfPart> &parts) const
{
	if (depth >= 32 || g.len < 10)
		return true;
	const size_t end = g.off + g.len;
	const int16_t n_contours = i16(g.off);
	size_t pos = g.off + 10;
	if (n_contours > 0) {
		const size_t nc = (size_t)n_contours;
		if (pos + 2 * nc + 2 > end)
			return true; // outline_impl: no callbacks
		const uint32_t n_points = (uint32_t)u16(pos + 2 * (nc - 1)) + 1;
		if (n_points == 1)
			return true;
		if (n_points > 2048) // B200SDF_GLYF_MAX_POINTS
			return false;
		GlyfPart p;
		p.off = (uint32_t)(g.off - glyf_.off);
		p.len = (uint32_t)g.len;
		p.ox = t.e, p.oy = t.f;
		p.points = n_points;
		p.xmin = i16(g.off + 2), p.ymin = i16(g.off + 4), p.xmax = i16(g.off + 6), p.ymax = i16(g.off + 8);
		parts.push_back(p);
	} else if (n_contours < 0) {
		enum : uint16_t { ARG_WORDS = 0x0001, ARGS_ARE_XY = 0x0002, HAVE_SCALE = 0x0008, MORE_COMPONENTS = 0x0020, HAVE_XY_SCALE = 0x0040, HAVE_2X2 = 0x0080 };
		for (;;) {
			if (pos + 4 > end)
				return true;
			const uint16_t fl = u16(pos), child = u16(pos + 2);
			pos += 4;
			Transform ct;
			if (fl & ARG_WORDS) {
				if (pos + 4 > end)
					return true;
				if (fl & ARGS_ARE_XY) {
					ct.e = (float)i16(pos);
					ct.f = (float)i16(pos + 2);
				}
				pos += 4;
			} else {
				if (pos + 2 > end)
					return true;
				if (fl & ARGS_ARE_XY) {
					ct.e = (float)(int8_t)data_[pos];
					ct.f = (float)(int8_t)data_[pos + 1];
				}
				pos += 2;
			}
			auto f2dot14 = [&](size_t q) { return (float)i16(q) / 16384.0f; };
			if (fl & HAVE_2X2) {
				if (pos + 8 > end)
					return true;
				ct.a = f2dot14(pos), ct.b = f2dot14(pos + 2), ct.c = f2dot14(pos + 4), ct.d = f2dot14(pos + 6);
				pos += 8;
			} else if (fl & HAVE_XY_SCALE) {
				if (pos + 4 > end)
					return true;
				ct.a = f2dot14(pos), ct.d = f2dot14(pos + 2);
				pos += 4;
			} else if (fl & HAVE_SCALE) {
				if (pos + 2 > end)
					return true;
				ct.a = f2dot14(pos);
				ct.d = ct.a;
				pos += 2;
			}
			Span cg;
			if (glyph_range(child, cg)) {
				const Transform ts = Transform::combine(t, ct);
				if (!(ts.a == 1.f && ts.b == 0.f && ts.c == 0.f && ts.d == 1.f))
					return false;
				if (!parts_impl(cg, depth + 1, ts, parts))
					return false;
			}
			if (!(fl & MORE_COMPONENTS))
				break;
		}
	}
	return true;
}

Face::GlyfPlan Face::glyf_parts(uint16_t gid, std::vector<GlyfPart> &parts) const
{
	if (glyf_.len == 0 || loca_.len == 0)
		return cff_ || cff2_ ? GlyfPlan::Host : GlyfPlan::None; // outline_glyph's order: glyf, then CFF, then CFF2
	Span g;
	if (!glyph_range(gid, g))
		return GlyfPlan::None;
	const size_t before = parts.size();
	if (!parts_impl(g, 0, Transform(), parts)) {
		parts.resize(before);
		return GlyfPlan::Host;
	}
	return parts.size() == before ? GlyfPlan::None : GlyfPlan::Parts;
}

bool Face::outline_glyph(uint16_t gid, OutlineBuilder &builder) const
{
	// ttf-parser's order (lib.rs, Face::outline_glyph): a face with glyf + loca never looks at `CFF `
	if (glyf_.len == 0 || loca_.len == 0)
		return cff_ ? cff_->outline(gid, builder) : (cff2_ ? cff2_->outline(gid, builder) : false);
	Span g;
	if (!glyph_range(gid, g))
		return false;
	outline_impl(g, 0, Transform(), builder);
	return true;
}

} // namespace vgb
