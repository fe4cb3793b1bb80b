/*
 * vg_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see vg_oracle.h).
 *
 * A literal f64 restatement of the reference's per-glyph SDF path.  Every function cites the
 * reference file:line (relative to the versatiles-glyphs-rs v0.9.1 tree) it follows.  The
 * ttf-parser 0.25.1 / rstar 0.13.0 / prost 0.14.4 crates are not vendored in the reference, so
 * their published behaviour is restated from SURVEY.md Appendix C and anchored on the
 * reference's own call sites and golden tests (tests/test_oracle_goldens.py).
 *
 * Build: -O3 -ffp-contract=off (no FMA contraction: the reference's f64 op order is the spec).
 */
#define _POSIX_C_SOURCE 200809L
#include "vg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* render/mod.rs:52-68 */
#define GLYPH_SIZE 24
#define BUFFER 3
#define SDF_RADIUS 8.0
#define CUTOFF (0.25 * 256.0)
#define PRECISION 0.01 /* render/ring_builder.rs:62 (tolerance_sq, font units squared) */

/* ------------------------------------------------------------------------------------------ */
/* small growable arrays                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	double *v;
	size_t n, cap;
} dvec;

static void dvec_push2(dvec *d, double a, double b)
{
	if (d->n + 2 > d->cap) {
		d->cap = d->cap ? d->cap * 2 : 256;
		d->v = (double *)realloc(d->v, d->cap * sizeof(double));
	}
	d->v[d->n++] = a;
	d->v[d->n++] = b;
}

typedef struct {
	uint32_t *v;
	size_t n, cap;
} uvec;

static void uvec_push(uvec *d, uint32_t a)
{
	if (d->n + 1 > d->cap) {
		d->cap = d->cap ? d->cap * 2 : 64;
		d->v = (uint32_t *)realloc(d->v, d->cap * sizeof(uint32_t));
	}
	d->v[d->n++] = a;
}

typedef struct {
	uint8_t *v;
	size_t n, cap;
} bvec;

static void bvec_reserve(bvec *b, size_t extra)
{
	if (b->n + extra > b->cap) {
		size_t c = b->cap ? b->cap * 2 : 1024;
		while (c < b->n + extra)
			c *= 2;
		b->v = (uint8_t *)realloc(b->v, c);
		b->cap = c;
	}
}
static void bvec_put(bvec *b, const void *p, size_t n)
{
	bvec_reserve(b, n);
	memcpy(b->v + b->n, p, n);
	b->n += n;
}
static void bvec_byte(bvec *b, uint8_t x) { bvec_put(b, &x, 1); }

void vgo_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* Font: sfnt directory + cmap + hmtx + glyf (restated ttf-parser subset)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	const uint8_t *p;
	size_t len;
} span;

typedef struct {
	uint16_t platform, encoding, format;
	span data;
} cmap_sub;

typedef struct vgo_cff vgo_cff;
static vgo_cff *cff_parse(span t);
static vgo_cff *cff2_parse(span t, uint32_t axes);

struct vgo_font {
	uint8_t *data;
	size_t len;
	uint16_t upm;
	int16_t loca_long;
	uint16_t num_glyphs;
	uint16_t num_hmetrics;
	span hmtx, loca, glyf, cmap;
	cmap_sub *subs;
	uint32_t n_subs;
	struct vgo_cff *cff;  /* parsed `CFF ` table, NULL if absent or rejected */
	struct vgo_cff *cff2; /* parsed `CFF2` table */
};

static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }
static inline int16_t rds16(const uint8_t *p) { return (int16_t)rd16(p); }
static inline uint32_t rd32(const uint8_t *p)
{
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

static int find_table(const vgo_font *f, const char *tag, span *out)
{
	if (f->len < 12)
		return 0;
	uint16_t n = rd16(f->data + 4);
	for (uint16_t i = 0; i < n; i++) {
		size_t rec = 12 + (size_t)i * 16;
		if (rec + 16 > f->len)
			return 0;
		if (memcmp(f->data + rec, tag, 4) == 0) {
			uint32_t off = rd32(f->data + rec + 8), len = rd32(f->data + rec + 12);
			if ((size_t)off + len > f->len)
				return 0;
			out->p = f->data + off;
			out->len = len;
			return 1;
		}
	}
	return 0;
}

/* cmap::Subtable::is_unicode (Appendix C: platform 0 | (3,1) | (3,10) with format 12/13) */
static int sub_is_unicode(const cmap_sub *s)
{
	if (s->platform == 0)
		return 1;
	if (s->platform == 3 && s->encoding == 1)
		return 1;
	if (s->platform == 3 && s->encoding == 10 && (s->format == 12 || s->format == 13))
		return 1;
	return 0;
}

/* Face::parse(data, 0) — src/font/file_entry.rs:48 */
vgo_font *vgo_font_parse(const uint8_t *data, size_t len)
{
	vgo_font *f = (vgo_font *)calloc(1, sizeof(*f));
	f->data = (uint8_t *)malloc(len ? len : 1);
	memcpy(f->data, data, len);
	f->len = len;
	span head, maxp, hhea;
	if (len < 12 || !find_table(f, "head", &head) || head.len < 54 || !find_table(f, "maxp", &maxp) || maxp.len < 6 ||
	    !find_table(f, "hhea", &hhea) || hhea.len < 36 || !find_table(f, "cmap", &f->cmap)) {
		vgo_font_free(f);
		return NULL;
	}
	f->upm = rd16(head.p + 18);
	f->loca_long = rds16(head.p + 50);
	f->num_glyphs = rd16(maxp.p + 4);
	f->num_hmetrics = rd16(hhea.p + 34);
	if (!find_table(f, "hmtx", &f->hmtx))
		f->hmtx.len = 0;
	if (!find_table(f, "loca", &f->loca))
		f->loca.len = 0;
	if (!find_table(f, "glyf", &f->glyf))
		f->glyf.len = 0;
	span cff;
	if (find_table(f, "CFF ", &cff))
		f->cff = cff_parse(cff);
	span cff2, fvar;
	if (find_table(f, "CFF2", &cff2)) {
		uint32_t axes = 0;
		if (find_table(f, "fvar", &fvar) && fvar.len >= 10)
			axes = rd16(fvar.p + 8) < 64 ? rd16(fvar.p + 8) : 64;
		f->cff2 = cff2_parse(cff2, axes);
	}
	/* cmap subtable records, in table order */
	if (f->cmap.len >= 4) {
		uint16_t n = rd16(f->cmap.p + 2);
		f->subs = (cmap_sub *)calloc(n ? n : 1, sizeof(cmap_sub));
		for (uint16_t i = 0; i < n; i++) {
			size_t rec = 4 + (size_t)i * 8;
			if (rec + 8 > f->cmap.len)
				break;
			uint32_t off = rd32(f->cmap.p + rec + 4);
			if ((size_t)off + 2 > f->cmap.len)
				continue;
			cmap_sub *s = &f->subs[f->n_subs++];
			s->platform = rd16(f->cmap.p + rec);
			s->encoding = rd16(f->cmap.p + rec + 2);
			s->format = rd16(f->cmap.p + off);
			s->data.p = f->cmap.p + off;
			s->data.len = f->cmap.len - off;
		}
	}
	return f;
}

void vgo_font_free(vgo_font *f)
{
	if (!f)
		return;
	free(f->subs);
	free(f->cff);
	free(f->cff2);
	free(f->data);
	free(f);
}

uint32_t vgo_font_units_per_em(const vgo_font *f) { return f->upm; }
uint32_t vgo_font_num_glyphs(const vgo_font *f) { return f->num_glyphs; }

/* cmap format 4 lookup (Appendix C) */
static int32_t fmt4_lookup(span d, uint32_t cp)
{
	if (cp > 0xFFFF || d.len < 16)
		return -1;
	uint16_t segx2 = rd16(d.p + 6);
	uint32_t nseg = segx2 / 2;
	size_t end_off = 14, start_off = 16 + (size_t)segx2, delta_off = start_off + segx2, range_off = delta_off + segx2;
	if (range_off + segx2 > d.len)
		return -1;
	/* binary search: first segment with end >= cp */
	uint32_t lo = 0, hi = nseg;
	while (lo < hi) {
		uint32_t mid = (lo + hi) / 2;
		if (rd16(d.p + end_off + 2 * mid) < cp)
			lo = mid + 1;
		else
			hi = mid;
	}
	if (lo >= nseg)
		return -1;
	uint16_t start = rd16(d.p + start_off + 2 * lo);
	if (start > cp)
		return -1;
	uint16_t delta = rd16(d.p + delta_off + 2 * lo);
	uint16_t ro = rd16(d.p + range_off + 2 * lo);
	if (ro == 0)
		return (int32_t)((cp + delta) & 0xFFFF);
	if (ro == 0xFFFF)
		return -1;
	size_t pos = range_off + 2 * (size_t)lo + ro + 2 * (size_t)(cp - start);
	if (pos + 2 > d.len)
		return -1;
	uint16_t v = rd16(d.p + pos);
	if (v == 0)
		return -1;
	return (int32_t)((v + delta) & 0xFFFF);
}

/* cmap format 12 lookup (Appendix C) */
static int32_t fmt12_lookup(span d, uint32_t cp)
{
	if (d.len < 16)
		return -1;
	uint32_t n = rd32(d.p + 12);
	if (16 + (size_t)n * 12 > d.len)
		return -1;
	uint32_t lo = 0, hi = n;
	while (lo < hi) {
		uint32_t mid = (lo + hi) / 2;
		const uint8_t *g = d.p + 16 + (size_t)mid * 12;
		uint32_t s = rd32(g), e = rd32(g + 4);
		if (cp < s)
			hi = mid;
		else if (cp > e)
			lo = mid + 1;
		else {
			uint32_t gid = rd32(g + 8) + (cp - s);
			if (gid > 0xFFFF)
				return -1;
			return (int32_t)gid;
		}
	}
	return -1;
}

/* cmap format 2 (ttf-parser 0.25.1 cmap/format2.rs): number of subheaders, 0 = the subtable does not parse */
static uint32_t fmt2_subheaders(span d)
{
	if (d.len < 518)
		return 0;
	uint32_t max_key = 0;
	for (uint32_t k = 0; k < 256; k++) {
		uint32_t key = rd16(d.p + 6 + 2 * k);
		if (key > max_key)
			max_key = key;
	}
	uint32_t n = max_key / 8 + 1;
	return 518 + (size_t)n * 8 > d.len ? 0 : n;
}

static int32_t fmt2_lookup(span d, uint32_t cp)
{
	uint32_t n_sub = fmt2_subheaders(d);
	if (cp > 0xFFFF || n_sub == 0)
		return -1;
	uint32_t high = cp >> 8, low = cp & 0xFF;
	uint32_t i = cp < 0xFF ? 0 : rd16(d.p + 6 + 2 * high) / 8u;
	if (i >= n_sub)
		return -1;
	const uint8_t *sh = d.p + 518 + (size_t)i * 8;
	uint32_t first = rd16(sh), count = rd16(sh + 2), range_offset = rd16(sh + 6);
	int32_t delta = rds16(sh + 4);
	if (first + count > 0xFFFF || low < first || low >= first + count)
		return -1;
	size_t pos = (size_t)518 + 8 * ((size_t)i + 1) - 2 + range_offset + 2 * (size_t)(low - first);
	if (pos + 2 > d.len)
		return -1;
	int32_t glyph = rd16(d.p + pos);
	if (glyph == 0)
		return -1;
	int32_t v = (glyph + delta) % 65536;
	return v < 0 ? -1 : v;
}

static int32_t sub_lookup(const cmap_sub *s, uint32_t cp)
{
	switch (s->format) {
	case 2:
		return fmt2_lookup(s->data, cp);
	case 4:
		return fmt4_lookup(s->data, cp);
	case 12:
		return fmt12_lookup(s->data, cp);
	case 0:
		if (cp < 256 && s->data.len >= 6 + 256) {
			uint8_t g = s->data.p[6 + cp];
			return g ? (int32_t)g : -1;
		}
		return -1;
	case 6: {
		if (s->data.len < 10)
			return -1;
		uint16_t first = rd16(s->data.p + 6), cnt = rd16(s->data.p + 8);
		if (cp < first || cp >= (uint32_t)first + cnt || 10 + 2 * (size_t)(cp - first) + 2 > s->data.len)
			return -1;
		return (int32_t)rd16(s->data.p + 10 + 2 * (cp - first));
	}
	case 10: { /* trimmed array, 32-bit: glyphs[cp - first] as stored (ttf-parser 0.25.1 cmap/format10.rs) */
		if (s->data.len < 20)
			return -1;
		uint32_t first = rd32(s->data.p + 12), cnt = rd32(s->data.p + 16);
		if (cp < first || cp - first >= cnt || 20 + 2 * (size_t)(cp - first) + 2 > s->data.len)
			return -1;
		return (int32_t)rd16(s->data.p + 20 + 2 * (size_t)(cp - first));
	}
	case 13: { /* many-to-one ranges: every cp of a group maps to the group's glyph (cmap/format13.rs) */
		if (s->data.len < 16)
			return -1;
		uint32_t n = rd32(s->data.p + 12);
		if (16 + (size_t)n * 12 > s->data.len)
			return -1;
		for (uint32_t i = 0; i < n; i++) {
			const uint8_t *g = s->data.p + 16 + (size_t)i * 12;
			if (cp >= rd32(g) && cp <= rd32(g + 4)) {
				uint32_t gid = rd32(g + 8);
				return gid > 0xFFFF ? -1 : (int32_t)gid;
			}
		}
		return -1;
	}
	default:
		return -1; /* formats 2, 8 and 14: no fixture and never a unicode lookup in practice; contribute nothing */
	}
}

/* Face::glyph_index — src/render/renderer.rs:106: first unicode subtable that maps cp */
int32_t vgo_font_glyph_index(const vgo_font *f, uint32_t cp)
{
	for (uint32_t i = 0; i < f->n_subs; i++) {
		if (!sub_is_unicode(&f->subs[i]))
			continue;
		int32_t g = sub_lookup(&f->subs[i], cp);
		if (g >= 0)
			return g;
	}
	return -1;
}

/* Face::glyph_hor_advance — src/render/renderer.rs:115 */
int32_t vgo_font_hor_advance(const vgo_font *f, uint32_t gid)
{
	if (gid >= f->num_glyphs || f->num_hmetrics == 0)
		return -1;
	uint32_t i = gid < f->num_hmetrics ? gid : (uint32_t)f->num_hmetrics - 1;
	if ((size_t)i * 4 + 2 > f->hmtx.len)
		return -1;
	return rd16(f->hmtx.p + (size_t)i * 4);
}

static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
	return x < y ? -1 : x > y;
}

/* FontMetadata::try_from codepoint set — src/font/metadata.rs:104-118 */
size_t vgo_font_codepoints(const vgo_font *f, uint32_t *out, size_t cap)
{
	uvec cps = {0};
	for (uint32_t si = 0; si < f->n_subs; si++) {
		const cmap_sub *s = &f->subs[si];
		if (!sub_is_unicode(s))
			continue;
		if (s->format == 4 && s->data.len >= 16) {
			uint16_t segx2 = rd16(s->data.p + 6);
			uint32_t nseg = segx2 / 2;
			size_t end_off = 14, start_off = 16 + (size_t)segx2;
			if (start_off + segx2 > s->data.len)
				continue;
			for (uint32_t i = 0; i < nseg; i++) {
				uint32_t st = rd16(s->data.p + start_off + 2 * i), en = rd16(s->data.p + end_off + 2 * i);
				if (st == 0xFFFF && en == 0xFFFF)
					break;
				for (uint32_t cp = st; cp <= en; cp++)
					if (sub_lookup(s, cp) >= 0)
						uvec_push(&cps, cp);
			}
		} else if ((s->format == 12 || s->format == 13) && s->data.len >= 16) {
			uint32_t n = rd32(s->data.p + 12);
			if (16 + (size_t)n * 12 > s->data.len)
				continue;
			for (uint32_t i = 0; i < n; i++) {
				const uint8_t *g = s->data.p + 16 + (size_t)i * 12;
				uint32_t st = rd32(g), en = rd32(g + 4);
				for (uint32_t cp = st; cp <= en && cp >= st; cp++)
					if (sub_lookup(s, cp) >= 0)
						uvec_push(&cps, cp);
			}
		} else if (s->format == 2) {
			uint32_t n_sub = fmt2_subheaders(s->data);
			int stop = n_sub == 0;
			for (uint32_t fb = 0; fb < 256 && !stop; fb++) {
				uint32_t i = rd16(s->data.p + 6 + 2 * fb) / 8u;
				if (i >= n_sub)
					break;
				const uint8_t *sh = s->data.p + 518 + (size_t)i * 8;
				uint32_t first = rd16(sh), count = rd16(sh + 2);
				if (i == 0) {
					if (first + count > 0xFFFF)
						break;
					if (fb >= first && fb < first + count && sub_lookup(s, fb) >= 0)
						uvec_push(&cps, fb);
				} else {
					uint32_t b = first + (fb << 8);
					if (b > 0xFFFF)
						break;
					for (uint32_t k = 0; k < count; k++) {
						if (b + k > 0xFFFF) {
							stop = 1;
							break;
						}
						if (sub_lookup(s, b + k) >= 0)
							uvec_push(&cps, b + k);
					}
				}
			}
		} else if (s->format == 0) {
			for (uint32_t cp = 0; cp < 256; cp++)
				if (sub_lookup(s, cp) >= 0)
					uvec_push(&cps, cp);
		} else if (s->format == 6 && s->data.len >= 10) {
			uint32_t first = rd16(s->data.p + 6), cnt = rd16(s->data.p + 8);
			for (uint32_t cp = first; cp < first + cnt; cp++)
				if (sub_lookup(s, cp) >= 0)
					uvec_push(&cps, cp);
		} else if (s->format == 10 && s->data.len >= 20) {
			uint32_t first = rd32(s->data.p + 12), cnt = rd32(s->data.p + 16);
			for (uint32_t i = 0; i < cnt && first + i >= first; i++)
				if (sub_lookup(s, first + i) >= 0)
					uvec_push(&cps, first + i);
		}
	}
	size_t n = 0;
	if (cps.n) {
		qsort(cps.v, cps.n, sizeof(uint32_t), cmp_u32);
		for (size_t i = 0; i < cps.n; i++)
			if (i == 0 || cps.v[i] != cps.v[i - 1]) {
				if (out && n < cap)
					out[n] = cps.v[i];
				n++;
			}
	}
	free(cps.v);
	return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Geometry: Ring / flattening (src/geometry/ring.rs, point.rs)                                */
/* ------------------------------------------------------------------------------------------ */

/* Ring::add_quadratic_bezier — src/geometry/ring.rs:119-144 (explicit stack, right half pushed first) */
static void flatten_quad(dvec *ring, double sx, double sy, double cx, double cy, double ex, double ey, double tol_sq)
{
	double stack[64][6];
	int sp = 0;
	stack[sp][0] = sx, stack[sp][1] = sy, stack[sp][2] = cx, stack[sp][3] = cy, stack[sp][4] = ex, stack[sp][5] = ey;
	sp++;
	while (sp > 0) {
		sp--;
		double s0 = stack[sp][0], s1 = stack[sp][1], c0 = stack[sp][2], c1 = stack[sp][3], e0 = stack[sp][4],
		       e1 = stack[sp][5];
		double dx = s0 + e0 - c0 * 2.0;
		double dy = s1 + e1 - c1 * 2.0;
		if (dx * dx + dy * dy <= tol_sq || sp + 2 > 64) {
			dvec_push2(ring, e0, e1);
			continue;
		}
		/* Point::midpoint = (a + b) / 2.0 — src/geometry/point.rs:29-31 */
		double m1x = (s0 + c0) / 2.0, m1y = (s1 + c1) / 2.0;
		double m2x = (c0 + e0) / 2.0, m2y = (c1 + e1) / 2.0;
		double mx = (m1x + m2x) / 2.0, my = (m1y + m2y) / 2.0;
		stack[sp][0] = mx, stack[sp][1] = my, stack[sp][2] = m2x, stack[sp][3] = m2y, stack[sp][4] = e0, stack[sp][5] = e1;
		sp++;
		stack[sp][0] = s0, stack[sp][1] = s1, stack[sp][2] = m1x, stack[sp][3] = m1y, stack[sp][4] = mx, stack[sp][5] = my;
		sp++;
	}
}

/* Ring::add_cubic_bezier — src/geometry/ring.rs:159-187 */
static void flatten_cubic(dvec *ring, const double s[2], const double c1[2], const double c2[2], const double e[2],
                          double tol_sq)
{
	double stack[64][8];
	int sp = 0;
	double *t = stack[sp++];
	t[0] = s[0], t[1] = s[1], t[2] = c1[0], t[3] = c1[1], t[4] = c2[0], t[5] = c2[1], t[6] = e[0], t[7] = e[1];
	while (sp > 0) {
		sp--;
		double a[8];
		memcpy(a, stack[sp], sizeof(a));
		double dx = (a[4] + a[2]) - (a[0] + a[6]);
		double dy = (a[5] + a[3]) - (a[1] + a[7]);
		if (dx * dx + dy * dy <= tol_sq || sp + 2 > 64) {
			dvec_push2(ring, a[6], a[7]);
			continue;
		}
		double p01x = (a[0] + a[2]) / 2.0, p01y = (a[1] + a[3]) / 2.0;
		double p12x = (a[2] + a[4]) / 2.0, p12y = (a[3] + a[5]) / 2.0;
		double p23x = (a[4] + a[6]) / 2.0, p23y = (a[5] + a[7]) / 2.0;
		double p012x = (p01x + p12x) / 2.0, p012y = (p01y + p12y) / 2.0;
		double p123x = (p12x + p23x) / 2.0, p123y = (p12y + p23y) / 2.0;
		double mx = (p012x + p123x) / 2.0, my = (p012y + p123y) / 2.0;
		double *r = stack[sp++];
		r[0] = mx, r[1] = my, r[2] = p123x, r[3] = p123y, r[4] = p23x, r[5] = p23y, r[6] = a[6], r[7] = a[7];
		double *l = stack[sp++];
		l[0] = a[0], l[1] = a[1], l[2] = p01x, l[3] = p01y, l[4] = p012x, l[5] = p012y, l[6] = mx, l[7] = my;
	}
}

size_t vgo_flatten_quad(const double s[2], const double c[2], const double e[2], double tol_sq, double *out_xy, size_t cap)
{
	dvec r = {0};
	flatten_quad(&r, s[0], s[1], c[0], c[1], e[0], e[1], tol_sq);
	size_t n = r.n / 2;
	for (size_t i = 0; i < r.n && i < 2 * cap; i++)
		out_xy[i] = r.v[i];
	free(r.v);
	return n;
}

size_t vgo_flatten_cubic(const double s[2], const double c1[2], const double c2[2], const double e[2], double tol_sq,
                         double *out_xy, size_t cap)
{
	dvec r = {0};
	flatten_cubic(&r, s, c1, c2, e, tol_sq);
	size_t n = r.n / 2;
	for (size_t i = 0; i < r.n && i < 2 * cap; i++)
		out_xy[i] = r.v[i];
	free(r.v);
	return n;
}

/* RingBuilder — src/render/ring_builder.rs:8-117 */
typedef struct {
	dvec pts;      /* all saved rings, concatenated */
	uvec starts;   /* ring start offsets (points) */
	dvec ring;     /* ring under construction */
} ring_builder;

/* save_ring — ring_builder.rs:33-54, Ring::close — geometry/ring.rs:53-63 */
static void rb_save_ring(ring_builder *b)
{
	size_t n = b->ring.n / 2;
	if (n < 3) {
		b->ring.n = 0;
		return;
	}
	double fx = b->ring.v[0], fy = b->ring.v[1];
	double lx = b->ring.v[2 * n - 2], ly = b->ring.v[2 * n - 1];
	if (fabs(fx - lx) > 2.220446049250313e-16 || fabs(fy - ly) > 2.220446049250313e-16) {
		dvec_push2(&b->ring, fx, fy);
		n++;
	}
	if (n < 4) {
		b->ring.n = 0;
		return;
	}
	uvec_push(&b->starts, (uint32_t)(b->pts.n / 2));
	for (size_t i = 0; i < n; i++)
		dvec_push2(&b->pts, b->ring.v[2 * i], b->ring.v[2 * i + 1]);
	b->ring.n = 0;
}
/* OutlineBuilder impl — ring_builder.rs:67-117; callbacks arrive as f32, widened to f64 (point.rs:108-112) */
static void rb_move_to(ring_builder *b, float x, float y)
{
	rb_save_ring(b);
	dvec_push2(&b->ring, (double)x, (double)y);
}
static void rb_line_to(ring_builder *b, float x, float y) { dvec_push2(&b->ring, (double)x, (double)y); }
static void rb_quad_to(ring_builder *b, float x1, float y1, float x, float y)
{
	if (b->ring.n == 0)
		return;
	double sx = b->ring.v[b->ring.n - 2], sy = b->ring.v[b->ring.n - 1];
	flatten_quad(&b->ring, sx, sy, (double)x1, (double)y1, (double)x, (double)y, PRECISION);
}
static void rb_close(ring_builder *b) { rb_save_ring(b); }

/* ------------------------------------------------------------------------------------------ */
/* glyf outline → callbacks (restated ttf-parser glyf::Table::outline)      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	float a, b, c, d, e, f;
} xform;
static const xform XF_ID = {1.f, 0.f, 0.f, 1.f, 0.f, 0.f};
static int xf_is_id(const xform *t)
{
	return t->a == 1.f && t->b == 0.f && t->c == 0.f && t->d == 1.f && t->e == 0.f && t->f == 0.f;
}
/* Transform::combine, all f32 */
static xform xf_combine(xform p, xform q)
{
	xform r;
	r.a = p.a * q.a + p.c * q.b;
	r.b = p.b * q.a + p.d * q.b;
	r.c = p.a * q.c + p.c * q.d;
	r.d = p.b * q.c + p.d * q.d;
	r.e = p.a * q.e + p.c * q.f + p.e;
	r.f = p.b * q.e + p.d * q.f + p.f;
	return r;
}
static void xf_apply(const xform *t, float *x, float *y)
{
	if (xf_is_id(t))
		return;
	float tx = *x, ty = *y;
	*x = t->a * tx + t->c * ty + t->e;
	*y = t->b * tx + t->d * ty + t->f;
}

typedef struct {
	ring_builder *rb;
	xform t;
	int has_on, has_first_off, has_last_off;
	float on_x, on_y, foff_x, foff_y, loff_x, loff_y;
} contour_builder;

static void cb_move(contour_builder *c, float x, float y)
{
	xf_apply(&c->t, &x, &y);
	rb_move_to(c->rb, x, y);
}
static void cb_line(contour_builder *c, float x, float y)
{
	xf_apply(&c->t, &x, &y);
	rb_line_to(c->rb, x, y);
}
static void cb_quad(contour_builder *c, float x1, float y1, float x, float y)
{
	xf_apply(&c->t, &x1, &y1);
	xf_apply(&c->t, &x, &y);
	rb_quad_to(c->rb, x1, y1, x, y);
}
static inline float lerp_half(float a, float b) { return a + 0.5f * (b - a); }

static void cb_finish(contour_builder *c)
{
	if (c->has_first_off && c->has_last_off) {
		float lx = c->loff_x, ly = c->loff_y;
		c->has_last_off = 0;
		cb_quad(c, lx, ly, lerp_half(lx, c->foff_x), lerp_half(ly, c->foff_y));
	}
	if (c->has_on && c->has_first_off)
		cb_quad(c, c->foff_x, c->foff_y, c->on_x, c->on_y);
	else if (c->has_on && c->has_last_off)
		cb_quad(c, c->loff_x, c->loff_y, c->on_x, c->on_y);
	else if (c->has_on)
		cb_line(c, c->on_x, c->on_y);
	c->has_on = c->has_first_off = c->has_last_off = 0;
	rb_close(c->rb);
}

static void cb_point(contour_builder *c, float x, float y, int on, int last)
{
	if (!c->has_on) {
		if (on) {
			c->has_on = 1, c->on_x = x, c->on_y = y;
			cb_move(c, x, y);
		} else if (c->has_first_off) {
			float mx = lerp_half(c->foff_x, x), my = lerp_half(c->foff_y, y);
			c->has_on = 1, c->on_x = mx, c->on_y = my;
			c->has_last_off = 1, c->loff_x = x, c->loff_y = y;
			cb_move(c, mx, my);
		} else {
			c->has_first_off = 1, c->foff_x = x, c->foff_y = y;
		}
	} else if (c->has_last_off) {
		float ox = c->loff_x, oy = c->loff_y;
		if (on) {
			c->has_last_off = 0;
			cb_quad(c, ox, oy, x, y);
		} else {
			c->loff_x = x, c->loff_y = y;
			cb_quad(c, ox, oy, lerp_half(ox, x), lerp_half(oy, y));
		}
	} else if (on) {
		cb_line(c, x, y);
	} else {
		c->has_last_off = 1, c->loff_x = x, c->loff_y = y;
	}
	if (last)
		cb_finish(c);
}

static int glyph_span(const vgo_font *f, uint32_t gid, span *out)
{
	if (gid >= f->num_glyphs)
		return 0;
	uint32_t a, b;
	if (f->loca_long) {
		if (((size_t)gid + 2) * 4 > f->loca.len)
			return 0;
		a = rd32(f->loca.p + (size_t)gid * 4);
		b = rd32(f->loca.p + (size_t)gid * 4 + 4);
	} else {
		if (((size_t)gid + 2) * 2 > f->loca.len)
			return 0;
		a = 2u * rd16(f->loca.p + (size_t)gid * 2);
		b = 2u * rd16(f->loca.p + (size_t)gid * 2 + 2);
	}
	if (b <= a || b > f->glyf.len)
		return 0;
	out->p = f->glyf.p + a;
	out->len = b - a;
	return 1;
}

static void outline_impl(const vgo_font *f, span g, int depth, xform t, ring_builder *rb)
{
	if (depth >= 32 || g.len < 10)
		return;
	int16_t ncont = rds16(g.p);
	size_t pos = 10;
	if (ncont > 0) {
		size_t n = (size_t)ncont;
		if (pos + 2 * n + 2 > g.len)
			return;
		const uint8_t *ends = g.p + pos;
		uint32_t total = (uint32_t)rd16(ends + 2 * (n - 1)) + 1;
		if (total == 1)
			return; /* single-point glyphs carry no outline */
		pos += 2 * n;
		uint16_t ilen = rd16(g.p + pos);
		pos += 2 + (size_t)ilen;
		if (pos > g.len)
			return;
		/* flags with REPEAT */
		uint8_t *flags = (uint8_t *)malloc(total);
		uint32_t k = 0;
		while (k < total) {
			if (pos >= g.len) {
				free(flags);
				return;
			}
			uint8_t fl = g.p[pos++];
			flags[k++] = fl;
			if (fl & 0x08) {
				if (pos >= g.len) {
					free(flags);
					return;
				}
				uint8_t rep = g.p[pos++];
				while (rep-- && k < total)
					flags[k++] = fl;
			}
		}
		size_t xlen = 0;
		for (k = 0; k < total; k++)
			xlen += (flags[k] & 0x02) ? 1 : ((flags[k] & 0x10) ? 0 : 2);
		size_t xpos = pos, ypos = pos + xlen;
		if (ypos > g.len) {
			free(flags);
			return;
		}
		contour_builder cb;
		memset(&cb, 0, sizeof(cb));
		cb.rb = rb;
		cb.t = t;
		int16_t x = 0, y = 0;
		uint32_t contour = 0;
		uint32_t cend = rd16(ends);
		for (k = 0; k < total; k++) {
			uint8_t fl = flags[k];
			if (fl & 0x02) {
				if (xpos >= g.len)
					break;
				int16_t dx = g.p[xpos++];
				x = (int16_t)(x + ((fl & 0x10) ? dx : -dx));
			} else if (!(fl & 0x10)) {
				if (xpos + 2 > g.len)
					break;
				x = (int16_t)(x + rds16(g.p + xpos));
				xpos += 2;
			}
			if (fl & 0x04) {
				if (ypos >= g.len)
					break;
				int16_t dy = g.p[ypos++];
				y = (int16_t)(y + ((fl & 0x20) ? dy : -dy));
			} else if (!(fl & 0x20)) {
				if (ypos + 2 > g.len)
					break;
				y = (int16_t)(y + rds16(g.p + ypos));
				ypos += 2;
			}
			int last = (k == cend);
			cb_point(&cb, (float)x, (float)y, fl & 0x01, last);
			if (last) {
				contour++;
				if (contour < n)
					cend = rd16(ends + 2 * contour);
			}
		}
		free(flags);
	} else if (ncont < 0) {
		for (;;) {
			if (pos + 4 > g.len)
				return;
			uint16_t fl = rd16(g.p + pos), child = rd16(g.p + pos + 2);
			pos += 4;
			xform ct = XF_ID;
			if (fl & 0x0001) {
				if (pos + 4 > g.len)
					return;
				if (fl & 0x0002) {
					ct.e = (float)rds16(g.p + pos);
					ct.f = (float)rds16(g.p + pos + 2);
				}
				pos += 4;
			} else {
				if (pos + 2 > g.len)
					return;
				if (fl & 0x0002) {
					ct.e = (float)(int8_t)g.p[pos];
					ct.f = (float)(int8_t)g.p[pos + 1];
				}
				pos += 2;
			}
			if (fl & 0x0080) {
				if (pos + 8 > g.len)
					return;
				ct.a = (float)rds16(g.p + pos) / 16384.0f;
				ct.b = (float)rds16(g.p + pos + 2) / 16384.0f;
				ct.c = (float)rds16(g.p + pos + 4) / 16384.0f;
				ct.d = (float)rds16(g.p + pos + 6) / 16384.0f;
				pos += 8;
			} else if (fl & 0x0040) {
				if (pos + 4 > g.len)
					return;
				ct.a = (float)rds16(g.p + pos) / 16384.0f;
				ct.d = (float)rds16(g.p + pos + 2) / 16384.0f;
				pos += 4;
			} else if (fl & 0x0008) {
				if (pos + 2 > g.len)
					return;
				ct.a = (float)rds16(g.p + pos) / 16384.0f;
				ct.d = ct.a;
				pos += 2;
			}
			span cg;
			if (glyph_span(f, child, &cg))
				outline_impl(f, cg, depth + 1, xf_combine(t, ct), rb);
			if (!(fl & 0x0020))
				break;
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* CFF 1 outlines → callbacks (restated ttf-parser 0.25.1 cff1.rs / charstring.rs / dict.rs /   */
/* index.rs; reached from renderer.rs:110 when the face has no glyf + loca).  Parity unpinned:  */
/* no fixture of the reference is CFF; tests use synthetic known-answer fonts.                  */
/* ------------------------------------------------------------------------------------------ */
static void rb_curve_to(ring_builder *b, float x1, float y1, float x2, float y2, float x, float y)
{
	if (b->ring.n == 0) /* ring_builder.rs:99 */
		return;
	double s[2] = {b->ring.v[b->ring.n - 2], b->ring.v[b->ring.n - 1]};
	double c1[2] = {(double)x1, (double)y1}, c2[2] = {(double)x2, (double)y2}, e[2] = {(double)x, (double)y};
	flatten_cubic(&b->ring, s, c1, c2, e, PRECISION);
}

typedef struct {
	uint32_t count;
	uint32_t osz;
	const uint8_t *offs;
	span data;
} cff_index;

static uint32_t cff_off(const uint8_t *p, uint32_t osz)
{
	uint32_t v = 0;
	while (osz--)
		v = (v << 8) | *p++;
	return v;
}

/* parse_index at *pos of s; advances *pos; 0 = None */
static int cff_read_index_w(span s, size_t *pos, cff_index *ix, int wide);
static int cff_read_index(span s, size_t *pos, cff_index *ix)
{
	return cff_read_index_w(s, pos, ix, 0);
}
/* wide: CFF 2's parse_index::<u32> */
static int cff_read_index_w(span s, size_t *pos, cff_index *ix, int wide)
{
	memset(ix, 0, sizeof(*ix));
	const size_t cb = wide ? 4 : 2;
	if (*pos > s.len || s.len - *pos < cb)
		return 0;
	uint32_t count = wide ? rd32(s.p + *pos) : rd16(s.p + *pos);
	*pos += cb;
	if (count == 0xffffffffu)
		return 0;
	if (count == 0)
		return 1;
	if (s.len - *pos < 1)
		return 0;
	uint32_t osz = s.p[(*pos)++];
	if (osz < 1 || osz > 4)
		return 0;
	size_t olen = ((size_t)count + 1) * osz;
	if (s.len - *pos < olen)
		return 0;
	const uint8_t *offs = s.p + *pos;
	*pos += olen;
	uint32_t last = cff_off(offs + (size_t)count * osz, osz);
	if (last == 0)
		return 1; /* empty index */
	if (s.len - *pos < (size_t)last - 1)
		return 0;
	ix->count = count, ix->osz = osz, ix->offs = offs;
	ix->data.p = s.p + *pos, ix->data.len = (size_t)last - 1;
	*pos += (size_t)last - 1;
	return 1;
}

static int cff_index_get(const cff_index *ix, uint32_t i, span *out)
{
	if (i >= ix->count)
		return 0;
	uint32_t a = cff_off(ix->offs + (size_t)i * ix->osz, ix->osz), b = cff_off(ix->offs + ((size_t)i + 1) * ix->osz, ix->osz);
	if (a == 0 || b == 0 || a > b || (size_t)b - 1 > ix->data.len)
		return 0;
	out->p = ix->data.p + (a - 1), out->len = b - a;
	return 1;
}

/* DICT walker: returns the next operator (two-byte: 1200 + b) or -1; operands in ops[0..*n) (at most 48 kept) */
static int cff_dict_next(span d, size_t *pos, double *ops, int *n)
{
	*n = 0;
	while (*pos < d.len) {
		uint8_t b = d.p[(*pos)++];
		if (b <= 27 || b == 31 || b == 255) {
			if (b != 12)
				return b;
			if (*pos >= d.len)
				return -1;
			return 1200 + d.p[(*pos)++];
		}
		double v = 0.0;
		if (b == 28) {
			if (d.len - *pos < 2)
				return -1;
			v = rds16(d.p + *pos), *pos += 2;
		} else if (b == 29) {
			if (d.len - *pos < 4)
				return -1;
			v = (int32_t)rd32(d.p + *pos), *pos += 4;
		} else if (b == 30) { /* packed-BCD real */
			char txt[80];
			int k = 0, end = 0;
			while (!end) {
				if (*pos >= d.len)
					return -1;
				uint8_t q = d.p[(*pos)++];
				for (int h = 1; h >= 0 && !end; h--) {
					int nib = (q >> (4 * h)) & 15;
					if (nib == 15) {
						end = 1;
					} else if (k > 70 || nib == 13) {
						return -1;
					} else if (nib < 10) {
						txt[k++] = (char)('0' + nib);
					} else if (nib == 10) {
						txt[k++] = '.';
					} else if (nib == 14) {
						txt[k++] = '-';
					} else {
						txt[k++] = 'e';
						if (nib == 12)
							txt[k++] = '-';
					}
				}
			}
			txt[k] = 0;
			char *stop;
			v = strtod(txt, &stop);
			if (k == 0 || *stop)
				return -1;
		} else if (b <= 246) {
			v = (int)b - 139;
		} else {
			if (*pos >= d.len)
				return -1;
			int w = d.p[(*pos)++];
			v = b <= 250 ? (b - 247) * 256 + w + 108 : -(b - 251) * 256 - w - 108;
		}
		if (*n < 48)
			ops[(*n)++] = v;
	}
	return -1;
}

static int cff_as_offset(const double *ops, int n, size_t *out)
{
	if (n != 1 || (int32_t)ops[0] < 0)
		return 0;
	*out = (size_t)(int32_t)ops[0];
	return 1;
}
static int cff_as_range(const double *ops, int n, size_t *start, size_t *len)
{
	if (n != 2 || (int32_t)ops[0] < 0 || (int32_t)ops[1] < 0)
		return 0;
	*len = (size_t)(int32_t)ops[0], *start = (size_t)(int32_t)ops[1];
	return 1;
}

struct vgo_cff {
	span table;
	cff_index gsubrs, charstrings, lsubrs, fdarray;
	int cid, fdsel_format;
	span fdsel;
	int charset_format; /* -1 ISOAdobe, -2 Expert, -3 ExpertSubset, else 0 / 1 / 2 with the records in charset */
	span charset;
	uint32_t charset_n; /* records */
	/* CFF 2 (cff2.rs): ItemVariationStore of the `vstore` operator; axes = the face's variation coordinates (all 0) */
	int cff2, has_vstore;
	uint32_t axes, region_axes, n_regions, n_vdata;
	span vstore, vdata_offs, regions;
};

/* StandardEncoding (Adobe): code -> SID, as runs {first code, last code, first SID} */
static const uint16_t STD_ENC_RUNS[][3] = {{32, 126, 1},   {161, 175, 96},  {177, 180, 111}, {182, 189, 115}, {191, 191, 123},
                                           {193, 200, 124}, {202, 203, 132}, {205, 208, 134}, {225, 225, 138}, {227, 227, 139},
                                           {232, 235, 140}, {241, 241, 144}, {245, 245, 145}, {248, 251, 146}};

/* seac_code_to_glyph_id — cff1.rs; -1 = None */
static int32_t cff_seac_gid(const struct vgo_cff *c, float code)
{
	if (!(code > -1.0f && code < 256.0f))
		return -1;
	uint32_t ch = (uint32_t)(int32_t)code, sid = 0;
	for (size_t i = 0; i < sizeof(STD_ENC_RUNS) / sizeof(STD_ENC_RUNS[0]); i++)
		if (ch >= STD_ENC_RUNS[i][0] && ch <= STD_ENC_RUNS[i][1])
			sid = STD_ENC_RUNS[i][2] + (ch - STD_ENC_RUNS[i][0]);
	if (c->charset_format == -1)
		return ch <= 228 ? (int32_t)sid : -1;
	if (c->charset_format < 0)
		return -1;
	if (sid == 0)
		return 0;
	const uint8_t *r = c->charset.p;
	if (c->charset_format == 0) {
		for (uint32_t g = 0; g < c->charset_n; g++)
			if (rd16(r + 2 * g) == sid)
				return (int32_t)g + 1;
		return -1;
	}
	uint32_t gid = 1;
	for (uint32_t i = 0; i < c->charset_n; i++) {
		uint32_t first = rd16(r), left = c->charset_format == 1 ? r[2] : rd16(r + 2);
		if (sid >= first && sid - first <= left)
			return (int32_t)((gid + (sid - first)) & 0xffff);
		gid += left + 1;
		r += c->charset_format == 1 ? 3 : 4;
	}
	return -1;
}

/* Subrs INDEX of the Private DICT at [start, start + len) of the table; 1 = found and parsed, 0 = none, -1 = malformed */
static int cff_private_subrs(span table, size_t start, size_t len, cff_index *out)
{
	if (start > table.len || table.len - start < len)
		return -1;
	span priv = {table.p + start, len};
	size_t pos = 0, off = 0;
	double ops[48];
	int n, op, have = 0;
	while ((op = cff_dict_next(priv, &pos, ops, &n)) >= 0)
		if (op == 19)
			have = cff_as_offset(ops, n, &off);
	if (!have)
		return 0;
	size_t at = start + off;
	if (at > table.len)
		return -1;
	return cff_read_index(table, &at, out) ? 1 : -1;
}

static vgo_cff *cff_parse(span t)
{
	if (t.len < 4 || t.p[0] != 1)
		return NULL;
	size_t pos = t.p[2] > 4 ? t.p[2] : 4;
	cff_index names, top, strings;
	vgo_cff c;
	memset(&c, 0, sizeof(c));
	c.table = t;
	if (!cff_read_index(t, &pos, &names) || !cff_read_index(t, &pos, &top))
		return NULL;
	span td;
	if (!cff_index_get(&top, 0, &td))
		return NULL;
	size_t charset = 0, encoding = 0, cs = 0, pstart = 0, plen = 0, fda = 0, fds = 0, dpos = 0;
	int has_charset = 0, has_encoding = 0, has_priv = 0, ros = 0, has_fda = 0, has_fds = 0, n, op;
	double ops[48];
	while ((op = cff_dict_next(td, &dpos, ops, &n)) >= 0) {
		if (op == 15)
			has_charset = cff_as_offset(ops, n, &charset);
		else if (op == 16)
			has_encoding = cff_as_offset(ops, n, &encoding);
		else if (op == 17) {
			if (!cff_as_offset(ops, n, &cs))
				return NULL;
		} else if (op == 18)
			has_priv = cff_as_range(ops, n, &pstart, &plen);
		else if (op == 1230)
			ros = 1;
		else if (op == 1236)
			has_fda = cff_as_offset(ops, n, &fda);
		else if (op == 1237)
			has_fds = cff_as_offset(ops, n, &fds);
	}
	if (cs == 0)
		return NULL;
	if (!cff_read_index(t, &pos, &strings) || !cff_read_index(t, &pos, &c.gsubrs))
		return NULL;
	if (cs > t.len || !cff_read_index(t, &cs, &c.charstrings) || c.charstrings.count == 0)
		return NULL;
	uint32_t ng = c.charstrings.count;
	c.charset_format = -1;
	if (has_charset && charset <= 2)
		c.charset_format = -1 - (int)charset;
	else if (has_charset) { /* parse_charset must succeed */
		if (charset >= t.len)
			return NULL;
		size_t p = charset + 1;
		uint8_t fmt = t.p[charset];
		c.charset.p = t.p + p;
		if (fmt == 0) {
			if (t.len - p < ((size_t)ng - 1) * 2)
				return NULL;
			c.charset_n = ng - 1;
		} else if (fmt == 1 || fmt == 2) {
			uint32_t left = ng - 1;
			size_t rec = fmt == 1 ? 3 : 4;
			while (left > 0) {
				if (t.len - p < rec)
					return NULL;
				uint32_t cnt = (fmt == 1 ? t.p[p + 2] : rd16(t.p + p + 2)) + 1u;
				p += rec;
				if (cnt > left)
					return NULL;
				left -= cnt;
				c.charset_n++;
			}
		} else
			return NULL;
		c.charset_format = fmt;
	}
	if (ros) {
		if (!has_charset || !has_fda || !has_fds || charset == 0 || fda == 0 || fds == 0)
			return NULL;
		c.cid = 1;
		if (fda > t.len || !cff_read_index(t, &fda, &c.fdarray))
			return NULL;
		if (fds >= t.len)
			return NULL;
		c.fdsel_format = t.p[fds];
		c.fdsel.p = t.p + fds + 1, c.fdsel.len = t.len - fds - 1;
		if (c.fdsel_format == 0) {
			if (c.fdsel.len < ng)
				return NULL;
			c.fdsel.len = ng;
		} else if (c.fdsel_format != 3)
			return NULL;
	} else {
		if (has_encoding && encoding > 1) { /* parse_encoding must succeed */
			size_t p = encoding;
			if (t.len < 2 || p > t.len - 2)
				return NULL;
			uint8_t fmt = t.p[p], cnt = t.p[p + 1];
			p += 2;
			size_t body = (fmt & 0x7f) == 0 ? cnt : (fmt & 0x7f) == 1 ? (size_t)cnt * 2 : (size_t)-1;
			if (body == (size_t)-1 || t.len - p < body)
				return NULL;
			p += body;
			if (fmt & 0x80) {
				if (p >= t.len || t.len - p - 1 < (size_t)t.p[p] * 3)
					return NULL;
			}
		}
		if (has_priv && cff_private_subrs(t, pstart, plen, &c.lsubrs) < 0)
			return NULL;
	}
	vgo_cff *out = (vgo_cff *)malloc(sizeof(c));
	*out = c;
	return out;
}

/* local subroutines of a CID glyph: FDSelect → FDArray[fd] → Private → Subrs */
/* cff2::Table::parse; axes = min(fvar axis count, 64) */
static vgo_cff *cff2_parse(span t, uint32_t axes)
{
	if (t.len < 5 || t.p[0] != 2 || t.p[2] < 5)
		return NULL;
	size_t pos = t.p[2];
	uint32_t tdlen = rd16(t.p + 3);
	if (pos > t.len || t.len - pos < tdlen)
		return NULL;
	span td = {t.p + pos, tdlen};
	pos += tdlen;
	vgo_cff c;
	memset(&c, 0, sizeof(c));
	c.table = t, c.cff2 = 1, c.axes = axes;
	size_t cs = 0, vs = 0, fda = 0, dpos = 0;
	int has_vs = 0, has_fda = 0, n, op;
	double ops[48];
	while ((op = cff_dict_next(td, &dpos, ops, &n)) >= 0) {
		if (op == 17) {
			if (!cff_as_offset(ops, n, &cs))
				return NULL;
		} else if (op == 24)
			has_vs = cff_as_offset(ops, n, &vs);
		else if (op == 1236)
			has_fda = cff_as_offset(ops, n, &fda);
	}
	if (cs == 0)
		return NULL;
	if (!cff_read_index_w(t, &pos, &c.gsubrs, 1))
		return NULL;
	size_t at = cs;
	if (!cff_read_index_w(t, &at, &c.charstrings, 1))
		return NULL;
	if (has_vs) {
		if (vs > t.len || t.len - vs < 2 + 8)
			return NULL;
		size_t base = vs + 2;
		c.vstore.p = t.p + base, c.vstore.len = t.len - base;
		uint32_t format = rd16(c.vstore.p), roff = rd32(c.vstore.p + 2);
		c.n_vdata = rd16(c.vstore.p + 6);
		if (format != 1 || c.vstore.len < 8 + (size_t)c.n_vdata * 4)
			return NULL;
		c.vdata_offs.p = c.vstore.p + 8, c.vdata_offs.len = (size_t)c.n_vdata * 4;
		if ((size_t)roff + 4 > c.vstore.len)
			return NULL;
		c.region_axes = rd16(c.vstore.p + roff), c.n_regions = rd16(c.vstore.p + roff + 2);
		if ((size_t)roff + 4 + (size_t)c.region_axes * c.n_regions * 6 > c.vstore.len)
			return NULL;
		c.regions.p = c.vstore.p + roff + 4, c.regions.len = (size_t)c.region_axes * c.n_regions * 6;
		c.has_vstore = 1;
	}
	if (has_fda) {
		cff_index fonts;
		at = fda;
		if (!cff_read_index_w(t, &at, &fonts, 1))
			return NULL;
		for (uint32_t i = 0; i < fonts.count; i++) {
			span fd;
			if (!cff_index_get(&fonts, i, &fd))
				return NULL;
			size_t pstart = 0, plen = 0, fpos = 0;
			int has_priv = 0;
			while ((op = cff_dict_next(fd, &fpos, ops, &n)) >= 0)
				if (op == 18)
					has_priv = cff_as_range(ops, n, &pstart, &plen);
			if (!has_priv || pstart > t.len || t.len - pstart < plen)
				return NULL;
			/* the first Font DICT whose Private DICT names local subroutines supplies them (no FDSelect) */
			span priv = {t.p + pstart, plen};
			size_t ppos = 0, soff = 0;
			int have = 0;
			while ((op = cff_dict_next(priv, &ppos, ops, &n)) >= 0)
				if (op == 19)
					have = cff_as_offset(ops, n, &soff);
			if (have) {
				size_t lat = pstart + soff;
				if (lat > t.len || !cff_read_index_w(t, &lat, &c.lsubrs, 1))
					return NULL;
				break;
			}
		}
	}
	vgo_cff *out = malloc(sizeof(c));
	if (out)
		*out = c;
	return out;
}

static int cff_cid_lsubrs(const vgo_cff *c, uint32_t gid, cff_index *out)
{
	uint32_t fd;
	if (c->fdsel_format == 0) {
		if (gid >= c->fdsel.len)
			return 0;
		fd = c->fdsel.p[gid];
	} else {
		span s = c->fdsel;
		if (s.len < 2)
			return 0;
		uint32_t nr = rd16(s.p);
		if (nr == 0 || nr == 0xffff)
			return 0;
		/* nr records of (first u16, fd u8) followed by a sentinel u16 */
		size_t p = 2;
		int hit = 0;
		fd = 0;
		for (uint32_t i = 0; i < nr; i++, p += 3) {
			if (s.len - p < 5)
				return 0;
			uint32_t first = rd16(s.p + p), next = rd16(s.p + p + 3);
			if (gid >= first && gid < next) {
				fd = s.p[p + 2];
				hit = 1;
				break;
			}
		}
		if (!hit)
			return 0;
	}
	span fdict;
	if (!cff_index_get(&c->fdarray, fd, &fdict))
		return 0;
	size_t pos = 0, start = 0, len = 0;
	double ops[48];
	int n, op, have = 0;
	while ((op = cff_dict_next(fdict, &pos, ops, &n)) >= 0)
		if (op == 18) {
			have = cff_as_range(ops, n, &start, &len);
			break;
		}
	if (!have)
		return 0;
	return cff_private_subrs(c->table, start, len, out) == 1;
}

typedef struct {
	const vgo_cff *c;
	ring_builder *rb;
	uint32_t gid;
	float st[513];
	int n, max_n;
	float x, y;
	int moved, first_move, width_seen, endchar, seac, stems;
	int had_vsindex, had_blend, n_scalars;
	float scalars[64];
	int lsubrs_ready;
	cff_index lsubrs;
} cs_state;

static void cs_curve(cs_state *s, float x1, float y1, float x2, float y2, float x, float y)
{
	s->x = x, s->y = y;
	rb_curve_to(s->rb, x1, y1, x2, y2, x, y);
}
static void cs_rrcurve(cs_state *s, const float *a)
{
	float x1 = s->x + a[0], y1 = s->y + a[1], x2 = x1 + a[2], y2 = y1 + a[3];
	cs_curve(s, x1, y1, x2, y2, x2 + a[4], y2 + a[5]);
}
static void cs_rline(cs_state *s, float dx, float dy)
{
	s->x += dx, s->y += dy;
	rb_line_to(s->rb, s->x, s->y);
}

/* update_scalars at all-zero coordinates (var_store.rs evaluate_region / evaluate_axis); 0 = error */
static int cff2_scalars(const vgo_cff *c, uint32_t index, float *scalars, int *count)
{
	*count = 0;
	if (!c->has_vstore || index >= c->n_vdata)
		return 0;
	uint32_t off = rd32(c->vdata_offs.p + (size_t)index * 4);
	if ((size_t)off + 6 > c->vstore.len)
		return 0;
	uint32_t n = rd16(c->vstore.p + off + 4);
	if ((size_t)off + 6 + (size_t)n * 2 > c->vstore.len)
		return 0;
	for (uint32_t k = 0; k < n; k++) {
		uint32_t region = rd16(c->vstore.p + off + 6 + (size_t)k * 2);
		float v = 1.0f;
		for (uint32_t axis = 0; axis < c->axes; axis++) {
			if (region >= c->n_regions || axis >= c->region_axes) {
				v = 0.0f;
				break;
			}
			const uint8_t *r = c->regions.p + ((size_t)region * c->region_axes + axis) * 6;
			int start = rds16(r), peak = rds16(r + 2), end = rds16(r + 4);
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
		if (*count == 64)
			return 0;
		scalars[(*count)++] = v;
	}
	return 1;
}

/* 1 = keep going / finished normally, 0 = error (CFFError) */
static int cs_exec(cs_state *s, span code, int depth)
{
	size_t pc = 0;
	while (pc < code.len) {
		uint8_t op = code.p[pc++];
		float *a = s->st;
		int n = s->n;
		if (op >= 32 || op == 28) { /* operands */
			float v;
			if (op == 28) {
				if (code.len - pc < 2)
					return 0;
				v = (float)rds16(code.p + pc), pc += 2;
			} else if (op == 255) {
				if (code.len - pc < 4)
					return 0;
				v = (float)(int32_t)rd32(code.p + pc) / 65536.0f, pc += 4;
			} else if (op <= 246) {
				v = (float)((int)op - 139);
			} else {
				if (pc >= code.len)
					return 0;
				int w = code.p[pc++];
				v = (float)(op <= 250 ? (op - 247) * 256 + w + 108 : -(op - 251) * 256 - w - 108);
			}
			if (s->n == s->max_n)
				return 0;
			s->st[s->n++] = v;
			continue;
		}
		if (s->c->cff2) {
			if (op == 11 || op == 14)
				return 0; /* no return / endchar in CFF 2 */
			if (op == 15) { /* vsindex */
				if (s->had_blend || s->had_vsindex || n != 1)
					return 0;
				double v = (double)a[0];
				if (!(v > -2147483649.0 && v < 2147483648.0) || (int32_t)v < 0 || (int32_t)v > 65535)
					return 0;
				if (!cff2_scalars(s->c, (uint32_t)(int32_t)v, s->scalars, &s->n_scalars))
					return 0;
				s->had_vsindex = 1;
				s->n = 0;
				continue;
			}
			if (op == 16) { /* blend */
				s->had_blend = 1;
				if (n == 0)
					return 0;
				double v = (double)a[n - 1];
				s->n = --n;
				if (!(v > -2147483649.0 && v < 2147483648.0) || (int32_t)v < 0 || (int32_t)v > 65535)
					return 0;
				int cnt = (int32_t)v, k = s->n_scalars;
				int need = cnt * (k + 1);
				if (n < need)
					return 0;
				int start = n - need;
				for (int i = cnt - 1; i >= 0; i--)
					for (int j = 0; j < k; j++) {
						float delta = s->st[--s->n];
						s->st[start + i] += delta * s->scalars[k - j - 1];
					}
				continue;
			}
			if ((op == 21 || op == 22 || op == 4) && n != (op == 21 ? 2 : 1))
				return 0; /* no width in CFF 2 */
		}
		switch (op) {
		case 1: case 3: case 18: case 23: /* h/vstem(hm) */
			if ((n & 1) && !s->width_seen)
				s->width_seen = 1, n--;
			s->stems += n >> 1;
			s->n = 0;
			break;
		case 19: case 20: /* hintmask, cntrmask */
			s->n = 0;
			if (n & 1)
				s->width_seen = 1, n--;
			s->stems += n >> 1;
			if (code.len - pc < (size_t)((s->stems + 7) >> 3))
				return 0;
			pc += (size_t)((s->stems + 7) >> 3);
			break;
		case 21: case 22: case 4: { /* rmoveto, hmoveto, vmoveto */
			int want = op == 21 ? 2 : 1, i = 0;
			if (n == want + 1) /* a surplus first argument is the width, also in a seac component */
				s->width_seen = 1, i = 1;
			if (n != want + i)
				return 0;
			if (s->first_move)
				s->first_move = 0;
			else
				rb_close(s->rb);
			s->moved = 1;
			if (op == 21)
				s->x += a[i], s->y += a[i + 1];
			else if (op == 22)
				s->x += a[i];
			else
				s->y += a[i];
			rb_move_to(s->rb, s->x, s->y);
			s->n = 0;
			break;
		}
		case 5: /* rlineto */
			if (!s->moved || (n & 1))
				return 0;
			for (int i = 0; i < n; i += 2)
				cs_rline(s, a[i], a[i + 1]);
			s->n = 0;
			break;
		case 6: case 7: /* hlineto, vlineto: alternate */
			if (!s->moved || n == 0)
				return 0;
			for (int i = 0; i < n; i++) {
				if (((i & 1) == 0) == (op == 6))
					cs_rline(s, a[i], 0.0f);
				else
					cs_rline(s, 0.0f, a[i]);
			}
			s->n = 0;
			break;
		case 8: /* rrcurveto */
			if (!s->moved || n % 6)
				return 0;
			for (int i = 0; i < n; i += 6)
				cs_rrcurve(s, a + i);
			s->n = 0;
			break;
		case 24: { /* rcurveline */
			if (!s->moved || n < 8 || (n - 2) % 6)
				return 0;
			int i = 0;
			for (; i < n - 2; i += 6)
				cs_rrcurve(s, a + i);
			cs_rline(s, a[i], a[i + 1]);
			s->n = 0;
			break;
		}
		case 25: { /* rlinecurve */
			if (!s->moved || n < 8 || ((n - 6) & 1))
				return 0;
			int i = 0;
			for (; i < n - 6; i += 2)
				cs_rline(s, a[i], a[i + 1]);
			cs_rrcurve(s, a + i);
			s->n = 0;
			break;
		}
		case 26: case 27: { /* vvcurveto, hhcurveto */
			if (!s->moved)
				return 0;
			int i = 0;
			if (n & 1) {
				if (op == 27)
					s->y += a[0];
				else
					s->x += a[0];
				i = 1;
			}
			if ((n - i) % 4)
				return 0;
			for (; i < n; i += 4) {
				if (op == 27) {
					float x1 = s->x + a[i], y1 = s->y, x2 = x1 + a[i + 1], y2 = y1 + a[i + 2];
					cs_curve(s, x1, y1, x2, y2, x2 + a[i + 3], y2);
				} else {
					float x1 = s->x, y1 = s->y + a[i], x2 = x1 + a[i + 1], y2 = y1 + a[i + 2];
					cs_curve(s, x1, y1, x2, y2, x2, y2 + a[i + 3]);
				}
			}
			s->n = 0;
			break;
		}
		case 30: case 31: { /* vhcurveto, hvcurveto */
			if (!s->moved || n < 4)
				return 0;
			int i = 0, horiz = op == 31;
			while (i < n) {
				int rest = n - i;
				if (rest < 4)
					return 0;
				float extra = rest == 5 ? a[i + 4] : 0.0f;
				if (horiz) {
					float x1 = s->x + a[i], y1 = s->y, x2 = x1 + a[i + 1], y2 = y1 + a[i + 2];
					float ex = rest == 5 ? x2 + extra : x2;
					cs_curve(s, x1, y1, x2, y2, ex, y2 + a[i + 3]);
				} else {
					float x1 = s->x, y1 = s->y + a[i], x2 = x1 + a[i + 1], y2 = y1 + a[i + 2];
					float ey = rest == 5 ? y2 + extra : y2;
					cs_curve(s, x1, y1, x2, y2, x2 + a[i + 3], ey);
				}
				i += rest == 5 ? 5 : 4;
				horiz = !horiz;
			}
			s->n = 0;
			break;
		}
		case 12: { /* flex family */
			if (pc >= code.len)
				return 0;
			uint8_t op2 = code.p[pc++];
			if (op2 < 34 || op2 > 37 || !s->moved)
				return 0;
			float sx = s->x, sy = s->y;
			if (op2 == 35) {
				if (n != 13)
					return 0;
				cs_rrcurve(s, a);
				cs_rrcurve(s, a + 6);
			} else if (op2 == 34) {
				if (n != 7)
					return 0;
				float x1 = sx + a[0], x2 = x1 + a[1], y2 = sy + a[2], x3 = x2 + a[3];
				cs_curve(s, x1, sy, x2, y2, x3, y2);
				float x4 = x3 + a[4], x5 = x4 + a[5];
				cs_curve(s, x4, y2, x5, sy, x5 + a[6], sy);
			} else if (op2 == 36) {
				if (n != 9)
					return 0;
				float x1 = sx + a[0], y1 = sy + a[1], x2 = x1 + a[2], y2 = y1 + a[3], x3 = x2 + a[4];
				cs_curve(s, x1, y1, x2, y2, x3, y2);
				float x4 = x3 + a[5], x5 = x4 + a[6], y5 = y2 + a[7];
				cs_curve(s, x4, y2, x5, y5, x5 + a[8], sy);
			} else {
				if (n != 11)
					return 0;
				float px[6], py[6];
				px[0] = sx, py[0] = sy;
				for (int k = 1; k <= 5; k++)
					px[k] = px[k - 1] + a[2 * k - 2], py[k] = py[k - 1] + a[2 * k - 1];
				float ex = sx, ey = sy;
				if (fabsf(px[5] - sx) > fabsf(py[5] - sy))
					ex = px[5] + a[10];
				else
					ey = py[5] + a[10];
				cs_curve(s, px[1], py[1], px[2], py[2], px[3], py[3]);
				cs_curve(s, px[4], py[4], px[5], py[5], ex, ey);
			}
			s->n = 0;
			break;
		}
		case 10: case 29: { /* callsubr, callgsubr */
			if (n == 0 || depth == 10)
				return 0;
			const cff_index *ix = &s->c->gsubrs;
			if (op == 10) {
				if (!s->lsubrs_ready) {
					if (!s->c->cid)
						s->lsubrs = s->c->lsubrs, s->lsubrs_ready = 1;
					else if (cff_cid_lsubrs(s->c, s->gid, &s->lsubrs))
						s->lsubrs_ready = 1;
					else
						return 0;
				}
				ix = &s->lsubrs;
			}
			int bias = ix->count < 1240 ? 107 : ix->count < 33900 ? 1131 : 32768;
			double v = (double)s->st[--s->n];
			if (!(v > -2147483649.0 && v < 2147483648.0))
				return 0;
			int64_t k = (int64_t)(int32_t)v + bias;
			span body;
			if (k < 0 || !cff_index_get(ix, (uint32_t)k, &body))
				return 0;
			if (!cs_exec(s, body, depth + 1))
				return 0;
			if (s->endchar && !s->seac)
				return pc >= code.len;
			break;
		}
		case 11: /* return */
			return 1;
		case 14: /* endchar */
			if (n == 4 || (n == 5 && !s->width_seen)) { /* seac: [w] adx ady bchar achar */
				int32_t accent = cff_seac_gid(s->c, a[n - 1]);
				int32_t base = accent < 0 ? -1 : cff_seac_gid(s->c, a[n - 2]);
				if (base < 0)
					return 0;
				float ady = a[n - 3], adx = a[n - 4];
				if (n == 5)
					s->width_seen = 1;
				s->n = 0;
				s->seac = 1;
				if (depth == 10)
					return 0;
				span part;
				if (!cff_index_get(&s->c->charstrings, (uint32_t)base, &part) || !cs_exec(s, part, depth + 1))
					return 0;
				s->x = adx, s->y = ady;
				if (!cff_index_get(&s->c->charstrings, (uint32_t)accent, &part) || !cs_exec(s, part, depth + 1))
					return 0;
			} else if (n == 1 && !s->width_seen)
				s->width_seen = 1, s->n = 0;
			if (!s->first_move) {
				s->first_move = 1;
				rb_close(s->rb);
			}
			if (pc < code.len)
				return 0;
			s->endchar = 1;
			return 1;
		default: /* 0, 2, 9, 13, 15, 16, 17: reserved */
			return 0;
		}
	}
	return 1;
}

static void cff_outline(const vgo_cff *c, uint32_t gid, ring_builder *rb)
{
	span code;
	if (!cff_index_get(&c->charstrings, gid, &code))
		return;
	cs_state s;
	memset(&s, 0, sizeof(s));
	s.c = c, s.rb = rb, s.gid = gid, s.first_move = 1, s.max_n = 48;
	if (c->cff2) {
		s.max_n = 513;
		s.width_seen = 1; /* no widths in CFF 2 */
		if (!cff2_scalars(c, 0, s.scalars, &s.n_scalars))
			return;
	}
	cs_exec(&s, code, 0); /* the reference ignores outline_glyph's result: what was emitted stays */
}

/* face.outline_glyph(gid, &mut RingBuilder) + into_rings — src/render/renderer.rs:109-111 */
int vgo_outline_rings(const vgo_font *f, uint32_t gid, vgo_rings *out)
{
	ring_builder rb;
	memset(&rb, 0, sizeof(rb));
	span g;
	if (f->glyf.len == 0 || f->loca.len == 0) { /* ttf-parser: glyf first, then cff */
		if (f->cff)
			cff_outline(f->cff, gid, &rb);
		else if (f->cff2)
			cff_outline(f->cff2, gid, &rb);
	} else if (glyph_span(f, gid, &g))
		outline_impl(f, g, 0, XF_ID, &rb);
	rb_save_ring(&rb); /* into_rings — ring_builder.rs:26-29 */
	free(rb.ring.v);
	uvec_push(&rb.starts, (uint32_t)(rb.pts.n / 2));
	out->xy = rb.pts.v;
	out->ring_start = rb.starts.v;
	out->n_rings = (uint32_t)rb.starts.n - 1;
	out->n_points = (uint32_t)(rb.pts.n / 2);
	return (int)out->n_rings;
}

void vgo_rings_free(vgo_rings *r)
{
	free(r->xy);
	free(r->ring_start);
	memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------------------------------ */
/* Segment distance — src/geometry/segment.rs:54-99, point.rs:38-42                            */
/* ------------------------------------------------------------------------------------------ */
static inline double sqdist_pt(double ax, double ay, double bx, double by)
{
	double dx = bx - ax, dy = by - ay; /* self=a, other=b: other - self */
	return dx * dx + dy * dy;
}

double vgo_segment_sqdist(double vx, double vy, double wx, double wy, double px, double py)
{
	double l2 = sqdist_pt(vx, vy, wx, wy);
	double qx, qy;
	if (l2 == 0.0) {
		qx = vx, qy = vy;
	} else {
		double t = ((px - vx) * (wx - vx) + (py - vy) * (wy - vy)) / l2;
		if (t < 0.0) {
			qx = vx, qy = vy;
		} else if (t > 1.0) {
			qx = wx, qy = wy;
		} else {
			qx = vx + t * (wx - vx);
			qy = vy + t * (wy - vy);
		}
	}
	return sqdist_pt(px, py, qx, qy);
}

/* ------------------------------------------------------------------------------------------ */
/* R-tree stand-in for rstar 0.13 (render/rtree_segments.rs:20-68).                            */
/* Sort-tile-recursive bulk load, fanout 6 (rstar's default max node size); the query returns  */
/* every entry whose AABB intersects the query AABB (inclusive).  Output-neutral.  */
/* ------------------------------------------------------------------------------------------ */
#define RT_M 6
typedef struct {
	double minx, miny, maxx, maxy;
	uint32_t first, count; /* children (node ids) or, for leaves, entry ids */
	int leaf;
} rt_node;
typedef struct {
	rt_node *nodes;
	uint32_t n_nodes, root;
	uint32_t *entries; /* segment ids, leaf order */
	const double *segs;
} rtree;

typedef struct {
	double cx, cy;
	uint32_t id;
} rt_item;
static int cmp_cx(const void *a, const void *b)
{
	double x = ((const rt_item *)a)->cx, y = ((const rt_item *)b)->cx;
	return x < y ? -1 : x > y;
}
static int cmp_cy(const void *a, const void *b)
{
	double x = ((const rt_item *)a)->cy, y = ((const rt_item *)b)->cy;
	return x < y ? -1 : x > y;
}

static void str_order(rt_item *items, uint32_t n)
{
	/* tile: sort by x, cut into vertical slices of S*M items, sort each slice by y */
	qsort(items, n, sizeof(rt_item), cmp_cx);
	uint32_t leaves = (n + RT_M - 1) / RT_M;
	uint32_t slices = (uint32_t)ceil(sqrt((double)leaves));
	uint32_t per = slices * RT_M;
	for (uint32_t i = 0; i < n; i += per) {
		uint32_t c = n - i < per ? n - i : per;
		qsort(items + i, c, sizeof(rt_item), cmp_cy);
	}
}

static void rtree_build(rtree *t, const double *segs, uint32_t n)
{
	memset(t, 0, sizeof(*t));
	t->segs = segs;
	if (n == 0)
		return;
	/* node count bound: sum of ceil(n/M^k) */
	uint32_t bound = 0, c = n;
	do {
		c = (c + RT_M - 1) / RT_M;
		bound += c;
	} while (c > 1);
	t->nodes = (rt_node *)malloc(sizeof(rt_node) * bound);
	t->entries = (uint32_t *)malloc(sizeof(uint32_t) * n);
	rt_item *items = (rt_item *)malloc(sizeof(rt_item) * n);
	for (uint32_t i = 0; i < n; i++) {
		const double *s = segs + 4 * (size_t)i;
		items[i].cx = 0.5 * (s[0] + s[2]);
		items[i].cy = 0.5 * (s[1] + s[3]);
		items[i].id = i;
	}
	str_order(items, n);
	/* leaves; envelope — rtree_segments.rs:20-32 */
	uint32_t level_first = 0, level_count = 0;
	for (uint32_t i = 0; i < n; i += RT_M) {
		rt_node *nd = &t->nodes[t->n_nodes++];
		nd->leaf = 1;
		nd->first = i;
		nd->count = n - i < RT_M ? n - i : RT_M;
		nd->minx = nd->miny = INFINITY;
		nd->maxx = nd->maxy = -INFINITY;
		for (uint32_t k = 0; k < nd->count; k++) {
			uint32_t id = items[i + k].id;
			t->entries[i + k] = id;
			const double *s = segs + 4 * (size_t)id;
			nd->minx = fmin(nd->minx, fmin(s[0], s[2]));
			nd->maxx = fmax(nd->maxx, fmax(s[0], s[2]));
			nd->miny = fmin(nd->miny, fmin(s[1], s[3]));
			nd->maxy = fmax(nd->maxy, fmax(s[1], s[3]));
		}
		level_count++;
	}
	/* upper levels */
	while (level_count > 1) {
		rt_item *up = (rt_item *)malloc(sizeof(rt_item) * level_count);
		for (uint32_t i = 0; i < level_count; i++) {
			rt_node *nd = &t->nodes[level_first + i];
			up[i].cx = 0.5 * (nd->minx + nd->maxx);
			up[i].cy = 0.5 * (nd->miny + nd->maxy);
			up[i].id = level_first + i;
		}
		str_order(up, level_count);
		/* children must be contiguous: remap by copying nodes into sorted order */
		rt_node *tmp = (rt_node *)malloc(sizeof(rt_node) * level_count);
		for (uint32_t i = 0; i < level_count; i++)
			tmp[i] = t->nodes[up[i].id];
		memcpy(&t->nodes[level_first], tmp, sizeof(rt_node) * level_count);
		free(tmp);
		free(up);
		uint32_t next_first = t->n_nodes, next_count = 0;
		for (uint32_t i = 0; i < level_count; i += RT_M) {
			rt_node *nd = &t->nodes[t->n_nodes++];
			nd->leaf = 0;
			nd->first = level_first + i;
			nd->count = level_count - i < RT_M ? level_count - i : RT_M;
			nd->minx = nd->miny = INFINITY;
			nd->maxx = nd->maxy = -INFINITY;
			for (uint32_t k = 0; k < nd->count; k++) {
				rt_node *ch = &t->nodes[nd->first + k];
				nd->minx = fmin(nd->minx, ch->minx);
				nd->maxx = fmax(nd->maxx, ch->maxx);
				nd->miny = fmin(nd->miny, ch->miny);
				nd->maxy = fmax(nd->maxy, ch->maxy);
			}
			next_count++;
		}
		level_first = next_first;
		level_count = next_count;
	}
	t->root = t->n_nodes - 1;
	free(items);
}

static void rtree_free(rtree *t)
{
	free(t->nodes);
	free(t->entries);
}

/* min_distance_to_line_segment — render/rtree_segments.rs:40-68 */
static double rtree_min_distance(const rtree *t, double px, double py, double radius)
{
	double best = INFINITY;
	if (t->n_nodes == 0)
		return sqrt(best);
	double qx0 = px - radius, qy0 = py - radius, qx1 = px + radius, qy1 = py + radius;
	uint32_t stack[128];
	int sp = 0;
	stack[sp++] = t->root;
	while (sp > 0) {
		const rt_node *nd = &t->nodes[stack[--sp]];
		if (nd->minx > qx1 || nd->maxx < qx0 || nd->miny > qy1 || nd->maxy < qy0)
			continue;
		if (nd->leaf) {
			for (uint32_t k = 0; k < nd->count; k++) {
				const double *s = t->segs + 4 * (size_t)t->entries[nd->first + k];
				double sx0 = s[0] < s[2] ? s[0] : s[2], sx1 = s[0] < s[2] ? s[2] : s[0];
				double sy0 = s[1] < s[3] ? s[1] : s[3], sy1 = s[1] < s[3] ? s[3] : s[1];
				if (sx0 > qx1 || sx1 < qx0 || sy0 > qy1 || sy1 < qy0)
					continue;
				double d = vgo_segment_sqdist(s[0], s[1], s[2], s[3], px, py);
				if (d < best)
					best = d;
			}
		} else {
			for (uint32_t k = 0; k < nd->count && sp < 128; k++)
				stack[sp++] = nd->first + k;
		}
	}
	return sqrt(best);
}

double vgo_min_distance(const double *segs, uint32_t n, double px, double py, double radius)
{
	rtree t;
	rtree_build(&t, segs, n);
	double d = rtree_min_distance(&t, px, py, radius);
	rtree_free(&t);
	return d;
}

/* ------------------------------------------------------------------------------------------ */
/* renderer_precise — src/render/renderer_precise.rs:8-84                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	double x;
	int32_t sign;
	uint32_t order;
} crossing;
static int cmp_cross(const void *a, const void *b)
{
	const crossing *p = (const crossing *)a, *q = (const crossing *)b;
	if (p->x < q->x)
		return -1;
	if (p->x > q->x)
		return 1;
	return p->order < q->order ? -1 : p->order > q->order; /* stable, like sort_by */
}

/* segs: n quads (x0,y0,x1,y1) in pixel space = rings.get_segments() (rings.rs:75-81) */
static void render_precise_segs(int32_t gx0, int32_t gy0, uint32_t W, uint32_t H, const double *segs, uint32_t n,
                                uint8_t *bitmap)
{
	rtree rt;
	rtree_build(&rt, segs, n);
	const double radius_by_256 = 256.0 / SDF_RADIUS;
	const double x0 = (double)gx0 + 0.5;
	const double y0 = (double)gy0 + 0.5;
	crossing *cr = (crossing *)malloc(sizeof(crossing) * (n ? n : 1));
	for (uint32_t y = 0; y < H; y++) {
		double py = (double)y + y0;
		uint32_t nc = 0;
		for (uint32_t i = 0; i < n; i++) {
			const double *s = segs + 4 * (size_t)i;
			double sx = s[0], sy = s[1], ex = s[2], ey = s[3];
			if (sy <= py && ey > py) {
				double t = (py - sy) / (ey - sy);
				cr[nc].x = sx + t * (ex - sx), cr[nc].sign = 1, cr[nc].order = nc;
				nc++;
			} else if (sy > py && ey <= py) {
				double t = (py - sy) / (ey - sy);
				cr[nc].x = sx + t * (ex - sx), cr[nc].sign = -1, cr[nc].order = nc;
				nc++;
			}
		}
		qsort(cr, nc, sizeof(crossing), cmp_cross);
		int32_t wn = 0;
		uint32_t idx = 0;
		for (uint32_t x = 0; x < W; x++) {
			double px = (double)x + x0;
			while (idx < nc && cr[idx].x <= px) {
				wn -= cr[idx].sign;
				idx++;
			}
			int inside = wn != 0;
			double d = rtree_min_distance(&rt, px, py, SDF_RADIUS);
			if (inside)
				d = -d;
			d = d * radius_by_256 + CUTOFF;
			double v = 255.0 - d;
			v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v); /* clamp(0,255) */
			bitmap[(size_t)(H - 1 - y) * W + x] = (uint8_t)round(v); /* round = half away from zero */
		}
	}
	free(cr);
	rtree_free(&rt);
}

static uint32_t rings_to_segments(const double *xy, const uint32_t *ring_start, uint32_t n_rings, double **out)
{
	uint32_t n = 0;
	for (uint32_t r = 0; r < n_rings; r++) {
		uint32_t c = ring_start[r + 1] - ring_start[r];
		if (c >= 2)
			n += c - 1;
	}
	double *segs = (double *)malloc(sizeof(double) * 4 * (n ? n : 1));
	uint32_t k = 0;
	for (uint32_t r = 0; r < n_rings; r++)
		for (uint32_t i = ring_start[r]; i + 1 < ring_start[r + 1]; i++) {
			segs[4 * k + 0] = xy[2 * i], segs[4 * k + 1] = xy[2 * i + 1];
			segs[4 * k + 2] = xy[2 * i + 2], segs[4 * k + 3] = xy[2 * i + 3];
			k++;
		}
	*out = segs;
	return n;
}

int vgo_renderer_precise(int32_t x0, int32_t y0, uint32_t W, uint32_t H, const double *xy, const uint32_t *ring_start,
                         uint32_t n_rings, uint8_t *bitmap)
{
	double *segs;
	uint32_t n = rings_to_segments(xy, ring_start, n_rings, &segs);
	render_precise_segs(x0, y0, W, H, segs, n, bitmap);
	free(segs);
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Renderer::render_glyph — src/render/renderer.rs:103-149                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	int some;          /* Option<PbfGlyph> */
	int empty;         /* PbfGlyph::empty */
	uint32_t advance;
	int32_t x0, y0, x1, y1;
	uint32_t W, H;
	vgo_rings rings;   /* pixel space */
} prepared;

static void prepare(const vgo_font *f, uint32_t index, prepared *p)
{
	memset(p, 0, sizeof(*p));
	/* char::from_u32(index)? — surrogates and > 0x10FFFF are None (renderer.rs:104) */
	if ((index >= 0xD800 && index <= 0xDFFF) || index > 0x10FFFF)
		return;
	int32_t gid = vgo_font_glyph_index(f, index);
	if (gid < 0)
		return;
	p->some = 1;
	double scale = (double)GLYPH_SIZE / (double)f->upm;
	vgo_outline_rings(f, (uint32_t)gid, &p->rings);
	int32_t adv = vgo_font_hor_advance(f, (uint32_t)gid);
	double advance_float = (double)(adv < 0 ? 0 : adv) * scale * 0.95;
	double ar = round(advance_float);
	p->advance = ar <= 0.0 ? 0u : (ar >= 4294967295.0 ? 4294967295u : (uint32_t)ar); /* saturating `as u32` */
	if (p->rings.n_rings == 0) {
		p->empty = 1;
		return;
	}
	/* rings.scale(scale); rings.translate((dx, 0)) — renderer.rs:122-131, point.rs:83-99 */
	double dx = ((double)p->advance - advance_float) / 2.0;
	double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
	for (uint32_t i = 0; i < p->rings.n_points; i++) {
		double x = p->rings.xy[2 * i], y = p->rings.xy[2 * i + 1];
		x *= scale;
		y *= scale;
		x += dx;
		y += 0.0;
		p->rings.xy[2 * i] = x, p->rings.xy[2 * i + 1] = y;
		/* BBox::include_point — bbox.rs:64-69 */
		minx = fmin(minx, x), miny = fmin(miny, y), maxx = fmax(maxx, x), maxy = fmax(maxy, y);
	}
	/* prepare_glyph — renderer.rs:64-91; BBox::is_empty — bbox.rs:56-58 */
	if (maxx <= minx && maxy <= miny) {
		p->empty = 1;
		return;
	}
	p->x0 = (int32_t)floor(minx) - BUFFER;
	p->y0 = (int32_t)floor(miny) - BUFFER;
	p->x1 = (int32_t)ceil(maxx) + BUFFER;
	p->y1 = (int32_t)ceil(maxy) + BUFFER;
	p->W = (uint32_t)(p->x1 - p->x0);
	p->H = (uint32_t)(p->y1 - p->y0);
}

int vgo_render_glyph(const vgo_font *f, uint32_t index, int mode, vgo_glyph *out)
{
	prepared p;
	prepare(f, index, &p);
	memset(out, 0, sizeof(*out));
	if (!p.some)
		return 0;
	out->id = index;
	out->advance = p.advance;
	if (p.empty) { /* PbfGlyph::empty — protobuf/glyph.rs:60-70 */
		vgo_rings_free(&p.rings);
		return 1;
	}
	size_t n = (size_t)p.W * p.H;
	uint8_t *bm = (uint8_t *)calloc(n ? n : 1, 1);
	double *segs;
	uint32_t ns = rings_to_segments(p.rings.xy, p.rings.ring_start, p.rings.n_rings, &segs);
	if (mode == VGO_MODE_PRECISE)
		render_precise_segs(p.x0, p.y0, p.W, p.H, segs, ns, bm);
	/* VGO_MODE_DUMMY: zero bitmap of the right size — render/renderer_dummy.rs:3-5 */
	free(segs);
	/* glyph.y1 -= GLYPH_SIZE; into_pbf_glyph — renderer.rs:146-148, result.rs:66-76 */
	out->has_bitmap = 1;
	out->bitmap = bm;
	out->bitmap_len = n;
	out->width = p.W - 2 * BUFFER;
	out->height = p.H - 2 * BUFFER;
	out->left = p.x0 + BUFFER;
	out->top = (p.y1 - GLYPH_SIZE) - BUFFER;
	out->n_segments = ns;
	vgo_rings_free(&p.rings);
	return 1;
}

void vgo_glyph_free(vgo_glyph *g)
{
	free(g->bitmap);
	g->bitmap = NULL;
}

uint32_t vgo_glyph_segments(const vgo_font *f, uint32_t codepoint, double **segs, int32_t frame[4])
{
	prepared p;
	prepare(f, codepoint, &p);
	*segs = NULL;
	frame[0] = frame[1] = frame[2] = frame[3] = 0;
	if (!p.some || p.empty) {
		vgo_rings_free(&p.rings);
		return 0;
	}
	uint32_t ns = rings_to_segments(p.rings.xy, p.rings.ring_start, p.rings.n_rings, segs);
	frame[0] = p.x0, frame[1] = p.y0, frame[2] = (int32_t)p.W, frame[3] = (int32_t)p.H;
	vgo_rings_free(&p.rings);
	return ns;
}

/* ------------------------------------------------------------------------------------------ */
/* PBF (prost encoding of protobuf/{glyph,fontstack,glyphs}.rs) — SURVEY Appendix A-12         */
/* ------------------------------------------------------------------------------------------ */
static void pb_varint(bvec *b, uint64_t v)
{
	while (v >= 0x80) {
		bvec_byte(b, (uint8_t)(v | 0x80));
		v >>= 7;
	}
	bvec_byte(b, (uint8_t)v);
}
static size_t varint_len(uint64_t v)
{
	size_t n = 1;
	while (v >= 0x80) {
		v >>= 7;
		n++;
	}
	return n;
}
static inline uint32_t zigzag(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }

static size_t glyph_body_len(const vgo_glyph *g)
{
	size_t n = 1 + varint_len(g->id);
	if (g->has_bitmap)
		n += 1 + varint_len(g->bitmap_len) + g->bitmap_len;
	n += 1 + varint_len(g->width) + 1 + varint_len(g->height);
	n += 1 + varint_len(zigzag(g->left)) + 1 + varint_len(zigzag(g->top));
	n += 1 + varint_len(g->advance);
	return n;
}
static void glyph_encode(bvec *b, const vgo_glyph *g)
{
	bvec_byte(b, 0x08), pb_varint(b, g->id);
	if (g->has_bitmap) {
		bvec_byte(b, 0x12), pb_varint(b, g->bitmap_len);
		bvec_put(b, g->bitmap, g->bitmap_len);
	}
	bvec_byte(b, 0x18), pb_varint(b, g->width);
	bvec_byte(b, 0x20), pb_varint(b, g->height);
	bvec_byte(b, 0x28), pb_varint(b, zigzag(g->left));
	bvec_byte(b, 0x30), pb_varint(b, zigzag(g->top));
	bvec_byte(b, 0x38), pb_varint(b, g->advance);
}

/* PbfGlyphs{stacks:[Fontstack{name, range, glyphs}]}.encode — protobuf/glyphs.rs:66-70 */
static void pbf_encode_block(bvec *out, const char *name, uint32_t start, const vgo_glyph *glyphs, uint32_t n)
{
	char range[32];
	int rl = 0;
	{ /* GlyphBlock::range — font/glyph_block.rs:52-58 */
		char tmp[32];
		uint32_t vals[2] = {start, start + 255};
		for (int k = 0; k < 2; k++) {
			int tl = 0;
			uint32_t v = vals[k];
			do {
				tmp[tl++] = (char)('0' + v % 10);
				v /= 10;
			} while (v);
			while (tl)
				range[rl++] = tmp[--tl];
			if (k == 0)
				range[rl++] = '-';
		}
	}
	size_t nl = strlen(name);
	size_t stack_len = 1 + varint_len(nl) + nl + 1 + varint_len((size_t)rl) + (size_t)rl;
	for (uint32_t i = 0; i < n; i++) {
		size_t gl = glyph_body_len(&glyphs[i]);
		stack_len += 1 + varint_len(gl) + gl;
	}
	bvec_byte(out, 0x0A), pb_varint(out, stack_len);
	bvec_byte(out, 0x0A), pb_varint(out, nl), bvec_put(out, name, nl);
	bvec_byte(out, 0x12), pb_varint(out, (uint64_t)rl), bvec_put(out, range, (size_t)rl);
	for (uint32_t i = 0; i < n; i++) {
		bvec_byte(out, 0x1A), pb_varint(out, glyph_body_len(&glyphs[i]));
		glyph_encode(out, &glyphs[i]);
	}
}

/* ------------------------------------------------------------------------------------------ */
/* FontWrapper / GlyphBlock / FontManager restated                                             */
/* ------------------------------------------------------------------------------------------ */
struct vgo_fontset {
	char *id;
	vgo_font **fonts;
	uint32_t n_fonts;
	int16_t *owner; /* 65536 entries: index of the first file that has the code point, -1 none */
};

/* name_to_id — src/font/manager.rs:141-147: lowercase, collapse [-_\s]+ to one '_', trim */
const char *vgo_name_to_id(const char *name, char *buf, size_t cap)
{
	size_t n = 0;
	int pending = 0;
	for (const char *p = name; *p; p++) {
		char c = *p;
		if (c >= 'A' && c <= 'Z')
			c = (char)(c - 'A' + 'a');
		if (c == '-' || c == '_' || c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') {
			pending = 1;
			continue;
		}
		if (pending && n > 0 && n + 1 < cap)
			buf[n++] = '_';
		pending = 0;
		if (n + 1 < cap)
			buf[n++] = c;
	}
	buf[n] = 0;
	return buf;
}

vgo_fontset *vgo_fontset_new(const char *font_id)
{
	vgo_fontset *s = (vgo_fontset *)calloc(1, sizeof(*s));
	s->id = strdup(font_id);
	return s;
}
void vgo_fontset_free(vgo_fontset *s)
{
	if (!s)
		return;
	free(s->id);
	free(s->fonts);
	free(s->owner);
	free(s);
}
void vgo_fontset_add(vgo_fontset *s, vgo_font *f)
{
	s->fonts = (vgo_font **)realloc(s->fonts, sizeof(vgo_font *) * (s->n_fonts + 1));
	s->fonts[s->n_fonts++] = f;
	free(s->owner);
	s->owner = NULL;
}

/* FontWrapper::get_blocks — src/font/wrapper.rs:53-76; set_glyph_font = or_insert (glyph_block.rs:34-36) */
static void fontset_index(vgo_fontset *s)
{
	if (s->owner)
		return;
	s->owner = (int16_t *)malloc(sizeof(int16_t) * 65536);
	for (uint32_t i = 0; i < 65536; i++)
		s->owner[i] = -1;
	for (uint32_t fi = 0; fi < s->n_fonts; fi++) {
		size_t n = vgo_font_codepoints(s->fonts[fi], NULL, 0);
		uint32_t *cps = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
		vgo_font_codepoints(s->fonts[fi], cps, n);
		for (size_t i = 0; i < n; i++) {
			if (cps[i] > 0xFFFF)
				continue;
			if (s->owner[cps[i]] < 0)
				s->owner[cps[i]] = (int16_t)fi;
		}
		free(cps);
	}
}

void vgo_fontset_block_population(vgo_fontset *s, uint32_t out[256])
{
	fontset_index(s);
	for (uint32_t b = 0; b < 256; b++) {
		uint32_t c = 0;
		for (uint32_t i = 0; i < 256; i++)
			c += s->owner[b * 256 + i] >= 0;
		out[b] = c;
	}
}

static void render_block(vgo_fontset *s, uint32_t block, int mode, bvec *out, vgo_stats *st)
{
	vgo_glyph glyphs[256];
	uint32_t n = 0;
	/* GlyphBlock::render — font/glyph_block.rs:69-80 (ascending id instead of HashMap order) */
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t cp = block * 256 + i;
		int16_t o = s->owner[cp];
		if (o < 0)
			continue;
		if (vgo_render_glyph(s->fonts[o], cp, mode, &glyphs[n])) {
			if (st) {
				st->glyphs++;
				if (glyphs[n].has_bitmap) {
					st->bitmaps++;
					st->pixels += glyphs[n].bitmap_len;
					st->segments += glyphs[n].n_segments;
					st->pairs += glyphs[n].bitmap_len * glyphs[n].n_segments;
				}
			}
			n++;
		}
	}
	pbf_encode_block(out, s->id, block * 256, glyphs, n);
	for (uint32_t i = 0; i < n; i++)
		vgo_glyph_free(&glyphs[i]);
}

int vgo_fontset_render_block(vgo_fontset *s, uint32_t block, int mode, uint8_t **pbf, uint64_t *len)
{
	if (block >= 256)
		return -1;
	fontset_index(s);
	bvec b = {0};
	render_block(s, block, mode, &b, NULL);
	*pbf = b.v;
	*len = b.n;
	return 0;
}

typedef struct {
	vgo_fontset *s;
	int mode;
	uint32_t lo, hi, stride;
	uint32_t next;
	pthread_mutex_t mu;
	bvec *outs; /* per block */
	vgo_stats *per_thread;
} job_ctx;
typedef struct {
	job_ctx *c;
	int tid;
} worker_arg;

static void *worker(void *p)
{
	worker_arg *a = (worker_arg *)p;
	job_ctx *c = a->c;
	for (;;) {
		pthread_mutex_lock(&c->mu);
		uint32_t b = c->next;
		c->next += c->stride;
		pthread_mutex_unlock(&c->mu);
		if (b >= c->hi)
			break;
		render_block(c->s, b, c->mode, &c->outs[b - c->lo], &c->per_thread[a->tid]);
	}
	return NULL;
}

/* FontManager::render_glyphs — src/font/manager.rs:81-125: one task per block, worker pool */
int vgo_fontset_render_all(vgo_fontset *s, int mode, int threads, uint32_t block_lo, uint32_t block_hi, vgo_stats *st)
{
	return vgo_fontset_render_strided(s, mode, threads, block_lo, block_hi, 1, st);
}

/* Same, visiting only blocks block_lo, block_lo+stride, ... (< block_hi): bounded samples for bench.py. */
int vgo_fontset_render_strided(vgo_fontset *s, int mode, int threads, uint32_t block_lo, uint32_t block_hi, uint32_t stride,
                               vgo_stats *st)
{
	if (stride < 1)
		stride = 1;
	if (block_hi > 256)
		block_hi = 256;
	if (block_lo > block_hi)
		block_lo = block_hi;
	if (threads < 1)
		threads = 1;
	fontset_index(s);
	job_ctx c;
	memset(&c, 0, sizeof(c));
	c.s = s, c.mode = mode, c.lo = block_lo, c.hi = block_hi, c.next = block_lo, c.stride = stride;
	pthread_mutex_init(&c.mu, NULL);
	c.outs = (bvec *)calloc(block_hi - block_lo + 1, sizeof(bvec));
	c.per_thread = (vgo_stats *)calloc((size_t)threads, sizeof(vgo_stats));
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
	worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)threads);
	for (int i = 0; i < threads; i++) {
		args[i].c = &c, args[i].tid = i;
		if (threads == 1)
			worker(&args[i]);
		else
			pthread_create(&th[i], NULL, worker, &args[i]);
	}
	if (threads > 1)
		for (int i = 0; i < threads; i++)
			pthread_join(th[i], NULL);
	memset(st, 0, sizeof(*st));
	for (int i = 0; i < threads; i++) {
		st->glyphs += c.per_thread[i].glyphs, st->bitmaps += c.per_thread[i].bitmaps;
		st->pixels += c.per_thread[i].pixels, st->segments += c.per_thread[i].segments;
		st->pairs += c.per_thread[i].pairs;
	}
	uint64_t h = 1469598103934665603ull;
	for (uint32_t b = block_lo; b < block_hi; b += stride) {
		bvec *o = &c.outs[b - block_lo];
		st->pbf_bytes += o->n;
		for (size_t i = 0; i < o->n; i++)
			h = (h ^ o->v[i]) * 1099511628211ull;
		free(o->v);
	}
	st->pbf_checksum = h;
	free(c.outs), free(c.per_thread), free(th), free(args);
	pthread_mutex_destroy(&c.mu);
	return 0;
}
