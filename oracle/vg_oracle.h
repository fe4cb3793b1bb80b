/*
 * vg_oracle.h — CPU oracle for the per-glyph SDF rendering path of versatiles_glyphs v0.9.1.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, f64 restatement of the reference's CPU
 * algorithm (see vg_oracle.c for the file:line each function follows).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (versatiles_glyphs_rs_b200/) never links, imports or calls anything here.
 *
 * Parity pin: every known-answer test the reference holds for this path is reproduced by
 * tests/test_oracle_goldens.py.  Still unpinned (no reference golden
 * exists): scaled / nested composite glyphs, the cubic path beyond the 17-point count, CFF
 * charstring interpretation (no CFF fixture; synthetic known-answer fonts only), and exact u8
 * values of real glyphs — see DESIGN.md "Oracle".
 */
#ifndef VG_ORACLE_H
#define VG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgo_font vgo_font;
typedef struct vgo_fontset vgo_fontset;

/* Flattened outline: closed rings of f64 points (x,y interleaved). */
typedef struct {
	double *xy;           /* 2 * n_points */
	uint32_t *ring_start; /* n_rings + 1 offsets (in points) */
	uint32_t n_rings;
	uint32_t n_points;
} vgo_rings;

/* Mirror of protobuf::PbfGlyph (src/protobuf/glyph.rs:10-41) plus bookkeeping. */
typedef struct {
	uint32_t id;
	int32_t has_bitmap;
	uint32_t width, height;
	int32_t left, top;
	uint32_t advance;
	uint8_t *bitmap; /* (width+6)*(height+6) bytes when has_bitmap */
	uint64_t bitmap_len;
	uint32_t n_segments; /* segments fed to renderer_precise (0 for empty glyphs) */
} vgo_glyph;

enum { VGO_MODE_PRECISE = 0, VGO_MODE_DUMMY = 1 };

/* ---- font (restated ttf-parser 0.25.1 subset) ---- */
vgo_font *vgo_font_parse(const uint8_t *data, size_t len);
void vgo_font_free(vgo_font *f);
uint32_t vgo_font_units_per_em(const vgo_font *f);
uint32_t vgo_font_num_glyphs(const vgo_font *f);
int32_t vgo_font_glyph_index(const vgo_font *f, uint32_t cp); /* -1 = None */
int32_t vgo_font_hor_advance(const vgo_font *f, uint32_t gid); /* -1 = None */
/* sorted union of code points of all unicode cmap subtables; returns the count */
size_t vgo_font_codepoints(const vgo_font *f, uint32_t *out, size_t cap);

/* ---- geometry ---- */
/* RingBuilder over outline_glyph: flattened rings in FONT UNITS. */
int vgo_outline_rings(const vgo_font *f, uint32_t gid, vgo_rings *out);
void vgo_rings_free(vgo_rings *r);
/* Ring::add_quadratic_bezier / add_cubic_bezier: appends points after `start`; returns count written. */
size_t vgo_flatten_quad(const double s[2], const double c[2], const double e[2], double tol_sq, double *out_xy, size_t cap);
size_t vgo_flatten_cubic(const double s[2], const double c1[2], const double c2[2], const double e[2], double tol_sq,
                         double *out_xy, size_t cap);
/* Segment::squared_distance_to_point */
double vgo_segment_sqdist(double vx, double vy, double wx, double wy, double px, double py);
/* min_distance_to_line_segment over a set of segments (x0,y0,x1,y1)*n with the +-radius AABB filter */
double vgo_min_distance(const double *segs, uint32_t n, double px, double py, double radius);

/* ---- render ---- */
/* renderer_precise on explicit rings (pixel space). bitmap: W*H bytes. */
int vgo_renderer_precise(int32_t x0, int32_t y0, uint32_t W, uint32_t H, const double *xy, const uint32_t *ring_start,
                         uint32_t n_rings, uint8_t *bitmap);
/* Renderer::render_glyph. Returns 1 = Some(glyph), 0 = None (skip). */
int vgo_render_glyph(const vgo_font *f, uint32_t codepoint, int mode, vgo_glyph *out);
void vgo_glyph_free(vgo_glyph *g);
/* Pixel-space segments + integer frame of a glyph exactly as renderer_precise receives them.
 * Returns segment count (0 for empty glyph / None); segs = malloc'd (x0,y0,x1,y1) f64 quads. */
uint32_t vgo_glyph_segments(const vgo_font *f, uint32_t codepoint, double **segs, int32_t frame[4] /*x0,y0,W,H*/);
void vgo_free(void *p);

/* ---- font set = FontWrapper (first file wins) + GlyphBlock + PBF ---- */
vgo_fontset *vgo_fontset_new(const char *font_id);
void vgo_fontset_free(vgo_fontset *s);
void vgo_fontset_add(vgo_fontset *s, vgo_font *f); /* borrowed; caller keeps ownership */
/* glyph count per 256-block (256 entries) */
void vgo_fontset_block_population(vgo_fontset *s, uint32_t out[256]);
/* GlyphBlock::render with glyphs in ascending id order. pbf is malloc'd. */
int vgo_fontset_render_block(vgo_fontset *s, uint32_t block, int mode, uint8_t **pbf, uint64_t *len);

typedef struct {
	uint64_t glyphs;        /* Some(glyph) count */
	uint64_t bitmaps;       /* glyphs with a bitmap */
	uint64_t pixels;        /* sum W*H */
	uint64_t segments;      /* sum S */
	uint64_t pairs;         /* sum W*H*S (brute-force-equivalent pairs) */
	uint64_t pbf_bytes;     /* sum of all 256 encoded blocks */
	uint64_t pbf_checksum;  /* FNV-1a over all blocks in block order */
} vgo_stats;
/* FontManager::render_glyphs restated: all 256 blocks, `threads` workers pulling blocks
 * (rayon par_iter granularity, src/font/manager.rs:117-121). block_lo/hi restrict to a sub-range. */
int vgo_fontset_render_all(vgo_fontset *s, int mode, int threads, uint32_t block_lo, uint32_t block_hi, vgo_stats *st);
/* Same over blocks block_lo, block_lo+stride, ... : bounded samples of a workload for bench.py's CPU legs. */
int vgo_fontset_render_strided(vgo_fontset *s, int mode, int threads, uint32_t block_lo, uint32_t block_hi, uint32_t stride,
                               vgo_stats *st);

const char *vgo_name_to_id(const char *name, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
