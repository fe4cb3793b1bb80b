// capi.cc — extern "C" surface of the host library (include/vgb200_host.h).
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../../include/vgb200_host.h"
#include "font.h"
#include "pbf.h"
#include "render.h"

using namespace vgb;

struct vgb_font {
	std::unique_ptr<FontFileEntry> e;
};
struct vgb_renderer {
	std::unique_ptr<Renderer> r;
};
struct vgb_batch {
	std::unique_ptr<GlyphBatch> b;
};
struct vgb_manager {
	FontManager m;
	std::vector<std::string> ids; // stable storage for vgb_manager_font_id
	explicit vgb_manager(bool parallel) : m(parallel) {}
};
struct vgb_writer {
	Writer w;
	explicit vgb_writer(Writer x) : w(std::move(x)) {}
};

namespace {
thread_local std::string g_err;
int fail(const std::string &msg, int code = -1)
{
	g_err = msg;
	return code;
}
void fill_glyph(const PbfGlyph &g, uint32_t n_segments, vgb_glyph *out)
{
	std::memset(out, 0, sizeof(*out));
	out->id = g.id;
	out->has_bitmap = g.has_bitmap ? 1 : 0;
	out->width = g.width;
	out->height = g.height;
	out->left = g.left;
	out->top = g.top;
	out->advance = g.advance;
	out->n_segments = n_segments;
	if (g.has_bitmap) {
		out->bitmap_len = g.bitmap.size();
		out->bitmap = (uint8_t *)std::malloc(g.bitmap.size() ? g.bitmap.size() : 1);
		if (out->bitmap && !g.bitmap.empty())
			std::memcpy(out->bitmap, g.bitmap.data(), g.bitmap.size());
	}
}
} // namespace

extern "C" {

const char *vgb_last_error(void) { return g_err.c_str(); }
void vgb_free(void *p) { std::free(p); }

// ---- font ------------------------------------------------------------------------------------------
vgb_font *vgb_font_from_bytes(const uint8_t *data, size_t len)
{
	std::string err;
	auto e = FontFileEntry::from_bytes(std::vector<uint8_t>(data, data + len), &err);
	if (!e) {
		g_err = err;
		return nullptr;
	}
	return new vgb_font{std::move(e)};
}
vgb_font *vgb_font_from_path(const char *path)
{
	std::string err;
	auto e = FontFileEntry::from_path(path, &err);
	if (!e) {
		g_err = err;
		return nullptr;
	}
	return new vgb_font{std::move(e)};
}
void vgb_font_free(vgb_font *f) { delete f; }
uint32_t vgb_font_units_per_em(const vgb_font *f) { return f->e->face->units_per_em(); }
uint32_t vgb_font_number_of_glyphs(const vgb_font *f) { return f->e->face->number_of_glyphs(); }
int32_t vgb_font_glyph_index(const vgb_font *f, uint32_t cp)
{
	auto g = f->e->face->glyph_index(cp);
	return g ? (int32_t)*g : -1;
}
int32_t vgb_font_hor_advance(const vgb_font *f, uint32_t gid)
{
	if (gid > 0xFFFF)
		return -1;
	auto a = f->e->face->glyph_hor_advance((uint16_t)gid);
	return a ? (int32_t)*a : -1;
}
size_t vgb_font_codepoints(const vgb_font *f, uint32_t *out, size_t cap)
{
	const auto &c = f->e->codepoints;
	if (out)
		for (size_t i = 0; i < c.size() && i < cap; ++i)
			out[i] = c[i];
	return c.size();
}
namespace {
// the raw callback stream of Face::outline_glyph: 7 floats per command (kind, x1, y1, x2, y2, x, y)
class CommandRecorder : public OutlineBuilder {
  public:
	std::vector<float> v;
	void put(float k, float a, float b, float c, float d, float x, float y)
	{
		const float r[7] = {k, a, b, c, d, x, y};
		v.insert(v.end(), r, r + 7);
	}
	void move_to(float x, float y) override { put(0, 0, 0, 0, 0, x, y); }
	void line_to(float x, float y) override { put(1, 0, 0, 0, 0, x, y); }
	void quad_to(float x1, float y1, float x, float y) override { put(2, x1, y1, 0, 0, x, y); }
	void curve_to(float x1, float y1, float x2, float y2, float x, float y) override { put(3, x1, y1, x2, y2, x, y); }
	void close() override { put(4, 0, 0, 0, 0, 0, 0); }
};
} // namespace

int32_t vgb_font_outline_commands(const vgb_font *f, uint32_t glyph_id, float **cmds)
{
	CommandRecorder rec;
	if (glyph_id > 0xFFFF)
		return fail("glyph id out of range");
	f->e->face->outline_glyph((uint16_t)glyph_id, rec);
	*cmds = nullptr;
	if (!rec.v.empty()) {
		*cmds = (float *)std::malloc(rec.v.size() * sizeof(float));
		if (!*cmds)
			return fail("out of memory");
		std::memcpy(*cmds, rec.v.data(), rec.v.size() * sizeof(float));
	}
	return (int32_t)(rec.v.size() / 7);
}

int32_t vgb_font_outline_rings(const vgb_font *f, uint32_t gid, double **xy, uint32_t **ring_start, uint32_t *n_points)
{
	RingSet rs;
	RingBuilder rb(rs);
	if (gid <= 0xFFFF)
		f->e->face->outline_glyph((uint16_t)gid, rb);
	rb.finish();
	const size_t np = rs.point_count(), nr = rs.ring_count();
	*xy = (double *)std::malloc(sizeof(double) * 2 * (np ? np : 1));
	*ring_start = (uint32_t *)std::malloc(sizeof(uint32_t) * (nr + 1));
	for (size_t i = 0; i < np; ++i) {
		(*xy)[2 * i] = rs.points()[i].x;
		(*xy)[2 * i + 1] = rs.points()[i].y;
	}
	for (size_t r = 0; r < nr; ++r)
		(*ring_start)[r] = (uint32_t)rs.ring_begin(r);
	(*ring_start)[nr] = (uint32_t)np;
	*n_points = (uint32_t)np;
	return (int32_t)nr;
}

// ---- geometry --------------------------------------------------------------------------------------
size_t vgb_flatten_quad(const double s[2], const double c[2], const double e[2], double tol_sq, double *out_xy, size_t cap)
{
	RingSet rs;
	rs.open_add_quadratic_bezier(Point(s[0], s[1]), Point(c[0], c[1]), Point(e[0], e[1]), tol_sq);
	const size_t n = rs.open_len();
	rs.open_commit();
	for (size_t i = 0; i < n && i < cap; ++i) {
		out_xy[2 * i] = rs.points()[i].x;
		out_xy[2 * i + 1] = rs.points()[i].y;
	}
	return n;
}
size_t vgb_flatten_cubic(const double s[2], const double c1[2], const double c2[2], const double e[2], double tol_sq,
                         double *out_xy, size_t cap)
{
	RingSet rs;
	rs.open_add_cubic_bezier(Point(s[0], s[1]), Point(c1[0], c1[1]), Point(c2[0], c2[1]), Point(e[0], e[1]), tol_sq);
	const size_t n = rs.open_len();
	rs.open_commit();
	for (size_t i = 0; i < n && i < cap; ++i) {
		out_xy[2 * i] = rs.points()[i].x;
		out_xy[2 * i + 1] = rs.points()[i].y;
	}
	return n;
}
double vgb_segment_sqdist(double vx, double vy, double wx, double wy, double px, double py)
{
	return Segment{Point(vx, vy), Point(wx, wy)}.squared_distance_to_point(Point(px, py));
}
const char *vgb_name_to_id(const char *name, char *buf, size_t cap)
{
	const std::string id = FontManager::name_to_id(name);
	if (cap == 0)
		return buf;
	const size_t n = std::min(id.size(), cap - 1);
	std::memcpy(buf, id.data(), n);
	buf[n] = 0;
	return buf;
}

// ---- font naming + index files --------------------------------------------------------------------
static bool put_str(char *buf, size_t cap, const std::string &v)
{
	if (!buf || v.size() + 1 > cap)
		return false;
	std::memcpy(buf, v.c_str(), v.size() + 1);
	return true;
}
int vgb_parse_font_name(const char *family, const char *ps_name, char *out_family, char *out_style, uint16_t *out_weight,
                        char *out_width, size_t cap)
{
	std::string fam, style, width;
	uint16_t weight = 400;
	parse_font_name(family, ps_name, fam, style, weight, width);
	if (!put_str(out_family, cap, fam) || !put_str(out_style, cap, style) || !put_str(out_width, cap, width))
		return fail("buffer too small");
	*out_weight = weight;
	return 0;
}
size_t vgb_encode_codeblocks(const uint32_t *codepoints, size_t n, char *buf, size_t cap)
{
	const std::string s = encode_codeblocks(std::vector<uint32_t>(codepoints, codepoints + n));
	if (buf && cap) {
		const size_t k = std::min(s.size(), cap - 1);
		std::memcpy(buf, s.data(), k);
		buf[k] = 0;
	}
	return s.size();
}
int vgb_font_metadata(const vgb_font *f, char *name, char *family, char *style, uint16_t *weight, char *width, char *generated,
                      size_t cap)
{
	const FontMetadata &m = f->e->metadata;
	if (!put_str(name, cap, m.name) || !put_str(family, cap, m.family) || !put_str(style, cap, m.style) ||
	    !put_str(width, cap, m.width) || !put_str(generated, cap, m.generate_name()))
		return fail("buffer too small");
	*weight = m.weight;
	return 0;
}

// ---- renderer --------------------------------------------------------------------------------------
vgb_renderer *vgb_renderer_new(int dummy, int device, uint32_t n_slots)
{
	std::string err;
	auto r = Renderer::create(dummy != 0, device, n_slots, &err);
	if (!r) {
		g_err = err;
		return nullptr;
	}
	return new vgb_renderer{std::move(r)};
}
void vgb_renderer_free(vgb_renderer *r) { delete r; }
int vgb_renderer_is_dummy(const vgb_renderer *r) { return r->r->mode() == Renderer::Mode::Dummy; }
b200sdf_ctx *vgb_renderer_context(const vgb_renderer *r) { return r->r->context(); }

int vgb_renderer_render_glyph(const vgb_renderer *r, const vgb_font *f, uint32_t codepoint, vgb_glyph *out)
{
	std::unique_ptr<GlyphBatch> batch = r->r->new_batch();
	if (!batch->add_glyph(*f->e->face, codepoint))
		return batch->failed() ? fail(batch->failure()) : 0;
	std::string err;
	if (!r->r->render_batch(*batch, &err))
		return fail(err);
	fill_glyph(batch->take_glyph(0), (uint32_t)batch->total_segments(), out);
	return 1;
}

// ---- batch -----------------------------------------------------------------------------------------
vgb_batch *vgb_batch_new(const vgb_renderer *r) { return new vgb_batch{r->r->new_batch()}; }
void vgb_batch_free(vgb_batch *b) { delete b; }
void vgb_batch_clear(vgb_batch *b) { b->b->clear(); }
int vgb_batch_add_glyph(vgb_batch *b, const vgb_font *f, uint32_t codepoint)
{
	if (b->b->add_glyph(*f->e->face, codepoint))
		return 1;
	return b->b->failed() ? fail(b->b->failure()) : 0;
}
int vgb_batch_add_rings(vgb_batch *b, uint32_t id, int32_t x0, int32_t y0, uint32_t width, uint32_t height, const double *xy,
                        const uint32_t *ring_start, uint32_t n_rings)
{
	RingSet rs;
	for (uint32_t r = 0; r < n_rings; ++r) {
		for (uint32_t i = ring_start[r]; i < ring_start[r + 1]; ++i)
			rs.open_add(Point(xy[2 * i], xy[2 * i + 1]));
		rs.open_commit();
	}
	RenderResult fr;
	fr.x0 = x0;
	fr.y0 = y0;
	fr.x1 = x0 + (int32_t)width;
	fr.y1 = y0 + (int32_t)height;
	fr.width = width;
	fr.height = height;
	return b->b->add_rings(id, 0, fr, rs) ? 1 : fail("add_rings: out of memory");
}
uint32_t vgb_batch_glyph_count(const vgb_batch *b) { return (uint32_t)b->b->glyphs().size(); }
int vgb_batch_glyph_info(const vgb_batch *b, uint32_t i, vgb_batch_glyph *out)
{
	if (i >= b->b->glyphs().size())
		return fail("glyph index out of range");
	const BatchGlyph &g = b->b->glyphs()[i];
	std::memset(out, 0, sizeof(*out));
	out->id = g.id;
	out->advance = g.advance;
	out->has_bitmap = g.has_bitmap ? 1 : 0;
	if (g.has_bitmap) {
		const b200sdf_outline_job &j = b->b->jobs()[g.job];
		const PbfGlyph p = g.frame.into_pbf_glyph(g.id, g.advance);
		out->x0 = g.frame.x0;
		out->y0 = g.frame.y0;
		out->bm_width = g.frame.width;
		out->bm_height = g.frame.height;
		out->width = p.width;
		out->height = p.height;
		out->left = p.left;
		out->top = p.top;
		out->kind = j.kind;
		out->src_off = j.src_off;
		out->src_cnt = j.src_cnt;
		out->seg_cnt = j.seg_cnt;
		out->out_off = j.out_off;
	}
	return 0;
}
const b200sdf_segment *vgb_batch_segments(const vgb_batch *b, uint32_t *n_seg)
{
	*n_seg = b->b->segment_count();
	return b->b->segments();
}
const b200sdf_outline_job *vgb_batch_jobs(const vgb_batch *b, uint32_t *n_jobs)
{
	*n_jobs = b->b->job_count();
	return b->b->jobs();
}
const b200sdf_curve *vgb_batch_curves(const vgb_batch *b, uint32_t *n_curves)
{
	*n_curves = b->b->curve_count();
	return b->b->curves();
}
uint64_t vgb_batch_total_segments(const vgb_batch *b) { return b->b->total_segments(); }
uint32_t vgb_batch_fallback_glyphs(const vgb_batch *b) { return b->b->fallback_glyphs(); }
void vgb_renderer_set_flatten(vgb_renderer *r, int mode)
{
	r->r->set_flatten(mode == 2 ? Flatten::Glyf : mode ? Flatten::Device : Flatten::Host);
}
int vgb_renderer_flatten(const vgb_renderer *r)
{
	return r->r->flatten() == Flatten::Glyf ? 2 : r->r->flatten() == Flatten::Device ? 1 : 0;
}
int vgb_batch_finalize(const vgb_renderer *r, vgb_batch *b)
{
	std::string err;
	return b->b->finalize(*r->r, &err) ? 0 : fail(err);
}
const b200sdf_glyph_req *vgb_batch_requests(const vgb_batch *b, uint32_t *n)
{
	*n = b->b->mode() == Flatten::Glyf ? b->b->job_count() : 0;
	return b->b->reqs();
}
const b200sdf_glyph_part *vgb_batch_parts(const vgb_batch *b, uint32_t *n)
{
	*n = b->b->part_count();
	return b->b->parts();
}
uint32_t vgb_batch_curve_slots(const vgb_batch *b) { return b->b->curve_slots(); }
uint32_t vgb_batch_tile_cap(const vgb_batch *b) { return b->b->tile_cap(); }
uint64_t vgb_batch_est_cost(const vgb_batch *b) { return b->b->est_cost(); }
uint32_t vgb_batch_handed_back(const vgb_batch *b) { return b->b->handed_back(); }
uint32_t vgb_batch_path_glyphs(const vgb_batch *b) { return b->b->path_glyphs(); }
const uint8_t *vgb_batch_glyph_bitmap(const vgb_batch *b, uint32_t i, uint64_t *len)
{
	*len = 0;
	if (i >= b->b->glyphs().size())
		return nullptr;
	const BatchGlyph &g = b->b->glyphs()[i];
	if (!g.has_bitmap)
		return nullptr;
	*len = (uint64_t)g.frame.width * g.frame.height;
	return b->b->bitmap_of(g);
}
const uint8_t *vgb_batch_bitmaps(const vgb_batch *b, uint64_t *bytes)
{
	*bytes = b->b->bitmap_bytes();
	return b->b->bitmaps();
}
uint64_t vgb_batch_pairs(const vgb_batch *b) { return b->b->pairs(); }
int vgb_renderer_render_batch(const vgb_renderer *r, vgb_batch *b)
{
	std::string err;
	return r->r->render_batch(*b->b, &err) ? 0 : fail(err);
}
int vgb_renderer_submit_batch(const vgb_renderer *r, vgb_batch *b, uint64_t *ticket)
{
	std::string err;
	return r->r->submit_batch(*b->b, ticket, &err) ? 0 : fail(err);
}
int vgb_renderer_prepare_batch(const vgb_renderer *r, vgb_batch *b)
{
	std::string err;
	return r->r->prepare_batch(*b->b, &err) ? 0 : fail(err);
}
int vgb_renderer_poll_batch(const vgb_renderer *r, uint64_t ticket)
{
	std::string err;
	bool done = false;
	if (!r->r->poll_batch(ticket, &done, &err))
		return fail(err);
	return done ? 1 : 0;
}
int vgb_renderer_wait_batch(const vgb_renderer *r, uint64_t ticket)
{
	std::string err;
	return r->r->wait_batch(ticket, &err) ? 0 : fail(err);
}

// ---- writer ----------------------------------------------------------------------------------------
vgb_writer *vgb_writer_new_file(const char *folder) { return new vgb_writer(Writer::new_file(folder)); }
vgb_writer *vgb_writer_new_memory(void) { return new vgb_writer(Writer::new_memory()); }
vgb_writer *vgb_writer_new_tar(const char *path) { return new vgb_writer(Writer::new_tar(path)); }
vgb_writer *vgb_writer_new_tar_memory(void) { return new vgb_writer(Writer::new_tar_memory()); }
int vgb_writer_write_file(vgb_writer *w, const char *filename, const uint8_t *bytes, uint64_t len)
{
	std::string err;
	return w->w.write_file(filename, bytes, (size_t)len, &err) ? 0 : fail(err);
}
int vgb_writer_write_directory(vgb_writer *w, const char *dirname)
{
	std::string err;
	return w->w.write_directory(dirname, &err) ? 0 : fail(err);
}
int vgb_writer_finish(vgb_writer *w)
{
	std::string err;
	return w->w.finish(&err) ? 0 : fail(err);
}
const uint8_t *vgb_writer_tar_bytes(const vgb_writer *w, uint64_t *len)
{
	*len = w->w.tar_bytes().size();
	return w->w.tar_bytes().data();
}
void vgb_writer_free(vgb_writer *w)
{
	if (w) { // Drop for Writer (writer/mod.rs:83-96): best-effort finalisation
		std::string err;
		if (!w->w.finish(&err))
			std::fprintf(stderr, "warning: writer finalize failed during drop: %s\n", err.c_str());
	}
	delete w;
}
uint32_t vgb_writer_entry_count(const vgb_writer *w) { return (uint32_t)w->w.entries().size(); }
int vgb_writer_entry(const vgb_writer *w, uint32_t i, const char **name, int32_t *is_dir, const uint8_t **bytes, uint64_t *len)
{
	if (i >= w->w.entries().size())
		return fail("entry index out of range");
	const Writer::Entry &e = w->w.entries()[i];
	*name = e.name.c_str();
	*is_dir = e.is_dir ? 1 : 0;
	*bytes = e.bytes.data();
	*len = e.bytes.size();
	return 0;
}

// ---- manager ---------------------------------------------------------------------------------------
vgb_manager *vgb_manager_new(int parallel) { return new vgb_manager(parallel != 0); }
void vgb_manager_free(vgb_manager *m) { delete m; }
int vgb_manager_add_path(vgb_manager *m, const char *path)
{
	std::string err;
	return m->m.add_path(path, &err) ? 0 : fail(err);
}
size_t vgb_manager_font_file_names(const vgb_manager *m, const char *font_id, char *buf, size_t cap)
{
	std::string out;
	const auto &fonts = m->m.fonts();
	const auto it = fonts.find(font_id);
	if (it != fonts.end())
		for (const auto &f : it->second.files()) {
			if (!out.empty())
				out.push_back('\n');
			out += f->metadata.name;
		}
	if (buf && cap) {
		const size_t k = std::min(out.size(), cap - 1);
		std::memcpy(buf, out.data(), k);
		buf[k] = 0;
	}
	return out.size();
}
int vgb_manager_scan(vgb_manager *m, const char *path)
{
	std::string err;
	return m->m.scan(path, &err) ? 0 : fail(err);
}
int vgb_manager_add_font_with_name(vgb_manager *m, const char *name, const char *const *sources, uint32_t n)
{
	std::vector<std::string> v;
	for (uint32_t i = 0; i < n; ++i)
		v.emplace_back(sources[i]);
	std::string err;
	return m->m.add_font_with_name(name, v, &err) ? 0 : fail(err);
}
int vgb_manager_add_font_bytes_with_name(vgb_manager *m, const char *name, const uint8_t *data, size_t len)
{
	std::string err;
	return m->m.add_font_bytes_with_name(name, std::vector<uint8_t>(data, data + len), &err) ? 0 : fail(err);
}
uint32_t vgb_manager_font_count(const vgb_manager *m) { return (uint32_t)m->m.fonts().size(); }
const char *vgb_manager_font_id(const vgb_manager *m, uint32_t i)
{
	vgb_manager *mm = const_cast<vgb_manager *>(m);
	mm->ids.clear();
	for (const auto &kv : m->m.fonts())
		mm->ids.push_back(kv.first);
	return i < mm->ids.size() ? mm->ids[i].c_str() : nullptr;
}
int vgb_manager_block_population(const vgb_manager *m, const char *font_id, uint32_t out[256])
{
	auto it = m->m.fonts().find(font_id);
	if (it == m->m.fonts().end())
		return fail(std::string("unknown font id ") + font_id);
	const std::vector<GlyphBlock> blocks = it->second.get_blocks();
	for (size_t i = 0; i < 256; ++i)
		out[i] = (uint32_t)blocks[i].len();
	return 0;
}
int vgb_manager_render_block(const vgb_manager *m, const char *font_id, uint32_t block, const vgb_renderer *r, uint8_t **pbf,
                             uint64_t *len)
{
	auto it = m->m.fonts().find(font_id);
	if (it == m->m.fonts().end())
		return fail(std::string("unknown font id ") + font_id);
	if (block >= 256)
		return fail("block out of range");
	const std::vector<GlyphBlock> blocks = it->second.get_blocks();
	std::vector<uint8_t> out;
	std::string err;
	if (!blocks[block].render(font_id, *r->r, out, &err))
		return fail(err);
	*pbf = (uint8_t *)std::malloc(out.size() ? out.size() : 1);
	std::memcpy(*pbf, out.data(), out.size());
	*len = out.size();
	return 0;
}
int vgb_manager_render_glyphs(const vgb_manager *m, vgb_writer *w, const vgb_renderer *r, uint32_t shard, uint32_t n_shards,
                              int threads, vgb_stats *stats)
{
	std::string err;
	RenderStats st;
	if (!m->m.render_glyphs(w->w, *r->r, &err, &st, shard, n_shards, threads))
		return fail(err);
	if (stats) {
		stats->glyphs = st.glyphs;
		stats->bitmaps = st.bitmaps;
		stats->pixels = st.pixels;
		stats->segments = st.segments;
		stats->pairs = st.pairs;
		stats->pbf_bytes = st.pbf_bytes;
		stats->blocks = st.blocks;
		stats->outline_ns = st.outline_ns;
		stats->submit_ns = st.submit_ns;
		stats->wait_ns = st.wait_ns;
		stats->encode_ns = st.encode_ns;
		stats->write_ns = st.write_ns;
		stats->wall_ns = st.wall_ns;
		stats->submits = st.submits;
		stats->workers = st.workers;
		stats->handed_back = st.handed_back;
		stats->h2d_bytes = st.h2d_bytes;
		stats->cost_total = st.cost_total;
		stats->cost_shard = st.cost_shard;
	}
	return 0;
}
int vgb_manager_shard_owners(const vgb_manager *m, uint32_t n_shards, uint16_t *owner, size_t cap, uint64_t *loads)
{
	std::vector<uint16_t> o;
	std::vector<uint64_t> l;
	m->m.shard_owners(n_shards, o, &l);
	if (cap < o.size())
		return fail("shard_owners: buffer too small (fonts x 256 entries)");
	std::memcpy(owner, o.data(), o.size() * sizeof(uint16_t));
	if (loads)
		for (size_t k = 0; k < l.size(); ++k)
			loads[k] = l[k];
	return (int)o.size();
}
int vgb_manager_write_index_json(const vgb_manager *m, vgb_writer *w)
{
	std::string err;
	return m->m.write_index_json(w->w, &err) ? 0 : fail(err);
}

int vgb_manager_write_families_json(const vgb_manager *m, vgb_writer *w)
{
	std::string err;
	return m->m.write_families_json(w->w, &err) ? 0 : fail(err);
}

// ---- pbf decode ------------------------------------------------------------------------------------
int32_t vgb_pbf_decode(const uint8_t *data, size_t len, char *name, size_t name_cap, char *range, size_t range_cap,
                       vgb_glyph **glyphs)
{
	std::string n, r;
	std::vector<PbfGlyph> g;
	if (!pbf_decode(data, len, n, r, g))
		return fail("malformed glyphs PBF");
	if (name && name_cap) {
		const size_t k = std::min(n.size(), name_cap - 1);
		std::memcpy(name, n.data(), k);
		name[k] = 0;
	}
	if (range && range_cap) {
		const size_t k = std::min(r.size(), range_cap - 1);
		std::memcpy(range, r.data(), k);
		range[k] = 0;
	}
	*glyphs = (vgb_glyph *)std::calloc(g.size() ? g.size() : 1, sizeof(vgb_glyph));
	for (size_t i = 0; i < g.size(); ++i)
		fill_glyph(g[i], 0, &(*glyphs)[i]);
	return (int32_t)g.size();
}
void vgb_glyphs_free(vgb_glyph *glyphs, int32_t n)
{
	if (!glyphs)
		return;
	for (int32_t i = 0; i < n; ++i)
		std::free(glyphs[i].bitmap);
	std::free(glyphs);
}

} // extern "C"
