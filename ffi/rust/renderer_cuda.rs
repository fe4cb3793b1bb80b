// UNVERIFIED: written without a Rust toolchain (no cargo/rustc in the build image); see INTEGRATION.md.
// renderer.rs — the patch (shape only)
enum RendererMode { Precise, Dummy, Cuda(std::sync::Arc<CudaContext>) }      // renderer.rs:11-15

impl Renderer {
    pub fn new_cuda(device: i32) -> anyhow::Result<Self> { /* b200sdf_create(device, 2 * rayon::current_num_threads(), ..) */ }

    /// First half of render_glyph (renderer.rs:103-137), unchanged arithmetic: glyph index, outline -> rings,
    /// advance, scale + translate, prepare_glyph.  Instead of calling renderer_precise it appends the glyph's
    /// segments (x - x0, y - y0 narrowed to f32) and one GlyphJob to `batch` and returns the metrics.
    pub fn prepare_into(&self, face: &Face, index: u32, batch: &mut CudaBatch) -> Option<PendingGlyph> { .. }
}

// glyph_block.rs:69-80 — GlyphBlock::render with RendererMode::Cuda
let mut batch = CudaBatch::new();                                 // pinned Vec-likes from b200sdf_alloc_pinned
let pending: Vec<PendingGlyph> = self.glyphs.iter()
    .filter_map(|(i, f)| renderer.prepare_into(&f.face, self.start_index + *i as u32, &mut batch)).collect();
let ticket = batch.submit(&ctx)?;                                 // b200sdf_submit, returns immediately
ctx.wait(ticket)?;                                                // b200sdf_wait
for g in pending {                                                // result.rs:66-76 unchanged
    glyphs.push(g.into_pbf_glyph(batch.bitmap(g.job).to_vec()));  // Vec<u8; W*H>, top row first
}
