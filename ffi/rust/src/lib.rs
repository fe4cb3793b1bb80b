//! UNVERIFIED: written without a Rust toolchain (this build image has no cargo / rustc).  The tested caller of the same
//! C ABI is the C++ host mirror (versatiles_glyphs_rs_b200/csrc/host/, exercised by tests/ through ctypes); this crate
//! is what a maintainer of versatiles_glyphs would add.  Every function cites the reference code it stands in for.
//!
//! `CudaRenderer` is the third arm of `RendererMode` (reference src/render/renderer.rs:11-15, matched at :140-143):
//! `GlyphBlock::render` (src/font/glyph_block.rs:69-80) appends its code points to a `GlyphBatch` and flushes it with
//! one `submit`; the bitmaps and the integer frames come back from the device.  See patches/renderer_mode_cuda.patch.
pub mod sys;

use anyhow::{anyhow, bail, Result};
use std::collections::HashMap;
use std::ffi::CStr;
use std::sync::{Arc, Mutex};
use ttf_parser::{Face, GlyphId, Tag};

pub const GLYPH_SIZE: i32 = 24; // src/render/mod.rs:52
pub const BUFFER: i32 = 3; // src/render/mod.rs:58
const MAX_POINTS: u32 = 2048; // B200SDF_GLYF_MAX_POINTS

/// What `Renderer::render_glyph` returns (src/protobuf/glyph.rs:10-41), minus the prost derive.
#[derive(Debug, Clone)]
pub struct RenderedGlyph {
    pub id: u32,
    pub bitmap: Option<Vec<u8>>,
    pub width: u32,
    pub height: u32,
    pub left: i32,
    pub top: i32,
    pub advance: u32,
}

struct Ctx(*mut sys::b200sdf_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::b200sdf_destroy(self.0) }
    }
}

/// One context <-> one GPU; `Clone + Send + Sync` like the reference's `Renderer` (renderer.rs:17-21).
#[derive(Clone)]
pub struct CudaRenderer {
    ctx: Arc<Ctx>,
    fonts: Arc<Mutex<HashMap<usize, u32>>>, // address of the font's data -> handle of its glyf table on the device
}

fn check(ctx: *mut sys::b200sdf_ctx, rc: i32, what: &str) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::b200sdf_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(anyhow!("{what}: error {rc}: {msg}"))
}

impl CudaRenderer {
    /// `Renderer::new_cuda(device)`: fails when no sm_100 GPU is usable — there is no CPU fallback in the library; the
    /// caller keeps `RendererMode::Precise` in that case.
    pub fn new(device: i32, batches_in_flight: u32) -> Result<Self> {
        if unsafe { sys::b200sdf_abi_version() } != sys::B200SDF_ABI_VERSION {
            bail!("libb200sdf.so has a different ABI version");
        }
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::b200sdf_create(device, batches_in_flight, &mut ctx) };
        if rc != 0 {
            bail!("b200sdf_create({device}) failed with {rc}: a B200 (sm_100) GPU is required");
        }
        Ok(Self { ctx: Arc::new(Ctx(ctx)), fonts: Arc::new(Mutex::new(HashMap::new())) })
    }

    /// Makes the face's `glyf` table resident in HBM (once per face; FontFileEntry::new is the place to call it,
    /// src/font/file_entry.rs:32-56).  Faces without glyf outlines (CFF) return None: their glyphs are recorded on the host.
    pub fn font_handle(&self, face: &Face) -> Result<Option<u32>> {
        let Some(glyf) = face.raw_face().table(Tag::from_bytes(b"glyf")) else { return Ok(None) };
        let key = glyf.as_ptr() as usize;
        let mut map = self.fonts.lock().map_err(|_| anyhow!("font table lock poisoned"))?;
        if let Some(h) = map.get(&key) {
            return Ok(Some(*h));
        }
        let mut h = 0u32;
        let rc = unsafe { sys::b200sdf_font_upload(self.ctx.0, glyf.as_ptr(), glyf.len() as u64, &mut h) };
        check(self.ctx.0, rc, "b200sdf_font_upload")?;
        map.insert(key, h);
        Ok(Some(h))
    }

    pub fn new_batch(&self) -> GlyphBatch {
        GlyphBatch::default()
    }

    /// `GlyphBlock::render`'s flush: one submission, then wait.  (A pipeline keeps several batches in flight: call
    /// `submit` from one thread and `b200sdf_poll` the tickets — csrc/host/pipeline.cc is that pipeline in C++.)
    pub fn render(&self, batch: &mut GlyphBatch) -> Result<Vec<RenderedGlyph>> {
        let ticket = self.submit(batch)?;
        check(self.ctx.0, unsafe { sys::b200sdf_wait(self.ctx.0, ticket) }, "b200sdf_wait")?;
        batch.finish()
    }

    pub fn submit(&self, batch: &mut GlyphBatch) -> Result<u64> {
        batch.frames.resize(batch.reqs.len() + 1, sys::b200sdf_glyph_frame::default());
        batch.out.resize(batch.out_bytes as usize + 16, 0);
        let mut ticket = 0u64;
        let rc = unsafe {
            sys::b200sdf_submit_glyphs(
                self.ctx.0, batch.reqs.as_ptr(), batch.reqs.len() as u32, batch.parts.as_ptr(), batch.parts.len() as u32,
                std::ptr::null(), 0, std::ptr::null(), 0, batch.curve_slots, batch.tile_cap.max(1), batch.est_cost,
                batch.frames.as_mut_ptr(), batch.out.as_mut_ptr(), batch.out_bytes, &mut ticket,
            )
        };
        check(self.ctx.0, rc, "b200sdf_submit_glyphs")?;
        Ok(ticket)
    }
}

impl CudaRenderer {
    /// Several batches in ONE submission (one decode launch, one SDF launch, one ticket): what a pipeline does with the
    /// batches that queued up while the previous submission was being made.  With more than one batch every array must
    /// live in memory from `b200sdf_alloc_pinned` (the library refuses pageable arrays here), so this is for hosts that
    /// build their batches in pinned buffers, as csrc/host/pipeline.cc does.
    pub fn submit_group(&self, batches: &mut [&mut GlyphBatch]) -> Result<u64> {
        let mut descs = Vec::with_capacity(batches.len());
        let mut est = 0u64;
        for b in batches.iter_mut() {
            b.frames.resize(b.reqs.len() + 1, sys::b200sdf_glyph_frame::default());
            b.out.resize(b.out_bytes as usize + 16, 0);
            est = est.max(b.est_cost);
            descs.push(sys::b200sdf_glyph_batch {
                reqs: b.reqs.as_ptr(), n_reqs: b.reqs.len() as u32, parts: b.parts.as_ptr(), n_parts: b.parts.len() as u32,
                curves: std::ptr::null(), n_curves: 0, segs: std::ptr::null(), n_seg: 0, curve_slots: b.curve_slots,
                tile_cap: b.tile_cap.max(1), frames: b.frames.as_mut_ptr(), out: b.out.as_mut_ptr(), out_bytes: b.out_bytes,
            });
        }
        let mut ticket = 0u64;
        let rc = unsafe { sys::b200sdf_submit_glyph_batches(self.ctx.0, descs.as_ptr(), descs.len() as u32, est, &mut ticket) };
        check(self.ctx.0, rc, "b200sdf_submit_glyph_batches")?;
        Ok(ticket)
    }

    /// Size every slot's device scratch for submissions of up to these totals, once, so that a long run never allocates
    /// device memory in mid-flight (`b200sdf_reserve_glyphs`).
    pub fn reserve(&self, glyphs: u32, segments: u32, curve_slots: u32, tile_jobs: u32) -> Result<()> {
        check(self.ctx.0, unsafe { sys::b200sdf_reserve_glyphs(self.ctx.0, glyphs, segments, curve_slots, tile_jobs) },
              "b200sdf_reserve_glyphs")
    }
}

struct Pending {
    id: u32,
    advance: u32,
    req: Option<usize>, // None: PbfGlyph::empty (renderer.rs:118-120)
}

/// The flat request buffer of one GlyphBlock ("packed into a flat buffer per GlyphBlock and uploaded once").
/// Plain `Vec`s take the staged-copy path; allocate the arrays with `b200sdf_alloc_pinned` to have the device read the
/// requests and write frames and bitmaps in place across PCIe (≈ 15 % faster end to end).
#[derive(Default)]
pub struct GlyphBatch {
    reqs: Vec<sys::b200sdf_glyph_req>,
    parts: Vec<sys::b200sdf_glyph_part>,
    frames: Vec<sys::b200sdf_glyph_frame>,
    out: Vec<u8>,
    out_bytes: u64,
    curve_slots: u32,
    tile_cap: u32,
    est_cost: u64,
    pending: Vec<Pending>,
    /// glyphs the device cannot take (scaled / rotated components, CFF): render these with RendererMode::Precise
    pub host_glyphs: Vec<u32>,
}

struct Part {
    off: u32,
    len: u32,
    ox: f32,
    oy: f32,
    points: u32,
    bbox: [i16; 4],
}

fn be16(d: &[u8], o: usize) -> Option<u16> {
    Some(u16::from_be_bytes([*d.get(o)?, *d.get(o + 1)?]))
}

/// loca lookup (ttf-parser's loca::Table::glyph_range): None for an empty or out-of-range glyph
fn glyph_range(face: &Face, glyf_len: usize, gid: u16) -> Option<(usize, usize)> {
    let raw = face.raw_face();
    let loca = raw.table(Tag::from_bytes(b"loca"))?;
    let head = raw.table(Tag::from_bytes(b"head"))?;
    let long = i16::from_be_bytes([*head.get(50)?, *head.get(51)?]) != 0;
    let g = gid as usize;
    let (a, b) = if long {
        let rd = |i: usize| Some(u32::from_be_bytes([*loca.get(i)?, *loca.get(i + 1)?, *loca.get(i + 2)?, *loca.get(i + 3)?]) as usize);
        (rd(4 * g)?, rd(4 * g + 4)?)
    } else {
        (2 * be16(loca, 2 * g)? as usize, 2 * be16(loca, 2 * g + 2)? as usize)
    };
    if b <= a || b > glyf_len {
        return None;
    }
    Some((a, b - a))
}

/// The composite walk of ttf-parser's glyf::Table::outline with the simple-glyph arm replaced by "remember the record"
/// (csrc/host/face.cc Face::parts_impl).  Ok(false) = a component carries more than a translation: host glyph.
fn collect_parts(face: &Face, glyf: &[u8], off: usize, len: usize, depth: u8, t: [f32; 6], out: &mut Vec<Part>) -> bool {
    if depth >= 32 || len < 10 {
        return true;
    }
    let g = &glyf[off..off + len];
    let n_contours = be16(g, 0).unwrap() as i16;
    if n_contours > 0 {
        let nc = n_contours as usize;
        if 10 + 2 * nc + 2 > len {
            return true;
        }
        let points = be16(g, 10 + 2 * (nc - 1)).unwrap() as u32 + 1;
        if points == 1 {
            return true;
        }
        if points > MAX_POINTS {
            return false;
        }
        let b = |o| be16(g, o).unwrap() as i16;
        out.push(Part { off: off as u32, len: len as u32, ox: t[4], oy: t[5], points, bbox: [b(2), b(4), b(6), b(8)] });
    } else if n_contours < 0 {
        let mut pos = 10usize;
        loop {
            let (Some(fl), Some(child)) = (be16(g, pos), be16(g, pos + 2)) else { return true };
            pos += 4;
            let mut ct = [1f32, 0., 0., 1., 0., 0.]; // a b c d e f
            if fl & 0x0001 != 0 {
                let (Some(a1), Some(a2)) = (be16(g, pos), be16(g, pos + 2)) else { return true };
                if fl & 0x0002 != 0 {
                    ct[4] = a1 as i16 as f32;
                    ct[5] = a2 as i16 as f32;
                }
                pos += 4;
            } else {
                let (Some(&a1), Some(&a2)) = (g.get(pos), g.get(pos + 1)) else { return true };
                if fl & 0x0002 != 0 {
                    ct[4] = a1 as i8 as f32;
                    ct[5] = a2 as i8 as f32;
                }
                pos += 2;
            }
            let f2 = |o: usize| be16(g, o).map(|v| v as i16 as f32 / 16384.0);
            if fl & 0x0080 != 0 {
                let (Some(a), Some(b), Some(c), Some(d)) = (f2(pos), f2(pos + 2), f2(pos + 4), f2(pos + 6)) else { return true };
                ct[0] = a; ct[1] = b; ct[2] = c; ct[3] = d;
                pos += 8;
            } else if fl & 0x0040 != 0 {
                let (Some(a), Some(d)) = (f2(pos), f2(pos + 2)) else { return true };
                ct[0] = a; ct[3] = d;
                pos += 4;
            } else if fl & 0x0008 != 0 {
                let Some(a) = f2(pos) else { return true };
                ct[0] = a; ct[3] = a;
                pos += 2;
            }
            if let Some((co, cl)) = glyph_range(face, glyf.len(), child) {
                // Transform::combine(t, ct), f32 like ttf-parser's
                let ts = [
                    t[0] * ct[0] + t[2] * ct[1], t[1] * ct[0] + t[3] * ct[1], t[0] * ct[2] + t[2] * ct[3], t[1] * ct[2] + t[3] * ct[3],
                    t[0] * ct[4] + t[2] * ct[5] + t[4], t[1] * ct[4] + t[3] * ct[5] + t[5],
                ];
                if !(ts[0] == 1.0 && ts[1] == 0.0 && ts[2] == 0.0 && ts[3] == 1.0) {
                    return false;
                }
                if !collect_parts(face, glyf, co, cl, depth + 1, ts, out) {
                    return false;
                }
            }
            if fl & 0x0020 == 0 {
                break;
            }
        }
    }
    true
}

impl GlyphBatch {
    /// First half of `Renderer::render_glyph` (renderer.rs:103-116): char check, cmap lookup, advance.  The outline itself
    /// is NOT walked: the request names the glyf records and the device does the rest.  Returns false for `None`
    /// (skip: glyph_block.rs:74-76).
    pub fn add_glyph(&mut self, r: &CudaRenderer, face: &Face, index: u32) -> Result<bool> {
        let Some(cp) = char::from_u32(index) else { return Ok(false) }; // :104
        let Some(gid) = face.glyph_index(cp) else { return Ok(false) }; // :106
        let scale = GLYPH_SIZE as f64 / face.units_per_em() as f64; // :107
        let advance_float = face.glyph_hor_advance(gid).unwrap_or(0) as f64 * scale * 0.95; // :115
        let advance = advance_float.round() as u32; // :116
        let dx = (advance as f64 - advance_float) / 2.0; // :130-131
        let Some(font) = r.font_handle(face)? else {
            self.host_glyphs.push(index);
            return Ok(true);
        };
        let glyf = face.raw_face().table(Tag::from_bytes(b"glyf")).unwrap();
        let mut parts = Vec::new();
        let Some((off, len)) = glyph_range(face, glyf.len(), gid.0) else {
            self.pending.push(Pending { id: index, advance, req: None }); // no outline: PbfGlyph::empty
            return Ok(true);
        };
        if !collect_parts(face, glyf, off, len, 0, [1., 0., 0., 1., 0., 0.], &mut parts) {
            self.host_glyphs.push(index);
            return Ok(true);
        }
        if parts.is_empty() {
            self.pending.push(Pending { id: index, advance, req: None });
            return Ok(true);
        }
        // the bitmap slot: prepare_glyph (renderer.rs:64-91) over the records' header boxes, one pixel of slack; the
        // device computes the real frame and hands the glyph back (NEEDS_HOST) if it does not fit
        let (mut x0, mut y0, mut x1, mut y1) = (f64::INFINITY, f64::INFINITY, f64::NEG_INFINITY, f64::NEG_INFINITY);
        let mut points = 0u32;
        let src_off = self.parts.len() as u32;
        for p in &parts {
            x0 = x0.min(p.bbox[0] as f64 + p.ox as f64);
            y0 = y0.min(p.bbox[1] as f64 + p.oy as f64);
            x1 = x1.max(p.bbox[2] as f64 + p.ox as f64);
            y1 = y1.max(p.bbox[3] as f64 + p.oy as f64);
            points += p.points;
            self.parts.push(sys::b200sdf_glyph_part { font, glyf_off: p.off, glyf_len: p.len, ox: p.ox, oy: p.oy });
        }
        let w = ((x1 * scale + dx).ceil() - (x0 * scale + dx).floor() + 2.0 * BUFFER as f64 + 2.0).clamp(8.0, 4096.0) as u32;
        let h = ((y1 * scale).ceil() - (y0 * scale).floor() + 2.0 * BUFFER as f64 + 2.0).clamp(8.0, 4096.0) as u32;
        let out_off = (self.out_bytes + 15) & !15;
        self.reqs.push(sys::b200sdf_glyph_req {
            kind: sys::B200SDF_KIND_GLYF, src_off, src_cnt: parts.len() as u32, scale, dx, out_off, out_cap: w * h,
            curve_off: self.curve_slots, curve_cap: points, ..Default::default()
        });
        self.pending.push(Pending { id: index, advance, req: Some(self.reqs.len() - 1) });
        self.out_bytes = out_off + (w * h) as u64;
        self.curve_slots += points;
        self.tile_cap += unsafe { sys::b200sdf_glyph_tile_bound(w, h) };
        self.est_cost += ((w as u64 + 3) / 4) * ((h as u64 + 3) / 4) * (points as u64 * 8 + 8);
        Ok(true)
    }

    /// Second half of `render_glyph` (renderer.rs:133-148, result.rs:66-76) from the frames the device wrote.
    fn finish(&mut self) -> Result<Vec<RenderedGlyph>> {
        let mut out = Vec::with_capacity(self.pending.len());
        for p in &self.pending {
            let empty = RenderedGlyph { id: p.id, bitmap: None, width: 0, height: 0, left: 0, top: 0, advance: p.advance };
            let Some(k) = p.req else {
                out.push(empty);
                continue;
            };
            let f = self.frames[k];
            match f.status {
                sys::B200SDF_GLYPH_OK => {
                    let o = self.reqs[k].out_off as usize;
                    let n = (f.width * f.height) as usize;
                    let y1 = f.y0 + f.height as i32 - GLYPH_SIZE; // renderer.rs:146
                    out.push(RenderedGlyph {
                        id: p.id, bitmap: Some(self.out[o..o + n].to_vec()), width: f.width - 2 * BUFFER as u32,
                        height: f.height - 2 * BUFFER as u32, left: f.x0 + BUFFER, top: y1 - BUFFER, advance: p.advance,
                    });
                }
                sys::B200SDF_GLYPH_EMPTY => out.push(empty), // renderer.rs:118-120 / :133-137
                sys::B200SDF_GLYPH_NEEDS_HOST => self.host_glyphs.push(p.id), // caller renders it with RendererMode::Precise
                s => bail!("glyph request for U+{:04X} rejected by the device (status {s})", p.id),
            }
        }
        Ok(out)
    }
}
