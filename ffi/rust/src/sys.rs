//! UNVERIFIED (no rustc in the build image).  Raw bindings: the bindgen-equivalent of include/b200sdf.h (ABI version 2).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct b200sdf_ctx {
    _private: [u8; 0],
}

pub const B200SDF_ABI_VERSION: c_int = 2;
pub const B200SDF_KIND_CURVES: u32 = 0;
pub const B200SDF_KIND_SEGMENTS: u32 = 1;
pub const B200SDF_KIND_GLYF: u32 = 2;
/// host-recorded outline with cubic curves: head + tail records, flattened by the decode kernel
pub const B200SDF_KIND_PATH: u32 = 3;
pub const B200SDF_CURVE_CUBIC: u32 = 0x8000_0000;
pub const B200SDF_CURVE_TAIL: u32 = 0x4000_0000;
pub const B200SDF_CUBIC_STACK: usize = 24;
pub const B200SDF_GLYPH_OK: u32 = 0;
pub const B200SDF_GLYPH_EMPTY: u32 = 1;
pub const B200SDF_GLYPH_NEEDS_HOST: u32 = 2;
pub const B200SDF_GLYPH_BAD_REQUEST: u32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_segment {
    pub x0: f32,
    pub y0: f32,
    pub x1: f32,
    pub y1: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_glyph_job {
    pub seg_off: u32,
    pub seg_cnt: u32,
    pub width: u32,
    pub height: u32,
    pub out_off: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_curve {
    pub sx: f32,
    pub sy: f32,
    pub cx: f32,
    pub cy: f32,
    pub ex: f32,
    pub ey: f32,
    pub seg_off: u32,
    pub depth: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_outline_job {
    pub kind: u32,
    pub src_off: u32,
    pub src_cnt: u32,
    pub seg_cnt: u32,
    pub width: u32,
    pub height: u32,
    pub x0: i32,
    pub y0: i32,
    pub scale: f64,
    pub dx: f64,
    pub out_off: u64,
}

/// One simple-glyph record of a glyph request (72-byte request + 20 bytes per part is all the device needs per glyph).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_glyph_part {
    pub font: u32,
    pub glyf_off: u32,
    pub glyf_len: u32,
    pub ox: f32,
    pub oy: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_glyph_req {
    pub kind: u32,
    pub src_off: u32,
    pub src_cnt: u32,
    pub seg_cnt: u32,
    pub width: u32,
    pub height: u32,
    pub x0: i32,
    pub y0: i32,
    pub scale: f64,
    pub dx: f64,
    pub out_off: u64,
    pub out_cap: u32,
    pub curve_off: u32,
    pub curve_cap: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200sdf_glyph_frame {
    pub x0: i32,
    pub y0: i32,
    pub width: u32,
    pub height: u32,
    pub seg_cnt: u32,
    pub status: u32,
}

/// One batch of a multi-batch submission (`b200sdf_submit_glyph_batches`); all buffers from `b200sdf_alloc_pinned`.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200sdf_glyph_batch {
    pub reqs: *const b200sdf_glyph_req,
    pub n_reqs: u32,
    pub parts: *const b200sdf_glyph_part,
    pub n_parts: u32,
    pub curves: *const b200sdf_curve,
    pub n_curves: u32,
    pub segs: *const b200sdf_segment,
    pub n_seg: u32,
    pub curve_slots: u32,
    pub tile_cap: u32,
    pub frames: *mut b200sdf_glyph_frame,
    pub out: *mut u8,
    pub out_bytes: u64,
}

pub const B200SDF_MAX_BATCHES: u32 = 16;

extern "C" {
    pub fn b200sdf_abi_version() -> c_int;
    pub fn b200sdf_device_count() -> c_int;
    pub fn b200sdf_create(device: c_int, n_slots: u32, out: *mut *mut b200sdf_ctx) -> c_int;
    pub fn b200sdf_destroy(ctx: *mut b200sdf_ctx);
    pub fn b200sdf_last_error(ctx: *const b200sdf_ctx) -> *const c_char;
    pub fn b200sdf_alloc_pinned(bytes: usize) -> *mut c_void;
    pub fn b200sdf_free_pinned(p: *mut c_void);

    // segment level: renderer_precise's own signature (src/render/renderer_precise.rs:8)
    pub fn b200sdf_submit(ctx: *mut b200sdf_ctx, segs: *const b200sdf_segment, n_seg: u32, jobs: *const b200sdf_glyph_job,
                          n_jobs: u32, out: *mut u8, out_bytes: u64, ticket: *mut u64) -> c_int;
    pub fn b200sdf_wait(ctx: *mut b200sdf_ctx, ticket: u64) -> c_int;
    pub fn b200sdf_poll(ctx: *mut b200sdf_ctx, ticket: u64) -> c_int;
    pub fn b200sdf_render(ctx: *mut b200sdf_ctx, segs: *const b200sdf_segment, n_seg: u32, jobs: *const b200sdf_glyph_job,
                          n_jobs: u32, out: *mut u8, out_bytes: u64) -> c_int;

    // outline level: the RingBuilder callbacks before flattening (src/render/ring_builder.rs:67-117)
    pub fn b200sdf_submit_outlines(ctx: *mut b200sdf_ctx, curves: *const b200sdf_curve, n_curves: u32, segs: *const b200sdf_segment,
                                   n_seg: u32, jobs: *const b200sdf_outline_job, n_jobs: u32, out: *mut u8, out_bytes: u64,
                                   ticket: *mut u64) -> c_int;

    // glyph level: Face::outline_glyph itself (src/render/renderer.rs:109-111) happens on the device
    pub fn b200sdf_font_upload(ctx: *mut b200sdf_ctx, glyf: *const u8, len: u64, handle: *mut u32) -> c_int;
    pub fn b200sdf_glyph_tile_bound(width: u32, height: u32) -> u32;
    pub fn b200sdf_submit_glyphs(ctx: *mut b200sdf_ctx, reqs: *const b200sdf_glyph_req, n_reqs: u32, parts: *const b200sdf_glyph_part,
                                 n_parts: u32, curves: *const b200sdf_curve, n_curves: u32, segs: *const b200sdf_segment, n_seg: u32,
                                 curve_slots: u32, tile_cap: u32, est_cost: u64, frames: *mut b200sdf_glyph_frame, out: *mut u8,
                                 out_bytes: u64, ticket: *mut u64) -> c_int;
    pub fn b200sdf_submit_glyph_batches(ctx: *mut b200sdf_ctx, batches: *const b200sdf_glyph_batch, n_batches: u32, est_cost: u64,
                                        ticket: *mut u64) -> c_int;
    pub fn b200sdf_reserve(ctx: *mut b200sdf_ctx) -> c_int;
    pub fn b200sdf_reserve_glyphs(ctx: *mut b200sdf_ctx, n_reqs: u32, n_seg: u32, curve_slots: u32, tile_cap: u32) -> c_int;
}
