// UNVERIFIED: written without a Rust toolchain (no cargo/rustc in the build image); see INTEGRATION.md.
// b200sdf_sys.rs — raw bindings (bindgen-equivalent of include/b200sdf.h)
#[repr(C)] pub struct B200sdfCtx { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy)] pub struct Segment { pub x0: f32, pub y0: f32, pub x1: f32, pub y1: f32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct GlyphJob { pub seg_off: u32, pub seg_cnt: u32, pub width: u32, pub height: u32, pub out_off: u64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct Curve { pub sx: f32, pub sy: f32, pub cx: f32, pub cy: f32, pub ex: f32, pub ey: f32, pub seg_off: u32, pub depth: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct OutlineJob { pub kind: u32, pub src_off: u32, pub src_cnt: u32, pub seg_cnt: u32,
    pub width: u32, pub height: u32, pub x0: i32, pub y0: i32, pub scale: f64, pub dx: f64, pub out_off: u64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct TileJob { pub seg_off: u32, pub seg_cnt: u32, pub out_off: u64, pub width: u16, pub height: u16,
    pub tx0: u16, pub ty0: u16, pub ntx: u16, pub nty: u16, pub job: u32 }   // 32 bytes, opaque to callers
#[link(name = "b200sdf")]
extern "C" {
    pub fn b200sdf_create(device: i32, n_slots: u32, out: *mut *mut B200sdfCtx) -> i32;
    pub fn b200sdf_destroy(ctx: *mut B200sdfCtx);
    pub fn b200sdf_last_error(ctx: *const B200sdfCtx) -> *const std::os::raw::c_char;
    pub fn b200sdf_alloc_pinned(bytes: usize) -> *mut std::ffi::c_void;
    pub fn b200sdf_free_pinned(p: *mut std::ffi::c_void);
    pub fn b200sdf_submit(ctx: *mut B200sdfCtx, segs: *const Segment, n_seg: u32, jobs: *const GlyphJob, n_jobs: u32,
                          out: *mut u8, out_bytes: u64, ticket: *mut u64) -> i32;
    pub fn b200sdf_submit_outlines(ctx: *mut B200sdfCtx, curves: *const Curve, n_curves: u32, segs: *const Segment, n_seg: u32,
                                   jobs: *const OutlineJob, n_jobs: u32, out: *mut u8, out_bytes: u64, ticket: *mut u64) -> i32;
    pub fn b200sdf_wait(ctx: *mut B200sdfCtx, ticket: u64) -> i32;
    /// non-blocking wait: 1 = finished (ticket consumed), 0 = still running, < 0 = error
    pub fn b200sdf_poll(ctx: *mut B200sdfCtx, ticket: u64) -> i32;
    /// tiles planned beforehand (b200sdf_plan_outline_tiles over the same jobs): the submitting thread only enqueues
    pub fn b200sdf_plan_outline_tiles(jobs: *const OutlineJob, n_jobs: u32, n_curves: u32, n_seg: u32, out_bytes: u64,
                                      tiles: *mut TileJob, cap: u32, n_tiles: *mut u32, pairs: *mut u64) -> i32;
    pub fn b200sdf_submit_planned(ctx: *mut B200sdfCtx, curves: *const Curve, n_curves: u32, segs: *const Segment, n_seg: u32,
                                  jobs: *const OutlineJob, n_jobs: u32, tiles: *const TileJob, n_tiles: u32,
                                  out: *mut u8, out_bytes: u64, ticket: *mut u64) -> i32;
}
