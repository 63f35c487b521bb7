//! Raw bindings to `include/ce_gpu.h`, one declaration per exported symbol the metric path needs.
//! Struct layouts are checked on the C side by tests/test_abi.py (5 / 48 / 56 bytes).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct ce_ctx { _p: [u8; 0] }
#[repr(C)] pub struct ce_ref { _p: [u8; 0] }

/// = codec_eval::metrics::MetricConfig (src/metrics/mod.rs:45-63), same field order
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ce_metric_config { pub dssim: u8, pub ssimulacra2: u8, pub butteraugli: u8, pub psnr: u8, pub xyb_roundtrip: u8 }

#[repr(C)] #[derive(Clone, Copy)]
pub struct ce_pair {
    pub reference: *const u8, pub dist: *const u8,
    pub ref_len: usize, pub dist_len: usize,
    pub width: u32, pub height: u32, pub ref_id: u32, pub reserved: u32,
}

/// = MetricResult (src/metrics/mod.rs:139-149) + raw SSE + libjxl 3-norm; `valid` bits <-> Some(..)
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ce_result {
    pub status: i32, pub valid: u32, pub sse: u64,
    pub dssim: f64, pub ssimulacra2: f64, pub butteraugli: f64, pub psnr: f64, pub butteraugli_pnorm3: f64,
}

pub const CE_OK: c_int = 0;
pub const CE_ERR_DIMENSION_MISMATCH: c_int = 1;   // Error::DimensionMismatch  (src/error.rs:33)
pub const CE_ERR_METRIC_CALCULATION: c_int = 2;   // Error::MetricCalculation  (src/error.rs:42)
pub const CE_ERR_INVALID_ARGUMENT: c_int = 3;
pub const CE_ERR_CUDA: c_int = 4;
pub const CE_ERR_OUT_OF_MEMORY: c_int = 5;

// ce_result.valid bits  <->  Some(..) in MetricResult
pub const CE_VALID_DSSIM: u32 = 1;
pub const CE_VALID_SSIMULACRA2: u32 = 2;
pub const CE_VALID_BUTTERAUGLI: u32 = 4;
pub const CE_VALID_PSNR: u32 = 8;

#[link(name = "ce_gpu")]
extern "C" {
    pub fn ce_ctx_create(out: *mut *mut ce_ctx, device: c_int, workspace_bytes: usize) -> c_int;
    pub fn ce_ctx_destroy(ctx: *mut ce_ctx);
    pub fn ce_ctx_set_stream(ctx: *mut ce_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn ce_last_error(ctx: *const ce_ctx) -> *const c_char;
    pub fn ce_launch_count(ctx: *const ce_ctx) -> u64;
    pub fn ce_version() -> *const c_char;
    // ---- pinned host memory (opt-in): register the session's decode buffers once, or allocate a pinned ring ----
    pub fn ce_host_register(ctx: *mut ce_ctx, ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn ce_host_unregister(ctx: *mut ce_ctx, ptr: *mut c_void) -> c_int;
    pub fn ce_host_alloc(ctx: *mut ce_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ce_host_free(ctx: *mut ce_ctx, ptr: *mut c_void);
    pub fn ce_sub_batch_capacity(ctx: *mut ce_ctx, cfg: *const ce_metric_config, width: u32, height: u32,
                                 pairs: *mut usize) -> c_int;
    pub fn ce_evaluate_batch(ctx: *mut ce_ctx, pairs: *const ce_pair, n: usize, cfg: *const ce_metric_config,
                             intensity_target: f32, out: *mut ce_result) -> c_int;
    pub fn ce_evaluate_batch_device_grouped(ctx: *mut ce_ctx, d_ref: *const u8, n_ref: usize, d_dist: *const u8, n: usize,
                                            ref_index: *const u32, width: u32, height: u32,
                                            cfg: *const ce_metric_config, intensity_target: f32,
                                            out: *mut ce_result) -> c_int;
    pub fn ce_evaluate_batch_device(ctx: *mut ce_ctx, d_ref: *const u8, d_dist: *const u8, n: usize, width: u32,
                                    height: u32, cfg: *const ce_metric_config, intensity_target: f32,
                                    out: *mut ce_result) -> c_int;
    pub fn ce_psnr(ctx: *mut ce_ctx, r: *const u8, r_len: usize, t: *const u8, t_len: usize, w: usize, h: usize,
                   psnr: *mut f64, sse: *mut u64) -> c_int;
    pub fn ce_ssimulacra2(ctx: *mut ce_ctx, r: *const u8, r_len: usize, t: *const u8, t_len: usize, w: usize,
                          h: usize, score: *mut f64) -> c_int;
    pub fn ce_butteraugli(ctx: *mut ce_ctx, r: *const u8, r_len: usize, t: *const u8, t_len: usize, w: usize,
                          h: usize, intensity_target: f32, score: *mut f64, pnorm3: *mut f64) -> c_int;
    pub fn ce_dssim_rgb8(ctx: *mut ce_ctx, r: *const u8, r_len: usize, t: *const u8, t_len: usize, w: usize,
                         h: usize, dssim: *mut f64) -> c_int;
    pub fn ce_dssim_rgbaf32(ctx: *mut ce_ctx, r: *const f32, rw: usize, rh: usize, rstride: usize,
                            t: *const f32, tw: usize, th: usize, tstride: usize, dssim: *mut f64) -> c_int;
    pub fn ce_rgb8_to_dssim_image(ctx: *mut ce_ctx, d: *const u8, len: usize, w: usize, h: usize, out: *mut f32) -> c_int;
    pub fn ce_rgba8_to_dssim_image(ctx: *mut ce_ctx, d: *const u8, len: usize, w: usize, h: usize, out: *mut f32) -> c_int;
    pub fn ce_xyb_roundtrip(ctx: *mut ce_ctx, rgb: *const u8, len: usize, w: usize, h: usize, out: *mut u8) -> c_int;
    pub fn ce_reference_create(ctx: *mut ce_ctx, r: *const u8, r_len: usize, w: usize, h: usize,
                               cfg: *const ce_metric_config, out: *mut *mut ce_ref) -> c_int;
    pub fn ce_reference_compare(ctx: *mut ce_ctx, r: *mut ce_ref, dist: *const u8, dist_len: usize,
                                intensity_target: f32, out: *mut ce_result) -> c_int;
    pub fn ce_reference_compare_many(ctx: *mut ce_ctx, r: *mut ce_ref, dists: *const *const u8,
                                     dist_lens: *const usize, n: usize, intensity_target: f32,
                                     out: *mut ce_result) -> c_int;
    pub fn ce_reference_destroy(r: *mut ce_ref);

    // ---- on-device distortion source (baseline JPEG round trip, bit-exact with libjpeg-turbo) ----
    pub fn ce_jpeg_roundtrip(ctx: *mut ce_ctx, rgb: *const u8, len: usize, width: usize, height: usize, quality: c_int,
                             subsampling: c_int, out: *mut u8) -> c_int;
    pub fn ce_jpeg_roundtrip_device(ctx: *mut ce_ctx, d_refs: *const u8, n_ref: usize, width: u32, height: u32,
                                    qualities: *const c_int, n_q: usize, subsampling: c_int, d_out: *mut u8) -> c_int;
    pub fn ce_evaluate_jpeg_sweep(ctx: *mut ce_ctx, refs: *const *const u8, n_ref: usize, width: u32, height: u32,
                                  qualities: *const c_int, n_q: usize, subsampling: c_int, cfg: *const ce_metric_config,
                                  intensity_target: f32, out: *mut ce_result) -> c_int;
    // ---- src/metrics/icc.rs:69-103 transform_to_srgb (matrix/TRC profiles); icc == null => ColorProfile::Srgb ----
    pub fn ce_transform_to_srgb(ctx: *mut ce_ctx, rgb: *const u8, len: usize, width: usize, height: usize, icc: *const u8,
                                icc_len: usize, out: *mut u8) -> c_int;
}
