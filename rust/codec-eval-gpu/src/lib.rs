//! Safe wrapper over `ce-gpu-sys` with codec-eval's types (src/metrics/mod.rs:45-149, src/error.rs:33-47).
//! `GpuMetrics::evaluate_batch` is the batched entry `EvalSession::evaluate_image` (src/eval/session.rs:368-434)
//! dispatches into; see `evaluate_image_batched` at the bottom for the session-side loop.
use ce_gpu_sys as sys;
use std::ffi::CStr;

#[derive(Debug, Clone, Copy, Default, PartialEq)]
pub struct MetricConfig { pub dssim: bool, pub ssimulacra2: bool, pub butteraugli: bool, pub psnr: bool, pub xyb_roundtrip: bool }

#[derive(Debug, Clone, Default, PartialEq)]
pub struct MetricResult { pub dssim: Option<f64>, pub ssimulacra2: Option<f64>, pub butteraugli: Option<f64>, pub psnr: Option<f64> }

#[derive(Debug)]
pub enum Error {
    DimensionMismatch { expected: (usize, usize), actual: (usize, usize) },
    MetricCalculation { metric: String, reason: String },
}
pub type Result<T> = std::result::Result<T, Error>;

/// One CUDA device + stream + workspace.  Single owner (`!Sync`), like `GpuSsim2`
/// (crates/codec-iter/src/gpu.rs:21-38).
pub struct GpuMetrics { ctx: *mut sys::ce_ctx }
unsafe impl Send for GpuMetrics {}

impl GpuMetrics {
    pub fn new(device: i32) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { sys::ce_ctx_create(&mut ctx, device, 0) };
        if st != sys::CE_OK {
            let reason = unsafe { CStr::from_ptr(sys::ce_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(Error::MetricCalculation { metric: "GPU".into(), reason });
        }
        Ok(Self { ctx })
    }

    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::ce_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }

    /// Batched `calculate_metrics` (src/eval/session.rs:437-497).  Pairs that borrow the same reference slice are
    /// grouped inside the library: the reference is uploaded and pre-processed once.
    pub fn evaluate_batch(&mut self, pairs: &[(&[u8], &[u8], u32, u32)], cfg: &MetricConfig) -> Vec<Result<MetricResult>> {
        // pairs that borrow the same reference slice get the same ref_id: the library treats equal (pointer, ref_id)
        // as ONE reference image
        let mut ids: std::collections::HashMap<*const u8, u32> = std::collections::HashMap::new();
        let c_pairs: Vec<sys::ce_pair> = pairs.iter().map(|(r, t, w, h)| {
            let next = ids.len() as u32;
            let ref_id = *ids.entry(r.as_ptr()).or_insert(next);
            sys::ce_pair { reference: r.as_ptr(), dist: t.as_ptr(), ref_len: r.len(), dist_len: t.len(),
                           width: *w, height: *h, ref_id, reserved: 0 }
        }).collect();
        let c_cfg = sys::ce_metric_config {
            dssim: cfg.dssim as u8, ssimulacra2: cfg.ssimulacra2 as u8, butteraugli: cfg.butteraugli as u8,
            psnr: cfg.psnr as u8, xyb_roundtrip: cfg.xyb_roundtrip as u8,
        };
        let mut out = vec![sys::ce_result::default(); pairs.len()];
        let st = unsafe { sys::ce_evaluate_batch(self.ctx, c_pairs.as_ptr(), c_pairs.len(), &c_cfg, 80.0, out.as_mut_ptr()) };
        if st != sys::CE_OK {
            let reason = self.last_error();
            return pairs.iter().map(|_| Err(Error::MetricCalculation { metric: "GPU".into(), reason: reason.clone() })).collect();
        }
        out.iter().zip(pairs).map(|(r, p)| match r.status {
            sys::CE_OK => Ok(MetricResult {
                dssim: (r.valid & sys::CE_VALID_DSSIM != 0).then_some(r.dssim),
                ssimulacra2: (r.valid & sys::CE_VALID_SSIMULACRA2 != 0).then_some(r.ssimulacra2),
                butteraugli: (r.valid & sys::CE_VALID_BUTTERAUGLI != 0).then_some(r.butteraugli),
                psnr: (r.valid & sys::CE_VALID_PSNR != 0).then_some(r.psnr),
            }),
            sys::CE_ERR_DIMENSION_MISMATCH => Err(Error::DimensionMismatch {        // src/metrics/ssimulacra2.rs:65-70
                expected: (p.2 as usize, p.3 as usize),
                actual: (p.1.len() / 3 / (p.3 as usize).max(1), p.3 as usize),
            }),
            _ => Err(Error::MetricCalculation { metric: "GPU".into(), reason: self.last_error() }),
        }).collect()
    }

    /// `calculate_ssimulacra2` (src/metrics/ssimulacra2.rs:59-100) for one pair.
    pub fn calculate_ssimulacra2(&mut self, reference: &[u8], test: &[u8], width: usize, height: usize) -> Result<f64> {
        let mut score = 0.0f64;
        let st = unsafe { sys::ce_ssimulacra2(self.ctx, reference.as_ptr(), reference.len(), test.as_ptr(), test.len(), width, height, &mut score) };
        match st {
            sys::CE_OK => Ok(score),
            sys::CE_ERR_DIMENSION_MISMATCH => Err(Error::DimensionMismatch { expected: (width, height), actual: (test.len() / 3 / height.max(1), height) }),
            _ => Err(Error::MetricCalculation { metric: "SSIMULACRA2".into(), reason: self.last_error() }),
        }
    }
}

impl GpuMetrics {
    /// `calculate_psnr` (src/metrics/mod.rs:312-331): asserts on the lengths like the reference, exact integer SSE on
    /// the device, `f64::INFINITY` for identical images.
    pub fn calculate_psnr(&mut self, reference: &[u8], test: &[u8], width: usize, height: usize) -> f64 {
        assert_eq!(reference.len(), test.len());
        assert_eq!(reference.len(), width * height * 3);
        let (mut psnr, mut sse) = (0.0f64, 0u64);
        let st = unsafe { sys::ce_psnr(self.ctx, reference.as_ptr(), reference.len(), test.as_ptr(), test.len(), width, height, &mut psnr, &mut sse) };
        assert_eq!(st, sys::CE_OK, "ce_psnr: {}", self.last_error());
        psnr
    }

    /// `rgb8_to_dssim_image` x2 + `calculate_dssim` (src/metrics/dssim.rs:102-114,40-71) as EvalSession calls them
    /// (src/eval/session.rs:467-476), fused on the device.
    pub fn calculate_dssim_rgb8(&mut self, reference: &[u8], test: &[u8], width: usize, height: usize) -> Result<f64> {
        let mut d = 0.0f64;
        let st = unsafe { sys::ce_dssim_rgb8(self.ctx, reference.as_ptr(), reference.len(), test.as_ptr(), test.len(), width, height, &mut d) };
        self.status_to_result(st, "DSSIM", d, width, height, test.len())
    }

    /// `calculate_dssim` (src/metrics/dssim.rs:40-71) on linear RGBA f32 images (`ImgVec<RGBA<f32>>` as flat
    /// `[r, g, b, a]` floats; strides in pixels).
    pub fn calculate_dssim(&mut self, reference: &[f32], ref_size: (usize, usize), ref_stride: usize, test: &[f32],
                           test_size: (usize, usize), test_stride: usize) -> Result<f64> {
        if ref_size != test_size {                                       // src/metrics/dssim.rs:45-50
            return Err(Error::DimensionMismatch { expected: ref_size, actual: test_size });
        }
        let need = |stride: usize, (w, h): (usize, usize)| if h == 0 { 0 } else { (stride * (h - 1) + w) * 4 };
        if ref_stride < ref_size.0 || test_stride < test_size.0 || reference.len() < need(ref_stride, ref_size)
            || test.len() < need(test_stride, test_size) {
            return Err(Error::MetricCalculation { metric: "DSSIM".into(), reason: "Failed to create reference image".into() });
        }
        let mut d = 0.0f64;
        let st = unsafe {
            sys::ce_dssim_rgbaf32(self.ctx, reference.as_ptr(), ref_size.0, ref_size.1, ref_stride, test.as_ptr(), test_size.0,
                                  test_size.1, test_stride, &mut d)
        };
        self.status_to_result(st, "DSSIM", d, ref_size.0, ref_size.1, test_size.0 * test_size.1 * 3)
    }

    /// `calculate_butteraugli` (src/metrics/butteraugli.rs:45-81): intensity target 80 nits.
    pub fn calculate_butteraugli(&mut self, reference: &[u8], test: &[u8], width: usize, height: usize) -> Result<f64> {
        self.calculate_butteraugli_with_intensity(reference, test, width, height, 80.0)
    }

    /// `calculate_butteraugli_with_intensity` (src/metrics/butteraugli.rs:99-136).
    pub fn calculate_butteraugli_with_intensity(&mut self, reference: &[u8], test: &[u8], width: usize, height: usize,
                                                intensity_target: f32) -> Result<f64> {
        let mut score = 0.0f64;
        let st = unsafe {
            sys::ce_butteraugli(self.ctx, reference.as_ptr(), reference.len(), test.as_ptr(), test.len(), width, height,
                                intensity_target, &mut score, std::ptr::null_mut())
        };
        self.status_to_result(st, "Butteraugli", score, width, height, test.len())
    }

    /// `xyb_roundtrip` (src/metrics/xyb.rs:225-253).
    pub fn xyb_roundtrip(&mut self, rgb: &[u8], width: usize, height: usize) -> Vec<u8> {
        assert_eq!(rgb.len(), width * height * 3, "Buffer size mismatch");
        let mut out = vec![0u8; rgb.len()];
        let st = unsafe { sys::ce_xyb_roundtrip(self.ctx, rgb.as_ptr(), rgb.len(), width, height, out.as_mut_ptr()) };
        assert_eq!(st, sys::CE_OK, "ce_xyb_roundtrip: {}", self.last_error());
        out
    }

    /// Page-lock a decode buffer the session keeps (cudaHostRegister): its copies then run asynchronously at PCIe rate.
    /// The slice must stay allocated (and not be reallocated) until `host_unregister`.
    pub fn host_register(&mut self, buf: &mut [u8]) -> Result<()> {
        let st = unsafe { sys::ce_host_register(self.ctx, buf.as_mut_ptr() as *mut std::ffi::c_void, buf.len()) };
        self.status_to_result(st, "GPU", (), 0, 0, 0)
    }
    pub fn host_unregister(&mut self, buf: &mut [u8]) -> Result<()> {
        let st = unsafe { sys::ce_host_unregister(self.ctx, buf.as_mut_ptr() as *mut std::ffi::c_void) };
        self.status_to_result(st, "GPU", (), 0, 0, 0)
    }

    fn status_to_result<T>(&self, st: i32, metric: &str, value: T, width: usize, height: usize, test_len: usize) -> Result<T> {
        match st {
            sys::CE_OK => Ok(value),
            sys::CE_ERR_DIMENSION_MISMATCH => Err(Error::DimensionMismatch { expected: (width, height), actual: (test_len / 3 / height.max(1), height) }),
            _ => Err(Error::MetricCalculation { metric: metric.into(), reason: self.last_error() }),
        }
    }
}

impl Drop for GpuMetrics {
    fn drop(&mut self) { unsafe { sys::ce_ctx_destroy(self.ctx) } }
}

/// The dispatch change inside `EvalSession::evaluate_image` (src/eval/session.rs:375-431): the caller has already
/// encoded/decoded every (codec, quality) output of one reference; this fills the metric slots in order with ONE
/// GPU call instead of one `calculate_metrics` call per output.
pub fn evaluate_image_batched(gpu: &mut GpuMetrics, reference_rgb: &[u8], decoded_rgb: &[Vec<u8>], width: u32, height: u32,
                              cfg: &MetricConfig) -> Result<Vec<MetricResult>> {
    let pairs: Vec<(&[u8], &[u8], u32, u32)> = decoded_rgb.iter().map(|d| (reference_rgb, d.as_slice(), width, height)).collect();
    gpu.evaluate_batch(&pairs, cfg).into_iter().collect()      // first per-pair error, like the `?` chain today
}

impl GpuMetrics {
    /// `transform_to_srgb` (src/metrics/icc.rs:69-103) for matrix/TRC profiles on the device; `None` = ColorProfile::Srgb.
    /// `Err(MetricCalculation{metric:"ICC"})` for profiles the device path does not cover (the caller falls back to moxcms).
    pub fn transform_to_srgb(&mut self, rgb: &[u8], width: usize, height: usize, icc: Option<&[u8]>) -> Result<Vec<u8>> {
        let mut out = vec![0u8; rgb.len()];
        let (p, n) = icc.map_or((std::ptr::null(), 0), |d| (d.as_ptr(), d.len()));
        let st = unsafe { sys::ce_transform_to_srgb(self.ctx, rgb.as_ptr(), rgb.len(), width, height, p, n, out.as_mut_ptr()) };
        match st {
            sys::CE_OK => Ok(out),
            _ => Err(Error::MetricCalculation { metric: "ICC".into(), reason: self.last_error() }),
        }
    }

    /// codec-iter's per-image quality sweep (crates/codec-iter/src/eval.rs:153-200) for baseline JPEG with the decoded
    /// images generated on the device (bit-exact with libjpeg-turbo): `out[r][k]` = metrics of `refs[r]` against its own
    /// JPEG(`qualities[k]`) round trip.  Only the references are uploaded.  `subsampling`: 0 = 4:4:4, 2 = 4:2:0.
    pub fn evaluate_jpeg_sweep(&mut self, refs: &[&[u8]], width: u32, height: u32, qualities: &[i32], subsampling: i32,
                               cfg: &MetricConfig) -> Result<Vec<Vec<MetricResult>>> {
        // the C side reads width*height*3 bytes from every pointer: a short slice must never reach it
        let need = (width as usize).checked_mul(height as usize).and_then(|n| n.checked_mul(3));
        for r in refs {
            if need != Some(r.len()) {
                return Err(Error::DimensionMismatch { expected: (width as usize, height as usize),
                                                      actual: (r.len() / 3 / (height as usize).max(1), height as usize) });
            }
        }
        if qualities.len() > 32 {
            return Err(Error::MetricCalculation { metric: "JPEG".into(), reason: "at most 32 quality levels per call".into() });
        }
        let ptrs: Vec<*const u8> = refs.iter().map(|r| r.as_ptr()).collect();
        let c_cfg = sys::ce_metric_config {
            dssim: cfg.dssim as u8, ssimulacra2: cfg.ssimulacra2 as u8, butteraugli: cfg.butteraugli as u8,
            psnr: cfg.psnr as u8, xyb_roundtrip: cfg.xyb_roundtrip as u8,
        };
        let mut out = vec![sys::ce_result::default(); refs.len() * qualities.len()];
        let st = unsafe {
            sys::ce_evaluate_jpeg_sweep(self.ctx, ptrs.as_ptr(), refs.len(), width, height, qualities.as_ptr(), qualities.len(),
                                        subsampling, &c_cfg, 80.0, out.as_mut_ptr())
        };
        if st != sys::CE_OK {
            return Err(Error::MetricCalculation { metric: "GPU".into(), reason: self.last_error() });
        }
        Ok(out.chunks(qualities.len().max(1)).map(|row| row.iter().map(|r| MetricResult {
            dssim: (r.valid & sys::CE_VALID_DSSIM != 0).then_some(r.dssim),
            ssimulacra2: (r.valid & sys::CE_VALID_SSIMULACRA2 != 0).then_some(r.ssimulacra2),
            butteraugli: (r.valid & sys::CE_VALID_BUTTERAUGLI != 0).then_some(r.butteraugli),
            psnr: (r.valid & sys::CE_VALID_PSNR != 0).then_some(r.psnr),
        }).collect()).collect())
    }
}
