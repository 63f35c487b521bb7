/*
 * ce_gpu.h -- C ABI of libce_gpu.so: codec-eval's metric hot path
 * (SSIMULACRA2, DSSIM, Butteraugli, PSNR, XYB round-trip) as sm_100a CUDA
 * kernels.  Plain pointers and sizes only; no exceptions cross the boundary.
 *
 * Each entry point names the reference interface it replaces (paths relative
 * to the imazen/codec-eval tree).  A Rust `-sys` crate binds these 1:1, see
 * INTEGRATION.md and rust/ce-gpu-sys.
 *
 * There is no CPU fallback: every compute entry returns CE_ERR_CUDA if no
 * sm_100 device is usable.
 */
#ifndef CE_GPU_H
#define CE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CE_API __attribute__((visibility("default")))
#else
#define CE_API
#endif

/* status codes.  1 and 2 map onto codec_eval::Error variants (src/error.rs:33,42). */
enum {
    CE_OK = 0,
    CE_ERR_DIMENSION_MISMATCH = 1, /* Error::DimensionMismatch  */
    CE_ERR_METRIC_CALCULATION = 2, /* Error::MetricCalculation  */
    CE_ERR_INVALID_ARGUMENT = 3,   /* null pointer / bad context (a Rust panic / assert in the reference) */
    CE_ERR_CUDA = 4,               /* CUDA runtime failure; text in ce_last_error() */
    CE_ERR_OUT_OF_MEMORY = 5       /* workspace too small even for one pair */
};

/* one CUDA device + stream + workspace; single owner (!Sync), like
 * GpuSsim2 (crates/codec-iter/src/gpu.rs:21-38). */
typedef struct ce_ctx ce_ctx;

/* = MetricConfig, src/metrics/mod.rs:45-63 (five bools, same order) */
typedef struct {
    uint8_t dssim, ssimulacra2, butteraugli, psnr, xyb_roundtrip;
} ce_metric_config;

/* one reference/distorted pair in HOST memory: RGB8, row-major, tight stride
 * (src/metrics/ssimulacra2.rs:41-44).  ref_len/dist_len are the slice lengths
 * the Rust caller holds (they drive the same validation, in the same order,
 * as ssimulacra2.rs:65-82).  ref_id groups pairs that share a reference
 * (evaluate_image's codecs x quality_levels, src/eval/session.rs:375-431):
 * pairs of one call whose `ref` pointer AND `ref_id` are both equal are
 * treated as one reference image, uploaded and pre-processed once.  Give
 * every pair its own ref_id (or its own buffer) to switch that off. */
typedef struct {
    const uint8_t* ref;
    const uint8_t* dist;
    size_t ref_len, dist_len;
    uint32_t width, height;
    uint32_t ref_id;
    uint32_t reserved;
} ce_pair;

/* = MetricResult (4 x Option<f64>, src/metrics/mod.rs:139-149) + raw SSE and
 * the libjxl 3-norm.  valid bit0 dssim, bit1 ssimulacra2, bit2 butteraugli,
 * bit3 psnr  <->  Some(..). */
typedef struct {
    int32_t status;
    uint32_t valid;
    uint64_t sse; /* exact sum of squared byte differences (PSNR bit-exactness) */
    double dssim, ssimulacra2, butteraugli, psnr, butteraugli_pnorm3;
} ce_result;

#define CE_VALID_DSSIM 1u
#define CE_VALID_SSIMULACRA2 2u
#define CE_VALID_BUTTERAUGLI 4u
#define CE_VALID_PSNR 8u

/* ---- context ------------------------------------------------------- */

/* replaces GpuSsim2::new (crates/codec-iter/src/gpu.rs:40-77).  workspace_bytes
 * = device scratch for intermediates (0 = default: half of free memory, at most
 * 64 GiB); batches larger than the workspace are processed in sub-batches. */
CE_API int ce_ctx_create(ce_ctx** out, int device, size_t workspace_bytes);
CE_API void ce_ctx_destroy(ce_ctx* ctx);
/* run all work on the caller's CUDA stream (cudaStream_t); NULL = the context's own stream */
CE_API int ce_ctx_set_stream(ce_ctx* ctx, void* cuda_stream);
/* text for Error::MetricCalculation.reason / CUDA failures; valid until the next call on ctx.
 * ctx == NULL returns the message of the last failed ce_ctx_create on this thread. */
CE_API const char* ce_last_error(const ce_ctx* ctx);
/* number of this library's kernels launched on ctx since creation */
CE_API uint64_t ce_launch_count(const ce_ctx* ctx);
/* "ce_gpu <version> (sm_100a) src:<hash>": the hash covers csrc/ and this header as they were when the
 * library was built (codec_eval_b200/build.py); the Python loader refuses a library whose hash differs
 * from the sources beside it. */
CE_API const char* ce_version(void);

/* ---- pinned host memory (opt-in) -------------------------------------- */
/* The Rust caller holds decoded images in pageable Vec<u8> (src/eval/session.rs:394).  ce_evaluate_batch
 * accepts those as they are -- the copy of chunk k+1 then blocks the calling thread under the kernels of
 * chunk k -- but page-locked buffers copy at full PCIe rate and fully asynchronously.  Two ways to opt in:
 * register an existing allocation (cudaHostRegister; the decode buffers of a session, once), or let the
 * library allocate a pinned ring the decoder writes into. */
CE_API int ce_host_register(ce_ctx* ctx, void* ptr, size_t bytes);
CE_API int ce_host_unregister(ce_ctx* ctx, void* ptr);
CE_API int ce_host_alloc(ce_ctx* ctx, size_t bytes, void** out);
CE_API void ce_host_free(ce_ctx* ctx, void* ptr);
/* opt-in per-kernel CUDA-event timing on the context's stream (bench.py's roofline evidence).
 * report: one line per kernel "name\tlaunches\tmilliseconds\talgorithmic_bytes\tbytes_per_pair\n"; returns
 * the length needed.  algorithmic_bytes counts each distinct input element and each output element once
 * (planes of a reference shared by several pairs once per reference); bytes_per_pair charges those shared
 * planes once per pair (the figure for batches without reference reuse).  Durations are collected when a
 * batch call synchronises. */
CE_API int ce_profile_enable(ce_ctx* ctx, int enable);
CE_API int ce_profile_reset(ce_ctx* ctx);
CE_API size_t ce_profile_report(ce_ctx* ctx, char* buf, size_t cap);

/* ---- batched entry points ------------------------------------------ */

/* The batched GPU entry point that EvalSession::calculate_metrics
 * (src/eval/session.rs:437-497) and evaluate_single (src/eval/helpers.rs:105-173)
 * dispatch into: n host pairs -> n results, same metric order and the same
 * "reference only" XYB round-trip rule (session.rs:447-456).  Host->device copies
 * happen inside.  A failing pair sets out[i].status and does not poison the batch;
 * the return value is CE_OK unless the call itself could not run. */
CE_API int ce_evaluate_batch(ce_ctx* ctx, const ce_pair* pairs, size_t n, const ce_metric_config* cfg,
                             float intensity_target, ce_result* out);

/* How many pairs of this size and metric set one sub-batch holds in the context's workspace (larger batches
 * are processed as several sub-batches; callers that stream a corpus can size their calls to a multiple of it).
 * CE_ERR_OUT_OF_MEMORY when not even one pair fits. */
CE_API int ce_sub_batch_capacity(ce_ctx* ctx, const ce_metric_config* cfg, uint32_t width, uint32_t height, size_t* pairs);

/* Same computation for a uniform-size batch already resident in device memory:
 * d_ref / d_dist hold n tightly packed RGB8 images of width x height (the
 * benchmark's timed entry; also what an on-device decoder would call). */
CE_API int ce_evaluate_batch_device(ce_ctx* ctx, const uint8_t* d_ref, const uint8_t* d_dist, size_t n, uint32_t width,
                                    uint32_t height, const ce_metric_config* cfg, float intensity_target,
                                    ce_result* out);

/* Device-resident batch with shared references: n_ref reference images, n distorted images, pair i compares
 * reference ref_index[i] (host array) with distorted image i.  This is evaluate_image's shape -- one reference
 * against every codec x quality output (src/eval/session.rs:375-431) -- and what Ssimulacra2Reference::new /
 * .compare exploits on the CPU (crates/codec-iter/src/eval.rs:138-149): reference-side work is done once per
 * distinct reference.  ce_evaluate_batch does the same automatically for pairs whose `ref` pointers are equal. */
CE_API int ce_evaluate_batch_device_grouped(ce_ctx* ctx, const uint8_t* d_ref, size_t n_ref, const uint8_t* d_dist, size_t n,
                                            const uint32_t* ref_index, uint32_t width, uint32_t height,
                                            const ce_metric_config* cfg, float intensity_target, ce_result* out);

/* ---- single-pair mirrors of src/metrics ----------------------------- */

/* calculate_psnr, src/metrics/mod.rs:312-331 (length asserts become CE_ERR_INVALID_ARGUMENT) */
CE_API int ce_psnr(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t width,
                   size_t height, double* psnr, uint64_t* sse);
/* calculate_ssimulacra2, src/metrics/ssimulacra2.rs:59-100 */
CE_API int ce_ssimulacra2(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len,
                          size_t width, size_t height, double* score);
/* calculate_butteraugli / calculate_butteraugli_with_intensity, src/metrics/butteraugli.rs:45-81,99-136
 * (default intensity_target 80.0); pnorm3 may be NULL */
CE_API int ce_butteraugli(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len,
                          size_t width, size_t height, float intensity_target, double* score, double* pnorm3);
/* rgb8_to_dssim_image x2 + calculate_dssim fused, src/metrics/dssim.rs:102-114,40-71 / session.rs:467-476 */
CE_API int ce_dssim_rgb8(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len,
                         size_t width, size_t height, double* dssim);
/* calculate_dssim on linear RGBA f32 (ImgVec<RGBA<f32>>), src/metrics/dssim.rs:40-71; strides in pixels */
CE_API int ce_dssim_rgbaf32(ce_ctx* ctx, const float* ref, size_t ref_w, size_t ref_h, size_t ref_stride,
                            const float* test, size_t test_w, size_t test_h, size_t test_stride, double* dssim);
/* rgb8_to_dssim_image / rgba8_to_dssim_image, src/metrics/dssim.rs:102-114,131-143: out = width*height*4 floats */
CE_API int ce_rgb8_to_dssim_image(ce_ctx* ctx, const uint8_t* data, size_t len, size_t width, size_t height, float* out);
CE_API int ce_rgba8_to_dssim_image(ce_ctx* ctx, const uint8_t* data, size_t len, size_t width, size_t height, float* out);
/* xyb_roundtrip, src/metrics/xyb.rs:225-253: out = width*height*3 bytes */
CE_API int ce_xyb_roundtrip(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height, uint8_t* out);

/* ---- reference reuse ------------------------------------------------ */
/* mirrors fast_ssim2::Ssimulacra2Reference::new / .compare as used by
 * crates/codec-iter/src/eval.rs:138-149,84-88: the reference image stays on the
 * device; each compare uploads only the distorted image. */
typedef struct ce_ref ce_ref;
CE_API int ce_reference_create(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, size_t width, size_t height,
                               const ce_metric_config* cfg, ce_ref** out);
CE_API int ce_reference_compare(ce_ctx* ctx, ce_ref* ref, const uint8_t* dist, size_t dist_len, float intensity_target,
                                ce_result* out);
/* all distortions of one reference in one launch set (dists: n_dist host pointers) */
CE_API int ce_reference_compare_many(ce_ctx* ctx, ce_ref* ref, const uint8_t* const* dists, const size_t* dist_lens,
                                     size_t n_dist, float intensity_target, ce_result* out);
CE_API void ce_reference_destroy(ce_ref* ref);

/* ---- on-device distortion source -------------- */
/* The step BEFORE the metric path in codec-iter's eval loop is encode -> decode per quality level
 * (crates/codec-iter/src/eval.rs:153-172, run_eval: `decode -> compare`).  These entry points produce the decoded
 * image of a baseline JPEG (IJG / libjpeg-turbo defaults: islow DCT, Annex-K tables scaled by `quality`,
 * subsampling 0 = 4:4:4 or 2 = 4:2:0 with fancy upsampling) directly on the device, bit-exact with
 * libjpeg-turbo's encode -> decode, so a quality sweep uploads only the reference images.  No bitstream is
 * produced (entropy coding is lossless and skipped): file sizes still come from the real encoder. */

/* one host image -> its JPEG(quality, subsampling) round trip; out = width*height*3 bytes */
CE_API int ce_jpeg_roundtrip(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height, int quality,
                             int subsampling, uint8_t* out);
/* device-resident: n_ref tightly packed RGB8 references -> d_out[(r * n_q + k)] = reference r at qualities[k]
 * (qualities: host array) */
CE_API int ce_jpeg_roundtrip_device(ce_ctx* ctx, const uint8_t* d_refs, size_t n_ref, uint32_t width, uint32_t height,
                                    const int* qualities, size_t n_q, int subsampling, uint8_t* d_out);
/* Quality sweep in one call: n_ref HOST references (refs[r], width*height*3 bytes each) x n_q qualities ->
 * out[r * n_q + k] = metrics of reference r against its own JPEG(qualities[k]) round trip.  Only the references
 * cross PCIe; reference-side metric work is shared by the n_q distortions of a reference
 * (Ssimulacra2Reference reuse, crates/codec-iter/src/eval.rs:138-149). */
CE_API int ce_evaluate_jpeg_sweep(ce_ctx* ctx, const uint8_t* const* refs, size_t n_ref, uint32_t width, uint32_t height,
                                  const int* qualities, size_t n_q, int subsampling, const ce_metric_config* cfg,
                                  float intensity_target, ce_result* out);

/* ---- ICC -> sRGB on the device ---------------- */
/* transform_to_srgb, src/metrics/icc.rs:69-103 (what ImageData::to_rgb8_srgb applies before the metrics,
 * src/eval/session.rs:143-147): RGB8 in the colour space of `icc` -> RGB8 sRGB; out = width*height*3 bytes.
 * icc == NULL or icc_len == 0 is ColorProfile::Srgb (bytes are copied, icc.rs:73).  Matrix/TRC RGB profiles
 * (rXYZ/gXYZ/bXYZ + rTRC/gTRC/bTRC, curveType or parametricCurveType) are supported; anything else returns
 * CE_ERR_METRIC_CALCULATION with the reason in ce_last_error(), i.e. Error::MetricCalculation{metric:"ICC"}. */
CE_API int ce_transform_to_srgb(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height,
                                const uint8_t* icc, size_t icc_len, uint8_t* out);

/* ---- stage-level entry points (parity tests; device does the work) -- */
/* Each runs ONE pipeline stage on the device for a single host image / pair
 * and returns the intermediate, so tests can localise a mismatch against the
 * oracle's same stage.  Not part of the reference's API. */
CE_API int ce_debug_ssim2_sums(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height,
                               double* sums /* 6*18 */, int* nscales);
CE_API int ce_debug_ssim2_scale0_planes(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width,
                                        size_t height, float* planes /* 3*7*h*w: i1,i2,mu1,mu2,s11,s22,s12 */);
CE_API int ce_debug_dssim_scales(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height,
                                 double* scale_scores /* 5 */, int* nscales, float* map0 /* h*w or NULL */);
CE_API int ce_debug_butteraugli_diffmap(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width,
                                        size_t height, float intensity_target, float* diffmap /* h*w */);
CE_API int ce_debug_butteraugli_psycho(ce_ctx* ctx, const uint8_t* rgb, size_t width, size_t height,
                                       float intensity_target, float* planes /* 10*h*w */);
CE_API int ce_debug_butteraugli_opsin(ce_ctx* ctx, const uint8_t* rgb, size_t width, size_t height,
                                      float intensity_target, float* planes /* 3*h*w */);
CE_API int ce_debug_ba_blur(ce_ctx* ctx, const float* plane, size_t width, size_t height, float sigma, float* out);

#ifdef __cplusplus
}
#endif
#endif /* CE_GPU_H */
