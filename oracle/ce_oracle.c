/*
 * ce_oracle.c -- CPU oracle for the codec-eval metric hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (codec_eval_b200/,
 * include/, the CUDA library) may link, import or execute this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED for SSIMULACRA2, DSSIM and Butteraugli: the arithmetic of
 * those three metrics lives in crates that are not vendored under
 * /root/reference (fast-ssim2 0.8.0, dssim-core 3.4.0, butteraugli 0.9.0;
 * Cargo.lock:410,356,132) and the reference's own tests hold inequalities
 * only (src/metrics/ssimulacra2.rs:154-174, dssim.rs:181-223,
 * butteraugli.rs:169-207).  Those sections restate the published upstream
 * algorithms (libjxl tools/ssimulacra2.cc, kornelski/dssim dssim-core,
 * libjxl lib/jxl/butteraugli/butteraugli.cc) as summarised in SURVEY.md
 * Appendix A.  PSNR, sRGB->linear and the XYB u8 round-trip are restated
 * from in-tree reference code and are pinned by its tests.
 *
 * All fp32 arithmetic is written as explicit IEEE operations (compile with
 * -ffp-contract=off; fmaf() only where the upstream code has mul_add), so
 * that the CUDA path can execute the same operation sequence and be compared
 * almost bit-for-bit.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CEO_API __attribute__((visibility("default")))

/* status codes = ce_gpu.h / src/error.rs:33,42 */
#define CEO_OK 0
#define CEO_DIMENSION_MISMATCH 1
#define CEO_METRIC_CALCULATION 2

static inline float asf(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t asu(float f) { uint32_t i; memcpy(&i, &f, 4); return i; }

/* ------------------------------------------------------------------ */
/* sRGB transfer functions                                             */
/* ------------------------------------------------------------------ */

/* src/metrics/dssim.rs:77-85 and src/metrics/xyb.rs:60-66,80-82 (same
 * expression): s = v/255; s <= 0.04045 ? s/12.92 : ((s+0.055)/1.055)^2.4 */
static float srgb_u8_to_linear(uint8_t v) {
    float s = (float)v / 255.0f;
    if (s <= 0.04045f) return s / 12.92f;
    return powf((s + 0.055f) / 1.055f, 2.4f);
}

static float g_lut[256];
static int g_lut_ready = 0;
static void rgauss_init(void);

__attribute__((constructor)) static void ceo_init(void) {
    for (int i = 0; i < 256; i++) g_lut[i] = srgb_u8_to_linear((uint8_t)i);
    g_lut_ready = 1;
    rgauss_init();
}

CEO_API void ceo_srgb_lut(float* out256) { memcpy(out256, g_lut, sizeof(g_lut)); }

/* ------------------------------------------------------------------ */
/* A.1 PSNR -- src/metrics/mod.rs:312-331                              */
/* ------------------------------------------------------------------ */

CEO_API uint64_t ceo_sse(const uint8_t* a, const uint8_t* b, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) {
        int d = (int)a[i] - (int)b[i];
        s += (uint64_t)(d * d);
    }
    return s;
}

/* Literal restatement of the reference loop (f64 accumulation of squared
 * f64 differences, mod.rs:316-322).  Exact because every partial sum is an
 * integer < 2^53. */
CEO_API double ceo_psnr(const uint8_t* ref, const uint8_t* test, size_t w, size_t h) {
    double mse_sum = 0.0;
    size_t n = w * h * 3;
    double pixel_count = (double)n;
    for (size_t i = 0; i < n; i++) {
        double diff = (double)ref[i] - (double)test[i];
        mse_sum += diff * diff;
    }
    double mse = mse_sum / pixel_count;
    if (mse == 0.0) return INFINITY;
    return 10.0 * log10(255.0 * 255.0 / mse);
}

/* the same final expression from an exact integer SSE (what the GPU path does) */
CEO_API double ceo_psnr_from_sse(uint64_t sse, size_t w, size_t h) {
    double mse = (double)sse / (double)(w * h * 3);
    if (mse == 0.0) return INFINITY;
    return 10.0 * log10(255.0 * 255.0 / mse);
}

/* ------------------------------------------------------------------ */
/* A.2 XYB u8 round-trip -- src/metrics/xyb.rs                         */
/* ------------------------------------------------------------------ */

/* Rust's f32::cbrt / f32::powf defer to the platform libm, whose results
 * differ between platforms in the last ulp (glibc 2.39 cbrtf is off by one
 * ulp from the correctly rounded value for ~11% of inputs).  The oracle
 * pins the platform-independent definition: the correctly rounded result,
 * obtained by evaluating in double and rounding once. */
/* g_xyb_libm (tests only, ceo_xyb_roundtrip_libm): call this platform's cbrtf / powf instead -- what the Rust
 * reference itself executes on a glibc host -- so a test can count how many output bytes the choice moves. */
static __thread int g_xyb_libm = 0;
static inline float cr_cbrtf(float v) { return g_xyb_libm ? cbrtf(v) : (float)cbrt((double)v); }
static inline float cr_powf(float v, float e) { return g_xyb_libm ? powf(v, e) : (float)pow((double)v, (double)e); }

static const float XYB_M[9] = {0.30f, 0.622f, 0.078f, 0.23f, 0.692f, 0.078f,
                               0.24342269f, 0.20476744f, 0.55180987f};  /* xyb.rs:33-43 */
static const float XYB_BIAS = 0.0037930733f;                            /* xyb.rs:45 */
static const float XYB_NEG_BIAS_CBRT = -0.15595412f;                    /* xyb.rs:47-51 */
static const float XYB_INV[9] = {11.031567f, -9.866944f, -0.164623f, -3.254147f, 4.41877f,
                                 -0.164623f, -3.658851f, 2.712923f, 1.945928f}; /* xyb.rs:53-56 */

static inline float xyb_srgb_to_linear_f32(float v) { /* xyb.rs:60-66 */
    if (v <= 0.04045f) return v / 12.92f;
    return cr_powf((v + 0.055f) / 1.055f, 2.4f);
}
static inline float xyb_linear_to_srgb_f32(float v) { /* xyb.rs:70-76 */
    if (v <= 0.0031308f) return v * 12.92f;
    return 1.055f * cr_powf(v, 1.0f / 2.4f) - 0.055f;
}
static inline float mixed_cbrt(float v) { return v < 0.0f ? -cr_cbrtf(-v) : cr_cbrtf(v); } /* :92-94 */
static inline float mixed_cube(float v) { /* xyb.rs:98-100, powi(3) = (v*v)*v */
    if (v < 0.0f) { float a = -v; return -((a * a) * a); }
    return (v * v) * v;
}
static inline float quantize_to_u8(float value, float mn, float mx) { /* xyb.rs:192-199 */
    float range = mx - mn;
    float normalized = (value - mn) / range;
    float r = roundf(normalized * 255.0f); /* half away from zero = Rust f32::round */
    if (r < 0.0f) r = 0.0f;
    if (r > 255.0f) r = 255.0f;
    float quantized = r / 255.0f;
    return quantized * range + mn;
}
static inline uint8_t linear_to_srgb_u8(float v) { /* xyb.rs:86-88 */
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t)roundf(xyb_linear_to_srgb_f32(v) * 255.0f);
}

CEO_API void ceo_xyb_roundtrip(const uint8_t* rgb, size_t w, size_t h, uint8_t* out) {
    size_t n = w * h;
    for (size_t i = 0; i < n; i++) { /* xyb.rs:225-253 */
        float r = xyb_srgb_to_linear_f32((float)rgb[i * 3] / 255.0f);
        float g = xyb_srgb_to_linear_f32((float)rgb[i * 3 + 1] / 255.0f);
        float b = xyb_srgb_to_linear_f32((float)rgb[i * 3 + 2] / 255.0f);
        /* linear_rgb_to_xyb, xyb.rs:104-131 (plain mul/add, left to right) */
        float o_r = XYB_M[0] * r + XYB_M[1] * g + XYB_M[2] * b + XYB_BIAS;
        float o_g = XYB_M[3] * r + XYB_M[4] * g + XYB_M[5] * b + XYB_BIAS;
        float o_b = XYB_M[6] * r + XYB_M[7] * g + XYB_M[8] * b + XYB_BIAS;
        float c_r = mixed_cbrt(o_r) + XYB_NEG_BIAS_CBRT;
        float c_g = mixed_cbrt(o_g) + XYB_NEG_BIAS_CBRT;
        float c_b = mixed_cbrt(o_b) + XYB_NEG_BIAS_CBRT;
        float x = 0.5f * (c_r - c_g);
        float y = 0.5f * (c_r + c_g);
        float bb = c_b;
        float xq = quantize_to_u8(x, -0.016f, 0.029f); /* xyb.rs:185-190 */
        float yq = quantize_to_u8(y, 0.0f, 0.846f);
        float bq = quantize_to_u8(bb, 0.0f, 0.846f);
        /* xyb_to_linear_rgb, xyb.rs:135-166 */
        float d_r = (yq + xq) - XYB_NEG_BIAS_CBRT;
        float d_g = (yq - xq) - XYB_NEG_BIAS_CBRT;
        float d_b = bq - XYB_NEG_BIAS_CBRT;
        float p_r = mixed_cube(d_r) - XYB_BIAS;
        float p_g = mixed_cube(d_g) - XYB_BIAS;
        float p_b = mixed_cube(d_b) - XYB_BIAS;
        float lr = XYB_INV[0] * p_r + XYB_INV[1] * p_g + XYB_INV[2] * p_b;
        float lg = XYB_INV[3] * p_r + XYB_INV[4] * p_g + XYB_INV[5] * p_b;
        float lb = XYB_INV[6] * p_r + XYB_INV[7] * p_g + XYB_INV[8] * p_b;
        out[i * 3] = linear_to_srgb_u8(lr);
        out[i * 3 + 1] = linear_to_srgb_u8(lg);
        out[i * 3 + 2] = linear_to_srgb_u8(lb);
    }
}

/* the same round trip through the platform libm's cbrtf / powf (src/metrics/xyb.rs:60-100 as Rust runs it on glibc) */
CEO_API void ceo_xyb_roundtrip_libm(const uint8_t* rgb, size_t w, size_t h, uint8_t* out) {
    g_xyb_libm = 1;
    ceo_xyb_roundtrip(rgb, w, h, out);
    g_xyb_libm = 0;
}

/* ------------------------------------------------------------------ */
/* helpers: planar fp32 images                                         */
/* ------------------------------------------------------------------ */

static float* falloc(size_t n) {
    float* p = (float*)malloc((n ? n : 1) * sizeof(float));
    return p;
}

/* RGB8 interleaved -> 3 linear planes through the 256-entry table */
static void rgb8_to_linear_planes(const uint8_t* rgb, size_t n, float* r, float* g, float* b) {
    for (size_t i = 0; i < n; i++) {
        r[i] = g_lut[rgb[3 * i]];
        g[i] = g_lut[rgb[3 * i + 1]];
        b[i] = g_lut[rgb[3 * i + 2]];
    }
}

/* ------------------------------------------------------------------ */
/* A.3 SSIMULACRA2 (libjxl tools/ssimulacra2.cc == rust-av ssimulacra2 */
/*     == fast-ssim2 0.8.0; call site src/metrics/ssimulacra2.rs:96)   */
/* ------------------------------------------------------------------ */

/* Cube root of the opsin mix.  fast-ssim2 0.8.0 depends on yuvxyb 0.5.0 and yuvxyb-math 0.1.1 (Cargo.lock:410-421,
 * 1310-1330), the crates rust-av's `ssimulacra2` takes its linear RGB -> XYB conversion from (CHANGELOG.md:43:
 * "identical results").  Their cbrtf is the FreeBSD msun / musl s_cbrtf.c algorithm without the special cases: a
 * bit-level seed, then two Halley steps t <- t (2x + t^3) / (x + 2 t^3) in double ("to 16 bits", "to 47 bits"),
 * rounded once to float.  Checked exhaustively over every float in [0.0035, 1.3] (71,370,277 values): the result is
 * the correctly rounded cube root for all of them.  (Round 1 used a cheaper 0.77-ulp fp32 sequence of its own here;
 * that differed from this in 9 % of the values and moved the score by up to 0.016 at 768x512 / q90.)
 * The CUDA path reaches the same value with an fp32 sequence whose last Newton step uses an error-free residual
 * (ce_common.cuh cbrt_pos: differs from this in 349 of the 71,370,277 values, by one ulp). */
static inline float ce_cbrtf(float x) {
    const uint32_t B1 = 709958130u; /* (127 - 127.0/3 - 0.03306235651) * 2^23 */
    uint32_t hx = (asu(x) & 0x7fffffffu) / 3u + B1;
    double t = (double)asf((asu(x) & 0x80000000u) | hx);
    const double xd = (double)x;
    double r = t * t * t;
    t = t * (xd + xd + r) / (xd + r + r);
    r = t * t * t;
    t = t * (xd + xd + r) / (xd + r + r);
    return (float)t;
}

CEO_API float ceo_cbrtf(float x) { return ce_cbrtf(x); }

/* yuvxyb 0.5.0 constants (same numbers as src/metrics/xyb.rs:33-51) */
#define S2_M00 0.30f
#define S2_M01 0.622f /* 1 - 0.078 - 0.30 */
#define S2_M02 0.078f
#define S2_M10 0.23f
#define S2_M11 0.692f
#define S2_M12 0.078f
#define S2_M20 0.24342269f
#define S2_M21 0.20476745f
#define S2_M22 0.55180986f /* 1 - M20 - M21 */
#define S2_B0 0.0037930734f
#define S2_B0_ROOT 0.1559542f

static inline void s2_xyb_positive(float r, float g, float b, float* X, float* Y, float* B) {
    float m0 = fmaf(S2_M00, r, fmaf(S2_M01, g, fmaf(S2_M02, b, S2_B0)));
    float m1 = fmaf(S2_M10, r, fmaf(S2_M11, g, fmaf(S2_M12, b, S2_B0)));
    float m2 = fmaf(S2_M20, r, fmaf(S2_M21, g, fmaf(S2_M22, b, S2_B0)));
    m0 = fmaxf(m0, 0.0f);
    m1 = fmaxf(m1, 0.0f);
    m2 = fmaxf(m2, 0.0f);
    /* m >= 0; bias keeps it >= 0.0037 for non-negative rgb, guard 0 anyway */
    float c0 = (m0 > 0.0f ? ce_cbrtf(m0) : 0.0f) - S2_B0_ROOT;
    float c1 = (m1 > 0.0f ? ce_cbrtf(m1) : 0.0f) - S2_B0_ROOT;
    float c2 = (m2 > 0.0f ? ce_cbrtf(m2) : 0.0f) - S2_B0_ROOT;
    float x = 0.5f * (c0 - c1);
    float y = 0.5f * (c0 + c1);
    float bb = c2;
    /* make_positive_xyb */
    *B = (bb - y) + 0.55f;
    *X = fmaf(x, 14.0f, 0.42f);
    *Y = y + 0.01f;
}

/* recursive Gaussian sigma = 1.5 (libjxl CreateRecursiveGaussian) */
static float RG_MUL_IN[3], RG_MUL_PREV[3], RG_MUL_PREV2[3];
#define RG_N 5

static void rgauss_init(void) {
    const double sigma = 1.5;
    const double radius = round(3.2795 * sigma + 0.2546); /* 5 */
    const double pi_div_2r = M_PI / (2.0 * radius);
    const double omega[3] = {pi_div_2r, 3.0 * pi_div_2r, 5.0 * pi_div_2r};
    const double p_1 = 1.0 / tan(0.5 * omega[0]);
    const double p_3 = -1.0 / tan(0.5 * omega[1]);
    const double p_5 = 1.0 / tan(0.5 * omega[2]);
    const double r_1 = p_1 * p_1 / sin(omega[0]);
    const double r_3 = -p_3 * p_3 / sin(omega[1]);
    const double r_5 = p_5 * p_5 / sin(omega[2]);
    const double neg_half_sigma2 = -0.5 * sigma * sigma;
    const double recip_radius = 1.0 / radius;
    double rho[3];
    for (int i = 0; i < 3; i++) rho[i] = exp(neg_half_sigma2 * omega[i] * omega[i]) * recip_radius;
    const double D_13 = p_1 * r_3 - r_1 * p_3;
    const double D_35 = p_3 * r_5 - r_3 * p_5;
    const double D_51 = p_5 * r_1 - r_5 * p_1;
    const double recip_d13 = 1.0 / D_13;
    const double zeta_15 = D_35 * recip_d13;
    const double zeta_35 = D_51 * recip_d13;
    double A[3][3] = {{p_1, p_3, p_5}, {r_1, r_3, r_5}, {zeta_15, zeta_35, 1.0}};
    /* invert 3x3 by cofactors */
    double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) -
                 A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                 A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    double inv[3][3];
    inv[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det;
    inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
    inv[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
    inv[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det;
    inv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
    inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
    inv[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det;
    inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
    inv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
    const double gamma[3] = {1.0, radius * radius - sigma * sigma,
                             zeta_15 * rho[0] + zeta_35 * rho[1] + rho[2]};
    double beta[3];
    for (int i = 0; i < 3; i++)
        beta[i] = inv[i][0] * gamma[0] + inv[i][1] * gamma[1] + inv[i][2] * gamma[2];
    for (int i = 0; i < 3; i++) {
        double n2 = -beta[i] * cos(omega[i] * (radius + 1.0));
        double d1 = -2.0 * cos(omega[i]);
        RG_MUL_IN[i] = (float)n2;
        RG_MUL_PREV[i] = (float)(-d1);
        RG_MUL_PREV2[i] = -1.0f;
    }
}

CEO_API void ceo_rgauss_coeffs(float* out9) {
    for (int i = 0; i < 3; i++) {
        out9[i] = RG_MUL_IN[i];
        out9[3 + i] = RG_MUL_PREV[i];
        out9[6 + i] = RG_MUL_PREV2[i];
    }
}

/* one line of the recurrence; `stride` in floats */
static void rgauss_line(const float* in, size_t in_stride, float* out, size_t out_stride, ptrdiff_t len) {
    float p1 = 0, p3 = 0, p5 = 0, q1 = 0, q3 = 0, q5 = 0; /* prev, prev2 */
    for (ptrdiff_t n = -RG_N + 1; n < len; n++) {
        ptrdiff_t left = n - RG_N - 1, right = n + RG_N - 1;
        float lv = left >= 0 ? in[(size_t)left * in_stride] : 0.0f;
        float rv = right < len ? in[(size_t)right * in_stride] : 0.0f;
        float sum = lv + rv;
        float o1 = sum * RG_MUL_IN[0];
        float o3 = sum * RG_MUL_IN[1];
        float o5 = sum * RG_MUL_IN[2];
        o1 = fmaf(RG_MUL_PREV2[0], q1, o1);
        o3 = fmaf(RG_MUL_PREV2[1], q3, o3);
        o5 = fmaf(RG_MUL_PREV2[2], q5, o5);
        q1 = p1; q3 = p3; q5 = p5;
        o1 = fmaf(RG_MUL_PREV[0], p1, o1);
        o3 = fmaf(RG_MUL_PREV[1], p3, o3);
        o5 = fmaf(RG_MUL_PREV[2], p5, o5);
        p1 = o1; p3 = o3; p5 = o5;
        if (n >= 0) out[(size_t)n * out_stride] = (o1 + o3) + o5;
    }
}

/* horizontal then vertical, zero outside the image */
static void rgauss_blur(const float* in, size_t w, size_t h, float* tmp, float* out) {
    for (size_t y = 0; y < h; y++) rgauss_line(in + y * w, 1, tmp + y * w, 1, (ptrdiff_t)w);
    /* vertical: all columns advance together (same per-column op sequence) */
    float* st = (float*)calloc(6 * w, sizeof(float));
    float *p1 = st, *p3 = st + w, *p5 = st + 2 * w, *q1 = st + 3 * w, *q3 = st + 4 * w, *q5 = st + 5 * w;
    ptrdiff_t len = (ptrdiff_t)h;
    for (ptrdiff_t n = -RG_N + 1; n < len; n++) {
        ptrdiff_t left = n - RG_N - 1, right = n + RG_N - 1;
        const float* lrow = left >= 0 ? tmp + (size_t)left * w : NULL;
        const float* rrow = right < len ? tmp + (size_t)right * w : NULL;
        float* orow = n >= 0 ? out + (size_t)n * w : NULL;
        for (size_t x = 0; x < w; x++) {
            float lv = lrow ? lrow[x] : 0.0f;
            float rv = rrow ? rrow[x] : 0.0f;
            float sum = lv + rv;
            float o1 = sum * RG_MUL_IN[0];
            float o3 = sum * RG_MUL_IN[1];
            float o5 = sum * RG_MUL_IN[2];
            o1 = fmaf(RG_MUL_PREV2[0], q1[x], o1);
            o3 = fmaf(RG_MUL_PREV2[1], q3[x], o3);
            o5 = fmaf(RG_MUL_PREV2[2], q5[x], o5);
            q1[x] = p1[x]; q3[x] = p3[x]; q5[x] = p5[x];
            o1 = fmaf(RG_MUL_PREV[0], p1[x], o1);
            o3 = fmaf(RG_MUL_PREV[1], p3[x], o3);
            o5 = fmaf(RG_MUL_PREV[2], p5[x], o5);
            p1[x] = o1; p3[x] = o3; p5[x] = o5;
            if (orow) orow[x] = (o1 + o3) + o5;
        }
    }
    free(st);
}

CEO_API void ceo_rgauss_blur(const float* in, size_t w, size_t h, float* out) {
    float* tmp = falloc(w * h);
    rgauss_blur(in, w, h, tmp, out);
    free(tmp);
}

/* down2 on linear RGB: ceil size, clamp coordinates, sum (iy,ix) order, *0.25 */
static void s2_down2(const float* in, size_t w, size_t h, float* out, size_t ow, size_t oh) {
    for (size_t oy = 0; oy < oh; oy++)
        for (size_t ox = 0; ox < ow; ox++) {
            float sum = 0.0f;
            for (size_t iy = 0; iy < 2; iy++)
                for (size_t ix = 0; ix < 2; ix++) {
                    size_t x = ox * 2 + ix; if (x > w - 1) x = w - 1;
                    size_t y = oy * 2 + iy; if (y > h - 1) y = h - 1;
                    sum += in[y * w + x];
                }
            out[oy * ow + ox] = sum * 0.25f;
        }
}

static const double S2_WEIGHT[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0, 0.0004371155730107379, 0.0,
    1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0,
    1.8422455520539298, 11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072, 0.9677937080626833, 0.0,
    0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0,
    0.0013648766163243398, 0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264,
    225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985,
    176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0008680556573291698, 0.0, 0.0,
    0.0, 0.0, 0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0,
    0.0004179171803251336, 0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862,
    23.19243343998926, 0.0, 95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493,
    171.2667255897307, 0.9807858872435379, 0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537};

/* per-scale averages: avg[scale*18 + 0..5] = avg_ssim[c*2+n], avg[scale*18 + 6..17] = avg_edgediff[c*4+k] */
CEO_API double ceo_ssimulacra2_score_from_avgs(const double* avg, int nscales) {
    double ssim = 0.0;
    size_t i = 0;
    for (int c = 0; c < 3; c++)
        for (int s = 0; s < nscales; s++)
            for (int n = 0; n < 2; n++) {
                const double* a = avg + (size_t)s * 18;
                ssim = fma(S2_WEIGHT[i++], fabs(a[c * 2 + n]), ssim);
                ssim = fma(S2_WEIGHT[i++], fabs(a[6 + c * 4 + n]), ssim);
                ssim = fma(S2_WEIGHT[i++], fabs(a[6 + c * 4 + n + 2]), ssim);
            }
    ssim *= 0.9562382616834844;
    ssim = fma(6.248496625763138e-5 * ssim * ssim, ssim,
               fma(2.326765642916932, ssim, -0.020884521182843837 * ssim * ssim));
    if (ssim > 0.0) ssim = fma(pow(ssim, 0.6276336467831387), -10.0, 100.0);
    else ssim = 100.0;
    return ssim;
}

/* sums (not yet divided by pixel count) of one scale: sums[c*6 + {0:d,1:d^4,2:art,3:art^4,4:det,5:det^4}] */
static void s2_scale_sums(float* const lin1[3], float* const lin2[3], size_t w, size_t h, double* sums,
                          float* dbg_planes /* optional: 3*(i1,i2,mu1,mu2,s11,s22,s12) */) {
    size_t n = w * h;
    float* x1[3]; float* x2[3];
    for (int c = 0; c < 3; c++) { x1[c] = falloc(n); x2[c] = falloc(n); }
    for (size_t i = 0; i < n; i++) {
        s2_xyb_positive(lin1[0][i], lin1[1][i], lin1[2][i], &x1[0][i], &x1[1][i], &x1[2][i]);
        s2_xyb_positive(lin2[0][i], lin2[1][i], lin2[2][i], &x2[0][i], &x2[1][i], &x2[2][i]);
    }
    float* mul = falloc(n); float* tmp = falloc(n);
    float* mu1 = falloc(n); float* mu2 = falloc(n);
    float* s11 = falloc(n); float* s22 = falloc(n); float* s12 = falloc(n);
    for (int c = 0; c < 3; c++) {
        const float* i1 = x1[c]; const float* i2 = x2[c];
        for (size_t i = 0; i < n; i++) mul[i] = i1[i] * i1[i];
        rgauss_blur(mul, w, h, tmp, s11);
        for (size_t i = 0; i < n; i++) mul[i] = i2[i] * i2[i];
        rgauss_blur(mul, w, h, tmp, s22);
        for (size_t i = 0; i < n; i++) mul[i] = i1[i] * i2[i];
        rgauss_blur(mul, w, h, tmp, s12);
        rgauss_blur(i1, w, h, tmp, mu1);
        rgauss_blur(i2, w, h, tmp, mu2);
        double sd = 0, sd4 = 0, sa = 0, sa4 = 0, sl = 0, sl4 = 0;
        for (size_t i = 0; i < n; i++) {
            /* ssim_map */
            float m1 = mu1[i], m2 = mu2[i];
            float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            float mdiff = m1 - m2;
            float num_m = fmaf(mdiff, -mdiff, 1.0f);
            float num_s = fmaf(2.0f, s12[i] - m12, 0.0009f);
            float denom_s = ((s11[i] - m11) + (s22[i] - m22)) + 0.0009f;
            double d = 1.0 - (double)((num_m * num_s) / denom_s);
            if (!(d > 0.0)) d = 0.0;
            double d2 = d * d;
            sd += d; sd4 += d2 * d2;
            /* edge_diff_map */
            double d1 = (1.0 + (double)fabsf(i2[i] - m2)) / (1.0 + (double)fabsf(i1[i] - m1)) - 1.0;
            double art = d1 > 0.0 ? d1 : 0.0;
            double det = d1 < 0.0 ? -d1 : 0.0;
            double a2 = art * art, l2 = det * det;
            sa += art; sa4 += a2 * a2; sl += det; sl4 += l2 * l2;
        }
        sums[c * 6 + 0] = sd; sums[c * 6 + 1] = sd4; sums[c * 6 + 2] = sa;
        sums[c * 6 + 3] = sa4; sums[c * 6 + 4] = sl; sums[c * 6 + 5] = sl4;
        if (dbg_planes) {
            float* d = dbg_planes + (size_t)c * 7 * n;
            memcpy(d, i1, n * 4); memcpy(d + n, i2, n * 4); memcpy(d + 2 * n, mu1, n * 4);
            memcpy(d + 3 * n, mu2, n * 4); memcpy(d + 4 * n, s11, n * 4); memcpy(d + 5 * n, s22, n * 4);
            memcpy(d + 6 * n, s12, n * 4);
        }
    }
    free(mul); free(tmp); free(mu1); free(mu2); free(s11); free(s22); free(s12);
    for (int c = 0; c < 3; c++) { free(x1[c]); free(x2[c]); }
}

/* sums -> the 18 averages of one scale, in the layout of ceo_ssimulacra2_score_from_avgs */
CEO_API void ceo_ssimulacra2_avgs_from_sums(const double* sums, size_t w, size_t h, double* avg) {
    double opp = 1.0 / (double)(w * h);
    for (int c = 0; c < 3; c++) {
        avg[c * 2 + 0] = opp * sums[c * 6 + 0];
        avg[c * 2 + 1] = sqrt(sqrt(opp * sums[c * 6 + 1]));
        avg[6 + c * 4 + 0] = opp * sums[c * 6 + 2];
        avg[6 + c * 4 + 1] = sqrt(sqrt(opp * sums[c * 6 + 3]));
        avg[6 + c * 4 + 2] = opp * sums[c * 6 + 4];
        avg[6 + c * 4 + 3] = sqrt(sqrt(opp * sums[c * 6 + 5]));
    }
}

/* Full metric.  sums_out (optional): 6 scales x 18 raw sums; nscales_out optional. */
CEO_API int ceo_ssimulacra2_ex(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, double* score,
                               double* sums_out, int* nscales_out) {
    if (w < 8 || h < 8) return CEO_METRIC_CALCULATION;
    size_t n = w * h;
    float* l1[3]; float* l2[3];
    for (int c = 0; c < 3; c++) { l1[c] = falloc(n); l2[c] = falloc(n); }
    rgb8_to_linear_planes(ref, n, l1[0], l1[1], l1[2]);
    rgb8_to_linear_planes(dist, n, l2[0], l2[1], l2[2]);
    double avgs[6 * 18];
    int ns = 0;
    size_t cw = w, ch = h;
    for (int scale = 0; scale < 6; scale++) {
        if (cw < 8 || ch < 8) break;
        if (scale > 0) {
            size_t ow = (cw + 1) / 2, oh = (ch + 1) / 2;
            for (int c = 0; c < 3; c++) {
                float* o1 = falloc(ow * oh); float* o2 = falloc(ow * oh);
                s2_down2(l1[c], cw, ch, o1, ow, oh);
                s2_down2(l2[c], cw, ch, o2, ow, oh);
                free(l1[c]); free(l2[c]);
                l1[c] = o1; l2[c] = o2;
            }
            cw = ow; ch = oh;
            /* NB upstream re-checks nothing here: a scale whose halved size fell below 8
             * is still evaluated (the check is on the pre-halving size). */
        }
        double sums[18];
        s2_scale_sums(l1, l2, cw, ch, sums, NULL);
        if (sums_out) memcpy(sums_out + (size_t)ns * 18, sums, sizeof(sums));
        ceo_ssimulacra2_avgs_from_sums(sums, cw, ch, avgs + (size_t)ns * 18);
        ns++;
    }
    for (int c = 0; c < 3; c++) { free(l1[c]); free(l2[c]); }
    if (nscales_out) *nscales_out = ns;
    *score = ceo_ssimulacra2_score_from_avgs(avgs, ns);
    return CEO_OK;
}

CEO_API int ceo_ssimulacra2(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, double* score) {
    return ceo_ssimulacra2_ex(ref, dist, w, h, score, NULL, NULL);
}

/* debug: planes of scale 0: 3 channels x (i1,i2,mu1,mu2,s11,s22,s12) */
CEO_API void ceo_ssimulacra2_scale0_planes(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, float* planes) {
    size_t n = w * h;
    float* l1[3]; float* l2[3];
    for (int c = 0; c < 3; c++) { l1[c] = falloc(n); l2[c] = falloc(n); }
    rgb8_to_linear_planes(ref, n, l1[0], l1[1], l1[2]);
    rgb8_to_linear_planes(dist, n, l2[0], l2[1], l2[2]);
    double sums[18];
    s2_scale_sums(l1, l2, w, h, sums, planes);
    for (int c = 0; c < 3; c++) { free(l1[c]); free(l2[c]); }
}

/* ------------------------------------------------------------------ */
/* A.4 DSSIM (dssim-core 3.4.0; call site src/metrics/dssim.rs:52-68)  */
/* ------------------------------------------------------------------ */

static const float DS_KERNEL[9] = {0.095332f, 0.118095f, 0.095332f, 0.118095f, 0.146293f,
                                   0.118095f, 0.095332f, 0.118095f, 0.095332f};
static const double DS_WEIGHTS[5] = {0.028, 0.197, 0.322, 0.298, 0.155};

/* one 3x3 pass, clamp-replicate edges, nine products summed row-major by rows */
static void ds_blur_pass(const float* src, size_t w, size_t h, float* dst) {
    for (size_t y = 0; y < h; y++) {
        const float* prev = src + (y > 0 ? y - 1 : 0) * w;
        const float* curr = src + y * w;
        const float* next = src + (y + 1 < h ? y + 1 : y) * w;
        for (size_t x = 0; x < w; x++) {
            size_t c0 = x > 0 ? x - 1 : 0, c1 = x, c2 = x + 1 < w ? x + 1 : x;
            float a = (prev[c0] * DS_KERNEL[0] + prev[c1] * DS_KERNEL[1]) + prev[c2] * DS_KERNEL[2];
            float b = (curr[c0] * DS_KERNEL[3] + curr[c1] * DS_KERNEL[4]) + curr[c2] * DS_KERNEL[5];
            float c = (next[c0] * DS_KERNEL[6] + next[c1] * DS_KERNEL[7]) + next[c2] * DS_KERNEL[8];
            dst[y * w + x] = (a + b) + c;
        }
    }
}
/* "blur" = the 3x3 pass applied twice */
static void ds_blur(const float* src, size_t w, size_t h, float* tmp, float* dst) {
    ds_blur_pass(src, w, h, tmp);
    ds_blur_pass(tmp, w, h, dst);
}
CEO_API void ceo_dssim_blur(const float* in, size_t w, size_t h, float* out) {
    float* tmp = falloc(w * h);
    ds_blur(in, w, h, tmp, out);
    free(tmp);
}

static inline float ds_cbrt_poly(float x) {
    float y = (-0.5f * x + 1.51f) * x + 0.2f;
    float y3 = (y * y) * y;
    y = (y * (y3 + 2.0f * x)) / (2.0f * y3 + x);
    y3 = (y * y) * y;
    y = (y * (y3 + 2.0f * x)) / (2.0f * y3 + x);
    return y;
}
static inline float ds_fma_matrix(float r, float rx, float g, float gx, float b, float bx) {
    return fmaf(b, bx, fmaf(g, gx, r * rx));
}
#define DS_D65X 0.9505f
#define DS_D65Z 1.089f
static inline void ds_to_lab(float r, float g, float b, float* L, float* A, float* B) {
    const float eps = 216.0f / 24389.0f;
    const float k = 24389.0f / (27.0f * 116.0f);
    float fx = ds_fma_matrix(r, 0.4124f / DS_D65X, g, 0.3576f / DS_D65X, b, 0.1805f / DS_D65X);
    float fy = ds_fma_matrix(r, 0.2126f, g, 0.7152f, b, 0.0722f);
    float fz = ds_fma_matrix(r, 0.0193f / DS_D65Z, g, 0.1192f / DS_D65Z, b, 0.9505f / DS_D65Z);
    float X = fx > eps ? ds_cbrt_poly(fx) - 16.0f / 116.0f : k * fx;
    float Y = fy > eps ? ds_cbrt_poly(fy) - 16.0f / 116.0f : k * fy;
    float Z = fz > eps ? ds_cbrt_poly(fz) - 16.0f / 116.0f : k * fz;
    *L = Y * 1.05f;
    *A = fmaf(500.0f / 220.0f, X - Y, 86.2f / 220.0f);
    *B = fmaf(200.0f / 220.0f, Y - Z, 107.9f / 220.0f);
}
CEO_API void ceo_dssim_lab(float r, float g, float b, float* out3) { ds_to_lab(r, g, b, out3, out3 + 1, out3 + 2); }

typedef struct {
    size_t w, h;
    float* img[3];  /* L, a(pre-blurred), b(pre-blurred) */
    float* mu[3];
    float* sq[3];   /* blur(img^2) */
} ds_scale;

static void ds_scale_free(ds_scale* s) {
    for (int c = 0; c < 3; c++) { free(s->img[c]); free(s->mu[c]); free(s->sq[c]); }
}

/* planes rgba[4] linear (alpha plane may be NULL = 1.0) at size w x h -> scale statistics */
static void ds_make_scale(float* const rgba[4], size_t w, size_t h, ds_scale* s) {
    size_t n = w * h;
    s->w = w; s->h = h;
    for (int c = 0; c < 3; c++) { s->img[c] = falloc(n); s->mu[c] = falloc(n); s->sq[c] = falloc(n); }
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            size_t i = y * w + x;
            float r = rgba[0][i], g = rgba[1][i], b = rgba[2][i];
            if (rgba[3]) { /* dssim-core ToRGB for RGBAPLU: blend on a position-dependent background */
                float a = rgba[3][i];
                if (a < 255.0f / 256.0f) {
                    size_t nn = (x + 11) ^ (y + 11);
                    if (nn & 16) r += 1.0f - a;
                    if (nn & 8) g += 1.0f - a;
                    if (nn & 32) b += 1.0f - a;
                }
            }
            ds_to_lab(r, g, b, &s->img[0][i], &s->img[1][i], &s->img[2][i]);
        }
    float* tmp = falloc(n); float* t2 = falloc(n);
    for (int c = 0; c < 3; c++) {
        if (c > 0) { /* chroma pre-blur, in place */
            ds_blur(s->img[c], w, h, tmp, t2);
            memcpy(s->img[c], t2, n * sizeof(float));
        }
        ds_blur(s->img[c], w, h, tmp, s->mu[c]);
        for (size_t i = 0; i < n; i++) t2[i] = s->img[c][i] * s->img[c][i];
        ds_blur(t2, w, h, tmp, s->sq[c]);
    }
    free(tmp); free(t2);
}

/* floor-halving, odd last row/column cropped, (a+b+c+d)*0.25 */
static int ds_down(float* rgba[4], size_t* pw, size_t* ph) {
    size_t w = *pw, h = *ph, hw = w / 2, hh = h / 2;
    if (hw < 4 || hh < 4) return 0;
    for (int c = 0; c < 4; c++) {
        if (!rgba[c]) continue;
        float* o = falloc(hw * hh);
        for (size_t y = 0; y < hh; y++)
            for (size_t x = 0; x < hw; x++) {
                const float* top = rgba[c] + (2 * y) * w + 2 * x;
                const float* bot = top + w;
                o[y * hw + x] = (((top[0] + top[1]) + bot[0]) + bot[1]) * 0.25f;
            }
        free(rgba[c]);
        rgba[c] = o;
    }
    *pw = hw; *ph = hh;
    return 1;
}

/* ssim map of one scale + its pooled score; map (optional) receives w*h floats */
static double ds_compare_scale(const ds_scale* a, const ds_scale* b, int scale_idx, float* map_out, double* mean_out) {
    size_t w = a->w, h = a->h, n = w * h;
    float* cross[3]; float* tmp = falloc(n); float* prod = falloc(n);
    for (int c = 0; c < 3; c++) {
        cross[c] = falloc(n);
        for (size_t i = 0; i < n; i++) prod[i] = a->img[c][i] * b->img[c][i];
        ds_blur(prod, w, h, tmp, cross[c]);
    }
    float* map = map_out ? map_out : falloc(n);
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    const float third = 1.0f / 3.0f;
    for (size_t i = 0; i < n; i++) {
        float m11[3], m12[3], m22[3], s1[3], s2[3], s12[3];
        for (int c = 0; c < 3; c++) {
            float mu1 = a->mu[c][i], mu2 = b->mu[c][i];
            m11[c] = mu1 * mu1; m12[c] = mu1 * mu2; m22[c] = mu2 * mu2;
            s1[c] = a->sq[c][i] - m11[c];
            s2[c] = b->sq[c][i] - m22[c];
            s12[c] = cross[c][i] - m12[c];
        }
        float mu1_sq = ((m11[0] + m11[1]) + m11[2]) * third;
        float mu2_sq = ((m22[0] + m22[1]) + m22[2]) * third;
        float mu1_mu2 = ((m12[0] + m12[1]) + m12[2]) * third;
        float sigma1_sq = ((s1[0] + s1[1]) + s1[2]) * third;
        float sigma2_sq = ((s2[0] + s2[1]) + s2[2]) * third;
        float sigma12 = ((s12[0] + s12[1]) + s12[2]) * third;
        map[i] = (fmaf(2.0f, mu1_mu2, c1) * fmaf(2.0f, sigma12, c2)) /
                 (((mu1_sq + mu2_sq) + c1) * ((sigma1_sq + sigma2_sq) + c2));
    }
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += (double)map[i];
    double len = (double)n;
    double mean = sum / len;
    double avg = pow(mean > 0.0 ? mean : 0.0, pow(0.5, (double)scale_idx));
    double dev = 0.0;
    for (size_t i = 0; i < n; i++) dev += fabs(avg - (double)map[i]);
    double score = 1.0 - dev / len;
    if (mean_out) *mean_out = mean;
    for (int c = 0; c < 3; c++) free(cross[c]);
    free(tmp); free(prod);
    if (!map_out) free(map);
    return score;
}

CEO_API double ceo_dssim_from_scale_scores(const double* scores, int nscales) {
    double ssim_sum = 0.0, weight_sum = 0.0;
    for (int i = 0; i < nscales; i++) { ssim_sum += scores[i] * DS_WEIGHTS[i]; weight_sum += DS_WEIGHTS[i]; }
    double ssim = ssim_sum / weight_sum;
    if (ssim < 2.220446049250313e-16) ssim = 2.220446049250313e-16;
    return 1.0 / ssim - 1.0;
}

/* core on planar linear RGBA (alpha planes may be NULL); consumes the planes */
static int ds_core(float* p1[4], float* p2[4], size_t w, size_t h, double* out, double* scale_scores, int* nscales_out,
                   float* map0) {
    double scores[5];
    int ns = 0;
    size_t cw = w, ch = h;
    for (int s = 0; s < 5; s++) {
        if (s > 0) {
            size_t w2 = cw, h2 = ch;
            if (!ds_down(p1, &cw, &ch)) break;
            ds_down(p2, &w2, &h2);
        }
        ds_scale a, b;
        ds_make_scale(p1, cw, ch, &a);
        ds_make_scale(p2, cw, ch, &b);
        scores[ns] = ds_compare_scale(&a, &b, s, (s == 0) ? map0 : NULL, NULL);
        ds_scale_free(&a); ds_scale_free(&b);
        ns++;
    }
    for (int c = 0; c < 4; c++) { free(p1[c]); free(p2[c]); }
    if (scale_scores) memcpy(scale_scores, scores, sizeof(double) * (size_t)ns);
    if (nscales_out) *nscales_out = ns;
    *out = ceo_dssim_from_scale_scores(scores, ns);
    return CEO_OK;
}

CEO_API int ceo_dssim_ex(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, double* out,
                         double* scale_scores, int* nscales_out, float* map0) {
    if (w == 0 || h == 0) return CEO_METRIC_CALCULATION;
    size_t n = w * h;
    float* p1[4] = {falloc(n), falloc(n), falloc(n), NULL};
    float* p2[4] = {falloc(n), falloc(n), falloc(n), NULL};
    rgb8_to_linear_planes(ref, n, p1[0], p1[1], p1[2]);
    rgb8_to_linear_planes(dist, n, p2[0], p2[1], p2[2]);
    return ds_core(p1, p2, w, h, out, scale_scores, nscales_out, map0);
}
CEO_API int ceo_dssim(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, double* out) {
    return ceo_dssim_ex(ref, dist, w, h, out, NULL, NULL, NULL);
}
/* linear RGBA f32 interleaved, stride in pixels (src/metrics/dssim.rs:40-71 input type) */
CEO_API int ceo_dssim_rgbaf32(const float* ref, const float* dist, size_t w, size_t h, size_t stride, double* out) {
    if (w == 0 || h == 0) return CEO_METRIC_CALCULATION;
    size_t n = w * h;
    float* p1[4]; float* p2[4];
    for (int c = 0; c < 4; c++) { p1[c] = falloc(n); p2[c] = falloc(n); }
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++)
            for (int c = 0; c < 4; c++) {
                p1[c][y * w + x] = ref[(y * stride + x) * 4 + (size_t)c];
                p2[c][y * w + x] = dist[(y * stride + x) * 4 + (size_t)c];
            }
    return ds_core(p1, p2, w, h, out, NULL, NULL, NULL);
}

/* src/metrics/dssim.rs:102-114 / :131-143: RGB8 / RGBA8 -> linear RGBA f32 interleaved */
CEO_API void ceo_rgb8_to_dssim_image(const uint8_t* data, size_t w, size_t h, float* out_rgba) {
    for (size_t i = 0; i < w * h; i++) {
        out_rgba[4 * i] = g_lut[data[3 * i]];
        out_rgba[4 * i + 1] = g_lut[data[3 * i + 1]];
        out_rgba[4 * i + 2] = g_lut[data[3 * i + 2]];
        out_rgba[4 * i + 3] = 1.0f;
    }
}
CEO_API void ceo_rgba8_to_dssim_image(const uint8_t* data, size_t w, size_t h, float* out_rgba) {
    for (size_t i = 0; i < w * h; i++) {
        out_rgba[4 * i] = g_lut[data[4 * i]];
        out_rgba[4 * i + 1] = g_lut[data[4 * i + 1]];
        out_rgba[4 * i + 2] = g_lut[data[4 * i + 2]];
        out_rgba[4 * i + 3] = (float)data[4 * i + 3] / 255.0f;
    }
}

/* ------------------------------------------------------------------ */
/* A.5 Butteraugli (libjxl butteraugli.cc == butteraugli 0.9.0;        */
/*     call site src/metrics/butteraugli.rs:72-80,127-135)             */
/* ------------------------------------------------------------------ */

#define BA_MAX_TAPS 65

typedef struct { int radius; float w[BA_MAX_TAPS]; } ba_kernel;

static void ba_make_kernel(float sigma, ba_kernel* k) {
    const float m = 2.25f;
    const double scaler = -1.0 / (2.0 * (double)sigma * (double)sigma);
    int diff = (int)(m * fabsf(sigma));
    if (diff < 1) diff = 1;
    k->radius = diff;
    for (int i = -diff; i <= diff; i++) k->w[i + diff] = (float)exp(scaler * (double)i * (double)i);
}
CEO_API int ceo_ba_kernel(float sigma, float* w_out) {
    ba_kernel k; ba_make_kernel(sigma, &k);
    memcpy(w_out, k.w, sizeof(float) * (size_t)(2 * k.radius + 1));
    return k.radius;
}

/* 1-D pass along a line: truncated kernel, renormalised by the in-range tap sum */
static void ba_conv_line(const float* in, size_t istride, float* out, size_t ostride, ptrdiff_t len, const ba_kernel* k,
                         const float* inv_weight /* per position 1/sum(in-range taps) */) {
    int R = k->radius;
    for (ptrdiff_t x = 0; x < len; x++) {
        ptrdiff_t lo = x - R < 0 ? 0 : x - R;
        ptrdiff_t hi = x + R > len - 1 ? len - 1 : x + R;
        float sum = 0.0f;
        for (ptrdiff_t j = lo; j <= hi; j++) sum = fmaf(in[(size_t)j * istride], k->w[j - x + R], sum); /* libjxl MulAdd */
        out[(size_t)x * ostride] = sum * inv_weight[x];
    }
}
static void ba_inv_weights(ptrdiff_t len, const ba_kernel* k, float* inv) {
    int R = k->radius;
    for (ptrdiff_t x = 0; x < len; x++) {
        ptrdiff_t lo = x - R < 0 ? 0 : x - R;
        ptrdiff_t hi = x + R > len - 1 ? len - 1 : x + R;
        float wsum = 0.0f;
        for (ptrdiff_t j = lo; j <= hi; j++) wsum += k->w[j - x + R];
        inv[x] = 1.0f / wsum;
    }
}
static inline ptrdiff_t ba_mirror(ptrdiff_t x, ptrdiff_t n) {
    while (x < 0 || x >= n) { if (x < 0) x = -x - 1; else x = 2 * n - 1 - x; }
    return x;
}
/* 5-tap special case (sigma 1.2): mirror padding, pre-normalised weights */
static void ba_conv5_line(const float* in, size_t istride, float* out, size_t ostride, ptrdiff_t len, float w0, float w1, float w2) {
    for (ptrdiff_t x = 0; x < len; x++) {
        float c = in[(size_t)x * istride];
        float l1 = in[(size_t)ba_mirror(x - 1, len) * istride], r1 = in[(size_t)ba_mirror(x + 1, len) * istride];
        float l2 = in[(size_t)ba_mirror(x - 2, len) * istride], r2 = in[(size_t)ba_mirror(x + 2, len) * istride];
        out[(size_t)x * ostride] = fmaf(l2 + r2, w2, fmaf(l1 + r1, w1, c * w0)); /* libjxl Separable5 MulAdd chain */
    }
}

/* separable blur, horizontal then vertical */
static void ba_blur(const float* in, size_t w, size_t h, float sigma, float* out) {
    ba_kernel k; ba_make_kernel(sigma, &k);
    float* tmp = falloc(w * h);
    if (k.radius == 2) {
        float sw = 0.0f;
        for (int i = 0; i < 5; i++) sw += k.w[i];
        float scale = 1.0f / sw;
        float w0 = k.w[2] * scale, w1 = k.w[1] * scale, w2 = k.w[0] * scale;
        for (size_t y = 0; y < h; y++) ba_conv5_line(in + y * w, 1, tmp + y * w, 1, (ptrdiff_t)w, w0, w1, w2);
        for (size_t x = 0; x < w; x++) ba_conv5_line(tmp + x, w, out + x, w, (ptrdiff_t)h, w0, w1, w2);
    } else {
        float* invx = falloc(w); float* invy = falloc(h);
        ba_inv_weights((ptrdiff_t)w, &k, invx);
        ba_inv_weights((ptrdiff_t)h, &k, invy);
        for (size_t y = 0; y < h; y++) ba_conv_line(in + y * w, 1, tmp + y * w, 1, (ptrdiff_t)w, &k, invx);
        /* vertical: row-major friendly accumulation, same per-pixel op order (taps ascending) */
        int R = k.radius;
        for (size_t y = 0; y < h; y++) {
            ptrdiff_t lo = (ptrdiff_t)y - R < 0 ? 0 : (ptrdiff_t)y - R;
            ptrdiff_t hi = (ptrdiff_t)y + R > (ptrdiff_t)h - 1 ? (ptrdiff_t)h - 1 : (ptrdiff_t)y + R;
            float* o = out + y * w;
            for (size_t x = 0; x < w; x++) o[x] = 0.0f;
            for (ptrdiff_t j = lo; j <= hi; j++) {
                float wt = k.w[j - (ptrdiff_t)y + R];
                const float* r = tmp + (size_t)j * w;
                for (size_t x = 0; x < w; x++) o[x] = fmaf(r[x], wt, o[x]);
            }
            for (size_t x = 0; x < w; x++) o[x] *= invy[y];
        }
        free(invx); free(invy);
    }
    free(tmp);
}
CEO_API void ceo_ba_blur(const float* in, size_t w, size_t h, float sigma, float* out) { ba_blur(in, w, h, sigma, out); }

/* libjxl FastLog2f: (2,2) rational approximation of log2 on [2/3, 4/3] + exponent */
static inline float ba_fast_log2f(float x) {
    int32_t xb = (int32_t)asu(x);
    int32_t eb = xb - 0x3f2aaaab;
    int32_t es = eb >> 23;
    float mant = asf((uint32_t)(xb - (es << 23)));
    float ev = (float)es;
    float t = mant - 1.0f;
    float yp = fmaf(fmaf(7.4245873327820566E-01f, t, 1.4287160470083755E+00f), t, -1.8503833400518310E-06f);
    float yq = fmaf(fmaf(1.7409343003366853E-01f, t, 1.0096718572241148E+00f), t, 9.9032814277590719E-01f);
    return yp / yq + ev;
}
CEO_API float ceo_ba_fast_log2f(float x) { return ba_fast_log2f(x); }

static inline float ba_gamma(float v) {
    const float kRetMul = 19.245013259874995f * 0.693147181f;
    const float kRetAdd = -23.16046239805755f;
    if (v < 0.0f) v = 0.0f;
    float biased = v + 9.9710635769299145f;
    return fmaf(kRetMul, ba_fast_log2f(biased), kRetAdd);
}
static inline void ba_opsin_absorbance(float r, float g, float b, float* o0, float* o1, float* o2) {
    *o0 = fmaf(0.29956550340058319f, r, fmaf(0.63373087833825936f, g, fmaf(0.077705617820981968f, b, 1.7557483643287353f)));
    *o1 = fmaf(0.22158691104574774f, r, fmaf(0.69391388044116142f, g, fmaf(0.0987313588422f, b, 1.7557483643287353f)));
    *o2 = fmaf(0.02f, r, fmaf(0.02f, g, fmaf(0.20480129041026129f, b, 12.226454707163354f)));
}

/* rgb: linear planes (1.0 = white); xyb out */
static void ba_opsin_dynamics(float* const rgb[3], size_t w, size_t h, float intensity, float* xyb[3]) {
    size_t n = w * h;
    float* bl[3];
    for (int c = 0; c < 3; c++) { bl[c] = falloc(n); ba_blur(rgb[c], w, h, 1.2f, bl[c]); }
    for (size_t i = 0; i < n; i++) {
        float p0, p1, p2;
        ba_opsin_absorbance(bl[0][i] * intensity, bl[1][i] * intensity, bl[2][i] * intensity, &p0, &p1, &p2);
        p0 = fmaxf(p0, 1e-4f); p1 = fmaxf(p1, 1e-4f); p2 = fmaxf(p2, 1e-4f);
        float s0 = fmaxf(ba_gamma(p0) / p0, 1e-4f);
        float s1 = fmaxf(ba_gamma(p1) / p1, 1e-4f);
        float s2 = fmaxf(ba_gamma(p2) / p2, 1e-4f);
        float c0, c1, c2;
        ba_opsin_absorbance(rgb[0][i] * intensity, rgb[1][i] * intensity, rgb[2][i] * intensity, &c0, &c1, &c2);
        c0 *= s0; c1 *= s1; c2 *= s2;
        c0 = fmaxf(c0, 1.7557483643287353f);
        c1 = fmaxf(c1, 1.7557483643287353f);
        c2 = fmaxf(c2, 12.226454707163354f);
        xyb[0][i] = c0 - c1;
        xyb[1][i] = c0 + c1;
        xyb[2][i] = c2;
    }
    for (int c = 0; c < 3; c++) free(bl[c]);
}

static inline float ba_remove_range(float v, float w) { return v > w ? v - w : (v < -w ? v + w : 0.0f); }
static inline float ba_amplify_range(float v, float w) { return v > w ? v + w : (v < -w ? v - w : v + v); }
static inline float ba_max_clamp(float v, float maxval) {
    const float kMul = 0.724216145665f;
    float if_pos = fmaf(v - maxval, kMul, maxval);
    float if_neg = fmaf(v + maxval, kMul, -maxval);
    float pos_or_v = v >= maxval ? if_pos : v;
    return v < -maxval ? if_neg : pos_or_v;
}

typedef struct { size_t w, h; float* lf[3]; float* mf[3]; float* hf[2]; float* uhf[2]; } ba_psycho;

static void ba_psycho_free(ba_psycho* p) {
    for (int c = 0; c < 3; c++) { free(p->lf[c]); free(p->mf[c]); }
    for (int c = 0; c < 2; c++) { free(p->hf[c]); free(p->uhf[c]); }
}

static void ba_separate_frequencies(float* const xyb[3], size_t w, size_t h, ba_psycho* ps) {
    size_t n = w * h;
    ps->w = w; ps->h = h;
    for (int c = 0; c < 3; c++) { ps->lf[c] = falloc(n); ps->mf[c] = falloc(n); }
    for (int c = 0; c < 2; c++) { ps->hf[c] = falloc(n); ps->uhf[c] = falloc(n); }
    float* t = falloc(n);
    for (int i = 0; i < 3; i++) {
        ba_blur(xyb[i], w, h, 7.15593339443f, ps->lf[i]);
        for (size_t k = 0; k < n; k++) t[k] = xyb[i][k] - ps->lf[i][k];     /* mf (unblurred) */
        ba_blur(t, w, h, 3.22489901262f, ps->mf[i]);
        if (i == 2) break;
        for (size_t k = 0; k < n; k++) {
            float mf = ps->mf[i][k];
            float hf = t[k] - mf;
            mf = (i == 0) ? ba_remove_range(mf, 0.29f) : ba_amplify_range(mf, 0.1f);
            ps->mf[i][k] = mf;
            ps->hf[i][k] = hf;
        }
    }
    /* SuppressXByY(in_y = hf[1], inout_x = hf[0]) */
    for (size_t k = 0; k < n; k++) {
        float vx = ps->hf[0][k], vy = ps->hf[1][k];
        float scaler = fmaf(46.0f / fmaf(vy, vy, 46.0f), (float)(1.0 - 0.653020556257), 0.653020556257f);
        ps->hf[0][k] = scaler * vx;
    }
    for (int i = 0; i < 2; i++) {
        memcpy(t, ps->hf[i], n * sizeof(float));                              /* uhf <- hf */
        ba_blur(t, w, h, 1.56416327805f, ps->hf[i]);
        for (size_t k = 0; k < n; k++) {
            float hf = ps->hf[i][k];
            float uhf;
            if (i == 0) {
                uhf = t[k] - hf;
                hf = ba_remove_range(hf, 1.5f);
                uhf = ba_remove_range(uhf, 0.04f);
            } else {
                hf = ba_max_clamp(hf, 28.4691806922f);
                uhf = t[k] - hf;
                uhf = ba_max_clamp(uhf, 5.19175294647f);
                uhf = uhf * 2.69313763794f;
                hf = hf * 2.155f;
                hf = ba_amplify_range(hf, 0.132f);
            }
            ps->hf[i][k] = hf;
            ps->uhf[i][k] = uhf;
        }
    }
    /* XybLowFreqToVals */
    for (size_t k = 0; k < n; k++) {
        float x = ps->lf[0][k], y = ps->lf[1][k], b = ps->lf[2][k];
        float bb = fmaf(-0.362267051518f, y, b);
        ps->lf[2][k] = bb * 49.87984651440f;
        ps->lf[0][k] = x * 33.832837186260f;
        ps->lf[1][k] = y * 14.458268100570f;
    }
    free(t);
}

/* the 16 line patterns; offsets (dy,dx); counts 9/7/8 for HF, 5 for LF */
static const int8_t MALTA_HF[16][9][2] = {
    {{0,-4},{0,-3},{0,-2},{0,-1},{0,0},{0,1},{0,2},{0,3},{0,4}},
    {{-4,0},{-3,0},{-2,0},{-1,0},{0,0},{1,0},{2,0},{3,0},{4,0}},
    {{-3,-3},{-2,-2},{-1,-1},{0,0},{1,1},{2,2},{3,3},{0,0},{0,0}},
    {{-3,3},{-2,2},{-1,1},{0,0},{1,-1},{2,-2},{3,-3},{0,0},{0,0}},
    {{-4,1},{-3,1},{-2,1},{-1,0},{0,0},{1,0},{2,-1},{3,-1},{4,-1}},
    {{-4,-1},{-3,-1},{-2,-1},{-1,0},{0,0},{1,0},{2,1},{3,1},{4,1}},
    {{-1,-4},{-1,-3},{-1,-2},{0,-1},{0,0},{0,1},{1,2},{1,3},{1,4}},
    {{1,-4},{1,-3},{1,-2},{0,-1},{0,0},{0,1},{-1,2},{-1,3},{-1,4}},
    {{-3,-2},{-2,-1},{-1,-1},{0,0},{1,1},{2,1},{3,2},{0,0},{0,0}},
    {{-3,2},{-2,1},{-1,1},{0,0},{1,-1},{2,-1},{3,-2},{0,0},{0,0}},
    {{-2,-3},{-1,-2},{-1,-1},{0,0},{1,1},{1,2},{2,3},{0,0},{0,0}},
    {{-2,3},{-1,2},{-1,1},{0,0},{1,-1},{1,-2},{2,-3},{0,0},{0,0}},
    {{2,-4},{2,-3},{1,-2},{1,-1},{0,0},{0,1},{-1,2},{-1,3},{0,0}},
    {{-2,-4},{-2,-3},{-1,-2},{-1,-1},{0,0},{0,1},{1,2},{1,3},{0,0}},
    {{-4,-2},{-3,-2},{-2,-1},{-1,-1},{0,0},{1,0},{2,1},{3,1},{0,0}},
    {{-4,2},{-3,2},{-2,1},{-1,1},{0,0},{1,0},{2,-1},{3,-1},{0,0}}};
static const int8_t MALTA_HF_N[16] = {9, 9, 7, 7, 9, 9, 9, 9, 7, 7, 7, 7, 8, 8, 8, 8};
static const int8_t MALTA_LF[16][5][2] = {
    {{0,-4},{0,-2},{0,0},{0,2},{0,4}},
    {{-4,0},{-2,0},{0,0},{2,0},{4,0}},
    {{-3,-3},{-2,-2},{0,0},{2,2},{3,3}},
    {{-3,3},{-2,2},{0,0},{2,-2},{3,-3}},
    {{-4,1},{-2,1},{0,0},{2,-1},{4,-1}},
    {{-4,-1},{-2,-1},{0,0},{2,1},{4,1}},
    {{-1,-4},{-1,-2},{0,0},{1,2},{1,4}},
    {{1,-4},{1,-2},{0,0},{-1,2},{-1,4}},
    {{-3,-2},{-2,-1},{0,0},{2,1},{3,2}},
    {{-3,2},{-2,1},{0,0},{2,-1},{3,-2}},
    {{-2,-3},{-1,-2},{0,0},{1,2},{2,3}},
    {{-2,3},{-1,2},{0,0},{1,-2},{2,-3}},
    {{2,-4},{1,-2},{0,0},{-1,2},{-2,4}},
    {{-2,-4},{-1,-2},{0,0},{1,2},{2,4}},
    {{-4,-2},{-2,-1},{0,0},{2,1},{4,2}},
    {{-4,2},{-2,1},{0,0},{2,-1},{4,-2}}};

CEO_API void ceo_malta_patterns(int8_t* hf /*16*9*2*/, int8_t* hf_n /*16*/, int8_t* lf /*16*5*2*/) {
    memcpy(hf, MALTA_HF, sizeof(MALTA_HF));
    memcpy(hf_n, MALTA_HF_N, sizeof(MALTA_HF_N));
    memcpy(lf, MALTA_LF, sizeof(MALTA_LF));
}

/* MaltaDiffMap / MaltaDiffMapLF: ac += sum over 16 patterns of (line sum of d)^2 */
static void ba_malta(const float* l0, const float* l1, size_t w, size_t h, double w_0gt1, double w_0lt1, double norm1,
                     int lf_patterns, float* diffs, float* ac) {
    const double len = 3.75;
    const double mulli = lf_patterns ? 0.611612573796 : 0.39905817637;
    const float kWeight0 = 0.5f, kWeight1 = 0.33f;
    const double w_pre0gt1 = mulli * sqrt((double)kWeight0 * w_0gt1) / (len * 2 + 1);
    const double w_pre0lt1 = mulli * sqrt((double)kWeight1 * w_0lt1) / (len * 2 + 1);
    const float norm2_0gt1 = (float)(w_pre0gt1 * norm1);
    const float norm2_0lt1 = (float)(w_pre0lt1 * norm1);
    const float norm1f = (float)norm1;
    size_t n = w * h;
    for (size_t i = 0; i < n; i++) {
        float v0 = l0[i], v1 = l1[i];
        float absval = 0.5f * (fabsf(v0) + fabsf(v1));
        float diff = v0 - v1;
        float scaler = norm2_0gt1 / (norm1f + absval);
        float d = scaler * diff;
        float scaler2 = norm2_0lt1 / (norm1f + absval);
        float fabs0 = fabsf(v0);
        float too_small = 0.55f * fabs0;
        float too_big = 1.05f * fabs0;
        if (v0 < 0.0f) {
            if (v1 > -too_small) d -= scaler2 * (v1 + too_small);
            else if (v1 < -too_big) d += scaler2 * (-v1 - too_big);
        } else {
            if (v1 < too_small) d += scaler2 * (too_small - v1);
            else if (v1 > too_big) d -= scaler2 * (v1 - too_big);
        }
        diffs[i] = d;
    }
    for (ptrdiff_t y = 0; y < (ptrdiff_t)h; y++)
        for (ptrdiff_t x = 0; x < (ptrdiff_t)w; x++) {
            float acc = 0.0f;
            for (int p = 0; p < 16; p++) {
                int cnt = lf_patterns ? 5 : MALTA_HF_N[p];
                float s = 0.0f;
                for (int t = 0; t < cnt; t++) {
                    ptrdiff_t yy = y + (lf_patterns ? MALTA_LF[p][t][0] : MALTA_HF[p][t][0]);
                    ptrdiff_t xx = x + (lf_patterns ? MALTA_LF[p][t][1] : MALTA_HF[p][t][1]);
                    float v = (yy >= 0 && yy < (ptrdiff_t)h && xx >= 0 && xx < (ptrdiff_t)w) ? diffs[(size_t)yy * w + (size_t)xx] : 0.0f;
                    s += v;
                }
                acc = fmaf(s, s, acc);
            }
            ac[(size_t)y * w + (size_t)x] += acc;
        }
}

static void ba_l2_diff_asym(const float* i0, const float* i1, size_t n, float w_0gt1, float w_0lt1, float* ac) {
    float vw0 = w_0gt1 * 0.8f, vw1 = w_0lt1 * 0.8f;
    for (size_t i = 0; i < n; i++) {
        float v0 = i0[i], v1 = i1[i];
        float diff = v0 - v1;
        float total = fmaf(diff * diff, vw0, ac[i]);
        float fabs0 = fabsf(v0);
        float too_small = 0.4f * fabs0, too_big = fabs0;
        float if_neg = v1 > -too_small ? v1 + too_small : (v1 < -too_big ? -v1 - too_big : 0.0f);
        float if_pos = v1 < too_small ? too_small - v1 : (v1 > too_big ? v1 - too_big : 0.0f);
        float v = v0 < 0.0f ? if_neg : if_pos;
        total = fmaf(vw1, v * v, total);
        ac[i] = total;
    }
}

static inline void ba_store_min3(float v, float* m0, float* m1, float* m2) {
    if (v < *m2) {
        if (v < *m0) { *m2 = *m1; *m1 = *m0; *m0 = v; }
        else if (v < *m1) { *m2 = *m1; *m1 = v; }
        else *m2 = v;
    }
}
static void ba_fuzzy_erosion(const float* from, size_t w, size_t h, float* to) {
    const ptrdiff_t S = 3, W = (ptrdiff_t)w, H = (ptrdiff_t)h;
    for (ptrdiff_t y = 0; y < H; y++)
        for (ptrdiff_t x = 0; x < W; x++) {
            float m0 = from[y * W + x], m1 = 2.0f * m0, m2 = m1;
            if (x >= S) {
                ba_store_min3(from[y * W + x - S], &m0, &m1, &m2);
                if (y >= S) ba_store_min3(from[(y - S) * W + x - S], &m0, &m1, &m2);
                if (y < H - S) ba_store_min3(from[(y + S) * W + x - S], &m0, &m1, &m2);
            }
            if (x < W - S) {
                ba_store_min3(from[y * W + x + S], &m0, &m1, &m2);
                if (y >= S) ba_store_min3(from[(y - S) * W + x + S], &m0, &m1, &m2);
                if (y < H - S) ba_store_min3(from[(y + S) * W + x + S], &m0, &m1, &m2);
            }
            if (y >= S) ba_store_min3(from[(y - S) * W + x], &m0, &m1, &m2);
            if (y < H - S) ba_store_min3(from[(y + S) * W + x], &m0, &m1, &m2);
            to[y * W + x] = (0.45f * m0 + 0.3f * m1) + 0.25f * m2;
        }
}

static inline float ba_mask_y(float delta) {
    const float offset = 0.829591754942f, scaler = 0.451936922203f, mul = 2.5485944793f;
    const float gs = (float)(1.0 / 17.83);
    float c = mul / (scaler * delta + offset);
    float r = gs * (1.0f + c);
    return r * r;
}
static inline float ba_mask_dc_y(float delta) {
    const float offset = 0.20025578522f, scaler = 3.87449418804f, mul = 0.505054525019f;
    const float gs = (float)(1.0 / 17.83);
    float c = mul / (scaler * delta + offset);
    float r = gs * (1.0f + c);
    return r * r;
}

/* diffmap of one resolution from the two psycho images (DiffmapPsychoImage) */
static void ba_diffmap_psycho(const ba_psycho* p0, const ba_psycho* p1, float hf_asym, float xmul, float* diffmap) {
    size_t w = p0->w, h = p0->h, n = w * h;
    float* ac[3]; float* dc[3];
    for (int c = 0; c < 3; c++) { ac[c] = (float*)calloc(n, sizeof(float)); dc[c] = (float*)calloc(n, sizeof(float)); }
    float* diffs = falloc(n);
    const double sq = sqrt((double)hf_asym);
    ba_malta(p0->uhf[1], p1->uhf[1], w, h, 1.10039032555 * hf_asym, 1.10039032555 / hf_asym, 71.7800275169, 0, diffs, ac[1]);
    ba_malta(p0->uhf[0], p1->uhf[0], w, h, 173.5 * hf_asym, 173.5 / hf_asym, 5.0, 0, diffs, ac[0]);
    ba_malta(p0->hf[1], p1->hf[1], w, h, 18.7237414387 * sq, 18.7237414387 / sq, 4498534.45232, 1, diffs, ac[1]);
    ba_malta(p0->hf[0], p1->hf[0], w, h, 6923.99476109 * sq, 6923.99476109 / sq, 8051.15833247, 1, diffs, ac[0]);
    ba_malta(p0->mf[1], p1->mf[1], w, h, 37.0819870399, 37.0819870399, 130262059.556, 1, diffs, ac[1]);
    ba_malta(p0->mf[0], p1->mf[0], w, h, 8246.75321353, 8246.75321353, 1009002.70582, 1, diffs, ac[0]);
    static const float wmul[9] = {400.0f, 1.50815703118f, 0.0f, 2150.0f, 10.6195433239f, 16.2176043152f,
                                  29.2353797994f, 0.844626970982f, 0.703646627719f};
    for (int c = 0; c < 3; c++) {
        if (c < 2) ba_l2_diff_asym(p0->hf[c], p1->hf[c], n, wmul[c] * hf_asym, wmul[c] / hf_asym, ac[c]);
        for (size_t i = 0; i < n; i++) {
            float d = p0->mf[c][i] - p1->mf[c][i];
            ac[c][i] = fmaf(d * d, wmul[3 + c], ac[c][i]);
            float e = p0->lf[c][i] - p1->lf[c][i];
            dc[c][i] = (e * e) * wmul[6 + c];
        }
    }
    /* MaskPsychoImage */
    float* m0 = falloc(n); float* m1 = falloc(n); float* b0 = falloc(n); float* b1 = falloc(n); float* mask = falloc(n);
    const float kMul = 6.19424080439f, kBias = 12.61050594197f;
    const float bias = kMul * kBias;
    const float sqrt_bias = sqrtf(bias);
    for (int im = 0; im < 2; im++) {
        const ba_psycho* p = im ? p1 : p0;
        float* m = im ? m1 : m0;
        for (size_t i = 0; i < n; i++) {
            float xd = (p->uhf[0][i] + p->hf[0][i]) * 2.5f;
            float yd = p->uhf[1][i] * 0.4f + p->hf[1][i] * 0.4f;
            float v = sqrtf(xd * xd + yd * yd);
            m[i] = sqrtf(kMul * fabsf(v) + bias) - sqrt_bias;   /* DiffPrecompute */
        }
    }
    ba_blur(m0, w, h, 2.7f, b0);
    ba_blur(m1, w, h, 2.7f, b1);
    ba_fuzzy_erosion(b0, w, h, mask);
    for (size_t i = 0; i < n; i++) {
        float d = b0[i] - b1[i];
        ac[1][i] += (10.0f * d) * d;
    }
    /* CombineChannelsToDiffmap */
    for (size_t i = 0; i < n; i++) {
        float val = mask[i];
        float maskval = ba_mask_y(val), dc_maskval = ba_mask_dc_y(val);
        float dsum = ((dc[0][i] * xmul) * dc_maskval + dc[1][i] * dc_maskval) + dc[2][i] * dc_maskval;
        float asum = ((ac[0][i] * xmul) * maskval + ac[1][i] * maskval) + ac[2][i] * maskval;
        diffmap[i] = sqrtf(dsum + asum);
    }
    for (int c = 0; c < 3; c++) { free(ac[c]); free(dc[c]); }
    free(diffs); free(m0); free(m1); free(b0); free(b1); free(mask);
}

static void ba_make_psycho(float* const lin[3], size_t w, size_t h, float intensity, ba_psycho* ps) {
    size_t n = w * h;
    float* xyb[3] = {falloc(n), falloc(n), falloc(n)};
    ba_opsin_dynamics(lin, w, h, intensity, xyb);
    ba_separate_frequencies(xyb, w, h, ps);
    for (int c = 0; c < 3; c++) free(xyb[c]);
}

static void ba_subsample2x(float* const in[3], size_t w, size_t h, float* out[3], size_t* ow, size_t* oh) {
    size_t xs = (w + 1) / 2, ys = (h + 1) / 2;
    for (int c = 0; c < 3; c++) {
        out[c] = (float*)calloc(xs * ys, sizeof(float));
        for (size_t y = 0; y < h; y++)
            for (size_t x = 0; x < w; x++) out[c][(y / 2) * xs + x / 2] += 0.25f * in[c][y * w + x];
        if (w & 1) for (size_t y = 0; y < ys; y++) out[c][y * xs + xs - 1] *= 2.0f;
        if (h & 1) for (size_t x = 0; x < xs; x++) out[c][(ys - 1) * xs + x] *= 2.0f;
    }
    *ow = xs; *oh = ys;
}

/* diffmap (w*h floats, optional) + max + 3-norm */
CEO_API int ceo_butteraugli_ex(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, float intensity,
                               double* max_out, double* pnorm_out, float* diffmap_out) {
    if (w < 8 || h < 8) return CEO_METRIC_CALCULATION;
    const float hf_asym = 1.0f, xmul = 1.0f;
    size_t n = w * h;
    float* l0[3] = {falloc(n), falloc(n), falloc(n)};
    float* l1[3] = {falloc(n), falloc(n), falloc(n)};
    rgb8_to_linear_planes(ref, n, l0[0], l0[1], l0[2]);
    rgb8_to_linear_planes(dist, n, l1[0], l1[1], l1[2]);
    float* diffmap = diffmap_out ? diffmap_out : falloc(n);
    {
        ba_psycho p0, p1;
        ba_make_psycho(l0, w, h, intensity, &p0);
        ba_make_psycho(l1, w, h, intensity, &p1);
        ba_diffmap_psycho(&p0, &p1, hf_asym, xmul, diffmap);
        ba_psycho_free(&p0); ba_psycho_free(&p1);
    }
    size_t sw = (w + 1) / 2, sh = (h + 1) / 2;
    if (sw >= 8 && sh >= 8) {
        float* s0[3]; float* s1[3];
        ba_subsample2x(l0, w, h, s0, &sw, &sh);
        ba_subsample2x(l1, w, h, s1, &sw, &sh);
        ba_psycho p0, p1;
        ba_make_psycho(s0, sw, sh, intensity, &p0);
        ba_make_psycho(s1, sw, sh, intensity, &p1);
        float* sub = falloc(sw * sh);
        ba_diffmap_psycho(&p0, &p1, hf_asym, xmul, sub);
        ba_psycho_free(&p0); ba_psycho_free(&p1);
        /* AddSupersampled2x(sub, 0.5, diffmap) */
        const float keep = (float)(1.0 - 0.3 * 0.5);
        for (size_t y = 0; y < h; y++)
            for (size_t x = 0; x < w; x++) {
                float v = diffmap[y * w + x] * keep;
                diffmap[y * w + x] = v + 0.5f * sub[(y / 2) * sw + x / 2];
            }
        free(sub);
        for (int c = 0; c < 3; c++) { free(s0[c]); free(s1[c]); }
    }
    float mx = 0.0f;
    double s3 = 0, s6 = 0, s12 = 0;
    for (size_t i = 0; i < n; i++) {
        float v = diffmap[i];
        if (v > mx) mx = v;
        double d = (double)v;
        double d3 = d * d * d;
        s3 += d3;
        double d6 = d3 * d3;
        s6 += d6;
        s12 += d6 * d6;
    }
    double opp = 1.0 / (double)n;
    double pn = pow(opp * s3, 1.0 / 3.0) + pow(opp * s6, 1.0 / 6.0) + pow(opp * s12, 1.0 / 12.0);
    if (max_out) *max_out = (double)mx;
    if (pnorm_out) *pnorm_out = pn / 3.0;
    for (int c = 0; c < 3; c++) { free(l0[c]); free(l1[c]); }
    if (!diffmap_out) free(diffmap);
    return CEO_OK;
}
CEO_API int ceo_butteraugli(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, float intensity,
                            double* max_out, double* pnorm_out) {
    return ceo_butteraugli_ex(ref, dist, w, h, intensity, max_out, pnorm_out, NULL);
}

/* debug: 10 psycho planes (lf0..2, mf0..2, hf0..1, uhf0..1) of one image at full resolution */
CEO_API void ceo_butteraugli_psycho(const uint8_t* rgb, size_t w, size_t h, float intensity, float* planes) {
    size_t n = w * h;
    float* l[3] = {falloc(n), falloc(n), falloc(n)};
    rgb8_to_linear_planes(rgb, n, l[0], l[1], l[2]);
    ba_psycho p;
    ba_make_psycho(l, w, h, intensity, &p);
    for (int c = 0; c < 3; c++) { memcpy(planes + (size_t)c * n, p.lf[c], n * 4); memcpy(planes + (size_t)(3 + c) * n, p.mf[c], n * 4); }
    for (int c = 0; c < 2; c++) { memcpy(planes + (size_t)(6 + c) * n, p.hf[c], n * 4); memcpy(planes + (size_t)(8 + c) * n, p.uhf[c], n * 4); }
    ba_psycho_free(&p);
    for (int c = 0; c < 3; c++) free(l[c]);
}
/* debug: opsin-dynamics XYB of one image */
CEO_API void ceo_butteraugli_opsin(const uint8_t* rgb, size_t w, size_t h, float intensity, float* planes) {
    size_t n = w * h;
    float* l[3] = {falloc(n), falloc(n), falloc(n)};
    rgb8_to_linear_planes(rgb, n, l[0], l[1], l[2]);
    float* xyb[3] = {planes, planes + n, planes + 2 * n};
    ba_opsin_dynamics(l, w, h, intensity, xyb);
    for (int c = 0; c < 3; c++) free(l[c]);
}

/* ------------------------------------------------------------------ */
/* batch entry (the "rayon par_iter over images" shape of               */
/* crates/codec-compare/src/full_comparison.rs:319-328) -- used as the  */
/* timed CPU baseline.  One pair per worker, metrics back to back.      */
/* ------------------------------------------------------------------ */

typedef struct {
    int32_t status;
    uint32_t valid;
    uint64_t sse;
    double dssim, ssimulacra2, butteraugli, psnr, butteraugli_pnorm3;
} ceo_result;

/* flags: bit0 dssim, bit1 ssimulacra2, bit2 butteraugli, bit3 psnr, bit4 xyb_roundtrip (MetricConfig, mod.rs:45-63) */
CEO_API int ceo_evaluate_pair(const uint8_t* ref, const uint8_t* dist, size_t w, size_t h, uint32_t flags, float intensity,
                              ceo_result* out) {
    memset(out, 0, sizeof(*out));
    size_t nb = w * h * 3;
    uint8_t* rt = NULL;
    const uint8_t* r = ref;
    if (flags & 16u) { /* src/eval/session.rs:447-456: reference only */
        rt = (uint8_t*)malloc(nb ? nb : 1);
        ceo_xyb_roundtrip(ref, w, h, rt);
        r = rt;
    }
    int st = CEO_OK;
    if (flags & 8u) { /* session.rs:458-465 */
        out->sse = ceo_sse(r, dist, nb);
        out->psnr = ceo_psnr(r, dist, w, h);
        out->valid |= 8u;
    }
    if (st == CEO_OK && (flags & 1u)) { /* session.rs:467-476 */
        st = ceo_dssim(r, dist, w, h, &out->dssim);
        if (st == CEO_OK) out->valid |= 1u;
    }
    if (st == CEO_OK && (flags & 2u)) { /* session.rs:478-485 */
        st = ceo_ssimulacra2(r, dist, w, h, &out->ssimulacra2);
        if (st == CEO_OK) out->valid |= 2u;
    }
    if (st == CEO_OK && (flags & 4u)) { /* session.rs:487-494 */
        st = ceo_butteraugli(r, dist, w, h, intensity, &out->butteraugli, &out->butteraugli_pnorm3);
        if (st == CEO_OK) out->valid |= 4u;
    }
    out->status = st;
    free(rt);
    return st;
}

/* refs/dists: n tightly packed RGB8 images of identical size */
CEO_API int ceo_evaluate_batch(const uint8_t* refs, const uint8_t* dists, size_t n, size_t w, size_t h, uint32_t flags,
                               float intensity, int threads, ceo_result* out) {
    size_t nb = w * h * 3;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++)
        ceo_evaluate_pair(refs + (size_t)i * nb, dists + (size_t)i * nb, w, h, flags, intensity, &out[i]);
    return CEO_OK;
}

CEO_API int ceo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
