// k_icc.cu -- ICC matrix/TRC profile -> sRGB on the device: the step before the metric
// path for inputs that carry a non-sRGB profile (src/metrics/icc.rs:69-103 transform_to_srgb, used by
// ImageData::to_rgb8_srgb, src/eval/session.rs:143-147).
//
// The reference delegates to moxcms (un-vendored; parity UNPINNED against it).  This file implements the
// matrix-shaper transform the ICC specification defines for RGB display profiles: per-channel TRC -> 3x3 colorant
// matrix into the D50 PCS -> inverse sRGB colorant matrix -> sRGB OETF -> round to 8 bits.  LUT-only profiles
// (no colorant / TRC tags) are refused with CE_ERR_METRIC_CALCULATION, like a profile the reference cannot
// build a transform for.
//
// Everything profile-dependent is folded on the host into three 256-entry fp32 input tables, one fp32 3x3 matrix
// and one 65536-entry u8 output table (linear sRGB in 1/65535 steps -> encoded byte).  The kernel is pointwise:
// 3 table reads, 9 un-fused multiply/adds in a fixed order, 3 table reads -- HBM-bound at 6 B / pixel, and
// reproducible bit-for-bit by a float32 CPU restatement (the parity tests do exactly that).
#include <math.h>
#include <string.h>

#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

struct IccDeviceTables {   // one upload per transform: [3][256] input tables, 9 matrix entries (+ padding)
    float in[3][256];
    float m[12];
};

// thread = 4 pixels (12 bytes in, 12 bytes out); tables read through the read-only path
__global__ void __launch_bounds__(256) k_icc_to_srgb(const uint8_t* __restrict__ rgb, size_t npix,
                                                      const IccDeviceTables* __restrict__ t, const uint8_t* __restrict__ out_lut,
                                                      uint8_t* __restrict__ out) {
    __shared__ float s_in[3][256];
    for (int i = threadIdx.x; i < 768; i += 256) s_in[i >> 8][i & 255] = t->in[i >> 8][i & 255];
    __shared__ float s_m[9];
    if (threadIdx.x < 9) s_m[threadIdx.x] = t->m[threadIdx.x];
    __syncthreads();
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
        const uint8_t* s = rgb + p * 3;
        const float r = s_in[0][s[0]], g = s_in[1][s[1]], b = s_in[2][s[2]];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float v = (s_m[3 * c] * r + s_m[3 * c + 1] * g) + s_m[3 * c + 2] * b;   // un-fused (-fmad=false), fixed order
            v = fminf(fmaxf(v, 0.0f), 1.0f);
            out[p * 3 + c] = out_lut[(int)(v * 65535.0f + 0.5f)];
        }
    }
}

// ---- host: ICC parsing ------------------------------------------------------------------------------------------
namespace {

struct IccError : std::runtime_error {
    explicit IccError(const std::string& s) : std::runtime_error(s) {}
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }
double s15f16(const uint8_t* p) { return (double)(int32_t)be32(p) / 65536.0; }

struct Tag {
    const uint8_t* p = nullptr;
    size_t n = 0;
};

Tag find_tag(const uint8_t* icc, size_t len, const char* sig) {
    const uint32_t count = be32(icc + 128);
    if (132 + (size_t)count * 12 > len) throw IccError("Failed to parse ICC profile: tag table exceeds the profile");
    for (uint32_t i = 0; i < count; i++) {
        const uint8_t* e = icc + 132 + (size_t)i * 12;
        if (memcmp(e, sig, 4) == 0) {
            const size_t off = be32(e + 4), sz = be32(e + 8);
            if (off + sz > len || sz < 8) throw IccError(std::string("Failed to parse ICC profile: tag ") + sig + " out of bounds");
            Tag t;
            t.p = icc + off;
            t.n = sz;
            return t;
        }
    }
    return Tag();
}

void read_xyz(const Tag& t, const char* name, double* out3) {
    if (!t.p) throw IccError(std::string("Failed to create ICC transform: profile has no ") + name + " tag (not a matrix/TRC profile)");
    if (memcmp(t.p, "XYZ ", 4) != 0 || t.n < 20) throw IccError(std::string("Failed to parse ICC profile: bad ") + name);
    for (int i = 0; i < 3; i++) out3[i] = s15f16(t.p + 8 + 4 * i);
}

// tone reproduction curve evaluated at x in [0,1] (ICC.1:2010 10.5 curveType, 10.15 parametricCurveType)
double eval_trc(const Tag& t, const char* name, double x) {
    if (!t.p) throw IccError(std::string("Failed to create ICC transform: profile has no ") + name + " tag (not a matrix/TRC profile)");
    if (memcmp(t.p, "curv", 4) == 0) {
        if (t.n < 12) throw IccError(std::string("Failed to parse ICC profile: bad ") + name);
        const uint32_t cnt = be32(t.p + 8);
        if (12 + (size_t)cnt * 2 > t.n) throw IccError(std::string("Failed to parse ICC profile: truncated ") + name);
        if (cnt == 0) return x;
        if (cnt == 1) return pow(x, (double)be16(t.p + 12) / 256.0);
        const double pos = x * (double)(cnt - 1);
        uint32_t i0 = (uint32_t)pos;
        if (i0 >= cnt - 1) i0 = cnt - 2;
        const double f = pos - (double)i0;
        const double a = (double)be16(t.p + 12 + 2 * i0) / 65535.0, b = (double)be16(t.p + 12 + 2 * (i0 + 1)) / 65535.0;
        return a + (b - a) * f;
    }
    if (memcmp(t.p, "para", 4) == 0) {
        if (t.n < 16) throw IccError(std::string("Failed to parse ICC profile: bad ") + name);
        const uint32_t ft = be16(t.p + 8);
        static const int nparam[5] = {1, 3, 4, 5, 7};
        if (ft > 4 || 12 + (size_t)nparam[ft] * 4 > t.n) throw IccError(std::string("Failed to parse ICC profile: bad parametric ") + name);
        double q[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < nparam[ft]; i++) q[i] = s15f16(t.p + 12 + 4 * i);
        const double g = q[0], a = q[1], b = q[2], c = q[3], d = q[4], e = q[5], f = q[6];
        switch (ft) {
            case 0: return pow(x, g);
            case 1: return x >= -b / a ? pow(a * x + b, g) : 0.0;
            case 2: return x >= -b / a ? pow(a * x + b, g) + c : c;
            case 3: return x >= d ? pow(a * x + b, g) : c * x;
            default: return x >= d ? pow(a * x + b, g) + e : c * x + f;
        }
    }
    throw IccError(std::string("Failed to parse ICC profile: unsupported curve type in ") + name);
}

double srgb_oetf(double l) { return l <= 0.0031308 ? 12.92 * l : 1.055 * pow(l, 1.0 / 2.4) - 0.055; }

// 65536-entry output table, the same for every profile
const uint8_t* srgb_out_lut() {
    static uint8_t lut[65536];
    static bool ready = false;
    if (!ready) {
        for (int i = 0; i < 65536; i++) {
            double v = srgb_oetf((double)i / 65535.0) * 255.0 + 0.5;
            lut[i] = (uint8_t)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : floor(v)));
        }
        ready = true;
    }
    return lut;
}

void build_tables(const uint8_t* icc, size_t len, IccDeviceTables& T) {
    if (len < 132) throw IccError("Failed to parse ICC profile: shorter than its header");
    if (memcmp(icc + 36, "acsp", 4) != 0) throw IccError("Failed to parse ICC profile: missing 'acsp' signature");
    if (memcmp(icc + 16, "RGB ", 4) != 0) throw IccError("Failed to create ICC transform: not an RGB profile");
    if (memcmp(icc + 20, "XYZ ", 4) != 0) throw IccError("Failed to create ICC transform: profile connection space is not XYZ");
    double P[3][3];   // columns = colorants: XYZ(D50) = P * rgb_linear
    double col[3];
    const char* xyz_tags[3] = {"rXYZ", "gXYZ", "bXYZ"};
    const char* trc_tags[3] = {"rTRC", "gTRC", "bTRC"};
    for (int c = 0; c < 3; c++) {
        read_xyz(find_tag(icc, len, xyz_tags[c]), xyz_tags[c], col);
        for (int r = 0; r < 3; r++) P[r][c] = col[r];
    }
    for (int c = 0; c < 3; c++) {
        const Tag t = find_tag(icc, len, trc_tags[c]);
        for (int i = 0; i < 256; i++) T.in[c][i] = (float)eval_trc(t, trc_tags[c], (double)i / 255.0);
    }
    // sRGB colorants in the D50 PCS (the sRGB v2 / v4 profile values); XYZ -> linear sRGB is the inverse
    const double S[3][3] = {{0.4360747, 0.3850649, 0.1430804}, {0.2225045, 0.7168786, 0.0606169}, {0.0139322, 0.0971045, 0.7141733}};
    const double det = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0]) +
                       S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]);
    double Si[3][3];
    Si[0][0] = (S[1][1] * S[2][2] - S[1][2] * S[2][1]) / det;
    Si[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) / det;
    Si[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) / det;
    Si[1][0] = (S[1][2] * S[2][0] - S[1][0] * S[2][2]) / det;
    Si[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) / det;
    Si[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) / det;
    Si[2][0] = (S[1][0] * S[2][1] - S[1][1] * S[2][0]) / det;
    Si[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) / det;
    Si[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) / det;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) T.m[3 * r + c] = (float)((Si[r][0] * P[0][c] + Si[r][1] * P[1][c]) + Si[r][2] * P[2][c]);
    T.m[9] = T.m[10] = T.m[11] = 0.0f;
}

}  // namespace

// d_rgb [npix][3] -> d_out [npix][3].  Throws std::runtime_error (text for Error::MetricCalculation.reason) when the
// profile cannot be used; arena temporaries; stream-ordered.
void icc_to_srgb_run(Context& c, const uint8_t* d_rgb, size_t npix, const uint8_t* icc, size_t icc_len, uint8_t* d_out) {
    IccDeviceTables T;
    build_tables(icc, icc_len, T);
    IccDeviceTables* d_t = c.arena.alloc<IccDeviceTables>(1);
    uint8_t* d_lut = c.arena.alloc<uint8_t>(65536);
    CE_CUDA(cudaMemcpyAsync(d_t, &T, sizeof(T), cudaMemcpyHostToDevice, c.stream));
    CE_CUDA(cudaMemcpyAsync(d_lut, srgb_out_lut(), 65536, cudaMemcpyHostToDevice, c.stream));
    CE_CUDA(cudaStreamSynchronize(c.stream));   // T lives on this stack frame
    const unsigned blocks = std::min<unsigned>(cdiv(npix, 256), (unsigned)c.sm_count * 16);
    CE_LAUNCH(c, "k_icc_to_srgb", (double)npix * 6, k_icc_to_srgb<<<blocks, 256, 0, c.stream>>>(d_rgb, npix, d_t, d_lut, d_out));
    CE_CUDA(cudaGetLastError());
}

bool icc_is_usable(const uint8_t* icc, size_t icc_len, std::string* why) {
    try {
        IccDeviceTables T;
        build_tables(icc, icc_len, T);
        return true;
    } catch (const std::exception& e) {
        if (why) *why = e.what();
        return false;
    }
}

}  // namespace ce
