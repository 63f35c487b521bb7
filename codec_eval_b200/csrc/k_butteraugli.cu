// k_butteraugli.cu -- Butteraugli (butteraugli 0.9.0 == libjxl butteraugli.cc; reference call
// site src/metrics/butteraugli.rs:72-80) for a batch of B pairs: max and libjxl 3-norm of the
// two-resolution diffmap.
//
// Image index convention: NI = 2B images, image i = which*B + b (which 0 = reference).
// Stages per resolution:
//   blur sigma 1.2 (5-tap mirror) -> opsin dynamics -> blur 7.16 (LF) -> blur 3.22 (MF)
//   -> blur 1.56 (UHF) with the range/clamp epilogues -> fused Malta (3 bands, 16 oriented line
//   sums over a 9x9 window in shared memory) + L2 diffs -> mask (DiffPrecompute, blur 2.7,
//   fuzzy erosion) -> combine -> diffmap; half-resolution diffmap supersample-added, then
//   max / sum d^3,d^6,d^12 reduced in fp64 (warp shuffles -> block partials -> fixed order).
// All separable blurs: shared-memory tiles, zero padding + per-coordinate 1/sum(in-range taps)
// tables (the renormalised borders of libjxl's ConvolveBorderColumn).
#include "ce_common.cuh"
#include "ce_internal.h"

#include <math.h>

namespace ce {

// ---------------------------------------------------------------- constants
struct BaConst {
    float w[4][33];   // slot 0: sigma 7.156 (R16), 1: 3.225 (R7), 2: 1.564 (R3), 3: 2.7 (R6)
    float w5[3];      // sigma 1.2: w0 (centre), w1, w2 pre-normalised
};
__constant__ BaConst c_ba;

static const float kSigmas[4] = {7.15593339443f, 3.22489901262f, 1.56416327805f, 2.7f};
static const int kRadii[4] = {16, 7, 3, 6};

static void ba_host_kernel(float sigma, int* radius, float* w) {
    const float m = 2.25f;
    const double scaler = -1.0 / (2.0 * (double)sigma * (double)sigma);
    int diff = (int)(m * fabsf(sigma));
    if (diff < 1) diff = 1;
    *radius = diff;
    for (int i = -diff; i <= diff; i++) w[i + diff] = (float)exp(scaler * (double)i * (double)i);
}

static BaConst g_ba_host;

void butteraugli_init(Context& c) {
    (void)c;
    BaConst k;
    memset(&k, 0, sizeof(k));
    for (int s = 0; s < 4; s++) {
        int r;
        ba_host_kernel(kSigmas[s], &r, k.w[s]);
        if (r != kRadii[s]) throw CudaError("butteraugli kernel radius mismatch");
    }
    float w5[5];
    int r5;
    ba_host_kernel(1.2f, &r5, w5);
    float sw = 0.0f;
    for (int i = 0; i < 5; i++) sw += w5[i];
    float scale = 1.0f / sw;
    k.w5[0] = w5[2] * scale;
    k.w5[1] = w5[1] * scale;
    k.w5[2] = w5[0] * scale;
    g_ba_host = k;
    CE_CUDA(cudaMemcpyToSymbol(c_ba, &k, sizeof(k), 0, cudaMemcpyHostToDevice));
}

// inv[x] = 1 / sum of in-range taps (ascending tap order, fp32) for a line of length len
static void ba_host_inv_weights(int slot, size_t len, std::vector<float>& inv) {
    const int R = kRadii[slot];
    const float* w = g_ba_host.w[slot];
    inv.resize(len);
    float full = 0.0f;
    for (int t = 0; t <= 2 * R; t++) full += w[t];
    const float inv_full = 1.0f / full;
    for (size_t x = 0; x < len; x++) {
        ptrdiff_t lo = (ptrdiff_t)x - R < 0 ? 0 : (ptrdiff_t)x - R;
        ptrdiff_t hi = (ptrdiff_t)x + R > (ptrdiff_t)len - 1 ? (ptrdiff_t)len - 1 : (ptrdiff_t)x + R;
        if (lo == (ptrdiff_t)x - R && hi == (ptrdiff_t)x + R) { inv[x] = inv_full; continue; }
        float ws = 0.0f;
        for (ptrdiff_t j = lo; j <= hi; j++) ws += w[j - (ptrdiff_t)x + R];
        inv[x] = 1.0f / ws;
    }
}

// device-resident tables, cached per (slot, len) for the life of the context
static float* ba_inv_table(Context& c, int slot, size_t len) {
    const uint64_t key = ((uint64_t)slot << 56) | (uint64_t)len;
    auto it = c.ba_inv_cache.find(key);
    if (it != c.ba_inv_cache.end()) return it->second;
    std::vector<float> inv;
    ba_host_inv_weights(slot, len, inv);
    float* d = nullptr;
    CE_CUDA(cudaMalloc(&d, len * sizeof(float)));
    CE_CUDA(cudaMemcpy(d, inv.data(), len * sizeof(float), cudaMemcpyHostToDevice));
    c.ba_inv_cache[key] = d;
    return d;
}

// ---------------------------------------------------------------- 5-tap mirror blur (sigma 1.2)
CE_DEVINL int mirror(int x, int n) {
    while (x < 0 || x >= n) { if (x < 0) x = -x - 1; else x = 2 * n - 1 - x; }
    return x;
}
// dir 0: along x, 1: along y.  planes [np][n]
__global__ void __launch_bounds__(256) k_ba_blur5(const float* __restrict__ in, int w, int h, size_t n, size_t total, int dir,
                                                   float* __restrict__ out) {
    const float w0 = c_ba.w5[0], w1 = c_ba.w5[1], w2 = c_ba.w5[2];
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t pl = t / n, i = t - pl * n;
        int y = (int)(i / w), x = (int)(i - (size_t)y * w);
        const float* p = in + pl * n;
        float c, l1, r1, l2, r2;
        if (dir == 0) {
            const float* row = p + (size_t)y * w;
            c = row[x]; l1 = row[mirror(x - 1, w)]; r1 = row[mirror(x + 1, w)]; l2 = row[mirror(x - 2, w)]; r2 = row[mirror(x + 2, w)];
        } else {
            c = p[i]; l1 = p[(size_t)mirror(y - 1, h) * w + x]; r1 = p[(size_t)mirror(y + 1, h) * w + x];
            l2 = p[(size_t)mirror(y - 2, h) * w + x]; r2 = p[(size_t)mirror(y + 2, h) * w + x];
        }
        out[t] = __fmaf_rn(l2 + r2, w2, __fmaf_rn(l1 + r1, w1, c * w0));  // libjxl Separable5 MulAdd chain
    }
}

// ---------------------------------------------------------------- generic separable blur
// horizontal: tile 128 x 8, thread = 4 consecutive outputs of one row, LDS.128 staging
template <int SLOT, int R>
__global__ void __launch_bounds__(256) k_ba_blur_h(const float* __restrict__ in, int w, int h, size_t n,
                                                    const float* __restrict__ inv, float* __restrict__ out) {
    constexpr int RUP = (R + 3) & ~3;
    constexpr int PITCH = 128 + 2 * RUP;
    __shared__ __align__(16) float s[8][PITCH];
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * 8;
    const float* p = in + (size_t)blockIdx.z * n;
    float* o = out + (size_t)blockIdx.z * n;
    for (int e = threadIdx.x; e < 8 * PITCH; e += 256) {
        int ry = e / PITCH, i = e - ry * PITCH;
        int x = x0 - RUP + i, y = y0 + ry;
        s[ry][i] = (x >= 0 && x < w && y < h) ? p[(size_t)y * w + x] : 0.0f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int y = y0 + ty;
    float v[4 + 2 * RUP];
#pragma unroll
    for (int q = 0; q < (4 + 2 * RUP) / 4; q++) {
        float4 f = *reinterpret_cast<const float4*>(&s[ty][tx * 4 + q * 4]);
        v[q * 4] = f.x; v[q * 4 + 1] = f.y; v[q * 4 + 2] = f.z; v[q * 4 + 3] = f.w;
    }
    if (y < h) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = x0 + tx * 4 + k;
            if (x < w) {
                float sum = 0.0f;
#pragma unroll
                for (int t = 0; t <= 2 * R; t++) sum = __fmaf_rn(v[k + (RUP - R) + t], c_ba.w[SLOT][t], sum);
                o[(size_t)y * w + x] = sum * inv[x];
            }
        }
    }
}

// vertical: tile 32 x 64, thread = 8 consecutive outputs of one column
template <int SLOT, int R>
__global__ void __launch_bounds__(256) k_ba_blur_v(const float* __restrict__ in, int w, int h, size_t n,
                                                    const float* __restrict__ inv, float* __restrict__ out) {
    constexpr int ROWS = 64 + 2 * R;
    __shared__ float s[ROWS][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 64;
    const float* p = in + (size_t)blockIdx.z * n;
    float* o = out + (size_t)blockIdx.z * n;
    for (int e = threadIdx.x; e < ROWS * 32; e += 256) {
        int ry = e >> 5, cx = e & 31;
        int x = x0 + cx, y = y0 - R + ry;
        s[ry][cx] = (x < w && y >= 0 && y < h) ? p[(size_t)y * w + x] : 0.0f;
    }
    __syncthreads();
    const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int x = x0 + cx;
    float v[8 + 2 * R];
#pragma unroll
    for (int q = 0; q < 8 + 2 * R; q++) v[q] = s[g * 8 + q][cx];
    if (x < w) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int y = y0 + g * 8 + k;
            if (y < h) {
                float sum = 0.0f;
#pragma unroll
                for (int t = 0; t <= 2 * R; t++) sum = __fmaf_rn(v[k + t], c_ba.w[SLOT][t], sum);
                o[(size_t)y * w + x] = sum * inv[y];
            }
        }
    }
}

struct BlurTables {
    float* inv_x[4];
    float* inv_y[4];
};

template <int SLOT, int R>
static void blur_slot(Context& c, const float* in, float* tmp, float* out, size_t np, size_t w, size_t h, const BlurTables& t) {
    const size_t n = w * h;
    for (size_t p0 = 0; p0 < np; p0 += 65535) {
        unsigned z = (unsigned)std::min<size_t>(65535, np - p0);
        dim3 gh(cdiv(w, 128), cdiv(h, 8), z);
        static const char* const hn[4] = {"k_ba_blur_h<R16>", "k_ba_blur_h<R7>", "k_ba_blur_h<R3>", "k_ba_blur_h<R6>"};
        static const char* const vn[4] = {"k_ba_blur_v<R16>", "k_ba_blur_v<R7>", "k_ba_blur_v<R3>", "k_ba_blur_v<R6>"};
        CE_LAUNCH(c, hn[SLOT], (double)z * n * 8,
                  k_ba_blur_h<SLOT, R><<<gh, 256, 0, c.stream>>>(in + p0 * n, (int)w, (int)h, n, t.inv_x[SLOT], tmp + p0 * n));
        dim3 gv(cdiv(w, 32), cdiv(h, 64), z);
        CE_LAUNCH(c, vn[SLOT], (double)z * n * 8,
                  k_ba_blur_v<SLOT, R><<<gv, 256, 0, c.stream>>>(tmp + p0 * n, (int)w, (int)h, n, t.inv_y[SLOT], out + p0 * n));
    }
    CE_CUDA(cudaGetLastError());
}
static void blur_planes(Context& c, int slot, const float* in, float* tmp, float* out, size_t np, size_t w, size_t h,
                        const BlurTables& t) {
    switch (slot) {
        case 0: blur_slot<0, 16>(c, in, tmp, out, np, w, h, t); break;
        case 1: blur_slot<1, 7>(c, in, tmp, out, np, w, h, t); break;
        case 2: blur_slot<2, 3>(c, in, tmp, out, np, w, h, t); break;
        default: blur_slot<3, 6>(c, in, tmp, out, np, w, h, t); break;
    }
}

// ---------------------------------------------------------------- pointwise stages
// opsin dynamics: lin [NI][3][n], blurred [NI][3][n] -> xyb [NI][3][n]
__global__ void __launch_bounds__(256) k_ba_opsin(const float* __restrict__ lin, const float* __restrict__ blurred, size_t n,
                                                   size_t total, float intensity, float* __restrict__ xyb) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t im = t / n, i = t - im * n;
        const float* L = lin + im * 3 * n + i;
        const float* Bl = blurred + im * 3 * n + i;
        float p0, p1, p2;
        ba_opsin_absorbance(Bl[0] * intensity, Bl[n] * intensity, Bl[2 * n] * intensity, p0, p1, p2);
        p0 = fmaxf(p0, 1e-4f); p1 = fmaxf(p1, 1e-4f); p2 = fmaxf(p2, 1e-4f);
        float s0 = fmaxf(ba_gamma(p0) / p0, 1e-4f);
        float s1 = fmaxf(ba_gamma(p1) / p1, 1e-4f);
        float s2 = fmaxf(ba_gamma(p2) / p2, 1e-4f);
        float c0, c1, c2;
        ba_opsin_absorbance(L[0] * intensity, L[n] * intensity, L[2 * n] * intensity, c0, c1, c2);
        c0 *= s0; c1 *= s1; c2 *= s2;
        c0 = fmaxf(c0, 1.7557483643287353f);
        c1 = fmaxf(c1, 1.7557483643287353f);
        c2 = fmaxf(c2, 12.226454707163354f);
        float* o = xyb + im * 3 * n + i;
        o[0] = c0 - c1; o[n] = c0 + c1; o[2 * n] = c2;
    }
}

// out = a - b elementwise
__global__ void __launch_bounds__(256) k_ba_sub(const float* __restrict__ a, const float* __restrict__ b, size_t total,
                                                 float* __restrict__ out) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
        out[t] = a[t] - b[t];
}

// t = unblurred mf [NI][3][n]; mf = blurred mf [NI][3][n] (in/out); hf [NI][2][n] out
__global__ void __launch_bounds__(256) k_ba_split_hf(const float* __restrict__ t_, float* __restrict__ mf, size_t n,
                                                      size_t total, float* __restrict__ hf) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t im = t / n, i = t - im * n;
        const float* T = t_ + im * 3 * n + i;
        float* M = mf + im * 3 * n + i;
        float bx = M[0], by = M[n];
        float hfx = T[0] - bx, hfy = T[n] - by;
        M[0] = ba_remove_range(bx, 0.29f);
        M[n] = ba_amplify_range(by, 0.1f);
        float scaler = __fmaf_rn(46.0f / __fmaf_rn(hfy, hfy, 46.0f), (float)(1.0 - 0.653020556257), 0.653020556257f);
        float* H = hf + im * 2 * n + i;
        H[0] = scaler * hfx;
        H[n] = hfy;
    }
}

// hf [NI][2][n] (orig in, final out), hfb blurred [NI][2][n], uhf out
__global__ void __launch_bounds__(256) k_ba_split_uhf(float* __restrict__ hf, const float* __restrict__ hfb, size_t n,
                                                       size_t total, float* __restrict__ uhf) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t im = t / n, i = t - im * n;
        float* H = hf + im * 2 * n + i;
        const float* Bq = hfb + im * 2 * n + i;
        float* U = uhf + im * 2 * n + i;
        {
            float h = Bq[0];
            float u = H[0] - h;
            H[0] = ba_remove_range(h, 1.5f);
            U[0] = ba_remove_range(u, 0.04f);
        }
        {
            float h = ba_max_clamp(Bq[n], 28.4691806922f);
            float u = H[n] - h;
            u = ba_max_clamp(u, 5.19175294647f);
            u = u * 2.69313763794f;
            h = h * 2.155f;
            h = ba_amplify_range(h, 0.132f);
            H[n] = h;
            U[n] = u;
        }
    }
}

__global__ void __launch_bounds__(256) k_ba_lf_vals(float* __restrict__ lf, size_t n, size_t total) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t im = t / n, i = t - im * n;
        float* Lp = lf + im * 3 * n + i;
        float x = Lp[0], y = Lp[n], b = Lp[2 * n];
        float bb = __fmaf_rn(-0.362267051518f, y, b);
        Lp[2 * n] = bb * 49.87984651440f;
        Lp[0] = x * 33.832837186260f;
        Lp[n] = y * 14.458268100570f;
    }
}

// mask input per image: m = DiffPrecompute(sqrt(((uhf_x+hf_x)*2.5)^2 + (uhf_y*0.4+hf_y*0.4)^2))
__global__ void __launch_bounds__(256) k_ba_mask_pre(const float* __restrict__ hf, const float* __restrict__ uhf, size_t n,
                                                      size_t total, float* __restrict__ m) {
    const float kMul = 6.19424080439f, kBias = 12.61050594197f;
    const float bias = kMul * kBias;
    const float sqrt_bias = sqrtf(bias);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t im = t / n, i = t - im * n;
        const float* H = hf + im * 2 * n + i;
        const float* U = uhf + im * 2 * n + i;
        float xd = (U[0] + H[0]) * 2.5f;
        float yd = U[n] * 0.4f + H[n] * 0.4f;
        float v = sqrtf(xd * xd + yd * yd);
        m[t] = sqrtf(kMul * fabsf(v) + bias) - sqrt_bias;
    }
}

// ---------------------------------------------------------------- Malta
#define MT_TW 32
#define MT_TH 16
#define MT_P (MT_TW + 8)
#define MT_ROWS (MT_TH + 8)
#define D(dy, dx) s[(dy) * MT_P + (dx)]

CE_DEVINL float malta_hf(const float* s) {  // s -> centre of the window, pitch MT_P
    float acc = 0.0f, t;
    t = D(0,-4) + D(0,-3) + D(0,-2) + D(0,-1) + D(0,0) + D(0,1) + D(0,2) + D(0,3) + D(0,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,0) + D(-3,0) + D(-2,0) + D(-1,0) + D(0,0) + D(1,0) + D(2,0) + D(3,0) + D(4,0);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,-3) + D(-2,-2) + D(-1,-1) + D(0,0) + D(1,1) + D(2,2) + D(3,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,3) + D(-2,2) + D(-1,1) + D(0,0) + D(1,-1) + D(2,-2) + D(3,-3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,1) + D(-3,1) + D(-2,1) + D(-1,0) + D(0,0) + D(1,0) + D(2,-1) + D(3,-1) + D(4,-1);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,-1) + D(-3,-1) + D(-2,-1) + D(-1,0) + D(0,0) + D(1,0) + D(2,1) + D(3,1) + D(4,1);
    acc = __fmaf_rn(t, t, acc);
    t = D(-1,-4) + D(-1,-3) + D(-1,-2) + D(0,-1) + D(0,0) + D(0,1) + D(1,2) + D(1,3) + D(1,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(1,-4) + D(1,-3) + D(1,-2) + D(0,-1) + D(0,0) + D(0,1) + D(-1,2) + D(-1,3) + D(-1,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,-2) + D(-2,-1) + D(-1,-1) + D(0,0) + D(1,1) + D(2,1) + D(3,2);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,2) + D(-2,1) + D(-1,1) + D(0,0) + D(1,-1) + D(2,-1) + D(3,-2);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,-3) + D(-1,-2) + D(-1,-1) + D(0,0) + D(1,1) + D(1,2) + D(2,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,3) + D(-1,2) + D(-1,1) + D(0,0) + D(1,-1) + D(1,-2) + D(2,-3);
    acc = __fmaf_rn(t, t, acc);
    t = D(2,-4) + D(2,-3) + D(1,-2) + D(1,-1) + D(0,0) + D(0,1) + D(-1,2) + D(-1,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,-4) + D(-2,-3) + D(-1,-2) + D(-1,-1) + D(0,0) + D(0,1) + D(1,2) + D(1,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,-2) + D(-3,-2) + D(-2,-1) + D(-1,-1) + D(0,0) + D(1,0) + D(2,1) + D(3,1);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,2) + D(-3,2) + D(-2,1) + D(-1,1) + D(0,0) + D(1,0) + D(2,-1) + D(3,-1);
    acc = __fmaf_rn(t, t, acc);
    return acc;
}
CE_DEVINL float malta_lf(const float* s) {  // s -> centre of the window, pitch MT_P
    float acc = 0.0f, t;
    t = D(0,-4) + D(0,-2) + D(0,0) + D(0,2) + D(0,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,0) + D(-2,0) + D(0,0) + D(2,0) + D(4,0);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,-3) + D(-2,-2) + D(0,0) + D(2,2) + D(3,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,3) + D(-2,2) + D(0,0) + D(2,-2) + D(3,-3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,1) + D(-2,1) + D(0,0) + D(2,-1) + D(4,-1);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,-1) + D(-2,-1) + D(0,0) + D(2,1) + D(4,1);
    acc = __fmaf_rn(t, t, acc);
    t = D(-1,-4) + D(-1,-2) + D(0,0) + D(1,2) + D(1,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(1,-4) + D(1,-2) + D(0,0) + D(-1,2) + D(-1,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,-2) + D(-2,-1) + D(0,0) + D(2,1) + D(3,2);
    acc = __fmaf_rn(t, t, acc);
    t = D(-3,2) + D(-2,1) + D(0,0) + D(2,-1) + D(3,-2);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,-3) + D(-1,-2) + D(0,0) + D(1,2) + D(2,3);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,3) + D(-1,2) + D(0,0) + D(1,-2) + D(2,-3);
    acc = __fmaf_rn(t, t, acc);
    t = D(2,-4) + D(1,-2) + D(0,0) + D(-1,2) + D(-2,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-2,-4) + D(-1,-2) + D(0,0) + D(1,2) + D(2,4);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,-2) + D(-2,-1) + D(0,0) + D(2,1) + D(4,2);
    acc = __fmaf_rn(t, t, acc);
    t = D(-4,2) + D(-2,1) + D(0,0) + D(2,-1) + D(4,-2);
    acc = __fmaf_rn(t, t, acc);
    return acc;
}
#undef D

struct MaltaBand {
    float norm2_0gt1, norm2_0lt1, norm1;
};
struct MaltaParams {
    MaltaBand band[3];   // uhf (HF patterns), hf (LF patterns), mf (LF patterns)
    float l2_hf_gt, l2_hf_lt;   // L2DiffAsymmetric weights (already * 0.8)
    float l2_mf;                // wmul[3+c]
};

CE_DEVINL float malta_diff(float v0, float v1, const MaltaBand& p) {
    float absval = 0.5f * (fabsf(v0) + fabsf(v1));
    float diff = v0 - v1;
    float scaler = p.norm2_0gt1 / (p.norm1 + absval);
    float d = scaler * diff;
    float scaler2 = p.norm2_0lt1 / (p.norm1 + absval);
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
    return d;
}

// grid (tiles_x, tiles_y, B).  Planes of channel C for image 0 / image 1 of each pair:
// uhf,hf: [NI][2][n]; mf: [NI][3][n].  ac out: [B][2][n] plane C.
template <int C>
__global__ void __launch_bounds__(256) k_ba_malta(const float* __restrict__ uhf, const float* __restrict__ hf,
                                                   const float* __restrict__ mf, int w, int h, size_t n, size_t B,
                                                   MaltaParams prm, float* __restrict__ ac) {
    __shared__ float s_d[MT_ROWS * MT_P];
    const size_t b = blockIdx.z;
    const int tx0 = blockIdx.x * MT_TW, ty0 = blockIdx.y * MT_TH;
    const int ox = threadIdx.x & 31, oy0 = threadIdx.x >> 5;
    const float* band0[3] = {uhf + (b * 2 + C) * n, hf + (b * 2 + C) * n, mf + (b * 3 + C) * n};
    const float* band1[3] = {uhf + ((B + b) * 2 + C) * n, hf + ((B + b) * 2 + C) * n, mf + ((B + b) * 3 + C) * n};
    float acc[2] = {0.0f, 0.0f};
#pragma unroll
    for (int bd = 0; bd < 3; bd++) {
        __syncthreads();
        for (int e = threadIdx.x; e < MT_ROWS * MT_P; e += 256) {
            int ry = e / MT_P, rx = e - ry * MT_P;
            int x = tx0 - 4 + rx, y = ty0 - 4 + ry;
            float d = 0.0f;
            if (x >= 0 && x < w && y >= 0 && y < h) {
                size_t idx = (size_t)y * w + x;
                d = malta_diff(band0[bd][idx], band1[bd][idx], prm.band[bd]);
            }
            s_d[e] = d;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float* ctr = s_d + (oy0 + k * 8 + 4) * MT_P + ox + 4;
            float m = (bd == 0) ? malta_hf(ctr) : malta_lf(ctr);
            acc[k] += m;
        }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int x = tx0 + ox, y = ty0 + oy0 + k * 8;
        if (x < w && y < h) {
            size_t idx = (size_t)y * w + x;
            float total = acc[k];
            {   // L2DiffAsymmetric on hf
                float v0 = band0[1][idx], v1 = band1[1][idx];
                float diff = v0 - v1;
                total = __fmaf_rn(diff * diff, prm.l2_hf_gt, total);
                float fabs0 = fabsf(v0);
                float too_small = 0.4f * fabs0, too_big = fabs0;
                float if_neg = v1 > -too_small ? v1 + too_small : (v1 < -too_big ? -v1 - too_big : 0.0f);
                float if_pos = v1 < too_small ? too_small - v1 : (v1 > too_big ? v1 - too_big : 0.0f);
                float v = v0 < 0.0f ? if_neg : if_pos;
                total = __fmaf_rn(prm.l2_hf_lt, v * v, total);
            }
            {   // L2Diff on mf
                float dm = band0[2][idx] - band1[2][idx];
                total = __fmaf_rn(dm * dm, prm.l2_mf, total);
            }
            ac[(b * 2 + C) * n + idx] = total;
        }
    }
}

// ---------------------------------------------------------------- combine
CE_DEVINL void store_min3(float v, float& m0, float& m1, float& m2) {
    if (v < m2) {
        if (v < m0) { m2 = m1; m1 = m0; m0 = v; }
        else if (v < m1) { m2 = m1; m1 = v; }
        else m2 = v;
    }
}

// bl: [NI][n] blurred mask inputs; ac: [B][2][n]; mf, lf: [NI][3][n]; diffmap out [B][n]
__global__ void __launch_bounds__(256) k_ba_combine(const float* __restrict__ bl, const float* __restrict__ ac,
                                                     const float* __restrict__ mf, const float* __restrict__ lf, int w, int h,
                                                     size_t n, size_t B, float xmul, float* __restrict__ diffmap) {
    const size_t total = B * n;
    const int S = 3;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t b = t / n, i = t - b * n;
        int y = (int)(i / w), x = (int)(i - (size_t)y * w);
        const float* from = bl + b * n;
        float m0 = from[i], m1 = 2.0f * m0, m2 = m1;
        if (x >= S) {
            store_min3(from[i - S], m0, m1, m2);
            if (y >= S) store_min3(from[i - (size_t)S * w - S], m0, m1, m2);
            if (y < h - S) store_min3(from[i + (size_t)S * w - S], m0, m1, m2);
        }
        if (x < w - S) {
            store_min3(from[i + S], m0, m1, m2);
            if (y >= S) store_min3(from[i - (size_t)S * w + S], m0, m1, m2);
            if (y < h - S) store_min3(from[i + (size_t)S * w + S], m0, m1, m2);
        }
        if (y >= S) store_min3(from[i - (size_t)S * w], m0, m1, m2);
        if (y < h - S) store_min3(from[i + (size_t)S * w], m0, m1, m2);
        float mask = (0.45f * m0 + 0.3f * m1) + 0.25f * m2;

        float dmk = from[i] - bl[(B + b) * n + i];
        float ac0 = ac[(b * 2 + 0) * n + i];
        float ac1 = ac[(b * 2 + 1) * n + i];
        ac1 += (10.0f * dmk) * dmk;
        const float* M0 = mf + b * 3 * n + i;
        const float* M1 = mf + (B + b) * 3 * n + i;
        const float* L0 = lf + b * 3 * n + i;
        const float* L1 = lf + (B + b) * 3 * n + i;
        float d2 = M0[2 * n] - M1[2 * n];
        float ac2 = (d2 * d2) * 16.2176043152f;
        float e0 = L0[0] - L1[0], e1 = L0[n] - L1[n], e2 = L0[2 * n] - L1[2 * n];
        float dc0 = (e0 * e0) * 29.2353797994f;
        float dc1 = (e1 * e1) * 0.844626970982f;
        float dc2 = (e2 * e2) * 0.703646627719f;
        float maskval = ba_mask_y(mask), dc_maskval = ba_mask_dc_y(mask);
        float dsum = ((dc0 * xmul) * dc_maskval + dc1 * dc_maskval) + dc2 * dc_maskval;
        float asum = ((ac0 * xmul) * maskval + ac1 * maskval) + ac2 * maskval;
        diffmap[t] = sqrtf(dsum + asum);
    }
}

// ---------------------------------------------------------------- multi-resolution
// SubSample2x on linear planes: [np][n] -> [np][on]; ((a+b)+c)+d)*0.25 with the x2 fix-ups
__global__ void __launch_bounds__(256) k_ba_subsample(const float* __restrict__ in, int w, int h, size_t n, int ow, int oh,
                                                       size_t on, size_t total, float* __restrict__ out) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t pl = t / on, i = t - pl * on;
        int oy = (int)(i / ow), ox = (int)(i - (size_t)oy * ow);
        const float* p = in + pl * n;
        int x0 = 2 * ox, y0 = 2 * oy;
        bool vx = x0 + 1 < w, vy = y0 + 1 < h;
        float s = 0.0f;
        s += 0.25f * p[(size_t)y0 * w + x0];
        if (vx) s += 0.25f * p[(size_t)y0 * w + x0 + 1];
        if (vy) {
            s += 0.25f * p[(size_t)(y0 + 1) * w + x0];
            if (vx) s += 0.25f * p[(size_t)(y0 + 1) * w + x0 + 1];
        }
        if ((w & 1) && ox == ow - 1) s *= 2.0f;
        if ((h & 1) && oy == oh - 1) s *= 2.0f;
        out[t] = s;
    }
}

#define BA_RED_BLOCKS 128
// diffmap [B][n] (+ optional sub [B][sn] supersample-added, result written back) -> block partials
// partial: [B][BA_RED_BLOCKS][4] = max, sum d^3, sum d^6, sum d^12
__global__ void __launch_bounds__(256) k_ba_finish(float* __restrict__ diffmap, const float* __restrict__ sub, int w, int sw,
                                                    size_t n, size_t sn, double* __restrict__ partial) {
    __shared__ double scratch[4 * 32];
    const size_t b = blockIdx.y;
    float* dm = diffmap + b * n;
    const float* sb = sub ? sub + b * sn : nullptr;
    const float keep = (float)(1.0 - 0.3 * 0.5);
    float mx = 0.0f;
    double s3 = 0, s6 = 0, s12 = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = dm[i];
        if (sb) {
            int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            v = v * keep;
            v = v + 0.5f * sb[(size_t)(y >> 1) * sw + (x >> 1)];
            dm[i] = v;
        }
        mx = fmaxf(mx, v);
        double d = (double)v;
        double d3 = d * d * d;
        s3 += d3;
        double d6 = d3 * d3;
        s6 += d6;
        s12 += d6 * d6;
    }
    mx = warp_max(mx);
    double v3[3] = {s3, s6, s12};
    block_sum<3>(v3, scratch);
    __shared__ float s_mx[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_mx[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.0f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) m = fmaxf(m, s_mx[i]);
        double* o = partial + (b * gridDim.x + blockIdx.x) * 4;
        o[0] = (double)m; o[1] = v3[0]; o[2] = v3[1]; o[3] = v3[2];
    }
}
__global__ void k_ba_finish_reduce(const double* __restrict__ partial, int nblk, size_t B, double* __restrict__ out) {
    size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m = 0, s3 = 0, s6 = 0, s12 = 0;
    for (int i = 0; i < nblk; i++) {
        const double* p = partial + (b * nblk + i) * 4;
        m = fmax(m, p[0]); s3 += p[1]; s6 += p[2]; s12 += p[3];
    }
    out[b * 4 + 0] = m; out[b * 4 + 1] = s3; out[b * 4 + 2] = s6; out[b * 4 + 3] = s12;
}

// ---------------------------------------------------------------- host orchestration
static MaltaParams make_malta_params(int c, float hf_asym) {
    const double len = 3.75;
    const double mulli_hf = 0.39905817637, mulli_lf = 0.611612573796;
    const float kWeight0 = 0.5f, kWeight1 = 0.33f;
    const double sq = sqrt((double)hf_asym);
    // {w_0gt1, w_0lt1, norm1} per band, per channel (c = 0: X, 1: Y)
    const double W[2][3][3] = {
        {{173.5 * hf_asym, 173.5 / hf_asym, 5.0},
         {6923.99476109 * sq, 6923.99476109 / sq, 8051.15833247},
         {8246.75321353, 8246.75321353, 1009002.70582}},
        {{1.10039032555 * hf_asym, 1.10039032555 / hf_asym, 71.7800275169},
         {18.7237414387 * sq, 18.7237414387 / sq, 4498534.45232},
         {37.0819870399, 37.0819870399, 130262059.556}}};
    MaltaParams p;
    for (int bd = 0; bd < 3; bd++) {
        const double mulli = bd == 0 ? mulli_hf : mulli_lf;
        const double w_pre0gt1 = mulli * sqrt((double)kWeight0 * W[c][bd][0]) / (len * 2 + 1);
        const double w_pre0lt1 = mulli * sqrt((double)kWeight1 * W[c][bd][1]) / (len * 2 + 1);
        p.band[bd].norm2_0gt1 = (float)(w_pre0gt1 * W[c][bd][2]);
        p.band[bd].norm2_0lt1 = (float)(w_pre0lt1 * W[c][bd][2]);
        p.band[bd].norm1 = (float)W[c][bd][2];
    }
    static const float wmul[9] = {400.0f, 1.50815703118f, 0.0f, 2150.0f, 10.6195433239f, 16.2176043152f,
                                  29.2353797994f, 0.844626970982f, 0.703646627719f};
    p.l2_hf_gt = (wmul[c] * hf_asym) * 0.8f;
    p.l2_hf_lt = (wmul[c] / hf_asym) * 0.8f;
    p.l2_mf = wmul[3 + c];
    return p;
}

struct BaLevelBufs {
    float *tmpA, *tmpB, *xyb, *lf, *mf, *hf, *uhf, *ac, *m, *bl;
    BlurTables tables;
};

static size_t ba_level_floats_per_pair(size_t n) {
    // per image: tmpA 3, tmpB 3, xyb 3, lf 3, mf 3, hf 2, uhf 2 = 19 ; x2 images = 38; pair: ac 2, m 2, bl 2 = 6
    return 44 * n;
}

static unsigned ew_blocks(Context& c, size_t total) {
    return (unsigned)std::min<size_t>(cdiv(total, 256), (size_t)c.sm_count * 32);
}

static void ba_alloc_level(Context& c, size_t B, size_t w, size_t h, BaLevelBufs& L) {
    const size_t n = w * h, NI = 2 * B;
    L.tmpA = c.arena.alloc<float>(NI * 3 * n);
    L.tmpB = c.arena.alloc<float>(NI * 3 * n);
    L.xyb = c.arena.alloc<float>(NI * 3 * n);
    L.lf = c.arena.alloc<float>(NI * 3 * n);
    L.mf = c.arena.alloc<float>(NI * 3 * n);
    L.hf = c.arena.alloc<float>(NI * 2 * n);
    L.uhf = c.arena.alloc<float>(NI * 2 * n);
    L.ac = c.arena.alloc<float>(B * 2 * n);
    L.m = c.arena.alloc<float>(NI * n);
    L.bl = c.arena.alloc<float>(NI * n);
    for (int s = 0; s < 4; s++) {
        L.tables.inv_x[s] = ba_inv_table(c, s, w);
        L.tables.inv_y[s] = ba_inv_table(c, s, h);
    }
}

// lin: [NI][3][n] -> psycho planes in L (lf, mf, hf, uhf)
static void ba_psycho_level(Context& c, const float* lin, size_t NI, size_t w, size_t h, float intensity, BaLevelBufs& L,
                            float* dbg_opsin) {
    const size_t n = w * h;
    size_t t3 = NI * 3 * n, t1 = NI * n;
    CE_LAUNCH(c, "k_ba_blur5", (double)t3 * 8, k_ba_blur5<<<ew_blocks(c, t3), 256, 0, c.stream>>>(lin, (int)w, (int)h, n, t3, 0, L.tmpA));
    CE_LAUNCH(c, "k_ba_blur5", (double)t3 * 8, k_ba_blur5<<<ew_blocks(c, t3), 256, 0, c.stream>>>(L.tmpA, (int)w, (int)h, n, t3, 1, L.tmpB));
    CE_LAUNCH(c, "k_ba_opsin", (double)t1 * 36, k_ba_opsin<<<ew_blocks(c, t1), 256, 0, c.stream>>>(lin, L.tmpB, n, t1, intensity, L.xyb));
    if (dbg_opsin) CE_CUDA(cudaMemcpyAsync(dbg_opsin, L.xyb, 3 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    blur_planes(c, 0, L.xyb, L.tmpA, L.lf, NI * 3, w, h, L.tables);
    CE_LAUNCH(c, "k_ba_sub", (double)t3 * 12, k_ba_sub<<<ew_blocks(c, t3), 256, 0, c.stream>>>(L.xyb, L.lf, t3, L.tmpB));
    blur_planes(c, 1, L.tmpB, L.tmpA, L.mf, NI * 3, w, h, L.tables);
    CE_LAUNCH(c, "k_ba_split_hf", (double)t1 * 32, k_ba_split_hf<<<ew_blocks(c, t1), 256, 0, c.stream>>>(L.tmpB, L.mf, n, t1, L.hf));
    blur_planes(c, 2, L.hf, L.tmpA, L.tmpB, NI * 2, w, h, L.tables);
    CE_LAUNCH(c, "k_ba_split_uhf", (double)t1 * 32, k_ba_split_uhf<<<ew_blocks(c, t1), 256, 0, c.stream>>>(L.hf, L.tmpB, n, t1, L.uhf));
    CE_LAUNCH(c, "k_ba_lf_vals", (double)t1 * 24, k_ba_lf_vals<<<ew_blocks(c, t1), 256, 0, c.stream>>>(L.lf, n, t1));
    CE_CUDA(cudaGetLastError());
}

// full diffmap of one resolution for B pairs; lin: [2B][3][n]
static void ba_diffmap_level(Context& c, const float* lin, size_t B, size_t w, size_t h, float intensity, float* diffmap) {
    const size_t n = w * h, NI = 2 * B;
    const float hf_asym = 1.0f, xmul = 1.0f;
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, B, w, h, L);
    ba_psycho_level(c, lin, NI, w, h, intensity, L, nullptr);
    const unsigned tx = cdiv(w, MT_TW), ty = cdiv(h, MT_TH);
    MaltaParams p0 = make_malta_params(0, hf_asym), p1 = make_malta_params(1, hf_asym);
    for (size_t b0 = 0; b0 < B; b0 += 32768) {
        if (b0 != 0) throw CudaError("butteraugli sub-batch too large");  // sub-batches are far smaller than 32768 pairs
        dim3 grid(tx, ty, (unsigned)B);
        CE_LAUNCH(c, "k_ba_malta", (double)B * n * 28,
                  k_ba_malta<1><<<grid, 256, 0, c.stream>>>(L.uhf, L.hf, L.mf, (int)w, (int)h, n, B, p1, L.ac));
        CE_LAUNCH(c, "k_ba_malta", (double)B * n * 28,
                  k_ba_malta<0><<<grid, 256, 0, c.stream>>>(L.uhf, L.hf, L.mf, (int)w, (int)h, n, B, p0, L.ac));
    }
    size_t t1 = NI * n;
    CE_LAUNCH(c, "k_ba_mask_pre", (double)t1 * 20, k_ba_mask_pre<<<ew_blocks(c, t1), 256, 0, c.stream>>>(L.hf, L.uhf, n, t1, L.m));
    blur_planes(c, 3, L.m, L.tmpA, L.bl, NI, w, h, L.tables);
    CE_LAUNCH(c, "k_ba_combine", (double)B * n * 52,
              k_ba_combine<<<ew_blocks(c, B * n), 256, 0, c.stream>>>(L.bl, L.ac, L.mf, L.lf, (int)w, (int)h, n, B, xmul, diffmap));
    CE_CUDA(cudaGetLastError());
    c.arena.release(mark);
}

size_t butteraugli_workspace_per_pair(size_t w, size_t h) {
    size_t n = w * h;
    size_t sn = ((w + 1) / 2) * ((h + 1) / 2);
    // level buffers are released between levels: max(full) dominates; + diffmap n + sub lin 6*sn + sub diffmap sn
    return (ba_level_floats_per_pair(n) + n + 7 * sn) * 4 + BA_RED_BLOCKS * 4 * 8 + 65536;
}

void butteraugli_run(Context& c, const float* lin, const float* lin2, size_t B, size_t w, size_t h, float intensity,
                     double* d_out, float* dbg_diffmap) {
    const size_t n = w * h;
    if (lin2 != lin + B * 3 * n) throw CudaError("butteraugli_run expects lin2 == lin1 + B*3*n");
    size_t mark = c.arena.mark();
    float* diffmap = c.arena.alloc<float>(B * n);
    double* partial = c.arena.alloc<double>(B * BA_RED_BLOCKS * 4);
    ba_diffmap_level(c, lin, B, w, h, intensity, diffmap);
    const size_t sw = (w + 1) / 2, sh = (h + 1) / 2, sn = sw * sh;
    float* sub = nullptr;
    if (sw >= 8 && sh >= 8) {
        float* slin = c.arena.alloc<float>(2 * B * 3 * sn);
        sub = c.arena.alloc<float>(B * sn);
        size_t total = 2 * B * 3 * sn;
        CE_LAUNCH(c, "k_ba_subsample", (double)total * 20,
                  k_ba_subsample<<<ew_blocks(c, total), 256, 0, c.stream>>>(lin, (int)w, (int)h, n, (int)sw, (int)sh, sn, total, slin));
        ba_diffmap_level(c, slin, B, sw, sh, intensity, sub);
    }
    for (size_t b0 = 0; b0 < B; b0 += 32768) {
        unsigned nb = (unsigned)std::min<size_t>(32768, B - b0);
        dim3 grid(BA_RED_BLOCKS, nb);
        CE_LAUNCH(c, "k_ba_finish", (double)nb * (sub ? 8 * n + 4 * sn : 4 * n),
                  k_ba_finish<<<grid, 256, 0, c.stream>>>(diffmap + b0 * n, sub ? sub + b0 * sn : nullptr, (int)w, (int)sw, n, sn,
                                                           partial + b0 * BA_RED_BLOCKS * 4));
    }
    CE_LAUNCH(c, "k_ba_finish_reduce", (double)B * (BA_RED_BLOCKS + 1) * 32,
              k_ba_finish_reduce<<<cdiv(B, 128), 128, 0, c.stream>>>(partial, BA_RED_BLOCKS, B, d_out));
    if (dbg_diffmap) CE_CUDA(cudaMemcpyAsync(dbg_diffmap, diffmap, n * 4, cudaMemcpyDeviceToDevice, c.stream));
    CE_CUDA(cudaGetLastError());
    c.arena.release(mark);
}

// ---- stage-level debug entries (single image) ----
void butteraugli_debug_psycho(Context& c, const float* lin, size_t w, size_t h, float intensity, float* d_planes10) {
    // run as a "batch" of one image by treating NI = 1 (B buffers sized for one pair)
    const size_t n = w * h;
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, 1, w, h, L);
    ba_psycho_level(c, lin, 1, w, h, intensity, L, nullptr);
    CE_CUDA(cudaMemcpyAsync(d_planes10, L.lf, 3 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    CE_CUDA(cudaMemcpyAsync(d_planes10 + 3 * n, L.mf, 3 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    CE_CUDA(cudaMemcpyAsync(d_planes10 + 6 * n, L.hf, 2 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    CE_CUDA(cudaMemcpyAsync(d_planes10 + 8 * n, L.uhf, 2 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    c.arena.release(mark);
}
void butteraugli_debug_opsin(Context& c, const float* lin, size_t w, size_t h, float intensity, float* d_planes3) {
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, 1, w, h, L);
    ba_psycho_level(c, lin, 1, w, h, intensity, L, d_planes3);
    c.arena.release(mark);
}
void butteraugli_debug_blur(Context& c, const float* in, size_t w, size_t h, float sigma, float* out) {
    const size_t n = w * h;
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, 1, w, h, L);
    if (fabsf(sigma - 1.2f) < 1e-6f) {
        CE_LAUNCH(c, "k_ba_blur5", (double)n * 8, k_ba_blur5<<<ew_blocks(c, n), 256, 0, c.stream>>>(in, (int)w, (int)h, n, n, 0, L.tmpA));
        CE_LAUNCH(c, "k_ba_blur5", (double)n * 8, k_ba_blur5<<<ew_blocks(c, n), 256, 0, c.stream>>>(L.tmpA, (int)w, (int)h, n, n, 1, out));
    } else {
        int slot = -1;
        for (int s = 0; s < 4; s++)
            if (fabsf(sigma - kSigmas[s]) < 1e-5f) slot = s;
        if (slot < 0) throw CudaError("unsupported sigma");
        blur_planes(c, slot, in, L.tmpA, out, 1, w, h, L.tables);
    }
    CE_CUDA(cudaGetLastError());
    c.arena.release(mark);
}

}  // namespace ce
