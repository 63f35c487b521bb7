// k_butteraugli.cu -- Butteraugli (butteraugli 0.9.0 == libjxl butteraugli.cc; reference call
// site src/metrics/butteraugli.rs:72-80) for a batch of B pairs: max and libjxl 3-norm of the
// two-resolution diffmap.
//
// Image index convention: NI = R + B images: the R distinct references of the sub-batch, then the B
// distorted images; pair b compares image ridx[b] with image R + b.  Per-image stages (everything up to
// the Malta filters) run once per image, so a reference shared by several pairs is processed once.
// Kernels per resolution (each stage's pointwise epilogue is fused into the blur that feeds it):
//   k_ba_opsin   : sigma-1.2 blur (5-tap mirror, H+V through a smem tile) + opsin dynamics      lin -> xyb
//   k_ba_blur_h  : sigma-7.16 (33 taps) along x                                               xyb -> tmp
//   k_ba_blur_v  : sigma-7.16 along y + LF epilogue (mf_pre = xyb - lf, LF scaling)            tmp -> lf, mf_pre
//   k_ba_blur2d<MF> : sigma-3.22 H+V through a smem tile + HF split epilogue                   mf_pre -> mf, hf_pre
//   k_ba_blur2d<HF> : sigma-1.56 H+V + UHF split + mask precompute                             hf_pre -> hf, uhf, m
//   k_ba_blur2d<NONE>: sigma-2.7 H+V of the mask input                                         m -> bl
//   k_ba_malta   : Malta filters, 3 bands x {X,Y}: diffs staged in smem, 9x12 register window per
//                  thread (4 pixels), 16 oriented line sums each (built from shared sub-sums),
//                  + L2 diffs                                                                   -> ac
//   k_ba_combine : fuzzy erosion, mask, DC/AC combine                                          -> diffmap
// then the half-resolution diffmap is supersample-added and max / sum d^3,d^6,d^12 reduced in
// fp64 (warp shuffles -> block partials -> fixed order).
// Blurs use zero padding + per-coordinate 1/sum(in-range taps) tables (the renormalised
// borders of libjxl's ConvolveBorderColumn).  Tiles are staged by TMA (cp.async.bulk.tensor.3d, one
// box per plane group, out-of-image elements zero-filled by the copy engine, completion on an
// mbarrier) when the row pitch is a multiple of 16 bytes; otherwise by 128-/32-bit loads.  Every
// consumer thread issues fence.proxy.async before the barrier that precedes a refill of a slot it
// has read with ordinary shared loads.
#include <cuda.h>

#include "ce_common.cuh"
#include "ce_internal.h"
#include "ba_weights.inc"

#include <math.h>

namespace ce {

// ---------------------------------------------------------------- constants
struct BaConst {
    float w[4][33];   // slot 0: sigma 7.156 (R16), 1: 3.225 (R7), 2: 1.564 (R3), 3: 2.7 (R6)
    float w5[3];      // sigma 1.2: w0 (centre), w1, w2 pre-normalised
};
__constant__ BaConst c_ba;

// blur weights as compile-time literals (FFMA immediates); SLOT as in c_ba.w
template <int SLOT>
CE_DEVINL constexpr float baw(int t) {
    return SLOT == 0 ? kBaW16[t < 33 ? t : 0] : SLOT == 1 ? kBaW7[t < 15 ? t : 0] : SLOT == 2 ? kBaW3[t < 7 ? t : 0] : kBaW6[t < 13 ? t : 0];
}

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

static void ba_set_kernel_attributes();
void butteraugli_init(Context& c) {
    (void)c;
    ba_set_kernel_attributes();
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
    // the kernels use the generated literals of ba_weights.inc; they must equal this run-time computation
    const float* lit[4] = {hBaW16, hBaW7, hBaW3, hBaW6};
    for (int s = 0; s < 4; s++)
        for (int t = 0; t <= 2 * kRadii[s]; t++)
            if (lit[s][t] != k.w[s][t]) throw CudaError("ba_weights.inc is stale: regenerate with tools/gen_ba_weights.py");
    for (int t = 0; t < 3; t++)
        if (hBaW5[t] != k.w5[t]) throw CudaError("ba_weights.inc is stale: regenerate with tools/gen_ba_weights.py");
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
    inv.resize(((len + 3) & ~size_t(3)) + 4, 0.0f);   // zero padding: kernels read 4 entries at a time
    float* d = nullptr;
    CE_CUDA(cudaMalloc(&d, inv.size() * sizeof(float)));
    CE_CUDA(cudaMemcpy(d, inv.data(), inv.size() * sizeof(float), cudaMemcpyHostToDevice));
    c.ba_inv_cache[key] = d;
    return d;
}


// ---------------------------------------------------------------- sigma-1.2 blur (+ opsin dynamics)
#define OP_TW 64
#define OP_TH 16
#define OP_P (OP_TW + 8)      // columns x0-4 .. x0+67
#define OP_ROWS (OP_TH + 4)   // rows y0-2 .. y0+17

CE_DEVINL float blur5(float l2, float l1, float c, float r1, float r2) {
    return __fmaf_rn(l2 + r2, kBaW5[2], __fmaf_rn(l1 + r1, kBaW5[1], c * kBaW5[0]));  // libjxl Separable5 MulAdd chain
}

// OPSIN = true : lin [NI][3][n] -> xyb [NI][3][n] (blur + OpsinDynamicsImage), grid.z = image
// OPSIN = false: one plane per grid.z, out = blurred plane (stage test entry)
// TMA: the NPL plane tiles arrive as one bulk tensor copy (box OP_P x OP_ROWS x NPL, zeros outside the image); tiles on
// the image border then overwrite their out-of-image entries with the mirrored in-image values (Separable5's rule).
template <bool OPSIN, bool TMA>
__global__ void __launch_bounds__(256) k_ba_opsin(const float* __restrict__ lin, int w, int h, size_t n, float intensity,
                                                   float* __restrict__ out, int vec, const __grid_constant__ CUtensorMap map) {
    constexpr int NPL = OPSIN ? 3 : 1;
    __shared__ __align__(128) float s_in[NPL][OP_ROWS * OP_P];
    __shared__ __align__(16) float s_h[NPL][OP_ROWS * OP_TW];
    __shared__ __align__(8) unsigned long long s_bar;
    const int x0 = blockIdx.x * OP_TW, y0 = blockIdx.y * OP_TH;
    const float* src = lin + (size_t)blockIdx.z * NPL * n;
    float* dst = out + (size_t)blockIdx.z * NPL * n;
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&s_bar, NPL * OP_ROWS * OP_P * 4);
            tma_load_3d(&s_in[0][0], &map, x0 - 4, y0 - 2, (int)(blockIdx.z * NPL), &s_bar);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
        const bool border = x0 - 4 < 0 || x0 - 4 + OP_P > w || y0 - 2 < 0 || y0 - 2 + OP_ROWS > h;   // block-uniform
        if (border) {
            for (int e = threadIdx.x; e < OP_ROWS * OP_P; e += 256) {
                const int r = e / OP_P, cx = e - r * OP_P;
                const int x = x0 - 4 + cx, y = y0 - 2 + r;
                if (x < 0 || x >= w || y < 0 || y >= h) {
                    const int mx = mirror(x, w), my = mirror(y, h);
                    const int tr = my - (y0 - 2), tc = mx - (x0 - 4);
                    const bool in_tile = tr >= 0 && tr < OP_ROWS && tc >= 0 && tc < OP_P;
#pragma unroll
                    for (int c = 0; c < NPL; c++)
                        s_in[c][e] = in_tile ? s_in[c][tr * OP_P + tc] : src[(size_t)c * n + (size_t)my * w + mx];
                }
            }
            __syncthreads();
        }
    } else {
#pragma unroll
        for (int c = 0; c < NPL; c++) load_tile<1, OP_P / 4, OP_ROWS, 256>(s_in[c], OP_P, src + (size_t)c * n, w, h, x0 - 4, y0 - 2, vec != 0);
        __syncthreads();
    }
    // horizontal pass: (row, 4-column group) items
    for (int e = threadIdx.x; e < OP_ROWS * (OP_TW / 4); e += 256) {
        const int r = e >> 4, g = e & 15;
#pragma unroll
        for (int c = 0; c < NPL; c++) {
            const float4* row = reinterpret_cast<const float4*>(s_in[c] + r * OP_P + 4 * g);
            const float4 a = row[0], b = row[1], d = row[2];
            const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
            float4 o;
            o.x = blur5(v[2], v[3], v[4], v[5], v[6]);
            o.y = blur5(v[3], v[4], v[5], v[6], v[7]);
            o.z = blur5(v[4], v[5], v[6], v[7], v[8]);
            o.w = blur5(v[5], v[6], v[7], v[8], v[9]);
            *reinterpret_cast<float4*>(s_h[c] + r * OP_TW + 4 * g) = o;
        }
    }
    __syncthreads();
    const int g = threadIdx.x & 15, r = threadIdx.x >> 4;
    const int x = x0 + 4 * g, y = y0 + r;
    if (y >= h || x >= w) return;
    float bl[NPL][4];
#pragma unroll
    for (int c = 0; c < NPL; c++) {
        float4 q[5];
#pragma unroll
        for (int j = 0; j < 5; j++) q[j] = *reinterpret_cast<const float4*>(s_h[c] + (r + j) * OP_TW + 4 * g);
        bl[c][0] = blur5(q[0].x, q[1].x, q[2].x, q[3].x, q[4].x);
        bl[c][1] = blur5(q[0].y, q[1].y, q[2].y, q[3].y, q[4].y);
        bl[c][2] = blur5(q[0].z, q[1].z, q[2].z, q[3].z, q[4].z);
        bl[c][3] = blur5(q[0].w, q[1].w, q[2].w, q[3].w, q[4].w);
    }
    float res[NPL][4];
    if (OPSIN) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float p0, p1, p2;
            ba_opsin_absorbance(bl[0][k] * intensity, bl[NPL > 1 ? 1 : 0][k] * intensity, bl[NPL > 2 ? 2 : 0][k] * intensity, p0, p1, p2);
            p0 = fmaxf(p0, 1e-4f); p1 = fmaxf(p1, 1e-4f); p2 = fmaxf(p2, 1e-4f);
            const float s0 = fmaxf(__fdividef(ba_gamma(p0), p0), 1e-4f);
            const float s1 = fmaxf(__fdividef(ba_gamma(p1), p1), 1e-4f);
            const float s2 = fmaxf(__fdividef(ba_gamma(p2), p2), 1e-4f);
            const int ci = (r + 2) * OP_P + 4 * g + 4 + k;
            float c0, c1, c2;
            ba_opsin_absorbance(s_in[0][ci] * intensity, s_in[NPL > 1 ? 1 : 0][ci] * intensity, s_in[NPL > 2 ? 2 : 0][ci] * intensity, c0, c1, c2);
            c0 *= s0; c1 *= s1; c2 *= s2;
            c0 = fmaxf(c0, 1.7557483643287353f);
            c1 = fmaxf(c1, 1.7557483643287353f);
            c2 = fmaxf(c2, 12.226454707163354f);
            res[0][k] = c0 - c1; res[NPL > 1 ? 1 : 0][k] = c0 + c1; res[NPL > 2 ? 2 : 0][k] = c2;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) res[0][k] = bl[0][k];
    }
    const size_t idx = (size_t)y * w + x;
#pragma unroll
    for (int c = 0; c < NPL; c++) {
        float* o = dst + (size_t)c * n + idx;
        if (vec) *reinterpret_cast<float4*>(o) = make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x + k < w) o[k] = res[c][k];
        }
    }
}

// ---------------------------------------------------------------- wide separable blur (sigma 7.16)
// horizontal: tile 128 x 8, thread = 4 consecutive outputs of one row.  A block walks down BH_NT vertically adjacent
// tiles through a ring of BH_SLOTS buffers with BH_SLOTS - 1 tiles in flight ahead of the one being computed: with one
// tile per block (round 1) or two buffers the kernel moved 3.9 TB/s and its warps spun on the tile barrier (ncu:
// 55 executed instructions per pixel against 42 in the loop body) -- too few bytes in flight per SM, not arithmetic.
// TMA: a tile is one bulk tensor copy (map = in as [planes][h][w], box (PITCH, 8, 1)); otherwise per-thread loads.
#define BH_NT 16
#define BH_SLOTS 4
template <int SLOT, int R, bool TMA>
__global__ void __launch_bounds__(256) k_ba_blur_h(const float* __restrict__ in, int w, int h, size_t n,
                                                    const float* __restrict__ inv, float* __restrict__ out, int vec,
                                                    const __grid_constant__ CUtensorMap map) {
    constexpr int RUP = (R + 3) & ~3;
    constexpr int PITCH = 128 + 2 * RUP;
    constexpr int NS = TMA ? BH_SLOTS : 1;
    __shared__ __align__(128) float s[NS][8 * PITCH];
    __shared__ __align__(8) unsigned long long s_bar[NS];
    const int x0 = blockIdx.x * 128, yb = blockIdx.y * (8 * BH_NT);
    const int nt = min(BH_NT, (h - yb + 7) / 8);
    const float* p = in + (size_t)blockIdx.z * n;
    float* o = out + (size_t)blockIdx.z * n;
    auto issue = [&](int t) {   // thread 0: tile t -> slot t % NS
        mbar_expect_tx(&s_bar[t % NS], 8 * PITCH * 4);
        tma_load_3d(s[t % NS], &map, x0 - RUP, yb + 8 * t, (int)blockIdx.z, &s_bar[t % NS]);
    };
    if (TMA) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NS; i++) mbar_init(&s_bar[i], 1);
            mbar_fence_init();
#pragma unroll
            for (int i = 0; i < NS; i++)
                if (i < nt) issue(i);
        }
        __syncthreads();   // barriers initialised before anyone polls them
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = x0 + tx * 4;
    float4 iv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < w) iv = *reinterpret_cast<const float4*>(inv + x);   // zero-padded table
    const float ivk[4] = {iv.x, iv.y, iv.z, iv.w};
    for (int t = 0; t < nt; t++) {
        const float* sb = s[t % NS];
        if (TMA) {
            mbar_wait(&s_bar[t % NS], (unsigned)(t / NS) & 1u);
        } else {
            __syncthreads();   // the previous tile has been consumed
            load_tile<0, PITCH / 4, 8, 256>(s[0], PITCH, p, w, h, x0 - RUP, yb + 8 * t, vec != 0);
            __syncthreads();
        }
        const int y = yb + 8 * t + ty;
        float res[4] = {0.f, 0.f, 0.f, 0.f};
        if (y < h && x < w) {
            float v[4 + 2 * RUP];
#pragma unroll
            for (int q = 0; q < (4 + 2 * RUP) / 4; q++) {
                float4 f = *reinterpret_cast<const float4*>(&sb[ty * PITCH + tx * 4 + q * 4]);
                v[q * 4] = f.x; v[q * 4 + 1] = f.y; v[q * 4 + 2] = f.z; v[q * 4 + 3] = f.w;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float sum = 0.0f;
#pragma unroll
                for (int tt = 0; tt <= 2 * R; tt++) sum = __fmaf_rn(v[k + (RUP - R) + tt], baw<SLOT>(tt), sum);
                res[k] = sum * ivk[k];
            }
        }
        if (TMA && t + NS < nt) {   // block-uniform: the slot is refilled once every thread is done reading it
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) issue(t + NS);
        }
        if (y >= h || x >= w) continue;
        float* d = o + ((size_t)y * w + x);
        if (vec) *reinterpret_cast<float4*>(d) = make_float4(res[0], res[1], res[2], res[3]);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x + k < w) d[k] = res[k];
        }
    }
}

// splat of a blur weight over both lanes of a packed pair
template <int SLOT>
CE_DEVINL f32x2 baw2(int t) {
    return pk2(baw<SLOT>(t), baw<SLOT>(t));
}

// vertical: tile 32 x 64, thread = 4 consecutive outputs of TWO adjacent columns held as packed fp32x2 pairs
// (LDS.64 delivers the pair; FFMA2 = the scalar fused multiply-add in each lane, half the issue slots).  The NPL
// planes of an image go through one block; plane c+1 is staged with cp.async while plane c is computed.
// EPI 0: out[pl] = blurred plane (grid.z counts plane groups of NPL)
// EPI 1 (NPL = 3): LF epilogue -- lf = blurred xyb; mf_pre = xyb - lf; lf scaled (XybLowFreqToVals)
template <int SLOT, int R, int NPL, int EPI, bool TMA>
__global__ void __launch_bounds__(256) k_ba_blur_v(const float* __restrict__ in, int w, int h, size_t n,
                                                    const float* __restrict__ inv, float* __restrict__ out,
                                                    const float* __restrict__ xyb, float* __restrict__ mf_pre,
                                                    const __grid_constant__ CUtensorMap map) {
    constexpr int ROWS = 64 + 2 * R;
    __shared__ __align__(128) float s[2][ROWS * 32];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 64;
    const int cp = threadIdx.x & 15, g = threadIdx.x >> 4;   // 16 column pairs x 16 groups of 4 rows
    const int x = x0 + 2 * cp;
    const size_t base = (size_t)blockIdx.z * NPL * n;
    const bool vec = (w & 3) == 0;
    float res[NPL][4][2];
    auto tma_issue = [&](int c) {   // thread 0: plane c -> buffer c & 1 (map = in as [planes][h][w], box (32, ROWS, 1))
        mbar_expect_tx(&s_bar[c & 1], ROWS * 32 * 4);
        tma_load_3d(s[c & 1], &map, x0, y0 - R, (int)(blockIdx.z * NPL + c), &s_bar[c & 1]);
    };
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
            tma_issue(0);
            if (NPL > 1) tma_issue(1);
        }
        __syncthreads();
    } else {
        load_tile_async<8, ROWS, 256>(s[0], 32, in + base, w, h, x0, y0 - R, vec);
        cp_async_commit();
    }
#pragma unroll
    for (int c = 0; c < NPL; c++) {
        if (TMA) {
            mbar_wait(&s_bar[c & 1], (unsigned)(c >> 1) & 1u);
        } else {
            if (c + 1 < NPL) {
                load_tile_async<8, ROWS, 256>(s[(c + 1) & 1], 32, in + base + (size_t)(c + 1) * n, w, h, x0, y0 - R, vec);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
        }
        const float* sc = s[c & 1];
        f32x2 v[4 + 2 * R];
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; q++) v[q] = *reinterpret_cast<const f32x2*>(&sc[(g * 4 + q) * 32 + 2 * cp]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int y = y0 + g * 4 + k;
            f32x2 sum = 0ULL;   // (+0, +0)
#pragma unroll
            for (int t = 0; t <= 2 * R; t++) sum = fma2(v[k + t], baw2<SLOT>(t), sum);
            const float iy = y < h ? inv[y] : 0.0f;
            float lo, hi;
            unpk2(sum, lo, hi);
            res[c][k][0] = lo * iy;
            res[c][k][1] = hi * iy;
        }
        if (TMA && c + 2 < NPL) fence_proxy_async();
        __syncthreads();   // plane c's buffer is refilled two planes later
        if (TMA && c + 2 < NPL && threadIdx.x == 0) tma_issue(c + 2);
    }
    if (x >= w) return;
    const bool two = x + 1 < w;
    if ((w & 1) == 0) {
        // even widths: the lane's two columns leave (and the xyb values arrive) as one 64-bit access, so a half warp
        // writes a whole 128-byte line per instruction instead of every other word of it
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int y = y0 + g * 4 + k;
            if (y >= h) break;
            const size_t idx = base + (size_t)y * w + x;
            if (EPI == 0) {
#pragma unroll
                for (int c = 0; c < NPL; c++) *reinterpret_cast<float2*>(out + idx + (size_t)c * n) = make_float2(res[c][k][0], res[c][k][1]);
            } else {
                const float2 xx = *reinterpret_cast<const float2*>(xyb + idx), xy = *reinterpret_cast<const float2*>(xyb + idx + n),
                             xb = *reinterpret_cast<const float2*>(xyb + idx + 2 * n);
                const float lx[2] = {res[0][k][0], res[0][k][1]}, ly[2] = {res[NPL > 1 ? 1 : 0][k][0], res[NPL > 1 ? 1 : 0][k][1]},
                            lb[2] = {res[NPL > 2 ? 2 : 0][k][0], res[NPL > 2 ? 2 : 0][k][1]};
                *reinterpret_cast<float2*>(mf_pre + idx) = make_float2(xx.x - lx[0], xx.y - lx[1]);
                *reinterpret_cast<float2*>(mf_pre + idx + n) = make_float2(xy.x - ly[0], xy.y - ly[1]);
                *reinterpret_cast<float2*>(mf_pre + idx + 2 * n) = make_float2(xb.x - lb[0], xb.y - lb[1]);
                const float bb0 = __fmaf_rn(-0.362267051518f, ly[0], lb[0]), bb1 = __fmaf_rn(-0.362267051518f, ly[1], lb[1]);
                *reinterpret_cast<float2*>(out + idx + 2 * n) = make_float2(bb0 * 49.87984651440f, bb1 * 49.87984651440f);
                *reinterpret_cast<float2*>(out + idx) = make_float2(lx[0] * 33.832837186260f, lx[1] * 33.832837186260f);
                *reinterpret_cast<float2*>(out + idx + n) = make_float2(ly[0] * 14.458268100570f, ly[1] * 14.458268100570f);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = y0 + g * 4 + k;
        if (y >= h) break;
        const size_t idx = base + (size_t)y * w + x;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (j == 1 && !two) break;
            if (EPI == 0) {
#pragma unroll
                for (int c = 0; c < NPL; c++) out[idx + j + (size_t)c * n] = res[c][k][j];
            } else {
                const float lx = res[0][k][j], ly = res[NPL > 1 ? 1 : 0][k][j], lb = res[NPL > 2 ? 2 : 0][k][j];
                mf_pre[idx + j] = xyb[idx + j] - lx;
                mf_pre[idx + j + n] = xyb[idx + j + n] - ly;
                mf_pre[idx + j + 2 * n] = xyb[idx + j + 2 * n] - lb;
                const float bb = __fmaf_rn(-0.362267051518f, ly, lb);
                out[idx + j + 2 * n] = bb * 49.87984651440f;
                out[idx + j] = lx * 33.832837186260f;
                out[idx + j + n] = ly * 14.458268100570f;
            }
        }
    }
}

// ---------------------------------------------------------------- fused 2-D blur (H then V through smem) + epilogues
// tile 64 x 32 outputs; the horizontal pass runs on the 32 + 2R rows the vertical pass needs.
// EPI 0 (NPL 1)  : out = blurred plane
// EPI 2 (NPL 3)  : MF -- in = mf_pre [img][3][n]; out_a = mf (range ops on X,Y; B as blurred) [img][3][n];
//                   out_b = hf_pre [img][2][n] (X suppressed by Y)
// EPI 3 (NPL 2)  : HF -- in = hf_pre [img][2][n]; out_a = hf [img][2][n]; out_b = uhf [img][2][n];
//                   out_c = mask input m [img][n] = DiffPrecompute(combine(hf, uhf))
#define B2_TW 64
#define B2_TH 32
#define B2_NT 4      // a block walks down B2_NT vertically adjacent tiles: plane c of tile t is work item t * NPL + c
#ifndef B2_MINB3
#define B2_MINB3 3   // minimum blocks / SM asked of ptxas: three-plane (MF) instance; 4 for the others
#endif
// Items go through two input buffers: item i + 2 is requested as soon as the horizontal pass of item i has left its
// buffer, so only the block's very first load is exposed (with one tile per block the first plane of EVERY tile was:
// ncu had 24 % of the HF kernel's stall samples in the wait for it, and all of the one-plane mask blur's loads).
template <int SLOT, int R, int NPL, int EPI, bool TMA>
__global__ void __launch_bounds__(256, (NPL == 3 ? B2_MINB3 : 4)) k_ba_blur2d(const float* __restrict__ in, int w, int h, size_t n,
                                                    const float* __restrict__ inv_x, const float* __restrict__ inv_y,
                                                    float* __restrict__ out_a, float* __restrict__ out_b,
                                                    float* __restrict__ out_c, const __grid_constant__ CUtensorMap map) {
    constexpr int RUP = (R + 3) & ~3;
    constexpr int PITCH = B2_TW + 2 * RUP;
    constexpr int ROWS = B2_TH + 2 * R;
    constexpr int TILE = (ROWS * PITCH + 31) & ~31;   // 128-byte multiple so both buffers are valid TMA destinations
    __shared__ __align__(128) float s_in2[2][TILE];   // item i+1 is staged (TMA / cp.async) while item i is computed
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(16) float s_h[ROWS * B2_TW];
    const int x0 = blockIdx.x * B2_TW;
    const int tiles_y = (h + B2_TH - 1) / B2_TH;
    const int t_begin = blockIdx.y * B2_NT, nt = min(B2_NT, tiles_y - t_begin);
    const int nitems = nt * NPL;
    const size_t img = blockIdx.z;
    const bool vec = (w & 3) == 0;
    // vertical pass: thread = 4 rows (g*4 ..) of the two adjacent columns 2*cp, 2*cp+1, held as packed fp32x2 pairs
    const int cp = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int x = x0 + 2 * cp;
    float res[NPL][4][2], ctr[NPL][4][2];
    auto tma_issue = [&](int i) {   // thread 0: item i -> buffer i & 1 (map = in as [planes][h][w], box (PITCH, ROWS, 1))
        const int t = i / NPL, c = i - t * NPL;
        mbar_expect_tx(&s_bar[i & 1], ROWS * PITCH * 4);
        tma_load_3d(s_in2[i & 1], &map, x0 - RUP, (t_begin + t) * B2_TH - R, (int)(img * NPL + c), &s_bar[i & 1]);
    };
    auto async_issue = [&](int i) {   // all threads: item i -> buffer i & 1 by cp.async
        const int t = i / NPL, c = i - t * NPL;
        load_tile_async<PITCH / 4, ROWS, 256>(s_in2[i & 1], PITCH, in + (img * NPL + c) * n, w, h, x0 - RUP, (t_begin + t) * B2_TH - R, vec);
        cp_async_commit();
    };
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
            tma_issue(0);
            if (nitems > 1) tma_issue(1);
        }
        __syncthreads();
    } else {
        async_issue(0);
    }
    for (int t = 0; t < nt; t++) {
        const int y0 = (t_begin + t) * B2_TH;
#pragma unroll
        for (int c = 0; c < NPL; c++) {
            const int i = t * NPL + c;
            if (TMA) {
                mbar_wait(&s_bar[i & 1], (unsigned)(i >> 1) & 1u);
            } else if (i + 1 < nitems) {
                async_issue(i + 1);   // buffer (i+1) & 1 was released by the second barrier of item i-1
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();   // item i staged; the previous item's vertical pass is done with s_h
            const float* s_in = s_in2[i & 1];
            for (int e = threadIdx.x; e < ROWS * (B2_TW / 4); e += 256) {
                const int r = e >> 4, q4 = e & 15;
                float v[4 + 2 * RUP];
#pragma unroll
                for (int q = 0; q < (4 + 2 * RUP) / 4; q++) {
                    float4 f = *reinterpret_cast<const float4*>(&s_in[r * PITCH + q4 * 4 + q * 4]);
                    v[q * 4] = f.x; v[q * 4 + 1] = f.y; v[q * 4 + 2] = f.z; v[q * 4 + 3] = f.w;
                }
                const float4 iv = *reinterpret_cast<const float4*>(inv_x + min(x0 + q4 * 4, (w + 3) & ~3));   // zero-padded table
                const float ivk[4] = {iv.x, iv.y, iv.z, iv.w};
                float o4[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float sum = 0.0f;
#pragma unroll
                    for (int tt = 0; tt <= 2 * R; tt++) sum = __fmaf_rn(v[k + (RUP - R) + tt], baw<SLOT>(tt), sum);
                    o4[k] = sum * ivk[k];
                }
                *reinterpret_cast<float4*>(&s_h[r * B2_TW + q4 * 4]) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            }
            if (EPI != 0) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 tc = *reinterpret_cast<const float2*>(&s_in[(g * 4 + k + R) * PITCH + 2 * cp + RUP]);
                    ctr[c][k][0] = tc.x; ctr[c][k][1] = tc.y;
                }
            }
            if (TMA && i + 2 < nitems) fence_proxy_async();   // the ctr loads above may still be in flight
            __syncthreads();
            if (TMA && i + 2 < nitems && threadIdx.x == 0) tma_issue(i + 2);   // everyone is done with buffer i & 1
            f32x2 v[4 + 2 * R];
#pragma unroll
            for (int q = 0; q < 4 + 2 * R; q++) v[q] = *reinterpret_cast<const f32x2*>(&s_h[(g * 4 + q) * B2_TW + 2 * cp]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int y = y0 + g * 4 + k;
                f32x2 sum = 0ULL;
#pragma unroll
                for (int tt = 0; tt <= 2 * R; tt++) sum = fma2(v[k + tt], baw2<SLOT>(tt), sum);
                const float iy = y < h ? inv_y[y] : 0.0f;
                float lo, hi;
                unpk2(sum, lo, hi);
                res[c][k][0] = lo * iy;
                res[c][k][1] = hi * iy;
            }
        }
        // epilogue of tile t from registers.  Even widths: the lane's two columns leave as one 64-bit store (a warp then
        // writes whole 128-byte lines instead of every other word of them)
        if (x >= w) continue;
        const bool two = x + 1 < w, pair = (w & 1) == 0;
        auto put = [&](float* d, float v0, float v1) {
            if (pair) *reinterpret_cast<float2*>(d) = make_float2(v0, v1);
            else {
                d[0] = v0;
                if (two) d[1] = v1;
            }
        };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int y = y0 + g * 4 + k;
            if (y >= h) break;
            const size_t i = (size_t)y * w + x;
            if (EPI == 0) {
#pragma unroll
                for (int c = 0; c < NPL; c++) put(out_a + (img * NPL + c) * n + i, res[c][k][0], res[c][k][1]);
            } else if (EPI == 2) {
                float m0[2], m1[2], h0[2], h1[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const float bx = res[0][k][j], by = res[NPL > 1 ? 1 : 0][k][j];
                    const float hfx = ctr[0][k][j] - bx, hfy = ctr[NPL > 1 ? 1 : 0][k][j] - by;
                    m0[j] = ba_remove_range(bx, 0.29f);
                    m1[j] = ba_amplify_range(by, 0.1f);
                    const float scaler = __fmaf_rn(46.0f / __fmaf_rn(hfy, hfy, 46.0f), (float)(1.0 - 0.653020556257), 0.653020556257f);
                    h0[j] = scaler * hfx;
                    h1[j] = hfy;
                }
                float* M = out_a + img * 3 * n + i;
                put(M, m0[0], m0[1]);
                put(M + n, m1[0], m1[1]);
                put(M + 2 * n, res[NPL > 2 ? 2 : 0][k][0], res[NPL > 2 ? 2 : 0][k][1]);
                float* H = out_b + img * 2 * n + i;
                put(H, h0[0], h0[1]);
                put(H + n, h1[0], h1[1]);
            } else {
                float hx[2], ux[2], hy[2], uy[2], mk[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    {
                        const float hb = res[0][k][j];
                        ux[j] = ba_remove_range(ctr[0][k][j] - hb, 0.04f);
                        hx[j] = ba_remove_range(hb, 1.5f);
                    }
                    {
                        float hb = ba_max_clamp(res[NPL > 1 ? 1 : 0][k][j], 28.4691806922f);
                        float u = ctr[NPL > 1 ? 1 : 0][k][j] - hb;
                        u = ba_max_clamp(u, 5.19175294647f);
                        uy[j] = u * 2.69313763794f;
                        hb = hb * 2.155f;
                        hy[j] = ba_amplify_range(hb, 0.132f);
                    }
                    // mask input: DiffPrecompute(sqrt(((uhf_x+hf_x)*2.5)^2 + (uhf_y*0.4+hf_y*0.4)^2))
                    const float kMul = 6.19424080439f, kBias = 12.61050594197f;
                    const float bias = kMul * kBias;
                    const float xd = (ux[j] + hx[j]) * 2.5f;
                    const float yd = uy[j] * 0.4f + hy[j] * 0.4f;
                    const float vv = sqrt_rn_nonneg(xd * xd + yd * yd);
                    mk[j] = sqrt_rn_nonneg(kMul * fabsf(vv) + bias) - sqrtf(bias);
                }
                float* H = out_a + img * 2 * n + i;
                float* U = out_b + img * 2 * n + i;
                put(H, hx[0], hx[1]);
                put(H + n, hy[0], hy[1]);
                put(U, ux[0], ux[1]);
                put(U + n, uy[0], uy[1]);
                put(out_c + img * n + i, mk[0], mk[1]);
            }
        }
    }
    if (!TMA) cp_async_wait<0>();
}

struct BlurTables {
    float* inv_x[4];
    float* inv_y[4];
};

// ---------------------------------------------------------------- Malta
#define MT_TW 64
#define MT_TH 16
#define MT_P (MT_TW + 8)
#define MT_ROWS (MT_TH + 8)
#define MT_WIN_ROWS 10   // a thread's register window: its 2 output rows + 4 rows of halo above and below
// The 16 line sums of a pixel share sub-sums with one another and with the sums of the thread's other pixels; they
// are generated (tools/gen_malta.py).  Mathematically the upstream sums; the association differs, i.e. ~1e-7 relative.
#include "malta_sums.inc"


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
    // upstream divides twice by the same denominator; one IEEE reciprocal serves both scalers here (each product is
    // within an ulp of the quotient upstream forms; the diffmap moves by ~1e-7 relative, the contract is 1e-3).
    // The kernel was XU / issue bound on the two divisions (47 % XU pipe): 1.11 -> 1.00 ms on the 192-pair batch.
    const float inv = 1.0f / (p.norm1 + absval);
    float scaler = p.norm2_0gt1 * inv;
    float d = scaler * diff;
    float scaler2 = p.norm2_0lt1 * inv;
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


struct MaltaParams2 {
    MaltaParams ch[2];
};

// Pointwise: the asymmetric, masked difference the Malta filters sum over, for the 3 bands x {X,Y} of every
// pair.  uhf,hf: [NI][2][n]; mf: [NI][3][n]  ->  diff [B][2 ch][3 bands][n].  One thread = 4 pixels (VEC) or 1.
template <bool VEC>
__global__ void __launch_bounds__(256) k_ba_malta_diff(const float* __restrict__ uhf, const float* __restrict__ hf,
                                                        const float* __restrict__ mf, size_t n, size_t B, size_t R,
                                                        const int* __restrict__ ridx,
                                                        const __grid_constant__ MaltaParams2 prm2, float* __restrict__ diff) {
    // grid (blocks over the plane, 2B): blockIdx.y = b*2 + C (a 1-D grid-stride loop needed a 64-bit division per step)
    const size_t per = VEC ? n / 4 : n;
    const size_t bc = blockIdx.y;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < per; q += (size_t)gridDim.x * blockDim.x) {
        const size_t b = bc >> 1;
        const int C = (int)(bc & 1);
        const size_t i = VEC ? q * 4 : q;
        const size_t i0 = (size_t)ridx[b], i1 = R + b;
        const float* p0[3] = {uhf + (i0 * 2 + C) * n + i, hf + (i0 * 2 + C) * n + i, mf + (i0 * 3 + C) * n + i};
        const float* p1[3] = {uhf + (i1 * 2 + C) * n + i, hf + (i1 * 2 + C) * n + i, mf + (i1 * 3 + C) * n + i};
        float* o = diff + bc * 3 * n + i;
        if (VEC) {
            float4 a[3], d[3];
#pragma unroll
            for (int bd = 0; bd < 3; bd++) {
                a[bd] = *reinterpret_cast<const float4*>(p0[bd]);
                d[bd] = *reinterpret_cast<const float4*>(p1[bd]);
            }
#pragma unroll
            for (int bd = 0; bd < 3; bd++) {
                const MaltaBand& mb = prm2.ch[C].band[bd];
                *reinterpret_cast<float4*>(o + (size_t)bd * n) =
                    make_float4(malta_diff(a[bd].x, d[bd].x, mb), malta_diff(a[bd].y, d[bd].y, mb),
                                malta_diff(a[bd].z, d[bd].z, mb), malta_diff(a[bd].w, d[bd].w, mb));
            }
        } else {
#pragma unroll
            for (int bd = 0; bd < 3; bd++) o[(size_t)bd * n] = malta_diff(*p0[bd], *p1[bd], prm2.ch[C].band[bd]);
        }
    }
}

// grid (tiles_x, ceil(tiles_y / MT_NT), 2B): blockIdx.z = b*2 + C.  diff: [B][2][3][n]; hf: [NI][2][n]; mf: [NI][3][n];
// ac out: [B][2][n] plane C.  A block of 128 threads walks down MT_NT vertically adjacent 64x16 tiles.  Per tile the
// three band tiles (+ halo 4, zero outside the image) and the four pointwise inputs of the L2 terms are staged by TMA
// (cp.async for widths TMA cannot describe) into one of two buffers, so the loads of tile t+1 run under the arithmetic
// of tile t.  A thread owns 4 x 2 pixels: it pulls their 10x12 window into registers (30 LDS.128 per band for 8
// pixels; round 1's 4 x 1 layout needed 27 for 4 pixels and was bound by shared-memory bandwidth: 81 LDS.128 = 324
// shared-memory cycles per warp and band set against 234 cycles of FP32 issue) and evaluates the 16 oriented line sums
// of each pixel from there.
#ifndef MT_NT
#define MT_NT 8
#endif
#define MT_THREADS 128
#define MT_BAND_FLOATS (MT_ROWS * MT_P)
#define MT_BUF_FLOATS (3 * MT_BAND_FLOATS + 4 * MT_TH * MT_TW)
#define MT_SMEM (2 * MT_BUF_FLOATS * 4)
struct MaltaMaps {   // TMA descriptors (used by the TMA variant only): planes as the third tensor dimension
    CUtensorMap diff;   // [B*2*3 planes][h][w]
    CUtensorMap hf;     // [NI*2 planes][h][w]
    CUtensorMap mf;     // [NI*3 planes][h][w]
};
template <bool TMA>
__global__ void __launch_bounds__(MT_THREADS, 3) k_ba_malta(const float* __restrict__ diff, const float* __restrict__ hf,
                                                             const float* __restrict__ mf, int w, int h, size_t n, size_t R,
                                                             const int* __restrict__ ridx,
                                                             const __grid_constant__ MaltaParams2 prm2,
                                                             const __grid_constant__ MaltaMaps maps, float* __restrict__ ac) {
    extern __shared__ __align__(128) float s_mt[];   // [2 buffers][3 band tiles | hf0 hf1 mf0 mf1 tiles]
    __shared__ __align__(8) unsigned long long s_bar[2];
    const size_t b = blockIdx.z >> 1;
    const int C = blockIdx.z & 1;
    const MaltaParams& prm = prm2.ch[C];
    const int tx0 = blockIdx.x * MT_TW;
    const int tiles_y = (h + MT_TH - 1) / MT_TH;
    const int t_begin = blockIdx.y * MT_NT, nt = min(MT_NT, tiles_y - t_begin);
    const int g = threadIdx.x & 15, oy = (threadIdx.x >> 4) * 2;   // columns 4g .. 4g+3, rows oy and oy+1 of the tile
    const bool vec = (w & 3) == 0;
    const size_t im0 = (size_t)ridx[b], im1 = R + b;
    const float* e_src[4] = {hf + (im0 * 2 + C) * n, hf + (im1 * 2 + C) * n, mf + (im0 * 3 + C) * n, mf + (im1 * 3 + C) * n};
    auto issue = [&](int t) {
        if (TMA) {
            // one thread arms the buffer's barrier with the byte count and issues seven bulk tensor copies
            if (threadIdx.x == 0 && t < nt) {
                float* buf = s_mt + (t & 1) * MT_BUF_FLOATS;
                const int ty0 = (t_begin + t) * MT_TH;
                unsigned long long* bar = &s_bar[t & 1];
                mbar_expect_tx(bar, MT_BUF_FLOATS * 4);
#pragma unroll
                for (int bd = 0; bd < 3; bd++)
                    tma_load_3d(buf + bd * MT_BAND_FLOATS, &maps.diff, tx0 - 4, ty0 - 4, (int)blockIdx.z * 3 + bd, bar);
                float* e = buf + 3 * MT_BAND_FLOATS;
                tma_load_3d(e, &maps.hf, tx0, ty0, (int)(im0 * 2 + C), bar);
                tma_load_3d(e + MT_TH * MT_TW, &maps.hf, tx0, ty0, (int)(im1 * 2 + C), bar);
                tma_load_3d(e + 2 * MT_TH * MT_TW, &maps.mf, tx0, ty0, (int)(im0 * 3 + C), bar);
                tma_load_3d(e + 3 * MT_TH * MT_TW, &maps.mf, tx0, ty0, (int)(im1 * 3 + C), bar);
            }
            return;
        }
        if (t < nt) {
            float* buf = s_mt + (t & 1) * MT_BUF_FLOATS;
            const int ty0 = (t_begin + t) * MT_TH;
#pragma unroll
            for (int bd = 0; bd < 3; bd++)
                load_tile_async<MT_P / 4, MT_ROWS, MT_THREADS>(buf + bd * MT_BAND_FLOATS, MT_P, diff + ((size_t)blockIdx.z * 3 + bd) * n,
                                                               w, h, tx0 - 4, ty0 - 4, vec);
#pragma unroll
            for (int e = 0; e < 4; e++)
                load_tile_async<MT_TW / 4, MT_TH, MT_THREADS>(buf + 3 * MT_BAND_FLOATS + e * MT_TH * MT_TW, MT_TW, e_src[e], w, h, tx0,
                                                              ty0, vec);
        }
        cp_async_commit();
    };
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    issue(0);
    issue(1);
    const int x = tx0 + 4 * g;
    for (int t = 0; t < nt; t++) {
        if (TMA) {
            mbar_wait(&s_bar[t & 1], (unsigned)(t >> 1) & 1u);   // the bytes of tile t have landed
        } else {
            cp_async_wait<1>();
            __syncthreads();   // tile t visible to all
        }
        const float* buf = s_mt + (t & 1) * MT_BUF_FLOATS;
        const int y = (t_begin + t) * MT_TH + oy;
        float acc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
#pragma unroll
        for (int bd = 0; bd < 3; bd++) {
            float win[MT_WIN_ROWS][12];
            const float* sb = buf + bd * MT_BAND_FLOATS;
#pragma unroll
            for (int r = 0; r < MT_WIN_ROWS; r++) {
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const float4 f = *reinterpret_cast<const float4*>(&sb[(oy + r) * MT_P + 4 * g + 4 * q]);
                    win[r][4 * q] = f.x; win[r][4 * q + 1] = f.y; win[r][4 * q + 2] = f.z; win[r][4 * q + 3] = f.w;
                }
            }
            float accb[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};   // a band's squares sum from 0, as upstream
            if (bd == 0) malta_hf8(win, accb);
            else malta_lf8(win, accb);
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) acc[j][k] += accb[j][k];
        }
        float4 pw[2][4];   // hf(ref), hf(dist), mf(ref), mf(dist) of the thread's two rows
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float* se = buf + 3 * MT_BAND_FLOATS + (oy + j) * MT_TW + 4 * g;
#pragma unroll
            for (int e = 0; e < 4; e++) pw[j][e] = *reinterpret_cast<const float4*>(se + e * MT_TH * MT_TW);
        }
        if (TMA) fence_proxy_async();
        __syncthreads();   // everyone is done reading buffer t & 1
        issue(t + 2);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float hv0[4] = {pw[j][0].x, pw[j][0].y, pw[j][0].z, pw[j][0].w}, hv1[4] = {pw[j][1].x, pw[j][1].y, pw[j][1].z, pw[j][1].w};
            const float mv0[4] = {pw[j][2].x, pw[j][2].y, pw[j][2].z, pw[j][2].w}, mv1[4] = {pw[j][3].x, pw[j][3].y, pw[j][3].z, pw[j][3].w};
            float tot[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // the three L2 products, added after the Malta sums in the upstream order
                const float v0 = hv0[k], v1 = hv1[k];
                const float df = v0 - v1;
                const float l2a = df * df;
                const float fabs0 = fabsf(v0);
                const float too_small = 0.4f * fabs0, too_big = fabs0;
                const float if_neg = v1 > -too_small ? v1 + too_small : (v1 < -too_big ? -v1 - too_big : 0.0f);
                const float if_pos = v1 < too_small ? too_small - v1 : (v1 > too_big ? v1 - too_big : 0.0f);
                const float v = v0 < 0.0f ? if_neg : if_pos;
                const float l2b = v * v;
                const float dm = mv0[k] - mv1[k];
                const float l2c = dm * dm;
                float total = acc[j][k];
                total = __fmaf_rn(l2a, prm.l2_hf_gt, total);   // L2DiffAsymmetric on hf
                total = __fmaf_rn(prm.l2_hf_lt, l2b, total);
                total = __fmaf_rn(l2c, prm.l2_mf, total);      // L2Diff on mf
                tot[k] = total;
            }
            if (y + j < h && x < w) {
                float* o = ac + (b * 2 + C) * n + (size_t)(y + j) * w + x;
                if (vec) *reinterpret_cast<float4*>(o) = make_float4(tot[0], tot[1], tot[2], tot[3]);
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x + k < w) o[k] = tot[k];
                }
            }
        }
    }
    if (!TMA) cp_async_wait<0>();
}

static void ba_set_kernel_attributes() {
    CE_CUDA(cudaFuncSetAttribute(k_ba_malta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
    CE_CUDA(cudaFuncSetAttribute(k_ba_malta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
}

// ---------------------------------------------------------------- combine
CE_DEVINL void store_min3(float v, float& m0, float& m1, float& m2) {
    if (v < m2) {
        if (v < m0) { m2 = m1; m1 = m0; m0 = v; }
        else if (v < m1) { m2 = m1; m1 = v; }
        else m2 = v;
    }
}

// The mask (fuzzy erosion of the REFERENCE's blurred mask input, then MaskY / MaskDcY) does not depend on the distorted
// image: it is formed once per distinct reference of the launch and the per-pair kernel reads the two factors.
// bl: [NI][n] blurred mask inputs (references first); mask out: [R][2][n] = {MaskY, MaskDcY}.
__global__ void __launch_bounds__(256) k_ba_mask(const float* __restrict__ bl, int w, int h, size_t n, float* __restrict__ mask) {
    const int S = 3;
    const size_t r = blockIdx.y;
    const float* from = bl + r * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int y = (int)(i / w), x = (int)(i - (size_t)y * w);
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
        const float mk = (0.45f * m0 + 0.3f * m1) + 0.25f * m2;
        mask[(r * 2 + 0) * n + i] = ba_mask_y(mk);
        mask[(r * 2 + 1) * n + i] = ba_mask_dc_y(mk);
    }
}

// sorted insert into the three smallest (m0 <= m1 <= m2), branch free: the values store_min3 leaves
CE_DEVINL void insert_min3(float v, float& m0, float& m1, float& m2) {
    const float a = fminf(m0, v), v1 = fmaxf(m0, v);
    const float b = fminf(m1, v1), v2 = fmaxf(m1, v1);
    m0 = a; m1 = b; m2 = fminf(m2, v2);
}

// 4 pixels per thread (w % 4 == 0): the 3x3 stride-3 neighbourhood comes from three aligned 128-bit loads per row
// (columns x-4 .. x+7); positions outside the image hold +inf, which an insert never selects (the blurred mask input
// is >= 0, so the start triple {c, 2c, 2c} is sorted).
__global__ void __launch_bounds__(256) k_ba_mask4(const float* __restrict__ bl, int w, int h, size_t n, float* __restrict__ mask) {
    const int S = 3;
    const int w4 = w >> 2;
    const int per = (int)(n >> 2);
    const size_t r = blockIdx.y;
    const float* from = bl + r * n;
    const float inf = __int_as_float(0x7f800000);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
        const int y = q / w4, x = (q - y * w4) * 4;
        const int i = y * w + x;
        float win[3][12];
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            const int yy = y + (rr - 1) * S;
            const bool yok = yy >= 0 && yy < h;
#pragma unroll
            for (int c4 = 0; c4 < 3; c4++) {
                const int xx = x - 4 + 4 * c4;
                float4 v = make_float4(inf, inf, inf, inf);
                if (yok && xx >= 0 && xx < w) v = *reinterpret_cast<const float4*>(from + yy * w + xx);
                win[rr][4 * c4] = v.x; win[rr][4 * c4 + 1] = v.y; win[rr][4 * c4 + 2] = v.z; win[rr][4 * c4 + 3] = v.w;
            }
        }
        float my[4], mdc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = 4 + k;   // centre column in the window
            float mn0 = win[1][c], mn1 = 2.0f * mn0, mn2 = mn1;
            insert_min3(win[1][c - S], mn0, mn1, mn2);
            insert_min3(win[0][c - S], mn0, mn1, mn2);
            insert_min3(win[2][c - S], mn0, mn1, mn2);
            insert_min3(win[1][c + S], mn0, mn1, mn2);
            insert_min3(win[0][c + S], mn0, mn1, mn2);
            insert_min3(win[2][c + S], mn0, mn1, mn2);
            insert_min3(win[0][c], mn0, mn1, mn2);
            insert_min3(win[2][c], mn0, mn1, mn2);
            const float mk = (0.45f * mn0 + 0.3f * mn1) + 0.25f * mn2;
            my[k] = ba_mask_y(mk);
            mdc[k] = ba_mask_dc_y(mk);
        }
        *reinterpret_cast<float4*>(mask + (r * 2 + 0) * n + i) = make_float4(my[0], my[1], my[2], my[3]);
        *reinterpret_cast<float4*>(mask + (r * 2 + 1) * n + i) = make_float4(mdc[0], mdc[1], mdc[2], mdc[3]);
    }
}

// bl: [NI][n] blurred mask inputs; mask: [R][2][n]; ac: [B][2][n]; mf, lf: [NI][3][n]; diffmap out [B][n]
__global__ void __launch_bounds__(256) k_ba_combine(const float* __restrict__ bl, const float* __restrict__ mask,
                                                     const float* __restrict__ ac, const float* __restrict__ mf,
                                                     const float* __restrict__ lf, size_t n, size_t B, size_t R,
                                                     const int* __restrict__ ridx, float xmul, float* __restrict__ diffmap) {
    const size_t b = blockIdx.y;
    const size_t i0 = (size_t)ridx[b], i1 = R + b;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float dmk = bl[i0 * n + i] - bl[i1 * n + i];
        float ac0 = ac[(b * 2 + 0) * n + i];
        float ac1 = ac[(b * 2 + 1) * n + i];
        ac1 += (10.0f * dmk) * dmk;
        const float* M0 = mf + i0 * 3 * n + i;
        const float* M1 = mf + i1 * 3 * n + i;
        const float* L0 = lf + i0 * 3 * n + i;
        const float* L1 = lf + i1 * 3 * n + i;
        float d2 = M0[2 * n] - M1[2 * n];
        float ac2 = (d2 * d2) * 16.2176043152f;
        float e0 = L0[0] - L1[0], e1 = L0[n] - L1[n], e2 = L0[2 * n] - L1[2 * n];
        float dc0 = (e0 * e0) * 29.2353797994f;
        float dc1 = (e1 * e1) * 0.844626970982f;
        float dc2 = (e2 * e2) * 0.703646627719f;
        float maskval = mask[(i0 * 2 + 0) * n + i], dc_maskval = mask[(i0 * 2 + 1) * n + i];
        float dsum = ((dc0 * xmul) * dc_maskval + dc1 * dc_maskval) + dc2 * dc_maskval;
        float asum = ((ac0 * xmul) * maskval + ac1 * maskval) + ac2 * maskval;
        diffmap[b * n + i] = sqrtf(dsum + asum);
    }
}

// 4 pixels per thread (n % 4 == 0): one 128-bit load per plane.  grid (blocks over the plane, B)
__global__ void __launch_bounds__(256) k_ba_combine4(const float* __restrict__ bl, const float* __restrict__ mask,
                                                      const float* __restrict__ ac, const float* __restrict__ mf,
                                                      const float* __restrict__ lf, size_t n, size_t B, size_t R,
                                                      const int* __restrict__ ridx, float xmul, float* __restrict__ diffmap) {
    const int per = (int)(n >> 2);
    const size_t b = blockIdx.y;
    const size_t i0 = (size_t)ridx[b], i1 = R + b;
    auto ld = [](const float* p) { const float4 v = *reinterpret_cast<const float4*>(p); return v; };
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
        const size_t i = (size_t)q * 4;
        const float4 b0 = ld(bl + i0 * n + i), b1 = ld(bl + i1 * n + i);
        const float4 a0 = ld(ac + (b * 2 + 0) * n + i), a1 = ld(ac + (b * 2 + 1) * n + i);
        const float4 m0 = ld(mf + (i0 * 3 + 2) * n + i), m1 = ld(mf + (i1 * 3 + 2) * n + i);
        const float4 ky = ld(mask + (i0 * 2 + 0) * n + i), kdc = ld(mask + (i0 * 2 + 1) * n + i);
        float4 l0[3], l1[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            l0[c] = ld(lf + (i0 * 3 + c) * n + i);
            l1[c] = ld(lf + (i1 * 3 + c) * n + i);
        }
        const float b0v[4] = {b0.x, b0.y, b0.z, b0.w}, b1v[4] = {b1.x, b1.y, b1.z, b1.w};
        const float a0v[4] = {a0.x, a0.y, a0.z, a0.w}, a1v[4] = {a1.x, a1.y, a1.z, a1.w};
        const float m0v[4] = {m0.x, m0.y, m0.z, m0.w}, m1v[4] = {m1.x, m1.y, m1.z, m1.w};
        const float kyv[4] = {ky.x, ky.y, ky.z, ky.w}, kdcv[4] = {kdc.x, kdc.y, kdc.z, kdc.w};
        const float l0v[3][4] = {{l0[0].x, l0[0].y, l0[0].z, l0[0].w}, {l0[1].x, l0[1].y, l0[1].z, l0[1].w}, {l0[2].x, l0[2].y, l0[2].z, l0[2].w}};
        const float l1v[3][4] = {{l1[0].x, l1[0].y, l1[0].z, l1[0].w}, {l1[1].x, l1[1].y, l1[1].z, l1[1].w}, {l1[2].x, l1[2].y, l1[2].z, l1[2].w}};
        float res[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float dmk = b0v[k] - b1v[k];
            const float ac0 = a0v[k];
            float ac1 = a1v[k];
            ac1 += (10.0f * dmk) * dmk;
            const float d2 = m0v[k] - m1v[k];
            const float ac2 = (d2 * d2) * 16.2176043152f;
            const float e0 = l0v[0][k] - l1v[0][k], e1 = l0v[1][k] - l1v[1][k], e2 = l0v[2][k] - l1v[2][k];
            const float dc0 = (e0 * e0) * 29.2353797994f;
            const float dc1 = (e1 * e1) * 0.844626970982f;
            const float dc2 = (e2 * e2) * 0.703646627719f;
            const float maskval = kyv[k], dc_maskval = kdcv[k];
            const float dsum = ((dc0 * xmul) * dc_maskval + dc1 * dc_maskval) + dc2 * dc_maskval;
            const float asum = ((ac0 * xmul) * maskval + ac1 * maskval) + ac2 * maskval;
            res[k] = sqrt_rn_nonneg(dsum + asum);
        }
        *reinterpret_cast<float4*>(diffmap + b * n + i) = make_float4(res[0], res[1], res[2], res[3]);
    }
}

// ---------------------------------------------------------------- multi-resolution
// SubSample2x on linear planes: [np][n] -> [np][on]; ((a+b)+c)+d)*0.25 with the x2 fix-ups
// grid (blocks over a plane, planes): 32-bit index arithmetic inside a plane
__global__ void __launch_bounds__(256) k_ba_subsample(const float* __restrict__ in, int w, int h, size_t n, int ow, int oh,
                                                       size_t on, size_t total, float* __restrict__ out) {
    const size_t pl = blockIdx.y;
    const float* p = in + pl * n;
    if ((w & 3) == 0) {
        // two outputs per thread from one 128-bit load per input row (w % 4 == 0: no odd column, ow is even)
        const unsigned ow2 = (unsigned)ow >> 1, per = (unsigned)(on >> 1);
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
            const unsigned oy = q / ow2, ox = (q - oy * ow2) * 2;
            const unsigned y0 = 2 * oy;
            const bool vy = (int)y0 + 1 < h;
            const float4 a = *reinterpret_cast<const float4*>(p + (size_t)y0 * w + 2 * ox);
            float s0 = 0.0f, s1 = 0.0f;
            s0 += 0.25f * a.x; s0 += 0.25f * a.y;
            s1 += 0.25f * a.z; s1 += 0.25f * a.w;
            if (vy) {
                const float4 b = *reinterpret_cast<const float4*>(p + (size_t)(y0 + 1) * w + 2 * ox);
                s0 += 0.25f * b.x; s0 += 0.25f * b.y;
                s1 += 0.25f * b.z; s1 += 0.25f * b.w;
            }
            if ((h & 1) && (int)oy == oh - 1) { s0 *= 2.0f; s1 *= 2.0f; }
            *reinterpret_cast<float2*>(out + pl * on + (size_t)oy * ow + ox) = make_float2(s0, s1);
        }
        return;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)on; i += gridDim.x * blockDim.x) {
        const size_t t = pl * on + i;
        int oy = (int)(i / (unsigned)ow), ox = (int)(i - (unsigned)oy * (unsigned)ow);
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
    auto pool = [&](float v) {
        mx = fmaxf(mx, v);
        const double d = (double)v;
        const double d3 = d * d * d;
        s3 += d3;
        const double d6 = d3 * d3;
        s6 += d6;
        s12 += d6 * d6;
    };
    if ((w & 3) == 0) {
        // 4 pixels per thread: one 128-bit load / store of the map, one 64-bit load of the two half-resolution values
        const int w4 = w >> 2, per = (int)(n >> 2);
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
            float4 v = *reinterpret_cast<const float4*>(dm + (size_t)q * 4);
            if (sb) {
                const int y = q / w4, x = (q - y * w4) * 4;
                const float2 u = *reinterpret_cast<const float2*>(sb + (size_t)(y >> 1) * sw + (x >> 1));
                v.x = v.x * keep; v.y = v.y * keep; v.z = v.z * keep; v.w = v.w * keep;
                v.x = v.x + 0.5f * u.x; v.y = v.y + 0.5f * u.x; v.z = v.z + 0.5f * u.y; v.w = v.w + 0.5f * u.y;
                *reinterpret_cast<float4*>(dm + (size_t)q * 4) = v;
            }
            pool(v.x); pool(v.y); pool(v.z); pool(v.w);
        }
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            float v = dm[i];
            if (sb) {
                int y = (int)(i / w), x = (int)(i - (size_t)y * w);
                v = v * keep;
                v = v + 0.5f * sb[(size_t)(y >> 1) * sw + (x >> 1)];
                dm[i] = v;
            }
            pool(v);
        }
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


// per-level planes.  Aliasing (a buffer is reused once its producer's consumers have all run):
//   hf_pre = tmp[0..1]            (tmp is dead after the LF vertical pass)
//   hf, m  = xyb[0..1], xyb[2]    (xyb is dead after the LF vertical pass)
//   uhf, bl = mf_pre[0..1], [2]   (mf_pre is dead after the MF blur)
struct BaLevelBufs {
    float *xyb, *tmp, *lf, *mf_pre, *mf, *ac;
    float *hf_pre, *hf, *uhf, *m, *bl;   // views, strides below
    float* mdiff;                        // [B][2][3][n] Malta difference planes
    BlurTables tables;
};

static size_t ba_level_floats_per_pair(size_t n) {
    // per image: xyb 3, tmp 3, lf 3, mf_pre 3, mf 3 = 15; at most 2 images per pair = 30; pair: ac 2, Malta diffs 6
    return 38 * n;
}

[[maybe_unused]] static unsigned ew_blocks(Context& c, size_t total) {
    return (unsigned)std::min<size_t>(cdiv(total, 256), (size_t)c.sm_count * 32);
}
// 2-D variant: grid.y = units (pairs or pair-channels), grid.x = blocks of 256 threads striding over `per` items of a unit,
// about sm_count * 32 blocks in all
static dim3 ew_grid2(Context& c, size_t per, size_t units) {
    const size_t want = std::max<size_t>(1, cdiv((size_t)c.sm_count * 32, units));
    return dim3((unsigned)std::min<size_t>(cdiv(per, 256), want), (unsigned)units);
}

static void ba_alloc_level(Context& c, size_t NI, size_t B, size_t w, size_t h, BaLevelBufs& L) {
    const size_t n = w * h;
    L.xyb = c.arena.alloc<float>(NI * 3 * n);
    L.tmp = c.arena.alloc<float>(NI * 3 * n);
    L.lf = c.arena.alloc<float>(NI * 3 * n);
    L.mf_pre = c.arena.alloc<float>(NI * 3 * n);
    L.mf = c.arena.alloc<float>(NI * 3 * n);
    L.ac = c.arena.alloc<float>(B * 2 * n);
    L.mdiff = c.arena.alloc<float>(B * 6 * n);
    L.hf_pre = L.tmp;                  // [NI][2][n]
    L.hf = L.xyb;                      // [NI][2][n]
    L.m = L.xyb + NI * 2 * n;          // [NI][n]
    L.uhf = L.mf_pre;                  // [NI][2][n]
    L.bl = L.mf_pre + NI * 2 * n;      // [NI][n]
    for (int s = 0; s < 4; s++) {
        L.tables.inv_x[s] = ba_inv_table(c, s, w);
        L.tables.inv_y[s] = ba_inv_table(c, s, h);
    }
}

static void check_grid_z(size_t z) {
    if (z > 65535) throw CudaError("butteraugli sub-batch too large for one launch");
}

// lin: [NI][3][n] -> psycho planes in L (lf, mf, hf, uhf) and the mask input m
static void ba_psycho_level(Context& c, const float* lin, size_t NI, size_t w, size_t h, float intensity, BaLevelBufs& L,
                            float* dbg_opsin) {
    const size_t n = w * h;
    const int vec = (w % 4 == 0) ? 1 : 0;
    check_grid_z(NI * 3);
    {
        dim3 grid(cdiv(w, OP_TW), cdiv(h, OP_TH), (unsigned)NI);
        CUtensorMap mo;
        memset(&mo, 0, sizeof(mo));
        if (tma_enabled(5) && tma_plane_map(&mo, lin, w, h, NI * 3, OP_P, OP_ROWS, 3))
            CE_LAUNCH(c, "k_ba_opsin", (double)NI * n * 24,
                      k_ba_opsin<true, true><<<grid, 256, 0, c.stream>>>(lin, (int)w, (int)h, n, intensity, L.xyb, vec, mo));
        else
            CE_LAUNCH(c, "k_ba_opsin", (double)NI * n * 24,
                      k_ba_opsin<true, false><<<grid, 256, 0, c.stream>>>(lin, (int)w, (int)h, n, intensity, L.xyb, vec, mo));
    }
    if (dbg_opsin) CE_CUDA(cudaMemcpyAsync(dbg_opsin, L.xyb, 3 * n * 4, cudaMemcpyDeviceToDevice, c.stream));
    {
        dim3 gh(cdiv(w, 128), cdiv(h, 8 * BH_NT), (unsigned)(NI * 3));
        CUtensorMap mh, mv;
        memset(&mh, 0, sizeof(mh));
        memset(&mv, 0, sizeof(mv));
        const bool tma = tma_enabled(2) && tma_plane_map(&mh, L.xyb, w, h, NI * 3, 128 + 2 * 16, 8, 1) && tma_plane_map(&mv, L.tmp, w, h, NI * 3, 32, 64 + 2 * 16, 1);
        if (tma)
            CE_LAUNCH(c, "k_ba_blur_h<R16>", (double)NI * n * 24,
                      k_ba_blur_h<0, 16, true><<<gh, 256, 0, c.stream>>>(L.xyb, (int)w, (int)h, n, L.tables.inv_x[0], L.tmp, vec, mh));
        else
            CE_LAUNCH(c, "k_ba_blur_h<R16>", (double)NI * n * 24,
                      k_ba_blur_h<0, 16, false><<<gh, 256, 0, c.stream>>>(L.xyb, (int)w, (int)h, n, L.tables.inv_x[0], L.tmp, vec, mh));
        dim3 gv(cdiv(w, 32), cdiv(h, 64), (unsigned)NI);
        if (tma)
            CE_LAUNCH(c, "k_ba_blur_v<R16>+lf", (double)NI * n * 48,
                      k_ba_blur_v<0, 16, 3, 1, true><<<gv, 256, 0, c.stream>>>(L.tmp, (int)w, (int)h, n, L.tables.inv_y[0], L.lf, L.xyb, L.mf_pre, mv));
        else
            CE_LAUNCH(c, "k_ba_blur_v<R16>+lf", (double)NI * n * 48,
                      k_ba_blur_v<0, 16, 3, 1, false><<<gv, 256, 0, c.stream>>>(L.tmp, (int)w, (int)h, n, L.tables.inv_y[0], L.lf, L.xyb, L.mf_pre, mv));
    }
    dim3 g2(cdiv(w, B2_TW), cdiv(cdiv(h, B2_TH), B2_NT), (unsigned)NI);
    {
        CUtensorMap m7, m3;
        memset(&m7, 0, sizeof(m7));
        memset(&m3, 0, sizeof(m3));
        const bool tma = tma_enabled(3) && tma_plane_map(&m7, L.mf_pre, w, h, NI * 3, B2_TW + 16, B2_TH + 14, 1) &&
                         tma_plane_map(&m3, L.hf_pre, w, h, NI * 2, B2_TW + 8, B2_TH + 6, 1);
        if (tma) {
            CE_LAUNCH(c, "k_ba_blur2d<R7>+hf_split", (double)NI * n * 32,
                      k_ba_blur2d<1, 7, 3, 2, true><<<g2, 256, 0, c.stream>>>(L.mf_pre, (int)w, (int)h, n, L.tables.inv_x[1], L.tables.inv_y[1], L.mf,
                                                                            L.hf_pre, nullptr, m7));
            CE_LAUNCH(c, "k_ba_blur2d<R3>+uhf_split", (double)NI * n * 28,
                      k_ba_blur2d<2, 3, 2, 3, true><<<g2, 256, 0, c.stream>>>(L.hf_pre, (int)w, (int)h, n, L.tables.inv_x[2], L.tables.inv_y[2], L.hf,
                                                                            L.uhf, L.m, m3));
        } else {
            CE_LAUNCH(c, "k_ba_blur2d<R7>+hf_split", (double)NI * n * 32,
                      k_ba_blur2d<1, 7, 3, 2, false><<<g2, 256, 0, c.stream>>>(L.mf_pre, (int)w, (int)h, n, L.tables.inv_x[1], L.tables.inv_y[1], L.mf,
                                                                             L.hf_pre, nullptr, m7));
            CE_LAUNCH(c, "k_ba_blur2d<R3>+uhf_split", (double)NI * n * 28,
                      k_ba_blur2d<2, 3, 2, 3, false><<<g2, 256, 0, c.stream>>>(L.hf_pre, (int)w, (int)h, n, L.tables.inv_x[2], L.tables.inv_y[2], L.hf,
                                                                             L.uhf, L.m, m3));
        }
    }
    CE_CUDA(cudaGetLastError());
}

// full diffmap of one resolution for B pairs; lin: [2B][3][n]
static void ba_diffmap_level(Context& c, const float* lin, size_t R, const int* ridx, size_t B, size_t w, size_t h,
                             float intensity, float* diffmap) {
    const size_t n = w * h, NI = R + B;
    const float hf_asym = 1.0f, xmul = 1.0f;
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, NI, B, w, h, L);
    ba_psycho_level(c, lin, NI, w, h, intensity, L, nullptr);
    {
        dim3 g2(cdiv(w, B2_TW), cdiv(cdiv(h, B2_TH), B2_NT), (unsigned)NI);
        CUtensorMap m6;
        memset(&m6, 0, sizeof(m6));
        if (tma_enabled(3) && tma_plane_map(&m6, L.m, w, h, NI, B2_TW + 16, B2_TH + 12, 1))
            CE_LAUNCH(c, "k_ba_blur2d<R6>", (double)NI * n * 8,
                      k_ba_blur2d<3, 6, 1, 0, true><<<g2, 256, 0, c.stream>>>(L.m, (int)w, (int)h, n, L.tables.inv_x[3], L.tables.inv_y[3], L.bl,
                                                                            nullptr, nullptr, m6));
        else
            CE_LAUNCH(c, "k_ba_blur2d<R6>", (double)NI * n * 8,
                      k_ba_blur2d<3, 6, 1, 0, false><<<g2, 256, 0, c.stream>>>(L.m, (int)w, (int)h, n, L.tables.inv_x[3], L.tables.inv_y[3], L.bl,
                                                                             nullptr, nullptr, m6));
    }
    {
        MaltaParams2 mp;
        mp.ch[0] = make_malta_params(0, hf_asym);
        mp.ch[1] = make_malta_params(1, hf_asym);
        if (n % 4 == 0)
            CE_LAUNCH_SHARED(c, "k_ba_malta_diff", ((double)B * 48 + (double)R * 24) * n, (double)B * n * 72,
                      k_ba_malta_diff<true><<<ew_grid2(c, n / 4, 2 * B), 256, 0, c.stream>>>(L.uhf, L.hf, L.mf, n, B, R, ridx, mp, L.mdiff));
        else
            CE_LAUNCH_SHARED(c, "k_ba_malta_diff", ((double)B * 48 + (double)R * 24) * n, (double)B * n * 72,
                      k_ba_malta_diff<false><<<ew_grid2(c, n, 2 * B), 256, 0, c.stream>>>(L.uhf, L.hf, L.mf, n, B, R, ridx, mp, L.mdiff));
        dim3 grid(cdiv(w, MT_TW), cdiv(cdiv(h, MT_TH), MT_NT), (unsigned)(2 * B));
        MaltaMaps maps;
        memset(&maps, 0, sizeof(maps));
        const bool tma = tma_enabled(0) && tma_plane_map(&maps.diff, L.mdiff, w, h, B * 6, MT_P, MT_ROWS, 1) &&
                         tma_plane_map(&maps.hf, L.hf, w, h, NI * 2, MT_TW, MT_TH, 1) &&
                         tma_plane_map(&maps.mf, L.mf, w, h, NI * 3, MT_TW, MT_TH, 1);
        if (tma)
            CE_LAUNCH_SHARED(c, "k_ba_malta", ((double)B * 48 + (double)R * 16) * n, (double)B * n * 64,
                      k_ba_malta<true><<<grid, MT_THREADS, MT_SMEM, c.stream>>>(L.mdiff, L.hf, L.mf, (int)w, (int)h, n, R, ridx, mp, maps, L.ac));
        else   // widths that are not a multiple of 4 cannot be described by a tensor map (16-byte row stride): cp.async tiles
            CE_LAUNCH_SHARED(c, "k_ba_malta", ((double)B * 48 + (double)R * 16) * n, (double)B * n * 64,
                      k_ba_malta<false><<<grid, MT_THREADS, MT_SMEM, c.stream>>>(L.mdiff, L.hf, L.mf, (int)w, (int)h, n, R, ridx, mp, maps, L.ac));
    }
    // the mask factors of the R distinct references go where hf_pre lived (L.tmp is dead after the R3 blur)
    float* mask = L.tmp;
    if (w % 4 == 0) {
        CE_LAUNCH(c, "k_ba_mask", (double)R * n * 12, k_ba_mask4<<<ew_grid2(c, n / 4, R), 256, 0, c.stream>>>(L.bl, (int)w, (int)h, n, mask));
        CE_LAUNCH_SHARED(c, "k_ba_combine", ((double)B * 32 + (double)R * 28) * n, (double)B * n * 60,
                  k_ba_combine4<<<ew_grid2(c, n / 4, B), 256, 0, c.stream>>>(L.bl, mask, L.ac, L.mf, L.lf, n, B, R, ridx, xmul, diffmap));
    } else {
        CE_LAUNCH(c, "k_ba_mask", (double)R * n * 12, k_ba_mask<<<ew_grid2(c, n, R), 256, 0, c.stream>>>(L.bl, (int)w, (int)h, n, mask));
        CE_LAUNCH_SHARED(c, "k_ba_combine", ((double)B * 32 + (double)R * 28) * n, (double)B * n * 60,
                  k_ba_combine<<<ew_grid2(c, n, B), 256, 0, c.stream>>>(L.bl, mask, L.ac, L.mf, L.lf, n, B, R, ridx, xmul, diffmap));
    }
    CE_CUDA(cudaGetLastError());
    c.arena.release(mark);
}

size_t butteraugli_workspace_per_pair(size_t w, size_t h) {
    size_t n = w * h;
    size_t sn = ((w + 1) / 2) * ((h + 1) / 2);
    // level buffers are released between levels: max(full) dominates; + diffmap n + sub lin 6*sn + sub diffmap sn
    return (ba_level_floats_per_pair(n) + n + 7 * sn) * 4 + BA_RED_BLOCKS * 4 * 8 + 65536;
}

void butteraugli_run(Context& c, const float* lin, size_t R, const int* ridx, size_t B, size_t w, size_t h, float intensity,
                     double* d_out, float* dbg_diffmap) {
    const size_t n = w * h, NI = R + B;
    check_grid_z(NI * 3);
    size_t mark = c.arena.mark();
    float* diffmap = c.arena.alloc<float>(B * n);
    double* partial = c.arena.alloc<double>(B * BA_RED_BLOCKS * 4);
    ba_diffmap_level(c, lin, R, ridx, B, w, h, intensity, diffmap);
    const size_t sw = (w + 1) / 2, sh = (h + 1) / 2, sn = sw * sh;
    float* sub = nullptr;
    if (sw >= 8 && sh >= 8) {
        float* slin = c.arena.alloc<float>(NI * 3 * sn);
        sub = c.arena.alloc<float>(B * sn);
        size_t total = NI * 3 * sn;
        CE_LAUNCH(c, "k_ba_subsample", (double)total * 20,
                  k_ba_subsample<<<ew_grid2(c, sn, total / sn), 256, 0, c.stream>>>(lin, (int)w, (int)h, n, (int)sw, (int)sh, sn, total, slin));
        ba_diffmap_level(c, slin, R, ridx, B, sw, sh, intensity, sub);
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
    ba_alloc_level(c, 2, 1, w, h, L);
    // with NI = 1 the alias views must be re-based on one image
    L.m = L.xyb + 2 * n;
    L.bl = L.mf_pre + 2 * n;
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
    ba_alloc_level(c, 2, 1, w, h, L);
    L.m = L.xyb + 2 * w * h;
    L.bl = L.mf_pre + 2 * w * h;
    ba_psycho_level(c, lin, 1, w, h, intensity, L, d_planes3);
    c.arena.release(mark);
}
// one plane through the production blur of that sigma (identity epilogue)
void butteraugli_debug_blur(Context& c, const float* in, size_t w, size_t h, float sigma, float* out) {
    const size_t n = w * h;
    size_t mark = c.arena.mark();
    BaLevelBufs L;
    ba_alloc_level(c, 2, 1, w, h, L);
    const int vec = (w % 4 == 0) ? 1 : 0;
    dim3 g2(cdiv(w, B2_TW), cdiv(cdiv(h, B2_TH), B2_NT), 1);
    if (fabsf(sigma - 1.2f) < 1e-6f) {
        dim3 grid(cdiv(w, OP_TW), cdiv(h, OP_TH), 1);
        CUtensorMap m5;
        memset(&m5, 0, sizeof(m5));
        if (tma_plane_map(&m5, in, w, h, 1, OP_P, OP_ROWS, 1))
            CE_LAUNCH(c, "k_ba_blur5", (double)n * 8, k_ba_opsin<false, true><<<grid, 256, 0, c.stream>>>(in, (int)w, (int)h, n, 0.0f, out, vec, m5));
        else
            CE_LAUNCH(c, "k_ba_blur5", (double)n * 8, k_ba_opsin<false, false><<<grid, 256, 0, c.stream>>>(in, (int)w, (int)h, n, 0.0f, out, vec, m5));
    } else if (fabsf(sigma - kSigmas[0]) < 1e-5f) {
        dim3 gh(cdiv(w, 128), cdiv(h, 8 * BH_NT), 1);
        dim3 gv(cdiv(w, 32), cdiv(h, 64), 1);
        CUtensorMap mh, mv;
        memset(&mh, 0, sizeof(mh));
        memset(&mv, 0, sizeof(mv));
        // the same staging path (TMA when the width allows it) as the production launches
        if (tma_plane_map(&mh, in, w, h, 1, 128 + 2 * 16, 8, 1) && tma_plane_map(&mv, L.tmp, w, h, 1, 32, 64 + 2 * 16, 1)) {
            CE_LAUNCH(c, "k_ba_blur_h<R16>", (double)n * 8,
                      k_ba_blur_h<0, 16, true><<<gh, 256, 0, c.stream>>>(in, (int)w, (int)h, n, L.tables.inv_x[0], L.tmp, vec, mh));
            CE_LAUNCH(c, "k_ba_blur_v<R16>", (double)n * 8,
                      k_ba_blur_v<0, 16, 1, 0, true><<<gv, 256, 0, c.stream>>>(L.tmp, (int)w, (int)h, n, L.tables.inv_y[0], out, nullptr, nullptr, mv));
        } else {
            CE_LAUNCH(c, "k_ba_blur_h<R16>", (double)n * 8,
                      k_ba_blur_h<0, 16, false><<<gh, 256, 0, c.stream>>>(in, (int)w, (int)h, n, L.tables.inv_x[0], L.tmp, vec, mh));
            CE_LAUNCH(c, "k_ba_blur_v<R16>", (double)n * 8,
                      k_ba_blur_v<0, 16, 1, 0, false><<<gv, 256, 0, c.stream>>>(L.tmp, (int)w, (int)h, n, L.tables.inv_y[0], out, nullptr, nullptr, mv));
        }
    } else if (fabsf(sigma - kSigmas[1]) < 1e-5f || fabsf(sigma - kSigmas[2]) < 1e-5f || fabsf(sigma - kSigmas[3]) < 1e-5f) {
        const int slot = fabsf(sigma - kSigmas[1]) < 1e-5f ? 1 : fabsf(sigma - kSigmas[2]) < 1e-5f ? 2 : 3;
        const int R = kRadii[slot], RUP = (R + 3) & ~3;
        CUtensorMap m;
        memset(&m, 0, sizeof(m));
        const bool tma = tma_plane_map(&m, in, w, h, 1, B2_TW + 2 * RUP, B2_TH + 2 * R, 1);
        const float* ix = L.tables.inv_x[slot];
        const float* iy = L.tables.inv_y[slot];
#define BLUR2D_DBG(S, RR)                                                                                                         \
    do {                                                                                                                          \
        if (tma) CE_LAUNCH(c, "k_ba_blur2d<dbg>", (double)n * 8,                                                                  \
                           k_ba_blur2d<S, RR, 1, 0, true><<<g2, 256, 0, c.stream>>>(in, (int)w, (int)h, n, ix, iy, out, nullptr, nullptr, m));  \
        else CE_LAUNCH(c, "k_ba_blur2d<dbg>", (double)n * 8,                                                                      \
                       k_ba_blur2d<S, RR, 1, 0, false><<<g2, 256, 0, c.stream>>>(in, (int)w, (int)h, n, ix, iy, out, nullptr, nullptr, m)); \
    } while (0)
        if (slot == 1) BLUR2D_DBG(1, 7);
        else if (slot == 2) BLUR2D_DBG(2, 3);
        else BLUR2D_DBG(3, 6);
#undef BLUR2D_DBG
    } else {
        throw CudaError("unsupported sigma");
    }
    CE_CUDA(cudaGetLastError());
    c.arena.release(mark);
}

}  // namespace ce
