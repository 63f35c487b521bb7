// k_dssim.cu -- DSSIM (dssim-core 3.4.0; reference call site src/metrics/dssim.rs:52-68)
// for a batch of B pairs.
//
// Per scale (<= 5 scales, floor-halving with crop on LINEAR rgb(a)):
//   k_ds_lab     : pointwise, thread = 2x2 block: linear -> L (final), a/b (to be pre-blurred), and the
//                  2x2 average of the linear planes for the next scale                 [HBM-bound]
//   k_ds_blur2   : chroma pre-blur: the 3x3 kernel applied twice (clamp-replicate per pass) through a
//                  shared-memory tile, 4 positions per thread from 128-bit shared loads [HBM-bound]
//   k_ds_stats   : per channel the five double-3x3 blurs {ch1,ch2,ch1^2,ch2^2,ch1*ch2} of a 64x16 tile,
//                  4 positions per thread from 128-bit shared loads; channel-averaged SSIM map
//                  written once + fp64 block partial of its sum                        [FP32-issue bound]
//   k_ds_mean    : fixed-order reduce -> sum(map), avg = max(mean,0)^(0.5^scale)
//   k_ds_mad     : sum |avg - map_i| in fp64 -> block partials; k_ds_mad_reduce fixes the order
#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

#define DS_TW 64
#define DS_TH 16
#define DS_IW (DS_TW + 8)   // staged columns x0-4 .. x0+67 (16-B aligned start; the blurs need x0-2 .. x0+65)
#define DS_IH (DS_TH + 4)   // rows y0-2 .. y0+17
#define DS_FW (DS_TW + 4)   // first-pass positions x0-2 .. x0+65 (column j <-> x = x0-2+j; j < 66 used)
#define DS_FH (DS_TH + 2)   // first-pass rows y0-1 .. y0+16
#define DS_FG (DS_FW / 4)   // 17 four-column groups per first-pass row

__constant__ float c_dsk[9] = {0.095332f, 0.118095f, 0.095332f, 0.118095f, 0.146293f,
                               0.118095f, 0.095332f, 0.118095f, 0.095332f};

CE_DEVINL float ds_k9(float v00, float v01, float v02, float v10, float v11, float v12, float v20, float v21, float v22) {
    float a = (v00 * c_dsk[0] + v01 * c_dsk[1]) + v02 * c_dsk[2];
    float b = (v10 * c_dsk[3] + v11 * c_dsk[4]) + v12 * c_dsk[5];
    float c = (v20 * c_dsk[6] + v21 * c_dsk[7]) + v22 * c_dsk[8];
    return (a + b) + c;
}

// 3 rows x 8 staged columns (i = 4q .. 4q+7) -> the 3x3 blur at the 4 positions j = 4q .. 4q+3 (centre column i = j+2)
CE_DEVINL float4 ds_k9x4(const float (&r0)[8], const float (&r1)[8], const float (&r2)[8]) {
    float4 o;
    o.x = ds_k9(r0[1], r0[2], r0[3], r1[1], r1[2], r1[3], r2[1], r2[2], r2[3]);
    o.y = ds_k9(r0[2], r0[3], r0[4], r1[2], r1[3], r1[4], r2[2], r2[3], r2[4]);
    o.z = ds_k9(r0[3], r0[4], r0[5], r1[3], r1[4], r1[5], r2[3], r2[4], r2[5]);
    o.w = ds_k9(r0[4], r0[5], r0[6], r1[4], r1[5], r1[6], r2[4], r2[5], r2[6]);
    return o;
}
CE_DEVINL void ds_ld8(const float* s, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// The second 3x3 pass reads its input with clamped coordinates, so a first-pass value at a position outside
// the image is the first-pass value at the nearest inside position.  Border tiles patch those entries.
CE_DEVINL void ds_fix_border(float* __restrict__ s_f, int nplanes, int plane_stride, int w, int h, int x0, int y0) {
    const bool border = x0 - 2 < 0 || x0 + DS_TW + 1 >= w || y0 - 1 < 0 || y0 + DS_TH >= h;
    if (!border) return;   // block-uniform
    for (int e = threadIdx.x; e < DS_FH * DS_FW; e += blockDim.x) {
        const int ry = e / DS_FW, j = e - ry * DS_FW;
        const int x = x0 - 2 + j, y = y0 - 1 + ry;
        const int cx = min(max(x, 0), w - 1), cy = min(max(y, 0), h - 1);
        if (cx != x || cy != y) {
            const int cj = cx - (x0 - 2), cr = cy - (y0 - 1);
            if (cj >= 0 && cj < DS_FW && cr >= 0 && cr < DS_FH)
                for (int f = 0; f < nplanes; f++) s_f[f * plane_stride + e] = s_f[f * plane_stride + cr * DS_FW + cj];
        }
    }
}

// ------------------------------------------------------------------ downsample (alpha plane / generic)
// planes: [nplanes][n] -> [nplanes][on]; floor size, (a+b+c+d)*0.25
__global__ void __launch_bounds__(256) k_ds_down(const float* __restrict__ in, int w, size_t n, int ow, int oh, size_t on,
                                                  size_t total, float* __restrict__ out) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t pl = t / on, i = t - pl * on;
        int y = (int)(i / ow), x = (int)(i - (size_t)y * ow);
        const float* top = in + pl * n + (size_t)(2 * y) * w + 2 * x;
        const float* bot = top + w;
        out[t] = (((top[0] + top[1]) + bot[0]) + bot[1]) * 0.25f;
    }
}

// ------------------------------------------------------------------ Lab (+ next scale)
// thread = one 2x2 block of the current scale; grid.z = image.  linear rgb [NI][3][n] (+ alpha [NI][n]) ->
//   img[i*3 + 0] = L;  chroma[i*2 + {0,1}] = a, b (un-blurred);
//   nlin [NI][3][on] = 2x2 average of the linear planes (floor size: a trailing odd row / column is dropped).
__global__ void __launch_bounds__(256) k_ds_lab(const float* __restrict__ lin, const float* __restrict__ alpha, int w, int h,
                                                 size_t n, float* __restrict__ img, float* __restrict__ chroma,
                                                 int has_next, int ow, int oh, size_t on, float* __restrict__ nlin) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const size_t b = blockIdx.z;
    const int x0 = 2 * ox, y0 = 2 * oy;
    if (x0 >= w || y0 >= h) return;
    const bool vx = x0 + 1 < w, vy = y0 + 1 < h;
    const float* src = lin + b * 3 * n;
    const size_t i00 = (size_t)y0 * w + x0;
    const size_t off[4] = {i00, i00 + (vx ? 1 : 0), i00 + (vy ? (size_t)w : 0), i00 + (vy ? (size_t)w : 0) + (vx ? 1 : 0)};
    float p[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float* pl = src + (size_t)c * n;
#pragma unroll
        for (int k = 0; k < 4; k++) p[c][k] = pl[off[k]];
    }
    if (has_next && vx && vy && ox < ow && oy < oh) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            nlin[(b * 3 + c) * on + (size_t)oy * ow + ox] = (((p[c][0] + p[c][1]) + p[c][2]) + p[c][3]) * 0.25f;
    }
    float* Lp = img + (b * 3) * n;
    float* Ap = chroma + (b * 2) * n;
    float* Bp = Ap + n;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if ((k & 1) && !vx) continue;
        if ((k & 2) && !vy) continue;
        float r = p[0][k], g = p[1][k], bl = p[2][k];
        if (alpha) {
            const float a = alpha[b * n + off[k]];
            if (a < 255.0f / 256.0f) {
                const unsigned x = (unsigned)(x0 + (k & 1)), y = (unsigned)(y0 + (k >> 1));
                const unsigned nn = (x + 11u) ^ (y + 11u);
                if (nn & 16u) r += 1.0f - a;
                if (nn & 8u) g += 1.0f - a;
                if (nn & 32u) bl += 1.0f - a;
            }
        }
        float L, A, Bv;
        ds_to_lab(r, g, bl, L, A, Bv);
        Lp[off[k]] = L; Ap[off[k]] = A; Bp[off[k]] = Bv;
    }
}

// ------------------------------------------------------------------ chroma pre-blur (3x3 kernel applied twice)
// grid (tiles_x, tiles_y, NI): blockIdx.z = image; chroma [NI][2][n] -> img[(z*3) + 1 + {0,1}].
// Both planes of a 64x16 tile (+ halo 2, clamped) are staged; each pass evaluates 4 positions per thread
// from 128-bit shared loads.
__global__ void __launch_bounds__(256) k_ds_blur2(const float* __restrict__ chroma, int w, int h, size_t n,
                                                   float* __restrict__ img) {
    __shared__ __align__(16) float s_ab[2][DS_IH * DS_IW];
    __shared__ __align__(16) float s_f[2][DS_FH * DS_FW];
    const int x0 = blockIdx.x * DS_TW, y0 = blockIdx.y * DS_TH;
    const size_t z = blockIdx.z;
    const bool vec = (w & 3) == 0;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) load_tile<2, DS_IW / 4, DS_IH, 256>(s_ab[pl], DS_IW, chroma + (z * 2 + pl) * n, w, h, x0 - 4, y0 - 2, vec);
    __syncthreads();
    // first 3x3 pass over the positions the second pass needs
    for (int e = threadIdx.x; e < 2 * DS_FH * DS_FG; e += 256) {
        const int pl = e / (DS_FH * DS_FG), r = e - pl * (DS_FH * DS_FG);
        const int ry = r / DS_FG, q = r - ry * DS_FG;
        float r0[8], r1[8], r2[8];
        const float* base = s_ab[pl] + ry * DS_IW + 4 * q;
        ds_ld8(base, r0); ds_ld8(base + DS_IW, r1); ds_ld8(base + 2 * DS_IW, r2);
        *reinterpret_cast<float4*>(&s_f[pl][ry * DS_FW + 4 * q]) = ds_k9x4(r0, r1, r2);
    }
    __syncthreads();
    ds_fix_border(&s_f[0][0], 2, DS_FH * DS_FW, w, h, x0, y0);
    __syncthreads();
    // second pass -> the 4 pixels of this thread
    const int g = threadIdx.x & 15, oy = threadIdx.x >> 4;
    const int x = x0 + 4 * g, y = y0 + oy;
    if (x >= w || y >= h) return;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        float r0[8], r1[8], r2[8];
        const float* base = s_f[pl] + oy * DS_FW + 4 * g;
        ds_ld8(base, r0); ds_ld8(base + DS_FW, r1); ds_ld8(base + 2 * DS_FW, r2);
        const float4 o = ds_k9x4(r0, r1, r2);
        float* d = img + (z * 3 + 1 + pl) * n + (size_t)y * w + x;
        if (vec) *reinterpret_cast<float4*>(d) = o;
        else {
            d[0] = o.x;
            if (x + 1 < w) d[1] = o.y;
            if (x + 2 < w) d[2] = o.z;
            if (x + 3 < w) d[3] = o.w;
        }
    }
}

// ------------------------------------------------------------------ statistics + SSIM map
// The five double-3x3 blurs {ch1, ch2, ch1^2, ch2^2, ch1*ch2} per channel split by who owns them:
//   k_ds_stats<0> (grid.z = distinct reference): mu1 = blur2(ch1), e11 = blur2(ch1^2) -> refstat [R][3][2][n],
//                  once per reference however many distortions it is compared with;
//   k_ds_stats<1> (grid.z = pair): blur2 of {ch2, ch2^2, ch1*ch2}, then the channel-averaged SSIM map
//                  (written once) + fp64 block partial of its sum.  partial: [B][tiles] doubles.
// Block 320, 64x16 tile: the image tile(s) (+ halo 2, clamped) are staged per channel; the first 3x3 pass is
// evaluated 4 positions per thread from 128-bit shared loads, the second pass likewise for the thread's 4
// pixels; every quantity keeps the upstream operation sequence, channel sums accumulate in upstream order.
#define DS_ST_THREADS 320
template <int MODE>
__global__ void __launch_bounds__(DS_ST_THREADS, 3) k_ds_stats(const float* __restrict__ img, size_t R,
                                                                const int* __restrict__ ridx, int w, int h, size_t n,
                                                                float* __restrict__ refstat, float* __restrict__ map,
                                                                double* __restrict__ partial) {
    constexpr int NQ = MODE == 0 ? 2 : 3;
    constexpr int NIMG = MODE == 0 ? 1 : 2;
    __shared__ __align__(16) float s_in[2][NIMG][DS_IH * DS_IW];   // [channel parity][image]: next channel prefetched
    __shared__ __align__(16) float s_f[NQ][DS_FH * DS_FW];
    __shared__ double scratch[32];
    const int x0 = blockIdx.x * DS_TW, y0 = blockIdx.y * DS_TH;
    const size_t b = blockIdx.z;
    const size_t im1 = MODE == 0 ? b : (size_t)ridx[b];   // reference image
    const size_t im2 = R + b;                             // distorted image (pair mode)
    const bool vec = (w & 3) == 0;
    const int g = threadIdx.x & 15, oy = threadIdx.x >> 4;   // second pass: threads 0..255
    const bool p2 = threadIdx.x < 256;
    const int x = x0 + 4 * g, y = y0 + oy;
    const bool live = p2 && x < w && y < h;
    const int pix = live ? y * w + x : 0;
    float sm11[4], sm12[4], sm22[4], ss1[4], ss2[4], ss12[4];
    // Tiles whose halo lies inside the image (block-uniform) are staged with cp.async one channel ahead; tiles on
    // the image border need clamped coordinates and are loaded synchronously.
    const bool interior = vec && x0 - 4 >= 0 && x0 - 4 + DS_IW <= w && y0 - 2 >= 0 && y0 - 2 + DS_IH <= h;
    auto prefetch = [&](int c) {
        if (interior) {
            load_tile_async<DS_IW / 4, DS_IH, DS_ST_THREADS>(s_in[c & 1][0], DS_IW, img + (im1 * 3 + c) * n, w, h, x0 - 4, y0 - 2, true);
            if (MODE == 1)
                load_tile_async<DS_IW / 4, DS_IH, DS_ST_THREADS>(s_in[c & 1][NIMG - 1], DS_IW, img + (im2 * 3 + c) * n, w, h, x0 - 4,
                                                                 y0 - 2, true);
        }
        cp_async_commit();
    };
    prefetch(0);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // everyone is past the first pass of channel c-1 (it read buffer (c+1) & 1) -- see the barriers below
        if (c + 1 < 3) prefetch(c + 1);
        if (interior) {
            if (c + 1 < 3) cp_async_wait<1>(); else cp_async_wait<0>();
        } else {
            load_tile<2, DS_IW / 4, DS_IH, DS_ST_THREADS>(s_in[c & 1][0], DS_IW, img + (im1 * 3 + c) * n, w, h, x0 - 4, y0 - 2, vec);
            if (MODE == 1)
                load_tile<2, DS_IW / 4, DS_IH, DS_ST_THREADS>(s_in[c & 1][NIMG - 1], DS_IW, img + (im2 * 3 + c) * n, w, h, x0 - 4,
                                                              y0 - 2, vec);
        }
        __syncthreads();   // tiles of channel c visible; previous channel's second pass is done with s_f
        if (threadIdx.x < DS_FH * DS_FG) {
            const int ry = threadIdx.x / DS_FG, q = threadIdx.x - ry * DS_FG;
            float u[3][8], t[3][8];
#pragma unroll
            for (int r = 0; r < 3; r++) ds_ld8(s_in[c & 1][0] + (ry + r) * DS_IW + 4 * q, u[r]);
            float* o = &s_f[0][ry * DS_FW + 4 * q];
            if (MODE == 0) {
                *reinterpret_cast<float4*>(o) = ds_k9x4(u[0], u[1], u[2]);
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int i = 1; i < 7; i++) t[r][i] = u[r][i] * u[r][i];
                *reinterpret_cast<float4*>(o + DS_FH * DS_FW) = ds_k9x4(t[0], t[1], t[2]);
            } else {
                float v[3][8];
#pragma unroll
                for (int r = 0; r < 3; r++) ds_ld8(s_in[c & 1][NIMG - 1] + (ry + r) * DS_IW + 4 * q, v[r]);
                *reinterpret_cast<float4*>(o) = ds_k9x4(v[0], v[1], v[2]);
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int i = 1; i < 7; i++) t[r][i] = v[r][i] * v[r][i];
                *reinterpret_cast<float4*>(o + DS_FH * DS_FW) = ds_k9x4(t[0], t[1], t[2]);
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int i = 1; i < 7; i++) t[r][i] = u[r][i] * v[r][i];
                *reinterpret_cast<float4*>(o + 2 * DS_FH * DS_FW) = ds_k9x4(t[0], t[1], t[2]);
            }
        }
        __syncthreads();
        ds_fix_border(&s_f[0][0], NQ, DS_FH * DS_FW, w, h, x0, y0);
        __syncthreads();
        if (p2) {
            float q[NQ][4];
#pragma unroll
            for (int f = 0; f < NQ; f++) {
                float r0[8], r1[8], r2[8];
                const float* base = s_f[f] + oy * DS_FW + 4 * g;
                ds_ld8(base, r0); ds_ld8(base + DS_FW, r1); ds_ld8(base + 2 * DS_FW, r2);
                const float4 o = ds_k9x4(r0, r1, r2);
                q[f][0] = o.x; q[f][1] = o.y; q[f][2] = o.z; q[f][3] = o.w;
            }
            if (MODE == 0) {
                if (live) {
                    float* d0 = refstat + ((b * 3 + c) * 2) * n + pix;
                    if (vec) {
                        *reinterpret_cast<float4*>(d0) = make_float4(q[0][0], q[0][1], q[0][2], q[0][3]);
                        *reinterpret_cast<float4*>(d0 + n) = make_float4(q[1][0], q[1][1], q[1][2], q[1][3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (x + k < w) { d0[k] = q[0][k]; d0[n + k] = q[1][k]; }
                    }
                }
            } else {
                float r_mu[4] = {0, 0, 0, 0}, r_e[4] = {0, 0, 0, 0};
                if (live) {
                    const float* s0 = refstat + ((im1 * 3 + c) * 2) * n + pix;
                    if (vec) {
                        const float4 a = *reinterpret_cast<const float4*>(s0), e = *reinterpret_cast<const float4*>(s0 + n);
                        r_mu[0] = a.x; r_mu[1] = a.y; r_mu[2] = a.z; r_mu[3] = a.w;
                        r_e[0] = e.x; r_e[1] = e.y; r_e[2] = e.z; r_e[3] = e.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (x + k < w) { r_mu[k] = s0[k]; r_e[k] = s0[n + k]; }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float mu1 = r_mu[k], mu2 = q[0][k];
                    const float m11 = mu1 * mu1, m12 = mu1 * mu2, m22 = mu2 * mu2;
                    const float s1 = r_e[k] - m11, s2 = q[1][k] - m22, s12 = q[NQ - 1][k] - m12;
                    if (c == 0) { sm11[k] = m11; sm12[k] = m12; sm22[k] = m22; ss1[k] = s1; ss2[k] = s2; ss12[k] = s12; }
                    else { sm11[k] += m11; sm12[k] += m12; sm22[k] += m22; ss1[k] += s1; ss2[k] += s2; ss12[k] += s12; }
                }
            }
        }
    }
    if (MODE == 0) return;
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f, third = 1.0f / 3.0f;
    double acc = 0.0;
    if (live) {
        float vv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float mu1_sq = sm11[k] * third, mu2_sq = sm22[k] * third, mu1_mu2 = sm12[k] * third;
            const float sigma1_sq = ss1[k] * third, sigma2_sq = ss2[k] * third, sigma12 = ss12[k] * third;
            vv[k] = (__fmaf_rn(2.0f, mu1_mu2, c1) * __fmaf_rn(2.0f, sigma12, c2)) /
                    (((mu1_sq + mu2_sq) + c1) * ((sigma1_sq + sigma2_sq) + c2));
        }
        float* d = map + b * n + pix;
        if (vec) {
            *reinterpret_cast<float4*>(d) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            acc = (((double)vv[0] + (double)vv[1]) + (double)vv[2]) + (double)vv[3];
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x + k < w) { d[k] = vv[k]; acc += (double)vv[k]; }
        }
    }
    double a1[1] = {acc};
    block_sum<1>(a1, scratch);
    if (threadIdx.x == 0) partial[(b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = a1[0];
}

// one thread per pair: sum partials in fixed order; out[b][scale][0] = sum; avg[b] = max(mean,0)^(0.5^scale)
__global__ void k_ds_mean(const double* __restrict__ partial, int ntiles, size_t B, size_t n, int scale, double* __restrict__ out,
                          double* __restrict__ avg) {
    size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int i = 0; i < ntiles; i++) s += partial[b * ntiles + i];
    out[(b * 5 + scale) * 2 + 0] = s;
    double mean = s / (double)n;
    if (!(mean > 0.0)) mean = 0.0;
    avg[b] = pow(mean, pow(0.5, (double)scale));
}

#define DS_MAD_BLOCKS 64
// grid (DS_MAD_BLOCKS, B)
__global__ void __launch_bounds__(256) k_ds_mad(const float* __restrict__ map, size_t n, const double* __restrict__ avg,
                                                 double* __restrict__ partial) {
    __shared__ double scratch[32];
    const size_t b = blockIdx.y;
    const double a = avg[b];
    const float* m = map + b * n;
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += fabs(a - (double)m[i]);
    double a1[1] = {acc};
    block_sum<1>(a1, scratch);
    if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = a1[0];
}
__global__ void k_ds_mad_reduce(const double* __restrict__ partial, int nblk, size_t B, int scale, double* __restrict__ out) {
    size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int i = 0; i < nblk; i++) s += partial[b * nblk + i];
    out[(b * 5 + scale) * 2 + 1] = s;
}

int dssim_num_scales(size_t w, size_t h, size_t* ws, size_t* hs) {
    int ns = 0;
    size_t cw = w, ch = h;
    for (int s = 0; s < 5; s++) {
        if (s > 0) {
            size_t hw = cw / 2, hh = ch / 2;
            if (hw < 4 || hh < 4) break;
            cw = hw; ch = hh;
        }
        if (ws) ws[ns] = cw;
        if (hs) hs[ns] = ch;
        ns++;
    }
    return ns;
}

size_t dssim_workspace_per_pair(size_t w, size_t h) {
    size_t n = w * h;
    size_t tiles = (size_t)cdiv(w, DS_TW) * cdiv(h, DS_TH);
    // img 6n, chroma 4n, map n, reference statistics 6n, next-scale rgb(a) ping-pong 2*2*4*(n/4)
    return (6 * n + 4 * n + n + 6 * n + 4 * n) * 4 + (tiles + DS_MAD_BLOCKS + 4) * 8 + 8192;
}

int dssim_run(Context& c, const float* lin_in, const float* alpha_in, size_t R, const int* ridx, size_t B, size_t w, size_t h,
              double* d_out, float* dbg_map0) {
    size_t ws[5], hs[5];
    const int ns = dssim_num_scales(w, h, ws, hs);
    size_t mark = c.arena.mark();
    const size_t n0 = w * h, NI = R + B;
    if (NI > 65535) throw CudaError("dssim sub-batch too large for one launch");
    float* img = c.arena.alloc<float>(NI * 3 * n0);
    float* chroma = c.arena.alloc<float>(NI * 2 * n0);
    float* map = c.arena.alloc<float>(B * n0);
    float* refstat = c.arena.alloc<float>(R * 6 * n0);   // per reference: [3 channels][mu1, blur2(ch1^2)]
    const size_t tiles0 = (size_t)cdiv(w, DS_TW) * cdiv(h, DS_TH);
    double* partial = c.arena.alloc<double>(B * std::max<size_t>(tiles0, DS_MAD_BLOCKS));
    double* avg = c.arena.alloc<double>(B);
    const bool has_alpha = alpha_in != nullptr;
    // next-scale planes, ping-pong: [NI][3][n/4] (+ alpha [NI][n/4])
    float* nl[2] = {nullptr, nullptr};
    float* nal[2] = {nullptr, nullptr};
    if (ns > 1)
        for (int i = 0; i < 2; i++) {
            nl[i] = c.arena.alloc<float>(NI * 3 * ws[1] * hs[1]);
            if (has_alpha) nal[i] = c.arena.alloc<float>(NI * ws[1] * hs[1]);
        }

    const float* l = lin_in;
    const float* al = alpha_in;
    const unsigned wave = (unsigned)c.sm_count * 8;
    for (int s = 0; s < ns; s++) {
        const size_t cw = ws[s], ch = hs[s], n = cw * ch;
        const bool has_next = s + 1 < ns;
        const size_t nw = has_next ? ws[s + 1] : 0, nh = has_next ? hs[s + 1] : 0, nn = nw * nh;
        const unsigned tx = cdiv(cw, DS_TW), ty = cdiv(ch, DS_TH);
        {
            float* dst = has_next ? nl[(s + 1) & 1] : nullptr;
            dim3 grid(cdiv((cw + 1) / 2, 64), cdiv((ch + 1) / 2, 4), (unsigned)NI);
            CE_LAUNCH(c, "k_ds_lab", (double)NI * (n * (has_alpha ? 28 : 24) + nn * 12),
                      k_ds_lab<<<grid, 256, 0, c.stream>>>(l, has_alpha ? al : nullptr, (int)cw, (int)ch, n, img, chroma,
                                                          has_next ? 1 : 0, (int)nw, (int)nh, nn, dst));
            if (has_next && has_alpha) {
                float* adst = nal[(s + 1) & 1];
                size_t atotal = NI * nn;
                CE_LAUNCH(c, "k_ds_down", (double)atotal * 20,
                          k_ds_down<<<std::min<unsigned>(cdiv(atotal, 256), wave * 4), 256, 0, c.stream>>>(
                              al, (int)cw, n, (int)nw, (int)nh, nn, atotal, adst));
                al = adst;
            }
            if (has_next) l = dst;
        }
        {
            dim3 grid(tx, ty, (unsigned)NI);
            CE_LAUNCH(c, "k_ds_blur2", (double)NI * n * 16, k_ds_blur2<<<grid, 256, 0, c.stream>>>(chroma, (int)cw, (int)ch, n, img));
        }
        const int ntiles = (int)(tx * ty);
        {
            dim3 grid(tx, ty, (unsigned)R);
            CE_LAUNCH(c, "k_ds_stats<ref>", (double)R * n * 36,
                      k_ds_stats<0><<<grid, DS_ST_THREADS, 0, c.stream>>>(img, R, nullptr, (int)cw, (int)ch, n, refstat, nullptr, nullptr));
        }
        for (size_t b0 = 0; b0 < B; b0 += 32768) {
            unsigned nb = (unsigned)std::min<size_t>(32768, B - b0);
            dim3 grid(tx, ty, nb);
            CE_LAUNCH(c, "k_ds_stats<pair>", (double)nb * n * 52,
                      k_ds_stats<1><<<grid, DS_ST_THREADS, 0, c.stream>>>(img, R + b0, ridx + b0, (int)cw, (int)ch, n, refstat,
                                                                           map + b0 * n, partial + b0 * ntiles));
        }
        CE_LAUNCH(c, "k_ds_mean", (double)B * (ntiles + 2) * 8,
                  k_ds_mean<<<cdiv(B, 128), 128, 0, c.stream>>>(partial, ntiles, B, n, s, d_out, avg));
        if (dbg_map0 && s == 0) CE_CUDA(cudaMemcpyAsync(dbg_map0, map, n * 4, cudaMemcpyDeviceToDevice, c.stream));
        for (size_t b0 = 0; b0 < B; b0 += 32768) {
            unsigned nb = (unsigned)std::min<size_t>(32768, B - b0);
            dim3 grid(DS_MAD_BLOCKS, nb);
            CE_LAUNCH(c, "k_ds_mad", (double)nb * n * 4,
                      k_ds_mad<<<grid, 256, 0, c.stream>>>(map + b0 * n, n, avg + b0, partial + b0 * DS_MAD_BLOCKS));
        }
        CE_LAUNCH(c, "k_ds_mad_reduce", (double)B * (DS_MAD_BLOCKS + 1) * 8,
                  k_ds_mad_reduce<<<cdiv(B, 128), 128, 0, c.stream>>>(partial, DS_MAD_BLOCKS, B, s, d_out));
        CE_CUDA(cudaGetLastError());
    }
    c.arena.release(mark);
    return ns;
}

}  // namespace ce
