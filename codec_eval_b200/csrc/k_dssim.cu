// k_dssim.cu -- DSSIM (dssim-core 3.4.0; reference call site src/metrics/dssim.rs:52-68)
// for a batch of B pairs.
//
// Per scale (<= 5 scales, floor-halving with crop on LINEAR rgb(a)):
//   k_ds_lab     : pointwise, thread = 2x2 block: linear -> L (final), a/b (to be pre-blurred), and the
//                  2x2 average of the linear planes for the next scale                 [issue: 6 IEEE divisions / px]
//   k_ds_blur2   : chroma pre-blur: the 3x3 kernel applied twice (clamp-replicate per pass) through a
//                  shared-memory tile, 4 positions per thread from 128-bit shared loads [HBM / latency]
//                  (a streaming-warp version like k_ds_stream measured 7-25 % slower: 256-B row segments)
//   k_ds_stream  : per channel the five double-3x3 blurs {ch1,ch2,ch1^2,ch2^2,ch1*ch2}: a warp streams down a
//                  60-column strip with both 3x3 passes as register windows (shared products and row sums);
//                  channel-averaged SSIM map written once + fp64 partial of its sum    [FP32-issue bound]
//   k_ds_mean    : fixed-order reduce -> sum(map), avg = max(mean,0)^(0.5^scale)
//   k_ds_mad     : sum |avg - map_i| in fp64 -> block partials; k_ds_mad_reduce fixes the order
#include <cuda.h>

#include "ce_common.cuh"
#include "ce_internal.h"

#include <type_traits>

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

// ------------------------------------------------------------------ chroma pre-blur (3x3 kernel applied twice)
// grid (tiles_x, tiles_y, NI): blockIdx.z = image; chroma [NI][2][n] -> img[(z*3) + 1 + {0,1}].
// Both planes of a 64x16 tile (+ halo 2, clamped) are staged; each pass evaluates 4 positions per thread
// from 128-bit shared loads.
// TMA: both plane tiles arrive as one bulk tensor copy (box DS_IW x DS_IH x 2, zeros outside the image); tiles on the
// image border then overwrite their out-of-image entries with the nearest in-image value (clamp-replicate).
template <bool TMA>
__global__ void __launch_bounds__(256) k_ds_blur2(const float* __restrict__ chroma, int w, int h, size_t n,
                                                   float* __restrict__ img, const __grid_constant__ CUtensorMap map) {
    __shared__ __align__(128) float s_ab[2][DS_IH * DS_IW];
    __shared__ __align__(16) float s_f[2][DS_FH * DS_FW];
    __shared__ __align__(8) unsigned long long s_bar;
    const int x0 = blockIdx.x * DS_TW, y0 = blockIdx.y * DS_TH;
    const size_t z = blockIdx.z;
    const bool vec = (w & 3) == 0;
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&s_bar, 2 * DS_IH * DS_IW * 4);
            tma_load_3d(&s_ab[0][0], &map, x0 - 4, y0 - 2, (int)(z * 2), &s_bar);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
        const bool border = x0 - 4 < 0 || x0 - 4 + DS_IW > w || y0 - 2 < 0 || y0 - 2 + DS_IH > h;   // block-uniform
        if (border) {
            for (int e = threadIdx.x; e < DS_IH * DS_IW; e += 256) {
                const int r = e / DS_IW, cx = e - r * DS_IW;
                const int x = x0 - 4 + cx, y = y0 - 2 + r;
                if (x < 0 || x >= w || y < 0 || y >= h) {
                    const int mx = min(max(x, 0), w - 1), my = min(max(y, 0), h - 1);
                    const int tr = my - (y0 - 2), tc = mx - (x0 - 4);
                    const bool in_tile = tr >= 0 && tr < DS_IH && tc >= 0 && tc < DS_IW;
#pragma unroll
                    for (int pl = 0; pl < 2; pl++)
                        s_ab[pl][e] = in_tile ? s_ab[pl][tr * DS_IW + tc] : chroma[(z * 2 + pl) * n + (size_t)my * w + mx];
                }
            }
            __syncthreads();
        }
    } else {
#pragma unroll
        for (int pl = 0; pl < 2; pl++) load_tile<2, DS_IW / 4, DS_IH, 256>(s_ab[pl], DS_IW, chroma + (z * 2 + pl) * n, w, h, x0 - 4, y0 - 2, vec);
        __syncthreads();
    }
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

// ------------------------------------------------------------------ statistics + SSIM map (streaming)
// The five double-3x3 blurs {ch1, ch2, ch1^2, ch2^2, ch1*ch2} per channel split by who owns them:
//   k_ds_stream<0> (grid.z = distinct reference): mu1 = blur2(ch1), e11 = blur2(ch1^2) -> refstat [R][3][2][n],
//                  once per reference however many distortions it is compared with;
//   k_ds_stream<1> (grid.z = pair): blur2 of {ch2, ch2^2, ch1*ch2}, then the channel-averaged SSIM map
//                  (written once) + fp64 partial of its sum.  partial: [B][strips] doubles.
//
// One warp owns a strip of 60 output columns (lane = 2 adjacent columns; lanes 0 and 31 are halo lanes) and
// walks down DSS rows, all three channels in registers, no shared memory and no block barrier.  The 3x3 pass
//     out = ((v00*k0 + v01*k1) + v02*k0) + ((v10*k1 + v11*k4) + v12*k1) + ((v20*k0 + v21*k1) + v22*k0)
// is evaluated with the upstream operation sequence, but every product v*k is formed once per input element and
// the row sums  A = (p0[x-1] + p1[x]) + p0[x+1]  (top and bottom rows use the same weights) and
// B = (p1[x-1] + p4[x]) + p1[x+1]  once per row:  out(r) = (A(r-1) + B(r)) + A(r+1).  That is 11 instead of 17
// fp32 instructions per 3x3 evaluation with bit-identical results.  Both passes run as two chained 3-row
// windows (6 registers per column and quantity); the clamp-replicate rule between the passes (a first-pass
// value outside the image is the nearest inside first-pass value) is applied explicitly: horizontally by a
// shuffle from the lane holding column 0 / w-1, vertically by pushing the first / last row twice.
#define DSS_OUT 60
#define DSS_PITCH 68     // floats per staged plane row: 2 pad + 64 window columns + 2 pad
#define DSS_SLOTS 4      // ring depth (three rows in flight ahead of the one being consumed)
// slot layout (floats): the three plane groups start on 128-byte boundaries so that each can be the destination of one
// TMA box (68 columns x 1 row x planes):  [ch1 x3 | pad][ch2 x3 | pad][reference statistics x6 | pad]
#define DSS_OFF_P2 224   // 3 * 68 = 204 floats -> 224 (896 B)
#define DSS_OFF_ST 448   // (1792 B)

// two-row window of one quantity: a[par] = A(r-2), a[par ^ 1] = A(r-1) for the tick parity par, b = B(r-1).
// Pushing writes the new A over the oldest one, so with the tick loop unrolled by two no register moves remain.
struct DsWin {
    f32x2 a[2], b;   // each holds the lane's two columns
};
// row sums of the two columns of this lane from the four values at columns c0-1 .. c0+2
// The products are scalar FMULs, the sums packed FADD2s (one issue slot for both columns); a packed multiply
// feeding a packed add would be contracted into FFMA2 by ptxas and round differently from dssim-core's chain.
CE_DEVINL void ds_rowsums(float vm, float v0, float v1, float v2, f32x2& A, f32x2& B) {
    const float k0 = 0.095332f, k1 = 0.118095f, k4 = 0.146293f;
    const float p0m = vm * k0, p00 = v0 * k0, p01 = v1 * k0, p02 = v2 * k0;
    const float p1m = vm * k1, p10 = v0 * k1, p11 = v1 * k1, p12 = v2 * k1;
    const float p40 = v0 * k4, p41 = v1 * k4;
    A = add2(add2(pk2(p0m, p00), pk2(p10, p11)), pk2(p01, p02));
    B = add2(add2(pk2(p1m, p10), pk2(p40, p41)), pk2(p11, p12));
}
CE_DEVINL void cp_async8(float* smem, const float* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}

template <int MODE>
struct DsStream {
    static constexpr int NQ = MODE == 0 ? 2 : 3;
    static constexpr int NPL = MODE == 0 ? 3 : 12;   // staged planes per tick: inputs (+ the reference statistics)
    static constexpr int SLOT = MODE == 0 ? 224 : 864;   // floats per slot (896 B / 3456 B: multiples of 128 B)
    static constexpr unsigned FULL = 0xffffffffu;

    DsWin s1[3][NQ], s2[3][NQ];
    float sm11[2], sm12[2], sm22[2], ss1[2], ss2[2], ss12[2];
    double acc;
    int lane, c0, w;
    bool left_edge, right_edge, st0, st1, v2ok;
    int rsrc, rel;

    // one channel's terms of the output row: o = second-pass values {mu2, e22, e12}, rs = staged {mu1, e11}
    CE_DEVINL void terms(int c, const float (&o)[NQ][2], const float (&rmu)[2], const float (&re)[2]) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float mu1 = rmu[e], mu2 = o[0][e];
            const float m11 = mu1 * mu1, m12 = mu1 * mu2, m22 = mu2 * mu2;
            const float t1 = re[e] - m11, t2 = o[1][e] - m22, t12 = o[NQ - 1][e] - m12;
            if (c == 0) { sm11[e] = m11; sm12[e] = m12; sm22[e] = m22; ss1[e] = t1; ss2[e] = t2; ss12[e] = t12; }
            else { sm11[e] += m11; sm12[e] += m12; sm22[e] += m22; ss1[e] += t1; ss2[e] += t2; ss12[e] += t12; }
        }
    }
    CE_DEVINL void store_ref(float* __restrict__ d0, size_t n, const float (&o)[NQ][2]) {   // d0: plane mu1 at the row start
        if (v2ok && st0) {
            *reinterpret_cast<float2*>(d0 + c0) = make_float2(o[0][0], o[0][1]);
            *reinterpret_cast<float2*>(d0 + n + c0) = make_float2(o[1][0], o[1][1]);
        } else {
            if (st0 && c0 >= 0) { d0[c0] = o[0][0]; d0[n + c0] = o[1][0]; }
            if (st1 && c0 + 1 >= 0) { d0[c0 + 1] = o[0][1]; d0[n + c0 + 1] = o[1][1]; }
        }
    }
    CE_DEVINL void finish_row(float* __restrict__ d, bool emit) {   // d: map row start
        const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f, third = 1.0f / 3.0f;
        float vv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float mu1_sq = sm11[e] * third, mu2_sq = sm22[e] * third, mu1_mu2 = sm12[e] * third;
            const float sigma1_sq = ss1[e] * third, sigma2_sq = ss2[e] * third, sigma12 = ss12[e] * third;
            vv[e] = div_rn_normal(__fmaf_rn(2.0f, mu1_mu2, c1) * __fmaf_rn(2.0f, sigma12, c2),
                                  ((mu1_sq + mu2_sq) + c1) * ((sigma1_sq + sigma2_sq) + c2));
        }
        if (!emit) return;
        if (v2ok && st0) {
            *reinterpret_cast<float2*>(d + c0) = make_float2(vv[0], vv[1]);
            acc += (double)vv[0] + (double)vv[1];
        } else {
            if (st0 && c0 >= 0) { d[c0] = vv[0]; acc += (double)vv[0]; }
            if (st1 && c0 + 1 >= 0) { d[c0 + 1] = vv[1]; acc += (double)vv[1]; }
        }
    }

    // One tick: consume the staged input row (slot), advance the first pass, push the finished first-pass row into
    // the second pass and form the output row.  Straight-line code: during the first ticks of a strip the windows
    // are still filling and the values are meaningless; only the stores and the pooled sum are predicated (emit).
    // TMA-staged edge strips: columns outside the image arrived as zeros; give them the value of the nearest inside
    // column (clamp-replicate), in registers.  Left edge (window starts at column -4): lane 0 holds columns -3 .. 0,
    // lane 1 columns -1 .. 2.  Right edge: column w-1 sits in lane rsrc, element rel of its pair.
    CE_DEVINL void patch_edges(float& vm, float& v0, float& v1, float& v2) const {
        if (left_edge) {
            if (lane == 0) { vm = v2; v0 = v2; v1 = v2; }
            else if (lane == 1) vm = v0;
        }
        if (right_edge) {
            const float ve = __shfl_sync(FULL, rel ? v1 : v0, rsrc);
            const int e = w - 1;
            if (c0 - 1 > e) vm = ve;
            if (c0 > e) v0 = ve;
            if (c0 + 1 > e) v1 = ve;
            if (c0 + 2 > e) v2 = ve;
        }
    }

    template <int PAR, bool PATCH>
    CE_DEVINL void tick(const float* __restrict__ slot, bool emit, float* __restrict__ ref_row, size_t n,
                        float* __restrict__ map_row) {
        const float* sl = slot + 2 * lane + 1;   // column c0 - 1 of plane 0
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float in[NQ][4];
            {
                const float* pa = sl + c * DSS_PITCH;
                float am = pa[0], a3 = pa[3];
                float2 a12 = *reinterpret_cast<const float2*>(pa + 1);
                if (PATCH) patch_edges(am, a12.x, a12.y, a3);
                if (MODE == 0) {
                    in[0][0] = am; in[0][1] = a12.x; in[0][2] = a12.y; in[0][3] = a3;
                    in[1][0] = am * am; in[1][1] = a12.x * a12.x; in[1][2] = a12.y * a12.y; in[1][3] = a3 * a3;
                } else {
                    const float* pb = sl + DSS_OFF_P2 + c * DSS_PITCH;
                    float bm = pb[0], b3 = pb[3];
                    float2 b12 = *reinterpret_cast<const float2*>(pb + 1);
                    if (PATCH) patch_edges(bm, b12.x, b12.y, b3);
                    in[0][0] = bm; in[0][1] = b12.x; in[0][2] = b12.y; in[0][3] = b3;
                    in[1][0] = bm * bm; in[1][1] = b12.x * b12.x; in[1][2] = b12.y * b12.y; in[1][3] = b3 * b3;
                    in[NQ - 1][0] = am * bm; in[NQ - 1][1] = a12.x * b12.x; in[NQ - 1][2] = a12.y * b12.y; in[NQ - 1][3] = a3 * b3;
                }
            }
            float f[NQ][2];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                f32x2 A, B;
                ds_rowsums(in[q][0], in[q][1], in[q][2], in[q][3], A, B);
                DsWin& s = s1[c][q];
                unpk2(add2(add2(s.a[PAR], s.b), A), f[q][0], f[q][1]);
                s.a[PAR] = A;
                s.b = B;
            }
            // clamp-replicate of the first pass in x: columns < 0 take column 0, columns > w-1 take column w-1
            if (left_edge) {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float t = __shfl_sync(FULL, f[q][0], 1);
                    if (lane == 0) { f[q][0] = t; f[q][1] = t; }
                }
            }
            if (right_edge) {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float t = __shfl_sync(FULL, rel ? f[q][1] : f[q][0], rsrc);
                    if (c0 > w - 1) f[q][0] = t;
                    if (c0 + 1 > w - 1) f[q][1] = t;
                }
            }
            float o[NQ][2];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const float fm = __shfl_up_sync(FULL, f[q][1], 1), f2 = __shfl_down_sync(FULL, f[q][0], 1);
                f32x2 A, B;
                ds_rowsums(fm, f[q][0], f[q][1], f2, A, B);
                DsWin& s = s2[c][q];
                unpk2(add2(add2(s.a[PAR], s.b), A), o[q][0], o[q][1]);
                s.a[PAR] = A;
                s.b = B;
            }
            if (MODE == 0) {
                if (emit) store_ref(ref_row + (size_t)c * 2 * n, n, o);
            } else {
                const float2 mu = *reinterpret_cast<const float2*>(sl + DSS_OFF_ST + (2 * c) * DSS_PITCH + 1);
                const float2 ee = *reinterpret_cast<const float2*>(sl + DSS_OFF_ST + (2 * c + 1) * DSS_PITCH + 1);
                const float rmu[2] = {mu.x, mu.y}, re[2] = {ee.x, ee.y};
                terms(c, o, rmu, re);
            }
        }
        if (MODE == 1) finish_row(map_row, emit);
    }
    // row 0 also stands for row -1: the second-pass window holds its row sums twice
    CE_DEVINL void dup_top() {
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int q = 0; q < NQ; q++) s2[c][q].a[1] = s2[c][q].a[0];
    }
};

// grid (column strips, row strips, units); block = one warp; unit = distinct reference (MODE 0) or pair (MODE 1)
// TMA (strips whose 68-column window lies inside the image, widths TMA can describe): lane 0 arms the slot's mbarrier
// and issues one bulk tensor copy per plane group of the row; the per-lane cp.async path (12 copies and their 64-bit
// address arithmetic per lane and row, ~10 % of the instruction stream) remains for the two edge strips, whose columns
// are clamped, and for odd widths.
struct DsMaps {
    CUtensorMap img;   // [NI*3 planes][h][w], box (68, 1, 3)
    CUtensorMap st;    // reference statistics [R*6 planes][h][w], box (68, 1, 6)
};
// Two launches per scale when TMA can describe the planes: the interior strips 1 .. s_hi (EDGE = false: blockIdx.x + 1)
// and the edge strips 0 and s_hi+1 .. sx-1 (EDGE = true: blockIdx.x == 0 -> strip 0, else s_hi + blockIdx.x), so that
// the interior variant carries none of the clamping; the edge variant fetches the same boxes (columns outside the image
// arrive as zeros) and patches the clamped columns in registers.  TMA = false: one launch over all strips with per-lane
// cp.async copies (odd widths, narrow scales).
template <int MODE, bool TMA, bool EDGE>
__global__ void __launch_bounds__(32, 12) k_ds_stream(const float* __restrict__ img, size_t R, const int* __restrict__ ridx, int w,
                                                   int h, size_t n, int rows_per_strip, float* __restrict__ refstat,
                                                   float* __restrict__ map, double* __restrict__ partial, int sx_total, int s_hi,
                                                   const __grid_constant__ DsMaps maps) {
    typedef DsStream<MODE> S;
    constexpr int NQ = S::NQ;
    __shared__ __align__(128) float ring[DSS_SLOTS * S::SLOT];
    __shared__ __align__(8) unsigned long long s_bar[TMA ? DSS_SLOTS : 1];
    S st;
    const int lane = threadIdx.x;
    const int strip = (TMA && !EDGE) ? 1 + (int)blockIdx.x : (blockIdx.x == 0 ? 0 : s_hi + (int)blockIdx.x);
    const int xw = strip * DSS_OUT - 2;   // first column of the warp's 64-column window
    const int c0 = xw + 2 * lane;                   // this lane's columns c0, c0 + 1 (may lie outside the image)
    const int ys = (int)blockIdx.y * rows_per_strip, ye = min(ys + rows_per_strip, h);
    const size_t b = blockIdx.z;
    const size_t im1 = MODE == 0 ? b : (size_t)ridx[b];   // reference image
    const size_t im2 = R + b;                             // distorted image (pair mode)
    const int cc0 = min(max(c0, 0), w - 1), cc1 = min(max(c0 + 1, 0), w - 1);
    constexpr bool INTERIOR = TMA && !EDGE;
    const bool allvec = TMA || ((w & 1) == 0 && xw >= 0 && xw + 63 < w);   // warp-uniform: every lane's pair is inside and 8-B aligned
    st.lane = lane; st.c0 = c0; st.w = w;
    st.v2ok = INTERIOR || ((w & 1) == 0 && c0 >= 0 && c0 + 1 < w);
    st.left_edge = !INTERIOR && xw < 0; st.right_edge = !INTERIOR && xw + 63 > w - 1;
    st.rsrc = (w - 1 - xw) >> 1; st.rel = (w - 1 - xw) & 1;   // lane / element holding column w-1
    st.st0 = lane >= 1 && lane <= 30 && c0 < w; st.st1 = lane >= 1 && lane <= 30 && c0 + 1 < w;
    st.acc = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int q = 0; q < NQ; q++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                st.s1[c][q].a[k] = 0ull; st.s2[c][q].a[k] = 0ull;
                st.s1[c][q].b = 0ull; st.s2[c][q].b = 0ull;
            }

    // first-pass rows needed: r_lo .. r_hi; input rows r_lo-1 .. r_hi+1 (clamped); tick k consumes input row r_lo-1+k,
    // finishes first-pass row r_lo+k-2 and (k >= kfirst) output row ys + k - kfirst
    const int r_lo = max(ys - 1, 0), r_hi = min(ye, h - 1);
    const int nin = r_hi - r_lo + 3;
    const int kfirst = ys == 0 ? 3 : 4;
    const float* p1 = img + im1 * 3 * n;
    const float* p2 = img + im2 * 3 * n;                 // unused in ref mode
    const float* prs = refstat + im1 * 6 * n;            // pair mode: [3][2][n] statistics of the reference
    float* my = ring + 2 + 2 * lane;                     // this lane's two columns of plane 0, slot 0
    // TMA strips: the whole 68-column window (pads included) lies inside the image, nothing to clamp
    constexpr bool tma = TMA;
    if (tma) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < DSS_SLOTS; i++) mbar_init(&s_bar[i], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    auto issue = [&](int k) {
        if (tma) {
            if (lane == 0 && k < nin) {
                float* d = ring + (k & (DSS_SLOTS - 1)) * S::SLOT;
                unsigned long long* bar = &s_bar[k & (DSS_SLOTS - 1)];
                const int row = min(max(r_lo - 1 + k, 0), h - 1);
                const int y = ys + k - kfirst;
                const bool stats = MODE == 1 && k >= kfirst && y < h;
                mbar_expect_tx(bar, (MODE == 1 ? 2 : 1) * 3 * DSS_PITCH * 4 + (stats ? 6 * DSS_PITCH * 4 : 0));
                tma_load_3d(d, &maps.img, xw - 2, row, (int)(im1 * 3), bar);
                if (MODE == 1) {
                    tma_load_3d(d + DSS_OFF_P2, &maps.img, xw - 2, row, (int)(im2 * 3), bar);
                    if (stats) tma_load_3d(d + DSS_OFF_ST, &maps.st, xw - 2, min(max(y, 0), h - 1), (int)(im1 * 6), bar);
                }
            }
            return;
        }
        if (k < nin) {
            float* d = my + (k & (DSS_SLOTS - 1)) * S::SLOT;
            const size_t ro = (size_t)min(max(r_lo - 1 + k, 0), h - 1) * w;
            const int y = ys + k - kfirst;
            const bool stats = MODE == 1 && k >= kfirst && y < h;
            const size_t so = (size_t)min(max(y, 0), h - 1) * w;
            if (allvec) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    cp_async8(d + c * DSS_PITCH, p1 + c * n + ro + c0);
                    if (MODE == 1) cp_async8(d + DSS_OFF_P2 + c * DSS_PITCH, p2 + c * n + ro + c0);
                }
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 6; j++) cp_async8(d + DSS_OFF_ST + j * DSS_PITCH, prs + j * n + so + c0);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    cp_async4(d + c * DSS_PITCH, p1 + c * n + ro + cc0, true);
                    cp_async4(d + c * DSS_PITCH + 1, p1 + c * n + ro + cc1, true);
                    if (MODE == 1) {
                        cp_async4(d + DSS_OFF_P2 + c * DSS_PITCH, p2 + c * n + ro + cc0, true);
                        cp_async4(d + DSS_OFF_P2 + c * DSS_PITCH + 1, p2 + c * n + ro + cc1, true);
                    }
                }
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        cp_async4(d + DSS_OFF_ST + j * DSS_PITCH, prs + j * n + so + cc0, true);
                        cp_async4(d + DSS_OFF_ST + j * DSS_PITCH + 1, prs + j * n + so + cc1, true);
                    }
                }
            }
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    issue(2);
    float* ref_out = MODE == 0 ? refstat + b * 6 * n : nullptr;
    float* map_out = MODE == 1 ? map + b * n : nullptr;
    // One step = wait for row kk, refill the slot freed by row kk-1, advance both passes.  The loop body is two
    // unconditional steps (parities 0 and 1) so that every window register holds the same quantity at the back edge as
    // at the loop head; an odd last row is peeled.  (With the steps guarded by `kk < nin` inside the loop ptxas merged
    // the two paths with a block of ~75 register moves per iteration: MOV was 18 % of the executed instructions.)
    auto step = [&](int kk, auto par_tag) {
        constexpr int PAR = decltype(par_tag)::value;
        if (tma) {
            mbar_wait(&s_bar[kk & (DSS_SLOTS - 1)], (unsigned)(kk / DSS_SLOTS) & 1u);
            fence_proxy_async();   // this lane's reads of slot (kk - 1) & 3 precede its refill below
        } else {
            cp_async_wait<2>();
        }
        __syncwarp();   // row kk landed for every lane; every lane is past its reads of slot (kk - 1) & 3
        issue(kk + 3);
        const int y = ys + kk - kfirst;
        const bool emit = kk >= kfirst;
        const float* slot = ring + (kk & (DSS_SLOTS - 1)) * S::SLOT;
        float* rr = MODE == 0 ? ref_out + (size_t)max(y, 0) * w : nullptr;
        float* mr = MODE == 1 ? map_out + (size_t)max(y, 0) * w : nullptr;
        st.template tick<PAR, TMA && EDGE>(slot, emit, rr, n, mr);
        if (PAR == 0 && kk == 2 && ys == 0) st.dup_top();   // tick 2 (parity 0) pushed first-pass row 0
    };
    int k = 0;
    for (; k + 1 < nin; k += 2) {
        step(k, std::integral_constant<int, 0>());
        step(k + 1, std::integral_constant<int, 1>());
    }
    if (k < nin) step(k, std::integral_constant<int, 0>());
    cp_async_wait<0>();
    if (ye == h) {
        // the last first-pass row also stands for row h: one more output row, y = h-1, from the window as it is:
        // A(h-2) = a[par], B(h-1) = b, A(h) := A(h-1) = a[par ^ 1] where par is the parity after the last tick
        const int par = nin & 1;
        const int y = h - 1;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float o[NQ][2];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const DsWin& s = st.s2[c][q];
                const f32x2 a_old = par ? s.a[1] : s.a[0], a_new = par ? s.a[0] : s.a[1];
                unpk2(add2(add2(a_old, s.b), a_new), o[q][0], o[q][1]);
            }
            if (MODE == 0) st.store_ref(ref_out + (size_t)y * w + (size_t)c * 2 * n, n, o);
            else {
                float rmu[2], re[2];
                const float* q0 = prs + (size_t)(2 * c) * n + (size_t)y * w;
                rmu[0] = q0[cc0]; rmu[1] = q0[cc1]; re[0] = q0[n + cc0]; re[1] = q0[n + cc1];
                st.terms(c, o, rmu, re);
            }
        }
        if (MODE == 1) st.finish_row(map_out + (size_t)y * w, true);
    }
    if (MODE == 0) return;
    const double tot = warp_sum(st.acc);
    if (lane == 0) partial[(b * gridDim.y + blockIdx.y) * sx_total + strip] = tot;
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
    if ((n & 3) == 0) {   // 128-bit loads (the map of pair b starts at a multiple of n floats, 16-byte aligned)
        const float4* m4 = reinterpret_cast<const float4*>(m);
        const size_t n4 = n >> 2;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = m4[i];
            acc += (fabs(a - (double)v.x) + fabs(a - (double)v.y)) + (fabs(a - (double)v.z) + fabs(a - (double)v.w));
        }
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            acc += fabs(a - (double)m[i]);
    }
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
    size_t tiles = (size_t)cdiv(w, DSS_OUT) * cdiv(h, 16);
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
    const size_t strips0 = (size_t)cdiv(w, DSS_OUT) * cdiv(h, 16);   // most strips any scale can have
    double* partial = c.arena.alloc<double>(B * std::max<size_t>(strips0, DS_MAD_BLOCKS));
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
        const unsigned sx = cdiv(cw, DSS_OUT);
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
            dim3 grid(cdiv(cw, DS_TW), cdiv(ch, DS_TH), (unsigned)NI);
            CUtensorMap mc;
            memset(&mc, 0, sizeof(mc));
            if (tma_enabled(6) && tma_plane_map(&mc, chroma, cw, ch, NI * 2, DS_IW, DS_IH, 2))
                CE_LAUNCH(c, "k_ds_blur2", (double)NI * n * 16, k_ds_blur2<true><<<grid, 256, 0, c.stream>>>(chroma, (int)cw, (int)ch, n, img, mc));
            else
                CE_LAUNCH(c, "k_ds_blur2", (double)NI * n * 16, k_ds_blur2<false><<<grid, 256, 0, c.stream>>>(chroma, (int)cw, (int)ch, n, img, mc));
        }
        // row strips: 64 rows per warp (128 measured no faster: the 4 window-filling ticks per strip it saves are offset
        // by the coarser tail), fewer when the launch would not fill the machine
        int rows = 64;
        while (rows > 16 && (size_t)sx * cdiv(ch, rows) * B < (size_t)c.sm_count * 32) rows /= 2;
        const unsigned sy = cdiv(ch, rows);
        const int ntiles = (int)(sx * sy);
        DsMaps dm;
        memset(&dm, 0, sizeof(dm));
        const bool ds_tma = tma_enabled(7) && cw >= 124 && tma_plane_map(&dm.img, img, cw, ch, NI * 3, DSS_PITCH, 1, 3) &&
                            tma_plane_map(&dm.st, refstat, cw, ch, R * 6, DSS_PITCH, 1, 6);
        // strips 1 .. s_hi have their 68-column window (60 s - 4 .. 60 s + 63) inside the image; s_hi = 0: none
        const int s_hi = ds_tma ? std::min<int>((int)sx - 1, ((int)cw - 64) / DSS_OUT) : 0;
        const unsigned n_edge = sx - (unsigned)s_hi;   // strip 0 and the strips after s_hi
        {
            int rrows = 64;   // the reference launch has R, not B, images to spread over the machine
            while (rrows > 16 && (size_t)sx * cdiv(ch, rrows) * R < (size_t)c.sm_count * 32) rrows /= 2;
            const double bytes = (double)R * n * 36;
            if (s_hi > 0)
                CE_LAUNCH(c, "k_ds_stats<ref>", bytes * s_hi / sx,
                          k_ds_stream<0, true, false><<<dim3((unsigned)s_hi, cdiv(ch, rrows), (unsigned)R), 32, 0, c.stream>>>(
                              img, R, nullptr, (int)cw, (int)ch, n, rrows, refstat, nullptr, nullptr, (int)sx, s_hi, dm));
            if (s_hi > 0)
                CE_LAUNCH(c, "k_ds_stats<ref>", bytes * n_edge / sx,
                          k_ds_stream<0, true, true><<<dim3(n_edge, cdiv(ch, rrows), (unsigned)R), 32, 0, c.stream>>>(
                              img, R, nullptr, (int)cw, (int)ch, n, rrows, refstat, nullptr, nullptr, (int)sx, s_hi, dm));
            else
                CE_LAUNCH(c, "k_ds_stats<ref>", bytes,
                          k_ds_stream<0, false, true><<<dim3(sx, cdiv(ch, rrows), (unsigned)R), 32, 0, c.stream>>>(
                              img, R, nullptr, (int)cw, (int)ch, n, rrows, refstat, nullptr, nullptr, (int)sx, 0, dm));
        }
        for (size_t b0 = 0; b0 < B; b0 += 32768) {
            unsigned nb = (unsigned)std::min<size_t>(32768, B - b0);
            // per pair: ch2 (12 B) + map out (4 B); per distinct reference: ch1 (12 B) + its statistics (24 B)
            const double bytes = ((double)nb * 16 + (double)std::min<size_t>(R, nb) * 36) * n, bytes_pp = (double)nb * n * 52;
            if (s_hi > 0) {
                CE_LAUNCH_SHARED(c, "k_ds_stats<pair>", bytes * s_hi / sx, bytes_pp * s_hi / sx,
                                 k_ds_stream<1, true, false><<<dim3((unsigned)s_hi, sy, nb), 32, 0, c.stream>>>(
                                     img, R + b0, ridx + b0, (int)cw, (int)ch, n, rows, refstat, map + b0 * n, partial + b0 * ntiles,
                                     (int)sx, s_hi, dm));
                CE_LAUNCH_SHARED(c, "k_ds_stats<pair>", bytes * n_edge / sx, bytes_pp * n_edge / sx,
                                 k_ds_stream<1, true, true><<<dim3(n_edge, sy, nb), 32, 0, c.stream>>>(
                                     img, R + b0, ridx + b0, (int)cw, (int)ch, n, rows, refstat, map + b0 * n, partial + b0 * ntiles,
                                     (int)sx, s_hi, dm));
            } else {
                CE_LAUNCH_SHARED(c, "k_ds_stats<pair>", bytes, bytes_pp,
                                 k_ds_stream<1, false, true><<<dim3(sx, sy, nb), 32, 0, c.stream>>>(
                                     img, R + b0, ridx + b0, (int)cw, (int)ch, n, rows, refstat, map + b0 * n, partial + b0 * ntiles,
                                     (int)sx, 0, dm));
            }
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
