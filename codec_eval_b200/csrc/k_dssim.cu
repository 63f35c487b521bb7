// k_dssim.cu -- DSSIM (dssim-core 3.4.0; reference call site src/metrics/dssim.rs:52-68)
// for a batch of B pairs.
//
// Per scale (<= 5 scales, floor-halving with crop on LINEAR rgb(a)):
//   k_ds_down    : 2x2 average to the next scale                               [HBM-bound]
//   k_ds_lab     : linear -> L (final) and a,b (to be pre-blurred)              [HBM-bound, pointwise]
//   k_ds_blur2   : chroma pre-blur: the 3x3 kernel applied twice, fused through a
//                  shared-memory tile (halo 2, clamp-replicate per pass)         [HBM-bound]
//   k_ds_stats   : per channel the five double-3x3 blurs {ch1,ch2,ch1^2,ch2^2,ch1*ch2}
//                  of a 32x16 tile fused in shared memory, channel-averaged SSIM map
//                  written once + fp64 block partial of its sum                   [HBM<->ALU ridge]
//   k_ds_mean    : fixed-order reduce -> sum(map), avg = max(mean,0)^(0.5^scale)
//   k_ds_mad     : sum |avg - map_i| in fp64 -> block partials; k_ds_mad_reduce fixes the order
#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

#define DS_TW 32
#define DS_TH 16
#define DS_IW (DS_TW + 4)
#define DS_IH (DS_TH + 4)
#define DS_FW (DS_TW + 2)
#define DS_FH (DS_TH + 2)

__constant__ float c_dsk[9] = {0.095332f, 0.118095f, 0.095332f, 0.118095f, 0.146293f,
                               0.118095f, 0.095332f, 0.118095f, 0.095332f};

CE_DEVINL float ds_k9(float v00, float v01, float v02, float v10, float v11, float v12, float v20, float v21, float v22) {
    float a = (v00 * c_dsk[0] + v01 * c_dsk[1]) + v02 * c_dsk[2];
    float b = (v10 * c_dsk[3] + v11 * c_dsk[4]) + v12 * c_dsk[5];
    float c = (v20 * c_dsk[6] + v21 * c_dsk[7]) + v22 * c_dsk[8];
    return (a + b) + c;
}

// ------------------------------------------------------------------ downsample
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

// ------------------------------------------------------------------ Lab
// lin: [B][3][n] per image, alpha nullable [B][n]; writes img[(b*2+which)*3 + 0] = L,
// chroma[(b*2+which)*2 + {0,1}] = a,b (un-blurred)
__global__ void __launch_bounds__(256) k_ds_lab(const float* __restrict__ lin, const float* __restrict__ alpha, int w,
                                                 size_t n, size_t total, int which, float* __restrict__ img,
                                                 float* __restrict__ chroma) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t b = t / n, i = t - b * n;
        const float* p = lin + b * 3 * n + i;
        float r = p[0], g = p[n], bl = p[2 * n];
        if (alpha) {
            float a = alpha[b * n + i];
            if (a < 255.0f / 256.0f) {
                unsigned y = (unsigned)(i / w), x = (unsigned)(i - (size_t)y * w);
                unsigned nn = (x + 11u) ^ (y + 11u);
                if (nn & 16u) r += 1.0f - a;
                if (nn & 8u) g += 1.0f - a;
                if (nn & 32u) bl += 1.0f - a;
            }
        }
        float L, A, Bv;
        ds_to_lab(r, g, bl, L, A, Bv);
        img[((b * 2 + which) * 3) * n + i] = L;
        chroma[((b * 2 + which) * 2 + 0) * n + i] = A;
        chroma[((b * 2 + which) * 2 + 1) * n + i] = Bv;
    }
}

// ------------------------------------------------------------------ shared tile helpers
// s_in[(iy,ix)] = plane[clamp(ty0-2+iy)][clamp(tx0-2+ix)]
CE_DEVINL void ds_load_tile(const float* __restrict__ plane, int w, int h, int tx0, int ty0, float* s_in) {
    for (int e = threadIdx.x; e < DS_IW * DS_IH; e += blockDim.x) {
        int iy = e / DS_IW, ix = e - iy * DS_IW;
        int gx = min(max(tx0 - 2 + ix, 0), w - 1), gy = min(max(ty0 - 2 + iy, 0), h - 1);
        s_in[e] = plane[(size_t)gy * w + gx];
    }
}

// ------------------------------------------------------------------ chroma pre-blur (two passes)
// grid (tiles_x, tiles_y, nplanes); in: chroma[(b*2+img)*2 + k], out: img[(b*2+img)*3 + 1 + k]
__global__ void __launch_bounds__(256) k_ds_blur2(const float* __restrict__ chroma, int w, int h, size_t n,
                                                   float* __restrict__ img) {
    __shared__ float s_in[DS_IW * DS_IH];
    __shared__ float s_f[DS_FW * DS_FH];
    const int tx0 = blockIdx.x * DS_TW, ty0 = blockIdx.y * DS_TH;
    const size_t pl = blockIdx.z;  // (b*2+img)*2 + k
    const float* src = chroma + pl * n;
    float* dst = img + ((pl >> 1) * 3 + 1 + (pl & 1)) * n;
    ds_load_tile(src, w, h, tx0, ty0, s_in);
    __syncthreads();
    for (int e = threadIdx.x; e < DS_FW * DS_FH; e += blockDim.x) {
        int ry = e / DS_FW, rx = e - ry * DS_FW;
        int gx = min(max(tx0 - 1 + rx, 0), w - 1), gy = min(max(ty0 - 1 + ry, 0), h - 1);
        int ix = gx - tx0 + 2, iy = gy - ty0 + 2;
        const float* r0 = s_in + (iy - 1) * DS_IW + ix;
        const float* r1 = r0 + DS_IW;
        const float* r2 = r1 + DS_IW;
        s_f[e] = ds_k9(r0[-1], r0[0], r0[1], r1[-1], r1[0], r1[1], r2[-1], r2[0], r2[1]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < DS_TW * DS_TH; e += blockDim.x) {
        int oy = e / DS_TW, ox = e - oy * DS_TW;
        int x = tx0 + ox, y = ty0 + oy;
        if (x < w && y < h) {
            const float* r0 = s_f + oy * DS_FW + ox + 1;
            const float* r1 = r0 + DS_FW;
            const float* r2 = r1 + DS_FW;
            dst[(size_t)y * w + x] = ds_k9(r0[-1], r0[0], r0[1], r1[-1], r1[0], r1[1], r2[-1], r2[0], r2[1]);
        }
    }
}

// ------------------------------------------------------------------ statistics + SSIM map
// grid (tiles_x, tiles_y, B); img: [B][2][3][n]; map: [B][n]; partial: [B][tiles] doubles
__global__ void __launch_bounds__(256) k_ds_stats(const float* __restrict__ img, int w, int h, size_t n,
                                                   float* __restrict__ map, double* __restrict__ partial) {
    __shared__ float s_in1[DS_IW * DS_IH];
    __shared__ float s_in2[DS_IW * DS_IH];
    __shared__ float s_f[5][DS_FW * DS_FH];
    __shared__ double scratch[32];
    const int tx0 = blockIdx.x * DS_TW, ty0 = blockIdx.y * DS_TH;
    const size_t b = blockIdx.z;
    // per-thread outputs: (ox, oy0) and (ox, oy0 + 8)
    const int ox = threadIdx.x & 31, oy0 = threadIdx.x >> 5;
    float ch_m11[2][3], ch_m12[2][3], ch_m22[2][3], ch_s1[2][3], ch_s2[2][3], ch_s12[2][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float* p1 = img + ((b * 2 + 0) * 3 + c) * n;
        const float* p2 = img + ((b * 2 + 1) * 3 + c) * n;
        __syncthreads();  // previous channel's s_f / s_in no longer read
        ds_load_tile(p1, w, h, tx0, ty0, s_in1);
        ds_load_tile(p2, w, h, tx0, ty0, s_in2);
        __syncthreads();
        for (int e = threadIdx.x; e < DS_FW * DS_FH; e += blockDim.x) {
            int ry = e / DS_FW, rx = e - ry * DS_FW;
            int gx = min(max(tx0 - 1 + rx, 0), w - 1), gy = min(max(ty0 - 1 + ry, 0), h - 1);
            int ix = gx - tx0 + 2, iy = gy - ty0 + 2;
            float u[9], v[9];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    int o = (iy - 1 + dy) * DS_IW + ix - 1 + dx;
                    u[dy * 3 + dx] = s_in1[o];
                    v[dy * 3 + dx] = s_in2[o];
                }
            s_f[0][e] = ds_k9(u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7], u[8]);
            s_f[1][e] = ds_k9(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
            s_f[2][e] = ds_k9(u[0] * u[0], u[1] * u[1], u[2] * u[2], u[3] * u[3], u[4] * u[4], u[5] * u[5], u[6] * u[6],
                              u[7] * u[7], u[8] * u[8]);
            s_f[3][e] = ds_k9(v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3], v[4] * v[4], v[5] * v[5], v[6] * v[6],
                              v[7] * v[7], v[8] * v[8]);
            s_f[4][e] = ds_k9(u[0] * v[0], u[1] * v[1], u[2] * v[2], u[3] * v[3], u[4] * v[4], u[5] * v[5], u[6] * v[6],
                              u[7] * v[7], u[8] * v[8]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int oy = oy0 + k * 8;
            float q[5];
#pragma unroll
            for (int f = 0; f < 5; f++) {
                const float* r0 = s_f[f] + oy * DS_FW + ox + 1;
                const float* r1 = r0 + DS_FW;
                const float* r2 = r1 + DS_FW;
                q[f] = ds_k9(r0[-1], r0[0], r0[1], r1[-1], r1[0], r1[1], r2[-1], r2[0], r2[1]);
            }
            float mu1 = q[0], mu2 = q[1];
            float m11 = mu1 * mu1, m12 = mu1 * mu2, m22 = mu2 * mu2;
            ch_m11[k][c] = m11; ch_m12[k][c] = m12; ch_m22[k][c] = m22;
            ch_s1[k][c] = q[2] - m11;
            ch_s2[k][c] = q[3] - m22;
            ch_s12[k][c] = q[4] - m12;
        }
    }
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f, third = 1.0f / 3.0f;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int x = tx0 + ox, y = ty0 + oy0 + k * 8;
        if (x < w && y < h) {
            float mu1_sq = ((ch_m11[k][0] + ch_m11[k][1]) + ch_m11[k][2]) * third;
            float mu2_sq = ((ch_m22[k][0] + ch_m22[k][1]) + ch_m22[k][2]) * third;
            float mu1_mu2 = ((ch_m12[k][0] + ch_m12[k][1]) + ch_m12[k][2]) * third;
            float sigma1_sq = ((ch_s1[k][0] + ch_s1[k][1]) + ch_s1[k][2]) * third;
            float sigma2_sq = ((ch_s2[k][0] + ch_s2[k][1]) + ch_s2[k][2]) * third;
            float sigma12 = ((ch_s12[k][0] + ch_s12[k][1]) + ch_s12[k][2]) * third;
            float v = (__fmaf_rn(2.0f, mu1_mu2, c1) * __fmaf_rn(2.0f, sigma12, c2)) /
                      (((mu1_sq + mu2_sq) + c1) * ((sigma1_sq + sigma2_sq) + c2));
            map[b * n + (size_t)y * w + x] = v;
            acc += (double)v;
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
    // img 6n, chroma 4n, map n, next-scale rgba ping-pong 2*2*4*(n/4)
    return (6 * n + 4 * n + n + 4 * n) * 4 + (tiles + DS_MAD_BLOCKS + 4) * 8 + 8192;
}

int dssim_run(Context& c, const float* lin1_in, const float* lin2_in, const float* alpha1_in, const float* alpha2_in, size_t B,
              size_t w, size_t h, double* d_out, float* dbg_map0) {
    size_t ws[5], hs[5];
    const int ns = dssim_num_scales(w, h, ws, hs);
    size_t mark = c.arena.mark();
    const size_t n0 = w * h;
    float* img = c.arena.alloc<float>(B * 6 * n0);
    float* chroma = c.arena.alloc<float>(B * 4 * n0);
    float* map = c.arena.alloc<float>(B * n0);
    const size_t tiles0 = (size_t)cdiv(w, DS_TW) * cdiv(h, DS_TH);
    double* partial = c.arena.alloc<double>(B * std::max<size_t>(tiles0, DS_MAD_BLOCKS));
    double* avg = c.arena.alloc<double>(B);
    const bool has_alpha = alpha1_in != nullptr;
    const int npl = has_alpha ? 4 : 3;
    // next-scale planes: [pingpong][img] each [B][npl][n/4]; alpha stored as plane 3 of each pair
    float* nl[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    if (ns > 1)
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) nl[i][j] = c.arena.alloc<float>(B * 3 * ws[1] * hs[1]);
    float* nal[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    if (ns > 1 && has_alpha)
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) nal[i][j] = c.arena.alloc<float>(B * ws[1] * hs[1]);
    (void)npl;

    const float* l[2] = {lin1_in, lin2_in};
    const float* al[2] = {alpha1_in, alpha2_in};
    const unsigned wave = (unsigned)c.sm_count * 8;
    for (int s = 0; s < ns; s++) {
        const size_t cw = ws[s], ch = hs[s], n = cw * ch;
        if (s > 0) {
            const size_t pw = ws[s - 1], pn = ws[s - 1] * hs[s - 1];
            for (int im = 0; im < 2; im++) {
                float* dst = nl[s & 1][im];
                size_t total = B * 3 * n;
                CE_LAUNCH(c, "k_ds_down", (double)total * 20,
                          k_ds_down<<<std::min<unsigned>(cdiv(total, 256), wave * 4), 256, 0, c.stream>>>(l[im], (int)pw, pn, (int)cw,
                                                                                                        (int)ch, n, total, dst));
                if (has_alpha) {
                    float* adst = nal[s & 1][im];
                    size_t atotal = B * n;
                    CE_LAUNCH(c, "k_ds_down", (double)atotal * 20,
                              k_ds_down<<<std::min<unsigned>(cdiv(atotal, 256), wave * 4), 256, 0, c.stream>>>(
                                  al[im], (int)pw, pn, (int)cw, (int)ch, n, atotal, adst));
                    al[im] = adst;
                }
                l[im] = dst;
            }
        }
        for (int im = 0; im < 2; im++) {
            size_t total = B * n;
            CE_LAUNCH(c, "k_ds_lab", (double)total * (has_alpha ? 28 : 24),
                      k_ds_lab<<<std::min<unsigned>(cdiv(total, 256), wave * 4), 256, 0, c.stream>>>(
                          l[im], has_alpha ? al[im] : nullptr, (int)cw, n, total, im, img, chroma));
        }
        const unsigned tx = cdiv(cw, DS_TW), ty = cdiv(ch, DS_TH);
        for (size_t p0 = 0; p0 < B * 4; p0 += 32768) {  // gridDim.z <= 65535; keep pairs whole (4 planes per pair)
            unsigned np = (unsigned)std::min<size_t>(32768, B * 4 - p0);
            dim3 grid(tx, ty, np);
            CE_LAUNCH(c, "k_ds_blur2", (double)np * n * 8,
                      k_ds_blur2<<<grid, 256, 0, c.stream>>>(chroma + p0 * n, (int)cw, (int)ch, n, img + (p0 / 2) * 3 * n));
        }
        const int ntiles = (int)(tx * ty);
        for (size_t b0 = 0; b0 < B; b0 += 32768) {
            unsigned nb = (unsigned)std::min<size_t>(32768, B - b0);
            dim3 grid(tx, ty, nb);
            CE_LAUNCH(c, "k_ds_stats", (double)nb * n * 28,
                      k_ds_stats<<<grid, 256, 0, c.stream>>>(img + b0 * 6 * n, (int)cw, (int)ch, n, map + b0 * n, partial + b0 * ntiles));
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
