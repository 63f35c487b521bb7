// k_ssim2.cu -- SSIMULACRA2 (fast-ssim2 0.8.0 == libjxl tools/ssimulacra2.cc;
// reference call site src/metrics/ssimulacra2.rs:96) for a batch of B pairs.
//
// Per scale (6 scales, ceil-halving with clamp on LINEAR rgb):
//   k_s2_xyb_down : linear(s) -> positive-XYB(s) planes + linear(s+1)        [HBM-bound, pointwise]
//   k_s2_hpass    : recursive Gaussian along x of {i1,i2,i1^2,i2^2,i1*i2};   [HBM-bound, rows staged in smem]
//                   lane = row, warp = product, 32x32 tiles staged through shared memory
//   k_s2_vpass    : column-parallel recurrence along y (thread = column, 10-row register
//                   delay line) fused with the SSIM / edge-artifact / detail-loss maps and
//                   L1/L4 pooling in fp64 (warp shuffles -> per-block partials)
//   k_s2_reduce   : fixed-order sum of the block partials -> 18 sums per (pair, scale)
// The recurrence is the exact operation sequence of the upstream code so the
// result does not depend on the tiling.
#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

__constant__ RGaussCoef c_rg;

// ------------------------------------------------------------------ xyb + down2
// thread = one 2x2 block of the current scale (= one pixel of the next scale)
__global__ void __launch_bounds__(256) k_s2_xyb_down(const float* __restrict__ lin1, const float* __restrict__ lin2,
                                                      int w, int h, int ow, int oh, size_t n, size_t on,
                                                      float* __restrict__ xyb, float* __restrict__ nlin1,
                                                      float* __restrict__ nlin2, int write_down) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const size_t b = blockIdx.z >> 1;
    const int img = blockIdx.z & 1;
    if (ox >= ow || oy >= oh) return;
    const float* lin = (img ? lin2 : lin1) + b * 3 * n;
    float* xo = xyb + (b * 2 + img) * 3 * n;
    float* no = (img ? nlin2 : nlin1) + b * 3 * on;
    const int x0 = 2 * ox, y0 = 2 * oy;
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const bool vx = (x0 + 1 < w), vy = (y0 + 1 < h);
    const size_t i00 = (size_t)y0 * w + x0, i01 = (size_t)y0 * w + x1, i10 = (size_t)y1 * w + x0, i11 = (size_t)y1 * w + x1;
    float p[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float* pl = lin + (size_t)c * n;
        p[c][0] = pl[i00]; p[c][1] = pl[i01]; p[c][2] = pl[i10]; p[c][3] = pl[i11];
    }
    if (write_down) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float s = 0.0f;
            s += p[c][0]; s += p[c][1]; s += p[c][2]; s += p[c][3];
            no[(size_t)c * on + (size_t)oy * ow + ox] = s * 0.25f;
        }
    }
    float X, Y, Bv;
    xyb_positive(p[0][0], p[1][0], p[2][0], X, Y, Bv);
    xo[i00] = X; xo[n + i00] = Y; xo[2 * n + i00] = Bv;
    if (vx) {
        xyb_positive(p[0][1], p[1][1], p[2][1], X, Y, Bv);
        xo[i01] = X; xo[n + i01] = Y; xo[2 * n + i01] = Bv;
    }
    if (vy) {
        xyb_positive(p[0][2], p[1][2], p[2][2], X, Y, Bv);
        xo[i10] = X; xo[n + i10] = Y; xo[2 * n + i10] = Bv;
        if (vx) {
            xyb_positive(p[0][3], p[1][3], p[2][3], X, Y, Bv);
            xo[i11] = X; xo[n + i11] = Y; xo[2 * n + i11] = Bv;
        }
    }
}

// ------------------------------------------------------------------ horizontal pass
struct RGState {
    float p1, p3, p5, q1, q3, q5;
};
CE_DEVINL float rg_step(RGState& s, float sum) {
    float o1 = sum * c_rg.mul_in[0];
    float o3 = sum * c_rg.mul_in[1];
    float o5 = sum * c_rg.mul_in[2];
    o1 = __fmaf_rn(c_rg.mul_prev2[0], s.q1, o1);
    o3 = __fmaf_rn(c_rg.mul_prev2[1], s.q3, o3);
    o5 = __fmaf_rn(c_rg.mul_prev2[2], s.q5, o5);
    s.q1 = s.p1; s.q3 = s.p3; s.q5 = s.p5;
    o1 = __fmaf_rn(c_rg.mul_prev[0], s.p1, o1);
    o3 = __fmaf_rn(c_rg.mul_prev[1], s.p3, o3);
    o5 = __fmaf_rn(c_rg.mul_prev[2], s.p5, o5);
    s.p1 = o1; s.p3 = o3; s.p5 = o5;
    return (o1 + o3) + o5;
}

#define HP_ROWS 32
#define HP_COLS 32
#define HP_PITCH 33

CE_DEVINL float hp_product(int p, const float* s1, const float* s2, int idx) {
    // p is warp-uniform: 0 i1, 1 i2, 2 i1*i1, 3 i2*i2, 4 i1*i2
    if (p == 0) return s1[idx];
    if (p == 1) return s2[idx];
    if (p == 2) { float a = s1[idx]; return a * a; }
    if (p == 3) { float b = s2[idx]; return b * b; }
    return s1[idx] * s2[idx];
}

// grid (ceil(h/32), 3*B); block 160 = 5 warps (product) x 32 lanes (row)
__global__ void __launch_bounds__(160) k_s2_hpass(const float* __restrict__ xyb, float* __restrict__ hb, int w, int h,
                                                   size_t n) {
    __shared__ float s_in[2][3][HP_ROWS * HP_PITCH];
    __shared__ float s_out[5][HP_ROWS * HP_PITCH];
    const int lane = threadIdx.x & 31, p = threadIdx.x >> 5;
    const size_t b = blockIdx.y / 3;
    const int c = blockIdx.y % 3;
    const int row0 = blockIdx.x * HP_ROWS;
    const float* i1 = xyb + ((b * 2 + 0) * 3 + c) * n;
    const float* i2 = xyb + ((b * 2 + 1) * 3 + c) * n;
    float* outp = hb + ((b * 3 + c) * 5) * n;
    const int nchunks = (w + HP_COLS - 1) / HP_COLS;

    auto load_chunk = [&](int k, int slot) {
        for (int e = threadIdx.x; e < 2 * HP_ROWS * HP_COLS; e += 160) {
            int pl = e >> 10, rr = (e >> 5) & 31, cc = e & 31;
            int y = row0 + rr, x = k * HP_COLS + cc;
            float v = 0.0f;
            if (y < h && x < w) v = (pl ? i2 : i1)[(size_t)y * w + x];
            s_in[pl][slot][rr * HP_PITCH + cc] = v;
        }
    };
    // slot 2 plays chunk -1 (zeros), slot 0 = chunk 0
    for (int e = threadIdx.x; e < 2 * HP_ROWS * HP_PITCH; e += 160) {
        int pl = e / (HP_ROWS * HP_PITCH), r = e % (HP_ROWS * HP_PITCH);
        s_in[pl][2][r] = 0.0f;
    }
    load_chunk(0, 0);
    __syncthreads();
    RGState st = {0, 0, 0, 0, 0, 0};
    const int rbase = lane * HP_PITCH;
    // warm-up steps n = -4..-1: right = in[0..3], left = 0
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++) {
        float r = hp_product(p, s_in[0][0], s_in[1][0], rbase + cidx);
        rg_step(st, r);
    }
    for (int k = 0; k < nchunks; k++) {
        const int s_cur = k % 3, s_next = (k + 1) % 3, s_prev = (k + 2) % 3;
        load_chunk(k + 1, s_next);
        __syncthreads();
        const float* c1 = s_in[0][s_cur];  const float* c2 = s_in[1][s_cur];
        const float* n1 = s_in[0][s_next]; const float* n2 = s_in[1][s_next];
        const float* p1 = s_in[0][s_prev]; const float* p2 = s_in[1][s_prev];
#pragma unroll
        for (int cc = 0; cc < HP_COLS; cc++) {
            float r = (cc + 4 < HP_COLS) ? hp_product(p, c1, c2, rbase + cc + 4) : hp_product(p, n1, n2, rbase + cc + 4 - HP_COLS);
            float l = (cc - 6 >= 0) ? hp_product(p, c1, c2, rbase + cc - 6) : hp_product(p, p1, p2, rbase + cc - 6 + HP_COLS);
            float o = rg_step(st, l + r);
            s_out[p][rbase + cc] = o;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 5 * HP_ROWS * HP_COLS; e += 160) {
            int pp = e >> 10, rr = (e >> 5) & 31, cc = e & 31;
            int y = row0 + rr, x = k * HP_COLS + cc;
            if (y < h && x < w) outp[(size_t)pp * n + (size_t)y * w + x] = s_out[pp][rr * HP_PITCH + cc];
        }
    }
}

// ------------------------------------------------------------------ vertical pass + maps
#define VP_THREADS 128

// grid (ceil(w/128), 3*B); thread = column.  partials: [(b*3+c)][gridDim.x][6]
__global__ void __launch_bounds__(VP_THREADS) k_s2_vpass(const float* __restrict__ xyb, const float* __restrict__ hb,
                                                          int w, int h, size_t n, double* __restrict__ partials,
                                                          float* __restrict__ dbg) {
    __shared__ double scratch[6 * 32];
    const int x = blockIdx.x * VP_THREADS + threadIdx.x;
    const size_t b = blockIdx.y / 3;
    const int c = blockIdx.y % 3;
    const bool active = x < w;
    const float* i1 = xyb + ((b * 2 + 0) * 3 + c) * n;
    const float* i2 = xyb + ((b * 2 + 1) * 3 + c) * n;
    const float* hp = hb + ((b * 3 + c) * 5) * n;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (active) {
        RGState st[5];
        float ring[5][10];
#pragma unroll
        for (int p = 0; p < 5; p++) {
            st[p] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 10; j++) ring[p][j] = 0.0f;
        }
        const int total = h + 4;  // input rows j = 0 .. h+3 (rows >= h are zero); output row n = j - 4
        for (int j0 = 0; j0 < total; j0 += 10) {
#pragma unroll
            for (int jj = 0; jj < 10; jj++) {
                const int j = j0 + jj;
                if (j < total) {
                    float o[5];
#pragma unroll
                    for (int p = 0; p < 5; p++) {
                        float r = (j < h) ? hp[(size_t)p * n + (size_t)j * w + x] : 0.0f;
                        float l = ring[p][jj];
                        ring[p][jj] = r;
                        o[p] = rg_step(st[p], l + r);
                    }
                    const int y = j - 4;
                    if (y >= 0) {
                        const size_t idx = (size_t)y * w + x;
                        const float a1 = i1[idx], a2 = i2[idx];
                        const float m1 = o[0], m2 = o[1], s11 = o[2], s22 = o[3], s12 = o[4];
                        if (dbg) {
                            float* d = dbg + (size_t)c * 7 * n + idx;
                            d[2 * n] = m1; d[3 * n] = m2; d[4 * n] = s11; d[5 * n] = s22; d[6 * n] = s12;
                        }
                        // ssim_map
                        float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
                        float mdiff = m1 - m2;
                        float num_m = __fmaf_rn(mdiff, -mdiff, 1.0f);
                        float num_s = __fmaf_rn(2.0f, s12 - m12, 0.0009f);
                        float denom_s = ((s11 - m11) + (s22 - m22)) + 0.0009f;
                        double d = 1.0 - (double)((num_m * num_s) / denom_s);
                        if (!(d > 0.0)) d = 0.0;
                        double d2 = d * d;
                        acc[0] += d;
                        acc[1] += d2 * d2;
                        // edge_diff_map
                        double d1 = (1.0 + (double)fabsf(a2 - m2)) / (1.0 + (double)fabsf(a1 - m1)) - 1.0;
                        double art = d1 > 0.0 ? d1 : 0.0;
                        double det = d1 < 0.0 ? -d1 : 0.0;
                        double a2_ = art * art, l2 = det * det;
                        acc[2] += art;
                        acc[3] += a2_ * a2_;
                        acc[4] += det;
                        acc[5] += l2 * l2;
                    }
                }
            }
        }
    }
    block_sum<6>(acc, scratch);
    if (threadIdx.x == 0) {
        double* o = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 6;
#pragma unroll
        for (int i = 0; i < 6; i++) o[i] = acc[i];
    }
}

// sums[b][scale][c*6+k] = sum over blocks, fixed order
__global__ void k_s2_reduce(const double* __restrict__ partials, int nblk, size_t total /* B*3*6 */, int scale,
                            double* __restrict__ sums) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    size_t bc = t / 6;
    int k = (int)(t % 6);
    const double* p = partials + bc * nblk * 6 + k;
    double s = 0.0;
    for (int i = 0; i < nblk; i++) s += p[(size_t)i * 6];
    size_t b = bc / 3;
    int c = (int)(bc % 3);
    sums[(b * 6 + scale) * 18 + c * 6 + k] = s;
}

void ssim2_init(Context& c) {
    CE_CUDA(cudaMemcpyToSymbol(c_rg, &c.rg, sizeof(RGaussCoef), 0, cudaMemcpyHostToDevice));
}

size_t ssim2_workspace_per_pair(size_t w, size_t h) {
    size_t n = w * h;
    // next-scale linear (2*3*n/4 .. geometric), xyb 6n, hb 15n, partials
    size_t bytes = (6 * n + 15 * n) * 4 + (2 * 3 * ((w + 1) / 2) * ((h + 1) / 2)) * 4 * 2 + 3 * cdiv(w, VP_THREADS) * 6 * 8 + 4096;
    return bytes;
}

int ssim2_run(Context& c, const float* lin1_in, const float* lin2_in, size_t B, size_t w, size_t h, double* d_sums,
              float* dbg_planes) {
    size_t mark = c.arena.mark();
    const size_t n0 = w * h;
    float* xyb = c.arena.alloc<float>(B * 6 * n0);
    float* hb = c.arena.alloc<float>(B * 15 * n0);
    const size_t ow0 = (w + 1) / 2, oh0 = (h + 1) / 2;
    float* nl[2][2];  // ping-pong next-scale linear buffers [pingpong][img]
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) nl[i][j] = c.arena.alloc<float>(B * 3 * ow0 * oh0);
    const int nblk0 = cdiv(w, VP_THREADS);
    double* partials = c.arena.alloc<double>(B * 3 * nblk0 * 6);

    const float* l1 = lin1_in;
    const float* l2 = lin2_in;
    size_t cw = w, ch = h;
    int ns = 0;
    for (int scale = 0; scale < 6; scale++) {
        const size_t n = cw * ch;
        const size_t ow = (cw + 1) / 2, oh = (ch + 1) / 2;
        // does a next scale exist? (check on the current size, as upstream does at the top of its loop)
        const bool has_next = (scale + 1 < 6) && !(cw < 8 || ch < 8);
        float* d1 = nl[scale & 1][0];
        float* d2 = nl[scale & 1][1];
        {
            dim3 grid(cdiv(ow, 64), cdiv(oh, 4), (unsigned)(B * 2));
            CE_LAUNCH(c, "k_s2_xyb_down", (double)B * 4 * (12 * n + (has_next ? 6 * ow * oh : 0)),
                      k_s2_xyb_down<<<grid, 256, 0, c.stream>>>(l1, l2, (int)cw, (int)ch, (int)ow, (int)oh, n, ow * oh, xyb, d1,
                                                                 d2, has_next ? 1 : 0));
        }
        {
            dim3 grid(cdiv(ch, HP_ROWS), (unsigned)(B * 3));
            CE_LAUNCH(c, "k_s2_hpass", (double)B * 3 * 7 * n * 4, k_s2_hpass<<<grid, 160, 0, c.stream>>>(xyb, hb, (int)cw, (int)ch, n));
        }
        const int nblk = cdiv(cw, VP_THREADS);
        float* dbg = (dbg_planes && scale == 0) ? dbg_planes : nullptr;
        {
            dim3 grid(nblk, (unsigned)(B * 3));
            CE_LAUNCH(c, "k_s2_vpass", (double)B * 3 * 7 * n * 4,
                      k_s2_vpass<<<grid, VP_THREADS, 0, c.stream>>>(xyb, hb, (int)cw, (int)ch, n, partials, dbg));
        }
        {
            size_t total = B * 3 * 6;
            CE_LAUNCH(c, "k_s2_reduce", (double)total * (nblk + 1) * 8,
                      k_s2_reduce<<<cdiv(total, 128), 128, 0, c.stream>>>(partials, nblk, total, scale, d_sums));
        }
        if (dbg) {
            for (int cc = 0; cc < 3; cc++) {
                CE_CUDA(cudaMemcpyAsync(dbg + (size_t)cc * 7 * n, xyb + (size_t)cc * n, n * 4, cudaMemcpyDeviceToDevice, c.stream));
                CE_CUDA(cudaMemcpyAsync(dbg + (size_t)cc * 7 * n + n, xyb + (size_t)(3 + cc) * n, n * 4, cudaMemcpyDeviceToDevice, c.stream));
            }
        }
        CE_CUDA(cudaGetLastError());
        ns++;
        if (!has_next) break;
        l1 = d1; l2 = d2; cw = ow; ch = oh;
    }
    c.arena.release(mark);
    return ns;
}

}  // namespace ce
