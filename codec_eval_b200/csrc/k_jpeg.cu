// k_jpeg.cu -- on-device distortion source: the sample-domain effect of a baseline JPEG
// encode -> decode with the IJG / libjpeg-turbo defaults, so a quality sweep (codec-iter's eval loop,
// crates/codec-iter/src/eval.rs:153-172: encode -> decode -> compare per quality) needs only the REFERENCE images
// on the device; the distorted images are produced here and never cross PCIe.  Entropy coding is lossless and
// is skipped, so there is no bitstream and no file size.
//
// Integer arithmetic throughout, bit-exact with libjpeg-turbo's baseline path (checked against Pillow):
//   k_jpg_ycc  : RGB8 -> Y, Cb, Cr planes padded by edge replication to whole MCUs (jccolor.c rgb_ycc_convert;
//                4:2:0: jcsample.c h2v2_downsample with the alternating bias, bottom rows as jcprepct.c pads them)
//   k_jpg_dct  : thread = one 8x8 block: jfdctint.c forward DCT, quantise / dequantise with the Annex-K tables
//                scaled by `quality` (jcparam.c, jcdctmgr.c), jidctint.c inverse DCT, range limit
//   k_jpg_rgb  : jdsample.c h2v2_fancy_upsample (4:2:0) + jdcolor.c ycc_rgb_convert -> RGB8 interleaved
// All three are byte-granular and HBM-bound; they cost ~1 % of the metrics they feed.
#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

#define JFIX16(x) ((int)((x) * 65536.0 + 0.5))
#define JFIX13(x) ((int)((x) * 8192.0 + 0.5))
#define JDESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

CE_DEVINL void jpg_rgb_to_ycc(int r, int g, int b, int& y, int& cb, int& cr) {
    const int half = 1 << 15, off = 128 << 16;
    y = (JFIX16(0.29900) * r + JFIX16(0.58700) * g + JFIX16(0.11400) * b + half) >> 16;
    cb = (-JFIX16(0.16874) * r - JFIX16(0.33126) * g + JFIX16(0.50000) * b + off + half - 1) >> 16;
    cr = (JFIX16(0.50000) * r - JFIX16(0.41869) * g - JFIX16(0.08131) * b + off + half - 1) >> 16;
}
CE_DEVINL int jpg_clamp8(int v) { return min(max(v, 0), 255); }

struct JpgGeom {
    int w, h;       // image
    int pw, ph;     // padded luma plane (multiple of 8, or 16 for 4:2:0)
    int cw, ch;     // valid chroma size
    int cpw, cph;   // padded chroma plane
    size_t plane_stride;   // bytes between the Y, Cb, Cr planes of one image (>= pw*ph, 256-B multiple)
};

// grid (ceil(X/32), ceil(Y/8), n_img), block (32, 8).  SS == 0: thread = one padded pixel; SS == 2: thread = one
// padded chroma sample + its 2x2 luma block.  ycc: [n_img][3][plane_stride]
template <int SS>
__global__ void __launch_bounds__(256) k_jpg_ycc(const uint8_t* __restrict__ rgb, JpgGeom g, uint8_t* __restrict__ ycc) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    const size_t img = blockIdx.z;
    const uint8_t* src = rgb + img * (size_t)g.w * g.h * 3;
    uint8_t* Y = ycc + img * 3 * g.plane_stride;
    uint8_t* Cb = Y + g.plane_stride;
    uint8_t* Cr = Cb + g.plane_stride;
    if (SS == 0) {
        if (i >= g.pw || j >= g.ph) return;
        const uint8_t* s = src + ((size_t)min(j, g.h - 1) * g.w + min(i, g.w - 1)) * 3;
        int y, cb, cr;
        jpg_rgb_to_ycc(s[0], s[1], s[2], y, cb, cr);
        const size_t o = (size_t)j * g.pw + i;
        Y[o] = (uint8_t)y; Cb[o] = (uint8_t)cb; Cr[o] = (uint8_t)cr;
    } else {
        if (i >= g.cpw || j >= g.cph) return;
        const int x0 = min(2 * i, g.w - 1), x1 = min(2 * i + 1, g.w - 1);
        // luma: plain edge replication
        {
            const int y0 = min(2 * j, g.h - 1), y1 = min(2 * j + 1, g.h - 1);
            const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const uint8_t* s = src + ((size_t)ys[a] * g.w + xs[b]) * 3;
                    int y, cb, cr;
                    jpg_rgb_to_ycc(s[0], s[1], s[2], y, cb, cr);
                    Y[(size_t)(2 * j + a) * g.pw + 2 * i + b] = (uint8_t)y;
                }
        }
        // chroma: the colour buffer is padded to an even number of rows, then downsampled; the DOWNSAMPLED rows
        // are replicated to a whole iMCU
        const int jj = min(j, g.ch - 1);
        const int y0 = 2 * jj, y1 = min(2 * jj + 1, g.h - 1);
        const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
        int sb = 0, sr = 0;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const uint8_t* s = src + ((size_t)ys[a] * g.w + xs[b]) * 3;
                int y, cb, cr;
                jpg_rgb_to_ycc(s[0], s[1], s[2], y, cb, cr);
                sb += cb; sr += cr;
            }
        const int bias = (i & 1) ? 2 : 1;
        Cb[(size_t)j * g.cpw + i] = (uint8_t)((sb + bias) >> 2);
        Cr[(size_t)j * g.cpw + i] = (uint8_t)((sr + bias) >> 2);
    }
}

// jfdctint.c, one 1-D pass over 8 values (FIRST: row pass, results scaled up by 2^PASS1_BITS)
template <bool FIRST>
CE_DEVINL void jpg_fdct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6, t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int n = FIRST ? 11 : 15;
    if (FIRST) { d0 = (t10 + t11) << 2; d4 = (t10 - t11) << 2; }
    else { d0 = JDESCALE(t10 + t11, 2); d4 = JDESCALE(t10 - t11, 2); }
    int z1 = (t12 + t13) * JFIX13(0.541196100);
    d2 = JDESCALE(z1 + t13 * JFIX13(0.765366865), n);
    d6 = JDESCALE(z1 + t12 * (-JFIX13(1.847759065)), n);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * JFIX13(1.175875602);
    t4 *= JFIX13(0.298631336); t5 *= JFIX13(2.053119869); t6 *= JFIX13(3.072711026); t7 *= JFIX13(1.501321110);
    z1 *= -JFIX13(0.899976223); z2 *= -JFIX13(2.562915447); z3 *= -JFIX13(1.961570560); z4 *= -JFIX13(0.390180644);
    z3 += z5; z4 += z5;
    d7 = JDESCALE(t4 + z1 + z3, n);
    d5 = JDESCALE(t5 + z2 + z4, n);
    d3 = JDESCALE(t6 + z2 + z3, n);
    d1 = JDESCALE(t7 + z1 + z4, n);
}
// jidctint.c, one 1-D pass; N = 11 for the column pass, 18 for the row pass
template <int N>
CE_DEVINL void jpg_idct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int z2 = d2, z3 = d6;
    int z1 = (z2 + z3) * JFIX13(0.541196100);
    int tmp2 = z1 + z3 * (-JFIX13(1.847759065)), tmp3 = z1 + z2 * JFIX13(0.765366865);
    z2 = d0; z3 = d4;
    int tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * JFIX13(1.175875602);
    tmp0 *= JFIX13(0.298631336); tmp1 *= JFIX13(2.053119869); tmp2 *= JFIX13(3.072711026); tmp3 *= JFIX13(1.501321110);
    z1 *= -JFIX13(0.899976223); z2 *= -JFIX13(2.562915447); z3 *= -JFIX13(1.961570560); z4 *= -JFIX13(0.390180644);
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    d0 = JDESCALE(tmp10 + tmp3, N); d7 = JDESCALE(tmp10 - tmp3, N);
    d1 = JDESCALE(tmp11 + tmp2, N); d6 = JDESCALE(tmp11 - tmp2, N);
    d2 = JDESCALE(tmp12 + tmp1, N); d5 = JDESCALE(tmp12 - tmp1, N);
    d3 = JDESCALE(tmp13 + tmp0, N); d4 = JDESCALE(tmp13 - tmp0, N);
}

// Annex-K base tables (natural order): [0] luminance, [1] chrominance
__constant__ uint8_t c_jpg_base[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,  69,  56,
     14, 17, 22, 29, 51,  87,  80,  62,  18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

#define JPG_MAXQ 32
struct JpgQualities {
    int q[JPG_MAXQ];
};

// grid (ceil(blocks/128), 3 components, n_ref * n_q), block 128; thread = one 8x8 block.
// ycc: [n_ref][3][plane_stride] -> rt: [n_ref * n_q][3][plane_stride].  The quantisation table of (quality,
// component) is formed per block: jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE).
__global__ void __launch_bounds__(128) k_jpg_dct(const uint8_t* __restrict__ ycc, JpgGeom g, const __grid_constant__ JpgQualities ql,
                                                  int n_q, uint8_t* __restrict__ rt) {
    __shared__ int s_q[64];
    const int comp = blockIdx.y;
    const size_t rk = blockIdx.z, r = rk / n_q;
    const int k = (int)(rk - r * n_q);
    if (threadIdx.x < 64) {
        const int quality = min(max(ql.q[k], 1), 100);
        const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
        s_q[threadIdx.x] = min(max(((int)c_jpg_base[comp ? 1 : 0][threadIdx.x] * scale + 50) / 100, 1), 255);
    }
    __syncthreads();
    const int pw = comp ? g.cpw : g.pw, ph = comp ? g.cph : g.ph;
    const int bw = pw >> 3, nblk = bw * (ph >> 3);
    const int blk = blockIdx.x * 128 + threadIdx.x;
    if (blk >= nblk) return;
    const int by = blk / bw, bx = blk - by * bw;
    const uint8_t* src = ycc + (r * 3 + comp) * g.plane_stride + ((size_t)by * 8) * pw + bx * 8;
    uint8_t* dst = rt + (rk * 3 + comp) * g.plane_stride + ((size_t)by * 8) * pw + bx * 8;
    int d[8][8];
#pragma unroll
    for (int y = 0; y < 8; y++) {
        const uint2 v = *reinterpret_cast<const uint2*>(src + (size_t)y * pw);
#pragma unroll
        for (int x = 0; x < 4; x++) {
            d[y][x] = (int)((v.x >> (8 * x)) & 255u) - 128;
            d[y][4 + x] = (int)((v.y >> (8 * x)) & 255u) - 128;
        }
    }
#pragma unroll
    for (int y = 0; y < 8; y++) jpg_fdct8<true>(d[y][0], d[y][1], d[y][2], d[y][3], d[y][4], d[y][5], d[y][6], d[y][7]);
#pragma unroll
    for (int x = 0; x < 8; x++) jpg_fdct8<false>(d[0][x], d[1][x], d[2][x], d[3][x], d[4][x], d[5][x], d[6][x], d[7][x]);
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const int q = s_q[y * 8 + x], qv = q << 3;
            const int t = d[y][x], a = (abs(t) + (qv >> 1)) / qv;
            d[y][x] = (t < 0 ? -a : a) * q;
        }
#pragma unroll
    for (int x = 0; x < 8; x++) jpg_idct8<11>(d[0][x], d[1][x], d[2][x], d[3][x], d[4][x], d[5][x], d[6][x], d[7][x]);
#pragma unroll
    for (int y = 0; y < 8; y++) {
        jpg_idct8<18>(d[y][0], d[y][1], d[y][2], d[y][3], d[y][4], d[y][5], d[y][6], d[y][7]);
        uint2 v = make_uint2(0u, 0u);
#pragma unroll
        for (int x = 0; x < 4; x++) {
            v.x |= (unsigned)jpg_clamp8(d[y][x] + 128) << (8 * x);
            v.y |= (unsigned)jpg_clamp8(d[y][4 + x] + 128) << (8 * x);
        }
        *reinterpret_cast<uint2*>(dst + (size_t)y * pw) = v;
    }
}

CE_DEVINL void jpg_ycc_to_rgb(int y, int cb, int cr, uint8_t* o) {
    const int half = 1 << 15;
    const int x = cr - 128, xb = cb - 128;
    o[0] = (uint8_t)jpg_clamp8(y + ((JFIX16(1.40200) * x + half) >> 16));
    o[1] = (uint8_t)jpg_clamp8(y + ((-JFIX16(0.34414) * xb + half - JFIX16(0.71414) * x) >> 16));
    o[2] = (uint8_t)jpg_clamp8(y + ((JFIX16(1.77200) * xb + half) >> 16));
}

// grid (ceil(X/32), ceil(Y/8), n_img), block (32, 8).  SS == 0: thread = one pixel; SS == 2: thread = one valid chroma
// sample and its (up to) 2x2 output pixels.  rt: [n_img][3][plane_stride] -> out: [n_img][h][w][3]
template <int SS>
__global__ void __launch_bounds__(256) k_jpg_rgb(const uint8_t* __restrict__ rt, JpgGeom g, uint8_t* __restrict__ out) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    const size_t img = blockIdx.z;
    const uint8_t* Y = rt + img * 3 * g.plane_stride;
    const uint8_t* Cb = Y + g.plane_stride;
    const uint8_t* Cr = Cb + g.plane_stride;
    uint8_t* dst = out + img * (size_t)g.w * g.h * 3;
    if (SS == 0) {
        if (i >= g.w || j >= g.h) return;
        const size_t o = (size_t)j * g.pw + i;
        jpg_ycc_to_rgb(Y[o], Cb[o], Cr[o], dst + ((size_t)j * g.w + i) * 3);
    } else {
        if (i >= g.cw || j >= g.ch) return;
        const int ja = max(j - 1, 0), jb = min(j + 1, g.ch - 1);
        const int il = max(i - 1, 0), ir = min(i + 1, g.cw - 1);
        int up[2][2][2];   // [plane][output row][output column]
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* c = pl ? Cr : Cb;
            const int cl = c[(size_t)j * g.cpw + il], cc = c[(size_t)j * g.cpw + i], cr = c[(size_t)j * g.cpw + ir];
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const size_t nrow = (size_t)(v ? jb : ja) * g.cpw;
                const int s_l = 3 * cl + c[nrow + il], s_c = 3 * cc + c[nrow + i], s_r = 3 * cr + c[nrow + ir];
                up[pl][v][0] = i == 0 ? (s_c * 4 + 8) >> 4 : (3 * s_c + s_l + 8) >> 4;
                up[pl][v][1] = i == g.cw - 1 ? (s_c * 4 + 7) >> 4 : (3 * s_c + s_r + 7) >> 4;
            }
        }
#pragma unroll
        for (int v = 0; v < 2; v++)
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int x = 2 * i + u, y = 2 * j + v;
                if (x < g.w && y < g.h) jpg_ycc_to_rgb(Y[(size_t)y * g.pw + x], up[0][v][u], up[1][v][u], dst + ((size_t)y * g.w + x) * 3);
            }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
static JpgGeom jpg_geom(size_t w, size_t h, int ss) {
    JpgGeom g;
    const size_t m = ss ? 16 : 8;
    g.w = (int)w; g.h = (int)h;
    g.pw = (int)((w + m - 1) / m * m); g.ph = (int)((h + m - 1) / m * m);
    g.cw = ss ? (int)((w + 1) / 2) : (int)w; g.ch = ss ? (int)((h + 1) / 2) : (int)h;
    g.cpw = ss ? g.pw / 2 : g.pw; g.cph = ss ? g.ph / 2 : g.ph;
    g.plane_stride = ((size_t)g.pw * g.ph + 255) & ~size_t(255);
    return g;
}

size_t jpeg_workspace_bytes(size_t n_ref, size_t n_q, size_t w, size_t h, int ss) {
    const JpgGeom g = jpg_geom(w, h, ss);
    return (n_ref + n_ref * n_q) * 3 * g.plane_stride + 4096;
}

// d_refs: [n_ref][h][w][3] -> d_out: [n_ref * n_q][h][w][3], image r * n_q + k = reference r at qualities[k].
// Temporaries come from the arena (the caller resets / releases it); everything is stream-ordered, no host sync.
void jpeg_roundtrip_run(Context& c, const uint8_t* d_refs, size_t n_ref, size_t w, size_t h, const int* qualities, size_t n_q,
                        int ss, uint8_t* d_out) {
    const JpgGeom g = jpg_geom(w, h, ss);
    if (n_ref * n_q > 65535) throw CudaError("jpeg sweep sub-batch too large for one launch");
    uint8_t* ycc = c.arena.alloc<uint8_t>(n_ref * 3 * g.plane_stride);
    uint8_t* rt = c.arena.alloc<uint8_t>(n_ref * n_q * 3 * g.plane_stride);
    if (n_q > JPG_MAXQ) throw CudaError("at most 32 quality levels per call");
    JpgQualities ql;
    for (size_t k = 0; k < JPG_MAXQ; k++) ql.q[k] = k < n_q ? qualities[k] : 1;
    const dim3 blk(32, 8);
    const double in_bytes = (double)n_ref * w * h * 3, plane_bytes = (double)g.pw * g.ph + 2.0 * g.cpw * g.cph;
    if (ss) {
        dim3 grid(cdiv(g.cpw, 32), cdiv(g.cph, 8), (unsigned)n_ref);
        CE_LAUNCH(c, "k_jpg_ycc", in_bytes + n_ref * plane_bytes, k_jpg_ycc<2><<<grid, blk, 0, c.stream>>>(d_refs, g, ycc));
    } else {
        dim3 grid(cdiv(g.pw, 32), cdiv(g.ph, 8), (unsigned)n_ref);
        CE_LAUNCH(c, "k_jpg_ycc", in_bytes + n_ref * plane_bytes, k_jpg_ycc<0><<<grid, blk, 0, c.stream>>>(d_refs, g, ycc));
    }
    {
        const unsigned nblk = (unsigned)((g.pw / 8) * (g.ph / 8));
        dim3 grid(cdiv(nblk, 128), 3, (unsigned)(n_ref * n_q));
        CE_LAUNCH(c, "k_jpg_dct", 2.0 * n_ref * n_q * plane_bytes, k_jpg_dct<<<grid, 128, 0, c.stream>>>(ycc, g, ql, (int)n_q, rt));
    }
    if (ss) {
        dim3 grid(cdiv(g.cw, 32), cdiv(g.ch, 8), (unsigned)(n_ref * n_q));
        CE_LAUNCH(c, "k_jpg_rgb", (double)n_ref * n_q * (plane_bytes + (double)w * h * 3), k_jpg_rgb<2><<<grid, blk, 0, c.stream>>>(rt, g, d_out));
    } else {
        dim3 grid(cdiv(g.w, 32), cdiv(g.h, 8), (unsigned)(n_ref * n_q));
        CE_LAUNCH(c, "k_jpg_rgb", (double)n_ref * n_q * (plane_bytes + (double)w * h * 3), k_jpg_rgb<0><<<grid, blk, 0, c.stream>>>(rt, g, d_out));
    }
    CE_CUDA(cudaGetLastError());
}

}  // namespace ce
