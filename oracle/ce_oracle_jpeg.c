/*
 * ce_oracle_jpeg.c -- CPU oracle of the on-device distortion source.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as ce_oracle.c).
 *
 * What it restates: the sample-domain effect of a baseline JPEG encode -> decode with the IJG / libjpeg-turbo
 * defaults (what codec-iter's sweep does per quality level, crates/codec-iter/src/eval.rs:153-172, minus the
 * lossless entropy coding): RGB -> YCbCr (jccolor.c rgb_ycc_convert), optional h2v2 chroma downsampling
 * (jcsample.c h2v2_downsample, edge replication as in jcprepct.c), forward DCT (jfdctint.c, "islow"),
 * quantisation with the Annex-K tables scaled by `quality` (jcparam.c jpeg_set_quality, force_baseline;
 * jcdctmgr.c rounding), dequantisation + inverse DCT (jidctint.c), fancy h2v2 upsampling (jdsample.c
 * h2v2_fancy_upsample), YCbCr -> RGB (jdcolor.c).  All integer arithmetic.
 *
 * PINNED: tests/test_jpeg_source.py checks this file bit-for-bit against Pillow's JPEG save -> load
 * (libjpeg-turbo) for 4:4:4 and 4:2:0 on sizes from 1x1 to 768x512, qualities 5..100.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CEO_API __attribute__((visibility("default")))

static const uint8_t STD_LUM[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t STD_CHR[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                    99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline) */
CEO_API void ceo_jpeg_qtable(int chroma, int quality, uint16_t* out64) {
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    const uint8_t* base = chroma ? STD_CHR : STD_LUM;
    for (int i = 0; i < 64; i++) {
        long t = ((long)base[i] * scale + 50L) / 100L;
        if (t < 1) t = 1;
        if (t > 255) t = 255;
        out64[i] = (uint16_t)t;
    }
}

#define FIX16(x) ((int32_t)((x) * 65536.0 + 0.5))
#define FIX13(x) ((int32_t)((x) * 8192.0 + 0.5))
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static inline void rgb_to_ycc(int r, int g, int b, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    const int32_t half = 1 << 15, off = 128 << 16;
    *y = (uint8_t)((FIX16(0.29900) * r + FIX16(0.58700) * g + FIX16(0.11400) * b + half) >> 16);
    *cb = (uint8_t)((-FIX16(0.16874) * r - FIX16(0.33126) * g + FIX16(0.50000) * b + off + half - 1) >> 16);
    *cr = (uint8_t)((FIX16(0.50000) * r - FIX16(0.41869) * g - FIX16(0.08131) * b + off + half - 1) >> 16);
}
static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static inline void ycc_to_rgb(int y, int cb, int cr, uint8_t* rgb) {
    const int32_t half = 1 << 15;
    const int x = cr - 128, xb = cb - 128;
    rgb[0] = clamp8(y + ((FIX16(1.40200) * x + half) >> 16));
    rgb[1] = clamp8(y + ((-FIX16(0.34414) * xb + half - FIX16(0.71414) * x) >> 16));
    rgb[2] = clamp8(y + ((FIX16(1.77200) * xb + half) >> 16));
}

/* jfdctint.c: one 1-D pass over d[0..7] (stride s); first = row pass (results scaled up by 2^PASS1_BITS) */
static void fdct_1d(int32_t* d, int s, int first) {
    int32_t t0 = d[0] + d[7 * s], t7 = d[0] - d[7 * s], t1 = d[s] + d[6 * s], t6 = d[s] - d[6 * s];
    int32_t t2 = d[2 * s] + d[5 * s], t5 = d[2 * s] - d[5 * s], t3 = d[3 * s] + d[4 * s], t4 = d[3 * s] - d[4 * s];
    int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    const int n = first ? 11 : 15;
    if (first) { d[0] = (t10 + t11) << 2; d[4 * s] = (t10 - t11) << 2; }
    else { d[0] = DESCALE(t10 + t11, 2); d[4 * s] = DESCALE(t10 - t11, 2); }
    int32_t z1 = (t12 + t13) * FIX13(0.541196100);
    d[2 * s] = DESCALE(z1 + t13 * FIX13(0.765366865), n);
    d[6 * s] = DESCALE(z1 + t12 * (-FIX13(1.847759065)), n);
    z1 = t4 + t7;
    int32_t z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * FIX13(1.175875602);
    t4 *= FIX13(0.298631336); t5 *= FIX13(2.053119869); t6 *= FIX13(3.072711026); t7 *= FIX13(1.501321110);
    z1 *= -FIX13(0.899976223); z2 *= -FIX13(2.562915447); z3 *= -FIX13(1.961570560); z4 *= -FIX13(0.390180644);
    z3 += z5; z4 += z5;
    d[7 * s] = DESCALE(t4 + z1 + z3, n);
    d[5 * s] = DESCALE(t5 + z2 + z4, n);
    d[3 * s] = DESCALE(t6 + z2 + z3, n);
    d[s] = DESCALE(t7 + z1 + z4, n);
}
/* jidctint.c: one 1-D pass; n = descale shift (CONST_BITS - PASS1_BITS = 11 for columns, CONST_BITS + PASS1_BITS + 3 = 18 for rows) */
static void idct_1d(const int32_t* in, int si, int32_t* out, int so, int n) {
    int32_t z2 = in[2 * si], z3 = in[6 * si];
    int32_t z1 = (z2 + z3) * FIX13(0.541196100);
    int32_t tmp2 = z1 + z3 * (-FIX13(1.847759065)), tmp3 = z1 + z2 * FIX13(0.765366865);
    z2 = in[0]; z3 = in[4 * si];
    int32_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7 * si]; tmp1 = in[5 * si]; tmp2 = in[3 * si]; tmp3 = in[si];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3, z5 = (z3 + z4) * FIX13(1.175875602);
    tmp0 *= FIX13(0.298631336); tmp1 *= FIX13(2.053119869); tmp2 *= FIX13(3.072711026); tmp3 *= FIX13(1.501321110);
    z1 *= -FIX13(0.899976223); z2 *= -FIX13(2.562915447); z3 *= -FIX13(1.961570560); z4 *= -FIX13(0.390180644);
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    out[0] = DESCALE(tmp10 + tmp3, n); out[7 * so] = DESCALE(tmp10 - tmp3, n);
    out[so] = DESCALE(tmp11 + tmp2, n); out[6 * so] = DESCALE(tmp11 - tmp2, n);
    out[2 * so] = DESCALE(tmp12 + tmp1, n); out[5 * so] = DESCALE(tmp12 - tmp1, n);
    out[3 * so] = DESCALE(tmp13 + tmp0, n); out[4 * so] = DESCALE(tmp13 - tmp0, n);
}

/* quantisation round trip of one padded plane (pw, ph multiples of 8), in place */
static void plane_roundtrip(uint8_t* p, size_t pw, size_t ph, const uint16_t* qt) {
    for (size_t by = 0; by < ph; by += 8)
        for (size_t bx = 0; bx < pw; bx += 8) {
            int32_t d[64], ws[64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) d[y * 8 + x] = (int32_t)p[(by + (size_t)y) * pw + bx + (size_t)x] - 128;
            for (int y = 0; y < 8; y++) fdct_1d(d + y * 8, 1, 1);
            for (int x = 0; x < 8; x++) fdct_1d(d + x, 8, 0);
            for (int i = 0; i < 64; i++) { /* jcdctmgr.c quantize + jddctmgr dequantize */
                const int32_t qv = (int32_t)qt[i] << 3;
                int32_t t = d[i];
                if (t < 0) { t = -t; t += qv >> 1; t = t >= qv ? t / qv : 0; t = -t; }
                else { t += qv >> 1; t = t >= qv ? t / qv : 0; }
                d[i] = t * (int32_t)qt[i];
            }
            for (int x = 0; x < 8; x++) idct_1d(d + x, 8, ws + x, 8, 11);
            for (int y = 0; y < 8; y++) {
                int32_t o[8];
                idct_1d(ws + y * 8, 1, o, 1, 18);
                for (int x = 0; x < 8; x++) p[(by + (size_t)y) * pw + bx + (size_t)x] = clamp8(o[x] + 128);
            }
        }
}

static inline size_t up(size_t v, size_t m) { return (v + m - 1) / m * m; }
static inline size_t minz(size_t a, size_t b) { return a < b ? a : b; }

/* rgb [h][w][3] -> out [h][w][3]; subsampling 0 = 4:4:4, 2 = 4:2:0 (Pillow's numbering) */
CEO_API int ceo_jpeg_roundtrip(const uint8_t* rgb, size_t w, size_t h, int quality, int subsampling, uint8_t* out) {
    if (w == 0 || h == 0 || (subsampling != 0 && subsampling != 2)) return 3;
    uint16_t ql[64], qc[64];
    ceo_jpeg_qtable(0, quality, ql);
    ceo_jpeg_qtable(1, quality, qc);
    const size_t m = subsampling ? 16 : 8;
    const size_t pw = up(w, m), ph = up(h, m);
    uint8_t* Y = (uint8_t*)malloc(pw * ph);
    const size_t cw = subsampling ? (w + 1) / 2 : w, chh = subsampling ? (h + 1) / 2 : h;   /* valid chroma size */
    const size_t cpw = subsampling ? pw / 2 : pw, cph = subsampling ? up(chh, 8) : ph;
    uint8_t* Cb = (uint8_t*)malloc(cpw * cph);
    uint8_t* Cr = (uint8_t*)malloc(cpw * cph);
    /* luma (and full-resolution chroma): edge replication to the padded size */
    for (size_t y = 0; y < ph; y++)
        for (size_t x = 0; x < pw; x++) {
            const uint8_t* s = rgb + (minz(y, h - 1) * w + minz(x, w - 1)) * 3;
            uint8_t yy, cb, cr;
            rgb_to_ycc(s[0], s[1], s[2], &yy, &cb, &cr);
            Y[y * pw + x] = yy;
            if (!subsampling) { Cb[y * pw + x] = cb; Cr[y * pw + x] = cr; }
        }
    if (subsampling) {
        /* jcprepct.c: the colour buffer is padded to an even number of rows, jcsample.c pads the right edge to
         * 2 * output_cols, h2v2_downsample averages with the alternating bias 1,2,..; the DOWNSAMPLED rows are
         * then replicated up to a full iMCU */
        for (size_t j = 0; j < cph; j++)
            for (size_t i = 0; i < cpw; i++) {
                const size_t jj = minz(j, chh - 1);
                const size_t y0 = 2 * jj, y1 = minz(2 * jj + 1, h - 1), x0 = minz(2 * i, w - 1), x1 = minz(2 * i + 1, w - 1);
                int sb = 0, sr = 0;
                const size_t ys[2] = {y0, y1}, xs[2] = {x0, x1};
                for (int a = 0; a < 2; a++)
                    for (int b = 0; b < 2; b++) {
                        const uint8_t* s = rgb + (ys[a] * w + xs[b]) * 3;
                        uint8_t yy, cb, cr;
                        rgb_to_ycc(s[0], s[1], s[2], &yy, &cb, &cr);
                        sb += cb; sr += cr;
                    }
                const int bias = (i & 1) ? 2 : 1;
                Cb[j * cpw + i] = (uint8_t)((sb + bias) >> 2);
                Cr[j * cpw + i] = (uint8_t)((sr + bias) >> 2);
            }
    }
    plane_roundtrip(Y, pw, ph, ql);
    plane_roundtrip(Cb, cpw, cph, qc);
    plane_roundtrip(Cr, cpw, cph, qc);
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            int cb, cr;
            if (!subsampling) { cb = Cb[y * cpw + x]; cr = Cr[y * cpw + x]; }
            else {
                /* jdsample.c h2v2_fancy_upsample on the valid chroma area [chh][cw]; rows above / below replicate */
                const size_t j = y >> 1, i = x >> 1;
                const size_t jn = (y & 1) ? minz(j + 1, chh - 1) : (j > 0 ? j - 1 : 0);
                int v[2];
                for (int pl = 0; pl < 2; pl++) {
                    const uint8_t* c = pl ? Cr : Cb;
                    const int s_this = 3 * c[j * cpw + i] + c[jn * cpw + i];
                    if ((x & 1) == 0) {
                        if (i == 0) v[pl] = (s_this * 4 + 8) >> 4;
                        else v[pl] = (3 * s_this + (3 * c[j * cpw + i - 1] + c[jn * cpw + i - 1]) + 8) >> 4;
                    } else {
                        if (i == cw - 1) v[pl] = (s_this * 4 + 7) >> 4;
                        else v[pl] = (3 * s_this + (3 * c[j * cpw + i + 1] + c[jn * cpw + i + 1]) + 7) >> 4;
                    }
                }
                cb = v[0]; cr = v[1];
            }
            ycc_to_rgb(Y[y * pw + x], cb, cr, out + (y * w + x) * 3);
        }
    free(Y); free(Cb); free(Cr);
    return 0;
}
