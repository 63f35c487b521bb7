// k_color.cu -- pointwise byte kernels: sRGB8 -> linear planes, integer SSE (PSNR),
// XYB u8 round-trip.  All HBM-bound streaming kernels: 128-bit loads/stores,
// grid sized in multiples of the SM count.
#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

// ---------------------------------------------------------------------------
// sRGB8 interleaved -> 3 linear fp32 planes (shared by SSIMULACRA2 / DSSIM /
// Butteraugli).  Reference: src/metrics/dssim.rs:77-85 (per byte powf); here a
// 256-entry table holding exactly those fp32 values.
// A warp converts 512 consecutive pixels in four rounds of 128: lane l takes pixels 4l .. 4l+3 of the round, i.e. three
// 32-bit loads at a 12-byte lane stride (the warp reads 384 contiguous bytes, through L1) and one 128-bit store per
// plane, so that every store instruction of the warp writes 512 contiguous bytes.  (Round 1 gave a thread 16 consecutive
// pixels: its four stores per plane were 64 bytes apart across lanes, 16 lines per instruction instead of 4.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_srgb8_to_linear_v16(const uint8_t* __restrict__ rgb, const int* __restrict__ src_index,
                                                              const float* __restrict__ lut, size_t n_groups,
                                                              size_t groups_per_img, size_t npix,
                                                              float* __restrict__ planes) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const size_t total_px = n_groups * 16;
    const int lane = threadIdx.x & 31;
    (void)groups_per_img;
    for (size_t wbase = (blockIdx.x * (size_t)blockDim.x + threadIdx.x - lane) * 16; wbase < total_px;
         wbase += (size_t)gridDim.x * blockDim.x * 16) {
        size_t img = wbase / npix;
        size_t i = wbase - img * npix + (size_t)lane * 4;
#pragma unroll
        for (int q = 0; q < 4; q++, i += 128) {
            while (i >= npix) { i -= npix; img++; }   // npix % 16 == 0: the 4 pixels of a lane never straddle images
            if (img * npix + i >= total_px) break;
            const size_t simg = src_index ? (size_t)src_index[img] : img;   // output image `img` reads source image simg
            const uint32_t* src = reinterpret_cast<const uint32_t*>(rgb + (simg * npix + i) * 3);
            const uint32_t wds[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
            float ch[3][4];
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const uint32_t byte = (wds[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                ch[k % 3][k / 3] = s_lut[byte];
            }
            float* base = planes + img * 3 * npix + i;
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
                *reinterpret_cast<float4*>(base + (size_t)cc * npix) = make_float4(ch[cc][0], ch[cc][1], ch[cc][2], ch[cc][3]);
        }
    }
}

// generic fallback for npix % 16 != 0: one thread per pixel
__global__ void __launch_bounds__(256) k_srgb8_to_linear_px(const uint8_t* __restrict__ rgb, const int* __restrict__ src_index,
                                                             const float* __restrict__ lut, size_t n_total, size_t npix,
                                                             float* __restrict__ planes) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n_total; p += (size_t)gridDim.x * blockDim.x) {
        size_t img = p / npix, i = p - img * npix;
        const size_t simg = src_index ? (size_t)src_index[img] : img;
        const uint8_t* q = rgb + (simg * npix + i) * 3;
        float* base = planes + img * 3 * npix + i;
        base[0] = s_lut[q[0]];
        base[npix] = s_lut[q[1]];
        base[2 * npix] = s_lut[q[2]];
    }
}

void launch_srgb8_to_linear(Context& c, const uint8_t* d_rgb, const int* src_index, size_t n_img, size_t npix, float* d_planes) {
    if (n_img == 0 || npix == 0) return;
    const int wave = c.sm_count * 8;
    if (npix % 16 == 0 && (reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0) {
        size_t gpi = npix / 16, ng = gpi * n_img;
        unsigned blocks = (unsigned)std::min<size_t>(cdiv(ng, 256), (size_t)wave * 4);
        CE_LAUNCH(c, "k_srgb8_to_linear_v16", n_img * npix * 15,
                  k_srgb8_to_linear_v16<<<blocks, 256, 0, c.stream>>>(d_rgb, src_index, c.d_lut, ng, gpi, npix, d_planes));
    } else {
        size_t nt = npix * n_img;
        unsigned blocks = (unsigned)std::min<size_t>(cdiv(nt, 256), (size_t)wave * 4);
        CE_LAUNCH(c, "k_srgb8_to_linear_px", n_img * npix * 15,
                  k_srgb8_to_linear_px<<<blocks, 256, 0, c.stream>>>(d_rgb, src_index, c.d_lut, nt, npix, d_planes));
    }
    CE_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// PSNR: exact integer sum of squared byte differences per pair.
// Reference: calculate_psnr, src/metrics/mod.rs:312-331 (f64 loop; integer-valued).
// grid = (chunks, pairs).  16 B per load per image, |a-b| by vabsdiffu4, squares
// accumulated with dp4a into u32 (flushed to u64 before it can overflow), warp
// shuffle, one u64 atomicAdd per block (integer => order independent => exact).
// ---------------------------------------------------------------------------
CE_DEVINL uint32_t sq4(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc);
}

__global__ void __launch_bounds__(256) k_sse(const uint8_t* __restrict__ ref, const int* __restrict__ ref_index,
                                              const uint8_t* __restrict__ dist, size_t bytes_per_img,
                                              unsigned long long* __restrict__ out) {
    const size_t pair = blockIdx.y;
    const uint8_t* r = ref + (ref_index ? (size_t)ref_index[pair] : pair) * bytes_per_img;
    const uint8_t* d = dist + pair * bytes_per_img;
    // head bytes until 16-B alignment (ref and dist arrays share the same offset modulo 16
    // only if both bases are 16-B aligned; checked by the launcher)
    size_t mis = (16 - (reinterpret_cast<uintptr_t>(r) & 15)) & 15;
    if (mis > bytes_per_img) mis = bytes_per_img;
    size_t nvec = (bytes_per_img - mis) / 16;
    size_t tail0 = mis + nvec * 16;
    unsigned long long total = 0;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (tid < mis) {
        int df = (int)r[tid] - (int)d[tid];
        total += (unsigned long long)(df * df);
    }
    if (tid < bytes_per_img - tail0) {
        size_t i = tail0 + tid;
        int df = (int)r[i] - (int)d[i];
        total += (unsigned long long)(df * df);
    }
    const uint4* rv = reinterpret_cast<const uint4*>(r + mis);
    const uint4* dv = reinterpret_cast<const uint4*>(d + mis);
    uint32_t acc = 0;
    int since_flush = 0;
    size_t i = tid;
    // 4 independent 16-B loads per image in flight
    for (; i + 3 * nthreads < nvec; i += 4 * nthreads) {
        uint4 a0 = ldg_stream_u4(rv + i), a1 = ldg_stream_u4(rv + i + nthreads), a2 = ldg_stream_u4(rv + i + 2 * nthreads),
              a3 = ldg_stream_u4(rv + i + 3 * nthreads);
        uint4 b0 = ldg_stream_u4(dv + i), b1 = ldg_stream_u4(dv + i + nthreads), b2 = ldg_stream_u4(dv + i + 2 * nthreads),
              b3 = ldg_stream_u4(dv + i + 3 * nthreads);
        acc = sq4(a0.x, b0.x, acc); acc = sq4(a0.y, b0.y, acc); acc = sq4(a0.z, b0.z, acc); acc = sq4(a0.w, b0.w, acc);
        acc = sq4(a1.x, b1.x, acc); acc = sq4(a1.y, b1.y, acc); acc = sq4(a1.z, b1.z, acc); acc = sq4(a1.w, b1.w, acc);
        acc = sq4(a2.x, b2.x, acc); acc = sq4(a2.y, b2.y, acc); acc = sq4(a2.z, b2.z, acc); acc = sq4(a2.w, b2.w, acc);
        acc = sq4(a3.x, b3.x, acc); acc = sq4(a3.y, b3.y, acc); acc = sq4(a3.z, b3.z, acc); acc = sq4(a3.w, b3.w, acc);
        // 64 bytes * 65025 = 4.2e6 per iteration; flush every 512 iterations (< 2^32)
        if (++since_flush == 512) { total += acc; acc = 0; since_flush = 0; }
    }
    for (; i < nvec; i += nthreads) {
        uint4 a = ldg_stream_u4(rv + i), b = ldg_stream_u4(dv + i);
        acc = sq4(a.x, b.x, acc); acc = sq4(a.y, b.y, acc); acc = sq4(a.z, b.z, acc); acc = sq4(a.w, b.w, acc);
    }
    total += acc;
    total = warp_sum_u64(total);
    __shared__ unsigned long long s[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s[warp] = total;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < (blockDim.x >> 5) ? s[lane] : 0ull;
        v = warp_sum_u64(v);
        if (lane == 0 && v) atomicAdd(out + pair, v);
    }
}

// scalar kernel for the (never expected) case of mutually misaligned arrays
__global__ void __launch_bounds__(256) k_sse_scalar(const uint8_t* __restrict__ ref, const int* __restrict__ ref_index,
                                                     const uint8_t* __restrict__ dist, size_t bytes_per_img,
                                                     unsigned long long* __restrict__ out) {
    const size_t pair = blockIdx.y;
    const uint8_t* r = ref + (ref_index ? (size_t)ref_index[pair] : pair) * bytes_per_img;
    const uint8_t* d = dist + pair * bytes_per_img;
    unsigned long long total = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < bytes_per_img; i += (size_t)gridDim.x * blockDim.x) {
        int df = (int)r[i] - (int)d[i];
        total += (unsigned long long)(df * df);
    }
    total = warp_sum_u64(total);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(out + pair, total);
}

void launch_sse(Context& c, const uint8_t* d_ref, const int* ref_index, const uint8_t* d_dist, size_t n, size_t bytes_per_img,
                unsigned long long* d_sse, size_t n_distinct_ref) {
    if (n == 0) return;
    if (n_distinct_ref == 0 || n_distinct_ref > n) n_distinct_ref = n;
    CE_CUDA(cudaMemsetAsync(d_sse, 0, n * sizeof(unsigned long long), c.stream));
    if (bytes_per_img == 0) return;
    // each thread should see >= 4 vectors; cap chunks so that grid ~ a few waves
    size_t nvec = bytes_per_img / 16;
    unsigned chunks = (unsigned)std::max<size_t>(1, std::min<size_t>(cdiv(nvec, 256 * 8), 1024));
    // keep total blocks >= 4 waves where possible but bounded
    while ((size_t)chunks * n > (size_t)c.sm_count * 64 && chunks > 1) chunks = (chunks + 1) / 2;
    for (size_t p0 = 0; p0 < n; p0 += 65535) {
        unsigned np = (unsigned)std::min<size_t>(65535, n - p0);
        dim3 grid(chunks, np);
        const uint8_t* r = ref_index ? d_ref : d_ref + p0 * bytes_per_img;
        const int* ri = ref_index ? ref_index + p0 : nullptr;
        const uint8_t* d = d_dist + p0 * bytes_per_img;
        // the vector kernel needs every ref / dist image pair to share its offset modulo 16
        bool same_align = ((reinterpret_cast<uintptr_t>(d_ref) ^ reinterpret_cast<uintptr_t>(d_dist)) & 15) == 0 &&
                          (bytes_per_img % 16 == 0 || !ref_index);
        const double nr = (double)std::min<size_t>(np, n_distinct_ref);   // distinct references behind this launch
        if (same_align)
            CE_LAUNCH_SHARED(c, "k_sse", (np + nr) * bytes_per_img + np * 8.0, (double)np * (2 * bytes_per_img + 8),
                             k_sse<<<grid, 256, 0, c.stream>>>(r, ri, d, bytes_per_img, d_sse + p0));
        else
            CE_LAUNCH_SHARED(c, "k_sse_scalar", (np + nr) * bytes_per_img + np * 8.0, (double)np * (2 * bytes_per_img + 8),
                             k_sse_scalar<<<grid, 256, 0, c.stream>>>(r, ri, d, bytes_per_img, d_sse + p0));
    }
    CE_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// XYB u8 round-trip, src/metrics/xyb.rs:225-253.  Pointwise; transcendental work
// evaluated in double and rounded once (= correctly rounded cbrtf / powf), see
// DESIGN.md "libm".  One thread per pixel; ALU/SFU-bound pre-pass (reference only).
// ---------------------------------------------------------------------------
CE_DEVINL float cr_cbrtf(float v) { return (float)cbrt((double)v); }
CE_DEVINL float cr_powf(float v, float e) { return (float)pow((double)v, (double)e); }
CE_DEVINL float mixed_cbrt(float v) { return v < 0.0f ? -cr_cbrtf(-v) : cr_cbrtf(v); }
CE_DEVINL float mixed_cube(float v) {
    if (v < 0.0f) { float a = -v; return -((a * a) * a); }
    return (v * v) * v;
}
CE_DEVINL float quantize_to_u8(float value, float mn, float mx) {
    float range = mx - mn;
    float normalized = (value - mn) / range;
    float r = roundf(normalized * 255.0f);
    r = fminf(fmaxf(r, 0.0f), 255.0f);
    float q = r / 255.0f;
    return q * range + mn;
}
CE_DEVINL float xyb_srgb_to_linear(float v) {
    if (v <= 0.04045f) return v / 12.92f;
    return cr_powf((v + 0.055f) / 1.055f, 2.4f);
}
CE_DEVINL uint8_t linear_to_srgb_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    float s = v <= 0.0031308f ? v * 12.92f : 1.055f * cr_powf(v, 1.0f / 2.4f) - 0.055f;
    return (uint8_t)roundf(s * 255.0f);
}

__global__ void __launch_bounds__(256) k_xyb_roundtrip(const uint8_t* __restrict__ rgb, size_t npix,
                                                        uint8_t* __restrict__ out) {
    // forward sRGB->linear only takes 256 distinct inputs: tabulate once per block
    __shared__ float s_lin[256];
    s_lin[threadIdx.x] = xyb_srgb_to_linear((float)threadIdx.x / 255.0f);
    __syncthreads();
    const float M[9] = {0.30f, 0.622f, 0.078f, 0.23f, 0.692f, 0.078f, 0.24342269f, 0.20476744f, 0.55180987f};
    const float BIAS = 0.0037930733f, NEGB = -0.15595412f;
    const float INV[9] = {11.031567f, -9.866944f, -0.164623f, -3.254147f, 4.41877f, -0.164623f, -3.658851f, 2.712923f, 1.945928f};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        float r = s_lin[rgb[3 * i]], g = s_lin[rgb[3 * i + 1]], b = s_lin[rgb[3 * i + 2]];
        float o_r = M[0] * r + M[1] * g + M[2] * b + BIAS;
        float o_g = M[3] * r + M[4] * g + M[5] * b + BIAS;
        float o_b = M[6] * r + M[7] * g + M[8] * b + BIAS;
        float c_r = mixed_cbrt(o_r) + NEGB, c_g = mixed_cbrt(o_g) + NEGB, c_b = mixed_cbrt(o_b) + NEGB;
        float x = 0.5f * (c_r - c_g), y = 0.5f * (c_r + c_g);
        float xq = quantize_to_u8(x, -0.016f, 0.029f);
        float yq = quantize_to_u8(y, 0.0f, 0.846f);
        float bq = quantize_to_u8(c_b, 0.0f, 0.846f);
        float d_r = (yq + xq) - NEGB, d_g = (yq - xq) - NEGB, d_b = bq - NEGB;
        float p_r = mixed_cube(d_r) - BIAS, p_g = mixed_cube(d_g) - BIAS, p_b = mixed_cube(d_b) - BIAS;
        float lr = INV[0] * p_r + INV[1] * p_g + INV[2] * p_b;
        float lg = INV[3] * p_r + INV[4] * p_g + INV[5] * p_b;
        float lb = INV[6] * p_r + INV[7] * p_g + INV[8] * p_b;
        out[3 * i] = linear_to_srgb_u8(lr);
        out[3 * i + 1] = linear_to_srgb_u8(lg);
        out[3 * i + 2] = linear_to_srgb_u8(lb);
    }
}

void launch_xyb_roundtrip(Context& c, const uint8_t* d_rgb, size_t npix_total, uint8_t* d_out) {
    if (npix_total == 0) return;
    unsigned blocks = (unsigned)std::min<size_t>(cdiv(npix_total, 256), (size_t)c.sm_count * 32);
    CE_LAUNCH(c, "k_xyb_roundtrip", npix_total * 6, k_xyb_roundtrip<<<blocks, 256, 0, c.stream>>>(d_rgb, npix_total, d_out));
    CE_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// rgb8_to_dssim_image / rgba8_to_dssim_image (src/metrics/dssim.rs:102-114,131-143):
// RGB8 / RGBA8 -> linear RGBA f32 interleaved.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rgb8_to_rgba_linear(const uint8_t* __restrict__ in,
                                                              const float* __restrict__ lut, size_t npix, int ch,
                                                              float4* __restrict__ out) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t* p = in + i * ch;
        float a = ch == 4 ? (float)p[3] / 255.0f : 1.0f;
        out[i] = make_float4(s_lut[p[0]], s_lut[p[1]], s_lut[p[2]], a);
    }
}
void launch_rgb8_to_rgba_linear(Context& c, const uint8_t* d_in, size_t npix, int in_channels, float* d_out) {
    if (npix == 0) return;
    unsigned blocks = (unsigned)std::min<size_t>(cdiv(npix, 256), (size_t)c.sm_count * 32);
    CE_LAUNCH(c, "k_rgb8_to_rgba_linear", npix * (in_channels + 16),
              k_rgb8_to_rgba_linear<<<blocks, 256, 0, c.stream>>>(d_in, c.d_lut, npix, in_channels, reinterpret_cast<float4*>(d_out)));
    CE_CUDA(cudaGetLastError());
}

// linear RGBA f32 interleaved (row stride in pixels) -> 3 planes [r,g,b][h*w] + alpha plane [h*w]
__global__ void __launch_bounds__(256) k_rgba_to_planar(const float4* __restrict__ in, size_t w, size_t h, size_t stride,
                                                         float* __restrict__ planes, float* __restrict__ alpha) {
    size_t n = w * h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / w, x = i - y * w;
        float4 v = in[y * stride + x];
        planes[i] = v.x;
        planes[n + i] = v.y;
        planes[2 * n + i] = v.z;
        alpha[i] = v.w;
    }
}
void launch_rgba_to_planar(Context& c, const float* d_rgba, size_t w, size_t h, size_t stride, float* d_planes3, float* d_alpha) {
    if (w * h == 0) return;
    unsigned blocks = (unsigned)std::min<size_t>(cdiv(w * h, 256), (size_t)c.sm_count * 32);
    CE_LAUNCH(c, "k_rgba_to_planar", w * h * 32,
              k_rgba_to_planar<<<blocks, 256, 0, c.stream>>>(reinterpret_cast<const float4*>(d_rgba), w, h, stride, d_planes3, d_alpha));
    CE_CUDA(cudaGetLastError());
}

}  // namespace ce
