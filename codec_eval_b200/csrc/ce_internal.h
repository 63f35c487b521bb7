// ce_internal.h -- host-side internals of libce_gpu: context, workspace arena,
// and the launchers each kernel file exports.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include "../../include/ce_gpu.h"
#include "ce_copy_pool.h"

namespace ce {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
struct OomError : std::runtime_error {
    explicit OomError(const std::string& s) : std::runtime_error(s) {}
};

#define CE_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            throw ::ce::CudaError(_b);                                                           \
        }                                                                                        \
    } while (0)

// bump allocator over one cudaMalloc'd workspace
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, high = 0;
    template <class T>
    T* alloc(size_t n) {
        size_t o = (off + 255) & ~size_t(255);
        size_t bytes = n * sizeof(T);
        if (o + bytes > cap) throw OomError("workspace exhausted");
        off = o + bytes;
        if (off > high) high = off;
        return reinterpret_cast<T*>(base + o);
    }
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
    void reset() { off = 0; }
};

// recursive-Gaussian coefficients (sigma 1.5), fp32
struct RGaussCoef {
    float mul_in[3], mul_prev[3], mul_prev2[3];
};

// truncated-Gaussian kernel for Butteraugli blurs
struct BaKernel {
    int radius;
    float w[33];
};

// per-kernel CUDA-event timing (opt-in; bench.py's roofline numbers come from here)
struct KernelStat {
    std::string name;
    uint64_t launches = 0;
    double ms = 0.0;      // summed device time of the timed launches
    double bytes = 0.0;   // summed ALGORITHMIC bytes: each distinct input element once + each output element once
                          // (planes of a reference shared by several pairs count once per reference)
    double bytes_pp = 0.0;  // the same with shared reference planes charged once per PAIR (the no-reuse figure)
};
struct Profiler {
    bool enabled = false;
    struct Pending { int stat; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, int> index;
    std::vector<KernelStat> stats;
    int cur = -1;
    cudaEvent_t cur_a = nullptr;
};

struct Context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    Arena arena;
    float* d_lut = nullptr;       // 256-entry sRGB->linear table
    void* h_pinned = nullptr;     // pinned result staging
    size_t h_pinned_bytes = 0;
    void* d_results = nullptr;    // device result staging
    size_t d_results_bytes = 0;
    uint8_t* d_in = nullptr;      // device input staging for host-pointer entries
    size_t d_in_bytes = 0;
    // double-buffered staging of ce_evaluate_batch: chunk k+1 is copied on copy_stream while chunk k computes
    uint8_t* d_stage[2] = {nullptr, nullptr};
    size_t d_stage_bytes[2] = {0, 0};
    uint8_t* h_stage[2] = {nullptr, nullptr};   // pinned twins of d_stage, used only for pageable caller memory
    size_t h_stage_bytes[2] = {0, 0};
    CopyPool copy_pool;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    int* h_idx = nullptr;         // pinned + device index staging (reference index tables of a sub-batch)
    int* d_idx = nullptr;
    size_t idx_cap = 0;
    RGaussCoef rg;
    std::string last_error;
    uint64_t launches = 0;
    std::map<uint64_t, float*> ba_inv_cache;  // Butteraugli border-renormalisation tables
    int sm_count = 148;

    // Small sub-batches (a single pair, a handful of thumbnails) cannot fill 148 SMs with one metric's kernels, so
    // there the three perceptual metrics run concurrently on the main stream + two side streams.  Large sub-batches
    // run the metrics back to back on the main stream (measured round 1: 17.26 ms forked vs 17.09 ms serialised on
    // the 192-pair batch) and reuse one workspace region, which doubles the pairs per sub-batch.
    // fork_mode: -1 = auto (fork when the sub-batch has at most FORK_MAX_PIXELS pixel-pairs), 0 = never, 1 = always
    // (env CE_FORK).  Never while the per-kernel profiler is on, so its event pairs bracket one kernel each.
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    int fork_mode = -1;
    static constexpr size_t FORK_MAX_PIXELS = (size_t)4 << 20;

    Profiler prof;
    void prof_begin(const char* name, double bytes, double bytes_per_pair);
    void prof_end();
    void prof_collect();   // call after the stream is synchronised
    void ensure_input(size_t bytes);
    void ensure_stage(int slot, size_t bytes);
    void ensure_host_stage(int slot, size_t bytes);
    void ensure_results(size_t bytes);
    void ensure_idx(size_t count);
};

// every kernel launch of the library goes through this: counts it, and (when profiling) brackets it with events
#define CE_LAUNCH(ctx, name, bytes, ...)                           \
    do {                                                           \
        (ctx).prof_begin(name, (double)(bytes), (double)(bytes));  \
        __VA_ARGS__;                                               \
        (ctx).prof_end();                                          \
    } while (0)
// pair kernels that read planes of a shared reference: bytes = distinct elements, bytes_pp = charged per pair
#define CE_LAUNCH_SHARED(ctx, name, bytes, bytes_pp, ...)             \
    do {                                                              \
        (ctx).prof_begin(name, (double)(bytes), (double)(bytes_pp));  \
        __VA_ARGS__;                                                  \
        (ctx).prof_end();                                             \
    } while (0)

// TMA descriptor of fp32 planes [nplanes][h][w] as a rank-3 tensor with box (bw, bh, bz) and zero fill outside
// (cuTensorMapEncodeTiled through the runtime's driver entry point; libcuda is not linked).  False when the shape
// cannot be described (w % 4 != 0: the row stride must be a multiple of 16 bytes) -- callers then use cp.async tiles.
bool tma_enabled(int group);   // 0 Malta, 1 SSIMULACRA2 column pass, 2 wide blur, 3 fused 2-D blurs, 4 SSIMULACRA2 row pass, 5 opsin, 6 DSSIM chroma blur, 7 DSSIM statistics rows
// swizzle128: 128-byte swizzle (box rows of exactly 128 bytes; the destination must be 1024-byte aligned): 16-byte chunk j
// of box row r lands at chunk j ^ (r & 7), which makes lane = row accesses bank-conflict free without padding.
bool tma_plane_map(CUtensorMap* m, const float* base, size_t w, size_t h, size_t nplanes, unsigned bw, unsigned bh, unsigned bz,
                   bool swizzle128 = false);

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---------------- k_color.cu ----------------
// RGB8 interleaved [n_img][h*w][3] -> planar fp32 [n_img][3][h*w] through the LUT
// src_index (device, nullable): output image i is converted from source image src_index[i]
void launch_srgb8_to_linear(Context& c, const uint8_t* d_rgb, const int* src_index, size_t n_img, size_t npix, float* d_planes);
// exact integer SSE per pair: d_sse[n] (zeroed inside)
// ref_index (device, nullable): pair i compares reference image ref_index[i] (identity when null)
// n_distinct_ref: distinct reference images behind ref_index (byte accounting only; 0 = n)
void launch_sse(Context& c, const uint8_t* d_ref, const int* ref_index, const uint8_t* d_dist, size_t n, size_t bytes_per_img,
                unsigned long long* d_sse, size_t n_distinct_ref = 0);
void launch_xyb_roundtrip(Context& c, const uint8_t* d_rgb, size_t npix_total, uint8_t* d_out);
void launch_rgb8_to_rgba_linear(Context& c, const uint8_t* d_in, size_t npix, int in_channels, float* d_out);
void launch_rgba_to_planar(Context& c, const float* d_rgba, size_t w, size_t h, size_t stride, float* d_planes3, float* d_alpha);

// Image convention of the three perceptual metrics: one array of NI = R + B images, the R distinct
// references of the sub-batch first, then the B distorted images; pair b compares image ridx[b] (< R)
// with image R + b.  ridx is a device array.  Reference-side work is done once per distinct reference.

// ---------------- k_ssim2.cu ----------------
// lin: [NI][3][h*w] linear planes; d_sums: [B][6][18] doubles.  Returns number of scales.
int ssim2_run(Context& c, const float* lin, size_t R, const int* ridx, size_t B, size_t w, size_t h, double* d_sums,
              float* dbg_planes /* nullable, B==1: 3*7*h*w */);
size_t ssim2_workspace_per_pair(size_t w, size_t h);
void ssim2_init(Context& c);  // uploads the recursive-Gaussian coefficients

// ---------------- k_dssim.cu ----------------
// lin: [NI][3][h*w]; alpha nullable [NI][h*w]; d_out: [B][5][2] doubles (sum_map, sum_absdev)
int dssim_run(Context& c, const float* lin, const float* alpha, size_t R, const int* ridx, size_t B, size_t w, size_t h,
              double* d_out, float* dbg_map0 /* nullable */);
size_t dssim_workspace_per_pair(size_t w, size_t h);
int dssim_num_scales(size_t w, size_t h, size_t* ws, size_t* hs);

// ---------------- k_butteraugli.cu ----------------
// lin: [NI][3][h*w]; d_out: [B][4] doubles (max, sum d^3, sum d^6, sum d^12)
void butteraugli_run(Context& c, const float* lin, size_t R, const int* ridx, size_t B, size_t w, size_t h, float intensity,
                     double* d_out, float* dbg_diffmap /* nullable, B==1 */);
size_t butteraugli_workspace_per_pair(size_t w, size_t h);
void butteraugli_init(Context& c);  // uploads the blur kernels
void butteraugli_debug_psycho(Context& c, const float* lin, size_t w, size_t h, float intensity, float* d_planes10);
void butteraugli_debug_opsin(Context& c, const float* lin, size_t w, size_t h, float intensity, float* d_planes3);
void butteraugli_debug_blur(Context& c, const float* in, size_t w, size_t h, float sigma, float* out);

// ---------------- k_jpeg.cu ----------------
// baseline-JPEG sample-domain round trip: d_refs [n_ref][h][w][3] -> d_out [n_ref * n_q][h][w][3]; temporaries from
// the arena; stream-ordered (no host synchronisation); at most 32 qualities per call
void jpeg_roundtrip_run(Context& c, const uint8_t* d_refs, size_t n_ref, size_t w, size_t h, const int* qualities, size_t n_q,
                        int subsampling, uint8_t* d_out);
size_t jpeg_workspace_bytes(size_t n_ref, size_t n_q, size_t w, size_t h, int subsampling);

// ---------------- k_icc.cu ----------------
// matrix/TRC ICC profile -> sRGB, d_rgb [npix][3] -> d_out [npix][3]; throws std::runtime_error with the reason when
// the profile is not a usable matrix/TRC RGB profile
void icc_to_srgb_run(Context& c, const uint8_t* d_rgb, size_t npix, const uint8_t* icc, size_t icc_len, uint8_t* d_out);
bool icc_is_usable(const uint8_t* icc, size_t icc_len, std::string* why);

}  // namespace ce
