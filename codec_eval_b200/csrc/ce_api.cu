// ce_api.cu -- the C ABI of libce_gpu (include/ce_gpu.h): context management, validation in
// the reference's order, sub-batch scheduling over the workspace, host-side score
// finalisation in fp64.  No CPU fallback: every compute entry needs a CUDA device.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>

#include "ce_internal.h"

using namespace ce;

struct ce_ctx {
    Context c;
};

struct ce_ref {
    ce_ctx* owner;   // identity check only (ce_reference_compare*); never dereferenced by ce_reference_destroy
    int device;      // so the handle can be destroyed after its context
    uint8_t* d_ref;  // device RGB8 (already XYB-round-tripped if the config asked for it)
    size_t width, height;
    ce_metric_config cfg;
};

static thread_local std::string g_create_error;

// ------------------------------------------------------------------ context helpers
void Context::ensure_input(size_t bytes) {
    if (bytes <= d_in_bytes) return;
    if (d_in) CE_CUDA(cudaFree(d_in));
    d_in = nullptr;
    d_in_bytes = 0;
    CE_CUDA(cudaMalloc(&d_in, bytes));
    d_in_bytes = bytes;
}
void Context::ensure_stage(int slot, size_t bytes) {
    if (bytes <= d_stage_bytes[slot]) return;
    if (d_stage[slot]) CE_CUDA(cudaFree(d_stage[slot]));
    d_stage[slot] = nullptr;
    d_stage_bytes[slot] = 0;
    CE_CUDA(cudaMalloc(&d_stage[slot], bytes));
    d_stage_bytes[slot] = bytes;
}
void Context::ensure_host_stage(int slot, size_t bytes) {
    if (bytes <= h_stage_bytes[slot]) return;
    if (h_stage[slot]) CE_CUDA(cudaFreeHost(h_stage[slot]));
    h_stage[slot] = nullptr;
    h_stage_bytes[slot] = 0;
    CE_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_stage[slot]), bytes, cudaHostAllocDefault));
    h_stage_bytes[slot] = bytes;
}

// true when the driver would have to bounce this host pointer (plain malloc / Vec<u8> memory)
static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// results of a sub-batch: one device buffer, TWO pinned halves (slot 0 / 1) so that the host can still be evaluating
// the scores of sub-batch k while sub-batch k+1's copy lands (h_results(slot))
void Context::ensure_results(size_t bytes) {
    if (bytes <= d_results_bytes) return;
    if (h_pinned) CE_CUDA(cudaFreeHost(h_pinned));
    if (d_results) CE_CUDA(cudaFree(d_results));
    h_pinned = nullptr;
    d_results = nullptr;
    h_pinned_bytes = d_results_bytes = 0;
    size_t cap = (std::max<size_t>(bytes, 1 << 20) + 255) & ~size_t(255);
    CE_CUDA(cudaMallocHost(&h_pinned, 2 * cap));
    CE_CUDA(cudaMalloc(&d_results, cap));
    h_pinned_bytes = 2 * cap;
    d_results_bytes = cap;
}

void Context::ensure_idx(size_t count) {
    if (count <= idx_cap) return;
    if (h_idx) CE_CUDA(cudaFreeHost(h_idx));
    if (d_idx) CE_CUDA(cudaFree(d_idx));
    h_idx = nullptr;
    d_idx = nullptr;
    idx_cap = 0;
    size_t cap = std::max<size_t>(count, 1 << 16);
    CE_CUDA(cudaMallocHost(&h_idx, cap * sizeof(int)));
    CE_CUDA(cudaMalloc(&d_idx, cap * sizeof(int)));
    idx_cap = cap;
}

void Context::prof_begin(const char* name, double bytes, double bytes_per_pair) {
    launches++;
    if (!prof.enabled) return;
    auto it = prof.index.find(name);
    int idx;
    if (it == prof.index.end()) {
        idx = (int)prof.stats.size();
        prof.index[name] = idx;
        KernelStat ks;
        ks.name = name;
        prof.stats.push_back(ks);
    } else idx = it->second;
    prof.stats[idx].launches++;
    prof.stats[idx].bytes += bytes;
    prof.stats[idx].bytes_pp += bytes_per_pair;
    auto get = [&]() {
        cudaEvent_t e;
        if (!prof.pool.empty()) { e = prof.pool.back(); prof.pool.pop_back(); }
        else CE_CUDA(cudaEventCreate(&e));
        return e;
    };
    prof.cur = idx;
    prof.cur_a = get();
    CE_CUDA(cudaEventRecord(prof.cur_a, stream));
}
void Context::prof_end() {
    if (!prof.enabled || prof.cur < 0) return;
    cudaEvent_t b;
    if (!prof.pool.empty()) { b = prof.pool.back(); prof.pool.pop_back(); }
    else CE_CUDA(cudaEventCreate(&b));
    CE_CUDA(cudaEventRecord(b, stream));
    prof.pending.push_back({prof.cur, prof.cur_a, b});
    prof.cur = -1;
}
void Context::prof_collect() {
    for (auto& p : prof.pending) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) prof.stats[p.stat].ms += ms;
        prof.pool.push_back(p.a);
        prof.pool.push_back(p.b);
    }
    prof.pending.clear();
}

// recursive Gaussian sigma 1.5 (libjxl CreateRecursiveGaussian), evaluated in double, stored as fp32
static void make_rgauss(RGaussCoef& rg) {
    const double sigma = 1.5;
    const double radius = round(3.2795 * sigma + 0.2546);
    const double pi_div_2r = M_PI / (2.0 * radius);
    const double omega[3] = {pi_div_2r, 3.0 * pi_div_2r, 5.0 * pi_div_2r};
    const double p_1 = 1.0 / tan(0.5 * omega[0]);
    const double p_3 = -1.0 / tan(0.5 * omega[1]);
    const double p_5 = 1.0 / tan(0.5 * omega[2]);
    const double r_1 = p_1 * p_1 / sin(omega[0]);
    const double r_3 = -p_3 * p_3 / sin(omega[1]);
    const double r_5 = p_5 * p_5 / sin(omega[2]);
    const double neg_half_sigma2 = -0.5 * sigma * sigma;
    const double recip_radius = 1.0 / radius;
    double rho[3];
    for (int i = 0; i < 3; i++) rho[i] = exp(neg_half_sigma2 * omega[i] * omega[i]) * recip_radius;
    const double D_13 = p_1 * r_3 - r_1 * p_3;
    const double D_35 = p_3 * r_5 - r_3 * p_5;
    const double D_51 = p_5 * r_1 - r_5 * p_1;
    const double recip_d13 = 1.0 / D_13;
    const double zeta_15 = D_35 * recip_d13;
    const double zeta_35 = D_51 * recip_d13;
    double A[3][3] = {{p_1, p_3, p_5}, {r_1, r_3, r_5}, {zeta_15, zeta_35, 1.0}};
    double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                 A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    double inv[3][3];
    inv[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det;
    inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
    inv[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
    inv[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det;
    inv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
    inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
    inv[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det;
    inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
    inv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
    const double gamma[3] = {1.0, radius * radius - sigma * sigma, zeta_15 * rho[0] + zeta_35 * rho[1] + rho[2]};
    for (int i = 0; i < 3; i++) {
        double beta = inv[i][0] * gamma[0] + inv[i][1] * gamma[1] + inv[i][2] * gamma[2];
        rg.mul_in[i] = (float)(-beta * cos(omega[i] * (radius + 1.0)));
        rg.mul_prev[i] = (float)(2.0 * cos(omega[i]));
        rg.mul_prev2[i] = -1.0f;
    }
}

// sRGB8 -> linear, the reference's expression (src/metrics/dssim.rs:77-85) evaluated once per code value
static void make_srgb_lut(float* lut) {
    for (int i = 0; i < 256; i++) {
        float s = (float)i / 255.0f;
        lut[i] = s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f);
    }
}

// ------------------------------------------------------------------ host-side finalisation
static const double S2_WEIGHT[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0, 0.0004371155730107379, 0.0,
    1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0,
    1.8422455520539298, 11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072, 0.9677937080626833, 0.0,
    0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0,
    0.0013648766163243398, 0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264,
    225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985,
    176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0008680556573291698, 0.0, 0.0,
    0.0, 0.0, 0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0,
    0.0004179171803251336, 0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862,
    23.19243343998926, 0.0, 95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493,
    171.2667255897307, 0.9807858872435379, 0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537};

// sums: [nscales][18] raw sums, layout c*6 + {d, d^4, art, art^4, det, det^4}
static double finalize_ssim2(const double* sums, int ns, size_t w, size_t h) {
    double avg_ssim[6][6], avg_edge[6][12];
    size_t cw = w, ch = h;
    for (int s = 0; s < ns; s++) {
        if (s > 0) { cw = (cw + 1) / 2; ch = (ch + 1) / 2; }
        const double opp = 1.0 / (double)(cw * ch);
        const double* S = sums + (size_t)s * 18;
        for (int c = 0; c < 3; c++) {
            avg_ssim[s][c * 2 + 0] = opp * S[c * 6 + 0];
            avg_ssim[s][c * 2 + 1] = sqrt(sqrt(opp * S[c * 6 + 1]));
            avg_edge[s][c * 4 + 0] = opp * S[c * 6 + 2];
            avg_edge[s][c * 4 + 1] = sqrt(sqrt(opp * S[c * 6 + 3]));
            avg_edge[s][c * 4 + 2] = opp * S[c * 6 + 4];
            avg_edge[s][c * 4 + 3] = sqrt(sqrt(opp * S[c * 6 + 5]));
        }
    }
    double ssim = 0.0;
    size_t i = 0;
    for (int c = 0; c < 3; c++)
        for (int s = 0; s < ns; s++)
            for (int n = 0; n < 2; n++) {
                ssim = fma(S2_WEIGHT[i++], fabs(avg_ssim[s][c * 2 + n]), ssim);
                ssim = fma(S2_WEIGHT[i++], fabs(avg_edge[s][c * 4 + n]), ssim);
                ssim = fma(S2_WEIGHT[i++], fabs(avg_edge[s][c * 4 + n + 2]), ssim);
            }
    ssim *= 0.9562382616834844;
    ssim = fma(6.248496625763138e-5 * ssim * ssim, ssim, fma(2.326765642916932, ssim, -0.020884521182843837 * ssim * ssim));
    if (ssim > 0.0) ssim = fma(pow(ssim, 0.6276336467831387), -10.0, 100.0);
    else ssim = 100.0;
    return ssim;
}

static double finalize_dssim(const double* ds /* [5][2] */, int ns, size_t w, size_t h, double* scale_scores) {
    static const double W[5] = {0.028, 0.197, 0.322, 0.298, 0.155};
    size_t ws[5], hs[5];
    dssim_num_scales(w, h, ws, hs);
    double ssim_sum = 0.0, weight_sum = 0.0;
    for (int s = 0; s < ns; s++) {
        double len = (double)(ws[s] * hs[s]);
        double score = 1.0 - ds[s * 2 + 1] / len;
        if (scale_scores) scale_scores[s] = score;
        ssim_sum += score * W[s];
        weight_sum += W[s];
    }
    double ssim = ssim_sum / weight_sum;
    if (ssim < 2.220446049250313e-16) ssim = 2.220446049250313e-16;
    return 1.0 / ssim - 1.0;
}

static void finalize_butteraugli(const double* ba /* max, s3, s6, s12 */, size_t w, size_t h, double* mx, double* pn) {
    double opp = 1.0 / (double)(w * h);
    *mx = ba[0];
    *pn = (pow(opp * ba[1], 1.0 / 3.0) + pow(opp * ba[2], 1.0 / 6.0) + pow(opp * ba[3], 1.0 / 12.0)) / 3.0;
}

static double finalize_psnr(uint64_t sse, size_t w, size_t h) {
    // src/metrics/mod.rs:324-330
    double mse = (double)sse / (double)(w * h * 3);
    if (mse == 0.0) return INFINITY;
    return 10.0 * log10(255.0 * 255.0 / mse);
}

// ------------------------------------------------------------------ the device batch
struct DebugOut {
    double* s2_sums = nullptr;   // [6*18]
    int* s2_nscales = nullptr;
    float* s2_planes = nullptr;  // device pointer
    double* ds_scores = nullptr; // [5]
    int* ds_nscales = nullptr;
    float* ds_map0 = nullptr;    // device pointer
    float* ba_diffmap = nullptr; // device pointer
};

// workspace bytes of one pair: `forked` = the three perceptual metrics overlap on separate streams, so their
// workspaces coexist; otherwise they run back to back and share one region (each *_run releases what it took)
static size_t per_pair_bytes(const ce_metric_config& cfg, size_t w, size_t h, bool forked) {
    size_t n = w * h;
    size_t bytes = 1024;
    if (cfg.xyb_roundtrip) bytes += n * 3 + 256;
    bool perceptual = cfg.dssim || cfg.ssimulacra2 || cfg.butteraugli;
    if (perceptual) bytes += 2 * 3 * n * 4 + 512;
    size_t sum = 0, mx = 0;
    auto add = [&](size_t b) { sum += b; mx = std::max(mx, b); };
    if (cfg.dssim) add(dssim_workspace_per_pair(w, h));
    if (cfg.ssimulacra2) add(ssim2_workspace_per_pair(w, h));
    if (cfg.butteraugli) add(butteraugli_workspace_per_pair(w, h));
    return bytes + (forked ? sum : mx);
}

static const size_t kArenaFixed = 1 << 20;
static const size_t kRawDoubles = 1 + 108 + 10 + 4;   // per pair: sse (u64), SSIMULACRA2 sums, DSSIM sums, Butteraugli

// most pairs of this size one sub-batch can hold (metrics back to back)
static size_t sub_batch_capacity(Context& c, const ce_metric_config& cfg, size_t w, size_t h) {
    const size_t ppb = per_pair_bytes(cfg, w, h, false);
    if (c.arena.cap < ppb + kArenaFixed) throw OomError("workspace too small for one pair of this size");
    // 10000 keeps every grid.z (up to 3 planes x 2 images per pair) under 65535
    return std::min<size_t>((c.arena.cap - kArenaFixed) / ppb, 10000);
}

static bool fork_metrics(const Context& c, const ce_metric_config& cfg, size_t B, size_t w, size_t h) {
    if (c.prof.enabled || c.fork_mode == 0) return false;
    if ((int)(cfg.dssim != 0) + (int)(cfg.ssimulacra2 != 0) + (int)(cfg.butteraugli != 0) < 2) return false;
    if (c.fork_mode < 0 && B * w * h > Context::FORK_MAX_PIXELS) return false;
    return B * per_pair_bytes(cfg, w, h, true) + kArenaFixed <= c.arena.cap;
}

// One sub-batch in flight: launch_sub_batch queues every kernel and the result copy on c.stream and returns without
// waiting; finish_sub_batch waits for it and evaluates the scores on the host.  Between the two the caller may
// stage the next chunk's host->device copies (ce_evaluate_batch), which is what lets a blocking copy from pageable
// memory overlap the kernels of the previous chunk.
struct SubBatch {
    size_t p0 = 0, B = 0;
    int s2_ns = 0, ds_ns = 0;
    int slot = 0;          // pinned result half
    bool small = false;
};
static inline char* h_results(Context& c, int slot) { return static_cast<char*>(c.h_pinned) + (size_t)slot * c.d_results_bytes; }

// d_ref: n_ref device images, d_dist: device images (tight RGB8, w x h).  Pair p0 + i compares reference
// ref_of[p0 + i] (host array; nullptr = identity) with distorted image p0 + i.  Reference-side work of a
// sub-batch is done once per distinct reference (fast_ssim2's Ssimulacra2Reference reuse,
// crates/codec-iter/src/eval.rs:138-149, generalised to all metrics).  local_of: scratch of n_ref ints, all -1.
static SubBatch launch_sub_batch(Context& c, const uint8_t* d_ref, const uint8_t* d_dist, const uint32_t* ref_of,
                                 std::vector<int>& local_of, size_t p0, size_t B, size_t w, size_t h,
                                 const ce_metric_config& cfg, float intensity, const DebugOut* dbg, int slot) {
    const size_t npix = w * h, img_bytes = npix * 3;
    SubBatch sb;
    sb.p0 = p0;
    sb.B = B;
    sb.small = (w < 8 || h < 8);
    sb.slot = slot;
    const bool small = sb.small;
    const bool perceptual = cfg.dssim || cfg.ssimulacra2 || cfg.butteraugli;
    c.arena.reset();
    // the caller sized the result staging for its largest sub-batch before the first launch: growing it here would
    // free the pinned half that still holds the previous sub-batch's unread results
    if (B * kRawDoubles * 8 > c.d_results_bytes) throw CudaError("internal: result staging not sized before launch");
    // ---- index tables of the sub-batch: ridx[B] local reference of each pair, uniq[R] global reference
    // image of each local one, gref[B] global reference image of each pair
    c.ensure_idx(3 * B);
    int* h_ridx = c.h_idx;
    int* h_uniq = c.h_idx + B;
    int* h_gref = c.h_idx + 2 * B;
    size_t R = 0;
    if (ref_of) {
        for (size_t i = 0; i < B; i++) {
            const uint32_t g = ref_of[p0 + i];
            if (local_of[g] < 0) { local_of[g] = (int)R; h_uniq[R++] = (int)g; }
            h_ridx[i] = local_of[g];
            h_gref[i] = (int)g;
        }
        for (size_t r = 0; r < R; r++) local_of[h_uniq[r]] = -1;
    } else {
        R = B;
        for (size_t i = 0; i < B; i++) { h_ridx[i] = (int)i; h_uniq[i] = (int)(p0 + i); h_gref[i] = (int)(p0 + i); }
    }
    CE_CUDA(cudaMemcpyAsync(c.d_idx, c.h_idx, 3 * B * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    const int* d_ridx = c.d_idx;
    const int* d_uniq = c.d_idx + B;
    const int* d_gref = c.d_idx + 2 * B;
    bool contiguous = true;
    for (size_t r = 1; r < R; r++) contiguous = contiguous && h_uniq[r] == h_uniq[r - 1] + 1;

    double* d_raw = reinterpret_cast<double*>(c.d_results);
    unsigned long long* d_sse = reinterpret_cast<unsigned long long*>(d_raw);
    double* d_s2 = d_raw + B;
    double* d_ds = d_s2 + B * 108;
    double* d_ba = d_ds + B * 10;
    const uint8_t* dist = d_dist + p0 * img_bytes;
    // reference images of this sub-batch: (ref_base, ref_src) with ref_src a device index array or
    // nullptr for "image r of ref_base"; (ref_base, sse_idx) the same per pair
    const uint8_t* ref_base = d_ref;
    const int* ref_src = d_uniq;
    const int* sse_idx = d_gref;
    if (contiguous) { ref_base = d_ref + (size_t)h_uniq[0] * img_bytes; ref_src = nullptr; sse_idx = d_ridx; }
    if (cfg.xyb_roundtrip) {  // distinct references only, before every metric (src/eval/session.rs:447-456)
        uint8_t* rt = c.arena.alloc<uint8_t>(R * img_bytes);
        for (size_t r = 0; r < R;) {   // one launch per run of consecutive source images
            size_t e = r + 1;
            while (e < R && h_uniq[e] == h_uniq[e - 1] + 1) e++;
            launch_xyb_roundtrip(c, d_ref + (size_t)h_uniq[r] * img_bytes, (e - r) * npix, rt + r * img_bytes);
            r = e;
        }
        ref_base = rt; ref_src = nullptr; sse_idx = d_ridx;
    }
    if (cfg.psnr) launch_sse(c, ref_base, sse_idx, dist, B, img_bytes, d_sse, R);
    if (perceptual) {
        const size_t NI = R + B;
        float* lin = c.arena.alloc<float>(NI * 3 * npix);
        launch_srgb8_to_linear(c, ref_base, ref_src, R, npix, lin);
        launch_srgb8_to_linear(c, dist, nullptr, B, npix, lin + R * 3 * npix);
        // fork (small sub-batches only): DSSIM and SSIMULACRA2 on the side streams, Butteraugli on the main one.
        // Each forked metric keeps its arena region (no release between them) because their kernels overlap in time.
        const bool fork = fork_metrics(c, cfg, B, w, h);
        cudaStream_t main_stream = c.stream;
        if (fork) CE_CUDA(cudaEventRecord(c.ev_fork, main_stream));
        int used = 0;
        auto on_side = [&](auto&& fn) {
            if (!fork) { fn(); return; }
            cudaStream_t s = c.side[used];
            CE_CUDA(cudaStreamWaitEvent(s, c.ev_fork, 0));
            c.stream = s;
            c.arena.high = c.arena.off;
            try { fn(); } catch (...) { c.stream = main_stream; throw; }
            c.arena.off = c.arena.high;   // keep the metric's workspace reserved until the join
            c.stream = main_stream;
            CE_CUDA(cudaEventRecord(c.ev_join[used], s));
            used++;
        };
        try {
            if (cfg.dssim) on_side([&] { sb.ds_ns = dssim_run(c, lin, nullptr, R, d_ridx, B, w, h, d_ds, dbg ? dbg->ds_map0 : nullptr); });
            if (cfg.ssimulacra2 && !small) on_side([&] { sb.s2_ns = ssim2_run(c, lin, R, d_ridx, B, w, h, d_s2, dbg ? dbg->s2_planes : nullptr); });
            if (cfg.butteraugli && !small) butteraugli_run(c, lin, R, d_ridx, B, w, h, intensity, d_ba, dbg ? dbg->ba_diffmap : nullptr);
            for (int i = 0; i < used; i++) CE_CUDA(cudaStreamWaitEvent(main_stream, c.ev_join[i], 0));
        } catch (...) {
            // kernels already queued on the side streams still use the arena: nothing may reset it before they end
            if (fork) { cudaStreamSynchronize(c.side[0]); cudaStreamSynchronize(c.side[1]); }
            cudaStreamSynchronize(main_stream);
            throw;
        }
    }
    CE_CUDA(cudaMemcpyAsync(h_results(c, slot), d_raw, B * kRawDoubles * 8, cudaMemcpyDeviceToHost, c.stream));
    return sb;
}

// wait for the sub-batch in flight (after this the workspace, the index tables and the device result buffer are free
// for the next launch; the raw results sit in the sub-batch's pinned half)
static void wait_sub_batch(Context& c) {
    CE_CUDA(cudaStreamSynchronize(c.stream));
    c.prof_collect();
}

// host-side score evaluation (fp64) of a sub-batch that has been waited for; may run while the next one computes
static void finalize_sub_batch(Context& c, const SubBatch& sb, size_t w, size_t h, const ce_metric_config& cfg, ce_result* out,
                               const DebugOut* dbg) {
    const size_t B = sb.B;
    const bool small = sb.small;
    const double* h_raw = reinterpret_cast<const double*>(h_results(c, sb.slot));
    const uint64_t* h_sse = reinterpret_cast<const uint64_t*>(h_raw);
    const double* h_s2 = h_raw + B;
    const double* h_ds = h_s2 + B * 108;
    const double* h_ba = h_ds + B * 10;
    for (size_t i = 0; i < B; i++) {
        ce_result& r = out[sb.p0 + i];
        memset(&r, 0, sizeof(r));
        // reference order: psnr, dssim, ssimulacra2, butteraugli; first failure ends the pair
        if (cfg.psnr) {
            r.sse = h_sse[i];
            r.psnr = finalize_psnr(r.sse, w, h);
            r.valid |= CE_VALID_PSNR;
        }
        if (cfg.dssim) {
            r.dssim = finalize_dssim(h_ds + i * 10, sb.ds_ns, w, h, (dbg && i == 0) ? dbg->ds_scores : nullptr);
            if (dbg && dbg->ds_nscales) *dbg->ds_nscales = sb.ds_ns;
            r.valid |= CE_VALID_DSSIM;
        }
        if (cfg.ssimulacra2) {
            if (small) { r.status = CE_ERR_METRIC_CALCULATION; continue; }
            r.ssimulacra2 = finalize_ssim2(h_s2 + i * 108, sb.s2_ns, w, h);
            if (dbg && i == 0 && dbg->s2_sums) memcpy(dbg->s2_sums, h_s2, sizeof(double) * 18 * (size_t)sb.s2_ns);
            if (dbg && dbg->s2_nscales) *dbg->s2_nscales = sb.s2_ns;
            r.valid |= CE_VALID_SSIMULACRA2;
        }
        if (cfg.butteraugli) {
            if (small) { r.status = CE_ERR_METRIC_CALCULATION; continue; }
            finalize_butteraugli(h_ba + i * 4, w, h, &r.butteraugli, &r.butteraugli_pnorm3);
            r.valid |= CE_VALID_BUTTERAUGLI;
        }
    }
    if (small && (cfg.ssimulacra2 || cfg.butteraugli))
        c.last_error = cfg.ssimulacra2 ? "SSIMULACRA2: images must be at least 8x8 pixels" : "Butteraugli: images must be at least 8x8 pixels";
}

static void run_device_batch(Context& c, const uint8_t* d_ref, size_t n_ref, const uint8_t* d_dist, size_t n,
                             const uint32_t* ref_of, size_t w, size_t h, const ce_metric_config& cfg, float intensity,
                             ce_result* out, const DebugOut* dbg) {
    const size_t Bmax = sub_batch_capacity(c, cfg, w, h);
    std::vector<int> local_of(ref_of ? n_ref : 0, -1);
    if (n == 0) return;
    c.ensure_results(std::min(Bmax, n) * kRawDoubles * 8);
    // launch k+1 as soon as k has drained, THEN evaluate k's scores on the host: the fp64 finalisation (108-term
    // polynomial, roots and pow per pair: ~0.5 ms per 262-pair sub-batch) runs under the next sub-batch's kernels
    int slot = 0;
    try {
        SubBatch cur = launch_sub_batch(c, d_ref, d_dist, ref_of, local_of, 0, std::min(Bmax, n), w, h, cfg, intensity, dbg, slot);
        for (size_t p0 = 0; p0 < n; p0 += Bmax) {
            wait_sub_batch(c);
            const SubBatch done = cur;
            const size_t q0 = p0 + Bmax;
            if (q0 < n) {
                slot ^= 1;
                cur = launch_sub_batch(c, d_ref, d_dist, ref_of, local_of, q0, std::min(Bmax, n - q0), w, h, cfg, intensity, dbg, slot);
            }
            finalize_sub_batch(c, done, w, h, cfg, out, dbg);
        }
    } catch (...) {
        cudaStreamSynchronize(c.stream);   // nothing of a failed call may still be running when the next one resets the workspace
        throw;
    }
}

// ------------------------------------------------------------------ TMA descriptors
namespace ce {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}
bool tma_enabled(int group) {   // CE_TMA_MASK (debugging): bit per kernel group, default all on
    static int mask = -2;
    if (mask == -2) {
        const char* e = getenv("CE_TMA_MASK");
        mask = e ? atoi(e) : -1;
    }
    return (mask & (1 << group)) != 0;
}
bool tma_plane_map(CUtensorMap* m, const float* base, size_t w, size_t h, size_t nplanes, unsigned bw, unsigned bh, unsigned bz,
                   bool swizzle128) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc || (w & 3) || nplanes == 0 || !base) return false;
    if (swizzle128 && bw * 4 > 128) return false;   // the swizzle span is the box row
    const cuuint64_t dims[3] = {w, h, nplanes};
    const cuuint64_t strides[2] = {w * 4, w * h * 4};
    const cuuint32_t box[3] = {bw, bh, bz};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace ce

// ------------------------------------------------------------------ C ABI
#define CE_TRY(ctx_expr, body)                                   \
    try {                                                        \
        body                                                     \
    } catch (const OomError& e) {                                \
        (ctx_expr).last_error = e.what();                        \
        return CE_ERR_OUT_OF_MEMORY;                             \
    } catch (const CudaError& e) {                               \
        (ctx_expr).last_error = e.what();                        \
        return CE_ERR_CUDA;                                      \
    } catch (const std::exception& e) {                          \
        (ctx_expr).last_error = e.what();                        \
        return CE_ERR_CUDA;                                      \
    }

extern "C" {

#ifndef CE_SOURCE_HASH
#define CE_SOURCE_HASH "unhashed"
#endif
// "... src:<sha256/16 of csrc/* + include/ce_gpu.h>": codec_eval_b200.build stamps it, _lib.load() compares it with the
// sources beside the library, so a stale binary cannot be taken for the committed code.
CE_API const char* ce_version(void) { return "ce_gpu 0.2.0 (sm_100a) src:" CE_SOURCE_HASH; }

// ---- pinned host memory for the callers' image buffers (opt-in) ---------------------------------------
CE_API int ce_host_register(ce_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr || !bytes) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        CE_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    })
    return CE_OK;
}
CE_API int ce_host_unregister(ce_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        CE_CUDA(cudaHostUnregister(ptr));
    })
    return CE_OK;
}
CE_API int ce_host_alloc(ce_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out || !bytes) return CE_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        CE_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    })
    return CE_OK;
}
CE_API void ce_host_free(ce_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return;
    cudaSetDevice(ctx->c.device);
    cudaFreeHost(ptr);
}

CE_API int ce_ctx_create(ce_ctx** out, int device, size_t workspace_bytes) {
    if (!out) return CE_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    ce_ctx* ctx = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libce_gpu has no CPU fallback)";
            return CE_ERR_CUDA;
        }
        if (device < 0 || device >= ndev) {
            g_create_error = "device index out of range";
            return CE_ERR_INVALID_ARGUMENT;
        }
        CE_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        CE_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            g_create_error = "libce_gpu is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
            return CE_ERR_CUDA;
        }
        ctx = new ce_ctx();
        Context& c = ctx->c;
        c.device = device;
        c.sm_count = prop.multiProcessorCount;
        if (const char* e = getenv("CE_FORK")) c.fork_mode = atoi(e) > 0 ? 1 : (atoi(e) == 0 ? 0 : -1);
        CE_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
        c.stream = c.own_stream;
        for (int i = 0; i < 2; i++) {
            CE_CUDA(cudaStreamCreateWithFlags(&c.side[i], cudaStreamNonBlocking));
            CE_CUDA(cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming));
        }
        CE_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        CE_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) CE_CUDA(cudaEventCreateWithFlags(&c.ev_copy[i], cudaEventDisableTiming));
        if (workspace_bytes == 0) {
            size_t free_b = 0, total_b = 0;
            CE_CUDA(cudaMemGetInfo(&free_b, &total_b));
            workspace_bytes = std::min<size_t>(free_b / 2, (size_t)64 << 30);
        }
        CE_CUDA(cudaMalloc(&c.arena.base, workspace_bytes));
        c.arena.cap = workspace_bytes;
        float lut[256];
        make_srgb_lut(lut);
        CE_CUDA(cudaMalloc(&c.d_lut, sizeof(lut)));
        CE_CUDA(cudaMemcpy(c.d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
        make_rgauss(c.rg);
        ssim2_init(c);
        butteraugli_init(c);
        c.ensure_results(1 << 20);
        *out = ctx;
        return CE_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        if (ctx) ce_ctx_destroy(ctx);
        return CE_ERR_CUDA;
    }
}

CE_API void ce_ctx_destroy(ce_ctx* ctx) {
    if (!ctx) return;
    Context& c = ctx->c;
    cudaSetDevice(c.device);
    if (c.own_stream) cudaStreamSynchronize(c.own_stream);
    if (c.arena.base) cudaFree(c.arena.base);
    if (c.d_lut) cudaFree(c.d_lut);
    if (c.h_pinned) cudaFreeHost(c.h_pinned);
    if (c.d_results) cudaFree(c.d_results);
    if (c.d_in) cudaFree(c.d_in);
    if (c.h_idx) cudaFreeHost(c.h_idx);
    if (c.d_idx) cudaFree(c.d_idx);
    for (auto& kv : c.ba_inv_cache) cudaFree(kv.second);
    c.prof_collect();
    for (auto e : c.prof.pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c.side[i]) { cudaStreamSynchronize(c.side[i]); cudaStreamDestroy(c.side[i]); }
        if (c.ev_join[i]) cudaEventDestroy(c.ev_join[i]);
    }
    if (c.ev_fork) cudaEventDestroy(c.ev_fork);
    if (c.copy_stream) { cudaStreamSynchronize(c.copy_stream); cudaStreamDestroy(c.copy_stream); }
    for (int i = 0; i < 2; i++) {
        if (c.ev_copy[i]) cudaEventDestroy(c.ev_copy[i]);
        if (c.d_stage[i]) cudaFree(c.d_stage[i]);
        if (c.h_stage[i]) cudaFreeHost(c.h_stage[i]);
    }
    if (c.own_stream) cudaStreamDestroy(c.own_stream);
    delete ctx;
}

CE_API int ce_ctx_set_stream(ce_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CE_ERR_INVALID_ARGUMENT;
    ctx->c.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->c.own_stream;
    return CE_OK;
}

CE_API const char* ce_last_error(const ce_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str(); }
CE_API uint64_t ce_launch_count(const ce_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

CE_API int ce_profile_enable(ce_ctx* ctx, int enable) {
    if (!ctx) return CE_ERR_INVALID_ARGUMENT;
    ctx->c.prof.enabled = enable != 0;
    return CE_OK;
}
CE_API int ce_profile_reset(ce_ctx* ctx) {
    if (!ctx) return CE_ERR_INVALID_ARGUMENT;
    for (auto& k : ctx->c.prof.stats) { k.launches = 0; k.ms = 0.0; k.bytes = 0.0; k.bytes_pp = 0.0; }
    return CE_OK;
}
CE_API size_t ce_profile_report(ce_ctx* ctx, char* buf, size_t cap) {
    if (!ctx) return 0;
    std::string out;
    char line[512];
    for (auto& k : ctx->c.prof.stats) {
        if (!k.launches) continue;
        snprintf(line, sizeof(line), "%s\t%llu\t%.6f\t%.0f\t%.0f\n", k.name.c_str(), (unsigned long long)k.launches, k.ms, k.bytes,
                 k.bytes_pp);
        out += line;
    }
    if (buf && cap) {
        size_t n = std::min(cap - 1, out.size());
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}

CE_API int ce_sub_batch_capacity(ce_ctx* ctx, const ce_metric_config* cfg, uint32_t width, uint32_t height, size_t* pairs) {
    if (!ctx || !cfg || !pairs || width == 0 || height == 0) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    CE_TRY(c, { *pairs = sub_batch_capacity(c, *cfg, width, height); })
    return CE_OK;
}

CE_API int ce_evaluate_batch_device(ce_ctx* ctx, const uint8_t* d_ref, const uint8_t* d_dist, size_t n, uint32_t width,
                                    uint32_t height, const ce_metric_config* cfg, float intensity_target, ce_result* out) {
    if (!ctx || !cfg || (!out && n) || ((!d_ref || !d_dist) && n)) return CE_ERR_INVALID_ARGUMENT;
    if (n == 0) return CE_OK;
    if (width == 0 || height == 0) {
        ctx->c.last_error = "zero-sized image";
        return CE_ERR_INVALID_ARGUMENT;
    }
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        run_device_batch(c, d_ref, n, d_dist, n, nullptr, width, height, *cfg, intensity_target, out, nullptr);
    })
    return CE_OK;
}

CE_API int ce_evaluate_batch_device_grouped(ce_ctx* ctx, const uint8_t* d_ref, size_t n_ref, const uint8_t* d_dist, size_t n,
                                            const uint32_t* ref_index, uint32_t width, uint32_t height,
                                            const ce_metric_config* cfg, float intensity_target, ce_result* out) {
    if (!ctx || !cfg || (n && (!out || !d_ref || !d_dist || !ref_index || n_ref == 0))) return CE_ERR_INVALID_ARGUMENT;
    if (n == 0) return CE_OK;
    if (width == 0 || height == 0) {
        ctx->c.last_error = "zero-sized image";
        return CE_ERR_INVALID_ARGUMENT;
    }
    for (size_t i = 0; i < n; i++)
        if (ref_index[i] >= n_ref) {
            ctx->c.last_error = "ref_index out of range";
            return CE_ERR_INVALID_ARGUMENT;
        }
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        run_device_batch(c, d_ref, n_ref, d_dist, n, ref_index, width, height, *cfg, intensity_target, out, nullptr);
    })
    return CE_OK;
}

// validation in the order of src/metrics/ssimulacra2.rs:65-82 / butteraugli.rs:51-68
static int validate_pair(const ce_pair& p, std::string* why) {
    if (!p.ref || !p.dist) return CE_ERR_INVALID_ARGUMENT;
    if (p.ref_len != p.dist_len) {
        if (why) *why = "reference and test buffers differ in length";
        return CE_ERR_DIMENSION_MISMATCH;
    }
    size_t expected = (size_t)p.width * p.height * 3;
    if (p.ref_len != expected || expected == 0) {
        if (why) *why = "Invalid image size: expected " + std::to_string(expected) + " bytes, got " + std::to_string(p.ref_len);
        return CE_ERR_METRIC_CALCULATION;
    }
    return CE_OK;
}

static void evaluate_host_pairs(Context& c, const ce_pair* pairs, size_t n, const ce_metric_config& cfg, float intensity,
                                ce_result* out, const DebugOut* dbg) {
    // group valid pairs by size, keeping submission order inside a group
    std::map<std::pair<uint32_t, uint32_t>, std::vector<size_t>> groups;
    for (size_t i = 0; i < n; i++) {
        memset(&out[i], 0, sizeof(ce_result));
        std::string why;
        int st = validate_pair(pairs[i], &why);
        if (st != CE_OK) {
            out[i].status = st;
            if (!why.empty()) c.last_error = why;
            continue;
        }
        groups[{pairs[i].width, pairs[i].height}].push_back(i);
    }
    std::vector<ce_result> tmp;
    struct Staged {
        size_t k0 = 0, B = 0, Ru = 0;
        std::vector<uint32_t> ref_of;
    } st[2];
    // Two pairs share a reference image when their `ref` pointer AND their `ref_id` are equal (evaluate_image compares
    // one reference with every codec x quality output, src/eval/session.rs:375-431): it is uploaded and pre-processed once.
    unsigned staging_threads = 4;   // CE_STAGING_THREADS=0 leaves pageable memory to the driver's own bounce path
    if (const char* e = getenv("CE_STAGING_THREADS")) staging_threads = (unsigned)std::max(0, atoi(e));
    staging_threads = std::min<unsigned>(staging_threads, std::max(1u, std::thread::hardware_concurrency()));
    auto same_ref = [&](size_t a, size_t b) { return pairs[a].ref == pairs[b].ref && pairs[a].ref_id == pairs[b].ref_id; };
    for (auto& g : groups) {
        const size_t w = g.first.first, h = g.first.second, img_bytes = w * h * 3;
        const std::vector<size_t>& idx = g.second;
        // The group is cut into chunks of at most one sub-batch (what the workspace holds, and at most 2 GiB of
        // distorted input): chunk k+1 is copied to the device on the copy stream while chunk k computes.  Only the first
        // chunk's copy is exposed, so the chunks ramp up -- 1/12 of the group (at most 1/8 of a sub-batch), then 1.5 times
        // the previous one until the cap.  A copy hides under the previous chunk's compute only while the growth factor
        // stays below (copy rate) / (compute rate): 55 GB/s against 4.5 pairs/ms allows 3.6 on one GPU, but with eight
        // ranks sharing the host each gets 23 GB/s and the bound is 1.5 (a factor of 3 stalled 27 ms of a 280 ms step).
        // Boundaries prefer the end of a run of pairs that share a reference.  CE_HOST_CHUNKS=1 disables the cutting.
        size_t max_chunks = 0;
        if (const char* e = getenv("CE_HOST_CHUNKS")) max_chunks = (size_t)std::max(0, atoi(e));
        const size_t cap = std::min<size_t>(sub_batch_capacity(c, cfg, w, h), std::max<size_t>(1, ((size_t)2 << 30) / std::max<size_t>(img_bytes, 1)));
        std::vector<size_t> bounds(1, 0);
        {
            const size_t total = idx.size();
            auto snap = [&](size_t b) {   // move a boundary to the end of the current same-reference run
                while (b > 0 && b < total && same_ref(idx[b], idx[b - 1])) b++;
                return b;
            };
            if (total >= 16 && max_chunks != 1) {
                size_t size = std::max<size_t>(1, std::min(std::max<size_t>(1, cap / 8), total / 12));
                auto snap_down = [&](size_t from, size_t b) {   // the last run boundary in (from, b], or b if there is none
                    size_t d = b;
                    while (d > from && d < total && same_ref(idx[d], idx[d - 1])) d--;
                    return d > from ? d : b;
                };
                for (size_t b = 0;;) {
                    const size_t from = b;
                    b = snap(from + size);
                    if (b - from > cap) b = snap_down(from, from + cap);
                    if (b >= total) break;
                    bounds.push_back(b);
                    size = std::min(cap, size + (size + 1) / 2);
                    if (total - b <= size + size / 4) break;   // no tiny tail chunk: the rest goes as one
                }
            }
            // enforce the size cap (a snapped boundary, or the tail, may have overshot it)
            std::vector<size_t> capped(1, 0);
            for (size_t k = 1; k <= bounds.size(); k++) {
                const size_t end = k < bounds.size() ? bounds[k] : total;
                while (end - capped.back() > cap) capped.push_back(capped.back() + cap);
                if (end < total) capped.push_back(end);
            }
            bounds = capped;
            bounds.push_back(total);
        }
        const size_t nchunks = bounds.size() - 1;
        auto stage = [&](size_t ci) {
            Staged& s = st[ci & 1];
            s.k0 = bounds[ci];
            s.B = bounds[ci + 1] - bounds[ci];
            std::map<std::pair<const uint8_t*, uint32_t>, uint32_t> seen;
            std::vector<const uint8_t*> urefs;
            s.ref_of.resize(s.B);
            for (size_t k = 0; k < s.B; k++) {
                const ce_pair& p = pairs[idx[s.k0 + k]];
                const auto key = std::make_pair(p.ref, p.ref_id);
                auto it = seen.find(key);
                if (it == seen.end()) { it = seen.emplace(key, (uint32_t)urefs.size()).first; urefs.push_back(p.ref); }
                s.ref_of[k] = it->second;
            }
            s.Ru = urefs.size();
            const size_t slot_bytes = (s.Ru + s.B) * img_bytes;
            c.ensure_stage((int)(ci & 1), slot_bytes);
            uint8_t* d_ref = c.d_stage[ci & 1];
            uint8_t* d_dist = d_ref + s.Ru * img_bytes;
            // Pageable caller memory (what the Rust side holds today, src/eval/session.rs:394) in a chunk worth the
            // trouble: gather the images into the slot's pinned twin with a few memcpy threads, then ONE asynchronous
            // copy.  (The previous user of this pinned slot, chunk ci-2, has completed: its copy was awaited by the
            // compute stream before chunk ci-2 ran, and chunk ci-2 has been finished.)
            bool bounce = slot_bytes >= ((size_t)8 << 20) && staging_threads > 0;
            if (bounce) {
                bounce = is_pageable(pairs[idx[s.k0]].dist) || is_pageable(urefs[0]);
                for (size_t k = 1; bounce == false && k < s.B; k += 16) bounce = is_pageable(pairs[idx[s.k0 + k]].dist);
            }
            if (bounce) {
                c.ensure_host_stage((int)(ci & 1), slot_bytes);
                c.copy_pool.start(staging_threads);
                uint8_t* hs = c.h_stage[ci & 1];
                const size_t Ru = s.Ru, k0 = s.k0;
                const std::function<void(size_t)> job = [&](size_t i) {
                    const uint8_t* src = i < Ru ? urefs[i] : pairs[idx[k0 + (i - Ru)]].dist;
                    memcpy(hs + i * img_bytes, src, img_bytes);
                };
                c.copy_pool.run(s.Ru + s.B, job);
                CE_CUDA(cudaMemcpyAsync(d_ref, hs, slot_bytes, cudaMemcpyHostToDevice, c.copy_stream));
            } else {
                // one copy per image, except that images which follow one another in host memory (a decoder writing
                // into one arena, the distortions of a sweep) go as ONE copy: with eight ranks sharing the host's memory
                // system, 3 MB copies reached 14 GB/s per GPU where 25 MB ones reach 24
                for (size_t r = 0; r < s.Ru;) {
                    size_t e = r + 1;
                    while (e < s.Ru && urefs[e] == urefs[e - 1] + img_bytes) e++;
                    CE_CUDA(cudaMemcpyAsync(d_ref + r * img_bytes, urefs[r], (e - r) * img_bytes, cudaMemcpyHostToDevice, c.copy_stream));
                    r = e;
                }
                for (size_t k = 0; k < s.B;) {
                    size_t e = k + 1;
                    while (e < s.B && pairs[idx[s.k0 + e]].dist == pairs[idx[s.k0 + e - 1]].dist + img_bytes) e++;
                    CE_CUDA(cudaMemcpyAsync(d_dist + k * img_bytes, pairs[idx[s.k0 + k]].dist, (e - k) * img_bytes,
                                            cudaMemcpyHostToDevice, c.copy_stream));
                    k = e;
                }
            }
            CE_CUDA(cudaEventRecord(c.ev_copy[ci & 1], c.copy_stream));
        };
        std::vector<int> local_of;
        {
            size_t largest = 0;
            for (size_t ci = 0; ci < nchunks; ci++) largest = std::max(largest, bounds[ci + 1] - bounds[ci]);
            c.ensure_results(largest * kRawDoubles * 8);
        }
        auto launch = [&](size_t ci) {
            Staged& s = st[ci & 1];
            CE_CUDA(cudaStreamWaitEvent(c.stream, c.ev_copy[ci & 1], 0));
            local_of.assign(s.Ru, -1);
            return launch_sub_batch(c, c.d_stage[ci & 1], c.d_stage[ci & 1] + s.Ru * img_bytes, s.ref_of.data(), local_of, 0, s.B, w, h,
                                    cfg, intensity, dbg, (int)(ci & 1));
        };
        try {
            // Per chunk: [chunk ci is computing] stage ci+1 -> wait for ci -> launch ci+1 -> evaluate ci's scores.
            // Staging after the launch matters for pageable memory: gathering it (or the driver's own blocking copy)
            // holds this thread, and that time runs under the kernels of chunk ci.  Staging slot (ci+1)&1 was last
            // read by chunk ci-1, which has completed.  The host-side score evaluation of ci runs under chunk ci+1.
            stage(0);
            SubBatch cur = launch(0);
            for (size_t ci = 0; ci < nchunks; ci++) {
                const Staged& s = st[ci & 1];
                const size_t k0 = s.k0, B = s.B;
                if (ci + 1 < nchunks) stage(ci + 1);
                wait_sub_batch(c);
                const SubBatch done = cur;
                if (ci + 1 < nchunks) cur = launch(ci + 1);
                tmp.resize(B);
                finalize_sub_batch(c, done, w, h, cfg, tmp.data(), dbg);
                for (size_t k = 0; k < B; k++) out[idx[k0 + k]] = tmp[k];
            }
        } catch (...) {
            cudaStreamSynchronize(c.copy_stream);   // no copy may outlive the caller's buffers
            cudaStreamSynchronize(c.stream);
            throw;
        }
    }
}

CE_API int ce_evaluate_batch(ce_ctx* ctx, const ce_pair* pairs, size_t n, const ce_metric_config* cfg, float intensity_target,
                             ce_result* out) {
    if (!ctx || !cfg || (n && (!pairs || !out))) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        evaluate_host_pairs(c, pairs, n, *cfg, intensity_target, out, nullptr);
    })
    return CE_OK;
}

static int single_pair(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t w, size_t h,
                       const ce_metric_config& cfg, float intensity, ce_result* r, const DebugOut* dbg = nullptr) {
    if (!ctx || !ref || !test) return CE_ERR_INVALID_ARGUMENT;
    if (w > 0xffffffffull || h > 0xffffffffull) return CE_ERR_INVALID_ARGUMENT;
    ce_pair p;
    memset(&p, 0, sizeof(p));
    p.ref = ref; p.dist = test; p.ref_len = ref_len; p.dist_len = test_len;
    p.width = (uint32_t)w; p.height = (uint32_t)h;
    Context& c = ctx->c;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        evaluate_host_pairs(c, &p, 1, cfg, intensity, r, dbg);
    })
    return r->status;
}

CE_API int ce_psnr(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t width,
                   size_t height, double* psnr, uint64_t* sse) {
    // the reference asserts (panics) on both length checks, src/metrics/mod.rs:313-314
    if (!ctx || !ref || !test) return CE_ERR_INVALID_ARGUMENT;
    if (ref_len != test_len || ref_len != width * height * 3 || ref_len == 0) {
        ctx->c.last_error = "calculate_psnr: buffer length does not match width*height*3";
        return CE_ERR_INVALID_ARGUMENT;
    }
    ce_metric_config cfg = {0, 0, 0, 1, 0};
    ce_result r;
    int st = single_pair(ctx, ref, ref_len, test, test_len, width, height, cfg, 80.0f, &r);
    if (st != CE_OK) return st;
    if (psnr) *psnr = r.psnr;
    if (sse) *sse = r.sse;
    return CE_OK;
}

CE_API int ce_ssimulacra2(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t width,
                          size_t height, double* score) {
    ce_metric_config cfg = {0, 1, 0, 0, 0};
    ce_result r;
    int st = single_pair(ctx, ref, ref_len, test, test_len, width, height, cfg, 80.0f, &r);
    if (st != CE_OK) return st;
    if (score) *score = r.ssimulacra2;
    return CE_OK;
}

CE_API int ce_butteraugli(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t width,
                          size_t height, float intensity_target, double* score, double* pnorm3) {
    ce_metric_config cfg = {0, 0, 1, 0, 0};
    ce_result r;
    int st = single_pair(ctx, ref, ref_len, test, test_len, width, height, cfg, intensity_target, &r);
    if (st != CE_OK) return st;
    if (score) *score = r.butteraugli;
    if (pnorm3) *pnorm3 = r.butteraugli_pnorm3;
    return CE_OK;
}

CE_API int ce_dssim_rgb8(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, const uint8_t* test, size_t test_len, size_t width,
                         size_t height, double* dssim) {
    ce_metric_config cfg = {1, 0, 0, 0, 0};
    ce_result r;
    int st = single_pair(ctx, ref, ref_len, test, test_len, width, height, cfg, 80.0f, &r);
    if (st != CE_OK) return st;
    if (dssim) *dssim = r.dssim;
    return CE_OK;
}

CE_API int ce_dssim_rgbaf32(ce_ctx* ctx, const float* ref, size_t ref_w, size_t ref_h, size_t ref_stride, const float* test,
                            size_t test_w, size_t test_h, size_t test_stride, double* dssim) {
    if (!ctx || !ref || !test) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    if (ref_w != test_w || ref_h != test_h) {  // src/metrics/dssim.rs:45-50
        c.last_error = "DSSIM: dimension mismatch";
        return CE_ERR_DIMENSION_MISMATCH;
    }
    if (ref_w == 0 || ref_h == 0 || ref_stride < ref_w || test_stride < test_w) {
        c.last_error = "DSSIM: Failed to create reference image";
        return CE_ERR_METRIC_CALCULATION;
    }
    const size_t w = ref_w, h = ref_h, n = w * h;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.arena.reset();
        const size_t need = (ref_stride + test_stride) * h * 16 + 8 * n * 4 + dssim_workspace_per_pair(w, h) + (1 << 20);
        if (c.arena.cap < need) throw OomError("workspace too small for one pair of this size");
        float* d_r = c.arena.alloc<float>(ref_stride * h * 4);
        float* d_t = c.arena.alloc<float>(test_stride * h * 4);
        CE_CUDA(cudaMemcpyAsync(d_r, ref, ref_stride * (h - 1) * 16 + w * 16, cudaMemcpyHostToDevice, c.stream));
        CE_CUDA(cudaMemcpyAsync(d_t, test, test_stride * (h - 1) * 16 + w * 16, cudaMemcpyHostToDevice, c.stream));
        float* lin = c.arena.alloc<float>(2 * 3 * n);   // [2 images][3][n]
        float* alpha = c.arena.alloc<float>(2 * n);     // [2 images][n]
        launch_rgba_to_planar(c, d_r, w, h, ref_stride, lin, alpha);
        launch_rgba_to_planar(c, d_t, w, h, test_stride, lin + 3 * n, alpha + n);
        c.ensure_results(10 * 8);
        c.ensure_idx(1);
        c.h_idx[0] = 0;
        CE_CUDA(cudaMemcpyAsync(c.d_idx, c.h_idx, sizeof(int), cudaMemcpyHostToDevice, c.stream));
        double* d_ds = reinterpret_cast<double*>(c.d_results);
        int ns = dssim_run(c, lin, alpha, 1, c.d_idx, 1, w, h, d_ds, nullptr);
        CE_CUDA(cudaMemcpyAsync(c.h_pinned, d_ds, 10 * 8, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
        if (dssim) *dssim = finalize_dssim(reinterpret_cast<const double*>(c.h_pinned), ns, w, h, nullptr);
    })
    return CE_OK;
}

static int to_dssim_image(ce_ctx* ctx, const uint8_t* data, size_t len, size_t width, size_t height, int ch, float* out) {
    if (!ctx || !data || !out) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    if (len != n * ch) {
        c.last_error = "buffer length does not match width*height*channels";
        return CE_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return CE_OK;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.arena.reset();
        uint8_t* d_in = c.arena.alloc<uint8_t>(len);
        float* d_out = c.arena.alloc<float>(n * 4);
        CE_CUDA(cudaMemcpyAsync(d_in, data, len, cudaMemcpyHostToDevice, c.stream));
        launch_rgb8_to_rgba_linear(c, d_in, n, ch, d_out);
        CE_CUDA(cudaMemcpyAsync(out, d_out, n * 16, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}
CE_API int ce_rgb8_to_dssim_image(ce_ctx* ctx, const uint8_t* data, size_t len, size_t width, size_t height, float* out) {
    return to_dssim_image(ctx, data, len, width, height, 3, out);
}
CE_API int ce_rgba8_to_dssim_image(ce_ctx* ctx, const uint8_t* data, size_t len, size_t width, size_t height, float* out) {
    return to_dssim_image(ctx, data, len, width, height, 4, out);
}

// ------------------------------------------------------------------ on-device distortion source
static int jpeg_args_ok(Context& c, int subsampling, const int* qualities, size_t n_q) {
    if (subsampling != 0 && subsampling != 2) {
        c.last_error = "subsampling must be 0 (4:4:4) or 2 (4:2:0)";
        return 0;
    }
    for (size_t k = 0; k < n_q; k++)
        if (qualities[k] < 1 || qualities[k] > 100) {
            c.last_error = "quality must be in 1..100";
            return 0;
        }
    if (n_q > 32) {
        c.last_error = "at most 32 quality levels per call";
        return 0;
    }
    return 1;
}

CE_API int ce_jpeg_roundtrip_device(ce_ctx* ctx, const uint8_t* d_refs, size_t n_ref, uint32_t width, uint32_t height,
                                    const int* qualities, size_t n_q, int subsampling, uint8_t* d_out) {
    if (!ctx || ((n_ref && n_q) && (!d_refs || !d_out || !qualities))) return CE_ERR_INVALID_ARGUMENT;
    if (n_ref == 0 || n_q == 0) return CE_OK;
    Context& c = ctx->c;
    if (width == 0 || height == 0) {
        c.last_error = "zero-sized image";
        return CE_ERR_INVALID_ARGUMENT;
    }
    if (!jpeg_args_ok(c, subsampling, qualities, n_q)) return CE_ERR_INVALID_ARGUMENT;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        const size_t img_bytes = (size_t)width * height * 3;
        const size_t per_ref = jpeg_workspace_bytes(1, n_q, width, height, subsampling);
        if (c.arena.cap < per_ref + (1 << 20)) throw OomError("workspace too small for one reference of this size");
        const size_t step = std::max<size_t>(1, std::min<size_t>((c.arena.cap - (1 << 20)) / per_ref, 65535 / n_q));
        for (size_t r0 = 0; r0 < n_ref; r0 += step) {
            const size_t nr = std::min(step, n_ref - r0);
            c.arena.reset();
            jpeg_roundtrip_run(c, d_refs + r0 * img_bytes, nr, width, height, qualities, n_q, subsampling,
                               d_out + r0 * n_q * img_bytes);
        }
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}

CE_API int ce_jpeg_roundtrip(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height, int quality,
                             int subsampling, uint8_t* out) {
    if (!ctx || !rgb || !out) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    if (width == 0 || height == 0 || len != width * height * 3) {
        c.last_error = "Buffer size mismatch";
        return CE_ERR_INVALID_ARGUMENT;
    }
    if (!jpeg_args_ok(c, subsampling, &quality, 1)) return CE_ERR_INVALID_ARGUMENT;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.ensure_input(2 * len);
        CE_CUDA(cudaMemcpyAsync(c.d_in, rgb, len, cudaMemcpyHostToDevice, c.stream));
        c.arena.reset();
        jpeg_roundtrip_run(c, c.d_in, 1, width, height, &quality, 1, subsampling, c.d_in + len);
        CE_CUDA(cudaMemcpyAsync(out, c.d_in + len, len, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}

CE_API int ce_evaluate_jpeg_sweep(ce_ctx* ctx, const uint8_t* const* refs, size_t n_ref, uint32_t width, uint32_t height,
                                  const int* qualities, size_t n_q, int subsampling, const ce_metric_config* cfg,
                                  float intensity_target, ce_result* out) {
    if (!ctx || !cfg || ((n_ref && n_q) && (!refs || !qualities || !out))) return CE_ERR_INVALID_ARGUMENT;
    if (n_ref == 0 || n_q == 0) return CE_OK;
    Context& c = ctx->c;
    if (width == 0 || height == 0) {
        c.last_error = "zero-sized image";
        return CE_ERR_INVALID_ARGUMENT;
    }
    for (size_t r = 0; r < n_ref; r++)
        if (!refs[r]) return CE_ERR_INVALID_ARGUMENT;
    if (!jpeg_args_ok(c, subsampling, qualities, n_q)) return CE_ERR_INVALID_ARGUMENT;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        const size_t w = width;
        const size_t h = height;
        const size_t img_bytes = w * h * 3;
        const size_t per_ref = jpeg_workspace_bytes(1, n_q, w, h, subsampling);
        if (c.arena.cap < per_ref + (1 << 20)) throw OomError("workspace too small for one reference of this size");
        // references per chunk: only the references are uploaded, so chunks exist to overlap that copy when it is
        // large (one chunk per 64 MB of references, at most 4), to keep the generated images of a chunk under 2 GiB and
        // the JPEG temporaries inside the arena
        size_t nchunks = std::min<size_t>(4, std::max<size_t>(1, n_ref * img_bytes / ((size_t)64 << 20)));
        size_t chunk = (n_ref + nchunks - 1) / nchunks;
        chunk = std::min<size_t>(chunk, std::max<size_t>(1, ((size_t)2 << 30) / (n_q * img_bytes)));
        chunk = std::min<size_t>(chunk, (c.arena.cap - (1 << 20)) / per_ref);
        chunk = std::max<size_t>(1, std::min<size_t>(chunk, 65535 / n_q));
        nchunks = (n_ref + chunk - 1) / chunk;
        auto stage = [&](size_t ci) {
            const size_t r0 = ci * chunk;
            const size_t nr = std::min(chunk, n_ref - r0);
            c.ensure_stage((int)(ci & 1), nr * (1 + n_q) * img_bytes);
            for (size_t r = 0; r < nr; r++)
                CE_CUDA(cudaMemcpyAsync(c.d_stage[ci & 1] + r * img_bytes, refs[r0 + r], img_bytes, cudaMemcpyHostToDevice,
                                        c.copy_stream));
            CE_CUDA(cudaEventRecord(c.ev_copy[ci & 1], c.copy_stream));
        };
        std::vector<uint32_t> ref_of;
        try {
            stage(0);
            for (size_t ci = 0; ci < nchunks; ci++) {
                if (ci + 1 < nchunks) stage(ci + 1);
                const size_t r0 = ci * chunk;
                const size_t nr = std::min(chunk, n_ref - r0);
                uint8_t* d_ref = c.d_stage[ci & 1];
                uint8_t* d_dist = d_ref + nr * img_bytes;
                CE_CUDA(cudaStreamWaitEvent(c.stream, c.ev_copy[ci & 1], 0));
                c.arena.reset();
                jpeg_roundtrip_run(c, d_ref, nr, w, h, qualities, n_q, subsampling, d_dist);
                ref_of.resize(nr * n_q);
                for (size_t i = 0; i < nr * n_q; i++) ref_of[i] = (uint32_t)(i / n_q);
                run_device_batch(c, d_ref, nr, d_dist, nr * n_q, ref_of.data(), w, h, *cfg, intensity_target, out + r0 * n_q, nullptr);
            }
        } catch (...) {
            cudaStreamSynchronize(c.copy_stream);
            throw;
        }
    })
    return CE_OK;
}

// transform_to_srgb, src/metrics/icc.rs:69-103
CE_API int ce_transform_to_srgb(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height, const uint8_t* icc,
                                size_t icc_len, uint8_t* out) {
    if (!ctx || !rgb || !out) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    if (len != width * height * 3) {
        c.last_error = "Buffer size mismatch";
        return CE_ERR_INVALID_ARGUMENT;
    }
    if (len == 0) return CE_OK;
    if (!icc || icc_len == 0) {   // ColorProfile::Srgb => rgb.to_vec()
        memmove(out, rgb, len);
        return CE_OK;
    }
    std::string why;
    if (!icc_is_usable(icc, icc_len, &why)) {
        c.last_error = why;
        return CE_ERR_METRIC_CALCULATION;
    }
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.ensure_input(2 * len);
        CE_CUDA(cudaMemcpyAsync(c.d_in, rgb, len, cudaMemcpyHostToDevice, c.stream));
        c.arena.reset();
        icc_to_srgb_run(c, c.d_in, width * height, icc, icc_len, c.d_in + len);
        CE_CUDA(cudaMemcpyAsync(out, c.d_in + len, len, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}

CE_API int ce_xyb_roundtrip(ce_ctx* ctx, const uint8_t* rgb, size_t len, size_t width, size_t height, uint8_t* out) {
    if (!ctx || !rgb || !out) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    if (len != n * 3) {  // assert_eq! in src/metrics/xyb.rs:227
        c.last_error = "Buffer size mismatch";
        return CE_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return CE_OK;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.arena.reset();
        uint8_t* d_in = c.arena.alloc<uint8_t>(len);
        uint8_t* d_out = c.arena.alloc<uint8_t>(len);
        CE_CUDA(cudaMemcpyAsync(d_in, rgb, len, cudaMemcpyHostToDevice, c.stream));
        launch_xyb_roundtrip(c, d_in, n, d_out);
        CE_CUDA(cudaMemcpyAsync(out, d_out, len, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}

// ---- reference reuse ------------------------------------------------------
CE_API int ce_reference_create(ce_ctx* ctx, const uint8_t* ref, size_t ref_len, size_t width, size_t height,
                               const ce_metric_config* cfg, ce_ref** out) {
    if (!ctx || !ref || !cfg || !out) return CE_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    Context& c = ctx->c;
    if (ref_len != width * height * 3 || ref_len == 0) {
        c.last_error = "Invalid image size: expected " + std::to_string(width * height * 3) + " bytes, got " + std::to_string(ref_len);
        return CE_ERR_METRIC_CALCULATION;
    }
    ce_ref* r = nullptr;
    try {
        CE_CUDA(cudaSetDevice(c.device));
        r = new ce_ref();
        r->owner = ctx; r->device = c.device; r->width = width; r->height = height; r->cfg = *cfg; r->d_ref = nullptr;
        CE_CUDA(cudaMalloc(&r->d_ref, ref_len));
        if (cfg->xyb_roundtrip) {
            c.arena.reset();
            uint8_t* d_in = c.arena.alloc<uint8_t>(ref_len);
            CE_CUDA(cudaMemcpyAsync(d_in, ref, ref_len, cudaMemcpyHostToDevice, c.stream));
            launch_xyb_roundtrip(c, d_in, width * height, r->d_ref);
        } else {
            CE_CUDA(cudaMemcpyAsync(r->d_ref, ref, ref_len, cudaMemcpyHostToDevice, c.stream));
        }
        CE_CUDA(cudaStreamSynchronize(c.stream));
        r->cfg.xyb_roundtrip = 0;  // already applied
        *out = r;
    } catch (const std::exception& e) {   // nothing of a half-built handle may leak
        c.last_error = e.what();
        if (r) {
            cudaStreamSynchronize(c.stream);
            if (r->d_ref) cudaFree(r->d_ref);
            delete r;
        }
        return dynamic_cast<const OomError*>(&e) ? CE_ERR_OUT_OF_MEMORY : CE_ERR_CUDA;
    }
    return CE_OK;
}

CE_API int ce_reference_compare_many(ce_ctx* ctx, ce_ref* ref, const uint8_t* const* dists, const size_t* dist_lens, size_t n_dist,
                                     float intensity_target, ce_result* out) {
    if (!ctx || !ref || ref->owner != ctx || (n_dist && (!dists || !dist_lens || !out))) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t img_bytes = ref->width * ref->height * 3;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        std::vector<size_t> ok;
        for (size_t i = 0; i < n_dist; i++) {
            memset(&out[i], 0, sizeof(ce_result));
            if (!dists[i]) { out[i].status = CE_ERR_INVALID_ARGUMENT; continue; }
            if (dist_lens[i] != img_bytes) {  // src/metrics/ssimulacra2.rs:65-70
                out[i].status = CE_ERR_DIMENSION_MISMATCH;
                c.last_error = "reference and test buffers differ in length";
                continue;
            }
            ok.push_back(i);
        }
        const size_t B = ok.size();
        if (B) {
            c.ensure_input(B * img_bytes);
            uint8_t* d_dist = c.d_in;
            for (size_t k = 0; k < B; k++)
                CE_CUDA(cudaMemcpyAsync(d_dist + k * img_bytes, dists[ok[k]], img_bytes, cudaMemcpyHostToDevice, c.stream));
            std::vector<ce_result> tmp(B);
            std::vector<uint32_t> ref_of(B, 0u);   // every distortion compares against the one resident reference
            run_device_batch(c, ref->d_ref, 1, d_dist, B, ref_of.data(), ref->width, ref->height, ref->cfg, intensity_target,
                             tmp.data(), nullptr);
            for (size_t k = 0; k < B; k++) out[ok[k]] = tmp[k];
        }
    })
    return CE_OK;
}

CE_API int ce_reference_compare(ce_ctx* ctx, ce_ref* ref, const uint8_t* dist, size_t dist_len, float intensity_target,
                                ce_result* out) {
    if (!out) return CE_ERR_INVALID_ARGUMENT;
    int st = ce_reference_compare_many(ctx, ref, &dist, &dist_len, 1, intensity_target, out);
    return st != CE_OK ? st : out->status;
}

CE_API void ce_reference_destroy(ce_ref* ref) {
    if (!ref) return;
    if (ref->d_ref) {   // the owning context may already be gone: only the handle's own fields are used
        cudaSetDevice(ref->device);
        cudaFree(ref->d_ref);
    }
    delete ref;
}

// ---- stage-level entry points ----------------------------------------------
CE_API int ce_debug_ssim2_sums(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height, double* sums,
                               int* nscales) {
    ce_metric_config cfg = {0, 1, 0, 0, 0};
    ce_result r;
    DebugOut d;
    d.s2_sums = sums;
    d.s2_nscales = nscales;
    return single_pair(ctx, ref, width * height * 3, dist, width * height * 3, width, height, cfg, 80.0f, &r, &d);
}

CE_API int ce_debug_ssim2_scale0_planes(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height,
                                        float* planes) {
    if (!ctx || !planes) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    float* d_planes = nullptr;
    int st;
    try {
        CE_CUDA(cudaSetDevice(c.device));
        CE_CUDA(cudaMalloc(&d_planes, 21 * n * 4));
    } catch (const std::exception& e) { c.last_error = e.what(); return CE_ERR_CUDA; }
    ce_metric_config cfg = {0, 1, 0, 0, 0};
    ce_result r;
    DebugOut d;
    d.s2_planes = d_planes;
    st = single_pair(ctx, ref, n * 3, dist, n * 3, width, height, cfg, 80.0f, &r, &d);
    if (st == CE_OK) cudaMemcpy(planes, d_planes, 21 * n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_planes);
    return st;
}

CE_API int ce_debug_dssim_scales(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height,
                                 double* scale_scores, int* nscales, float* map0) {
    if (!ctx) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    float* d_map = nullptr;
    if (map0) {
        try {
            CE_CUDA(cudaSetDevice(c.device));
            CE_CUDA(cudaMalloc(&d_map, n * 4));
        } catch (const std::exception& e) { c.last_error = e.what(); return CE_ERR_CUDA; }
    }
    ce_metric_config cfg = {1, 0, 0, 0, 0};
    ce_result r;
    DebugOut d;
    d.ds_scores = scale_scores;
    d.ds_nscales = nscales;
    d.ds_map0 = d_map;
    int st = single_pair(ctx, ref, n * 3, dist, n * 3, width, height, cfg, 80.0f, &r, &d);
    if (st == CE_OK && map0) cudaMemcpy(map0, d_map, n * 4, cudaMemcpyDeviceToHost);
    if (d_map) cudaFree(d_map);
    return st;
}

CE_API int ce_debug_butteraugli_diffmap(ce_ctx* ctx, const uint8_t* ref, const uint8_t* dist, size_t width, size_t height,
                                        float intensity_target, float* diffmap) {
    if (!ctx || !diffmap) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    float* d_map = nullptr;
    try {
        CE_CUDA(cudaSetDevice(c.device));
        CE_CUDA(cudaMalloc(&d_map, n * 4));
    } catch (const std::exception& e) { c.last_error = e.what(); return CE_ERR_CUDA; }
    ce_metric_config cfg = {0, 0, 1, 0, 0};
    ce_result r;
    DebugOut d;
    d.ba_diffmap = d_map;
    int st = single_pair(ctx, ref, n * 3, dist, n * 3, width, height, cfg, intensity_target, &r, &d);
    if (st == CE_OK) cudaMemcpy(diffmap, d_map, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_map);
    return st;
}

static int debug_ba_image(ce_ctx* ctx, const uint8_t* rgb, size_t width, size_t height, float intensity, float* planes, int nplanes,
                          bool opsin) {
    if (!ctx || !rgb || !planes) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    if (width < 8 || height < 8) return CE_ERR_METRIC_CALCULATION;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.arena.reset();
        uint8_t* d_in = c.arena.alloc<uint8_t>(n * 3);
        float* lin = c.arena.alloc<float>(3 * n);
        float* d_out = c.arena.alloc<float>((size_t)nplanes * n);
        CE_CUDA(cudaMemcpyAsync(d_in, rgb, n * 3, cudaMemcpyHostToDevice, c.stream));
        launch_srgb8_to_linear(c, d_in, nullptr, 1, n, lin);
        if (opsin) butteraugli_debug_opsin(c, lin, width, height, intensity, d_out);
        else butteraugli_debug_psycho(c, lin, width, height, intensity, d_out);
        CE_CUDA(cudaMemcpyAsync(planes, d_out, (size_t)nplanes * n * 4, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}
CE_API int ce_debug_butteraugli_psycho(ce_ctx* ctx, const uint8_t* rgb, size_t width, size_t height, float intensity_target,
                                       float* planes) {
    return debug_ba_image(ctx, rgb, width, height, intensity_target, planes, 10, false);
}
CE_API int ce_debug_butteraugli_opsin(ce_ctx* ctx, const uint8_t* rgb, size_t width, size_t height, float intensity_target,
                                      float* planes) {
    return debug_ba_image(ctx, rgb, width, height, intensity_target, planes, 3, true);
}
CE_API int ce_debug_ba_blur(ce_ctx* ctx, const float* plane, size_t width, size_t height, float sigma, float* out) {
    if (!ctx || !plane || !out) return CE_ERR_INVALID_ARGUMENT;
    Context& c = ctx->c;
    const size_t n = width * height;
    CE_TRY(c, {
        CE_CUDA(cudaSetDevice(c.device));
        c.arena.reset();
        float* d_in = c.arena.alloc<float>(n);
        float* d_out = c.arena.alloc<float>(n);
        CE_CUDA(cudaMemcpyAsync(d_in, plane, n * 4, cudaMemcpyHostToDevice, c.stream));
        butteraugli_debug_blur(c, d_in, width, height, sigma, d_out);
        CE_CUDA(cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, c.stream));
        CE_CUDA(cudaStreamSynchronize(c.stream));
    })
    return CE_OK;
}

}  // extern "C"
