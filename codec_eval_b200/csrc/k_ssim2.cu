// k_ssim2.cu -- SSIMULACRA2 (fast-ssim2 0.8.0 == libjxl tools/ssimulacra2.cc;
// reference call site src/metrics/ssimulacra2.rs:96) for a batch of B pairs.
//
// Per scale (6 scales, ceil-halving with clamp on LINEAR rgb):
//   k_s2_xyb_down : linear(s) -> positive-XYB(s) planes + linear(s+1)        [HBM-bound, pointwise]
//   k_s2_hpass    : recursive Gaussian along x of {i1,i2,i1^2,i2^2,i1*i2};   [rows staged in smem by TMA]
//                   lane = row, warp = product, 32-row x 64-column input slots double-buffered
//                   through shared memory (TMA boxes with the 128-byte swizzle so lane = row reads
//                   are conflict-free; cp.async + padded pitch when the row pitch is not 16 B aligned)
//   k_s2_vpass    : column-parallel recurrence along y (lane = column, warp = product, 10-row
//                   register delay line, 5-row slots staged by TMA) fused with the SSIM /
//                   edge-artifact / detail-loss maps
//                   and L1/L4 pooling in fp64 (warp shuffles -> per-block partials)
//   k_s2_reduce   : fixed-order sum of the block partials -> 18 sums per (pair, scale)
// The recurrence is the exact operation sequence of the upstream code so the
// result does not depend on the tiling.
#include <cuda.h>

#include "ce_common.cuh"
#include "ce_internal.h"

namespace ce {

__constant__ RGaussCoef c_rg;

// ------------------------------------------------------------------ xyb + down2
// thread = one 2x2 block of the current scale (= one pixel of the next scale); grid.z = image
__global__ void __launch_bounds__(256) k_s2_xyb_down(const float* __restrict__ lin_all, int w, int h, int ow, int oh,
                                                      size_t n, size_t on, float* __restrict__ xyb,
                                                      float* __restrict__ nlin, int write_down) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const size_t im = blockIdx.z;
    if (ox >= ow || oy >= oh) return;
    const float* lin = lin_all + im * 3 * n;
    float* xo = xyb + im * 3 * n;
    float* no = nlin + im * 3 * on;
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
#ifndef HP_COLS
#define HP_COLS 64
#endif
#define HP_C4 (HP_COLS / 4)
#define HP_PITCH (HP_COLS + 4)   // 16-B aligned rows; lane = row reads LDS.128 at chunk (C4+1)*row + q: conflict-free per quarter warp
#ifndef HP_SLOTS
#define HP_SLOTS 2
#endif
// Output stage: a warp collects HP_OC finished columns of its 32 rows (lane = row) and writes them out transposed
// (lane = 16-byte group).  The recurrence runs 4 columns behind the loads, so a chunk finishes columns 64k-4 .. 64k+59;
// the stage is used as a ring over the ALIGNED block of HP_OC columns (column g at g mod HP_OC) and a block leaves as
// soon as the group that completes it has been stored: every row segment of a store is then whole 128-byte lines
// (with the stores starting at 64k-4 each 256-byte row segment touched three lines, and narrower unaligned stages --
// 32 / 16 columns, 4 / 5 blocks per SM instead of 3 -- were measured slower on the corpus step, 236 / 304 ms against
// 194: the kernel waits for its stores, not for occupancy).
#ifndef HP_OC
#define HP_OC 32
#endif
#define HP_OC4 (HP_OC / 4)
#define HP_OPITCH (HP_OC + 4)    // lane = row writes STS.128 at row pitch OC+4 floats: conflict free per quarter warp

// Which blurs a launch computes.  The reference-side statistics (mu1 = blur(i1), blur(i1^2)) do not depend on the
// distorted image, so when a sub-batch has shared references they are computed once per distinct reference
// (S2_REF) and the per-pair launch (S2_PAIR) does the other three -- fast_ssim2's Ssimulacra2Reference split.
enum { S2_ALL = 0, S2_REF = 1, S2_PAIR = 2 };
template <int MODE> struct S2Mode {
    static constexpr int NW = MODE == S2_ALL ? 5 : MODE == S2_REF ? 2 : 3;     // products = warps per block
    static constexpr int NIN = MODE == S2_REF ? 1 : 2;                          // image planes staged by the row pass
    static constexpr int HP_SMEM = (HP_SLOTS * NIN * HP_PITCH + NW * HP_OPITCH) * HP_ROWS * 4;
    static constexpr int HP_SMEM_TMA = (2 * NIN * (HP_COLS / 32) * 32 * 32 + NW * HP_ROWS * HP_OPITCH) * 4;   // HP_SLOTS_TMA = 2
};

// grid (ceil(h/32), 3*units); block NW warps (product) x 32 lanes (row); unit = pair (S2_ALL, S2_PAIR) or distinct
// reference (S2_REF).  Warp p blurs one product of the band's 32 rows:
//   S2_ALL: i1, i2, i1*i1, i2*i2, i1*i2      S2_REF: i1, i1*i1      S2_PAIR: i2, i2*i2, i1*i2
// (one code path for all warps so the loop body stays inside the instruction cache).
// The row is walked in HP_COLS-column chunks staged by cp.async (three slots, two chunks in flight,
// zero-filled past the image, one block barrier per chunk).  Output n needs in[n+4] and in[n-6], so the
// recurrence runs 4 columns behind the loads: chunk k (input columns k*HP_COLS ..) produces output columns
// k*HP_COLS-4 .. ; the first four steps are the upstream warm-up (n = -4 .. -1), and
// ceil((w+4)/HP_COLS) chunks reach the last column.  Each warp writes out the plane it produced.
// hb: [unit][3][NW][n].
// TMA variant: a chunk of one plane is two 32-column x 32-row boxes written with the 128-byte swizzle (4 KB each,
// no padding): lane = row r reads its 16-byte group j at chunk j ^ (r & 7), conflict free.  Plane stride in a slot =
// HP_TPLANE floats.
#define HP_TPLANE ((HP_COLS / 32) * 32 * 32)
#define HP_SLOTS_TMA 2
template <int MODE, bool TMA>
__global__ void __launch_bounds__(S2Mode<MODE>::NW * 32) k_s2_hpass(const float* __restrict__ xyb, size_t R,
                                                                     const int* __restrict__ ridx, float* __restrict__ hb,
                                                                     int w, int h, size_t n, int vec,
                                                                     const __grid_constant__ CUtensorMap map) {
    constexpr int NW = S2Mode<MODE>::NW, NIN = S2Mode<MODE>::NIN, NT = NW * 32;
    constexpr int PLANE = TMA ? HP_TPLANE : HP_ROWS * HP_PITCH;     // floats per staged plane
    constexpr int SLOTS = TMA ? HP_SLOTS_TMA : HP_SLOTS;            // ring depth
    extern __shared__ __align__(1024) float s_dyn[];   // opt-in dynamic shared memory (> 48 KB)
    __shared__ __align__(8) unsigned long long s_bar[HP_SLOTS_TMA];
    float* s_in = s_dyn;
    const int lane = threadIdx.x & 31, p = threadIdx.x >> 5;
    float* so = s_dyn + SLOTS * NIN * PLANE + p * (HP_ROWS * HP_OPITCH);
    const size_t u = blockIdx.y / 3;
    const int c = blockIdx.y % 3;
    const int row0 = blockIdx.x * HP_ROWS;
    const float* i1 = xyb + ((MODE == S2_REF ? u : (size_t)ridx[u]) * 3 + c) * n;
    const float* i2 = xyb + ((R + u) * 3 + c) * n;   // unused by S2_REF
    float* op = hb + ((u * 3 + c) * NW + p) * n;
    const int nchunks = (w + 4 + HP_COLS - 1) / HP_COLS;
    // product roles: plane of the first / second factor, and whether there is a second factor
    int offA, offB;
    bool mul;
    if (MODE == S2_ALL) { offA = (p == 1 || p == 3) ? PLANE : 0; offB = (p >= 3) ? PLANE : 0; mul = p >= 2; }
    else if (MODE == S2_REF) { offA = 0; offB = 0; mul = p == 1; }
    else { offA = (p == 2) ? 0 : PLANE; offB = PLANE; mul = p >= 1; }
    const size_t pl1 = (MODE == S2_REF ? u : (size_t)ridx[u]) * 3 + c, pl2 = (R + u) * 3 + c;   // planes of the xyb tensor

    auto issue = [&](int k) {
        if (TMA) {
            if (threadIdx.x == 0 && k < nchunks) {
                float* slot = s_in + (k % SLOTS) * NIN * PLANE;
                unsigned long long* bar = &s_bar[k % SLOTS];
                mbar_expect_tx(bar, NIN * PLANE * 4);
#pragma unroll
                for (int half = 0; half < HP_COLS / 32; half++) {
                    tma_load_3d(slot + half * 1024, &map, k * HP_COLS + 32 * half, row0, (int)pl1, bar);
                    if (NIN == 2) tma_load_3d(slot + PLANE + half * 1024, &map, k * HP_COLS + 32 * half, row0, (int)pl2, bar);
                }
            }
            return;
        }
        if (k < nchunks) {
            constexpr int ITEMS = NIN * HP_ROWS * HP_C4;
            float* slot = s_in + (k % SLOTS) * NIN * PLANE;
#pragma unroll
            for (int it = 0; it < (ITEMS + NT - 1) / NT; it++) {
                const int e = (int)threadIdx.x + it * NT;
                if (ITEMS % NT != 0 && e >= ITEMS) break;
                const int pl = e / (HP_ROWS * HP_C4), rem = e - pl * (HP_ROWS * HP_C4);
                const int rr = rem / HP_C4, c4 = rem - rr * HP_C4;
                const int y = row0 + rr, x = k * HP_COLS + 4 * c4;
                const float* base = (NIN == 2 && pl) ? i2 : i1;
                float* dst = slot + (pl * HP_ROWS + rr) * HP_PITCH + 4 * c4;
                if (vec) {
                    const bool ok = y < h && x < w;
                    cp_async16(dst, ok ? base + (size_t)y * w + x : base, ok);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const bool ok = y < h && x + j < w;
                        cp_async4(dst + j, ok ? base + (size_t)y * w + x + j : base, ok);
                    }
                }
            }
        }
        cp_async_commit();
    };
    if (TMA) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < SLOTS; i++) mbar_init(&s_bar[i], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    issue(0);
    if (SLOTS >= 3) issue(1);
    RGState st = {0, 0, 0, 0, 0, 0};
    // products of the previous 16 columns: Q0 = c-16..c-13, Q1 = c-12..c-9, Q2 = c-8..c-5, Q3 = c-4..c-1
    float Q0[4] = {0, 0, 0, 0}, Q1[4] = {0, 0, 0, 0}, Q2[4] = {0, 0, 0, 0}, Q3[4] = {0, 0, 0, 0};
    // this warp writes out aligned block kb (HP_OC columns) of the plane it produced:
    // lane -> (row lane/OC4 + (32/OC4)*i, 16-B group lane%OC4)
    auto flush = [&](int kb) {
        __syncwarp();
        const int x = kb * HP_OC + 4 * (lane % HP_OC4);
        if (x < w) {
#pragma unroll
            for (int i = 0; i < HP_OC4; i++) {
                const int rr = (32 / HP_OC4) * i + lane / HP_OC4;
                const int y = row0 + rr;
                if (y < h) {
                    const float4 v = *reinterpret_cast<const float4*>(so + rr * HP_OPITCH + 4 * (lane % HP_OC4));
                    float* d = op + (size_t)y * w + x;
                    if (vec) *reinterpret_cast<float4*>(d) = v;
                    else {
                        d[0] = v.x;
                        if (x + 1 < w) d[1] = v.y;
                        if (x + 2 < w) d[2] = v.z;
                        if (x + 3 < w) d[3] = v.w;
                    }
                }
            }
        }
        __syncwarp();
    };
    for (int k = 0; k < nchunks; k++) {
        if (TMA) {
            mbar_wait(&s_bar[k % SLOTS], (unsigned)(k / SLOTS) & 1u);
            fence_proxy_async();   // this thread's reads of slot (k-1) % HP_SLOTS precede its refill below
        } else if (SLOTS >= 3) {
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();   // chunk k visible to all; every warp is past its reads of slot (k-1) % HP_SLOTS
        issue(k + SLOTS - 1);      // -> slot (k-1) % SLOTS
        // lane = row: padded rows (cp.async) or 128-byte rows with the TMA swizzle
        const float* a = TMA ? s_in + (k % SLOTS) * NIN * PLANE + lane * 32 : s_in + ((k % SLOTS) * NIN * HP_ROWS + lane) * HP_PITCH;
        const int sw = (lane & 7) << 2;   // swizzle: float offset of 16-byte group j is (4 j) ^ sw
#pragma unroll 1
        for (int m0 = 0; m0 < HP_C4; m0 += 4) {
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                const int m = m0 + mm;
                const int off = TMA ? ((m >> 3) * 1024 + (((m & 7) << 2) ^ sw)) : 4 * m;
                float4 g4 = *reinterpret_cast<const float4*>(a + offA + off);
                if (mul) {
                    const float4 v = *reinterpret_cast<const float4*>(a + offB + off);
                    g4.x *= v.x; g4.y *= v.y; g4.z *= v.z; g4.w *= v.w;
                }
                // steps n = c-4+j: right = in[n+4] = g4[j]; left = in[n-6] = column c-10+j
                float4 o;
                float* L0 = (mm == 0) ? Q1 : (mm == 1) ? Q2 : (mm == 2) ? Q3 : Q0;   // group holding c-12..c-9
                float* L1 = (mm == 0) ? Q2 : (mm == 1) ? Q3 : (mm == 2) ? Q0 : Q1;   // group holding c-8..c-5
                float* NWQ = (mm == 0) ? Q0 : (mm == 1) ? Q1 : (mm == 2) ? Q2 : Q3;  // oldest group, replaced by the new one
                o.x = rg_step(st, L0[2] + g4.x);
                o.y = rg_step(st, L0[3] + g4.y);
                o.z = rg_step(st, L1[0] + g4.z);
                o.w = rg_step(st, L1[1] + g4.w);
                // columns 64k-4+4m .. +3 -> ring position (4m-4) mod HP_OC
                *reinterpret_cast<float4*>(so + lane * HP_OPITCH + ((4 * m + HP_OC - 4) & (HP_OC - 1))) = o;
                NWQ[0] = g4.x; NWQ[1] = g4.y; NWQ[2] = g4.z; NWQ[3] = g4.w;
                // a group that ends at a multiple of HP_OC columns completed the block before it
                if (mm == 0 && (m0 % HP_OC4) == 0 && (k > 0 || m0 > 0)) flush((k * HP_COLS + 4 * m0) / HP_OC - 1);
            }
        }
    }
    flush(nchunks * (HP_COLS / HP_OC) - 1);   // what the last, incomplete block holds of the image (often nothing)
}

// Tried in round 2 and dropped (measurements on the 192-pair batch, block version above = 1.39 ms):
//   * 32-column chunks (HP_COLS 32: 30 KB per block, 7 blocks / SM instead of 3): 1.89 ms -- twice the block barriers
//     and half-length inner loops cost more than the occupancy gives;
//   * one warp per row band running all three products (nine independent chains per thread, no block barrier, 24 KB
//     per warp): 2.57 ms with a coalescing output stage, 4.16 ms storing each lane's float4 straight to global memory
//     (32 distinct lines per store instruction).  ncu: 8 warps / SM, 17 % issue-active, the stalls are the shared-memory
//     loads at the head of every 4-column group (short scoreboard 4.8 per issue) and the wait for the next chunk -- one
//     warp cannot hide them, three warps sharing a barrier can.

// ------------------------------------------------------------------ vertical pass + maps
#define VP_COLS 32
#define VP_BATCH 5
#define VP_SLOTS 4
#define VP_MAXITEMS 3   // staged 16-B items per thread and batch

// grid (ceil(w/32), 3*units); block NW warps x 32 lanes.  lane = column.  Each warp runs the recurrence of ONE
// product down the column strip (10-row register delay line).  Rows arrive in batches of 5 through a 4-slot
// cp.async ring (two batches in flight; every thread owns fixed copy slots whose source pointers just advance).
//   S2_ALL : ring planes = 5 row-pass planes + i1 + i2; the five blurred values of each pixel meet in shared memory
//            and, one batch later, warp r evaluates the SSIM / edge-artifact / detail-loss terms of row r.
//   S2_REF : ring planes = 2 row-pass planes; the finished mu1 / blur(i1^2) rows go to vref [r][3][2][n].
//   S2_PAIR: ring planes = 3 row-pass planes + i1 + i2 + the reference's mu1 and blur(i1^2) rows (vref); the
//            map rows of a batch are dealt over the three warps.
// ONE block barrier per 5 rows.  partials: [(unit*3+c)][gridDim.x][6]
struct VpassMaps {   // TMA descriptors of the column pass (TMA variant): planes are the third tensor dimension
    CUtensorMap hb;    // row-pass planes [units*3*NW][h][w], box (32, 5, NW)
    CUtensorMap xyb;   // [NI*3][h][w], box (32, 5, 1)
    CUtensorMap vr;    // reference statistics [R*3*2][h][w], box (32, 5, 2)
};
template <int MODE, bool TMA>
__global__ void __launch_bounds__(S2Mode<MODE>::NW * 32) k_s2_vpass(const float* __restrict__ xyb, size_t R,
                                                                     const int* __restrict__ ridx,
                                                                     const float* __restrict__ hb, float* __restrict__ vref,
                                                                     int w, int h, size_t n, double* __restrict__ partials,
                                                                     float* __restrict__ dbg, int vec,
                                                                     const __grid_constant__ VpassMaps maps) {
    constexpr int NW = S2Mode<MODE>::NW, NT = NW * 32;
    constexpr int NPL = MODE == S2_REF ? 2 : 7;          // ring planes
    constexpr int SLOT_FLOATS = VP_BATCH * NPL * VP_COLS;
    __shared__ __align__(128) float s_ld[VP_SLOTS * SLOT_FLOATS];   // [slot][plane][row r][col]
    __shared__ __align__(8) unsigned long long s_bar[VP_SLOTS];
    __shared__ float s_v[MODE == S2_REF ? 1 : 2][VP_BATCH][NW][VP_COLS];
    __shared__ double scratch[6 * 32];
    const int lane = threadIdx.x & 31, p = threadIdx.x >> 5;
    const int x0 = blockIdx.x * VP_COLS;
    const int x = x0 + lane;
    const size_t u = blockIdx.y / 3;
    const int c = blockIdx.y % 3;
    const bool active = x < w;
    const size_t iref = MODE == S2_REF ? u : (size_t)ridx[u];
    const float* i1 = xyb + (iref * 3 + c) * n;
    const float* i2 = xyb + ((R + u) * 3 + c) * n;
    const float* hp = hb + ((u * 3 + c) * NW) * n;
    const float* vr = MODE == S2_PAIR ? vref + ((iref * 3 + c) * 2) * n : nullptr;
    const int total = h + 4;  // input rows j = 0 .. h+3 (rows >= h are zero); output row y = j - 4
    const int nbatch = (total + VP_BATCH - 1) / VP_BATCH;

    // copy roles: item e -> (row r of the batch, ring plane pl, 16-B group c4); planes < NW are row-pass planes
    // (row j0 + r), the others are image-space planes (row j0 + r - 4).
    constexpr int ITEMS = VP_BATCH * NPL * (VP_COLS / 4);
    // Per copy slot: the source address for batch row 0 and the range of batch start rows jn for which the source
    // row lies inside the image (so the per-batch work is two compares, one 64-bit add and the copy).
    const float* src0[VP_MAXITEMS];
    int jlo[VP_MAXITEMS], jhi[VP_MAXITEMS], sdst[VP_MAXITEMS];
    static_assert((ITEMS + NT - 1) / NT <= VP_MAXITEMS, "copy roles");
#pragma unroll
    for (int it = 0; it < VP_MAXITEMS; it++) {
        const int e = (int)threadIdx.x + it * NT;
        src0[it] = hp; jlo[it] = 0x7fffffff; jhi[it] = 0; sdst[it] = 0;
        if (e < ITEMS) {
            const int r = e / (NPL * 8), rem = e - r * (NPL * 8), pl = rem >> 3, c4 = rem & 7;
            const float* base;
            if (pl < NW) base = hp + (size_t)pl * n;
            else if (MODE == S2_ALL) base = (pl == 5) ? i1 : i2;
            else base = (pl == 3) ? i1 : (pl == 4) ? i2 : vr + (size_t)(pl - 5) * n;
            const int srow = (pl < NW) ? r : r - 4;   // source row of this slot when the batch starts at row 0
            sdst[it] = (pl * VP_BATCH + r) * VP_COLS + 4 * c4;
            if (x0 + 4 * c4 < w) {
                src0[it] = base + (ptrdiff_t)srow * w + x0 + 4 * c4;
                jlo[it] = -srow; jhi[it] = h - srow;
            }
        }
    }
    auto issue = [&](int t) {   // called for t = 0, 1, 2, ... in order
        if (TMA) {
            // one thread: arm the slot's barrier with the byte count, then one bulk tensor copy per source tensor
            // (rows / columns outside the image arrive as zeros).  The ~60 instructions of the issue go to the LAST warp:
            // in pair mode it evaluates one map row per batch where the other two evaluate two, so the three warps
            // reach the batch barrier together (on warp 0 the issue made it the slowest by a third).
            if (threadIdx.x == NT - 32 && t < nbatch) {
                float* slot = s_ld + (t & (VP_SLOTS - 1)) * SLOT_FLOATS;
                unsigned long long* bar = &s_bar[t & (VP_SLOTS - 1)];
                const int jn = t * VP_BATCH;
                mbar_expect_tx(bar, SLOT_FLOATS * 4);
                tma_load_3d(slot, &maps.hb, x0, jn, (int)((u * 3 + c) * NW), bar);
                if (MODE != S2_REF) {
                    tma_load_3d(slot + NW * VP_BATCH * VP_COLS, &maps.xyb, x0, jn - 4, (int)(iref * 3 + c), bar);
                    tma_load_3d(slot + (NW + 1) * VP_BATCH * VP_COLS, &maps.xyb, x0, jn - 4, (int)((R + u) * 3 + c), bar);
                }
                if (MODE == S2_PAIR) tma_load_3d(slot + (NW + 2) * VP_BATCH * VP_COLS, &maps.vr, x0, jn - 4, (int)((iref * 3 + c) * 2), bar);
            }
            return;
        }
        if (t < nbatch) {
            float* slot = s_ld + (t & (VP_SLOTS - 1)) * SLOT_FLOATS;
            const int jn = t * VP_BATCH;   // first row of the batch
            if (vec) {
                const size_t ro = (size_t)jn * w;
#pragma unroll
                for (int it = 0; it < VP_MAXITEMS; it++) {
                    if ((int)threadIdx.x + it * NT < ITEMS) {
                        const bool ok = jn >= jlo[it] && jn < jhi[it];
                        cp_async16(slot + sdst[it], ok ? src0[it] + ro : hp, ok);
                    }
                }
            } else {
                for (int e = threadIdx.x; e < VP_BATCH * NPL * VP_COLS; e += NT) {
                    const int cx = e & 31, pl = (e >> 5) % NPL, r = e / (NPL * 32);
                    const int xx = x0 + cx;
                    const float* base;
                    if (pl < NW) base = hp + (size_t)pl * n;
                    else if (MODE == S2_ALL) base = (pl == 5) ? i1 : i2;
                    else base = (pl == 3) ? i1 : (pl == 4) ? i2 : vr + (size_t)(pl - 5) * n;
                    const int row = (pl < NW) ? jn + r : jn + r - 4;
                    const bool ok = row >= 0 && row < h && xx < w;
                    cp_async4(slot + (pl * VP_BATCH + r) * VP_COLS + cx, ok ? base + (size_t)row * w + xx : base, ok);
                }
            }
        }
        cp_async_commit();
    };

    // Pooling in fp64: a sum of fp32 terms in fp64 is exact at these counts and magnitudes, so the pooled sums -- and
    // with them the score -- do not depend on how rows are dealt over warps, i.e. on the launch mode or the batch a
    // pair travels in (test_full_size_1024: the same pair alone and inside a shared-reference batch gives the same
    // bits).  fp32 partials flushed every few rows were measured 4 % faster on this kernel and rejected for that.
    double acc[6] = {0, 0, 0, 0, 0, 0};
    // SSIM / edge terms of row r of batch t (blurred values in s_v[t & 1], image-space rows in slot t % VP_SLOTS)
    auto map_row = [&](int t, int r) {
        const int y = t * VP_BATCH + r - 4;
        if (y >= 0 && y < h && active) {
            const int bf = t & 1;
            const float* slot = s_ld + (t & (VP_SLOTS - 1)) * SLOT_FLOATS + r * VP_COLS + lane;   // + plane * VP_BATCH * VP_COLS
            float m1, m2, s11, s22, s12, a1, a2;
            if (MODE == S2_ALL) {
                m1 = s_v[bf][r][0][lane]; m2 = s_v[bf][r][1][lane]; s11 = s_v[bf][r][2][lane];
                s22 = s_v[bf][r][NW > 3 ? 3 : 0][lane]; s12 = s_v[bf][r][NW > 4 ? 4 : 0][lane];
                a1 = slot[5 * VP_BATCH * VP_COLS]; a2 = slot[6 * VP_BATCH * VP_COLS];
            } else {
                m2 = s_v[bf][r][0][lane]; s22 = s_v[bf][r][1][lane]; s12 = s_v[bf][r][NW > 2 ? 2 : 0][lane];
                a1 = slot[3 * VP_BATCH * VP_COLS]; a2 = slot[4 * VP_BATCH * VP_COLS];
                m1 = slot[(NPL > 5 ? 5 : 0) * VP_BATCH * VP_COLS]; s11 = slot[(NPL > 6 ? 6 : 0) * VP_BATCH * VP_COLS];
            }
            if (dbg) {
                float* d = dbg + (size_t)c * 7 * n + (size_t)y * w + x;
                d[0] = a1; d[n] = a2; d[2 * n] = m1; d[3 * n] = m2; d[4 * n] = s11; d[5 * n] = s22; d[6 * n] = s12;
            }
            // ssim_map.  Upstream forms these terms in f64 from the f32 blurs; here they are formed in fp32 with
            // cancellation-free expressions (1 - q is exact for q in [0.5, 2]; (1+x)/(1+y) - 1 = (x-y)/(1+y)) and only
            // the pooled sums are fp64: per-pixel relative error ~1e-7, sums agree with the f64 forms to ~1e-7.
            const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            const float mdiff = m1 - m2;
            const float num_m = __fmaf_rn(mdiff, -mdiff, 1.0f);
            const float num_s = __fmaf_rn(2.0f, s12 - m12, 0.0009f);
            const float denom_s = ((s11 - m11) + (s22 - m22)) + 0.0009f;
            // IEEE quotient like upstream (a 2-ulp __fdividef here was measured 4.6 % faster on this kernel, 0.1 % on the
            // step, and moved scores by ~2e-5: not taken)
            const float d = fmaxf(1.0f - div_rn_normal(num_m * num_s, denom_s), 0.0f);   // NaN -> 0 like !(d > 0)
            const float d2 = d * d;
            acc[0] += (double)d;
            acc[1] += (double)(d2 * d2);
            // edge_diff_map
            const float ex = fabsf(a2 - m2), ey = fabsf(a1 - m1);
            const float d1 = __fdividef(ex - ey, 1.0f + ey);
            const float art = fmaxf(d1, 0.0f), det = fmaxf(-d1, 0.0f);
            const float a2_ = art * art, l2 = det * det;
            acc[2] += (double)art;
            acc[3] += (double)(a2_ * a2_);
            acc[4] += (double)det;
            acc[5] += (double)(l2 * l2);
        }
    };
    // rows of a batch are dealt over the warps: warp p takes rows p, p + NW, ...
    auto map_rows = [&](int t) {
#pragma unroll
        for (int r0 = 0; r0 < VP_BATCH; r0 += NW) {
            const int r = r0 + p;
            if (r < VP_BATCH) map_row(t, r);
        }
    };

    if (TMA) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < VP_SLOTS; i++) mbar_init(&s_bar[i], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    issue(0);
    issue(1);
    RGState st = {0, 0, 0, 0, 0, 0};
    float ring[2 * VP_BATCH];
#pragma unroll
    for (int j = 0; j < 2 * VP_BATCH; j++) ring[j] = 0.0f;
    float* vout = MODE == S2_REF ? vref + ((u * 3 + c) * 2 + p) * n : nullptr;
    for (int t0 = 0; t0 < nbatch; t0 += 2) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int t = t0 + half;
            if (t < nbatch) {   // block-uniform
                if (TMA) {
                    mbar_wait(&s_bar[t & (VP_SLOTS - 1)], (unsigned)(t / VP_SLOTS) & 1u);
                    fence_proxy_async();   // this thread's reads of slot (t-2) % 4 precede its TMA refill below
                } else {
                    cp_async_wait<1>();
                }
                __syncthreads();   // batch t landed; s_v of batch t-1 complete; slot (t-2) % 4 free
                issue(t + 2);
                const float* in = s_ld + (t & (VP_SLOTS - 1)) * SLOT_FLOATS + p * VP_BATCH * VP_COLS + lane;
#pragma unroll
                for (int r = 0; r < VP_BATCH; r++) {
                    const float rv = in[r * VP_COLS];
                    const float l = ring[half * VP_BATCH + r];
                    ring[half * VP_BATCH + r] = rv;
                    const float o = rg_step(st, l + rv);
                    if (MODE == S2_REF) {
                        const int y = t * VP_BATCH + r - 4;
                        if (y >= 0 && y < h && active) vout[(size_t)y * w + x] = o;
                    } else {
                        s_v[half][r][p][lane] = o;
                    }
                }
                if (MODE != S2_REF && t > 0) map_rows(t - 1);
            }
        }
    }
    if (MODE == S2_REF) return;
    __syncthreads();
    map_rows(nbatch - 1);
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
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_ALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_ALL>::HP_SMEM));
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_REF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_REF>::HP_SMEM));
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_PAIR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_PAIR>::HP_SMEM));
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_ALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_ALL>::HP_SMEM_TMA));
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_REF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_REF>::HP_SMEM_TMA));
    CE_CUDA(cudaFuncSetAttribute(k_s2_hpass<S2_PAIR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2Mode<S2_PAIR>::HP_SMEM_TMA));
}

size_t ssim2_workspace_per_pair(size_t w, size_t h) {
    size_t n = w * h;
    // xyb 6n (two images), row-pass planes 15n, next-scale linear ping-pong 2 x 2 x 3 x n/4, partials
    size_t bytes = (6 * n + 15 * n) * 4 + (2 * 3 * ((w + 1) / 2) * ((h + 1) / 2)) * 4 * 2 + 3 * cdiv(w, VP_COLS) * 6 * 8 + 4096;
    return bytes;
}

int ssim2_run(Context& c, const float* lin_in, size_t R, const int* ridx, size_t B, size_t w, size_t h, double* d_sums,
              float* dbg_planes) {
    size_t mark = c.arena.mark();
    const size_t n0 = w * h, NI = R + B;
    if (NI > 65535 || B * 3 > 65535) throw CudaError("ssimulacra2 sub-batch too large for one launch");
    float* xyb = c.arena.alloc<float>(NI * 3 * n0);
    // shared references (on average >= 2 distortions each): reference-side blurs once per distinct reference
    const bool split = 2 * R <= B;
    float* hb = nullptr;    // [B][3][5][n]  (S2_ALL)
    float* hbr = nullptr;   // [R][3][2][n]  row pass of {i1, i1^2}
    float* vref = nullptr;  // [R][3][2][n]  finished mu1, blur(i1^2)
    float* hbp = nullptr;   // [B][3][3][n]  row pass of {i2, i2^2, i1*i2}
    if (split) {
        hbr = c.arena.alloc<float>(R * 6 * n0);
        vref = c.arena.alloc<float>(R * 6 * n0);
        hbp = c.arena.alloc<float>(B * 9 * n0);
    } else {
        hb = c.arena.alloc<float>(B * 15 * n0);
    }
    const size_t ow0 = (w + 1) / 2, oh0 = (h + 1) / 2;
    float* nl[2];  // ping-pong next-scale linear buffers, [NI][3][n/4]
    for (int i = 0; i < 2; i++) nl[i] = c.arena.alloc<float>(NI * 3 * ow0 * oh0);
    const int nblk0 = cdiv(w, VP_COLS);
    double* partials = c.arena.alloc<double>(B * 3 * nblk0 * 6);

    const float* l = lin_in;
    size_t cw = w, ch = h;
    int ns = 0;
    for (int scale = 0; scale < 6; scale++) {
        const size_t n = cw * ch;
        const size_t ow = (cw + 1) / 2, oh = (ch + 1) / 2;
        // does a next scale exist? (check on the current size, as upstream does at the top of its loop)
        const bool has_next = (scale + 1 < 6) && !(cw < 8 || ch < 8);
        float* d = nl[scale & 1];
        {
            dim3 grid(cdiv(ow, 64), cdiv(oh, 4), (unsigned)NI);
            CE_LAUNCH(c, "k_s2_xyb_down", (double)NI * 4 * (6 * n + (has_next ? 3 * ow * oh : 0)),
                      k_s2_xyb_down<<<grid, 256, 0, c.stream>>>(l, (int)cw, (int)ch, (int)ow, (int)oh, n, ow * oh, xyb, d,
                                                                 has_next ? 1 : 0));
        }
        const int vecf = (cw % 4 == 0) ? 1 : 0;
        const int nblk = cdiv(cw, VP_COLS);
        float* dbg = (dbg_planes && scale == 0) ? dbg_planes : nullptr;
        if (split) {
            dim3 ghr(cdiv(ch, HP_ROWS), (unsigned)(R * 3)), ghp(cdiv(ch, HP_ROWS), (unsigned)(B * 3));
            dim3 gvr(nblk, (unsigned)(R * 3)), gvp(nblk, (unsigned)(B * 3));
            CUtensorMap mx;
            memset(&mx, 0, sizeof(mx));
            const bool htma = tma_enabled(4) && tma_plane_map(&mx, xyb, cw, ch, NI * 3, 32, HP_ROWS, 1, true);
            if (htma)
                CE_LAUNCH(c, "k_s2_hpass<ref>", (double)R * 3 * 3 * n * 4,
                          k_s2_hpass<S2_REF, true><<<ghr, 64, S2Mode<S2_REF>::HP_SMEM_TMA, c.stream>>>(xyb, R, ridx, hbr, (int)cw, (int)ch, n, vecf, mx));
            else
                CE_LAUNCH(c, "k_s2_hpass<ref>", (double)R * 3 * 3 * n * 4,
                          k_s2_hpass<S2_REF, false><<<ghr, 64, S2Mode<S2_REF>::HP_SMEM, c.stream>>>(xyb, R, ridx, hbr, (int)cw, (int)ch, n, vecf, mx));
            VpassMaps mr, mp;
            memset(&mr, 0, sizeof(mr));
            memset(&mp, 0, sizeof(mp));
            const bool tma = tma_enabled(1) && tma_plane_map(&mr.hb, hbr, cw, ch, R * 6, VP_COLS, VP_BATCH, 2) &&
                             tma_plane_map(&mp.hb, hbp, cw, ch, B * 9, VP_COLS, VP_BATCH, 3) &&
                             tma_plane_map(&mp.xyb, xyb, cw, ch, NI * 3, VP_COLS, VP_BATCH, 1) &&
                             tma_plane_map(&mp.vr, vref, cw, ch, R * 6, VP_COLS, VP_BATCH, 2);
            if (tma)
                CE_LAUNCH(c, "k_s2_vpass<ref>", (double)R * 3 * 4 * n * 4,
                          k_s2_vpass<S2_REF, true><<<gvr, 64, 0, c.stream>>>(xyb, R, ridx, hbr, vref, (int)cw, (int)ch, n, nullptr, nullptr, vecf, mr));
            else
                CE_LAUNCH(c, "k_s2_vpass<ref>", (double)R * 3 * 4 * n * 4,
                          k_s2_vpass<S2_REF, false><<<gvr, 64, 0, c.stream>>>(xyb, R, ridx, hbr, vref, (int)cw, (int)ch, n, nullptr, nullptr, vecf, mr));
            if (htma)
                CE_LAUNCH_SHARED(c, "k_s2_hpass<pair>", (double)(B * 4 + R) * 3 * n * 4, (double)B * 3 * 5 * n * 4,
                          k_s2_hpass<S2_PAIR, true><<<ghp, 96, S2Mode<S2_PAIR>::HP_SMEM_TMA, c.stream>>>(xyb, R, ridx, hbp, (int)cw, (int)ch, n, vecf, mx));
            else
                CE_LAUNCH_SHARED(c, "k_s2_hpass<pair>", (double)(B * 4 + R) * 3 * n * 4, (double)B * 3 * 5 * n * 4,
                          k_s2_hpass<S2_PAIR, false><<<ghp, 96, S2Mode<S2_PAIR>::HP_SMEM, c.stream>>>(xyb, R, ridx, hbp, (int)cw, (int)ch, n, vecf, mx));
            if (tma)
                CE_LAUNCH_SHARED(c, "k_s2_vpass<pair>", (double)(B * 4 + R * 3) * 3 * n * 4, (double)B * 3 * 7 * n * 4,
                          k_s2_vpass<S2_PAIR, true><<<gvp, 96, 0, c.stream>>>(xyb, R, ridx, hbp, vref, (int)cw, (int)ch, n, partials, dbg, vecf, mp));
            else
                CE_LAUNCH_SHARED(c, "k_s2_vpass<pair>", (double)(B * 4 + R * 3) * 3 * n * 4, (double)B * 3 * 7 * n * 4,
                          k_s2_vpass<S2_PAIR, false><<<gvp, 96, 0, c.stream>>>(xyb, R, ridx, hbp, vref, (int)cw, (int)ch, n, partials, dbg, vecf, mp));
        } else {
            dim3 gh(cdiv(ch, HP_ROWS), (unsigned)(B * 3)), gv(nblk, (unsigned)(B * 3));
            CUtensorMap mx;
            memset(&mx, 0, sizeof(mx));
            if (tma_enabled(4) && tma_plane_map(&mx, xyb, cw, ch, NI * 3, 32, HP_ROWS, 1, true))
                CE_LAUNCH(c, "k_s2_hpass", (double)B * 3 * 7 * n * 4,
                          k_s2_hpass<S2_ALL, true><<<gh, 160, S2Mode<S2_ALL>::HP_SMEM_TMA, c.stream>>>(xyb, R, ridx, hb, (int)cw, (int)ch, n, vecf, mx));
            else
                CE_LAUNCH(c, "k_s2_hpass", (double)B * 3 * 7 * n * 4,
                          k_s2_hpass<S2_ALL, false><<<gh, 160, S2Mode<S2_ALL>::HP_SMEM, c.stream>>>(xyb, R, ridx, hb, (int)cw, (int)ch, n, vecf, mx));
            VpassMaps ma;
            memset(&ma, 0, sizeof(ma));
            const bool tma = tma_enabled(1) && tma_plane_map(&ma.hb, hb, cw, ch, B * 15, VP_COLS, VP_BATCH, 5) &&
                             tma_plane_map(&ma.xyb, xyb, cw, ch, NI * 3, VP_COLS, VP_BATCH, 1);
            if (tma)
                CE_LAUNCH(c, "k_s2_vpass", (double)B * 3 * 7 * n * 4,
                          k_s2_vpass<S2_ALL, true><<<gv, 160, 0, c.stream>>>(xyb, R, ridx, hb, nullptr, (int)cw, (int)ch, n, partials, dbg, vecf, ma));
            else
                CE_LAUNCH(c, "k_s2_vpass", (double)B * 3 * 7 * n * 4,
                          k_s2_vpass<S2_ALL, false><<<gv, 160, 0, c.stream>>>(xyb, R, ridx, hb, nullptr, (int)cw, (int)ch, n, partials, dbg, vecf, ma));
        }
        {
            size_t total = B * 3 * 6;
            CE_LAUNCH(c, "k_s2_reduce", (double)total * (nblk + 1) * 8,
                      k_s2_reduce<<<cdiv(total, 128), 128, 0, c.stream>>>(partials, nblk, total, scale, d_sums));
        }
        CE_CUDA(cudaGetLastError());
        ns++;
        if (!has_next) break;
        l = d; cw = ow; ch = oh;
    }
    c.arena.release(mark);
    return ns;
}

}  // namespace ce
