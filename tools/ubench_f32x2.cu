// Micro-benchmark: packed fp32x2 arithmetic on sm_100a (add/mul/fma .f32x2) vs scalar, alone and mixed with
// integer work.  Reports warp-instructions per clock per SM and fp32 lane-ops per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
    unsigned long long a[8];
    float f[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = pk(seed + threadIdx.x * 1e-3f + i, seed - i); f[i] = seed + i; q[i] = threadIdx.x + i; }
    unsigned long long b = pk(seed * 0.5f + 1.0f, 0.999f), c = pk(0.25f, 0.125f);
    float fb = seed * 0.5f + 1.0f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) a[i] = add2(a[i], b);
                else if (MODE == 1) a[i] = mul2(a[i], b);
                else if (MODE == 2) a[i] = fma2(a[i], b, c);
                else if (MODE == 3) { a[i] = add2(a[i], b); q[i] = q[i] * 3 + it; }          // packed add + IMAD
                else if (MODE == 4) { f[i] = f[i] + fb; q[i] = q[i] * 3 + it; }             // scalar add + IMAD
                else if (MODE == 5) { a[i] = add2(a[i], b); q[i] = (q[i] ^ it) + i; }        // packed add + LOP3/IADD
                else if (MODE == 6) { f[i] = f[i] + fb; q[i] = (q[i] ^ it) + i; }           // scalar add + LOP3/IADD
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi + f[i] + q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double fp_per_iter_instr, double instr_per_slot) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 2000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int bps = 4;
    k<MODE><<<148 * bps, 256>>>(out, 1.0f, 10);
    cudaEventRecord(e0); k<MODE><<<148 * bps, 256>>>(out, 1.0f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double slots = (double)148 * bps * 8 * iters * 64;    // (warp, inner op) slots
    double per_clk = slots / (ms * 1e-3) / 148 / (clk * 1e3);
    printf("%-34s %.2f op-slots/clk/SM -> %.2f warp-instr/clk/SM, %.1f fp32 lane-ops/clk/SM\n", name, per_clk, per_clk * instr_per_slot,
           per_clk * fp_per_iter_instr * 32);
    cudaFree(out);
}
int main() {
    run<0>("add.f32x2", 2, 1); run<1>("mul.f32x2", 2, 1); run<2>("fma.f32x2", 2, 1);
    run<3>("add.f32x2 + IMAD", 2, 2); run<4>("add.f32 + IMAD", 1, 2);
    run<5>("add.f32x2 + LOP3+IADD", 2, 3); run<6>("add.f32 + LOP3+IADD", 1, 3);
    return 0;
}
