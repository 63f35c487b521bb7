// Micro-benchmark: FP32 issue rates on sm_100a (warp-instructions per clock per SM) for the instruction
// forms the metric kernels are made of.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_w[64];
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-3f + i;
    float b = seed * 0.5f + 1.0f, c = seed * 0.25f + 0.5f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) a[i] = __fmaf_rn(a[i], b, c);                 // FFMA r,r,r
                else if (MODE == 1) a[i] = __fmaf_rn(a[i], c_w[u * 8 + i], c); // FFMA r,c[],r
                else if (MODE == 2) a[i] = a[i] + b;                          // FADD r,r
                else if (MODE == 3) a[i] = a[i] * b;                          // FMUL r,r
                else if (MODE == 4) a[i] = __fmaf_rn(a[i], 1.0001f, c);       // FFMA r,imm,r
                else if (MODE == 5) a[i] = __fmaf_rn(a[i], b, a[(i + 1) & 7]); // FFMA 3 distinct regs
                else if (MODE == 6) a[i] = a[i] + a[(i + 3) & 7];             // FADD two varying regs
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm = 1; blocks_per_sm <= 8; blocks_per_sm *= 2) {
        k<MODE><<<148 * blocks_per_sm, 256>>>(out, 1.0f, 10);
        cudaEventRecord(e0);
        k<MODE><<<148 * blocks_per_sm, 256>>>(out, 1.0f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winstr = (double)148 * blocks_per_sm * 8 * iters * 64;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-28s warps/SM %2d: %.3f G warp-instr/s  = %.2f per clk per SM @ %d MHz nominal\n", name, blocks_per_sm * 8,
               winstr / ms / 1e6, winstr / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    }
    cudaFree(out);
}
int main() {
    float w[64]; for (int i = 0; i < 64; i++) w[i] = 1.0f + i * 1e-4f;
    cudaMemcpyToSymbol(c_w, w, sizeof(w));
    run<0>("FFMA r,r,r (2 loop-inv)");
    run<1>("FFMA r,c[],r");
    run<2>("FADD r,r");
    run<3>("FMUL r,r");
    run<4>("FFMA r,imm,r");
    run<5>("FFMA 3 varying regs");
    run<6>("FADD 2 varying regs");
    return 0;
}
