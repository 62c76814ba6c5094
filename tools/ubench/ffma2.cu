// Micro-benchmark: issue rate of FFMA vs FFMA2 (packed fp32, sm_100) per SM sub-partition, with 8 independent chains per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }   // 2 FFMA
            else x[i] = __ffma2_rn(x[i], aa, bb);                                              // 1 FFMA2
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
    float* d; cudaMalloc(&d, 1 << 22);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        float c[2];
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
            if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
            cudaMemcpy(&c[mode], d, 4, cudaMemcpyDeviceToHost);
        }
        // per SM: warps * iters * 16 fp32 FMAs per lane-pair-group; report FMA lanes per clock per SM
        const double fma = (double)warps * 32 * iters * 16;
        printf("warps/SM %2d: FFMA %.0f cycles = %.1f FMA/clk/SM | FFMA2 %.0f cycles = %.1f FMA/clk/SM\n", warps, c[0], fma / c[0], c[1], fma / c[1]);
    }
    return 0;
}
