// Micro-benchmark: issue rate of packed FP32 FMA (FFMA2) vs scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float c, float2 v, float2 a) {
    float2 cc = make_float2(c, c), r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(cc)), "l"(reinterpret_cast<unsigned long long&>(v)),
          "l"(reinterpret_cast<unsigned long long&>(a)));
    return r;
}
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, const float* taps, int iters) {
    float2 acc[16], v[4];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(i, -i);
    for (int i = 0; i < 4; ++i) v[i] = make_float2(threadIdx.x + i + 1.5f, threadIdx.x - i - 0.5f);
    float t0 = taps[0], t1 = taps[1], t2 = taps[2], t3 = taps[3];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {            // FFMA2: 4 per accumulator
                acc[i] = fma2(t0, v[0], acc[i]); acc[i] = fma2(t1, v[1], acc[i]);
                acc[i] = fma2(t2, v[2], acc[i]); acc[i] = fma2(t3, v[3], acc[i]);
            } else {                    // scalar FFMA: the same 8 FMAs per accumulator
                acc[i].x = fmaf(t0, v[0].x, acc[i].x); acc[i].y = fmaf(t0, v[0].y, acc[i].y);
                acc[i].x = fmaf(t1, v[1].x, acc[i].x); acc[i].y = fmaf(t1, v[1].y, acc[i].y);
                acc[i].x = fmaf(t2, v[2].x, acc[i].x); acc[i].y = fmaf(t2, v[2].y, acc[i].y);
                acc[i].x = fmaf(t3, v[3].x, acc[i].x); acc[i].y = fmaf(t3, v[3].y, acc[i].y);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out, *taps;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    cudaMalloc(&taps, 16);
    float h[4] = {0.5f, 0.25f, -0.125f, 0.0625f};
    cudaMemcpy(taps, h, 16, cudaMemcpyHostToDevice);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int warps = 4; warps <= 16; warps *= 2)
    for (int mode = 0; mode < 2; ++mode) {
        const int iters = 4000, threads = warps * 32, blocks = 148;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(out, taps, iters); else k<1><<<blocks, threads>>>(out, taps, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = (double)blocks * threads * iters * 16 * 8;
        printf("%s warps/SM=%2d  %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", mode == 0 ? "FFMA2" : "FFMA ",
               warps, ms, fmas / ms / 1e9, fmas / ms / 1e3 / 148 / (clk / 1e3) , clk / 1000);
    }
    return 0;
}
