// Micro-benchmark: throughput of cp.async.bulk (1-D TMA copies, global -> shared) as a function of the copy size, with
// the bytes in flight per CTA held constant.  Question it answers: is there a per-request cost that makes many 1 KB
// row segments slower than fewer, larger copies?  (round 2: the staged level-1 inverse issues 8 x 1 KB per quad row.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulkcopy.bin bulkcopy.cu && ./bulkcopy.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool trywait(uint64_t* b, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
// NST stages of STAGE bytes; each stage is filled by STAGE / SZ copies of SZ bytes taken from `nstreams` different rows
template <int NST, int STAGE>
__global__ void __launch_bounds__(256) k(const char* src, size_t row_pitch, int sz, int iters, float* sink) {
    extern __shared__ __align__(128) char sm[];
    __shared__ uint64_t full[NST];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int ncopy = STAGE / sz;
    const char* base = src + (size_t)blockIdx.x * iters * sz;            // a CTA walks along its own columns of the rows
    float acc = 0.f;
    for (int it = 0; it < iters + NST - 1; ++it) {
        if (tid == 0 && it < iters) {
            const int s = it % NST;
            expect(&full[s], STAGE);
            for (int c = 0; c < ncopy; ++c) bulk(sm + s * STAGE + c * sz, base + (size_t)c * row_pitch + (size_t)it * sz, sz, &full[s]);
        }
        const int j = it - (NST - 1);
        if (j >= 0) {
            const int s = j % NST;
            while (!trywait(&full[s], (j / NST) & 1)) {}
            acc += reinterpret_cast<const float*>(sm + s * STAGE)[tid];
            __syncthreads();                                               // stage free again
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}
int main() {
    const size_t row_pitch = 256u << 20;                                   // 8 rows of 256 MiB
    char* src; float* sink;
    cudaMalloc(&src, 8 * row_pitch); cudaMalloc(&sink, 4);
    cudaMemset(src, 1, 8 * row_pitch);
    const int ctas = 296, STAGE = 8192, NST = 4;
    cudaFuncSetAttribute(k<NST, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STAGE);
    for (int sz = 1024; sz <= STAGE; sz *= 2) {
        const int iters = (int)(row_pitch / ctas / sz / 2);                // stay inside a row
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<NST, STAGE><<<ctas, 256, NST * STAGE>>>(src, row_pitch, sz, iters, sink);
        cudaEventRecord(e0);
        k<NST, STAGE><<<ctas, 256, NST * STAGE>>>(src, row_pitch, sz, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)ctas * iters * STAGE;
        printf("copy size %5d B x %d per 8 KB stage, %d stages in flight per CTA, 2 CTAs/SM: %.0f GB/s (%s)\n", sz, STAGE / sz, NST - 1, bytes / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
