// Asynchronous bulk copies (global -> shared, completion on an mbarrier) and the mbarrier primitives the staged
// kernels use.  Device build: PTX (cp.async.bulk = SASS UBLKCP, the TMA unit without a tensor map).  Host emulator
// (DTCWT_EMU): the copy is a memcpy and the barriers are no-ops -- tests/emu runs the threads of a CTA in turn and
// places the producer calls where their data is needed.
#pragma once
#include "common.cuh"

#ifdef DTCWT_EMU
#include <string.h>
#endif

namespace dtcwt {

#ifdef DTCWT_EMU
typedef uint64_t Mbar;
inline void mbar_init(Mbar*, uint32_t) {}
inline void mbar_expect_tx(Mbar*, uint32_t) {}
inline void mbar_arrive(Mbar*) {}
inline void mbar_wait(Mbar*, uint32_t) {}
inline void bulk_copy(void* dst, const void* src, uint32_t bytes, Mbar*) { memcpy(dst, src, bytes); }
inline void warp_release(Mbar*, int) {}
#else
typedef uint64_t Mbar;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(Mbar* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(Mbar* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(Mbar* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(Mbar* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(Mbar* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();      // a lost transaction must not hang the GPU
    }
}
// 1-D bulk copy: dst (shared), src (global) and bytes are multiples of 16
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, Mbar* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// every lane of the warp has finished reading a stage (its shared-memory loads have returned): lane 0 arrives for all
__device__ __forceinline__ void warp_release(Mbar* bar, int tid) {
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

}  // namespace dtcwt
