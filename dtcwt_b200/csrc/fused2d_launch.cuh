// Device-side launch of the fused 2-D kernels: TMA staging of the forward input tile,
// phase sequencing, dynamic shared memory opt-in.  Device build only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "async_copy.cuh"
#include "fused2d.cuh"
#include "stream2d.cuh"

namespace dtcwt {

// ------------------------------------------------------------------ TMA (PTX); mbarrier primitives: async_copy.cuh
// box [1][rows][cols] of a [n][rows][cols] tensor -> dense smem tile; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// the same box, only as far as L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <class K, int PH, bool DONE = (PH >= K::kPhases)>
struct PhaseRunner {
    static __device__ __forceinline__ void run(const typename K::Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (PH > 0) __syncthreads();
        K::template phase<PH>(a, sm, bx, by, bz, tid);
        PhaseRunner<K, PH + 1>::run(a, sm, bx, by, bz, tid);
    }
};
template <class K, int PH>
struct PhaseRunner<K, PH, true> {
    static __device__ __forceinline__ void run(const typename K::Args&, float*, int, int, int, int) {}
};

extern __shared__ __align__(128) float fused_smem[];

// Persistent forward tile kernel: a CTA walks over tiles blockIdx.x, blockIdx.x + gridDim.x, ... (column index fastest,
// so CTAs running side by side work on neighbouring tiles).  The input tile is only needed until the row pass has
// turned it into A / B, so the TMA load of the NEXT tile is issued right after the row pass and lands in the same
// buffer while the column pass -- 60 % of the arithmetic -- runs.
template <class K>
__global__ void __launch_bounds__(kFusedThreads, K::kMinBlocks)
fwd2d_kernel(const __grid_constant__ typename K::Args a, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    if (!K::kPersistent) {
        // one CTA per tile on a 3-D grid: the tile coordinates are the block indices -- no integer divisions in front of
        // the TMA issue (they were 10 % of the level-1 kernel's stall samples, profiles/r2_04)
        const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
        if (a.use_tma) {
            if (tid == 0) {
                mbar_init(&bar, 1);
                fence_mbar_init();
                mbar_expect_tx(&bar, (uint32_t)(K::RX * K::CX * sizeof(float)));
                tma_load_3d(fused_smem, &tmap, K::col0(bx) - a.pc_lo, K::row0(by) - a.pr_lo, bz, &bar);
                if (a.prefetch > 0) {
                    // the CTA that will take over this one's slot loads the tile about one resident wave further on: fetch it
                    // into L2 now, so that its TMA load finds it there (DRAM latency off the head of every CTA's life)
                    const int tc = gridDim.x, tr = gridDim.y;
                    const int lin = bx + tc * (by + tr * bz) + a.prefetch;
                    if (lin < tc * tr * (int)gridDim.z) {
                        const int q = lin / tc;
                        tma_prefetch_3d(&tmap, K::col0(lin - q * tc) - a.pc_lo, K::row0(q % tr) - a.pr_lo, q / tr);
                    }
                }
            }
            __syncthreads();
            mbar_wait(&bar, 0);            // every thread observes the completion itself: the tile is visible to it
        } else {
            K::template phase<0>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
        }
        // interior tiles (the CTA-uniform common case) have nothing to patch and skip those two barriers
        if (K::tile_on_edge(a, bx, by)) {
            K::template phase<1>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
            K::template phase<2>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
        }
        K::template phase<3>(a, fused_smem, bx, by, bz, tid);
        __syncthreads();
        K::template phase<4>(a, fused_smem, bx, by, bz, tid);
        return;
    }
    const int tc = K::tiles_c(a), tr = K::tiles_r(a);
    const int ntiles = tc * tr * a.n;
    if (a.use_tma) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0 && (int)blockIdx.x < ntiles) {
            const int t = blockIdx.x, bx = t % tc, by = (t / tc) % tr, bz = t / (tc * tr);
            mbar_expect_tx(&bar, (uint32_t)(K::RX * K::CX * sizeof(float)));
            tma_load_3d(fused_smem, &tmap, K::col0(bx) - a.pc_lo, K::row0(by) - a.pr_lo, bz, &bar);
        }
    }
    uint32_t parity = 0;
    // tile coordinates advance incrementally (the stride gridDim.x decomposed once): no division per tile, in particular
    // none in front of the prefetch of the next tile
    const int sx = (int)(gridDim.x % tc), sq = (int)(gridDim.x / tc), sy = sq % tr, sz = sq / tr;
    int bx = (int)(blockIdx.x % tc), by = (int)((blockIdx.x / tc) % tr), bz = (int)(blockIdx.x / (tc * tr));
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int nx = bx + sx, ny = by + sy, nz = bz + sz;
        if (nx >= tc) { nx -= tc; ++ny; }
        if (ny >= tr) { ny -= tr; ++nz; }
        if (a.use_tma) {
            mbar_wait(&bar, parity);
            parity ^= 1u;
        } else {
            K::template phase<0>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
        }
        if (K::tile_on_edge(a, bx, by)) {
            K::template phase<1>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
            K::template phase<2>(a, fused_smem, bx, by, bz, tid);
            __syncthreads();
        }
        K::template phase<3>(a, fused_smem, bx, by, bz, tid);
        if (a.use_tma) fence_proxy_async_smem();      // this thread's reads / patch writes of the tile precede the next TMA fill
        __syncthreads();
        if (a.use_tma && tid == 0 && t + (int)gridDim.x < ntiles) {
            mbar_expect_tx(&bar, (uint32_t)(K::RX * K::CX * sizeof(float)));
            tma_load_3d(fused_smem, &tmap, K::col0(nx) - a.pc_lo, K::row0(ny) - a.pr_lo, nz, &bar);
        }
        K::template phase<4>(a, fused_smem, bx, by, bz, tid);
        __syncthreads();
        bx = nx; by = ny; bz = nz;
    }
}

template <class K>
__global__ void __launch_bounds__(kFusedThreads, K::kMinBlocks) inv2d_kernel(const __grid_constant__ typename K::Args a) {
    PhaseRunner<K, 0>::run(a, fused_smem, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
}

// streaming level-1 inverse (stream2d.cuh): warm-up period, then column pass / row pass per period
template <class K>
__global__ void __launch_bounds__(kStreamThreads, K::kMinBlocks) invs1_kernel(const __grid_constant__ typename K::Args a) {
    const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, tid = threadIdx.x;
    typename K::Thread th;
    K::init(a, th, bx, by, bz, tid, fused_smem);
    const int np = K::run_periods(a, by);
    for (int p = 0; p < np; ++p) {
        K::cols(a, th, fused_smem, bx, by, bz, tid, p);
        if (p > 0) {
            __syncthreads();
            K::rows(a, fused_smem, bx, by, bz, tid, p);
            __syncthreads();
        }
    }
}

// streaming level-1 inverse with bulk-copy staged inputs (stream2d.cuh: InvS1T)
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kStreamThreads) : "memory"); }

template <class K>
__global__ void __launch_bounds__(K::kLaunchThreads, K::kMinBlocks) invs1t_kernel(const __grid_constant__ typename K::Args a) {
    __shared__ __align__(8) Mbar full[K::NSTAGE], empty[K::NSTAGE];
    const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < K::NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kStreamThreads / 32);      // one arrival per consumer warp
        }
        fence_mbar_init();
    }
    __syncthreads();
    typename K::Pipe pipe;
    pipe.full = full;
    pipe.empty = empty;
    if (tid >= kStreamThreads) {                 // producer warp: one lane walks over the steps of the run, as far ahead as the ring allows
        if (tid == kStreamThreads) {             // (eight lanes with one stream each measured slower: 1.72 vs 1.35 ms, profiles/r2_02)
            typename K::Stream st[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) K::stream_setup(a, st[i], bx, bz, i);
            const int total = K::total_steps(a, by);
            for (int g = 0; g < total; ++g) {
                K::step_begin(st[0], pipe, g);
#pragma unroll
                for (int i = 0; i < 8; ++i) K::step_copy(a, st[i], fused_smem, pipe, by, g);
            }
        }
        return;
    }
    typename K::Thread th;
    K::init(a, th, fused_smem, pipe, bx, by, bz, tid);
    const int np = K::run_periods(a, by);
    for (int p = 0; p < np; ++p) {
        K::cols(a, th, fused_smem, pipe, bx, by, bz, tid, p);
        if (p > 0) {
            consumer_barrier();                  // the eight consumer warps only (named barrier 1)
            K::rows(a, fused_smem, bx, by, bz, tid, p);
            consumer_barrier();
        }
    }
}

// streaming level-1 forward (stream2d.cuh).  Input rows of period p+1 are fetched by TMA while period p computes:
// two buffers, one mbarrier each.  Periods that reach above or below the image are staged with plain loads at
// the symmetrically extended row indices instead (two or three periods per run of an image's first / last run).
template <class K>
__device__ __forceinline__ void fwds1_issue(const typename K::Args& a, float* sm, uint64_t* bar, const CUtensorMap* box,
                                            int bx, int by, int bz, int p) {
    fence_proxy_async_smem();
    mbar_expect_tx(&bar[p & 1], (uint32_t)(K::XBUF * sizeof(float)));
    tma_load_3d(sm + (p & 1) * K::XBUF, box, K::col_base(bx), K::row_base(a, by, p), bz, &bar[p & 1]);
}

template <class K>
__global__ void __launch_bounds__(K::kThreads) __maxnreg__(K::kMaxRegs)
fwds1_kernel(const __grid_constant__ typename K::Args a, const __grid_constant__ CUtensorMap tm_box) {
    __shared__ __align__(8) uint64_t bar[2];
    const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, tid = threadIdx.x;
    typename K::Thread th;
    K::init(th);
    const int np = K::run_periods(a, by);
    if (a.use_tma) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0 && K::rows_inside(a, by, 0)) fwds1_issue<K>(a, fused_smem, bar, &tm_box, bx, by, bz, 0);
    }
    const bool patch = !K::cols_inside(a, bx);
    uint32_t phase = 0;                       // bit b: parity the next wait on bar[b] uses
    for (int p = 0; p < np; ++p) {
        const bool tma = a.use_tma && K::rows_inside(a, by, p);
        if (a.use_tma && tid == 0 && p + 1 < np && K::rows_inside(a, by, p + 1))
            fwds1_issue<K>(a, fused_smem, bar, &tm_box, bx, by, bz, p + 1);
        if (tma) {
            mbar_wait(&bar[p & 1], (phase >> (p & 1)) & 1u);
            phase ^= 1u << (p & 1);
            if (patch) {
                K::patch_cols(a, fused_smem, bx, p, tid);
                __syncthreads();
            }
        } else {
            K::load_plain(a, fused_smem, bx, by, bz, p, tid);
            __syncthreads();
        }
        K::rows(a, fused_smem, p, tid);
        __syncthreads();
        K::cols(a, th, fused_smem, bx, by, bz, tid, p);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                           CUtensorMapFloatOOBfill);

static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = []() -> TensorMapEncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (TensorMapEncodeTiledFn)p;
    }();
    return fn;
}

static bool tma_disabled_by_env() {
    static const bool off = []() {
        const char* e = getenv("DTCWT_B200_NO_TMA");
        return e && e[0] && e[0] != '0';
    }();
    return off;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel instance and device instead of once per launch (the
// attribute is sticky; a small transform is bound by exactly this kind of host work).  TAG makes one flag set per call site.
template <class TAG>
static cudaError_t smem_opt_in(const void* kernel, size_t smem) {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

template <class K>
static int launch_fwd2d(typename K::Args& a, void* stream) {
    const size_t smem = (size_t)K::kSmemFloats * sizeof(float);
    cudaError_t e = smem_opt_in<K>((const void*)fwd2d_kernel<K>, smem);
    if (e != cudaSuccess) return (int)e;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    a.use_tma = 0;
    a.prefetch = 0;
    // TMA needs a 16-byte aligned base and row pitch, and every box must START on a 16-byte boundary of its row: the
    // first column of a tile is a multiple of 4 minus pc_lo, so two replicated columns (ext_mode 8 of the 3-D transform)
    // rule it out -- an unaligned start never completes its transaction (measured: the mbarrier wait times out).
    // Other shapes take the plain-load staging phase.
    if (!tma_disabled_by_env() && (a.cols % 4) == 0 && (a.pc_lo % 4) == 0 && ((uintptr_t)a.x % 16) == 0 && tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t)a.cols, (cuuint64_t)a.rows, (cuuint64_t)(a.n > 0 ? a.n : 1)};
        const cuuint64_t strides[2] = {(cuuint64_t)a.cols * 4, (cuuint64_t)a.cols * a.rows * 4};
        const cuuint32_t box[3] = {(cuuint32_t)K::CX, (cuuint32_t)K::RX, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.x, dims, strides, box,
                                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) a.use_tma = 1;
    }
    if (a.n == 0) return DTCWT_B200_OK;
    const int64_t ntiles = (int64_t)K::tiles_c(a) * K::tiles_r(a) * a.n;
    if (ntiles > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    {
        static int c_sms[64] = {}, c_per_sm[64] = {};         // per kernel instance and device: asked once
        if (dev >= 0 && dev < 64 && c_sms[dev] > 0) {
            sms = c_sms[dev]; per_sm = c_per_sm[dev];
        } else {
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd2d_kernel<K>, kFusedThreads, smem)) != cudaSuccess)
                return (int)e;
            if (dev >= 0 && dev < 64) { c_per_sm[dev] = per_sm; c_sms[dev] = sms; }
        }
    }
    if (!K::kPersistent) {
        if (K::tiles_r(a) > 65535 || a.n > 65535) return DTCWT_B200_EUNSUPPORTED;      // grid.y / grid.z limits
        // L2 prefetch distance in percent of a resident wave of CTAs (DTCWT_B200_FWD_PREFETCH, 0: off).  Measured on the level-1
        // forward, 16 x 4096^2: off 1.021 ms, 50 % 0.991, 100 % 0.995, 200 % 1.026 (profiles/r3_01)
        const char* pfe = getenv("DTCWT_B200_FWD_PREFETCH");           // read per call, like the other experiment switches
        const int pf = pfe ? atoi(pfe) : 50;
        a.prefetch = (a.use_tma && ntiles <= 0x3fffffff) ? (int)((int64_t)sms * (per_sm > 0 ? per_sm : 1) * pf / 100) : 0;
        const dim3 grid((unsigned)K::tiles_c(a), (unsigned)K::tiles_r(a), (unsigned)a.n);
        fwd2d_kernel<K><<<grid, kFusedThreads, smem, (cudaStream_t)stream>>>(a, map);
        return (int)cudaGetLastError();
    }
    int64_t ctas = (int64_t)sms * (per_sm > 0 ? per_sm : 1);          // one wave of resident CTAs, each walks over tiles
    if (ctas > ntiles) ctas = ntiles;
    fwd2d_kernel<K><<<(unsigned)ctas, kFusedThreads, smem, (cudaStream_t)stream>>>(a, map);
    return (int)cudaGetLastError();
}

template <class K>
static int launch_inv2d(typename K::Args& a, void* stream) {
    const size_t smem = (size_t)K::kSmemFloats * sizeof(float);
    cudaError_t e = smem_opt_in<K>((const void*)inv2d_kernel<K>, smem);
    if (e != cudaSuccess) return (int)e;
    if (a.n == 0) return DTCWT_B200_OK;
    if (K::tiles_r(a) > 65535 || a.n > 65535) return DTCWT_B200_EUNSUPPORTED;      // grid.y / grid.z limits
    const dim3 grid((unsigned)K::tiles_c(a), (unsigned)K::tiles_r(a), (unsigned)a.n);
    inv2d_kernel<K><<<grid, kFusedThreads, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

template <class K>
static int launch_fwds1(typename K::Args& a, void* stream) {
    const size_t smem = (size_t)K::kSmemFloats * sizeof(float);
    cudaError_t e = smem_opt_in<K>((const void*)fwds1_kernel<K>, smem);
    if (e != cudaSuccess) return (int)e;
    CUtensorMap box;
    memset(&box, 0, sizeof(box));
    a.use_tma = 0;
    // TMA needs a 16-byte aligned base and row pitch; other shapes are staged with plain loads
    if (!tma_disabled_by_env() && (a.cols % 4) == 0 && ((uintptr_t)a.x % 16) == 0 && tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t)a.cols, (cuuint64_t)a.rows, (cuuint64_t)(a.n > 0 ? a.n : 1)};
        const cuuint64_t strides[2] = {(cuuint64_t)a.cols * 4, (cuuint64_t)a.cols * a.rows * 4};
        const cuuint32_t bbox[3] = {(cuuint32_t)K::CXS, (cuuint32_t)K::RING, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = tensor_map_encoder()(&box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.x, dims, strides, bbox,
                                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) a.use_tma = 1;
    }
    if (a.n == 0) return DTCWT_B200_OK;
    if (K::tiles_r(a) > 65535 || a.n > 65535) return DTCWT_B200_EUNSUPPORTED;      // grid.y / grid.z limits
    const dim3 grid((unsigned)K::tiles_c(a), (unsigned)K::tiles_r(a), (unsigned)a.n);
    fwds1_kernel<K><<<grid, K::kThreads, smem, (cudaStream_t)stream>>>(a, box);
    return (int)cudaGetLastError();
}

template <class K>
static int launch_invs1t(typename K::Args& a, void* stream) {
    const size_t smem = (size_t)K::kSmemFloats * sizeof(float);
    cudaError_t e = smem_opt_in<K>((const void*)invs1t_kernel<K>, smem);
    if (e != cudaSuccess) return (int)e;
    if (a.n == 0) return DTCWT_B200_OK;
    if (K::tiles_r(a) > 65535 || a.n > 65535) return DTCWT_B200_EUNSUPPORTED;      // grid.y / grid.z limits
    const dim3 grid((unsigned)K::tiles_c(a), (unsigned)K::tiles_r(a), (unsigned)a.n);
    invs1t_kernel<K><<<grid, K::kLaunchThreads, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

template <class K>
static int launch_invs1(typename K::Args& a, void* stream) {
    const size_t smem = (size_t)K::kSmemFloats * sizeof(float);
    cudaError_t e = smem_opt_in<K>((const void*)invs1_kernel<K>, smem);
    if (e != cudaSuccess) return (int)e;
    if (a.n == 0) return DTCWT_B200_OK;
    if (K::tiles_r(a) > 65535 || a.n > 65535) return DTCWT_B200_EUNSUPPORTED;      // grid.y / grid.z limits
    const dim3 grid((unsigned)K::tiles_c(a), (unsigned)K::tiles_r(a), (unsigned)a.n);
    invs1_kernel<K><<<grid, kStreamThreads, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace dtcwt
