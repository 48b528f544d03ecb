// Host-side kernel-logic emulator -- TEST INFRASTRUCTURE, never loaded by the package.
//
// The build container has nvcc but no GPU.  This file compiles the SAME kernel
// bodies (dtcwt_b200/csrc/*.cuh, DTCWT_EMU) and the SAME C-ABI wrappers
// (abi_generic.inl, ...) with g++ and runs every "thread" in a loop, so the index
// arithmetic of the CUDA kernels and the whole Python host layer can be exercised
// by `pytest -m "not gpu"`.  "Device" pointers are host pointers here; `stream`
// is ignored.  dtcwt_b200_is_device_build() returns 0 so nothing can mistake this
// for the product library.
#define DTCWT_EMU 1
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <vector>

#include "../../dtcwt_b200/csrc/generic_kernels.cuh"
#include "../../dtcwt_b200/csrc/fused2d.cuh"
#include "../../dtcwt_b200/csrc/stream2d.cuh"
#include "../../dtcwt_b200/csrc/fused3d.cuh"
#include "../../dtcwt_b200/csrc/registration.cuh"
#include "../../dtcwt_b200/csrc/keypoint.cuh"
#include "../../dtcwt_b200/csrc/axis_pass.cuh"

namespace dtcwt {

template <class Elem>
static int launch_1d(const typename Elem::Args& a, void* /*stream*/) {
    const int64_t total = Elem::total(a);
    for (int64_t gid = 0; gid < total; ++gid) Elem::run(a, gid);
    return DTCWT_B200_OK;
}

template <class Elem>
static int launch_1d_2(const typename Elem::Args& a, void* stream) { return launch_1d<Elem>(a, stream); }
template <class T>
static int launch_qtilde(const QtildeArgs<T>& a, void* stream) { return launch_1d<QtildeElem<T> >(a, stream); }   // one thread per pixel on the host

template <class K>
static int launch_axis(const AxisArgs& a, void* /*stream*/) {
    const int64_t total = K::total(a);
    for (int64_t gid = total - 1; gid >= 0; --gid) K::run(a, gid);     // descending: exposes writes outside a thread's outputs
    return DTCWT_B200_OK;
}

template <class K>
static int launch_z3(const Z3Args& a, void* /*stream*/) {
    const int64_t total = K::total(a);
    for (int64_t gid = total - 1; gid >= 0; --gid) K::run(a, gid);
    return DTCWT_B200_OK;
}

// A fused kernel is a sequence of phases separated by block barriers: run every phase for every
// thread of a block before the next one, block by block.  Shared memory is a scratch vector filled
// with NaN so that a read of something no phase wrote shows up in the results.
template <class K, int PH, bool DONE = (PH >= K::kPhases)>
struct EmuPhases {
    static void run(const typename K::Args& a, float* sm, int bx, int by, int bz) {
        for (int tid = 0; tid < K::kThreads; ++tid) K::template phase<PH>(a, sm, bx, by, bz, tid);
        EmuPhases<K, PH + 1>::run(a, sm, bx, by, bz);
    }
};
template <class K, int PH>
struct EmuPhases<K, PH, true> {
    static void run(const typename K::Args&, float*, int, int, int) {}
};

// Tiles run in DESCENDING order: a tile that writes outside its own outputs (a GPU race between CTAs) then
// clobbers results of tiles that already ran instead of being silently overwritten by them.
template <class K>
static int emu_launch(typename K::Args& a) {
    std::vector<float> sm(K::kSmemFloats);
    for (int bz = a.n - 1; bz >= 0; --bz)
        for (int by = K::tiles_r(a) - 1; by >= 0; --by)
            for (int bx = K::tiles_c(a) - 1; bx >= 0; --bx) {
                for (size_t i = 0; i < sm.size(); ++i) sm[i] = NAN;
                EmuPhases<K, 0>::run(a, sm.data(), bx, by, bz);
            }
    return DTCWT_B200_OK;
}

template <class K>
static int launch_fwd2d(typename K::Args& a, void* /*stream*/) {
    a.use_tma = 0;
    return emu_launch<K>(a);
}
template <class K>
static int launch_inv2d(typename K::Args& a, void* /*stream*/) { return emu_launch<K>(a); }

// Streaming kernels keep per-thread state (accumulator ring, prefetched rows) across barriers: one Thread
// object per emulated thread, the period loop of the device kernel replayed with every thread run in turn.
template <class K>
static int launch_invs1(typename K::Args& a, void* /*stream*/) {
    std::vector<float> sm(K::kSmemFloats);
    std::vector<typename K::Thread> th(K::kThreads);
    for (int bz = 0; bz < a.n; ++bz)
        for (int by = 0; by < K::tiles_r(a); ++by)
            for (int bx = 0; bx < K::tiles_c(a); ++bx) {
                for (size_t i = 0; i < sm.size(); ++i) sm[i] = NAN;
                for (int tid = 0; tid < K::kThreads; ++tid) K::init(a, th[tid], bx, by, bz, tid, sm.data());
                const int np = K::run_periods(a, by);
                for (int p = 0; p < np; ++p) {
                    for (int tid = 0; tid < K::kThreads; ++tid) K::cols(a, th[tid], sm.data(), bx, by, bz, tid, p);
                    if (p > 0)
                        for (int tid = 0; tid < K::kThreads; ++tid) K::rows(a, sm.data(), bx, by, bz, tid, p);
                }
            }
    return DTCWT_B200_OK;
}

// Staged streaming kernel: on the device a producer warp fills the ring as far ahead as the empty barriers allow.  Here
// every step of a period is run for all threads in turn and the producer is replayed at its MAXIMUM run-ahead: before
// step g it fills step g + NSTAGE - 1, whose stage was last read in step g - 1 (complete), so every stage index and
// the wrap of the ring are exercised exactly as on the device.
template <class K>
static int launch_invs1t(typename K::Args& a, void* /*stream*/) {
    std::vector<float> sm(K::kSmemFloats);
    std::vector<typename K::Thread> th(K::kThreads);
    typename K::Pipe pipe;
    pipe.full = nullptr;
    pipe.empty = nullptr;
    for (int bz = 0; bz < a.n; ++bz)
        for (int by = 0; by < K::tiles_r(a); ++by)
            for (int bx = 0; bx < K::tiles_c(a); ++bx) {
                for (size_t i = 0; i < sm.size(); ++i) sm[i] = NAN;
                for (int tid = 0; tid < K::kThreads; ++tid) K::init(a, th[tid], sm.data(), pipe, bx, by, bz, tid);
                const int np = K::run_periods(a, by), total = K::total_steps(a, by);
                for (int g = 0; g < K::NSTAGE - 1 && g < total; ++g) K::produce(a, sm.data(), pipe, bx, by, bz, g);
                for (int p = 0; p < np; ++p) {
                    for (int u = 0; u < K::PER; ++u) {
                        const int g = p * K::PER + u + K::NSTAGE - 1;
                        if (g < total) K::produce(a, sm.data(), pipe, bx, by, bz, g);
                        for (int tid = 0; tid < K::kThreads; ++tid) K::step(a, th[tid], sm.data(), pipe, bx, by, bz, tid, p, u);
                    }
                    if (p > 0)
                        for (int tid = 0; tid < K::kThreads; ++tid) K::rows(a, sm.data(), bx, by, bz, tid, p);
                }
            }
    return DTCWT_B200_OK;
}

template <class K>
static int launch_fwds1(typename K::Args& a, void* /*stream*/) {
    std::vector<float> sm(K::kSmemFloats);
    std::vector<typename K::Thread> th(K::kThreads);
    a.use_tma = 0;
    for (int bz = 0; bz < a.n; ++bz)
        for (int by = 0; by < K::tiles_r(a); ++by)
            for (int bx = 0; bx < K::tiles_c(a); ++bx) {
                for (size_t i = 0; i < sm.size(); ++i) sm[i] = NAN;
                for (int tid = 0; tid < K::kThreads; ++tid) K::init(th[tid]);
                const int np = K::run_periods(a, by);
                for (int p = 0; p < np; ++p) {
                    for (int tid = 0; tid < K::kThreads; ++tid) K::load_plain(a, sm.data(), bx, by, bz, p, tid);
                    for (int tid = 0; tid < K::kThreads; ++tid) K::rows(a, sm.data(), p, tid);
                    for (int tid = 0; tid < K::kThreads; ++tid) K::cols(a, th[tid], sm.data(), bx, by, bz, tid, p);
                }
            }
    return DTCWT_B200_OK;
}

}  // namespace dtcwt

#include "../../dtcwt_b200/csrc/abi_generic.inl"
#include "../../dtcwt_b200/csrc/abi_reg.inl"
#include "../../dtcwt_b200/csrc/abi_fused2d.inl"
#include "../../dtcwt_b200/csrc/abi_axis.inl"
#include "../../dtcwt_b200/csrc/abi_fused3d.inl"
#include "../../dtcwt_b200/csrc/abi_chain.inl"

extern "C" {

int dtcwt_b200_is_device_build(void) { return 0; }

const char* dtcwt_b200_error_string(int code) {
    if (code == DTCWT_B200_OK) return "ok";
    if (code == DTCWT_B200_EINVAL) return "dtcwt_b200: invalid argument (shape, tap count or NULL pointer)";
    if (code == DTCWT_B200_EUNSUPPORTED) return "dtcwt_b200: request not supported by this build";
    return "dtcwt_b200[emulator]: unknown error";
}

}  // extern "C"
