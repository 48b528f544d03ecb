// Shared definitions for the dtcwt_b200 kernels.
//
// The kernel BODIES in this directory are written so that they also compile as
// plain C++ (DTCWT_EMU): tests/emu builds them with g++ into a host library that
// executes the same index arithmetic thread by thread.  That emulator is test
// infrastructure (no GPU in the build container); the package never loads it.
#pragma once
#include <stdint.h>

// The device library is compiled as four translation units in parallel (build.py: -DDTCWT_PART=0..4), each of which
// emits one group of C-ABI entry points and instantiates only the kernels that group launches.  Without DTCWT_PART
// (the host emulator, or a plain one-file nvcc build) everything is emitted.
#if !defined(DTCWT_PART) || DTCWT_PART == 0
#define DTCWT_EMIT_GENERIC 1          /* version / error strings, generic + per-axis filters, packers */
#endif
#if !defined(DTCWT_PART) || DTCWT_PART == 1
#define DTCWT_EMIT_FWD2D 1            /* fwd2d_level1, fwd2d_levelq */
#endif
#if !defined(DTCWT_PART) || DTCWT_PART == 2
#define DTCWT_EMIT_INV2D_Q 1          /* inv2d_levelq */
#endif
#if !defined(DTCWT_PART) || DTCWT_PART == 3
#define DTCWT_EMIT_INV2D_1 1          /* inv2d_level1 */
#endif
#if !defined(DTCWT_PART) || DTCWT_PART == 4
#define DTCWT_EMIT_FUSED3D 1          /* fwd3d_*, inv3d_* */
#endif

#ifdef DTCWT_EMU
#define DTCWT_HD inline
#define DTCWT_D inline
#include <cmath>
#else
#include <cuda_runtime.h>
#define DTCWT_HD __host__ __device__ __forceinline__
#define DTCWT_D __device__ __forceinline__
#endif

#include "../../include/dtcwt_b200.h"

namespace dtcwt {

constexpr int kMaxTaps = DTCWT_B200_MAX_TAPS;

// Filter taps travel by value inside the kernel parameter block (constant bank).
template <typename T>
struct Taps {
    T v[kMaxTaps];
    int m;
};

// Half-sample symmetric fold of p onto [0, L): ... 1 0 | 0 1 .. L-1 | L-1 L-2 ...
// (reference dtcwt/utils.py:136-153 with minx = -0.5, maxx = L - 0.5).  Any p.
DTCWT_HD int reflect_any(int p, int L) {
    const int P = 2 * L;
    int q = p % P;
    if (q < 0) q += P;
    return (q >= L) ? (P - 1 - q) : q;
}

// Logical (edge-replicated) index -> index into the stored array of `len` samples.
DTCWT_HD int unpad(int p, int pad_lo, int len) {
    int s = p - pad_lo;
    s = s < 0 ? 0 : s;
    return s >= len ? len - 1 : s;
}

template <typename T> DTCWT_HD T fma_t(T a, T b, T c);
template <> DTCWT_HD float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> DTCWT_HD double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

}  // namespace dtcwt
