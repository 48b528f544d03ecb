// Fast float32 versions of the three reference filters along one axis of an [outer][len][inner] array:
//   colfilter (lowlevel.py:47-80), coldfilt (:82-154), colifilt (:156-260), symmetric extension (utils.py:136-153).
// They serve everything the fused 2-D kernels do not cover: the 1-D and 3-D transforms, the `_bp` variants, the
// low-level API.  One thread owns V adjacent samples across `inner` (V = 2 when inner is even: packed FFMA2) and NG
// output groups along the axis: it walks down the Q*NG + halo input samples its outputs depend on -- each loaded
// once per thread, coalesced across the threads of a warp -- and scatters them into P*NG register accumulators
// (the same polyphase scatter as the fused kernels, fused2d.cuh).  The generic one-thread-per-output kernels
// (generic_kernels.cuh) remain the fallback for float64, even-length colfilter taps and tap counts without an instance.
#pragma once
#include "fused2d.cuh"

namespace dtcwt {

struct AxisArgs {
    const float* x;
    float* y;
    int64_t outer;
    int inner;
    int len, pad_lo, L;            // stored length, replicated samples before it, logical (padded) length
    int Lout, crop;                // stored output length; logical output `crop` is stored at index 0
    int accumulate;
    PhaseTaps t;
};

DTCWT_D float axis_fma(float c, float v, float acc) { return fmaf(c, v, acc); }
DTCWT_D F2 axis_fma(float c, F2 v, F2 acc) { return fma2(c, v, acc); }
DTCWT_D void axis_zero(float& v) { v = 0.f; }
DTCWT_D void axis_zero(F2& v) { v.x = 0.f; v.y = 0.f; }
DTCWT_D float axis_add(float a, float b) { return a + b; }
DTCWT_D F2 axis_add(F2 a, F2 b) { F2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }

template <class F, int NG_, class VT>
struct AxisPass {
    typedef AxisArgs Args;
    static constexpr int P = F::P, Q = F::Q, NG = NG_;
    static constexpr int V = sizeof(VT) / sizeof(float);
    static constexpr int HL = spec_lo<F>(), HR = spec_hi<F>();
    static constexpr int NR = Q * NG + HL + HR;              // input samples a thread walks over
    static constexpr int NOUT = P * NG;

    static DTCWT_HD int64_t blocks_along(const Args& a) {     // threads along the axis
        const int logical_out = a.Lout + 2 * a.crop;
        return (logical_out + NOUT - 1) / NOUT;
    }
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * blocks_along(a) * (a.inner / V); }

    template <bool INSIDE>
    static DTCWT_D void accumulate(const Args& a, const float* xo, int l0, int s0, VT (&acc)[NOUT]) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            if (fir_row_used<F, NG, HL>(j)) {
                const int s = INSIDE ? s0 + j : unpad(reflect_any(l0 + j, a.L), a.pad_lo, a.len);
                const VT v = *reinterpret_cast<const VT*>(xo + (int64_t)s * a.inner);
#pragma unroll
                for (int ii = 0; ii < NG; ++ii) {
#pragma unroll
                    for (int ph = 0; ph < P; ++ph) {
                        const int num = j - HL - Q * ii - F::b(ph);
                        if (num >= 0 && (num % F::S) == 0 && (num / F::S) < F::K && F::on(ph, num / F::S))
                            acc[P * ii + ph] = axis_fma(a.t.t[ph][num / F::S], v, acc[P * ii + ph]);
                    }
                }
            }
        }
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
        const int ip_n = a.inner / V;
        const int ip = (int)(gid % ip_n);
        const int64_t r = gid / ip_n;
        const int64_t nb = blocks_along(a);
        const int gb = (int)(r % nb);
        const int64_t o = r / nb;
        const float* xo = a.x + o * (int64_t)a.len * a.inner + V * ip;
        VT acc[NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) axis_zero(acc[i]);
        const int l0 = Q * NG * gb - HL;                     // logical index of window sample 0
        // a window that lies inside the stored samples needs no symmetric extension: one add per sample instead of the
        // modulo of reflect_any.  Two copies of the unrolled loop behind one (warp-uniform) branch -- a per-sample select
        // still evaluates the modulo and measured SLOWER than the plain loop (profiles/r2_05).
        const int s0 = l0 - a.pad_lo;
        if (s0 >= 0 && s0 + NR <= a.len) accumulate<true>(a, xo, l0, s0, acc);
        else accumulate<false>(a, xo, l0, s0, acc);
        float* yo = a.y + o * (int64_t)a.Lout * a.inner + V * ip;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const int q = NOUT * gb + i - a.crop;
            if (q >= 0 && q < a.Lout) {
                VT* d = reinterpret_cast<VT*>(yo + (int64_t)q * a.inner);
                *d = a.accumulate ? axis_add(*d, acc[i]) : acc[i];
            }
        }
    }
};

}  // namespace dtcwt
