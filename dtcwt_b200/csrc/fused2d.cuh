// Fused per-level 2-D DT-CWT kernels (float32).
//
// One CTA transforms one tile of one image through a whole level:
//
//   forward  (reference transform2d.py:112-160)   X tile --TMA--> smem
//            row pass   A = H:h0(X), B = H:h1(X)/sqrt2            (smem -> registers -> smem)
//            column pass LoLo = V:h0(A), q2c(V:h1(A)/sqrt2), q2c(V:h0(B)), q2c(V:h1(B))
//            -> LoLo tile and the six complex sub-bands go straight from registers to HBM
//   inverse  (reference transform2d.py:240-293)   c2q is applied while the tile is loaded,
//            row pass   p1 = H:g0(Z) + H:g1(hl),  p2 = H:g0(lh) + H:g1(hh)
//            column pass out = V:g0(p1) + V:g1(p2)
//
// (The reference filters columns first; the passes commute, only rounding differs.)
// Every input sample is read from HBM once per level (plus the tile halo, which
// neighbouring CTAs find in L2) and every output is written once.
//
// All three reference filters are instances of one polyphase form
//     out[P*i + ph] = sum_{k<K} t[ph][k] * in[Q*i + b(ph) + S*k]
//   colfilter (lowlevel.py:47-80)    P=1 Q=1 S=1 K=m      b = -(m-1)/2
//   coldfilt  (lowlevel.py:82-154)   P=2 Q=4 S=2 K=m      b = -m+2+delta(ph)
//   colifilt  (lowlevel.py:156-260)  P=4 Q=2 S=2 K=m/2    b = -m/2+2+off(ph)
// with the structure (P,Q,S,K,b) fixed at compile time and the taps t[ph][k]
// prepared on the host (abi_fused2d.inl) and passed by value: they live in the
// constant bank and feed FFMA directly.
//
// The bodies are split into phases separated by block barriers; tests/emu runs
// the same phases thread by thread on the host (DTCWT_EMU).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "baked_taps.h"

namespace dtcwt {

#ifdef DTCWT_EMU
struct alignas(8) F2 { float x, y; };
struct alignas(16) F4 { float x, y, z, w; };
#else
typedef float2 F2;
typedef float4 F4;
#endif

constexpr int kFusedThreads = 256;
constexpr int kMaxPhases = 4;
constexpr int kMaxPhaseTaps = 19;

struct PhaseTaps {
    float t[kMaxPhases][kMaxPhaseTaps];
};

// acc += c * v on a pair of floats.  sm_100a has a packed FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2) whose
// multiplier may be a scalar broadcast from a uniform register: one issue slot, two FMAs.
DTCWT_D F2 fma2(const float c, const F2 v, const F2 acc) {
#ifdef DTCWT_EMU
    F2 r;
    r.x = fmaf(c, v.x, acc.x);
    r.y = fmaf(c, v.y, acc.y);
    return r;
#else
    F2 cc = make_float2(c, c), r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<const unsigned long long&>(cc)), "l"(reinterpret_cast<const unsigned long long&>(v)),
          "l"(reinterpret_cast<const unsigned long long&>(acc)));
    return r;
#endif
}

// a + b on a pair of floats (PTX add.rn.f32x2, SASS FADD2)
DTCWT_D F2 add2(const F2 a, const F2 b) {
#ifdef DTCWT_EMU
    F2 r;
    r.x = a.x + b.x;
    r.y = a.y + b.y;
    return r;
#else
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
    return r;
#endif
}

// Per-thread asynchronous copies global -> shared (cp.async, 8 bytes): a thread stages its own future inputs in a private slice
// of shared memory and waits on its own copy groups; used by the kernels whose loads are per-thread 8-byte streams.
DTCWT_D void async_copy8(void* smem_dst, const void* gmem_src) {
#ifdef DTCWT_EMU
    *reinterpret_cast<F2*>(smem_dst) = *reinterpret_cast<const F2*>(gmem_src);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
DTCWT_D void async_commit() {
#ifndef DTCWT_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
DTCWT_D void async_wait() {            // at most N of this thread's copy groups still pending
#ifndef DTCWT_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

DTCWT_HD constexpr int cmax(int a, int b) { return a > b ? a : b; }
DTCWT_HD constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------ packed-FMA helpers shared with stream2d.cuh
constexpr int kStreamThreads = 256;
constexpr int kStreamMaxTaps = 19;

struct ColTaps { float t[kStreamMaxTaps]; };           // t[k'] = h[m-1-k'] centred in K slots (taps_col)
struct PairTab { F2 p[kStreamMaxTaps + 1]; };          // p[k] = (t[k], t[k-1]), k = 0..K; t[-1] = t[K] = 0

DTCWT_D F2 zero2() { F2 z; z.x = 0.f; z.y = 0.f; return z; }
DTCWT_HD constexpr uint32_t full_mask(int K) { return (K >= 32) ? 0xffffffffu : ((1u << K) - 1u); }
DTCWT_HD constexpr bool tap_on(uint32_t mask, int k, int K) { return k >= 0 && k < K && ((mask >> k) & 1u); }
DTCWT_HD constexpr int pmod(int a, int m) { return ((a % m) + m) % m; }

// Row pass over a register window: w holds NW consecutive samples, sample j of the window sits at
// output-relative column j + W0 (output 0 of the task is column 0), filter centre C, acc[e/2] holds outputs
// (e, e+1).  out[e] += t[k] w[j] with k = j + W0 - e + C.
template <int K, uint32_t MASK, int C, int W0, int NOUT, int NW>
DTCWT_D void pair_gather4(const int j0, const F4 v, const PairTab& pt, F2 (&acc)[NOUT / 2]) {
    const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int ep = 0; ep < NOUT / 2; ++ep) {
            const int k = j0 + i + W0 - 2 * ep + C;       // tap of output 2ep; output 2ep+1 takes tap k-1
            if (k >= 0 && k <= K && (tap_on(MASK, k, K) || tap_on(MASK, k - 1, K))) acc[ep] = fma2(w[i], pt.p[k], acc[ep]);
        }
    }
}

// ------------------------------------------------------------------ filter structure
template <int K_, uint32_t MASK_ = 0xffffffffu>
struct SpecCol {                       // colfilter, K odd (shorter filters are zero-padded, centred); MASK: taps that may be non-zero
    static constexpr int P = 1, Q = 1, S = 1, K = K_;
    static constexpr uint32_t MASK = MASK_ & full_mask(K_);
    static DTCWT_HD constexpr int b(int) { return -(K_ - 1) / 2; }
    static DTCWT_HD constexpr bool on(int, int k) { return (MASK >> k) & 1u; }
};

template <int M, bool POS>
struct SpecDec {                       // coldfilt: phase 0 is Ya (delta 0) when POS else Yb (delta 1)
    static constexpr int P = 2, Q = 4, S = 2, K = M;
    static DTCWT_HD constexpr int b(int ph) { return -M + 2 + (((ph == 0) == POS) ? 0 : 1); }
    static DTCWT_HD constexpr bool on(int, int) { return true; }
};

template <int M, bool POS>
struct SpecInt {                       // colifilt (phase tables: abi_generic.inl colifilt_phase_tables)
    static constexpr int P = 4, Q = 2, S = 2, K = M / 2;
    static DTCWT_HD constexpr int off(int ph) {
        return ((M / 2) & 1) ? (POS ? ((ph & 1) ? 0 : -1) : ((ph & 1) ? -1 : 0))
                             : (POS ? ph - 2 : (ph == 0 ? -1 : ph == 1 ? -2 : ph == 2 ? 1 : 0));
    }
    static DTCWT_HD constexpr int b(int ph) { return -(M / 2) + 2 + off(ph); }
    static DTCWT_HD constexpr bool on(int, int) { return true; }
};

template <class F>
DTCWT_HD constexpr int spec_lo() {     // input samples needed before the first sample of a group
    int m = 0;
    for (int ph = 0; ph < F::P; ++ph) m = cmax(m, -F::b(ph));
    return m;
}
template <class F>
DTCWT_HD constexpr int spec_hi() {     // ... and after its last sample
    int m = 0;
    for (int ph = 0; ph < F::P; ++ph) m = cmax(m, F::b(ph) + F::S * (F::K - 1) - (F::Q - 1));
    return m;
}

// acc[P*ii+ph] += sum_k t[ph][k] * w[Q*ii + b(ph) + S*k + HALO]   (register window, all indices static)
template <class F, int NG, int HALO, int WN>
DTCWT_D void fir_gather(const float (&w)[WN], const PhaseTaps& t, float (&acc)[F::P * NG]) {
#pragma unroll
    for (int ii = 0; ii < NG; ++ii) {
#pragma unroll
        for (int ph = 0; ph < F::P; ++ph) {
            float s = acc[F::P * ii + ph];
#pragma unroll
            for (int k = 0; k < F::K; ++k)
                if (F::on(ph, k)) s = fmaf(t.t[ph][k], w[F::Q * ii + F::b(ph) + F::S * k + HALO], s);
            acc[F::P * ii + ph] = s;
        }
    }
}

// The same on a window of ROW PAIRS: w[j] holds column j of two adjacent rows, one FFMA2 with a scalar tap advances the
// output of both rows (TS: tap source, see below -- immediates for a baked table)
template <class F, int NG, int HALO, int WN, class TS>
DTCWT_D void fir_gather2(const F2 (&w)[WN], const PhaseTaps& t, F2 (&acc)[F::P * NG]) {
#pragma unroll
    for (int ii = 0; ii < NG; ++ii) {
#pragma unroll
        for (int ph = 0; ph < F::P; ++ph) {
            F2 s = acc[F::P * ii + ph];
#pragma unroll
            for (int k = 0; k < F::K; ++k)
                if (F::on(ph, k)) s = fma2(TS::get(t, ph, k), w[F::Q * ii + F::b(ph) + F::S * k + HALO], s);
            acc[F::P * ii + ph] = s;
        }
    }
}

// Tap source of a scatter: the kernel arguments, or a table baked into the instance (FFMA2 immediates).
struct RtPhase {
    static DTCWT_D float get(const PhaseTaps& t, int ph, int k) { return t.t[ph][k]; }
};
template <class B>
struct BakedPhase {
    static DTCWT_D float get(const PhaseTaps&, int, int k) { return B::get(k); }
};
template <class B>
struct BakedPhaseQ {                       // phase tables t[ph][k] (colifilt), baked
    static DTCWT_D float get(const PhaseTaps&, int ph, int k) { return B::get(ph, k); }
    static bool same(const PhaseTaps& t) {
        for (int ph = 0; ph < 4; ++ph)
            for (int k = 0; k < B::K; ++k)
                if (!(t.t[ph][k] == B::get(ph, k))) return false;
        return true;
    }
};

// Input row j (relative to the window start, HALO rows before the first group) contributes to
// the outputs whose support contains it; j is a compile-time constant after unrolling.
template <class F, int NG, int HALO, class TS = RtPhase>
DTCWT_D void fir_scatter(const int j, const F2 v, const PhaseTaps& t, F2 (&acc)[F::P * NG]) {
#pragma unroll
    for (int ii = 0; ii < NG; ++ii) {
#pragma unroll
        for (int ph = 0; ph < F::P; ++ph) {
            const int num = j - HALO - F::Q * ii - F::b(ph);
            if (num >= 0 && (num % F::S) == 0 && (num / F::S) < F::K && F::on(ph, num / F::S)) {
                acc[F::P * ii + ph] = fma2(TS::get(t, ph, num / F::S), v, acc[F::P * ii + ph]);
            }
        }
    }
}
template <class F, int NG, int HALO>
DTCWT_HD constexpr bool fir_row_used(const int j) {
    for (int ii = 0; ii < NG; ++ii)
        for (int ph = 0; ph < F::P; ++ph) {
            const int num = j - HALO - F::Q * ii - F::b(ph);
            if (num >= 0 && (num % F::S) == 0 && (num / F::S) < F::K && F::on(ph, num / F::S)) return true;
        }
    return false;
}

// =============================================================================== forward level
struct Fwd2dArgs {
    const float* x;                 // [n][rows][cols]
    float* lolo;                    // [n][out_rows][out_cols]
    float* yh;                      // complex, planar: (b, band, i, j) at 2*(b*zs_n + band*zs_band + i*zs_row + j)
    int n, rows, cols;              // stored input size
    int pr_lo, pc_lo;               // replicate padding on the low side (high side implied by Lr/Lc)
    int Lr, Lc;                     // logical (padded) size
    int out_rows, out_cols;         // P*Lr/Q, P*Lc/Q
    int use_tma;
    int prefetch;                   // one CTA per tile: L2 prefetch of the tile this many tiles ahead (0: none), see fwd2d_kernel
    int64_t zs_n, zs_band, zs_row;
    PhaseTaps h0, h1s;              // row pass: lowpass, highpass/sqrt2
    PhaseTaps v0, v1, v1s;          // column pass: lowpass, highpass, highpass/sqrt2
    PairTab ph0, ph1s;              // level 1 only: the row-pass taps as pairs (t[k], t[k-1]) for the packed row pass
};

// MODE: what leaves the column pass
//   kFwdQ2c   (2-D transform)  LoLo + the six complex sub-bands (q2c in registers)
//   kFwdRaw   (3-D transform, x/y passes of one slice; transform3d.py:256-273, 353-369)  the four REAL images
//             s0 = V:h0 H:h0, s1 = V:h1 H:h0, s2 = V:h0 H:h1, s3 = V:h1 H:h1, image s at lolo + s * zs_band (floats);
//             the host passes unscaled taps (the packers' scale belongs to cube2c, fused3d.cuh)
//   kFwdLow   (3-D level 1 without highpasses, transform3d.py:291-315, 442-456)  V:h0 H:h0 only
//   kFwdSym   kFwdQ2c for a level-1 pair whose two filters are SYMMETRIC (t[k] = t[K-1-k], all shipped biorthogonal
//             families): the column pass gathers instead of scattering and forms the sums x[c-k] + x[c+k] once for BOTH
//             filters, 8 FADD2 + 15 FFMA2 per output row pair instead of 28 FFMA2 for near_sym_b (the host checks the
//             symmetry of the taps it is given bit for bit)
//   kFwdHH    bands 1 and 4 only, from ONE filter in both directions: the band-pass filter h2 of the `_bp` families
//             (transform2d.py:116-127, 145-157), passed in the H1 slot.  A `_bp` level is the ordinary launch followed by
//             this one, which overwrites the two diagonal sub-bands.
//   kFwdSymP  kFwdSym as a persistent kernel (one resident wave of CTAs, next tile prefetched during the column pass): experiment
constexpr int kFwdQ2c = 0, kFwdRaw = 1, kFwdLow = 2, kFwdSym = 3, kFwdHH = 4, kFwdSymP = 5;

template <class H0, class H1, int GH_, int GW_, int NGV_, class TV0 = RtPhase, class TV1S = RtPhase, class TV1 = RtPhase,
          int MODE_ = kFwdQ2c, bool SPLIT_ = true>
struct Fwd2d {
    typedef Fwd2dArgs Args;
    static constexpr int MODE = MODE_;
    static constexpr int P = H0::P, Q = H0::Q;
    static constexpr int GH = GH_, GW = GW_, NGV = NGV_;
    static constexpr int NGH = 4 / Q;                       // a row task covers 4 input columns
    static constexpr int HL = cmax(spec_lo<H0>(), spec_lo<H1>());
    static constexpr int HR = cmax(spec_hi<H0>(), spec_hi<H1>());
    static constexpr int HLA = round_up(HL, 4), HRA = round_up(HR, 4);
    static constexpr int RX = Q * GH + HL + HR;             // input tile rows
    // input tile pitch (multiple of 4) and pitch of the row-pass outputs; the q-shift row pass reads 32-byte segments
    // with adjacent lanes on adjacent rows, which is free of bank conflicts when the pitch is 4 (mod 8) floats
    static constexpr int CX = Q * GW + HLA + HRA + ((P == 2 && (Q * GW + HLA + HRA) % 8 == 0) ? 4 : 0);
    static constexpr int CA = P * GW + ((P == 2) ? 4 : 0);
    static constexpr int WN = 4 + HLA + HRA;                // register window of a row task
    static constexpr int NSEG = GW / NGH;
    static constexpr int NR = Q * NGV + HL + HR;            // input rows of a column task
    static constexpr int NOUT = P * NGV;                    // output rows of a column task (even)
    static constexpr int kSmemFloats = RX * CX + (MODE == kFwdLow ? 1 : 2) * RX * CA;
    static constexpr int kThreads = kFusedThreads;
    static constexpr int kPhases = 5;
    // q-shift levels: one resident wave of CTAs walks over the tiles and prefetches the next tile during the column
    // pass (measured 4 % faster); level 1 is faster with one CTA per tile (profiles/r1_03)
    static constexpr bool kPersistent = (P != 1) || (MODE_ == kFwdSymP);
    static constexpr int kMinBlocks = (kSmemFloats * 4 * 3 <= 220 * 1024) ? 3 : 2;      // CTAs per SM the shared memory allows
    static_assert(P == H1::P && Q == H1::Q, "filter pair must share its rate");
    static_assert((NOUT % 2) == 0 && (GH % NGV) == 0 && (GW % NGH) == 0 && (CA % 2) == 0, "tile shape");
    static_assert(P * NGH == 4 || P * NGH == 2, "row task writes a float4 or float2");

    static DTCWT_HD int tiles_r(const Args& a) { return (a.Lr + Q * GH - 1) / (Q * GH); }
    static DTCWT_HD int tiles_c(const Args& a) { return (a.Lc + Q * GW - 1) / (Q * GW); }
    // logical coordinates of smem element (0, 0)
    static DTCWT_HD int row0(int by) { return Q * GH * by - HL; }
    static DTCWT_HD int col0(int bx) { return Q * GW * bx - HLA; }

    // phase 0: plain loads of the tile, zero outside the stored array (what TMA produces; used when
    // the row pitch is not a multiple of 16 bytes, and by the host emulator)
    static DTCWT_D void phase_load(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (a.use_tma) return;
        const float* img = a.x + (int64_t)bz * a.rows * a.cols;
        const int r0 = row0(by) - a.pr_lo, c0 = col0(bx) - a.pc_lo;
        for (int e = tid; e < RX * CX; e += kThreads) {
            const int lr = e / CX, lc = e - lr * CX;
            const int r = r0 + lr, c = c0 + lc;
            sm[e] = (r >= 0 && r < a.rows && c >= 0 && c < a.cols) ? img[(int64_t)r * a.cols + c] : 0.f;
        }
    }

    // smem index that holds the sample logical index L (outside the stored range) mirrors, or -1
    static DTCWT_HD int mirror_src(int L, int Ltot, int pad_lo, int len, int L0, int extent) {
        const int s = L - pad_lo;
        if (s >= 0 && s < len) return -2;                                  // directly loaded
        // one fold covers every halo that is shorter than the image; reflect_any's modulo only for tinier images
        int r = L < 0 ? -1 - L : (L >= Ltot ? 2 * Ltot - 1 - L : L);
        if ((unsigned)r >= (unsigned)Ltot) r = reflect_any(L, Ltot);
        const int l = unpad(r, pad_lo, len) + pad_lo - L0;
        return (l >= 0 && l < extent) ? l : -1;
    }
    static DTCWT_HD bool touches_edge(int L0, int extent, int pad_lo, int len) {
        return (L0 - pad_lo < 0) || (L0 + extent - pad_lo > len);
    }

    // the tile reaches outside the stored image (uniform over the CTA): only then do the patch phases have work
    static DTCWT_HD bool tile_on_edge(const Args& a, int bx, int by) {
        return touches_edge(row0(by), RX, a.pr_lo, a.rows) || touches_edge(col0(bx), CX, a.pc_lo, a.cols);
    }

    // phase 1 / 2: symmetric extension (utils.py:136-153) of tiles on the image border, inside smem
    // Only the rows / columns of the tile that lie OUTSIDE the stored array are visited (the first n_lo and the ones from
    // hi0 on): on the small slices of a 3-D volume every tile touches a border, and a sweep over the whole tile cost
    // more than the filtering itself.  Sources are stored samples, which no patch writes, so there is no ordering hazard.
    static DTCWT_HD void outside_range(int L0, int extent, int pad_lo, int len, int& n_lo, int& hi0) {
        n_lo = pad_lo - L0;
        n_lo = n_lo < 0 ? 0 : (n_lo > extent ? extent : n_lo);
        hi0 = pad_lo + len - L0;
        hi0 = hi0 < n_lo ? n_lo : (hi0 > extent ? extent : hi0);
    }
    static DTCWT_D void phase_patch_rows(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const int L0 = row0(by);
        if (!touches_edge(L0, RX, a.pr_lo, a.rows)) return;
        int n_lo, hi0;
        outside_range(L0, RX, a.pr_lo, a.rows, n_lo, hi0);
        const int n_out = n_lo + (RX - hi0);
        for (int e = tid; e < n_out * CX; e += kThreads) {
            const int k = e / CX, lc = e - k * CX;
            const int lr = k < n_lo ? k : hi0 + (k - n_lo);
            const int src = mirror_src(L0 + lr, a.Lr, a.pr_lo, a.rows, L0, RX);
            if (src >= 0) sm[lr * CX + lc] = sm[src * CX + lc];
        }
    }
    static DTCWT_D void phase_patch_cols(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const int L0 = col0(bx);
        if (!touches_edge(L0, CX, a.pc_lo, a.cols)) return;
        int n_lo, hi0;
        outside_range(L0, CX, a.pc_lo, a.cols, n_lo, hi0);
        const int n_out = n_lo + (CX - hi0);
        if (n_out == 0) return;
        if (n_out <= 32) {
            // the usual case (the halo of one or both sides): rows x outside columns with the columns padded to a power of
            // two, so that no division by a run-time count is needed
            const int sh = n_out <= 8 ? 3 : (n_out <= 16 ? 4 : 5);
            for (int e = tid; e < (RX << sh); e += kThreads) {
                const int lr = e >> sh, k = e & ((1 << sh) - 1);
                if (k >= n_out) continue;
                const int lc = k < n_lo ? k : hi0 + (k - n_lo);
                const int src = mirror_src(L0 + lc, a.Lc, a.pc_lo, a.cols, L0, CX);
                if (src >= 0) sm[lr * CX + lc] = sm[lr * CX + src];
            }
            return;
        }
        for (int e = tid; e < RX * n_out; e += kThreads) {         // images narrower than a tile
            const int lr = e / n_out, k = e - lr * n_out;
            const int lc = k < n_lo ? k : hi0 + (k - n_lo);
            const int src = mirror_src(L0 + lc, a.Lc, a.pc_lo, a.cols, L0, CX);
            if (src >= 0) sm[lr * CX + lc] = sm[lr * CX + src];
        }
    }

    // level 1 (P = Q = 1): scalar sample x tap pair (t[k], t[k-1]) accumulates the output pair (c, c+1) in one FFMA2
    template <class HH = H0>
    static DTCWT_D typename std::enable_if<HH::P == 1 && HH::Q == 1>::type phase_rows_packed(const Args& a, float* sm, int tid) {
        const float* Xs = sm;
        float* As = sm + RX * CX;
        float* Bs = As + RX * CA;
        constexpr int C0 = (H0::K - 1) / 2, C1 = (H1::K - 1) / 2;
        for (int task = tid; task < RX * NSEG; task += kThreads) {
            const int lr = task / NSEG, seg = task - lr * NSEG;
            F2 oa[2], ob[2];
            oa[0] = zero2(); oa[1] = zero2(); ob[0] = zero2(); ob[1] = zero2();
            const F4* src = reinterpret_cast<const F4*>(Xs + lr * CX + seg * 4);
#pragma unroll
            for (int c = 0; c < WN / 4; ++c) {
                const F4 v = src[c];
                if (MODE != kFwdHH) pair_gather4<H0::K, H0::MASK, C0, -HLA, 4, WN>(4 * c, v, a.ph0, oa);
                if (MODE != kFwdLow) pair_gather4<H1::K, H1::MASK, C1, -HLA, 4, WN>(4 * c, v, a.ph1s, ob);
            }
            F4 va, vb;
            va.x = oa[0].x; va.y = oa[0].y; va.z = oa[1].x; va.w = oa[1].y;
            vb.x = ob[0].x; vb.y = ob[0].y; vb.z = ob[1].x; vb.w = ob[1].y;
            if (MODE != kFwdHH) *reinterpret_cast<F4*>(As + lr * CA + seg * 4) = va;
            if (MODE != kFwdLow) *reinterpret_cast<F4*>(Bs + lr * CA + seg * 4) = vb;
        }
    }
    template <class HH = H0>
    static DTCWT_D typename std::enable_if<!(HH::P == 1 && HH::Q == 1)>::type phase_rows_packed(const Args&, float*, int) {}

    // q-shift levels (P = 2, Q = 4): lowpass phase 0 and highpass phase 1 read the same samples (phases 1 and 0 the
    // neighbouring ones), so one FFMA2 of a scalar sample with the tap pair (h0[ph][k], h1[1-ph][k]) advances both
    // filters (pairs prepared by the host in ph0 / ph1s).  One task = one tile row x NGD groups (4 * NGD input columns).
    static constexpr int NGD = 2;
    template <class HH = H0>
    static DTCWT_D typename std::enable_if<HH::P == 2 && HH::Q == 4>::type phase_rows_dec(const Args& a, float* sm, int tid) {
        constexpr int M = H0::K;
        constexpr int OFF = HLA - (M - 2);                                  // window index of the first even-phase sample
        constexpr int WND = round_up(4 * (NGD - 1) + OFF + 2 * (M - 1) + 2, 4);
        constexpr int NSD = GW / NGD;
        static_assert(H0::b(0) == -M + 2 && H1::b(1) == -M + 2 && H0::b(1) == -M + 3 && H1::b(0) == -M + 3, "phase pairing");
        static_assert((GW % NGD) == 0 && (RX % 2) == 0 && 4 * NGD * (NSD - 1) + WND <= CX && NGD == 2, "q-shift row task");
        const float* Xs = sm;
        float* As = sm + RX * CX;
        float* Bs = As + RX * CA;
        for (int task = tid; task < RX * NSD; task += kThreads) {
            const int half = task >> 1;
            const int lr = 2 * (half / NSD) + (task & 1), seg = half % NSD;
            F2 m1[NGD], m2[NGD];                                            // (A[2g], B[2g+1]) and (A[2g+1], B[2g])
#pragma unroll
            for (int g = 0; g < NGD; ++g) { m1[g] = zero2(); m2[g] = zero2(); }
            const F4* src = reinterpret_cast<const F4*>(Xs + lr * CX + 4 * NGD * seg);
#pragma unroll
            for (int c = 0; c < WND / 4; ++c) {
                const F4 v = src[c];
                const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int g = 0; g < NGD; ++g) {
                        const int num = 4 * c + i - 4 * g - OFF;
                        if (num >= 0 && (num & 1) == 0 && num / 2 < M) m1[g] = fma2(w[i], a.ph0.p[num / 2], m1[g]);
                        if (num >= 1 && (num & 1) == 1 && (num - 1) / 2 < M) m2[g] = fma2(w[i], a.ph1s.p[(num - 1) / 2], m2[g]);
                    }
                }
            }
            F4 va, vb;
            va.x = m1[0].x; va.y = m2[0].x; va.z = m1[1].x; va.w = m2[1].x;
            vb.x = m2[0].y; vb.y = m1[0].y; vb.z = m2[1].y; vb.w = m1[1].y;
            *reinterpret_cast<F4*>(As + lr * CA + 2 * NGD * seg) = va;
            *reinterpret_cast<F4*>(Bs + lr * CA + 2 * NGD * seg) = vb;
        }
    }
    template <class HH = H0>
    static DTCWT_D typename std::enable_if<!(HH::P == 2 && HH::Q == 4)>::type phase_rows_dec(const Args&, float*, int) {}

    // phase 3: row pass, one task = one tile row x 4 input columns
    static DTCWT_D void phase_rows(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (P == 1 && Q == 1) {
            phase_rows_packed(a, sm, tid);
            return;
        }
        if (P == 2 && Q == 4) {
            phase_rows_dec(a, sm, tid);
            return;
        }
        const float* Xs = sm;
        float* As = sm + RX * CX;
        float* Bs = As + RX * CA;
        for (int task = tid; task < RX * NSEG; task += kThreads) {
            const int lr = task / NSEG, seg = task - lr * NSEG;
            float w[WN];
            const F4* src = reinterpret_cast<const F4*>(Xs + lr * CX + seg * 4);
#pragma unroll
            for (int c = 0; c < WN / 4; ++c) {
                const F4 v = src[c];
                w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
            }
            float oa[P * NGH], ob[P * NGH];
#pragma unroll
            for (int i = 0; i < P * NGH; ++i) { oa[i] = 0.f; ob[i] = 0.f; }
            fir_gather<H0, NGH, HLA, WN>(w, a.h0, oa);
            fir_gather<H1, NGH, HLA, WN>(w, a.h1s, ob);
            float* da = As + lr * CA + seg * (P * NGH);
            float* db = Bs + lr * CA + seg * (P * NGH);
            if (P * NGH == 4) {
                F4 va, vb;
                va.x = oa[0]; va.y = oa[1]; va.z = oa[2 % (P * NGH)]; va.w = oa[3 % (P * NGH)];
                vb.x = ob[0]; vb.y = ob[1]; vb.z = ob[2 % (P * NGH)]; vb.w = ob[3 % (P * NGH)];
                *reinterpret_cast<F4*>(da) = va;
                *reinterpret_cast<F4*>(db) = vb;
            } else {
                F2 va, vb;
                va.x = oa[0]; va.y = oa[1];
                vb.x = ob[0]; vb.y = ob[1];
                *reinterpret_cast<F2*>(da) = va;
                *reinterpret_cast<F2*>(db) = vb;
            }
        }
    }

    // q2c of NOUT rows x 2 columns (transform2d.py:301-322; the 1/sqrt2 is already in the taps): z0 / z1 point at the
    // first quad row of the two sub-bands, rs = floats per quad row, nq = quad rows that lie inside the image
    static DTCWT_D void store_q2c(const F2 (&y)[NOUT], float* z0, float* z1, int64_t rs, int nq) {
#pragma unroll
        for (int q = 0; q < NOUT / 2; ++q) {
            if (q < nq) {
                const float A = y[2 * q].x, B = y[2 * q].y, C = y[2 * q + 1].x, D = y[2 * q + 1].y;
                F2 w0, w1;
                w0.x = A - D; w0.y = B + C;
                w1.x = A + D; w1.y = B - C;
                *reinterpret_cast<F2*>(z0 + q * rs) = w0;
                *reinterpret_cast<F2*>(z1 + q * rs) = w1;
            }
        }
    }

    // NOUT rows x 2 columns of a real image (kFwdRaw / kFwdLow)
    static DTCWT_D void store_rows(const F2 (&y)[NOUT], float* dst, int64_t rs, int nrow) {
#pragma unroll
        for (int i = 0; i < NOUT; ++i)
            if (i < nrow) *reinterpret_cast<F2*>(dst + (int64_t)i * rs) = y[i];
    }

    // Column tasks of a tile: NCP column pairs x GH / NGV strips.  The q-shift tiles have only half as many as the CTA has
    // threads; there the A half (LoLo + bands 0, 5 / images s0, s1) and the B half (bands 2, 3 and 1, 4 / images s2, s3)
    // of a task go to different warps, so every warp has work in the column pass (kSplitAB).
    static constexpr int NTASK = (P * GW / 2) * (GH / NGV);
    static constexpr bool kSplitAB = SPLIT_ && (2 * NTASK <= kThreads) && MODE_ != kFwdLow && MODE_ != kFwdHH;

    // column pass of the 3-D modes: real images out
    template <bool DO_A, bool DO_B>
    static DTCWT_D void cols_real_task(const Args& a, float* sm, int bx, int by, int bz, int task) {
        const float* As = sm + RX * CX;
        const float* Bs = As + RX * CA;
        constexpr int NCP = P * GW / 2;
        const int strip = task / NCP, cp = task - strip * NCP;
        const int lrow = Q * NGV * strip;
        const int orow = P * (GH * by + NGV * strip);
        const int ocol = P * GW * bx + 2 * cp;
        const int nrow = (ocol < a.out_cols) ? a.out_rows - orow : 0;
        float* dst = a.lolo + ((int64_t)bz * a.out_rows + orow) * a.out_cols + ocol;
        F2 lo[NOUT], hi[NOUT];
        if (DO_A) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) { lo[i].x = lo[i].y = 0.f; hi[i].x = hi[i].y = 0.f; }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const F2 v = *reinterpret_cast<const F2*>(As + (lrow + j) * CA + 2 * cp);
                fir_scatter<H0, NGV, HL, TV0>(j, v, a.v0, lo);
                if (MODE == kFwdRaw) fir_scatter<H1, NGV, HL, TV1>(j, v, a.v1, hi);
            }
            store_rows(lo, dst, a.out_cols, nrow);
            if (MODE == kFwdRaw) store_rows(hi, dst + a.zs_band, a.out_cols, nrow);
        }
        if (DO_B && MODE == kFwdRaw) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) { lo[i].x = lo[i].y = 0.f; hi[i].x = hi[i].y = 0.f; }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const F2 v = *reinterpret_cast<const F2*>(Bs + (lrow + j) * CA + 2 * cp);
                fir_scatter<H0, NGV, HL, TV0>(j, v, a.v0, lo);
                fir_scatter<H1, NGV, HL, TV1>(j, v, a.v1, hi);
            }
            store_rows(lo, dst + 2 * a.zs_band, a.out_cols, nrow);
            store_rows(hi, dst + 3 * a.zs_band, a.out_cols, nrow);
        }
    }
    static DTCWT_D void phase_cols_real(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (kSplitAB) {
            const int half = tid / (kThreads / 2);                                           // uniform within a warp
            for (int task = tid - half * (kThreads / 2); task < NTASK; task += kThreads / 2) {
                if (half == 0) cols_real_task<true, false>(a, sm, bx, by, bz, task);
                else cols_real_task<false, true>(a, sm, bx, by, bz, task);
            }
            return;
        }
        for (int task = tid; task < NTASK; task += kThreads) cols_real_task<true, true>(a, sm, bx, by, bz, task);
    }

    // kFwdSym: one output row pair of BOTH symmetric filters from the register window w (centre row c): the sums
    // w[c-k] + w[c+k] are shared; taps known to be zero (MASK) are compiled out, sums nobody needs are never formed
    template <class TLO, class THI>
    static DTCWT_D void sym_row(const F2 (&w)[NR], const int c, const PhaseTaps& tlo, const PhaseTaps& thi, F2& lo, F2& hi) {
        constexpr int C0 = (H0::K - 1) / 2, C1 = (H1::K - 1) / 2;
        lo = fma2(TLO::get(tlo, 0, C0), w[c], zero2());
        hi = fma2(THI::get(thi, 0, C1), w[c], zero2());
#pragma unroll
        for (int k = 1; k <= cmax(C0, C1); ++k) {
            const bool ul = k <= C0 && H0::on(0, C0 + k), uh = k <= C1 && H1::on(0, C1 + k);
            if (ul || uh) {
                const F2 s = add2(w[c - k], w[c + k]);
                if (ul) lo = fma2(TLO::get(tlo, 0, C0 + k), s, lo);
                if (uh) hi = fma2(THI::get(thi, 0, C1 + k), s, hi);
            }
        }
    }

    template <class HH = H0>
    static DTCWT_D typename std::enable_if<HH::P == 1 && HH::Q == 1>::type phase_cols_sym(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const float* As = sm + RX * CX;
        const float* Bs = As + RX * CA;
        constexpr int NCP = GW / 2;
        static_assert(HL == HR && NR == NGV + 2 * HL && (NGV % 2) == 0, "centred window");
        for (int task = tid; task < NCP * (GH / NGV); task += kThreads) {
            const int strip = task / NCP, cp = task - strip * NCP;
            const int lrow = NGV * strip;
            const int orow = GH * by + NGV * strip;
            const int ocol = GW * bx + 2 * cp;
            const int nrow = (ocol < a.out_cols) ? a.out_rows - orow : 0;
            const int nq = nrow / 2;
            float* zb = a.yh + 2 * ((int64_t)bz * a.zs_n + (int64_t)(orow / 2) * a.zs_row + ocol / 2);
            const int64_t bs = 2 * a.zs_band, rs = 2 * a.zs_row;
            float* dst = a.lolo + ((int64_t)bz * a.out_rows + orow) * a.out_cols + ocol;
            F2 w[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) w[j] = *reinterpret_cast<const F2*>(As + (lrow + j) * CA + 2 * cp);
#pragma unroll
            for (int q = 0; q < NGV / 2; ++q) {
                F2 l0, h0, l1, h1;
                sym_row<TV0, TV1S>(w, 2 * q + HL, a.v0, a.v1s, l0, h0);
                sym_row<TV0, TV1S>(w, 2 * q + 1 + HL, a.v0, a.v1s, l1, h1);
                if (2 * q < nrow) *reinterpret_cast<F2*>(dst + (int64_t)(2 * q) * a.out_cols) = l0;
                if (2 * q + 1 < nrow) *reinterpret_cast<F2*>(dst + (int64_t)(2 * q + 1) * a.out_cols) = l1;
                if (q < nq) store_quad(h0, h1, zb + q * rs, zb + 5 * bs + q * rs);               // bands 0, 5
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) w[j] = *reinterpret_cast<const F2*>(Bs + (lrow + j) * CA + 2 * cp);
#pragma unroll
            for (int q = 0; q < NGV / 2; ++q) {
                F2 l0, h0, l1, h1;
                sym_row<TV0, TV1>(w, 2 * q + HL, a.v0, a.v1, l0, h0);
                sym_row<TV0, TV1>(w, 2 * q + 1 + HL, a.v0, a.v1, l1, h1);
                if (q < nq) {
                    store_quad(l0, l1, zb + 2 * bs + q * rs, zb + 3 * bs + q * rs);              // bands 2, 3
                    store_quad(h0, h1, zb + 1 * bs + q * rs, zb + 4 * bs + q * rs);              // bands 1, 4
                }
            }
        }
    }
    template <class HH = H0>
    static DTCWT_D typename std::enable_if<!(HH::P == 1 && HH::Q == 1)>::type phase_cols_sym(const Args&, float*, int, int, int, int) {}

    // q2c of one quad (rows e0 / e1, two columns) -> sub-bands z0, z1 (the 1/sqrt2 is in the taps)
    static DTCWT_D void store_quad(const F2 e0, const F2 e1, float* z0, float* z1) {
        F2 w0, w1;
        w0.x = e0.x - e1.y; w0.y = e0.y + e1.x;
        w1.x = e0.x + e1.y; w1.y = e0.y - e1.x;
        *reinterpret_cast<F2*>(z0) = w0;
        *reinterpret_cast<F2*>(z1) = w1;
    }

    // kFwdHH: V:h2 of B = H:h2(X)/sqrt2 -> q2c -> bands 1, 4
    static DTCWT_D void phase_cols_hh(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const float* Bs = sm + RX * CX + RX * CA;
        constexpr int NCP = P * GW / 2;
        for (int task = tid; task < NCP * (GH / NGV); task += kThreads) {
            const int strip = task / NCP, cp = task - strip * NCP;
            const int lrow = Q * NGV * strip;
            const int orow = P * (GH * by + NGV * strip);
            const int ocol = P * GW * bx + 2 * cp;
            const int nrow = (ocol < a.out_cols) ? a.out_rows - orow : 0;
            float* zb = a.yh + 2 * ((int64_t)bz * a.zs_n + (int64_t)(orow / 2) * a.zs_row + ocol / 2);
            const int64_t bs = 2 * a.zs_band, rs = 2 * a.zs_row;
            F2 hi[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; ++i) hi[i] = zero2();
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const F2 v = *reinterpret_cast<const F2*>(Bs + (lrow + j) * CA + 2 * cp);
                fir_scatter<H1, NGV, HL, TV1>(j, v, a.v1, hi);
            }
            store_q2c(hi, zb + 1 * bs, zb + 4 * bs, rs, nrow / 2);
        }
    }

    // phase 4: column pass, one task = 2 adjacent columns x NGV groups of rows; results leave from registers
    static DTCWT_D void phase_cols(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (MODE == kFwdHH) {
            phase_cols_hh(a, sm, bx, by, bz, tid);
            return;
        }
        if (MODE == kFwdSym || MODE == kFwdSymP) {
            phase_cols_sym(a, sm, bx, by, bz, tid);
            return;
        }
        if (MODE != kFwdQ2c) {
            phase_cols_real(a, sm, bx, by, bz, tid);
            return;
        }
        if (kSplitAB) {
            const int half = tid / (kThreads / 2);                                           // uniform within a warp
            for (int task = tid - half * (kThreads / 2); task < NTASK; task += kThreads / 2) {
                if (half == 0) cols_q2c_task<true, false>(a, sm, bx, by, bz, task);
                else cols_q2c_task<false, true>(a, sm, bx, by, bz, task);
            }
            return;
        }
        for (int task = tid; task < NTASK; task += kThreads) cols_q2c_task<true, true>(a, sm, bx, by, bz, task);
    }

    // one column task of the 2-D mode: DO_A = LoLo + bands 0, 5 from A, DO_B = bands 2, 3 and 1, 4 from B
    template <bool DO_A, bool DO_B>
    static DTCWT_D void cols_q2c_task(const Args& a, float* sm, int bx, int by, int bz, int task) {
        const float* As = sm + RX * CX;
        const float* Bs = As + RX * CA;
        constexpr int NCP = P * GW / 2;                           // column pairs of the tile (CA may be padded)
        const int strip = task / NCP, cp = task - strip * NCP;
        const int lrow = Q * NGV * strip;
        const int orow = P * (GH * by + NGV * strip);        // first output row (LoLo coordinates)
        const int ocol = P * GW * bx + 2 * cp;
        // output rows / quad rows of this task that lie inside the image (none when its columns lie outside)
        const int nrow = (ocol < a.out_cols) ? a.out_rows - orow : 0;
        const int nq = nrow / 2;                              // out_rows is even
        float* zb = a.yh + 2 * ((int64_t)bz * a.zs_n + (int64_t)(orow / 2) * a.zs_row + ocol / 2);
        const int64_t bs = 2 * a.zs_band, rs = 2 * a.zs_row;
        F2 lo[NOUT], hi[NOUT];
        if (DO_A) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) { lo[i].x = lo[i].y = 0.f; hi[i].x = hi[i].y = 0.f; }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const F2 v = *reinterpret_cast<const F2*>(As + (lrow + j) * CA + 2 * cp);
                fir_scatter<H0, NGV, HL, TV0>(j, v, a.v0, lo);
                fir_scatter<H1, NGV, HL, TV1S>(j, v, a.v1s, hi);
            }
            float* dst = a.lolo + ((int64_t)bz * a.out_rows + orow) * a.out_cols + ocol;
#pragma unroll
            for (int i = 0; i < NOUT; ++i)
                if (i < nrow) *reinterpret_cast<F2*>(dst + (int64_t)i * a.out_cols) = lo[i];
            store_q2c(hi, zb, zb + 5 * bs, rs, nq);              // vertical high x horizontal low -> bands 0, 5
        }
        if (DO_B) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) { lo[i].x = lo[i].y = 0.f; hi[i].x = hi[i].y = 0.f; }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const F2 v = *reinterpret_cast<const F2*>(Bs + (lrow + j) * CA + 2 * cp);
                fir_scatter<H0, NGV, HL, TV0>(j, v, a.v0, lo);
                fir_scatter<H1, NGV, HL, TV1>(j, v, a.v1, hi);
            }
            store_q2c(lo, zb + 2 * bs, zb + 3 * bs, rs, nq);     // vertical low x horizontal high -> bands 2, 3
            store_q2c(hi, zb + 1 * bs, zb + 4 * bs, rs, nq);     // high x high -> bands 1, 4
        }
    }

    template <int PH>
    static DTCWT_D void phase(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (PH == 0) phase_load(a, sm, bx, by, bz, tid);
        if (PH == 1) phase_patch_rows(a, sm, bx, by, bz, tid);
        if (PH == 2) phase_patch_cols(a, sm, bx, by, bz, tid);
        if (PH == 3) phase_rows(a, sm, bx, by, bz, tid);
        if (PH == 4) phase_cols(a, sm, bx, by, bz, tid);
    }
};

// =============================================================================== inverse level
struct Inv2dArgs {
    const float* z;                 // lowpass [n][rows][cols]
    const float* yh;                // complex planar sub-bands [n][6][rows/2][cols/2] (strides below)
    float* out;                     // [n][out_rows][out_cols]
    int n, rows, cols;
    int crop_r, crop_c;             // 1: drop the first and last output row / column (transform2d.py:263-268)
    int out_rows, out_cols;         // P*rows/Q - 2*crop_r, ...
    int out_vec4;                   // rows of `out` are 16-byte aligned and not cropped: float4 stores
    int64_t zs_n, zs_band, zs_row;
    float gain[6];                  // gain_mask column of this level, times 1/sqrt2
    PhaseTaps g0, g1;
};

// Column pass first: each thread owns one quad column (two adjacent real columns), NGV groups of rows
// and ONE of the two intermediate images (its warp's role):
//     role 0: y1 = V:g0(Z)  + V:g1(lh)        role 1: y2 = V:g0(hl) + V:g1(hh)   (transform2d.py:248-256, 279-285)
// It walks down the quad rows its outputs depend on, loads the lowpass quad / the complex coefficients
// straight from global memory (coalesced 8-byte loads, next quad row prefetched), applies c2q in
// registers and scatters the real rows into its accumulators, which go to shared memory; the row pass
// out = H:g0(y1) + H:g1(y2) reads them back with a register window and stores float4s.
// No input staging, one block barrier.  Tiles whose halo lies inside the image take a variant compiled
// without the symmetric-extension logic.
// RAW (3-D transform, y/x passes of one slice, transform3d.py:485-490): the inputs are the four REAL images
// s0..s3 of Fwd2d's kFwdRaw mode, image s at z + s * zs_band (floats), instead of lowpass + complex sub-bands.
// HH (`_bp` families, transform2d.py:254-262): the second launch of a level -- the lowpass counts as zero and the result is
// ADDED to `out`; with the gains of the other four sub-bands zero and g2 in the G1 slot that is H:g2(V:g2(c2q(bands 1, 4))).
// ROWPAIR: y1 / y2 are kept in shared memory as interleaved ROW PAIRS ((row 2i, row 2i+1) of a column adjacent), so the row
// pass runs on pairs of rows with packed FFMA2 and scalar taps -- immediates when the taps are baked (TS0 / TS1), which
// also takes the constant loads (LDCU) out of the column pass.  The kernel is issue-bound (profiles/r2_04: 74 % issue
// active, a third of its instructions unpacked FFMA of the row pass), so halving the row pass's FMA instructions pays.
// ASYNC > 0: the column pass stages its next ASYNC quad rows in a thread-private slice of shared memory with cp.async instead
// of holding one prefetched quad row in registers (async_copy8 above): more loads in flight, no barrier.
template <class G0, class G1, int NGV_, int NSTRIP_, int NWIDE_, bool RAW_ = false, bool HH_ = false, class TS0 = RtPhase,
          class TS1 = RtPhase, bool ROWPAIR_ = false, int ASYNC_ = 0>
struct Inv2d {
    typedef Inv2dArgs Args;
    static constexpr bool RAW = RAW_, HH = HH_, ROWPAIR = ROWPAIR_;
    static constexpr int ASYNC = ASYNC_;
    static_assert(!(ROWPAIR_ && HH_), "the accumulate-into-out variant keeps the single-row row pass");
    static constexpr int P = G0::P, Q = G0::Q;
    static constexpr int NGV = NGV_, NSTRIP = NSTRIP_, NWIDE = NWIDE_;
    static constexpr int NGH = 4 / Q;
    static constexpr int HL = cmax(spec_lo<G0>(), spec_lo<G1>());
    static constexpr int HR = cmax(spec_hi<G0>(), spec_hi<G1>());
    static constexpr int HLR = round_up(HL, 2), HRR = round_up(HR, 2);   // rows: whole quads
    static constexpr int HLC = round_up(HL, 2), HRC = round_up(HR, 2);   // columns: whole quads
    static constexpr int kThreads = kFusedThreads;
    static constexpr int QCOLS = 32 * NWIDE;                             // quad columns of a tile (whole warps)
    static constexpr int CY = 2 * QCOLS;                                 // columns of y1 / y2 in smem
    static constexpr int TWI = (CY - HLC - HRC) / 4 * 4;                 // input columns a tile produces (multiple of 4)
    static constexpr int GH = NGV * NSTRIP;                              // row groups of a tile
    static constexpr int RY = P * GH;                                    // rows of y1 / y2
    static constexpr int NQR = (Q * NGV + HLR + HRR) / 2;                // quad rows a column task reads
    static constexpr int NOUT = P * NGV;
    static constexpr int WN = round_up(HLC + 4 + HR, 4);                 // register window of a row task
    static constexpr int NSEG = TWI / 4;
    // ROWPAIR: pitch of a pair row (2 CY floats) padded by 4, so that adjacent pair rows sit one 16-byte bank group apart
    static constexpr int PYP = 2 * CY + 4;
    static constexpr int kImgFloats = ROWPAIR_ ? (RY / 2) * PYP : RY * CY;      // one of y1 / y2
    static constexpr int kSmemFloats = 2 * kImgFloats + ASYNC_ * 4 * 2 * kFusedThreads;       // y1 / y2 + the copy stages
    static constexpr int kPhases = 2;
    static constexpr int kMinBlocks = 3;                                 // register budget: 3 CTAs (24 warps) per SM
    static_assert(P == G1::P && Q == G1::Q, "filter pair must share its rate");
    static_assert(2 * NSTRIP * NWIDE * 32 == kThreads && ((Q * NGV) % 2) == 0 && (P * NGH) % 4 == 0, "tile shape");
    static_assert(WN + 4 * (NSEG - 1) <= CY, "row-pass window must stay inside the smem row");

    static DTCWT_HD int tiles_r(const Args& a) { return (a.rows + Q * GH - 1) / (Q * GH); }
    static DTCWT_HD int tiles_c(const Args& a) { return (a.cols + TWI - 1) / TWI; }

    struct Raw { F2 v[4]; };       // role 0: Z top row, Z bottom row, band 0, band 5;  role 1: bands 2, 3, 1, 4

    static DTCWT_HD int reflect_quad(int g, int n, bool& flip) {
        flip = false;
        if (g < 0) { g = -1 - g; flip = true; } else if (g >= n) { g = 2 * n - 1 - g; flip = true; }
        return g;
    }

    // EDGE: the tile touches the top or bottom of the image (row reflection).  ptr[i] + qrow * stride[i] is this thread's
    // (already column-reflected) quad column in quad row `qrow` of its four inputs: one IMAD.WIDE per load.
    template <bool EDGE>
    static DTCWT_D void load_quad_row(const Args& a, const char* const (&ptr)[4], const int (&stride)[4], int qrow, Raw& r) {
        int gi = qrow;
        if (EDGE) {
            bool fr;
            gi = reflect_quad(qrow, a.rows / 2, fr);
            if (!(gi >= 0 && gi < a.rows / 2)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) r.v[i].x = r.v[i].y = 0.f;
                return;
            }
        }
#pragma unroll
        for (int i = (HH ? 2 : 0); i < 4; ++i) r.v[i] = *reinterpret_cast<const F2*>(ptr[i] + (int64_t)gi * stride[i]);
    }

    // ASYNC: the four copies of quad row `qrow` into one stage (element i at st[i * kThreads]), one copy group.  Quad rows
    // that lie outside the image even after mirroring are copied from a clamped row and zeroed when they are read.
    template <bool EDGE>
    static DTCWT_D void issue_quad_row(const Args& a, const char* const (&ptr)[4], const int (&stride)[4], int qrow, F2* st) {
        int gi = qrow;
        if (EDGE) {
            bool fr;
            gi = reflect_quad(qrow, a.rows / 2, fr);
            gi = gi < 0 ? 0 : (gi >= a.rows / 2 ? a.rows / 2 - 1 : gi);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) async_copy8(st + i * kThreads, ptr[i] + (int64_t)gi * stride[i]);
        async_commit();
    }

    // c2q (transform2d.py:324-350), gains pre-scaled by 1/sqrt2:  top row (A, B), bottom row (C, D)
    static DTCWT_D void c2q_rows(const F2 w0, const F2 w1, float g0, float g1, F2& top, F2& bot) {
        const float r0 = w0.x * g0, i0 = w0.y * g0;
        top.x = fmaf(w1.x, g1, r0); top.y = fmaf(w1.y, g1, i0);
        bot.x = fmaf(-w1.y, g1, i0); bot.y = fmaf(w1.x, g1, -r0);
    }
    // symmetric extension at quad granularity: a mirrored quad has its rows and / or columns exchanged
    static DTCWT_D void flip_quad(bool fr, bool fc, F2& top, F2& bot) {
        if (fc) { float t; t = top.x; top.x = top.y; top.y = t; t = bot.x; bot.x = bot.y; bot.y = t; }
        if (fr) { const F2 t = top; top = bot; bot = t; }
    }

    template <int ROLE, bool EDGE>
    static DTCWT_D void cols_body(const Args& a, float* sm, int bx, int by, int bz, int qc, int strip) {
        // HH (band-pass launch of a `_bp` level): the lowpass and the sub-bands 0, 5, 2, 3 count as zero -- role 0 has no work,
        // role 1 filters c2q(bands 1, 4) with the band-pass pair only, the row pass reads y2 only
        if (HH && ROLE == 0) return;
        // Columns: strips on the left / right border mirror their outer quad columns (fc: swap the two columns of the
        // quad); quad columns further out feed outputs that are never stored and are clamped.  `cedge` is uniform over the
        // CTA, so interior strips skip the swaps with one branch -- keeping border strips on the same code as interior
        // ones matters: a tile is only 5 strips wide at level 2 and four code variants thrash the instruction cache.
        const bool cedge = (TWI * bx - HLC < 0) || (TWI * bx - HLC + CY > a.cols);
        int gj = (TWI * bx - HLC) / 2 + qc;                      // exact: TWI and HLC are even
        bool fc = false;
        if (cedge) {
            gj = reflect_quad(gj, a.cols / 2, fc);
            gj = gj < 0 ? 0 : (gj >= a.cols / 2 ? a.cols / 2 - 1 : gj);
        }
        const int qr0 = (Q * (GH * by + NGV * strip) - HLR) / 2;
        const float* zimg = a.z + (int64_t)bz * a.rows * a.cols + 2 * gj;
        const float* zb = a.yh + 2 * ((int64_t)bz * a.zs_n + gj);
        const float* f[4];
        if (RAW) {           // role 0: rows of s0 (filtered with g0) and s1 (g1); role 1: s2 and s3
            const float* sa = zimg + (ROLE == 0 ? 0 : 2) * a.zs_band;
            f[0] = sa; f[1] = sa + a.cols; f[2] = sa + a.zs_band; f[3] = sa + a.zs_band + a.cols;
        } else if (ROLE == 0) { f[0] = zimg; f[1] = zimg + a.cols; f[2] = zb; f[3] = zb + 2 * 5 * a.zs_band; }
        else { f[0] = zb + 2 * 2 * a.zs_band; f[1] = zb + 2 * 3 * a.zs_band; f[2] = zb + 2 * 1 * a.zs_band; f[3] = zb + 2 * 4 * a.zs_band; }
        const char* ptr[4];
        int stride[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ptr[i] = reinterpret_cast<const char*>(f[i]);
            stride[i] = (RAW || (ROLE == 0 && i < 2)) ? 8 * a.cols : 8 * (int)a.zs_row;      // bytes per quad row (the ABI bounds both)
        }
        const float ga0 = a.gain[ROLE == 0 ? 0 : 2], ga1 = a.gain[ROLE == 0 ? 5 : 3];
        const float gb0 = a.gain[1], gb1 = a.gain[4];
        F2 acc[NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) acc[i].x = acc[i].y = 0.f;
        Raw cur, nxt;
        F2* stage = reinterpret_cast<F2*>(sm + 2 * kImgFloats) + (ROLE * NSTRIP + strip) * QCOLS + qc;      // + threadIdx.x
        if (ASYNC > 0) {
#pragma unroll
            for (int d = 0; d < ASYNC; ++d)
                if (d < NQR) issue_quad_row<EDGE>(a, ptr, stride, qr0 + d, stage + d * 4 * kThreads);
        } else {
            load_quad_row<EDGE>(a, ptr, stride, qr0, cur);
        }
#pragma unroll
        for (int jq = 0; jq < NQR; ++jq) {
            if (ASYNC > 0) {
                F2* st = stage + (jq % (ASYNC > 0 ? ASYNC : 1)) * 4 * kThreads;
                async_wait<(ASYNC > 0 ? ASYNC - 1 : 0)>();               // the oldest pending group -- quad row jq -- has landed
#pragma unroll
                for (int i = 0; i < 4; ++i) cur.v[i] = st[i * kThreads];
                if (EDGE) {
                    bool fr;
                    const int gi = reflect_quad(qr0 + jq, a.rows / 2, fr);
                    if (!(gi >= 0 && gi < a.rows / 2)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) cur.v[i].x = cur.v[i].y = 0.f;
                    }
                }
                if (jq + ASYNC < NQR) issue_quad_row<EDGE>(a, ptr, stride, qr0 + jq + ASYNC, st);      // refill the slot just read
                else async_commit();                                     // an empty group keeps the count uniform
            } else if (jq + 1 < NQR) {
                load_quad_row<EDGE>(a, ptr, stride, qr0 + jq + 1, nxt);
            }
            F2 at, ab, bt, bb;         // image A (filtered with g0) and image B (g1): top / bottom real rows
            if (RAW) {
                at = cur.v[0]; ab = cur.v[1]; bt = cur.v[2]; bb = cur.v[3];
            } else if (ROLE == 0) {
                at = cur.v[0]; ab = cur.v[1];
                if (HH) { at = zero2(); ab = zero2(); }
                c2q_rows(cur.v[2], cur.v[3], ga0, ga1, bt, bb);
            } else {
                if (HH) { at = zero2(); ab = zero2(); }
                else c2q_rows(cur.v[0], cur.v[1], ga0, ga1, at, ab);
                c2q_rows(cur.v[2], cur.v[3], gb0, gb1, bt, bb);
            }
            if (EDGE) {
                bool fr;
                reflect_quad(qr0 + jq, a.rows / 2, fr);
                flip_quad(fr, false, at, ab);
                flip_quad(fr, false, bt, bb);
            }
            if (cedge) {
                flip_quad(false, fc, at, ab);
                flip_quad(false, fc, bt, bb);
            }
            if (!HH) fir_scatter<G0, NGV, HLR, TS0>(2 * jq, at, a.g0, acc);
            fir_scatter<G1, NGV, HLR, TS1>(2 * jq, bt, a.g1, acc);
            if (!HH) fir_scatter<G0, NGV, HLR, TS0>(2 * jq + 1, ab, a.g0, acc);
            fir_scatter<G1, NGV, HLR, TS1>(2 * jq + 1, bb, a.g1, acc);
            if (ASYNC == 0) cur = nxt;
        }
        if (ROWPAIR) {
            // pair row p = rows (2p, 2p+1): element (p, col, parity) at p * 2CY + 2 col + parity
            float* y = sm + ROLE * kImgFloats + (NOUT / 2 * strip) * PYP + 4 * qc;
#pragma unroll
            for (int i = 0; i < NOUT; i += 2) {
                F4 v;
                v.x = acc[i].x; v.y = acc[i + 1].x; v.z = acc[i].y; v.w = acc[i + 1].y;
                *reinterpret_cast<F4*>(y + (i / 2) * PYP) = v;
            }
            return;
        }
        float* y = sm + ROLE * RY * CY + (NOUT * strip) * CY + 2 * qc;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) *reinterpret_cast<F2*>(y + i * CY) = acc[i];
    }

    // phase 0: column pass from global memory into y1 / y2
    static DTCWT_D void phase_cols(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const int qc = tid % QCOLS;
        const int t = tid / QCOLS;
        const int role = t % 2, strip = t / 2;                   // uniform within a warp
        // quad columns right of the last window a stored output reads are nobody's input: on images narrower than a tile
        // (the 128- and 64-wide slices of a 3-D level) that is up to two thirds of the threads
        if (2 * ((TWI * bx - HLC) / 2 + qc) >= a.cols + HRC) return;
        const bool edge = (Q * GH * by - HLR < 0) || (Q * GH * (by + 1) + HRR > a.rows);      // rows only, see cols_body
        if (role == 0) {
            if (edge) cols_body<0, true>(a, sm, bx, by, bz, qc, strip);
            else cols_body<0, false>(a, sm, bx, by, bz, qc, strip);
        } else {
            if (edge) cols_body<1, true>(a, sm, bx, by, bz, qc, strip);
            else cols_body<1, false>(a, sm, bx, by, bz, qc, strip);
        }
    }

    // one row of a row-pair task: P * NGH outputs from c0 on
    static DTCWT_D void store_pair_row(const Args& a, float* img, int r, int c0, const float (&o)[P * NGH]) {
        if (r < 0 || r >= a.out_rows) return;
        float* d = img + (int64_t)r * a.out_cols + c0;
        if (a.out_vec4 && c0 + P * NGH <= a.out_cols) {                // 16-byte aligned rows, no crop
#pragma unroll
            for (int c = 0; c < (P * NGH) / 4; ++c) {
                F4 v;
                v.x = o[4 * c]; v.y = o[4 * c + 1]; v.z = o[4 * c + 2]; v.w = o[4 * c + 3];
                reinterpret_cast<F4*>(d)[c] = v;
            }
        } else if (a.crop_c == 0) {                                    // rows are 8-byte aligned
#pragma unroll
            for (int i = 0; i < P * NGH; i += 2)
                if (c0 + i < a.out_cols) {
                    F2 v;
                    v.x = o[i]; v.y = o[i + 1];
                    *reinterpret_cast<F2*>(d + i) = v;
                }
        } else {
#pragma unroll
            for (int i = 0; i < P * NGH; ++i)
                if (c0 + i >= 0 && c0 + i < a.out_cols) d[i] = o[i];
        }
    }

    // phase 1, ROWPAIR: one task = one PAIR of output rows x 4 input columns, every multiply-add a packed FFMA2
    static DTCWT_D void phase_rows_pair(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        const float* y1 = sm;
        const float* y2 = sm + kImgFloats;
        float* img = a.out + (int64_t)bz * a.out_rows * a.out_cols;
        static_assert(((RY / 2) % 2) == 0, "pair rows are handed out two at a time");
        // adjacent lanes take adjacent pair rows of the same segment (one bank group apart: PYP), lane pairs walk along the
        // segments (two bank groups apart): a quarter-warp of 16-byte loads touches eight different bank groups
        for (int task = tid; task < (RY / 2) * NSEG; task += kThreads) {
            const int half = task >> 1;
            const int lp = 2 * (half / NSEG) + (task & 1), seg = half % NSEG;
            const int r = P * GH * by + 2 * lp - a.crop_r;             // first row of the pair
            if (r + 1 < 0 || r >= a.out_rows) continue;
            if ((P / Q) * (TWI * bx + 4 * seg) - a.crop_c >= a.out_cols) continue;      // segment right of the image
            F2 acc[P * NGH];
#pragma unroll
            for (int i = 0; i < P * NGH; ++i) acc[i] = zero2();
            F2 w[WN];
            {
                const F4* src = reinterpret_cast<const F4*>(y1 + lp * PYP + 2 * (seg * 4));
#pragma unroll
                for (int c = 0; c < WN / 2; ++c) {
                    const F4 v = src[c];
                    w[2 * c].x = v.x; w[2 * c].y = v.y; w[2 * c + 1].x = v.z; w[2 * c + 1].y = v.w;
                }
                fir_gather2<G0, NGH, HLC, WN, TS0>(w, a.g0, acc);
            }
            {
                const F4* src = reinterpret_cast<const F4*>(y2 + lp * PYP + 2 * (seg * 4));
#pragma unroll
                for (int c = 0; c < WN / 2; ++c) {
                    const F4 v = src[c];
                    w[2 * c].x = v.x; w[2 * c].y = v.y; w[2 * c + 1].x = v.z; w[2 * c + 1].y = v.w;
                }
                fir_gather2<G1, NGH, HLC, WN, TS1>(w, a.g1, acc);
            }
            const int c0 = (P / Q) * (TWI * bx + 4 * seg) - a.crop_c;     // first output column of the task
            float o[P * NGH];
#pragma unroll
            for (int i = 0; i < P * NGH; ++i) o[i] = acc[i].x;
            store_pair_row(a, img, r, c0, o);
#pragma unroll
            for (int i = 0; i < P * NGH; ++i) o[i] = acc[i].y;
            store_pair_row(a, img, r + 1, c0, o);
        }
    }

    // phase 1: row pass out = H:g0(y1) + H:g1(y2); one task = one output row x 4 input columns
    static DTCWT_D void phase_rows(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (ROWPAIR) {
            phase_rows_pair(a, sm, bx, by, bz, tid);
            return;
        }
        const float* y1 = sm;
        const float* y2 = sm + RY * CY;
        float* img = a.out + (int64_t)bz * a.out_rows * a.out_cols;
        for (int task = tid; task < RY * NSEG; task += kThreads) {
            const int lr = task / NSEG, seg = task - lr * NSEG;
            const int r = P * GH * by + lr - a.crop_r;
            if (r < 0 || r >= a.out_rows) continue;
            if ((P / Q) * (TWI * bx + 4 * seg) - a.crop_c >= a.out_cols) continue;      // segment right of the image
            float acc[P * NGH];
#pragma unroll
            for (int i = 0; i < P * NGH; ++i) acc[i] = 0.f;
            float w[WN];
            if (!HH) {
                const F4* src = reinterpret_cast<const F4*>(y1 + lr * CY + seg * 4);
#pragma unroll
                for (int c = 0; c < WN / 4; ++c) {
                    const F4 v = src[c];
                    w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
                }
                fir_gather<G0, NGH, HLC, WN>(w, a.g0, acc);
            }
            {
                const F4* src = reinterpret_cast<const F4*>(y2 + lr * CY + seg * 4);
#pragma unroll
                for (int c = 0; c < WN / 4; ++c) {
                    const F4 v = src[c];
                    w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
                }
                fir_gather<G1, NGH, HLC, WN>(w, a.g1, acc);
            }
            const int c0 = (P / Q) * (TWI * bx + 4 * seg) - a.crop_c;     // first output column of the task
            float* d = img + (int64_t)r * a.out_cols + c0;
            if (HH) {
#pragma unroll
                for (int i = 0; i < P * NGH; ++i)
                    if (c0 + i >= 0 && c0 + i < a.out_cols) acc[i] += d[i];
            }
            if (a.out_vec4 && c0 + P * NGH <= a.out_cols) {                // 16-byte aligned rows, no crop
#pragma unroll
                for (int c = 0; c < (P * NGH) / 4; ++c) {
                    F4 v;
                    v.x = acc[4 * c]; v.y = acc[4 * c + 1]; v.z = acc[4 * c + 2]; v.w = acc[4 * c + 3];
                    reinterpret_cast<F4*>(d)[c] = v;
                }
            } else if (a.crop_c == 0) {                                    // rows are 8-byte aligned
#pragma unroll
                for (int i = 0; i < P * NGH; i += 2)
                    if (c0 + i < a.out_cols) {
                        F2 v;
                        v.x = acc[i]; v.y = acc[i + 1];
                        *reinterpret_cast<F2*>(d + i) = v;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < P * NGH; ++i)
                    if (c0 + i >= 0 && c0 + i < a.out_cols) d[i] = acc[i];
            }
        }
    }

    template <int PH>
    static DTCWT_D void phase(const Args& a, float* sm, int bx, int by, int bz, int tid) {
        if (PH == 0) phase_cols(a, sm, bx, by, bz, tid);
        if (PH == 1) phase_rows(a, sm, bx, by, bz, tid);
    }
};

}  // namespace dtcwt
