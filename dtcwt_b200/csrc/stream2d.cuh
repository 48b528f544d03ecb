// Streaming level-1 kernels of the 2-D DT-CWT (float32): the undecimated biorthogonal level, which is
// 2/3 of the arithmetic and of the HBM traffic of a whole transform.
//
// A CTA owns a strip of columns of one image and walks DOWN a run of rows.  The vertical filter is a
// register-resident scatter: every input row a thread loads is multiplied into a ring of RING partial
// output rows (static register indices: the loop over one period of RING rows is fully unrolled), an output
// row leaves the ring when its last contributor has passed.  Nothing is re-read vertically, there is no
// halo in the vertical direction except RING rows of warm-up per run, and every multiply-add of both passes
// is a packed FFMA2 (fma.rn.f32x2):
//   column pass   data pair (col c, col c+1)  x  scalar tap           (FFMA2 R, R.F32x2, UR.F32, R)
//   row pass      scalar data w[j]            x  tap pair (t[k], t[k-1]) accumulating outputs (c, c+1)
//                                                                     (FFMA2 R, R.F32, UR.F32x2, R)
// so neither pass needs a shuffle or a re-layout.  Taps known to be exactly zero (near_sym_b has two in each
// filter) are compiled out through the MASK template arguments; the host picks an instance whose masks
// cover the taps it was given.
//
//   inverse (reference transform2d.py:275-293, c2q :324-350)
//       thread = one quad column (2 real columns) x one role; role 0 builds y1 = V:g0(Z) + V:g1(lh),
//       role 1 builds y2 = V:g0(hl) + V:g1(hh).  Inputs come straight from global memory (8-byte coalesced
//       loads, NST quad rows prefetched in registers), c2q runs in registers.  Finished rows of y1 / y2 go
//       to shared memory; after every period the CTA runs the row pass out = H:g0(y1) + H:g1(y2) on them.
//   forward (reference transform2d.py:112-130, q2c :301-322)
//       per period the input rows are staged in shared memory (TMA box, one period ahead; periods that touch
//       the top or bottom of the image are staged with plain loads at mirrored row indices), the row pass
//       writes A = H:h0(X), B = H:h1(X)/sqrt2 to shared memory and the column pass streams them: thread =
//       one column pair x one of four roles (LoLo = V:h0(A); q2c(V:h1(A)/sqrt2); q2c(V:h0(B)); q2c(V:h1(B))).
//
// The bodies compile as plain C++ (DTCWT_EMU) for tests/emu, which runs them thread by thread.
#pragma once
#include <type_traits>

#include "async_copy.cuh"
#include "fused2d.cuh"

namespace dtcwt {

// Where the scalar taps of a column pass come from: the kernel arguments (constant bank -> uniform registers), or a
// table baked into the instance (baked_taps.h), which the compiler encodes as FFMA2 immediates -- no uniform-register
// pressure and no constant loads inside the unrolled period.
struct ArgTaps {
    static DTCWT_D float get(const ColTaps& rt, int k) { return rt.t[k]; }
};
template <class B>
struct BakedTaps {
    static DTCWT_D float get(const ColTaps&, int k) { return B::get(k); }
    static bool same(const ColTaps& rt) {
        for (int k = 0; k < B::K; ++k)
            if (!(rt.t[k] == B::get(k))) return false;
        return true;
    }
};

// Input row j (relative, static after unrolling) of a C-centred K-tap filter goes to the output rows
// j + C - k, k = 0..K-1; the k = 0 contribution is the first one an output row ever receives.
template <class TS, int K, uint32_t MASK, int C, bool FIRST, int RING>
DTCWT_D void ring_scatter(const int j, const F2 v, const ColTaps& t, F2 (&acc)[RING]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if ((MASK >> k) & 1u) {
            const int slot = pmod(j + C - k, RING);
            if (FIRST && k == 0) acc[slot] = fma2(TS::get(t, 0), v, zero2());
            else acc[slot] = fma2(TS::get(t, k), v, acc[slot]);
        }
    }
}

DTCWT_HD int fold_quad(int q, int nq, bool& flip) {
    flip = false;
    if (q < 0) { q = -1 - q; flip = true; } else if (q >= nq) { q = 2 * nq - 1 - q; flip = true; }
    return q < 0 ? 0 : (q >= nq ? nq - 1 : q);
}

// =============================================================================== inverse level 1
struct InvS1Args {
    const float* z;                 // lowpass [n][rows][cols]
    const float* yh;                // complex planar sub-bands (strides below, in complex elements)
    float* out;                     // [n][rows][cols]
    int n, rows, cols;              // even
    int periods;                    // emitting periods per run: a run covers RING * periods output rows
    int out_vec4;                   // rows of `out` are 16-byte aligned
    int64_t zs_n, zs_band, zs_row;
    float gain[6];                  // gain_mask column of this level, times 1/sqrt2
    ColTaps g0, g1;                 // column pass (scalar taps)
    PairTab p0, p1;                 // row pass (tap pairs)
};

// DBG (diagnosis builds only, results are wrong): 1 = memory traffic without the arithmetic, 2 = arithmetic without the loads
// HH: the second launch of a `_bp` level (transform2d.py:279-292): lowpass counted as zero, result ADDED to `out`
// ASYNC > 0: the prefetched quad rows live in a thread-private slice of shared memory filled by cp.async (ASYNC stages of
// four 8-byte copies per thread) instead of NST register stages: deeper prefetch at a lower register count, and still no
// barrier -- a thread only ever waits on its own copy groups (fused2d.cuh: async_copy8 / async_wait).
// UNI: both roles run ONE instruction stream (the role is a warp-uniform run-time value: role 0 discards a c2q it does not
// need) instead of two specialised copies of the unrolled period -- 19 KB of hot code instead of 38 KB, inside the 32 KB
// L1.5 instruction cache (the two-copy form showed 11 % `no_instruction` stalls, profiles/r3_02).
template <int K0, int K1, uint32_t M0, uint32_t M1, int RING_, int NST_, class T0 = ArgTaps, class T1 = ArgTaps, int MINB_ = 2, int DBG = 0,
          bool HH = false, int ASYNC_ = 0, bool UNI_ = false>
struct InvS1 {
    typedef InvS1Args Args;
    static constexpr int ASYNC = ASYNC_;
    static constexpr int RING = RING_, PER = RING_ / 2, NST = ASYNC_ > 0 ? ASYNC_ : NST_;
    static constexpr int C0 = (K0 - 1) / 2, C1 = (K1 - 1) / 2, CQ = round_up(C0, 2);
    static constexpr int kThreads = kStreamThreads;
    static constexpr int QC = kThreads / 2;                    // quad columns of a strip
    static constexpr int CY = 2 * QC;                          // columns of y1 / y2
    static constexpr int CYP = CY + 4;                         // padded pitch: odd and even rows hit disjoint banks
    static constexpr int TWI = (CY - 2 * CQ) / 8 * 8;          // output columns of a strip
    static constexpr int NSEG = TWI / 8;                       // row task = 8 output columns of one row
    static constexpr int WS0 = (CQ - C0) / 4 * 4, WE0 = round_up(CQ + 8 + C0, 4);   // y1 window of a row task
    static constexpr int WS1 = (CQ - C1) / 4 * 4, WE1 = round_up(CQ + 8 + C1, 4);   // y2 window
    static constexpr int kYFloats = 2 * RING * CYP;
    static constexpr int kSmemFloats = kYFloats + (ASYNC_ > 0 ? ASYNC_ * 4 * 2 * kThreads : 0);      // y1 / y2 + the copy stages
    static constexpr int kMinBlocks = MINB_;
    static_assert(ASYNC_ == 0 || DBG == 0, "diagnosis builds use the register stages");
    static_assert(K0 >= K1 && (K0 & 1) && (K1 & 1) && K0 <= kStreamMaxTaps && (M0 & 1u), "filter pair");
    static_assert(RING >= CQ + C0 + 1 && (RING % 2) == 0 && (PER % NST) == 0, "ring");
    static_assert(8 * (NSEG - 1) + WE0 <= CY && 8 * (NSEG - 1) + WE1 <= CY, "row-pass window inside the smem row");

    struct Raw { F2 v[4]; };       // role 0: Z top row, Z bottom row, band 0, band 5;  role 1: bands 2, 3, 1, 4
    struct Thread {
        F2 acc[RING];
        Raw st[ASYNC_ > 0 ? 1 : NST];
        const char* ptr[4];        // byte addresses of this thread's column in quad row 0 of its four inputs
        int stride[4];             // bytes per quad row
        int fc;
        int idle;                  // HH: role 0 (lowpass + bands 0, 5: all zero in the band-pass launch) has nothing to do
        F2* stage;                 // ASYNC: this thread's slice of the copy stages (element i of stage s at stage[(4 s + i) * kThreads])
    };
    static_assert(!HH || K0 == K1, "HH: the band-pass filter sits in both slots (first touch of a ring row comes from it)");

    static DTCWT_HD int run_rows(const Args& a) { return RING * a.periods; }
    static DTCWT_HD int tiles_c(const Args& a) { return (a.cols + TWI - 1) / TWI; }
    static DTCWT_HD int tiles_r(const Args& a) { return (a.rows + run_rows(a) - 1) / run_rows(a); }
    // periods a run executes: one of warm-up plus the emitting ones that still start inside the image
    static DTCWT_HD int run_periods(const Args& a, int by) {
        const int left = a.rows - run_rows(a) * by;
        const int e = (left + RING - 1) / RING;
        return 1 + (e < a.periods ? e : a.periods);
    }
    // quad row consumed by step 0 of period p (negative above the image)
    static DTCWT_HD int quad_base(const Args& a, int by, int p) { return (run_rows(a) * by - RING + CQ) / 2 + PER * p; }
    static DTCWT_HD bool edge_period(const Args& a, int bx, int by, int p) {
        const int qb = quad_base(a, by, p);
        return qb < 0 || (qb + PER + NST > a.rows / 2);           // rows only; columns: col_edge()
    }
    // strips on the left / right border mirror their outer quad columns; uniform over the CTA
    static DTCWT_HD bool col_edge(const Args& a, int bx) { return (TWI * bx - CQ < 0) || (TWI * bx - CQ + CY > a.cols); }

    template <bool EDGE>
    static DTCWT_D void load_stage(const Args& a, const Thread& th, Raw& r, int q) {
        if (EDGE) {
            bool f;
            q = fold_quad(q, a.rows / 2, f);
        }
        if (DBG == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { r.v[i].x = (float)q; r.v[i].y = (float)(q + i); }
            return;
        }
        if (HH && th.idle) return;
#pragma unroll
        for (int i = (HH ? 2 : 0); i < 4; ++i) r.v[i] = *reinterpret_cast<const F2*>(th.ptr[i] + (int64_t)q * th.stride[i]);   // one IMAD.WIDE
    }

    // ASYNC: the four copies of quad row q into stage slot `slot`, one copy group (HH: only bands 1 and 4 of role 1)
    template <bool EDGE>
    static DTCWT_D void issue_stage(const Args& a, const Thread& th, int slot, int q) {
        if (EDGE) {
            bool f;
            q = fold_quad(q, a.rows / 2, f);
        }
        if (!(HH && th.idle)) {
#pragma unroll
            for (int i = (HH ? 2 : 0); i < 4; ++i)
                async_copy8(th.stage + (4 * slot + i) * kThreads, th.ptr[i] + (int64_t)q * th.stride[i]);
        }
        async_commit();
    }
    static DTCWT_D void read_stage(const Thread& th, int slot, Raw& r) {
        async_wait<NST - 1>();                                     // the oldest pending group has landed
#pragma unroll
        for (int i = (HH ? 2 : 0); i < 4; ++i) r.v[i] = th.stage[(4 * slot + i) * kThreads];
    }

    static DTCWT_D void init(const Args& a, Thread& th, int bx, int by, int bz, int tid, float* sm = nullptr) {
        const int qc = tid % QC, role = tid / QC;
        bool fc;
        const int gj = fold_quad((TWI * bx - CQ) / 2 + qc, a.cols / 2, fc);
        th.fc = fc ? 1 : 0;
        th.idle = (HH && role == 0) ? 1 : 0;
        const float* zimg = a.z + (int64_t)bz * a.rows * a.cols + 2 * gj;
        const float* yb = a.yh + 2 * ((int64_t)bz * a.zs_n + gj);
        const int sz = 8 * a.cols, sb = 8 * (int)a.zs_row;       // bytes per quad row (the ABI bounds both)
        const float* f[4];
        if (role == 0) {
            f[0] = zimg; f[1] = zimg + a.cols; f[2] = yb; f[3] = yb + 2 * 5 * a.zs_band;
            th.stride[0] = sz; th.stride[1] = sz; th.stride[2] = sb; th.stride[3] = sb;
        } else {
            f[0] = yb + 2 * 2 * a.zs_band; f[1] = yb + 2 * 3 * a.zs_band;
            f[2] = yb + 2 * 1 * a.zs_band; f[3] = yb + 2 * 4 * a.zs_band;
            th.stride[0] = sb; th.stride[1] = sb; th.stride[2] = sb; th.stride[3] = sb;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) th.ptr[i] = reinterpret_cast<const char*>(f[i]);
#pragma unroll
        for (int i = 0; i < RING; ++i) th.acc[i] = zero2();
        const int q0 = quad_base(a, by, 0);
        if (ASYNC > 0) {
            th.stage = reinterpret_cast<F2*>(sm + kYFloats) + tid;
#pragma unroll
            for (int s = 0; s < NST; ++s) issue_stage<true>(a, th, s, q0 + s);
            return;
        }
#pragma unroll
        for (int s = 0; s < (ASYNC_ > 0 ? 1 : NST); ++s) load_stage<true>(a, th, th.st[s], q0 + s);
    }

    // c2q (transform2d.py:324-350), gains pre-scaled by 1/sqrt2:  top row (A, B), bottom row (C, D)
    static DTCWT_D void c2q_rows(const F2 w0, const F2 w1, float g0, float g1, F2& top, F2& bot) {
        const float r0 = w0.x * g0, i0 = w0.y * g0;
        top.x = fmaf(w1.x, g1, r0); top.y = fmaf(w1.y, g1, i0);
        bot.x = fmaf(-w1.y, g1, i0); bot.y = fmaf(w1.x, g1, -r0);
    }
    static DTCWT_D void flip_quad(bool fr, bool fc, F2& top, F2& bot) {
        if (fc) { float t; t = top.x; top.x = top.y; top.y = t; t = bot.x; bot.x = bot.y; bot.y = t; }
        if (fr) { const F2 t = top; top = bot; bot = t; }
    }

    // ROLE 0 / 1: specialised copies; ROLE 2 (UNI): one copy, `role` decides at run time (uniform within a warp)
    template <int ROLE, bool EDGE>
    static DTCWT_D void cols_role(const Args& a, Thread& th, float* sm, int by, int p, int qc, bool cedge, int role = ROLE) {
        const bool r0 = (ROLE == 2) ? (role == 0) : (ROLE == 0);
        // HH (the band-pass launch of a `_bp` level): the lowpass and the sub-bands 0, 5, 2, 3 count as zero, so role 0 has no
        // work at all and role 1 only filters c2q(bands 1, 4) with the band-pass filter; the row pass reads y2 only
        if (HH && r0) return;
        const int qb = quad_base(a, by, p);
        const bool emit = p > 0;
        const float ga0 = a.gain[r0 ? 0 : 2], ga1 = a.gain[r0 ? 5 : 3];
        const float gb0 = (ROLE == 2 && r0) ? a.gain[0] : a.gain[1], gb1 = (ROLE == 2 && r0) ? a.gain[5] : a.gain[4];
        float* y = sm + (r0 ? 0 : 1) * (RING * CYP) + 2 * qc;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            Raw cur;
            if (ASYNC > 0) {
                read_stage(th, u % NST, cur);
                issue_stage<EDGE>(a, th, u % NST, qb + u + NST);       // refill the slot just read (same thread: program order)
            } else {
                cur = th.st[(ASYNC_ > 0 ? 0 : u % NST)];
                load_stage<EDGE>(a, th, th.st[(ASYNC_ > 0 ? 0 : u % NST)], qb + u + NST);
            }
            F2 at, ab, bt, bb;         // image A (filtered with g0) and image B (g1): top / bottom real rows
            if (ROLE == 2) {
                // one stream for both roles: role 1 needs c2q(bands 2, 3) for A, role 0 takes the two lowpass rows as they are
                F2 ct, cb;
                c2q_rows(cur.v[0], cur.v[1], a.gain[2], a.gain[3], ct, cb);
                at = r0 ? cur.v[0] : ct; ab = r0 ? cur.v[1] : cb;
                if (HH) { at = zero2(); ab = zero2(); }
                c2q_rows(cur.v[2], cur.v[3], gb0, gb1, bt, bb);
            } else if (ROLE == 0) {
                at = cur.v[0]; ab = cur.v[1];
                c2q_rows(cur.v[2], cur.v[3], ga0, ga1, bt, bb);
            } else {
                if (HH) { at = zero2(); ab = zero2(); }
                else c2q_rows(cur.v[0], cur.v[1], ga0, ga1, at, ab);
                c2q_rows(cur.v[2], cur.v[3], gb0, gb1, bt, bb);
            }
            if (EDGE) {
                bool fr;
                fold_quad(qb + u, a.rows / 2, fr);
                flip_quad(fr, false, at, ab);
                flip_quad(fr, false, bt, bb);
            }
            if (cedge) {
                flip_quad(false, th.fc != 0, at, ab);
                flip_quad(false, th.fc != 0, bt, bb);
            }
            if (DBG == 1) {
                th.acc[pmod(2 * u - CQ, RING)].x = at.x + bt.y; th.acc[pmod(2 * u - CQ, RING)].y = at.y + bt.x;
                th.acc[pmod(2 * u + 1 - CQ, RING)].x = ab.x + bb.y; th.acc[pmod(2 * u + 1 - CQ, RING)].y = ab.y + bb.x;
            } else {
                if (!HH) ring_scatter<T0, K0, M0, C0, true, RING>(2 * u, at, a.g0, th.acc);
                ring_scatter<T1, K1, M1, C1, HH, RING>(2 * u, bt, a.g1, th.acc);         // HH: the first touch of a ring row is its own
            }
            if (emit) {                // rows 2u and 2u+1 of this period's block are complete
                *reinterpret_cast<F2*>(y + (2 * u) * CYP) = th.acc[pmod(2 * u - CQ, RING)];
                *reinterpret_cast<F2*>(y + (2 * u + 1) * CYP) = th.acc[pmod(2 * u + 1 - CQ, RING)];
            }
            if (DBG != 1) {
                if (!HH) ring_scatter<T0, K0, M0, C0, true, RING>(2 * u + 1, ab, a.g0, th.acc);
                ring_scatter<T1, K1, M1, C1, HH, RING>(2 * u + 1, bb, a.g1, th.acc);
            }
        }
    }

    // column pass of period p: consumes PER quad rows; for p > 0 leaves RING rows of y1 / y2 in shared memory
    static DTCWT_D void cols(const Args& a, Thread& th, float* sm, int bx, int by, int bz, int tid, int p) {
        const int qc = tid % QC, role = tid / QC;                // role is uniform within a warp
        const bool edge = edge_period(a, bx, by, p), cedge = col_edge(a, bx);
        if (UNI_) {
            if (edge) cols_role<2, true>(a, th, sm, by, p, qc, cedge, role);
            else cols_role<2, false>(a, th, sm, by, p, qc, cedge, role);
        } else if (role == 0) {
            if (edge) cols_role<0, true>(a, th, sm, by, p, qc, cedge);
            else cols_role<0, false>(a, th, sm, by, p, qc, cedge);
        } else {
            if (edge) cols_role<1, true>(a, th, sm, by, p, qc, cedge);
            else cols_role<1, false>(a, th, sm, by, p, qc, cedge);
        }
    }

    // row pass of period p (p > 0): out = H:g0(y1) + H:g1(y2) on the RING rows the column pass just finished
    static DTCWT_D void rows(const Args& a, float* sm, int bx, int by, int bz, int tid, int p) {
        const float* y1 = sm;
        const float* y2 = sm + RING * CYP;
        float* img = a.out + (int64_t)bz * a.rows * a.cols;
        const int r0 = run_rows(a) * by + RING * (p - 1);
        // task t = tid + kThreads * round covers row 2 * (t/2 / NSEG) + (t & 1), segment (t/2) % NSEG: adjacent lanes take
        // the two rows of a pair (their smem rows sit 4 banks apart), so a quarter-warp of 16-byte loads is conflict-free
        int rp = (tid >> 1) / NSEG, seg = (tid >> 1) % NSEG;
#pragma unroll 1
        for (int task = tid; task < RING * NSEG; task += kThreads, rp += (kThreads / 2) / NSEG, seg += (kThreads / 2) % NSEG) {
            if (seg >= NSEG) { seg -= NSEG; ++rp; }
            const int yr = 2 * rp + (tid & 1);
            const int r = r0 + yr, c0 = TWI * bx + 8 * seg;
            if (r >= a.rows || c0 >= a.cols) continue;
            F2 acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = zero2();
            const F4* s1 = reinterpret_cast<const F4*>(y1 + yr * CYP + 8 * seg + WS0);
            const F4* s2 = reinterpret_cast<const F4*>(y2 + yr * CYP + 8 * seg + WS1);
            if (DBG == 1) {
                const F4 p = s1[1], q = s1[2], r = s2[1], t = s2[2];
                acc[0].x = p.x + r.x; acc[0].y = p.y + r.y; acc[1].x = p.z + r.z; acc[1].y = p.w + r.w;
                acc[2].x = q.x + t.x; acc[2].y = q.y + t.y; acc[3].x = q.z + t.z; acc[3].y = q.w + t.w;
            } else {
                if (!HH) {
#pragma unroll
                    for (int c = 0; c < (WE0 - WS0) / 4; ++c)
                        pair_gather4<K0, M0, C0, WS0 - CQ, 8, WE0 - WS0>(4 * c, s1[c], a.p0, acc);
                }
#pragma unroll
                for (int c = 0; c < (WE1 - WS1) / 4; ++c)
                    pair_gather4<K1, M1, C1, WS1 - CQ, 8, WE1 - WS1>(4 * c, s2[c], a.p1, acc);
            }
            float* d = img + (int64_t)r * a.cols + c0;
            if (HH) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c0 + 2 * i < a.cols) { const F2 o = *reinterpret_cast<const F2*>(d + 2 * i); acc[i].x += o.x; acc[i].y += o.y; }
            }
            if (a.out_vec4 && c0 + 8 <= a.cols) {
                F4 v;
                v.x = acc[0].x; v.y = acc[0].y; v.z = acc[1].x; v.w = acc[1].y;
                reinterpret_cast<F4*>(d)[0] = v;
                v.x = acc[2].x; v.y = acc[2].y; v.z = acc[3].x; v.w = acc[3].y;
                reinterpret_cast<F4*>(d)[1] = v;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c0 + 2 * i < a.cols) *reinterpret_cast<F2*>(d + 2 * i) = acc[i];      // cols is even
            }
        }
    }
};

// =============================================================================== inverse level 1, staged inputs
// InvS1 with its eight input streams (two lowpass rows, six sub-band rows per quad row) staged in shared memory by
// asynchronous bulk copies (cp.async.bulk, the TMA unit) instead of per-thread global loads: a ninth, PRODUCER warp
// issues the eight 1 KB row segments of every quad row into a ring of NSTAGE stages and runs as far ahead as the ring
// allows; a `full` mbarrier per stage counts the bytes as they land, an `empty` one the consumer warps that have taken
// their four values.  The eight column-pass warps never wait on a global load, up to NSTAGE x 8 KB per CTA are in
// flight whatever the register budget, and the prefetch registers of InvS1 are gone.  (A first version let consumer
// threads issue the copies in turn: the address arithmetic of one lane then sat on the critical path of all eight
// warps, which the ring couples to within NSTAGE - DEPTH steps of each other -- 2.2 ms instead of 1.2, profiles/r2.)  Segments are 16-byte aligned because CQ is a multiple of 4
// (the strip's first quad column is even); on the left / right image border only the part of the segment that lies
// inside the image is copied and the threads read their mirrored quad column (index th.idx) with the two columns
// exchanged, exactly as InvS1 does.  The host falls back to InvS1 when rows are not 16-byte aligned.
template <int K0, int K1, uint32_t M0, uint32_t M1, int RING_, int NSTAGE_, class T0 = ArgTaps, class T1 = ArgTaps>
struct InvS1T {
    typedef InvS1Args Args;
    static constexpr int RING = RING_, PER = RING_ / 2, NSTAGE = NSTAGE_;
    static constexpr int kLaunchThreads = kStreamThreads + 32;        // eight consumer warps + the producer warp
    static constexpr int C0 = (K0 - 1) / 2, C1 = (K1 - 1) / 2, CQ = round_up(C0, 4);
    static constexpr int kThreads = kStreamThreads;
    static constexpr int QC = kThreads / 2;
    static constexpr int CY = 2 * QC;
    static constexpr int CYP = CY + 4;
    static constexpr int TWI = (CY - 2 * CQ) / 8 * 8;
    static constexpr int NSEG = TWI / 8;
    static constexpr int WS0 = (CQ - C0) / 4 * 4, WE0 = round_up(CQ + 8 + C0, 4);
    static constexpr int WS1 = (CQ - C1) / 4 * 4, WE1 = round_up(CQ + 8 + C1, 4);
    static constexpr int kYFloats = 2 * RING * CYP;
    static constexpr int kStreamFloats = 2 * QC;                // one staged row segment: QC complex values / 2 QC reals
    static constexpr int kStageFloats = 8 * kStreamFloats;
    static constexpr int kSmemFloats = kYFloats + NSTAGE * kStageFloats;
    static constexpr int kMinBlocks = 2;
    static_assert(K0 >= K1 && (K0 & 1) && (K1 & 1) && K0 <= kStreamMaxTaps && (M0 & 1u), "filter pair");
    static_assert(RING >= CQ + C0 + 1 && (RING % 2) == 0 && (PER % NSTAGE) == 0 && NSTAGE >= 2, "ring / pipeline");
    static_assert(8 * (NSEG - 1) + WE0 <= CY && 8 * (NSEG - 1) + WE1 <= CY, "row-pass window inside the smem row");
    static_assert((kYFloats % 4) == 0 && (TWI % 4) == 0 && (CQ % 4) == 0, "16-byte aligned stages and segments");

    struct Raw { F2 v[4]; };       // role 0: Z top row, Z bottom row, band 0, band 5;  role 1: bands 2, 3, 1, 4
    struct Thread {
        F2 acc[RING];
        int idx;                   // this thread's (mirrored) quad column inside a staged segment
        int fc;
    };
    struct Pipe { Mbar* full; Mbar* empty; };

    static DTCWT_HD int run_rows(const Args& a) { return RING * a.periods; }
    static DTCWT_HD int tiles_c(const Args& a) { return (a.cols + TWI - 1) / TWI; }
    static DTCWT_HD int tiles_r(const Args& a) { return (a.rows + run_rows(a) - 1) / run_rows(a); }
    static DTCWT_HD int run_periods(const Args& a, int by) {
        const int left = a.rows - run_rows(a) * by;
        const int e = (left + RING - 1) / RING;
        return 1 + (e < a.periods ? e : a.periods);
    }
    static DTCWT_HD int quad_base(const Args& a, int by, int p) { return (run_rows(a) * by - RING + CQ) / 2 + PER * p; }
    static DTCWT_HD bool edge_period(const Args& a, int by, int p) {
        const int qb = quad_base(a, by, p);
        return qb < 0 || (qb + PER > a.rows / 2);
    }
    static DTCWT_HD bool col_edge(const Args& a, int bx) { return (TWI * bx - CQ < 0) || (TWI * bx - CQ + CY > a.cols); }
    static DTCWT_HD int first_quad(int bx) { return (TWI * bx - CQ) / 2; }        // exact: both even; may be negative

    // Producer warp, lanes 0..7: lane i owns input stream i (0, 1: the two lowpass rows of a quad row; 2..7: sub-bands 0, 5,
    // 2, 3, 1, 4).  Everything that does not change along the run -- the clipped segment, its first byte in quad row 0,
    // the row pitch, the slot inside a stage -- is set up once, so a step costs a fold, one multiply-add and one copy.
    struct Stream {
        const float* base;         // first float of the (clipped) segment in quad row 0
        int64_t pitch;             // floats per quad row
        int dst;                   // float offset of the segment inside a stage
        uint32_t bytes;
    };
    static DTCWT_D void stream_setup(const Args& a, Stream& s, int bx, int bz, int i) {
        const int g0 = first_quad(bx), nq = a.cols / 2;
        const int s0 = g0 < 0 ? 0 : g0;
        const int s1 = (g0 + QC < nq) ? g0 + QC : nq;
        s.bytes = (uint32_t)(s1 - s0) * 8u;
        s.dst = i * kStreamFloats + 2 * (s0 - g0);
        if (i < 2) {
            s.base = a.z + (int64_t)bz * a.rows * a.cols + (int64_t)i * a.cols + 2 * s0;
            s.pitch = 2 * (int64_t)a.cols;
        } else {
            const int band = (i == 2) ? 0 : (i == 3) ? 5 : (i == 4) ? 2 : (i == 5) ? 3 : (i == 6) ? 1 : 4;
            s.base = a.yh + 2 * ((int64_t)bz * a.zs_n + (int64_t)band * a.zs_band + s0);
            s.pitch = 2 * a.zs_row;
        }
    }
    // lane 0, before the copies of step g: every consumer warp has released the stage's previous content
    static DTCWT_D void step_begin(const Stream& s, const Pipe& pipe, int g) {
        const int stage = g % NSTAGE;
        if (g >= NSTAGE) mbar_wait(&pipe.empty[stage], (uint32_t)((g / NSTAGE) - 1) & 1u);
        mbar_expect_tx(&pipe.full[stage], 8u * s.bytes);
    }
    // lane i: the row segment of stream i for step g (quad row quad_base(by, 0) + g, folded into the image)
    static DTCWT_D void step_copy(const Args& a, const Stream& s, float* sm, const Pipe& pipe, int by, int g) {
        const int stage = g % NSTAGE;
        bool f;
        const int q = fold_quad(quad_base(a, by, 0) + g, a.rows / 2, f);
        bulk_copy(sm + kYFloats + stage * kStageFloats + s.dst, s.base + (int64_t)q * s.pitch, s.bytes, &pipe.full[stage]);
    }
    // all eight streams of step g by one thread (host emulator)
    static DTCWT_D void produce(const Args& a, float* sm, const Pipe& pipe, int bx, int by, int bz, int g) {
        for (int i = 0; i < 8; ++i) {
            Stream s;
            stream_setup(a, s, bx, bz, i);
            if (i == 0) step_begin(s, pipe, g);
            step_copy(a, s, sm, pipe, by, g);
        }
    }

    static DTCWT_D void init(const Args& a, Thread& th, float* sm, const Pipe& pipe, int bx, int by, int bz, int tid) {
        const int qc = tid % QC;
        bool fc;
        int gj = fold_quad(first_quad(bx) + qc, a.cols / 2, fc);
        th.fc = fc ? 1 : 0;
        // quad columns far outside the image only feed outputs that are never stored: keep them inside the copied part
        const int g0 = first_quad(bx), nq = a.cols / 2;
        const int s0 = g0 < 0 ? 0 : g0, s1 = (g0 + QC < nq) ? g0 + QC : nq;
        gj = gj < s0 ? s0 : (gj >= s1 ? s1 - 1 : gj);
        th.idx = gj - g0;
#pragma unroll
        for (int i = 0; i < RING; ++i) th.acc[i] = zero2();
    }
    static DTCWT_HD int total_steps(const Args& a, int by) { return run_periods(a, by) * PER; }

    static DTCWT_D void c2q_rows(const F2 w0, const F2 w1, float g0, float g1, F2& top, F2& bot) {
        const float r0 = w0.x * g0, i0 = w0.y * g0;
        top.x = fmaf(w1.x, g1, r0); top.y = fmaf(w1.y, g1, i0);
        bot.x = fmaf(-w1.y, g1, i0); bot.y = fmaf(w1.x, g1, -r0);
    }
    static DTCWT_D void flip_quad(bool fr, bool fc, F2& top, F2& bot) {
        if (fc) { float t; t = top.x; top.x = top.y; top.y = t; t = bot.x; bot.x = bot.y; bot.y = t; }
        if (fr) { const F2 t = top; top = bot; bot = t; }
    }

    // step u of period p for one thread: take the staged quad row, c2q, scatter into the ring, emit two finished rows
    template <int ROLE, bool EDGE>
    static DTCWT_D void step_role(const Args& a, Thread& th, float* sm, const Pipe& pipe, int bx, int by, int bz, int tid, int p,
                                  const int u, bool cedge) {
        const int g = p * PER + u;
        const int stage = u % NSTAGE;                              // PER is a multiple of NSTAGE
        mbar_wait(&pipe.full[stage], (uint32_t)(g / NSTAGE) & 1u);
        const float* st = sm + kYFloats + stage * kStageFloats + (4 * ROLE) * kStreamFloats + 2 * th.idx;
        Raw cur;
#pragma unroll
        for (int i = 0; i < 4; ++i) cur.v[i] = *reinterpret_cast<const F2*>(st + i * kStreamFloats);
        warp_release(&pipe.empty[stage], tid);                  // ONE arrival per warp: 256 arrivals on one barrier word serialise
        const int qb = quad_base(a, by, p);
        const bool emit = p > 0;
        const float ga0 = a.gain[ROLE == 0 ? 0 : 2], ga1 = a.gain[ROLE == 0 ? 5 : 3];
        const float gb0 = a.gain[1], gb1 = a.gain[4];
        float* y = sm + ROLE * (RING * CYP) + 2 * (tid % QC);
        F2 at, ab, bt, bb;
        if (ROLE == 0) {
            at = cur.v[0]; ab = cur.v[1];
            c2q_rows(cur.v[2], cur.v[3], ga0, ga1, bt, bb);
        } else {
            c2q_rows(cur.v[0], cur.v[1], ga0, ga1, at, ab);
            c2q_rows(cur.v[2], cur.v[3], gb0, gb1, bt, bb);
        }
        if (EDGE) {
            bool fr;
            fold_quad(qb + u, a.rows / 2, fr);
            flip_quad(fr, false, at, ab);
            flip_quad(fr, false, bt, bb);
        }
        if (cedge) {
            flip_quad(false, th.fc != 0, at, ab);
            flip_quad(false, th.fc != 0, bt, bb);
        }
        ring_scatter<T0, K0, M0, C0, true, RING>(2 * u, at, a.g0, th.acc);
        ring_scatter<T1, K1, M1, C1, false, RING>(2 * u, bt, a.g1, th.acc);
        if (emit) {
            *reinterpret_cast<F2*>(y + (2 * u) * CYP) = th.acc[pmod(2 * u - CQ, RING)];
            *reinterpret_cast<F2*>(y + (2 * u + 1) * CYP) = th.acc[pmod(2 * u + 1 - CQ, RING)];
        }
        ring_scatter<T0, K0, M0, C0, true, RING>(2 * u + 1, ab, a.g0, th.acc);
        ring_scatter<T1, K1, M1, C1, false, RING>(2 * u + 1, bb, a.g1, th.acc);
    }

    static DTCWT_D void step(const Args& a, Thread& th, float* sm, const Pipe& pipe, int bx, int by, int bz, int tid, int p, const int u) {
        const int role = tid / QC;                                 // uniform within a warp
        const bool edge = edge_period(a, by, p), cedge = col_edge(a, bx);
        if (role == 0) {
            if (edge) step_role<0, true>(a, th, sm, pipe, bx, by, bz, tid, p, u, cedge);
            else step_role<0, false>(a, th, sm, pipe, bx, by, bz, tid, p, u, cedge);
        } else {
            if (edge) step_role<1, true>(a, th, sm, pipe, bx, by, bz, tid, p, u, cedge);
            else step_role<1, false>(a, th, sm, pipe, bx, by, bz, tid, p, u, cedge);
        }
    }

    template <int ROLE, bool EDGE>
    static DTCWT_D void cols_role(const Args& a, Thread& th, float* sm, const Pipe& pipe, int bx, int by, int bz, int tid, int p, bool cedge) {
#pragma unroll
        for (int u = 0; u < PER; ++u) step_role<ROLE, EDGE>(a, th, sm, pipe, bx, by, bz, tid, p, u, cedge);
    }
    // column pass of period p (device: the PER steps unrolled, ring indices static)
    static DTCWT_D void cols(const Args& a, Thread& th, float* sm, const Pipe& pipe, int bx, int by, int bz, int tid, int p) {
        const int role = tid / QC;
        const bool edge = edge_period(a, by, p), cedge = col_edge(a, bx);
        if (role == 0) {
            if (edge) cols_role<0, true>(a, th, sm, pipe, bx, by, bz, tid, p, cedge);
            else cols_role<0, false>(a, th, sm, pipe, bx, by, bz, tid, p, cedge);
        } else {
            if (edge) cols_role<1, true>(a, th, sm, pipe, bx, by, bz, tid, p, cedge);
            else cols_role<1, false>(a, th, sm, pipe, bx, by, bz, tid, p, cedge);
        }
    }

    // row pass of period p (p > 0): out = H:g0(y1) + H:g1(y2) on the RING rows the column pass just finished
    static DTCWT_D void rows(const Args& a, float* sm, int bx, int by, int bz, int tid, int p) {
        const float* y1 = sm;
        const float* y2 = sm + RING * CYP;
        float* img = a.out + (int64_t)bz * a.rows * a.cols;
        const int r0 = run_rows(a) * by + RING * (p - 1);
        int rp = (tid >> 1) / NSEG, seg = (tid >> 1) % NSEG;
#pragma unroll 1
        for (int task = tid; task < RING * NSEG; task += kThreads, rp += (kThreads / 2) / NSEG, seg += (kThreads / 2) % NSEG) {
            if (seg >= NSEG) { seg -= NSEG; ++rp; }
            const int yr = 2 * rp + (tid & 1);
            const int r = r0 + yr, c0 = TWI * bx + 8 * seg;
            if (r >= a.rows || c0 >= a.cols) continue;
            F2 acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = zero2();
            const F4* s1 = reinterpret_cast<const F4*>(y1 + yr * CYP + 8 * seg + WS0);
            const F4* s2 = reinterpret_cast<const F4*>(y2 + yr * CYP + 8 * seg + WS1);
#pragma unroll
            for (int c = 0; c < (WE0 - WS0) / 4; ++c)
                pair_gather4<K0, M0, C0, WS0 - CQ, 8, WE0 - WS0>(4 * c, s1[c], a.p0, acc);
#pragma unroll
            for (int c = 0; c < (WE1 - WS1) / 4; ++c)
                pair_gather4<K1, M1, C1, WS1 - CQ, 8, WE1 - WS1>(4 * c, s2[c], a.p1, acc);
            float* d = img + (int64_t)r * a.cols + c0;
            if (a.out_vec4 && c0 + 8 <= a.cols) {
                F4 v;
                v.x = acc[0].x; v.y = acc[0].y; v.z = acc[1].x; v.w = acc[1].y;
                reinterpret_cast<F4*>(d)[0] = v;
                v.x = acc[2].x; v.y = acc[2].y; v.z = acc[3].x; v.w = acc[3].y;
                reinterpret_cast<F4*>(d)[1] = v;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c0 + 2 * i < a.cols) *reinterpret_cast<F2*>(d + 2 * i) = acc[i];
            }
        }
    }
};

// =============================================================================== inverse, q-shift levels
// Levels >= 2 of the inverse (reference transform2d.py:240-273): colifilt (lowlevel.py:156-260) interpolates 1:2, so a quad
// row of input (2 rows) finishes 4 rows of y1 / y2.  Same streaming structure as InvS1; the ring holds m/2 groups of 4
// output rows.  An input row of parity e feeds output phases (0, 2) of the positive-correlation filter g0 and (1, 3) of
// the negative one g1 when e is even, and the other way round when it is odd (SpecInt::b); even rows arrive first, so
// they carry the first touch of a group.
struct InvSqArgs {
    const float* z;                 // lowpass [n][rows][cols]
    const float* yh;                // complex planar sub-bands (strides below, complex elements)
    float* out;                     // [n][out_rows][out_cols]
    int n, rows, cols;              // even
    int crop_r, crop_c;             // 1: drop the first and last output row / column (transform2d.py:263-268)
    int out_rows, out_cols;         // 2*rows - 2*crop_r, 2*cols - 2*crop_c
    int periods;                    // emitting periods per run (a run covers RING * periods uncropped output rows)
    int out_vec4;                   // output rows 16-byte aligned and not cropped
    int64_t zs_n, zs_band, zs_row;
    float gain[6];
    PhaseTaps g0, g1;               // column pass: t[ph][k], out[4i+ph] = sum_k t[ph][k] in[2i + b(ph) + 2k]
    PairTab q[4];                   // row pass tap pairs: (g0 phases 0,2) (g1 phases 0,2) (g0 phases 1,3) (g1 phases 1,3)
};

struct ArgPhase {
    static DTCWT_D float get(const PhaseTaps& t, int ph, int k) { return t.t[ph][k]; }
};
template <class B>
struct BakedPhase2 {
    static DTCWT_D float get(const PhaseTaps&, int ph, int k) { return B::get(ph, k); }
    static bool same(const PhaseTaps& t) {
        for (int ph = 0; ph < 4; ++ph)
            for (int k = 0; k < B::K; ++k)
                if (!(t.t[ph][k] == B::get(ph, k))) return false;
        return true;
    }
};

template <int M, int RING_, int NST_, class T0 = ArgPhase, class T1 = ArgPhase>
struct InvSq {
    typedef InvSqArgs Args;
    typedef SpecInt<M, true> G0;
    typedef SpecInt<M, false> G1;
    static constexpr int K = M / 2;                            // taps per output phase
    static constexpr int RING = RING_, PER = RING_ / 4, NST = NST_;
    static constexpr int HG = (K - 1) / 2;                     // groups between first touch and the input quad row
    static constexpr int kThreads = kStreamThreads;
    static constexpr int QC = kThreads / 2, CY = 2 * QC, CYP = CY + 4;
    static constexpr int HC = K - 1;                           // y columns left of a strip's first input column (even)
    static constexpr int NGR = 4;                              // groups of a row task: 8 input columns, 16 outputs
    static constexpr int TWI = (CY - HC - (K + 1)) / (2 * NGR) * (2 * NGR);   // input columns of a strip
    static constexpr int NSEG = TWI / (2 * NGR);
    static constexpr int WN = round_up(2 * NGR + 2 * (K - 1), 4);              // y window of a row task
    static constexpr int kSmemFloats = 2 * RING * CYP;
    static constexpr int kMinBlocks = 2;
    static_assert((K & 1) && (RING % 4) == 0 && RING >= 4 * (K + 1) && PER >= K - 1 && (PER % NST) == 0, "ring");
    static_assert(G0::b(0) == -K + 1 && G0::b(1) == -K + 2 && G1::b(0) == -K + 2 && G1::b(1) == -K + 1, "phase pairing");
    static_assert(WN + 2 * NGR * (NSEG - 1) <= CY, "row-pass window inside the smem row");

    struct Raw { F2 v[4]; };
    struct Thread {
        F2 acc[RING];
        Raw st[NST];
        const char* ptr[4];
        int stride[4];
        int fc;
    };

    static DTCWT_HD int run_rows(const Args& a) { return RING * a.periods; }          // uncropped output rows per run
    static DTCWT_HD int tiles_c(const Args& a) { return (a.cols + TWI - 1) / TWI; }
    static DTCWT_HD int tiles_r(const Args& a) { return (2 * a.rows + run_rows(a) - 1) / run_rows(a); }
    static DTCWT_HD int run_periods(const Args& a, int by) {
        const int left = 2 * a.rows - run_rows(a) * by;
        const int e = (left + RING - 1) / RING;
        return 1 + (e < a.periods ? e : a.periods);
    }
    // input quad row consumed by step 0 of period p (negative above the image)
    static DTCWT_HD int quad_base(const Args& a, int by, int p) { return run_rows(a) * by / 4 - PER + HG + PER * p; }
    static DTCWT_HD bool edge_period(const Args& a, int bx, int by, int p) {
        const int qb = quad_base(a, by, p);
        return (TWI * bx - HC < 0) || (TWI * bx - HC + CY > a.cols) || qb < 0 || (qb + PER + NST > a.rows / 2);
    }

    template <bool EDGE>
    static DTCWT_D void load_stage(const Args& a, const Thread& th, Raw& r, int q) {
        if (EDGE) {
            bool f;
            q = fold_quad(q, a.rows / 2, f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) r.v[i] = *reinterpret_cast<const F2*>(th.ptr[i] + (int64_t)q * th.stride[i]);
    }

    static DTCWT_D void init(const Args& a, Thread& th, int bx, int by, int bz, int tid, float* = nullptr) {
        const int qc = tid % QC, role = tid / QC;
        bool fc;
        const int gj = fold_quad((TWI * bx - HC) / 2 + qc, a.cols / 2, fc);
        th.fc = fc ? 1 : 0;
        const float* zimg = a.z + (int64_t)bz * a.rows * a.cols + 2 * gj;
        const float* yb = a.yh + 2 * ((int64_t)bz * a.zs_n + gj);
        const int sz = 8 * a.cols, sb = 8 * (int)a.zs_row;
        const float* f[4];
        if (role == 0) {
            f[0] = zimg; f[1] = zimg + a.cols; f[2] = yb; f[3] = yb + 2 * 5 * a.zs_band;
            th.stride[0] = sz; th.stride[1] = sz; th.stride[2] = sb; th.stride[3] = sb;
        } else {
            f[0] = yb + 2 * 2 * a.zs_band; f[1] = yb + 2 * 3 * a.zs_band;
            f[2] = yb + 2 * 1 * a.zs_band; f[3] = yb + 2 * 4 * a.zs_band;
            th.stride[0] = sb; th.stride[1] = sb; th.stride[2] = sb; th.stride[3] = sb;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) th.ptr[i] = reinterpret_cast<const char*>(f[i]);
#pragma unroll
        for (int i = 0; i < RING; ++i) th.acc[i] = zero2();
        const int q0 = quad_base(a, by, 0);
#pragma unroll
        for (int s = 0; s < NST; ++s) load_stage<true>(a, th, th.st[s], q0 + s);
    }

    static DTCWT_D void c2q_rows(const F2 w0, const F2 w1, float g0, float g1, F2& top, F2& bot) {
        const float r0 = w0.x * g0, i0 = w0.y * g0;
        top.x = fmaf(w1.x, g1, r0); top.y = fmaf(w1.y, g1, i0);
        bot.x = fmaf(-w1.y, g1, i0); bot.y = fmaf(w1.x, g1, -r0);
    }
    static DTCWT_D void flip_quad(bool fr, bool fc, F2& top, F2& bot) {
        if (fc) { float t; t = top.x; top.x = top.y; top.y = t; t = bot.x; bot.x = bot.y; bot.y = t; }
        if (fr) { const F2 t = top; top = bot; bot = t; }
    }

    // Input row j of the period (static) of an image filtered with spec G goes to out[2 (j - b(ph)) - 4k + ph] for the
    // phases ph whose b(ph) has the parity of j; FIRST: the k = 0 term is the first contribution of its output row.
    template <class G, class TS, bool FIRST>
    static DTCWT_D void scatter(const int j, const F2 v, const PhaseTaps& t, F2 (&acc)[RING]) {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            if (((j - G::b(ph)) & 1) == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int slot = pmod(2 * (j - G::b(ph)) - 4 * k + ph, RING);
                    if (FIRST && k == 0) acc[slot] = fma2(TS::get(t, ph, 0), v, zero2());
                    else acc[slot] = fma2(TS::get(t, ph, k), v, acc[slot]);
                }
            }
        }
    }

    template <int ROLE, bool EDGE>
    static DTCWT_D void cols_role(const Args& a, Thread& th, float* sm, int by, int p, int qc) {
        const int qb = quad_base(a, by, p);
        const bool emit = p > 0;
        const float ga0 = a.gain[ROLE == 0 ? 0 : 2], ga1 = a.gain[ROLE == 0 ? 5 : 3];
        const float gb0 = a.gain[1], gb1 = a.gain[4];
        float* y = sm + ROLE * (RING * CYP) + 2 * qc;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const Raw cur = th.st[u % NST];
            load_stage<EDGE>(a, th, th.st[u % NST], qb + u + NST);
            F2 at, ab, bt, bb;
            if (ROLE == 0) {
                at = cur.v[0]; ab = cur.v[1];
                c2q_rows(cur.v[2], cur.v[3], ga0, ga1, bt, bb);
            } else {
                c2q_rows(cur.v[0], cur.v[1], ga0, ga1, at, ab);
                c2q_rows(cur.v[2], cur.v[3], gb0, gb1, bt, bb);
            }
            if (EDGE) {
                bool fr;
                fold_quad(qb + u, a.rows / 2, fr);
                flip_quad(fr, th.fc != 0, at, ab);
                flip_quad(fr, th.fc != 0, bt, bb);
            }
            // ring coordinates: input row j of the period feeds the output rows 2 (j + HGOFF) ..., see scatter()
            scatter<G0, T0, true>(2 * u, at, a.g0, th.acc);       // even row: first touch of phases 0, 2
            scatter<G1, T1, true>(2 * u, bt, a.g1, th.acc);       //           first touch of phases 1, 3
            scatter<G0, T0, false>(2 * u + 1, ab, a.g0, th.acc);
            scatter<G1, T1, false>(2 * u + 1, bb, a.g1, th.acc);
            if (emit) {                // the group first touched K - 1 quad rows ago is complete: rows 4u .. 4u+3 of the block
#pragma unroll
                for (int ph = 0; ph < 4; ++ph)
                    *reinterpret_cast<F2*>(y + (4 * u + ph) * CYP) = th.acc[pmod(4 * (u + HG - (K - 1)) + ph, RING)];
            }
        }
    }

    static DTCWT_D void cols(const Args& a, Thread& th, float* sm, int bx, int by, int bz, int tid, int p) {
        const int qc = tid % QC, role = tid / QC;
        const bool edge = edge_period(a, bx, by, p);
        if (role == 0) {
            if (edge) cols_role<0, true>(a, th, sm, by, p, qc);
            else cols_role<0, false>(a, th, sm, by, p, qc);
        } else {
            if (edge) cols_role<1, true>(a, th, sm, by, p, qc);
            else cols_role<1, false>(a, th, sm, by, p, qc);
        }
    }

    // row pass of period p (p > 0): out = H:g0(y1) + H:g1(y2) on the RING rows just finished; a scalar sample times the
    // tap pair (t[ph][k], t[ph+2][k]) advances output phases ph and ph+2 of a group in one FFMA2
    static DTCWT_D void rows(const Args& a, float* sm, int bx, int by, int bz, int tid, int p) {
        const float* y1 = sm;
        const float* y2 = sm + RING * CYP;
        float* img = a.out + (int64_t)bz * a.out_rows * a.out_cols;
        const int r0 = run_rows(a) * by + RING * (p - 1) - a.crop_r;
        int rp = (tid >> 1) / NSEG, seg = (tid >> 1) % NSEG;
#pragma unroll 1
        for (int task = tid; task < RING * NSEG; task += kThreads, rp += (kThreads / 2) / NSEG, seg += (kThreads / 2) % NSEG) {
            if (seg >= NSEG) { seg -= NSEG; ++rp; }
            const int yr = 2 * rp + (tid & 1);
            const int r = r0 + yr;
            const int c0 = 2 * (TWI * bx + 2 * NGR * seg) - a.crop_c;       // first output column of the task
            if (r < 0 || r >= a.out_rows || c0 >= a.out_cols) continue;
            F2 pe[NGR], po[NGR];                 // outputs (0, 2) and (1, 3) of each group
#pragma unroll
            for (int g = 0; g < NGR; ++g) { pe[g] = zero2(); po[g] = zero2(); }
            const F4* s1 = reinterpret_cast<const F4*>(y1 + yr * CYP + 2 * NGR * seg);
            const F4* s2 = reinterpret_cast<const F4*>(y2 + yr * CYP + 2 * NGR * seg);
#pragma unroll
            for (int c = 0; c < WN / 4; ++c) {
                const F4 v1 = s1[c], v2 = s2[c];
                const float w1[4] = {v1.x, v1.y, v1.z, v1.w}, w2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int g = 0; g < NGR; ++g) {
                        const int num = 4 * c + i - 2 * g;
                        if (num >= 0 && (num & 1) == 0 && num / 2 < K) {
                            pe[g] = fma2(w1[i], a.q[0].p[num / 2], pe[g]);
                            po[g] = fma2(w2[i], a.q[3].p[num / 2], po[g]);
                        }
                        if (num >= 1 && (num & 1) == 1 && (num - 1) / 2 < K) {
                            pe[g] = fma2(w2[i], a.q[1].p[(num - 1) / 2], pe[g]);
                            po[g] = fma2(w1[i], a.q[2].p[(num - 1) / 2], po[g]);
                        }
                    }
                }
            }
            float o[4 * NGR];
#pragma unroll
            for (int g = 0; g < NGR; ++g) { o[4 * g] = pe[g].x; o[4 * g + 1] = po[g].x; o[4 * g + 2] = pe[g].y; o[4 * g + 3] = po[g].y; }
            float* d = img + (int64_t)r * a.out_cols + c0;
            if (a.out_vec4 && c0 + 4 * NGR <= a.out_cols) {
#pragma unroll
                for (int c = 0; c < NGR; ++c) {
                    F4 v;
                    v.x = o[4 * c]; v.y = o[4 * c + 1]; v.z = o[4 * c + 2]; v.w = o[4 * c + 3];
                    reinterpret_cast<F4*>(d)[c] = v;
                }
            } else if (a.crop_c == 0) {
#pragma unroll
                for (int i = 0; i < 4 * NGR; i += 2)
                    if (c0 + i < a.out_cols) {
                        F2 v;
                        v.x = o[i]; v.y = o[i + 1];
                        *reinterpret_cast<F2*>(d + i) = v;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4 * NGR; ++i)
                    if (c0 + i >= 0 && c0 + i < a.out_cols) d[i] = o[i];
            }
        }
    }
};

// =============================================================================== forward level 1
struct FwdS1Args {
    const float* x;                 // [n][rows][cols]
    float* lolo;                    // [n][Lr][Lc]
    float* yh;                      // complex planar sub-bands [n][6][Lr/2][Lc/2] (strides below, complex elements)
    int n, rows, cols;              // stored size
    int Lr, Lc;                     // logical size: odd sizes repeat the last row / column (transform2d.py:86-94)
    int periods;                    // emitting periods per run
    int use_tma;
    int64_t zs_n, zs_band, zs_row;
    PairTab ph0, ph1s;              // row pass: h0, h1/sqrt2
    ColTaps v0, v1s, v1;            // column pass: h0, h1/sqrt2 (on A), h1 (on B, which already carries 1/sqrt2)
};

template <int K0, int K1, uint32_t M0, uint32_t M1, int RING_, int NT_, int MINB_, class T0 = ArgTaps, class T1S = ArgTaps, class T1 = ArgTaps>
struct FwdS1 {
    typedef FwdS1Args Args;
    static constexpr int RING = RING_, PER = RING_ / 2;
    static constexpr int C0 = (K0 - 1) / 2, C1 = (K1 - 1) / 2, CH = cmax(C0, C1), CQ = round_up(CH, 2);
    static constexpr int kThreads = NT_;
    static constexpr int NCP = kThreads / 2;                    // column pairs of a strip (2 roles)
    static constexpr int TW = 2 * NCP;                          // output columns of a strip
    static constexpr int NSEG = TW / 4;                         // row task = 4 output columns of one row
    static constexpr int HLA = round_up(CH, 4);                 // staged columns left of the strip (TMA boxes start 16-byte aligned)
    static constexpr int NWX = round_up(HLA + 4 + CH, 4);       // register window of a row task
    static constexpr int CXS = round_up(4 * (NSEG - 1) + NWX, 32);   // pitch of the staged input rows (128-byte rows for TMA)
    static constexpr int PA = TW + 4;                           // pitch of A / B
    static constexpr int XBUF = RING * CXS;                     // floats per input buffer (two of them)
    static constexpr int kSmemFloats = 2 * XBUF + 2 * RING * PA;
    static constexpr int kMinBlocks = MINB_;
    static constexpr int kMaxRegs = (65536 / (MINB_ * NT_)) / 8 * 8;     // registers per thread that still let MINB CTAs share an SM
    static_assert((K0 & 1) && (K1 & 1) && K0 <= kStreamMaxTaps && K1 <= kStreamMaxTaps && (M0 & 1u) && (M1 & 1u), "filters");
    static_assert(RING >= CQ + CH + 1 && (RING % 2) == 0 && (RING * NSEG) % kThreads == 0, "ring");
    static_assert(CXS <= 256 && (NCP % 32) == 0, "TMA box width; roles are whole warps");

    struct Thread { F2 lo[RING], hi[RING]; };                  // the two vertical filters of this thread's image

    static DTCWT_HD int run_rows(const Args& a) { return RING * a.periods; }
    static DTCWT_HD int tiles_c(const Args& a) { return (a.Lc + TW - 1) / TW; }
    static DTCWT_HD int tiles_r(const Args& a) { return (a.Lr + run_rows(a) - 1) / run_rows(a); }
    static DTCWT_HD int run_periods(const Args& a, int by) {
        const int left = a.Lr - run_rows(a) * by;
        const int e = (left + RING - 1) / RING;
        return 1 + (e < a.periods ? e : a.periods);
    }
    // logical input row consumed first by period p, logical column of element 0 of a staged row
    static DTCWT_HD int row_base(const Args& a, int by, int p) { return run_rows(a) * by - RING + CQ + RING * p; }
    static DTCWT_HD int col_base(int bx) { return TW * bx - HLA; }
    static DTCWT_HD int stored_row(const Args& a, int L) { return unpad(reflect_any(L, a.Lr), 0, a.rows); }
    static DTCWT_HD int stored_col(const Args& a, int L) { return unpad(reflect_any(L, a.Lc), 0, a.cols); }
    static DTCWT_HD bool rows_inside(const Args& a, int by, int p) {
        const int rb = row_base(a, by, p);
        return rb >= 0 && rb + RING <= a.rows;
    }
    static DTCWT_HD bool cols_inside(const Args& a, int bx) { return col_base(bx) >= 0 && col_base(bx) + CXS <= a.cols; }

    static DTCWT_D void init(Thread& th) {
#pragma unroll
        for (int i = 0; i < RING; ++i) { th.lo[i] = zero2(); th.hi[i] = zero2(); }
    }

    // staging without TMA (row pitch not a multiple of 16 bytes, and the host emulator): symmetric extension
    // (utils.py:136-153) resolved per element
    static DTCWT_D void load_plain(const Args& a, float* sm, int bx, int by, int bz, int p, int tid) {
        float* XS = sm + (p & 1) * XBUF;
        const float* img = a.x + (int64_t)bz * a.rows * a.cols;
        const int rb = row_base(a, by, p), cb = col_base(bx);
        for (int e = tid; e < XBUF; e += kThreads) {
            const int lr = e / CXS, lc = e - lr * CXS;
            XS[e] = img[(int64_t)stored_row(a, rb + lr) * a.cols + stored_col(a, cb + lc)];
        }
    }
    // TMA staging delivers zeros left and right of the image: mirror those columns inside shared memory
    static DTCWT_D void patch_cols(const Args& a, float* sm, int bx, int p, int tid) {
        float* XS = sm + (p & 1) * XBUF;
        const int cb = col_base(bx);
        for (int e = tid; e < XBUF; e += kThreads) {
            const int lr = e / CXS, lc = e - lr * CXS;
            const int L = cb + lc;
            if (L < 0 || L >= a.cols) {
                const int src = stored_col(a, L) - cb;
                if (src >= 0 && src < CXS) XS[e] = XS[lr * CXS + src];
            }
        }
    }

    // row pass of period p: A = H:h0(X), B = H:h1(X)/sqrt2 for the RING staged rows
    static DTCWT_D void rows(const Args& a, float* sm, int p, int tid) {
        const float* XS = sm + (p & 1) * XBUF;
        float* As = sm + 2 * XBUF;
        float* Bs = As + RING * PA;
        // kThreads / NSEG whole rows per round: a thread keeps its segment and steps down the rows
        static_assert(kThreads % NSEG == 0, "row pass steps whole rows");
        const int seg = tid % NSEG;
#pragma unroll 1
        for (int lr = tid / NSEG; lr < RING; lr += kThreads / NSEG) {
            F2 oa[2], ob[2];
            oa[0] = zero2(); oa[1] = zero2(); ob[0] = zero2(); ob[1] = zero2();
            const F4* src = reinterpret_cast<const F4*>(XS + lr * CXS + 4 * seg);
#pragma unroll
            for (int c = 0; c < NWX / 4; ++c) {
                const F4 v = src[c];
                pair_gather4<K0, M0, C0, -HLA, 4, NWX>(4 * c, v, a.ph0, oa);
                pair_gather4<K1, M1, C1, -HLA, 4, NWX>(4 * c, v, a.ph1s, ob);
            }
            F4 va, vb;
            va.x = oa[0].x; va.y = oa[0].y; va.z = oa[1].x; va.w = oa[1].y;
            vb.x = ob[0].x; vb.y = ob[0].y; vb.z = ob[1].x; vb.w = ob[1].y;
            *reinterpret_cast<F4*>(As + lr * PA + 4 * seg) = va;
            *reinterpret_cast<F4*>(Bs + lr * PA + 4 * seg) = vb;
        }
    }

    // q2c (transform2d.py:301-322; the 1/sqrt2 is already in the taps) of the quad (e0 / e1) -> bands b0, b1
    static DTCWT_D void store_q2c(const F2 e0, const F2 e1, float* z0, float* z1) {
        F2 w0, w1;
        w0.x = e0.x - e1.y; w0.y = e0.y + e1.x;
        w1.x = e0.x + e1.y; w1.y = e0.y - e1.x;
        *reinterpret_cast<F2*>(z0) = w0;
        *reinterpret_cast<F2*>(z1) = w1;
    }

    // ROLE 0 streams A: LoLo = V:h0(A) and q2c(V:h1(A)/sqrt2) -> bands 0,5.
    // ROLE 1 streams B: q2c(V:h0(B)) -> bands 2,3 and q2c(V:h1(B)) -> bands 1,4.
    template <int ROLE, bool FULL>
    static DTCWT_D void cols_role(const Args& a, Thread& th, const float* sm, int bx, int by, int bz, int p, int cp) {
        typedef typename std::conditional<ROLE == 0, T1S, T1>::type THI;
        const ColTaps& thi = (ROLE == 0) ? a.v1s : a.v1;
        const float* src = sm + 2 * XBUF + ROLE * (RING * PA) + 2 * cp;
        const int r0 = run_rows(a) * by + RING * (p - 1);          // first output row of this period's block
        const int c = TW * bx + 2 * cp;                            // first of the two output columns
        const int rows_ok = (p > 0 && c < a.Lc) ? a.Lr - r0 : 0;   // rows 2u, 2u+1 of the block are stored while 2u < rows_ok (Lr is even)
        float* lo = a.lolo + ((int64_t)bz * a.Lr + r0) * a.Lc + c;
        float* zb = a.yh + 2 * ((int64_t)bz * a.zs_n + (int64_t)(r0 / 2) * a.zs_row + c / 2);
        const int64_t bs = 2 * a.zs_band;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const F2 top = *reinterpret_cast<const F2*>(src + (2 * u) * PA);
            const F2 bot = *reinterpret_cast<const F2*>(src + (2 * u + 1) * PA);
            ring_scatter<T0, K0, M0, C0, true, RING>(2 * u, top, a.v0, th.lo);
            ring_scatter<THI, K1, M1, C1, true, RING>(2 * u, top, thi, th.hi);
            if (FULL || 2 * u < rows_ok) {
                const int s0 = pmod(2 * u - CQ, RING), s1 = pmod(2 * u + 1 - CQ, RING);
                float* z = zb + 2 * (int64_t)u * a.zs_row;
                if (ROLE == 0) {
                    *reinterpret_cast<F2*>(lo + (int64_t)(2 * u) * a.Lc) = th.lo[s0];
                    *reinterpret_cast<F2*>(lo + (int64_t)(2 * u + 1) * a.Lc) = th.lo[s1];
                    store_q2c(th.hi[s0], th.hi[s1], z, z + 5 * bs);
                } else {
                    store_q2c(th.lo[s0], th.lo[s1], z + 2 * bs, z + 3 * bs);
                    store_q2c(th.hi[s0], th.hi[s1], z + 1 * bs, z + 4 * bs);
                }
            }
            ring_scatter<T0, K0, M0, C0, true, RING>(2 * u + 1, bot, a.v0, th.lo);
            ring_scatter<THI, K1, M1, C1, true, RING>(2 * u + 1, bot, thi, th.hi);
        }
    }

    static DTCWT_D void cols(const Args& a, Thread& th, const float* sm, int bx, int by, int bz, int tid, int p) {
        const int cp = tid % NCP, role = tid / NCP;                // role is uniform within a warp
        // whole block of RING rows inside the image and this column pair inside too: no per-row store guards
        const bool full = p > 0 && TW * bx + 2 * cp < a.Lc && run_rows(a) * by + RING * p <= a.Lr;
        if (role == 0) {
            if (full) cols_role<0, true>(a, th, sm, bx, by, bz, p, cp);
            else cols_role<0, false>(a, th, sm, bx, by, bz, p, cp);
        } else {
            if (full) cols_role<1, true>(a, th, sm, bx, by, bz, p, cp);
            else cols_role<1, false>(a, th, sm, bx, by, bz, p, cp);
        }
    }
};

}  // namespace dtcwt
