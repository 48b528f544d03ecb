// Fused per-level 3-D DT-CWT kernels (float32): the DEPTH pass with the 2x2x2 packers in registers.
//
// A 3-D level of the reference (transform3d.py:208-383 forward, :385-526 inverse) filters a volume
// [D0][D1][D2] along axis 2, 1 and 0 with a (lowpass, highpass) pair and packs seven of the eight
// octants into 4 complex channels each (cube2c :532-579, c2cube :581-619).  Here a level is two launches:
//
//   forward   slices   Fwd2d<..., kFwdRaw>  (fused2d.cuh): axes 2 and 1 of every slice [D1][D2] in one tile kernel,
//                      four real images s0..s3 per slice  ->  scratch [4][n*D0][D1'][D2']
//             depth    Z3Fwd (this file): axis 0 of the four scratch volumes, both filters, cube2c in registers
//                      ->  LLL and the 28 complex channels straight from registers
//   inverse   depth    Z3Inv: c2cube in registers while the channels are loaded, axis 0 with both filters
//                      ->  scratch [4][n*D0''][a1][a2]
//             slices   Inv2d<..., RAW>: axes 1 and 2
//
// (The passes commute; only rounding differs from the reference's 2, 1, 0 / 1, 0, 2 order.)
// Per level every sample is read twice and written twice: 16 B per input voxel instead of the 14 round
// trips of the one-axis-per-launch composition it replaces.
//
// One depth thread owns a 2 x 2 patch of (axis 1, axis 2) -- two rows, one float2 each -- and NG groups
// along axis 0, so every 2x2x2 octet of the packers lives in its registers.  Lanes run along axis 2
// (8-byte coalesced loads / stores).  The filter is the same polyphase scatter as the 2-D kernels
// (fir_scatter, packed FFMA2).  The packers' 1/2 is folded into the taps; the one real output (LLL)
// and the one real input (Yl) are scaled by 2 instead, which is exact in binary floating point.
#pragma once
#include "fused2d.cuh"

namespace dtcwt {

// channel block (4 channels) of the octant with filter types (t0, t1, t2) along axes (0, 1, 2), 1 = highpass:
// order HLL LHL HHL LLH HLH LHH HHH = (0,1,0) (1,0,0) (1,1,0) (0,0,1) (0,1,1) (1,0,1) (1,1,1)  (transform3d.py:280-288)
DTCWT_HD constexpr int octant_block(int t0, int t1, int t2) { return t1 + 2 * t0 + 4 * t2 - 1; }

struct Z3Args {
    const float* s;                 // forward: scratch in; inverse: lowpass Yl [n][a0][a1][a2]
    float* lll;                     // forward: LLL out [n][L0 * P / Q][h][w]; inverse: scratch out
    float* yh;                      // complex planar channels (strides below); forward: out, inverse: in
    int n;
    int d0, pad0, L0;               // forward: stored slices, replicated slices before them, logical length
    int h, w;                       // size of one (axis 1, axis 2) image of the depth pass (even)
    int out_d0, crop0;              // inverse: stored output slices and cropped slices per side (transform3d.py:505-524)
    int64_t sub_stride, vol_stride; // scratch: floats between the four images s0..s3 / between volumes
    int64_t zs_n, zs_chan, zs_0, zs_1, zs_2;     // complex strides of yh
    PhaseTaps lo, hi;               // taps of the pair, times 1/2
};

// ------------------------------------------------------------------------------- forward depth pass
template <class FLO, class FHI, int NG_>
struct Z3Fwd {
    typedef Z3Args Args;
    static constexpr int P = FLO::P, Q = FLO::Q, NG = NG_;
    static constexpr int HL = cmax(spec_lo<FLO>(), spec_lo<FHI>());
    static constexpr int HR = cmax(spec_hi<FLO>(), spec_hi<FHI>());
    static constexpr int NR = Q * NG + HL + HR;
    static constexpr int NOUT = P * NG;
    static_assert(P == FHI::P && Q == FHI::Q && (NOUT % 2) == 0, "filter pair");

    static DTCWT_HD int64_t groups(const Args& a) { return (a.L0 * P / Q + NOUT - 1) / NOUT; }
    // gid -> (x pair, y pair, depth group, image s, volume)
    static DTCWT_HD int64_t total(const Args& a) { return (int64_t)(a.w / 2) * (a.h / 2) * groups(a) * 4 * a.n; }

    template <bool INSIDE>
    static DTCWT_D void accumulate(const Args& a, const float* src, int64_t plane, int l0, int z0, F2 (&lo)[2][NOUT], F2 (&hi)[2][NOUT]) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int z = INSIDE ? z0 + j : unpad(reflect_any(l0 + j, a.L0), a.pad0, a.d0);
            const float* p = src + (int64_t)z * plane;
            const F2 v0 = *reinterpret_cast<const F2*>(p);
            const F2 v1 = *reinterpret_cast<const F2*>(p + a.w);
            fir_scatter<FLO, NG, HL>(j, v0, a.lo, lo[0]);
            fir_scatter<FLO, NG, HL>(j, v1, a.lo, lo[1]);
            fir_scatter<FHI, NG, HL>(j, v0, a.hi, hi[0]);
            fir_scatter<FHI, NG, HL>(j, v1, a.hi, hi[1]);
        }
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
        const int wp = a.w / 2, hp = a.h / 2;
        const int xp = (int)(gid % wp);
        int64_t r = gid / wp;
        const int yp = (int)(r % hp);
        r /= hp;
        const int64_t ng = groups(a);
        const int gz = (int)(r % ng);
        r /= ng;
        const int sub = (int)(r % 4);
        const int b = (int)(r / 4);
        const float* src = a.s + sub * a.sub_stride + b * a.vol_stride + (int64_t)(2 * yp) * a.w + 2 * xp;
        const int64_t plane = (int64_t)a.h * a.w;
        F2 lo[2][NOUT], hi[2][NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) { lo[0][i] = zero2(); lo[1][i] = zero2(); hi[0][i] = zero2(); hi[1][i] = zero2(); }
        const int l0 = Q * NG * gz - HL;
        const int z0 = l0 - a.pad0;
        // no symmetric extension needed when the window lies inside the stored slices: skip reflect_any's modulo (two copies
        // of the unrolled loop behind one warp-uniform branch)
        if (z0 >= 0 && z0 + NR <= a.d0) accumulate<true>(a, src, plane, l0, z0, lo, hi);
        else accumulate<false>(a, src, plane, l0, z0, lo, hi);
        const int t1 = sub & 1, t2 = sub >> 1;                   // filter types of this image along axes 1 and 2
        const int Lout = a.L0 * P / Q;
        const int zo = NOUT * gz;                                 // first output slice
        if (sub == 0) {                                           // LLL: real, undo the folded 1/2
            float* d = a.lll + ((int64_t)b * Lout + zo) * plane + (int64_t)(2 * yp) * a.w + 2 * xp;
#pragma unroll
            for (int i = 0; i < NOUT; ++i) {
                if (zo + i < Lout) {
                    F2 u0, u1;
                    u0.x = 2.f * lo[0][i].x; u0.y = 2.f * lo[0][i].y;
                    u1.x = 2.f * lo[1][i].x; u1.y = 2.f * lo[1][i].y;
                    *reinterpret_cast<F2*>(d + (int64_t)i * plane) = u0;
                    *reinterpret_cast<F2*>(d + (int64_t)i * plane + a.w) = u1;
                }
            }
        } else {
            pack(a, lo, octant_block(0, t1, t2), b, zo, yp, xp, Lout);
        }
        pack(a, hi, octant_block(1, t1, t2), b, zo, yp, xp, Lout);
    }

    // cube2c (transform3d.py:532-579; the 1/2 is in the taps): octet corners by (axis 0, axis 1, axis 2) parity
    //   A=(0,0,0) B=(0,1,0) C=(1,0,0) D=(1,1,0) E=(0,0,1) F=(0,1,1) G=(1,0,1) H=(1,1,1)
    static DTCWT_D void pack(const Args& a, const F2 (&y)[2][NOUT], int block, int b, int zo, int yp, int xp, int Lout) {
        float* z = a.yh + 2 * ((int64_t)b * a.zs_n + (int64_t)(4 * block) * a.zs_chan + (int64_t)(zo / 2) * a.zs_0 +
                               (int64_t)yp * a.zs_1 + (int64_t)xp * a.zs_2);
        const int64_t cs = 2 * a.zs_chan;
#pragma unroll
        for (int q = 0; q < NOUT / 2; ++q) {
            if (zo + 2 * q < Lout) {
                // with X = (A, E), Y = (-G, C), Z = (D, H), W = (-F, B) as float pairs:
                //   p = X + Y - Z + W   q = X + Y + Z - W   r = X - Y + Z + W   s = X - Y - Z - W
                // (eight packed additions instead of 24 scalar ones; they associate differently, which moves the last bit)
                const F2 X = y[0][2 * q], Z = y[1][2 * q + 1];
                F2 Y, W;
                Y.x = -y[0][2 * q + 1].y; Y.y = y[0][2 * q + 1].x;
                W.x = -y[1][2 * q].y; W.y = y[1][2 * q].x;
                const F2 u1 = add2(X, Y), u2 = fma2(-1.f, Y, X), v1 = fma2(-1.f, W, Z), v2 = add2(Z, W);
                float* zz = z + 2 * (int64_t)q * a.zs_0;
                *reinterpret_cast<F2*>(zz) = fma2(-1.f, v1, u1);                  // p
                *reinterpret_cast<F2*>(zz + cs) = add2(u1, v1);                   // q
                *reinterpret_cast<F2*>(zz + 2 * cs) = add2(u2, v2);               // r
                *reinterpret_cast<F2*>(zz + 3 * cs) = fma2(-1.f, v2, u2);         // s
            }
        }
    }
};

// ------------------------------------------------------------------------------- inverse depth pass
// scratch image s (axis-1 type s & 1, axis-2 type s >> 1) = G0 along axis 0 of the lowpass-in-depth octant
//                                                          + G1 along axis 0 of the highpass-in-depth octant
template <class GLO, class GHI, int NG_>
struct Z3Inv {
    typedef Z3Args Args;
    static constexpr int P = GLO::P, Q = GLO::Q, NG = NG_;
    static constexpr int HL = round_up(cmax(spec_lo<GLO>(), spec_lo<GHI>()), 2);     // whole octets
    static constexpr int HR = round_up(cmax(spec_hi<GLO>(), spec_hi<GHI>()), 2);
    static constexpr int NR = Q * NG + HL + HR;
    static constexpr int NOUT = P * NG;
    static_assert(P == GHI::P && Q == GHI::Q && ((Q * NG) % 2) == 0, "filter pair");

    // d0 = slices of the lowpass / real octants (even)
    static DTCWT_HD int64_t groups(const Args& a) { return (a.d0 + Q * NG - 1) / (Q * NG); }
    static DTCWT_HD int64_t total(const Args& a) { return (int64_t)(a.w / 2) * (a.h / 2) * groups(a) * 4 * a.n; }

    struct Oct { F2 v[2][2]; };      // [axis-0 parity][axis-1 parity], .x / .y = axis-2 parity

    // c2cube (transform3d.py:581-619; the 1/2 is in the taps)
    static DTCWT_D void unpack(const Args& a, int block, int b, int oz, int yp, int xp, Oct& o) {
        const float* z = a.yh + 2 * ((int64_t)b * a.zs_n + (int64_t)(4 * block) * a.zs_chan + (int64_t)oz * a.zs_0 +
                                     (int64_t)yp * a.zs_1 + (int64_t)xp * a.zs_2);
        const int64_t cs = 2 * a.zs_chan;
        const F2 p = *reinterpret_cast<const F2*>(z), q = *reinterpret_cast<const F2*>(z + cs);
        const F2 r = *reinterpret_cast<const F2*>(z + 2 * cs), s = *reinterpret_cast<const F2*>(z + 3 * cs);
        o.v[0][0].x = p.x + q.x + r.x + s.x;        // A
        o.v[0][0].y = p.y + q.y + r.y + s.y;        // E
        o.v[0][1].x = p.y - q.y + r.y - s.y;        // B
        o.v[0][1].y = -p.x + q.x - r.x + s.x;       // F
        o.v[1][0].x = p.y + q.y - r.y - s.y;        // C
        o.v[1][0].y = -p.x - q.x + r.x + s.x;       // G
        o.v[1][1].x = -p.x + q.x + r.x - s.x;       // D
        o.v[1][1].y = -p.y + q.y + r.y - s.y;       // H
    }

    // the octets (lowpass-in-depth, highpass-in-depth) of logical octet index o, mirrored into the volume; flip: the mirror
    // exchanges the two axis-0 parities
    static DTCWT_D void fetch(const Args& a, int sub, int b, int yp, int xp, int o, int noct, int64_t plane, int t1, int t2,
                              Oct& lo, Oct& hi, bool& flip) {
        int oz = o;
        flip = false;
        if (oz < 0) { oz = -1 - oz; flip = true; } else if (oz >= noct) { oz = 2 * noct - 1 - oz; flip = true; }
        oz = oz < 0 ? 0 : (oz >= noct ? noct - 1 : oz);         // further out only feeds outputs that are never stored
        if (sub == 0) {
            const float* p = a.s + ((int64_t)b * a.d0 + 2 * oz) * plane + (int64_t)(2 * yp) * a.w + 2 * xp;
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const F2 v = *reinterpret_cast<const F2*>(p + (int64_t)e * plane + (int64_t)f * a.w);
                    lo.v[e][f].x = 2.f * v.x; lo.v[e][f].y = 2.f * v.y;
                }
        } else {
            unpack(a, octant_block(0, t1, t2), b, oz, yp, xp, lo);
        }
        unpack(a, octant_block(1, t1, t2), b, oz, yp, xp, hi);
    }

    // (Variants of this loop that did not pay, profiles/r2_02: a second copy without the mirror logic for interior windows,
    // 1.10 vs 0.99 ms per step; loading octet jq + 1 before scattering octet jq, 96 registers / 2 CTAs per SM, 1.03 ms.)
    static DTCWT_D void accumulate(const Args& a, F2 (&acc)[2][NOUT], int sub, int b, int yp, int xp, int o0, int noct, int64_t plane,
                                   int t1, int t2) {
#pragma unroll
        for (int jq = 0; jq < NR / 2; ++jq) {
            Oct lo, hi;
            bool flip;
            fetch(a, sub, b, yp, xp, o0 + jq, noct, plane, t1, t2, lo, hi, flip);
            if (flip) {                                          // selects, not indexed: the octets stay in registers
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    F2 t = lo.v[0][f]; lo.v[0][f] = lo.v[1][f]; lo.v[1][f] = t;
                    t = hi.v[0][f]; hi.v[0][f] = hi.v[1][f]; hi.v[1][f] = t;
                }
            }
            fir_scatter<GLO, NG, HL>(2 * jq, lo.v[0][0], a.lo, acc[0]);
            fir_scatter<GLO, NG, HL>(2 * jq, lo.v[0][1], a.lo, acc[1]);
            fir_scatter<GHI, NG, HL>(2 * jq, hi.v[0][0], a.hi, acc[0]);
            fir_scatter<GHI, NG, HL>(2 * jq, hi.v[0][1], a.hi, acc[1]);
            fir_scatter<GLO, NG, HL>(2 * jq + 1, lo.v[1][0], a.lo, acc[0]);
            fir_scatter<GLO, NG, HL>(2 * jq + 1, lo.v[1][1], a.lo, acc[1]);
            fir_scatter<GHI, NG, HL>(2 * jq + 1, hi.v[1][0], a.hi, acc[0]);
            fir_scatter<GHI, NG, HL>(2 * jq + 1, hi.v[1][1], a.hi, acc[1]);
        }
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
        const int wp = a.w / 2, hp = a.h / 2;
        const int xp = (int)(gid % wp);
        int64_t r = gid / wp;
        const int yp = (int)(r % hp);
        r /= hp;
        const int64_t ng = groups(a);
        const int gz = (int)(r % ng);
        r /= ng;
        const int sub = (int)(r % 4);
        const int b = (int)(r / 4);
        const int t1 = sub & 1, t2 = sub >> 1;
        const int64_t plane = (int64_t)a.h * a.w;
        const int noct = a.d0 / 2;
        F2 acc[2][NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) { acc[0][i] = zero2(); acc[1][i] = zero2(); }
        const int o0 = (Q * NG * gz - HL) / 2;                   // first octet along axis 0 (may be negative)
        accumulate(a, acc, sub, b, yp, xp, o0, noct, plane, t1, t2);
        float* d = a.lll + sub * a.sub_stride + b * a.vol_stride + (int64_t)(2 * yp) * a.w + 2 * xp;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const int zo = NOUT * gz + i - a.crop0;
            if (zo >= 0 && zo < a.out_d0) {
                *reinterpret_cast<F2*>(d + (int64_t)zo * plane) = acc[0][i];
                *reinterpret_cast<F2*>(d + (int64_t)zo * plane + a.w) = acc[1][i];
            }
        }
    }
};

// ------------------------------------------------------------------------------- depth passes, one image ROW per thread
// Z3Fwd / Z3Inv give a thread both rows of its 2 x 2 patch, which leaves registers for only NG = 2 groups along axis 0: a
// thread then reads 34 slices to produce 8 (forward) or 8 octets to produce 2 (inverse) -- four times the compulsory traffic
// between L2 and the SM, and the launch is bound by exactly that (profiles/r3_01).  Here ADJACENT LANES take the two rows
// of a patch and each thread covers NG = 4 groups with the same accumulator count: 42 slices for 16 (10 octets for 4),
// 2.6 x.  The inverse needs nothing else (c2cube gives each lane its own row's corners from the same four complex loads,
// which the pair of lanes shares inside one request); the forward exchanges its outputs with the neighbouring lane
// (shuffles) for cube2c and each lane of a pair stores two of the four complex channels.
DTCWT_D F2 pair_exchange(const unsigned mask, const F2 v) {
#ifdef DTCWT_EMU
    return v;                                   // the emulator computes the partner row itself (Z3FwdS::run)
#else
    F2 r;
    r.x = __shfl_xor_sync(mask, v.x, 1);
    r.y = __shfl_xor_sync(mask, v.y, 1);
    return r;
#endif
}

template <class FLO, class FHI, int NG_>
struct Z3FwdS {
    typedef Z3Args Args;
    static constexpr int P = FLO::P, Q = FLO::Q, NG = NG_;
    static constexpr int HL = cmax(spec_lo<FLO>(), spec_lo<FHI>());
    static constexpr int HR = cmax(spec_hi<FLO>(), spec_hi<FHI>());
    static constexpr int NR = Q * NG + HL + HR;
    static constexpr int NOUT = P * NG;
    static_assert(P == FHI::P && Q == FHI::Q && (NOUT % 2) == 0, "filter pair");

    static DTCWT_HD int64_t groups(const Args& a) { return (a.L0 * P / Q + NOUT - 1) / NOUT; }
    // gid -> (row parity, x pair, y pair, depth group, image s, volume); an even total keeps lane pairs together
    static DTCWT_HD int64_t total(const Args& a) { return (int64_t)a.w * (a.h / 2) * groups(a) * 4 * a.n; }

    template <bool INSIDE>
    static DTCWT_D void accumulate(const Args& a, const float* src, int64_t plane, int l0, int z0, F2 (&lo)[NOUT], F2 (&hi)[NOUT]) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int z = INSIDE ? z0 + j : unpad(reflect_any(l0 + j, a.L0), a.pad0, a.d0);
            const F2 v = *reinterpret_cast<const F2*>(src + (int64_t)z * plane);
            fir_scatter<FLO, NG, HL>(j, v, a.lo, lo);
            fir_scatter<FHI, NG, HL>(j, v, a.hi, hi);
        }
    }
    static DTCWT_D void row_outputs(const Args& a, const float* src, int64_t plane, int gz, F2 (&lo)[NOUT], F2 (&hi)[NOUT]) {
#pragma unroll
        for (int i = 0; i < NOUT; ++i) { lo[i] = zero2(); hi[i] = zero2(); }
        const int l0 = Q * NG * gz - HL;
        const int z0 = l0 - a.pad0;
        if (z0 >= 0 && z0 + NR <= a.d0) accumulate<true>(a, src, plane, l0, z0, lo, hi);
        else accumulate<false>(a, src, plane, l0, z0, lo, hi);
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
        const int f = (int)(gid & 1);                             // row of the patch: adjacent lanes
        int64_t r = gid >> 1;
        const int wp = a.w / 2, hp = a.h / 2;
        const int xp = (int)(r % wp);
        r /= wp;
        const int yp = (int)(r % hp);
        r /= hp;
        const int64_t ng = groups(a);
        const int gz = (int)(r % ng);
        r /= ng;
        const int sub = (int)(r % 4);
        const int b = (int)(r / 4);
        const int64_t plane = (int64_t)a.h * a.w;
        const float* src = a.s + sub * a.sub_stride + b * a.vol_stride + (int64_t)(2 * yp + f) * a.w + 2 * xp;
        F2 lo[NOUT], hi[NOUT];
        row_outputs(a, src, plane, gz, lo, hi);
        const int t1 = sub & 1, t2 = sub >> 1;                   // filter types of this image along axes 1 and 2
        const int Lout = a.L0 * P / Q;
        const int zo = NOUT * gz;                                 // first output slice
        // lanes of this warp that run at all (the launch guards gid < total; total is even, so a lane and its partner are
        // either both in or both out): the exchange names exactly those
        const int64_t left = total(a) - (gid & ~(int64_t)31);
        const unsigned mask = left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u);
#ifdef DTCWT_EMU
        F2 plo[NOUT], phi[NOUT];                                  // the neighbouring lane's row, recomputed on the host
        row_outputs(a, src + (f ? -(int64_t)a.w : (int64_t)a.w), plane, gz, plo, phi);
#else
        const F2 (&plo)[NOUT] = lo;                               // placeholders: pack() fetches the partner's values itself
        const F2 (&phi)[NOUT] = hi;
#endif
        if (sub == 0) {                                           // LLL: real, undo the folded 1/2
            float* d = a.lll + ((int64_t)b * Lout + zo) * plane + (int64_t)(2 * yp + f) * a.w + 2 * xp;
#pragma unroll
            for (int i = 0; i < NOUT; ++i) {
                if (zo + i < Lout) {
                    F2 u;
                    u.x = 2.f * lo[i].x; u.y = 2.f * lo[i].y;
                    *reinterpret_cast<F2*>(d + (int64_t)i * plane) = u;
                }
            }
        }
        // every running lane goes through both exchanges (a warp may straddle two images `sub`); only the stores are predicated
        pack(a, mask, sub != 0, lo, plo, f, sub ? octant_block(0, t1, t2) : 0, b, zo, yp, xp, Lout);
        pack(a, mask, true, hi, phi, f, octant_block(1, t1, t2), b, zo, yp, xp, Lout);
    }

    // cube2c (transform3d.py:532-579; the 1/2 is in the taps) with the two rows of an octet in two lanes: y = this lane's
    // row (f), o = the other row (device: fetched from the neighbouring lane; emulator: passed in).  Lane f = 0 stores
    // channels p and q, lane f = 1 channels r and s.
    static DTCWT_D void pack(const Args& a, const unsigned mask, const bool store, const F2 (&y)[NOUT], const F2 (&o)[NOUT], int f,
                             int block, int b, int zo, int yp, int xp, int Lout) {
        float* z = a.yh + 2 * ((int64_t)b * a.zs_n + (int64_t)(4 * block + 2 * f) * a.zs_chan + (int64_t)(zo / 2) * a.zs_0 +
                               (int64_t)yp * a.zs_1 + (int64_t)xp * a.zs_2);
        const int64_t cs = 2 * a.zs_chan;
#pragma unroll
        for (int q = 0; q < NOUT / 2; ++q) {
#ifdef DTCWT_EMU
            const F2 o0 = o[2 * q], o1 = o[2 * q + 1];
#else
            const F2 o0 = pair_exchange(mask, y[2 * q]), o1 = pair_exchange(mask, y[2 * q + 1]);      // all running lanes take part
#endif
            const F2 r0z0 = f ? o0 : y[2 * q], r1z0 = f ? y[2 * q] : o0;
            const F2 r0z1 = f ? o1 : y[2 * q + 1], r1z1 = f ? y[2 * q + 1] : o1;
            const float A = r0z0.x, E = r0z0.y, B = r1z0.x, F = r1z0.y;
            const float C = r0z1.x, G = r0z1.y, D = r1z1.x, H = r1z1.y;
            if (store && zo + 2 * q < Lout) {
                float* zz = z + 2 * (int64_t)q * a.zs_0;
                F2 c0, c1;
                if (f == 0) {
                    c0.x = A - G - D - F; c0.y = B - H + C + E;                   // p
                    c1.x = A - G + D + F; c1.y = -B + H + C + E;                  // q
                } else {
                    c0.x = A + G + D - F; c0.y = B + H - C + E;                   // r
                    c1.x = A + G - D + F; c1.y = -B - H - C + E;                  // s
                }
                *reinterpret_cast<F2*>(zz) = c0;
                *reinterpret_cast<F2*>(zz + cs) = c1;
            }
        }
    }
};

template <class GLO, class GHI, int NG_>
struct Z3InvS {
    typedef Z3Args Args;
    static constexpr int P = GLO::P, Q = GLO::Q, NG = NG_;
    static constexpr int HL = round_up(cmax(spec_lo<GLO>(), spec_lo<GHI>()), 2);     // whole octets
    static constexpr int HR = round_up(cmax(spec_hi<GLO>(), spec_hi<GHI>()), 2);
    static constexpr int NR = Q * NG + HL + HR;
    static constexpr int NOUT = P * NG;
    static_assert(P == GHI::P && Q == GHI::Q && ((Q * NG) % 2) == 0, "filter pair");

    static DTCWT_HD int64_t groups(const Args& a) { return (a.d0 + Q * NG - 1) / (Q * NG); }
    static DTCWT_HD int64_t total(const Args& a) { return (int64_t)a.w * (a.h / 2) * groups(a) * 4 * a.n; }

    // row f of the octet: v[e] = axis-0 parity e, .x / .y = axis-2 parity   (c2cube, transform3d.py:581-619; 1/2 in the taps)
    static DTCWT_D void unpack_row(const Args& a, int f, int block, int b, int oz, int yp, int xp, F2 (&v)[2]) {
        const float* z = a.yh + 2 * ((int64_t)b * a.zs_n + (int64_t)(4 * block) * a.zs_chan + (int64_t)oz * a.zs_0 +
                                     (int64_t)yp * a.zs_1 + (int64_t)xp * a.zs_2);
        const int64_t cs = 2 * a.zs_chan;
        const F2 p = *reinterpret_cast<const F2*>(z), q = *reinterpret_cast<const F2*>(z + cs);
        const F2 r = *reinterpret_cast<const F2*>(z + 2 * cs), s = *reinterpret_cast<const F2*>(z + 3 * cs);
        if (f == 0) {
            v[0].x = p.x + q.x + r.x + s.x;         // A
            v[0].y = p.y + q.y + r.y + s.y;         // E
            v[1].x = p.y + q.y - r.y - s.y;         // C
            v[1].y = -p.x - q.x + r.x + s.x;        // G
        } else {
            v[0].x = p.y - q.y + r.y - s.y;         // B
            v[0].y = -p.x + q.x - r.x + s.x;        // F
            v[1].x = -p.x + q.x + r.x - s.x;        // D
            v[1].y = -p.y + q.y + r.y - s.y;        // H
        }
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
        const int f = (int)(gid & 1);
        int64_t r = gid >> 1;
        const int wp = a.w / 2, hp = a.h / 2;
        const int xp = (int)(r % wp);
        r /= wp;
        const int yp = (int)(r % hp);
        r /= hp;
        const int64_t ng = groups(a);
        const int gz = (int)(r % ng);
        r /= ng;
        const int sub = (int)(r % 4);
        const int b = (int)(r / 4);
        const int t1 = sub & 1, t2 = sub >> 1;
        const int64_t plane = (int64_t)a.h * a.w;
        const int noct = a.d0 / 2;
        F2 acc[NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) acc[i] = zero2();
        const int o0 = (Q * NG * gz - HL) / 2;                   // first octet along axis 0 (may be negative)
#pragma unroll
        for (int jq = 0; jq < NR / 2; ++jq) {
            int oz = o0 + jq;
            bool flip = false;                                   // a mirrored octet has its two axis-0 parities exchanged
            if (oz < 0) { oz = -1 - oz; flip = true; } else if (oz >= noct) { oz = 2 * noct - 1 - oz; flip = true; }
            oz = oz < 0 ? 0 : (oz >= noct ? noct - 1 : oz);      // further out only feeds outputs that are never stored
            F2 lo[2], hi[2];
            if (sub == 0) {
                const float* p = a.s + ((int64_t)b * a.d0 + 2 * oz) * plane + (int64_t)(2 * yp + f) * a.w + 2 * xp;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const F2 v = *reinterpret_cast<const F2*>(p + (int64_t)e * plane);
                    lo[e].x = 2.f * v.x; lo[e].y = 2.f * v.y;
                }
            } else {
                unpack_row(a, f, octant_block(0, t1, t2), b, oz, yp, xp, lo);
            }
            unpack_row(a, f, octant_block(1, t1, t2), b, oz, yp, xp, hi);
            if (flip) {
                F2 t = lo[0]; lo[0] = lo[1]; lo[1] = t;
                t = hi[0]; hi[0] = hi[1]; hi[1] = t;
            }
            fir_scatter<GLO, NG, HL>(2 * jq, lo[0], a.lo, acc);
            fir_scatter<GHI, NG, HL>(2 * jq, hi[0], a.hi, acc);
            fir_scatter<GLO, NG, HL>(2 * jq + 1, lo[1], a.lo, acc);
            fir_scatter<GHI, NG, HL>(2 * jq + 1, hi[1], a.hi, acc);
        }
        float* d = a.lll + sub * a.sub_stride + b * a.vol_stride + (int64_t)(2 * yp + f) * a.w + 2 * xp;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const int zo = NOUT * gz + i - a.crop0;
            if (zo >= 0 && zo < a.out_d0) *reinterpret_cast<F2*>(d + (int64_t)zo * plane) = acc[i];
        }
    }
};

// ------------------------------------------------------------------------------- inverse depth pass, staged loads
// Z3Inv is bound by the latency of its global loads (profiles/r2_04: long-scoreboard stalls, 35 % of the warps resident, a
// third of the DRAM rate): a thread has one octet (eight 8-byte loads) in flight, and fetching further ahead in registers
// costs the occupancy it was meant to replace.  Here every thread stages its next DEPTH octets in a PRIVATE slice of shared
// memory with asynchronous copies (cp.async, 8 bytes each; nobody else reads the slice, so there is no barrier anywhere --
// the thread waits on its own copy groups): DEPTH x 8 loads per thread in flight at the register count of Z3Inv.
template <class GLO, class GHI, int NG_, int DEPTH_>
struct Z3InvA {
    typedef Z3Inv<GLO, GHI, NG_> Base;
    typedef Z3Args Args;
    static constexpr int P = Base::P, Q = Base::Q, NG = NG_, DEPTH = DEPTH_;
    static constexpr int HL = Base::HL, HR = Base::HR, NR = Base::NR, NOUT = Base::NOUT, NQ = NR / 2;
    static constexpr int kThreads = 256;
    static_assert(DEPTH >= 1 && DEPTH <= NQ && DEPTH * 8 * kThreads * 8 <= 48 * 1024, "stages fit static shared memory");

    static DTCWT_HD int64_t groups(const Args& a) { return Base::groups(a); }
    static DTCWT_HD int64_t total(const Args& a) { return Base::total(a); }

    struct Src {                    // what does not change along the run: the eight streams of this thread
        const float* lo;            // sub == 0: lowpass volume at (row 2 yp, column 2 xp) of slice 0; else channel block of the lowpass-in-depth octant
        const float* hi;            // channel block of the highpass-in-depth octant
        int64_t lo_step, hi_step;   // floats per octet
        int64_t lo_i1, lo_i2;       // sub == 0: +row, +slice;  else: channel stride (lo_i2 unused)
        int64_t cs;
        bool real;
    };

    static DTCWT_D int fold(int o, int noct, bool& flip) {
        flip = false;
        if (o < 0) { o = -1 - o; flip = true; } else if (o >= noct) { o = 2 * noct - 1 - o; flip = true; }
        return o < 0 ? 0 : (o >= noct ? noct - 1 : o);          // further out only feeds outputs that are never stored
    }

    // the eight copies of octet o into stage slot `st` (element i of the stage at st[i * kThreads])
    static DTCWT_D void issue(const Src& s, F2* st, int o, int noct) {
        bool flip;
        const int oz = fold(o, noct, flip);
        const float* pl = s.lo + (int64_t)oz * s.lo_step;
        const float* ph = s.hi + (int64_t)oz * s.hi_step;
        if (s.real) {
            async_copy8(st + 0 * kThreads, pl);                            // slice 2 oz,     row 2 yp
            async_copy8(st + 1 * kThreads, pl + s.lo_i1);                  //                 row 2 yp + 1
            async_copy8(st + 2 * kThreads, pl + s.lo_i2);                  // slice 2 oz + 1, row 2 yp
            async_copy8(st + 3 * kThreads, pl + s.lo_i2 + s.lo_i1);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) async_copy8(st + i * kThreads, pl + i * s.cs);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) async_copy8(st + (4 + i) * kThreads, ph + i * s.cs);
    }

    // c2cube on complex values as float pairs: with u = p + q, v = r + s, w = q - p, z = s - r (packed adds, FADD2)
    //   (A, E) = u + v     (G, -C) = v - u     (D, H) = w - z     (F, -B) = w + z
    // eight packed additions and two sign flips instead of 24 scalar additions (the kernel is issue-bound once its loads are
    // staged); the sums associate differently from the scalar form, which only moves the last bit
    static DTCWT_D F2 sub2(const F2 a, const F2 b) { return fma2(-1.f, b, a); }      // a - b, exact: one FFMA2
    static DTCWT_D void unpack_vals(const F2 p, const F2 q, const F2 r, const F2 s, typename Base::Oct& o) {
        const F2 u = add2(p, q), v = add2(r, s), w = sub2(q, p), z = sub2(s, r);
        const F2 ae = add2(u, v), gc = sub2(v, u), dh = sub2(w, z), fb = add2(w, z);
        o.v[0][0] = ae;                              // A, E
        o.v[0][1].x = -fb.y; o.v[0][1].y = fb.x;     // B, F
        o.v[1][0].x = -gc.y; o.v[1][0].y = gc.x;     // C, G
        o.v[1][1] = dh;                              // D, H
    }

    static DTCWT_D void run(const Args& a, int64_t gid) {
#ifdef DTCWT_EMU
        static F2 stage[DEPTH * 8 * kThreads];
        const int tid = (int)(gid % kThreads);
#else
        __shared__ F2 stage[DEPTH * 8 * kThreads];
        const int tid = threadIdx.x;
#endif
        const int wp = a.w / 2, hp = a.h / 2;
        const int xp = (int)(gid % wp);
        int64_t r = gid / wp;
        const int yp = (int)(r % hp);
        r /= hp;
        const int64_t ng = groups(a);
        const int gz = (int)(r % ng);
        r /= ng;
        const int sub = (int)(r % 4);
        const int b = (int)(r / 4);
        const int t1 = sub & 1, t2 = sub >> 1;
        const int64_t plane = (int64_t)a.h * a.w;
        const int noct = a.d0 / 2;
        Src s;
        s.real = (sub == 0);
        s.cs = 2 * a.zs_chan;
        const int64_t chan_off = 2 * ((int64_t)b * a.zs_n + (int64_t)yp * a.zs_1 + (int64_t)xp * a.zs_2);
        s.hi = a.yh + chan_off + 2 * (int64_t)(4 * octant_block(1, t1, t2)) * a.zs_chan;
        s.hi_step = 2 * a.zs_0;
        if (s.real) {
            s.lo = a.s + (int64_t)b * a.d0 * plane + (int64_t)(2 * yp) * a.w + 2 * xp;
            s.lo_step = 2 * plane; s.lo_i1 = a.w; s.lo_i2 = plane;
        } else {
            s.lo = a.yh + chan_off + 2 * (int64_t)(4 * octant_block(0, t1, t2)) * a.zs_chan;
            s.lo_step = 2 * a.zs_0; s.lo_i1 = 0; s.lo_i2 = 0;
        }
        F2 acc[2][NOUT];
#pragma unroll
        for (int i = 0; i < NOUT; ++i) { acc[0][i] = zero2(); acc[1][i] = zero2(); }
        const int o0 = (Q * NG * gz - HL) / 2;                   // first octet along axis 0 (may be negative)
        F2* mine = stage + tid;
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            issue(s, mine + d * 8 * kThreads, o0 + d, noct);
            async_commit();
        }
#pragma unroll
        for (int jq = 0; jq < NQ; ++jq) {
            async_wait<DEPTH - 1>();                             // the oldest pending group -- octet jq -- has landed
            F2* st = mine + (jq % DEPTH) * 8 * kThreads;
            F2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = st[i * kThreads];
            if (jq + DEPTH < NQ) issue(s, st, o0 + jq + DEPTH, noct);          // refill the slot just read (same thread: program order)
            async_commit();                                      // one group per step, empty ones included, keeps the count uniform
            typename Base::Oct lo, hi;
            if (s.real) {
                lo.v[0][0].x = 2.f * v[0].x; lo.v[0][0].y = 2.f * v[0].y;
                lo.v[0][1].x = 2.f * v[1].x; lo.v[0][1].y = 2.f * v[1].y;
                lo.v[1][0].x = 2.f * v[2].x; lo.v[1][0].y = 2.f * v[2].y;
                lo.v[1][1].x = 2.f * v[3].x; lo.v[1][1].y = 2.f * v[3].y;
            } else {
                unpack_vals(v[0], v[1], v[2], v[3], lo);
            }
            unpack_vals(v[4], v[5], v[6], v[7], hi);
            bool flip;
            fold(o0 + jq, noct, flip);
            if (flip) {
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    F2 t = lo.v[0][f]; lo.v[0][f] = lo.v[1][f]; lo.v[1][f] = t;
                    t = hi.v[0][f]; hi.v[0][f] = hi.v[1][f]; hi.v[1][f] = t;
                }
            }
            fir_scatter<GLO, NG, HL>(2 * jq, lo.v[0][0], a.lo, acc[0]);
            fir_scatter<GLO, NG, HL>(2 * jq, lo.v[0][1], a.lo, acc[1]);
            fir_scatter<GHI, NG, HL>(2 * jq, hi.v[0][0], a.hi, acc[0]);
            fir_scatter<GHI, NG, HL>(2 * jq, hi.v[0][1], a.hi, acc[1]);
            fir_scatter<GLO, NG, HL>(2 * jq + 1, lo.v[1][0], a.lo, acc[0]);
            fir_scatter<GLO, NG, HL>(2 * jq + 1, lo.v[1][1], a.lo, acc[1]);
            fir_scatter<GHI, NG, HL>(2 * jq + 1, hi.v[1][0], a.hi, acc[0]);
            fir_scatter<GHI, NG, HL>(2 * jq + 1, hi.v[1][1], a.hi, acc[1]);
        }
        float* d = a.lll + sub * a.sub_stride + b * a.vol_stride + (int64_t)(2 * yp) * a.w + 2 * xp;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const int zo = NOUT * gz + i - a.crop0;
            if (zo >= 0 && zo < a.out_d0) {
                *reinterpret_cast<F2*>(d + (int64_t)zo * plane) = acc[0][i];
                *reinterpret_cast<F2*>(d + (int64_t)zo * plane + a.w) = acc[1][i];
            }
        }
    }
};

}  // namespace dtcwt
