// C-ABI entry points of the fused per-level 2-D kernels (fused2d.cuh).  Included by
// dtcwt_b200.cu (device build) and tests/emu/emu.cpp (host emulator); the including
// file provides
//   template <class K> int launch_fwd2d(typename K::Args&, void* stream);
//   template <class K> int launch_inv2d(typename K::Args&, void* stream);
// Requests outside what the fused kernels cover return DTCWT_B200_EUNSUPPORTED; the
// host layer then composes the level from the generic CUDA kernels instead.

namespace dtcwt {

static const double kInvSqrt2 = 0.70710678118654752440;
static const int kFusedMinSide = 32;      // smaller images: reflection may wrap twice -> generic kernels

static void clear_taps(PhaseTaps& t) {
    for (int p = 0; p < kMaxPhases; ++p)
        for (int k = 0; k < kMaxPhaseTaps; ++k) t.t[p][k] = 0.f;
}

// colfilter (lowlevel.py:69-78): Y[i] = sum_k h[k] X[i + c - k], c = (m-1)/2; with k' = K-1-k (K >= m, centred)
// out[i] = sum_k' t[k'] in[i - (K-1)/2 + k'].
static void taps_col(PhaseTaps& t, const double* h, int m, int K, double scale) {
    clear_taps(t);
    const int z = (K - m) / 2;
    for (int k = 0; k < m; ++k) t.t[0][z + (m - 1 - k)] = (float)((double)(float)h[k] * scale);
}

// coldfilt (lowlevel.py:131-152): Ya[i] = sum_j ha[j] X[4i + m - 2j], Yb[i] = sum_j hb[j] X[4i + m + 1 - 2j];
// phase 0 of the output pair is Ya when pos else Yb.  Reversed: t[j'] = f[m-1-j'], in[4i - m + 2 + delta + 2j'].
static void taps_dec(PhaseTaps& t, const double* ha, const double* hb, int m, bool pos, double scale) {
    clear_taps(t);
    for (int ph = 0; ph < 2; ++ph) {
        const double* f = ((ph == 0) == pos) ? ha : hb;
        for (int j = 0; j < m; ++j) t.t[ph][j] = (float)((double)(float)f[m - 1 - j] * scale);
    }
}

// colifilt (lowlevel.py:205-258): Y[4i+ph] = sum_k f_ph[2k + tp_ph] X[2i + m2 - 2k + off_ph], f_ph = hb for odd ph.
// Reversed: t[ph][k'] = f_ph[2(m2-1-k') + tp_ph], in[2i - m2 + 2 + off_ph + 2k'].
static void taps_int(PhaseTaps& t, const double* ha, const double* hb, int m, bool pos, double scale = 1.0) {
    clear_taps(t);
    int tp[4], off[4];
    colifilt_phase_tables(m, pos, tp, off);
    const int m2 = m / 2;
    for (int ph = 0; ph < 4; ++ph) {
        const double* f = (ph & 1) ? hb : ha;
        for (int k = 0; k < m2; ++k) t.t[ph][k] = (float)((double)(float)f[2 * (m2 - 1 - k) + tp[ph]] * scale);
    }
}

static bool aligned_to(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

// Streaming kernels (stream2d.cuh): reversed, centred taps as for taps_col; returns the mask of non-zero taps.
static uint32_t taps_col_s(ColTaps& t, const double* h, int m, int K, double scale) {
    for (int k = 0; k < kStreamMaxTaps; ++k) t.t[k] = 0.f;
    const int z = (K - m) / 2;
    uint32_t mask = 0;
    for (int k = 0; k < m; ++k) {
        const float v = (float)((double)(float)h[k] * scale);
        t.t[z + (m - 1 - k)] = v;
        if (v != 0.f) mask |= 1u << (z + (m - 1 - k));
    }
    return mask;
}
static void pair_tab(PairTab& p, const ColTaps& t, int K) {
    for (int k = 0; k <= kStreamMaxTaps; ++k) {
        p.p[k].x = (k < K) ? t.t[k] : 0.f;
        p.p[k].y = (k >= 1 && k - 1 < K) ? t.t[k - 1] : 0.f;
    }
}
static int env_int(const char* name, int dflt) {      // tuning experiments only
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
// Emitting periods per run: enough runs to fill the GPU several times over (148 SMs x 2 CTAs), but runs long
// enough that the one warm-up period stays a few per cent of the work.
static int choose_periods(int rows, int ring, int64_t strips_times_n) {
    const int total = (rows + ring - 1) / ring;
    int64_t runs = (4 * 296 + strips_times_n - 1) / (strips_times_n > 0 ? strips_times_n : 1);
    if (runs < 1) runs = 1;
    int per = (int)((total + runs - 1) / runs);
    if (per < 16) per = 16;
    if (per > total) per = total;
    return per;
}

// ------------------------------------------------------------------ kernel instances
// level 1: (lowpass taps, highpass taps) of the shipped biorthogonal families, longest first
// level-1 forward: tile kernels (TMA-staged tile, packed row pass into shared memory, column pass from shared memory)
constexpr uint32_t kMask19 = 0x7ffffu & ~(1u << 1) & ~(1u << 17);    // near_sym_b: taps 1 and m-2 are exactly zero
constexpr uint32_t kMask13 = 0x1fffu & ~(1u << 1) & ~(1u << 11);
typedef Fwd2d<SpecCol<13, kMask13>, SpecCol<19, kMask19>, 64, 64, 8, BakedPhase<NearSymB_h0>, BakedPhase<NearSymB_h1s>,
              BakedPhase<NearSymB_h1> > FwdT1_nsb;    // near_sym_b: exact-zero taps compiled out, column taps as immediates
typedef Fwd2d<SpecCol<13, kMask13>, SpecCol<19, kMask19>, 64, 64, 8, BakedPhase<NearSymB_h0>, BakedPhase<NearSymB_h1s>,
              BakedPhase<NearSymB_h1>, kFwdSym> FwdT1_nsb_sym;    // the same, column pass with shared symmetric sums: the default (measured 5 % faster)
typedef Fwd2d<SpecCol<13, kMask13>, SpecCol<19, kMask19>, 64, 64, 8, BakedPhase<NearSymB_h0>, BakedPhase<NearSymB_h1s>,
              BakedPhase<NearSymB_h1>, kFwdSymP> FwdT1_nsb_symp;   // experiment: DTCWT_B200_FWD_PERSIST=1
typedef Fwd2d<SpecCol<5>, SpecCol<7>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdSym> FwdT1_5_7_sym;
typedef Fwd2d<SpecCol<19>, SpecCol<19>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdSym> FwdT1_19_19_sym;
typedef Fwd2d<SpecCol<5>, SpecCol<7>, 64, 64, 8> FwdT1_5_7;                         // near_sym_a (+ legall 5/3)
typedef Fwd2d<SpecCol<19>, SpecCol<19>, 64, 64, 8> FwdT1_19_19;                     // any odd pair up to 19 taps
// level-1 forward: streaming kernels, selected with DTCWT_B200_FWD_STREAM=1 (h0 taps, h1 taps, masks of taps that may be non-zero, ring)
typedef FwdS1<13, 19, kMask13, kMask19, 20, 192, 3, BakedTaps<NearSymB_h0>, BakedTaps<NearSymB_h1s>, BakedTaps<NearSymB_h1> > FwdL1_nsb;
typedef FwdS1<19, 19, 0x7ffffu, 0x7ffffu, 20, 192, 2> FwdL1_19_19;          // any odd pair up to 19 taps (zero-padded)
typedef FwdS1<5, 7, 0x1fu, 0x7fu, 8, 192, 3> FwdL1_5_7;                     // near_sym_a (+ legall 5/3)
// level-1 inverse: streaming kernels (g0 taps, g1 taps, masks of taps that may be non-zero, ring, prefetch depth)
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1> > InvL1_nsb;   // near_sym_b, taps as immediates
#ifdef DTCWT_DIAGNOSIS     // python build.py --diagnosis: builds that give WRONG results on purpose, never part of the default library
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 2, 1> InvL1_nsbB;   // memory traffic only
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 2, 2> InvL1_nsbC;   // arithmetic only
#endif
typedef InvS1<19, 19, 0x7ffffu, 0x7ffffu, 24, 3> InvL1_19_19;       // any odd pair up to 19 taps (zero-padded)
// the same with the prefetched quad rows staged in shared memory by per-thread cp.async (last argument: stages)
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 2, 0, false, 6> InvL1_nsbA;
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 2, 0, false, 4> InvL1_nsbA4;
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 2, 0, false, 6, true> InvL1_nsbAU;   // one instruction stream for both roles
typedef InvS1<19, 13, kMask19, kMask13, 24, 3, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1>, 3, 0, false, 3> InvL1_nsbA3;   // 3 CTAs per SM
typedef InvS1<19, 19, 0x7ffffu, 0x7ffffu, 24, 3, ArgTaps, ArgTaps, 2, 0, false, 6> InvL1_19_19A;
typedef InvS1<7, 5, 0x7fu, 0x1fu, 8, 2, ArgTaps, ArgTaps, 2, 0, false, 4> InvL1_7_5A;
// the same with the inputs staged by bulk copies (ring, stages); opt-in (DTCWT_B200_INV_STAGED=1), see dtcwt_b200_inv2d_level1_f32
typedef InvS1T<19, 13, kMask19, kMask13, 24, 6, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1> > InvT1_nsb;
typedef InvS1T<19, 13, kMask19, kMask13, 24, 4, BakedTaps<NearSymB_g0>, BakedTaps<NearSymB_g1> > InvT1_nsb4;   // experiment: DTCWT_B200_INV_NSTAGE=4
typedef InvS1T<19, 19, 0x7ffffu, 0x7ffffu, 24, 6> InvT1_19_19;
typedef InvS1T<7, 5, 0x7fu, 0x1fu, 8, 4> InvT1_7_5;
typedef InvS1<7, 5, 0x7fu, 0x1fu, 8, 2> InvL1_7_5;                  // near_sym_a (+ legall 3/5)
// `_bp` families: the second launch of a level, bands 1 and 4 from the band-pass pair (h2 / g2) in the H1 / G1 slot
typedef Fwd2d<SpecCol<19>, SpecCol<19>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdHH> FwdT1_hh;
typedef InvS1<19, 19, 0x7ffffu, 0x7ffffu, 24, 3, ArgTaps, ArgTaps, 2, 0, true> InvL1_hh;
typedef InvS1<19, 19, 0x7ffffu, 0x7ffffu, 24, 3, ArgTaps, ArgTaps, 2, 0, true, 6> InvL1_hhA;      // cp.async stages (no spills at 128 registers)
template <int M> struct FwdLqHH { typedef Fwd2d<SpecDec<M, true>, SpecDec<M, false>, 32, 16, 4, RtPhase, RtPhase, RtPhase, kFwdHH> type; };
template <int M> struct InvLqHH { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, false, true> type; };
// levels >= 2: q-shift pairs; every shipped family has a positive lowpass and a negative highpass tap correlation
template <int M> struct FwdLq { typedef Fwd2d<SpecDec<M, true>, SpecDec<M, false>, 32, 16, 4> type; };
template <int M> struct InvLq { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4> type; };
// the same with the column pass's quad rows staged by per-thread cp.async (last argument: stages)
template <int M> struct InvLqA { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, false, false, RtPhase, RtPhase, false, 4> type; };
template <int M> struct InvLqA2 { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, false, false, RtPhase, RtPhase, false, 2> type; };
// the same with the row pass on interleaved row pairs (all FFMA2); qshift_b: taps as immediates
template <int M> struct InvLqP { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, false, false, RtPhase, RtPhase, true> type; };
typedef Inv2d<SpecInt<14, true>, SpecInt<14, false>, 4, 1, 4, false, false, BakedPhaseQ<QshiftB_g0>, BakedPhaseQ<QshiftB_g1>, true> InvLqP_qb;
// levels >= 2 inverse: streaming kernels for the 10- and 14-tap families (ring of 4 * (m/2 + 1) output rows); opt-in
typedef InvSq<14, 32, 2, BakedPhase2<QshiftB_g0>, BakedPhase2<QshiftB_g1> > InvSq_qb;     // qshift_b, taps as immediates
typedef InvSq<14, 32, 2> InvSq_14;
typedef InvSq<10, 24, 2> InvSq_10;

static int fwd_common(Fwd2dArgs& a, const float* x, float* lolo, float* yh, int64_t n, int64_t rows, int64_t cols,
                      int pr_lo, int pr_hi, int pc_lo, int pc_hi, int P, int Q, int64_t zs_n, int64_t zs_band,
                      int64_t zs_row) {
    if (n < 0 || rows < 1 || cols < 1 || (n > 0 && (!x || !lolo || !yh))) return DTCWT_B200_EINVAL;   // an empty batch has no buffers
    if (n > 65535 || rows < kFusedMinSide || cols < kFusedMinSide || rows > (1 << 24) || cols > (1 << 24))
        return DTCWT_B200_EUNSUPPORTED;
    if (!aligned_to(lolo, 8) || !aligned_to(yh, 8) || !aligned_to(x, 4)) return DTCWT_B200_EUNSUPPORTED;
    a.x = x; a.lolo = lolo; a.yh = yh;
    a.n = (int)n; a.rows = (int)rows; a.cols = (int)cols;
    a.pr_lo = pr_lo; a.pc_lo = pc_lo;
    a.Lr = (int)rows + pr_lo + pr_hi; a.Lc = (int)cols + pc_lo + pc_hi;
    if ((a.Lr % (2 * Q / P)) || (a.Lc % (2 * Q / P))) return DTCWT_B200_EINVAL;   // level 1: even; level q: multiple of 4
    a.out_rows = P * a.Lr / Q; a.out_cols = P * a.Lc / Q;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row;
    a.use_tma = 0;
    a.prefetch = 0;
    return DTCWT_B200_OK;
}

static int inv_common(Inv2dArgs& a, const float* z, const float* yh, float* out, int64_t n, int64_t rows,
                      int64_t cols, int crop_r, int crop_c, int P, int Q, const double* gain, int64_t zs_n,
                      int64_t zs_band, int64_t zs_row) {
    if (!gain || n < 0 || rows < 2 || cols < 2 || (rows & 1) || (cols & 1) || (n > 0 && (!z || !yh || !out)))
        return DTCWT_B200_EINVAL;
    if (crop_r < 0 || crop_r > 1 || crop_c < 0 || crop_c > 1) return DTCWT_B200_EINVAL;
    if (n > 65535 || rows < kFusedMinSide || cols < kFusedMinSide || rows > (1 << 24) || cols > (1 << 24))
        return DTCWT_B200_EUNSUPPORTED;
    if (!aligned_to(z, 8) || !aligned_to(yh, 8) || !aligned_to(out, 8)) return DTCWT_B200_EUNSUPPORTED;
    // the kernels index one image's lowpass / sub-bands with 32-bit element offsets
    if (rows * cols >= (1LL << 30) || zs_band < 0 || zs_row < 0 || 6 * zs_band + (rows / 2) * zs_row + cols >= (1LL << 29))
        return DTCWT_B200_EUNSUPPORTED;
    a.z = z; a.yh = yh; a.out = out;
    a.n = (int)n; a.rows = (int)rows; a.cols = (int)cols;
    a.crop_r = crop_r; a.crop_c = crop_c;
    a.out_rows = P * (int)rows / Q - 2 * crop_r; a.out_cols = P * (int)cols / Q - 2 * crop_c;
    a.out_vec4 = (crop_c == 0 && (a.out_cols % 4) == 0 && aligned_to(out, 16)) ? 1 : 0;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row;
    for (int b = 0; b < 6; ++b) a.gain[b] = (float)(gain[b] * kInvSqrt2);
    return DTCWT_B200_OK;
}

}  // namespace dtcwt

using namespace dtcwt;

extern "C" {

#ifdef DTCWT_EMIT_FWD2D
// transform2d.py:112-130 (level 1 of Transform2d.forward), 4-tuple biort
int dtcwt_b200_fwd2d_level1_f32(const float* x, float* lolo, float* yh, int64_t n, int64_t rows, int64_t cols,
                                int pad_r_hi, int pad_c_hi, const double* h0o, int m0, const double* h1o, int m1,
                                int64_t zs_n, int64_t zs_band, int64_t zs_row, void* stream) {
    if (!h0o || !h1o || m0 < 1 || m1 < 1 || pad_r_hi < 0 || pad_r_hi > 1 || pad_c_hi < 0 || pad_c_hi > 1)
        return DTCWT_B200_EINVAL;
    if (!(m0 & 1) || !(m1 & 1) || m0 > 19 || m1 > 19) return DTCWT_B200_EUNSUPPORTED;
    Fwd2dArgs c;                      // argument checks shared with the q-shift levels
    const int rc = fwd_common(c, x, lolo, yh, n, rows, cols, 0, pad_r_hi, 0, pad_c_hi, 1, 1, zs_n, zs_band, zs_row);
    if (rc) return rc;
    FwdS1Args a;
    a.x = x; a.lolo = lolo; a.yh = yh;
    a.n = c.n; a.rows = c.rows; a.cols = c.cols; a.Lr = c.Lr; a.Lc = c.Lc;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row;
    a.use_tma = 0;
    const bool small = (m0 <= 5 && m1 <= 7);
    const int K1 = small ? 7 : 19;
    const int K0 = small ? 5 : ((m0 <= 13 && m1 > m0) ? 13 : 19);
    ColTaps t0, t1s;
    uint32_t nz0 = taps_col_s(t0, h0o, m0, K0, 1.0);
    const uint32_t nz1 = taps_col_s(t1s, h1o, m1, K1, kInvSqrt2);
    taps_col_s(a.v1, h1o, m1, K1, 1.0);
    const bool nsb = !small && K0 == 13 && BakedTaps<NearSymB_h0>::same(t0) && BakedTaps<NearSymB_h1s>::same(t1s) &&
                     BakedTaps<NearSymB_h1>::same(a.v1);
    if (!small && K0 == 13 && !nsb) nz0 = taps_col_s(t0, h0o, m0, 19, 1.0);      // general instance: two 19-slot filters
    const int K0e = (small || nsb) ? K0 : 19;
    pair_tab(a.ph0, t0, K0e);
    pair_tab(a.ph1s, t1s, K1);
    a.v0 = t0;
    a.v1s = t1s;
    if (!env_int("DTCWT_B200_FWD_STREAM", 0)) {
        // default: the tile kernel (measured faster than the streaming one for the forward direction, profiles/)
        const int KT0 = small ? 5 : (nsb ? 13 : 19);
        taps_col(c.h0, h0o, m0, KT0, 1.0);
        taps_col(c.h1s, h1o, m1, K1, kInvSqrt2);
        taps_col(c.v0, h0o, m0, KT0, 1.0);
        taps_col(c.v1, h1o, m1, K1, 1.0);
        taps_col(c.v1s, h1o, m1, K1, kInvSqrt2);
        ColTaps r0;
        taps_col_s(r0, h0o, m0, KT0, 1.0);
        pair_tab(c.ph0, r0, KT0);
        pair_tab(c.ph1s, t1s, K1);
        // symmetric pairs (every shipped biorthogonal family) share the sums x[c-k] + x[c+k] between the two column filters;
        // checked bit for bit on the taps as the kernels will see them (DTCWT_B200_FWD_SYM=0: scatter form)
        bool sym = env_int("DTCWT_B200_FWD_SYM", 1) != 0;
        for (int k = 0; k < KT0 && sym; ++k) sym = c.v0.t[0][k] == c.v0.t[0][KT0 - 1 - k];
        for (int k = 0; k < K1 && sym; ++k) sym = c.v1.t[0][k] == c.v1.t[0][K1 - 1 - k] && c.v1s.t[0][k] == c.v1s.t[0][K1 - 1 - k];
        if (small) return sym ? launch_fwd2d<FwdT1_5_7_sym>(c, stream) : launch_fwd2d<FwdT1_5_7>(c, stream);
        if (nsb && sym && env_int("DTCWT_B200_FWD_PERSIST", 0)) return launch_fwd2d<FwdT1_nsb_symp>(c, stream);
        if (nsb) return sym ? launch_fwd2d<FwdT1_nsb_sym>(c, stream) : launch_fwd2d<FwdT1_nsb>(c, stream);
        return sym ? launch_fwd2d<FwdT1_19_19_sym>(c, stream) : launch_fwd2d<FwdT1_19_19>(c, stream);
    }
    if (small) {
        a.periods = choose_periods(a.Lr, FwdL1_5_7::RING, (int64_t)FwdL1_5_7::tiles_c(a) * a.n);
        return launch_fwds1<FwdL1_5_7>(a, stream);
    }
    if (nsb) {
        a.periods = choose_periods(a.Lr, FwdL1_nsb::RING, (int64_t)FwdL1_nsb::tiles_c(a) * a.n);
        return launch_fwds1<FwdL1_nsb>(a, stream);
    }
    a.periods = choose_periods(a.Lr, FwdL1_19_19::RING, (int64_t)FwdL1_19_19::tiles_c(a) * a.n);
    return launch_fwds1<FwdL1_19_19>(a, stream);
}

// transform2d.py:132-160 (levels >= 2 of Transform2d.forward), 8-tuple qshift.  (lo_a, lo_b) and (hi_a, hi_b) are
// coldfilt's (ha, hb) arguments, i.e. the reference passes (h0b, h0a) and (h1b, h1a).  pad_r / pad_c = 1 extends
// the input by one replicated sample on each side of that axis (transform2d.py:134-140).
int dtcwt_b200_fwd2d_levelq_f32(const float* x, float* lolo, float* yh, int64_t n, int64_t rows, int64_t cols,
                                int pad_r, int pad_c, const double* lo_a, const double* lo_b, const double* hi_a,
                                const double* hi_b, int m, int64_t zs_n, int64_t zs_band, int64_t zs_row,
                                void* stream) {
    if (!lo_a || !lo_b || !hi_a || !hi_b || m < 2 || (m & 1) || pad_r < 0 || pad_r > 1 || pad_c < 0 || pad_c > 1)
        return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 16 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (!(tap_dot(lo_a, lo_b, m) > 0) || (tap_dot(hi_a, hi_b, m) > 0)) return DTCWT_B200_EUNSUPPORTED;
    Fwd2dArgs a;
    const int rc = fwd_common(a, x, lolo, yh, n, rows, cols, pad_r, pad_r, pad_c, pad_c, 2, 4, zs_n, zs_band, zs_row);
    if (rc) return rc;
    taps_dec(a.h0, lo_a, lo_b, m, true, 1.0);
    taps_dec(a.h1s, hi_a, hi_b, m, false, kInvSqrt2);
    taps_dec(a.v0, lo_a, lo_b, m, true, 1.0);
    taps_dec(a.v1, hi_a, hi_b, m, false, 1.0);
    taps_dec(a.v1s, hi_a, hi_b, m, false, kInvSqrt2);
    for (int k = 0; k <= kStreamMaxTaps; ++k) {          // row pass: (lowpass phase ph, highpass phase 1 - ph) tap pairs
        a.ph0.p[k].x = (k < m) ? a.h0.t[0][k] : 0.f;
        a.ph0.p[k].y = (k < m) ? a.h1s.t[1][k] : 0.f;
        a.ph1s.p[k].x = (k < m) ? a.h0.t[1][k] : 0.f;
        a.ph1s.p[k].y = (k < m) ? a.h1s.t[0][k] : 0.f;
    }
    if (m == 10) return launch_fwd2d<FwdLq<10>::type>(a, stream);
    if (m == 14) {
        const int v = env_int("DTCWT_B200_FWDQ_VARIANT", 0);       // tile-shape experiments (profiles/r2_05)
        if (v == 1) return launch_fwd2d<Fwd2d<SpecDec<14, true>, SpecDec<14, false>, 32, 16, 2> >(a, stream);
        if (v == 2) return launch_fwd2d<Fwd2d<SpecDec<14, true>, SpecDec<14, false>, 16, 16, 4> >(a, stream);
        if (v == 3) return launch_fwd2d<Fwd2d<SpecDec<14, true>, SpecDec<14, false>, 16, 16, 2> >(a, stream);
        if (v == 5) return launch_fwd2d<Fwd2d<SpecDec<14, true>, SpecDec<14, false>, 20, 16, 4> >(a, stream);     // 71 KB: 3 CTAs per SM
        if (v == 4)      // column tasks not split into A / B halves (half of the warps idle in the column pass)
            return launch_fwd2d<Fwd2d<SpecDec<14, true>, SpecDec<14, false>, 32, 16, 4, RtPhase, RtPhase, RtPhase, kFwdQ2c, false> >(a, stream);
        return launch_fwd2d<FwdLq<14>::type>(a, stream);
    }
    if (m == 16) return launch_fwd2d<FwdLq<16>::type>(a, stream);
    return launch_fwd2d<FwdLq<18>::type>(a, stream);
}
// transform2d.py:116-121 (`_bp` 6-tuple biort): bands 1 and 4 of level 1 are q2c(V:h2o(H:h2o(X))).  Call AFTER
// dtcwt_b200_fwd2d_level1_f32 on the same yh: the two diagonal sub-bands are overwritten.
int dtcwt_b200_fwd2d_level1_hh_f32(const float* x, float* yh, int64_t n, int64_t rows, int64_t cols, int pad_r_hi, int pad_c_hi,
                                   const double* h2o, int m2, int64_t zs_n, int64_t zs_band, int64_t zs_row, void* stream) {
    if (!h2o || m2 < 1 || pad_r_hi < 0 || pad_r_hi > 1 || pad_c_hi < 0 || pad_c_hi > 1) return DTCWT_B200_EINVAL;
    if (!(m2 & 1) || m2 > 19) return DTCWT_B200_EUNSUPPORTED;
    Fwd2dArgs c;
    const int rc = fwd_common(c, x, yh, yh, n, rows, cols, 0, pad_r_hi, 0, pad_c_hi, 1, 1, zs_n, zs_band, zs_row);
    if (rc) return rc;
    taps_col(c.h0, h2o, m2, 19, 1.0);
    taps_col(c.h1s, h2o, m2, 19, kInvSqrt2);
    c.v0 = c.h0;
    taps_col(c.v1, h2o, m2, 19, 1.0);
    c.v1s = c.h1s;
    ColTaps t;
    taps_col_s(t, h2o, m2, 19, kInvSqrt2);
    pair_tab(c.ph1s, t, 19);
    c.ph0 = c.ph1s;
    return launch_fwd2d<FwdT1_hh>(c, stream);
}

// transform2d.py:145-157 (`_bp` 12-tuple qshift): bands 1 and 4 of a level >= 2 from the band-pass pair; (h2_a, h2_b) are
// coldfilt's (ha, hb), i.e. the reference passes (h2b, h2a).  Call AFTER dtcwt_b200_fwd2d_levelq_f32 on the same yh.
int dtcwt_b200_fwd2d_levelq_hh_f32(const float* x, float* yh, int64_t n, int64_t rows, int64_t cols, int pad_r, int pad_c,
                                   const double* h2_a, const double* h2_b, int m, int64_t zs_n, int64_t zs_band, int64_t zs_row,
                                   void* stream) {
    if (!h2_a || !h2_b || m < 2 || (m & 1) || pad_r < 0 || pad_r > 1 || pad_c < 0 || pad_c > 1) return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (tap_dot(h2_a, h2_b, m) > 0) return DTCWT_B200_EUNSUPPORTED;
    Fwd2dArgs a;
    const int rc = fwd_common(a, x, yh, yh, n, rows, cols, pad_r, pad_r, pad_c, pad_c, 2, 4, zs_n, zs_band, zs_row);
    if (rc) return rc;
    taps_dec(a.h1s, h2_a, h2_b, m, false, kInvSqrt2);
    taps_dec(a.v1, h2_a, h2_b, m, false, 1.0);
    a.h0 = a.h1s; a.v0 = a.v1; a.v1s = a.h1s;
    for (int k = 0; k <= kStreamMaxTaps; ++k) {          // the packed row pass pairs a lowpass phase with a highpass one: h2 in both
        a.ph0.p[k].x = 0.f;
        a.ph0.p[k].y = (k < m) ? a.h1s.t[1][k] : 0.f;
        a.ph1s.p[k].x = 0.f;
        a.ph1s.p[k].y = (k < m) ? a.h1s.t[0][k] : 0.f;
    }
    if (m == 10) return launch_fwd2d<FwdLqHH<10>::type>(a, stream);
    if (m == 14) return launch_fwd2d<FwdLqHH<14>::type>(a, stream);
    return launch_fwd2d<FwdLqHH<18>::type>(a, stream);
}
#endif  // DTCWT_EMIT_FWD2D

#ifdef DTCWT_EMIT_INV2D_Q
// transform2d.py:240-273 (levels >= 2 of Transform2d.inverse).  (lo_a, lo_b), (hi_a, hi_b) are colifilt's (ha, hb):
// the reference passes (g0b, g0a) and (g1b, g1a).  gain[6] is this level's gain_mask column.
int dtcwt_b200_inv2d_levelq_f32(const float* z, const float* yh, float* out, int64_t n, int64_t rows, int64_t cols,
                                int crop_r, int crop_c, const double* lo_a, const double* lo_b, const double* hi_a,
                                const double* hi_b, int m, const double* gain, int64_t zs_n, int64_t zs_band,
                                int64_t zs_row, void* stream) {
    if (!lo_a || !lo_b || !hi_a || !hi_b || m < 2 || (m & 1)) return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 16 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (!(tap_dot(lo_a, lo_b, m) > 0) || (tap_dot(hi_a, hi_b, m) > 0)) return DTCWT_B200_EUNSUPPORTED;
    Inv2dArgs a;
    const int rc = inv_common(a, z, yh, out, n, rows, cols, crop_r, crop_c, 4, 2, gain, zs_n, zs_band, zs_row);
    if (rc) return rc;
    taps_int(a.g0, lo_a, lo_b, m, true);
    taps_int(a.g1, hi_a, hi_b, m, false);
    if (cols >= (1 << 27) || zs_row >= (1 << 27)) return DTCWT_B200_EUNSUPPORTED;   // byte strides are 32-bit
    // The streaming kernel is parity-tested but measured no faster than the tile kernel on level 2 and slower on the small
    // levels (few, long CTAs): profiles/r1_04.  It is selected with DTCWT_B200_INV_STREAM=1.
    if ((m == 10 || m == 14) && env_int("DTCWT_B200_INV_STREAM", 0) && cols < (1 << 27) && zs_row < (1 << 27)) {
        InvSqArgs s;
        s.z = z; s.yh = yh; s.out = out;
        s.n = a.n; s.rows = a.rows; s.cols = a.cols;
        s.crop_r = a.crop_r; s.crop_c = a.crop_c; s.out_rows = a.out_rows; s.out_cols = a.out_cols;
        s.out_vec4 = a.out_vec4;
        s.zs_n = zs_n; s.zs_band = zs_band; s.zs_row = zs_row;
        for (int b = 0; b < 6; ++b) s.gain[b] = a.gain[b];
        s.g0 = a.g0; s.g1 = a.g1;
        for (int k = 0; k <= kStreamMaxTaps; ++k) {
            const bool in = k < m / 2;
            s.q[0].p[k].x = in ? a.g0.t[0][k] : 0.f; s.q[0].p[k].y = in ? a.g0.t[2][k] : 0.f;
            s.q[1].p[k].x = in ? a.g1.t[0][k] : 0.f; s.q[1].p[k].y = in ? a.g1.t[2][k] : 0.f;
            s.q[2].p[k].x = in ? a.g0.t[1][k] : 0.f; s.q[2].p[k].y = in ? a.g0.t[3][k] : 0.f;
            s.q[3].p[k].x = in ? a.g1.t[1][k] : 0.f; s.q[3].p[k].y = in ? a.g1.t[3][k] : 0.f;
        }
        if (m == 10) {
            s.periods = choose_periods(2 * s.rows, InvSq_10::RING, (int64_t)InvSq_10::tiles_c(s) * s.n);
            return launch_invs1<InvSq_10>(s, stream);
        }
        if (BakedPhase2<QshiftB_g0>::same(s.g0) && BakedPhase2<QshiftB_g1>::same(s.g1)) {
            s.periods = choose_periods(2 * s.rows, InvSq_qb::RING, (int64_t)InvSq_qb::tiles_c(s) * s.n);
            return launch_invs1<InvSq_qb>(s, stream);
        }
        s.periods = choose_periods(2 * s.rows, InvSq_14::RING, (int64_t)InvSq_14::tiles_c(s) * s.n);
        return launch_invs1<InvSq_14>(s, stream);
    }
    // DTCWT_B200_INVQ_VARIANT: 0 (default) = single-row row pass (unpacked FFMA), 1 = row pairs (all FFMA2) with runtime taps,
    // 2 = row pairs with baked immediates for qshift_b.  Measured 0.736 / 0.766 / 0.754 ms per 16 x 4096^2 step
    // (profiles/r3_01): a third fewer instructions did not make the kernel faster, so the variants stay opt-in.
    const int v = env_int("DTCWT_B200_INVQ_VARIANT", 0);
    if (m == 14 && v == 2 && BakedPhaseQ<QshiftB_g0>::same(a.g0) && BakedPhaseQ<QshiftB_g1>::same(a.g1))
        return launch_inv2d<InvLqP_qb>(a, stream);
    if (v >= 1) {
        if (m == 10) return launch_inv2d<InvLqP<10>::type>(a, stream);
        if (m == 14) return launch_inv2d<InvLqP<14>::type>(a, stream);
        if (m == 16) return launch_inv2d<InvLqP<16>::type>(a, stream);
        return launch_inv2d<InvLqP<18>::type>(a, stream);
    }
    // stages of a per-thread cp.async prefetch in the column pass (0, the default: one quad row ahead in registers).  Measured
    // SLOWER on this issue-bound kernel, 0.80 (4 stages) / 0.83 (2) vs 0.73 ms per 16 x 4096^2 step (profiles/r3_01): opt-in
    const int nasync = env_int("DTCWT_B200_INVQ_ASYNC", 0);
    if (nasync == 2 && m == 14) return launch_inv2d<InvLqA2<14>::type>(a, stream);
    if (nasync > 0) {
        if (m == 10) return launch_inv2d<InvLqA<10>::type>(a, stream);
        if (m == 14) return launch_inv2d<InvLqA<14>::type>(a, stream);
        if (m == 16) return launch_inv2d<InvLqA<16>::type>(a, stream);
        return launch_inv2d<InvLqA<18>::type>(a, stream);
    }
    if (m == 10) return launch_inv2d<InvLq<10>::type>(a, stream);
    if (m == 14) return launch_inv2d<InvLq<14>::type>(a, stream);
    if (m == 16) return launch_inv2d<InvLq<16>::type>(a, stream);
    return launch_inv2d<InvLq<18>::type>(a, stream);
}
// transform2d.py:254-262 (`_bp`): out += H:g2(V:g2(c2q(bands 1, 4))); (g2_a, g2_b) are colifilt's (ha, hb), i.e. the reference
// passes (g2b, g2a).  gain[6] is the level's gain_mask column (entries 1 and 4 are used).  Call AFTER
// dtcwt_b200_inv2d_levelq_f32 with gain[1] = gain[4] = 0 on the same out.
int dtcwt_b200_inv2d_levelq_hh_f32(const float* yh, float* out, int64_t n, int64_t rows, int64_t cols, int crop_r, int crop_c,
                                   const double* g2_a, const double* g2_b, int m, const double* gain, int64_t zs_n, int64_t zs_band,
                                   int64_t zs_row, void* stream) {
    if (!g2_a || !g2_b || m < 2 || (m & 1)) return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (tap_dot(g2_a, g2_b, m) > 0) return DTCWT_B200_EUNSUPPORTED;
    if (!gain) return DTCWT_B200_EINVAL;
    const double g[6] = {0.0, gain[1], 0.0, 0.0, gain[4], 0.0};
    Inv2dArgs a;
    const int rc = inv_common(a, out, yh, out, n, rows, cols, crop_r, crop_c, 4, 2, g, zs_n, zs_band, zs_row);   // the lowpass pointer is never used for data
    if (rc) return rc;
    a.out_vec4 = 0;
    if (cols >= (1 << 27) || zs_row >= (1 << 27)) return DTCWT_B200_EUNSUPPORTED;
    clear_taps(a.g0);
    taps_int(a.g1, g2_a, g2_b, m, false);
    if (m == 10) return launch_inv2d<InvLqHH<10>::type>(a, stream);
    if (m == 14) return launch_inv2d<InvLqHH<14>::type>(a, stream);
    return launch_inv2d<InvLqHH<18>::type>(a, stream);
}
#endif  // DTCWT_EMIT_INV2D_Q

#ifdef DTCWT_EMIT_INV2D_1
// transform2d.py:275-293 (level 1 of Transform2d.inverse), 4-tuple biort
int dtcwt_b200_inv2d_level1_f32(const float* z, const float* yh, float* out, int64_t n, int64_t rows, int64_t cols,
                                const double* g0o, int m0, const double* g1o, int m1, const double* gain,
                                int64_t zs_n, int64_t zs_band, int64_t zs_row, void* stream) {
    if (!g0o || !g1o || m0 < 1 || m1 < 1) return DTCWT_B200_EINVAL;
    if (!(m0 & 1) || !(m1 & 1) || m0 > 19 || m1 > 19) return DTCWT_B200_EUNSUPPORTED;
    Inv2dArgs c;                      // argument checks shared with the q-shift levels
    const int rc = inv_common(c, z, yh, out, n, rows, cols, 0, 0, 1, 1, gain, zs_n, zs_band, zs_row);
    if (rc) return rc;
    if (cols >= (1 << 27) || zs_row >= (1 << 27)) return DTCWT_B200_EUNSUPPORTED;   // byte strides are 32-bit
    InvS1Args a;
    a.z = z; a.yh = yh; a.out = out;
    a.n = c.n; a.rows = c.rows; a.cols = c.cols;
    a.out_vec4 = ((a.cols % 4) == 0 && aligned_to(out, 16)) ? 1 : 0;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row;
    for (int b = 0; b < 6; ++b) a.gain[b] = c.gain[b];
    const bool small = (m0 <= 7 && m1 <= 5);
    const int K0 = small ? 7 : 19;
    const int K1 = small ? 5 : ((m1 <= 13 && m0 > m1) ? 13 : 19);
    const uint32_t nz0 = taps_col_s(a.g0, g0o, m0, K0, 1.0);
    const uint32_t nz1 = taps_col_s(a.g1, g1o, m1, K1, 1.0);
    pair_tab(a.p0, a.g0, K0);
    pair_tab(a.p1, a.g1, K1);
    // DTCWT_B200_INV_STAGED=1: inputs staged by asynchronous bulk copies (needs 16-byte aligned row segments).  Parity-tested,
    // but measured SLOWER than the per-thread loads (1.35 vs 1.20 ms per 16 x 4096^2, profiles/r2_02): opt-in.
    const bool staged = env_int("DTCWT_B200_INV_STAGED", 0) != 0 && (cols % 4) == 0 && (zs_row % 2) == 0 && (zs_band % 2) == 0 &&
                        (zs_n % 2) == 0 && aligned_to(z, 16) && aligned_to(yh, 16);
    if (staged && small) {
        a.periods = choose_periods(a.rows, InvT1_7_5::RING, (int64_t)InvT1_7_5::tiles_c(a) * a.n);
        return launch_invs1t<InvT1_7_5>(a, stream);
    }
    if (staged && K1 == 13 && BakedTaps<NearSymB_g0>::same(a.g0) && BakedTaps<NearSymB_g1>::same(a.g1)) {
        a.periods = choose_periods(a.rows, InvT1_nsb::RING, (int64_t)InvT1_nsb::tiles_c(a) * a.n);
        if (env_int("DTCWT_B200_INV_NSTAGE", 6) == 4) return launch_invs1t<InvT1_nsb4>(a, stream);
        return launch_invs1t<InvT1_nsb>(a, stream);
    }
    if (staged) {
        if (K1 == 13) {                  // the general instance takes two 19-slot filters
            taps_col_s(a.g1, g1o, m1, 19, 1.0);
            pair_tab(a.p1, a.g1, 19);
        }
        a.periods = choose_periods(a.rows, InvT1_19_19::RING, (int64_t)InvT1_19_19::tiles_c(a) * a.n);
        return launch_invs1t<InvT1_19_19>(a, stream);
    }
    // DTCWT_B200_INV_ASYNC: stages of the per-thread cp.async prefetch (0: register stages, the round-2 kernel)
    const int nasync = env_int("DTCWT_B200_INV_ASYNC", 6);
    if (small && nasync > 0) {
        a.periods = choose_periods(a.rows, InvL1_7_5A::RING, (int64_t)InvL1_7_5A::tiles_c(a) * a.n);
        return launch_invs1<InvL1_7_5A>(a, stream);
    }
    if (small) {
        a.periods = choose_periods(a.rows, InvL1_7_5::RING, (int64_t)InvL1_7_5::tiles_c(a) * a.n);
        return launch_invs1<InvL1_7_5>(a, stream);
    }
    if (nasync > 0 && K1 == 13 && BakedTaps<NearSymB_g0>::same(a.g0) && BakedTaps<NearSymB_g1>::same(a.g1)) {
        a.periods = choose_periods(a.rows, InvL1_nsbA::RING, (int64_t)InvL1_nsbA::tiles_c(a) * a.n);
        if (env_int("DTCWT_B200_INV_UNI", 1)) return launch_invs1<InvL1_nsbAU>(a, stream);
        if (nasync == 4) return launch_invs1<InvL1_nsbA4>(a, stream);
        if (nasync == 3) return launch_invs1<InvL1_nsbA3>(a, stream);
        return launch_invs1<InvL1_nsbA>(a, stream);
    }
    if (K1 == 13 && BakedTaps<NearSymB_g0>::same(a.g0) && BakedTaps<NearSymB_g1>::same(a.g1)) {
#ifdef DTCWT_DIAGNOSIS
        const int variant = env_int("DTCWT_B200_INV_VARIANT", 0);
        if (variant == 1) {
            a.periods = choose_periods(a.rows, InvL1_nsbB::RING, (int64_t)InvL1_nsbB::tiles_c(a) * a.n);
            return launch_invs1<InvL1_nsbB>(a, stream);
        }
        if (variant == 2) {
            a.periods = choose_periods(a.rows, InvL1_nsbC::RING, (int64_t)InvL1_nsbC::tiles_c(a) * a.n);
            return launch_invs1<InvL1_nsbC>(a, stream);
        }
#endif
        a.periods = choose_periods(a.rows, InvL1_nsb::RING, (int64_t)InvL1_nsb::tiles_c(a) * a.n);
        return launch_invs1<InvL1_nsb>(a, stream);
    }
    if (K1 == 13) {                  // the general instance takes two 19-slot filters
        taps_col_s(a.g1, g1o, m1, 19, 1.0);
        pair_tab(a.p1, a.g1, 19);
    }
    if (nasync > 0) {
        a.periods = choose_periods(a.rows, InvL1_19_19A::RING, (int64_t)InvL1_19_19A::tiles_c(a) * a.n);
        return launch_invs1<InvL1_19_19A>(a, stream);
    }
    a.periods = choose_periods(a.rows, InvL1_19_19::RING, (int64_t)InvL1_19_19::tiles_c(a) * a.n);
    return launch_invs1<InvL1_19_19>(a, stream);
}

// transform2d.py:279-292 (`_bp`): out += H:g2o(V:g2o(c2q(bands 1, 4))).  Call AFTER dtcwt_b200_inv2d_level1_f32 with
// gain[1] = gain[4] = 0 on the same out.
int dtcwt_b200_inv2d_level1_hh_f32(const float* yh, float* out, int64_t n, int64_t rows, int64_t cols, const double* g2o, int m2,
                                   const double* gain, int64_t zs_n, int64_t zs_band, int64_t zs_row, void* stream) {
    if (!g2o || m2 < 1 || !gain) return DTCWT_B200_EINVAL;
    if (!(m2 & 1) || m2 > 19) return DTCWT_B200_EUNSUPPORTED;
    const double g[6] = {0.0, gain[1], 0.0, 0.0, gain[4], 0.0};
    Inv2dArgs c;
    const int rc = inv_common(c, out, yh, out, n, rows, cols, 0, 0, 1, 1, g, zs_n, zs_band, zs_row);
    if (rc) return rc;
    if (cols >= (1 << 27) || zs_row >= (1 << 27)) return DTCWT_B200_EUNSUPPORTED;
    InvS1Args a;
    a.z = out; a.yh = yh; a.out = out;                 // the lowpass pointer is only an address that may be read; its data is ignored
    a.n = c.n; a.rows = c.rows; a.cols = c.cols;
    a.out_vec4 = ((a.cols % 4) == 0 && aligned_to(out, 16)) ? 1 : 0;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row;
    for (int b = 0; b < 6; ++b) a.gain[b] = c.gain[b];
    for (int k = 0; k < kStreamMaxTaps; ++k) a.g0.t[k] = 0.f;
    taps_col_s(a.g1, g2o, m2, 19, 1.0);
    pair_tab(a.p0, a.g0, 19);
    pair_tab(a.p1, a.g1, 19);
    a.periods = choose_periods(a.rows, InvL1_hh::RING, (int64_t)InvL1_hh::tiles_c(a) * a.n);
    if (env_int("DTCWT_B200_INV_ASYNC", 6) > 0) return launch_invs1<InvL1_hhA>(a, stream);
    return launch_invs1<InvL1_hh>(a, stream);
}
#endif  // DTCWT_EMIT_INV2D_1

}  // extern "C"
