// C-ABI entry points of the fused per-level 3-D kernels (fused3d.cuh + the 3-D modes of fused2d.cuh).
// Included after abi_fused2d.inl / abi_axis.inl (it reuses their tap preparation); the including file provides
//   template <class K> int launch_fwd2d(typename K::Args&, void* stream);
//   template <class K> int launch_inv2d(typename K::Args&, void* stream);
//   template <class K> int launch_z3(const Z3Args&, void* stream);
// Requests outside what these kernels cover return DTCWT_B200_EUNSUPPORTED; the host layer then composes the
// level from the per-axis kernels.

#ifdef DTCWT_EMIT_FUSED3D
namespace dtcwt {

// shortest in-slice side: the forward slice kernels patch their halo once inside a 32-wide tile (kFusedMinSide); the
// inverse ones and both depth kernels reflect once, which needs a side no shorter than the longest halo (10)
static const int kFused3dMinSide = 32, kFused3dMinSideInv = 16;

// slice kernels of the 3-D levels
template <int M> struct FwdLqRaw { typedef Fwd2d<SpecDec<M, true>, SpecDec<M, false>, 32, 16, 4, RtPhase, RtPhase, RtPhase, kFwdRaw> type; };
template <int M> struct InvLqRaw { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, true> type; };
template <int M> struct InvLqRawA { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, true, false, RtPhase, RtPhase, false, 4> type; };   // cp.async stages
// row pass on interleaved row pairs (fused2d.cuh: ROWPAIR); qshift_b with its taps as immediates
template <int M> struct InvLqRawP { typedef Inv2d<SpecInt<M, true>, SpecInt<M, false>, 4, 1, 4, true, false, RtPhase, RtPhase, true> type; };
typedef Inv2d<SpecInt<14, true>, SpecInt<14, false>, 4, 1, 4, true, false, BakedPhaseQ<QshiftB_g0>, BakedPhaseQ<QshiftB_g1>, true> InvLqRawP_qb;
typedef Fwd2d<SpecCol<19>, SpecCol<19>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdRaw> FwdT1Raw;
typedef Inv2d<SpecCol<19>, SpecCol<19>, 8, 1, 4, true> InvT1Raw;
typedef Fwd2d<SpecCol<13, kMask13>, SpecCol<13>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdLow> FwdLow13m;   // near_sym_b h0o
typedef Fwd2d<SpecCol<19, kMask19>, SpecCol<19>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdLow> FwdLow19m;   // near_sym_b g0o
typedef Fwd2d<SpecCol<7>, SpecCol<7>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdLow> FwdLow7;
typedef Fwd2d<SpecCol<19>, SpecCol<19>, 64, 64, 8, RtPhase, RtPhase, RtPhase, kFwdLow> FwdLow19;
// depth kernels
template <int M> struct Z3FwdQ { typedef Z3Fwd<SpecDec<M, true>, SpecDec<M, false>, 2> type; };      // NG = 2: 72 registers, 3 CTAs per SM (measured 17 % faster than NG = 4)
template <int M> struct Z3InvQ { typedef Z3Inv<SpecInt<M, true>, SpecInt<M, false>, 2> type; };
typedef Z3Fwd<SpecCol<19>, SpecCol<19>, 8> Z3Fwd1;
typedef Z3Inv<SpecCol<19>, SpecCol<19>, 8> Z3Inv1;

static bool chan_strides_ok(int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2) {
    return zs_n >= 0 && zs_chan >= 0 && zs_0 >= 0 && zs_1 >= 0 && zs_2 >= 1;
}

// arguments of a slice launch whose "images" are the n*d0 slices of a volume batch
static int slices_fwd(Fwd2dArgs& a, const float* x, float* out, int64_t slices, int64_t d1, int64_t d2, int pad1, int pad2,
                      int P, int Q, int64_t sub_stride) {
    if (slices > 0x7fffffffLL / 4) return DTCWT_B200_EUNSUPPORTED;
    a.x = x; a.lolo = out; a.yh = out;
    a.n = (int)slices; a.rows = (int)d1; a.cols = (int)d2;
    a.pr_lo = pad1; a.pc_lo = pad2;
    a.Lr = (int)d1 + 2 * pad1; a.Lc = (int)d2 + 2 * pad2;
    a.out_rows = P * a.Lr / Q; a.out_cols = P * a.Lc / Q;
    a.zs_n = 0; a.zs_band = sub_stride; a.zs_row = 0;
    a.use_tma = 0;
    a.prefetch = 0;
    return DTCWT_B200_OK;
}

static int volume_check(int64_t n, int64_t d0, int64_t d1, int64_t d2, const void* p0, const void* p1, const void* p2,
                        const void* p3, int min_side = kFused3dMinSide) {
    if (n < 0 || d0 < 1 || d1 < 1 || d2 < 1) return DTCWT_B200_EINVAL;
    if (n > 0 && (!p0 || !p1 || !p2 || !p3)) return DTCWT_B200_EINVAL;
    if (d1 < min_side || d2 < min_side || d0 < 16 || d1 * d2 > 0x3fffffff || d0 > (1 << 20) || d1 > (1 << 20) || d2 > (1 << 20) || n > 65535)
        return DTCWT_B200_EUNSUPPORTED;
    if (!aligned_to(p0, 16) || !aligned_to(p1, 16) || !aligned_to(p2, 16) || !aligned_to(p3, 16)) return DTCWT_B200_EUNSUPPORTED;
    return DTCWT_B200_OK;
}

static int lowpass3d(const float* x, float* y, float* scratch, int64_t n, int64_t d0, int64_t d1, int64_t d2,
                     const double* h, int m, void* stream) {
    if (!h || m < 1) return DTCWT_B200_EINVAL;
    if (!(m & 1) || m > 19) return DTCWT_B200_EUNSUPPORTED;
    int rc = volume_check(n, d0, d1, d2, x, y, scratch, scratch);
    if (rc) return rc;
    if (n == 0) return DTCWT_B200_OK;
    Fwd2dArgs a;
    rc = slices_fwd(a, x, scratch, n * d0, d1, d2, 0, 0, 1, 1, 0);
    if (rc) return rc;
    // axes 1 and 2 of every slice in one tile kernel, then axis 0 as one coalesced pass (inner = d1 * d2)
    const int K = (m <= 7) ? 7 : 19;
    ColTaps t;
    const uint32_t nz = taps_col_s(t, h, m, K, 1.0);
    if (K == 7) {
        taps_col(a.h0, h, m, 7, 1.0); taps_col(a.v0, h, m, 7, 1.0);
        pair_tab(a.ph0, t, 7);
        a.h1s = a.h0; a.v1 = a.v0; a.v1s = a.v0; a.ph1s = a.ph0;
        rc = launch_fwd2d<FwdLow7>(a, stream);
    } else {
        ColTaps t13;
        const bool fits13 = m <= 13;
        const uint32_t nz13 = fits13 ? taps_col_s(t13, h, m, 13, 1.0) : 0;
        if (fits13 && (nz13 & ~kMask13) == 0) {
            taps_col(a.h0, h, m, 13, 1.0); taps_col(a.v0, h, m, 13, 1.0);
            pair_tab(a.ph0, t13, 13);
            a.h1s = a.h0; a.v1 = a.v0; a.v1s = a.v0; a.ph1s = a.ph0;
            rc = launch_fwd2d<FwdLow13m>(a, stream);
        } else {
            taps_col(a.h0, h, m, 19, 1.0); taps_col(a.v0, h, m, 19, 1.0);
            pair_tab(a.ph0, t, 19);
            a.h1s = a.h0; a.v1 = a.v0; a.v1s = a.v0; a.ph1s = a.ph0;
            rc = ((nz & ~kMask19) == 0) ? launch_fwd2d<FwdLow19m>(a, stream) : launch_fwd2d<FwdLow19>(a, stream);
        }
    }
    if (rc) return rc;
    AxisArgs ax;
    if (!axis_common(ax, scratch, y, n, d0, d1 * d2)) return DTCWT_B200_EUNSUPPORTED;
    ax.pad_lo = 0; ax.L = (int)d0; ax.Lout = (int)d0; ax.crop = 0; ax.accumulate = 0;
    const int KA = (m <= 7) ? 7 : (m <= 13 ? 13 : 19);
    taps_col(ax.t, h, m, KA, 1.0);
    if (ax.L >= 128 && env_int("DTCWT_B200_AXIS_NG", 16) == 16) {      // long axes: 16 outputs per thread (abi_axis.inl)
        if (KA == 13) return axis_launch_v<SpecCol<13>, 16>(ax, stream);
        if (KA == 19) return axis_launch_v<SpecCol<19>, 16>(ax, stream);
    }
    if (KA == 7) return axis_launch_v<SpecCol<7>, 8>(ax, stream);
    if (KA == 13) return axis_launch_v<SpecCol<13>, 8>(ax, stream);
    return axis_launch_v<SpecCol<19>, 8>(ax, stream);
}

static void z3_common(Z3Args& z, int64_t n, int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2) {
    z.n = (int)n;
    z.zs_n = zs_n; z.zs_chan = zs_chan; z.zs_0 = zs_0; z.zs_1 = zs_1; z.zs_2 = zs_2;
    z.pad0 = 0; z.crop0 = 0; z.out_d0 = 0; z.L0 = 0;
}

}  // namespace dtcwt

using namespace dtcwt;

extern "C" {

// transform3d.py:291-315 (_level1_xfm_no_highpass): h0o along all three axes.  scratch: n*d0*d1*d2 floats.
int dtcwt_b200_fwd3d_level1_lo_f32(const float* x, float* y, float* scratch, int64_t n, int64_t d0, int64_t d1, int64_t d2,
                                   const double* h0o, int m0, void* stream) {
    return lowpass3d(x, y, scratch, n, d0, d1, d2, h0o, m0, stream);
}

// transform3d.py:442-456 (_level1_ifm_no_highpass): g0o along all three axes.  scratch: n*d0*d1*d2 floats.
int dtcwt_b200_inv3d_level1_lo_f32(const float* yl, float* out, float* scratch, int64_t n, int64_t d0, int64_t d1, int64_t d2,
                                   const double* g0o, int m0, void* stream) {
    return lowpass3d(yl, out, scratch, n, d0, d1, d2, g0o, m0, stream);
}

// transform3d.py:208-289 (_level1_xfm), odd-length biort: lll [n][d0][d1][d2], yh planar [n][28][d0/2][d1/2][d2/2].
// scratch: 4*n*d0*d1*d2 floats.
int dtcwt_b200_fwd3d_level1_f32(const float* x, float* lll, float* yh, float* scratch, int64_t n, int64_t d0, int64_t d1,
                                int64_t d2, const double* h0o, int m0, const double* h1o, int m1, int64_t zs_n,
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void* stream) {
    if (!h0o || !h1o || m0 < 1 || m1 < 1 || !chan_strides_ok(zs_n, zs_chan, zs_0, zs_1, zs_2)) return DTCWT_B200_EINVAL;
    if (!(m0 & 1) || !(m1 & 1) || m0 > 19 || m1 > 19) return DTCWT_B200_EUNSUPPORTED;
    if ((d0 & 1) || (d1 & 1) || (d2 & 1)) return DTCWT_B200_EINVAL;
    int rc = volume_check(n, d0, d1, d2, x, lll, yh, scratch);
    if (rc) return rc;
    if (n == 0) return DTCWT_B200_OK;
    const int64_t sub = n * d0 * d1 * d2;
    Fwd2dArgs a;
    rc = slices_fwd(a, x, scratch, n * d0, d1, d2, 0, 0, 1, 1, sub);
    if (rc) return rc;
    taps_col(a.h0, h0o, m0, 19, 1.0); taps_col(a.h1s, h1o, m1, 19, 1.0);
    taps_col(a.v0, h0o, m0, 19, 1.0); taps_col(a.v1, h1o, m1, 19, 1.0); a.v1s = a.v1;
    ColTaps t0, t1;
    taps_col_s(t0, h0o, m0, 19, 1.0); taps_col_s(t1, h1o, m1, 19, 1.0);
    pair_tab(a.ph0, t0, 19); pair_tab(a.ph1s, t1, 19);
    rc = launch_fwd2d<FwdT1Raw>(a, stream);
    if (rc) return rc;
    Z3Args z;
    z3_common(z, n, zs_n, zs_chan, zs_0, zs_1, zs_2);
    z.s = scratch; z.lll = lll; z.yh = yh;
    z.d0 = (int)d0; z.L0 = (int)d0; z.h = (int)d1; z.w = (int)d2;
    z.sub_stride = sub; z.vol_stride = d0 * d1 * d2;
    taps_col(z.lo, h0o, m0, 19, 0.5); taps_col(z.hi, h1o, m1, 19, 0.5);
    return launch_z3<Z3Fwd1>(z, stream);
}

// transform3d.py:317-383 (_level2_xfm).  (lo_a, lo_b) / (hi_a, hi_b) are coldfilt's (ha, hb): the reference passes
// (h0b, h0a) and (h1b, h1a).  pad_i = replicated samples attached to EACH side of axis i (ext_mode 4: 0 or 1,
// ext_mode 8: 0 or 2; :322-335).  With L_i = d_i + 2 pad_i: lll [n][L0/2][L1/2][L2/2], yh planar
// [n][28][L0/4][L1/4][L2/4].  scratch: n*d0*L1*L2 floats.
int dtcwt_b200_fwd3d_levelq_f32(const float* x, float* lll, float* yh, float* scratch, int64_t n, int64_t d0, int64_t d1,
                                int64_t d2, int pad0, int pad1, int pad2, const double* lo_a, const double* lo_b,
                                const double* hi_a, const double* hi_b, int m, int64_t zs_n, int64_t zs_chan,
                                int64_t zs_0, int64_t zs_1, int64_t zs_2, void* stream) {
    if (!lo_a || !lo_b || !hi_a || !hi_b || m < 2 || (m & 1) || pad0 < 0 || pad0 > 2 || pad1 < 0 || pad1 > 2 || pad2 < 0 ||
        pad2 > 2 || !chan_strides_ok(zs_n, zs_chan, zs_0, zs_1, zs_2))
        return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 16 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (!(tap_dot(lo_a, lo_b, m) > 0) || (tap_dot(hi_a, hi_b, m) > 0)) return DTCWT_B200_EUNSUPPORTED;
    const int64_t L0 = d0 + 2 * pad0, L1 = d1 + 2 * pad1, L2 = d2 + 2 * pad2;
    if ((L0 % 4) || (L1 % 4) || (L2 % 4)) return DTCWT_B200_EINVAL;
    int rc = volume_check(n, d0, d1, d2, x, lll, yh, scratch);
    if (rc) return rc;
    if (n == 0) return DTCWT_B200_OK;
    const int64_t sub = n * d0 * (L1 / 2) * (L2 / 2);
    Fwd2dArgs a;
    rc = slices_fwd(a, x, scratch, n * d0, d1, d2, pad1, pad2, 2, 4, sub);
    if (rc) return rc;
    taps_dec(a.h0, lo_a, lo_b, m, true, 1.0);
    taps_dec(a.h1s, hi_a, hi_b, m, false, 1.0);
    a.v0 = a.h0; a.v1 = a.h1s; a.v1s = a.h1s;
    for (int k = 0; k <= kStreamMaxTaps; ++k) {          // row pass: (lowpass phase ph, highpass phase 1 - ph) tap pairs
        a.ph0.p[k].x = (k < m) ? a.h0.t[0][k] : 0.f;
        a.ph0.p[k].y = (k < m) ? a.h1s.t[1][k] : 0.f;
        a.ph1s.p[k].x = (k < m) ? a.h0.t[1][k] : 0.f;
        a.ph1s.p[k].y = (k < m) ? a.h1s.t[0][k] : 0.f;
    }
    if (m == 10) rc = launch_fwd2d<FwdLqRaw<10>::type>(a, stream);
    else if (m == 14) rc = launch_fwd2d<FwdLqRaw<14>::type>(a, stream);
    else if (m == 16) rc = launch_fwd2d<FwdLqRaw<16>::type>(a, stream);
    else rc = launch_fwd2d<FwdLqRaw<18>::type>(a, stream);
    if (rc) return rc;
    Z3Args z;
    z3_common(z, n, zs_n, zs_chan, zs_0, zs_1, zs_2);
    z.s = scratch; z.lll = lll; z.yh = yh;
    z.d0 = (int)d0; z.pad0 = pad0; z.L0 = (int)L0; z.h = (int)(L1 / 2); z.w = (int)(L2 / 2);
    z.sub_stride = sub; z.vol_stride = d0 * (L1 / 2) * (L2 / 2);
    taps_dec(z.lo, lo_a, lo_b, m, true, 0.5);
    taps_dec(z.hi, hi_a, hi_b, m, false, 0.5);
    // DTCWT_B200_Z3_SPLIT=0: both rows of a 2 x 2 patch in one thread, NG = 2 (the round-2 kernels)
    if (env_int("DTCWT_B200_Z3_SPLIT", 0)) {
        if (m == 10) return launch_z3<Z3FwdS<SpecDec<10, true>, SpecDec<10, false>, 4> >(z, stream);
        if (m == 14) return launch_z3<Z3FwdS<SpecDec<14, true>, SpecDec<14, false>, 4> >(z, stream);
        if (m == 16) return launch_z3<Z3FwdS<SpecDec<16, true>, SpecDec<16, false>, 4> >(z, stream);
        return launch_z3<Z3FwdS<SpecDec<18, true>, SpecDec<18, false>, 4> >(z, stream);
    }
    if (m == 10) return launch_z3<Z3FwdQ<10>::type>(z, stream);
    if (m == 14 && env_int("DTCWT_B200_Z3_NG", 2) == 4) return launch_z3<Z3Fwd<SpecDec<14, true>, SpecDec<14, false>, 4> >(z, stream);
    if (m == 14) return launch_z3<Z3FwdQ<14>::type>(z, stream);
    if (m == 16) return launch_z3<Z3FwdQ<16>::type>(z, stream);
    return launch_z3<Z3FwdQ<18>::type>(z, stream);
}

// transform3d.py:458-526 (_level2_ifm).  yl [n][a0][a1][a2], yh planar [n][28][a0/2][a1/2][a2/2]; (lo_a, lo_b) / (hi_a,
// hi_b) are colifilt's (ha, hb): the reference passes (g0b, g0a) and (g1b, g1a).  crop_i = samples dropped at EACH end
// of axis i (:505-524).  out [n][2 a0 - 2 crop0][2 a1 - 2 crop1][2 a2 - 2 crop2].  scratch: 4*n*(2 a0 - 2 crop0)*a1*a2 floats.
int dtcwt_b200_inv3d_levelq_f32(const float* yl, const float* yh, float* out, float* scratch, int64_t n, int64_t a0,
                                int64_t a1, int64_t a2, int crop0, int crop1, int crop2, const double* lo_a,
                                const double* lo_b, const double* hi_a, const double* hi_b, int m, int64_t zs_n,
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void* stream) {
    if (!lo_a || !lo_b || !hi_a || !hi_b || m < 2 || (m & 1) || crop0 < 0 || crop0 > 2 || crop1 < 0 || crop1 > 2 || crop2 < 0 ||
        crop2 > 2 || !chan_strides_ok(zs_n, zs_chan, zs_0, zs_1, zs_2))
        return DTCWT_B200_EINVAL;
    if ((a0 & 1) || (a1 & 1) || (a2 & 1)) return DTCWT_B200_EINVAL;
    if (m != 10 && m != 14 && m != 16 && m != 18) return DTCWT_B200_EUNSUPPORTED;
    if (!(tap_dot(lo_a, lo_b, m) > 0) || (tap_dot(hi_a, hi_b, m) > 0)) return DTCWT_B200_EUNSUPPORTED;
    int rc = volume_check(n, a0, a1, a2, yl, yh, out, scratch, kFused3dMinSideInv);
    if (rc) return rc;
    if (n == 0) return DTCWT_B200_OK;
    const int64_t od0 = 2 * a0 - 2 * crop0;
    const int64_t sub = n * od0 * a1 * a2;
    if (n * od0 > 0x7fffffffLL / 4 || a1 * a2 >= (1LL << 28) || 4 * sub >= (1LL << 40)) return DTCWT_B200_EUNSUPPORTED;
    Z3Args z;
    z3_common(z, n, zs_n, zs_chan, zs_0, zs_1, zs_2);
    z.s = yl; z.lll = scratch; z.yh = const_cast<float*>(yh);
    z.d0 = (int)a0; z.h = (int)a1; z.w = (int)a2; z.out_d0 = (int)od0; z.crop0 = crop0;
    z.sub_stride = sub; z.vol_stride = od0 * a1 * a2;
    taps_int(z.lo, lo_a, lo_b, m, true, 0.5);
    taps_int(z.hi, hi_a, hi_b, m, false, 0.5);
    const int zdepth = env_int("DTCWT_B200_Z3_ASYNC", 2);      // octets staged ahead by cp.async (0: plain loads, the round-2 kernel)
    if (m == 14 && zdepth == 2) rc = launch_z3<Z3InvA<SpecInt<14, true>, SpecInt<14, false>, 2, 2> >(z, stream);
    else if (m == 14 && zdepth == 3) rc = launch_z3<Z3InvA<SpecInt<14, true>, SpecInt<14, false>, 2, 3> >(z, stream);
    else if (env_int("DTCWT_B200_Z3_SPLIT", 0)) {
        if (m == 10) rc = launch_z3<Z3InvS<SpecInt<10, true>, SpecInt<10, false>, 4> >(z, stream);
        else if (m == 14) rc = launch_z3<Z3InvS<SpecInt<14, true>, SpecInt<14, false>, 4> >(z, stream);
        else if (m == 16) rc = launch_z3<Z3InvS<SpecInt<16, true>, SpecInt<16, false>, 4> >(z, stream);
        else rc = launch_z3<Z3InvS<SpecInt<18, true>, SpecInt<18, false>, 4> >(z, stream);
    }
    else if (m == 10) rc = launch_z3<Z3InvQ<10>::type>(z, stream);
    else if (m == 14 && env_int("DTCWT_B200_Z3_NG", 2) == 4) rc = launch_z3<Z3Inv<SpecInt<14, true>, SpecInt<14, false>, 4> >(z, stream);
    else if (m == 14) rc = launch_z3<Z3InvQ<14>::type>(z, stream);
    else if (m == 16) rc = launch_z3<Z3InvQ<16>::type>(z, stream);
    else rc = launch_z3<Z3InvQ<18>::type>(z, stream);
    if (rc) return rc;
    Inv2dArgs a;
    a.z = scratch; a.yh = scratch; a.out = out;
    a.n = (int)(n * od0); a.rows = (int)a1; a.cols = (int)a2;
    a.crop_r = crop1; a.crop_c = crop2;
    a.out_rows = 2 * (int)a1 - 2 * crop1; a.out_cols = 2 * (int)a2 - 2 * crop2;
    a.out_vec4 = (crop2 == 0 && (a.out_cols % 4) == 0) ? 1 : 0;
    a.zs_n = 0; a.zs_band = sub; a.zs_row = 0;
    for (int b = 0; b < 6; ++b) a.gain[b] = 1.f;
    taps_int(a.g0, lo_a, lo_b, m, true);
    taps_int(a.g1, hi_a, hi_b, m, false);
    const int v = env_int("DTCWT_B200_INVQ_VARIANT", 0);       // see dtcwt_b200_inv2d_levelq_f32
    if (m == 14 && v == 2 && BakedPhaseQ<QshiftB_g0>::same(a.g0) && BakedPhaseQ<QshiftB_g1>::same(a.g1))
        return launch_inv2d<InvLqRawP_qb>(a, stream);
    if (v >= 1) {
        if (m == 10) return launch_inv2d<InvLqRawP<10>::type>(a, stream);
        if (m == 14) return launch_inv2d<InvLqRawP<14>::type>(a, stream);
        if (m == 16) return launch_inv2d<InvLqRawP<16>::type>(a, stream);
        return launch_inv2d<InvLqRawP<18>::type>(a, stream);
    }
    if (env_int("DTCWT_B200_INVQ_ASYNC", 0) > 0) {
        if (m == 10) return launch_inv2d<InvLqRawA<10>::type>(a, stream);
        if (m == 14) return launch_inv2d<InvLqRawA<14>::type>(a, stream);
        if (m == 16) return launch_inv2d<InvLqRawA<16>::type>(a, stream);
        return launch_inv2d<InvLqRawA<18>::type>(a, stream);
    }
    if (m == 10) return launch_inv2d<InvLqRaw<10>::type>(a, stream);
    if (m == 14) return launch_inv2d<InvLqRaw<14>::type>(a, stream);
    if (m == 16) return launch_inv2d<InvLqRaw<16>::type>(a, stream);
    return launch_inv2d<InvLqRaw<18>::type>(a, stream);
}

// transform3d.py:385-440 (_level1_ifm), odd-length biort: yl [n][a0][a1][a2], yh planar [n][28][a0/2][a1/2][a2/2] ->
// out [n][a0][a1][a2].  scratch: 4*n*a0*a1*a2 floats.
int dtcwt_b200_inv3d_level1_f32(const float* yl, const float* yh, float* out, float* scratch, int64_t n, int64_t a0,
                                int64_t a1, int64_t a2, const double* g0o, int m0, const double* g1o, int m1,
                                int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void* stream) {
    if (!g0o || !g1o || m0 < 1 || m1 < 1 || !chan_strides_ok(zs_n, zs_chan, zs_0, zs_1, zs_2)) return DTCWT_B200_EINVAL;
    if (!(m0 & 1) || !(m1 & 1) || m0 > 19 || m1 > 19) return DTCWT_B200_EUNSUPPORTED;
    if ((a0 & 1) || (a1 & 1) || (a2 & 1)) return DTCWT_B200_EINVAL;
    int rc = volume_check(n, a0, a1, a2, yl, yh, out, scratch, kFused3dMinSideInv);
    if (rc) return rc;
    if (n == 0) return DTCWT_B200_OK;
    const int64_t sub = n * a0 * a1 * a2;
    if (n * a0 > 0x7fffffffLL / 4 || a1 * a2 >= (1LL << 28) || 4 * sub >= (1LL << 40)) return DTCWT_B200_EUNSUPPORTED;
    Z3Args z;
    z3_common(z, n, zs_n, zs_chan, zs_0, zs_1, zs_2);
    z.s = yl; z.lll = scratch; z.yh = const_cast<float*>(yh);
    z.d0 = (int)a0; z.h = (int)a1; z.w = (int)a2; z.out_d0 = (int)a0; z.crop0 = 0;
    z.sub_stride = sub; z.vol_stride = a0 * a1 * a2;
    taps_col(z.lo, g0o, m0, 19, 0.5); taps_col(z.hi, g1o, m1, 19, 0.5);
    rc = launch_z3<Z3Inv1>(z, stream);
    if (rc) return rc;
    Inv2dArgs a;
    a.z = scratch; a.yh = scratch; a.out = out;
    a.n = (int)(n * a0); a.rows = (int)a1; a.cols = (int)a2;
    a.crop_r = 0; a.crop_c = 0; a.out_rows = (int)a1; a.out_cols = (int)a2;
    a.out_vec4 = ((a.out_cols % 4) == 0) ? 1 : 0;
    a.zs_n = 0; a.zs_band = sub; a.zs_row = 0;
    for (int b = 0; b < 6; ++b) a.gain[b] = 1.f;
    taps_col(a.g0, g0o, m0, 19, 1.0); taps_col(a.g1, g1o, m1, 19, 1.0);
    return launch_inv2d<InvT1Raw>(a, stream);
}

}  // extern "C"
#endif  // DTCWT_EMIT_FUSED3D
