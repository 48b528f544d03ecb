// Host dispatch of the fast float32 axis passes (axis_pass.cuh).  Included after abi_fused2d.inl (it reuses its tap
// preparation); the including file provides  template <class K> int launch_axis(const AxisArgs&, void* stream).
// Every function returns DTCWT_B200_EUNSUPPORTED when it declines; the caller then runs the generic kernel.

namespace dtcwt {

template <class F, int NG>
static int axis_launch_v(AxisArgs& a, void* stream) {
    const bool vec2 = (a.inner % 2) == 0 && ((uintptr_t)a.x % 8) == 0 && ((uintptr_t)a.y % 8) == 0;
    if (vec2) return launch_axis<AxisPass<F, NG, F2> >(a, stream);
    return launch_axis<AxisPass<F, NG, float> >(a, stream);
}

static bool axis_common(AxisArgs& a, const float* x, float* y, int64_t outer, int64_t len, int64_t inner) {
    if (inner > 0x3fffffff || len > 0x3fffffff) return false;
    a.x = x; a.y = y; a.outer = outer; a.inner = (int)inner; a.len = (int)len;
    return true;
}

}  // namespace dtcwt

#ifdef DTCWT_EMIT_GENERIC
namespace dtcwt {

static int axis_colfilter(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                          const double* h, int m, int accumulate, void* stream) {
    if (!(m & 1) || m > 19 || env_int("DTCWT_B200_NO_AXIS", 0)) return DTCWT_B200_EUNSUPPORTED;
    AxisArgs a;
    if (!axis_common(a, x, y, outer, len, inner)) return DTCWT_B200_EUNSUPPORTED;
    a.pad_lo = pad_lo; a.L = (int)len + pad_lo + pad_hi; a.Lout = a.L; a.crop = 0; a.accumulate = accumulate;
    const int K = (m <= 7) ? 7 : (m <= 13 ? 13 : 19);
    taps_col(a.t, h, m, K, 1.0);
    // long axes: 16 outputs per thread (34 loads for 16 outputs instead of 26 for 8 with 19 taps) -- DTCWT_B200_AXIS_NG=8 restores 8
    if (a.L >= 128 && env_int("DTCWT_B200_AXIS_NG", 16) == 16) {
        if (K == 13) return axis_launch_v<SpecCol<13>, 16>(a, stream);
        if (K == 19) return axis_launch_v<SpecCol<19>, 16>(a, stream);
    }
    if (K == 7) return axis_launch_v<SpecCol<7>, 8>(a, stream);
    if (K == 13) return axis_launch_v<SpecCol<13>, 8>(a, stream);
    return axis_launch_v<SpecCol<19>, 8>(a, stream);
}

template <int M>
static int axis_dec_m(AxisArgs& a, bool pos, void* stream) {
    if (pos) return axis_launch_v<SpecDec<M, true>, 4>(a, stream);
    return axis_launch_v<SpecDec<M, false>, 4>(a, stream);
}

static int axis_coldfilt(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                         const double* ha, const double* hb, int m, int accumulate, void* stream) {
    if ((m != 10 && m != 14 && m != 16 && m != 18) || env_int("DTCWT_B200_NO_AXIS", 0)) return DTCWT_B200_EUNSUPPORTED;
    AxisArgs a;
    if (!axis_common(a, x, y, outer, len, inner)) return DTCWT_B200_EUNSUPPORTED;
    a.pad_lo = pad_lo; a.L = (int)len + pad_lo + pad_hi; a.Lout = a.L / 2; a.crop = 0; a.accumulate = accumulate;
    if (a.L % 4) return DTCWT_B200_EUNSUPPORTED;
    const bool pos = tap_dot(ha, hb, m) > 0;
    taps_dec(a.t, ha, hb, m, pos, 1.0);
    if (m == 10) return axis_dec_m<10>(a, pos, stream);
    if (m == 14) return axis_dec_m<14>(a, pos, stream);
    if (m == 16) return axis_dec_m<16>(a, pos, stream);
    return axis_dec_m<18>(a, pos, stream);
}

template <int M>
static int axis_int_m(AxisArgs& a, bool pos, void* stream) {
    if (pos) return axis_launch_v<SpecInt<M, true>, 4>(a, stream);
    return axis_launch_v<SpecInt<M, false>, 4>(a, stream);
}

static int axis_colifilt(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int crop,
                         const double* ha, const double* hb, int m, int accumulate, void* stream) {
    if ((m != 10 && m != 14 && m != 16 && m != 18) || env_int("DTCWT_B200_NO_AXIS", 0)) return DTCWT_B200_EUNSUPPORTED;
    AxisArgs a;
    if (!axis_common(a, x, y, outer, len, inner)) return DTCWT_B200_EUNSUPPORTED;
    a.pad_lo = 0; a.L = (int)len; a.Lout = 2 * (int)len - 2 * crop; a.crop = crop; a.accumulate = accumulate;
    const bool pos = tap_dot(ha, hb, m) > 0;
    taps_int(a.t, ha, hb, m, pos);
    if (m == 10) return axis_int_m<10>(a, pos, stream);
    if (m == 14) return axis_int_m<14>(a, pos, stream);
    if (m == 16) return axis_int_m<16>(a, pos, stream);
    return axis_int_m<18>(a, pos, stream);
}

}  // namespace dtcwt
#endif  // DTCWT_EMIT_GENERIC
