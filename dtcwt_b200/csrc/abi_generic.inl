// C-ABI entry points of the generic kernels.  Included by dtcwt_b200.cu (device
// build) and by tests/emu/emu.cpp (host emulator); the including file provides
//   template <class Elem> int launch_1d(const typename Elem::Args&, void* stream);
// Argument validation mirrors the reference's ValueError contracts
// (dtcwt/numpy/lowlevel.py:118-125, 189-196); the Python layer raises the same
// errors before calling, so a negative return here means a host-layer bug.

namespace dtcwt {

template <typename T>
static bool load_taps(Taps<T>& t, const double* h, int m) {
    if (h == nullptr || m < 1 || m > kMaxTaps) return false;
    for (int k = 0; k < kMaxTaps; ++k) t.v[k] = (k < m) ? (T)h[k] : T(0);
    t.m = m;
    return true;
}

static bool fits_int(int64_t v) { return v >= 0 && v < (int64_t)0x3fffffff; }

// Fast float32 paths (axis_pass.cuh, defined in abi_axis.inl further down the translation unit); they return
// DTCWT_B200_EUNSUPPORTED when they decline, and there is none for float64.
static int axis_colfilter(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                          const double* h, int m, int accumulate, void* stream);
static int axis_coldfilt(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                         const double* ha, const double* hb, int m, int accumulate, void* stream);
static int axis_colifilt(const float* x, float* y, int64_t outer, int64_t len, int64_t inner, int crop,
                         const double* ha, const double* hb, int m, int accumulate, void* stream);
static inline int axis_colfilter(const double*, double*, int64_t, int64_t, int64_t, int, int, const double*, int, int, void*) {
    return DTCWT_B200_EUNSUPPORTED;
}
static inline int axis_coldfilt(const double*, double*, int64_t, int64_t, int64_t, int, int, const double*, const double*,
                                int, int, void*) {
    return DTCWT_B200_EUNSUPPORTED;
}
static inline int axis_colifilt(const double*, double*, int64_t, int64_t, int64_t, int, const double*, const double*, int,
                                int, void*) {
    return DTCWT_B200_EUNSUPPORTED;
}

template <typename T>
static int colfilter_impl(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                          const double* h, int m, int accumulate, void* stream) {
    ColfilterArgs<T> a;
    if (outer < 0 || inner < 0 || len < 1 || pad_lo < 0 || pad_hi < 0) return DTCWT_B200_EINVAL;
    if (outer == 0 || inner == 0) return DTCWT_B200_OK;          // empty batch / no columns: nothing to launch
    if (!x || !y) return DTCWT_B200_EINVAL;
    if (!load_taps(a.h, h, m)) return DTCWT_B200_EINVAL;
    if (!fits_int(len + pad_lo + pad_hi + 1)) return DTCWT_B200_EUNSUPPORTED;
    a.x = x; a.y = y; a.outer = outer; a.inner = inner;
    a.len = (int)len; a.pad_lo = pad_lo; a.L = (int)len + pad_lo + pad_hi;
    a.Lout = a.L + ((m & 1) ? 0 : 1);
    a.accumulate = accumulate;
    if (outer > 0) {
        const int rc = axis_colfilter(x, y, outer, len, inner, pad_lo, pad_hi, h, m, accumulate, stream);
        if (rc != DTCWT_B200_EUNSUPPORTED) return rc;
    }
    return launch_1d<ColfilterElem<T> >(a, stream);
}

static double tap_dot(const double* a, const double* b, int m) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += a[k] * b[k];
    return s;
}

template <typename T>
static int coldfilt_impl(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int pad_lo, int pad_hi,
                         const double* ha, const double* hb, int m, int accumulate, void* stream) {
    ColdfiltArgs<T> a;
    if (outer < 0 || inner < 0 || len < 1 || pad_lo < 0 || pad_hi < 0) return DTCWT_B200_EINVAL;
    if (outer == 0 || inner == 0) return DTCWT_B200_OK;          // empty batch / no columns: nothing to launch
    if (!x || !y) return DTCWT_B200_EINVAL;
    if ((m & 1) || !load_taps(a.ha, ha, m) || !load_taps(a.hb, hb, m)) return DTCWT_B200_EINVAL;
    if (!fits_int(len + pad_lo + pad_hi)) return DTCWT_B200_EUNSUPPORTED;
    a.x = x; a.y = y; a.outer = outer; a.inner = inner;
    a.len = (int)len; a.pad_lo = pad_lo; a.L = (int)len + pad_lo + pad_hi;
    if (a.L % 4) return DTCWT_B200_EINVAL;
    a.Lout = a.L / 2;
    a.accumulate = accumulate;
    a.pos = tap_dot(ha, hb, m) > 0;
    if (outer > 0) {
        const int rc = axis_coldfilt(x, y, outer, len, inner, pad_lo, pad_hi, ha, hb, m, accumulate, stream);
        if (rc != DTCWT_B200_EUNSUPPORTED) return rc;
    }
    return launch_1d<ColdfiltElem<T> >(a, stream);
}

// Output phase tables of colifilt (reference lowlevel.py:205-258).
static void colifilt_phase_tables(int m, bool pos, int tp[4], int off[4]) {
    const int m2 = m / 2;
    if (m2 & 1) {
        // phases 0,2 read the "b" index set (2i+m2-1-2k), phases 1,3 the "a" set (2i+m2-2k);
        // a negative tap correlation exchanges the two index sets.
        const int ob = pos ? -1 : 0, oa = pos ? 0 : -1;
        tp[0] = 0; tp[1] = 0; tp[2] = 1; tp[3] = 1;
        off[0] = ob; off[1] = oa; off[2] = ob; off[3] = oa;
    } else {
        tp[0] = 1; tp[1] = 1; tp[2] = 0; tp[3] = 0;
        if (pos) { off[0] = -2; off[1] = -1; off[2] = 0; off[3] = 1; }
        else     { off[0] = -1; off[1] = -2; off[2] = 1; off[3] = 0; }
    }
}

template <typename T>
static int colifilt_impl(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int crop,
                         const double* ha, const double* hb, int m, int accumulate, void* stream) {
    ColifiltArgs<T> a;
    if (outer < 0 || inner < 0 || len < 2 || (len & 1) || crop < 0 || 2 * crop >= 2 * len) return DTCWT_B200_EINVAL;
    if (outer == 0 || inner == 0) return DTCWT_B200_OK;
    if (!x || !y) return DTCWT_B200_EINVAL;
    if ((m & 1) || !load_taps(a.ha, ha, m) || !load_taps(a.hb, hb, m)) return DTCWT_B200_EINVAL;
    if (!fits_int(2 * len)) return DTCWT_B200_EUNSUPPORTED;
    a.x = x; a.y = y; a.outer = outer; a.inner = inner;
    a.len = (int)len; a.crop = crop; a.Lout = 2 * (int)len - 2 * crop;
    a.accumulate = accumulate;
    colifilt_phase_tables(m, tap_dot(ha, hb, m) > 0, a.tp, a.off);
    if (outer > 0) {
        const int rc = axis_colifilt(x, y, outer, len, inner, crop, ha, hb, m, accumulate, stream);
        if (rc != DTCWT_B200_EUNSUPPORTED) return rc;
    }
    return launch_1d<ColifiltElem<T> >(a, stream);
}

template <typename T, template <typename> class Elem>
static int quad_impl(const T* src, T* dst, int64_t n, int64_t h, int64_t w, int64_t zs_n, int64_t zs_band,
                     int64_t zs_row, int64_t zs_col, int band0, int band1, double g0, double g1, void* stream) {
    if (n < 0 || h < 1 || w < 1 || band0 < 0 || band1 < 0 || (n > 0 && (!src || !dst))) return DTCWT_B200_EINVAL;
    QuadArgs<T> a;
    a.src = src; a.dst = dst; a.n = n; a.h = h; a.w = w;
    a.zs_n = zs_n; a.zs_band = zs_band; a.zs_row = zs_row; a.zs_col = zs_col;
    a.band0 = band0; a.band1 = band1;
    const double s = 0.70710678118654752440;
    a.g0 = (T)(g0 * s); a.g1 = (T)(g1 * s);
    return launch_1d<Elem<T> >(a, stream);
}

template <typename T, template <typename> class Elem>
static int pack1d_impl(const T* src, T* dst, int64_t outer, int64_t k, int64_t inner, double gain, void* stream) {
    if (outer < 0 || k < 1 || inner < 0) return DTCWT_B200_EINVAL;
    if (outer == 0 || inner == 0) return DTCWT_B200_OK;
    if (!src || !dst) return DTCWT_B200_EINVAL;
    Pack1dArgs<T> a;
    a.src = src; a.dst = dst; a.outer = outer; a.k = k; a.inner = inner; a.gain = (T)gain;
    return launch_1d<Elem<T> >(a, stream);
}

template <typename T, template <typename> class Elem>
static int cube_impl(const T* src, T* dst, int64_t n, int64_t a_, int64_t b_, int64_t c_, int64_t zs_n,
                     int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, int chan0, void* stream) {
    if (n < 0 || a_ < 1 || b_ < 1 || c_ < 1 || chan0 < 0 || (n > 0 && (!src || !dst))) return DTCWT_B200_EINVAL;
    CubeArgs<T> g;
    g.src = src; g.dst = dst; g.n = n; g.a = a_; g.b = b_; g.c = c_;
    g.zs_n = zs_n; g.zs_chan = zs_chan; g.zs_0 = zs_0; g.zs_1 = zs_1; g.zs_2 = zs_2; g.chan0 = chan0;
    return launch_1d<Elem<T> >(g, stream);
}

}  // namespace dtcwt

using namespace dtcwt;

#ifdef DTCWT_EMIT_GENERIC
extern "C" {

int dtcwt_b200_version(void) { return DTCWT_B200_VERSION; }

#define DTCWT_FILTERS(SUF, T)                                                                                     \
    int dtcwt_b200_colfilter_##SUF(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int pad_lo,      \
                                   int pad_hi, const double* h, int m, int accumulate, void* stream) {           \
        return colfilter_impl<T>(x, y, outer, len, inner, pad_lo, pad_hi, h, m, accumulate, stream);             \
    }                                                                                                             \
    int dtcwt_b200_coldfilt_##SUF(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int pad_lo,       \
                                  int pad_hi, const double* ha, const double* hb, int m, int accumulate,         \
                                  void* stream) {                                                                 \
        return coldfilt_impl<T>(x, y, outer, len, inner, pad_lo, pad_hi, ha, hb, m, accumulate, stream);         \
    }                                                                                                             \
    int dtcwt_b200_colifilt_##SUF(const T* x, T* y, int64_t outer, int64_t len, int64_t inner, int crop,         \
                                  const double* ha, const double* hb, int m, int accumulate, void* stream) {     \
        return colifilt_impl<T>(x, y, outer, len, inner, crop, ha, hb, m, accumulate, stream);                   \
    }                                                                                                             \
    int dtcwt_b200_q2c_##SUF(const T* y, T* z, int64_t n, int64_t h, int64_t w, int64_t zs_n, int64_t zs_band,   \
                             int64_t zs_row, int64_t zs_col, int band0, int band1, void* stream) {               \
        return quad_impl<T, Q2cElem>(y, z, n, h, w, zs_n, zs_band, zs_row, zs_col, band0, band1, 1, 1, stream);  \
    }                                                                                                             \
    int dtcwt_b200_c2q_##SUF(const T* z, T* y, int64_t n, int64_t h, int64_t w, int64_t zs_n, int64_t zs_band,   \
                             int64_t zs_row, int64_t zs_col, int band0, int band1, double gain0, double gain1,   \
                             void* stream) {                                                                      \
        return quad_impl<T, C2qElem>(z, y, n, h, w, zs_n, zs_band, zs_row, zs_col, band0, band1, gain0, gain1,   \
                                     stream);                                                                     \
    }                                                                                                             \
    int dtcwt_b200_pack1d_##SUF(const T* hi, T* z, int64_t outer, int64_t k, int64_t inner, void* stream) {      \
        return pack1d_impl<T, Pack1dElem>(hi, z, outer, k, inner, 1.0, stream);                                  \
    }                                                                                                             \
    int dtcwt_b200_unpack1d_##SUF(const T* z, T* hi, int64_t outer, int64_t k, int64_t inner, double gain,       \
                                  void* stream) {                                                                 \
        return pack1d_impl<T, Unpack1dElem>(z, hi, outer, k, inner, gain, stream);                               \
    }                                                                                                             \
    int dtcwt_b200_cube2c_##SUF(const T* y, T* z, int64_t n, int64_t a, int64_t b, int64_t c, int64_t zs_n,      \
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, int chan0,            \
                                void* stream) {                                                                   \
        return cube_impl<T, Cube2cElem>(y, z, n, a, b, c, zs_n, zs_chan, zs_0, zs_1, zs_2, chan0, stream);       \
    }                                                                                                             \
    int dtcwt_b200_c2cube_##SUF(const T* z, T* y, int64_t n, int64_t a, int64_t b, int64_t c, int64_t zs_n,      \
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, int chan0,            \
                                void* stream) {                                                                   \
        return cube_impl<T, C2cubeElem>(z, y, n, a, b, c, zs_n, zs_chan, zs_0, zs_1, zs_2, chan0, stream);       \
    }

DTCWT_FILTERS(f32, float)
DTCWT_FILTERS(f64, double)
#undef DTCWT_FILTERS

}  // extern "C"
#endif  // DTCWT_EMIT_GENERIC
