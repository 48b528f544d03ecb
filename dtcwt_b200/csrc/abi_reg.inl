// C-ABI entry points of the registration / re-sampling kernels (registration.cuh).  Included after abi_generic.inl;
// the including file provides  template <class Elem> int launch_1d(const typename Elem::Args&, void* stream).
#ifdef DTCWT_EMIT_GENERIC
#include <math.h>
namespace dtcwt {

static const double kExpectedShift = 3.14159265358979323846 / 2.15;
static const double kExpectedShifts[6][2] = {{-1, -3}, {-3, -3}, {-3, -1}, {-3, 1}, {-3, 3}, {-1, 3}};   // registration.py:30

template <typename T>
static int reg_qtilde_impl(const T* src, const T* ref, double* qt, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band,
                           int64_t s_row, int64_t s_col, int64_t r_n, int64_t r_band, int64_t r_row, int64_t r_col, int reduce,
                           void* stream) {
    if (n < 0 || h < 1 || w < 1) return DTCWT_B200_EINVAL;
    if (n == 0) return DTCWT_B200_OK;
    if (!src || !ref || !qt) return DTCWT_B200_EINVAL;
    if (h > 0x3fffffff || w > 0x3fffffff) return DTCWT_B200_EUNSUPPORTED;
    QtildeArgs<T> a;
    a.src = src; a.ref = ref; a.qt = qt; a.n = n; a.h = h; a.w = w;
    a.s_n = s_n; a.s_band = s_band; a.s_row = s_row; a.s_col = s_col;
    a.r_n = r_n; a.r_band = r_band; a.r_row = r_row; a.r_col = r_col;
    a.reduce = reduce;
    a.epsilon = 1e-6;                                    // registration.py:84
    for (int b = 0; b < 6; ++b) {
        a.shift[b][0] = kExpectedShifts[b][0] * kExpectedShift;
        a.shift[b][1] = kExpectedShifts[b][1] * kExpectedShift;
        for (int ax = 0; ax < 2; ++ax) {
            a.rot[b][ax][0] = cos(-a.shift[b][ax]);
            a.rot[b][ax][1] = sin(-a.shift[b][ax]);
        }
    }
    return launch_qtilde<T>(a, stream);
}

template <typename T>
static int sample_impl(const T* im, T* out, const double* xs, const double* ys, int64_t n, int64_t h, int64_t w, int64_t C,
                       int64_t oh, int64_t ow, int64_t i_n, int64_t i_y, int64_t i_x, int64_t i_c, int64_t o_n, int64_t o_y,
                       int64_t o_x, int64_t o_c, int64_t coord_n, int is_complex, int method, int coords, const double* wx,
                       const double* wy, void* stream) {
    if (n < 0 || h < 1 || w < 1 || C < 0 || oh < 0 || ow < 0 || method < 0 || method > 3 || coords < 0 || coords > 1)
        return DTCWT_B200_EINVAL;
    if (method == 3 && (coords != 1 || oh != 2 * h || ow != 2 * w)) return DTCWT_B200_EINVAL;     // upsample's kernel
    if ((wx == nullptr) != (wy == nullptr) || (wx && (C > 8 || !is_complex))) return DTCWT_B200_EINVAL;
    if (n == 0 || C == 0 || oh == 0 || ow == 0) return DTCWT_B200_OK;
    if (!im || !out || (coords == 0 && (!xs || !ys))) return DTCWT_B200_EINVAL;
    if (h > 0x3fffffff || w > 0x3fffffff) return DTCWT_B200_EUNSUPPORTED;
    SampleArgs<T> a;
    a.im = im; a.out = out; a.xs = xs; a.ys = ys;
    a.n = n; a.h = h; a.w = w; a.C = C; a.oh = oh; a.ow = ow;
    a.i_n = i_n; a.i_y = i_y; a.i_x = i_x; a.i_c = i_c; a.o_n = o_n; a.o_y = o_y; a.o_x = o_x; a.o_c = o_c;
    a.coord_n = coord_n;
    a.ncomp = is_complex ? 2 : 1; a.method = method; a.coords = coords; a.phase = wx ? 1 : 0;
    for (int c = 0; c < 8; ++c) {
        a.wx[c] = (wx && c < C) ? wx[c] : 0.0; a.wy[c] = (wy && c < C) ? wy[c] : 0.0;
        a.rx[c][0] = cos(-a.wx[c]); a.rx[c][1] = sin(-a.wx[c]);
        a.ry[c][0] = cos(-a.wy[c]); a.ry[c][1] = sin(-a.wy[c]);
    }
    return launch_1d<SampleElem<T> >(a, stream);
}

template <typename T>
static int kp_energy_impl(const T* yh, double* e, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band, int64_t s_row,
                          int64_t s_col, int method, double scale_gain, double beta, double kappa, void* stream) {
    if (n < 0 || h < 1 || w < 1 || method < 0 || method > 2) return DTCWT_B200_EINVAL;
    if (n == 0) return DTCWT_B200_OK;
    if (!yh || !e) return DTCWT_B200_EINVAL;
    KpEnergyArgs<T> a;
    a.yh = yh; a.e = e; a.n = n; a.h = h; a.w = w; a.s_n = s_n; a.s_band = s_band; a.s_row = s_row; a.s_col = s_col;
    a.method = method; a.scale_gain = scale_gain; a.beta = beta; a.kappa = kappa;
    return launch_1d<KpEnergyElem<T> >(a, stream);
}

}  // namespace dtcwt

using namespace dtcwt;

extern "C" {

// keypoint.py:143-156: keypoint energy of one level's six sub-bands; method 0 fauqueur (scale_gain = alpha**(scale+1),
// beta), 1 bendale, 2 kingsbury (kappa).  e is [n][h][w] float64.
int dtcwt_b200_kp_energy_f32(const float* yh, double* e, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band,
                             int64_t s_row, int64_t s_col, int method, double scale_gain, double beta, double kappa, void* stream) {
    return kp_energy_impl<float>(yh, e, n, h, w, s_n, s_band, s_row, s_col, method, scale_gain, beta, kappa, stream);
}
int dtcwt_b200_kp_energy_f64(const double* yh, double* e, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band,
                             int64_t s_row, int64_t s_col, int method, double scale_gain, double beta, double kappa, void* stream) {
    return kp_energy_impl<double>(yh, e, n, h, w, s_n, s_band, s_row, s_col, method, scale_gain, beta, kappa, stream);
}

// keypoint.py:201-260 (_kp_energy_maxima): out [n][h][w][4] = (1 if the pixel is a kept local maximum else 0, refined row,
// refined column, energy); refine != 0 fits the quadratic patch and drops maxima that move by more than half a pixel.
int dtcwt_b200_kp_maxima(const double* x, double* out, int64_t n, int64_t h, int64_t w, double threshold, int refine, void* stream) {
    if (n < 0 || h < 1 || w < 1) return DTCWT_B200_EINVAL;
    if (n == 0) return DTCWT_B200_OK;
    if (!x || !out) return DTCWT_B200_EINVAL;
    if (h > 0x3fffffff || w > 0x3fffffff) return DTCWT_B200_EUNSUPPORTED;
    KpMaximaArgs a;
    a.x = x; a.out = out; a.n = n; a.h = h; a.w = w; a.threshold = threshold; a.refine = refine;
    return launch_1d<KpMaximaElem>(a, stream);
}

// registration.py:141-212 (qtildematrices) for ONE level: src / ref are the level's six complex sub-bands of the (warped)
// source and of the reference image, element (b, band, i, j) at 2*(b*x_n + band*x_band + i*x_row + j*x_col).
// reduce == 0: qt is [n][h][w][27] float64; reduce != 0: qt is [n][27], the sum over the image, ADDED to its content.
int dtcwt_b200_reg_qtilde_f32(const float* src, const float* ref, double* qt, int64_t n, int64_t h, int64_t w, int64_t s_n,
                              int64_t s_band, int64_t s_row, int64_t s_col, int64_t r_n, int64_t r_band, int64_t r_row,
                              int64_t r_col, int reduce, void* stream) {
    return reg_qtilde_impl<float>(src, ref, qt, n, h, w, s_n, s_band, s_row, s_col, r_n, r_band, r_row, r_col, reduce, stream);
}
int dtcwt_b200_reg_qtilde_f64(const double* src, const double* ref, double* qt, int64_t n, int64_t h, int64_t w, int64_t s_n,
                              int64_t s_band, int64_t s_row, int64_t s_col, int64_t r_n, int64_t r_band, int64_t r_row,
                              int64_t r_col, int reduce, void* stream) {
    return reg_qtilde_impl<double>(src, ref, qt, n, h, w, s_n, s_band, s_row, s_col, r_n, r_band, r_row, r_col, reduce, stream);
}

// registration.py:357-362: out (+)= rescale(_boxfilter(qt, 3), (H, W), 'bilinear'); qt [n][h][w][27], out [n][H][W][27]
int dtcwt_b200_reg_boxrescale(const double* qt, double* out, int64_t n, int64_t h, int64_t w, int64_t H, int64_t W,
                              int accumulate, void* stream) {
    if (n < 0 || h < 1 || w < 1 || H < 1 || W < 1) return DTCWT_B200_EINVAL;
    if (n == 0) return DTCWT_B200_OK;
    if (!qt || !out) return DTCWT_B200_EINVAL;
    if (h > 0x3fffffff || w > 0x3fffffff) return DTCWT_B200_EUNSUPPORTED;
    BoxRescaleArgs a;
    a.qt = qt; a.out = out; a.n = n; a.h = h; a.w = w; a.H = H; a.W = W; a.accumulate = accumulate;
    return launch_1d<BoxRescaleElem>(a, stream);
}

// registration.py:214-257 (solvetransform): avecs[i] (+)= solve(triu(Q_i), -q_i) for `count` 27-vectors
int dtcwt_b200_reg_solve(const double* qt, double* avecs, int64_t count, int accumulate, void* stream) {
    if (count < 0) return DTCWT_B200_EINVAL;
    if (count == 0) return DTCWT_B200_OK;
    if (!qt || !avecs) return DTCWT_B200_EINVAL;
    SolveArgs a;
    a.qt = qt; a.avecs = avecs; a.count = count; a.accumulate = accumulate;
    return launch_1d<SolveElem>(a, stream);
}

// registration.py:374-423: avecs [n][H][W][6] -> xs, ys [n][h][w].  mode 0: velocityfield(avecs, (h, w), 'bilinear');
// mode 1: the pixel coordinates warp() / warphighpass() sample at, (X + vx) * w and (Y + vy) * h.
int dtcwt_b200_reg_coords(const double* avecs, double* xs, double* ys, int64_t n, int64_t H, int64_t W, int64_t h, int64_t w,
                          int mode, void* stream) {
    if (n < 0 || H < 1 || W < 1 || h < 1 || w < 1 || mode < 0 || mode > 1) return DTCWT_B200_EINVAL;
    if (n == 0) return DTCWT_B200_OK;
    if (!avecs || !xs || !ys) return DTCWT_B200_EINVAL;
    if (H > 0x3fffffff || W > 0x3fffffff) return DTCWT_B200_EUNSUPPORTED;
    CoordsArgs a;
    a.avecs = avecs; a.xs = xs; a.ys = ys; a.n = n; a.H = H; a.W = W; a.h = h; a.w = w; a.mode = mode;
    return launch_1d<CoordsElem>(a, stream);
}

// sampling.py:105-129 (sample), :131-165 (rescale), :192-222 (sample_highpass), :224-278 (rescale_highpass).
// im element (b, y, x, c) at k*(b*i_n + y*i_y + x*i_x + c*i_c), out likewise with o_*; k = 2 for complex (interleaved).
// method 0 nearest / 1 bilinear / 2 lanczos / 3 upsample()'s 7-tap lanczos (doubled rescale grid only).  coords 0: positions from xs / ys ([oh][ow] planes, coord_n elements apart
// per batch item, 0 = shared); coords 1: the rescale grid.  wx / wy non-NULL (host, C <= 8 entries): the image is complex
// sub-bands, phase un-rolled by exp(-j(wx x + wy y)) before sampling and re-rolled at the sample position.
int dtcwt_b200_sample_f32(const float* im, float* out, const double* xs, const double* ys, int64_t n, int64_t h, int64_t w,
                          int64_t C, int64_t oh, int64_t ow, int64_t i_n, int64_t i_y, int64_t i_x, int64_t i_c, int64_t o_n,
                          int64_t o_y, int64_t o_x, int64_t o_c, int64_t coord_n, int is_complex, int method, int coords,
                          const double* wx, const double* wy, void* stream) {
    return sample_impl<float>(im, out, xs, ys, n, h, w, C, oh, ow, i_n, i_y, i_x, i_c, o_n, o_y, o_x, o_c, coord_n, is_complex,
                              method, coords, wx, wy, stream);
}
int dtcwt_b200_sample_f64(const double* im, double* out, const double* xs, const double* ys, int64_t n, int64_t h, int64_t w,
                          int64_t C, int64_t oh, int64_t ow, int64_t i_n, int64_t i_y, int64_t i_x, int64_t i_c, int64_t o_n,
                          int64_t o_y, int64_t o_x, int64_t o_c, int64_t coord_n, int is_complex, int method, int coords,
                          const double* wx, const double* wy, void* stream) {
    return sample_impl<double>(im, out, xs, ys, n, h, w, C, oh, ow, i_n, i_y, i_x, i_c, o_n, o_y, o_x, o_c, coord_n, is_complex,
                               method, coords, wx, wy, stream);
}

}  // extern "C"
#endif  // DTCWT_EMIT_GENERIC
