// Device kernels of dtcwt.keypoint.find_keypoints (reference dtcwt/keypoint.py): the keypoint-energy maps of one
// level's six sub-bands and the local-maximum search with sub-pixel refinement.  One thread per pixel, float64.
#pragma once
#include "common.cuh"

namespace dtcwt {

constexpr int kKpFauqueur = 0, kKpBendale = 1, kKpKingsbury = 2;

template <typename T>
struct KpEnergyArgs {
    const T* yh;                    // complex, element (b, band, i, j) at 2*(b*s_n + band*s_band + i*s_row + j*s_col)
    double* e;                      // [n][h][w]
    int64_t n, h, w, s_n, s_band, s_row, s_col;
    int method;
    double scale_gain;              // fauqueur: alpha ** (scale + 1)
    double beta, kappa;
};

template <typename T>
struct KpEnergyElem {
    typedef KpEnergyArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t j = gid % a.w, t = gid / a.w, i = t % a.h, b = t / a.h;
        double m[6];
        for (int band = 0; band < 6; ++band) {
            const T* p = a.yh + 2 * (b * a.s_n + band * a.s_band + i * a.s_row + j * a.s_col);
            m[band] = sqrt((double)p[0] * (double)p[0] + (double)p[1] * (double)p[1]);
        }
        double e;
        if (a.method == kKpFauqueur) {              // keypoint.py:143-144
            double prod = 1.0;
            for (int band = 0; band < 6; ++band) prod *= m[band];
            e = a.scale_gain * pow(prod > 0.0 ? prod : 0.0, a.beta);
        } else if (a.method == kKpBendale) {        // :146-147
            e = m[0];
            for (int band = 1; band < 6; ++band) e = m[band] < e ? m[band] : e;
        } else {                                    // kingsbury, :149-156
            double A = 0.0, B = 0.0;
            for (int band = 0; band < 6; ++band) A += m[band] * m[band];
            A = sqrt(A);
            for (int band = 0; band < 3; ++band) B += m[band] * m[band + 3];
            e = B / (A > 1e-8 ? A : 1e-8) - a.kappa * A;
            e = e > 0.0 ? e : 0.0;
        }
        a.e[gid] = e;
    }
};

struct KpMaximaArgs {
    const double* x;                // energy map [n][h][w]
    double* out;                    // [n][h][w][4]: flag (1 = keypoint), refined row, refined column, value
    int64_t n, h, w;
    double threshold;
    int refine;
};

struct KpMaximaElem {
    typedef KpMaximaArgs Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }
    // numpy.gradient along one axis: central differences inside, one-sided at the two ends
    static DTCWT_HD double X(const Args& a, const double* im, int r, int c) { return im[(int64_t)r * a.w + c]; }
    static DTCWT_HD double gx(const Args& a, const double* im, int r, int c) {
        const int w = (int)a.w;
        if (w < 2) return 0.0;
        if (c == 0) return X(a, im, r, 1) - X(a, im, r, 0);
        if (c == w - 1) return X(a, im, r, w - 1) - X(a, im, r, w - 2);
        return 0.5 * (X(a, im, r, c + 1) - X(a, im, r, c - 1));
    }
    static DTCWT_HD double gy(const Args& a, const double* im, int r, int c) {
        const int h = (int)a.h;
        if (h < 2) return 0.0;
        if (r == 0) return X(a, im, 1, c) - X(a, im, 0, c);
        if (r == h - 1) return X(a, im, h - 1, c) - X(a, im, h - 2, c);
        return 0.5 * (X(a, im, r + 1, c) - X(a, im, r - 1, c));
    }
    struct GX { DTCWT_HD double operator()(const Args& a, const double* im, int r, int c) const { return gx(a, im, r, c); } };
    struct GY { DTCWT_HD double operator()(const Args& a, const double* im, int r, int c) const { return gy(a, im, r, c); } };
    // second derivatives = numpy.gradient of the gradient images (keypoint.py:229-231)
    template <class G>
    static DTCWT_HD double d_dx(const Args& a, const double* im, int r, int c, G g) {
        const int w = (int)a.w;
        if (w < 2) return 0.0;
        if (c == 0) return g(a, im, r, 1) - g(a, im, r, 0);
        if (c == w - 1) return g(a, im, r, w - 1) - g(a, im, r, w - 2);
        return 0.5 * (g(a, im, r, c + 1) - g(a, im, r, c - 1));
    }
    template <class G>
    static DTCWT_HD double d_dy(const Args& a, const double* im, int r, int c, G g) {
        const int h = (int)a.h;
        if (h < 2) return 0.0;
        if (r == 0) return g(a, im, 1, c) - g(a, im, 0, c);
        if (r == h - 1) return g(a, im, h - 1, c) - g(a, im, h - 2, c);
        return 0.5 * (g(a, im, r + 1, c) - g(a, im, r - 1, c));
    }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int c = (int)(gid % a.w);
        const int64_t t = gid / a.w;
        const int r = (int)(t % a.h);
        const int64_t b = t / a.h;
        const double* im = a.x + b * a.h * a.w;
        double* o = a.out + gid * 4;
        o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0;
        // keypoint.py:207-218: the 3x3 maximum is only formed on rows 1..h-3 and columns 1..w-3
        double mx = a.threshold;
        if (r >= 1 && r < a.h - 2 && c >= 1 && c < a.w - 2)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const double v = X(a, im, r + dy, c + dx);
                    mx = v > mx ? v : mx;
                }
        const double v0 = X(a, im, r, c);
        if (!(mx == v0)) return;
        double x = 0.0, y = 0.0, val = v0;
        if (a.refine) {
            // quadratic patch (keypoint.py:221-252): [2 a0, a2; a2, 2 a1] (x, y) = -(a3, a4), the null vector of the
            // reference's 2 x 3 system normalised to a unit third component
            const double a0 = d_dx(a, im, r, c, GX()), a1 = d_dy(a, im, r, c, GY()), a2 = d_dy(a, im, r, c, GX());
            const double a3 = gx(a, im, r, c), a4 = gy(a, im, r, c);
            const double det = 4.0 * a0 * a1 - a2 * a2;
            x = (-2.0 * a1 * a3 + a2 * a4) / det;
            y = (a2 * a3 - 2.0 * a0 * a4) / det;
            if (fabs(x) > 0.5 || fabs(y) > 0.5) return;
            val = a0 * x * x + a1 * y * y + a2 * x * y + a3 * x + a4 * y + v0;
        }
        o[0] = 1.0; o[1] = (double)r + y; o[2] = (double)c + x; o[3] = val;
    }
};

}  // namespace dtcwt
