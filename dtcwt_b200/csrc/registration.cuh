// Device kernels of the DTCWT-based image registration (reference dtcwt/registration.py) and of the
// re-sampling helpers it needs (reference dtcwt/sampling.py).  One thread per output element; all
// arithmetic in float64 whatever the sub-band storage type (the reference accumulates its Q~ matrices
// in float64, registration.py:181, and the affine solve is ill-conditioned where the confidence is low).
//
//   QtildeElem      confidence (:84-139), phasegradient (:32-76) and the 27-element Q~ vector (:141-212) of one
//                   level, summed over its six sub-bands; optionally reduced over the image (atomics) for the
//                   global first estimate (:333-338)
//   BoxRescaleElem  rescale(_boxfilter(Q~, 3), avecs.shape, 'bilinear') accumulated over levels (:357-362, :425-446)
//   SolveElem       a = -Q^-1 q with ONLY the upper triangle of Q filled, as the reference does (:229-243)
//   CoordsElem      velocityfield (:374-393) of an affine-parameter grid resampled to a target shape, and the
//                   sample coordinates warp / warphighpass (:395-423) feed to the sampler
//   SampleElem      sample / rescale / sample_highpass / rescale_highpass (sampling.py:36-278): nearest, bilinear
//                   and Lanczos-3 taps with symmetric extension, optional phase un-rolling of complex sub-bands
#pragma once
#include "common.cuh"

namespace dtcwt {

// reflect(x, -0.5, n - 0.5).astype(int) of the reference (utils.py:136-153, sampling.py:36-40) for a real x
DTCWT_HD int reflect_coord(double x, int n) {
    const double rng = (double)n, rng2 = 2.0 * rng;
    double mod = fmod(x + 0.5, rng2);
    if (mod < 0) mod += rng2;
    const double out = ((mod >= rng) ? (rng2 - mod) : mod) - 0.5;
    int i = (int)out;                                // truncation toward zero, like ndarray.astype(int)
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
DTCWT_HD int clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

struct Cplx { double re, im; };
DTCWT_HD Cplx cmul(Cplx a, Cplx b) { Cplx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
DTCWT_HD Cplx cconj(Cplx a) { a.im = -a.im; return a; }
DTCWT_HD Cplx cadd(Cplx a, Cplx b) { a.re += b.re; a.im += b.im; return a; }
DTCWT_HD double cabs2(Cplx a) { return a.re * a.re + a.im * a.im; }
DTCWT_HD double cangle(Cplx a) { return atan2(a.im, a.re); }
DTCWT_HD Cplx cexp(double ph) {
    Cplx r;
#if defined(__CUDA_ARCH__)
    sincos(ph, &r.im, &r.re);           // one argument reduction for both
#else
    r.re = cos(ph); r.im = sin(ph);
#endif
    return r;
}
// reflect_coord for a coordinate that is a whole number (every tap position of the samplers is): the same fold in integer
// arithmetic -- no double-precision fmod per tap.  Falls back to reflect_coord for positions beyond the int range.
DTCWT_HD int reflect_whole(double x, int n) {
    if (!(x > -1.0e9 && x < 1.0e9)) return reflect_coord(x, n);
    return reflect_any((int)x, n);
}
// |z|^3 = s sqrt(s), s = |z|^2.  Device: single-precision reciprocal square root as the seed of one Newton step in double
// (relative error ~1e-14) instead of the double-precision square root sequence -- 48 of them per pixel sat in Q~'s inner loop.
DTCWT_HD double cabs3(Cplx a) {
    const double s = cabs2(a);
#if defined(__CUDA_ARCH__)
    if (!(s > 1e-30 && s < 1e30)) return s * sqrt(s);           // zeros, denormals, overflow of the float seed
    double r = (double)rsqrtf((float)s);
    r = r * (1.5 - 0.5 * s * r * r);
    r = r * (1.5 - 0.5 * s * r * r);
    return s * (s * r);
#else
    return s * sqrt(s);
#endif
}

// ------------------------------------------------------------------ Q~ matrices
template <typename T>
struct QtildeArgs {
    const T* src;                   // "t_ref" of the reference's qtildematrices: the (warped) source pyramid level
    const T* ref;                   // "t_target": the reference image's level
    double* qt;                     // reduce == 0: [n][h][w][27]; reduce != 0: [n][27], zeroed by the caller
    int64_t n, h, w;
    int64_t s_n, s_band, s_row, s_col;     // complex strides of src
    int64_t r_n, r_band, r_row, r_col;     // complex strides of ref
    int reduce;
    double shift[6][2];             // EXPECTED_SHIFTS (registration.py:30)
    double rot[6][2][2];            // exp(-i shift[band][axis]) as (re, im): the same for every pixel, prepared by the host
    double epsilon;
};

template <typename T>
struct QtildeElem {
    typedef QtildeArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }

    static DTCWT_HD Cplx ld(const T* p, int64_t off) { Cplx c; c.re = (double)p[2 * off]; c.im = (double)p[2 * off + 1]; return c; }

    // Q += the contribution of one sub-band at pixel (i, j) of image b
    static DTCWT_HD void add_band(const Args& a, int64_t b, int i, int j, int band, double (&Q)[27]) {
        const int h = (int)a.h, w = (int)a.w;
        const double xs = (double)j * (1.0 / (double)w), ys = (double)i * (1.0 / (double)h);
        {
            const T* u = a.src + 2 * (b * a.s_n + band * a.s_band);
            const T* v = a.ref + 2 * (b * a.r_n + band * a.r_band);
            auto U = [&](int y, int x) { return ld(u, (int64_t)y * a.s_row + (int64_t)x * a.s_col); };
            auto V = [&](int y, int x) { return ld(v, (int64_t)y * a.r_row + (int64_t)x * a.r_col); };
            // confidence: the four diagonal neighbours of the edge-replicated sub-bands (registration.py:101-139)
            Cplx num; num.re = 0; num.im = 0;
            double den = a.epsilon;
            for (int dy = -1; dy <= 1; dy += 2)
                for (int dx = -1; dx <= 1; dx += 2) {
                    const int y = clampi(i + dy, h), x = clampi(j + dx, w);
                    const Cplx uu = U(y, x), vv = V(y, x);
                    num = cadd(num, cmul(cconj(uu), vv));
                    den += cabs3(uu) + cabs3(vv);
                }
            const double C = cabs2(num) / den;
            // phase gradients (registration.py:52-76): conjugate products across horizontal / vertical pairs, de-rotated
            // by the expected shift, averaged between the two pairs that straddle the pixel
            Cplx rot0, rot1;
            rot0.re = a.rot[band][0][0]; rot0.im = a.rot[band][0][1];
            rot1.re = a.rot[band][1][0]; rot1.im = a.rot[band][1][1];
            auto Sx = [&](int x) {      // pair (x, x+1) of row i
                return cmul(cadd(cmul(U(i, x + 1), cconj(U(i, x))), cmul(V(i, x + 1), cconj(V(i, x)))), rot0);
            };
            auto Sy = [&](int y) {
                return cmul(cadd(cmul(U(y + 1, j), cconj(U(y, j))), cmul(V(y + 1, j), cconj(V(y, j)))), rot1);
            };
            double dx, dy;
            if (w < 2) dx = 0.0;
            else if (j == 0) dx = cangle(Sx(0));
            else if (j == w - 1) dx = cangle(Sx(w - 2));
            else { Cplx s = cadd(Sx(j - 1), Sx(j)); s.re *= 0.5; s.im *= 0.5; dx = cangle(s); }
            if (h < 2) dy = 0.0;
            else if (i == 0) dy = cangle(Sy(0));
            else if (i == h - 1) dy = cangle(Sy(h - 2));
            else { Cplx s = cadd(Sy(i - 1), Sy(i)); s.re *= 0.5; s.im *= 0.5; dy = cangle(s); }
            dx = (dx + a.shift[band][0]) * (double)w;
            dy = (dy + a.shift[band][1]) * (double)h;
            const double dt = cangle(cmul(V(i, j), cconj(U(i, j))));
            const double tmp[7] = {dx, dy, xs * dx, xs * dy, ys * dx, ys * dy, -dt};
            const double c2 = C * C;
            int e = 0;
            for (int r = 0; r < 6; ++r)
                for (int c = r; c < 6; ++c) Q[e++] += tmp[r] * tmp[c] * c2;
            for (int r = 0; r < 6; ++r) Q[e++] += tmp[r] * tmp[6] * c2;
        }
    }

#if !defined(DTCWT_EMU) && defined(__CUDACC__)
    // Device launch shape (dtcwt_b200.cu: qtilde_kernel): TWO adjacent lanes per pixel with three sub-bands each -- twice the
    // threads for the coarse pyramid levels, whose few hundred CTAs of long dependent float64 chains do not fill the GPU, and
    // half the chain per thread.  The two lanes add their 27 values with one shuffle; each then writes every other element
    // (or feeds the image-wide reduction).  Every thread of the grid takes part; lanes beyond the last pixel carry zeros.
    static __device__ __forceinline__ void run_lanes(const Args& a, int64_t gid2) {
        const int half = (int)(gid2 & 1);
        const int64_t gid = gid2 >> 1;
        const bool live = gid < total(a);
        const int64_t g = live ? gid : 0;
        const int j = (int)(g % a.w);
        const int64_t t = g / a.w;
        const int i = (int)(t % a.h);
        const int64_t b = t / a.h;
        double Q[27];
#pragma unroll
        for (int e = 0; e < 27; ++e) Q[e] = 0.0;
        if (live)
            for (int band = 3 * half; band < 3 * half + 3; ++band) add_band(a, b, i, j, band, Q);
        // (b0 + b1 + b2) + (b3 + b4 + b5): float64, differs from the reference's sequential sum in the last bits only
#pragma unroll
        for (int e = 0; e < 27; ++e) Q[e] += __shfl_xor_sync(0xffffffffu, Q[e], 1);
        if (a.reduce) {
            // image-wide sum: the sixteen pixels of a warp, then one atomic per warp and element
            const unsigned same = __match_any_sync(0xffffffffu, live ? (int)b : -1);
            if (same == 0xffffffffu) {
#pragma unroll
                for (int e = 0; e < 27; ++e) {
                    double v = Q[e];
#pragma unroll
                    for (int off = 2; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if ((threadIdx.x & 31) == 0 && live) atomicAdd(a.qt + b * 27 + e, v);
                }
            } else if (live && half == 0) {
                for (int e = 0; e < 27; ++e) atomicAdd(a.qt + b * 27 + e, Q[e]);
            }
            return;
        }
        if (!live) return;
        double* d = a.qt + gid * 27;
#pragma unroll
        for (int e = 0; e < 27; ++e)
            if ((e & 1) == half) d[e] = Q[e];
    }
#endif

    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int j = (int)(gid % a.w);
        const int64_t t = gid / a.w;
        const int i = (int)(t % a.h);
        const int64_t b = t / a.h;
        double Q[27];
        for (int e = 0; e < 27; ++e) Q[e] = 0.0;
        for (int band = 0; band < 6; ++band) add_band(a, b, i, j, band, Q);
        if (a.reduce) {
#if defined(DTCWT_EMU) || !defined(__CUDA_ARCH__)
            for (int e = 0; e < 27; ++e) a.qt[b * 27 + e] += Q[e];
#else
            // sum over the image: first inside the warp (shuffles), then one atomic per warp and element -- 27 atomics per
            // THREAD on the same 27 addresses made this launch ten times longer than the arithmetic (profiles/r3_01).
            // A warp whose lanes belong to two images, or that is not full, falls back to per-thread atomics.
            const unsigned act = __activemask();
            const unsigned same = __match_any_sync(act, (int)b);
            if (same == 0xffffffffu) {
#pragma unroll
                for (int e = 0; e < 27; ++e) {
                    double v = Q[e];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if ((threadIdx.x & 31) == 0) atomicAdd(a.qt + b * 27 + e, v);
                }
            } else {
                for (int e = 0; e < 27; ++e) atomicAdd(a.qt + b * 27 + e, Q[e]);
            }
#endif
        } else {
            double* d = a.qt + gid * 27;
            for (int e = 0; e < 27; ++e) d[e] = Q[e];
        }
    }
};

// ------------------------------------------------------------------ box filter + bilinear rescale of the Q~ field
struct BoxRescaleArgs {
    const double* qt;               // [n][h][w][27]
    double* out;                    // [n][H][W][27]
    int64_t n, h, w, H, W;
    int accumulate;
};

struct BoxRescaleElem {
    typedef BoxRescaleArgs Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.H * a.W * 27; }
    // _boxfilter(X, 3) at (y, x), element e (registration.py:425-446): mean over the 3 x 3 symmetric-extended patch
    static DTCWT_HD double box(const Args& a, int64_t b, int y, int x, int e) {
        const int h = (int)a.h, w = (int)a.w;
        double rows = 0.0;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = clampi(y + dy, h);
            const double* p = a.qt + ((b * a.h + yy) * a.w) * 27 + e;
            rows += (p[(int64_t)clampi(x - 1, w) * 27] + p[(int64_t)x * 27] + p[(int64_t)clampi(x + 1, w) * 27]) / 3.0;
        }
        return rows / 3.0;
    }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int e = (int)(gid % 27);
        int64_t t = gid / 27;
        const int X = (int)(t % a.W);
        t /= a.W;
        const int Y = (int)(t % a.H);
        const int64_t b = t / a.H;
        const double sx = ((double)a.w / (double)a.W) * ((double)X + 0.5) - 0.5;     // sampling.py:155-161
        const double sy = ((double)a.h / (double)a.H) * ((double)Y + 0.5) - 0.5;
        const double fx0 = floor(sx), fy0 = floor(sy), fx = sx - fx0, fy = sy - fy0;
        // the rescale grid reaches at most one sample outside the array, where the half-sample symmetric extension
        // (reflect_coord) is a clamp: no fmod per element
        const int x0 = clampi((int)fx0, (int)a.w), x1 = clampi((int)fx0 + 1, (int)a.w);
        const int y0 = clampi((int)fy0, (int)a.h), y1 = clampi((int)fy0 + 1, (int)a.h);
        double b00, b01, b10, b11;
        if (x1 == x0 + 1 && y1 == y0 + 1) {
            // the four 3 x 3 windows overlap in a 4 x 4 patch: 16 loads instead of 36, the same additions in the same order
            const int h = (int)a.h, w = (int)a.w;
            const int64_t c0 = (int64_t)clampi(x0 - 1, w) * 27, c1 = (int64_t)x0 * 27, c2 = (int64_t)x1 * 27, c3 = (int64_t)clampi(x1 + 1, w) * 27;
            const int ys[4] = {clampi(y0 - 1, h), y0, y1, clampi(y1 + 1, h)};
            double ra[4], rb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double* p = a.qt + ((b * a.h + ys[r]) * a.w) * 27 + e;
                const double q0 = p[c0], q1 = p[c1], q2 = p[c2], q3 = p[c3];
                ra[r] = (q0 + q1 + q2) / 3.0;
                rb[r] = (q1 + q2 + q3) / 3.0;
            }
            b00 = (ra[0] + ra[1] + ra[2]) / 3.0; b01 = (rb[0] + rb[1] + rb[2]) / 3.0;
            b10 = (ra[1] + ra[2] + ra[3]) / 3.0; b11 = (rb[1] + rb[2] + rb[3]) / 3.0;
        } else {
            b00 = box(a, b, y0, x0, e); b01 = box(a, b, y0, x1, e);
            b10 = box(a, b, y1, x0, e); b11 = box(a, b, y1, x1, e);
        }
        const double lower = (1.0 - fx) * b00 + fx * b01;
        const double upper = (1.0 - fx) * b10 + fx * b11;
        const double v = (1.0 - fy) * lower + fy * upper;
        if (a.accumulate) a.out[gid] += v; else a.out[gid] = v;
    }
};

// ------------------------------------------------------------------ affine parameters from Q~
struct SolveArgs {
    const double* qt;               // [count][27]
    double* avecs;                  // [count][6]
    int64_t count;
    int accumulate;
};

struct SolveElem {
    typedef SolveArgs Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.count; }
    // registration.py:229-243: Q holds the 21 upper-triangle elements and ZEROS below the diagonal, then solve(Q, -q)
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const double* v = a.qt + gid * 27;
        double Q[6][6], x[6];
        int e = 0;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Q[r][c] = (c >= r) ? v[e++] : 0.0;
        for (int r = 5; r >= 0; --r) {
            double s = -v[21 + r];
            for (int c = r + 1; c < 6; ++c) s -= Q[r][c] * x[c];
            x[r] = s / Q[r][r];
        }
        double* d = a.avecs + gid * 6;
        for (int r = 0; r < 6; ++r) d[r] = a.accumulate ? d[r] + x[r] : x[r];
    }
};

// ------------------------------------------------------------------ velocity field / warp coordinates
struct CoordsArgs {
    const double* avecs;            // [n][H][W][6]
    double* xs;                     // [n][h][w]
    double* ys;
    int64_t n, H, W, h, w;
    int mode;                       // 0: velocity field (vx, vy) in normalised units; 1: sample coordinates in pixels
};

struct CoordsElem {
    typedef CoordsArgs Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }
    static DTCWT_HD void field(const Args& a, int64_t b, int Y, int X, double& vx, double& vy) {
        const double* v = a.avecs + ((b * a.H + Y) * a.W + X) * 6;
        const double px = (double)X / (double)a.W, py = (double)Y / (double)a.H;      // registration.py:385-389
        vx = v[0] + v[2] * px + v[4] * py;
        vy = v[1] + v[3] * px + v[5] * py;
    }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int x = (int)(gid % a.w);
        const int64_t t = gid / a.w;
        const int y = (int)(t % a.h);
        const int64_t b = t / a.h;
        // rescale(..., shape, 'bilinear') of the two velocity components (sampling.py:131-165)
        const double sx = ((double)a.W / (double)a.w) * ((double)x + 0.5) - 0.5;
        const double sy = ((double)a.H / (double)a.h) * ((double)y + 0.5) - 0.5;
        const double fx0 = floor(sx), fy0 = floor(sy), fx = sx - fx0, fy = sy - fy0;
        const int x0 = reflect_coord(fx0, (int)a.W), x1 = reflect_coord(fx0 + 1.0, (int)a.W);
        const int y0 = reflect_coord(fy0, (int)a.H), y1 = reflect_coord(fy0 + 1.0, (int)a.H);
        double ax, ay, bx, by, cx, cy, dx, dy;
        field(a, b, y0, x0, ax, ay); field(a, b, y0, x1, bx, by);
        field(a, b, y1, x0, cx, cy); field(a, b, y1, x1, dx, dy);
        const double vx = (1.0 - fy) * ((1.0 - fx) * ax + fx * bx) + fy * ((1.0 - fx) * cx + fx * dx);
        const double vy = (1.0 - fy) * ((1.0 - fx) * ay + fx * by) + fy * ((1.0 - fx) * cy + fx * dy);
        if (a.mode == 0) {
            a.xs[gid] = vx; a.ys[gid] = vy;
        } else {                     // registration.py:401-404, 416-423: (X + vx) * width, (Y + vy) * height
            a.xs[gid] = ((double)x / (double)a.w + vx) * (double)a.w;
            a.ys[gid] = ((double)y / (double)a.h + vy) * (double)a.h;
        }
    }
};

// ------------------------------------------------------------------ sampling
// kSampleLanczosUp: the reference's upsample() (sampling.py:280-370) convolves with SEVEN un-windowed Lanczos taps at
// offsets -3..3 around the source pixel, one more than sample()'s six; only valid on the doubled rescale grid
constexpr int kSampleNearest = 0, kSampleBilinear = 1, kSampleLanczos = 2, kSampleLanczosUp = 3;

template <typename T>
struct SampleArgs {
    const T* im;                    // (b, y, x, c) at ncomp * (b*i_n + y*i_y + x*i_x + c*i_c)
    T* out;                         // (b, y, x, c) at ncomp * (b*o_n + y*o_y + x*o_x + c*o_c)
    const double* xs;               // coords == 0: [nc][oh][ow] sample positions (nc = 1: shared by the batch)
    const double* ys;
    int64_t n, h, w, C, oh, ow;
    int64_t i_n, i_y, i_x, i_c, o_n, o_y, o_x, o_c;
    int64_t coord_n;                // elements between the coordinate planes of consecutive batch items (0 = shared)
    int ncomp;                      // 1 real, 2 complex (interleaved)
    int method;
    int coords;                     // 0: xs / ys arrays; 1: rescale to [oh][ow] (sampling.py:155-161)
    int phase;                      // sample_highpass / rescale_highpass (sampling.py:192-278): un-roll, sample, re-roll
    double wx[8], wy[8];            // phase advance per channel (DTHETA_DX_2D / DTHETA_DY_2D of the selected sub-bands)
    double rx[8][2], ry[8][2];      // exp(-i wx[c]), exp(-i wy[c]) as (re, im), prepared by the host
};

template <typename T>
struct SampleElem {
    typedef SampleArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.oh * a.ow * a.C; }

    static DTCWT_HD double lanczos(double x) {      // np.sinc(x) * np.sinc(x / 3)
        if (x == 0.0) return 1.0;
        const double pi = 3.14159265358979323846, px = pi * x;
        return (sin(px) / px) * (sin(px / 3.0) / (px / 3.0));
    }

    static DTCWT_HD Cplx raw(const Args& a, int64_t b, int c, int x, int y) {      // the stored sample, no phase handling
        const T* p = a.im + a.ncomp * (b * a.i_n + (int64_t)y * a.i_y + (int64_t)x * a.i_x + (int64_t)c * a.i_c);
        Cplx v;
        v.re = (double)p[0];
        v.im = a.ncomp == 2 ? (double)p[1] : 0.0;
        return v;
    }

    static DTCWT_HD Cplx tap(const Args& a, int64_t b, int c, double fx, double fy) {
        const int x = reflect_whole(fx, (int)a.w), y = reflect_whole(fy, (int)a.h);       // fx, fy are whole numbers
        const T* p = a.im + a.ncomp * (b * a.i_n + (int64_t)y * a.i_y + (int64_t)x * a.i_x + (int64_t)c * a.i_c);
        Cplx v;
        v.re = (double)p[0];
        v.im = a.ncomp == 2 ? (double)p[1] : 0.0;
        if (a.phase) v = cmul(v, cexp(-(a.wx[c] * (double)x + a.wy[c] * (double)y)));
        return v;
    }

    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int c = (int)(gid % a.C);
        int64_t t = gid / a.C;
        const int X = (int)(t % a.ow);
        t /= a.ow;
        const int Y = (int)(t % a.oh);
        const int64_t b = t / a.oh;
        double sx, sy;
        if (a.coords == 1) {
            sx = ((double)a.w / (double)a.ow) * ((double)X + 0.5) - 0.5;
            sy = ((double)a.h / (double)a.oh) * ((double)Y + 0.5) - 0.5;
        } else {
            const int64_t o = b * a.coord_n + (int64_t)Y * a.ow + X;
            sx = a.xs[o]; sy = a.ys[o];
        }
        Cplx acc; acc.re = 0; acc.im = 0;
        if (a.method == kSampleNearest) {
            acc = tap(a, b, c, rint(sx), rint(sy));                      // np.round: half to even
        } else if (a.method == kSampleBilinear) {
            const double fx0 = floor(sx), fy0 = floor(sy), fx = sx - fx0, fy = sy - fy0;
            Cplx p00, p10, p01, p11;
            if (a.phase) {
                // phase un-rolling of the four taps: one sincos for the first, the neighbours by the angle-addition
                // theorem (a complex multiply with the per-channel constants exp(-i wx), exp(-i wy)) wherever the mirrored
                // coordinates are still adjacent -- everywhere except on the image border
                const int x0 = reflect_whole(fx0, (int)a.w), x1 = reflect_whole(fx0 + 1.0, (int)a.w);
                const int y0 = reflect_whole(fy0, (int)a.h), y1 = reflect_whole(fy0 + 1.0, (int)a.h);
                Cplx rx, ry;
                rx.re = a.rx[c][0]; rx.im = a.rx[c][1]; ry.re = a.ry[c][0]; ry.im = a.ry[c][1];
                const Cplx e00 = cexp(-(a.wx[c] * (double)x0 + a.wy[c] * (double)y0));
                const Cplx e10 = (x1 == x0 + 1) ? cmul(e00, rx) : cexp(-(a.wx[c] * (double)x1 + a.wy[c] * (double)y0));
                const Cplx e01 = (y1 == y0 + 1) ? cmul(e00, ry) : cexp(-(a.wx[c] * (double)x0 + a.wy[c] * (double)y1));
                const Cplx e11 = (y1 == y0 + 1) ? cmul(e10, ry) : cexp(-(a.wx[c] * (double)x1 + a.wy[c] * (double)y1));
                p00 = cmul(raw(a, b, c, x0, y0), e00); p10 = cmul(raw(a, b, c, x1, y0), e10);
                p01 = cmul(raw(a, b, c, x0, y1), e01); p11 = cmul(raw(a, b, c, x1, y1), e11);
            } else {
                p00 = tap(a, b, c, fx0, fy0); p10 = tap(a, b, c, fx0 + 1.0, fy0);
                p01 = tap(a, b, c, fx0, fy0 + 1.0); p11 = tap(a, b, c, fx0 + 1.0, fy0 + 1.0);
            }
            const double lr = (1.0 - fx) * p00.re + fx * p10.re, li = (1.0 - fx) * p00.im + fx * p10.im;
            const double ur = (1.0 - fx) * p01.re + fx * p11.re, ui = (1.0 - fx) * p01.im + fx * p11.im;
            acc.re = (1.0 - fy) * lr + fy * ur;
            acc.im = (1.0 - fy) * li + fy * ui;
        } else if (a.method == kSampleLanczosUp) {
            const double ox = (X & 1) ? 0.25 : -0.25, oy = (Y & 1) ? 0.25 : -0.25;      // sampling.py:312-320
            for (int dx = -3; dx <= 3; ++dx) {
                const double Lx = lanczos(ox - (double)dx);
                for (int dy = -3; dy <= 3; ++dy) {
                    const double wgt = Lx * lanczos(oy - (double)dy);
                    const Cplx p = tap(a, b, c, (double)((X >> 1) + dx), (double)((Y >> 1) + dy));
                    acc.re += wgt * p.re; acc.im += wgt * p.im;
                }
            }
        } else {
            const double fx0 = floor(sx), fy0 = floor(sy), fx = sx - fx0, fy = sy - fy0;
            for (int dx = -2; dx <= 3; ++dx) {
                const double Lx = lanczos(fx - (double)dx);
                for (int dy = -2; dy <= 3; ++dy) {
                    const double wgt = Lx * lanczos(fy - (double)dy);
                    const Cplx p = tap(a, b, c, fx0 + (double)dx, fy0 + (double)dy);
                    acc.re += wgt * p.re; acc.im += wgt * p.im;
                }
            }
        }
        if (a.phase) acc = cmul(acc, cexp(a.wx[c] * sx + a.wy[c] * sy));
        T* d = a.out + a.ncomp * (b * a.o_n + (int64_t)Y * a.o_y + (int64_t)X * a.o_x + (int64_t)c * a.o_c);
        d[0] = (T)acc.re;
        if (a.ncomp == 2) d[1] = (T)acc.im;
    }
};

}  // namespace dtcwt
