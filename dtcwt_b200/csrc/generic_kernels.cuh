// Generic (any size, any tap count <= 32, float or double) one-thread-per-output
// kernels.  They are the correctness baseline of the library and the path taken
// for shapes / wavelets the fused kernels do not cover (tiny images where the
// reflection wraps more than once, float64, 1-D and 3-D transforms, custom taps).
//
// Index maps (verified against the reference for every
// shipped wavelet):
//   colfilter  Y[i]  = sum_k h[k]  X[refl(i + m-1-k - m/2)]        lowlevel.py:69-78
//   coldfilt   Ya[i] = sum_j ha[j] X[refl(4i + m   - 2j)]          lowlevel.py:131-152
//              Yb[i] = sum_j hb[j] X[refl(4i + m+1 - 2j)]
//   colifilt   Y[4i+ph] = sum_k f_ph[2k+tp_ph] X[refl(2i + m/2 - 2k + off_ph)]   lowlevel.py:205-258
#pragma once
#include "common.cuh"

namespace dtcwt {

// ------------------------------------------------------------------ colfilter
template <typename T>
struct ColfilterArgs {
    const T* x;
    T* y;
    int64_t outer, inner;
    int len, pad_lo, L, Lout;   // stored length, padding, logical length, output length
    int accumulate;
    Taps<T> h;
};

template <typename T>
struct ColfilterElem {
    typedef ColfilterArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * a.Lout * a.inner; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t c = gid % a.inner;
        const int64_t t = gid / a.inner;
        const int i = (int)(t % a.Lout);
        const int64_t o = t / a.Lout;
        const T* xo = a.x + o * (int64_t)a.len * a.inner + c;
        const int m = a.h.m;
        T acc = a.accumulate ? a.y[gid] : T(0);
        const int base = i + (m - 1) - m / 2;
        for (int k = 0; k < m; ++k) {
            const int s = unpad(reflect_any(base - k, a.L), a.pad_lo, a.len);
            acc = fma_t<T>(a.h.v[k], xo[(int64_t)s * a.inner], acc);
        }
        a.y[gid] = acc;
    }
};

// ------------------------------------------------------------------ coldfilt
template <typename T>
struct ColdfiltArgs {
    const T* x;
    T* y;
    int64_t outer, inner;
    int len, pad_lo, L, Lout;
    int accumulate;
    int pos;            // sum(ha*hb) > 0: (Ya, Yb) interleave order, else (Yb, Ya)
    Taps<T> ha, hb;
};

template <typename T>
struct ColdfiltElem {
    typedef ColdfiltArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * a.Lout * a.inner; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t c = gid % a.inner;
        const int64_t t = gid / a.inner;
        const int i2 = (int)(t % a.Lout);
        const int64_t o = t / a.Lout;
        const T* xo = a.x + o * (int64_t)a.len * a.inner + c;
        const int m = a.ha.m;
        const int i = i2 >> 1;
        const bool use_a = ((i2 & 1) == 0) == (a.pos != 0);
        const T* f = use_a ? a.ha.v : a.hb.v;
        const int base = 4 * i + m + (use_a ? 0 : 1);
        T acc = a.accumulate ? a.y[gid] : T(0);
        for (int j = 0; j < m; ++j) {
            const int s = unpad(reflect_any(base - 2 * j, a.L), a.pad_lo, a.len);
            acc = fma_t<T>(f[j], xo[(int64_t)s * a.inner], acc);
        }
        a.y[gid] = acc;
    }
};

// ------------------------------------------------------------------ colifilt
template <typename T>
struct ColifiltArgs {
    const T* x;
    T* y;
    int64_t outer, inner;
    int len, crop, Lout;   // Lout = 2*len - 2*crop
    int accumulate;
    int tp[4];             // tap parity used by output phase ph
    int off[4];            // input offset used by output phase ph
    Taps<T> ha, hb;
};

template <typename T>
struct ColifiltElem {
    typedef ColifiltArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * a.Lout * a.inner; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t c = gid % a.inner;
        const int64_t t = gid / a.inner;
        const int q = (int)(t % a.Lout) + a.crop;     // logical output index
        const int64_t o = t / a.Lout;
        const T* xo = a.x + o * (int64_t)a.len * a.inner + c;
        const int m2 = a.ha.m >> 1;
        const int i = q >> 2, ph = q & 3;
        const T* f = (ph & 1) ? a.hb.v : a.ha.v;
        const int tp = a.tp[ph];
        const int base = 2 * i + m2 + a.off[ph];
        T acc = a.accumulate ? a.y[gid] : T(0);
        for (int k = 0; k < m2; ++k) {
            const int s = reflect_any(base - 2 * k, a.len);
            acc = fma_t<T>(f[2 * k + tp], xo[(int64_t)s * a.inner], acc);
        }
        a.y[gid] = acc;
    }
};

// ------------------------------------------------------------------ q2c / c2q
template <typename T>
struct QuadArgs {
    const T* src;
    T* dst;
    int64_t n, h, w;                       // complex sub-band size; the real image is [n][2h][2w]
    int64_t zs_n, zs_band, zs_row, zs_col; // complex strides
    int band0, band1;
    T g0, g1;                              // c2q: gain/sqrt(2) per band; q2c: unused
};

// a b / c d -> z0 = ((a-d) + j(b+c))/sqrt2, z1 = ((a+d) + j(b-c))/sqrt2   (transform2d.py:301-322)
template <typename T>
struct Q2cElem {
    typedef QuadArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t j = gid % a.w;
        const int64_t t = gid / a.w;
        const int64_t i = t % a.h;
        const int64_t b = t / a.h;
        const T* y = a.src + (b * 2 * a.h + 2 * i) * (2 * a.w) + 2 * j;
        const T s = T(0.70710678118654752440);
        const T A = y[0] * s, B = y[1] * s, C = y[2 * a.w] * s, D = y[2 * a.w + 1] * s;
        const int64_t e = b * a.zs_n + i * a.zs_row + j * a.zs_col;
        T* z0 = a.dst + 2 * (e + a.band0 * a.zs_band);
        T* z1 = a.dst + 2 * (e + a.band1 * a.zs_band);
        z0[0] = A - D; z0[1] = B + C;
        z1[0] = A + D; z1[1] = B - C;
    }
};

// P = w0 g0 + w1 g1, Q = w0 g0 - w1 g1 (gains pre-scaled by 1/sqrt2):
// a = Re P, b = Im P, c = Im Q, d = -Re Q                                 (transform2d.py:324-350)
template <typename T>
struct C2qElem {
    typedef QuadArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.n * a.h * a.w; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t j = gid % a.w;
        const int64_t t = gid / a.w;
        const int64_t i = t % a.h;
        const int64_t b = t / a.h;
        const int64_t e = b * a.zs_n + i * a.zs_row + j * a.zs_col;
        const T* z0 = a.src + 2 * (e + a.band0 * a.zs_band);
        const T* z1 = a.src + 2 * (e + a.band1 * a.zs_band);
        const T r0 = z0[0] * a.g0, i0 = z0[1] * a.g0, r1 = z1[0] * a.g1, i1 = z1[1] * a.g1;
        T* y = a.dst + (b * 2 * a.h + 2 * i) * (2 * a.w) + 2 * j;
        y[0] = r0 + r1;
        y[1] = i0 + i1;
        y[2 * a.w] = i0 - i1;
        y[2 * a.w + 1] = r1 - r0;
    }
};

// ------------------------------------------------------------------ 1-D pack / unpack
template <typename T>
struct Pack1dArgs {
    const T* src;
    T* dst;
    int64_t outer, k, inner;
    T gain;
};

// z[o][i][c] = hi[o][2i][c] + j hi[o][2i+1][c]                          (transform1d.py:86-88)
template <typename T>
struct Pack1dElem {
    typedef Pack1dArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * a.k * a.inner; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t c = gid % a.inner;
        const int64_t t = gid / a.inner;
        const int64_t i = t % a.k;
        const int64_t o = t / a.k;
        const T* hi = a.src + (o * 2 * a.k + 2 * i) * a.inner + c;
        a.dst[2 * gid] = hi[0];
        a.dst[2 * gid + 1] = hi[a.inner];
    }
};

// inverse of the above, scaled by the level's gain                       (transform1d.py:161,186-196)
template <typename T>
struct Unpack1dElem {
    typedef Pack1dArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& a) { return a.outer * a.k * a.inner; }
    static DTCWT_HD void run(const Args& a, int64_t gid) {
        const int64_t c = gid % a.inner;
        const int64_t t = gid / a.inner;
        const int64_t i = t % a.k;
        const int64_t o = t / a.k;
        T* hi = a.dst + (o * 2 * a.k + 2 * i) * a.inner + c;
        hi[0] = a.src[2 * gid] * a.gain;
        hi[a.inner] = a.src[2 * gid + 1] * a.gain;
    }
};

// ------------------------------------------------------------------ cube2c / c2cube
template <typename T>
struct CubeArgs {
    const T* src;
    T* dst;
    int64_t n, a, b, c;                          // complex sub-band size; the real cube is [n][2a][2b][2c]
    int64_t zs_n, zs_chan, zs_0, zs_1, zs_2;     // complex strides
    int chan0;
};

// octet corners (transform3d.py:555-568):
//   A=y[0,0,0] B=y[0,1,0] C=y[1,0,0] D=y[1,1,0] E=y[0,0,1] F=y[0,1,1] G=y[1,0,1] H=y[1,1,1]
template <typename T>
struct Cube2cElem {
    typedef CubeArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& g) { return g.n * g.a * g.b * g.c; }
    static DTCWT_HD void run(const Args& g, int64_t gid) {
        const int64_t k = gid % g.c;
        int64_t t = gid / g.c;
        const int64_t j = t % g.b;
        t /= g.b;
        const int64_t i = t % g.a;
        const int64_t v = t / g.a;
        const int64_t s1 = 2 * g.c, s0 = 4 * g.b * g.c;
        const T* y = g.src + v * 2 * g.a * s0 + 2 * i * s0 + 2 * j * s1 + 2 * k;
        const T A = y[0], E = y[1], B = y[s1], F = y[s1 + 1];
        const T C = y[s0], G = y[s0 + 1], D = y[s0 + s1], H = y[s0 + s1 + 1];
        const T hf = T(0.5);
        T* z = g.dst + 2 * (v * g.zs_n + g.chan0 * g.zs_chan + i * g.zs_0 + j * g.zs_1 + k * g.zs_2);
        const int64_t cs = 2 * g.zs_chan;
        z[0]          = (A - G - D - F) * hf;  z[1]          = (B - H + C + E) * hf;    // p
        z[cs]         = (A - G + D + F) * hf;  z[cs + 1]     = (-B + H + C + E) * hf;   // q
        z[2 * cs]     = (A + G + D - F) * hf;  z[2 * cs + 1] = (B + H - C + E) * hf;    // r
        z[3 * cs]     = (A + G - D + F) * hf;  z[3 * cs + 1] = (-B - H - C + E) * hf;   // s
    }
};

template <typename T>
struct C2cubeElem {
    typedef CubeArgs<T> Args;
    static DTCWT_HD int64_t total(const Args& g) { return g.n * g.a * g.b * g.c; }
    static DTCWT_HD void run(const Args& g, int64_t gid) {
        const int64_t k = gid % g.c;
        int64_t t = gid / g.c;
        const int64_t j = t % g.b;
        t /= g.b;
        const int64_t i = t % g.a;
        const int64_t v = t / g.a;
        const T* z = g.src + 2 * (v * g.zs_n + g.chan0 * g.zs_chan + i * g.zs_0 + j * g.zs_1 + k * g.zs_2);
        const int64_t cs = 2 * g.zs_chan;
        const T pr = z[0], pi = z[1], qr = z[cs], qi = z[cs + 1];
        const T rr = z[2 * cs], ri = z[2 * cs + 1], sr = z[3 * cs], si = z[3 * cs + 1];
        const int64_t s1 = 2 * g.c, s0 = 4 * g.b * g.c;
        T* y = g.dst + v * 2 * g.a * s0 + 2 * i * s0 + 2 * j * s1 + 2 * k;
        const T hf = T(0.5);
        y[0]           = (pr + qr + rr + sr) * hf;    // A
        y[s0 + 1]      = (-pr - qr + rr + sr) * hf;   // G
        y[s0 + s1]     = (-pr + qr + rr - sr) * hf;   // D
        y[s1 + 1]      = (-pr + qr - rr + sr) * hf;   // F
        y[s1]          = (pi - qi + ri - si) * hf;    // B
        y[s0 + s1 + 1] = (-pi + qi + ri - si) * hf;   // H
        y[s0]          = (pi + qi - ri - si) * hf;    // C
        y[1]           = (pi + qi + ri + si) * hf;    // E
    }
};

}  // namespace dtcwt
