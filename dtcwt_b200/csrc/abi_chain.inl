// Levels 1 and 2 of the 2-D transform chained image by image, with the level-1 lowpass kept in L2.
//
// Per-level launches over a whole batch send LoLo1 (as large as the input) through HBM twice per direction: 16 of the
// 61 B/pixel a 4-level forward + inverse moves.  Here a CHUNK of images goes through level 1
// and level 2 back to back, LoLo1 of the chunk lives in a scratch buffer that is reused by every chunk and is marked
// L2-persisting (stream access-policy window) -- it is written and read inside the 126 MB L2 and never needs to reach
// DRAM.  The kernels are the per-level ones; only their launch order and the cache policy of one buffer change, so the
// results are bit-identical to the per-level path.
//
// Emitted with the generic group; calls the per-level entry points through the C ABI.
#ifdef DTCWT_EMIT_GENERIC
#ifndef DTCWT_EMU
#include <cuda_runtime.h>
#endif

namespace dtcwt {

#ifndef DTCWT_EMU
// [base, base + bytes) becomes L2-persisting for the kernels launched on `stream` from now on; bytes == 0 ends it.
// Returns the number of bytes the device will actually keep (0: persistence unavailable, the chain still runs).
static int64_t l2_window(void* base, int64_t bytes, double hit_ratio, void* stream) {
    int dev = 0, max_persist = 0, max_window = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    if (bytes <= 0 || max_persist <= 0 || max_window <= 0) {
        v.accessPolicyWindow.base_ptr = nullptr;
        v.accessPolicyWindow.num_bytes = 0;
        v.accessPolicyWindow.hitRatio = 0.f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v);
        cudaGetLastError();
        return 0;
    }
    size_t want = (size_t)bytes < (size_t)max_persist ? (size_t)bytes : (size_t)max_persist;
    size_t have = 0;
    cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
    if (have < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
    const size_t win = (size_t)bytes < (size_t)max_window ? (size_t)bytes : (size_t)max_window;
    double ratio = hit_ratio;
    if ((double)have < ratio * (double)win) ratio = (double)have / (double)win;      // never ask for more lines than the set-aside holds
    v.accessPolicyWindow.base_ptr = base;
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = (float)ratio;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (int64_t)((double)win * ratio);
}
#else
static int64_t l2_window(void*, int64_t, double, void*) { return 0; }
#endif

}  // namespace dtcwt

extern "C" {

// diagnostics: {max persisting bytes, max window bytes, L2 bytes} of the current device
int dtcwt_b200_l2_info(int64_t* out3) {
    if (!out3) return DTCWT_B200_EINVAL;
    out3[0] = out3[1] = out3[2] = 0;
#ifndef DTCWT_EMU
    int dev = 0, a = 0, b = 0, c = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&c, cudaDevAttrL2CacheSize, dev);
    out3[0] = a; out3[1] = b; out3[2] = c;
#endif
    return DTCWT_B200_OK;
}

// transform2d.py:112-160, levels 1 and 2 of Transform2d.forward for a batch, `chunk` images at a time.
//   lolo1   scratch [chunk][rows + pad_r_hi][cols + pad_c_hi]; persist != 0: kept in L2 (hit ratio persist / 100)
//   lolo2   [n][..][..], yh1 / yh2: the sub-bands of levels 1 and 2 (strides zs1_* / zs2_*, complex elements)
// Level 2 pads its input by one replicated sample per side where the level-1 size is not a multiple of 4 (:134-140).
int dtcwt_b200_fwd2d_level12_f32(const float* x, float* lolo1, float* lolo2, float* yh1, float* yh2, int64_t n, int64_t rows,
                                 int64_t cols, int pad_r_hi, int pad_c_hi, const double* h0o, int m0, const double* h1o, int m1,
                                 const double* lo_a, const double* lo_b, const double* hi_a, const double* hi_b, int m,
                                 int64_t zs1_n, int64_t zs1_band, int64_t zs1_row, int64_t zs2_n, int64_t zs2_band,
                                 int64_t zs2_row, int64_t chunk, int persist, void* stream) {
    if (n < 0 || chunk < 1 || rows < 1 || cols < 1 || pad_r_hi < 0 || pad_r_hi > 1 || pad_c_hi < 0 || pad_c_hi > 1)
        return DTCWT_B200_EINVAL;
    const int64_t Lr = rows + pad_r_hi, Lc = cols + pad_c_hi;
    const int pr = (Lr % 4) ? 1 : 0, pc = (Lc % 4) ? 1 : 0;
    const int64_t r2 = (Lr + 2 * pr) / 2, c2 = (Lc + 2 * pc) / 2;
    if (persist > 0) dtcwt::l2_window(lolo1, chunk * Lr * Lc * (int64_t)sizeof(float), persist / 100.0, stream);
    int rc = DTCWT_B200_OK;
    for (int64_t i0 = 0; i0 < n && rc == DTCWT_B200_OK; i0 += chunk) {
        const int64_t cnt = (n - i0 < chunk) ? n - i0 : chunk;
        rc = dtcwt_b200_fwd2d_level1_f32(x + i0 * rows * cols, lolo1, yh1 + 2 * i0 * zs1_n, cnt, rows, cols, pad_r_hi, pad_c_hi,
                                         h0o, m0, h1o, m1, zs1_n, zs1_band, zs1_row, stream);
        if (rc != DTCWT_B200_OK) break;
        rc = dtcwt_b200_fwd2d_levelq_f32(lolo1, lolo2 + i0 * r2 * c2, yh2 + 2 * i0 * zs2_n, cnt, Lr, Lc, pr, pc, lo_a, lo_b, hi_a,
                                         hi_b, m, zs2_n, zs2_band, zs2_row, stream);
    }
    if (persist > 0) dtcwt::l2_window(nullptr, 0, 0.0, stream);
    return rc;
}

// transform2d.py:240-293, levels 2 and 1 of Transform2d.inverse for a batch, `chunk` images at a time.
//   z2      lowpass entering level 2 [n][rows2][cols2]; yh2 / yh1 the sub-bands of levels 2 and 1
//   z1      scratch [chunk][2 rows2 - 2 crop_r][2 cols2 - 2 crop_c]: the level-1 lowpass, kept in L2 when persist != 0
//   out     [n][rows1][cols1], rows1 = 2 rows2 - 2 crop_r
int dtcwt_b200_inv2d_level21_f32(const float* z2, const float* yh2, const float* yh1, float* z1, float* out, int64_t n,
                                 int64_t rows2, int64_t cols2, int crop_r, int crop_c, const double* lo_a, const double* lo_b,
                                 const double* hi_a, const double* hi_b, int m, const double* gain2, const double* g0o, int m0,
                                 const double* g1o, int m1, const double* gain1, int64_t zs2_n, int64_t zs2_band,
                                 int64_t zs2_row, int64_t zs1_n, int64_t zs1_band, int64_t zs1_row, int64_t chunk, int persist,
                                 void* stream) {
    if (n < 0 || chunk < 1 || rows2 < 2 || cols2 < 2 || crop_r < 0 || crop_r > 1 || crop_c < 0 || crop_c > 1)
        return DTCWT_B200_EINVAL;
    const int64_t r1 = 2 * rows2 - 2 * crop_r, c1 = 2 * cols2 - 2 * crop_c;
    if (persist > 0) dtcwt::l2_window(z1, chunk * r1 * c1 * (int64_t)sizeof(float), persist / 100.0, stream);
    int rc = DTCWT_B200_OK;
    for (int64_t i0 = 0; i0 < n && rc == DTCWT_B200_OK; i0 += chunk) {
        const int64_t cnt = (n - i0 < chunk) ? n - i0 : chunk;
        rc = dtcwt_b200_inv2d_levelq_f32(z2 + i0 * rows2 * cols2, yh2 + 2 * i0 * zs2_n, z1, cnt, rows2, cols2, crop_r, crop_c,
                                         lo_a, lo_b, hi_a, hi_b, m, gain2, zs2_n, zs2_band, zs2_row, stream);
        if (rc != DTCWT_B200_OK) break;
        rc = dtcwt_b200_inv2d_level1_f32(z1, yh1 + 2 * i0 * zs1_n, out + i0 * r1 * c1, cnt, r1, c1, g0o, m0, g1o, m1, gain1,
                                         zs1_n, zs1_band, zs1_row, stream);
    }
    if (persist > 0) dtcwt::l2_window(nullptr, 0, 0.0, stream);
    return rc;
}

}  // extern "C"
#endif  // DTCWT_EMIT_GENERIC
