/*
 * dtcwt_b200 -- C ABI of the Blackwell (sm_100a) DT-CWT hot path.
 *
 * This is the drop-in boundary: everything the host layer (dtcwt_b200/*.py, the
 * mirror of the reference's dtcwt.numpy backend) needs from the device goes
 * through the entry points below.  The reference (rjw57/dtcwt) is pure Python
 * and has no FFI of its own; each entry point therefore cites the reference
 * FUNCTION it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - All data pointers are DEVICE pointers (e.g. torch.Tensor.data_ptr()); tap
 *    pointers (const double* h...) are HOST pointers to float64 taps, which are
 *    rounded to the data type inside the call (reference lowlevel.py:33) and
 *    passed to the kernels by value -- the library keeps no global state, does
 *    not allocate, and is re-entrant across streams and devices.
 *  - `stream` is a cudaStream_t (NULL = legacy default stream).  Launches are
 *    asynchronous; nothing here synchronises.
 *  - A real array filtered along one axis is described as a C-contiguous
 *    [outer][len][inner] view: e.g. an image batch [N][H][W] is (N, H, W) for
 *    the vertical axis and (N*H, W, 1) for the horizontal one; a volume batch
 *    [N][D0][D1][D2] filtered along D1 is (N*D0, D1, D2).
 *  - `pad_lo`/`pad_hi`: the input is treated as if `pad_lo` copies of its first
 *    sample and `pad_hi` copies of its last sample had been attached along the
 *    filtered axis (the reference's odd-size / not-divisible-by-4 extension,
 *    transform2d.py:86-94,134-140; transform3d.py:322-335) -- no copy is made.
 *  - `crop`: the first and last `crop` output samples along the filtered axis
 *    are not produced (reference transform2d.py:263-268, transform3d.py:505-524).
 *  - `accumulate` != 0 adds the result to `y` instead of overwriting it (the
 *    inverse transforms sum two or three filtered arrays).
 *  - complex arrays are interleaved (re, im) pairs of the real type; complex
 *    strides are in COMPLEX elements.
 *  - Return value: 0 on success, < 0 for an argument error (DTCWT_B200_E*),
 *    > 0 a cudaError_t.  dtcwt_b200_error_string() describes either.
 */
#ifndef DTCWT_B200_H
#define DTCWT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTCWT_B200_VERSION 201          /* 0.2.1: + levels 1 and 2 chained through L2 */
#define DTCWT_B200_MAX_TAPS 32          /* longest filter accepted (qshift_32) */

#define DTCWT_B200_OK 0
#define DTCWT_B200_EINVAL (-1)          /* bad shape / tap count / NULL pointer */
#define DTCWT_B200_EUNSUPPORTED (-2)    /* valid request this build has no kernel for */

int dtcwt_b200_version(void);
const char *dtcwt_b200_error_string(int code);
/* 1 when the library was built from the device sources (nvcc, sm_100a); the
 * host-side kernel-logic emulator used by the CPU test-suite reports 0. */
int dtcwt_b200_is_device_build(void);

/* ---- the three 1-D filters -------------------------------------------------
 * replaces dtcwt/numpy/lowlevel.py: colfilter (:47-80), coldfilt (:82-154),
 * colifilt (:156-260).  L = len + pad_lo + pad_hi is the logical input length.
 *   colfilter: y is [outer][L or L+1 (m even)][inner]
 *   coldfilt : needs L % 4 == 0, m even;  y is [outer][L/2][inner]
 *   colifilt : needs len % 2 == 0, m even; y is [outer][2*len - 2*crop][inner]
 */
int dtcwt_b200_colfilter_f32(const float *x, float *y, int64_t outer, int64_t len, int64_t inner,
                             int pad_lo, int pad_hi, const double *h, int m,
                             int accumulate, void *stream);
int dtcwt_b200_colfilter_f64(const double *x, double *y, int64_t outer, int64_t len, int64_t inner,
                             int pad_lo, int pad_hi, const double *h, int m,
                             int accumulate, void *stream);
int dtcwt_b200_coldfilt_f32(const float *x, float *y, int64_t outer, int64_t len, int64_t inner,
                            int pad_lo, int pad_hi, const double *ha, const double *hb, int m,
                            int accumulate, void *stream);
int dtcwt_b200_coldfilt_f64(const double *x, double *y, int64_t outer, int64_t len, int64_t inner,
                            int pad_lo, int pad_hi, const double *ha, const double *hb, int m,
                            int accumulate, void *stream);
int dtcwt_b200_colifilt_f32(const float *x, float *y, int64_t outer, int64_t len, int64_t inner,
                            int crop, const double *ha, const double *hb, int m,
                            int accumulate, void *stream);
int dtcwt_b200_colifilt_f64(const double *x, double *y, int64_t outer, int64_t len, int64_t inner,
                            int crop, const double *ha, const double *hb, int m,
                            int accumulate, void *stream);

/* ---- 2-D sub-band packing ---------------------------------------------------
 * replaces dtcwt/numpy/transform2d.py: q2c (:301-322) and c2q (:324-350).
 * y is real [n][2h][2w]; z is complex, element (b, band, i, j) at
 * z + 2*(b*zs_n + band*zs_band + i*zs_row + j*zs_col) (real units), so both the
 * reference's interleaved (h, w, 6) layout and the planar [6][h][w] layout this
 * library prefers are expressible.  q2c writes bands band0 and band1; c2q reads
 * them and scales by gain0 / gain1 (the gain_mask entries, transform2d.py:243).
 */
int dtcwt_b200_q2c_f32(const float *y, float *z, int64_t n, int64_t h, int64_t w,
                       int64_t zs_n, int64_t zs_band, int64_t zs_row, int64_t zs_col,
                       int band0, int band1, void *stream);
int dtcwt_b200_q2c_f64(const double *y, double *z, int64_t n, int64_t h, int64_t w,
                       int64_t zs_n, int64_t zs_band, int64_t zs_row, int64_t zs_col,
                       int band0, int band1, void *stream);
int dtcwt_b200_c2q_f32(const float *z, float *y, int64_t n, int64_t h, int64_t w,
                       int64_t zs_n, int64_t zs_band, int64_t zs_row, int64_t zs_col,
                       int band0, int band1, double gain0, double gain1, void *stream);
int dtcwt_b200_c2q_f64(const double *z, double *y, int64_t n, int64_t h, int64_t w,
                       int64_t zs_n, int64_t zs_band, int64_t zs_row, int64_t zs_col,
                       int band0, int band1, double gain0, double gain1, void *stream);

/* ---- 1-D sub-band packing ---------------------------------------------------
 * replaces dtcwt/numpy/transform1d.py:86-88 (Hi[::2] + 1j*Hi[1::2]) and c2q1d
 * (:186-196).  hi is real [outer][2k][inner]; z is complex [outer][k][inner].
 * unpack multiplies by `gain` (the 1-D gain_mask entry, transform1d.py:161).
 */
int dtcwt_b200_pack1d_f32(const float *hi, float *z, int64_t outer, int64_t k, int64_t inner, void *stream);
int dtcwt_b200_pack1d_f64(const double *hi, double *z, int64_t outer, int64_t k, int64_t inner, void *stream);
int dtcwt_b200_unpack1d_f32(const float *z, float *hi, int64_t outer, int64_t k, int64_t inner,
                            double gain, void *stream);
int dtcwt_b200_unpack1d_f64(const double *z, double *hi, int64_t outer, int64_t k, int64_t inner,
                            double gain, void *stream);

/* ---- 3-D sub-band packing ---------------------------------------------------
 * replaces dtcwt/numpy/transform3d.py: cube2c (:532-579) and c2cube (:581-619).
 * y is real [n][2a][2b][2c]; z is complex, element (v, chan, i, j, k) at
 * z + 2*(v*zs_n + chan*zs_chan + i*zs_0 + j*zs_1 + k*zs_2).  The four complex
 * outputs p,q,r,s go to channels chan0 .. chan0+3.
 */
int dtcwt_b200_cube2c_f32(const float *y, float *z, int64_t n, int64_t a, int64_t b, int64_t c,
                          int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2,
                          int chan0, void *stream);
int dtcwt_b200_cube2c_f64(const double *y, double *z, int64_t n, int64_t a, int64_t b, int64_t c,
                          int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2,
                          int chan0, void *stream);
int dtcwt_b200_c2cube_f32(const float *z, float *y, int64_t n, int64_t a, int64_t b, int64_t c,
                          int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2,
                          int chan0, void *stream);
int dtcwt_b200_c2cube_f64(const double *z, double *y, int64_t n, int64_t a, int64_t b, int64_t c,
                          int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2,
                          int chan0, void *stream);

/* ---- fused per-level 2-D transform (float32) -----------------------------------
 * One launch per pyramid level: the separable filtering and the q2c / c2q packing
 * of a whole level happen in shared memory / registers; every input sample is read
 * from HBM once and every output written once.
 *
 *   fwd2d_level1  replaces level 1 of Transform2d.forward  (numpy/transform2d.py:112-130)
 *   fwd2d_levelq  replaces levels >= 2 of Transform2d.forward (:132-160)
 *   inv2d_levelq  replaces levels >= 2 of Transform2d.inverse (:240-273)
 *   inv2d_level1  replaces level 1 of Transform2d.inverse  (:275-293)
 *
 * x / z are real [n][rows][cols] (C-contiguous); lolo / out are C-contiguous real
 * outputs; yh is the complex sub-band array of that level, element (b, band, i, j)
 * at yh + 2*(b*zs_n + band*zs_band + i*zs_row + j) -- unit column stride, e.g. the
 * planar [n][6][h][w] layout.
 *   level1: lolo is [n][rows+pad_r_hi][cols+pad_c_hi], sub-bands half that size;
 *           pad_*_hi = 1 repeats the last row / column of an odd-sized image (:86-94).
 *   levelq: pad_r / pad_c = 1 extends that axis by one replicated sample on EACH side
 *           (:134-140); lolo is [n][(rows+2*pad_r)/2][(cols+2*pad_c)/2].
 *           (lo_a, lo_b), (hi_a, hi_b) are coldfilt's / colifilt's (ha, hb) arguments:
 *           the reference passes (h0b, h0a), (h1b, h1a) forward and (g0b, g0a),
 *           (g1b, g1a) inverse.
 *   inverse: z is [n][rows][cols] with rows, cols twice the sub-band size; gain[6] is
 *           the level's gain_mask column (:214-217, :243-245); crop_* = 1 drops the
 *           first and last output row / column (:263-268).  out is
 *           [n][2*rows-2*crop_r][2*cols-2*crop_c] (levelq) or [n][rows][cols] (level1).
 * Returns DTCWT_B200_EUNSUPPORTED for requests the fused kernels do not cover (sides
 * shorter than 32, tap counts other than odd <= 19 (level1) / 10, 14, 18 (levelq),
 * q-shift pairs whose lowpass (highpass) tap correlation is not positive (negative));
 * callers then compose the level from the primitives above.  Rows whose pitch is a
 * multiple of 16 bytes are staged by TMA; set DTCWT_B200_NO_TMA=1 to force plain loads.
 */
int dtcwt_b200_fwd2d_level1_f32(const float *x, float *lolo, float *yh, int64_t n, int64_t rows, int64_t cols,
                                int pad_r_hi, int pad_c_hi, const double *h0o, int m0, const double *h1o, int m1,
                                int64_t zs_n, int64_t zs_band, int64_t zs_row, void *stream);
int dtcwt_b200_fwd2d_levelq_f32(const float *x, float *lolo, float *yh, int64_t n, int64_t rows, int64_t cols,
                                int pad_r, int pad_c, const double *lo_a, const double *lo_b, const double *hi_a,
                                const double *hi_b, int m, int64_t zs_n, int64_t zs_band, int64_t zs_row,
                                void *stream);
int dtcwt_b200_inv2d_levelq_f32(const float *z, const float *yh, float *out, int64_t n, int64_t rows, int64_t cols,
                                int crop_r, int crop_c, const double *lo_a, const double *lo_b, const double *hi_a,
                                const double *hi_b, int m, const double *gain, int64_t zs_n, int64_t zs_band,
                                int64_t zs_row, void *stream);
int dtcwt_b200_inv2d_level1_f32(const float *z, const float *yh, float *out, int64_t n, int64_t rows, int64_t cols,
                                const double *g0o, int m0, const double *g1o, int m1, const double *gain,
                                int64_t zs_n, int64_t zs_band, int64_t zs_row, void *stream);
/* `_bp` families (6-tuple biort, 12-tuple qshift; transform2d.py:116-127, 145-157, 254-262, 279-292): the diagonal
 * sub-bands 1 and 4 use the band-pass pair h2 / g2 in both directions.  A `_bp` level is the ordinary launch above
 * followed by one of these on the same yh / out:
 *   fwd2d_level*_hh  overwrite bands 1 and 4 of yh with q2c(V:h2(H:h2(x)))
 *   inv2d_level*_hh  out += H:g2(V:g2(c2q(bands 1, 4) * gain[1], gain[4])); the ordinary inverse launch before it is given
 *                    gain[1] = gain[4] = 0
 * (h2_a, h2_b) / (g2_a, g2_b) are coldfilt's / colifilt's (ha, hb): the reference passes (h2b, h2a) / (g2b, g2a). */
int dtcwt_b200_fwd2d_level1_hh_f32(const float *x, float *yh, int64_t n, int64_t rows, int64_t cols, int pad_r_hi,
                                   int pad_c_hi, const double *h2o, int m2, int64_t zs_n, int64_t zs_band,
                                   int64_t zs_row, void *stream);
int dtcwt_b200_fwd2d_levelq_hh_f32(const float *x, float *yh, int64_t n, int64_t rows, int64_t cols, int pad_r, int pad_c,
                                   const double *h2_a, const double *h2_b, int m, int64_t zs_n, int64_t zs_band,
                                   int64_t zs_row, void *stream);
int dtcwt_b200_inv2d_levelq_hh_f32(const float *yh, float *out, int64_t n, int64_t rows, int64_t cols, int crop_r,
                                   int crop_c, const double *g2_a, const double *g2_b, int m, const double *gain,
                                   int64_t zs_n, int64_t zs_band, int64_t zs_row, void *stream);
int dtcwt_b200_inv2d_level1_hh_f32(const float *yh, float *out, int64_t n, int64_t rows, int64_t cols, const double *g2o,
                                   int m2, const double *gain, int64_t zs_n, int64_t zs_band, int64_t zs_row,
                                   void *stream);

/* Levels 1 and 2 chained CHUNK images at a time, the level-1 lowpass kept in L2 (same kernels, same results as the
 * per-level entry points above; what changes is the launch order and the cache policy of one scratch buffer).
 *   fwd2d_level12  = fwd2d_level1 + fwd2d_levelq per chunk (numpy/transform2d.py:112-160): lolo1 is scratch for
 *                    chunk images [chunk][rows+pad_r_hi][cols+pad_c_hi]; level 2 pads by one replicated sample per
 *                    side where that size is not a multiple of 4 (:134-140); lolo2 / yh1 / yh2 are the batch outputs.
 *   inv2d_level21  = inv2d_levelq + inv2d_level1 per chunk (:240-293): z2 [n][rows2][cols2] enters level 2, z1 is
 *                    scratch [chunk][2 rows2 - 2 crop_r][2 cols2 - 2 crop_c], out has that size per image.
 *   persist        0: plain launch order change; 1..100: the scratch is marked L2-persisting on `stream` for the
 *                    duration of the call (cudaStreamAttributeAccessPolicyWindow, hit ratio persist / 100), so it is
 *                    written and re-read inside L2 and the 8 B/pixel of its HBM round trip disappear.
 *   l2_info        {max persisting bytes, max access-policy window bytes, L2 bytes} of the current device. */
int dtcwt_b200_fwd2d_level12_f32(const float *x, float *lolo1, float *lolo2, float *yh1, float *yh2, int64_t n,
                                 int64_t rows, int64_t cols, int pad_r_hi, int pad_c_hi, const double *h0o, int m0,
                                 const double *h1o, int m1, const double *lo_a, const double *lo_b,
                                 const double *hi_a, const double *hi_b, int m, int64_t zs1_n, int64_t zs1_band,
                                 int64_t zs1_row, int64_t zs2_n, int64_t zs2_band, int64_t zs2_row, int64_t chunk,
                                 int persist, void *stream);
int dtcwt_b200_inv2d_level21_f32(const float *z2, const float *yh2, const float *yh1, float *z1, float *out, int64_t n,
                                 int64_t rows2, int64_t cols2, int crop_r, int crop_c, const double *lo_a,
                                 const double *lo_b, const double *hi_a, const double *hi_b, int m,
                                 const double *gain2, const double *g0o, int m0, const double *g1o, int m1,
                                 const double *gain1, int64_t zs2_n, int64_t zs2_band, int64_t zs2_row, int64_t zs1_n,
                                 int64_t zs1_band, int64_t zs1_row, int64_t chunk, int persist, void *stream);
int dtcwt_b200_l2_info(int64_t *out3);

/* ---- fused per-level 3-D transform (float32) -----------------------------------
 * A level of Transform3d is two launches: the two in-slice axes of every slice in one
 * tile kernel (the 2-D kernels above in a mode that keeps the four real images), and
 * the depth axis with the 2x2x2 packers cube2c / c2cube in registers.  Volumes are
 * C-contiguous [n][d0][d1][d2]; yh is the planar complex array of the level, element
 * (b, chan, i, j, k) at yh + 2*(b*zs_n + chan*zs_chan + i*zs_0 + j*zs_1 + k*zs_2), the
 * 28 channels in the reference's order (transform3d.py:280-288).  `scratch` is caller-
 * owned device memory of the stated size (the library never allocates).
 *
 *   fwd3d_level1_lo  replaces _level1_xfm_no_highpass  (numpy/transform3d.py:291-315)
 *   inv3d_level1_lo  replaces _level1_ifm_no_highpass  (:442-456)
 *                    y = colfilter(h) along all three axes; scratch n*d0*d1*d2 floats
 *   fwd3d_level1     replaces _level1_xfm (:208-289), odd-length taps; lll [n][d0][d1][d2],
 *                    yh [n][28][d0/2][d1/2][d2/2]; scratch 4*n*d0*d1*d2 floats
 *   inv3d_level1     replaces _level1_ifm (:385-440), odd-length taps; scratch 4*n*a0*a1*a2
 *   fwd3d_levelq     replaces _level2_xfm (:317-383); pad_i = replicated samples attached to
 *                    EACH side of axis i (ext_mode 4: 0 or 1, ext_mode 8: 0 or 2; :322-335);
 *                    with L_i = d_i + 2 pad_i: lll [n][L0/2][L1/2][L2/2], yh [n][28][L0/4][L1/4][L2/4];
 *                    scratch n*d0*L1*L2 floats
 *   inv3d_levelq     replaces _level2_ifm (:458-526); yl [n][a0][a1][a2]; crop_i = samples
 *                    dropped at EACH end of axis i (:505-524); out [n][2a0-2crop0][2a1-2crop1]
 *                    [2a2-2crop2]; scratch 4*n*(2a0-2crop0)*a1*a2 floats
 * (lo_a, lo_b), (hi_a, hi_b) as for the 2-D levels.  EUNSUPPORTED: in-slice sides < 32,
 * depth < 16, pointers not 16-byte aligned, tap counts / correlations as for the 2-D levels;
 * callers then compose the level from the per-axis primitives.
 */
int dtcwt_b200_fwd3d_level1_lo_f32(const float *x, float *y, float *scratch, int64_t n, int64_t d0, int64_t d1,
                                   int64_t d2, const double *h0o, int m0, void *stream);
int dtcwt_b200_inv3d_level1_lo_f32(const float *yl, float *out, float *scratch, int64_t n, int64_t d0, int64_t d1,
                                   int64_t d2, const double *g0o, int m0, void *stream);
int dtcwt_b200_fwd3d_level1_f32(const float *x, float *lll, float *yh, float *scratch, int64_t n, int64_t d0,
                                int64_t d1, int64_t d2, const double *h0o, int m0, const double *h1o, int m1,
                                int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void *stream);
int dtcwt_b200_inv3d_level1_f32(const float *yl, const float *yh, float *out, float *scratch, int64_t n, int64_t a0,
                                int64_t a1, int64_t a2, const double *g0o, int m0, const double *g1o, int m1,
                                int64_t zs_n, int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void *stream);
int dtcwt_b200_fwd3d_levelq_f32(const float *x, float *lll, float *yh, float *scratch, int64_t n, int64_t d0,
                                int64_t d1, int64_t d2, int pad0, int pad1, int pad2, const double *lo_a,
                                const double *lo_b, const double *hi_a, const double *hi_b, int m, int64_t zs_n,
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void *stream);
int dtcwt_b200_inv3d_levelq_f32(const float *yl, const float *yh, float *out, float *scratch, int64_t n, int64_t a0,
                                int64_t a1, int64_t a2, int crop0, int crop1, int crop2, const double *lo_a,
                                const double *lo_b, const double *hi_a, const double *hi_b, int m, int64_t zs_n,
                                int64_t zs_chan, int64_t zs_0, int64_t zs_1, int64_t zs_2, void *stream);

/* ---- registration and re-sampling (float64 arithmetic) ------------------------------
 * replaces dtcwt/registration.py (estimatereg :304-372 and its helpers) and dtcwt/sampling.py
 * (sample :105, rescale :131, sample_highpass :192, rescale_highpass :224).
 *
 *   reg_qtilde     qtildematrices (:141-212) for ONE level, with confidence (:84-139) and phasegradient
 *                  (:32-76) inside: src / ref are the level's six complex sub-bands of the (warped) source
 *                  and of the reference image, element (b, band, i, j) at 2*(b*x_n + band*x_band + i*x_row
 *                  + j*x_col).  reduce == 0: qt is [n][h][w][27] float64; reduce != 0: qt is [n][27], the
 *                  sum over the image ADDED to its content (the global estimate, :333-338).
 *   reg_boxrescale out (+)= rescale(_boxfilter(qt, 3), (H, W), 'bilinear')  (:357-362, :425-446)
 *   reg_solve      solvetransform (:214-257): avecs[i] (+)= solve(Q_i, -q_i), Q_i holding ONLY the upper
 *                  triangle of the 27-vector's 21 matrix elements, as the reference does (:229-232)
 *   reg_coords     avecs [n][H][W][6] -> xs, ys [n][h][w]; mode 0: velocityfield(avecs, (h, w), 'bilinear')
 *                  (:374-393), mode 1: the pixel coordinates warp / warphighpass sample at (:395-423)
 *   sample         im element (b, y, x, c) at k*(b*i_n + y*i_y + x*i_x + c*i_c), out likewise with o_*;
 *                  k = 2 for complex (interleaved).  method 0 nearest / 1 bilinear / 2 lanczos / 3 the
 *                  7-tap lanczos of upsample() (sampling.py:312-320; doubled rescale grid only).  coords 0:
 *                  positions from xs / ys ([oh][ow] planes, coord_n elements apart per batch item, 0 =
 *                  shared); coords 1: the rescale grid.  wx / wy non-NULL (HOST arrays, C <= 8): complex
 *                  sub-bands, phase un-rolled by exp(-j(wx x + wy y)) before sampling and re-rolled after.
 */
int dtcwt_b200_reg_qtilde_f32(const float *src, const float *ref, double *qt, int64_t n, int64_t h, int64_t w,
                              int64_t s_n, int64_t s_band, int64_t s_row, int64_t s_col, int64_t r_n,
                              int64_t r_band, int64_t r_row, int64_t r_col, int reduce, void *stream);
int dtcwt_b200_reg_qtilde_f64(const double *src, const double *ref, double *qt, int64_t n, int64_t h, int64_t w,
                              int64_t s_n, int64_t s_band, int64_t s_row, int64_t s_col, int64_t r_n,
                              int64_t r_band, int64_t r_row, int64_t r_col, int reduce, void *stream);
int dtcwt_b200_reg_boxrescale(const double *qt, double *out, int64_t n, int64_t h, int64_t w, int64_t H, int64_t W,
                              int accumulate, void *stream);
int dtcwt_b200_reg_solve(const double *qt, double *avecs, int64_t count, int accumulate, void *stream);
int dtcwt_b200_reg_coords(const double *avecs, double *xs, double *ys, int64_t n, int64_t H, int64_t W, int64_t h,
                          int64_t w, int mode, void *stream);
int dtcwt_b200_sample_f32(const float *im, float *out, const double *xs, const double *ys, int64_t n, int64_t h,
                          int64_t w, int64_t C, int64_t oh, int64_t ow, int64_t i_n, int64_t i_y, int64_t i_x,
                          int64_t i_c, int64_t o_n, int64_t o_y, int64_t o_x, int64_t o_c, int64_t coord_n,
                          int is_complex, int method, int coords, const double *wx, const double *wy, void *stream);
int dtcwt_b200_sample_f64(const double *im, double *out, const double *xs, const double *ys, int64_t n, int64_t h,
                          int64_t w, int64_t C, int64_t oh, int64_t ow, int64_t i_n, int64_t i_y, int64_t i_x,
                          int64_t i_c, int64_t o_n, int64_t o_y, int64_t o_x, int64_t o_c, int64_t coord_n,
                          int is_complex, int method, int coords, const double *wx, const double *wy, void *stream);

/* ---- keypoints (float64 arithmetic) ---------------------------------------------------
 * replaces the per-pixel work of dtcwt/keypoint.py (find_keypoints :9-141):
 *   kp_energy  the keypoint-energy map of one level's six sub-bands (:143-156): method 0 fauqueur (scale_gain =
 *              alpha**(scale+1), beta), 1 bendale, 2 kingsbury (kappa); yh element (b, band, i, j) at
 *              2*(b*s_n + band*s_band + i*s_row + j*s_col); e is [n][h][w] float64
 *   kp_maxima  _kp_energy_maxima (:201-260): out [n][h][w][4] = (1 if the pixel is a kept local maximum else 0,
 *              refined row, refined column, energy); refine != 0 fits the quadratic patch (:221-252)
 */
int dtcwt_b200_kp_energy_f32(const float *yh, double *e, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band,
                             int64_t s_row, int64_t s_col, int method, double scale_gain, double beta, double kappa,
                             void *stream);
int dtcwt_b200_kp_energy_f64(const double *yh, double *e, int64_t n, int64_t h, int64_t w, int64_t s_n, int64_t s_band,
                             int64_t s_row, int64_t s_col, int method, double scale_gain, double beta, double kappa,
                             void *stream);
int dtcwt_b200_kp_maxima(const double *x, double *out, int64_t n, int64_t h, int64_t w, double threshold, int refine,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DTCWT_B200_H */
