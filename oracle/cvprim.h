/*
 * oracle/cvprim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatements of the OpenCV primitives that the reference's ORB
 * front end calls (src/ORBextractor.cc:103,809-815,1090,1116,1124-1132 in the
 * reference tree).  OpenCV itself is not vendored by the reference and its
 * C++ library is absent from this image, so the arithmetic is restated from
 * OpenCV 4.13.0's published algorithms and pinned bit-exactly against the
 * Python cv2 4.13.0 wheel (tests/test_oracle_cvprim.py, tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may link or call this.  The CUDA product path never does.
 */
#ifndef ORACLE_CVPRIM_H
#define ORACLE_CVPRIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cvRound / cvFloor / cvCeil (round-half-to-even, like _mm_cvtss_si32). */
int cvp_round_f(float v);
int cvp_round_d(double v);
int cvp_floor_d(double v);
int cvp_ceil_d(double v);

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (11-bit fixed
 * point coefficients, 2-step rounding in the vertical pass). */
void cvp_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride);

/* cv::copyMakeBorder(..., BORDER_REFLECT_101) for CV_8UC1.  dst is
 * (w+left+right) x (h+top+bottom).  src may alias the interior of dst. */
void cvp_border_reflect101_u8(const uint8_t *src, int w, int h, int sstride,
                              uint8_t *dst, int dstride,
                              int top, int bottom, int left, int right);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for a
 * continuous CV_8UC1 image (8.8 fixed-point separable kernel). */
void cvp_gaussian7x7_s2_u8(const uint8_t *src, int w, int h, int sstride,
                           uint8_t *dst, int dstride);

/* The plain scalar form of the same blur (cross-check of the vectorised one). */
void cvp_gaussian7x7_s2_u8_scalar(const uint8_t *src, int w, int h, int sstride,
                                  uint8_t *dst, int dstride);

typedef struct { int x, y, score; } cvp_corner;

/* FAST-9/16 corner score of one pixel (OpenCV cornerScore<16>: the largest
 * threshold at which the pixel is still a corner; corner at t <=> score>=t).
 * Needs a 3-pixel ring around p. */
int cvp_fast_score(const uint8_t *p, int stride);

/* cv::FAST(img, kps, threshold, nonmaxSuppression, TYPE_9_16) on a w x h
 * 8-bit image.  Emits (x, y, score) in row-major order, returns the count
 * (may exceed cap; only the first cap are written). */
int cvp_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold,
                 int nms, cvp_corner *out, int cap);

/* The scalar form (compass pre-test, 16-bit ring masks, cornerScore per corner), as cv::FAST's generic path; the function above
 * scores 32 pixels at a time with AVX2 byte min / max when the build has it.  Identical outputs (tests/test_oracle_cvprim.py). */
int cvp_fast9_16_scalar(const uint8_t *img, int w, int h, int stride, int threshold,
                        int nms, cvp_corner *out, int cap);

/* cv::fastAtan2(y, x): degrees in [0, 360), 7th-order polynomial, fp32. */
float cvp_fast_atan2(float y, float x);

/* cv::cvtColor to grey, 8-bit, 3 or 4 channels; rgb_order != 0: first channel is R (CV_RGB[A]2GRAY), else B (CV_BGR[A]2GRAY) */
void cvp_cvt_gray_u8(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order, uint8_t *dst, int dstride);

#ifdef __cplusplus
}
#endif
#endif
