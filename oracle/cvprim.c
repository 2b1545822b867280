/*
 * oracle/cvprim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cvprim.h).
 *
 * Restated from OpenCV 4.13.0's published algorithms (the reference depends on
 * an un-vendored, un-pinned OpenCV: CMakeLists.txt:34-40 of the reference):
 *   resize      modules/imgproc/src/resize.cpp   (resizeGeneric_, HResizeLinear,
 *                                                 VResizeLinear, INTER_RESIZE_COEF_BITS=11)
 *   blur        modules/imgproc/src/smooth.dispatch.cpp + fixedpoint.inl.hpp
 *               (ufixedpoint16 bit-exact kernel [18,34,48,56,48,34,18]/256)
 *   FAST        modules/features2d/src/fast.cpp + fast_score.cpp
 *   fastAtan2   modules/core/src/mathfuncs_core.simd.hpp (atan_f32)
 * Pinned against cv2 4.13.0 by tests/test_oracle_cvprim.py.
 *
 * Build with -ffp-contract=off: every float expression below must round
 * exactly where the C source says it rounds.
 */
#include "cvprim.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

int cvp_round_f(float v)  { return (int)lrintf(v); }
int cvp_round_d(double v) { return (int)lrint(v); }
int cvp_floor_d(double v) { int i = (int)v; return i - (i > v); }
int cvp_ceil_d(double v)  { int i = (int)v; return i + (i < v); }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline int16_t sat_s16_from_float(float v)
{
    int i = cvp_round_f(v);
    return (int16_t)clampi(i, -32768, 32767);
}

/* ------------------------------------------------------------------ resize */

void cvp_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride)
{
    if (sw == dw && sh == dh) {             /* cv::resize copies in that case */
        for (int y = 0; y < dh; ++y) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, (size_t)dw);
        return;
    }
    /* cv::resize: inv_scale = dsize/ssize; hal::resize: scale = 1./inv_scale */
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;

    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    int16_t *ialpha = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)dw);
    int32_t *rows[2];
    rows[0] = (int32_t *)malloc(sizeof(int32_t) * (size_t)dw);
    rows[1] = (int32_t *)malloc(sizeof(int32_t) * (size_t)dw);

    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvp_floor_d(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx]     = sat_s16_from_float((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = sat_s16_from_float(fx * 2048.f);
    }

    int have[2] = {-1, -1};                 /* source row held by rows[0] / rows[1] (cv::resize keeps them between rows too) */
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvp_floor_d(fy);
        fy -= sy;
        const int16_t b0 = sat_s16_from_float((1.f - fy) * 2048.f);
        const int16_t b1 = sat_s16_from_float(fy * 2048.f);
        const int want[2] = {clampi(sy, 0, sh - 1), clampi(sy + 1, 0, sh - 1)};
        if (have[1] == want[0] && have[0] != want[0]) {        /* the lower row of the previous pair is the upper row now */
            int32_t *t = rows[0]; rows[0] = rows[1]; rows[1] = t;
            have[0] = have[1]; have[1] = -1;
        }
        for (int k = 0; k < 2; ++k) {
            if (have[k] == want[k]) continue;
            if (k == 1 && want[1] == want[0]) { memcpy(rows[1], rows[0], sizeof(int32_t) * (size_t)dw); have[1] = want[1]; continue; }
            const uint8_t *S = src + (size_t)want[k] * sstride;
            int32_t *r = rows[k];
            for (int dx = 0; dx < dw; ++dx) {
                const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                r[dx] = S[sx] * ialpha[2 * dx] + S[sx1] * ialpha[2 * dx + 1];
            }
            have[k] = want[k];
        }
        uint8_t *D = dst + (size_t)dy * dstride;
        const int32_t *r0 = rows[0], *r1 = rows[1];
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)clampi(v, 0, 255);
        }
    }
    free(xofs); free(ialpha); free(rows[0]); free(rows[1]);
}

/* ------------------------------------------------------------------ border */

static inline int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

void cvp_border_reflect101_u8(const uint8_t *src, int w, int h, int sstride,
                              uint8_t *dst, int dstride,
                              int top, int bottom, int left, int right)
{
    const int W = w + left + right, H = h + top + bottom;
    /* interior rows first (src may alias dst's interior, so go through a row copy) */
    uint8_t *row = (uint8_t *)malloc((size_t)w);
    for (int y = 0; y < h; ++y) {
        memcpy(row, src + (size_t)y * sstride, (size_t)w);
        uint8_t *d = dst + (size_t)(y + top) * dstride;
        for (int x = 0; x < left; ++x) d[x] = row[reflect101(x - left, w)];
        memcpy(d + left, row, (size_t)w);
        for (int x = left + w; x < W; ++x) d[x] = row[reflect101(x - left, w)];
    }
    free(row);
    for (int y = 0; y < H; ++y) {
        if (y >= top && y < top + h) continue;
        const int sy = reflect101(y - top, h);
        memcpy(dst + (size_t)y * dstride, dst + (size_t)(sy + top) * dstride, (size_t)W);
    }
}

/* -------------------------------------------------------------------- blur */

/* Same arithmetic laid out for the vectoriser: the row is extended by its reflect-101 margins once, then both passes are plain
 * loops over contiguous data (what OpenCV's SIMD path does; the scalar form below stays as the cross-check). */
void cvp_gaussian7x7_s2_u8(const uint8_t *src, int w, int h, int sstride,
                           uint8_t *dst, int dstride)
{
    uint16_t *H = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    uint8_t *pad = (uint8_t *)malloc((size_t)w + 6);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * sstride;
        for (int i = 0; i < 3; ++i) { pad[i] = s[reflect101(i - 3, w)]; pad[w + 3 + i] = s[reflect101(w + i, w)]; }
        memcpy(pad + 3, s, (size_t)w);
        uint16_t *hr = H + (size_t)y * w;
        for (int x = 0; x < w; ++x)
            hr[x] = (uint16_t)(18u * (pad[x] + pad[x + 6]) + 34u * (pad[x + 1] + pad[x + 5]) + 48u * (pad[x + 2] + pad[x + 4]) + 56u * pad[x + 3]);
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *r0 = H + (size_t)reflect101(y - 3, h) * w, *r1 = H + (size_t)reflect101(y - 2, h) * w, *r2 = H + (size_t)reflect101(y - 1, h) * w,
                       *r3 = H + (size_t)y * w, *r4 = H + (size_t)reflect101(y + 1, h) * w, *r5 = H + (size_t)reflect101(y + 2, h) * w,
                       *r6 = H + (size_t)reflect101(y + 3, h) * w;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = 18u * ((uint32_t)r0[x] + r6[x]) + 34u * ((uint32_t)r1[x] + r5[x]) + 48u * ((uint32_t)r2[x] + r4[x]) + 56u * (uint32_t)r3[x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(pad); free(H);
}

void cvp_gaussian7x7_s2_u8_scalar(const uint8_t *src, int w, int h, int sstride,
                                  uint8_t *dst, int dstride)
{
    static const uint32_t k[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t *H = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * sstride;
        uint16_t *hr = H + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, w)];
            hr[x] = (uint16_t)acc;            /* <= 255*256, no saturation needed */
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *r[7];
        for (int j = 0; j < 7; ++j) r[j] = H + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int j = 0; j < 7; ++j) acc += k[j] * r[j][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(H);
}

/* -------------------------------------------------------------------- FAST */

static const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int cvp_fast_score(const uint8_t *p, int stride)
{
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[RING_DY[k] * stride + RING_DX[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -256;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int j = 1; j < 9; ++j) {
            if (d[s + j] < mn) mn = d[s + j];
            if (d[s + j] > mx) mx = d[s + j];
        }
        if (mn > A) A = mn;
        if (-mx > A) A = -mx;
    }
    return A - 1;
}

static inline int has9(unsigned m)         /* 9 contiguous set bits in a 16-bit ring */
{
    m |= m << 16;
    m &= m >> 1;            /* runs of 2 */
    m &= m >> 2;            /* runs of 4 */
    m &= m >> 4;            /* runs of 8 */
    m &= m >> 1;            /* runs of 9 */
    return (m & 0xffffu) != 0;
}

int cvp_fast9_16_scalar(const uint8_t *img, int w, int h, int stride, int threshold,
                        int nms, cvp_corner *out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = RING_DY[k] * stride + RING_DX[k];
    uint8_t *S = (uint8_t *)calloc((size_t)w * (size_t)h, 1);   /* 0 = not a corner at threshold */
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t *p = row + x;
            const int v = p[0], hi = v + t, lo = v - t;
            /* any 9-arc holds two adjacent compass points of one polarity */
            const int n = p[off[8]], s = p[off[0]], e = p[off[4]], wv = p[off[12]];
            const int bn = n > hi, bs = s > hi, be = e > hi, bw = wv > hi;
            const int dn = n < lo, ds = s < lo, de = e < lo, dw_ = wv < lo;
            if (!((bn & be) | (be & bs) | (bs & bw) | (bw & bn) |
                  (dn & de) | (de & ds) | (ds & dw_) | (dw_ & dn)))
                continue;
            unsigned mb = 0, md = 0;
            for (int k = 0; k < 16; ++k) {
                const int r = p[off[k]];
                mb |= (unsigned)(r > hi) << k;
                md |= (unsigned)(r < lo) << k;
            }
            if (!has9(mb) && !has9(md)) continue;
            const int sc = cvp_fast_score(p, stride);
            S[(size_t)y * w + x] = (uint8_t)(sc > 255 ? 255 : (sc < 0 ? 0 : sc));
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *s0 = S + (size_t)(y - 1) * w, *s1 = S + (size_t)y * w, *s2 = S + (size_t)(y + 1) * w;
        for (int x = 3; x < w - 3; ++x) {
            const int sc = s1[x];
            if (sc < t || (t <= 0 && !sc)) continue;
            if (nms) {
                if (!(sc > s1[x - 1] && sc > s1[x + 1] &&
                      sc > s0[x - 1] && sc > s0[x] && sc > s0[x + 1] &&
                      sc > s2[x - 1] && sc > s2[x] && sc > s2[x + 1]))
                    continue;
            }
            if (n < cap) { out[n].x = x; out[n].y = y; out[n].score = sc; }
            ++n;
        }
    }
    free(S);
    return n;
}

#if defined(__AVX2__)
#include <immintrin.h>
/* 32 pixels at a time on unsigned bytes.  With sat(a - b) = max(a - b, 0):  B_k = sat(r_k - c), D_k = sat(c - r_k),
 * bright = max over the sixteen 9-arcs of min(B over the arc), dark likewise on D, and max(bright, dark) = max(A, 0) for the
 * A of cornerScore<16> -- exact wherever the score can reach a threshold >= 1.  The arcs are taken in pairs 2i, 2i+1 that share
 * the eight ring pixels 2i+1 .. 2i+8: max(min arc 2i, min arc 2i+1) = min(those eight, max(r_2i, r_2i+9)).  Every pixel is
 * scored (no pre-test: on these lanes the full score costs about as much as OpenCV's compass test does per pixel). */
static inline __m256i fast_arcs_u8(const __m256i *e)
{
    __m256i l2[8], l4[8], best = _mm256_setzero_si256();
    for (int i = 0; i < 8; ++i) l2[i] = _mm256_min_epu8(e[2 * i + 1], e[(2 * i + 2) & 15]);
    for (int i = 0; i < 8; ++i) l4[i] = _mm256_min_epu8(l2[i], l2[(i + 1) & 7]);
    for (int i = 0; i < 8; ++i) {
        const __m256i eight = _mm256_min_epu8(l4[i], l4[(i + 2) & 7]);
        best = _mm256_max_epu8(best, _mm256_min_epu8(eight, _mm256_max_epu8(e[2 * i], e[(2 * i + 9) & 15])));
    }
    return best;
}

int cvp_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold,
                 int nms, cvp_corner *out, int cap)
{
    if (w < 7 || h < 7) return 0;
    if (threshold < 1 || threshold > 254) return cvp_fast9_16_scalar(img, w, h, stride, threshold, nms, out, cap);
    /* padded copy: 32-byte loads may run past the row end, and the score rows need zeroed margins */
    const int ts = (w + 32 + 31) & ~31;
    uint8_t *T = (uint8_t *)calloc((size_t)ts * (size_t)(h + 1), 1), *S = (uint8_t *)calloc((size_t)ts * (size_t)(h + 1), 1);
    for (int y = 0; y < h; ++y) memcpy(T + (size_t)y * ts, img + (size_t)y * stride, (size_t)w);
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = RING_DY[k] * ts + RING_DX[k];
    const __m256i vt = _mm256_set1_epi8((char)threshold), one = _mm256_set1_epi8(1);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *row = T + (size_t)y * ts;
        uint8_t *srow = S + (size_t)y * ts;
        for (int x = 3; x < w - 3; x += 32) {
            const __m256i c = _mm256_loadu_si256((const __m256i *)(row + x));
            __m256i b[16], d[16];
            for (int k = 0; k < 16; ++k) {
                const __m256i r = _mm256_loadu_si256((const __m256i *)(row + x + off[k]));
                b[k] = _mm256_subs_epu8(r, c);
                d[k] = _mm256_subs_epu8(c, r);
            }
            const __m256i a = _mm256_max_epu8(fast_arcs_u8(b), fast_arcs_u8(d));     /* max(A, 0) */
            /* score = A - 1; corner at t <=> A - 1 >= t <=> A > t; S holds the score of corners, 0 elsewhere (as cv::FAST) */
            const __m256i is_corner = _mm256_cmpeq_epi8(_mm256_subs_epu8(a, vt), _mm256_setzero_si256());       /* 0xff where NOT a corner */
            const __m256i sc = _mm256_andnot_si256(is_corner, _mm256_sub_epi8(a, one));
            _mm256_storeu_si256((__m256i *)(srow + x), sc);
        }
        for (int x = w - 3; x < ts; ++x) srow[x] = 0;                 /* the last chunk spilled past the scored columns */
    }
    int n = 0;
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *s0 = S + (size_t)(y - 1) * ts, *s1 = S + (size_t)y * ts, *s2 = S + (size_t)(y + 1) * ts;
        for (int x = 3; x < w - 3; ++x) {
            const int sc = s1[x];
            if (sc < t) continue;
            if (nms) {
                if (!(sc > s1[x - 1] && sc > s1[x + 1] &&
                      sc > s0[x - 1] && sc > s0[x] && sc > s0[x + 1] &&
                      sc > s2[x - 1] && sc > s2[x] && sc > s2[x + 1]))
                    continue;
            }
            if (n < cap) { out[n].x = x; out[n].y = y; out[n].score = sc; }
            ++n;
        }
    }
    free(T); free(S);
    return n;
}
#else
int cvp_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold,
                 int nms, cvp_corner *out, int cap)
{
    return cvp_fast9_16_scalar(img, w, h, stride, threshold, nms, out, cap);
}
#endif

/* --------------------------------------------------------------- fastAtan2 */

float cvp_fast_atan2(float y, float x)
{
    static const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* cv::cvtColor(src, dst, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) on 8-bit images (called from
 * Tracking::GrabImage*, src/Tracking.cc:459-472): 15-bit fixed point, R 9798, G 19235, B 3735, round to nearest. */
void cvp_cvt_gray_u8(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order, uint8_t *dst, int dstride)
{
    const int c0 = rgb_order ? 9798 : 3735, c2 = rgb_order ? 3735 : 9798;
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * stride;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x, s += channels)
            d[x] = (uint8_t)((s[0] * c0 + s[1] * 19235 + s[2] * c2 + (1 << 14)) >> 15);
    }
}
