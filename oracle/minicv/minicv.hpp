// oracle/minicv/minicv.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal stand-in for the slice of the OpenCV C++ API that the reference's
// src/ORBextractor.cc, include/ORBextractor.h and ORBmatcher::DescriptorDistance
// use, so that those reference sources compile UNMODIFIED, from where they lie
// under /root/reference, into oracle/_ref/ (see oracle/Makefile).  The real
// OpenCV C++ library is not present in this image.  Image primitives forward to
// oracle/cvprim.c (bit-exact to cv2 4.13.0, tests/test_oracle_cvprim.py).
// CV_8UC1 matrices everywhere; CV_32F only for the 11x11 patch arithmetic of
// Frame::ComputeStereoMatches (src/Frame.cc:949-985: convertTo, Mat - float*Mat::ones, cv::norm L1).
#ifndef ORACLE_MINICV_HPP
#define ORACLE_MINICV_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "cvprim.h"

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

typedef unsigned char uchar;

static inline int cvRound(double v) { return cvp_round_d(v); }
static inline int cvRound(float v) { return cvp_round_f(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvp_floor_d(v); }
static inline int cvFloor(float v) { return cvp_floor_d(v); }
static inline int cvCeil(double v) { return cvp_ceil_d(v); }
static inline int cvCeil(float v) { return cvp_ceil_d(v); }

namespace cv {

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b)
{
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0,
             int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

// Mat::zeros() yields an expression object; assigning it to an existing Mat of the
// same size/type fills that Mat's buffer IN PLACE (cv::MatExpr semantics).  The
// reference relies on this: computeDescriptors() assigns Mat::zeros(...) to a
// row-range view of the output matrix (src/ORBextractor.cc:1037) and then
// writes the descriptors through it.
struct MatExpr { int rows, cols, type; };
class Mat;
struct MatTExpr { const Mat *a; double alpha; };      // alpha * A^T, see the product operators below

class Mat {
public:
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8UC1), esz_(1) {}
    Mat(const MatExpr &e) : Mat() { *this = e; }
    Mat &operator=(const MatExpr &e)
    {
        create(e.rows, e.cols, e.type);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, (size_t)cols * esz_);
        return *this;
    }
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    // external (non-owning) data, like cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void *ext, size_t _step = 0)
        : rows(r), cols(c), data((uchar *)ext), step(_step ? _step : (size_t)c), type_(type), esz_(1) { assert(type == CV_8UC1); }

    void create(int r, int c, int type)
    {
        assert(type == CV_8UC1 || type == CV_32F);
        if (data && r == rows && c == cols && type == type_) return;   // cv::Mat::create keeps a matching buffer
        esz_ = type == CV_32F ? 4 : 1;
        rows = r; cols = c; step = (size_t)c * esz_; type_ = type;
        buf_.reset(new uchar[(size_t)r * (size_t)c * esz_ + 64], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    static MatExpr zeros(int r, int c, int type) { MatExpr e = {r, c, type}; return e; }

    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz_);
        return m;
    }
    Mat rowRange(int a, int b) const { assert(0 <= a && a <= b && b <= rows); Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { assert(0 <= a && a <= b && b <= cols); Mat m(*this); m.data = data + (size_t)a * esz_; m.cols = b - a; return m; }
    // 8U -> 32F (or a copy); dst may alias *this, like cv::Mat::convertTo
    void convertTo(Mat &dst, int type) const
    {
        assert(type == CV_32F);
        Mat t(rows, cols, CV_32F);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                t.at<float>(y, x) = type_ == CV_32F ? at<float>(y, x) : (float)at<uchar>(y, x);
        dst = t;
    }
    static Mat ones(int r, int c, int type)
    {
        assert(type == CV_32F);
        Mat t(r, c, CV_32F);
        for (int y = 0; y < r; ++y) for (int x = 0; x < c; ++x) t.at<float>(y, x) = 1.0f;
        return t;
    }
    Mat operator()(const Rect &r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    MatTExpr t() const { MatTExpr e = {this, 1.0}; return e; }
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }              // vector element
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step; }
    bool isContinuous() const { return step == (size_t)cols || rows == 1; }
    Size size() const { return Size(cols, rows); }

    uchar *ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar *ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T *ptr(int y = 0) { return (T *)(data + (size_t)y * step); }
    template <typename T> const T *ptr(int y = 0) const { return (const T *)(data + (size_t)y * step); }
    template <typename T> T &at(int y, int x) { return ((T *)(data + (size_t)y * step))[x]; }
    template <typename T> const T &at(int y, int x) const { return ((const T *)(data + (size_t)y * step))[x]; }

    int rows, cols;
    uchar *data;
    size_t step;

private:
    int type_;
    size_t esz_;
    std::shared_ptr<uchar> buf_;
};

// CV_32F matrix products as the projection matchers write them (src/ORBmatcher.cc:1968-1976, 1990-1991):
//   A*B + C  and  A*B   -> cv::gemm with A not transposed: OpenCV 4.13 accumulates in FLOAT, k ascending, every product and
//                          sum rounded separately, alpha / beta applied in float (pinned against live cv2, tests/test_oracle_cvprim.py)
//   -A.t()*B            -> cv::gemm with GEMM_1_T and alpha = -1: DOUBLE accumulation, one rounding
// Compile users of these with -ffp-contract=off.
struct MatMulExpr { const Mat *a, *b; };
static inline MatTExpr operator-(const MatTExpr &e) { MatTExpr r = {e.a, -e.alpha}; return r; }
static inline Mat mat_mul_eval(const Mat &A, const Mat &B, const Mat *C)
{
    assert(A.type() == CV_32F && B.type() == CV_32F && A.cols == B.rows);
    Mat t(A.rows, B.cols, CV_32F);
    for (int i = 0; i < A.rows; ++i)
        for (int j = 0; j < B.cols; ++j) {
            float s = 0.f;
            for (int k = 0; k < A.cols; ++k) { const float p = A.at<float>(i, k) * B.at<float>(k, j); s = s + p; }
            t.at<float>(i, j) = C ? s + C->at<float>(i, j) : s;
        }
    return t;
}
static inline MatMulExpr operator*(const Mat &a, const Mat &b) { MatMulExpr e = {&a, &b}; return e; }
static inline Mat operator+(const MatMulExpr &e, const Mat &c) { return mat_mul_eval(*e.a, *e.b, &c); }
static inline Mat operator*(const MatTExpr &e, const Mat &B)
{
    const Mat &A = *e.a;
    assert(A.type() == CV_32F && B.type() == CV_32F && A.rows == B.rows);
    Mat t(A.cols, B.cols, CV_32F);
    for (int i = 0; i < A.cols; ++i)
        for (int j = 0; j < B.cols; ++j) {
            double s = 0;
            for (int k = 0; k < A.rows; ++k) s += (double)A.at<float>(k, i) * (double)B.at<float>(k, j);
            t.at<float>(i, j) = (float)(s * e.alpha);
        }
    return t;
}

// CV_32F element-wise arithmetic used by Frame::ComputeStereoMatches (:951, :969, :971)
enum { NORM_L1 = 2 };
static inline Mat operator*(float s, const Mat &m)
{
    assert(m.type() == CV_32F);
    Mat t(m.rows, m.cols, CV_32F);
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) t.at<float>(y, x) = s * m.at<float>(y, x);
    return t;
}
static inline Mat operator-(const Mat &a, const Mat &b)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat t(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) t.at<float>(y, x) = a.at<float>(y, x) - b.at<float>(y, x);
    return t;
}
static inline double norm(const Mat &a, const Mat &b, int normType)
{
    assert(normType == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    (void)normType;
    double s = 0;                                   // cv::norm accumulates |a-b| of CV_32F in double
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) s += std::fabs((double)a.at<float>(y, x) - (double)b.at<float>(y, x));
    return s;
}

class _InputArray {
public:
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
protected:
    Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat &m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz, type); }
    void release() const { m_->release(); }
    Mat &getMatRef() const { return *m_; }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

static inline float fastAtan2(float y, float x) { return cvp_fast_atan2(y, x); }

static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double = 0, double = 0,
                          int interpolation = INTER_LINEAR)
{
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    Mat src = _src.getMat();
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    cvp_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                                  int borderType)
{
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    Mat src = _src.getMat();
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    cvp_border_reflect101_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step, top, bottom, left, right);
}

static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy = 0,
                                int borderType = BORDER_DEFAULT)
{
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat src = _src.getMat();
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    cvp_gaussian7x7_s2_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);
}

static inline void FAST(InputArray _img, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true)
{
    Mat img = _img.getMat();
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    std::vector<cvp_corner> tmp((size_t)img.cols * (size_t)img.rows / 2 + 16);
    int n = cvp_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0,
                         tmp.data(), (int)tmp.size());
    if (n > (int)tmp.size()) {           // only possible without NMS
        tmp.resize((size_t)n);
        n = cvp_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0,
                         tmp.data(), (int)tmp.size());
    }
    keypoints.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        keypoints.push_back(KeyPoint((float)tmp[i].x, (float)tmp[i].y, 7.f, -1.f,
                                     nonmaxSuppression ? (float)tmp[i].score : 0.f));
}

struct KeyPointsFilter {
    // keep the n strongest responses plus ties with the n-th (cv::KeyPointsFilter::retainBest)
    static void retainBest(std::vector<KeyPoint> &kps, int n)
    {
        if (n < 0 || (int)kps.size() <= n) return;
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(),
                         [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
        const float cut = kps[(size_t)n - 1].response;
        auto it = std::partition(kps.begin() + n, kps.end(), [cut](const KeyPoint &k) { return k.response >= cut; });
        kps.resize((size_t)(it - kps.begin()));
    }
};

} // namespace cv

#endif
