// oracle/ref_wrap.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's own ORB_SLAM2::ORBextractor
// (src/ORBextractor.cc, include/ORBextractor.h, compiled unmodified against
// oracle/minicv) and ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2279-2295,
// excerpted at build time by oracle/Makefile).  Built into oracle/_ref/*.so.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "ORBextractor.h"

#include "stereo_shim.hpp"            // the ORBmatcher declaration shared by every oracle translation unit

namespace {
// DistributeOctTree is protected: reach it through a subclass, no source edits.
class RefExtractor : public ORB_SLAM2::ORBextractor {
public:
    using ORB_SLAM2::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint> &keys, int minX, int maxX, int minY,
                                         int maxY, int N, int level)
    {
        return DistributeOctTree(keys, minX, maxX, minY, maxY, N, level);
    }
    const std::vector<int> &featuresPerLevel() const { return mnFeaturesPerLevel; }
    const std::vector<int> &uMax() const { return umax; }
};
}

extern "C" {

void *orbref_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    return new RefExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}

void orbref_destroy(void *h) { delete (RefExtractor *)h; }

// Runs operator() (src/ORBextractor.cc:1046-1109).  kps: n x 7 floats
// (x, y, size, angle, response, octave, class_id).  Returns n (<= cap) or -n if cap too small.
int orbref_extract(void *h, const uint8_t *gray, int w, int hgt, int stride, float *kps, uint8_t *desc, int cap)
{
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img(hgt, w, CV_8UC1, (void *)gray, (size_t)stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    (*e)(img, cv::Mat(), keys, d);
    const int n = (int)keys.size();
    if (n > cap) return -n;
    for (int i = 0; i < n; ++i) {
        float *o = kps + 7 * (size_t)i;
        o[0] = keys[i].pt.x; o[1] = keys[i].pt.y; o[2] = keys[i].size; o[3] = keys[i].angle;
        o[4] = keys[i].response; o[5] = (float)keys[i].octave; o[6] = (float)keys[i].class_id;
    }
    for (int i = 0; i < n; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    return n;
}

// Level geometry / bytes of mvImagePyramid after the last extract.  with_border
// copies the (w+38)x(h+38) parent buffer (what Frame::ComputeStereoMatches may touch).
int orbref_pyramid_level(void *h, int level, uint8_t *dst, int dst_stride, int *w, int *hgt, int with_border)
{
    RefExtractor *e = (RefExtractor *)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size()) return -1;
    const cv::Mat &m = e->mvImagePyramid[level];
    if (m.empty()) return -2;
    const int b = with_border ? 19 : 0;
    *w = m.cols + 2 * b; *hgt = m.rows + 2 * b;
    if (dst)
        for (int y = 0; y < *hgt; ++y)
            std::memcpy(dst + (size_t)y * dst_stride, m.data + (ptrdiff_t)(y - b) * (ptrdiff_t)m.step - b, (size_t)*w);
    return 0;
}

int orbref_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *nfeat, int *umax16)
{
    RefExtractor *e = (RefExtractor *)h;
    const int L = e->GetLevels();
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                       d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < L; ++i) {
        scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i];
        nfeat[i] = e->featuresPerLevel()[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->uMax()[i];
    return L;
}

// DistributeOctTree alone (src/ORBextractor.cc:539-763).  in: n x 3 floats (x, y, response)
// relative to (minX, minY); out: m x 3 floats.  Returns m.
int orbref_distribute(void *h, const float *xyr, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                      float *out, int cap)
{
    RefExtractor *e = (RefExtractor *)h;
    std::vector<cv::KeyPoint> keys((size_t)n);
    for (int i = 0; i < n; ++i)
        keys[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]);
    std::vector<cv::KeyPoint> r = e->distribute(keys, minX, maxX, minY, maxY, N, level);
    const int m = (int)r.size();
    if (m > cap) return -m;
    for (int i = 0; i < m; ++i) { out[3 * i] = r[i].pt.x; out[3 * i + 1] = r[i].pt.y; out[3 * i + 2] = r[i].response; }
    return m;
}

int orbref_descriptor_distance(const uint8_t *a, const uint8_t *b)
{
    // 32-byte rows, read as 8 x int32 like the reference does; copy to aligned storage first.
    alignas(16) uint8_t ta[32], tb[32];
    std::memcpy(ta, a, 32); std::memcpy(tb, b, 32);
    cv::Mat ma(1, 32, CV_8UC1, ta), mb(1, 32, CV_8UC1, tb);
    return ORB_SLAM2::ORBmatcher::DescriptorDistance(ma, mb);
}

// mbCheckOrientation around the reference's own ComputeThreeMaxima.  The histogram fill and the pruning are call-site
// code inside SearchByBoW (src/ORBmatcher.cc:545, 610-620, 641-660); they are restated here line for line.
int orbref_rotation_filter(int nA, const int32_t *idx, uint8_t *accept, const float *angleA, const float *angleB,
                           int32_t *hist, int32_t *top3)
{
    const int HISTO_LENGTH = ORB_SLAM2::ORBmatcher::HISTO_LENGTH;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nA; ++i) {
        if (!accept[i]) continue;
        float rot = angleA[i] - angleB[idx[i]];
        if (rot < 0.0) rot += 360.0f;
        int bin = round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(i);
    }
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ORB_SLAM2::ORBmatcher m;
    m.ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    int kept = 0;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i == ind1 || i == ind2 || i == ind3) { kept += (int)rotHist[i].size(); continue; }
        for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) accept[rotHist[i][j]] = 0;
    }
    if (hist) for (int i = 0; i < HISTO_LENGTH; ++i) hist[i] = (int32_t)rotHist[i].size();
    if (top3) { top3[0] = ind1; top3[1] = ind2; top3[2] = ind3; }
    return kept;
}

void orbref_thresholds(int *th_low, int *th_high, int *histo_length)
{
    *th_low = ORB_SLAM2::ORBmatcher::TH_LOW;
    *th_high = ORB_SLAM2::ORBmatcher::TH_HIGH;
    *histo_length = ORB_SLAM2::ORBmatcher::HISTO_LENGTH;
}

} // extern "C"

// CPU baseline runner: one ORBextractor instance and one frame at a time per thread
// (the reference's own threading contract, src/Frame.cc:96-99).  Returns wall seconds.
#include <atomic>
#include <chrono>
#include <thread>
extern "C" double orbref_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                                      const uint8_t *frames, int nframes, int w, int h, int threads, long *total_kps)
{
    if (threads < 1) threads = 1;
    std::atomic<int> next(0);
    std::atomic<long> kps(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            RefExtractor e(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
            for (;;) {
                const int f = next.fetch_add(1);
                if (f >= nframes) break;
                cv::Mat img(h, w, CV_8UC1, (void *)(frames + (size_t)f * w * h), (size_t)w);
                std::vector<cv::KeyPoint> keys;
                cv::Mat d;
                e(img, cv::Mat(), keys, d);
                kps += (long)keys.size();
            }
        });
    for (auto &th : pool) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (total_kps) *total_kps = kps.load();
    return std::chrono::duration<double>(t1 - t0).count();
}
