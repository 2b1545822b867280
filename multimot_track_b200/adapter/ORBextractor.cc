// multimot_track_b200/adapter/ORBextractor.cc -- the reference's ORBextractor surface over
// the GPU library.  Error behaviour mirrors src/ORBextractor.cc: an empty image returns
// silently (:1049), a non-CV_8UC1 image trips the assert (:1053), zero keypoints release the
// descriptor matrix (:1068-1069).  A failing GPU call is not silent: it throws
// std::runtime_error with the library's message (the reference has no such failure mode).
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbx.h"

namespace ORB_SLAM2 {

static const int EDGE_THRESHOLD = 19;

static void check(orbx_handle *h, int rc, const char *what)
{
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + orbx_last_error(h));
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(NULL), mbExportPyramid(true)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "orbx_keypoint must have cv::KeyPoint's layout");
    orbx_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.nfeatures = _nfeatures; cfg.scale_factor = _scaleFactor; cfg.nlevels = _nlevels;
    cfg.ini_th_fast = _iniThFAST; cfg.min_th_fast = _minThFAST; cfg.device_id = -1;
    check(NULL, orbx_create(&cfg, &mpHandle), "orbx_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orbx_get_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                    mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_destroy(mpHandle); }

const char *ORBextractor::LastError() const { return orbx_last_error(mpHandle); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty())
        return;

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    const int cap = orbx_max_keypoints(mpHandle, image.cols, image.rows);
    check(mpHandle, cap, "orbx_max_keypoints");
    _keypoints.resize(cap);
    cv::Mat all(cap, 32, CV_8U);
    int n = 0;
    check(mpHandle, orbx_extract(mpHandle, image.data, image.cols, image.rows, (int)image.step,
                                 reinterpret_cast<orbx_keypoint *>(_keypoints.data()), all.data, cap, &n), "orbx_extract");
    _keypoints.resize(n);
    if (n == 0)
        _descriptors.release();
    else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), all.ptr(i), 32);
    }

    if (mbExportPyramid) {
        // mvImagePyramid (src/ORBextractor.cc:1111-1136): every level a view into a (w+38)x(h+38) parent with the reflect-101 frame.
        // All parents live side by side in ONE refcounted matrix and arrive with one synchronisation (orbx_get_pyramid_levels);
        // the views keep that matrix alive the way the reference's per-level temporaries are kept alive by their views.
        std::vector<int> lw(nlevels), lh(nlevels), xoff(nlevels), strides(nlevels);
        int total_w = 0, max_h = 0;
        for (int level = 0; level < nlevels; ++level) {
            check(mpHandle, orbx_get_level_size(mpHandle, level, &lw[level], &lh[level]), "orbx_get_level_size");
            xoff[level] = total_w;
            total_w += lw[level] + 2 * EDGE_THRESHOLD;
            if (lh[level] + 2 * EDGE_THRESHOLD > max_h) max_h = lh[level] + 2 * EDGE_THRESHOLD;
        }
        cv::Mat store(max_h, total_w, CV_8UC1);
        std::vector<uint8_t *> dst(nlevels);
        for (int level = 0; level < nlevels; ++level) { dst[level] = store.data + xoff[level]; strides[level] = (int)store.step; }
        check(mpHandle, orbx_get_pyramid_levels(mpHandle, 0, nlevels, dst.data(), strides.data(), 1), "orbx_get_pyramid_levels");
        for (int level = 0; level < nlevels; ++level)
            mvImagePyramid[level] = store(cv::Rect(xoff[level] + EDGE_THRESHOLD, EDGE_THRESHOLD, lw[level], lh[level]));
    }
}

} // namespace ORB_SLAM2
