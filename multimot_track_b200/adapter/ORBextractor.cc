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
        for (int level = 0; level < nlevels; ++level) {
            int w = 0, h = 0;
            check(mpHandle, orbx_get_level_size(mpHandle, level, &w, &h), "orbx_get_level_size");
            cv::Mat temp(h + 2 * EDGE_THRESHOLD, w + 2 * EDGE_THRESHOLD, CV_8UC1);
            check(mpHandle, orbx_get_pyramid_level(mpHandle, 0, level, temp.data, (int)temp.step, 1), "orbx_get_pyramid_level");
            mvImagePyramid[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }
    }
}

} // namespace ORB_SLAM2
