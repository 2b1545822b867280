// multimot_track_b200/adapter/ORBextractor.h -- drop-in replacement for the reference's
// include/ORBextractor.h.  Same namespace, class name, constructor, call operator,
// getters, enum and public mvImagePyramid (include/ORBextractor.h:45-111 of the
// reference), so Frame::ExtractORB (src/Frame.cc:618-632), the Frame constructors
// (src/Frame.cc:87-93,170-176) and Tracking (src/Tracking.cc:202-214) compile and run
// unchanged; the work happens on the GPU behind the C ABI of include/orbx.h.
//
// With real OpenCV this header includes <opencv2/core/core.hpp>; in this repository it is
// compile-checked against the oracle/minicv shim (tests/test_abi.py), because the OpenCV
// C++ library is not installed here.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>

#include <opencv2/core/core.hpp>

extern "C" {
struct orbx_handle;
}

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // Compute the ORB features and descriptors on an image (mask ignored, as in the reference).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                    cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Host copies of the pyramid levels of the last frame, each a view into a (w+38)x(h+38)
    // buffer with the reference's 19-pixel BORDER_REFLECT_101 frame, exactly what
    // Frame::ComputeStereoMatches reads (src/Frame.cc:859,960-977).
    std::vector<cv::Mat> mvImagePyramid;

    // B200 extension: only the stereo path reads mvImagePyramid on the host.  RGB-D and monocular
    // callers may switch the per-frame device-to-host copy of the pyramid off.
    void SetPyramidExport(bool on) { mbExportPyramid = on; }

    // Last error text from the GPU library (empty when the last call succeeded).
    const char *LastError() const;

    // B200 extension: the GPU handle that still holds the last frame's keypoints, descriptors and pyramid on the
    // device (for ComputeStereoMatchesGPU / BruteForceMatch without a round trip through the host).
    orbx_handle *Handle() const { return mpHandle; }

protected:
    ORBextractor(const ORBextractor &);              // a handle owns GPU memory: not copyable
    ORBextractor &operator=(const ORBextractor &);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbx_handle *mpHandle;
    bool mbExportPyramid;
};

} // namespace ORB_SLAM2

#endif
