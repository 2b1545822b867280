// oracle/stereo_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The slice of ORB_SLAM2::Frame (include/Frame.h:43-295 of the reference) that Frame::ComputeStereoMatches
// (src/Frame.cc:849-1038) touches, so that the function body compiles UNMODIFIED from an excerpt made at build time
// (oracle/Makefile).  Member names and types are the reference's; everything else of Frame is left out.
#ifndef ORACLE_STEREO_SHIM_HPP
#define ORACLE_STEREO_SHIM_HPP
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include "minicv.hpp"

namespace ORB_SLAM2 {

class ORBmatcher {                       // same declaration as in ref_wrap.cc / the matcher excerpt
public:
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);
};

struct PyramidHolder { std::vector<cv::Mat> mvImagePyramid; };   // the one ORBextractor member the function reads

class Frame {
public:
    void ComputeStereoMatches();         // src/Frame.cc:849-1038
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvuRight, mvDepth;
    std::vector<int> vDescIndex;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    PyramidHolder *mpORBextractorLeft, *mpORBextractorRight;
    float mbf;
};

}
#endif
