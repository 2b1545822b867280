// multimot_track_b200/adapter/ORBmatcher_core.h -- the hot core of the reference's ORBmatcher
// (include/ORBmatcher.h:44,94-96): the static DescriptorDistance, the thresholds, and the
// best / second-best / ratio scan every SearchBy* loop shares (src/ORBmatcher.cc:574-605) as
// one all-pairs GPU call.  A maintainer replaces the definitions at src/ORBmatcher.cc:41-43
// and :2279-2295 with ORBmatcher_core.cc and keeps the rest of ORBmatcher untouched.
#ifndef ORBMATCHER_CORE_H
#define ORBMATCHER_CORE_H

#include <vector>

#include <opencv2/core/core.hpp>

extern "C" {
struct orbx_handle;
}

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    // Computes the Hamming distance between two ORB descriptors (1x32 CV_8U rows).
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
};

// All-pairs nearest / second-nearest match of descriptor matrices (N x 32 CV_8U) on the GPU with the
// reference's update rule and acceptance test  best <= th && (float)best < ratio*(float)second.
// Outputs are resized to descA.rows; bestIdx[i] = -1 when descB is empty.  Returns the accepted count.
int BruteForceMatch(orbx_handle *handle, const cv::Mat &descA, const cv::Mat &descB, int th, float ratio,
                    std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist,
                    std::vector<unsigned char> &accepted);

// mbCheckOrientation of the matchers (src/ORBmatcher.cc:610-620 histogram, :2233-2274 ComputeThreeMaxima, :641-660
// pruning) for the matches of BruteForceMatch: accepted[i] is cleared when the rotation anglesA[i] - anglesB[bestIdx[i]]
// does not fall into one of the three most populated bins.  Returns the matches kept.
int RotationConsistencyFilter(orbx_handle *handle, const std::vector<int> &bestIdx, std::vector<unsigned char> &accepted,
                              const std::vector<float> &anglesA, const std::vector<float> &anglesB);

// Frame::ComputeStereoMatches (src/Frame.cc:849-1038) on the GPU, from the device-resident results of the two
// extractors that just processed the left and the right image (Frame::Frame stereo constructor, src/Frame.cc:96-104).
// Fills mvuRight, mvDepth and vDescIndex exactly as the reference does (N = number of left keypoints); mbf as in Frame.
// Returns the number of stereo matches kept.  The body of Frame::ComputeStereoMatches becomes
//     ComputeStereoMatchesGPU(mpORBextractorLeft->Handle(), mpORBextractorRight->Handle(), mbf, mvuRight, mvDepth, vDescIndex);
int ComputeStereoMatchesGPU(orbx_handle *left, orbx_handle *right, float mbf, std::vector<float> &mvuRight,
                            std::vector<float> &mvDepth, std::vector<int> &vDescIndex);

} // namespace ORB_SLAM2

#endif
