// multimot_track_b200/adapter/ORBmatcher_core.cc -- see ORBmatcher_core.h.
#include "ORBmatcher_core.h"

#include <stdexcept>
#include <string>

#include "orbx.h"

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = ORBX_TH_HIGH;
const int ORBmatcher::TH_LOW = ORBX_TH_LOW;
const int ORBmatcher::HISTO_LENGTH = ORBX_HISTO_LENGTH;

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    return orbx_hamming256(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

int BruteForceMatch(orbx_handle *handle, const cv::Mat &descA, const cv::Mat &descB, int th, float ratio,
                    std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist,
                    std::vector<unsigned char> &accepted)
{
    const int nA = descA.rows, nB = descB.rows;
    if ((nA && (descA.cols != 32 || !descA.isContinuous())) || (nB && (descB.cols != 32 || !descB.isContinuous())))
        throw std::invalid_argument("BruteForceMatch: descriptors must be continuous N x 32 CV_8U");
    bestIdx.assign(nA, -1); bestDist.assign(nA, 256); secondDist.assign(nA, 256); accepted.assign(nA, 0);
    if (nA == 0) return 0;
    const int rc = orbx_match(handle, descA.data, nA, descB.data, nB, th, ratio, bestIdx.data(), bestDist.data(), secondDist.data(),
                              accepted.data());
    if (rc < 0) throw std::runtime_error(std::string("orbx_match: ") + orbx_last_error(handle));
    return rc;
}

int RotationConsistencyFilter(orbx_handle *handle, const std::vector<int> &bestIdx, std::vector<unsigned char> &accepted,
                              const std::vector<float> &anglesA, const std::vector<float> &anglesB)
{
    const int nA = (int)bestIdx.size();
    if ((int)accepted.size() != nA || (int)anglesA.size() != nA)
        throw std::invalid_argument("RotationConsistencyFilter: bestIdx, accepted and anglesA need one entry per query");
    if (nA == 0) return 0;
    const int rc = orbx_rotation_filter(handle, nA, bestIdx.data(), accepted.data(), anglesA.data(), anglesB.data(), (int)anglesB.size(),
                                        nullptr, nullptr);
    if (rc < 0) throw std::runtime_error(std::string("orbx_rotation_filter: ") + orbx_last_error(handle));
    return rc;
}

int ComputeStereoMatchesGPU(orbx_handle *left, orbx_handle *right, float mbf, std::vector<float> &mvuRight,
                            std::vector<float> &mvDepth, std::vector<int> &vDescIndex)
{
    int n = 0;
    int rc = orbx_stereo_match(left, right, 0, 0, mbf, nullptr, nullptr, nullptr, 0, &n);      // N = left keypoints
    if (rc < 0) throw std::runtime_error(std::string("orbx_stereo_match: ") + orbx_last_error(left));
    mvuRight.assign(n, -1.0f); mvDepth.assign(n, -1.0f); vDescIndex.assign(n, -1);             // src/Frame.cc:852-854
    if (n == 0) return 0;
    rc = orbx_stereo_match(left, right, 0, 0, mbf, mvuRight.data(), mvDepth.data(), vDescIndex.data(), n, &n);
    if (rc < 0) throw std::runtime_error(std::string("orbx_stereo_match: ") + orbx_last_error(left));
    return rc;
}

} // namespace ORB_SLAM2
