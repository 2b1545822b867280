// Compile-time check that the adapter keeps the call shapes the reference uses:
//   (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors)          src/Frame.cc:621
//   mpORBextractorLeft->GetLevels() ... GetInverseScaleSigmaSquares()    src/Frame.cc:87-93
//   mpORBextractorLeft->mvImagePyramid[kpL.octave].rowRange(..)          src/Frame.cc:960
//   ORBmatcher::DescriptorDistance(dL, dR), ORBmatcher::TH_HIGH/TH_LOW   src/Frame.cc:868,931
// Built with -fsyntax-only against oracle/minicv by tests/test_abi.py.
#include "ORBextractor.h"
#include "ORBmatcher_core.h"

namespace {
void frame_like_usage(cv::Mat &im, std::vector<cv::KeyPoint> &mvKeys, cv::Mat &mDescriptors)
{
    ORB_SLAM2::ORBextractor *mpORBextractorLeft = new ORB_SLAM2::ORBextractor(2000, 1.2f, 8, 20, 7);
    (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
    int mnScaleLevels = mpORBextractorLeft->GetLevels();
    float mfScaleFactor = mpORBextractorLeft->GetScaleFactor();
    std::vector<float> mvScaleFactors = mpORBextractorLeft->GetScaleFactors();
    std::vector<float> mvInvScaleFactors = mpORBextractorLeft->GetInverseScaleFactors();
    std::vector<float> mvLevelSigma2 = mpORBextractorLeft->GetScaleSigmaSquares();
    std::vector<float> mvInvLevelSigma2 = mpORBextractorLeft->GetInverseScaleSigmaSquares();
    const int nRows = mpORBextractorLeft->mvImagePyramid[0].rows;
    const int thOrbDist = (ORB_SLAM2::ORBmatcher::TH_HIGH + ORB_SLAM2::ORBmatcher::TH_LOW) / 2;
    const cv::Mat &dL = mDescriptors.row(0), &dR = mDescriptors.row(1);
    const int dist = ORB_SLAM2::ORBmatcher::DescriptorDistance(dL, dR);
    cv::Mat IL = mpORBextractorLeft->mvImagePyramid[0].rowRange(20, 31).colRange(20, 31);
    (void)mnScaleLevels; (void)mfScaleFactor; (void)nRows; (void)thOrbDist; (void)dist; (void)IL;
    (void)ORB_SLAM2::ORBextractor::FAST_SCORE;
    mpORBextractorLeft->SetPyramidExport(false);
}
}
