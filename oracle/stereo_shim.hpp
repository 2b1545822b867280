// oracle/stereo_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The slices of ORB_SLAM2::Frame (include/Frame.h:43-295 of the reference), MapPoint (include/MapPoint.h) and ORBmatcher
// (include/ORBmatcher.h:37-109) that the excerpted reference functions touch -- Frame::ComputeStereoMatches
// (src/Frame.cc:849-1038), Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid (:601-616, 710-776),
// ORBmatcher::SearchByProjection(Frame&, const Frame&, ...) (src/ORBmatcher.cc:1958-2102), SearchForInitialization (:780-895),
// SearchByProjection(Frame&, const vector<MapPoint*>&, th) with RadiusByViewingCos (:418-511), SearchByBoW(KeyFrame*, Frame&, ...) (:532-663), SearchByBoW(KeyFrame*, KeyFrame*, ...) (:897-1030), ComputeThreeMaxima and
// DescriptorDistance -- so that those function bodies compile UNMODIFIED from excerpts made at build time
// (oracle/Makefile).  Member names and types are the reference's; everything else of the classes is left out.
#ifndef ORACLE_STEREO_SHIM_HPP
#define ORACLE_STEREO_SHIM_HPP
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "minicv.hpp"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {

class Frame;

class MapPoint {                         // what SearchByProjection reads of a map point (include/MapPoint.h:44-47, 59, 64)
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return bad; }
    cv::Mat mWorldPos, mDescriptor;      // 3x1 CV_32F, 1x32 CV_8U
    int nObs;
    // what Frame::isInFrustum leaves for SearchByProjection(Frame&, vector<MapPoint*>&, th) (include/MapPoint.h:88-94)
    float mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos;
    bool mbTrackInView = false, bad = false;
    int mnTrackScaleLevel = 0;
};

}
namespace DBoW2 {                        // Thirdparty/DBoW2/DBoW2/FeatureVector.h:24-27: node id -> indices of the local features under it
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}
namespace ORB_SLAM2 {

class KeyFrame {                         // what SearchByBoW(KeyFrame*, Frame&, ...) reads of a key frame (include/KeyFrame.h)
public:
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<MapPoint *> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
};

class ORBmatcher {                       // the one declaration every oracle translation unit uses
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono, std::vector<int> &TemperalMatch);
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3);
    float RadiusByViewingCos(const float &viewCos);
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches, std::vector<int> &TemperalMatch);
    int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12);
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

struct PyramidHolder { std::vector<cv::Mat> mvImagePyramid; };   // the one ORBextractor member the function reads

class Frame {
public:
    void ComputeStereoMatches();         // src/Frame.cc:849-1038
    void AssignFeaturesToGrid();         // :601-616
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const;   // :710-763
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);   // :765-776
    // static in the reference (include/Frame.h:120-125, 245-246, 269-272); plain members here, the access syntax is the same
    float fx, fy, cx, cy, mb;
    float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    cv::Mat mTcw;                        // 4x4 CV_32F
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvuRight, mvDepth;
    std::vector<int> vDescIndex;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    PyramidHolder *mpORBextractorLeft, *mpORBextractorRight;
    float mbf;
    DBoW2::FeatureVector mFeatVec;
};

}
#endif
