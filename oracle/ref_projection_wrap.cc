// oracle/ref_projection_wrap.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's own ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame,
// th, bMono, ...) (src/ORBmatcher.cc:1958-2102) and the Frame grid it searches (src/Frame.cc:601-616, 710-776), all
// excerpted unmodified at build time (oracle/Makefile) and compiled against oracle/stereo_shim.hpp + oracle/minicv.
#include <cstring>

#include "stereo_shim.hpp"

namespace {
void fill_frame(ORB_SLAM2::Frame &F, const float *cam, const float *Tcw, int n, const float *xy_un, const int32_t *octave, const float *angle,
                const float *uright, const uint8_t *desc, const float *scale, int nlevels)
{
    F.fx = cam[0]; F.fy = cam[1]; F.cx = cam[2]; F.cy = cam[3]; F.mbf = cam[4]; F.mb = cam[5];
    F.mnMinX = cam[6]; F.mnMaxX = cam[7]; F.mnMinY = cam[8]; F.mnMaxY = cam[9];
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (F.mnMaxX - F.mnMinX);      // src/Frame.cc:126-127
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (F.mnMaxY - F.mnMinY);
    F.mTcw = cv::Mat(4, 4, CV_32F);
    for (int i = 0; i < 16; ++i) F.mTcw.at<float>(i / 4, i % 4) = Tcw[i];
    F.N = n;
    F.mvKeys.resize((size_t)n); F.mvKeysUn.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        F.mvKeysUn[i] = cv::KeyPoint(xy_un[2 * i], xy_un[2 * i + 1], 31.f, angle[i], 0.f, octave[i], -1);
        F.mvKeys[i] = F.mvKeysUn[i];
    }
    F.mvuRight.assign(uright, uright + n);
    F.mDescriptors = cv::Mat(n, 32, CV_8UC1, (void *)desc, 32);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvpMapPoints.assign((size_t)n, (ORB_SLAM2::MapPoint *)0);
    F.mvbOutlier.assign((size_t)n, false);
}
}

extern "C" {

// cam: fx fy cx cy mbf mb mnMinX mnMaxX mnMinY mnMaxY.  Last frame: world_pos [nL][3], mp_desc [nL][32], valid[i] = has a map point
// and is not an outlier, nobs[i] = MapPoint::Observations().  cur_match[nC] receives, per current feature, the index of the last-frame
// point whose MapPoint it ends up holding (-1 = none).  Returns nmatches.
int orbref_search_by_projection(const float *cam, const float *Tcw_cur, const float *Tcw_last,
                                int nL, const float *world_pos, const uint8_t *mp_desc, const uint8_t *valid, const int32_t *nobs,
                                const int32_t *last_octave, const float *last_angle,
                                int nC, const float *cur_xy_un, const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                                const uint8_t *cur_desc, const float *scale, int nlevels, float th, int mono, int check_orientation,
                                int32_t *cur_match)
{
    ORB_SLAM2::Frame Cur, Last;
    std::vector<float> zeros((size_t)nL, -1.0f), xyl((size_t)nL * 2, 0.f);
    fill_frame(Cur, cam, Tcw_cur, nC, cur_xy_un, cur_octave, cur_angle, cur_uright, cur_desc, scale, nlevels);
    fill_frame(Last, cam, Tcw_last, nL, xyl.data(), last_octave, last_angle, zeros.data(), mp_desc, scale, nlevels);
    Cur.AssignFeaturesToGrid();
    std::vector<ORB_SLAM2::MapPoint> pts((size_t)nL);
    for (int i = 0; i < nL; ++i) {
        pts[i].mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) pts[i].mWorldPos.at<float>(k) = world_pos[3 * i + k];
        pts[i].mDescriptor = cv::Mat(1, 32, CV_8UC1, (void *)(mp_desc + 32 * (size_t)i), 32);
        pts[i].nObs = nobs[i];
        if (valid[i]) Last.mvpMapPoints[i] = &pts[i];
    }
    ORB_SLAM2::ORBmatcher matcher(0.9f, check_orientation != 0);
    std::vector<int> temporal;
    const int n = matcher.SearchByProjection(Cur, Last, th, mono != 0, temporal);
    for (int i = 0; i < nC; ++i) cur_match[i] = Cur.mvpMapPoints[i] ? (int32_t)(Cur.mvpMapPoints[i] - &pts[0]) : -1;
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:780-895).
// xy1 / xy2: mvKeysUn positions; prev_xy [n1][2]: vbPrevMatched, updated in place as the reference does.  Returns nmatches.
int orbref_search_for_initialization(const float *cam, int n1, const float *xy1, const int32_t *oct1, const float *ang1, const uint8_t *desc1,
                                     int n2, const float *xy2, const int32_t *oct2, const float *ang2, const uint8_t *desc2,
                                     float *prev_xy, int window, float nnratio, int check_orientation, int32_t *matches12)
{
    ORB_SLAM2::Frame F1, F2;
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, one[1] = {1.0f};
    std::vector<float> ur1((size_t)n1, -1.0f), ur2((size_t)n2, -1.0f);
    fill_frame(F1, cam, eye, n1, xy1, oct1, ang1, ur1.data(), desc1, one, 1);
    fill_frame(F2, cam, eye, n2, xy2, oct2, ang2, ur2.data(), desc2, one, 1);
    F2.AssignFeaturesToGrid();
    std::vector<cv::Point2f> prev((size_t)n1);
    for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORB_SLAM2::ORBmatcher matcher(nnratio, check_orientation != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return n;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (src/ORBmatcher.cc:418-502), the matcher of
// Tracking::SearchLocalPoints (src/Tracking.cc:3464).  proj [nP][3] = mTrackProjX, mTrackProjY, mTrackProjXR; valid[i] = mbTrackInView
// && !isBad(); feat_obs[j] = Observations() of the map point feature j already holds (<= 0: none / unobserved).  feat_match[j] receives
// the index of the map point the function assigns to feature j (-1 = left as it was).  Returns nmatches.
int orbref_search_local_points(const float *cam, int nP, const float *proj, const float *view_cos, const int32_t *level, const uint8_t *mp_desc,
                               const uint8_t *valid, const int32_t *nobs, int nF, const float *xy_un, const int32_t *octave, const float *uright,
                               const uint8_t *desc, const int32_t *feat_obs, const float *scale, int nlevels, float th, float nnratio,
                               int32_t *feat_match)
{
    ORB_SLAM2::Frame F;
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<float> ang((size_t)nF, 0.f);
    fill_frame(F, cam, eye, nF, xy_un, octave, ang.data(), uright, desc, scale, nlevels);
    F.AssignFeaturesToGrid();
    std::vector<ORB_SLAM2::MapPoint> pts((size_t)nP), held((size_t)nF);
    std::vector<ORB_SLAM2::MapPoint *> vp((size_t)nP);
    for (int i = 0; i < nP; ++i) {
        pts[i].mDescriptor = cv::Mat(1, 32, CV_8UC1, (void *)(mp_desc + 32 * (size_t)i), 32);
        pts[i].nObs = nobs[i];
        pts[i].mbTrackInView = valid[i] != 0;
        pts[i].mnTrackScaleLevel = level[i];
        pts[i].mTrackViewCos = view_cos[i];
        pts[i].mTrackProjX = proj[3 * i]; pts[i].mTrackProjY = proj[3 * i + 1]; pts[i].mTrackProjXR = proj[3 * i + 2];
        vp[i] = &pts[i];
    }
    for (int j = 0; j < nF; ++j)
        if (feat_obs[j] >= 0) { held[j].nObs = feat_obs[j]; F.mvpMapPoints[j] = &held[j]; }
    ORB_SLAM2::ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, vp, th);
    for (int j = 0; j < nF; ++j) {
        ORB_SLAM2::MapPoint *p = F.mvpMapPoints[j];
        feat_match[j] = (p && p >= &pts[0] && p < &pts[0] + nP) ? (int32_t)(p - &pts[0]) : -1;
    }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches, TemperalMatch) (src/ORBmatcher.cc:532-663).  Feature vectors
// flattened in map order: node j owns feats[off[j] .. off[j+1]).  kf_valid[i] = the key frame's feature i has a map point that is not
// bad.  f_match[j] receives the key-frame feature whose map point feature j of F ends up with (-1 = none).  Returns nmatches.
int orbref_search_by_bow(int nK, const float *kf_angle, const uint8_t *kf_desc, const uint8_t *kf_valid, int nk_nodes, const int32_t *kf_nodes,
                         const int32_t *kf_off, const int32_t *kf_feats, int nF, const float *f_angle, const uint8_t *f_desc, int nf_nodes,
                         const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_feats, float nnratio, int check_orientation, int32_t *f_match)
{
    ORB_SLAM2::KeyFrame KF;
    ORB_SLAM2::Frame F;
    std::vector<ORB_SLAM2::MapPoint> pts((size_t)nK);
    KF.mvpMapPoints.assign((size_t)nK, (ORB_SLAM2::MapPoint *)0);
    KF.mvKeysUn.resize((size_t)nK);
    for (int i = 0; i < nK; ++i) {
        KF.mvKeysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, kf_angle[i], 0.f, 0, -1);
        if (kf_valid[i]) KF.mvpMapPoints[i] = &pts[i];
    }
    KF.mDescriptors = cv::Mat(nK, 32, CV_8UC1, (void *)kf_desc, 32);
    for (int j = 0; j < nk_nodes; ++j)
        for (int k = kf_off[j]; k < kf_off[j + 1]; ++k) KF.mFeatVec[(unsigned)kf_nodes[j]].push_back((unsigned)kf_feats[k]);
    F.N = nF;
    F.mvKeys.resize((size_t)nF);
    for (int i = 0; i < nF; ++i) F.mvKeys[i] = cv::KeyPoint(0.f, 0.f, 31.f, f_angle[i], 0.f, 0, -1);
    F.mDescriptors = cv::Mat(nF, 32, CV_8UC1, (void *)f_desc, 32);
    for (int j = 0; j < nf_nodes; ++j)
        for (int k = f_off[j]; k < f_off[j + 1]; ++k) F.mFeatVec[(unsigned)f_nodes[j]].push_back((unsigned)f_feats[k]);
    ORB_SLAM2::ORBmatcher matcher(nnratio, check_orientation != 0);
    std::vector<ORB_SLAM2::MapPoint *> out;
    std::vector<int> temporal;
    const int n = matcher.SearchByBoW(&KF, F, out, temporal);
    for (int j = 0; j < nF; ++j) f_match[j] = out[j] ? (int32_t)(out[j] - &pts[0]) : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:897-1030).  valid1 / valid2: the feature has a
// map point that is not bad.  m12[i] receives the feature of key frame 2 whose map point feature i of key frame 1 is matched to.
int orbref_search_by_bow_kf(int n1, const float *ang1, const uint8_t *desc1, const uint8_t *valid1, int nn1, const int32_t *nodes1, const int32_t *off1,
                            const int32_t *feats1, int n2, const float *ang2, const uint8_t *desc2, const uint8_t *valid2, int nn2, const int32_t *nodes2,
                            const int32_t *off2, const int32_t *feats2, float nnratio, int check_orientation, int32_t *m12)
{
    ORB_SLAM2::KeyFrame K1, K2;
    std::vector<ORB_SLAM2::MapPoint> p1((size_t)n1), p2((size_t)n2);
    auto fill = [](ORB_SLAM2::KeyFrame &K, std::vector<ORB_SLAM2::MapPoint> &pts, int n, const float *ang, const uint8_t *desc, const uint8_t *valid,
                   int nn, const int32_t *nodes, const int32_t *off, const int32_t *feats) {
        K.mvpMapPoints.assign((size_t)n, (ORB_SLAM2::MapPoint *)0);
        K.mvKeysUn.resize((size_t)n);
        for (int i = 0; i < n; ++i) {
            K.mvKeysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, ang[i], 0.f, 0, -1);
            if (valid[i]) K.mvpMapPoints[i] = &pts[i];
        }
        K.mDescriptors = cv::Mat(n, 32, CV_8UC1, (void *)desc, 32);
        for (int j = 0; j < nn; ++j)
            for (int k = off[j]; k < off[j + 1]; ++k) K.mFeatVec[(unsigned)nodes[j]].push_back((unsigned)feats[k]);
    };
    fill(K1, p1, n1, ang1, desc1, valid1, nn1, nodes1, off1, feats1);
    fill(K2, p2, n2, ang2, desc2, valid2, nn2, nodes2, off2, feats2);
    ORB_SLAM2::ORBmatcher matcher(nnratio, check_orientation != 0);
    std::vector<ORB_SLAM2::MapPoint *> out;
    const int n = matcher.SearchByBoW(&K1, &K2, out);
    for (int i = 0; i < n1; ++i) m12[i] = out[i] ? (int32_t)(out[i] - &p2[0]) : -1;
    return n;
}

} // extern "C"
