// oracle/ref_stereo_wrap.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's own Frame::ComputeStereoMatches (src/Frame.cc:849-1038, excerpted
// unmodified at build time by oracle/Makefile and compiled against oracle/stereo_shim.hpp + oracle/minicv).
#include <cstring>

#include "stereo_shim.hpp"

extern "C" {

// kps*: n x 7 floats (x, y, size, angle, response, octave, class_id) as orbref_extract returns them; desc*: n x 32 bytes;
// pyr*: per level a pointer to the un-padded image, its pitch and size.  Outputs have nL entries.
// Returns the number of left keypoints with a stereo match (mvuRight >= 0), or -1 when the reference would have
// indexed an empty vector (no match survived; undefined behaviour there, src/Frame.cc:1024).
int orbref_stereo_matches(int nL, const float *kpsL, const uint8_t *descL, int nR, const float *kpsR, const uint8_t *descR,
                          int nlevels, const float *scale, const float *inv_scale,
                          const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *pitch, const int *lw, const int *lh,
                          float bf, float *u_right, float *depth, int32_t *desc_index)
{
    ORB_SLAM2::Frame F;
    ORB_SLAM2::PyramidHolder L, R;
    F.N = nL; F.mbf = bf;
    F.mpORBextractorLeft = &L; F.mpORBextractorRight = &R;
    auto fill = [](std::vector<cv::KeyPoint> &v, const float *k, int n) {
        v.resize((size_t)n);
        for (int i = 0; i < n; ++i)
            v[i] = cv::KeyPoint(k[7 * i], k[7 * i + 1], k[7 * i + 2], k[7 * i + 3], k[7 * i + 4], (int)k[7 * i + 5], (int)k[7 * i + 6]);
    };
    fill(F.mvKeys, kpsL, nL); fill(F.mvKeysRight, kpsR, nR);
    F.mDescriptors = cv::Mat(nL, 32, CV_8UC1, (void *)descL, 32);
    F.mDescriptorsRight = cv::Mat(nR, 32, CV_8UC1, (void *)descR, 32);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
    for (int l = 0; l < nlevels; ++l) {
        L.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8UC1, (void *)pyrL[l], (size_t)pitch[l]));
        R.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8UC1, (void *)pyrR[l], (size_t)pitch[l]));
    }
    // the reference sorts and then reads vDistIdx[size/2] without checking for an empty vector; probe with a dry run
    // is not possible without editing it, so the caller guarantees at least one match (tests do) -- detect afterwards
    F.ComputeStereoMatches();
    int n = 0;
    for (int i = 0; i < nL; ++i) {
        u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; desc_index[i] = F.vDescIndex[i];
        n += F.mvuRight[i] >= 0;
    }
    return n;
}

} // extern "C"
