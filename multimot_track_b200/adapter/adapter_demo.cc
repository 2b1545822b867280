// multimot_track_b200/adapter/adapter_demo.cc -- runs the drop-in C++ ORBextractor adapter the way
// Frame::ExtractORB does and dumps the result, so a test can compare it with the oracle.
//   adapter_demo <gray.raw> <width> <height> <nfeatures> <scale> <nlevels> <iniTh> <minTh> <out.bin>
// out.bin: int32 n, n x 28-byte cv::KeyPoint, n x 32 descriptor bytes, then per level int32 w, h and the
// (w+38)x(h+38) mvImagePyramid parent buffer.  Built against the oracle/minicv shim (no OpenCV here).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher_core.h"

int main(int argc, char **argv)
{
    if (argc != 10) { std::fprintf(stderr, "usage: %s gray.raw w h nfeatures scale nlevels ini min out.bin\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    std::vector<unsigned char> buf((size_t)w * h);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(buf.data(), 1, buf.size(), f) != buf.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::fclose(f);
    cv::Mat im(h, w, CV_8UC1, buf.data());
    ORB_SLAM2::ORBextractor *extractor = new ORB_SLAM2::ORBextractor(std::atoi(argv[4]), (float)std::atof(argv[5]), std::atoi(argv[6]),
                                                                    std::atoi(argv[7]), std::atoi(argv[8]));
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    (*extractor)(im, cv::Mat(), mvKeys, mDescriptors);                 // src/Frame.cc:621
    FILE *o = std::fopen(argv[9], "wb");
    const int n = (int)mvKeys.size();
    std::fwrite(&n, 4, 1, o);
    std::fwrite(mvKeys.data(), sizeof(cv::KeyPoint), n, o);
    for (int i = 0; i < n; ++i) std::fwrite(mDescriptors.ptr(i), 1, 32, o);
    for (int l = 0; l < extractor->GetLevels(); ++l) {
        const cv::Mat &m = extractor->mvImagePyramid[l];
        const int lw = m.cols, lh = m.rows;
        std::fwrite(&lw, 4, 1, o); std::fwrite(&lh, 4, 1, o);
        for (int y = -19; y < lh + 19; ++y) std::fwrite(m.data + (long)y * (long)m.step - 19, 1, lw + 38, o);
    }
    std::fclose(o);
    // stereo constructor shape (src/Frame.cc:96-104): a second extractor on the same image, then ComputeStereoMatches
    ORB_SLAM2::ORBextractor *right = new ORB_SLAM2::ORBextractor(std::atoi(argv[4]), (float)std::atof(argv[5]), std::atoi(argv[6]),
                                                                std::atoi(argv[7]), std::atoi(argv[8]));
    std::vector<cv::KeyPoint> mvKeysRight;
    cv::Mat mDescriptorsRight;
    (*right)(im, cv::Mat(), mvKeysRight, mDescriptorsRight);
    std::vector<float> mvuRight, mvDepth;
    std::vector<int> vDescIndex;
    const int nstereo = ORB_SLAM2::ComputeStereoMatchesGPU(extractor->Handle(), right->Handle(), 386.1448f, mvuRight, mvDepth, vDescIndex);
    std::vector<int> bi, bd, sd;
    std::vector<unsigned char> acc;
    const int nm = ORB_SLAM2::BruteForceMatch(extractor->Handle(), mDescriptors, mDescriptorsRight, ORB_SLAM2::ORBmatcher::TH_HIGH, 0.9f, bi, bd, sd, acc);
    std::vector<float> angL(n), angR(mvKeysRight.size());
    for (int i = 0; i < n; ++i) angL[i] = mvKeys[i].angle;
    for (size_t i = 0; i < mvKeysRight.size(); ++i) angR[i] = mvKeysRight[i].angle;
    const int nrot = ORB_SLAM2::RotationConsistencyFilter(extractor->Handle(), bi, acc, angL, angR);
    std::printf("stereo on identical images: %d of %d matched (%d sized outputs); %d brute-force matches, %d after the rotation check\n",
                nstereo, n, (int)mvuRight.size(), nm, nrot);
    int self = n > 1 ? ORB_SLAM2::ORBmatcher::DescriptorDistance(mDescriptors.row(0), mDescriptors.row(0)) : 0;
    std::printf("%d keypoints, self distance %d, TH_LOW %d TH_HIGH %d\n", n, self, ORB_SLAM2::ORBmatcher::TH_LOW, ORB_SLAM2::ORBmatcher::TH_HIGH);
    return 0;
}
