// multimot_track_b200/csrc/host_match.cpp -- the sequential part of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th,
// bMono) (src/ORBmatcher.cc:1958-2102).  The GPU (k_project_candidates) projects every last-frame map point, applies the window /
// level / stereo gates of Frame::GetFeaturesInArea and of :2032-2038 and computes the Hamming distances; what is left is the
// reference's greedy walk in point order -- a feature already claimed by an observed map point is skipped (:2028-2030), the
// first least distance in GetFeaturesInArea's candidate order wins (:2045-2049), a later temporal point overwrites -- and the
// rotation histogram (:2058-2067, ComputeThreeMaxima :2233-2274, pruning :2075-2094).  It is O(candidates), a few microseconds.
#include <algorithm>
#include <cmath>
#include <vector>

#include "orbx_internal.h"

namespace orbx {

namespace {
int rotation_bin_host(float a, float b)                              // :2060-2065, factor = 1.0f / HISTO_LENGTH
{
    const float factor = 1.0f / ORBX_HISTO_LENGTH;
    float rot = a - b;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == ORBX_HISTO_LENGTH) bin = 0;
    return bin;
}
}

// cand: compact lists of packed candidates (cell << 32 | feature index << 16 | distance); point i owns cand[offset[i] .. + count[i])
int resolve_projection_matches(int n_last, int n_cur, const unsigned long long *cand, const int *count, const int *offset, const int32_t *nobs,
                               const float *last_angle, const float *cur_angle, int check_orientation, int32_t *cur_match)
{
    std::vector<int> claim_obs((size_t)n_cur, 0), hist_item, hist_bin;
    for (int i = 0; i < n_cur; ++i) cur_match[i] = -1;
    int nmatches = 0;
    std::vector<unsigned long long> row;
    for (int i = 0; i < n_last; ++i) {
        const int c = count[i];
        if (c <= 0) continue;
        row.assign(cand + offset[i], cand + offset[i] + c);
        std::sort(row.begin(), row.end());                                   // GetFeaturesInArea order: cell column-major, then index
        int bestDist = 256, bestIdx2 = -1;
        for (int k = 0; k < c; ++k) {
            const int i2 = (int)((row[k] >> 16) & 0xffffu), dist = (int)(row[k] & 0xffffu);
            if (cur_match[i2] >= 0 && claim_obs[i2] > 0) continue;
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= ORBX_TH_HIGH) {
            cur_match[bestIdx2] = i; claim_obs[bestIdx2] = nobs[i];
            ++nmatches;
            if (check_orientation) { hist_item.push_back(bestIdx2); hist_bin.push_back(rotation_bin_host(last_angle[i], cur_angle[bestIdx2])); }
        }
    }
    if (check_orientation) {
        int cnt[ORBX_HISTO_LENGTH] = {0};
        for (size_t k = 0; k < hist_bin.size(); ++k) cnt[hist_bin[k]]++;
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < ORBX_HISTO_LENGTH; ++i) {
            const int s = cnt[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) i3 = -1;
        for (size_t k = 0; k < hist_bin.size(); ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { cur_match[hist_item[k]] = -1; --nmatches; }
    }
    return nmatches;
}

} // namespace orbx
