// multimot_track_b200/csrc/host_match.cpp -- the sequential part of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th,
// bMono) (src/ORBmatcher.cc:1958-2102).  The GPU (k_project_candidates) projects every last-frame map point, applies the window /
// level / stereo gates of Frame::GetFeaturesInArea and of :2032-2038 and computes the Hamming distances; what is left is the
// reference's greedy walk in point order -- a feature already claimed by an observed map point is skipped (:2028-2030), the
// first least distance in GetFeaturesInArea's candidate order wins (:2045-2049), a later temporal point overwrites -- and the
// rotation histogram (:2058-2067, ComputeThreeMaxima :2233-2274, pruning :2075-2094).  It is O(candidates), a few microseconds.
#include <algorithm>
#include <climits>
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
    // the reference asserts bin in [0, HISTO_LENGTH) (:618 and the other matchers); angles outside [0, 360) or NaN go to the overflow
    // bin HISTO_LENGTH, which ComputeThreeMaxima never selects, so such a match is pruned instead of indexing out of bounds
    if (!(rot >= 0.0f && rot < 915.0f)) return ORBX_HISTO_LENGTH;
    int bin = (int)roundf(rot * factor);
    if (bin == ORBX_HISTO_LENGTH) bin = 0;
    return bin;
}

// ComputeThreeMaxima (src/ORBmatcher.cc:2233-2274) on bin counts
void three_maxima_host(const int *cnt, int &i1, int &i2, int &i3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    i1 = i2 = i3 = -1;
    for (int i = 0; i < ORBX_HISTO_LENGTH; ++i) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
        else if (s > max3) { max3 = s; i3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) i3 = -1;
}
}

// The sequential part of ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:780-895) over the candidate lists of
// k_window_candidates: a candidate already matched at a distance <= ours is skipped (:819), best / second best in
// GetFeaturesInArea's order (:822-831), TH_LOW and the ratio test (:834-836), a better match steals the feature from its
// previous owner (:838-842); rotHist keeps every push, stolen or not (:856), and the pruning clears what is still set (:872-885).
int resolve_initialization_matches(int n1, int n2, const unsigned long long *cand, const int *count, const int *offset, const float *ang1,
                                   const float *ang2, float nnratio, int check_orientation, int32_t *m12)
{
    std::vector<int> matched_dist((size_t)n2, INT_MAX), m21((size_t)n2, -1), hist_item, hist_bin;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    int nmatches = 0;
    std::vector<unsigned long long> row;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int c = count[i1];
        if (c <= 0) continue;
        row.assign(cand + offset[i1], cand + offset[i1] + c);
        std::sort(row.begin(), row.end());
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < c; ++k) {
            const int i2 = (int)((row[k] >> 16) & 0xffffu), dist = (int)(row[k] & 0xffffu);
            if (matched_dist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= ORBX_TH_LOW && (float)bestDist < (float)bestDist2 * nnratio) {
            if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; --nmatches; }
            m12[i1] = bestIdx2; m21[bestIdx2] = i1; matched_dist[bestIdx2] = bestDist;
            ++nmatches;
            if (check_orientation) { hist_item.push_back(i1); hist_bin.push_back(rotation_bin_host(ang1[i1], ang2[bestIdx2])); }
        }
    }
    if (check_orientation) {
        int cnt[ORBX_HISTO_LENGTH + 1] = {0}, i1, i2, i3;
        for (size_t k = 0; k < hist_bin.size(); ++k) cnt[hist_bin[k]]++;
        three_maxima_host(cnt, i1, i2, i3);
        for (size_t k = 0; k < hist_bin.size(); ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3 && m12[hist_item[k]] >= 0) { m12[hist_item[k]] = -1; --nmatches; }
    }
    return nmatches;
}

// The sequential part of ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:418-502) over the candidate lists of
// k_local_candidates: a feature that holds an observed map point -- from before the call or assigned earlier in it -- is skipped
// (:457-459), best / second best with their octaves (:476-489), TH_HIGH, and the ratio test only when both are on one level (:493-496).
int resolve_local_matches(int n_mp, int n_feat, const unsigned long long *cand, const int *count, const int *offset, const int32_t *nobs,
                          const int32_t *feat_octave, const int32_t *feat_obs, float nnratio, int32_t *feat_match)
{
    std::vector<int> held((size_t)n_feat);
    for (int j = 0; j < n_feat; ++j) { held[j] = feat_obs[j] > 0 ? feat_obs[j] : 0; feat_match[j] = -1; }
    int nmatches = 0;
    std::vector<unsigned long long> row;
    for (int i = 0; i < n_mp; ++i) {
        const int c = count[i];
        if (c <= 0) continue;
        row.assign(cand + offset[i], cand + offset[i] + c);
        std::sort(row.begin(), row.end());
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int k = 0; k < c; ++k) {
            const int j = (int)((row[k] >> 16) & 0xffffu), dist = (int)(row[k] & 0xffffu);
            if (held[j] > 0) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = feat_octave[j]; bestIdx = j; }
            else if (dist < bestDist2) { bestLevel2 = feat_octave[j]; bestDist2 = dist; }
        }
        if (bestDist <= ORBX_TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            feat_match[bestIdx] = i; held[bestIdx] = nobs[i] > 0 ? nobs[i] : 0;
            ++nmatches;
        }
    }
    return nmatches;
}

// The sequential part of ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:532-663) over the distances of
// k_bow_pair_distances: entries in the order of the merge walk (node ascending, key-frame features in node order); a frame feature
// matched earlier is skipped (:582-583), best / second best (:589-598), TH_LOW and ratio (:601-603), rotation histogram whose
// pruning clears and decrements unconditionally (:641-654).  entries[e] = {key-frame feature, first slot in f_feats, count, first distance}.
int resolve_bow_matches(int n_entries, const int *entries, const uint16_t *dist, const int32_t *f_feats, int n_f, const float *kf_angle,
                        const float *f_angle, float nnratio, int check_orientation, int32_t *f_match)
{
    for (int j = 0; j < n_f; ++j) f_match[j] = -1;
    std::vector<int> hist_item, hist_bin;
    int nmatches = 0;
    for (int e = 0; e < n_entries; ++e) {
        const int ik = entries[4 * e], lo = entries[4 * e + 1], cnt = entries[4 * e + 2], at = entries[4 * e + 3];
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int t = 0; t < cnt; ++t) {
            const int jf = f_feats[lo + t];
            if (f_match[jf] >= 0) continue;
            const int d = dist[at + t];
            if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = jf; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist1 <= ORBX_TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
            f_match[bestIdxF] = ik;
            if (check_orientation) { hist_item.push_back(bestIdxF); hist_bin.push_back(rotation_bin_host(kf_angle[ik], f_angle[bestIdxF])); }
            ++nmatches;
        }
    }
    if (check_orientation) {
        int cnt[ORBX_HISTO_LENGTH + 1] = {0}, i1, i2, i3;
        for (size_t k = 0; k < hist_bin.size(); ++k) cnt[hist_bin[k]]++;
        three_maxima_host(cnt, i1, i2, i3);
        for (size_t k = 0; k < hist_bin.size(); ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { f_match[hist_item[k]] = -1; --nmatches; }
    }
    return nmatches;
}

// The same walk for ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:897-1030): side 2 must hold a good
// map point and be unmatched (:948-954), the threshold is strict (:973), the result is indexed by the features of key frame 1.
int resolve_bow_matches_kf(int n_entries, const int *entries, const uint16_t *dist, const int32_t *feats2, int n1, int n2, const uint8_t *valid2,
                           const float *ang1, const float *ang2, float nnratio, int check_orientation, int32_t *m12)
{
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    std::vector<uint8_t> matched2((size_t)n2, 0);
    std::vector<int> hist_item, hist_bin;
    int nmatches = 0;
    for (int e = 0; e < n_entries; ++e) {
        const int i1 = entries[4 * e], lo = entries[4 * e + 1], cnt = entries[4 * e + 2], at = entries[4 * e + 3];
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int t = 0; t < cnt; ++t) {
            const int i2 = feats2[lo + t];
            if (matched2[i2] || !valid2[i2]) continue;
            const int d = dist[at + t];
            if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = i2; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist1 < ORBX_TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
            m12[i1] = bestIdx2; matched2[bestIdx2] = 1;
            if (check_orientation) { hist_item.push_back(i1); hist_bin.push_back(rotation_bin_host(ang1[i1], ang2[bestIdx2])); }
            ++nmatches;
        }
    }
    if (check_orientation) {
        int cnt[ORBX_HISTO_LENGTH + 1] = {0}, i1, i2, i3;
        for (size_t k = 0; k < hist_bin.size(); ++k) cnt[hist_bin[k]]++;
        three_maxima_host(cnt, i1, i2, i3);
        for (size_t k = 0; k < hist_bin.size(); ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { m12[hist_item[k]] = -1; --nmatches; }
    }
    return nmatches;
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
        int cnt[ORBX_HISTO_LENGTH + 1] = {0};
        for (size_t k = 0; k < hist_bin.size(); ++k) cnt[hist_bin[k]]++;
        int i1, i2, i3;
        three_maxima_host(cnt, i1, i2, i3);
        for (size_t k = 0; k < hist_bin.size(); ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { cur_match[hist_item[k]] = -1; --nmatches; }
    }
    return nmatches;
}

} // namespace orbx
