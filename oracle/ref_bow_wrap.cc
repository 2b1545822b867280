// oracle/ref_bow_wrap.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's vendored DBoW2 (Thirdparty/DBoW2/DBoW2/*.cpp, TemplatedVocabulary.h, compiled
// UNMODIFIED against oracle/minicv + oracle/minicv/dbow_shim.hpp): ORBVocabulary::loadFromTextFile and the two transform()
// overloads that Frame::ComputeBoW uses (src/Frame.cc:778-785 -> TemplatedVocabulary.h:1127-1250).
#include <cstdint>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h
struct Voc : public ORBVocabulary {
    // the per-feature overload is protected in DBoW2 (TemplatedVocabulary.h:355)
    void one(const cv::Mat &f, DBoW2::WordId &id, DBoW2::WordValue &w, DBoW2::NodeId *nid, int levelsup) const { transform(f, id, w, nid, levelsup); }
    int nodes() const { return (int)m_nodes.size(); }
    int words() const { return (int)m_words.size(); }
    int k() const { return m_k; }
    int L() const { return m_L; }
};
std::vector<cv::Mat> rows(const uint8_t *desc, int n)
{
    std::vector<cv::Mat> v;
    for (int i = 0; i < n; ++i) v.push_back(cv::Mat(1, 32, CV_8UC1, (void *)(desc + 32 * (size_t)i), 32));
    return v;
}
}

extern "C" {

void *orbref_voc_load(const char *path)
{
    Voc *v = new Voc();
    if (!v->loadFromTextFile(path) || v->empty()) { delete v; return nullptr; }
    return v;
}
void orbref_voc_free(void *v) { delete (Voc *)v; }
void orbref_voc_info(void *v, int *k, int *L, int *nodes, int *words)
{
    const Voc *p = (const Voc *)v;
    *k = p->k(); *L = p->L(); *nodes = p->nodes(); *words = p->words();
}

// per feature: word id, node id at level L - levelsup, weight (TemplatedVocabulary.h:1205-1250)
void orbref_voc_transform_each(void *v, const uint8_t *desc, int n, int levelsup, int32_t *word, int32_t *node, double *weight)
{
    const Voc *p = (const Voc *)v;
    std::vector<cv::Mat> f = rows(desc, n);
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid = 0;
        p->one(f[i], id, w, &nid, levelsup);
        word[i] = (int32_t)id; node[i] = (int32_t)nid; weight[i] = w;
    }
}

// the BowVector / FeatureVector pair of Frame::ComputeBoW (TemplatedVocabulary.h:1127-1195), flattened in map order.
// Returns the number of words in the BowVector; *n_fv = nodes in the FeatureVector; fv_off has *n_fv + 1 entries.
int orbref_voc_transform(void *v, const uint8_t *desc, int n, int levelsup, int32_t *bow_ids, double *bow_vals,
                         int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv)
{
    const Voc *p = (const Voc *)v;
    std::vector<cv::Mat> f = rows(desc, n);
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    p->transform(f, bv, fv, levelsup);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i) { bow_ids[i] = (int32_t)it->first; bow_vals[i] = it->second; }
    int j = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++j) {
        fv_nodes[j] = (int32_t)it->first; fv_off[j] = o;
        for (size_t q = 0; q < it->second.size(); ++q) fv_feats[o++] = (int32_t)it->second[q];
    }
    fv_off[j] = o; *n_fv = j;
    return i;
}

} // extern "C"
