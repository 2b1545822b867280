// multimot_track_b200/csrc/vocabulary.cpp -- host side of the DBoW2 vocabulary the reference uses for Frame::ComputeBoW
// (src/Frame.cc:778-785 -> Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h): the ORBvoc text loader (loadFromTextFile,
// :1338-1418) and the BowVector / FeatureVector assembly of transform(features, v, fv, levelsup) (:1127-1195, with
// BowVector::addWeight / addIfNotExist / normalize, BowVector.cpp:32-86).  The descent itself runs on the GPU
// (k_bow_descent); what is here is the double-precision bookkeeping, kept in DBoW2's operation order and compiled
// without FMA contraction so the BowVector values are bit-identical.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "orbx_internal.h"

namespace orbx {

int voc_build(int k, int L, int scoring, int weighting, int nfile, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc,
              const double *weight, VocHost *v, std::string *err)
{
    if (k < 1 || k > 32 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3 || nfile < 1) {
        *err = "vocabulary header out of range (k in 1..32, L in 1..10, scoring in 0..5, weighting in 0..3, at least one node)";
        return ORBX_ERR_BAD_ARG;
    }
    const int n = nfile + 1;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = n; v->nwords = 0;
    v->child_off.assign((size_t)n + 1, 0); v->child_ids.assign((size_t)n, 0); v->word_id.assign((size_t)n, -1);
    v->desc.assign((size_t)n * 32, 0); v->weight.assign((size_t)n, 0.0);
    std::vector<int> cnt((size_t)n, 0);
    for (int i = 0; i < nfile; ++i) {
        if (parent[i] < 0 || parent[i] > i) { *err = "vocabulary: a node's parent must be an earlier node"; return ORBX_ERR_BAD_ARG; }
        cnt[parent[i]]++;
    }
    for (int i = 0; i < n; ++i) {
        if (cnt[i] > 32) { *err = "vocabulary: more than 32 children of one node"; return ORBX_ERR_UNSUPPORTED; }
        v->child_off[i + 1] = v->child_off[i] + cnt[i];
    }
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int i = 0; i < nfile; ++i) {
        const int nid = i + 1, pid = parent[i];
        v->child_ids[v->child_off[pid] + cnt[pid]++] = nid;                 // children in file order
        std::memcpy(&v->desc[(size_t)nid * 32], desc + (size_t)i * 32, 32);
        v->weight[nid] = weight[i];
        if (is_leaf[i]) v->word_id[nid] = v->nwords++;                      // word ids in the order of the leaf flags
    }
    // DBoW2 stops the descent at a node without children and then reads its word id: the flag and the shape must agree
    for (int nid = 1; nid < n; ++nid)
        if ((v->child_off[nid + 1] == v->child_off[nid]) != (v->word_id[nid] >= 0)) {
            *err = "vocabulary: leaf flag and tree shape disagree (DBoW2 would read an unset word id)"; return ORBX_ERR_UNSUPPORTED;
        }
    if (v->child_off[1] == 0) { *err = "vocabulary: the root has no children"; return ORBX_ERR_BAD_ARG; }
    return ORBX_OK;
}

int voc_load_text(const char *path, VocHost *v, std::string *err)
{
    std::ifstream f(path);
    if (!f) { *err = std::string("cannot open ") + path; return ORBX_ERR_BAD_ARG; }
    std::string line;
    std::getline(f, line);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    { std::stringstream ss(line); ss >> k >> L >> n1 >> n2; }
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {      // the reference's check, :1359
        *err = "Vocabulary loading failure: This is not a correct text file!"; return ORBX_ERR_BAD_ARG;
    }
    std::vector<int32_t> parent; std::vector<uint8_t> leaf, desc; std::vector<double> weight;
    while (std::getline(f, line)) {
        const char *p = line.c_str();
        char *e = nullptr;
        const long pid = std::strtol(p, &e, 10);
        if (e == p) continue;                                               // blank line (DBoW2's loop would misread it)
        p = e;
        const long lf = std::strtol(p, &e, 10); p = e;
        parent.push_back((int32_t)pid); leaf.push_back(lf > 0);
        for (int i = 0; i < 32; ++i) { const long b = std::strtol(p, &e, 10); p = e; desc.push_back((uint8_t)b); }
        weight.push_back(std::strtod(p, &e));
    }
    if (parent.empty()) { *err = "vocabulary file has no nodes"; return ORBX_ERR_BAD_ARG; }
    return voc_build(k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), v, err);
}

namespace {
struct KV { int32_t key, idx; };
bool kv_less(const KV &a, const KV &b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; }
}

int voc_bow(const VocHost &v, int n, const int32_t *word, const int32_t *node, const double *weight, int32_t *bow_ids, double *bow_vals,
            int *n_bow, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv)
{
    // ScoringObject.h:74-89: every scoring but DOT_PRODUCT normalises, L2 scoring with the L2 norm, the rest with L1
    const bool must = v.scoring != 5, l2 = v.scoring == 1, tf = v.weighting == 0 || v.weighting == 1;
    std::vector<KV> kw, kn;
    for (int i = 0; i < n; ++i)
        if (weight[i] > 0) { kw.push_back({word[i], i}); kn.push_back({node[i], i}); }      // "not stopped", :1153
    std::sort(kw.begin(), kw.end(), kv_less);                               // by word, then feature order: what addWeight sees
    std::sort(kn.begin(), kn.end(), kv_less);
    const int m = (int)kw.size();
    int nb = 0;
    for (int i = 0; i < m;) {
        int j = i + 1;
        double acc = weight[kw[i].idx];                                     // insert(id, v)
        for (; j < m && kw[j].key == kw[i].key; ++j)
            if (tf) acc += weight[kw[j].idx];                               // TF_IDF / TF: addWeight; IDF / BINARY: addIfNotExist
        bow_ids[nb] = kw[i].key; bow_vals[nb] = acc; ++nb;
        i = j;
    }
    if (tf && nb > 0 && !must) {
        const double nd = nb;
        for (int i = 0; i < nb; ++i) bow_vals[i] /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) for (int i = 0; i < nb; ++i) norm += std::fabs(bow_vals[i]);
        else { for (int i = 0; i < nb; ++i) norm += bow_vals[i] * bow_vals[i]; norm = std::sqrt(norm); }
        if (norm > 0.0) for (int i = 0; i < nb; ++i) bow_vals[i] /= norm;
    }
    int nf = 0, o = 0;
    for (int i = 0; i < m;) {
        int j = i;
        fv_nodes[nf] = kn[i].key; fv_off[nf] = o;
        for (; j < m && kn[j].key == kn[i].key; ++j) fv_feats[o++] = kn[j].idx;
        ++nf; i = j;
    }
    fv_off[nf] = o;
    *n_bow = nb; *n_fv = nf;
    return ORBX_OK;
}

} // namespace orbx
