/*
 * oracle/orb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the reference's ORB front end:
 *   ORBextractor   src/ORBextractor.cc:77-147,410-853,1034-1136
 *   ORBmatcher     src/ORBmatcher.cc:41-43,574-605,2279-2295
 * with the canonical octree tie-break of SURVEY.md section 8c (creation
 * sequence number instead of heap address).  Pinned against
 *   - oracle/_ref/liborbref_canon.so : the reference's own ORBextractor.cc compiled
 *     from /root/reference with only that tie-break edit (tests/test_oracle_vs_ref.py),
 *   - cv2 4.13.0 for the OpenCV primitives (tests/test_oracle_cvprim.py),
 *   - committed golden vectors (tests/golden/, scripts/make_golden.py).
 * The reference itself ships no tests or golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may use this.  The CUDA product path never does.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBO_MAX_LEVELS 32

/* cv::KeyPoint layout (28 bytes). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbo_keypoint;

/* FAST candidate, coordinates relative to (minBorderX, minBorderY) = (16,16). */
typedef struct { int32_t x, y, score; } orbo_cand;

typedef struct orbo_extractor orbo_extractor;

orbo_extractor *orbo_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orbo_destroy(orbo_extractor *e);

/* ORBextractor::operator().  Returns the keypoint count n (kps/desc may be NULL to
 * just run the stages), or -n if cap < n. */
int orbo_extract(orbo_extractor *e, const uint8_t *gray, int w, int h, int stride,
                 orbo_keypoint *kps, uint8_t *desc, int cap);

/* Constructor tables (src/ORBextractor.cc:410-470). umax16 gets 16 ints. */
int orbo_tables(const orbo_extractor *e, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                int *nfeat_per_level, int *umax16);

/* Stage outputs of the last orbo_extract call. */
int orbo_level_size(const orbo_extractor *e, int level, int *w, int *h);
const uint8_t *orbo_level_image(const orbo_extractor *e, int level);    /* un-padded, stride = w */
const uint8_t *orbo_level_blurred(const orbo_extractor *e, int level);  /* NULL if the level had no keypoints */
int orbo_level_candidates(const orbo_extractor *e, int level, const orbo_cand **out);
int orbo_level_min_cells(const orbo_extractor *e, int level, int *cells_visited); /* cells that fell back to minTh */
int orbo_level_nkeypoints(const orbo_extractor *e, int level);
/* copy the padded (w+38)x(h+38) level like mvImagePyramid's parent buffer */
int orbo_level_padded(const orbo_extractor *e, int level, uint8_t *dst, int dst_stride);

/* DistributeOctTree alone (canonical tie-break). cands relative coords; out gets selected
 * candidates in list order.  Returns count. */
int orbo_distribute(const orbo_cand *cands, int n, int minX, int maxX, int minY, int maxY, int N,
                    orbo_cand *out, int cap);

/* ORBmatcher core. */
int orbo_hamming256(const uint8_t *a, const uint8_t *b);           /* SWAR popcount, 8 x int32 */
/* All-pairs best / second-best scan in index order (src/ORBmatcher.cc:574-605).
 * idx[i] = best j or -1 when nB == 0; d1/d2 = best / second-best distance (256 if none);
 * accept[i] = d1 <= th && (float)d1 < ratio*(float)d2.  Returns the number accepted. */
int orbo_match(const uint8_t *descA, int nA, const uint8_t *descB, int nB, int th, float ratio,
               int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept);
/* Same, query rows split over `threads` pthreads (CPU baseline). */
/* mbCheckOrientation (src/ORBmatcher.cc:610-620, 641-660, 2233-2274): prunes accept[] in place, returns matches kept */
int orbo_rotation_bin(float angle_a, float angle_b);
int orbo_rotation_filter(int nA, const int32_t *idx, uint8_t *accept, const float *angleA, const float *angleB,
                         int32_t *hist, int32_t *top3);
/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1958-2102).
 * cam: fx fy cx cy mbf mb mnMinX mnMaxX mnMinY mnMaxY; Tc / Tl: mTcw of the current / last frame, row-major 4x4;
 * last frame point i: world_pos, descriptor of its map point, valid (map point present and not an outlier), nobs
 * (MapPoint::Observations()), octave and angle of its keypoint; current frame: undistorted positions, octaves, angles, mvuRight,
 * descriptors.  cur_match[nC]: per current feature the last-frame point whose map point it holds at the end (-1 none).
 * Returns nmatches. */
int orbo_search_by_projection(const float *cam, const float *Tc, const float *Tl,
                              int nL, const float *world_pos, const uint8_t *mp_desc, const uint8_t *valid, const int32_t *nobs,
                              const int32_t *last_octave, const float *last_angle,
                              int nC, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                              const uint8_t *cur_desc, const float *scale, int nlevels, float th, int mono, int check_orientation,
                              int32_t *cur_match);
/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:780-895).  cam as above (only the image bounds are used); frame 1 / 2:
 * mvKeysUn positions, octaves, angles, descriptors; prev_xy [n1][2] = vbPrevMatched (updated in place); m12 [n1] = vnMatches12. */
int orbo_search_for_initialization(const float *cam, int n1, const float *xy1, const int32_t *oct1, const float *ang1, const uint8_t *desc1,
                                   int n2, const float *xy2, const int32_t *oct2, const float *ang2, const uint8_t *desc2,
                                   float *prev_xy, int window, float nnratio, int check_orientation, int32_t *m12);
/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (src/ORBmatcher.cc:418-502) with
 * RadiusByViewingCos (:504-510).  proj [nP][3] = mTrackProjX, mTrackProjY, mTrackProjXR; valid = mbTrackInView && !isBad();
 * feat_obs[j] = Observations() of the map point feature j holds (< 0: no map point); feat_match[j] = index of the newly assigned point. */
int orbo_search_local_points(const float *cam, int nP, const float *proj, const float *view_cos, const int32_t *level, const uint8_t *mp_desc,
                             const uint8_t *valid, const int32_t *nobs, int nF, const float *xy, const int32_t *octave, const float *uright,
                             const uint8_t *desc, const int32_t *feat_obs, const float *scale, int nlevels, float th, float nnratio,
                             int32_t *feat_match);
/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:532-663).  Feature vectors flattened in map order (node j owns
 * feats[off[j] .. off[j+1])); kf_valid = map point present and not bad; f_match[j] = key-frame feature matched to feature j. */
int orbo_search_by_bow(int nK, const float *kf_angle, const uint8_t *kf_desc, const uint8_t *kf_valid, int nk_nodes, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_feats, int nF, const float *f_angle, const uint8_t *f_desc, int nf_nodes,
                       const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_feats, float nnratio, int check_orientation, int32_t *f_match);
/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:897-1030); m12[i] = feature of key frame 2 matched to feature i. */
int orbo_search_by_bow_kf(int n1, const float *ang1, const uint8_t *desc1, const uint8_t *valid1, int nn1, const int32_t *nodes1, const int32_t *off1,
                          const int32_t *feats1, int n2, const float *ang2, const uint8_t *desc2, const uint8_t *valid2, int nn2, const int32_t *nodes2,
                          const int32_t *off2, const int32_t *feats2, float nnratio, int check_orientation, int32_t *m12);
/* DBoW2 vocabulary tree as the reference vendors it (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h): built from the rows
 * of an ORBvoc text file (parent id, leaf flag, 32 descriptor bytes, weight per node, file order), descent per feature,
 * BowVector / FeatureVector assembly.  scoring: 0 L1, 1 L2, 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT;
 * weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY (BowVector.h). */
typedef struct orbo_voc orbo_voc;
orbo_voc *orbo_voc_create(int k, int L, int scoring, int weighting, int nfile, const int32_t *parent, const uint8_t *is_leaf,
                          const uint8_t *desc, const double *weight);
void orbo_voc_destroy(orbo_voc *v);
void orbo_voc_transform_each(const orbo_voc *v, const uint8_t *desc, int n, int levelsup, int32_t *word, int32_t *node, double *weight);
int orbo_voc_bow(const orbo_voc *v, int n, const int32_t *word, const int32_t *node, const double *weight,
                 int32_t *bow_ids, double *bow_vals, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv);
/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301) for one map point: index of the representative descriptor */
int orbo_distinctive_descriptor(const uint8_t *desc, int n, int32_t *best_median);
/* Frame::ComputeStereoMatches (src/Frame.cc:849-1038).  pyr*[l]: un-padded level images (pitch[l], lw[l] x lh[l]).
 * Outputs per left keypoint: mvuRight, mvDepth, vDescIndex; optional debug: best Hamming distance / index, SAD of the
 * pushed matches (-1 otherwise).  Returns the matches kept, or -1 when none was pushed (the reference then reads an
 * empty vector, :1024). */
int orbo_stereo_matches(int nL, const orbo_keypoint *kL, const uint8_t *dL, int nR, const orbo_keypoint *kR, const uint8_t *dR,
                        int nlevels, const float *scale, const float *inv_scale,
                        const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *pitch, const int *lw, const int *lh,
                        float bf, float *u_right, float *depth, int32_t *desc_index, int32_t *best_dist, int32_t *best_idx, int32_t *sad);
int orbo_match_mt(const uint8_t *descA, int nA, const uint8_t *descB, int nB, int th, float ratio,
                  int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept, int threads);

/* CPU baseline: extract `nframes` frames (contiguous, w*h each) with `threads` worker
 * threads, one extractor instance and one frame at a time per thread.  Returns wall
 * seconds; *total_kps = sum of keypoint counts. */
double orbo_extract_many(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                         const uint8_t *frames, int nframes, int w, int h, int threads, long *total_kps);

#ifdef __cplusplus
}
#endif
#endif
