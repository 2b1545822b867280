/*
 * include/orbx.h -- C ABI of the B200-native ORB front end (liborbx.so).
 *
 * This is the drop-in boundary for the reference's feature front end
 * (MultMotTracking / ORB-SLAM2): everything below replaces work the reference
 * does on the CPU inside
 *
 *   ORBextractor::operator()(image, mask, keypoints, descriptors)
 *                                   src/ORBextractor.cc:1046-1109, include/ORBextractor.h:56-61
 *   ORBmatcher::DescriptorDistance  src/ORBmatcher.cc:2279-2295, include/ORBmatcher.h:44
 *   the best / second-best / ratio loop shared by every SearchBy* matcher
 *                                   src/ORBmatcher.cc:574-605 (also :807-836, :941-975)
 *
 * Plain C: POD structs, raw pointers and sizes, integer status codes, no C++
 * exceptions, no OpenCV or torch types.  The C++ adapter that keeps the
 * reference's class surface (multimot_track_b200/adapter/ORBextractor.{h,cc}) and the
 * Python mirror (multimot_track_b200/extractor.py) are thin layers over these entry
 * points; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device
 * (sm_100a) and returns ORBX_ERR_CUDA when none is usable.
 *
 * Devices: every entry point runs on its handle's device and restores the calling
 * thread's current CUDA device before it returns.
 *
 * Threading: a handle owns one CUDA stream and all of its scratch memory.  One
 * handle must not be used from two threads at once; different handles may run
 * concurrently (the reference's stereo Frame constructor runs two extractor
 * instances on two threads, src/Frame.cc:96-99).  orbx_hamming256 is re-entrant.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK               0
#define ORBX_ERR_BAD_ARG     -1   /* NULL pointer, non-positive size, nlevels out of range ... */
#define ORBX_ERR_CAPACITY    -2   /* output buffer too small (see orbx_max_keypoints)          */
#define ORBX_ERR_CUDA        -3   /* CUDA runtime error or no usable device                     */
#define ORBX_ERR_OOM         -4   /* host or device allocation failed                           */
#define ORBX_ERR_UNSUPPORTED -5   /* shape the reference itself cannot process (see below)      */
#define ORBX_ERR_STATE       -6   /* collect without a pending submit, stage read before extract */

#define ORBX_MAX_LEVELS 16
#define ORBX_TH_LOW   50          /* ORBmatcher::TH_LOW,  src/ORBmatcher.cc:42 */
#define ORBX_TH_HIGH 100          /* ORBmatcher::TH_HIGH, src/ORBmatcher.cc:41 */
#define ORBX_HISTO_LENGTH 30      /* ORBmatcher::HISTO_LENGTH, src/ORBmatcher.cc:43 */

typedef struct orbx_handle orbx_handle;

/* The five ORBextractor constructor arguments (src/ORBextractor.cc:410-414,
 * read from ORBextractor.* in the settings yaml by src/Tracking.cc:202-208)
 * plus sizing hints.  max_* may be 0: buffers then grow on first use. */
typedef struct {
    int32_t nfeatures;
    float   scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t max_width;
    int32_t max_height;
    int32_t max_batch;
    int32_t device_id;       /* CUDA device ordinal; -1 = current device */
} orbx_config;

/* Binary layout of cv::KeyPoint (28 bytes), so a std::vector<cv::KeyPoint> can be
 * filled by memcpy.  Fields as the reference sets them (src/ORBextractor.cc:837-847,
 * :1098-1104): pt in level-0 pixels, size = (int)(31*scale[octave]), angle in
 * degrees [0,360) from fastAtan2, response = FAST score, class_id = -1. */
typedef struct {
    float   x, y;
    float   size;
    float   angle;
    float   response;
    int32_t octave;
    int32_t class_id;
} orbx_keypoint;

/* The same keypoint without its redundant fields, 12 bytes (ORBX_OPT_COMPACT_KEYPOINTS): integer pixel position IN ITS OWN
 * PYRAMID LEVEL, octave, FAST response (an integer <= 255) and the angle.  Lossless: orbx_expand_keypoints rebuilds the
 * cv::KeyPoint exactly (pt = position * mvScaleFactor[octave] in float, size from the octave, class_id = -1).  A batch's
 * keypoints shrink from 28 to 12 bytes on their way to the host (results are 44 instead of 60 bytes per keypoint). */
typedef struct {
    uint16_t x, y;
    uint8_t  octave, response;
    uint16_t reserved;
    float    angle;
} orbx_keypoint_compact;

/* ---- lifetime ---------------------------------------------------------- */

/* Replaces `new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST)`
 * (src/Tracking.cc:208).  Builds the constructor tables with the reference's exact
 * float/double arithmetic (src/ORBextractor.cc:415-469). */
int orbx_create(const orbx_config *cfg, orbx_handle **out);
void orbx_destroy(orbx_handle *h);

/* Message for the last failing call on this handle (h may be NULL: last
 * orbx_create failure of the calling thread).  Never NULL. */
const char *orbx_last_error(const orbx_handle *h);

/* ---- constructor tables (the getters Frame reads, src/Frame.cc:87-93) ---- */

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:63-85) and mnFeaturesPerLevel.
 * Any pointer may be NULL.  Returns nlevels (>0) or a negative status. */
int orbx_get_tables(const orbx_handle *h, float *scale, float *inv_scale, float *sigma2,
                    float *inv_sigma2, int32_t *nfeatures_per_level);

/* Upper bound on keypoints per frame for a width x height image (each level may
 * exceed its quota by at most 3: src/ORBextractor.cc:669,730).  Output buffers
 * passed to the extract calls must hold at least this many rows per frame. */
int orbx_max_keypoints(orbx_handle *h, int width, int height);

/* Host-only planning (no CUDA device needed): the constructor tables and the
 * per-level geometry the front end will use for a width x height image -- level
 * sizes (:1116), FAST cell grid (:771-806), feature quota per level (:435-447),
 * initial octree nodes (:543), and the keypoint capacity per frame.  Lets a caller
 * size its buffers, and lets the host arithmetic be checked without a GPU. */
typedef struct {
    int32_t nlevels;
    int32_t max_keypoints;                       /* == orbx_max_keypoints(width, height) */
    int32_t level_width[ORBX_MAX_LEVELS], level_height[ORBX_MAX_LEVELS];
    int32_t nfeatures_per_level[ORBX_MAX_LEVELS];
    int32_t cell_cols[ORBX_MAX_LEVELS], cell_rows[ORBX_MAX_LEVELS];   /* cells actually visited */
    int32_t cell_w[ORBX_MAX_LEVELS], cell_h[ORBX_MAX_LEVELS];         /* wCell, hCell */
    int32_t octree_roots[ORBX_MAX_LEVELS];       /* nIni */
    int32_t max_candidates[ORBX_MAX_LEVELS];     /* worst-case FAST candidates of the level */
    float   scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS];
    float   sigma2[ORBX_MAX_LEVELS], inv_sigma2[ORBX_MAX_LEVELS];
    float   keypoint_size[ORBX_MAX_LEVELS];      /* cv::KeyPoint::size per octave */
    int32_t umax[16];                            /* circular patch row ends (:453-469) */
} orbx_plan;
int orbx_make_plan(const orbx_config *cfg, int width, int height, orbx_plan *out);

/* ---- ORBextractor::operator() ------------------------------------------ */

/* One frame, host buffers, synchronous: the body of ORBextractor::operator().
 * gray: 8-bit single channel (the reference asserts CV_8UC1, :1053), stride_bytes
 * between rows.  kps[cap], desc[cap*32].  On return *n_out keypoints/descriptor rows
 * are valid, in the reference's order (level-major, octree list order).  An empty
 * image (NULL / zero size) returns ORBX_OK with *n_out = 0 and leaves the outputs
 * untouched, mirroring the silent return at :1049. */
int orbx_extract(orbx_handle *h, const uint8_t *gray, int width, int height, int stride_bytes,
                 orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out);

/* The colour image as the drivers hand it to Tracking::GrabImageRGBD / Stereo / Monocular: 3 or 4 interleaved 8-bit
 * channels, converted to grey ON THE DEVICE exactly like cvtColor(CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY |
 * CV_BGRA2GRAY) at src/Tracking.cc:459-472 (rgb_order = Camera.RGB of the settings file, :192-193), then extracted.
 * channels == 1 behaves like orbx_extract.  The grey image is level 0 of the pyramid (orbx_get_pyramid_level). */
int orbx_extract_color(orbx_handle *h, const uint8_t *image, int width, int height, int stride_bytes, int channels,
                       int rgb_order, orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out);

/* A batch of equally sized frames given as host pointers.  Outputs are laid out
 * frame-major: frame f writes kps[f*cap_per_frame ...], desc[f*cap_per_frame*32 ...],
 * n_out[f].  Pinned (page-locked) host buffers are used directly by the copy engine;
 * pageable ones go through the handle's pinned staging area. */
int orbx_extract_batch(orbx_handle *h, const uint8_t *const *frames, int nframes, int width, int height,
                       int stride_bytes, orbx_keypoint *kps, uint8_t *desc, int cap_per_frame, int *n_out);

/* Same, with the frames already resident in device memory (frame f starts at
 * d_frames + f*frame_stride_bytes).  Split into an asynchronous submit on the
 * handle's stream and a collect that waits for it, so a caller can keep several
 * handles in flight (results of batch i copy back while batch i+1 computes).
 * Lifetime of d_frames: when the pointer, stride_bytes and frame_stride_bytes are all multiples of 16 the frames are
 * read IN PLACE as pyramid level 0 -- by this submit and by later orbx_get_pyramid_level(level 0) / orbx_stereo_match
 * calls on the handle -- so they must stay unchanged until the next submit on this handle.  Otherwise (or with
 * ORBX_OPT_COPY_INPUT set) they are copied into the handle's own level-0 slots and may be reused after the collect. */
int orbx_submit_device(orbx_handle *h, const uint8_t *d_frames, int nframes, int width, int height,
                       int stride_bytes, size_t frame_stride_bytes);
int orbx_submit_host(orbx_handle *h, const uint8_t *const *frames, int nframes, int width, int height,
                     int stride_bytes);
/* Waits for the pending submit, then fills the caller's arrays (frame-major as above). */
int orbx_collect(orbx_handle *h, orbx_keypoint *kps, uint8_t *desc, int cap_per_frame, int *n_out);

/* Zero-copy variant: waits for the pending submit and returns pointers into the
 * handle's pinned result buffers (frame f at kps + f*cap, desc + f*cap*32, n[f]);
 * they stay valid until the next submit on this handle.  Any pointer may be NULL. */
int orbx_collect_view(orbx_handle *h, const orbx_keypoint **kps, const uint8_t **desc, const int **n_out,
                      int *cap_per_frame);

/* With ORBX_OPT_COMPACT_KEYPOINTS set the device-to-host copy carries orbx_keypoint_compact records: orbx_collect still fills
 * cv::KeyPoint-layout arrays (expanded on the host), orbx_collect_view is replaced by this call (frame f at ckps + f*cap). */
int orbx_collect_view_compact(orbx_handle *h, const orbx_keypoint_compact **ckps, const uint8_t **desc, const int **n_out,
                              int *cap_per_frame);
/* Rebuilds n cv::KeyPoint records from compact ones with the handle's constructor tables (exact: one float multiply each). */
int orbx_expand_keypoints(const orbx_handle *h, const orbx_keypoint_compact *in, int n, orbx_keypoint *out);

/* ---- mvImagePyramid and stage read-back -------------------------------- */

/* Size of pyramid level `level` for the last processed shape. */
int orbx_get_level_size(const orbx_handle *h, int level, int *width, int *height);

/* Copies level `level` of frame `frame` of the last batch to host memory.
 * with_border = 0: the w x h image (what mvImagePyramid[level] views);
 * with_border = 1: the (w+38) x (h+38) parent buffer with the 19-pixel
 * BORDER_REFLECT_101 frame of src/ORBextractor.cc:1126-1132, which
 * Frame::ComputeStereoMatches may read around a keypoint (src/Frame.cc:960-977). */
int orbx_get_pyramid_level(orbx_handle *h, int frame, int level, uint8_t *dst, int dst_stride, int with_border);

/* Levels 0 .. nlevels-1 of frame `frame` at once, level l into dst[l] with row stride dst_stride[l] (same meaning of
 * with_border): one synchronisation for the whole pyramid -- what the adapter fills mvImagePyramid with. */
int orbx_get_pyramid_levels(orbx_handle *h, int frame, int nlevels, uint8_t *const *dst, const int *dst_stride, int with_border);

/* The 7x7 sigma=2 blurred level the descriptors were sampled from (:1089-1090). */
int orbx_get_blurred_level(orbx_handle *h, int frame, int level, uint8_t *dst, int dst_stride);

/* FAST candidates of one level before the octree (vToDistributeKeys, :820-825):
 * xys[3*i..] = x, y (relative to (16,16) like the reference) and score.  The device
 * emits them unordered; this call returns them sorted into the reference's order
 * (cell row, cell column, y, x).  *n gets the count (may exceed cap). */
int orbx_get_candidates(orbx_handle *h, int frame, int level, int32_t *xys, int cap, int *n);

/* ---- ORBmatcher core ---------------------------------------------------- */

/* ORBmatcher::DescriptorDistance on two 32-byte descriptors (host scalar: a
 * one-pair GPU launch is meaningless; this keeps the static member's signature). */
int orbx_hamming256(const void *a, const void *b);

/* All-pairs nearest / second-nearest scan, queries descA[nA][32] against
 * descB[nB][32] (host pointers), with the reference's update rule in index order:
 *   if d < best1 { best2 = best1; best1 = d; idx = j } else if d < best2 { best2 = d }
 * idx[i] = -1 and d1 = d2 = 256 when nB == 0.  accept[i] (may be NULL) =
 *   best1 <= th && (float)best1 < ratio * (float)best2          (:601-603).
 * Returns the number of accepted matches (>= 0) or a negative status. */
int orbx_match(orbx_handle *h, const uint8_t *descA, int nA, const uint8_t *descB, int nB,
               int th, float ratio, int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept);

/* Same with all six arrays in device memory, asynchronous on the handle's stream
 * (orbx_sync waits).  d_accept may be NULL.  Returns a status only. */
int orbx_match_device(orbx_handle *h, const uint8_t *d_descA, int nA, const uint8_t *d_descB, int nB,
                      int th, float ratio, int32_t *d_idx, int32_t *d_d1, int32_t *d_d2, uint8_t *d_accept);

/* Rotation-consistency check of the matchers (mbCheckOrientation): src/ORBmatcher.cc:545 and 610-620 (histogram of
 * angleA[i] - angleB[idx[i]] over the accepted matches, 30 bins, bin = round(rot * (1.0f / HISTO_LENGTH)) as the
 * reference writes it), ComputeThreeMaxima :2233-2274, pruning :641-660.  accept[nA] is updated in place (host
 * pointers; angleB has nB entries); hist[30] and top3[3] (ind1..ind3, -1 = none) may be NULL.
 * Returns the number of matches kept (>= 0) or a negative status. */
int orbx_rotation_filter(orbx_handle *h, int nA, const int32_t *idx, uint8_t *accept, const float *angleA,
                         const float *angleB, int nB, int32_t *hist, int32_t *top3);

/* Same on device arrays, asynchronous on the handle's stream; d_kept (may be NULL) receives the count. */
int orbx_rotation_filter_device(orbx_handle *h, int nA, const int32_t *d_idx, uint8_t *d_accept, const float *d_angleA,
                                const float *d_angleB, int32_t *d_hist, int32_t *d_top3, int32_t *d_kept);

/* ---- windowed projection matcher ---------------------------- */

/* Camera / pose block of the two frames: Frame::fx, fy, cx, cy, mbf, mb, the undistorted image bounds mnMinX .. mnMaxY
 * (src/Frame.cc:113-136) and mTcw of the current and of the last frame, row-major 4x4. */
typedef struct orbx_projection_setup {
    float fx, fy, cx, cy, bf, b;
    float min_x, max_x, min_y, max_y;
    float Tcw_cur[16], Tcw_last[16];
} orbx_projection_setup;

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (src/ORBmatcher.cc:1958-2102), the
 * matcher of Tracking::TrackWithMotionModel (src/Tracking.cc:2986-2992), with the 64x48 grid of the current frame
 * (Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid, src/Frame.cc:601-616, 710-776).
 * Last frame, per keypoint i < n_last: world_pos[i][3] and mp_desc[i][32] of its map point (MapPoint::GetWorldPos /
 * GetDescriptor), valid[i] (a map point is there and mvbOutlier[i] is false), nobs[i] (MapPoint::Observations()),
 * last_octave[i], last_angle[i] (mvKeys / mvKeysUn).  Current frame, per feature j < n_cur: cur_xy[j][2] (mvKeysUn),
 * cur_octave, cur_angle, cur_uright (mvuRight, <= 0 when absent), cur_desc[j][32].  The scale factors are the handle's.
 * cur_match[j] receives the index i of the last-frame point whose map point feature j holds when the function returns
 * (CurrentFrame.mvpMapPoints[j]), -1 for none; with check_orientation the rotation histogram pruning is applied.
 * The projections, window tests and Hamming distances run on the GPU; the greedy claim bookkeeping, sequential in the
 * reference (:2028-2030, :2053), is replayed on the host over the candidate lists.  Host pointers.
 * Returns nmatches (>= 0) or a negative status.  A window may hold any number of candidates (up to 512 are staged in shared
 * memory; a fuller window makes the call run a second time with a global staging area sized from the counts). */
int orbx_search_by_projection(orbx_handle *h, const orbx_projection_setup *setup,
                              int n_last, const float *world_pos, const uint8_t *mp_desc, const uint8_t *valid, const int32_t *nobs,
                              const int32_t *last_octave, const float *last_angle,
                              int n_cur, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                              const uint8_t *cur_desc, float th, int mono, int check_orientation, int32_t *cur_match);

/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:780-895),
 * the matcher of Tracking::MonocularInitialization (src/Tracking.cc:2622), over the 64x48 grid of F2 with the image bounds
 * mnMinX .. mnMaxY (src/Frame.cc:113-136).  Frame 1, per keypoint i < n1: octave1, angle1 (mvKeysUn), desc1[i][32]; frame 2,
 * per keypoint j < n2: xy2[j][2], octave2, angle2 (mvKeysUn), desc2[j][32]; prev_matched[i][2] = vbPrevMatched, read as the
 * window centres and updated in place for the matched keypoints as the reference does (:888-891); nnratio = mfNNratio,
 * check_orientation = mbCheckOrientation.  matches12[i] receives vnMatches12.  The window / level gates and the Hamming
 * distances run on the GPU; the order-dependent vMatchedDistance / vnMatches21 bookkeeping (:819, :838-846) is replayed on the
 * host over the candidate lists.  Host pointers.  Returns nmatches (>= 0) or a negative status. */
int orbx_search_for_initialization(orbx_handle *h, float min_x, float max_x, float min_y, float max_y,
                                   int n1, const int32_t *octave1, const float *angle1, const uint8_t *desc1,
                                   int n2, const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2,
                                   float *prev_matched, int window_size, float nnratio, int check_orientation, int32_t *matches12);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (src/ORBmatcher.cc:418-502) with
 * RadiusByViewingCos (:504-510), the matcher of Tracking::SearchLocalPoints (src/Tracking.cc:3464), over the frame's 64x48 grid.
 * Local map points, per point i < n_mp: proj[i][3] = mTrackProjX, mTrackProjY, mTrackProjXR and view_cos[i] = mTrackViewCos,
 * level[i] = mnTrackScaleLevel (as Frame::isInFrustum left them), mp_desc[i][32] = GetDescriptor(), valid[i] = mbTrackInView &&
 * !isBad(), nobs[i] = Observations().  Frame, per feature j < n_feat: feat_xy[j][2], feat_octave (mvKeysUn), feat_uright (mvuRight,
 * <= 0 when absent), feat_desc[j][32], feat_obs[j] = Observations() of the map point the feature already holds (< 0: none).
 * The scale factors are the handle's.  feat_match[j] receives the index i of the map point the call assigns to feature j
 * (F.mvpMapPoints[j] = vpMapPoints[i]), -1 where the feature is left as it was.  The window / level / right-coordinate gates and
 * the Hamming distances run on the GPU; the claim bookkeeping, sequential in the reference (:457-459, :498), is replayed on the
 * host.  Host pointers.  Returns nmatches (>= 0; a point with no observations can be overwritten later, so nmatches may exceed the
 * number of assigned features, as in the reference) or a negative status. */
int orbx_search_local_points(orbx_handle *h, float min_x, float max_x, float min_y, float max_y,
                             int n_mp, const float *proj, const float *view_cos, const int32_t *level, const uint8_t *mp_desc,
                             const uint8_t *valid, const int32_t *nobs,
                             int n_feat, const float *feat_xy, const int32_t *feat_octave, const float *feat_uright,
                             const uint8_t *feat_desc, const int32_t *feat_obs, float th, float nnratio, int32_t *feat_match);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches, ..) (src/ORBmatcher.cc:532-663), the matcher of
 * Tracking::TrackReferenceKeyFrame (src/Tracking.cc:2853) and Relocalization (:3651): features are compared only inside the
 * vocabulary nodes the two DBoW2::FeatureVectors share.  Key frame, per feature i < n_kf: kf_angle (mvKeysUn), kf_desc[i][32],
 * kf_valid[i] (GetMapPointMatches()[i] is set and not bad); frame, per feature j < n_f: f_angle (mvKeys), f_desc[j][32].  The
 * feature vectors are flattened in map order as orbx_voc_bow returns them: node k of *_nnodes has id *_nodes[k] (ascending) and
 * owns the feature indices *_feats[*_off[k] .. *_off[k+1]).  f_match[j] receives the key-frame feature whose map point feature j
 * is matched to (vpMapPointMatches[j] = vpMapPointsKF[f_match[j]]), -1 for none; with check_orientation the rotation histogram
 * pruning is applied.  All pair distances inside the shared nodes are computed on the GPU; the walk in which an already matched
 * frame feature is skipped (:582-583) is replayed on the host.  Host pointers.  Returns nmatches or a negative status. */
int orbx_search_by_bow(orbx_handle *h, int n_kf, const float *kf_angle, const uint8_t *kf_desc, const uint8_t *kf_valid,
                       int kf_nnodes, const int32_t *kf_nodes, const int32_t *kf_off, const int32_t *kf_feats,
                       int n_f, const float *f_angle, const uint8_t *f_desc,
                       int f_nnodes, const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_feats,
                       float nnratio, int check_orientation, int32_t *f_match);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:897-1030), the matcher of
 * LoopClosing::ComputeSim3 (src/LoopClosing.cc:267): as orbx_search_by_bow between two key frames.  valid1 / valid2: the feature
 * holds a map point that is not bad; a feature of key frame 2 is used at most once (:948); the acceptance is bestDist1 < TH_LOW
 * (strict, :973) and the ratio test.  matches12[i] receives the feature of key frame 2 whose map point feature i is matched to
 * (vpMatches12[i] = vpMapPoints2[matches12[i]]), -1 for none.  Host pointers.  Returns nmatches or a negative status. */
int orbx_search_by_bow_keyframes(orbx_handle *h, int n1, const float *angle1, const uint8_t *desc1, const uint8_t *valid1,
                                 int nnodes1, const int32_t *nodes1, const int32_t *off1, const int32_t *feats1,
                                 int n2, const float *angle2, const uint8_t *desc2, const uint8_t *valid2,
                                 int nnodes2, const int32_t *nodes2, const int32_t *off2, const int32_t *feats2,
                                 float nnratio, int check_orientation, int32_t *matches12);

/* ---- bag-of-words vocabulary -------------------------------- */

/* The DBoW2 vocabulary tree of the reference (ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB>,
 * include/ORBVocabulary.h; loaded by System at src/System.cc:66-67, used by Frame::ComputeBoW, src/Frame.cc:778-785).
 * The tree lives on the handle's device; the per-descriptor descent (TemplatedVocabulary.h:1205-1250) runs on the GPU,
 * the BowVector / FeatureVector bookkeeping (:1127-1195) on the host in DBoW2's double arithmetic. */
typedef struct orbx_vocabulary orbx_vocabulary;

/* ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1338-1418): header `k L scoring weighting`, then one node per
 * line `parent isLeaf d0 .. d31 weight`.  Returns ORBX_OK or a negative status (message in orbx_last_error(h)). */
int orbx_voc_load_text(orbx_handle *h, const char *path, orbx_vocabulary **out);
/* The same tree from arrays: node i+1 of the file order has parent[i], is_leaf[i], desc[i][32], weight[i]. */
int orbx_voc_create(orbx_handle *h, int k, int L, int scoring, int weighting, int nnodes, const int32_t *parent,
                    const uint8_t *is_leaf, const uint8_t *desc, const double *weight, orbx_vocabulary **out);
void orbx_voc_destroy(orbx_vocabulary *voc);
int orbx_voc_info(const orbx_vocabulary *voc, int *k, int *L, int *nodes, int *words);

/* Per descriptor (host pointers, n x 32 bytes): word id, node id at level L - levelsup (0 when that level is <= 0) and
 * the word's weight, as TemplatedVocabulary::transform(feature, id, weight, &nid, levelsup) returns them. */
int orbx_voc_transform(orbx_handle *h, const orbx_vocabulary *voc, const uint8_t *desc, int n, int levelsup,
                       int32_t *word, int32_t *node, double *weight);

/* BowVector and FeatureVector of transform(features, v, fv, levelsup) from the per-descriptor results, flattened in map
 * order: bow_ids / bow_vals (ascending word id, *n_bow entries, at most n), fv_nodes[j] owns the feature indices
 * fv_feats[fv_off[j] .. fv_off[j+1]) (ascending node id, *n_fv nodes; fv_off needs n + 1 entries).  Host only. */
int orbx_voc_bow(const orbx_vocabulary *voc, int n, const int32_t *word, const int32_t *node, const double *weight,
                 int32_t *bow_ids, double *bow_vals, int *n_bow, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv);

/* ---- representative descriptors of map points -------------- */

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-306), batched: map point p has the observed descriptors
 * desc[offsets[p] .. offsets[p+1]) (32 bytes each, gathered by the caller as :263-270 does); for each point the
 * all-pairs Hamming distances, the median of every row (sorted row[(int)(0.5*(n-1))], :291) and the FIRST row with the
 * least median (:293-297).  best_idx[p] is relative to the point's first row (-1 for a point without descriptors);
 * best_median (may be NULL) receives that median.  Host pointers; at most 1024 observations per point.
 * Returns ORBX_OK or a negative status. */
int orbx_distinctive_descriptors(orbx_handle *h, const uint8_t *desc, const int32_t *offsets, int npoints,
                                 int32_t *best_idx, int32_t *best_median);

/* ---- stereo association ------------------------------------- */

/* Frame::ComputeStereoMatches (src/Frame.cc:849-1038) on the results of two extractors that are still resident on the
 * device: frame `frame_left` of the last batch of `left` against frame `frame_right` of the last batch of `right`
 * (keypoints, descriptors and pyramids; both handles on the same device, same constructor parameters and image size).
 * Per left keypoint i (i < *n_left <= cap): u_right[i] = mvuRight (sub-pixel right column, -1 = no match),
 * depth[i] = mvDepth = bf / disparity (-1 = none), desc_index[i] = vDescIndex (best right keypoint by Hamming distance,
 * with the reference's quirk that index 0 is never recorded, :948).  Gates as in the reference of this fork:
 * row band 1.2 * scale, octave +-1, disparity in [0, 200), Hamming < (TH_HIGH+TH_LOW)/2 for the 11x11 SAD refinement,
 * final cut at 1.5 * 1.4 * median SAD.  Any output pointer may be NULL.
 * Returns the number of matches kept (>= 0; 0 also when no match reached the median step, where the reference reads
 * an empty vector) or a negative status. */
int orbx_stereo_match(orbx_handle *left, orbx_handle *right, int frame_left, int frame_right, float bf,
                      float *u_right, float *depth, int32_t *desc_index, int cap, int *n_left);

/* ---- per-stage device timing ---------------------------------------------- */

/* With profiling on, every submit brackets its stages with CUDA events on the
 * handle's stream; after the matching collect, orbx_get_stage_ms returns the device
 * time of each stage of that batch in milliseconds (stage i named orbx_stage_name(i):
 * "input", "pyramid", "blur", "fast", "octree", "orient_desc", "d2h").  Returns the
 * number of stages.  Off by default (the events cost a few microseconds per batch). */
#define ORBX_NUM_STAGES 7
int orbx_set_profiling(orbx_handle *h, int on);
int orbx_get_stage_ms(orbx_handle *h, float *ms, int cap);
const char *orbx_stage_name(int stage);

/* ---- frame-sharded multi-GPU dispatcher ------------------------------ */

/* Frames are independent inside ORBextractor::operator() -- the reference itself runs two extractor instances on two threads
 * for a stereo pair (src/Frame.cc:96-99) -- so a batch or sequence shards frame-parallel across the GPUs of a box with no
 * exchange between them.  A pool owns, per device, one worker thread and `depth` extractor handles: a submit only queues the
 * shards, the worker threads issue the copies and launches (the caller's thread never sits in the CUDA launch path), `depth`
 * submits overlap on every GPU, and the results land in pinned host memory per shard. */
typedef struct orbx_pool orbx_pool;

typedef struct {
    orbx_config    extractor;   /* constructor arguments and sizing hints of every handle; max_batch = frames per device per
                                   submit; device_id is ignored                                                            */
    int32_t        ndevices;    /* entries of devices[]; 0 = one worker on the calling thread's current device             */
    const int32_t *devices;     /* CUDA ordinals, one worker each (an ordinal may repeat: several workers on one GPU)       */
    int32_t        depth;       /* submits in flight per device = handles per worker; 0 = 6                                 */
    int32_t        compact_keypoints;  /* ORBX_OPT_COMPACT_KEYPOINTS on every handle: shards carry ckps instead of kps       */
} orbx_pool_config;

/* One shard of a collected submit: nframes frames starting at frame first_frame of the submit; frame f of the shard has
 * n[f] keypoints at kps + f*cap_per_frame and descriptors at desc + f*cap_per_frame*32 (pinned host memory of the pool). */
typedef struct {
    const orbx_keypoint *kps;                 /* NULL when the pool delivers compact records */
    const uint8_t       *desc;
    const int32_t       *n;
    int32_t nframes, first_frame, cap_per_frame, device;
    const orbx_keypoint_compact *ckps;        /* frame f at ckps + f*cap_per_frame; NULL unless compact_keypoints */
} orbx_shard_result;

int orbx_pool_create(const orbx_pool_config *cfg, orbx_pool **out);
void orbx_pool_destroy(orbx_pool *p);
const char *orbx_pool_last_error(const orbx_pool *p);      /* p may be NULL: last orbx_pool_create failure of the thread */
int orbx_pool_devices(const orbx_pool *p);                  /* number of shards per submit                                 */
int orbx_pool_depth(const orbx_pool *p);
/* ORBX_OPT_* on every handle of the pool; only while no ticket is outstanding (ORBX_ERR_STATE otherwise). */
int orbx_pool_set_option(orbx_pool *p, int option, int value);
/* Handle `slot` (0 .. depth-1) of shard `shard`, e.g. for orbx_get_tables or stage read-back between submits. */
orbx_handle *orbx_pool_handle(orbx_pool *p, int shard, int slot);
/* The split every submit uses: shard g of G owns the contiguous frames [g*F/G, (g+1)*F/G). */
void orbx_pool_shard_range(int nframes, int nshards, int shard, int *first, int *count);

/* Host frames (as orbx_extract_batch takes them), split into contiguous blocks over the devices.  Returns a ticket >= 0
 * without waiting for the GPUs, or a negative status (ORBX_ERR_STATE when `depth` tickets are uncollected).  The frame
 * buffers must stay valid until the ticket is collected. */
long long orbx_pool_submit_host(orbx_pool *p, const uint8_t *const *frames, int nframes, int width, int height, int stride_bytes);
/* Device-resident shards (per-GPU replicas of a frame pool): shard g = nframes[g] frames at d_frames[g] on device g's GPU,
 * laid out as orbx_submit_device takes them (same lifetime rule). */
long long orbx_pool_submit_device(orbx_pool *p, const uint8_t *const *d_frames, const int32_t *nframes, int width, int height,
                                  int stride_bytes, size_t frame_stride_bytes);
/* Waits for every shard of the ticket and fills shards[orbx_pool_devices()] (may be NULL).  The views stay valid until the
 * submit that returns ticket + depth.  Tickets may be collected in any order. */
int orbx_pool_collect(orbx_pool *p, long long ticket, orbx_shard_result *shards);

/* ---- options --------------------------------------------------------------- */

/* Runtime options of a handle; set between batches (ORBX_ERR_STATE while a submit is pending). */
#define ORBX_OPT_TMA_STAGING 1   /* 1 (default): tiles / patches arrive by cp.async.bulk.tensor; 0: plain staging loads (same results) */
#define ORBX_OPT_FAST_TMA    2   /* 1: persistent TMA variant of the FAST kernel; default 0 (slower in the pipelined loop)           */
#define ORBX_OPT_COPY_INPUT  3   /* 1: orbx_submit_device always copies the frames into the handle's level-0 slots; default 0        */
#define ORBX_OPT_COMPACT_KEYPOINTS 4 /* 1: keypoints travel to the host as 12-byte orbx_keypoint_compact records; default 0            */
int orbx_set_option(orbx_handle *h, int option, int value);

/* ---- misc ---------------------------------------------------------------- */

int orbx_sync(orbx_handle *h);                 /* waits for everything queued on the handle's stream */
void *orbx_stream(orbx_handle *h);             /* the handle's cudaStream_t (for event timing by the caller) */
/* Number of kernel launches issued by this handle since creation (bench bookkeeping). */
long long orbx_launch_count(const orbx_handle *h);
const char *orbx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
