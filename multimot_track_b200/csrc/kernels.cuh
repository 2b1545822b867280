// multimot_track_b200/csrc/kernels.cuh -- device-side parameter block and launch API.
#ifndef ORBX_KERNELS_CUH
#define ORBX_KERNELS_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "orbx_internal.h"

namespace orbx {

// Lives in device memory (one per handle), rebuilt when the image shape changes.
struct DevParams {
    int nlevels, ini_th, min_th;
    int kp_frame_cap;
    long long pyr_frame_bytes;
    long long cand_frame_elems;
    LevelGeom lv[kMaxLevels];
    int umax[16];
    int xtab_off[kMaxLevels], ytab_off[kMaxLevels];
    // buffers (batch-major: frame f at base + f*stride)
    uint8_t *pyr;                 // [F][pyr_frame_bytes]   levels 0..L-1, un-padded, pitch = lv.pitch
    uint8_t *blur;                // [F][pyr_frame_bytes]   7x7 sigma=2 blurred levels
    uint32_t *cand;               // [F][cand_frame_elems]  packed x | y<<12 | score<<24 (relative to (16,16))
    uint32_t *cand_count;         // [F][nlevels]
    uint32_t *kp_stage;           // [F][kp_frame_cap]      octree winners, packed like cand, list order per level
    uint32_t *kp_count;           // [F][nlevels]
    unsigned long long *sort_scratch;   // [F][cand_frame_elems]  octree keys when a level does not fit shared memory
    orbx_keypoint *out_kps;       // [F][kp_frame_cap]
    uint32_t *out_ckps;           // [F][kp_frame_cap][3]   compact 12-byte records beside the full ones (ORBX_OPT_COMPACT_KEYPOINTS), else NULL
    uint8_t *out_desc;            // [F][kp_frame_cap][32]
    int *out_n;                   // [F]
    const ResizeTab *xtab, *ytab;
    const uint32_t *ffast_work; int n_ffast_work;      // k_fast_fused jobs: (level<<24 | cell_row<<12 | first cell col), 1-2 cells each
    const uint32_t *blur_work;  int n_blur_work;
    const int8_t *pattern;        // 512 x (x,y)
    const uint32_t *oct_lut;      // octree path-code / cell-order tables, LevelGeom::lut_off
};

// Level-0 source of the current batch (either the caller's device frames used in
// place, or the level-0 slot of DevParams::pyr).
struct Src0 {
    const uint8_t *ptr;
    long long frame_stride;
    int pitch;
};

struct LaunchStats { long long launches = 0; };

// TMA descriptors for the pyramid: src[l] describes the level l-1 images (bytes; x, y, frame) that level l is
// resized from, with a box large enough for the source footprint of one 128x32 destination tile.
struct TmaMaps {
    CUtensorMap src[kMaxLevels];
    int box_w[kMaxLevels], box_h[kMaxLevels];
    bool ok[kMaxLevels];
};
// Encodes one 3-D (x bytes, y rows, frame) uint8 tensor map; false when the driver entry point is missing or the
// shape breaks a TMA constraint (then the caller keeps the plain shared-memory staging path).
bool encode_image_map(CUtensorMap *out, const void *base, int pitch, int rows, long long frame_stride, int nframes, int box_w, int box_h);
void resize_box(const LevelGeom &src, const LevelGeom &dst, int *box_w, int *box_h);   // box for dst tiles of 128x32; box_w = 0 if it cannot fit

cudaError_t launch_pyramid(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls,
                           const TmaMaps *tma = nullptr);
cudaError_t launch_gray(const uint8_t *src, long long src_frame_stride, int src_pitch, int channels, int rgb_order, uint8_t *dst,
                        long long dst_frame_stride, int dst_pitch, int w, int h, int nframes, cudaStream_t st, LaunchStats *ls);
cudaError_t launch_repack(const uint8_t *src, long long src_frame_stride, int src_pitch, uint8_t *dst, long long dst_frame_stride,
                          int dst_pitch, int w, int h, int nframes, cudaStream_t st, LaunchStats *ls);
// maps->m[l]: the level-l images of this batch (level 0 = the caller's frames or our level-0 slots); bit l of tma_levels
// set = level l's blur tiles are staged by TMA (box kBlurBoxW x kBlurBoxH bytes).  Passed by value as a __grid_constant__.
struct BlurMaps { CUtensorMap m[kMaxLevels]; };
cudaError_t launch_blur(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls,
                        const BlurMaps *maps = nullptr, unsigned tma_levels = 0);
constexpr int kBlurBoxW = kBlurTileW + 32, kBlurBoxH = kBlurTileH + 6;   // 16 bytes of halo each side keep the box origin 16-byte aligned
// maps->m[l]: level-l images of the batch, box kFfBoxW bytes x (h_cell + 6) rows; nullptr = plain global staging
struct FastMaps { CUtensorMap m[kMaxLevels]; };
constexpr int kFfBoxW = 96;
cudaError_t launch_fast(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, int n_small_jobs, cudaStream_t st, LaunchStats *ls,
                        const FastMaps *maps = nullptr);
cudaError_t launch_octree(const DevParams *dP, const DevParams &hP, int nframes, int max_node_cap, int max_feat, cudaStream_t st, LaunchStats *ls);
// img[l] / blr[l]: level-l images of the batch and their blurred copies, boxes kOdIcBoxW x kOdIcBoxH and kOdBlBoxW x kOdBlBoxH
struct OdMaps { CUtensorMap img[kMaxLevels], blr[kMaxLevels]; };
constexpr int kOdIcBoxW = 48, kOdIcBoxH = 31, kOdBlBoxW = 64, kOdBlBoxH = 37;
cudaError_t launch_orient_desc(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls,
                               const OdMaps *maps = nullptr);     // maps == nullptr: plain shared-memory staging
cudaError_t launch_pad_level(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dst_pitch, cudaStream_t st, LaunchStats *ls);
cudaError_t launch_match(const uint32_t *dA, int nA, const uint32_t *dB, int nB, int th, float ratio,
                         int32_t *d_idx, int32_t *d_d1, int32_t *d_d2, uint8_t *d_accept, int *d_naccept,
                         int4 *d_partial, unsigned *d_arrive, int nchunks, cudaStream_t st, LaunchStats *ls);   // d_arrive: one zeroed counter per 128 query rows
constexpr int kMatchRowsPerBlock = 128;
cudaError_t launch_rotation_filter(int nA, const int32_t *d_idx, uint8_t *d_accept, const float *d_angleA, const float *d_angleB,
                                   int32_t *d_hist, int32_t *d_top3, int *d_kept, cudaStream_t st, LaunchStats *ls);
struct ProjSetup {                   // camera, current pose and window parameters of SearchByProjection, by value
    float fx, fy, cx, cy, bf, b, min_x, max_x, min_y, max_y, w_inv, h_inv;
    float Tc[16];
    float scale[kMaxLevels];
    float th;
    int forward, backward;
};
cudaError_t launch_project_candidates(const ProjSetup &S, int n_last, const float *d_world, const uint8_t *d_mp_desc, const uint8_t *d_valid,
                                      const int32_t *d_last_octave, int n_cur, const float *d_cur_xy, const int32_t *d_cur_octave,
                                      const float *d_cur_uright, const uint8_t *d_cur_desc, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count, int *d_offset,
                                      int *d_total, cudaStream_t st, LaunchStats *ls);
cudaError_t launch_window_candidates(const ProjSetup &S, int n1, const float *d_prev_xy, const int32_t *d_oct1, const uint8_t *d_desc1, int n2,
                                     const float *d_xy2, const int32_t *d_oct2, const uint8_t *d_desc2, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count,
                                     int *d_offset, int *d_total, cudaStream_t st, LaunchStats *ls);
cudaError_t launch_local_candidates(const ProjSetup &S, int n_mp, const float *d_proj, const float *d_view_cos, const int32_t *d_level,
                                    const uint8_t *d_mp_desc, const uint8_t *d_valid, int n_feat, const float *d_xy, const int32_t *d_octave,
                                    const float *d_uright, const uint8_t *d_desc, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count, int *d_offset,
                                    int *d_total, cudaStream_t st, LaunchStats *ls);
cudaError_t launch_bow_pair_distances(int n_entries, const int4 *d_entries, const uint8_t *d_kf_desc, const uint8_t *d_f_desc, const int32_t *d_f_feats,
                                      uint16_t *d_out, cudaStream_t st, LaunchStats *ls);
constexpr int kProjCap = 512;        // candidates one search window stages in shared memory; a fuller window makes the call rerun with a global staging area sized from the counts
cudaError_t launch_bow_descent(const uint8_t *d_feat, int n, const int32_t *d_child_off, const int32_t *d_child_ids, const uint8_t *d_node_desc,
                               const int32_t *d_word_id, int nid_level, int32_t *d_word, int32_t *d_node, int32_t *d_final, cudaStream_t st, LaunchStats *ls);
constexpr int kDistinctiveMaxObs = 1024;      // observations of one map point that fit the CTA's shared memory
cudaError_t launch_distinctive(const uint8_t *d_desc, const int32_t *d_offsets, int npoints, int max_obs, int32_t *d_best_idx,
                               int32_t *d_best_median, cudaStream_t st, LaunchStats *ls);
struct StereoScales { float scale[kMaxLevels], inv_scale[kMaxLevels]; };     // mvScaleFactors / mvInvScaleFactors
cudaError_t launch_stereo(const DevParams *dPL, Src0 s0L, int frameL, const DevParams *dPR, Src0 s0R, int frameR, const StereoScales &sc, float bf,
                          int capL, float *d_u_right, float *d_depth, int32_t *d_desc_index, int32_t *d_sad, int *d_kept, cudaStream_t st, LaunchStats *ls);
int match_chunks(int nA, int nB);       // number of B chunks launch_match will use (sizes d_partial: nchunks*nA)
size_t octree_smem_bytes(int max_node_cap, int max_feat, int *key_cap);

} // namespace orbx
#endif
