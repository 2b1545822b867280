// multimot_track_b200/csrc/orbx_internal.h -- shared between the host-side geometry
// builder (host_tables.cpp), the kernels (kernels.cu) and the C-ABI (orbx_api.cpp).
#ifndef ORBX_INTERNAL_H
#define ORBX_INTERNAL_H

#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/orbx.h"

namespace orbx {

constexpr int kMaxLevels = ORBX_MAX_LEVELS;
constexpr int kMaxRoots = 16;            // initial octree nodes (round(width/height)), :543
constexpr int kEdge = 19;                // EDGE_THRESHOLD, src/ORBextractor.cc:74
constexpr int kMinBorder = 16;           // EDGE_THRESHOLD-3, :772
constexpr int kMaxDim = 4095 + 2 * kMinBorder;   // candidate coordinates are packed in 12 bits

// Per-level geometry.  Everything that needs the reference's float arithmetic is
// evaluated once on the host (host_tables.cpp) so the device code is integer-only
// wherever the reference is.
struct LevelGeom {
    int w, h, pitch;                 // un-padded level image; pitch in bytes (multiple of 64)
    long long img_off;               // byte offset of the level inside one frame's pyramid block
    // per-cell FAST grid, src/ORBextractor.cc:771-806
    int n_cols, n_rows;              // nCols, nRows as the reference computes them
    int w_cell, h_cell;              // wCell, hCell
    int cols_vis, rows_vis;          // cells actually visited (after the two `continue`s)
    int x_end, y_end;                // exclusive end of the detection area in level coordinates
    int cell_work_off;               // first entry of this level in the FAST work list
    // DistributeOctTree, :539-763
    int n_feat;                      // mnFeaturesPerLevel[level]
    int n_ini;                       // initial nodes
    float h_x;                       // hX
    int root_ul[kMaxRoots], root_br[kMaxRoots];
    int region_h;                    // maxY - minY
    int depth;                       // splits until every node is one pixel
    int lut_off;                     // offset (uint32 units) of this level's octree tables in Geometry::oct_lut:
                                     //   [0,region_w) path bits of x (+ initial node), then [region_w, +region_h) path bits of y,
                                     //   then region_w cell columns, then region_h cell rows * n_cols
    int region_w;                    // maxX - minX
    // buffers
    int cand_cap;                    // worst-case FAST candidates of the level
    long long cand_off;              // element offset inside one frame's candidate block
    int kp_cap, kp_off;              // per-level slot in one frame's keypoint staging
    int node_cap;                    // octree node array capacity
    // output fields
    float scale;                     // mvScaleFactor[level]
    float kp_size;                   // (float)(int)(PATCH_SIZE*scale), :837
};

struct alignas(8) ResizeTab {        // cv::resize INTER_LINEAR coefficients, one per dst column / row (one 64-bit load)
    uint16_t s0, s1;                 // source indices (s1 clamped)
    int16_t c0, c1;                  // 11-bit fixed point weights
};

struct Geometry {
    int width = 0, height = 0, nlevels = 0;
    LevelGeom lv[kMaxLevels];
    long long pyr_frame_bytes = 0;   // bytes of one frame's pyramid block (levels 1..L-1; level 0 too when copied)
    long long cand_frame_elems = 0;  // candidate slots per frame
    int kp_frame_cap = 0;            // keypoint staging slots per frame == output capacity per frame
    int max_node_cap = 0, max_feat = 0, max_cand_cap = 0;
    std::vector<ResizeTab> xtab, ytab;   // concatenated per level (level 0 unused)
    int xtab_off[kMaxLevels], ytab_off[kMaxLevels];
    std::vector<uint32_t> blur_work; // (level<<24 | tile_y<<12 | tile_x)
    std::vector<uint32_t> ffast_work; // k_fast_fused jobs (level<<24 | cell_row<<12 | first cell col); small-cell levels first
    int n_ffast_small = 0;
    std::vector<uint32_t> oct_lut;   // per-level octree lookup tables (see LevelGeom::lut_off)
};

struct Tables {                      // ORBextractor constructor, :410-470
    int nfeatures, nlevels, ini_th, min_th;
    float scale_factor_f;
    float scale[kMaxLevels], inv_scale[kMaxLevels], sigma2[kMaxLevels], inv_sigma2[kMaxLevels];
    int nfeat[kMaxLevels];
    int umax[16];
};

// host_tables.cpp
void build_tables(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, Tables *t);
// returns ORBX_OK / ORBX_ERR_UNSUPPORTED (+ message)
int build_geometry(const Tables &t, int width, int height, Geometry *g, std::string *err);

constexpr int kBlurTileW = 128, kBlurTileH = 64;

// host_match.cpp: the sequential walk of SearchByProjection over the GPU's candidate lists
int resolve_projection_matches(int n_last, int n_cur, const unsigned long long *cand, const int *count, const int *offset, const int32_t *nobs,
                               const float *last_angle, const float *cur_angle, int check_orientation, int32_t *cur_match);

// ... and of SearchForInitialization (steal bookkeeping, ratio test, rotation histogram)
int resolve_initialization_matches(int n1, int n2, const unsigned long long *cand, const int *count, const int *offset, const float *ang1,
                                   const float *ang2, float nnratio, int check_orientation, int32_t *m12);

// ... and of SearchByProjection(F, vpMapPoints, th) (claims, best / second best with levels, ratio on equal levels)
int resolve_local_matches(int n_mp, int n_feat, const unsigned long long *cand, const int *count, const int *offset, const int32_t *nobs,
                          const int32_t *feat_octave, const int32_t *feat_obs, float nnratio, int32_t *feat_match);

// ... and of SearchByBoW(KeyFrame*, Frame&, ...) over the per-node pair distances
int resolve_bow_matches(int n_entries, const int *entries, const uint16_t *dist, const int32_t *f_feats, int n_f, const float *kf_angle,
                        const float *f_angle, float nnratio, int check_orientation, int32_t *f_match);

int resolve_bow_matches_kf(int n_entries, const int *entries, const uint16_t *dist, const int32_t *feats2, int n1, int n2, const uint8_t *valid2,
                           const float *ang1, const float *ang2, float nnratio, int check_orientation, int32_t *m12);

// DBoW2 vocabulary tree (vocabulary.cpp): node 0 is the root, children of node i are child_ids[child_off[i] .. child_off[i+1])
struct VocHost {
    int k = 0, L = 0, scoring = 0, weighting = 0, nnodes = 0, nwords = 0;
    std::vector<int32_t> child_off, child_ids, word_id;
    std::vector<uint8_t> desc;           // [nnodes][32]
    std::vector<double> weight;
};
int voc_build(int k, int L, int scoring, int weighting, int nfile, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc,
              const double *weight, VocHost *v, std::string *err);
int voc_load_text(const char *path, VocHost *v, std::string *err);
int voc_bow(const VocHost &v, int n, const int32_t *word, const int32_t *node, const double *weight, int32_t *bow_ids, double *bow_vals,
            int *n_bow, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv);

} // namespace orbx

#endif
