// multimot_track_b200/csrc/host_tables.cpp -- host-side tables of the ORB front end.
//
// Everything here depends only on the five constructor parameters and the image
// size, never on pixel data, so it is evaluated once on the host with exactly
// the float/double arithmetic of the reference (compile with -ffp-contract=off):
//   constructor tables          src/ORBextractor.cc:410-470
//   level sizes                 :1116   (cvRound((float)cols*mvInvScaleFactor[level]))
//   FAST cell grid              :771-806
//   octree roots                :543-560
//   cv::resize INTER_LINEAR coefficient tables (OpenCV resize.cpp, 11-bit fixed point)
// The device kernels then run integer-only wherever the reference is integer.
#include "orbx_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace orbx {

static inline int cv_round(float v) { return (int)lrintf(v); }      // round-half-even, like cvRound
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

void build_tables(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, Tables *t)
{
    std::memset(t, 0, sizeof(*t));
    t->nfeatures = nfeatures; t->nlevels = nlevels; t->ini_th = ini_th; t->min_th = min_th;
    t->scale_factor_f = scale_factor;
    const double sf = (double)scale_factor;           // the member is a double holding the float (ORBextractor.h:98)
    t->scale[0] = 1.0f; t->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        t->scale[i] = (float)(t->scale[i - 1] * sf);
        t->sigma2[i] = t->scale[i] * t->scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        t->inv_scale[i] = 1.0f / t->scale[i];
        t->inv_sigma2[i] = 1.0f / t->sigma2[i];
    }
    const float factor = (float)(1.0f / sf);
    float per_scale = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; ++level) {
        t->nfeat[level] = cv_round(per_scale);
        sum += t->nfeat[level];
        per_scale *= factor;
    }
    t->nfeat[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    // circular patch row ends (radius 15), made symmetric like :453-469
    const int R = 15;
    int umax[R + 2] = {0};
    const int vmax = cv_floor(R * std::sqrt(2.f) / 2 + 1);
    const int vmin = cv_ceil(R * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt((double)R * R - (double)v * v));
    for (int v = R, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    for (int v = 0; v <= R; ++v) t->umax[v] = umax[v];
}

static inline int16_t sat16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

static void resize_tables(int sw, int sh, int dw, int dh, std::vector<ResizeTab> *xt, std::vector<ResizeTab> *yt)
{
    const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        ResizeTab r;
        r.s0 = (uint16_t)sx; r.s1 = (uint16_t)(sx + 1 < sw ? sx + 1 : sw - 1);
        r.c0 = sat16(cv_round((1.f - fx) * 2048.f)); r.c1 = sat16(cv_round(fx * 2048.f));
        xt->push_back(r);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        ResizeTab r;
        const int s0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        const int s1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        r.s0 = (uint16_t)s0; r.s1 = (uint16_t)s1;
        r.c0 = sat16(cv_round((1.f - fy) * 2048.f)); r.c1 = sat16(cv_round(fy * 2048.f));
        yt->push_back(r);
    }
    // k_resize reads 4 column entries per thread: pad both tables to a multiple of 4
    while (xt->size() % 4) xt->push_back(xt->back());
    while (yt->size() % 4) yt->push_back(yt->back());
}

int build_geometry(const Tables &t, int width, int height, Geometry *g, std::string *err)
{
    char msg[256];
    if (width > kMaxDim || height > kMaxDim) {
        std::snprintf(msg, sizeof msg, "image %dx%d exceeds the supported maximum %d", width, height, kMaxDim);
        *err = msg; return ORBX_ERR_UNSUPPORTED;
    }
    g->width = width; g->height = height; g->nlevels = t.nlevels;
    g->blur_work.clear(); g->ffast_work.clear(); g->oct_lut.clear(); g->xtab.clear(); g->ytab.clear();
    long long img_off = 0, cand_off = 0;
    int kp_off = 0;
    g->max_node_cap = g->max_feat = g->max_cand_cap = 0;
    for (int l = 0; l < t.nlevels; ++l) {
        LevelGeom &L = g->lv[l];
        std::memset(&L, 0, sizeof(L));
        L.w = cv_round((float)width * t.inv_scale[l]);
        L.h = cv_round((float)height * t.inv_scale[l]);
        // The reference computes nCols = (int)((w-32)/30.f) and divides by it (:776-778):
        // any level narrower or lower than 62 pixels is a division by zero there.
        if (L.w < 62 || L.h < 62) {
            std::snprintf(msg, sizeof msg, "pyramid level %d is %dx%d; the reference needs every level >= 62x62", l, L.w, L.h);
            *err = msg; return ORBX_ERR_UNSUPPORTED;
        }
        L.pitch = (L.w + 63) & ~63;
        L.img_off = img_off;
        img_off += ((long long)L.pitch * L.h + 255) & ~255LL;
        L.scale = t.scale[l];
        L.kp_size = (float)(int)(31 * t.scale[l]);

        // ---- FAST cells (:771-806)
        const int minBX = kMinBorder, minBY = kMinBorder;
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float Wc = 30;
        const float fw = (float)(maxBX - minBX), fh = (float)(maxBY - minBY);
        L.n_cols = (int)(fw / Wc); L.n_rows = (int)(fh / Wc);
        L.w_cell = (int)std::ceil(fw / L.n_cols); L.h_cell = (int)std::ceil(fh / L.n_rows);
        L.cols_vis = 0; L.rows_vis = 0;
        for (int j = 0; j < L.n_cols; ++j) {
            const float iniX = (float)(minBX + j * L.w_cell);
            if (iniX >= maxBX - 6) continue;
            L.cols_vis = j + 1;
        }
        for (int i = 0; i < L.n_rows; ++i) {
            const float iniY = (float)(minBY + i * L.h_cell);
            if (iniY >= maxBY - 3) continue;
            float maxY = iniY + L.h_cell + 6;
            if (maxY > maxBY) maxY = (float)maxBY;
            if ((int)maxY - (int)iniY < 7) continue;       // cv::FAST finds nothing in < 7 rows
            L.rows_vis = i + 1;
        }
        L.x_end = std::min(kEdge + L.cols_vis * L.w_cell, maxBX - 3);
        L.y_end = std::min(kEdge + L.rows_vis * L.h_cell, maxBY - 3);
        if (L.w_cell > 64 || L.h_cell > 64) {          // cannot happen for levels >= 62 px (cell < 60)
            *err = "internal: FAST cell larger than 64"; return ORBX_ERR_UNSUPPORTED;
        }
        L.cell_work_off = 0;
        long long cap = 0;                           // worst case survivors of the cell-local NMS: ceil(w/2)*ceil(h/2) per cell
        for (int i = 0; i < L.rows_vis; ++i)
            for (int j = 0; j < L.cols_vis; ++j) {
                const int x0 = kEdge + j * L.w_cell, x1 = std::min(x0 + L.w_cell, L.x_end);
                const int y0 = kEdge + i * L.h_cell, y1 = std::min(y0 + L.h_cell, L.y_end);
                if (x1 <= x0 || y1 <= y0) continue;
                cap += (long long)((x1 - x0 + 1) / 2) * ((y1 - y0 + 1) / 2);
            }
        L.cand_cap = (int)cap;
        L.cand_off = cand_off;
        cand_off += (cap + 63) & ~63LL;

        // ---- octree roots (:543-560)
        const int ow = maxBX - minBX, oh = maxBY - minBY;
        L.n_feat = t.nfeat[l];
        L.n_ini = (int)std::round((float)ow / oh);
        if (L.n_ini < 1 || L.n_ini > kMaxRoots) {
            std::snprintf(msg, sizeof msg, "level %d aspect %d:%d gives %d initial octree nodes (supported 1..%d; "
                          "the reference divides by zero for 0)", l, ow, oh, L.n_ini, kMaxRoots);
            *err = msg; return ORBX_ERR_UNSUPPORTED;
        }
        L.h_x = (float)ow / L.n_ini;
        int maxdim = oh;
        for (int i = 0; i < L.n_ini; ++i) {
            L.root_ul[i] = (int)(L.h_x * (float)i);
            L.root_br[i] = (int)(L.h_x * (float)(i + 1));
            maxdim = std::max(maxdim, L.root_br[i] - L.root_ul[i]);
        }
        L.region_h = oh;
        L.depth = 0;
        while ((1 << L.depth) < maxdim) ++L.depth;
        // Path-code tables.  DivideNode (:481-537) halves x and y independently (ceil(w/2) to the left/upper
        // child), so the 2-bit child index at every depth splits into an x bit and a y bit that depend only on
        // x (and the initial node, :569) or only on y: code(x, y) = lut_x[x] | lut_y[y].
        L.region_w = ow;
        L.lut_off = (int)g->oct_lut.size();
        for (int x = 0; x < ow; ++x) {
            int r = (int)((float)x / L.h_x);
            if (r > L.n_ini - 1) r = L.n_ini - 1;
            int ul = L.root_ul[r], br = L.root_br[r];
            uint32_t code = (uint32_t)r << (2 * L.depth);
            for (int d = 0; d < L.depth; ++d) {
                const int mid = ul + ((br - ul + 1) >> 1);
                const bool right = x >= mid;
                if (right) ul = mid; else br = mid;
                code |= (uint32_t)right << (2 * (L.depth - 1 - d));
            }
            g->oct_lut.push_back(code);
        }
        for (int y = 0; y < oh; ++y) {
            int ul = 0, br = oh;
            uint32_t code = 0;
            for (int d = 0; d < L.depth; ++d) {
                const int mid = ul + ((br - ul + 1) >> 1);
                const bool down = y >= mid;
                if (down) ul = mid; else br = mid;
                code |= (uint32_t)down << (2 * (L.depth - 1 - d) + 1);
            }
            g->oct_lut.push_back(code);
        }
        // candidate order of the reference = (cell row, cell column, y, x) (:789-829): cell indices per coordinate
        for (int x = 0; x < ow; ++x) g->oct_lut.push_back(x >= 3 ? (uint32_t)((x - 3) / L.w_cell) : 0u);
        for (int y = 0; y < oh; ++y) g->oct_lut.push_back(y >= 3 ? (uint32_t)((y - 3) / L.h_cell * L.n_cols) : 0u);
        const int nmax = std::max(L.n_feat, 4 * L.n_ini);
        L.kp_cap = nmax + 3;
        L.kp_off = kp_off;
        kp_off += L.kp_cap;
        L.node_cap = 2 * nmax + 4 * L.n_ini + 16;
        g->max_node_cap = std::max(g->max_node_cap, L.node_cap);
        g->max_feat = std::max(g->max_feat, nmax);
        g->max_cand_cap = std::max(g->max_cand_cap, L.cand_cap);

        // ---- resize tables from level l-1
        g->xtab_off[l] = (int)g->xtab.size(); g->ytab_off[l] = (int)g->ytab.size();
        if (l > 0) resize_tables(g->lv[l - 1].w, g->lv[l - 1].h, L.w, L.h, &g->xtab, &g->ytab);

        // ---- blur tiles
        for (int ty = 0; ty < (L.h + kBlurTileH - 1) / kBlurTileH; ++ty)
            for (int tx = 0; tx < (L.w + kBlurTileW - 1) / kBlurTileW; ++tx)
                g->blur_work.push_back((uint32_t)l << 24 | (uint32_t)ty << 12 | (uint32_t)tx);
    }
    // FAST jobs: one warp per cell row x (64 / w_cell) adjacent cells; levels with cells <= 44 px first (k_fast_fused<44>)
    {
        std::vector<uint32_t> small, large;
        for (int l = 0; l < t.nlevels; ++l) {
            const LevelGeom &L = g->lv[l];
            const int cpw = std::max(1, 64 / L.w_cell);
            for (int i = 0; i < L.rows_vis; ++i) {
                const int y0 = kEdge + i * L.h_cell;
                if (std::min(y0 + L.h_cell, L.y_end) <= y0) continue;
                for (int j = 0; j < L.cols_vis; j += cpw)
                    (L.w_cell <= 44 && L.h_cell <= 44 ? small : large).push_back((uint32_t)l << 24 | (uint32_t)i << 12 | (uint32_t)j);
            }
        }
        g->n_ffast_small = (int)small.size();
        g->ffast_work = small;
        g->ffast_work.insert(g->ffast_work.end(), large.begin(), large.end());
    }
    g->pyr_frame_bytes = img_off;
    g->cand_frame_elems = cand_off;
    g->kp_frame_cap = kp_off;
    return ORBX_OK;
}

} // namespace orbx
