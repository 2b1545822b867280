/*
 * oracle/orb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * Scalar restatement of the reference ORB front end.  Citations are to the
 * reference tree (src/ORBextractor.cc unless another file is named).  Build with
 * -ffp-contract=off so every float expression rounds where the source says.
 */
#include "orb_oracle.h"
#include "cvprim.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19 };      /* :72-74 */

static const int8_t kPattern[512 * 2] = {
#include "orb_pattern.inc"
};

struct orbo_extractor {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;                      /* include/ORBextractor.h:98 stores the float as double */
    float scale[ORBO_MAX_LEVELS], inv_scale[ORBO_MAX_LEVELS], sigma2[ORBO_MAX_LEVELS], inv_sigma2[ORBO_MAX_LEVELS];
    int nfeat[ORBO_MAX_LEVELS];
    int umax[HALF_PATCH_SIZE + 2];
    /* stage outputs of the last call */
    int lw[ORBO_MAX_LEVELS], lh[ORBO_MAX_LEVELS];
    uint8_t *img[ORBO_MAX_LEVELS], *blur[ORBO_MAX_LEVELS];
    size_t img_cap[ORBO_MAX_LEVELS];
    int has_blur[ORBO_MAX_LEVELS];
    orbo_cand *cand[ORBO_MAX_LEVELS];
    int ncand[ORBO_MAX_LEVELS], cand_cap[ORBO_MAX_LEVELS];
    int min_cells[ORBO_MAX_LEVELS], cells[ORBO_MAX_LEVELS];
    orbo_keypoint *lkp[ORBO_MAX_LEVELS];
    int nlkp[ORBO_MAX_LEVELS], lkp_cap[ORBO_MAX_LEVELS];
};

/* ------------------------------------------------------------ constructor */

orbo_extractor *orbo_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    if (nlevels < 1 || nlevels > ORBO_MAX_LEVELS) return NULL;
    orbo_extractor *e = (orbo_extractor *)calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = ini_th; e->min_th = min_th;
    e->scale_factor = (double)scale_factor;
    /* :415-431 -- float * double -> double -> float */
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scale_factor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        e->inv_scale[i] = 1.0f / e->scale[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    /* :435-447 -- geometric split of nfeatures over the levels */
    float factor = (float)(1.0f / e->scale_factor);
    float n_desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; ++level) {
        e->nfeat[level] = cvp_round_f(n_desired);
        sum += e->nfeat[level];
        n_desired *= factor;
    }
    e->nfeat[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* :453-469 -- row ends of the circular patch */
    int v, v0;
    int vmax = cvp_floor_d(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = cvp_ceil_d(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cvp_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

void orbo_destroy(orbo_extractor *e)
{
    if (!e) return;
    for (int l = 0; l < ORBO_MAX_LEVELS; ++l) { free(e->img[l]); free(e->blur[l]); free(e->cand[l]); free(e->lkp[l]); }
    free(e);
}

int orbo_tables(const orbo_extractor *e, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                int *nfeat, int *umax16)
{
    for (int i = 0; i < e->nlevels; ++i) {
        if (scale) scale[i] = e->scale[i];
        if (inv_scale) inv_scale[i] = e->inv_scale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->inv_sigma2[i];
        if (nfeat) nfeat[i] = e->nfeat[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
    return e->nlevels;
}

/* ----------------------------------------------------------------- octree */
/* DistributeOctTree / ExtractorNode::DivideNode, :481-763, canonical tie-break. */

typedef struct {
    int ulx, uly, brx, bry;      /* UL and BR corners (UR.x = brx, BL.y = bry) */
    int *keys; int nkeys;        /* indices into the candidate array, parent order kept (:519-534) */
    int no_more;
    long seq;                    /* creation sequence number (canonical tie-break) */
    int prev, next;              /* list links; the std::list of :545 */
} onode;

typedef struct {
    onode *n; int count, cap;
    int head, tail, size;
    long seq_counter;
} olist;

static int ol_new(olist *L)
{
    if (L->count == L->cap) { L->cap = L->cap ? L->cap * 2 : 256; L->n = (onode *)realloc(L->n, sizeof(onode) * (size_t)L->cap); }
    memset(&L->n[L->count], 0, sizeof(onode));
    L->n[L->count].prev = L->n[L->count].next = -1;
    return L->count++;
}
static void ol_push_back(olist *L, int i)
{
    L->n[i].seq = L->seq_counter++;
    L->n[i].prev = L->tail; L->n[i].next = -1;
    if (L->tail >= 0) L->n[L->tail].next = i; else L->head = i;
    L->tail = i; L->size++;
}
static void ol_push_front(olist *L, int i)
{
    L->n[i].seq = L->seq_counter++;
    L->n[i].next = L->head; L->n[i].prev = -1;
    if (L->head >= 0) L->n[L->head].prev = i; else L->tail = i;
    L->head = i; L->size++;
}
static int ol_erase(olist *L, int i)       /* returns the following element */
{
    const int p = L->n[i].prev, q = L->n[i].next;
    if (p >= 0) L->n[p].next = q; else L->head = q;
    if (q >= 0) L->n[q].prev = p; else L->tail = p;
    L->size--;
    free(L->n[i].keys); L->n[i].keys = NULL;
    return q;
}

/* DivideNode (:481-537).  Children are created in the node pool; returns their ids. */
static void divide_node(olist *L, int p, const orbo_cand *c, int ch[4])
{
    const int ulx = L->n[p].ulx, uly = L->n[p].uly, brx = L->n[p].brx, bry = L->n[p].bry;
    const int halfX = (int)ceil((double)((float)(brx - ulx) / 2));
    const int halfY = (int)ceil((double)((float)(bry - uly) / 2));
    const int nk = L->n[p].nkeys;
    for (int k = 0; k < 4; ++k) {
        ch[k] = ol_new(L);
        L->n[ch[k]].keys = (int *)malloc(sizeof(int) * (size_t)(nk ? nk : 1));
    }
    onode *n1 = &L->n[ch[0]], *n2 = &L->n[ch[1]], *n3 = &L->n[ch[2]], *n4 = &L->n[ch[3]];
    const int midx = ulx + halfX, midy = uly + halfY;
    n1->ulx = ulx;  n1->uly = uly;  n1->brx = midx; n1->bry = midy;
    n2->ulx = midx; n2->uly = uly;  n2->brx = brx;  n2->bry = midy;
    n3->ulx = ulx;  n3->uly = midy; n3->brx = midx; n3->bry = bry;
    n4->ulx = midx; n4->uly = midy; n4->brx = brx;  n4->bry = bry;
    const int *keys = L->n[p].keys;
    for (int i = 0; i < nk; ++i) {
        const orbo_cand *kp = &c[keys[i]];
        onode *dst;
        if (kp->x < midx) dst = kp->y < midy ? n1 : n3;
        else dst = kp->y < midy ? n2 : n4;
        dst->keys[dst->nkeys++] = keys[i];
    }
    for (int k = 0; k < 4; ++k) if (L->n[ch[k]].nkeys == 1) L->n[ch[k]].no_more = 1;
}

typedef struct { int size; int node; long seq; } size_node;
static int cmp_size_node(const void *a, const void *b)
{
    const size_node *x = (const size_node *)a, *y = (const size_node *)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);       /* canonical: seq instead of pointer */
}

/* push non-empty children to the front in order n1..n4 and record the splittable ones */
static void adopt_children(olist *L, const int ch[4], size_node **rec, int *nrec, int *caprec, int *n_to_expand)
{
    for (int k = 0; k < 4; ++k) {
        const int c = ch[k];
        if (L->n[c].nkeys > 0) {
            ol_push_front(L, c);
            if (L->n[c].nkeys > 1) {
                if (n_to_expand) (*n_to_expand)++;
                if (*nrec == *caprec) { *caprec = *caprec ? *caprec * 2 : 256; *rec = (size_node *)realloc(*rec, sizeof(size_node) * (size_t)*caprec); }
                (*rec)[*nrec].size = L->n[c].nkeys; (*rec)[*nrec].node = c; (*rec)[*nrec].seq = L->n[c].seq;
                (*nrec)++;
            }
        } else {
            free(L->n[c].keys); L->n[c].keys = NULL;
        }
    }
}

int orbo_distribute(const orbo_cand *cands, int n, int minX, int maxX, int minY, int maxY, int N,
                    orbo_cand *out, int cap)
{
    /* :543-546 */
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return 0;                      /* the reference divides by zero here; callers reject such shapes */
    const float hX = (float)(maxX - minX) / nIni;

    olist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int *roots = (int *)malloc(sizeof(int) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) {                                   /* :553-563 */
        const int id = ol_new(&L);
        L.n[id].ulx = (int)(hX * (float)i);
        L.n[id].brx = (int)(hX * (float)(i + 1));
        L.n[id].uly = 0; L.n[id].bry = maxY - minY;
        L.n[id].keys = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
        ol_push_back(&L, id);
        roots[i] = id;
    }
    for (int i = 0; i < n; ++i) {                                      /* :566-570 */
        int r = (int)((float)cands[i].x / hX);
        if (r >= nIni) r = nIni - 1;             /* unreachable for FAST candidates (x <= width-4) */
        onode *nd = &L.n[roots[r]];
        nd->keys[nd->nkeys++] = i;
    }
    for (int it = L.head; it >= 0;) {                                  /* :572-585 */
        if (L.n[it].nkeys == 1) { L.n[it].no_more = 1; it = L.n[it].next; }
        else if (L.n[it].nkeys == 0) it = ol_erase(&L, it);
        else it = L.n[it].next;
    }

    int finish = 0;
    size_node *rec = NULL; int nrec = 0, caprec = 0;
    while (!finish) {                                                  /* :594-737 */
        int prev_size = L.size;
        int n_to_expand = 0;
        nrec = 0;
        for (int it = L.head; it >= 0;) {
            if (L.n[it].no_more) { it = L.n[it].next; continue; }
            int ch[4];
            divide_node(&L, it, cands, ch);
            adopt_children(&L, ch, &rec, &nrec, &caprec, &n_to_expand);
            it = ol_erase(&L, it);
        }
        if (L.size >= N || L.size == prev_size) {
            finish = 1;
        } else if (L.size + n_to_expand * 3 > N) {
            while (!finish) {                                          /* :676-735 */
                prev_size = L.size;
                size_node *prev = (size_node *)malloc(sizeof(size_node) * (size_t)(nrec ? nrec : 1));
                const int nprev = nrec;
                memcpy(prev, rec, sizeof(size_node) * (size_t)nrec);
                nrec = 0;
                qsort(prev, (size_t)nprev, sizeof(size_node), cmp_size_node);
                for (int j = nprev - 1; j >= 0; --j) {
                    int ch[4];
                    divide_node(&L, prev[j].node, cands, ch);
                    adopt_children(&L, ch, &rec, &nrec, &caprec, NULL);
                    ol_erase(&L, prev[j].node);
                    if (L.size >= N) break;
                }
                free(prev);
                if (L.size >= N || L.size == prev_size) finish = 1;
            }
        }
    }
    /* :740-760 -- first strictly greatest response per node, list order */
    int m = 0;
    for (int it = L.head; it >= 0; it = L.n[it].next) {
        const onode *nd = &L.n[it];
        int best = nd->keys[0];
        for (int k = 1; k < nd->nkeys; ++k)
            if (cands[nd->keys[k]].score > cands[best].score) best = nd->keys[k];
        if (m < cap) out[m] = cands[best];
        ++m;
    }
    for (int i = 0; i < L.count; ++i) free(L.n[i].keys);
    free(L.n); free(rec); free(roots);
    return m;
}

/* ---------------------------------------------------------- orientation */

static float ic_angle(const uint8_t *img, int stride, float ptx, float pty, const int *umax)   /* :77-104 */
{
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t)cvp_round_f(pty) * stride + cvp_round_f(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return cvp_fast_atan2((float)m_01, (float)m_10);
}

/* ------------------------------------------------------------ descriptor */

static void orb_descriptor(const orbo_keypoint *kpt, const uint8_t *img, int stride, uint8_t *desc)   /* :108-147 */
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = kpt->angle * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t *center = img + (size_t)cvp_round_f(kpt->y) * stride + cvp_round_f(kpt->x);
    const int8_t *p = kPattern;
    for (int i = 0; i < 32; ++i, p += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int x0 = p[4 * k], y0 = p[4 * k + 1], x1 = p[4 * k + 2], y1 = p[4 * k + 3];
            const int t0 = center[cvp_round_f(x0 * b + y0 * a) * stride + cvp_round_f(x0 * a - y0 * b)];
            const int t1 = center[cvp_round_f(x1 * b + y1 * a) * stride + cvp_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------- operator() */

static void ensure_img(orbo_extractor *e, int l, size_t bytes)
{
    if (e->img_cap[l] < bytes) {
        free(e->img[l]); free(e->blur[l]);
        e->img[l] = (uint8_t *)malloc(bytes); e->blur[l] = (uint8_t *)malloc(bytes);
        e->img_cap[l] = bytes;
    }
}

static void cand_push(orbo_extractor *e, int l, int x, int y, int s)
{
    if (e->ncand[l] == e->cand_cap[l]) {
        e->cand_cap[l] = e->cand_cap[l] ? e->cand_cap[l] * 2 : 4096;
        e->cand[l] = (orbo_cand *)realloc(e->cand[l], sizeof(orbo_cand) * (size_t)e->cand_cap[l]);
    }
    orbo_cand *c = &e->cand[l][e->ncand[l]++];
    c->x = x; c->y = y; c->score = s;
}

int orbo_extract(orbo_extractor *e, const uint8_t *gray, int w, int h, int stride,
                 orbo_keypoint *kps, uint8_t *desc, int cap)
{
    const int L = e->nlevels;
    for (int l = 0; l < L; ++l) { e->ncand[l] = 0; e->nlkp[l] = 0; e->has_blur[l] = 0; e->lw[l] = e->lh[l] = 0; }
    if (!gray || w <= 0 || h <= 0) return 0;                           /* :1049 */

    /* ComputePyramid :1111-1136 (levels kept un-padded; the border is synthesised on demand) */
    for (int l = 0; l < L; ++l) {
        const float s = e->inv_scale[l];
        const int lw = cvp_round_f((float)w * s), lh = cvp_round_f((float)h * s);
        e->lw[l] = lw; e->lh[l] = lh;
        if (lw <= 0 || lh <= 0) { e->lw[l] = e->lh[l] = 0; continue; }
        ensure_img(e, l, (size_t)lw * (size_t)lh);
        if (l == 0) for (int y = 0; y < h; ++y) memcpy(e->img[0] + (size_t)y * w, gray + (size_t)y * stride, (size_t)w);
        else cvp_resize_linear_u8(e->img[l - 1], e->lw[l - 1], e->lh[l - 1], e->lw[l - 1], e->img[l], lw, lh, lw);
    }

    /* ComputeKeyPointsOctTree :765-853 */
    const float W = 30;
    cvp_corner *cell_buf = (cvp_corner *)malloc(sizeof(cvp_corner) * 4096);
    int cell_cap = 4096;
    for (int level = 0; level < L; ++level) {
        const int cols = e->lw[level], rows = e->lh[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
        e->min_cells[level] = e->cells[level] = 0;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) continue;       /* the reference divides by zero for such tiny levels */
        const int wCell = (int)ceil((double)(width / nCols)), hCell = (int)ceil((double)(height / nRows));
        const uint8_t *img = e->img[level];
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
                if (cw * chh > cell_cap) { cell_cap = cw * chh; cell_buf = (cvp_corner *)realloc(cell_buf, sizeof(cvp_corner) * (size_t)cell_cap); }
                const uint8_t *roi = img + (size_t)y0 * cols + x0;
                e->cells[level]++;
                int nk = cvp_fast9_16(roi, cw, chh, cols, e->ini_th, 1, cell_buf, cell_cap);      /* :809 */
                if (nk == 0) {
                    nk = cvp_fast9_16(roi, cw, chh, cols, e->min_th, 1, cell_buf, cell_cap);      /* :814 */
                    e->min_cells[level]++;
                }
                for (int k = 0; k < nk; ++k)                                                       /* :820-825 */
                    cand_push(e, level, cell_buf[k].x + j * wCell, cell_buf[k].y + i * hCell, cell_buf[k].score);
            }
        }
        /* :834 DistributeOctTree, :837-847 */
        const int capk = e->nfeat[level] + 64 > e->ncand[level] ? e->ncand[level] + 1 : e->nfeat[level] + 64;
        orbo_cand *sel = (orbo_cand *)malloc(sizeof(orbo_cand) * (size_t)(capk > 0 ? capk : 1));
        int nsel = orbo_distribute(e->cand[level], e->ncand[level], minBorderX, maxBorderX, minBorderY, maxBorderY,
                                   e->nfeat[level], sel, capk);
        if (nsel > capk) nsel = capk;               /* cannot happen: the list never exceeds N+3 */
        if (e->lkp_cap[level] < nsel) { free(e->lkp[level]); e->lkp[level] = (orbo_keypoint *)malloc(sizeof(orbo_keypoint) * (size_t)nsel); e->lkp_cap[level] = nsel; }
        const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[level]);
        for (int i = 0; i < nsel; ++i) {
            orbo_keypoint *k = &e->lkp[level][i];
            k->x = (float)sel[i].x + minBorderX; k->y = (float)sel[i].y + minBorderY;
            k->size = (float)scaledPatchSize; k->angle = -1.f; k->response = (float)sel[i].score;
            k->octave = level; k->class_id = -1;
        }
        e->nlkp[level] = nsel;
        free(sel);
    }
    free(cell_buf);
    for (int level = 0; level < L; ++level)                                                         /* :851-852 */
        for (int i = 0; i < e->nlkp[level]; ++i)
            e->lkp[level][i].angle = ic_angle(e->img[level], e->lw[level], e->lkp[level][i].x, e->lkp[level][i].y, e->umax);

    /* :1062-1108 */
    int n = 0;
    for (int level = 0; level < L; ++level) n += e->nlkp[level];
    const int write = kps && desc && n <= cap;
    int offset = 0;
    for (int level = 0; level < L; ++level) {
        const int nl = e->nlkp[level];
        if (nl == 0) continue;
        cvp_gaussian7x7_s2_u8(e->img[level], e->lw[level], e->lh[level], e->lw[level], e->blur[level], e->lw[level]);
        e->has_blur[level] = 1;
        if (write) {
            for (int i = 0; i < nl; ++i) {
                orbo_keypoint k = e->lkp[level][i];
                orb_descriptor(&k, e->blur[level], e->lw[level], desc + 32 * (size_t)(offset + i));
                if (level != 0) { const float s = e->scale[level]; k.x = k.x * s; k.y = k.y * s; }
                kps[offset + i] = k;
            }
        }
        offset += nl;
    }
    return (kps && desc && n > cap) ? -n : n;
}

int orbo_level_size(const orbo_extractor *e, int level, int *w, int *h)
{
    if (level < 0 || level >= e->nlevels) return -1;
    *w = e->lw[level]; *h = e->lh[level];
    return 0;
}
const uint8_t *orbo_level_image(const orbo_extractor *e, int level) { return e->img[level]; }
const uint8_t *orbo_level_blurred(const orbo_extractor *e, int level) { return e->has_blur[level] ? e->blur[level] : NULL; }
int orbo_level_candidates(const orbo_extractor *e, int level, const orbo_cand **out) { *out = e->cand[level]; return e->ncand[level]; }
int orbo_level_min_cells(const orbo_extractor *e, int level, int *cells) { if (cells) *cells = e->cells[level]; return e->min_cells[level]; }
int orbo_level_nkeypoints(const orbo_extractor *e, int level) { return e->nlkp[level]; }
int orbo_level_padded(const orbo_extractor *e, int level, uint8_t *dst, int dst_stride)
{
    if (level < 0 || level >= e->nlevels || !e->lw[level]) return -1;
    cvp_border_reflect101_u8(e->img[level], e->lw[level], e->lh[level], e->lw[level], dst, dst_stride,
                             EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD);
    return 0;
}

/* ---------------------------------------------------------------- matcher */

int orbo_hamming256(const uint8_t *a, const uint8_t *b)        /* src/ORBmatcher.cc:2279-2295 */
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

static int match_rows(const uint8_t *A, int i0, int i1, const uint8_t *B, int nB, int th, float ratio,
                      int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept)
{
    int acc = 0;
    for (int i = i0; i < i1; ++i) {                            /* src/ORBmatcher.cc:574-605 */
        int best1 = 256, best2 = 256, bi = -1;
        for (int j = 0; j < nB; ++j) {
            const int d = orbo_hamming256(A + 32 * (size_t)i, B + 32 * (size_t)j);
            if (d < best1) { best2 = best1; best1 = d; bi = j; }
            else if (d < best2) best2 = d;
        }
        idx[i] = bi; d1[i] = best1; d2[i] = best2;
        const int ok = bi >= 0 && best1 <= th && (float)best1 < ratio * (float)best2;
        if (accept) accept[i] = (uint8_t)ok;
        acc += ok;
    }
    return acc;
}

int orbo_match(const uint8_t *A, int nA, const uint8_t *B, int nB, int th, float ratio,
               int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept)
{
    return match_rows(A, 0, nA, B, nB, th, ratio, idx, d1, d2, accept);
}

typedef struct { const uint8_t *A, *B; int i0, i1, nB, th; float ratio; int32_t *idx, *d1, *d2; uint8_t *accept; int acc; } match_job;
static void *match_worker(void *p)
{
    match_job *j = (match_job *)p;
    j->acc = match_rows(j->A, j->i0, j->i1, j->B, j->nB, j->th, j->ratio, j->idx, j->d1, j->d2, j->accept);
    return NULL;
}

int orbo_match_mt(const uint8_t *A, int nA, const uint8_t *B, int nB, int th, float ratio,
                  int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > nA) threads = nA > 0 ? nA : 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    match_job *jobs = (match_job *)malloc(sizeof(match_job) * (size_t)threads);
    for (int k = 0; k < threads; ++k) {
        match_job j = {A, B, (int)((long)nA * k / threads), (int)((long)nA * (k + 1) / threads), nB, th, ratio, idx, d1, d2, accept, 0};
        jobs[k] = j;
        pthread_create(&t[k], NULL, match_worker, &jobs[k]);
    }
    int acc = 0;
    for (int k = 0; k < threads; ++k) { pthread_join(t[k], NULL); acc += jobs[k].acc; }
    free(t); free(jobs);
    return acc;
}

/* Rotation-consistency check of the matchers (mbCheckOrientation): histogram fill as in src/ORBmatcher.cc:610-620
 * (factor = 1.0f/HISTO_LENGTH, :545), ComputeThreeMaxima :2233-2274, pruning :641-660. */
static void three_maxima(const int *cnt, int L, int *ind1, int *ind2, int *ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

int orbo_rotation_bin(float angle_a, float angle_b)
{
    const float factor = 1.0f / 30;
    float rot = angle_a - angle_b;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == 30) bin = 0;
    return bin;
}

int orbo_rotation_filter(int nA, const int32_t *idx, uint8_t *accept, const float *angleA, const float *angleB,
                         int32_t *hist, int32_t *top3)
{
    int cnt[30] = {0};
    for (int i = 0; i < nA; ++i)
        if (accept[i]) cnt[orbo_rotation_bin(angleA[i], angleB[idx[i]])]++;
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(cnt, 30, &i1, &i2, &i3);
    int kept = 0;
    for (int i = 0; i < nA; ++i)
        if (accept[i]) {
            const int b = orbo_rotation_bin(angleA[i], angleB[idx[i]]);
            if (b == i1 || b == i2 || b == i3) ++kept; else accept[i] = 0;
        }
    if (hist) for (int i = 0; i < 30; ++i) hist[i] = cnt[i];
    if (top3) { top3[0] = i1; top3[1] = i2; top3[2] = i3; }
    return kept;
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:272-301: all-pairs Hamming distances of a map point's n
 * observed descriptors, per-row median = sorted row[(int)(0.5*(n-1))], first row with the least median. */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
int orbo_distinctive_descriptor(const uint8_t *desc, int n, int32_t *best_median)
{
    int best = INT_MAX, best_idx = 0;
    int *row = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) row[j] = i == j ? 0 : orbo_hamming256(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
        qsort(row, (size_t)n, sizeof(int), cmp_int);
        const int median = row[(int)(0.5 * (n - 1))];
        if (median < best) { best = median; best_idx = i; }
    }
    free(row);
    if (best_median) *best_median = best;
    return best_idx;
}

/* ------------------------------------------------------- projection matcher
 * ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), src/ORBmatcher.cc:1958-2102, with the
 * grid of the current frame (Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, src/Frame.cc:601-616, 710-776).
 * Matrix products follow OpenCV 4.13's gemm for CV_32F (A*B [+ C]: float accumulation, k ascending; -A.t()*B: double
 * accumulation), every other float operation is one statement in the reference's order (-ffp-contract=off). */
static void three_maxima(const int *cnt, int L, int *ind1, int *ind2, int *ind3);

int orbo_search_by_projection(const float *cam, const float *Tc, const float *Tl,
                              int nL, const float *world_pos, const uint8_t *mp_desc, const uint8_t *valid, const int32_t *nobs,
                              const int32_t *last_octave, const float *last_angle,
                              int nC, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                              const uint8_t *cur_desc, const float *scale, int nlevels, float th, int mono, int check_orientation,
                              int32_t *cur_match)
{
    (void)nlevels;
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], mbf = cam[4], mb = cam[5];
    const float minX = cam[6], maxX = cam[7], minY = cam[8], maxY = cam[9];
    const float wInv = 64.0f / (maxX - minX), hInv = 48.0f / (maxY - minY);
    /* AssignFeaturesToGrid: cell lists in feature order */
    int *cell = (int *)malloc(sizeof(int) * (size_t)(nC > 0 ? nC : 1));
    int *off = (int *)calloc(64 * 48 + 1, sizeof(int)), *idx = (int *)malloc(sizeof(int) * (size_t)(nC > 0 ? nC : 1));
    for (int i = 0; i < nC; ++i) {
        const int px = (int)roundf((cur_xy[2 * i] - minX) * wInv), py = (int)roundf((cur_xy[2 * i + 1] - minY) * hInv);
        cell[i] = (px < 0 || px >= 64 || py < 0 || py >= 48) ? -1 : px * 48 + py;
        if (cell[i] >= 0) off[cell[i] + 1]++;
    }
    for (int c = 0; c < 64 * 48; ++c) off[c + 1] += off[c];
    {
        int *fill = (int *)calloc(64 * 48, sizeof(int));
        for (int i = 0; i < nC; ++i) if (cell[i] >= 0) idx[off[cell[i]] + fill[cell[i]]++] = i;
        free(fill);
    }
    /* camera geometry, :1968-1979 */
    float twc[3], tlc[3];
    for (int r = 0; r < 3; ++r) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)Tc[4 * k + r] * (double)Tc[4 * k + 3];
        twc[r] = (float)(s * -1.0);
    }
    for (int r = 0; r < 3; ++r) {
        float s = 0.f;
        for (int k = 0; k < 3; ++k) { const float p = Tl[4 * r + k] * twc[k]; s = s + p; }
        tlc[r] = s + Tl[4 * r + 3];
    }
    const int bForward = tlc[2] > mb && !mono, bBackward = -tlc[2] > mb && !mono;
    int *claim_obs = (int *)calloc((size_t)(nC > 0 ? nC : 1), sizeof(int));
    for (int i = 0; i < nC; ++i) cur_match[i] = -1;
    int *hist_items = (int *)malloc(sizeof(int) * (size_t)(nL > 0 ? nL : 1)), *hist_bin = (int *)malloc(sizeof(int) * (size_t)(nL > 0 ? nL : 1));
    int nh = 0, nmatches = 0;
    for (int i = 0; i < nL; ++i) {
        if (!valid[i]) continue;
        float xc3[3];
        for (int r = 0; r < 3; ++r) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) { const float p = Tc[4 * r + k] * world_pos[3 * i + k]; s = s + p; }
            xc3[r] = s + Tc[4 * r + 3];
        }
        const float xc = xc3[0], yc = xc3[1];
        const float invzc = (float)(1.0 / (double)xc3[2]);
        if (invzc < 0) continue;
        float u = fx * xc; u = u * invzc; u = u + cx;
        float v = fy * yc; v = v * invzc; v = v + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const int oct = last_octave[i];
        const float radius = th * scale[oct];
        int minLevel, maxLevel;
        if (bForward) { minLevel = oct; maxLevel = -1; }
        else if (bBackward) { minLevel = 0; maxLevel = oct; }
        else { minLevel = oct - 1; maxLevel = oct + 1; }
        /* GetFeaturesInArea, src/Frame.cc:710-763 */
        int nMinCellX = (int)floorf((u - minX - radius) * wInv); if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= 64) continue;
        int nMaxCellX = (int)ceilf((u - minX + radius) * wInv); if (nMaxCellX > 63) nMaxCellX = 63;
        if (nMaxCellX < 0) continue;
        int nMinCellY = (int)floorf((v - minY - radius) * hInv); if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= 48) continue;
        int nMaxCellY = (int)ceilf((v - minY + radius) * hInv); if (nMaxCellY > 47) nMaxCellY = 47;
        if (nMaxCellY < 0) continue;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        int bestDist = 256, bestIdx2 = -1, any = 0;
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int j = off[ix * 48 + iy]; j < off[ix * 48 + iy + 1]; ++j) {
                    const int i2 = idx[j];
                    if (bCheckLevels) {
                        if (cur_octave[i2] < minLevel) continue;
                        if (maxLevel >= 0 && cur_octave[i2] > maxLevel) continue;
                    }
                    const float distx = cur_xy[2 * i2] - u, disty = cur_xy[2 * i2 + 1] - v;
                    if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
                    any = 1;                                                    /* member of vIndices2 */
                    if (cur_match[i2] >= 0 && claim_obs[i2] > 0) continue;      /* :2028-2030 */
                    if (cur_uright[i2] > 0) {
                        const float t = mbf * invzc;
                        const float ur = u - t;
                        const float er = fabsf(ur - cur_uright[i2]);
                        if (er > radius) continue;
                    }
                    const int dist = orbo_hamming256(mp_desc + 32 * (size_t)i, cur_desc + 32 * (size_t)i2);
                    if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
                }
        if (!any) continue;
        if (bestDist <= 100) {
            cur_match[bestIdx2] = i; claim_obs[bestIdx2] = nobs[i];
            ++nmatches;
            if (check_orientation) { hist_items[nh] = bestIdx2; hist_bin[nh] = orbo_rotation_bin(last_angle[i], cur_angle[bestIdx2]); ++nh; }
        }
    }
    if (check_orientation) {
        int cnt[30] = {0}, i1 = -1, i2 = -1, i3 = -1;
        for (int k = 0; k < nh; ++k) cnt[hist_bin[k]]++;
        three_maxima(cnt, 30, &i1, &i2, &i3);
        for (int k = 0; k < nh; ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { cur_match[hist_items[k]] = -1; --nmatches; }
    }
    free(cell); free(off); free(idx); free(claim_obs); free(hist_items); free(hist_bin);
    return nmatches;
}

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:780-895: octave-0
 * keypoints of F1 search F2's grid around their previous match position; a candidate already matched at a distance <= ours is
 * skipped, a better match steals it (the old owner loses its match); ratio test, rotation histogram, vbPrevMatched update. */
int orbo_search_for_initialization(const float *cam, int n1, const float *xy1, const int32_t *oct1, const float *ang1, const uint8_t *desc1,
                                   int n2, const float *xy2, const int32_t *oct2, const float *ang2, const uint8_t *desc2,
                                   float *prev_xy, int window, float nnratio, int check_orientation, int32_t *m12)
{
    (void)xy1;
    const float minX = cam[6], maxX = cam[7], minY = cam[8], maxY = cam[9];
    const float wInv = 64.0f / (maxX - minX), hInv = 48.0f / (maxY - minY);
    int *cell = (int *)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    int *off = (int *)calloc(64 * 48 + 1, sizeof(int)), *idx = (int *)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; ++i) {
        const int px = (int)roundf((xy2[2 * i] - minX) * wInv), py = (int)roundf((xy2[2 * i + 1] - minY) * hInv);
        cell[i] = (px < 0 || px >= 64 || py < 0 || py >= 48) ? -1 : px * 48 + py;
        if (cell[i] >= 0) off[cell[i] + 1]++;
    }
    for (int c = 0; c < 64 * 48; ++c) off[c + 1] += off[c];
    {
        int *fill = (int *)calloc(64 * 48, sizeof(int));
        for (int i = 0; i < n2; ++i) if (cell[i] >= 0) idx[off[cell[i]] + fill[cell[i]]++] = i;
        free(fill);
    }
    int *matched_dist = (int *)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1)), *m21 = (int *)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; ++i) { matched_dist[i] = INT_MAX; m21[i] = -1; }
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    int *hist_item = (int *)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1)), *hist_bin = (int *)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    int nh = 0, nmatches = 0;
    const float r = (float)window;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = oct1[i1];
        if (level1 > 0) continue;
        const float x = prev_xy[2 * i1], y = prev_xy[2 * i1 + 1];
        int nMinCellX = (int)floorf((x - minX - r) * wInv); if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= 64) continue;
        int nMaxCellX = (int)ceilf((x - minX + r) * wInv); if (nMaxCellX > 63) nMaxCellX = 63;
        if (nMaxCellX < 0) continue;
        int nMinCellY = (int)floorf((y - minY - r) * hInv); if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= 48) continue;
        int nMaxCellY = (int)ceilf((y - minY + r) * hInv); if (nMaxCellY > 47) nMaxCellY = 47;
        if (nMaxCellY < 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int j = off[ix * 48 + iy]; j < off[ix * 48 + iy + 1]; ++j) {
                    const int i2 = idx[j];
                    if (oct2[i2] < level1 || oct2[i2] > level1) continue;          /* GetFeaturesInArea(.., level1, level1): bCheckLevels */
                    const float distx = xy2[2 * i2] - x, disty = xy2[2 * i2 + 1] - y;
                    if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
                    const int dist = orbo_hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
                    if (matched_dist[i2] <= dist) continue;
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
        if (bestDist <= 50) {
            if ((float)bestDist < (float)bestDist2 * nnratio) {
                if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; --nmatches; }
                m12[i1] = bestIdx2; m21[bestIdx2] = i1; matched_dist[bestIdx2] = bestDist;
                ++nmatches;
                if (check_orientation) { hist_item[nh] = i1; hist_bin[nh] = orbo_rotation_bin(ang1[i1], ang2[bestIdx2]); ++nh; }
            }
        }
    }
    if (check_orientation) {
        int cnt[30] = {0}, i1 = -1, i2 = -1, i3 = -1;
        for (int k = 0; k < nh; ++k) cnt[hist_bin[k]]++;
        three_maxima(cnt, 30, &i1, &i2, &i3);
        for (int b = 0; b < 30; ++b) {                                         /* bins in order, items in push order, :872-885 */
            if (b == i1 || b == i2 || b == i3) continue;
            for (int k = 0; k < nh; ++k)
                if (hist_bin[k] == b && m12[hist_item[k]] >= 0) { m12[hist_item[k]] = -1; --nmatches; }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (m12[i1] >= 0) { prev_xy[2 * i1] = xy2[2 * m12[i1]]; prev_xy[2 * i1 + 1] = xy2[2 * m12[i1] + 1]; }
    free(cell); free(off); free(idx); free(matched_dist); free(m21); free(hist_item); free(hist_bin);
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th), src/ORBmatcher.cc:418-502: every local map
 * point that isInFrustum looks in a window of RadiusByViewingCos (x th) x scale[predicted level] around its projection, levels
 * (predicted - 1, predicted); features that hold an observed map point are skipped (also those assigned earlier in this call);
 * best / second best with their octaves, TH_HIGH, ratio only when both are on the same level. */
int orbo_search_local_points(const float *cam, int nP, const float *proj, const float *view_cos, const int32_t *level, const uint8_t *mp_desc,
                             const uint8_t *valid, const int32_t *nobs, int nF, const float *xy, const int32_t *octave, const float *uright,
                             const uint8_t *desc, const int32_t *feat_obs, const float *scale, int nlevels, float th, float nnratio,
                             int32_t *feat_match)
{
    (void)nlevels;
    const float minX = cam[6], maxX = cam[7], minY = cam[8], maxY = cam[9];
    const float wInv = 64.0f / (maxX - minX), hInv = 48.0f / (maxY - minY);
    int *cell = (int *)malloc(sizeof(int) * (size_t)(nF > 0 ? nF : 1));
    int *off = (int *)calloc(64 * 48 + 1, sizeof(int)), *idx = (int *)malloc(sizeof(int) * (size_t)(nF > 0 ? nF : 1));
    for (int i = 0; i < nF; ++i) {
        const int px = (int)roundf((xy[2 * i] - minX) * wInv), py = (int)roundf((xy[2 * i + 1] - minY) * hInv);
        cell[i] = (px < 0 || px >= 64 || py < 0 || py >= 48) ? -1 : px * 48 + py;
        if (cell[i] >= 0) off[cell[i] + 1]++;
    }
    for (int c = 0; c < 64 * 48; ++c) off[c + 1] += off[c];
    {
        int *fill = (int *)calloc(64 * 48, sizeof(int));
        for (int i = 0; i < nF; ++i) if (cell[i] >= 0) idx[off[cell[i]] + fill[cell[i]]++] = i;
        free(fill);
    }
    int *held_obs = (int *)malloc(sizeof(int) * (size_t)(nF > 0 ? nF : 1));
    for (int j = 0; j < nF; ++j) { held_obs[j] = feat_obs[j] > 0 ? feat_obs[j] : 0; feat_match[j] = -1; }
    const int bFactor = th != 1.0;
    int nmatches = 0;
    for (int i = 0; i < nP; ++i) {
        if (!valid[i]) continue;
        const int lvl = level[i];
        float r = view_cos[i] > 0.998 ? 2.5f : 4.0f;                           /* RadiusByViewingCos, :504-510 (float vs double 0.998) */
        if (bFactor) r *= th;
        const float radius = r * scale[lvl];
        const float x = proj[3 * i], y = proj[3 * i + 1];
        const int minLevel = lvl - 1, maxLevel = lvl;
        int nMinCellX = (int)floorf((x - minX - radius) * wInv); if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= 64) continue;
        int nMaxCellX = (int)ceilf((x - minX + radius) * wInv); if (nMaxCellX > 63) nMaxCellX = 63;
        if (nMaxCellX < 0) continue;
        int nMinCellY = (int)floorf((y - minY - radius) * hInv); if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= 48) continue;
        int nMaxCellY = (int)ceilf((y - minY + radius) * hInv); if (nMaxCellY > 47) nMaxCellY = 47;
        if (nMaxCellY < 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int k = off[ix * 48 + iy]; k < off[ix * 48 + iy + 1]; ++k) {
                    const int j = idx[k];
                    if (octave[j] < minLevel) continue;                        /* bCheckLevels is always true here: maxLevel >= 0 */
                    if (octave[j] > maxLevel) continue;
                    const float distx = xy[2 * j] - x, disty = xy[2 * j + 1] - y;
                    if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
                    if (held_obs[j] > 0) continue;
                    if (uright[j] > 0) {
                        const float er = fabsf(proj[3 * i + 2] - uright[j]);
                        if (er > radius) continue;
                    }
                    const int dist = orbo_hamming256(mp_desc + 32 * (size_t)i, desc + 32 * (size_t)j);
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = octave[j]; bestIdx = j; }
                    else if (dist < bestDist2) { bestLevel2 = octave[j]; bestDist2 = dist; }
                }
        if (bestDist <= 100) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            feat_match[bestIdx] = i; held_obs[bestIdx] = nobs[i] > 0 ? nobs[i] : 0;
            ++nmatches;
        }
    }
    free(cell); free(off); free(idx); free(held_obs);
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, ...), src/ORBmatcher.cc:532-663: merge walk over the two feature vectors; in a shared
 * node every key-frame feature with a good map point scans the frame's features of that node that are still unmatched; TH_LOW, ratio,
 * rotation histogram whose pruning decrements unconditionally (:649-654). */
int orbo_search_by_bow(int nK, const float *kf_angle, const uint8_t *kf_desc, const uint8_t *kf_valid, int nk_nodes, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_feats, int nF, const float *f_angle, const uint8_t *f_desc, int nf_nodes,
                       const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_feats, float nnratio, int check_orientation, int32_t *f_match)
{
    (void)nK;
    for (int j = 0; j < nF; ++j) f_match[j] = -1;
    int *hist_item = (int *)malloc(sizeof(int) * (size_t)(nF > 0 ? nF : 1)), *hist_bin = (int *)malloc(sizeof(int) * (size_t)(nF > 0 ? nF : 1));
    int nh = 0, nmatches = 0, a = 0, b = 0;
    while (a < nk_nodes && b < nf_nodes) {
        if (kf_nodes[a] < f_nodes[b]) { ++a; continue; }                       /* lower_bound on a sorted map == advance */
        if (kf_nodes[a] > f_nodes[b]) { ++b; continue; }
        for (int p = kf_off[a]; p < kf_off[a + 1]; ++p) {
            const int ik = kf_feats[p];
            if (!kf_valid[ik]) continue;
            int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
            for (int q = f_off[b]; q < f_off[b + 1]; ++q) {
                const int jf = f_feats[q];
                if (f_match[jf] >= 0) continue;
                const int dist = orbo_hamming256(kf_desc + 32 * (size_t)ik, f_desc + 32 * (size_t)jf);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = jf; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestDist1 <= 50 && (float)bestDist1 < nnratio * (float)bestDist2) {
                f_match[bestIdxF] = ik;
                if (check_orientation) { hist_item[nh] = bestIdxF; hist_bin[nh] = orbo_rotation_bin(kf_angle[ik], f_angle[bestIdxF]); ++nh; }
                ++nmatches;
            }
        }
        ++a; ++b;
    }
    if (check_orientation) {
        int cnt[30] = {0}, i1 = -1, i2 = -1, i3 = -1;
        for (int k = 0; k < nh; ++k) cnt[hist_bin[k]]++;
        three_maxima(cnt, 30, &i1, &i2, &i3);
        for (int k = 0; k < nh; ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { f_match[hist_item[k]] = -1; --nmatches; }
    }
    free(hist_item); free(hist_bin);
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12), src/ORBmatcher.cc:897-1030: as above between two key frames;
 * side 2 must hold a good map point and not be matched yet (:948-954), the threshold is strict (bestDist1 < TH_LOW, :973). */
int orbo_search_by_bow_kf(int n1, const float *ang1, const uint8_t *desc1, const uint8_t *valid1, int nn1, const int32_t *nodes1, const int32_t *off1,
                          const int32_t *feats1, int n2, const float *ang2, const uint8_t *desc2, const uint8_t *valid2, int nn2, const int32_t *nodes2,
                          const int32_t *off2, const int32_t *feats2, float nnratio, int check_orientation, int32_t *m12)
{
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    unsigned char *matched2 = (unsigned char *)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
    int *hist_item = (int *)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1)), *hist_bin = (int *)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    int nh = 0, nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] < nodes2[b]) { ++a; continue; }
        if (nodes1[a] > nodes2[b]) { ++b; continue; }
        for (int p = off1[a]; p < off1[a + 1]; ++p) {
            const int i1 = feats1[p];
            if (!valid1[i1]) continue;
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int q = off2[b]; q < off2[b + 1]; ++q) {
                const int i2 = feats2[q];
                if (matched2[i2] || !valid2[i2]) continue;
                const int dist = orbo_hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = i2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestDist1 < 50 && (float)bestDist1 < nnratio * (float)bestDist2) {
                m12[i1] = bestIdx2; matched2[bestIdx2] = 1;
                if (check_orientation) { hist_item[nh] = i1; hist_bin[nh] = orbo_rotation_bin(ang1[i1], ang2[bestIdx2]); ++nh; }
                ++nmatches;
            }
        }
        ++a; ++b;
    }
    if (check_orientation) {
        int cnt[30] = {0}, i1 = -1, i2 = -1, i3 = -1;
        for (int k = 0; k < nh; ++k) cnt[hist_bin[k]]++;
        three_maxima(cnt, 30, &i1, &i2, &i3);
        for (int k = 0; k < nh; ++k)
            if (hist_bin[k] != i1 && hist_bin[k] != i2 && hist_bin[k] != i3) { m12[hist_item[k]] = -1; --nmatches; }
    }
    free(matched2); free(hist_item); free(hist_bin);
    return nmatches;
}

/* ------------------------------------------------------------- vocabulary
 * DBoW2 as vendored by the reference (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h): the tree built the way
 * loadFromTextFile builds it (:1338-1418: node ids in file order, children in file order, word ids in order of the
 * leaf flag), the per-feature descent (:1205-1250: first child with the least FORB::distance, stop at a node without
 * children, node id recorded at level L - levelsup) and the BowVector / FeatureVector assembly of transform(features)
 * (:1127-1195) with BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:32-86) in double. */
struct orbo_voc {
    int k, L, scoring, weighting, nnodes;        /* nnodes includes the root (id 0) */
    int *child_off, *child_ids, *word_id;
    uint8_t *desc; double *weight;
};

orbo_voc *orbo_voc_create(int k, int L, int scoring, int weighting, int nfile, const int32_t *parent, const uint8_t *is_leaf,
                          const uint8_t *desc, const double *weight)
{
    orbo_voc *v = (orbo_voc *)calloc(1, sizeof(orbo_voc));
    const int n = nfile + 1;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = n;
    v->child_off = (int *)calloc((size_t)n + 1, sizeof(int));
    v->child_ids = (int *)calloc((size_t)n, sizeof(int));
    v->word_id = (int *)malloc(sizeof(int) * (size_t)n);
    v->desc = (uint8_t *)calloc((size_t)n, 32);
    v->weight = (double *)calloc((size_t)n, sizeof(double));
    int *cnt = (int *)calloc((size_t)n, sizeof(int));
    for (int i = 0; i < nfile; ++i) cnt[parent[i]]++;
    for (int i = 0; i < n; ++i) v->child_off[i + 1] = v->child_off[i] + cnt[i];
    memset(cnt, 0, sizeof(int) * (size_t)n);
    int words = 0;
    v->word_id[0] = -1;
    for (int i = 0; i < nfile; ++i) {
        const int nid = i + 1, pid = parent[i];
        v->child_ids[v->child_off[pid] + cnt[pid]++] = nid;
        memcpy(v->desc + 32 * (size_t)nid, desc + 32 * (size_t)i, 32);
        v->weight[nid] = weight[i];
        v->word_id[nid] = is_leaf[i] ? words++ : -1;
    }
    free(cnt);
    return v;
}

void orbo_voc_destroy(orbo_voc *v)
{
    if (!v) return;
    free(v->child_off); free(v->child_ids); free(v->word_id); free(v->desc); free(v->weight); free(v);
}

void orbo_voc_transform_each(const orbo_voc *v, const uint8_t *desc, int n, int levelsup, int32_t *word, int32_t *node, double *weight)
{
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; ++i) {
        const uint8_t *f = desc + 32 * (size_t)i;
        int final_id = 0, current_level = 0, nid = 0;
        do {
            ++current_level;
            const int *ch = v->child_ids + v->child_off[final_id];
            const int nc = v->child_off[final_id + 1] - v->child_off[final_id];
            final_id = ch[0];
            int best_d = orbo_hamming256(f, v->desc + 32 * (size_t)final_id);
            for (int c = 1; c < nc; ++c) {
                const int d = orbo_hamming256(f, v->desc + 32 * (size_t)ch[c]);
                if (d < best_d) { best_d = d; final_id = ch[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (v->child_off[final_id + 1] - v->child_off[final_id] > 0);
        word[i] = v->word_id[final_id]; node[i] = nid; weight[i] = v->weight[final_id];
    }
}

typedef struct { int32_t key; int32_t idx; } kv_pair;
static int cmp_kv(const void *a, const void *b)
{
    const kv_pair *x = (const kv_pair *)a, *y = (const kv_pair *)b;
    return x->key != y->key ? (x->key < y->key ? -1 : 1) : (x->idx < y->idx ? -1 : (x->idx > y->idx));
}

int orbo_voc_bow(const orbo_voc *v, int n, const int32_t *word, const int32_t *node, const double *weight,
                 int32_t *bow_ids, double *bow_vals, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv)
{
    /* scoring -> (must normalize, norm): L1, L2, CHI_SQUARE, KL, BHATTACHARYYA normalize (L2 with the L2 norm), DOT_PRODUCT not */
    const int must = v->scoring != 5, l2 = v->scoring == 1;
    kv_pair *kw = (kv_pair *)malloc(sizeof(kv_pair) * (size_t)(n > 0 ? n : 1)), *kn = (kv_pair *)malloc(sizeof(kv_pair) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) if (weight[i] > 0) { kw[m].key = word[i]; kw[m].idx = i; kn[m].key = node[i]; kn[m].idx = i; ++m; }
    qsort(kw, (size_t)m, sizeof(kv_pair), cmp_kv);                          /* by word, then feature order: the order addWeight sees */
    qsort(kn, (size_t)m, sizeof(kv_pair), cmp_kv);
    int nb = 0;
    for (int i = 0; i < m;) {
        int j = i;
        double acc = weight[kw[i].idx];                                     /* insert(id, v) */
        for (j = i + 1; j < m && kw[j].key == kw[i].key; ++j)
            if (v->weighting == 0 || v->weighting == 1) acc += weight[kw[j].idx];   /* TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist */
        bow_ids[nb] = kw[i].key; bow_vals[nb] = acc; ++nb;
        i = j;
    }
    if ((v->weighting == 0 || v->weighting == 1) && nb > 0 && !must) {
        const double nd = nb;
        for (int i = 0; i < nb; ++i) bow_vals[i] /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) for (int i = 0; i < nb; ++i) norm += fabs(bow_vals[i]);
        else { for (int i = 0; i < nb; ++i) norm += bow_vals[i] * bow_vals[i]; norm = sqrt(norm); }
        if (norm > 0.0) for (int i = 0; i < nb; ++i) bow_vals[i] /= norm;
    }
    int nf = 0, o = 0;
    for (int i = 0; i < m;) {
        int j = i;
        fv_nodes[nf] = kn[i].key; fv_off[nf] = o;
        for (j = i; j < m && kn[j].key == kn[i].key; ++j) fv_feats[o++] = kn[j].idx;
        ++nf; i = j;
    }
    fv_off[nf] = o; *n_fv = nf;
    free(kw); free(kn);
    return nb;
}

/* ------------------------------------------------------------------ stereo
 * Frame::ComputeStereoMatches, src/Frame.cc:849-1038 of the reference (this fork: minD = 0, maxD = 200, row band
 * r = 1.2 * scale, patches normalised by their centre pixel, vDescIndex with its `bestIdxR != 0` quirk).
 * Float operations are written one per statement in the reference's order; compile with -ffp-contract=off. */
int orbo_stereo_matches(int nL, const orbo_keypoint *kL, const uint8_t *dL, int nR, const orbo_keypoint *kR, const uint8_t *dR,
                        int nlevels, const float *scale, const float *inv_scale,
                        const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *pitch, const int *lw, const int *lh,
                        float bf, float *u_right, float *depth, int32_t *desc_index, int32_t *best_dist, int32_t *best_idx, int32_t *sad)
{
    (void)nlevels; (void)lh;
    const int thOrbDist = (100 + 50) / 2;                                   /* :856 */
    int *minr = (int *)malloc(sizeof(int) * (size_t)(nR > 0 ? nR : 1)), *maxr = (int *)malloc(sizeof(int) * (size_t)(nR > 0 ? nR : 1));
    for (int iR = 0; iR < nR; ++iR) {                                       /* :870-881 */
        const float kpY = kR[iR].y;
        const float r = 1.2f * scale[kR[iR].octave];
        maxr[iR] = (int)ceilf(kpY + r);
        minr[iR] = (int)floorf(kpY - r);
    }
    const float minD = 0, maxD = 200;                                       /* :885-886 */
    int *vd = (int *)malloc(sizeof(int) * (size_t)(nL > 0 ? nL : 1)), nv = 0;
    for (int iL = 0; iL < nL; ++iL) {
        u_right[iL] = -1.0f; depth[iL] = -1.0f; desc_index[iL] = -1;
        if (best_dist) best_dist[iL] = 100;
        if (best_idx) best_idx[iL] = 0;
        if (sad) sad[iL] = -1;
    }
    for (int iL = 0; iL < nL; ++iL) {
        const int levelL = kL[iL].octave;
        const float vL = kL[iL].y, uL = kL[iL].x;
        const int row = (int)vL;                                            /* vRowIndices[vL], :902 */
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100, bestIdxR = 0, any = 0;
        for (int iR = 0; iR < nR; ++iR) {                                   /* candidates of the row, ascending iR */
            if (row < minr[iR] || row > maxr[iR]) continue;
            any = 1;
            if (kR[iR].octave < levelL - 1 || kR[iR].octave > levelL + 1) continue;
            const float uR = kR[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orbo_hamming256(dL + (size_t)iL * 32, dR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (!any) continue;                                                 /* vCandidates.empty(), :904 */
        if (best_dist) best_dist[iL] = bestDist;
        if (best_idx) best_idx[iL] = bestIdxR;
        if (bestIdxR != 0) desc_index[iL] = bestIdxR;                        /* :948 */
        if (bestDist < thOrbDist) {                                         /* :954 */
            const float uR0 = kR[bestIdxR].x;
            const float scaleFactor = inv_scale[levelL];
            const float scaleduL = roundf(kL[iL].x * scaleFactor);
            const float scaledvL = roundf(kL[iL].y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= (float)lw[levelL]) continue;           /* :975 */
            const uint8_t *IL = pyrL[levelL], *IR = pyrR[levelL];
            const int p = pitch[levelL], y0 = (int)scaledvL - w, xl = (int)scaleduL - w;
            const int cl = IL[(size_t)(y0 + w) * p + xl + w];
            int bestD = INT_MAX, bestincR = 0;
            float vDists[11];
            for (int incR = -L; incR <= L; ++incR) {
                const int xr = (int)scaleduR0 + incR - w;
                const int cr = IR[(size_t)(y0 + w) * p + xr + w];
                int acc = 0;                                                /* L1 norm of integer-valued floats: exact */
                for (int yy = 0; yy < 11; ++yy)
                    for (int xx = 0; xx < 11; ++xx)
                        acc += abs((IL[(size_t)(y0 + yy) * p + xl + xx] - cl) - (IR[(size_t)(y0 + yy) * p + xr + xx] - cr));
                const float dist = (float)acc;
                if (dist < (float)bestD) { bestD = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float num = dist1 - dist3;
            const float den = 2.0f * (dist1 + dist3 - 2.0f * dist2);
            const float deltaR = num / den;
            if (deltaR < -1 || deltaR > 1) continue;
            float t = scaleduR0 + (float)bestincR;
            t = t + deltaR;
            float bestuR = scale[levelL] * t;
            float disparity = uL - bestuR;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = bf / disparity;
                u_right[iL] = bestuR;
                if (sad) sad[iL] = bestD;
                vd[nv++] = bestD;
            }
        }
    }
    int kept = -1;
    if (nv > 0) {                                                           /* :1023-1037; the reference has UB when nv == 0 */
        int *sorted = (int *)malloc(sizeof(int) * (size_t)nv);
        memcpy(sorted, vd, sizeof(int) * (size_t)nv);
        for (int i = 1; i < nv; ++i) { const int v = sorted[i]; int j = i - 1; while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; } sorted[j + 1] = v; }
        const float median = (float)sorted[nv / 2];
        const float thDist = 1.5f * 1.4f * median;
        free(sorted);
        kept = 0;
        /* the reference walks the sorted list from the back and stops at the first SAD < thDist: every pushed keypoint
         * with SAD >= thDist is invalidated.  Pushed <=> depth was set (bf / disparity > 0), in ascending iL. */
        int k = 0;
        for (int iL = 0; iL < nL; ++iL) {
            if (depth[iL] < 0) continue;
            const int v = vd[k++];
            if (!((float)v < thDist)) { u_right[iL] = -1; depth[iL] = -1; desc_index[iL] = -1; }
            else ++kept;
        }
    }
    free(minr); free(maxr); free(vd);
    return kept;
}

/* ----------------------------------------------------------- CPU baseline */

typedef struct {
    int nfeatures, nlevels, ini_th, min_th; float sf;
    const uint8_t *frames; int nframes, w, h;
    int *next; pthread_mutex_t *mu; long kps;
} many_job;

static void *many_worker(void *p)
{
    many_job *j = (many_job *)p;
    orbo_extractor *e = orbo_create(j->nfeatures, j->sf, j->nlevels, j->ini_th, j->min_th);
    const int cap = j->nfeatures + 4 * j->nlevels + 64;
    orbo_keypoint *k = (orbo_keypoint *)malloc(sizeof(orbo_keypoint) * (size_t)cap);
    uint8_t *d = (uint8_t *)malloc(32 * (size_t)cap);
    for (;;) {
        pthread_mutex_lock(j->mu);
        const int f = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (f >= j->nframes) break;
        const int n = orbo_extract(e, j->frames + (size_t)f * j->w * j->h, j->w, j->h, j->w, k, d, cap);
        j->kps += n > 0 ? n : 0;
    }
    free(k); free(d); orbo_destroy(e);
    return NULL;
}

double orbo_extract_many(int nfeatures, float sf, int nlevels, int ini_th, int min_th,
                         const uint8_t *frames, int nframes, int w, int h, int threads, long *total_kps)
{
    if (threads < 1) threads = 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    many_job *jobs = (many_job *)malloc(sizeof(many_job) * (size_t)threads);
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < threads; ++k) {
        many_job j = {nfeatures, nlevels, ini_th, min_th, sf, frames, nframes, w, h, &next, &mu, 0};
        jobs[k] = j;
        pthread_create(&t[k], NULL, many_worker, &jobs[k]);
    }
    long kps = 0;
    for (int k = 0; k < threads; ++k) { pthread_join(t[k], NULL); kps += jobs[k].kps; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_kps) *total_kps = kps;
    free(t); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
