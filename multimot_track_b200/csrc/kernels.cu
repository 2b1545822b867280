// multimot_track_b200/csrc/kernels.cu -- hand-written sm_100a kernels of the ORB front end.
//
// Stage                         reference (src/ORBextractor.cc)       kernel
//   pyramid (INTER_LINEAR)      ComputePyramid :1111-1136             k_resize
//   per-cell FAST-9/16 + NMS    ComputeKeyPointsOctTree :765-829      k_fast
//   octree distribution         DistributeOctTree :539-763            k_octree
//   IC_Angle + rBRIEF           :77-147, :472-479, :1034-1041         k_orient_desc
//   7x7 sigma=2 blur            :1089-1090                            k_blur
//   border (mvImagePyramid)     :1126-1132                            k_pad_reflect101
//   Hamming best/second-best    src/ORBmatcher.cc:574-605,2279-2295   k_match_partial / k_match_merge
//
// Integer stages are bit-exact restatements of the OpenCV fixed-point arithmetic;
// float stages use explicit round-to-nearest intrinsics so nothing is contracted
// into an FMA.
#include "kernels.cuh"

#include <cstdio>

namespace orbx {

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ const uint8_t *level_ptr(const DevParams *P, const Src0 &s0, int frame, int level, int *pitch)
{
    if (level == 0) { *pitch = s0.pitch; return s0.ptr + (long long)frame * s0.frame_stride; }
    *pitch = P->lv[level].pitch;
    return P->pyr + (long long)frame * P->pyr_frame_bytes + P->lv[level].img_off;
}

__device__ __forceinline__ int reflect101(int p, int len)
{
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------ pyramid
// cv::resize INTER_LINEAR, 8UC1: H = S[s0]*a0 + S[s1]*a1 (11-bit weights), then
// out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.  The weight tables are
// built on the host with OpenCV's float arithmetic (host_tables.cpp).

__global__ void __launch_bounds__(256) k_resize(const DevParams *__restrict__ P, Src0 s0, int level)
{
    const LevelGeom &D = P->lv[level];
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int frame = blockIdx.z;
    if (x4 >= D.w || y >= D.h) return;
    int sp;
    const uint8_t *S = level_ptr(P, s0, frame, level - 1, &sp);
    uint8_t *dst = P->pyr + (long long)frame * P->pyr_frame_bytes + D.img_off;
    const ResizeTab ty = P->ytab[P->ytab_off[level] + y];
    const uint8_t *S0 = S + (long long)ty.s0 * sp, *S1 = S + (long long)ty.s1 * sp;
    const ResizeTab *tx = P->xtab + P->xtab_off[level] + x4;       // padded to a multiple of 4 entries
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ResizeTab t = tx[k];
        const int h0 = S0[t.s0] * t.c0 + S0[t.s1] * t.c1;
        const int h1 = S1[t.s0] * t.c0 + S1[t.s1] * t.c1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out |= (uint32_t)v << (8 * k);
    }
    *reinterpret_cast<uint32_t *>(dst + (long long)y * D.pitch + x4) = out;   // pitch % 64 == 0: padding absorbs the tail
}

cudaError_t launch_pyramid(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls)
{
    for (int l = 1; l < hP.nlevels; ++l) {
        const LevelGeom &D = hP.lv[l];
        dim3 block(32, 8), grid((D.w + 127) / 128, (D.h + 7) / 8, nframes);
        k_resize<<<grid, block, 0, st>>>(dP, s0, l);
        ls->launches++;
    }
    return cudaGetLastError();
}

// --------------------------------------------------------------------- blur
// cv::GaussianBlur 7x7 sigma 2 on 8-bit: 8.8 fixed-point kernel [18,34,48,56,48,34,18],
// horizontal pass to u16, vertical pass to u32, one rounding (+32768)>>16; reflect-101.

__global__ void __launch_bounds__(256) k_blur(const DevParams *__restrict__ P, Src0 s0)
{
    __shared__ uint8_t sin_[kBlurTileH + 6][kBlurTileW + 8];
    __shared__ uint16_t sh[kBlurTileH + 6][kBlurTileW];
    const uint32_t wk = P->blur_work[blockIdx.x];
    const int level = wk >> 24, ty0 = ((wk >> 12) & 0xfff) * kBlurTileH, tx0 = (wk & 0xfff) * kBlurTileW;
    const int frame = blockIdx.y;
    const LevelGeom &G = P->lv[level];
    int sp;
    const uint8_t *S = level_ptr(P, s0, frame, level, &sp);
    const int tid = threadIdx.x;
    for (int i = tid; i < (kBlurTileH + 6) * (kBlurTileW + 6); i += 256) {
        const int r = i / (kBlurTileW + 6), c = i - r * (kBlurTileW + 6);
        const int gy = reflect101(ty0 + r - 3, G.h), gx = reflect101(tx0 + c - 3, G.w);
        sin_[r][c] = S[(long long)gy * sp + gx];
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTileH + 6) * kBlurTileW; i += 256) {
        const int r = i / kBlurTileW, c = i - r * kBlurTileW;
        const uint8_t *p = &sin_[r][c];
        sh[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    const int r = tid >> 4, c4 = (tid & 15) * 4;
    const int gy = ty0 + r, gx = tx0 + c4;
    if (gy < G.h && gx < G.w) {
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c4 + k;
            const uint32_t acc = 18u * (sh[r][c] + sh[r + 6][c]) + 34u * (sh[r + 1][c] + sh[r + 5][c]) +
                                 48u * (sh[r + 2][c] + sh[r + 4][c]) + 56u * sh[r + 3][c];
            out |= ((acc + 32768u) >> 16) << (8 * k);
        }
        uint8_t *dst = P->blur + (long long)frame * P->pyr_frame_bytes + G.img_off;
        *reinterpret_cast<uint32_t *>(dst + (long long)gy * G.pitch + gx) = out;
    }
}

cudaError_t launch_blur(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls)
{
    dim3 grid(hP.n_blur_work, nframes);
    k_blur<<<grid, 256, 0, st>>>(dP, s0);
    ls->launches++;
    return cudaGetLastError();
}

// --------------------------------------------------------------------- FAST
// One 128-thread CTA per 30-pixel cell (the reference calls cv::FAST once per cell,
// :809-815).  The cell and its 3-pixel ring are staged in shared memory as 16-bit
// lanes, two horizontally adjacent pixels per 32-bit word, so every step works on a
// pixel PAIR with the packed-halfword integer pipe (VIADD.16x2 / VIMNMX3.S16x2):
//   A. compass pre-test (any 9-arc holds two adjacent compass points of one
//      polarity) over all pairs; pairs that can hold a bright / dark corner are
//      ballot-compacted into two lists;
//   B. exact score of the listed pairs, one polarity per list:
//         bright = max_k min(e_k..e_k+8) - 1,  dark = -min_k max(e_k..e_k+8) - 1,
//      e_k = ring_k - centre  (OpenCV cornerScore<16>: the largest threshold at which
//      the pixel is still a corner; corner at t <=> score >= t);
//   C. non-max suppression inside the cell's own rectangle (outside counts as 0,
//      like cv::FAST on the cell sub-image) and emission.
// A cell with no survivor at iniThFAST is redone at minThFAST (:812-816).

template <int CELL>       // largest detection-area side this instantiation handles
struct FastCfg {
    static constexpr int NPMAX = (CELL + 1) / 2;            // pixel pairs per row
    static constexpr int TPW = NPMAX + 5;                   // tile pitch in words: pairs -2 .. NPMAX+2
    static constexpr int TH = CELL + 6;
    static constexpr int SP = (CELL + 2 + 3) & ~3;          // score-map pitch (bytes)
    static constexpr int SH = CELL + 2;
    static constexpr int LN = NPMAX * CELL;                 // worst-case pairs per polarity list
    static constexpr int CN = CELL * CELL;                  // worst-case corners
    static constexpr int SMEM = TH * TPW * 4 + SH * SP + 2 * LN * 2 + CN * 2;
    static constexpr int THREADS = 128;
};

__device__ __forceinline__ unsigned add16x2(unsigned a, unsigned b) { return __vadd2(a, b); }

// ring pair at column offset ox (pixels) relative to the pair's first pixel; row pointer given
template <int OX>
__device__ __forceinline__ unsigned ring_pair(const uint32_t *row, int j)
{
    if (OX % 2 == 0) return row[j + OX / 2];
    const int a = j + (OX - 1) / 2;                 // floor division for odd OX of either sign
    return __funnelshift_r(row[a], row[a + 1], 16);
}

template <int CELL>
__global__ void __launch_bounds__(FastCfg<CELL>::THREADS) k_fast(const DevParams *__restrict__ P, Src0 s0, int work_off)
{
    using C = FastCfg<CELL>;
    extern __shared__ __align__(16) uint8_t fast_smem[];
    __shared__ int s_nb, s_nd, s_nc, s_emitted;
    uint32_t *tile = reinterpret_cast<uint32_t *>(fast_smem);                   // [TH][TPW] pixel pairs as 16x2
    uint8_t *smap = fast_smem + C::TH * C::TPW * 4;                             // [SH][SP] corner scores
    uint16_t *listB = reinterpret_cast<uint16_t *>(smap + C::SH * C::SP);       // pairs to score, bright polarity
    uint16_t *listD = listB + C::LN;
    uint16_t *corners = listD + C::LN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, frame = blockIdx.y;
    const uint32_t wk = P->fast_work[work_off + blockIdx.x];
    const int level = wk >> 24, ci = (wk >> 12) & 0xfff, cj = wk & 0xfff;
    const LevelGeom &G = P->lv[level];
    const int x0 = kEdge + cj * G.w_cell, x1 = min(x0 + G.w_cell, G.x_end);
    const int y0 = kEdge + ci * G.h_cell, y1 = min(y0 + G.h_cell, G.y_end);
    const int dw = x1 - x0, dh = y1 - y0, np = (dw + 1) >> 1;

    // ---- stage rows [y0-3, y1+3) x pixels [x0-4, x0-4+2*TPW) as 16-bit lanes (tile pixel u = x - x0 + 4)
    {
        int sp;
        const uint8_t *img = level_ptr(P, s0, frame, level, &sp);
        const int xb = (x0 - 4) & ~3;
        const int xlast = min(x0 - 4 + 2 * (np + 5), G.w);                      // exclusive; stays inside the row
        const int nwords = (xlast - xb + 3) >> 2, nrows = dh + 6;
        uint16_t *t16 = reinterpret_cast<uint16_t *>(tile);
        for (int i = tid; i < nrows * nwords; i += C::THREADS) {
            const int r = i / nwords, c = i - r * nwords;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(img + (long long)(y0 - 3 + r) * sp + xb) + c);
            const int u = xb + 4 * c - (x0 - 4);
            uint16_t *d = t16 + r * (2 * C::TPW) + u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((unsigned)(u + k) < (unsigned)(2 * C::TPW)) d[k] = (uint16_t)((v >> (8 * k)) & 0xff);
        }
        for (int i = tid; i < (C::SH * C::SP) / 4; i += C::THREADS) reinterpret_cast<uint32_t *>(smap)[i] = 0;
        if (tid == 0) { s_nb = 0; s_nd = 0; s_nc = 0; s_emitted = 0; }
    }
    __syncthreads();

    uint32_t *cand = P->cand + (long long)frame * P->cand_frame_elems + G.cand_off;
    uint32_t *cnt = P->cand_count + frame * P->nlevels + level;
    const int npairs = np * dh;
    const float inv_np = 1.0f / (float)np;
    const unsigned lt = lanemask_lt();

    int th = P->ini_th;
    for (int pass = 0; pass < 2; ++pass) {
        // ---- A: compass pre-test on every pair
        const unsigned tb = 0x00010001u * (unsigned)(th + 1), td = 0x00010001u * (unsigned)th;
        for (int base = warp * 32; base < npairs; base += C::THREADS) {
            const int i = base + lane;
            bool wb = false, wd = false;
            int dy = 0, p = 0;
            if (i < npairs) {
                dy = (int)(((float)i + 0.5f) * inv_np);
                p = i - dy * np;
                const uint32_t *row = tile + (dy + 3) * C::TPW;
                const int j = p + 2;
                const unsigned c = row[j];
                const unsigned negc = __vneg2(c);
                const unsigned nb = __vsub2(negc, tb);                      // -(c + t + 1)
                const unsigned nd = add16x2(negc, td);                      // -(c - t)
                const unsigned n = row[j - 3 * C::TPW], s = row[j + 3 * C::TPW];
                const unsigned e = ring_pair<3>(row, j), w = ring_pair<-3>(row, j);
                // sign bit clear in ring+nb  <=> ring > c+t ; sign bit set in ring+nd <=> ring < c-t
                const unsigned bn = add16x2(n, nb), bs = add16x2(s, nb), be = add16x2(e, nb), bw = add16x2(w, nb);
                const unsigned dn = add16x2(n, nd), ds = add16x2(s, nd), de = add16x2(e, nd), dwk = add16x2(w, nd);
                const unsigned bright = ~((bn & bs) | (be & bw)) & 0x80008000u;
                const unsigned dark = (dn | ds) & (de | dwk) & 0x80008000u;
                wb = bright != 0; wd = dark != 0;
            }
            const unsigned mb = __ballot_sync(0xffffffffu, wb), md = __ballot_sync(0xffffffffu, wd);
            int ob = 0, od = 0;
            if (lane == 0) { if (mb) ob = atomicAdd(&s_nb, __popc(mb)); if (md) od = atomicAdd(&s_nd, __popc(md)); }
            ob = __shfl_sync(0xffffffffu, ob, 0); od = __shfl_sync(0xffffffffu, od, 0);
            if (wb) listB[ob + __popc(mb & lt)] = (uint16_t)(dy << 8 | p);
            if (wd) listD[od + __popc(md & lt)] = (uint16_t)(dy << 8 | p);
        }
        __syncthreads();
        // ---- B: exact one-sided scores of the listed pairs
        const int nb_ = s_nb, nd_ = s_nd;
        for (int base = warp * 32; base < nb_ + nd_; base += C::THREADS) {
            const int i = base + lane;
            bool k0 = false, k1 = false;
            int dy = 0, p = 0, s0v = 0, s1v = 0;
            if (i < nb_ + nd_) {
                const bool is_b = i < nb_;
                const int ent = is_b ? listB[i] : listD[i - nb_];
                dy = ent >> 8; p = ent & 0xff;
                const uint32_t *row = tile + (dy + 3) * C::TPW;
                const int j = p + 2;
                const unsigned negc = __vneg2(row[j]);
                unsigned e[16];
                {
                    const uint32_t *r3 = row + 3 * C::TPW, *r2 = row + 2 * C::TPW, *r1 = row + C::TPW;
                    const uint32_t *m1 = row - C::TPW, *m2 = row - 2 * C::TPW, *m3 = row - 3 * C::TPW;
                    e[0] = r3[j];               e[1] = ring_pair<1>(r3, j);   e[15] = ring_pair<-1>(r3, j);
                    e[2] = r2[j + 1];           e[14] = r2[j - 1];
                    e[3] = ring_pair<3>(r1, j); e[13] = ring_pair<-3>(r1, j);
                    e[4] = ring_pair<3>(row, j); e[12] = ring_pair<-3>(row, j);
                    e[5] = ring_pair<3>(m1, j); e[11] = ring_pair<-3>(m1, j);
                    e[6] = m2[j + 1];           e[10] = m2[j - 1];
                    e[8] = m3[j];               e[7] = ring_pair<1>(m3, j);   e[9] = ring_pair<-1>(m3, j);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = add16x2(e[k], negc);
                unsigned res;
                if (is_b) {                                        // max over arcs of min9
                    unsigned t3[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) t3[k] = __vimin3_s16x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
                    unsigned a[6];
#pragma unroll
                    for (int k = 0; k < 16; ++k) e[k] = __vimin3_s16x2(t3[k], t3[(k + 3) & 15], t3[(k + 6) & 15]);
#pragma unroll
                    for (int k = 0; k < 5; ++k) a[k] = __vimax3_s16x2(e[3 * k], e[3 * k + 1], e[3 * k + 2]);
                    a[5] = e[15];
                    res = __vimax3_s16x2(__vimax3_s16x2(a[0], a[1], a[2]), __vimax3_s16x2(a[3], a[4], a[5]), a[5]);
                } else {                                           // -(min over arcs of max9)
                    unsigned t3[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) t3[k] = __vimax3_s16x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
                    unsigned a[6];
#pragma unroll
                    for (int k = 0; k < 16; ++k) e[k] = __vimax3_s16x2(t3[k], t3[(k + 3) & 15], t3[(k + 6) & 15]);
#pragma unroll
                    for (int k = 0; k < 5; ++k) a[k] = __vimin3_s16x2(e[3 * k], e[3 * k + 1], e[3 * k + 2]);
                    a[5] = e[15];
                    res = __vneg2(__vimin3_s16x2(__vimin3_s16x2(a[0], a[1], a[2]), __vimin3_s16x2(a[3], a[4], a[5]), a[5]));
                }
                s0v = (int)(short)(res & 0xffff) - 1;
                s1v = (int)(short)(res >> 16) - 1;
                k0 = s0v >= th;
                k1 = s1v >= th && 2 * p + 1 < dw;
                if (k0) smap[(dy + 1) * C::SP + 2 * p + 1] = (uint8_t)s0v;
                if (k1) smap[(dy + 1) * C::SP + 2 * p + 2] = (uint8_t)s1v;
            }
            const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
            if (m0 | m1) {
                int o = 0;
                if (lane == 0) o = atomicAdd(&s_nc, __popc(m0) + __popc(m1));
                o = __shfl_sync(0xffffffffu, o, 0);
                if (k0) corners[o + __popc(m0 & lt)] = (uint16_t)(dy << 8 | (2 * p));
                if (k1) corners[o + __popc(m0) + __popc(m1 & lt)] = (uint16_t)(dy << 8 | (2 * p + 1));
            }
        }
        __syncthreads();
        // ---- C: non-max suppression + emission
        const int nc = s_nc;
        for (int base = warp * 32; base < nc; base += C::THREADS) {
            const int i = base + lane;
            bool keep = false;
            int dy = 0, dx = 0, sc = 0;
            if (i < nc) {
                const int ent = corners[i];
                dy = ent >> 8; dx = ent & 0xff;
                const uint8_t *q = smap + (dy + 1) * C::SP + dx + 1;
                sc = q[0];
                keep = sc > q[-1] && sc > q[1] && sc > q[-C::SP - 1] && sc > q[-C::SP] && sc > q[-C::SP + 1] &&
                       sc > q[C::SP - 1] && sc > q[C::SP] && sc > q[C::SP + 1];
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                int slot = 0;
                if (lane == 0) { slot = (int)atomicAdd(cnt, (unsigned)__popc(m)); s_emitted = 1; }
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (keep) {
                    const uint32_t xr = (uint32_t)(x0 + dx - kMinBorder), yr = (uint32_t)(y0 + dy - kMinBorder);
                    const int at = slot + __popc(m & lt);
                    if (at < G.cand_cap) cand[at] = xr | yr << 12 | (uint32_t)sc << 24;
                }
            }
        }
        __syncthreads();
        if (s_emitted || pass == 1 || P->min_th == th) break;
        // nothing survived at iniThFAST: clear and redo the cell at minThFAST
        __syncthreads();
        th = P->min_th;
        for (int i = tid; i < (C::SH * C::SP) / 4; i += C::THREADS) reinterpret_cast<uint32_t *>(smap)[i] = 0;
        if (tid == 0) { s_nb = 0; s_nd = 0; s_nc = 0; }
        __syncthreads();
    }
}

cudaError_t launch_fast(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, int n_small, cudaStream_t st, LaunchStats *ls)
{
    if (n_small > 0) {
        using C = FastCfg<38>;
        cudaFuncSetAttribute(k_fast<38>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        k_fast<38><<<dim3(n_small, nframes), C::THREADS, C::SMEM, st>>>(dP, s0, 0);
        ls->launches++;
    }
    if (hP.n_fast_work > n_small) {
        using C = FastCfg<64>;
        cudaFuncSetAttribute(k_fast<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        k_fast<64><<<dim3(hP.n_fast_work - n_small, nframes), C::THREADS, C::SMEM, st>>>(dP, s0, n_small);
        ls->launches++;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------- octree
// DistributeOctTree (:539-763) for one (frame, level) per CTA.
//
// Every candidate gets a path code: its initial node, then one 2-bit child index
// (n1=UL 0, n2=UR 1, n3=BL 2, n4=BR 3, DivideNode :481-537) per split down to
// single pixels.  After sorting by that code every tree node is a contiguous key
// range, so splitting a node is three binary searches and the whole subdivision
// works on <= 2N+3 (lo, hi, depth) records in shared memory.
//
// The std::list of the reference is kept as an array in REVERSE list order
// (push_front == append).  A sweep (:606-665) splits every multi-key node, visiting
// in list order == reverse array order; the careful phase (:676-735) visits by
// (key count, creation order) descending and cuts off where the node count reaches N
// (prefix sum of children-1).  Creation order equals array position, which is the
// canonical replacement for the reference's pointer-value tie-break.
// The winner of a node is the key with the greatest response, first in the reference's
// candidate order (cell row, cell column, y, x) (:747-757).

constexpr int kOctThreads = 1024;
constexpr int kOctIPT = 8;                 // node records per thread in the block-wide passes
typedef unsigned long long u64;

__device__ __forceinline__ int block_scan_incl(int v, int *warp_sums, int *total)
{
    // inclusive scan over threadIdx.x order; *total = block sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();                       // protects warp_sums reuse
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    *total = warp_sums[31];
    return x + (warp > 0 ? warp_sums[warp - 1] : 0);
}

__device__ __forceinline__ void bitonic_sort(u64 *v, int n_pow2, bool descending)
{
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));       // lower index of the pair
                const int p = i | j;
                const u64 a = v[i], b = v[p];
                const bool up = ((i & k) == 0) != descending;
                if ((a > b) == up) { v[i] = b; v[p] = a; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int lower_child(const u64 *keys, int lo, int hi, int shift, unsigned c)
{
    // first index in [lo,hi) whose 2-bit child field at `shift` (of the high word) is >= c
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if ((((unsigned)(keys[m] >> 32) >> shift) & 3u) >= c) hi = m; else lo = m + 1;
    }
    return lo;
}

size_t octree_smem_bytes(int max_node_cap, int max_feat, int *key_cap)
{
    int sp2 = 1;
    while (sp2 < max_feat + 3) sp2 <<= 1;
    const size_t node_bytes = (size_t)max_node_cap * 12 + (size_t)max_node_cap * 4 /*eidx*/ + (size_t)sp2 * 8 /*skey*/ + 256;
    const size_t budget = 200 * 1024;
    int kc = 1024;
    while ((size_t)(kc * 2) * 8 + node_bytes <= budget) kc *= 2;
    if (key_cap) *key_cap = kc;
    return node_bytes + (size_t)kc * 8;
}

__global__ void __launch_bounds__(kOctThreads, 1)
k_octree(const DevParams *__restrict__ P, int node_cap, int skey_cap, int key_cap)
{
    extern __shared__ __align__(16) uint8_t oct_smem[];
    __shared__ int warp_sums[32];
    __shared__ int s_m, s_added;

    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom &G = P->lv[level];
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = (int)min(P->cand_count[frame * P->nlevels + level], (unsigned)G.cand_cap);
    const int N = G.n_feat, D = G.depth;

    u64 *skey = reinterpret_cast<u64 *>(oct_smem);                 // careful-phase sort buffer
    int *nlo = reinterpret_cast<int *>(skey + skey_cap);
    int *nhi = nlo + node_cap;
    int *ndep = nhi + node_cap;
    int *eidx = ndep + node_cap;
    u64 *keys = reinterpret_cast<u64 *>(eidx + node_cap);            // 4 int arrays = 16*node_cap bytes: stays 8-aligned

    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    if (n2 > key_cap)                                               // level too dense for shared memory: L2-resident scratch
        keys = P->sort_scratch + ((long long)frame * P->cand_frame_elems + G.cand_off) * 2;
    const uint32_t *cand = P->cand + (long long)frame * P->cand_frame_elems + G.cand_off;
    uint32_t *stage = P->kp_stage + (long long)frame * P->kp_frame_cap + G.kp_off;
    uint32_t *kp_count = P->kp_count + frame * P->nlevels + level;

    if (n == 0) { if (tid == 0) *kp_count = 0; return; }

    // ---- path codes (:566-570 root assignment, :481-534 child tests)
    for (int i = tid; i < n2; i += kOctThreads) {
        u64 kv = ~0ull;
        if (i < n) {
            const uint32_t c = cand[i];
            const int x = c & 0xfff, y = (c >> 12) & 0xfff;
            int r = (int)__fdiv_rn((float)x, G.h_x);
            r = min(r, G.n_ini - 1);
            int ulx = G.root_ul[r], brx = G.root_br[r], uly = 0, bry = G.region_h;
            uint32_t code = (uint32_t)r;
            for (int d = 0; d < D; ++d) {
                const int midx = ulx + ((brx - ulx + 1) >> 1), midy = uly + ((bry - uly + 1) >> 1);
                const bool right = x >= midx, down = y >= midy;
                if (right) ulx = midx; else brx = midx;
                if (down) uly = midy; else bry = midy;
                code = code << 2 | (uint32_t)down << 1 | (uint32_t)right;
            }
            kv = (u64)code << 32 | c;
        }
        keys[i] = kv;
    }
    __syncthreads();
    bitonic_sort(keys, n2, false);

    // ---- roots, reverse list order (:553-585)
    if (tid == 0) {
        int na = 0;
        for (int r = G.n_ini - 1; r >= 0; --r) {
            int lo = 0, hi = n;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (((unsigned)(keys[m] >> 32) >> (2 * D)) >= (unsigned)r) hi = m; else lo = m + 1; }
            const int a = lo;
            lo = a; hi = n;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (((unsigned)(keys[m] >> 32) >> (2 * D)) >= (unsigned)(r + 1)) hi = m; else lo = m + 1; }
            if (lo > a) { nlo[na] = a; nhi[na] = lo; ndep[na] = 0; ++na; }
        }
        int ne = 0;
        for (int i = 0; i < na; ++i) if (nhi[i] - nlo[i] > 1) eidx[ne++] = i;
        s_m = na; s_added = ne;
    }
    __syncthreads();
    int size = s_m, nE = s_added;
    __syncthreads();

    bool careful = false;
    for (;;) {
        const int prev_size = size;
        // ---- visit order
        int sp2 = 1;
        if (careful) {
            while (sp2 < nE) sp2 <<= 1;
            for (int i = tid; i < sp2; i += kOctThreads)
                skey[i] = i < nE ? ((u64)(unsigned)(nhi[eidx[i]] - nlo[eidx[i]]) << 32 | (unsigned)eidx[i]) : 0ull;
            __syncthreads();
            bitonic_sort(skey, sp2, true);
        }
        // ---- pass 1: children of the parents this thread owns (visit ranks tid*IPT ..)
        int pb[kOctIPT][3];
        int pc[kOctIPT];
        int csum = 0;
#pragma unroll
        for (int k = 0; k < kOctIPT; ++k) {
            const int r = tid * kOctIPT + k;
            pc[k] = 0;
            if (r < nE) {
                const int p = careful ? (int)(skey[r] & 0xffffffffu) : eidx[nE - 1 - r];
                const int lo = nlo[p], hi = nhi[p], shift = 2 * (D - 1 - ndep[p]);
                const int b1 = lower_child(keys, lo, hi, shift, 1);
                const int b2 = lower_child(keys, b1, hi, shift, 2);
                const int b3 = lower_child(keys, b2, hi, shift, 3);
                pb[k][0] = b1; pb[k][1] = b2; pb[k][2] = b3;
                pc[k] = (b1 > lo) + (b2 > b1) + (b3 > b2) + (hi > b3);
                csum += pc[k];
            }
        }
        int total_c;
        const int incl = block_scan_incl(csum, warp_sums, &total_c);
        // ---- cutoff m (careful phase: smallest m with size + sum_{i<m}(c_i - 1) >= N)
        if (tid == 0) { s_m = nE; s_added = total_c; }
        __syncthreads();
        if (careful) {
            int run = incl - csum;                                  // children of all earlier ranks
#pragma unroll
            for (int k = 0; k < kOctIPT; ++k) {
                const int r = tid * kOctIPT + k;
                if (r < nE) {
                    const int before = size + run - r;              // size + sum_{i<r}(c_i-1)
                    run += pc[k];
                    const int after = size + run - (r + 1);
                    if (after >= N && before < N) { s_m = r + 1; s_added = run; }
                }
            }
            __syncthreads();
        }
        const int m = s_m, added = s_added;
        // ---- pass 2: append children in visit order, retire the parents
        {
            int off = incl - csum;
#pragma unroll
            for (int k = 0; k < kOctIPT; ++k) {
                const int r = tid * kOctIPT + k;
                if (r < nE && r < m) {
                    const int p = careful ? (int)(skey[r] & 0xffffffffu) : eidx[nE - 1 - r];
                    const int lo = nlo[p], hi = nhi[p], dep = ndep[p] + 1;
                    const int b[5] = {lo, pb[k][0], pb[k][1], pb[k][2], hi};
                    int w = size + off;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (b[c + 1] > b[c]) { nlo[w] = b[c]; nhi[w] = b[c + 1]; ndep[w] = dep; ++w; }
                    ndep[p] = -1;
                }
                off += pc[k];
            }
        }
        __syncthreads();
        const int n_pre = size + added;
        size = size - m + added;
        // ---- stable compaction of the array + new expandable list
        {
            int rl[kOctIPT], rh[kOctIPT], rd[kOctIPT];
            int alive = 0, multi = 0;
#pragma unroll
            for (int k = 0; k < kOctIPT; ++k) {
                const int i = tid * kOctIPT + k;
                rd[k] = -1;
                if (i < n_pre) {
                    rl[k] = nlo[i]; rh[k] = nhi[i]; rd[k] = ndep[i];
                    if (rd[k] >= 0) { ++alive; multi += (rh[k] - rl[k] > 1); }
                }
            }
            int tot;
            const int inc2 = block_scan_incl(alive | multi << 16, warp_sums, &tot);
            int wa = (inc2 & 0xffff) - alive, we = (inc2 >> 16) - multi;
#pragma unroll
            for (int k = 0; k < kOctIPT; ++k) {
                if (rd[k] >= 0) {
                    nlo[wa] = rl[k]; nhi[wa] = rh[k]; ndep[wa] = rd[k];
                    if (rh[k] - rl[k] > 1) eidx[we++] = wa;
                    ++wa;
                }
            }
            nE = tot >> 16;
            __syncthreads();
        }
        // ---- termination (:667-737)
        if (size >= N || size == prev_size) break;
        if (!careful && size + 3 * nE > N) careful = true;
    }

    // ---- winners, list order front->back == reverse array order (:740-760)
    const int n_cols = G.n_cols, w_cell = G.w_cell, h_cell = G.h_cell;
    for (int j = tid >> 5; j < size; j += kOctThreads >> 5) {
        const int nd = size - 1 - j;
        u64 best = 0;
        for (int i = nlo[nd] + lane; i < nhi[nd]; i += 32) {
            const uint32_t c = (uint32_t)keys[i];
            const u64 x = c & 0xfff, y = (c >> 12) & 0xfff, s = c >> 24;
            const u64 order = ((u64)(((int)y - 3) / h_cell * n_cols + ((int)x - 3) / w_cell) << 24) | y << 12 | x;
            const u64 v = s << 40 | (~order & 0xffffffffffull);
            best = v > best ? v : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        if (lane == 0) {
            const u64 order = ~best & 0xffffffffffull;
            stage[j] = (uint32_t)(order & 0xffffff) | (uint32_t)(best >> 40) << 24;
        }
    }
    if (tid == 0) *kp_count = (uint32_t)size;
}

cudaError_t launch_octree(const DevParams *dP, const DevParams &hP, int nframes, int max_node_cap, int max_feat, cudaStream_t st, LaunchStats *ls)
{
    int key_cap = 0;
    const size_t smem = octree_smem_bytes(max_node_cap, max_feat, &key_cap);
    int sp2 = 1;
    while (sp2 < max_feat + 3) sp2 <<= 1;
    cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(hP.nlevels, nframes);
    k_octree<<<grid, kOctThreads, smem, st>>>(dP, max_node_cap, sp2, key_cap);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------- orientation + descriptor
// One warp per keypoint.  IC_Angle (:77-104): integer moments over the radius-15
// disc (lane = column, loop over rows), fastAtan2 polynomial in non-contracted
// fp32.  computeOrbDescriptor (:108-147): lane i produces byte i from 16 rotated
// samples of the blurred level; the 32 bytes leave as two 128-bit stores.

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    constexpr float kScale = (float)(180.0 / 3.1415926535897932384626433832795);
    constexpr float p1 = 0.9997878412794807f * kScale, p3 = -0.3258083974640975f * kScale;
    constexpr float p5 = 0.1555786518463281f * kScale, p7 = -0.04432655554792128f * kScale;
    constexpr float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int kOdWarps = 8;

__global__ void __launch_bounds__(kOdWarps * 32) k_orient_desc(const DevParams *__restrict__ P, Src0 s0)
{
    __shared__ int8_t spat[1024];
    for (int i = threadIdx.x; i < 1024; i += kOdWarps * 32) spat[i] = P->pattern[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kOdWarps + warp, frame = blockIdx.y;
    const int L = P->nlevels;
    // level of this staging slot and its row in the frame's output
    int level = -1, out_idx = 0, total = 0;
    const uint32_t *counts = P->kp_count + frame * L;
    for (int l = 0; l < L; ++l) {
        const int c = min((int)counts[l], P->lv[l].kp_cap);
        const int rel = slot - P->lv[l].kp_off;
        if (rel >= 0 && rel < c) { level = l; out_idx = total + rel; }
        total += c;
    }
    if (slot == 0 && lane == 0) P->out_n[frame] = total;
    if (level < 0) return;

    const LevelGeom &G = P->lv[level];
    const uint32_t c = P->kp_stage[(long long)frame * P->kp_frame_cap + slot];
    const int cx = (int)(c & 0xfff) + kMinBorder, cy = (int)((c >> 12) & 0xfff) + kMinBorder;
    int sp;
    const uint8_t *img = level_ptr(P, s0, frame, level, &sp);
    // ---- IC_Angle
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const uint8_t *ctr = img + (long long)cy * sp + cx + u;
        const int au = abs(u);
        for (int v = -15; v <= 15; ++v) {
            if (au <= P->umax[abs(v)]) {
                const int val = ctr[(long long)v * sp];
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
    m10 = warp_sum(m10);
    m01 = warp_sum(m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered rBRIEF on the blurred level
    const uint8_t *bl = P->blur + (long long)frame * P->pyr_frame_bytes + G.img_off;
    const int bp = G.pitch;
    constexpr float kFactorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, kFactorPI);
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const uint8_t *ctr = bl + (long long)cy * bp + cx;
    const int8_t *pp = spat + lane * 32;
    unsigned byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)pp[4 * k], y0 = (float)pp[4 * k + 1], x1 = (float)pp[4 * k + 2], y1 = (float)pp[4 * k + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = ctr[r0 * bp + c0], t1 = ctr[r1 * bp + c1];
        byte |= (unsigned)(t0 < t1) << k;
    }
    // gather 32 bytes -> 8 words -> two uint4 stores
    const int q = lane & 7;
    unsigned w = __shfl_sync(0xffffffffu, byte, 4 * q) | __shfl_sync(0xffffffffu, byte, 4 * q + 1) << 8 |
                 __shfl_sync(0xffffffffu, byte, 4 * q + 2) << 16 | __shfl_sync(0xffffffffu, byte, 4 * q + 3) << 24;
    const int h4 = (lane & 1) * 4;
    uint4 v;
    v.x = __shfl_sync(0xffffffffu, w, h4); v.y = __shfl_sync(0xffffffffu, w, h4 + 1);
    v.z = __shfl_sync(0xffffffffu, w, h4 + 2); v.w = __shfl_sync(0xffffffffu, w, h4 + 3);
    const long long row = (long long)frame * P->kp_frame_cap + out_idx;
    if (lane < 2) reinterpret_cast<uint4 *>(P->out_desc + row * 32)[lane] = v;

    // ---- cv::KeyPoint record (:837-847, :1098-1104)
    if (lane < 7) {
        float f;
        switch (lane) {
        case 0: f = __fmul_rn((float)cx, G.scale); break;
        case 1: f = __fmul_rn((float)cy, G.scale); break;
        case 2: f = G.kp_size; break;
        case 3: f = angle; break;
        case 4: f = (float)(c >> 24); break;
        case 5: f = __int_as_float(level); break;
        default: f = __int_as_float(-1); break;
        }
        reinterpret_cast<float *>(P->out_kps + row)[lane] = f;
    }
}

cudaError_t launch_orient_desc(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls)
{
    dim3 grid((hP.kp_frame_cap + kOdWarps - 1) / kOdWarps, nframes);
    k_orient_desc<<<grid, kOdWarps * 32, 0, st>>>(dP, s0);
    ls->launches++;
    return cudaGetLastError();
}

// ----------------------------------------------------- mvImagePyramid border

__global__ void k_pad_reflect101(const uint8_t *__restrict__ src, int w, int h, int pitch, uint8_t *__restrict__ dst, int dst_pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w + 2 * kEdge || y >= h + 2 * kEdge) return;
    dst[(long long)y * dst_pitch + x] = src[(long long)reflect101(y - kEdge, h) * pitch + reflect101(x - kEdge, w)];
}

cudaError_t launch_pad_level(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dst_pitch, cudaStream_t st, LaunchStats *ls)
{
    dim3 block(32, 8), grid((w + 2 * kEdge + 31) / 32, (h + 2 * kEdge + 7) / 8);
    k_pad_reflect101<<<grid, block, 0, st>>>(src, w, h, pitch, dst, dst_pitch);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ matcher
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2279-2295) is a 256-bit Hamming
// distance == 8 x __popc.  The scan of :574-605 in index order is equivalent to
// (minimum, lowest index attaining it, second smallest of the multiset), which
// merges associatively across chunks of B processed by different CTAs.

constexpr int kMatchThreads = 128;
constexpr int kMatchTile = 128;

__device__ __forceinline__ void match_update(int d, int j, int &b1, int &bi, int &b2)
{
    if (d < b1) { b2 = b1; b1 = d; bi = j; }
    else if (d < b2) b2 = d;
}

__global__ void __launch_bounds__(kMatchThreads)
k_match_partial(const uint32_t *__restrict__ A, int nA, const uint32_t *__restrict__ B, int nB, int chunk, int4 *__restrict__ partial)
{
    __shared__ uint4 sB[kMatchTile * 2];
    const int i = blockIdx.x * kMatchThreads + threadIdx.x;
    const int j0 = blockIdx.y * chunk, j1 = min(j0 + chunk, nB);
    uint32_t a[8];
    if (i < nA) {
        const uint4 lo = reinterpret_cast<const uint4 *>(A)[2 * (long long)i], hi = reinterpret_cast<const uint4 *>(A)[2 * (long long)i + 1];
        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 0;
    }
    int b1 = 256, b2 = 256, bi = -1;
    for (int t0 = j0; t0 < j1; t0 += kMatchTile) {
        const int cnt = min(kMatchTile, j1 - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 2; k += kMatchThreads) sB[k] = reinterpret_cast<const uint4 *>(B)[2 * (long long)t0 + k];
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const uint4 lo = sB[2 * j], hi = sB[2 * j + 1];
            const int d = __popc(a[0] ^ lo.x) + __popc(a[1] ^ lo.y) + __popc(a[2] ^ lo.z) + __popc(a[3] ^ lo.w) +
                          __popc(a[4] ^ hi.x) + __popc(a[5] ^ hi.y) + __popc(a[6] ^ hi.z) + __popc(a[7] ^ hi.w);
            match_update(d, t0 + j, b1, bi, b2);
        }
    }
    if (i < nA) partial[(long long)blockIdx.y * nA + i] = make_int4(b1, bi, b2, 0);
}

__global__ void k_match_merge(const int4 *__restrict__ partial, int nA, int nchunks, int th, float ratio,
                              int32_t *__restrict__ idx, int32_t *__restrict__ d1, int32_t *__restrict__ d2,
                              uint8_t *__restrict__ accept, int *__restrict__ naccept)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < nA) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int c = 0; c < nchunks; ++c) {              // chunks in ascending index order
            const int4 p = partial[(long long)c * nA + i];
            if (p.x < b1) { b2 = min(b1, p.z); b1 = p.x; bi = p.y; }
            else b2 = min(b2, p.x);
        }
        idx[i] = bi; d1[i] = b1; d2[i] = b2;
        ok = bi >= 0 && b1 <= th && (float)b1 < __fmul_rn(ratio, (float)b2);
        if (accept) accept[i] = ok;
    }
    if (naccept) {
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(naccept, __popc(m));
    }
}

int match_chunks(int nA, int nB)
{
    if (nB <= 0) return 1;
    const int row_blocks = (nA + kMatchThreads - 1) / kMatchThreads;
    int want = (148 * 4 + row_blocks - 1) / (row_blocks > 0 ? row_blocks : 1);     // ~4 CTAs per SM in flight
    const int max_chunks = (nB + kMatchTile - 1) / kMatchTile;
    if (want < 1) want = 1;
    if (want > max_chunks) want = max_chunks;
    return want;
}

cudaError_t launch_match(const uint32_t *dA, int nA, const uint32_t *dB, int nB, int th, float ratio,
                         int32_t *d_idx, int32_t *d_d1, int32_t *d_d2, uint8_t *d_accept, int *d_naccept,
                         int4 *d_partial, int nchunks, cudaStream_t st, LaunchStats *ls)
{
    if (nA <= 0) return cudaSuccess;
    int chunk = nB > 0 ? (nB + nchunks - 1) / nchunks : 1;
    chunk = (chunk + kMatchTile - 1) / kMatchTile * kMatchTile;
    dim3 grid((nA + kMatchThreads - 1) / kMatchThreads, nchunks);
    k_match_partial<<<grid, kMatchThreads, 0, st>>>(dA, nA, dB, nB, chunk, d_partial);
    k_match_merge<<<(nA + 127) / 128, 128, 0, st>>>(d_partial, nA, nchunks, th, ratio, d_idx, d_d1, d_d2, d_accept, d_naccept);
    ls->launches += 2;
    return cudaGetLastError();
}

} // namespace orbx
