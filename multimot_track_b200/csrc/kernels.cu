// multimot_track_b200/csrc/kernels.cu -- hand-written sm_100a kernels of the ORB front end.
//
// Stage                         reference (src/ORBextractor.cc)       kernel
//   pyramid (INTER_LINEAR)      ComputePyramid :1111-1136             k_resize
//   per-cell FAST-9/16 + NMS    ComputeKeyPointsOctTree :765-829      k_fast
//   octree distribution         DistributeOctTree :539-763            k_octree
//   IC_Angle + rBRIEF           :77-147, :472-479, :1034-1041         k_orient_desc
//   7x7 sigma=2 blur            :1089-1090                            k_blur
//   border (mvImagePyramid)     :1126-1132                            k_pad_reflect101
//   Hamming best/second-best    src/ORBmatcher.cc:574-605,2279-2295   k_match_partial (merge fused: last-arriving CTA)
//
// Integer stages are bit-exact restatements of the OpenCV fixed-point arithmetic;
// float stages use explicit round-to-nearest intrinsics so nothing is contracted
// into an FMA.
#include <type_traits>
#include <cooperative_groups.h>
#ifndef RZ_SEP_TH
#define RZ_SEP_TH 32
#endif
#include "kernels.cuh"

#include <cstdio>
#include <algorithm>
#include <cstdlib>

namespace orbx {

constexpr int kMaxOptInSmem = 224 * 1024;     // dynamic shared memory opt-in cap (227 KB per CTA on sm_100, minus room for static)

// Experiment knobs (scripts/exp_*.py) exist only in the -DORBX_DEBUG_KNOBS build (`make dbg` -> liborbx_dbg.so); the release
// library never reads the environment.
static inline int debug_knob(const char *name, int dflt)
{
#ifdef ORBX_DEBUG_KNOBS
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

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

// ---- TMA / mbarrier helpers (cp.async.bulk.tensor lands a 2-D box of image bytes in shared memory) ----------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, unsigned parity)
{
    const unsigned a = smem_u32(bar);
    for (int spin = 0; spin < (1 << 24); ++spin) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();                                                      // a lost TMA completion must fail loudly, never hang the GPU
}

// ------------------------------------------------------------------ pyramid
// cv::resize INTER_LINEAR, 8UC1: H = S[s0]*a0 + S[s1]*a1 (11-bit weights), then
// out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.  The weight tables are
// built on the host with OpenCV's float arithmetic (host_tables.cpp).  A CTA produces a
// 128x16 destination tile; the source footprint (<= 2x down-scaling) is staged in shared
// memory with aligned 32-bit loads, so global traffic is coalesced and each source byte
// is fetched once per tile.

constexpr int kRzTW = 128, kRzTH = 32, kRzSrcWords = 68, kRzSrcRows = 68;
constexpr int kRzSepTH = RZ_SEP_TH;              // tile height of the TMA / separable kernel (rows per thread: kRzSepTH / 8)

// Rows [y0, y_end) x columns [x0, x0+128) of `level` from a staged source window: sb points at source pixel
// (sx_lo, sy_lo), row pitch `spitch` bytes.  256 threads as 32x8, 4 pixels x 4 rows per thread.
// No clamp is needed on the result: the two weights of an axis sum to at most 2049, so
// ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) <= 1020 and (x + 2) >> 2 <= 255.
__device__ __forceinline__ void resize_rows(const DevParams *__restrict__ P, int level, const uint8_t *sb, int spitch, int sx_lo, int sy_lo,
                                            int x0, int y0, int y_end, uint8_t *dst)
{
    const LevelGeom &D = P->lv[level];
    const ResizeTab *xt = P->xtab + P->xtab_off[level], *yt = P->ytab + P->ytab_off[level];
    const int x4 = x0 + threadIdx.x * 4;
    if (x4 >= D.w) return;
    int o0[4], o1[4], c0[4], c1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                  // tables are padded to a multiple of 4 entries
        const ResizeTab t = xt[x4 + k];
        o0[k] = t.s0 - sx_lo; o1[k] = t.s1 - sx_lo; c0[k] = t.c0; c1[k] = t.c1;
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int y = y0 + threadIdx.y + 8 * rr;
        if (y >= y_end) break;
        const ResizeTab ty = yt[y];
        const uint8_t *S0 = sb + (ty.s0 - sy_lo) * spitch, *S1 = sb + (ty.s1 - sy_lo) * spitch;
        const int b0 = ty.c0, b1 = ty.c1;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h0 = S0[o0[k]] * c0[k] + S0[o1[k]] * c1[k];
            const int h1 = S1[o0[k]] * c0[k] + S1[o1[k]] * c1[k];
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)v << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(dst + (long long)y * D.pitch + x4) = out;   // pitch % 64 == 0: padding absorbs the tail
    }
}

// One destination tile [x0, x0+128) x [y0, y_end) (y_end - y0 <= 32) of `level` from level-1.
// Block-wide (256 threads as 32x8); contains two barriers, so call it uniformly.
__device__ __forceinline__ void resize_tile(const DevParams *__restrict__ P, const uint8_t *S, int sp, uint8_t *dst, int level,
                                            int x0, int y0, int y_end, uint32_t (*ssrc)[kRzSrcWords])
{
    const LevelGeom &D = P->lv[level];
    const ResizeTab *xt = P->xtab + P->xtab_off[level], *yt = P->ytab + P->ytab_off[level];
    const int xl = min(x0 + kRzTW, D.w) - 1, yl = y_end - 1;
    const int sx_lo = xt[x0].s0 & ~3, sx_hi = xt[xl].s1, sy_lo = yt[y0].s0, sy_hi = yt[yl].s1;
    const int nwords = ((sx_hi - sx_lo) >> 2) + 1, nrows = sy_hi - sy_lo + 1;
    __syncthreads();                                               // the previous tile is done with ssrc
    for (int r = threadIdx.y; r < nrows; r += 8) {
        // plain (coherent) loads: in the fused kernel the source level was written by this CTA moments ago
        const uint32_t *row = reinterpret_cast<const uint32_t *>(S + (long long)(sy_lo + r) * sp + sx_lo);
        for (int c = threadIdx.x; c < nwords; c += 32) ssrc[r][c] = row[c];
    }
    __syncthreads();
    resize_rows(P, level, reinterpret_cast<const uint8_t *>(&ssrc[0][0]), kRzSrcWords * 4, sx_lo, sy_lo, x0, y0, y_end, dst);
}

__global__ void __launch_bounds__(256) k_resize(const DevParams *__restrict__ P, Src0 s0, int level)
{
    __shared__ uint32_t ssrc[kRzSrcRows][kRzSrcWords];
    const LevelGeom &D = P->lv[level];
    int sp;
    const uint8_t *S = level_ptr(P, s0, blockIdx.z, level - 1, &sp);
    uint8_t *dst = P->pyr + (long long)blockIdx.z * P->pyr_frame_bytes + D.img_off;
    const int y0 = blockIdx.y * kRzTH;
    resize_tile(P, S, sp, dst, level, blockIdx.x * kRzTW, y0, min(y0 + kRzTH, D.h), ssrc);
}

// ---- TMA-staged, separable variant.  Persistent CTAs walk the 128x32 destination tiles of one frame; the source
// window of the NEXT tile is in flight as one cp.async.bulk.tensor box (double buffer) while this one is resampled.
// The horizontal interpolation is done ONCE per source row of the window (H >> 4, as cv::resize keeps it) into shared
// memory, then the vertical pass combines two of those rows per destination row -- half the multiplies and a third of
// the byte loads of the direct form, with identical integer results.
struct ResizeArgs {                  // everything the kernel needs, resolved on the host (no dependent parameter loads)
    const ResizeTab *xt, *yt;        // tables of this level
    uint8_t *dst;                    // level image of frame 0
    long long frame_stride;
    int w, h, pitch;                 // destination level
    int box_w, box_h, box_bytes;
    int ntx, nty, nchunks;           // tiles per row / column; a CTA owns column tx and the tile rows chunk, chunk + nchunks, ...
};

// a = two 16-bit values, b = four bytes: c + a.lo * b.byte0 + a.hi * b.byte1 (lo) or c + a.lo * b.byte2 + a.hi * b.byte3 (hi)
__device__ __forceinline__ unsigned dp2a_lo_uu(unsigned a, unsigned b, unsigned c = 0u)
{
    unsigned d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp2a_hi_uu(unsigned a, unsigned b, unsigned c = 0u)
{
    unsigned d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp4a_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(256)
k_resize_sep(const __grid_constant__ CUtensorMap tmap, const ResizeArgs A)
{
    extern __shared__ __align__(128) uint8_t rs_smem[];            // [2][box_bytes] source boxes, then int Hs[box_h][128]
    __shared__ __align__(8) uint64_t mbar[2];
    int *Hs = reinterpret_cast<int *>(rs_smem + 2 * A.box_bytes);
    const int frame = blockIdx.y, tid = threadIdx.x + threadIdx.y * 32, q = threadIdx.x, g = threadIdx.y;
    const int chunk = blockIdx.x / A.ntx, tx = blockIdx.x - chunk * A.ntx;
    const int x0 = tx * kRzTW, x4 = x0 + 4 * q;
    const bool col_ok = x4 < A.w;
    const int sx_lo = A.xt[x0].s0 & ~15;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int ty, int b) {                               // thread 0
        const int sy_lo = A.yt[ty * kRzSepTH].s0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar[b])), "r"(A.box_w * A.box_h) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(rs_smem + b * A.box_bytes)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(&mbar[b])),
                        "r"(sx_lo), "r"(sy_lo), "r"(frame) : "memory");
    };
    int ty = chunk;
    if (tid == 0 && ty < A.nty) issue(ty, 0);
    // the four destination columns of this thread (fixed for the CTA): their <= 8 source bytes start at byte `wb` (word aligned)
    // + `sh` / 8 of the box row; sel01 / sel23 pick (S[s0], S[s1]) of two columns each out of that window, wk = a0 | a1 << 16 are
    // the 11-bit weights of one column, so that H = S[s0]*a0 + S[s1]*a1 is one two-way dot product (IDP.2A)
    int wb = 0, sh = 0;
    unsigned sel01 = 0, sel23 = 0, wk[4] = {0, 0, 0, 0};
    if (col_ok) {
        int i0[4], i1[4];
        int first = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                  // tables are padded to a multiple of 4 entries
            const ResizeTab e = A.xt[x4 + k];
            if (k == 0) first = e.s0 - sx_lo;
            i0[k] = e.s0 - sx_lo - first; i1[k] = e.s1 - sx_lo - first;   // 0 .. 7 (<= 2x down-scaling)
            wk[k] = (unsigned)e.c0 | (unsigned)e.c1 << 16;
        }
        wb = first & ~3; sh = (first & 3) * 8;
        sel01 = (unsigned)(i0[0] | i1[0] << 4 | i0[1] << 8 | i1[1] << 12);
        sel23 = (unsigned)(i0[2] | i1[2] << 4 | i0[3] << 8 | i1[3] << 12);
    }
    uint8_t *dcol = A.dst + (long long)frame * A.frame_stride + x4;
    unsigned phase = 0;
    for (int n = 0; ty < A.nty; ty += A.nchunks, ++n) {
        const int b = n & 1;
        if (tid == 0 && ty + A.nchunks < A.nty) issue(ty + A.nchunks, b ^ 1);
        const int y0 = ty * kRzSepTH, y_end = min(y0 + kRzSepTH, A.h);
        const int sy_lo = A.yt[y0].s0, nrows = A.yt[y_end - 1].s1 - sy_lo + 1;
        ResizeTab ey[kRzSepTH / 8];
#pragma unroll
        for (int rr = 0; rr < kRzSepTH / 8; ++rr) ey[rr] = A.yt[min(y0 + g + 8 * rr, y_end - 1)];
        mbar_wait_parity(&mbar[b], (phase >> b) & 1u);
        phase ^= 1u << b;
        // ---- horizontal pass: source rows of the window -> Hs[r][x] = (S[s0]*a0 + S[s1]*a1) >> 4
        if (col_ok) {
            const uint8_t *row = rs_smem + b * A.box_bytes + g * A.box_w + wb;
            int *hrow = Hs + g * kRzTW + 4 * q;
            for (int r = g; r < nrows; r += 8, row += 8 * A.box_w, hrow += 8 * kRzTW) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(row);
                const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
                int4 h;
                h.x = (int)(dp2a_lo_uu(wk[0], p01) >> 4);
                h.y = (int)(dp2a_hi_uu(wk[1], p01) >> 4);
                h.z = (int)(dp2a_lo_uu(wk[2], p23) >> 4);
                h.w = (int)(dp2a_hi_uu(wk[3], p23) >> 4);
                *reinterpret_cast<int4 *>(hrow) = h;
            }
        }
        __syncthreads();
        // ---- vertical pass: out = (((b0*H0) >> 16) + ((b1*H1) >> 16) + 2) >> 2; no clamp needed (see resize_rows)
        if (col_ok) {
#pragma unroll
            for (int rr = 0; rr < kRzSepTH / 8; ++rr) {
                const int y = y0 + g + 8 * rr;
                if (y < y_end) {
                    const ResizeTab e = ey[rr];
                    const int4 h0 = *reinterpret_cast<const int4 *>(Hs + (e.s0 - sy_lo) * kRzTW + 4 * q);
                    const int4 h1 = *reinterpret_cast<const int4 *>(Hs + (e.s1 - sy_lo) * kRzTW + 4 * q);
                    const int b0 = e.c0, b1 = e.c1;
                    // (measured: the same sums as two IMAD.HI per pixel -- (b << 16) * H >> 32 -- are not faster: IMAD.HI issues at half rate)
                    const uint32_t v0 = (((b0 * h0.x) >> 16) + ((b1 * h1.x) >> 16) + 2) >> 2, v1 = (((b0 * h0.y) >> 16) + ((b1 * h1.y) >> 16) + 2) >> 2;
                    const uint32_t v2 = (((b0 * h0.z) >> 16) + ((b1 * h1.z) >> 16) + 2) >> 2, v3 = (((b0 * h0.w) >> 16) + ((b1 * h1.w) >> 16) + 2) >> 2;
                    *reinterpret_cast<uint32_t *>(dcol + (long long)y * A.pitch) = v0 | v1 << 8 | v2 << 16 | v3 << 24;   // pitch % 64 == 0
                }
            }
        }
        __syncthreads();                                               // Hs and box b are free again
    }
}

void resize_box(const LevelGeom &src, const LevelGeom &dst, int *box_w, int *box_h)
{
    // footprint of a 128x32 destination tile: ceil(128*ratio)+2 source columns, +15 for the 16-byte aligned origin
    const int fw = (int)(((long long)kRzTW * src.w + dst.w - 1) / dst.w) + 2 + 15 + 1;
    const int fh = (int)(((long long)kRzSepTH * src.h + dst.h - 1) / dst.h) + 3;
    const int bw = (fw + 15) & ~15;
    if (bw > 256 || fh > 256) { *box_w = 0; *box_h = 0; return; }
    *box_w = bw; *box_h = fh;
}

bool encode_image_map(CUtensorMap *out, const void *base, int pitch, int rows, long long frame_stride, int nframes, int box_w, int box_h)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_fn>(p);
    }
    if (!fn || box_w <= 0 || (reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frame_stride & 15)) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nframes};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Copies frames of arbitrary row stride / alignment into the 64-byte-pitched level-0 slots.
__global__ void k_repack(const uint8_t *__restrict__ src, long long src_frame_stride, int src_pitch, uint8_t *__restrict__ dst,
                         long long dst_frame_stride, int dst_pitch, int w, int h)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t *s = src + f * src_frame_stride + (long long)y * src_pitch + x4;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (x4 + k < w) v |= (uint32_t)s[k] << (8 * k);
    *reinterpret_cast<uint32_t *>(dst + f * dst_frame_stride + (long long)y * dst_pitch + x4) = v;
}

// cv::cvtColor(..., CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) of Tracking::GrabImage* (src/Tracking.cc:459-472)
// straight into the 64-byte-pitched level-0 slots: (c0*k0 + c1*19235 + c2*k2 + 2^14) >> 15 with (k0, k2) = (9798, 3735) for
// RGB order, swapped for BGR.  Four pixels per thread, one 32-bit store.
__global__ void k_gray(const uint8_t *__restrict__ src, long long src_frame_stride, int src_pitch, int channels, int k0, int k2,
                       uint8_t *__restrict__ dst, long long dst_frame_stride, int dst_pitch, int w, int h)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t *s = src + f * src_frame_stride + (long long)y * src_pitch + (long long)x4 * channels;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (x4 + k < w) {
            const uint8_t *p = s + k * channels;
            v |= (uint32_t)((p[0] * k0 + p[1] * 19235 + p[2] * k2 + (1 << 14)) >> 15) << (8 * k);
        }
    *reinterpret_cast<uint32_t *>(dst + f * dst_frame_stride + (long long)y * dst_pitch + x4) = v;
}

cudaError_t launch_gray(const uint8_t *src, long long src_frame_stride, int src_pitch, int channels, int rgb_order, uint8_t *dst,
                        long long dst_frame_stride, int dst_pitch, int w, int h, int nframes, cudaStream_t st, LaunchStats *ls)
{
    dim3 block(128), grid((w + 511) / 512, h, nframes);
    k_gray<<<grid, block, 0, st>>>(src, src_frame_stride, src_pitch, channels, rgb_order ? 9798 : 3735, rgb_order ? 3735 : 9798, dst,
                                   dst_frame_stride, dst_pitch, w, h);
    ls->launches++;
    return cudaGetLastError();
}

cudaError_t launch_repack(const uint8_t *src, long long src_frame_stride, int src_pitch, uint8_t *dst, long long dst_frame_stride,
                          int dst_pitch, int w, int h, int nframes, cudaStream_t st, LaunchStats *ls)
{
    dim3 block(128), grid((w + 511) / 512, h, nframes);
    k_repack<<<grid, block, 0, st>>>(src, src_frame_stride, src_pitch, dst, dst_frame_stride, dst_pitch, w, h);
    ls->launches++;
    return cudaGetLastError();
}

// Fallback for scale factors above 2 (source footprint larger than the shared tile): direct global reads.
__global__ void __launch_bounds__(256) k_resize_direct(const DevParams *__restrict__ P, Src0 s0, int level)
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
    const ResizeTab *tx = P->xtab + P->xtab_off[level] + x4;
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
    *reinterpret_cast<uint32_t *>(dst + (long long)y * D.pitch + x4) = out;
}

cudaError_t launch_pyramid(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls, const TmaMaps *tma)
{
    if (hP.nlevels < 2) return cudaSuccess;
    bool all_staged = true;
    for (int l = 1; l < hP.nlevels; ++l) {
        const LevelGeom &D = hP.lv[l], &Sg = hP.lv[l - 1];
        // staged path needs the source footprint of a 128x32 tile to fit 68 words x 68 rows
        const bool staged = (long long)Sg.w * (kRzTW + 2) <= (long long)D.w * (kRzSrcWords * 4 - 12) &&
                            (long long)Sg.h * (kRzTH + 2) <= (long long)D.h * (kRzSrcRows - 3);
        all_staged = all_staged && staged;
    }
    if (all_staged) {
        // one launch per level: thousands of independent tiles hide the load latency better than the
        // single-launch k_pyramid_fused (few long-running CTAs), which measured 1.6x slower at batch 32
        for (int l = 1; l < hP.nlevels; ++l) {
            const LevelGeom &D = hP.lv[l];
            const dim3 grid((D.w + kRzTW - 1) / kRzTW, (D.h + kRzTH - 1) / kRzTH, nframes);
            if (tma && tma->ok[l]) {                               // source windows by TMA, persistent separable kernel
                const int bw = tma->box_w[l], bh = tma->box_h[l];
                const int box_bytes = (bw * bh + 127) / 128 * 128;
                const size_t smem = 2 * (size_t)box_bytes + (size_t)bh * kRzTW * sizeof(int);
                if (smem > 40 * 1024) {                            // static + dynamic above 48 KB needs the opt-in (per device: set every time)
                    cudaError_t e = cudaFuncSetAttribute(k_resize_sep, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptInSmem);
                    if (e != cudaSuccess) return e;
                }
                int per_sm = (int)((227 * 1024) / (smem + 1024));
                per_sm = per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm);
                const int ntx = grid.x, nty = (D.h + kRzSepTH - 1) / kRzSepTH;
                const int cap = std::max(1, 148 * per_sm / nframes / ntx);   // resident CTAs available to one tile column of one frame
                const int iters = (nty + cap - 1) / cap;
                ResizeArgs A;
                A.xt = hP.xtab + hP.xtab_off[l]; A.yt = hP.ytab + hP.ytab_off[l];
                A.dst = hP.pyr + D.img_off; A.frame_stride = hP.pyr_frame_bytes;
                A.w = D.w; A.h = D.h; A.pitch = D.pitch; A.box_w = bw; A.box_h = bh; A.box_bytes = box_bytes;
                A.ntx = ntx; A.nty = nty; A.nchunks = (nty + iters - 1) / iters;
                k_resize_sep<<<dim3(ntx * A.nchunks, nframes), dim3(32, 8), smem, st>>>(tma->src[l], A);
                const cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return e;
            } else
                k_resize<<<grid, dim3(32, 8), 0, st>>>(dP, s0, l);
            ls->launches++;
        }
        return cudaGetLastError();
    }
    for (int l = 1; l < hP.nlevels; ++l) {
        const LevelGeom &D = hP.lv[l];
        dim3 block(32, 8), grid((D.w + 127) / 128, (D.h + 7) / 8, nframes);
        k_resize_direct<<<grid, block, 0, st>>>(dP, s0, l);
        ls->launches++;
    }
    return cudaGetLastError();
}

// --------------------------------------------------------------------- blur
// cv::GaussianBlur 7x7 sigma 2 on 8-bit: 8.8 fixed-point kernel [18,34,48,56,48,34,18],
// horizontal pass to u16, vertical pass to u32, one rounding (+32768)>>16; reflect-101.
// A 128x64 output tile per CTA.  The horizontal pass runs on pixel PAIRS packed as
// 16x2 in one register (18..56 * 255 summed stays below 65536, so a plain 32-bit
// multiply-add never carries between the two lanes); the vertical pass slides down
// 8 rows per thread on the 16-bit row sums kept in shared memory.

// Persistent CTAs walk the (tile, frame) work list.  The raw tile (rows ty0-3 .. ty0+66, columns tx0-16 .. tx0+143,
// 160-byte pitch: TMA needs a 16-byte aligned box origin) arrives as ONE cp.async.bulk.tensor box, double buffered so
// the box of the next tile is in flight while this one is filtered.  Pixels outside the image come back as zeros (or
// row padding); only the <= 3 the 7-tap kernel can reach on each side are overwritten with their reflect-101 sources.
constexpr int kBlurRawWords = kBlurBoxW / 4;

// 18*(p0+p6) + 34*(p1+p5) + 48*(p2+p4) + 56*p3 + acc as a chain of IMADs (kept from being re-associated into adds)
__device__ __forceinline__ uint32_t mad7(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t p5, uint32_t p6, uint32_t acc)
{
    asm("mad.lo.u32 %0, %1, 56, %0;" : "+r"(acc) : "r"(p3));
    asm("mad.lo.u32 %0, %1, 48, %0;" : "+r"(acc) : "r"(p2));
    asm("mad.lo.u32 %0, %1, 48, %0;" : "+r"(acc) : "r"(p4));
    asm("mad.lo.u32 %0, %1, 34, %0;" : "+r"(acc) : "r"(p1));
    asm("mad.lo.u32 %0, %1, 34, %0;" : "+r"(acc) : "r"(p5));
    asm("mad.lo.u32 %0, %1, 18, %0;" : "+r"(acc) : "r"(p0));
    asm("mad.lo.u32 %0, %1, 18, %0;" : "+r"(acc) : "r"(p6));
    return acc;
}

struct BlurItem { int level, tx0, ty0, frame; };

// item = frame * n_blur_work + widx; the persistent loop advances (frame, widx) incrementally (no division per tile)
__device__ __forceinline__ BlurItem blur_item(const DevParams *__restrict__ P, int frame, int widx)
{
    const uint32_t wk = P->blur_work[widx];
    BlurItem it;
    it.level = wk >> 24; it.ty0 = ((wk >> 12) & 0xfff) * kBlurTileH; it.tx0 = (wk & 0xfff) * kBlurTileW; it.frame = frame;
    return it;
}

__global__ void __launch_bounds__(256)
k_blur(const DevParams *__restrict__ P, Src0 s0, const __grid_constant__ BlurMaps maps, unsigned tma_levels, int total_items)
{
    constexpr int TW = kBlurTileW, TH = kBlurTileH, RWD = kBlurRawWords, PB = kBlurBoxW;
    struct __align__(128) RawBuf { uint32_t w[TH + 6][RWD]; };        // TMA destinations must be 128-byte aligned
    __shared__ RawBuf sraw[2];
    __shared__ uint4 shv[(TH + 6) / 2][TW / 4];                      // H of two tile rows interleaved: H(x, 2p) | H(x, 2p+1) << 16
    __shared__ __align__(8) uint64_t mbar[2];
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: start the box of `item` into buffer b (no-op for a level without a tensor map)
    const int nw = P->n_blur_work, gstep = (int)gridDim.x;
    const int step_q = gstep / nw, step_r = gstep - step_q * nw;      // once per CTA
    auto advance = [&](int &frame, int &widx) { frame += step_q; widx += step_r; if (widx >= nw) { widx -= nw; ++frame; } };
    auto issue = [&](int frame, int widx, int b) {
        const BlurItem it = blur_item(P, frame, widx);
        if (!((tma_levels >> it.level) & 1u)) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads/writes of this buffer come first
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar[b])), "r"((TH + 6) * PB) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(&sraw[b].w[0][0])), "l"(reinterpret_cast<unsigned long long>(&maps.m[it.level])), "r"(smem_u32(&mbar[b])),
                        "r"(it.tx0 - 16), "r"(it.ty0 - 3), "r"(it.frame) : "memory");
    };
    int item = blockIdx.x;
    int cur_frame = item / nw, cur_widx = item - cur_frame * nw;     // the only division: once per CTA
    if (tid == 0 && item < total_items) issue(cur_frame, cur_widx, 0);
    unsigned phase = 0;                                              // bit b = parity the next wait on buffer b uses
    for (int n = 0; item < total_items; item += gstep, ++n) {
        const int b = n & 1;
        int nxt_frame = cur_frame, nxt_widx = cur_widx;
        advance(nxt_frame, nxt_widx);
        if (tid == 0 && item + gstep < total_items) issue(nxt_frame, nxt_widx, b ^ 1);
        const BlurItem it = blur_item(P, cur_frame, cur_widx);
        cur_frame = nxt_frame; cur_widx = nxt_widx;
        const int level = it.level, tx0 = it.tx0, ty0 = it.ty0, frame = it.frame;
        const LevelGeom &G = P->lv[level];
        uint32_t (*raw)[RWD] = sraw[b].w;
        if ((tma_levels >> level) & 1u) {
            mbar_wait_parity(&mbar[b], (phase >> b) & 1u);
            phase ^= 1u << b;
            uint8_t *sb = reinterpret_cast<uint8_t *>(&raw[0][0]);
            // rows above / below the image: whole-row copies from their reflect-101 source rows (inside the tile)
            const int rb = G.h - ty0 + 3;                            // tile row of image row h
            if (ty0 == 0 || rb < TH + 6) {
                if (tid < 6 * RWD) {
                    const int k = tid / RWD, c = tid - k * RWD;
                    const int r = k < 3 ? k : rb + k - 3;
                    const int gy = ty0 - 3 + r;
                    if (r < TH + 6 && (unsigned)gy >= (unsigned)G.h && gy < G.h + 3)
                        raw[r][c] = raw[reflect101(gy, G.h) - (ty0 - 3)][c];
                }
                __syncthreads();
            }
            // columns left / right of the image
            const bool left = tx0 == 0, right = tx0 + TW + 3 > G.w;
            if (left || right) {
                if (tid < 2 * (TH + 6)) {
                    const int side = tid >= TH + 6, r = tid - side * (TH + 6);
                    uint8_t *row = sb + r * PB;
                    if (side == 0 && left) {
#pragma unroll
                        for (int k = 1; k <= 3; ++k) row[16 - k] = row[16 + k];
                    }
                    if (side == 1 && right) {
                        const int cl = G.w - 1 - (tx0 - 16);
#pragma unroll
                        for (int k = 1; k <= 3; ++k) if (cl + k < PB) row[cl + k] = row[cl - k];
                    }
                }
            }
        } else {
            int sp;
            const uint8_t *S = level_ptr(P, s0, frame, level, &sp);
            // ---- raw rows ty0-3 .. ty0+TH+2 (reflected), columns tx0-4 .. tx0+TW+3 (tile words 3 .. 36)
            for (int i = tid; i < (TH + 6) * (TW / 4 + 2); i += 256) {
                const int r = i / (TW / 4 + 2), c = i - r * (TW / 4 + 2);
                const int gy = reflect101(ty0 + r - 3, G.h), gx = tx0 - 4 + 4 * c;
                const uint8_t *row = S + (long long)gy * sp;
                uint32_t v;
                if (gx >= 0 && gx + 3 < G.w) v = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
                else v = (uint32_t)row[reflect101(gx, G.w)] | (uint32_t)row[reflect101(gx + 1, G.w)] << 8 |
                         (uint32_t)row[reflect101(gx + 2, G.w)] << 16 | (uint32_t)row[reflect101(gx + 3, G.w)] << 24;
                raw[r][c + 3] = v;
            }
        }
        __syncthreads();
        // ---- horizontal pass: item = (pair of tile rows, pixel quad).  H(x) = sum_i k[i] * p[x+i-3] as two four-way dot
        //      products (IDP.4A) on the byte windows x-3..x and x+1..x+4 (funnel shifts of the three aligned words around the quad);
        //      the two rows leave interleaved, H(x, r) | H(x, r+1) << 16, which is the operand form the vertical pass wants
        constexpr unsigned kWA = 18u | 34u << 8 | 48u << 16 | 56u << 24, kWB = 48u | 34u << 8 | 18u << 16;
        // edge tiles: only the row pairs / quads the vertical pass of the pixels inside the image will read
        const int pr_end = min((TH + 6) / 2, (min(TH, G.h - ty0) + 7) >> 1), q_end = min(TW / 4, (G.w - tx0 + 3) >> 2);
        for (int i = tid; i < pr_end * (TW / 4); i += 256) {
            const int pr = i / (TW / 4), q = i - pr * (TW / 4);
            if (q >= q_end) continue;
            uint32_t hrow[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint32_t *rw = &raw[2 * pr + rr][q + 3];
                const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];              // pixels x-4..x-1 | x..x+3 | x+4..x+7
                hrow[rr][0] = dp4a_uu(__funnelshift_r(w1, w2, 8), kWB, dp4a_uu(__funnelshift_r(w0, w1, 8), kWA, 0u));
                hrow[rr][1] = dp4a_uu(__funnelshift_r(w1, w2, 16), kWB, dp4a_uu(__funnelshift_r(w0, w1, 16), kWA, 0u));
                hrow[rr][2] = dp4a_uu(__funnelshift_r(w1, w2, 24), kWB, dp4a_uu(__funnelshift_r(w0, w1, 24), kWA, 0u));
                hrow[rr][3] = dp4a_uu(w2, kWB, dp4a_uu(w1, kWA, 0u));
            }
            shv[pr][q] = make_uint4(__byte_perm(hrow[0][0], hrow[1][0], 0x5410), __byte_perm(hrow[0][1], hrow[1][1], 0x5410),
                                    __byte_perm(hrow[0][2], hrow[1][2], 0x5410), __byte_perm(hrow[0][3], hrow[1][3], 0x5410));
        }
        __syncthreads();
        // ---- vertical pass: thread = (pixel quad, 8-row segment).  V(y) = sum_j k[j] * H(y+j) + 32768 over the row pairs as four
        //      two-way dot products (IDP.2A): even rows pair the taps (0,1)(2,3)(4,5)(6,-), odd rows (-,0)(1,2)(3,4)(5,6)
        const int q = tid & 31, seg = tid >> 5;
        const int gx = tx0 + 4 * q, gy0 = ty0 + seg * 8;
        if (gx < G.w && gy0 < G.h) {
            uint4 pv[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) pv[j] = shv[seg * 4 + j][q];
            uint32_t out[8];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const uint32_t *p0 = &pv[m].x, *p1 = &pv[m + 1].x, *p2 = &pv[m + 2].x, *p3 = &pv[m + 3].x;
                uint32_t ve[4], vo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    ve[c] = dp2a_lo_uu(p3[c], 18u, dp2a_lo_uu(p2[c], 48u | 34u << 8, dp2a_lo_uu(p1[c], 48u | 56u << 8, dp2a_lo_uu(p0[c], 18u | 34u << 8, 32768u))));
                    vo[c] = dp2a_lo_uu(p3[c], 34u | 18u << 8, dp2a_lo_uu(p2[c], 56u | 48u << 8, dp2a_lo_uu(p1[c], 34u | 48u << 8, dp2a_lo_uu(p0[c], 18u << 8, 32768u))));
                }
                // byte 2 of each sum is the rounded result (sums < 2^24)
                out[2 * m] = __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
                out[2 * m + 1] = __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
            }
            uint8_t *dst = P->blur + (long long)frame * P->pyr_frame_bytes + G.img_off + gx;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (gy0 + r < G.h) *reinterpret_cast<uint32_t *>(dst + (long long)(gy0 + r) * G.pitch) = out[r];
        }
        __syncthreads();                                               // shv and raw[b] are free again
    }
}

cudaError_t launch_blur(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls,
                        const BlurMaps *maps, unsigned tma_levels)
{
    static const BlurMaps zero_maps = {};
    const int total = hP.n_blur_work * nframes;
    if (total <= 0) return cudaSuccess;
    const int grid = total < 148 * 4 ? total : 148 * 4;               // 4 resident CTAs per SM (64 registers x 256 threads)
    k_blur<<<grid, 256, 0, st>>>(dP, s0, maps ? *maps : zero_maps, maps ? tma_levels : 0u, total);
    ls->launches++;
    return cudaGetLastError();
}

// --------------------------------------------------------------------- FAST
// The reference calls cv::FAST once per ~30-pixel cell with iniThFAST and again with
// minThFAST when the cell stayed empty (:789-829).  The corner score (OpenCV
// cornerScore<16>: the largest threshold at which the pixel is still a FAST-9 corner)
// does not depend on the threshold, and "corner at t" <=> score >= t, so one dense
// scoring pass serves both thresholds.  Scores are computed on pixel PAIRS held as 16x2
// lanes so the packed-halfword integer pipe does two pixels per instruction
// (VIMNMX3.S16x2):  bright = max_k min(r_k..r_k+8) - c,  dark = c - min_k max(r_k..r_k+8)
// over the 16 ring pixels r_k (min/max commute with subtracting the centre c),
// score = max(bright, dark) - 1.  Each lane owns one pair column and walks down the rows
// keeping the 7-row ring neighbourhood in registers (5 shared-memory loads + 4 funnel
// shifts per new row).

// ---- score + per-cell NMS + threshold choice + emission in one kernel ---------------------------------------
// One WARP per job = one cell row x up to two horizontally adjacent cells (<= 64 pixels = 32 pixel pairs, the
// pairs start at the cell boundary).  The warp walks down the cell's rows; the
// scores of the previous rows stay in registers, left / right neighbours come from the adjacent lanes by
// shuffle, and neighbours that belong to another cell are masked to 0 (cv::FAST runs on the cell sub-image,
// so pixels outside it do not take part in the non-max suppression).  Survivors with score >= min(ini,min)
// go to a per-warp list; when the job's rows are done the cell decides iniThFAST vs minThFAST (:812-816).
template <int CELL>
struct FfCfg {
    static constexpr int WARPS = 4;
    static constexpr int PITCH = 37;                                   // tile words per row: pairs -2 .. 33, +1
    static constexpr int ROWS = CELL + 6;
    static constexpr int LIST_LANES = 32 * ((CELL + 1) / 2);           // per-lane survivor slots: a column pair keeps <= one pixel every 2nd row
    static constexpr int LIST = LIST_LANES + CELL;                     // + the lane that straddles two cells (independent columns: up to one entry per row)
    static constexpr int WARP_WORDS = ROWS * PITCH + LIST;
    static constexpr int SMEM = WARPS * WARP_WORDS * 4;
};

// A job = one cell row x up to two adjacent cells of one level of one frame.
struct FfJob { int level, x0, y0, nrows, span, wc, frame; };

__device__ __forceinline__ FfJob ff_job(const DevParams *__restrict__ P, int widx, int frame)
{
    const uint32_t wk = P->ffast_work[widx];
    FfJob j;
    j.level = wk >> 24;
    const int ci = (wk >> 12) & 0xfff, cj = wk & 0xfff;
    const LevelGeom &G = P->lv[j.level];
    j.wc = G.w_cell;
    const int ncell = min(max(1, 64 / j.wc), G.cols_vis - cj);
    j.x0 = kEdge + cj * j.wc;
    const int x1 = min(j.x0 + ncell * j.wc, G.x_end);
    j.y0 = kEdge + ci * G.h_cell;
    const int y1 = min(j.y0 + G.h_cell, G.y_end);
    j.span = x1 - j.x0; j.nrows = y1 - j.y0; j.frame = frame;
    return j;
}

// Scores, cell-local non-max suppression, threshold choice and emission of one job whose 16-bit tile is staged.
// ~x as x * -1 + -1, and ~x - k in one go as x * -1 + (-1 - k): an IMAD, i.e. FMA-pipe work, in a kernel that is bound by the
// integer ALU pipe.  The multiplier has to be a register ptxas cannot see through (it folds a literal -1 into an IADD3, ALU pipe
// again): ff_process derives it from a kernel-parameter field.

__device__ __forceinline__ unsigned lds_u16(unsigned addr)          // kept apart from the word loads of the same row on purpose
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
// d = a & ~b and nz = (d != 0) as ONE LOP3 with a predicate output (written as two C operations the compiler emits two LOP3s)
__device__ __forceinline__ void and_not_p(unsigned &d, unsigned &nz, unsigned a, unsigned b)
{
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, 0, 0;\n\tlop3.or.b32 %0|p, %2, %3, 0, 0x30, q;\n\tselp.u32 %1, 1, 0, p;\n\t}"
        : "=r"(d), "=r"(nz) : "r"(a), "r"(b));
}
// prmt in its generic PTX mode: bit 3 of a selector nibble replicates the sign of the chosen byte.  (__byte_perm documents three
// bits per nibble only, and the compiler does drop the fourth from a selector it cannot see as a constant.)
__device__ __forceinline__ unsigned prmt_generic(unsigned a, unsigned b, unsigned sel)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned mad_fma(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// BOUND_ONLY (experiment builds, ORBX_FAST_BOUND=1): the exact score is replaced by its cheap upper bound from the eight opposing
// ring pairs, U = max(min(A, B) - c, c - max(A, B)) - 1 with A = min_k max(r_k, r_k+8), B = max_k min(r_k, r_k+8) (every 9-arc holds one
// pixel of every opposing pair and both pixels of one).  Results are then NOT the reference's; the variant exists to time the floor of
// any "pre-test first, exact score for the survivors" scheme: this is the work such a scheme still does for EVERY pixel.
template <int CELL, bool BOUND_ONLY = false>
__device__ __forceinline__ void ff_process(const DevParams *__restrict__ P, const FfJob &J, uint32_t *tile, uint32_t *list, int lane)
{
    using C = FfCfg<CELL>;
    const int level = J.level, x0 = J.x0, y0 = J.y0, nrows = J.nrows, span = J.span, wc = J.wc, frame = J.frame;
    const LevelGeom &G = P->lv[level];
    // ---- per-lane constants: which of my two pixels are inside the job, and which neighbours share their cell
    const int dx0 = 2 * lane, dx1 = dx0 + 1;
    const bool in0 = dx0 < span, in1 = dx1 < span;
    const int c0 = dx0 >= wc, c1 = dx1 >= wc;                          // cell (0/1) of my pixels
    // pixels outside the job keep whatever score their ring gives (they are never anybody's neighbour, see below) and are
    // dropped where survivors are recorded
    const unsigned sel_nm = (in0 ? 0x99u : 0xCCu) | (in1 ? 0xBB00u : 0xCC00u);     // sign of t per halfword, or "not a survivor" outside the job
    // Lv = (left neighbour of px0, px0 as left neighbour of px1); Rv = (px1 as right neighbour of px0, right neighbour of px1),
    // each half zero when that neighbour is in another cell or outside the job.  Scores S' are <= 255, so the high byte of
    // every half is zero and ONE byte permute both moves the halves and masks them (a masked half selects high bytes only).
    const bool mL0 = dx0 > 0 && (dx0 - 1 >= wc) == c0, mL1 = c0 == c1;
    const bool mR0 = in1 && c0 == c1, mR1 = dx1 + 1 < span && (dx1 + 1 >= wc) == c1;
    const unsigned selL = (mL0 ? 0x76u : 0x77u) | (mL1 ? 0x1000u : 0x1100u);       // __byte_perm(Cv, Pl, selL) = (Pl.hi | 0, Cv.lo | 0)
    const unsigned selR = (mR0 ? 0x32u : 0x33u) | (mR1 ? 0x5400u : 0x5500u);       // __byte_perm(Cv, Pr, selR) = (Cv.hi | 0, Pr.lo | 0)
    const bool straddle = in1 && c0 != c1;                             // my two pixels belong to different cells: independent columns
    const int th_store = min(P->min_th, P->ini_th);
    // S' = score - th_store + 1 where the pixel is a corner at th_store (>= 1), else 0.  VIADD.16x2 / VIADDMNMX.S16x2 are
    // the native packed forms; there is no packed subtract, so differences are written as x + ~y (= x - y - 1).
    // (x half by half) -x - th = ~x - (th - 1): no borrow between the halves for x <= 255, th <= 255, so one 32-bit IMAD does both
    const unsigned m1 = (unsigned)(P->nlevels >> 31) - 1u;             // 0xffffffff, opaque to the compiler
    const unsigned kv = 0xffffffffu - 0x00010001u * (unsigned)(th_store - 1);
    unsigned T2 = 0, T1 = 0, U1 = 0, C1 = 0;                           // T/U/centre of rows y-2 and y-1 (0 outside the cell)
    // a lane whose two pixels share a cell keeps at most one survivor every second row; the one lane that straddles two cells
    // (independent columns) may have one per row and gets its own region
    uint32_t *const mylist = straddle ? list + C::LIST_LANES : list + lane;
    const int lstride = straddle ? 1 : 32;
    const unsigned lp0 = smem_u32(mylist), lstride4 = 4u * lstride;
    unsigned lp = lp0;                                                 // 32-bit shared-memory address of the lane's next free slot

    unsigned a[7][3], o[7][4];
    const uint32_t *col = tile + lane + 2;
    const unsigned col_s = smem_u32(col);
    // the odd-offset pairs (w_k.hi, w_k+1.lo) as hi16(w_k) + w_k+1 * 65536: a 16-bit shared-memory load and an IMAD (FMA pipe)
    // instead of a funnel shift (ALU pipe, the one this kernel is bound by)
#define FF_LOAD(slot, r)                                                                        \
    {                                                                                           \
        const uint32_t *q = col + (r) * C::PITCH;                                               \
        const unsigned qa = col_s + (r) * (C::PITCH * 4);                                       \
        const unsigned w1_ = q[-1], w2_ = q[0], w3_ = q[1], w4_ = q[2];                         \
        const unsigned h0_ = lds_u16(qa - 6), h1_ = lds_u16(qa - 2), h2_ = lds_u16(qa + 2), h3_ = lds_u16(qa + 6); \
        a[slot][0] = w1_; a[slot][1] = w2_; a[slot][2] = w3_;                                   \
        o[slot][0] = mad_fma(w1_, 65536u, h0_); o[slot][1] = mad_fma(w2_, 65536u, h1_);         \
        o[slot][2] = mad_fma(w3_, 65536u, h2_); o[slot][3] = mad_fma(w4_, 65536u, h3_);         \
    }
    // finalize the non-max suppression of row `yy` whose centre is Cc, with max-of-3 of the rows above / below.
    // t = Cc + ~m8 = Cc - m8 - 1 >= 0  <=>  centre strictly greater than all 8 neighbours (and then Cc >= 1).  PRMT in sign-replicate
    // mode turns the sign of t into a halfword mask, and takes the mask of a pixel outside the job from a constant instead (sel_nm).
    // A row with a survivor leaves ONE entry in the lane's private list (slot-major, so no ballot / prefix is needed):
    // S' of pixel 0 | row << 9 | S' of pixel 1 << 16, S' = 0 for a pixel that did not survive.  Nearly every warp row has a survivor in
    // some lane, so the append is written to compile to predicated instructions, not a divergent branch.
#define FF_NMS(yy9, Tabove, Umid, Cc, Tbelow)                                                                    \
    {                                                                                                            \
        const unsigned m8 = __vimax3_s16x2(Tabove, Umid, Tbelow);                                                \
        const unsigned t_ = __vadd2(mad_fma(m8, m1, m1), (Cc));                                                  \
        unsigned ev, any_;                                                                                       \
        and_not_p(ev, any_, (Cc), prmt_generic(t_, 0x80808080u, sel_nm));                                        \
        if (any_) {                                                                                              \
            sts_u32(lp, ev + (yy9));                                                                             \
            lp += lstride4;                                                                                      \
        }                                                                                                        \
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) FF_LOAD(r, r)
    // one row of scores + the NMS of the row above; U = row index mod 7 (a compile-time constant: the 7-row window rotates through
    // fixed register slots).  Whole groups of seven rows run without per-row guards, the last partial group with them.
    auto ff_row = [&](auto U, const int g, const unsigned g9) {        // row g + u of the job; g is a multiple of 7, g9 = g << 9
        constexpr int u = decltype(U)::value;
        const int y = g + u;
        FF_LOAD((u + 6) % 7, y + 6)
        const int sc = (u + 3) % 7, sp3 = (u + 6) % 7, sp2 = (u + 5) % 7, sp1 = (u + 4) % 7;
        const int sm1 = (u + 2) % 7, sm2 = (u + 1) % 7, sm3 = u % 7;
        const unsigned cc = a[sc][1];
        unsigned e[16];
        e[0] = a[sp3][1]; e[1] = o[sp3][2]; e[15] = o[sp3][1];
        e[2] = a[sp2][2]; e[14] = a[sp2][0];
        e[3] = o[sp1][3]; e[13] = o[sp1][0];
        e[4] = o[sc][3];  e[12] = o[sc][0];
        e[5] = o[sm1][3]; e[11] = o[sm1][0];
        e[6] = a[sm2][2]; e[10] = a[sm2][0];
        e[7] = o[sm3][2]; e[8] = a[sm3][1]; e[9] = o[sm3][1];
        unsigned maxmin, minmax;
        if (BOUND_ONLY) {
            unsigned M[8], m[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { M[i] = __vmaxs2(e[i], e[i + 8]); m[i] = __vmins2(e[i], e[i + 8]); }
            const unsigned A = __vimin3_s16x2(__vimin3_s16x2(M[0], M[1], M[2]), __vimin3_s16x2(M[3], M[4], M[5]), __vmins2(M[6], M[7]));
            const unsigned B = __vimax3_s16x2(__vimax3_s16x2(m[0], m[1], m[2]), __vimax3_s16x2(m[3], m[4], m[5]), __vmaxs2(m[6], m[7]));
            maxmin = __vmins2(A, B); minmax = __vmaxs2(A, B);
        } else {
        // The 16 arcs of 9 in pairs: arcs 2j and 2j+1 share the 8 ring pixels 2j+1 .. 2j+8 (W_j), so
        // max(min(arc 2j), min(arc 2j+1)) = min(W_j, max(e[2j], e[2j+9])).  W_j is four consecutive pair minima lo2[j .. j+3]; the windows
        // of the pairs 2i and 2i+1 share the middle three (ring pixels 4i+3 .. 4i+8), taken once per i as a three-input min:
        // 8 + 8 + 4 + 8 + 4 = 32 packed min / max per polarity.
        unsigned lo2[8], hi2[8], mx[8], mn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            lo2[i] = __vmins2(e[2 * i + 1], e[(2 * i + 2) & 15]);
            hi2[i] = __vmaxs2(e[2 * i + 1], e[(2 * i + 2) & 15]);
            mx[i] = __vmaxs2(e[2 * i], e[(2 * i + 9) & 15]);
            mn[i] = __vmins2(e[2 * i], e[(2 * i + 9) & 15]);
        }
        unsigned bv[8], dv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned cmin = __vimin3_s16x2(lo2[2 * i + 1], lo2[(2 * i + 2) & 7], lo2[(2 * i + 3) & 7]);
            const unsigned cmax = __vimax3_s16x2(hi2[2 * i + 1], hi2[(2 * i + 2) & 7], hi2[(2 * i + 3) & 7]);
            bv[2 * i] = __vimin3_s16x2(cmin, lo2[2 * i], mx[2 * i]);
            bv[2 * i + 1] = __vimin3_s16x2(cmin, lo2[(2 * i + 4) & 7], mx[2 * i + 1]);
            dv[2 * i] = __vimax3_s16x2(cmax, hi2[2 * i], mn[2 * i]);
            dv[2 * i + 1] = __vimax3_s16x2(cmax, hi2[(2 * i + 4) & 7], mn[2 * i + 1]);
        }
        maxmin = __vimax3_s16x2(__vimax3_s16x2(bv[0], bv[1], bv[2]), __vimax3_s16x2(bv[3], bv[4], bv[5]), __vmaxs2(bv[6], bv[7]));
        minmax = __vimin3_s16x2(__vimin3_s16x2(dv[0], dv[1], dv[2]), __vimin3_s16x2(dv[3], dv[4], dv[5]), __vmins2(dv[6], dv[7]));
        }
        // S' = max(bright, dark) - th clamped at 0, bright = maxmin - c, dark = c - minmax: (-c - th) and (-minmax - th) by IMAD
        const unsigned Cv = __viaddmax_s16x2_relu(cc, mad_fma(minmax, m1, kv), __vadd2(maxmin, mad_fma(cc, m1, kv)));
        // neighbours in the same row, masked to the pixels' own cells
        const unsigned Pl = __shfl_up_sync(0xffffffffu, Cv, 1), Pr = __shfl_down_sync(0xffffffffu, Cv, 1);
        const unsigned Lv = __byte_perm(Cv, Pl, selL), Rv = __byte_perm(Cv, Pr, selR);
        const unsigned T0 = __vimax3_s16x2(Lv, Cv, Rv), U0 = __vmaxs2(Lv, Rv);
        if (u > 0 || g > 0) FF_NMS(g9 + (unsigned)((u - 1) * 512), T2, U1, C1, T0)
        T2 = T1; T1 = T0; U1 = U0; C1 = Cv;
    };
    int g = 0;
    unsigned g9 = 0;
    for (; g + 7 <= nrows; g += 7, g9 += 7u << 9) {
        ff_row(std::integral_constant<int, 0>{}, g, g9); ff_row(std::integral_constant<int, 1>{}, g, g9);
        ff_row(std::integral_constant<int, 2>{}, g, g9); ff_row(std::integral_constant<int, 3>{}, g, g9);
        ff_row(std::integral_constant<int, 4>{}, g, g9); ff_row(std::integral_constant<int, 5>{}, g, g9);
        ff_row(std::integral_constant<int, 6>{}, g, g9);
    }
    if (g < nrows) ff_row(std::integral_constant<int, 0>{}, g, g9);
    if (g + 1 < nrows) ff_row(std::integral_constant<int, 1>{}, g, g9);
    if (g + 2 < nrows) ff_row(std::integral_constant<int, 2>{}, g, g9);
    if (g + 3 < nrows) ff_row(std::integral_constant<int, 3>{}, g, g9);
    if (g + 4 < nrows) ff_row(std::integral_constant<int, 4>{}, g, g9);
    if (g + 5 < nrows) ff_row(std::integral_constant<int, 5>{}, g, g9);
    FF_NMS((unsigned)(nrows - 1) << 9, T2, U1, C1, 0u)
#undef FF_LOAD
#undef FF_NMS
    __syncwarp();
    // ---- threshold choice per cell (:812-816) and emission
    const int ini = P->ini_th, mn = P->min_th;
    const int s_ini = ini - th_store + 1;                              // S' >= s_ini  <=>  score >= iniThFAST
    // the lane's survivors at iniThFAST and at all, counted on both pixels at once (pixel 0 | pixel 1 << 16)
    const unsigned k_ini = 0x00010001u * (unsigned)((1 - s_ini) & 0xffff);
    unsigned n_str = 0, n_pos = 0;
    for (unsigned q = lp0; q != lp; q += lstride4) {
        const unsigned e2 = lds_u32(q) & 0x01ff01ffu;
        n_pos += __vminu2(e2, 0x00010001u);
        n_str += __viaddmin_s16x2_relu(e2, k_ini, 0x00010001u);
    }
    const int strA = n_str & 0xffffu, strB = n_str >> 16, weakA = (int)(n_pos & 0xffffu) - strA, weakB = (int)(n_pos >> 16) - strB;
    const int strong0 = (c0 ? 0 : strA) + (c1 ? 0 : strB), strong1 = (c0 ? strA : 0) + (c1 ? strB : 0);
    const int weak0 = (c0 ? 0 : weakA) + (c1 ? 0 : weakB), weak1 = (c0 ? weakA : 0) + (c1 ? weakB : 0);
    const bool has0 = __any_sync(0xffffffffu, strong0 > 0), has1 = __any_sync(0xffffffffu, strong1 > 0);
    // a cell with a corner at iniThFAST keeps only those; an empty one falls back to minThFAST (when that is lower)
    const int s0c = has0 ? s_ini : (mn < ini ? 1 : 0x7fffffff), s1c = has1 ? s_ini : (mn < ini ? 1 : 0x7fffffff);
    const int mine = strong0 + strong1 + (has0 || mn >= ini ? 0 : weak0) + (has1 || mn >= ini ? 0 : weak1);
    int incl = mine;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, ofs); if (lane >= ofs) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    uint32_t *cand = P->cand + (long long)frame * P->cand_frame_elems + G.cand_off;
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(P->cand_count + frame * P->nlevels + level, (unsigned)total);
    int at = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
    // candidate = x | y << 12 | score << 24, score = S' + th_store - 1
    const uint32_t yx0 = ((uint32_t)(x0 + dx0 - kMinBorder) | (uint32_t)(y0 - kMinBorder) << 12) + ((uint32_t)(th_store - 1) << 24);
    const int thr0 = max(1, c0 ? s1c : s0c), thr1 = max(1, c1 ? s1c : s0c);
    const int cap = G.cand_cap;
    for (unsigned q = lp0; q != lp; q += lstride4) {
        const uint32_t en = lds_u32(q);
        const uint32_t yx = yx0 + ((en & 0xfe00u) << 3);               // + row << 12
        const int sa = en & 0x1ffu, sb = en >> 16;
        if (sa >= thr0) {
            if (at < cap) cand[at] = yx + ((uint32_t)sa << 24);
            ++at;
        }
        if (sb >= thr1) {
            if (at < cap) cand[at] = yx + 1u + ((uint32_t)sb << 24);
            ++at;
        }
    }
}

// Rows of the job's tile from `src` (row pitch sp bytes, word index base wbase): a lane takes one aligned source word
// (4 pixels) and writes two whole tile words; when the cell starts at an odd column the pairs straddle source words
// and the missing byte comes from the previous lane.  SMEM_SRC: the source is the TMA-staged raw box in shared memory.
template <int CELL, bool SMEM_SRC, int UNR = 4>
__device__ __forceinline__ void ff_stage(const uint8_t *src, int sp, int x_org, int row_org, int row_max, int xmaxw,
                                         const FfJob &J, uint32_t *tile, int lane)
{
    using C = FfCfg<CELL>;
    const int xb = (J.x0 - 4) & ~3, ush = xb - (J.x0 - 4);          // tile pixel of the first byte of source word 0: -3..0
    const int srows = J.nrows + 6;
    constexpr int NW = 20;                                           // 72 pixels + 3 of misalignment (+1 spare)
    const bool odd = ush & 1;
    const int c = lane;                                              // lanes 0..19: the words of one row
    const int u = ush + 4 * c - (odd ? 1 : 0);                       // tile pixel of my first output pair (even)
    const int gw = min(((xb - x_org) >> 2) + c, xmaxw);
    uint32_t *d = tile + (u >> 1);
    const bool st0 = lane < NW && (unsigned)u < 72u, st1 = lane < NW && (unsigned)(u + 2) < 72u;
#pragma unroll UNR
    for (int r = 0; r < srows; ++r) {
        const int gy = min(J.y0 - 3 + r, row_max) - row_org;
        uint32_t v = 0;
        if (lane < NW) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(src + (long long)gy * sp) + gw;
            v = SMEM_SRC ? *q : __ldg(q);
        }
        const uint32_t pv = __shfl_up_sync(0xffffffffu, v, 1);
        // even shift: pairs (b0,b1),(b2,b3); odd shift: pairs (prev.b3,b0),(b1,b2)
        const uint32_t w_lo = odd ? (__byte_perm(pv, v, 0x0403) & 0x00ff00ffu) : __byte_perm(v, 0, 0x4140);
        const uint32_t w_hi = odd ? __byte_perm(v, 0, 0x4241) : __byte_perm(v, 0, 0x4342);
        if (st0) d[r * C::PITCH] = w_lo;
        if (st1) d[r * C::PITCH + 1] = w_hi;
    }
}

template <int CELL, int MINB = 4, int UNR = 4, bool BOUND_ONLY = false>
__global__ void __launch_bounds__(FfCfg<CELL>::WARPS * 32, MINB) k_fast_fused(const DevParams *__restrict__ P, Src0 s0, int work_off, int work_end)
{
    using C = FfCfg<CELL>;
    extern __shared__ __align__(16) uint32_t ff_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, frame = blockIdx.y;
    const int widx = work_off + blockIdx.x * C::WARPS + warp;
    if (widx >= work_end) return;                                      // warp-uniform; the kernel has no block barrier
    uint32_t *tile = ff_smem + (size_t)warp * C::WARP_WORDS;
    uint32_t *list = tile + C::ROWS * C::PITCH;
    const FfJob J = ff_job(P, widx, frame);
    int sp;
    const uint8_t *img = level_ptr(P, s0, frame, J.level, &sp);
    ff_stage<CELL, false, UNR>(img, sp, 0, 0, P->lv[J.level].h - 1, (sp >> 2) - 1, J, tile, lane);
    __syncwarp();
    ff_process<CELL, BOUND_ONLY>(P, J, tile, list, lane);
}

// ---- persistent TMA variant: every warp walks (job, frame) items; the raw pixel box of the NEXT item (96 bytes x
// h_cell+6 rows, 16-byte aligned origin) is fetched by cp.async.bulk.tensor while the current item is scored.
template <int CELL>
struct FfTmaCfg {
    using C = FfCfg<CELL>;
    static constexpr int RAW_BYTES = (kFfBoxW * C::ROWS + 127) / 128 * 128;
    static constexpr int WARP_BYTES = (RAW_BYTES + C::WARP_WORDS * 4 + 127) / 128 * 128;
    static constexpr int SMEM = C::WARPS * WARP_BYTES;
};

template <int CELL>
__global__ void __launch_bounds__(FfCfg<CELL>::WARPS * 32)
k_fast_fused_tma(const DevParams *__restrict__ P, const __grid_constant__ FastMaps maps, int work_off, int work_end, int nframes)
{
    using C = FfCfg<CELL>;
    using T = FfTmaCfg<CELL>;
    extern __shared__ __align__(128) uint8_t fft_smem[];
    __shared__ __align__(8) uint64_t s_bar[C::WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *raw = fft_smem + (size_t)warp * T::WARP_BYTES;
    uint32_t *tile = reinterpret_cast<uint32_t *>(raw + T::RAW_BYTES);
    uint32_t *list = tile + C::ROWS * C::PITCH;
    const int njobs = work_end - work_off, total = njobs * nframes;
    const int stride = gridDim.x * C::WARPS;
    int item = blockIdx.x * C::WARPS + warp;
    if (item >= total) return;                                         // warp-uniform; the kernel has no block barrier
    const unsigned bar = smem_u32(&s_bar[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](const FfJob &J) {                                 // lane 0
        const int rows = P->lv[J.level].h_cell + 6;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kFfBoxW * rows) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(raw)), "l"(reinterpret_cast<unsigned long long>(&maps.m[J.level])), "r"(bar),
                        "r"((J.x0 - 4) & ~15), "r"(J.y0 - 3), "r"(J.frame) : "memory");
    };
    FfJob J = ff_job(P, work_off + item % njobs, item / njobs);
    if (lane == 0) issue(J);
    unsigned phase = 0;
    for (;;) {
        mbar_wait_parity(&s_bar[warp], phase);
        phase ^= 1u;
        // rows past the image come back as zeros; they only feed pixels outside the detection area
        ff_stage<CELL, true>(raw, kFfBoxW, (J.x0 - 4) & ~15, J.y0 - 3, 0x7fffffff, kFfBoxW / 4 - 1, J, tile, lane);
        __syncwarp();                                                  // raw is consumed, the tile is complete
        const int next = item + stride;
        FfJob Jn = J;
        if (next < total) {
            Jn = ff_job(P, work_off + next % njobs, next / njobs);
            if (lane == 0) issue(Jn);
        }
        ff_process<CELL>(P, J, tile, list, lane);
        if (next >= total) break;
        __syncwarp();
        item = next; J = Jn;
    }
}

template <int CELL>
static cudaError_t launch_fast_tma(const DevParams *dP, const FastMaps &maps, int work_off, int work_end, int nframes, cudaStream_t st)
{
    using C = FfCfg<CELL>;
    using T = FfTmaCfg<CELL>;
    cudaError_t e = cudaFuncSetAttribute(k_fast_fused_tma<CELL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
    if (e != cudaSuccess) return e;
    const int per_sm = (227 * 1024) / (T::SMEM + 1024);
    const int total = (work_end - work_off) * nframes;
    int grid = 148 * (per_sm < 1 ? 1 : per_sm);
    if (grid > (total + C::WARPS - 1) / C::WARPS) grid = (total + C::WARPS - 1) / C::WARPS;
    k_fast_fused_tma<CELL><<<grid, C::WARPS * 32, T::SMEM, st>>>(dP, maps, work_off, work_end, nframes);
    return cudaGetLastError();
}

cudaError_t launch_fast(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, int n_small, cudaStream_t st, LaunchStats *ls,
                        const FastMaps *maps)
{
    if (hP.n_ffast_work == 0) return cudaSuccess;
    if (maps) {
        int max_cell = 0;
        for (int l = 0; l < hP.nlevels; ++l) max_cell = std::max(max_cell, std::max(hP.lv[l].h_cell, hP.lv[l].w_cell));
        if (max_cell <= 40 && n_small == hP.n_ffast_work) {
            // cells of at most 40 rows: raw box + 16-bit tile + lists take 13.75 KB per warp, so four CTAs (16 warps) stay resident
            cudaError_t e = launch_fast_tma<40>(dP, *maps, 0, n_small, nframes, st);
            if (e != cudaSuccess) return e;
            ls->launches++;
            return cudaSuccess;
        }
        if (n_small > 0) {
            cudaError_t e = launch_fast_tma<44>(dP, *maps, 0, n_small, nframes, st);
            if (e != cudaSuccess) return e;
            ls->launches++;
        }
        if (hP.n_ffast_work > n_small) {
            cudaError_t e = launch_fast_tma<64>(dP, *maps, n_small, hP.n_ffast_work, nframes, st);
            if (e != cudaSuccess) return e;
            ls->launches++;
        }
        return cudaSuccess;
    }
    if (n_small > 0) {
        using C = FfCfg<44>;
        const dim3 grid((n_small + C::WARPS - 1) / C::WARPS, nframes);
        // 4 CTAs / SM (104 registers) and a 12-deep staging unroll measured best; 5 CTAs at 96 registers, shallower unrolls,
        // persistent CTAs and the TMA variant above were all slower
        // experiment knob: extra dynamic shared memory per CTA caps the FAST CTAs per SM and leaves registers / shared memory to the
        // other handles' kernels (k_blur, k_resize_sep fit beside three of them).  Measured slower: 16 KB (3 CTAs / SM) 0.311 vs 0.291 ms.
        static const int pad = debug_knob("ORBX_FAST_PAD_KB", 0) * 1024;
#ifdef ORBX_DEBUG_KNOBS
        if (debug_knob("ORBX_FAST_BOUND", 0)) {                        // timing experiment only: results are not the reference's
            cudaFuncSetAttribute(k_fast_fused<44, 4, 12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + pad);
            k_fast_fused<44, 4, 12, true><<<grid, C::WARPS * 32, C::SMEM + pad, st>>>(dP, s0, 0, n_small);
        } else
#endif
        {
        cudaFuncSetAttribute(k_fast_fused<44, 4, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + pad);
        k_fast_fused<44, 4, 12><<<grid, C::WARPS * 32, C::SMEM + pad, st>>>(dP, s0, 0, n_small);
        }
        ls->launches++;
    }
    if (hP.n_ffast_work > n_small) {
        using C = FfCfg<64>;
        cudaFuncSetAttribute(k_fast_fused<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        k_fast_fused<64><<<dim3((hP.n_ffast_work - n_small + C::WARPS - 1) / C::WARPS, nframes), C::WARPS * 32, C::SMEM, st>>>(dP, s0, n_small, hP.n_ffast_work);
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

typedef unsigned long long u64;

template <int THREADS>
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
        int s = lane < THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    *total = warp_sums[31];
    return x + (warp > 0 ? warp_sums[warp - 1] : 0);
}

template <int THREADS>
__device__ __forceinline__ void bitonic_sort_desc(u64 *v, int n_pow2)
{
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));       // lower index of the pair
                const int p = i | j;
                const u64 a = v[i], b = v[p];
                const bool down = (i & k) == 0;                             // descending run
                if ((a < b) == down) { v[i] = b; v[p] = a; }
            }
            __syncthreads();
        }
    }
}

// path code of one packed candidate: initial node, then the DivideNode child index per split.
// x and y halve independently, so the code is the OR of two host-built tables (host_tables.cpp).
__device__ __forceinline__ uint32_t path_code(uint32_t c, const uint32_t *lut_x, const uint32_t *lut_y)
{
    return lut_x[c & 0xfff] | lut_y[(c >> 12) & 0xfff];              // the tables are in shared memory when they fit (k_octree)
}

// The three child boundaries of a node at once: first indices in [lo,hi) whose 2-bit child field at `shift` is >= 1, 2, 3.  The three
// binary searches run in lockstep over the whole range (fixed trip count), so their shared-memory loads overlap.
__device__ __forceinline__ void lower_children(const uint32_t *codes, int lo, int hi, int shift, int *b1, int *b2, int *b3)
{
    int l1 = lo, h1 = hi, l2 = lo, h2 = hi, l3 = lo, h3 = hi;
    for (int steps = 32 - __clz(hi - lo); steps > 0; --steps) {     // a range of n candidates needs at most floor(log2 n) + 1 halvings
        const int m1 = (l1 + h1) >> 1, m2 = (l2 + h2) >> 1, m3 = (l3 + h3) >> 1;
        const unsigned c1 = l1 < h1 ? (codes[m1] >> shift) & 3u : 0u, c2 = l2 < h2 ? (codes[m2] >> shift) & 3u : 0u,
                       c3 = l3 < h3 ? (codes[m3] >> shift) & 3u : 0u;
        if (l1 < h1) { if (c1 >= 1u) h1 = m1; else l1 = m1 + 1; }
        if (l2 < h2) { if (c2 >= 2u) h2 = m2; else l2 = m2 + 1; }
        if (l3 < h3) { if (c3 >= 3u) h3 = m3; else l3 = m3 + 1; }
    }
    *b1 = l1; *b2 = l2; *b3 = l3;
}

struct OctreeSmem { size_t bytes; int key_cap, skey_cap; };

static OctreeSmem octree_smem(int threads, int node_cap, int max_feat, size_t budget, int lut_cap = 0)
{
    OctreeSmem o;
    o.skey_cap = 1;
    while (o.skey_cap < max_feat + 3) o.skey_cap <<= 1;
    const size_t fixed = (size_t)o.skey_cap * 8 + (size_t)node_cap * 16 + (size_t)(threads / 32) * 256 * 4 + (size_t)lut_cap * 4;
    long long kc = ((long long)budget - (long long)fixed) / 8;
    kc = kc < 256 ? 256 : kc;
    o.key_cap = (int)(kc & ~31LL);
    o.bytes = fixed + (size_t)o.key_cap * 8;
    return o;
}

// shared-memory budget of the 256-thread octree CTA (experiment knob ORBX_OCT_BUDGET_KB)
static size_t oct_small_budget()
{
    static const size_t b = (size_t)debug_knob("ORBX_OCT_BUDGET_KB", 100) * 1024;
    return b;
}

size_t octree_smem_bytes(int max_node_cap, int max_feat, int *key_cap)
{
    const int threads = max_node_cap <= 1024 ? 256 : (max_node_cap <= 4096 ? 512 : 1024);
    const OctreeSmem o = octree_smem(threads, max_node_cap, max_feat, threads == 256 ? oct_small_budget() : 200 * 1024);
    if (key_cap) *key_cap = o.key_cap;
    return o.bytes;
}

// CLUSTER > 1 (levels of very large images, ~10^5 candidates): a thread-block cluster of CLUSTER CTAs shares one (frame, level).
// The radix sort -- two thirds of such a level's time in a single CTA -- and the path-code pass are split over the cluster's CTAs
// (every CTA owns a contiguous slice of the keys per pass; the per-warp digit counts are exchanged through the L2-resident
// scratch and ordered by cluster barriers, so the sort stays the same stable LSD sort); CTA 0 then builds the tree alone.
// MODE 0: the whole thing in one launch.  MODE 1: sort + path codes only (the clustered launch of 4K-class images).  MODE 2: the tree
// on keys / codes a MODE 1 launch left sorted in the scratch -- a plain launch with a small shared-memory footprint, so that the
// octree of one batch shares the SMs with the other batches' kernels instead of pinning idle cluster CTAs on them.
template <int THREADS, int IPT, int CLUSTER = 1, int MODE = 0, int MINB = ((CLUSTER > 1 || MODE == 2) ? 2 : 1)>
__global__ void __launch_bounds__(THREADS, MINB)
k_octree(const DevParams *__restrict__ P, int node_cap, int skey_cap, int key_cap, int level_off, int lut_cap, int nframes_lm)
{
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) uint8_t oct_smem[];
    __shared__ int warp_sums[32];
    __shared__ int s_m, s_added;

    const int crank = CLUSTER > 1 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
    // plain launches: grid (level, frame).  Clustered launches: a 1-D grid of clusters in level-major order (all frames' level 0
    // first, the longest-running ones), so that the clusters that wait for a free slot are the short ones
    const int cid = CLUSTER > 1 ? (int)blockIdx.x / CLUSTER : 0;
    const int level = (CLUSTER > 1 ? cid / nframes_lm : (int)blockIdx.x) + level_off, frame = CLUSTER > 1 ? cid % nframes_lm : (int)blockIdx.y;
    const LevelGeom &G = P->lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)min(P->cand_count[frame * P->nlevels + level], (unsigned)G.cand_cap);
    const int N = G.n_feat, D = G.depth;
    const uint32_t *glut = P->oct_lut + G.lut_off;                  // x codes [region_w], y codes [region_h], then the cell-order tables
    const uint32_t *lut_x = glut, *lut_y = glut + G.region_w;

    u64 *skey = reinterpret_cast<u64 *>(oct_smem);                 // careful-phase sort buffer
    int *nlo = reinterpret_cast<int *>(skey + skey_cap);
    int *nhi = nlo + node_cap;
    int *ndep = nhi + node_cap;
    int *eidx = ndep + node_cap;
    uint32_t *hist = reinterpret_cast<uint32_t *>(eidx + node_cap);  // [WARPS][256]
    uint32_t *bufA = hist + WARPS * 256, *bufB = bufA + key_cap;
    uint32_t *slut = bufB + key_cap;                                // path-code tables of this level, when they fit
    uint32_t *ghist = nullptr;                                      // CLUSTER > 1: per-warp digit counts of every CTA of the cluster
    if (n > key_cap || CLUSTER > 1 || MODE == 2) {                  // level too dense for shared memory: L2-resident scratch
        bufA = reinterpret_cast<uint32_t *>(P->sort_scratch + ((long long)frame * P->cand_frame_elems + G.cand_off) * 2);
        bufB = bufA + G.cand_cap;
        ghist = bufB + G.cand_cap;                                  // the slot holds 16 bytes per candidate: 8 * cand_cap bytes are free
    }
    const uint32_t *cand = P->cand + (long long)frame * P->cand_frame_elems + G.cand_off;
    uint32_t *stage = P->kp_stage + (long long)frame * P->kp_frame_cap + G.kp_off;
    uint32_t *kp_count = P->kp_count + frame * P->nlevels + level;

    if (n == 0) { if (tid == 0 && crank == 0) *kp_count = 0; return; }       // uniform over the cluster: nobody waits at a barrier
#ifdef ORBX_OCT_TIMING
    long long tk[8], acc[5] = {0, 0, 0, 0, 0}, tl = 0; int ntk = 0, nsweeps = 0, ncareful = 0;
#define OCT_LAP(i) do { if (tid == 0) { const long long t_ = clock64(); acc[i] += t_ - tl; tl = t_; } } while (0)
#define OCT_TICK() do { if (tid == 0 && ntk < 8) tk[ntk++] = clock64(); } while (0)
#else
#define OCT_TICK() do { } while (0)
#define OCT_LAP(i) do { } while (0)
#endif
    OCT_TICK();

    int code_bits = 2 * D;
    for (int t = G.n_ini - 1; t > 0; t >>= 1) ++code_bits;
    if (MODE == 2) {                                                // sorted by a MODE 1 launch: an odd number of passes leaves the keys in the second buffer
        if (((code_bits + 7) / 8) & 1) { uint32_t *t = bufA; bufA = bufB; bufB = t; }
    } else {
    // ---- LSD radix sort of the packed candidates by path code, 8 bits per pass (stable)
    if (G.region_w + G.region_h <= lut_cap) {                       // every pass looks both tables up per candidate: keep them close
        const int nl = G.region_w + G.region_h;
        for (int base = tid; base < nl; base += 4 * THREADS) {      // four loads in flight per thread
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = base + u * THREADS < nl ? __ldg(glut + base + u * THREADS) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (base + u * THREADS < nl) slut[base + u * THREADS] = v[u];
        }
        lut_x = slut; lut_y = slut + G.region_w;
    }
    constexpr int CW = WARPS * CLUSTER;                             // warps that share the sort; global warp gw owns the gw-th slice of every pass
    const int gw = crank * WARPS + warp;
    for (int base = crank * THREADS + tid; base < n; base += 8 * THREADS * CLUSTER) {           // eight loads in flight per thread
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = base + u * THREADS * CLUSTER < n ? __ldg(cand + base + u * THREADS * CLUSTER) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (base + u * THREADS * CLUSTER < n) bufA[base + u * THREADS * CLUSTER] = v[u];
    }
    if (CLUSTER > 1) { __threadfence(); cooperative_groups::this_cluster().sync(); }
    const int chunk = (((n + CW - 1) / CW) + 31) & ~31;
    const int c_lo = min(gw * chunk, n), c_hi = min(c_lo + chunk, n);
    const unsigned lt = lanemask_lt();
    for (int shift = 0; shift < code_bits; shift += 8) {
        for (int i = tid; i < WARPS * 256; i += THREADS) hist[i] = 0;
        __syncthreads();
        uint32_t *wh = hist + warp * 256;
        for (int base = c_lo + lane; base < c_hi; base += 256) {     // counts need no order: shared-memory atomics, eight loads in flight
            uint32_t k8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) k8[u] = base + 32 * u < c_hi ? bufA[base + 32 * u] : 0u;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + 32 * u < c_hi) atomicAdd(wh + ((path_code(k8[u], lut_x, lut_y) >> shift) & 255u), 1u);
        }
        __syncthreads();
        if (CLUSTER > 1) {
            // every CTA publishes its warps' counts, then derives the start of each of ITS warps' runs: digit major, then global warp
            for (int i = tid; i < WARPS * 256; i += THREADS) ghist[crank * WARPS * 256 + i] = hist[i];
            __threadfence();
            cooperative_groups::this_cluster().sync();
            int total_d = 0, before_cta = 0;
            if (tid < 256) {
                for (int g2 = 0; g2 < CW; ++g2) {
                    const int c = (int)__ldcg(ghist + g2 * 256 + tid);
                    if (g2 == crank * WARPS) before_cta = total_d;
                    total_d += c;
                }
            }
            int tot;
            const int digit_base = block_scan_incl<THREADS>(tid < 256 ? total_d : 0, warp_sums, &tot) - total_d;
            if (tid < 256) {
                int run = digit_base + before_cta;
                for (int w2 = 0; w2 < WARPS; ++w2) { const int c = (int)hist[w2 * 256 + tid]; hist[w2 * 256 + tid] = (uint32_t)run; run += c; }
            }
        } else
        {   // exclusive scan over (digit major, warp minor)
            constexpr int HPT = 256 * WARPS / THREADS;               // == 8
            uint32_t v[HPT];
            int sum = 0;
#pragma unroll
            for (int k = 0; k < HPT; ++k) { const int e = tid * HPT + k; v[k] = hist[(e % WARPS) * 256 + e / WARPS]; sum += (int)v[k]; }
            int tot;
            int run = block_scan_incl<THREADS>(sum, warp_sums, &tot) - sum;
#pragma unroll
            for (int k = 0; k < HPT; ++k) { const int e = tid * HPT + k; hist[(e % WARPS) * 256 + e / WARPS] = (uint32_t)run; run += (int)v[k]; }
        }
        __syncthreads();
        // stable scatter, 32 candidates at a time in order; four groups' keys and digits are fetched ahead so that only the
        // short offset chain through wh[] is sequential
        for (int base = c_lo; base < c_hi; base += 128) {
            uint32_t key[4]; unsigned dg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + 32 * u + lane;
                key[u] = i < c_hi ? bufA[i] : 0u;
                dg[u] = (path_code(key[u], lut_x, lut_y) >> shift) & 255u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool valid = base + 32 * u + lane < c_hi;
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                if (vm == 0) break;                                  // warp-uniform
                unsigned peers = 0; uint32_t off = 0;
                if (valid) {
                    peers = __match_any_sync(vm, dg[u]);
                    off = wh[dg[u]];
                }
                __syncwarp();
                if (valid) {
                    if ((peers & lt) == 0) wh[dg[u]] = off + __popc(peers);
                    bufB[off + __popc(peers & lt)] = key[u];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (CLUSTER > 1) { __threadfence(); cooperative_groups::this_cluster().sync(); }
        uint32_t *t = bufA; bufA = bufB; bufB = t;
    }
    OCT_TICK();
    for (int i = crank * THREADS + tid; i < n; i += THREADS * CLUSTER) bufB[i] = path_code(bufA[i], lut_x, lut_y);
    __syncthreads();
    if (CLUSTER > 1) {
        __threadfence();
        cooperative_groups::this_cluster().sync();
        if (crank != 0) return;                                     // the tree is built by CTA 0 of the cluster
    }
    }
    if (MODE == 1) return;                                          // keys (bufA) and codes (bufB) stay in the scratch for the MODE 2 launch
    uint32_t *keys = bufA, *codes = bufB;                           // sorted candidates and their path codes
    OCT_TICK();

    // ---- roots, reverse list order (:553-585): thread r finds root n_ini-1-r's range, thread 0 drops the empty ones
    if (tid < G.n_ini && tid < node_cap) {
        const int r = G.n_ini - 1 - tid;
        int lo = 0, hi = n;
        while (lo < hi) { const int m = (lo + hi) >> 1; if ((codes[m] >> (2 * D)) >= (unsigned)r) hi = m; else lo = m + 1; }
        const int a = lo;
        hi = n;
        while (lo < hi) { const int m = (lo + hi) >> 1; if ((codes[m] >> (2 * D)) >= (unsigned)(r + 1)) hi = m; else lo = m + 1; }
        reinterpret_cast<int *>(skey)[2 * tid] = a; reinterpret_cast<int *>(skey)[2 * tid + 1] = lo;
    }
    __syncthreads();
    if (tid == 0) {
        int na = 0;
        for (int t = 0; t < G.n_ini; ++t) {
            const int a = reinterpret_cast<int *>(skey)[2 * t], b = reinterpret_cast<int *>(skey)[2 * t + 1];
            if (b > a) { nlo[na] = a; nhi[na] = b; ndep[na] = 0; ++na; }
        }
        int ne = 0;
        for (int i = 0; i < na; ++i) if (nhi[i] - nlo[i] > 1) eidx[ne++] = i;
        s_m = na; s_added = ne;
    }
    __syncthreads();
    int size = s_m, nE = s_added;
    __syncthreads();

    bool careful = false;
    OCT_TICK();
    for (;;) {
#ifdef ORBX_OCT_TIMING
        ++nsweeps; ncareful += careful;
        if (tid == 0) tl = clock64();
#endif
        const int prev_size = size;
        // ---- visit order
        if (careful) {
            int sp2 = 1;
            while (sp2 < nE) sp2 <<= 1;
            for (int i = tid; i < sp2; i += THREADS)
                skey[i] = i < nE ? ((u64)(unsigned)(nhi[eidx[i]] - nlo[eidx[i]]) << 32 | (unsigned)eidx[i]) : 0ull;
            __syncthreads();
            bitonic_sort_desc<THREADS>(skey, sp2);
        }
        OCT_LAP(0);
        // ---- pass 1: children of the parents this thread owns (visit ranks tid*IPT ..)
        // visit ranks tid*ipt .. tid*ipt + ipt-1 with ipt = ceil(nE / THREADS) <= IPT: the first sweeps (a handful of parents with
        // long ranges) spread over the threads instead of queueing IPT deep on the first few
        const int ipt = min(IPT, max(1, (nE + THREADS - 1) / THREADS));
        int pb[IPT][3];
        int pc[IPT];
        int csum = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int r = tid * ipt + k;
            pc[k] = 0;
            if (k < ipt && r < nE) {
                const int p = careful ? (int)(skey[r] & 0xffffffffu) : eidx[nE - 1 - r];
                const int lo = nlo[p], hi = nhi[p], shift = 2 * (D - 1 - ndep[p]);
                int b1, b2, b3;
                lower_children(codes, lo, hi, shift, &b1, &b2, &b3);
                pb[k][0] = b1; pb[k][1] = b2; pb[k][2] = b3;
                pc[k] = (b1 > lo) + (b2 > b1) + (b3 > b2) + (hi > b3);
                csum += pc[k];
            }
        }
        OCT_LAP(1);
        int total_c;
        const int incl = block_scan_incl<THREADS>(csum, warp_sums, &total_c);
        // ---- cutoff m (careful phase: smallest m with size + sum_{i<m}(c_i - 1) >= N)
        if (tid == 0) { s_m = nE; s_added = total_c; }
        __syncthreads();
        if (careful) {
            int run = incl - csum;                                  // children of all earlier ranks
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int r = tid * ipt + k;
                if (k < ipt && r < nE) {
                    const int before = size + run - r;              // size + sum_{i<r}(c_i-1)
                    run += pc[k];
                    const int after = size + run - (r + 1);
                    if (after >= N && before < N) { s_m = r + 1; s_added = run; }
                }
            }
            __syncthreads();
        }
        const int m = s_m, added = s_added;
        OCT_LAP(2);
        // ---- pass 2: append children in visit order, retire the parents
        {
            int off = incl - csum;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int r = tid * ipt + k;
                if (k < ipt && r < nE && r < m) {
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
        OCT_LAP(3);
        // ---- stable compaction of the array + new expandable list
        {
            const int ipt2 = min(IPT, max(1, (n_pre + THREADS - 1) / THREADS));
            int rl[IPT], rh[IPT], rd[IPT];
            int alive = 0, multi = 0;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int i = tid * ipt2 + k;
                rd[k] = -1;
                if (k < ipt2 && i < n_pre) {
                    rl[k] = nlo[i]; rh[k] = nhi[i]; rd[k] = ndep[i];
                    if (rd[k] >= 0) { ++alive; multi += (rh[k] - rl[k] > 1); }
                }
            }
            int tot;
            const int inc2 = block_scan_incl<THREADS>(alive | multi << 16, warp_sums, &tot);
            int wa = (inc2 & 0xffff) - alive, we = (inc2 >> 16) - multi;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                if (rd[k] >= 0) {
                    nlo[wa] = rl[k]; nhi[wa] = rh[k]; ndep[wa] = rd[k];
                    if (rh[k] - rl[k] > 1) eidx[we++] = wa;
                    ++wa;
                }
            }
            nE = tot >> 16;
            __syncthreads();
        }
        OCT_LAP(4);
        // ---- termination (:667-737)
        if (size >= N || size == prev_size) break;
        if (!careful && size + 3 * nE > N) careful = true;
    }

    OCT_TICK();
    // ---- winners, list order front->back == reverse array order (:740-760): greatest response, first in the
    //      reference's candidate order (cell row, cell column, y, x).  One thread per node.
    const uint32_t *lut_cx = glut + G.region_w + G.region_h, *lut_cy = lut_cx + G.region_w;   // global copies (lut_x / lut_y may be the shared ones)
    if (n >= 16 * size) {
        // many candidates per node (large images): a warp per node, lanes stride over its candidates
        for (int j = warp; j < size; j += WARPS) {
            const int nd = size - 1 - j;
            u64 best = 0;
            for (int i = nlo[nd] + lane; i < nhi[nd]; i += 32) {
                const uint32_t c = keys[i];
                const uint32_t x = c & 0xfff, y = (c >> 12) & 0xfff;
                const u64 order = (u64)(__ldg(lut_cy + y) + __ldg(lut_cx + x)) << 24 | (c & 0xffffffu);
                const u64 v = (u64)(c >> 24) << 40 | (~order & 0xffffffffffull);
                best = v > best ? v : best;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
            const u64 order = ~best & 0xffffffffffull;
            if (lane == 0) stage[j] = (uint32_t)(order & 0xffffff) | (uint32_t)(best >> 40) << 24;
        }
    } else {
        for (int j = tid; j < size; j += THREADS) {
            const int nd = size - 1 - j;
            u64 best = 0;
            for (int i = nlo[nd]; i < nhi[nd]; ++i) {
                const uint32_t c = keys[i];
                const uint32_t x = c & 0xfff, y = (c >> 12) & 0xfff;
                const u64 order = (u64)(__ldg(lut_cy + y) + __ldg(lut_cx + x)) << 24 | (c & 0xffffffu);
                const u64 v = (u64)(c >> 24) << 40 | (~order & 0xffffffffffull);
                best = v > best ? v : best;
            }
            const u64 order = ~best & 0xffffffffffull;
            stage[j] = (uint32_t)(order & 0xffffff) | (uint32_t)(best >> 40) << 24;
        }
    }
    if (tid == 0) *kp_count = (uint32_t)size;
#ifdef ORBX_OCT_TIMING
    __syncthreads();
    OCT_TICK();
    if (tid == 0 && frame == 0)
        printf("octree level %d n %d N %d size %d sweeps %d careful %d | sort %lld codes %lld roots %lld sweeps %lld (order %lld pass1 %lld scan %lld pass2 %lld compact %lld) winners %lld cycles\n", level, n, N, size,
               nsweeps, ncareful, tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], acc[0], acc[1], acc[2], acc[3], acc[4], tk[5] - tk[4]);
#endif
#undef OCT_TICK
#undef OCT_LAP
}

template <int THREADS, int IPT, int CLUSTER = 1, int MODE = 0, int MINB = ((CLUSTER > 1 || MODE == 2) ? 2 : 1)>
static cudaError_t launch_octree_t(const DevParams *dP, const DevParams &hP, int nframes, int node_cap, int max_feat, size_t budget, cudaStream_t st,
                                   int level_off = 0, int level_cnt = -1)
{
    if (level_cnt < 0) level_cnt = hP.nlevels - level_off;
    if (level_cnt <= 0) return cudaSuccess;
    int lut_cap = 0;                                                 // path-code tables of the largest launched level: x codes + y codes
    for (int l = level_off; l < level_off + level_cnt; ++l) lut_cap = std::max(lut_cap, hP.lv[l].region_w + hP.lv[l].region_h);
    if ((size_t)lut_cap * 4 > budget / 4 || CLUSTER > 1 || MODE == 2) lut_cap = 0;   // too large for this budget (or a launch that keeps
                                                                     // several CTAs per SM): the kernel reads them from global memory (L1-resident)
    if (MODE == 1) { node_cap = 0; max_feat = -3; }                  // sort only: no node arrays, no careful-phase buffer, just the digit counters
    const OctreeSmem o = octree_smem(THREADS, node_cap, max_feat, budget, lut_cap);
    {                                                                // the attribute is a per-device maximum: cheap, set every time
        cudaError_t e = cudaFuncSetAttribute(k_octree<THREADS, IPT, CLUSTER, MODE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptInSmem);
        if (e != cudaSuccess) return e;
    }
    if (CLUSTER > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(level_cnt * nframes * CLUSTER); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = o.bytes; cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = CLUSTER; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_octree<THREADS, IPT, CLUSTER, MODE, MINB>, dP, node_cap, o.skey_cap, o.key_cap, level_off, lut_cap, nframes);
    }
    k_octree<THREADS, IPT, CLUSTER, MODE, MINB><<<dim3(level_cnt, nframes), THREADS, o.bytes, st>>>(dP, node_cap, o.skey_cap, o.key_cap, level_off, lut_cap, nframes);
    return cudaGetLastError();
}

cudaError_t launch_octree(const DevParams *dP, const DevParams &hP, int nframes, int max_node_cap, int max_feat, cudaStream_t st, LaunchStats *ls)
{
    ls->launches++;
    // CTA size by node-array capacity (THREADS*IPT records); small CTAs leave room for 2+ per SM
    if (max_node_cap <= 1024) {
        // Batches: the small levels first with a small shared-memory footprint, then the big ones.  Either launch has
        // about one CTA per SM, so the latency-bound octree leaves room for the other handles' kernels on every SM
        // (one launch of all levels at 100 KB per CTA pins two CTAs = 200 KB on most SMs for ~100 us).
        int split = 0;                                                // first level whose image has < 40 % of level 0's pixels
        const long long p0 = (long long)hP.lv[0].w * hP.lv[0].h;
        while (split < hP.nlevels && (long long)hP.lv[split].w * hP.lv[split].h * 5 >= p0 * 2) ++split;
        static const bool no_split = debug_knob("ORBX_OCT_NOSPLIT", 0) != 0;
        if (nframes >= 8 && split > 0 && split < hP.nlevels && !no_split) {
            ls->launches++;
            cudaError_t e = launch_octree_t<256, 4>(dP, hP, nframes, max_node_cap, max_feat, oct_small_budget() / 2 + 4096, st, split, hP.nlevels - split);
            if (e != cudaSuccess) return e;
            return launch_octree_t<256, 4>(dP, hP, nframes, max_node_cap, max_feat, oct_small_budget(), st, 0, split);
        }
        return launch_octree_t<256, 4>(dP, hP, nframes, max_node_cap, max_feat, oct_small_budget(), st);
    }
    // Large images (3840x2160: ~10^5 candidates on level 0, far beyond shared memory) sort through the L2-resident scratch and are
    // bound by the CTA's own throughput; with one CTA per (frame, level) there are fewer CTAs than SMs, so each gets a whole SM:
    // 1024 threads instead of 512 (debug knob ORBX_OCT_THREADS to compare)
    // 4K-class images (3840x2160: ~10^5 candidates on level 0, far beyond shared memory): the levels whose candidates cannot fit
    // a CTA's shared memory are sorted by a cluster of four CTAs each (k_octree<.., 4>), launched first; the small levels follow
    // as single CTAs.  ORBX_OCT_CLUSTER=0 (debug-knob build) keeps everything on single 1024-thread CTAs for comparison.
    int min_cap = hP.lv[0].cand_cap;
    for (int l = 1; l < hP.nlevels; ++l) min_cap = std::min(min_cap, hP.lv[l].cand_cap);
    // (the cluster's digit counts, 4 CTAs x 16 warps x 256 counters, live in the free half of a level's 16-byte-per-candidate scratch slot:
    // every level must have room for them, else the single-CTA path below takes the whole image)
    if (max_node_cap <= 4096 && hP.lv[0].cand_cap > 1000000 && (size_t)min_cap * 8 >= (size_t)4 * 16 * 256 * 4 && debug_knob("ORBX_OCT_CLUSTER", 1)) {
        // a clustered sort launch for all levels (level-major grid: the long sorts start first), then the tree launch on the sorted scratch
        cudaError_t e = launch_octree_t<512, 8, 4, 1>(dP, hP, nframes, max_node_cap, max_feat, 0, st, 0, hP.nlevels);
        if (e != cudaSuccess) return e;
        ls->launches++;
        return launch_octree_t<512, 8, 1, 2>(dP, hP, nframes, max_node_cap, max_feat, 0, st, 0, hP.nlevels);
    }
    const int big = debug_knob("ORBX_OCT_THREADS", hP.lv[0].cand_cap > 1000000 ? 1024 : 512);
    if (max_node_cap <= 4096 && big != 1024) {
        // 1920x1080-class images: 512 CTAs per 64-frame batch.  Two CTAs per SM (64 registers, ~105 KB of shared memory: the keys of the
        // big levels then sort through the L2 scratch, which costs the sort nothing -- it is bound by its own scatter chain, not by
        // where the keys live) instead of one with everything in 200 KB: fewer waves, and room for the other batches' kernels
        if (debug_knob("ORBX_OCT_TWO_PER_SM", 1) && nframes * hP.nlevels > 148)
            return launch_octree_t<512, 8, 1, 0, 2>(dP, hP, nframes, max_node_cap, max_feat, 105 * 1024, st);
        return launch_octree_t<512, 8>(dP, hP, nframes, max_node_cap, max_feat, 200 * 1024, st);
    }
    return launch_octree_t<1024, 8>(dP, hP, nframes, max_node_cap, max_feat, 200 * 1024, st);
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


// cos / sin of a float angle in [0, 2*pi], each correctly rounded to float in all but ~1e-8 of the cases (the same
// as rounding CUDA's double cos()/sin()): one Cody-Waite reduction by pi/2 and two Taylor polynomials in fp64,
// shared between the two results.  The reference calls cosf/sinf of glibc, which round the same way (:112-113).
__device__ __forceinline__ void sincos_f32_via_f64(float ang, float *sn, float *cs)
{
    const double x = (double)ang;
    const int q = __double2int_rn(x * 0.63661977236758134308);
    const double qd = (double)q;
    double r = fma(qd, -1.57079632679489655800e+00, x);
    r = fma(qd, -6.12323399573676603587e-17, r);
    const double r2 = r * r;
    double ps = -1.0 / 355687428096000.0;                          // sin: r - r^3/3! + ... - r^19-free tail below 1e-19
    ps = fma(ps, r2, 1.0 / 1307674368000.0);
    ps = fma(ps, r2, -1.0 / 6227020800.0);
    ps = fma(ps, r2, 1.0 / 39916800.0);
    ps = fma(ps, r2, -1.0 / 362880.0);
    ps = fma(ps, r2, 1.0 / 5040.0);
    ps = fma(ps, r2, -1.0 / 120.0);
    ps = fma(ps, r2, 1.0 / 6.0);
    const double sr = fma(-ps * r2, r, r);
    double pc = 1.0 / 20922789888000.0;                            // cos: 1 - r^2/2! + ... + r^16/16!
    pc = fma(pc, r2, -1.0 / 87178291200.0);
    pc = fma(pc, r2, 1.0 / 479001600.0);
    pc = fma(pc, r2, -1.0 / 3628800.0);
    pc = fma(pc, r2, 1.0 / 40320.0);
    pc = fma(pc, r2, -1.0 / 720.0);
    pc = fma(pc, r2, 1.0 / 24.0);
    pc = fma(pc, r2, -0.5);
    const double cr = fma(pc, r2, 1.0);
    const bool swap = q & 1;
    double s_ = swap ? cr : sr, c_ = swap ? sr : cr;
    if (q & 2) s_ = -s_;
    if ((q + 1) & 2) c_ = -c_;
    *sn = (float)s_; *cs = (float)c_;
}

constexpr int kOdWarps = 8;
constexpr int kOdPitchW = 10;                     // patch pitch in words (37 + up to 3 bytes of misalignment)

__global__ void __launch_bounds__(kOdWarps * 32) k_orient_desc(const DevParams *__restrict__ P, Src0 s0)
{
    __shared__ int8_t spat[1024];                 // pattern, transposed: component e (0..31) of byte-lane l at [e*32 + l]
    __shared__ int s_prefix[kMaxLevels + 1], s_count[kMaxLevels], s_kpoff[kMaxLevels + 1];
    __shared__ uint32_t spatch[kOdWarps][37 * kOdPitchW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame = blockIdx.y, L = P->nlevels;
    for (int i = threadIdx.x; i < 1024; i += kOdWarps * 32) spat[(i & 31) * 32 + (i >> 5)] = P->pattern[i];
    if (threadIdx.x < 32) {                       // keypoints per level of this frame and their running offsets
        int c = lane < L ? min((int)P->kp_count[frame * L + lane], P->lv[lane].kp_cap) : 0;
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < L) { s_count[lane] = c; s_prefix[lane] = x - c; s_kpoff[lane] = P->lv[lane].kp_off; }
        if (lane == L - 1) { s_prefix[L] = x; s_kpoff[L] = 0x7fffffff; }
    }
    __syncthreads();
    const int slot = blockIdx.x * kOdWarps + warp;
    if (slot == 0 && lane == 0) P->out_n[frame] = s_prefix[L];
    // level of this staging slot (slots are level-major with a fixed capacity per level)
    int level = 0;
    while (slot >= s_kpoff[level + 1]) ++level;
    const int rel = slot - s_kpoff[level];
    if (rel >= s_count[level]) return;
    const int out_idx = s_prefix[level] + rel;

    const LevelGeom &G = P->lv[level];
    const uint32_t c = P->kp_stage[(long long)frame * P->kp_frame_cap + slot];
    const int cx = (int)(c & 0xfff) + kMinBorder, cy = (int)((c >> 12) & 0xfff) + kMinBorder;
    uint32_t *patch = spatch[warp];
    const uint8_t *pb = reinterpret_cast<const uint8_t *>(patch);
    // ---- IC_Angle: stage the 31x31 neighbourhood with aligned word loads, then one column per lane
    {
        int sp;
        const uint8_t *img = level_ptr(P, s0, frame, level, &sp);
        const int xa = (cx - 15) & ~3, nw = ((cx + 15 - xa) >> 2) + 1;          // <= 9 words
        const int lr = lane / kOdPitchW, lk = lane - lr * kOdPitchW;           // lanes 0..29: 3 rows x 10 words per step
        {
            const uint8_t *src = img + (long long)(cy - 15 + lr) * sp + xa + 4 * lk;
            const bool on = lane < 30 && lk < nw;
#pragma unroll
            for (int it = 0; it < 11; ++it, src += 3 * (long long)sp)
                if (on && 3 * it + lr < 31) patch[it * 30 + lane] = __ldg(reinterpret_cast<const uint32_t *>(src));
        }
        __syncwarp();
        const int off = cx - 15 - xa;
        const int u = lane - 15, au = abs(u);
        // rows of this column inside the circular patch: |v| <= umax[|u|] (the patch is symmetric, :453-469)
        const int vm = lane < 31 ? P->umax[au] : -1;
        const uint8_t *col = pb + 15 * (kOdPitchW * 4) + off + lane;
        int sum = 0, m01 = 0;
#pragma unroll
        for (int v = -15; v <= 15; ++v) {
            const int val = (v < 0 ? -v : v) <= vm ? (int)col[v * (kOdPitchW * 4)] : 0;
            sum += val;
            m01 += v * val;
        }
        int m10 = u * sum;
        m10 = warp_sum(m10);
        m01 = warp_sum(m01);
        __syncwarp();
        // ---- angle
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- steered rBRIEF: stage the 37x37 neighbourhood of the blurred level the same way
        const uint8_t *bl = P->blur + (long long)frame * P->pyr_frame_bytes + G.img_off;
        const int bp = G.pitch;
        const int xb = (cx - 18) & ~3, nwb = ((cx + 18 - xb) >> 2) + 1;         // <= 10 words
        {
            const uint8_t *bsrc = bl + (long long)(cy - 18 + lr) * bp + xb + 4 * lk;
            const bool on = lane < 30 && lk < nwb;
#pragma unroll
            for (int it = 0; it < 13; ++it, bsrc += 3 * (long long)bp)
                if (on && 3 * it + lr < 37) patch[it * 30 + lane] = __ldg(reinterpret_cast<const uint32_t *>(bsrc));
        }
        __syncwarp();
        constexpr float kFactorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        const float ang = __fmul_rn(angle, kFactorPI);
        const float a = (float)cos((double)ang), b = (float)sin((double)ang);
        const uint8_t *ctr = pb + 18 * (kOdPitchW * 4) + (cx - 18 - xb) + 18;
        const int8_t *pp = spat + lane;
        unsigned byte = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float x0 = (float)pp[(4 * k) * 32], y0 = (float)pp[(4 * k + 1) * 32], x1 = (float)pp[(4 * k + 2) * 32], y1 = (float)pp[(4 * k + 3) * 32];
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = ctr[r0 * (kOdPitchW * 4) + c0], t1 = ctr[r1 * (kOdPitchW * 4) + c1];
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
        if (P->out_ckps && lane < 3)                                   // orbx_keypoint_compact: x | y << 16, octave | response << 8, angle
            P->out_ckps[3 * row + lane] = lane == 0 ? (uint32_t)cx | (uint32_t)cy << 16 : (lane == 1 ? (uint32_t)level | (c >> 24) << 8 : __float_as_uint(angle));
    }
}

// ---- TMA variant: persistent warps, one keypoint at a time per warp.  The 31x31 source neighbourhood and the
// 37x37 blurred neighbourhood of a keypoint arrive as two cp.async.bulk.tensor boxes (16-byte aligned origin, so
// 48 and 64 bytes wide) and are double buffered: the boxes of the warp's next keypoint are in flight while the
// current one is processed.  Moments: lane = patch row, four pixels per IDP.4A against constant weights, the
// circular mask (umax, :453-469) as byte masks.  rBRIEF: the 256 sample pairs as a float4 table in shared memory.
constexpr int kOdIcW = 48, kOdIcH = 31, kOdBlW = 64, kOdBlH = 37;
constexpr int kOdIcBytes = 1536, kOdBlBytes = 2432;                  // each rounded up to a multiple of 128 (TMA destination alignment)
constexpr int kOdBufBytes = kOdIcBytes + kOdBlBytes;
constexpr int kOdSmemBytes = 4096 + 1024 + kOdWarps * 2 * kOdBufBytes;

__device__ __forceinline__ unsigned lds_u8(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <bool TMA>
__global__ void __launch_bounds__(kOdWarps * 32) k_orient_desc_tma(const DevParams *__restrict__ P, const __grid_constant__ OdMaps maps, Src0 s0)
{
    extern __shared__ __align__(128) uint8_t od_smem[];
    __shared__ int s_prefix[kMaxLevels + 1], s_kpoff[kMaxLevels];
    __shared__ __align__(8) uint64_t s_bar[kOdWarps][2];
    float4 *spatf = reinterpret_cast<float4 *>(od_smem);               // [k][lane]: (x0, y0, x1, y1) of bit k of byte `lane`
    uint32_t *smask = reinterpret_cast<uint32_t *>(od_smem + 4096);    // [word k][row]: bytes of the row inside the circular patch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame = blockIdx.y, L = P->nlevels;
    uint8_t *wbuf = od_smem + 4096 + 1024 + warp * 2 * kOdBufBytes;
    for (int i = threadIdx.x; i < 256; i += kOdWarps * 32) {
        const int l = i >> 3, k = i & 7;
        const int8_t *pt = P->pattern + 32 * l + 4 * k;
        spatf[k * 32 + l] = make_float4((float)pt[0], (float)pt[1], (float)pt[2], (float)pt[3]);
    }
    {
        const int k = threadIdx.x >> 5, r = threadIdx.x & 31;          // 8 warps x 32 lanes == the 8 x 32 mask table
        uint32_t m = 0;
        if (r < 31) {
            const int d = P->umax[abs(r - 15)];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) if (abs(4 * k + bb - 15) <= d) m |= 0xffu << (8 * bb);
        }
        smask[k * 32 + r] = m;
    }
    if (threadIdx.x < 32) {                       // keypoints per level of this frame and their running offsets
        int c = lane < L ? min((int)P->kp_count[frame * L + lane], P->lv[lane].kp_cap) : 0;
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < L) { s_prefix[lane] = x - c; s_kpoff[lane] = P->lv[lane].kp_off; }
        if (lane == L - 1) s_prefix[L] = x;
    }
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[warp][0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[warp][1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n = s_prefix[L];
    if (blockIdx.x == 0 && threadIdx.x == 0) P->out_n[frame] = n;
    uint32_t rmask[8];                                               // this lane's (= patch row's) circular mask, fixed for the kernel
#pragma unroll
    for (int k = 0; k < 8; ++k) rmask[k] = smask[k * 32 + lane];
    float4 rpat[8];                                                  // and its eight rBRIEF sample pairs (byte `lane` of the descriptor)
#pragma unroll
    for (int k = 0; k < 8; ++k) rpat[k] = spatf[k * 32 + lane];
    const int step = gridDim.x * kOdWarps;
    int j = blockIdx.x * kOdWarps + warp;
    if (j >= n) return;
    const int my_pre = lane < L ? s_prefix[lane + 1] : 0x7fffffff;     // first output index of level lane+1
    const uint32_t *stage = P->kp_stage + (long long)frame * P->kp_frame_cap;
    // level and packed candidate of output index jj
    auto fetch = [&](int jj, int *level) -> uint32_t {
        const int lv = __popc(__ballot_sync(0xffffffffu, my_pre <= jj));
        *level = lv;
        return __ldg(stage + s_kpoff[lv] + jj - s_prefix[lv]);
    };
    // chunk q of a patch = 16 bytes at (row q / chunks_per_row, column chunk q % chunks_per_row): lane-constant
    int ic_row[3], ic_col[3], bl_row[5], bl_col[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int q = lane + 32 * i; ic_row[i] = q / 3; ic_col[i] = (q - ic_row[i] * 3) * 16; }
#pragma unroll
    for (int i = 0; i < 5; ++i) { const int q = lane + 32 * i; bl_row[i] = q >> 2; bl_col[i] = (q & 3) * 16; }
    auto issue = [&](uint32_t c, int level, int b) {                   // both neighbourhoods of one keypoint into buffer b
        const int cx = (int)(c & 0xfff) + kMinBorder, cy = (int)((c >> 12) & 0xfff) + kMinBorder;
        if (TMA) {
            if (lane != 0) return;
            const unsigned bar = smem_u32(&s_bar[warp][b]), dst = smem_u32(wbuf + b * kOdBufBytes);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kOdIcW * kOdIcH + kOdBlW * kOdBlH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(&maps.img[level])), "r"(bar),
                            "r"((cx - 15) & ~15), "r"(cy - 15), "r"(frame) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(dst + kOdIcBytes), "l"(reinterpret_cast<unsigned long long>(&maps.blr[level])), "r"(bar),
                            "r"((cx - 18) & ~15), "r"(cy - 18), "r"(frame) : "memory");
        } else {
            // LDGSTS: 93 + 148 sixteen-byte chunks, eight cp.async per lane; rows stay inside the level (keypoints are
            // >= 19 pixels from every edge) and the pitch is a multiple of 16, so every chunk is in bounds and aligned
            int sp;
            const uint8_t *img = level_ptr(P, s0, frame, level, &sp);
            const int bp = P->lv[level].pitch;
            const uint8_t *ics = img + (long long)(cy - 15) * sp + ((cx - 15) & ~15);
            const uint8_t *bls = P->blur + (long long)frame * P->pyr_frame_bytes + P->lv[level].img_off + (long long)(cy - 18) * bp + ((cx - 18) & ~15);
            const unsigned dst = smem_u32(wbuf + b * kOdBufBytes) + lane * 16;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < 2 || lane < 93 - 64)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + 512 * i), "l"(ics + ic_row[i] * sp + ic_col[i]) : "memory");
#pragma unroll
            for (int i = 0; i < 5; ++i)
                if (i < 4 || lane < 148 - 128)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + kOdIcBytes + 512 * i), "l"(bls + bl_row[i] * bp + bl_col[i]) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    int lv0, lv1 = 0, lv2 = 0;
    uint32_t c0 = fetch(j, &lv0), c1 = 0, c2 = 0;
    issue(c0, lv0, 0);
    bool v1 = j + step < n, v2 = false;
    if (v1) c1 = fetch(j + step, &lv1);
    unsigned phase = 0;
    for (int it = 0;; ++it) {
        const int b = it & 1;
        __syncwarp();                                                  // every lane is done reading buffer b^1
        if (v1) issue(c1, lv1, b ^ 1);
        v2 = j + 2 * step < n;
        if (v2) c2 = fetch(j + 2 * step, &lv2);
        if (TMA) {
            mbar_wait_parity(&s_bar[warp][b], (phase >> b) & 1u);
            phase ^= 1u << b;
        } else {
            if (v1) asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the group just committed
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }

        const uint32_t c = c0;
        const int level = lv0;
        const LevelGeom &G = P->lv[level];
        const int cx = (int)(c & 0xfff) + kMinBorder, cy = (int)((c >> 12) & 0xfff) + kMinBorder;
        const uint8_t *icp = wbuf + b * kOdBufBytes, *blp = icp + kOdIcBytes;
        // ---- IC_Angle (:77-104): lane = patch row v = lane-15; the 32 pixels u = -15..16 of the row as 8 aligned words
        int m10, m01;
        {
            const int xoff = cx - 15 - ((cx - 15) & ~15);
            // the whole 48-byte box row as three 128-bit loads; the patch starts at word xoff / 4 of it (uniform over the warp)
            const uint4 *row = reinterpret_cast<const uint4 *>(icp + (lane < 31 ? lane : 0) * kOdIcW);
            const uint4 q0 = row[0], q1 = row[1], q2 = row[2];
            const uint32_t rw[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            const int sh = (xoff & 3) * 8;
            int rowsum = 0, mu = 0;
#define OD_MOMENTS(WO)                                                                                       \
            _Pragma("unroll")                                                                                \
            for (int k = 0; k < 8; ++k) {                                                                    \
                const uint32_t v = __funnelshift_r(rw[WO + k], rw[WO + k + 1], sh) & rmask[k];               \
                constexpr int kOnes = 0x01010101;                                                            \
                const int u0 = 4 * k - 15;                                                                   \
                const int wu = (u0 & 0xff) | ((u0 + 1) & 0xff) << 8 | ((u0 + 2) & 0xff) << 16 | ((u0 + 3) & 0xff) << 24; \
                rowsum = dp4a_us(v, kOnes, rowsum);                                                          \
                mu = dp4a_us(v, wu, mu);                                                                     \
            }
            switch (xoff >> 2) {
            case 0: OD_MOMENTS(0) break;
            case 1: OD_MOMENTS(1) break;
            case 2: OD_MOMENTS(2) break;
            default: OD_MOMENTS(3) break;
            }
#undef OD_MOMENTS
            m10 = __reduce_add_sync(0xffffffffu, mu);                 // REDUX.SUM: one instruction per moment
            m01 = __reduce_add_sync(0xffffffffu, (lane - 15) * rowsum);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        // ---- steered rBRIEF (:108-147)
        constexpr float kFactorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        const float ang = __fmul_rn(angle, kFactorPI);
        float a, bs;
        sincos_f32_via_f64(ang, &bs, &a);
        // cvRound(s) for |s| < 2^22 is the low mantissa of fl(s + 1.5*2^23) (round-half-even, like cvRound); the bias
        // of the two integers is folded into the patch pointer, so no F2I is issued
        constexpr float kMagic = 12582912.f;
        constexpr int kMagicBits = 0x4B400000;
        const unsigned ctr = smem_u32(blp) + 18 * kOdBlW + (cx - 18 - ((cx - 18) & ~15)) + 18 - (unsigned)kMagicBits * (kOdBlW + 1);
        unsigned byte = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 pt = rpat[k];
            const int r0 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, bs), __fmul_rn(pt.y, a)), kMagic));
            const int q0 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, bs)), kMagic));
            const int r1 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.z, bs), __fmul_rn(pt.w, a)), kMagic));
            const int q1 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, bs)), kMagic));
            const unsigned t0 = lds_u8(ctr + (unsigned)r0 * kOdBlW + (unsigned)q0), t1 = lds_u8(ctr + (unsigned)r1 * kOdBlW + (unsigned)q1);
            byte |= (unsigned)(t0 < t1) << k;
        }
        // gather 32 bytes -> 8 words (lanes 0, 4, .., 28) -> two uint4 stores (lanes 0 and 16)
        unsigned wd = byte | __shfl_down_sync(0xffffffffu, byte, 1) << 8;
        wd = (wd & 0xffffu) | __shfl_down_sync(0xffffffffu, wd, 2) << 16;
        uint4 v;
        v.x = wd; v.y = __shfl_down_sync(0xffffffffu, wd, 4);
        v.z = __shfl_down_sync(0xffffffffu, wd, 8); v.w = __shfl_down_sync(0xffffffffu, wd, 12);
        const long long orow = (long long)frame * P->kp_frame_cap + j;
        if ((lane & 15) == 0) reinterpret_cast<uint4 *>(P->out_desc + orow * 32)[lane >> 4] = v;
        // ---- cv::KeyPoint record (:837-847, :1098-1104): lanes 0..6 = x, y, size, angle, response, octave, class_id
        {
            const float fx = __fmul_rn((float)cx, G.scale), fy = __fmul_rn((float)cy, G.scale);
            float f = lane == 0 ? fx : fy;
            f = lane == 2 ? G.kp_size : f;
            f = lane == 3 ? angle : f;
            f = lane == 4 ? (float)(c >> 24) : f;
            f = lane == 5 ? __int_as_float(level) : f;
            f = lane == 6 ? __int_as_float(-1) : f;
            if (lane < 7) reinterpret_cast<float *>(P->out_kps + orow)[lane] = f;
            if (P->out_ckps && lane < 3)                               // orbx_keypoint_compact: x | y << 16, octave | response << 8, angle
                P->out_ckps[3 * orow + lane] = lane == 0 ? (uint32_t)cx | (uint32_t)cy << 16 : (lane == 1 ? (uint32_t)level | (c >> 24) << 8 : __float_as_uint(angle));
        }
        if (!v1) break;
        c0 = c1; lv0 = lv1; c1 = c2; lv1 = lv2; v1 = v2; j += step;
    }
}

cudaError_t launch_orient_desc(const DevParams *dP, const DevParams &hP, Src0 s0, int nframes, cudaStream_t st, LaunchStats *ls, const OdMaps *maps)
{
    ls->launches++;
    if (maps) {
        {
            cudaError_t e = cudaFuncSetAttribute(k_orient_desc_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOdSmemBytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_orient_desc_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOdSmemBytes);
            if (e != cudaSuccess) return e;
        }
        // 3 CTAs of 8 warps per SM resident; every warp walks several keypoints of its frame
        int per_frame = 148 * 3 / nframes;                             // one wave: never more CTAs than resident slots
        const int max_useful = (hP.kp_frame_cap + kOdWarps - 1) / kOdWarps;
        if (per_frame > max_useful) per_frame = max_useful;
        if (per_frame < 1) per_frame = 1;
        static const int mode = debug_knob("ORBX_OD_MODE", 1);   // 0 = LDGSTS chunks (measured slower), 1 = TMA boxes
        if (mode == 1) k_orient_desc_tma<true><<<dim3(per_frame, nframes), kOdWarps * 32, kOdSmemBytes, st>>>(dP, *maps, s0);
        else k_orient_desc_tma<false><<<dim3(per_frame, nframes), kOdWarps * 32, kOdSmemBytes, st>>>(dP, *maps, s0);
        return cudaGetLastError();
    }
    dim3 grid((hP.kp_frame_cap + kOdWarps - 1) / kOdWarps, nframes);
    k_orient_desc<<<grid, kOdWarps * 32, 0, st>>>(dP, s0);
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

// One thread = one query row; the chunk's B rows pass through shared memory.  Per pair: 8 XOR, three carry-save adders (two LOP3
// each) that fold the eight XOR words into two "ones" and three "twos" words, five POPC (the slow instruction: 16 lanes / clk / SM)
// instead of eight, and the update rule of :589-598 on packed keys  d << 22 | j  (j = index inside the chunk): the smallest key is
// the least distance at its lowest index, the second smallest key carries the second smallest distance of the multiset, so the rule
// is  b2 = min(b2, max(b1, key)); b1 = min(b1, key)  -- three min / max, no compare / select chain.  The weighted sum of the five
// counts and the key are built with IMADs (FMA pipe), leaving the integer ALU pipe to the LOP3s.
constexpr int kMatchKeyShift = 22;                                   // d <= 256 -> keys < 2^31; chunks hold fewer than 2^22 rows
__device__ __forceinline__ unsigned imad_u(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(kMatchThreads)
k_match_partial(const uint32_t *__restrict__ A, int nA, const uint32_t *__restrict__ B, int nB, int chunk, int4 *partial,
                unsigned *__restrict__ arrive, int th, float ratio, int32_t *__restrict__ idx, int32_t *__restrict__ d1o, int32_t *__restrict__ d2o,
                uint8_t *__restrict__ accept, int *__restrict__ naccept)
{
    __shared__ uint4 sB[kMatchTile * 2];
    __shared__ unsigned s_ticket;
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
    constexpr unsigned kNone = 256u << kMatchKeyShift | ((1u << kMatchKeyShift) - 1u);      // above every real key
    unsigned b1 = kNone, b2 = kNone;
    for (int t0 = j0; t0 < j1; t0 += kMatchTile) {
        const int cnt = min(kMatchTile, j1 - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 2; k += kMatchThreads) sB[k] = reinterpret_cast<const uint4 *>(B)[2 * (long long)t0 + k];
        __syncthreads();
        unsigned jrel = (unsigned)(t0 - j0);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const uint4 lo = sB[2 * j], hi = sB[2 * j + 1];
            const uint32_t x0 = a[0] ^ lo.x, x1 = a[1] ^ lo.y, x2 = a[2] ^ lo.z, x3 = a[3] ^ lo.w;
            const uint32_t x4 = a[4] ^ hi.x, x5 = a[5] ^ hi.y, x6 = a[6] ^ hi.z, x7 = a[7] ^ hi.w;
            const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
            const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
            const uint32_t s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
            const unsigned twos = __popc(c0) + __popc(c1) + __popc(c2);                     // one IADD3
            const unsigned d = imad_u(twos, 2u, __popc(s2)) + __popc(x7);
            const unsigned key = imad_u(d, 1u << kMatchKeyShift, jrel);
            jrel = imad_u(jrel, 1u, 1u);
            b2 = min(b2, max(b1, key));
            b1 = min(b1, key);
        }
    }
    if (i < nA) {
        const int d1 = (int)(b1 >> kMatchKeyShift), d2 = (int)(b2 >> kMatchKeyShift);
        const bool any = b1 != kNone;
        partial[(long long)blockIdx.y * nA + i] = make_int4(any ? d1 : 256, any ? j0 + (int)(b1 & ((1u << kMatchKeyShift) - 1u)) : -1, b2 != kNone ? d2 : 256, 0);
    }
    // ---- the last chunk's CTA of this row block to arrive merges the row block's partial results (chunks in ascending index
    //      order, so "lowest index attaining the minimum" survives) and applies the threshold and the ratio test (:601-603)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(arrive + blockIdx.x, 1u);
    __syncthreads();
    if (s_ticket != gridDim.y - 1) return;
    if (threadIdx.x == 0) arrive[blockIdx.x] = 0;                    // ready for the next call on this stream
    __threadfence();
    bool ok = false;
    if (i < nA) {
        int m1 = 256, m2 = 256, mi = -1;
        for (int c = 0; c < (int)gridDim.y; ++c) {
            const int4 p = __ldcg(partial + (long long)c * nA + i);   // written by other CTAs: read through L2
            if (p.x < m1) { m2 = min(m1, p.z); m1 = p.x; mi = p.y; }
            else m2 = min(m2, p.x);
        }
        idx[i] = mi; d1o[i] = m1; d2o[i] = m2;
        ok = mi >= 0 && m1 <= th && (float)m1 < __fmul_rn(ratio, (float)m2);
        if (accept) accept[i] = ok;
    }
    if (naccept) {
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(naccept, __popc(m));
    }
}

// mbCheckOrientation (src/ORBmatcher.cc:545, 610-620, 641-660; ComputeThreeMaxima :2233-2274): 30-bin histogram of
// the angle differences of the accepted matches, bin = round(rot * (1.0f / HISTO_LENGTH)) exactly as the reference
// writes it (only bins 0..12 can fill; kept for parity), the three most populated bins survive (second / third only
// when >= 0.1 * first, ties to the lower bin as the reference's strict > scan does).  One CTA; counts are
// order-independent, so shared-memory atomics are exact.
__device__ __forceinline__ int rotation_bin(float a, float b)
{
    float rot = __fsub_rn(a, b);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    // the reference asserts bin in [0, HISTO_LENGTH) (:618); angles outside [0, 360) or NaN (uninitialised keypoints) land in the
    // overflow bin HISTO_LENGTH, which ComputeThreeMaxima never sees, so such a match is pruned instead of indexing out of bounds
    if (!(rot >= 0.0f && rot < 915.0f)) return ORBX_HISTO_LENGTH;
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / ORBX_HISTO_LENGTH));
    return bin == ORBX_HISTO_LENGTH ? 0 : bin;
}

__global__ void __launch_bounds__(1024)
k_rotation_filter(int nA, const int32_t *__restrict__ idx, uint8_t *__restrict__ accept, const float *__restrict__ angleA,
                  const float *__restrict__ angleB, int32_t *__restrict__ hist_out, int32_t *__restrict__ top3_out, int *__restrict__ kept_out)
{
    __shared__ int cnt[32], sel[3], kept;
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nA; i += blockDim.x)
        if (accept[i]) atomicAdd(&cnt[rotation_bin(angleA[i], angleB[idx[i]])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < ORBX_HISTO_LENGTH; ++i) {
            const int s = cnt[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        const float lim = __fmul_rn(0.1f, (float)max1);
        if ((float)max2 < lim) { i2 = -1; i3 = -1; }
        else if ((float)max3 < lim) i3 = -1;
        sel[0] = i1; sel[1] = i2; sel[2] = i3;
    }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < nA; i += blockDim.x)
        if (accept[i]) {
            const int b = rotation_bin(angleA[i], angleB[idx[i]]);
            if (b == sel[0] || b == sel[1] || b == sel[2]) ++mine; else accept[i] = 0;
        }
    mine = warp_sum(mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&kept, mine);
    __syncthreads();
    if (threadIdx.x < ORBX_HISTO_LENGTH && hist_out) hist_out[threadIdx.x] = cnt[threadIdx.x];
    if (threadIdx.x < 3 && top3_out) top3_out[threadIdx.x] = sel[threadIdx.x];
    if (threadIdx.x == 0 && kept_out) *kept_out = kept;
}

cudaError_t launch_rotation_filter(int nA, const int32_t *d_idx, uint8_t *d_accept, const float *d_angleA, const float *d_angleB,
                                   int32_t *d_hist, int32_t *d_top3, int *d_kept, cudaStream_t st, LaunchStats *ls)
{
    k_rotation_filter<<<1, 1024, 0, st>>>(nA, d_idx, d_accept, d_angleA, d_angleB, d_hist, d_top3, d_kept);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------ projection matcher
// The parallel part of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1958-2102): one warp
// per last-frame map point projects it into the current frame (OpenCV's CV_32F gemm semantics: float accumulation, k ascending,
// then + t; 1.0 / z in double), derives the search window and the cell range of Frame::GetFeaturesInArea (src/Frame.cc:710-763),
// and tests EVERY current feature against the cell range (PosInGrid rounding, :765-776), the level range, the |dx|, |dy| < r
// window and the stereo gate (:2032-2038); survivors leave as (cell << 32 | index << 16 | Hamming distance), the order the
// reference would visit them in.  The claim bookkeeping, which is sequential in the reference, follows on the host.
// One warp: every feature of the searched frame against one query window (cell range of GetFeaturesInArea, level range,
// |dx|, |dy| < r, optional stereo gate), Hamming distance of the survivors, compact write-out in arbitrary order (the packed
// key carries the reference's visiting order).  count[i] may exceed cap; the caller reports that.
__device__ __forceinline__ void window_scan(const ProjSetup &S, bool live, float u, float v, float radius, int min_cx, int max_cx, int min_cy, int max_cy,
                                            int min_level, int max_level, bool stereo_gate, float ur, const uint8_t *qdesc, int n_cur,
                                            const float *__restrict__ cur_xy, const int32_t *__restrict__ cur_octave, const float *__restrict__ cur_uright,
                                            const uint8_t *__restrict__ cur_desc, int cap, unsigned long long *list, int lane, int i,
                                            unsigned long long *__restrict__ cand, int *__restrict__ count, int *__restrict__ offset, int *__restrict__ total)
{
    int n_out = 0;
    if (live) {
        const bool check_levels = min_level > 0 || max_level >= 0;
        const uint4 a0 = reinterpret_cast<const uint4 *>(qdesc)[0], a1 = reinterpret_cast<const uint4 *>(qdesc)[1];
        for (int base = 0; base < n_cur; base += 32) {
            const int i2 = base + lane;
            bool ok = i2 < n_cur;
            unsigned long long key = 0;
            if (ok) {
                const float x = cur_xy[2 * i2], y = cur_xy[2 * i2 + 1];
                const int px = (int)roundf(__fmul_rn(__fsub_rn(x, S.min_x), S.w_inv)), py = (int)roundf(__fmul_rn(__fsub_rn(y, S.min_y), S.h_inv));
                ok = px >= min_cx && px <= max_cx && py >= min_cy && py <= max_cy && px < 64 && py < 48;     // in a visited grid cell (PosInGrid)
                if (ok && check_levels) {
                    const int lv = cur_octave[i2];
                    ok = !(lv < min_level) && !(max_level >= 0 && lv > max_level);
                }
                ok = ok && fabsf(__fsub_rn(x, u)) < radius && fabsf(__fsub_rn(y, v)) < radius;
                if (ok && stereo_gate) {
                    const float urt = cur_uright[i2];
                    if (urt > 0 && fabsf(__fsub_rn(ur, urt)) > radius) ok = false;
                }
                if (ok) {
                    const uint4 b0 = reinterpret_cast<const uint4 *>(cur_desc)[2 * (long long)i2], b1 = reinterpret_cast<const uint4 *>(cur_desc)[2 * (long long)i2 + 1];
                    const unsigned d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                       __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                    key = (unsigned long long)(px * 48 + py) << 32 | (unsigned long long)i2 << 16 | d;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int at = n_out + __popc(m & lanemask_lt());
                if (at < cap) list[at] = key;
            }
            n_out += __popc(m);
        }
    }
    __syncwarp();
    int base = 0;
    const int keep = min(n_out, cap);
    if (lane == 0) { base = keep ? atomicAdd(total, keep) : 0; count[i] = n_out; offset[i] = base; }
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < keep; k += 32) cand[base + k] = list[k];
}

__global__ void __launch_bounds__(256)
k_project_candidates(ProjSetup S, int n_last, const float *__restrict__ world_pos, const uint8_t *__restrict__ mp_desc,
                     const uint8_t *__restrict__ valid, const int32_t *__restrict__ last_octave, int n_cur,
                     const float *__restrict__ cur_xy, const int32_t *__restrict__ cur_octave, const float *__restrict__ cur_uright,
                     const uint8_t *__restrict__ cur_desc, int cap, unsigned long long *__restrict__ gstage, unsigned long long *__restrict__ cand, int *__restrict__ count,
                     int *__restrict__ offset, int *__restrict__ total)
{
    __shared__ unsigned long long s_list[8][kProjCap];               // per-warp staging; the lists leave compacted (one atomicAdd per point)
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i >= n_last) return;
    bool live = valid[i] != 0;
    float u = 0, v = 0, invzc = 0, radius = 0;
    int min_cx = 0, max_cx = -1, min_cy = 0, max_cy = -1, min_level = 0, max_level = -1;
    if (live) {
        const float x0 = world_pos[3 * i], x1 = world_pos[3 * i + 1], x2 = world_pos[3 * i + 2];
        float p[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float s = __fadd_rn(0.f, __fmul_rn(S.Tc[4 * r], x0));
            s = __fadd_rn(s, __fmul_rn(S.Tc[4 * r + 1], x1));
            s = __fadd_rn(s, __fmul_rn(S.Tc[4 * r + 2], x2));
            p[r] = __fadd_rn(s, S.Tc[4 * r + 3]);
        }
        invzc = __double2float_rn(__ddiv_rn(1.0, (double)p[2]));
        u = __fadd_rn(__fmul_rn(__fmul_rn(S.fx, p[0]), invzc), S.cx);
        v = __fadd_rn(__fmul_rn(__fmul_rn(S.fy, p[1]), invzc), S.cy);
        live = !(invzc < 0) && !(u < S.min_x || u > S.max_x) && !(v < S.min_y || v > S.max_y);
    }
    if (live) {
        const int oct = last_octave[i];
        radius = __fmul_rn(S.th, S.scale[oct]);
        if (S.forward) { min_level = oct; max_level = -1; }
        else if (S.backward) { min_level = 0; max_level = oct; }
        else { min_level = oct - 1; max_level = oct + 1; }
        min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        max_cx = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        max_cy = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        live = min_cx < 64 && max_cx >= 0 && min_cy < 48 && max_cy >= 0;
    }
    window_scan(S, live, u, v, radius, min_cx, max_cx, min_cy, max_cy, min_level, max_level, true, __fsub_rn(u, __fmul_rn(S.bf, invzc)),
                mp_desc + 32 * (long long)i, n_cur, cur_xy, cur_octave, cur_uright, cur_desc, cap, gstage ? gstage + (size_t)i * cap : s_list[warp], lane, i, cand, count, offset, total);
}

// The window search of ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:797): octave-0 keypoints of F1 look around
// vbPrevMatched[i1] in F2's grid, levels (level1, level1), radius = windowSize, no stereo gate.
__global__ void __launch_bounds__(256)
k_window_candidates(ProjSetup S, int n1, const float *__restrict__ prev_xy, const int32_t *__restrict__ oct1, const uint8_t *__restrict__ desc1,
                    int n2, const float *__restrict__ xy2, const int32_t *__restrict__ oct2, const uint8_t *__restrict__ desc2, int cap, unsigned long long *__restrict__ gstage,
                    unsigned long long *__restrict__ cand, int *__restrict__ count, int *__restrict__ offset, int *__restrict__ total)
{
    __shared__ unsigned long long s_list[8][kProjCap];
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i >= n1) return;
    const int level1 = oct1[i];
    bool live = !(level1 > 0);
    const float u = prev_xy[2 * i], v = prev_xy[2 * i + 1], radius = S.th;
    int min_cx = 0, max_cx = -1, min_cy = 0, max_cy = -1;
    if (live) {
        min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        max_cx = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        max_cy = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        live = min_cx < 64 && max_cx >= 0 && min_cy < 48 && max_cy >= 0;
    }
    window_scan(S, live, u, v, radius, min_cx, max_cx, min_cy, max_cy, level1, level1, false, 0.f, desc1 + 32 * (long long)i, n2, xy2, oct2,
                nullptr, desc2, cap, gstage ? gstage + (size_t)i * cap : s_list[warp], lane, i, cand, count, offset, total);
}

// The window search of ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (src/ORBmatcher.cc:418-502):
// a local map point looks around (mTrackProjX, mTrackProjY) with radius RadiusByViewingCos(mTrackViewCos) [x th] x scale[predicted level]
// on levels (predicted - 1, predicted); the right-coordinate gate uses the same radius (:463-468).  S.th carries th, S.forward bFactor.
__global__ void __launch_bounds__(256)
k_local_candidates(ProjSetup S, int n_mp, const float *__restrict__ proj, const float *__restrict__ view_cos, const int32_t *__restrict__ level,
                   const uint8_t *__restrict__ mp_desc, const uint8_t *__restrict__ valid, int n_feat, const float *__restrict__ xy,
                   const int32_t *__restrict__ octave, const float *__restrict__ uright, const uint8_t *__restrict__ desc, int cap, unsigned long long *__restrict__ gstage,
                   unsigned long long *__restrict__ cand, int *__restrict__ count, int *__restrict__ offset, int *__restrict__ total)
{
    __shared__ unsigned long long s_list[8][kProjCap];
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i >= n_mp) return;
    bool live = valid[i] != 0;
    // mnTrackScaleLevel is uninitialised for points outside the frustum (valid == 0): only a live point's level indexes the scale table
    const int lvl = live ? min(max(level[i], 0), kMaxLevels - 1) : 0;
    float r = (double)view_cos[i] > 0.998 ? 2.5f : 4.0f;                       // RadiusByViewingCos, :504-510
    if (S.forward) r = __fmul_rn(r, S.th);
    const float radius = __fmul_rn(r, S.scale[lvl]);
    const float u = proj[3 * i], v = proj[3 * i + 1];
    int min_cx = 0, max_cx = -1, min_cy = 0, max_cy = -1;
    if (live) {
        min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        max_cx = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, S.min_x), radius), S.w_inv)));
        min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        max_cy = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, S.min_y), radius), S.h_inv)));
        live = min_cx < 64 && max_cx >= 0 && min_cy < 48 && max_cy >= 0;
    }
    window_scan(S, live, u, v, radius, min_cx, max_cx, min_cy, max_cy, lvl - 1, lvl, true, proj[3 * i + 2], mp_desc + 32 * (long long)i, n_feat, xy,
                octave, uright, desc, cap, gstage ? gstage + (size_t)i * cap : s_list[warp], lane, i, cand, count, offset, total);
}

cudaError_t launch_local_candidates(const ProjSetup &S, int n_mp, const float *d_proj, const float *d_view_cos, const int32_t *d_level,
                                    const uint8_t *d_mp_desc, const uint8_t *d_valid, int n_feat, const float *d_xy, const int32_t *d_octave,
                                    const float *d_uright, const uint8_t *d_desc, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count, int *d_offset,
                                    int *d_total, cudaStream_t st, LaunchStats *ls)
{
    if (n_mp <= 0) return cudaSuccess;
    k_local_candidates<<<(n_mp + 7) / 8, 256, 0, st>>>(S, n_mp, d_proj, d_view_cos, d_level, d_mp_desc, d_valid, n_feat, d_xy, d_octave, d_uright, d_desc,
                                                       cap, d_stage, d_cand, d_count, d_offset, d_total);
    ls->launches++;
    return cudaGetLastError();
}

// The distances of ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:555-598): entry e = one key-frame feature with a
// good map point in a vocabulary node both feature vectors share, (x: key-frame feature, y: first slot of the node in the frame's
// flattened feature vector, z: number of frame features there, w: first output slot); one warp per entry, lane = frame feature.
__global__ void __launch_bounds__(256)
k_bow_pair_distances(int n_entries, const int4 *__restrict__ entries, const uint8_t *__restrict__ kf_desc, const uint8_t *__restrict__ f_desc,
                     const int32_t *__restrict__ f_feats, uint16_t *__restrict__ out)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n_entries) return;
    const int4 E = entries[e];
    const uint4 a0 = reinterpret_cast<const uint4 *>(kf_desc)[2 * (long long)E.x], a1 = reinterpret_cast<const uint4 *>(kf_desc)[2 * (long long)E.x + 1];
    for (int t = lane; t < E.z; t += 32) {
        const long long j = f_feats[E.y + t];
        const uint4 b0 = reinterpret_cast<const uint4 *>(f_desc)[2 * j], b1 = reinterpret_cast<const uint4 *>(f_desc)[2 * j + 1];
        out[E.w + t] = (uint16_t)(__popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w));
    }
}

cudaError_t launch_bow_pair_distances(int n_entries, const int4 *d_entries, const uint8_t *d_kf_desc, const uint8_t *d_f_desc, const int32_t *d_f_feats,
                                      uint16_t *d_out, cudaStream_t st, LaunchStats *ls)
{
    if (n_entries <= 0) return cudaSuccess;
    k_bow_pair_distances<<<(n_entries + 7) / 8, 256, 0, st>>>(n_entries, d_entries, d_kf_desc, d_f_desc, d_f_feats, d_out);
    ls->launches++;
    return cudaGetLastError();
}

cudaError_t launch_window_candidates(const ProjSetup &S, int n1, const float *d_prev_xy, const int32_t *d_oct1, const uint8_t *d_desc1, int n2,
                                     const float *d_xy2, const int32_t *d_oct2, const uint8_t *d_desc2, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count,
                                     int *d_offset, int *d_total, cudaStream_t st, LaunchStats *ls)
{
    if (n1 <= 0) return cudaSuccess;
    k_window_candidates<<<(n1 + 7) / 8, 256, 0, st>>>(S, n1, d_prev_xy, d_oct1, d_desc1, n2, d_xy2, d_oct2, d_desc2, cap, d_stage, d_cand, d_count, d_offset, d_total);
    ls->launches++;
    return cudaGetLastError();
}

cudaError_t launch_project_candidates(const ProjSetup &S, int n_last, const float *d_world, const uint8_t *d_mp_desc, const uint8_t *d_valid,
                                      const int32_t *d_last_octave, int n_cur, const float *d_cur_xy, const int32_t *d_cur_octave,
                                      const float *d_cur_uright, const uint8_t *d_cur_desc, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_count, int *d_offset,
                                      int *d_total, cudaStream_t st, LaunchStats *ls)
{
    if (n_last <= 0) return cudaSuccess;
    k_project_candidates<<<(n_last + 7) / 8, 256, 0, st>>>(S, n_last, d_world, d_mp_desc, d_valid, d_last_octave, n_cur, d_cur_xy, d_cur_octave,
                                                          d_cur_uright, d_cur_desc, cap, d_stage, d_cand, d_count, d_offset, d_total);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------------- BoW descent
// TemplatedVocabulary::transform(feature, word, weight, nid, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1205-1250),
// one warp per descriptor: lane c takes child c of the current node (k <= 32), FORB::distance by __popc (FORB.cpp:81-101),
// the FIRST child with the least distance wins (redux.min on distance << 8 | child position), down to a node without
// children; the node id of level L - levelsup is recorded on the way.
__global__ void __launch_bounds__(256)
k_bow_descent(const uint8_t *__restrict__ feat, int n, const int32_t *__restrict__ child_off, const int32_t *__restrict__ child_ids,
              const uint4 *__restrict__ node_desc, const int32_t *__restrict__ word_id, int nid_level,
              int32_t *__restrict__ word_out, int32_t *__restrict__ node_out, int32_t *__restrict__ final_out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const uint4 a0 = reinterpret_cast<const uint4 *>(feat)[2 * (long long)warp], a1 = reinterpret_cast<const uint4 *>(feat)[2 * (long long)warp + 1];
    int node = 0, level = 0, nid = 0;
    int off = __ldg(child_off), cnt = __ldg(child_off + 1) - off;
    do {
        ++level;
        unsigned key = 0xffffffffu;
        int mine = 0;
        if (lane < cnt) {
            mine = __ldg(child_ids + off + lane);
            const uint4 b0 = __ldg(node_desc + 2 * (long long)mine), b1 = __ldg(node_desc + 2 * (long long)mine + 1);
            const unsigned d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                               __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            key = d << 8 | (unsigned)lane;
        }
        key = __reduce_min_sync(0xffffffffu, key);
        node = __shfl_sync(0xffffffffu, mine, (int)(key & 0xffu));
        if (level == nid_level) nid = node;
        off = __ldg(child_off + node); cnt = __ldg(child_off + node + 1) - off;
    } while (cnt > 0);
    if (lane == 0) { word_out[warp] = __ldg(word_id + node); node_out[warp] = nid; final_out[warp] = node; }
}

cudaError_t launch_bow_descent(const uint8_t *d_feat, int n, const int32_t *d_child_off, const int32_t *d_child_ids, const uint8_t *d_node_desc,
                               const int32_t *d_word_id, int nid_level, int32_t *d_word, int32_t *d_node, int32_t *d_final, cudaStream_t st, LaunchStats *ls)
{
    if (n <= 0) return cudaSuccess;
    k_bow_descent<<<(n + 7) / 8, 256, 0, st>>>(d_feat, n, d_child_off, d_child_ids, reinterpret_cast<const uint4 *>(d_node_desc), d_word_id, nid_level,
                                                 d_word, d_node, d_final);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------- distinctive descriptors
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301), batched over map points: one CTA per map point,
// its n observed descriptors staged in shared memory; a warp owns a row i, builds the 257-bin histogram of the
// distances d(i, .) (d(i,i) = 0) with shared-memory atomics and reads the median off the running count
// (sorted[(int)(0.5*(n-1))]); the point keeps the FIRST row with the least median (atomicMin on median << 16 | i).
constexpr int kDdWarps = 4, kDdMaxObs = 1024;

__global__ void __launch_bounds__(kDdWarps * 32)
k_distinctive(const uint8_t *__restrict__ desc, const int32_t *__restrict__ offsets, int npoints, int32_t *__restrict__ best_idx,
              int32_t *__restrict__ best_median)
{
    extern __shared__ __align__(16) uint8_t dd_smem[];               // [n][32] descriptors, then kDdWarps x 264 histogram words
    __shared__ unsigned s_best;
    const int p = blockIdx.x;
    if (p >= npoints) return;
    const int o0 = offsets[p], n = offsets[p + 1] - o0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (n <= 0) { if (threadIdx.x == 0) { best_idx[p] = -1; if (best_median) best_median[p] = -1; } return; }
    uint4 *sd = reinterpret_cast<uint4 *>(dd_smem);
    int *hist = reinterpret_cast<int *>(dd_smem + (size_t)n * 32) + warp * 264;
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) sd[i] = reinterpret_cast<const uint4 *>(desc + (long long)o0 * 32)[i];
    if (threadIdx.x == 0) s_best = 0xffffffffu;
    __syncthreads();
    const int mpos = (int)(0.5 * (n - 1));                           // index of the median in the sorted row
    for (int i = warp; i < n; i += kDdWarps) {
        for (int k = lane; k < 264; k += 32) hist[k] = 0;
        __syncwarp();
        const uint4 a0 = sd[2 * i], a1 = sd[2 * i + 1];
        for (int j = lane; j < n; j += 32) {
            const uint4 b0 = sd[2 * j], b1 = sd[2 * j + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            atomicAdd(&hist[d], 1);                                   // d(i,i) = 0 falls out of the XOR
        }
        __syncwarp();
        // median = smallest v with count(d <= v) > mpos: lane l owns bins 9l .. 9l+8 (288 >= 257)
        int c[9], sum = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const int b = 9 * lane + k; c[k] = b < 257 ? hist[b] : 0; sum += c[k]; }
        int incl = sum;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, ofs); if (lane >= ofs) incl += t; }
        int run = incl - sum, med = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 9; ++k) { run += c[k]; if (run > mpos && med == 0x7fffffff) med = 9 * lane + k; }
        med = __reduce_min_sync(0xffffffffu, med);
        if (lane == 0) atomicMin(&s_best, (unsigned)med << 16 | (unsigned)i);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) { best_idx[p] = (int)(s_best & 0xffffu); if (best_median) best_median[p] = (int)(s_best >> 16); }
}

cudaError_t launch_distinctive(const uint8_t *d_desc, const int32_t *d_offsets, int npoints, int max_obs, int32_t *d_best_idx,
                               int32_t *d_best_median, cudaStream_t st, LaunchStats *ls)
{
    if (npoints <= 0) return cudaSuccess;
    const size_t smem = (size_t)max_obs * 32 + kDdWarps * 264 * sizeof(int);
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptInSmem);
        if (e != cudaSuccess) return e;
    }
    k_distinctive<<<npoints, kDdWarps * 32, smem, st>>>(d_desc, d_offsets, npoints, d_best_idx, d_best_median);
    ls->launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------- stereo
// Frame::ComputeStereoMatches (src/Frame.cc:849-1038) on the device-resident results of a left and a right extractor.
// One warp per left keypoint: (1) scan of ALL right keypoints with the reference's row-band / octave / disparity gates
// (the row table vRowIndices of :861-881 is a gate "row(int)vL in [floor(y-r), ceil(y+r)]", candidates in ascending iR,
// strict < keeps the first minimum, :914-941), Hamming by __popc; (2) the 11-position SAD search on the 11x11 patches of
// the two pyramids, patches normalised by their centre pixel (:949-985; integer-valued floats, so integer arithmetic is
// exact), parabola fit and disparity in the reference's float operation order (:990-1019).  A second, single-CTA kernel
// applies the median cut of :1022-1037 by rank selection (no sort needed: only the median VALUE enters thDist).
struct StereoSide {
    const DevParams *P; Src0 s0; int frame;
};

__global__ void __launch_bounds__(256)
k_stereo_match(StereoSide L, StereoSide R, StereoScales sc, float bf, float *__restrict__ u_right, float *__restrict__ depth,
               int32_t *__restrict__ desc_index, int32_t *__restrict__ sad_out)
{
    __shared__ int sIL[8][121 + 231];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + warp;
    const DevParams *PL = L.P, *PR = R.P;
    const int nL = PL->out_n[L.frame], nR = PR->out_n[R.frame];
    if (iL >= nL) return;
    const orbx_keypoint *kL = PL->out_kps + (long long)L.frame * PL->kp_frame_cap;
    const orbx_keypoint *kR = PR->out_kps + (long long)R.frame * PR->kp_frame_cap;
    const uint4 *dLp = reinterpret_cast<const uint4 *>(PL->out_desc + ((long long)L.frame * PL->kp_frame_cap + iL) * 32);
    const uint8_t *dRb = PR->out_desc + (long long)R.frame * PR->kp_frame_cap * 32;
    float ur_o = -1.0f, dp_o = -1.0f; int di_o = -1, sad_o = -1;
    const float uL = kL[iL].x, vL = kL[iL].y;
    const int levelL = kL[iL].octave;
    const int row = (int)vL;
    const float minU = __fsub_rn(uL, 200.0f), maxU = uL;
    bool live = !(maxU < 0);
    unsigned key = 100u << 16;                                        // bestDist = TH_HIGH, bestIdxR = 0
    bool any = false;
    if (live) {
        const uint4 a0 = dLp[0], a1 = dLp[1];
        for (int iR = lane; iR < nR; iR += 32) {
            const float y = kR[iR].y;
            const int oc = kR[iR].octave;
            const float r = __fmul_rn(1.2f, sc.scale[oc]);
            const int maxr = (int)ceilf(__fadd_rn(y, r)), minr = (int)floorf(__fsub_rn(y, r));
            if (row < minr || row > maxr) continue;
            any = true;
            if (oc < levelL - 1 || oc > levelL + 1) continue;
            const float uR = kR[iR].x;
            if (uR >= minU && uR <= maxU) {
                const uint4 b0 = reinterpret_cast<const uint4 *>(dRb + (long long)iR * 32)[0], b1 = reinterpret_cast<const uint4 *>(dRb + (long long)iR * 32)[1];
                const unsigned dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                      __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                const unsigned k2 = dist << 16 | (unsigned)iR;            // iR < 65536 (kp_frame_cap is far below)
                if (dist < (key >> 16)) key = k2;                        // ascending iR per lane: strict < keeps the first
            }
        }
        any = __any_sync(0xffffffffu, any);
        key = __reduce_min_sync(0xffffffffu, key);                       // lowest distance, then lowest iR
    }
    if (live && any) {
        const int bestDist = (int)(key >> 16), bestIdxR = (int)(key & 0xffffu);
        if (bestIdxR != 0) di_o = bestIdxR;                              // :948 (index 0 is never recorded)
        if (bestDist < (ORBX_TH_HIGH + ORBX_TH_LOW) / 2) {
            const float uR0 = kR[bestIdxR].x;
            const float sf = sc.inv_scale[levelL];
            const float suL = roundf(__fmul_rn(uL, sf)), svL = roundf(__fmul_rn(vL, sf)), suR0 = roundf(__fmul_rn(uR0, sf));
            const LevelGeom &G = PL->lv[levelL];
            const float iniu = suR0, endu = __fadd_rn(suR0, 11.0f);      // scaleduR0 + L - w, scaleduR0 + L + w + 1 with L = w = 5
            if (!(iniu < 0 || endu >= (float)G.w)) {
                int spL, spR;
                const uint8_t *imL = level_ptr(PL, L.s0, L.frame, levelL, &spL), *imR = level_ptr(PR, R.s0, R.frame, levelL, &spR);
                const int y0 = (int)svL - 5, xl = (int)suL - 5, xr0 = (int)suR0 - 10;
                int *pl = sIL[warp], *pr = pl + 121;
                for (int i = lane; i < 121; i += 32) { const int yy = i / 11, xx = i - yy * 11; pl[i] = imL[(long long)(y0 + yy) * spL + xl + xx]; }
                for (int i = lane; i < 231; i += 32) { const int yy = i / 21, xx = i - yy * 21; pr[i] = imR[(long long)(y0 + yy) * spR + xr0 + xx]; }
                __syncwarp();
                int acc = 0x7fffff;
                if (lane < 11) {                                         // lane = incR + 5
                    const int cl = pl[5 * 11 + 5], cr = pr[5 * 21 + 5 + lane];
                    acc = 0;
                    for (int yy = 0; yy < 11; ++yy)
#pragma unroll
                        for (int xx = 0; xx < 11; ++xx) acc += abs((pl[yy * 11 + xx] - cl) - (pr[yy * 21 + xx + lane] - cr));
                }
                const unsigned k3 = __reduce_min_sync(0xffffffffu, (unsigned)acc << 8 | (unsigned)lane);   // first minimum over incR
                const int bestk = (int)(k3 & 0xffu), bestD = (int)(k3 >> 8);
                const int d1i = __shfl_sync(0xffffffffu, acc, max(bestk - 1, 0)), d3i = __shfl_sync(0xffffffffu, acc, min(bestk + 1, 10));
                if (bestk != 0 && bestk != 10) {
                    const float dist1 = (float)d1i, dist2 = (float)bestD, dist3 = (float)d3i;
                    const float num = __fsub_rn(dist1, dist3);
                    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
                    const float deltaR = __fdiv_rn(num, den);
                    if (!(deltaR < -1 || deltaR > 1)) {
                        const float t = __fadd_rn(__fadd_rn(suR0, (float)(bestk - 5)), deltaR);
                        float bestuR = __fmul_rn(sc.scale[levelL], t);
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0 && disparity < 200.0f) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                            dp_o = __fdiv_rn(bf, disparity);
                            ur_o = bestuR;
                            sad_o = bestD;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { u_right[iL] = ur_o; depth[iL] = dp_o; desc_index[iL] = di_o; sad_out[iL] = sad_o; }
}

// Median cut (:1022-1037): thDist = 1.5f*1.4f*median of the SADs of the pushed matches (the element of rank n/2 of the
// sorted list); every pushed match with SAD >= thDist is invalidated.  SAD <= 121*510 < 2^16: two 256-bin histograms.
__global__ void __launch_bounds__(1024)
k_stereo_median(const DevParams *__restrict__ PL, int frame, float *__restrict__ u_right, float *__restrict__ depth,
                int32_t *__restrict__ desc_index, const int32_t *__restrict__ sad, int *__restrict__ kept_out)
{
    __shared__ int hist[256], s_bin, s_rank, s_total, s_kept;
    __shared__ float s_th;
    const int n = PL->out_n[frame];
    for (int pass = 0; pass < 2; ++pass) {
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        if (threadIdx.x == 0 && pass == 0) s_kept = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int v = sad[i];
            if (v < 0) continue;
            if (pass == 0) atomicAdd(&hist[v >> 8], 1);
            else if ((v >> 8) == s_bin) atomicAdd(&hist[v & 255], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (pass == 0) {
                int tot = 0;
                for (int b = 0; b < 256; ++b) tot += hist[b];
                s_total = tot;
                int rank = tot / 2, b = 0;
                while (b < 255 && rank >= hist[b]) { rank -= hist[b]; ++b; }
                s_bin = b; s_rank = rank;
            } else {
                int rank = s_rank, b = 0;
                while (b < 255 && rank >= hist[b]) { rank -= hist[b]; ++b; }
                const float median = (float)(s_bin << 8 | b);
                s_th = __fmul_rn(__fmul_rn(1.5f, 1.4f), median);
            }
        }
        __syncthreads();
        if (s_total == 0) { if (threadIdx.x == 0 && kept_out) *kept_out = -1; return; }   // the reference reads an empty vector here
    }
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int v = sad[i];
        if (v < 0) continue;
        if (!((float)v < s_th)) { u_right[i] = -1.0f; depth[i] = -1.0f; desc_index[i] = -1; }
        else ++mine;
    }
    mine = warp_sum(mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_kept, mine);
    __syncthreads();
    if (threadIdx.x == 0 && kept_out) *kept_out = s_kept;
}

cudaError_t launch_stereo(const DevParams *dPL, Src0 s0L, int frameL, const DevParams *dPR, Src0 s0R, int frameR, const StereoScales &sc, float bf,
                          int capL, float *d_u_right, float *d_depth, int32_t *d_desc_index, int32_t *d_sad, int *d_kept, cudaStream_t st, LaunchStats *ls)
{
    StereoSide L = {dPL, s0L, frameL}, R = {dPR, s0R, frameR};
    k_stereo_match<<<(capL + 7) / 8, 256, 0, st>>>(L, R, sc, bf, d_u_right, d_depth, d_desc_index, d_sad);
    k_stereo_median<<<1, 1024, 0, st>>>(dPL, frameL, d_u_right, d_depth, d_desc_index, d_sad, d_kept);
    ls->launches += 2;
    return cudaGetLastError();
}

int match_chunks(int nA, int nB)
{
    if (nB <= 0) return 1;
    const int row_blocks = (nA + kMatchThreads - 1) / kMatchThreads;
    // enough CTAs that the 148 SMs end together: at least ~16 CTAs per SM overall (5000 x 5000: 40 row blocks x 40 chunks of one
    // 128-row tile; with 15 chunks of three tiles the 600 CTAs left some SMs with 5 CTAs and others with 4)
    int want = (148 * 16 + row_blocks - 1) / (row_blocks > 0 ? row_blocks : 1);
    const int max_chunks = (nB + kMatchTile - 1) / kMatchTile;
    if (want < 1) want = 1;
    if (want > max_chunks) want = max_chunks;
    return want;
}

cudaError_t launch_match(const uint32_t *dA, int nA, const uint32_t *dB, int nB, int th, float ratio,
                         int32_t *d_idx, int32_t *d_d1, int32_t *d_d2, uint8_t *d_accept, int *d_naccept,
                         int4 *d_partial, unsigned *d_arrive, int nchunks, cudaStream_t st, LaunchStats *ls)
{
    if (nA <= 0) return cudaSuccess;
    int chunk = nB > 0 ? (nB + nchunks - 1) / nchunks : 1;
    chunk = (chunk + kMatchTile - 1) / kMatchTile * kMatchTile;
    dim3 grid((nA + kMatchThreads - 1) / kMatchThreads, nchunks);
    k_match_partial<<<grid, kMatchThreads, 0, st>>>(dA, nA, dB, nB, chunk, d_partial, d_arrive, th, ratio, d_idx, d_d1, d_d2, d_accept, d_naccept);
    ls->launches += 1;
    return cudaGetLastError();
}

} // namespace orbx
