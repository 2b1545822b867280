// multimot_track_b200/csrc/orbx_api.cu -- the C ABI of include/orbx.h: handle, buffers,
// stream orchestration.  No compute happens on the host: an entry point either
// runs the CUDA kernels of kernels.cu or fails with ORBX_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace orbx;

static const int8_t kPatternHost[1024] = {
#include "orb_pattern.inc"
};

struct orbx_handle {
    Tables tab;
    Geometry geo;
    bool geo_valid = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    int batch_cap = 0;                  // frames the per-batch buffers hold
    DevParams hp;                       // host copy of the device parameter block
    DevParams *d_params = nullptr;
    // geometry-sized device tables
    ResizeTab *d_xtab = nullptr, *d_ytab = nullptr;
    uint32_t *d_blur_work = nullptr, *d_ffast_work = nullptr, *d_oct_lut = nullptr;
    int8_t *d_pattern = nullptr;
    // batch-sized device buffers
    uint8_t *d_pyr = nullptr, *d_blur = nullptr;
    uint32_t *d_cand = nullptr, *d_cand_count = nullptr, *d_kp_stage = nullptr, *d_kp_count = nullptr;
    unsigned long long *d_sort = nullptr;
    orbx_keypoint *d_out_kps = nullptr;
    uint32_t *d_out_ckps = nullptr;     // compact records beside the full ones (opt_compact)
    orbx_keypoint_compact *p_ckps = nullptr;
    uint8_t *d_out_desc = nullptr;
    int *d_out_n = nullptr;
    uint8_t *d_pad = nullptr; size_t pad_bytes = 0;
    uint8_t *d_scratch = nullptr; size_t scratch_bytes = 0;      // grow-only scratch of the synchronous helper calls (matcher filters, stereo, BoW, ...)
    uint8_t *d_in = nullptr; size_t in_bytes = 0;       // H2D landing area for host frames before the repack kernel
    // pinned host staging of the results
    orbx_keypoint *p_kps = nullptr;
    uint8_t *p_desc = nullptr;
    int *p_n = nullptr;
    // state of the last / pending batch
    bool pending = false, have_batch = false;
    int last_nframes = 0;
    Src0 last_src0 = {nullptr, 0, 0};
    // matcher scratch
    uint32_t *m_A = nullptr, *m_B = nullptr; size_t m_capA = 0, m_capB = 0;
    int32_t *m_out = nullptr; uint8_t *m_acc = nullptr; int *m_nacc = nullptr; size_t m_cap_out = 0;
    int4 *m_partial = nullptr; size_t m_cap_partial = 0;
    unsigned *m_arrive = nullptr; size_t m_cap_arrive = 0;         // arrival counters of the matcher's fused merge (zero between calls)
    TmaMaps tma;                        // pyramid source descriptors (levels >= 2 from our buffer, level 1 from the batch's level 0)
    const void *tma_l0_base = nullptr; long long tma_l0_fs = 0; int tma_l0_pitch = 0, tma_l0_frames = 0;
    bool use_tma = true;
    bool opt_tma = true, opt_fast_tma = false, opt_copy_input = false, opt_compact = false;   // orbx_set_option
    FastMaps fast_maps; unsigned fast_ok = 0;               // FAST raw boxes (bit l = level l encoded)
    OdMaps od_maps; unsigned od_img_ok = 0, od_blr_ok = 0;  // orientation / descriptor patch boxes (bit l = level l encoded)
    BlurMaps blur_maps; unsigned blur_tma_levels = 0;       // blur source boxes: bit l set = maps.m[l] valid (level 0 per batch source)
    LaunchStats stats;
    bool profiling = false;
    cudaEvent_t ev[ORBX_NUM_STAGES + 1] = {};
    bool ev_valid = false;
    std::string err;
};

static thread_local std::string g_create_error;

namespace {

int fail(orbx_handle *h, int code, const std::string &msg) { if (h) h->err = msg; else g_create_error = msg; return code; }

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char b_[512];                                                                                \
            std::snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(h, e_ == cudaErrorMemoryAllocation ? ORBX_ERR_OOM : ORBX_ERR_CUDA, b_);          \
        }                                                                                                \
    } while (0)

template <typename T> void dfree(T *&p) { if (p) cudaFree(p); p = nullptr; }

// Entry points run on the handle's device and restore the caller's current device on the way out (multi-GPU callers keep
// their own device selection).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = err == cudaSuccess; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define ON_DEVICE(hh)                                                                                    \
    DeviceGuard device_guard_((hh)->device);                                                             \
    if (device_guard_.err != cudaSuccess)                                                                \
        return fail(h, ORBX_ERR_CUDA, std::string("cudaSetDevice failed: ") + cudaGetErrorString(device_guard_.err))

// Grow-only device scratch of a handle (the calls that use it are synchronous, so one buffer is enough).
int get_scratch(orbx_handle *h, size_t bytes, void **out)
{
    if (bytes > h->scratch_bytes) {
        cudaStreamSynchronize(h->stream);
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr; h->scratch_bytes = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        const cudaError_t e = cudaMalloc(&h->d_scratch, want);
        if (e != cudaSuccess) { h->err = std::string("scratch allocation failed: ") + cudaGetErrorString(e); return ORBX_ERR_OOM; }
        h->scratch_bytes = want;
    }
    *out = h->d_scratch;
    return ORBX_OK;
}
template <typename T> void hfree(T *&p) { if (p) cudaFreeHost(p); p = nullptr; }

void free_batch_buffers(orbx_handle *h)
{
    dfree(h->d_pyr); dfree(h->d_blur); dfree(h->d_cand); dfree(h->d_cand_count); dfree(h->d_kp_stage);
    dfree(h->d_kp_count); dfree(h->d_sort); dfree(h->d_out_kps); dfree(h->d_out_desc); dfree(h->d_out_n);
    dfree(h->d_out_ckps);
    hfree(h->p_kps); hfree(h->p_desc); hfree(h->p_n); hfree(h->p_ckps);
    h->batch_cap = 0;
}

void free_geo_tables(orbx_handle *h) { dfree(h->d_xtab); dfree(h->d_ytab); dfree(h->d_blur_work); dfree(h->d_ffast_work); dfree(h->d_oct_lut); }

int upload_params(orbx_handle *h)
{
    DevParams &P = h->hp;
    P.pyr = h->d_pyr; P.blur = h->d_blur; P.cand = h->d_cand; P.cand_count = h->d_cand_count;
    P.kp_stage = h->d_kp_stage; P.kp_count = h->d_kp_count; P.sort_scratch = h->d_sort;
    P.out_kps = h->d_out_kps; P.out_desc = h->d_out_desc; P.out_n = h->d_out_n; P.out_ckps = h->d_out_ckps;
    P.xtab = h->d_xtab; P.ytab = h->d_ytab; P.blur_work = h->d_blur_work; P.ffast_work = h->d_ffast_work; P.oct_lut = h->d_oct_lut;
    P.pattern = h->d_pattern;
    CU(cudaMemcpyAsync(h->d_params, &P, sizeof(DevParams), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));      // hp is reused; keep it simple: params change rarely
    return ORBX_OK;
}

int ensure_batch(orbx_handle *h, int nframes)
{
    if (nframes <= h->batch_cap) return ORBX_OK;
    CU(cudaStreamSynchronize(h->stream));
    free_batch_buffers(h);
    const Geometry &g = h->geo;
    const size_t F = (size_t)nframes, L = (size_t)g.nlevels;
    CU(cudaMalloc(&h->d_pyr, F * g.pyr_frame_bytes + 256));       // +256: word-granular tile staging may read a few bytes past the last row
    CU(cudaMalloc(&h->d_blur, F * g.pyr_frame_bytes + 256));
    CU(cudaMalloc(&h->d_cand, F * g.cand_frame_elems * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_sort, F * g.cand_frame_elems * 2 * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_cand_count, F * L * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_kp_count, F * L * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_kp_stage, F * g.kp_frame_cap * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_out_kps, F * g.kp_frame_cap * sizeof(orbx_keypoint)));
    CU(cudaMalloc(&h->d_out_desc, F * g.kp_frame_cap * 32));
    CU(cudaMalloc(&h->d_out_n, F * sizeof(int)));
    CU(cudaMallocHost(&h->p_kps, F * g.kp_frame_cap * sizeof(orbx_keypoint)));
    CU(cudaMallocHost(&h->p_desc, F * g.kp_frame_cap * 32));
    CU(cudaMallocHost(&h->p_n, F * sizeof(int)));
    if (h->opt_compact) {
        CU(cudaMalloc(&h->d_out_ckps, F * g.kp_frame_cap * sizeof(orbx_keypoint_compact)));
        CU(cudaMallocHost(&h->p_ckps, F * g.kp_frame_cap * sizeof(orbx_keypoint_compact)));
    }
    {   // landing area of the host-frame entry points (tightly packed grey frames; grows on demand for padded / colour input), so that
        // the first orbx_submit_host on a pre-sized handle does not allocate -- and synchronise the device -- inside a caller's pipeline
        const size_t want = F * (size_t)g.width * (size_t)g.height;
        if (want > h->in_bytes) { dfree(h->d_in); CU(cudaMalloc(&h->d_in, want)); h->in_bytes = want; }
    }
    // rows beyond a level's width are padding that kernels may write but never read as data;
    // clear once so read-back of padded rows is deterministic
    CU(cudaMemsetAsync(h->d_pyr, 0, F * g.pyr_frame_bytes, h->stream));
    CU(cudaMemsetAsync(h->d_blur, 0, F * g.pyr_frame_bytes, h->stream));
    h->batch_cap = nframes;
    // TMA descriptors of the pyramid levels that live in our own buffer (source of level l is level l-1 >= 1)
    h->use_tma = h->opt_tma;
    for (int l = 0; l < kMaxLevels; ++l) h->tma.ok[l] = false;
    h->tma_l0_base = nullptr;
    for (int l = 2; l < g.nlevels && h->use_tma; ++l) {
        resize_box(g.lv[l - 1], g.lv[l], &h->tma.box_w[l], &h->tma.box_h[l]);
        h->tma.ok[l] = encode_image_map(&h->tma.src[l], h->d_pyr + g.lv[l - 1].img_off, g.lv[l - 1].pitch, g.lv[l - 1].h,
                                        g.pyr_frame_bytes, nframes, h->tma.box_w[l], h->tma.box_h[l]);
    }
    // blur source boxes: one descriptor per level of our pyramid buffer (level 0 follows the batch's source, enqueue_pipeline)
    h->blur_tma_levels = 0; h->od_img_ok = 0; h->od_blr_ok = 0; h->fast_ok = 0;
    std::memset(&h->fast_maps, 0, sizeof(h->fast_maps));
    for (int l = 1; l < g.nlevels && h->use_tma; ++l)
        if (encode_image_map(&h->fast_maps.m[l], h->d_pyr + g.lv[l].img_off, g.lv[l].pitch, g.lv[l].h, g.pyr_frame_bytes, nframes, kFfBoxW, g.lv[l].h_cell + 6))
            h->fast_ok |= 1u << l;
    std::memset(&h->blur_maps, 0, sizeof(h->blur_maps));
    std::memset(&h->od_maps, 0, sizeof(h->od_maps));
    for (int l = 0; l < g.nlevels && h->use_tma; ++l) {
        if (l >= 1 && encode_image_map(&h->od_maps.img[l], h->d_pyr + g.lv[l].img_off, g.lv[l].pitch, g.lv[l].h, g.pyr_frame_bytes, nframes, kOdIcBoxW, kOdIcBoxH))
            h->od_img_ok |= 1u << l;
        if (encode_image_map(&h->od_maps.blr[l], h->d_blur + g.lv[l].img_off, g.lv[l].pitch, g.lv[l].h, g.pyr_frame_bytes, nframes, kOdBlBoxW, kOdBlBoxH))
            h->od_blr_ok |= 1u << l;
    }
    for (int l = 1; l < g.nlevels && h->use_tma; ++l)
        if (encode_image_map(&h->blur_maps.m[l], h->d_pyr + g.lv[l].img_off, g.lv[l].pitch, g.lv[l].h, g.pyr_frame_bytes, nframes, kBlurBoxW, kBlurBoxH))
            h->blur_tma_levels |= 1u << l;
    return upload_params(h);
}

int ensure_geometry(orbx_handle *h, int width, int height, int nframes)
{
    if (h->geo_valid && h->geo.width == width && h->geo.height == height) return ensure_batch(h, nframes);
    CU(cudaStreamSynchronize(h->stream));
    Geometry g;
    std::string err;
    const int rc = build_geometry(h->tab, width, height, &g, &err);
    if (rc != ORBX_OK) return fail(h, rc, err);
    if (g.max_node_cap > 8192) {
        char b[160];
        std::snprintf(b, sizeof b, "nfeatures too large: a level asks for %d octree nodes (limit 8192)", g.max_node_cap);
        return fail(h, ORBX_ERR_UNSUPPORTED, b);
    }
    h->geo_valid = false; h->have_batch = false;
    const int keep_cap = h->batch_cap;
    free_batch_buffers(h);
    free_geo_tables(h);
    h->geo = g;
    CU(cudaMalloc(&h->d_xtab, std::max<size_t>(g.xtab.size(), 4) * sizeof(ResizeTab)));
    CU(cudaMalloc(&h->d_ytab, std::max<size_t>(g.ytab.size(), 4) * sizeof(ResizeTab)));
    CU(cudaMalloc(&h->d_blur_work, std::max<size_t>(g.blur_work.size(), 1) * sizeof(uint32_t)));
    if (!g.xtab.empty()) CU(cudaMemcpyAsync(h->d_xtab, h->geo.xtab.data(), g.xtab.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, h->stream));
    if (!g.ytab.empty()) CU(cudaMemcpyAsync(h->d_ytab, h->geo.ytab.data(), g.ytab.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMalloc(&h->d_ffast_work, std::max<size_t>(g.ffast_work.size(), 1) * sizeof(uint32_t)));
    if (!g.ffast_work.empty()) CU(cudaMemcpyAsync(h->d_ffast_work, h->geo.ffast_work.data(), g.ffast_work.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMalloc(&h->d_oct_lut, std::max<size_t>(g.oct_lut.size(), 1) * sizeof(uint32_t)));
    if (!g.oct_lut.empty()) CU(cudaMemcpyAsync(h->d_oct_lut, h->geo.oct_lut.data(), g.oct_lut.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    if (!g.blur_work.empty()) CU(cudaMemcpyAsync(h->d_blur_work, h->geo.blur_work.data(), g.blur_work.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));

    DevParams &P = h->hp;
    std::memset(&P, 0, sizeof(P));
    P.nlevels = g.nlevels; P.ini_th = h->tab.ini_th; P.min_th = h->tab.min_th;
    P.kp_frame_cap = g.kp_frame_cap; P.pyr_frame_bytes = g.pyr_frame_bytes; P.cand_frame_elems = g.cand_frame_elems;
    for (int l = 0; l < g.nlevels; ++l) { P.lv[l] = g.lv[l]; P.xtab_off[l] = g.xtab_off[l]; P.ytab_off[l] = g.ytab_off[l]; }
    for (int i = 0; i < 16; ++i) P.umax[i] = h->tab.umax[i];
    P.n_blur_work = (int)g.blur_work.size(); P.n_ffast_work = (int)g.ffast_work.size();
    CU(cudaStreamSynchronize(h->stream));      // the table uploads above (sources live in h->geo) are complete before any kernel can read them
    h->geo_valid = true;
    return ensure_batch(h, std::max(nframes, keep_cap));
}

bool aligned16(const void *p, long long a, long long b) { return ((uintptr_t)p % 16 == 0) && a % 16 == 0 && b % 16 == 0; }

// Queue the whole front end for a batch whose level-0 images are described by s0.
int enqueue_pipeline(orbx_handle *h, Src0 s0, int nframes)
{
    const DevParams &P = h->hp;
    cudaStream_t st = h->stream;
    const bool prof = h->profiling;
#define MARK(i) do { if (prof) CU(cudaEventRecord(h->ev[i], st)); } while (0)
    MARK(1);                                     // ev[0] was recorded by the caller before the input copies
    // measurement aid only (scripts/exp_overlap.py): leave stages out to read their marginal cost in the overlapped pipeline;
    // the buffers keep the previous batch's contents, results are then meaningless.  1 pyramid, 2 blur, 4 FAST, 8 octree, 16 orient/desc, 32 D2H
#ifdef ORBX_DEBUG_KNOBS
    const char *skip_env = std::getenv("ORBX_DEBUG_SKIP");
    const int skip = skip_env ? std::atoi(skip_env) : 0;
#else
    constexpr int skip = 0;                      // the release library has no such knob: every stage always runs
#endif
    if (!(skip & 4))
    CU(cudaMemsetAsync(h->d_cand_count, 0, (size_t)nframes * P.nlevels * sizeof(uint32_t), st));
    if (h->use_tma &&
        (h->tma_l0_base != s0.ptr || h->tma_l0_fs != s0.frame_stride || h->tma_l0_pitch != s0.pitch || h->tma_l0_frames < nframes)) {
        // level 1 reads the batch's level-0 images: the caller's frames in place, or our level-0 slots
        if (P.nlevels > 1) resize_box(h->geo.lv[0], h->geo.lv[1], &h->tma.box_w[1], &h->tma.box_h[1]);
        if (P.nlevels > 1) h->tma.ok[1] = encode_image_map(&h->tma.src[1], s0.ptr, s0.pitch, h->geo.lv[0].h, s0.frame_stride, nframes, h->tma.box_w[1], h->tma.box_h[1]);
        if (encode_image_map(&h->blur_maps.m[0], s0.ptr, s0.pitch, h->geo.lv[0].h, s0.frame_stride, nframes, kBlurBoxW, kBlurBoxH)) h->blur_tma_levels |= 1u;
        else h->blur_tma_levels &= ~1u;
        if (encode_image_map(&h->fast_maps.m[0], s0.ptr, s0.pitch, h->geo.lv[0].h, s0.frame_stride, nframes, kFfBoxW, h->geo.lv[0].h_cell + 6)) h->fast_ok |= 1u;
        else h->fast_ok &= ~1u;
        if (encode_image_map(&h->od_maps.img[0], s0.ptr, s0.pitch, h->geo.lv[0].h, s0.frame_stride, nframes, kOdIcBoxW, kOdIcBoxH)) h->od_img_ok |= 1u;
        else h->od_img_ok &= ~1u;
        h->tma_l0_base = s0.ptr; h->tma_l0_fs = s0.frame_stride; h->tma_l0_pitch = s0.pitch; h->tma_l0_frames = nframes;
    }
    if (!(skip & 1))
    CU(launch_pyramid(h->d_params, P, s0, nframes, st, &h->stats, h->use_tma ? &h->tma : nullptr));
    MARK(2);
    if (!(skip & 2))
    CU(launch_blur(h->d_params, P, s0, nframes, st, &h->stats, h->use_tma ? &h->blur_maps : nullptr, h->blur_tma_levels));
    MARK(3);
    {
        const unsigned all = P.nlevels >= 32 ? 0xffffffffu : (1u << P.nlevels) - 1u;
        const bool fast_tma = h->use_tma && (h->fast_ok & all) == all && h->opt_fast_tma;   // opt-in (ORBX_OPT_FAST_TMA): measured slower in the pipelined loop
        if (!(skip & 4))
        CU(launch_fast(h->d_params, P, s0, nframes, h->geo.n_ffast_small, st, &h->stats, fast_tma ? &h->fast_maps : nullptr));
    }
    MARK(4);
    if (!(skip & 8))
    CU(launch_octree(h->d_params, P, nframes, h->geo.max_node_cap, h->geo.max_feat, st, &h->stats));
    MARK(5);
    {
        const unsigned all = P.nlevels >= 32 ? 0xffffffffu : (1u << P.nlevels) - 1u;
        const bool od_tma = h->use_tma && (h->od_img_ok & all) == all && (h->od_blr_ok & all) == all;
        if (!(skip & 16))
        CU(launch_orient_desc(h->d_params, P, s0, nframes, st, &h->stats, od_tma ? &h->od_maps : nullptr));
    }
    MARK(6);
    const size_t cap = (size_t)P.kp_frame_cap;
    CU(cudaMemcpyAsync(h->p_n, h->d_out_n, (size_t)nframes * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (!(skip & 32)) {
    if (h->opt_compact) CU(cudaMemcpyAsync(h->p_ckps, h->d_out_ckps, (size_t)nframes * cap * sizeof(orbx_keypoint_compact), cudaMemcpyDeviceToHost, st));
    else CU(cudaMemcpyAsync(h->p_kps, h->d_out_kps, (size_t)nframes * cap * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->p_desc, h->d_out_desc, (size_t)nframes * cap * 32, cudaMemcpyDeviceToHost, st));
    }
    MARK(7);
#undef MARK
    h->ev_valid = prof;
    h->pending = true; h->have_batch = true; h->last_nframes = nframes; h->last_src0 = s0;
    return ORBX_OK;
}

int check_shape(orbx_handle *h, int nframes, int width, int height, int stride)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (nframes <= 0 || width <= 0 || height <= 0 || stride < width) return fail(h, ORBX_ERR_BAD_ARG, "bad frame shape / stride / count");
    return ORBX_OK;
}

} // namespace

// ------------------------------------------------------------------ lifetime

extern "C" int orbx_create(const orbx_config *cfg, orbx_handle **out)
{
    if (!cfg || !out) return fail(nullptr, ORBX_ERR_BAD_ARG, "orbx_create: NULL argument");
    *out = nullptr;
    if (cfg->nlevels < 1 || cfg->nlevels > ORBX_MAX_LEVELS || cfg->nfeatures < 1 || !(cfg->scale_factor >= 1.0f) ||
        cfg->ini_th_fast < 1 || cfg->min_th_fast < 1 || cfg->ini_th_fast > 255 || cfg->min_th_fast > 255)
        return fail(nullptr, ORBX_ERR_BAD_ARG, "orbx_create: nlevels in 1..16, nfeatures >= 1, scale_factor >= 1, thresholds in 1..255 required");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ORBX_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    orbx_handle *h = new (std::nothrow) orbx_handle();
    if (!h) return fail(nullptr, ORBX_ERR_OOM, "host allocation failed");
    int dev = cfg->device_id;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= ndev) { delete h; return fail(nullptr, ORBX_ERR_BAD_ARG, "device_id out of range"); }
    h->device = dev;
    build_tables(cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th_fast, cfg->min_th_fast, &h->tab);
    int rc = ORBX_OK;
    DeviceGuard device_guard_(dev);
    auto init = [&]() -> int {
        if (device_guard_.err != cudaSuccess) return fail(h, ORBX_ERR_CUDA, std::string("cudaSetDevice failed: ") + cudaGetErrorString(device_guard_.err));
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&h->d_params, sizeof(DevParams)));
        CU(cudaMalloc(&h->d_pattern, sizeof(kPatternHost)));
        CU(cudaMemcpyAsync(h->d_pattern, kPatternHost, sizeof(kPatternHost), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (cfg->max_width > 0 && cfg->max_height > 0)
            return ensure_geometry(h, cfg->max_width, cfg->max_height, std::max(cfg->max_batch, 1));
        return ORBX_OK;
    };
    rc = init();
    if (rc != ORBX_OK) { g_create_error = h->err; orbx_destroy(h); return rc; }
    *out = h;
    return ORBX_OK;
}

extern "C" void orbx_destroy(orbx_handle *h)
{
    if (!h) return;
    DeviceGuard device_guard_(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_batch_buffers(h);
    free_geo_tables(h);
    dfree(h->d_params); dfree(h->d_pattern); dfree(h->d_pad); dfree(h->d_in); dfree(h->d_scratch);
    dfree(h->m_A); dfree(h->m_B); dfree(h->m_out); dfree(h->m_acc); dfree(h->m_nacc); dfree(h->m_partial); dfree(h->m_arrive);
    for (int i = 0; i <= ORBX_NUM_STAGES; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *orbx_last_error(const orbx_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" const char *orbx_version(void) { return "orbx 0.1 (sm_100a)"; }

extern "C" int orbx_get_tables(const orbx_handle *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int32_t *nfeat)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    for (int i = 0; i < h->tab.nlevels; ++i) {
        if (scale) scale[i] = h->tab.scale[i];
        if (inv_scale) inv_scale[i] = h->tab.inv_scale[i];
        if (sigma2) sigma2[i] = h->tab.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->tab.inv_sigma2[i];
        if (nfeat) nfeat[i] = h->tab.nfeat[i];
    }
    return h->tab.nlevels;
}

extern "C" int orbx_max_keypoints(orbx_handle *h, int width, int height)
{
    if (!h || width <= 0 || height <= 0) return ORBX_ERR_BAD_ARG;
    if (h->geo_valid && h->geo.width == width && h->geo.height == height) return h->geo.kp_frame_cap;
    Geometry g;
    std::string err;
    const int rc = build_geometry(h->tab, width, height, &g, &err);
    if (rc != ORBX_OK) return fail(h, rc, err);
    return g.kp_frame_cap;
}

extern "C" int orbx_make_plan(const orbx_config *cfg, int width, int height, orbx_plan *out)
{
    if (!cfg || !out || width <= 0 || height <= 0 || cfg->nlevels < 1 || cfg->nlevels > ORBX_MAX_LEVELS || cfg->nfeatures < 1 ||
        !(cfg->scale_factor >= 1.0f))
        return fail(nullptr, ORBX_ERR_BAD_ARG, "orbx_make_plan: bad argument");
    Tables t;
    build_tables(cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th_fast, cfg->min_th_fast, &t);
    Geometry g;
    std::string err;
    const int rc = build_geometry(t, width, height, &g, &err);
    if (rc != ORBX_OK) return fail(nullptr, rc, err);
    std::memset(out, 0, sizeof(*out));
    out->nlevels = g.nlevels; out->max_keypoints = g.kp_frame_cap;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom &L = g.lv[l];
        out->level_width[l] = L.w; out->level_height[l] = L.h; out->nfeatures_per_level[l] = L.n_feat;
        out->cell_cols[l] = L.cols_vis; out->cell_rows[l] = L.rows_vis; out->cell_w[l] = L.w_cell; out->cell_h[l] = L.h_cell;
        out->octree_roots[l] = L.n_ini; out->max_candidates[l] = L.cand_cap;
        out->scale[l] = t.scale[l]; out->inv_scale[l] = t.inv_scale[l]; out->sigma2[l] = t.sigma2[l]; out->inv_sigma2[l] = t.inv_sigma2[l];
        out->keypoint_size[l] = L.kp_size;
    }
    for (int i = 0; i < 16; ++i) out->umax[i] = t.umax[i];
    return ORBX_OK;
}

// ------------------------------------------------------------------ extract

extern "C" int orbx_submit_device(orbx_handle *h, const uint8_t *d_frames, int nframes, int width, int height,
                                  int stride_bytes, size_t frame_stride_bytes)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!d_frames) return fail(h, ORBX_ERR_BAD_ARG, "NULL frame pointer");
    int rc = check_shape(h, nframes, width, height, stride_bytes);
    if (rc != ORBX_OK) return rc;
    if (frame_stride_bytes < (size_t)stride_bytes * (size_t)(height - 1) + (size_t)width) return fail(h, ORBX_ERR_BAD_ARG, "frame_stride_bytes too small");
    ON_DEVICE(h);
    if (h->pending) return fail(h, ORBX_ERR_STATE, "a batch is already pending on this handle: call orbx_collect first");
    rc = ensure_geometry(h, width, height, nframes);
    if (rc != ORBX_OK) return rc;
    if (h->profiling) CU(cudaEventRecord(h->ev[0], h->stream));
    Src0 s0;
    if (!h->opt_copy_input && aligned16(d_frames, stride_bytes, (long long)frame_stride_bytes)) {
        s0.ptr = d_frames; s0.pitch = stride_bytes; s0.frame_stride = (long long)frame_stride_bytes;   // used in place
    } else {
        const LevelGeom &L0 = h->geo.lv[0];
        CU(launch_repack(d_frames, (long long)frame_stride_bytes, stride_bytes, h->d_pyr + L0.img_off, h->geo.pyr_frame_bytes, L0.pitch,
                         width, height, nframes, h->stream, &h->stats));
        s0.ptr = h->d_pyr + L0.img_off; s0.pitch = L0.pitch; s0.frame_stride = h->geo.pyr_frame_bytes;
    }
    return enqueue_pipeline(h, s0, nframes);
}

// channels == 1: grey frames; 3 / 4: colour frames converted on the device (rgb_order as Camera.RGB, src/Tracking.cc:192-193)
static int submit_host_frames(orbx_handle *h, const uint8_t *const *frames, int nframes, int width, int height, int stride_bytes,
                              int channels, int rgb_order)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!frames) return fail(h, ORBX_ERR_BAD_ARG, "NULL frame array");
    if (channels != 1 && channels != 3 && channels != 4) return fail(h, ORBX_ERR_BAD_ARG, "channels must be 1, 3 or 4");
    int rc = check_shape(h, nframes, width, height, stride_bytes / channels);
    if (rc != ORBX_OK) return rc;
    for (int f = 0; f < nframes; ++f) if (!frames[f]) return fail(h, ORBX_ERR_BAD_ARG, "NULL frame pointer");
    ON_DEVICE(h);
    if (h->pending) return fail(h, ORBX_ERR_STATE, "a batch is already pending on this handle: call orbx_collect first");
    rc = ensure_geometry(h, width, height, nframes);
    if (rc != ORBX_OK) return rc;
    if (h->profiling) CU(cudaEventRecord(h->ev[0], h->stream));
    const LevelGeom &L0 = h->geo.lv[0];
    // H2D as plain 1-D copies (one per frame, or one for the whole batch when the frames are
    // contiguous in host memory), then a device kernel re-pitches rows into the aligned level-0 slots
    const size_t frame_bytes = (size_t)stride_bytes * (size_t)(height - 1) + (size_t)width * channels;
    const size_t slot = (size_t)stride_bytes * (size_t)height;
    if (slot * (size_t)nframes > h->in_bytes) {
        CU(cudaStreamSynchronize(h->stream));
        dfree(h->d_in);
        CU(cudaMalloc(&h->d_in, slot * (size_t)nframes));
        h->in_bytes = slot * (size_t)nframes;
    }
    bool contiguous = true;
    for (int f = 1; f < nframes; ++f) contiguous = contiguous && frames[f] == frames[0] + (size_t)f * slot;
    if (contiguous)
        CU(cudaMemcpyAsync(h->d_in, frames[0], slot * (size_t)(nframes - 1) + frame_bytes, cudaMemcpyHostToDevice, h->stream));
    else
        for (int f = 0; f < nframes; ++f)
            CU(cudaMemcpyAsync(h->d_in + (size_t)f * slot, frames[f], frame_bytes, cudaMemcpyHostToDevice, h->stream));
    if (channels == 1)
        CU(launch_repack(h->d_in, (long long)slot, stride_bytes, h->d_pyr + L0.img_off, h->geo.pyr_frame_bytes, L0.pitch, width, height,
                         nframes, h->stream, &h->stats));
    else
        CU(launch_gray(h->d_in, (long long)slot, stride_bytes, channels, rgb_order, h->d_pyr + L0.img_off, h->geo.pyr_frame_bytes, L0.pitch,
                       width, height, nframes, h->stream, &h->stats));
    Src0 s0;
    s0.ptr = h->d_pyr + L0.img_off; s0.pitch = L0.pitch; s0.frame_stride = h->geo.pyr_frame_bytes;
    return enqueue_pipeline(h, s0, nframes);
}

extern "C" int orbx_submit_host(orbx_handle *h, const uint8_t *const *frames, int nframes, int width, int height, int stride_bytes)
{
    return submit_host_frames(h, frames, nframes, width, height, stride_bytes, 1, 0);
}

extern "C" int orbx_extract_color(orbx_handle *h, const uint8_t *image, int width, int height, int stride_bytes, int channels, int rgb_order,
                                  orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!n_out) return fail(h, ORBX_ERR_BAD_ARG, "NULL n_out");
    if (!image || width <= 0 || height <= 0) { *n_out = 0; return ORBX_OK; }      // empty image: silent return (src/ORBextractor.cc:1049)
    const uint8_t *frames[1] = {image};
    int rc = submit_host_frames(h, frames, 1, width, height, stride_bytes, channels, rgb_order);
    if (rc != ORBX_OK) return rc;
    return orbx_collect(h, kps, desc, cap, n_out);
}

static void expand_keypoints(const orbx_handle *h, const orbx_keypoint_compact *in, int n, orbx_keypoint *out)
{
    const int L = h->tab.nlevels;
    for (int i = 0; i < n; ++i) {
        const orbx_keypoint_compact c = in[i];
        const int o = c.octave < L ? c.octave : L - 1;
        orbx_keypoint k;
        k.x = (float)c.x * h->tab.scale[o]; k.y = (float)c.y * h->tab.scale[o];        // pt *= mvScaleFactor[level] (src/ORBextractor.cc:1098-1104); scale[0] == 1
        k.size = (float)(int)(31.0f * h->tab.scale[o]);                                  // PATCH_SIZE * mvScaleFactor[level] into the int scaledPatchSize (:841-846)
        k.angle = c.angle; k.response = (float)c.response; k.octave = c.octave; k.class_id = -1;
        out[i] = k;
    }
}

extern "C" int orbx_expand_keypoints(const orbx_handle *h, const orbx_keypoint_compact *in, int n, orbx_keypoint *out)
{
    if (!h || n < 0 || (n > 0 && (!in || !out))) return ORBX_ERR_BAD_ARG;
    expand_keypoints(h, in, n, out);
    return ORBX_OK;
}

extern "C" int orbx_collect_view_compact(orbx_handle *h, const orbx_keypoint_compact **ckps, const uint8_t **desc, const int **n_out, int *cap_per_frame)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!h->pending) return fail(h, ORBX_ERR_STATE, "orbx_collect without a pending submit");
    if (!h->opt_compact) return fail(h, ORBX_ERR_STATE, "orbx_collect_view_compact: ORBX_OPT_COMPACT_KEYPOINTS is off on this handle");
    ON_DEVICE(h);
    CU(cudaStreamSynchronize(h->stream));
    h->pending = false;
    if (ckps) *ckps = h->p_ckps;
    if (desc) *desc = h->p_desc;
    if (n_out) *n_out = h->p_n;
    if (cap_per_frame) *cap_per_frame = h->geo.kp_frame_cap;
    return ORBX_OK;
}

extern "C" int orbx_collect_view(orbx_handle *h, const orbx_keypoint **kps, const uint8_t **desc, const int **n_out, int *cap_per_frame)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!h->pending) return fail(h, ORBX_ERR_STATE, "orbx_collect without a pending submit");
    if (h->opt_compact && kps) return fail(h, ORBX_ERR_STATE, "orbx_collect_view: the handle delivers compact keypoints (use orbx_collect_view_compact or orbx_collect)");
    ON_DEVICE(h);
    CU(cudaStreamSynchronize(h->stream));
    h->pending = false;
    if (kps) *kps = h->p_kps;
    if (desc) *desc = h->p_desc;
    if (n_out) *n_out = h->p_n;
    if (cap_per_frame) *cap_per_frame = h->geo.kp_frame_cap;
    return ORBX_OK;
}

extern "C" int orbx_collect(orbx_handle *h, orbx_keypoint *kps, uint8_t *desc, int cap_per_frame, int *n_out)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!kps || !desc || !n_out) return fail(h, ORBX_ERR_BAD_ARG, "NULL output pointer");
    const int rc = orbx_collect_view(h, nullptr, nullptr, nullptr, nullptr);
    if (rc != ORBX_OK) return rc;
    const size_t cap = (size_t)h->geo.kp_frame_cap;
    int status = ORBX_OK;
    for (int f = 0; f < h->last_nframes; ++f) {
        const int n = h->p_n[f];
        n_out[f] = n;
        if (n > cap_per_frame) { status = fail(h, ORBX_ERR_CAPACITY, "cap_per_frame smaller than the keypoint count (see orbx_max_keypoints)"); continue; }
        if (h->opt_compact) expand_keypoints(h, h->p_ckps + f * cap, n, kps + (size_t)f * cap_per_frame);
        else std::memcpy(kps + (size_t)f * cap_per_frame, h->p_kps + f * cap, (size_t)n * sizeof(orbx_keypoint));
        std::memcpy(desc + (size_t)f * cap_per_frame * 32, h->p_desc + f * cap * 32, (size_t)n * 32);
    }
    return status;
}

extern "C" int orbx_extract_batch(orbx_handle *h, const uint8_t *const *frames, int nframes, int width, int height,
                                  int stride_bytes, orbx_keypoint *kps, uint8_t *desc, int cap_per_frame, int *n_out)
{
    const int rc = orbx_submit_host(h, frames, nframes, width, height, stride_bytes);
    if (rc != ORBX_OK) return rc;
    return orbx_collect(h, kps, desc, cap_per_frame, n_out);
}

extern "C" int orbx_extract(orbx_handle *h, const uint8_t *gray, int width, int height, int stride_bytes,
                            orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!n_out) return fail(h, ORBX_ERR_BAD_ARG, "NULL n_out");
    if (!gray || width <= 0 || height <= 0) { *n_out = 0; return ORBX_OK; }     // src/ORBextractor.cc:1049
    const uint8_t *frames[1] = {gray};
    return orbx_extract_batch(h, frames, 1, width, height, stride_bytes, kps, desc, cap, n_out);
}

// ----------------------------------------------------------- stage read-back

static int stage_ready(orbx_handle *h, int frame, int level)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!h->have_batch || !h->geo_valid) return fail(h, ORBX_ERR_STATE, "no batch has been processed yet");
    if (frame < 0 || frame >= h->last_nframes || level < 0 || level >= h->geo.nlevels) return fail(h, ORBX_ERR_BAD_ARG, "frame/level out of range");
    return ORBX_OK;
}

extern "C" int orbx_get_level_size(const orbx_handle *h, int level, int *width, int *height)
{
    if (!h || !h->geo_valid || level < 0 || level >= h->geo.nlevels) return ORBX_ERR_BAD_ARG;
    if (width) *width = h->geo.lv[level].w;
    if (height) *height = h->geo.lv[level].h;
    return ORBX_OK;
}

extern "C" int orbx_get_pyramid_level(orbx_handle *h, int frame, int level, uint8_t *dst, int dst_stride, int with_border)
{
    int rc = stage_ready(h, frame, level);
    if (rc != ORBX_OK) return rc;
    if (!dst) return fail(h, ORBX_ERR_BAD_ARG, "NULL dst");
    ON_DEVICE(h);
    const LevelGeom &L = h->geo.lv[level];
    const uint8_t *src; int pitch;
    if (level == 0) { src = h->last_src0.ptr + (size_t)frame * h->last_src0.frame_stride; pitch = h->last_src0.pitch; }
    else { src = h->d_pyr + (size_t)frame * h->geo.pyr_frame_bytes + L.img_off; pitch = L.pitch; }
    if (!with_border) {
        if (dst_stride < L.w) return fail(h, ORBX_ERR_BAD_ARG, "dst_stride too small");
        CU(cudaMemcpy2DAsync(dst, dst_stride, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
    } else {
        const int W = L.w + 2 * kEdge, H = L.h + 2 * kEdge;
        if (dst_stride < W) return fail(h, ORBX_ERR_BAD_ARG, "dst_stride too small");
        const size_t need = (size_t)W * H;
        if (need > h->pad_bytes) { dfree(h->d_pad); CU(cudaMalloc(&h->d_pad, need)); h->pad_bytes = need; }
        CU(launch_pad_level(src, L.w, L.h, pitch, h->d_pad, W, h->stream, &h->stats));
        CU(cudaMemcpy2DAsync(dst, dst_stride, h->d_pad, W, W, H, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// All levels of one frame with a single synchronisation (mvImagePyramid for the stereo path: the adapter used to pay a pad kernel,
// a copy and a synchronize per level).
extern "C" int orbx_get_pyramid_levels(orbx_handle *h, int frame, int nlevels, uint8_t *const *dst, const int *dst_stride, int with_border)
{
    int rc = stage_ready(h, frame, 0);
    if (rc != ORBX_OK) return rc;
    if (!dst || !dst_stride || nlevels < 1 || nlevels > h->geo.nlevels) return fail(h, ORBX_ERR_BAD_ARG, "orbx_get_pyramid_levels: bad argument");
    ON_DEVICE(h);
    const int b = with_border ? kEdge : 0;
    size_t need = 0;
    for (int l = 0; l < nlevels; ++l) {
        const LevelGeom &L = h->geo.lv[l];
        if (!dst[l] || dst_stride[l] < L.w + 2 * b) return fail(h, ORBX_ERR_BAD_ARG, "orbx_get_pyramid_levels: NULL level pointer or stride too small");
        need += (size_t)(L.w + 2 * b) * (L.h + 2 * b);
    }
    if (with_border && need > h->pad_bytes) { CU(cudaStreamSynchronize(h->stream)); dfree(h->d_pad); CU(cudaMalloc(&h->d_pad, need)); h->pad_bytes = need; }
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        const LevelGeom &L = h->geo.lv[l];
        const uint8_t *src; int pitch;
        if (l == 0) { src = h->last_src0.ptr + (size_t)frame * h->last_src0.frame_stride; pitch = h->last_src0.pitch; }
        else { src = h->d_pyr + (size_t)frame * h->geo.pyr_frame_bytes + L.img_off; pitch = L.pitch; }
        const int W = L.w + 2 * b, H = L.h + 2 * b;
        if (with_border) {
            CU(launch_pad_level(src, L.w, L.h, pitch, h->d_pad + off, W, h->stream, &h->stats));
            CU(cudaMemcpy2DAsync(dst[l], dst_stride[l], h->d_pad + off, W, W, H, cudaMemcpyDeviceToHost, h->stream));
        } else
            CU(cudaMemcpy2DAsync(dst[l], dst_stride[l], src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
        off += (size_t)W * H;
    }
    CU(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_get_blurred_level(orbx_handle *h, int frame, int level, uint8_t *dst, int dst_stride)
{
    int rc = stage_ready(h, frame, level);
    if (rc != ORBX_OK) return rc;
    const LevelGeom &L = h->geo.lv[level];
    if (!dst || dst_stride < L.w) return fail(h, ORBX_ERR_BAD_ARG, "bad dst / dst_stride");
    ON_DEVICE(h);
    CU(cudaMemcpy2DAsync(dst, dst_stride, h->d_blur + (size_t)frame * h->geo.pyr_frame_bytes + L.img_off, L.pitch, L.w, L.h,
                         cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_get_candidates(orbx_handle *h, int frame, int level, int32_t *xys, int cap, int *n)
{
    int rc = stage_ready(h, frame, level);
    if (rc != ORBX_OK) return rc;
    if (!n) return fail(h, ORBX_ERR_BAD_ARG, "NULL n");
    ON_DEVICE(h);
    const LevelGeom &L = h->geo.lv[level];
    uint32_t cnt = 0;
    CU(cudaMemcpyAsync(&cnt, h->d_cand_count + (size_t)frame * h->geo.nlevels + level, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (cnt > (uint32_t)L.cand_cap) cnt = (uint32_t)L.cand_cap;
    *n = (int)cnt;
    if (!xys || cap <= 0 || cnt == 0) return ORBX_OK;
    std::vector<uint32_t> raw(cnt);
    CU(cudaMemcpyAsync(raw.data(), h->d_cand + (size_t)frame * h->geo.cand_frame_elems + L.cand_off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // reference emission order: cell row, cell column, then y, x inside the cell (:789-829)
    auto key = [&](uint32_t c) {
        const unsigned long long x = c & 0xfff, y = (c >> 12) & 0xfff;
        return ((unsigned long long)((y - 3) / L.h_cell * L.n_cols + (x - 3) / L.w_cell) << 24) | y << 12 | x;
    };
    std::sort(raw.begin(), raw.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
    for (int i = 0; i < (int)cnt && i < cap; ++i) {
        xys[3 * i] = (int32_t)(raw[i] & 0xfff); xys[3 * i + 1] = (int32_t)((raw[i] >> 12) & 0xfff); xys[3 * i + 2] = (int32_t)(raw[i] >> 24);
    }
    return ORBX_OK;
}

// ------------------------------------------------------------------ matcher

extern "C" int orbx_hamming256(const void *a, const void *b)
{
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32); std::memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

static int ensure_partial(orbx_handle *h, int nA, int nchunks)
{
    const size_t need = (size_t)nA * nchunks;
    if (need > h->m_cap_partial) { dfree(h->m_partial); CU(cudaMalloc(&h->m_partial, need * sizeof(int4))); h->m_cap_partial = need; }
    if (!h->m_nacc) CU(cudaMalloc(&h->m_nacc, sizeof(int)));
    const size_t blocks = ((size_t)nA + kMatchRowsPerBlock - 1) / kMatchRowsPerBlock;
    if (blocks > h->m_cap_arrive) {
        CU(cudaStreamSynchronize(h->stream));
        dfree(h->m_arrive);
        CU(cudaMalloc(&h->m_arrive, blocks * sizeof(unsigned)));
        CU(cudaMemsetAsync(h->m_arrive, 0, blocks * sizeof(unsigned), h->stream));
        h->m_cap_arrive = blocks;
    }
    return ORBX_OK;
}

extern "C" int orbx_match_device(orbx_handle *h, const uint8_t *dA, int nA, const uint8_t *dB, int nB, int th, float ratio,
                                 int32_t *d_idx, int32_t *d_d1, int32_t *d_d2, uint8_t *d_accept)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (nA < 0 || nB < 0 || (nA > 0 && (!dA || !d_idx || !d_d1 || !d_d2)) || (nB > 0 && !dB)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_match_device: bad argument");
    if (((uintptr_t)dA | (uintptr_t)dB) % 16) return fail(h, ORBX_ERR_BAD_ARG, "descriptor arrays must be 16-byte aligned");
    if (nA == 0) return ORBX_OK;
    ON_DEVICE(h);
    const int nchunks = match_chunks(nA, nB);
    int rc = ensure_partial(h, nA, nchunks);
    if (rc != ORBX_OK) return rc;
    CU(launch_match((const uint32_t *)dA, nA, (const uint32_t *)dB, nB, th, ratio, d_idx, d_d1, d_d2, d_accept, nullptr,
                    h->m_partial, h->m_arrive, nchunks, h->stream, &h->stats));
    return ORBX_OK;
}

extern "C" int orbx_match(orbx_handle *h, const uint8_t *descA, int nA, const uint8_t *descB, int nB, int th, float ratio,
                          int32_t *idx, int32_t *d1, int32_t *d2, uint8_t *accept)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (nA < 0 || nB < 0 || (nA > 0 && (!descA || !idx || !d1 || !d2)) || (nB > 0 && !descB)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_match: bad argument");
    if (nA == 0) return 0;
    ON_DEVICE(h);
    if ((size_t)nA > h->m_capA) { dfree(h->m_A); CU(cudaMalloc(&h->m_A, (size_t)nA * 32)); h->m_capA = nA; }
    if ((size_t)std::max(nB, 1) > h->m_capB) { dfree(h->m_B); CU(cudaMalloc(&h->m_B, (size_t)std::max(nB, 1) * 32)); h->m_capB = std::max(nB, 1); }
    if ((size_t)nA > h->m_cap_out) {
        dfree(h->m_out); dfree(h->m_acc);
        CU(cudaMalloc(&h->m_out, (size_t)nA * 3 * sizeof(int32_t)));
        CU(cudaMalloc(&h->m_acc, (size_t)nA));
        h->m_cap_out = nA;
    }
    const int nchunks = match_chunks(nA, nB);
    int rc = ensure_partial(h, nA, nchunks);
    if (rc != ORBX_OK) return rc;
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(h->m_A, descA, (size_t)nA * 32, cudaMemcpyHostToDevice, st));
    if (nB > 0) CU(cudaMemcpyAsync(h->m_B, descB, (size_t)nB * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->m_nacc, 0, sizeof(int), st));
    CU(launch_match(h->m_A, nA, h->m_B, nB, th, ratio, h->m_out, h->m_out + nA, h->m_out + 2 * (size_t)nA, h->m_acc, h->m_nacc,
                    h->m_partial, h->m_arrive, nchunks, st, &h->stats));
    int nacc = 0;
    CU(cudaMemcpyAsync(idx, h->m_out, (size_t)nA * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d1, h->m_out + nA, (size_t)nA * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d2, h->m_out + 2 * (size_t)nA, (size_t)nA * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (accept) CU(cudaMemcpyAsync(accept, h->m_acc, (size_t)nA, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&nacc, h->m_nacc, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return nacc;
}

extern "C" int orbx_rotation_filter_device(orbx_handle *h, int nA, const int32_t *d_idx, uint8_t *d_accept, const float *d_angleA,
                                           const float *d_angleB, int32_t *d_hist, int32_t *d_top3, int32_t *d_kept)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (nA < 0 || (nA > 0 && (!d_idx || !d_accept || !d_angleA || !d_angleB))) return fail(h, ORBX_ERR_BAD_ARG, "orbx_rotation_filter_device: bad argument");
    ON_DEVICE(h);
    CU(launch_rotation_filter(nA, d_idx, d_accept, d_angleA, d_angleB, d_hist, d_top3, d_kept, h->stream, &h->stats));
    return ORBX_OK;
}

extern "C" int orbx_rotation_filter(orbx_handle *h, int nA, const int32_t *idx, uint8_t *accept, const float *angleA,
                                    const float *angleB, int nB, int32_t *hist, int32_t *top3)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (nA < 0 || nB < 0 || (nA > 0 && (!idx || !accept || !angleA)) || (nB > 0 && !angleB)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_rotation_filter: bad argument");
    for (int i = 0; i < nA; ++i)
        if (accept[i] && (idx[i] < 0 || idx[i] >= nB)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_rotation_filter: accepted match with idx outside [0, nB)");
    ON_DEVICE(h);
    cudaStream_t st = h->stream;
    // one scratch block: idx | angleA | angleB | hist[30] top3[3] kept[1] | accept
    const size_t words = (size_t)nA * 2 + (size_t)nB + 34;
    uint32_t *d = nullptr;
    { const int rcs = get_scratch(h, words * 4 + (size_t)nA + 16, (void **)&d); if (rcs != ORBX_OK) return rcs; }
    int32_t *d_idx = (int32_t *)d; float *d_a = (float *)(d + nA), *d_b = (float *)(d + 2 * (size_t)nA);
    int32_t *d_hist = (int32_t *)(d + 2 * (size_t)nA + nB); uint8_t *d_acc = (uint8_t *)(d + words);
    int32_t res[34];
    auto body = [&]() -> int {
        if (nA > 0) {
            CU(cudaMemcpyAsync(d_idx, idx, (size_t)nA * 4, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_a, angleA, (size_t)nA * 4, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_acc, accept, (size_t)nA, cudaMemcpyHostToDevice, st));
        }
        if (nB > 0) CU(cudaMemcpyAsync(d_b, angleB, (size_t)nB * 4, cudaMemcpyHostToDevice, st));
        CU(launch_rotation_filter(nA, d_idx, d_acc, d_a, d_b, d_hist, d_hist + 30, d_hist + 33, st, &h->stats));
        if (nA > 0) CU(cudaMemcpyAsync(accept, d_acc, (size_t)nA, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(res, d_hist, sizeof(res), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return ORBX_OK;
    };
    const int rc = body();
    if (rc != ORBX_OK) return rc;
    if (hist) std::memcpy(hist, res, 30 * sizeof(int32_t));
    if (top3) std::memcpy(top3, res + 30, 3 * sizeof(int32_t));
    return res[33];
}


// ------------------------------------------------- windowed candidate search
// The three windowed matchers share this: `enqueue(d, cap, stage, cand, cnt, off, tot)` copies the call's inputs into the first
// `in_bytes` of the scratch block d and launches its candidate kernel for n queries.  First run: the kernels stage a window's
// candidates in shared memory (kProjCap per window).  The reference has no such limit, so when a window holds more (wide th * 2
// retries, windowSize 100 on dense frames) the call runs once more with a global staging area of max(count) entries per query.
template <typename Enqueue>
static int run_window_search(orbx_handle *h, int n, size_t in_bytes, Enqueue enqueue, std::vector<int> *count, std::vector<int> *offset,
                             std::vector<unsigned long long> *cand)
{
    cudaStream_t st = h->stream;
    count->assign((size_t)n, 0); offset->assign((size_t)n, 0);
    int cap = kProjCap;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t in_al = (in_bytes + 255) / 256 * 256, meta = ((size_t)n * 8 + 16 + 255) / 256 * 256;
        const size_t cand_bytes = (size_t)n * cap * 8, stage_bytes = attempt ? cand_bytes : 0;
        uint8_t *d = nullptr;
        { const int rcs = get_scratch(h, in_al + meta + cand_bytes + stage_bytes, (void **)&d); if (rcs != ORBX_OK) return rcs; }
        int *d_cnt = (int *)(d + in_al), *d_off = d_cnt + n, *d_tot = d_off + n;
        unsigned long long *d_cand = (unsigned long long *)(d + in_al + meta), *d_stage = attempt ? d_cand + (size_t)n * cap : nullptr;
        int ncand = 0;
        CU(cudaMemsetAsync(d_tot, 0, 4, st));
        { const int rcl = enqueue(d, cap, d_stage, d_cand, d_cnt, d_off, d_tot); if (rcl != ORBX_OK) return rcl; }
        CU(cudaMemcpyAsync(count->data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(offset->data(), d_off, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&ncand, d_tot, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        int most = 0;
        for (int i = 0; i < n; ++i) most = std::max(most, (*count)[i]);
        if (most > cap) {
            if (attempt) return fail(h, ORBX_ERR_CUDA, "windowed search: candidate count changed between two runs");
            cap = (most + 63) / 64 * 64;
            continue;
        }
        cand->assign((size_t)std::max(ncand, 1), 0ull);
        if (ncand > 0) {
            CU(cudaMemcpyAsync(cand->data(), d_cand, (size_t)ncand * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        return ORBX_OK;
    }
    return ORBX_OK;
}

// ------------------------------------------------------- projection matcher

extern "C" int orbx_search_by_projection(orbx_handle *h, const orbx_projection_setup *cam,
                                         int n_last, const float *world_pos, const uint8_t *mp_desc, const uint8_t *valid, const int32_t *nobs,
                                         const int32_t *last_octave, const float *last_angle,
                                         int n_cur, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                                         const uint8_t *cur_desc, float th, int mono, int check_orientation, int32_t *cur_match)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!cam || n_last < 0 || n_cur < 0 || n_cur > 65535 ||
        (n_last > 0 && (!world_pos || !mp_desc || !valid || !nobs || !last_octave || !last_angle)) ||
        (n_cur > 0 && (!cur_xy || !cur_octave || !cur_angle || !cur_uright || !cur_desc || !cur_match)))
        return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_by_projection: bad argument (at most 65535 current features)");
    for (int i = 0; i < n_last; ++i)
        if (valid[i] && (last_octave[i] < 0 || last_octave[i] >= h->tab.nlevels)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_by_projection: octave out of range");
    for (int j = 0; j < n_cur; ++j) cur_match[j] = -1;
    if (n_last == 0 || n_cur == 0) return 0;
    ON_DEVICE(h);
    ProjSetup S;
    S.fx = cam->fx; S.fy = cam->fy; S.cx = cam->cx; S.cy = cam->cy; S.bf = cam->bf; S.b = cam->b;
    S.min_x = cam->min_x; S.max_x = cam->max_x; S.min_y = cam->min_y; S.max_y = cam->max_y;
    S.w_inv = 64.0f / (cam->max_x - cam->min_x); S.h_inv = 48.0f / (cam->max_y - cam->min_y);      // src/Frame.cc:126-127
    std::memcpy(S.Tc, cam->Tcw_cur, sizeof(S.Tc));
    for (int l = 0; l < kMaxLevels; ++l) S.scale[l] = l < h->tab.nlevels ? h->tab.scale[l] : 1.f;
    S.th = th;
    {   // bForward / bBackward, :1968-1979: twc = -Rcw.t()*tcw (gemm with GEMM_1_T: double accumulation), tlc = Rlw*twc + tlw (float)
        const float *Tc = cam->Tcw_cur, *Tl = cam->Tcw_last;
        float twc[3], tlc2;
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)Tc[4 * k + r] * (double)Tc[4 * k + 3];
            twc[r] = (float)(s * -1.0);
        }
        volatile float s = 0.f;                                            // keep the float accumulation un-contracted
        for (int k = 0; k < 3; ++k) { volatile float p = Tl[4 * 2 + k] * twc[k]; s = s + p; }
        tlc2 = s + Tl[4 * 2 + 3];
        S.forward = tlc2 > cam->b && !mono; S.backward = -tlc2 > cam->b && !mono;
    }
    // scratch inputs: world | last_octave | cur_xy | cur_octave | cur_uright | mp_desc | cur_desc | valid
    auto up = [&](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_world = 0, o_loct = o_world + up((size_t)n_last * 12), o_xy = o_loct + up((size_t)n_last * 4), o_coct = o_xy + up((size_t)n_cur * 8),
                 o_ur = o_coct + up((size_t)n_cur * 4), o_mpd = o_ur + up((size_t)n_cur * 4), o_cd = o_mpd + up((size_t)n_last * 32),
                 o_val = o_cd + up((size_t)n_cur * 32), in_bytes = o_val + up((size_t)n_last);
    cudaStream_t st = h->stream;
    std::vector<int> count, offset;
    std::vector<unsigned long long> cand;
    auto enqueue = [&](uint8_t *d, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_cnt, int *d_off, int *d_tot) -> int {
        CU(cudaMemcpyAsync(d + o_world, world_pos, (size_t)n_last * 12, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_loct, last_octave, (size_t)n_last * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_xy, cur_xy, (size_t)n_cur * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_coct, cur_octave, (size_t)n_cur * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_ur, cur_uright, (size_t)n_cur * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_mpd, mp_desc, (size_t)n_last * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_cd, cur_desc, (size_t)n_cur * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_val, valid, (size_t)n_last, cudaMemcpyHostToDevice, st));
        CU(launch_project_candidates(S, n_last, (const float *)(d + o_world), d + o_mpd, d + o_val, (const int32_t *)(d + o_loct), n_cur,
                                     (const float *)(d + o_xy), (const int32_t *)(d + o_coct), (const float *)(d + o_ur), d + o_cd,
                                     cap, d_stage, d_cand, d_cnt, d_off, d_tot, st, &h->stats));
        return ORBX_OK;
    };
    { const int rcw = run_window_search(h, n_last, in_bytes, enqueue, &count, &offset, &cand); if (rcw != ORBX_OK) return rcw; }
    return resolve_projection_matches(n_last, n_cur, cand.data(), count.data(), offset.data(), nobs, last_angle, cur_angle, check_orientation, cur_match);
}

extern "C" int orbx_search_for_initialization(orbx_handle *h, float min_x, float max_x, float min_y, float max_y,
                                              int n1, const int32_t *octave1, const float *angle1, const uint8_t *desc1,
                                              int n2, const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2,
                                              float *prev_matched, int window_size, float nnratio, int check_orientation, int32_t *matches12)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (n1 < 0 || n2 < 0 || n2 > 65535 || !(max_x > min_x) || !(max_y > min_y) ||
        (n1 > 0 && (!octave1 || !angle1 || !desc1 || !prev_matched || !matches12)) || (n2 > 0 && (!xy2 || !octave2 || !angle2 || !desc2)))
        return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_for_initialization: bad argument (at most 65535 features in the second frame)");
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return 0;
    ON_DEVICE(h);
    ProjSetup S{};
    S.min_x = min_x; S.max_x = max_x; S.min_y = min_y; S.max_y = max_y;
    S.w_inv = 64.0f / (max_x - min_x); S.h_inv = 48.0f / (max_y - min_y);                           // src/Frame.cc:126-127
    S.th = (float)window_size;                                                                      // const float &r of GetFeaturesInArea
    // scratch inputs: prev | oct1 | xy2 | oct2 | desc1 | desc2
    auto up = [&](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_prev = 0, o_o1 = o_prev + up((size_t)n1 * 8), o_xy = o_o1 + up((size_t)n1 * 4), o_o2 = o_xy + up((size_t)n2 * 8),
                 o_d1 = o_o2 + up((size_t)n2 * 4), o_d2 = o_d1 + up((size_t)n1 * 32), in_bytes = o_d2 + up((size_t)n2 * 32);
    cudaStream_t st = h->stream;
    std::vector<int> count, offset;
    std::vector<unsigned long long> cand;
    auto enqueue = [&](uint8_t *d, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_cnt, int *d_off, int *d_tot) -> int {
        CU(cudaMemcpyAsync(d + o_prev, prev_matched, (size_t)n1 * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_o1, octave1, (size_t)n1 * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_xy, xy2, (size_t)n2 * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_o2, octave2, (size_t)n2 * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_d1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_d2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, st));
        CU(launch_window_candidates(S, n1, (const float *)(d + o_prev), (const int32_t *)(d + o_o1), d + o_d1, n2, (const float *)(d + o_xy),
                                    (const int32_t *)(d + o_o2), d + o_d2, cap, d_stage, d_cand, d_cnt, d_off, d_tot, st, &h->stats));
        return ORBX_OK;
    };
    { const int rcw = run_window_search(h, n1, in_bytes, enqueue, &count, &offset, &cand); if (rcw != ORBX_OK) return rcw; }
    const int n = resolve_initialization_matches(n1, n2, cand.data(), count.data(), offset.data(), angle1, angle2, nnratio, check_orientation, matches12);
    for (int i = 0; i < n1; ++i)                                                                    // "Update prev matched", :888-891
        if (matches12[i] >= 0) { prev_matched[2 * i] = xy2[2 * matches12[i]]; prev_matched[2 * i + 1] = xy2[2 * matches12[i] + 1]; }
    return n;
}

extern "C" int orbx_search_local_points(orbx_handle *h, float min_x, float max_x, float min_y, float max_y,
                                        int n_mp, const float *proj, const float *view_cos, const int32_t *level, const uint8_t *mp_desc,
                                        const uint8_t *valid, const int32_t *nobs,
                                        int n_feat, const float *feat_xy, const int32_t *feat_octave, const float *feat_uright,
                                        const uint8_t *feat_desc, const int32_t *feat_obs, float th, float nnratio, int32_t *feat_match)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (n_mp < 0 || n_feat < 0 || n_feat > 65535 || !(max_x > min_x) || !(max_y > min_y) ||
        (n_mp > 0 && (!proj || !view_cos || !level || !mp_desc || !valid || !nobs)) ||
        (n_feat > 0 && (!feat_xy || !feat_octave || !feat_uright || !feat_desc || !feat_obs || !feat_match)))
        return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_local_points: bad argument (at most 65535 features)");
    for (int i = 0; i < n_mp; ++i)
        if (valid[i] && (level[i] < 0 || level[i] >= h->tab.nlevels)) return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_local_points: predicted level out of range");
    for (int j = 0; j < n_feat; ++j) feat_match[j] = -1;
    if (n_mp == 0 || n_feat == 0) return 0;
    ON_DEVICE(h);
    ProjSetup S{};
    S.min_x = min_x; S.max_x = max_x; S.min_y = min_y; S.max_y = max_y;
    S.w_inv = 64.0f / (max_x - min_x); S.h_inv = 48.0f / (max_y - min_y);                           // src/Frame.cc:126-127
    for (int l = 0; l < kMaxLevels; ++l) S.scale[l] = l < h->tab.nlevels ? h->tab.scale[l] : 1.f;
    S.th = th; S.forward = th != 1.0;                                                               // bFactor, :422
    // scratch inputs: proj | view_cos | level | xy | octave | uright | mp_desc | desc | valid
    auto up = [&](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_proj = 0, o_vc = o_proj + up((size_t)n_mp * 12), o_lvl = o_vc + up((size_t)n_mp * 4), o_xy = o_lvl + up((size_t)n_mp * 4),
                 o_oct = o_xy + up((size_t)n_feat * 8), o_ur = o_oct + up((size_t)n_feat * 4), o_mpd = o_ur + up((size_t)n_feat * 4),
                 o_fd = o_mpd + up((size_t)n_mp * 32), o_val = o_fd + up((size_t)n_feat * 32), in_bytes = o_val + up((size_t)n_mp);
    cudaStream_t st = h->stream;
    std::vector<int> count, offset;
    std::vector<unsigned long long> cand;
    auto enqueue = [&](uint8_t *d, int cap, unsigned long long *d_stage, unsigned long long *d_cand, int *d_cnt, int *d_off, int *d_tot) -> int {
        CU(cudaMemcpyAsync(d + o_proj, proj, (size_t)n_mp * 12, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_vc, view_cos, (size_t)n_mp * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_lvl, level, (size_t)n_mp * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_xy, feat_xy, (size_t)n_feat * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_oct, feat_octave, (size_t)n_feat * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_ur, feat_uright, (size_t)n_feat * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_mpd, mp_desc, (size_t)n_mp * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_fd, feat_desc, (size_t)n_feat * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + o_val, valid, (size_t)n_mp, cudaMemcpyHostToDevice, st));
        CU(launch_local_candidates(S, n_mp, (const float *)(d + o_proj), (const float *)(d + o_vc), (const int32_t *)(d + o_lvl), d + o_mpd, d + o_val,
                                   n_feat, (const float *)(d + o_xy), (const int32_t *)(d + o_oct), (const float *)(d + o_ur), d + o_fd,
                                   cap, d_stage, d_cand, d_cnt, d_off, d_tot, st, &h->stats));
        return ORBX_OK;
    };
    { const int rcw = run_window_search(h, n_mp, in_bytes, enqueue, &count, &offset, &cand); if (rcw != ORBX_OK) return rcw; }
    return resolve_local_matches(n_mp, n_feat, cand.data(), count.data(), offset.data(), nobs, feat_octave, feat_obs, nnratio, feat_match);
}

// Shared by the two SearchByBoW entry points: the merge walk over the two flattened feature vectors (src/ORBmatcher.cc:552-619,
// lower_bound on sorted maps == advancing the smaller side) and the GPU pair distances of every side-1 feature with valid1 set
// against all side-2 features of the same node.  entries[e] = {side-1 feature, first slot in feats2, count, first distance}.
static int bow_pair_distances(orbx_handle *h, const char *who, int n1, const uint8_t *desc1, const uint8_t *valid1, int nn1, const int32_t *nodes1,
                              const int32_t *off1, const int32_t *feats1, int n2, const uint8_t *desc2, int nn2, const int32_t *nodes2,
                              const int32_t *off2, const int32_t *feats2, std::vector<int> *entries, std::vector<uint16_t> *dist)
{
    entries->clear(); dist->clear();
    if (off1[nn1] > n1 || off2[nn2] > n2) return fail(h, ORBX_ERR_BAD_ARG, std::string(who) + ": feature vector longer than the feature count");
    for (int k = 0; k < off1[nn1]; ++k) if (feats1[k] < 0 || feats1[k] >= n1) return fail(h, ORBX_ERR_BAD_ARG, std::string(who) + ": feature index out of range (side 1)");
    for (int k = 0; k < off2[nn2]; ++k) if (feats2[k] < 0 || feats2[k] >= n2) return fail(h, ORBX_ERR_BAD_ARG, std::string(who) + ": feature index out of range (side 2)");
    long long npairs = 0;
    for (int a = 0, b = 0; a < nn1 && b < nn2;) {
        if (nodes1[a] < nodes2[b]) { ++a; continue; }
        if (nodes1[a] > nodes2[b]) { ++b; continue; }
        const int cnt = off2[b + 1] - off2[b];
        for (int p = off1[a]; p < off1[a + 1] && cnt > 0; ++p) {
            const int i1 = feats1[p];
            if (!valid1[i1]) continue;
            entries->push_back(i1); entries->push_back(off2[b]); entries->push_back(cnt); entries->push_back((int)npairs);
            npairs += cnt;
        }
        ++a; ++b;
    }
    const int ne = (int)(entries->size() / 4);
    if (ne == 0) return ORBX_OK;
    if (npairs > (1ll << 30)) return fail(h, ORBX_ERR_UNSUPPORTED, std::string(who) + ": more than 2^30 descriptor pairs");
    ON_DEVICE(h);
    auto up = [&](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_ent = 0, o_d1 = o_ent + up((size_t)ne * 16), o_d2 = o_d1 + up((size_t)n1 * 32), o_ff = o_d2 + up((size_t)n2 * 32),
                 o_out = o_ff + up((size_t)off2[nn2] * 4), total = o_out + up((size_t)npairs * 2);
    uint8_t *d = nullptr;
    { const int rcs = get_scratch(h, total, (void **)&d); if (rcs != ORBX_OK) return rcs; }
    cudaStream_t st = h->stream;
    dist->resize((size_t)npairs);
    CU(cudaMemcpyAsync(d + o_ent, entries->data(), (size_t)ne * 16, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_d1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_d2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_ff, feats2, (size_t)off2[nn2] * 4, cudaMemcpyHostToDevice, st));
    CU(launch_bow_pair_distances(ne, (const int4 *)(d + o_ent), d + o_d1, d + o_d2, (const int32_t *)(d + o_ff), (uint16_t *)(d + o_out), st, &h->stats));
    CU(cudaMemcpyAsync(dist->data(), d + o_out, (size_t)npairs * 2, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ORBX_OK;
}

extern "C" int orbx_search_by_bow(orbx_handle *h, int n_kf, const float *kf_angle, const uint8_t *kf_desc, const uint8_t *kf_valid,
                                  int kf_nnodes, const int32_t *kf_nodes, const int32_t *kf_off, const int32_t *kf_feats,
                                  int n_f, const float *f_angle, const uint8_t *f_desc,
                                  int f_nnodes, const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_feats,
                                  float nnratio, int check_orientation, int32_t *f_match)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (n_kf < 0 || n_f < 0 || kf_nnodes < 0 || f_nnodes < 0 || (n_kf > 0 && (!kf_angle || !kf_desc || !kf_valid)) ||
        (kf_nnodes > 0 && (!kf_nodes || !kf_off || !kf_feats)) || (n_f > 0 && (!f_angle || !f_desc || !f_match)) ||
        (f_nnodes > 0 && (!f_nodes || !f_off || !f_feats)))
        return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_by_bow: bad argument");
    for (int j = 0; j < n_f; ++j) f_match[j] = -1;
    if (n_kf == 0 || n_f == 0 || kf_nnodes == 0 || f_nnodes == 0) return 0;
    std::vector<int> entries;
    std::vector<uint16_t> dist;
    const int rc = bow_pair_distances(h, "orbx_search_by_bow", n_kf, kf_desc, kf_valid, kf_nnodes, kf_nodes, kf_off, kf_feats, n_f, f_desc, f_nnodes,
                                      f_nodes, f_off, f_feats, &entries, &dist);
    if (rc != ORBX_OK) return rc;
    return resolve_bow_matches((int)(entries.size() / 4), entries.data(), dist.data(), f_feats, n_f, kf_angle, f_angle, nnratio, check_orientation, f_match);
}

extern "C" int orbx_search_by_bow_keyframes(orbx_handle *h, int n1, const float *angle1, const uint8_t *desc1, const uint8_t *valid1,
                                            int nnodes1, const int32_t *nodes1, const int32_t *off1, const int32_t *feats1,
                                            int n2, const float *angle2, const uint8_t *desc2, const uint8_t *valid2,
                                            int nnodes2, const int32_t *nodes2, const int32_t *off2, const int32_t *feats2,
                                            float nnratio, int check_orientation, int32_t *matches12)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (n1 < 0 || n2 < 0 || nnodes1 < 0 || nnodes2 < 0 || (n1 > 0 && (!angle1 || !desc1 || !valid1 || !matches12)) ||
        (nnodes1 > 0 && (!nodes1 || !off1 || !feats1)) || (n2 > 0 && (!angle2 || !desc2 || !valid2)) || (nnodes2 > 0 && (!nodes2 || !off2 || !feats2)))
        return fail(h, ORBX_ERR_BAD_ARG, "orbx_search_by_bow_keyframes: bad argument");
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0 || nnodes1 == 0 || nnodes2 == 0) return 0;
    std::vector<int> entries;
    std::vector<uint16_t> dist;
    const int rc = bow_pair_distances(h, "orbx_search_by_bow_keyframes", n1, desc1, valid1, nnodes1, nodes1, off1, feats1, n2, desc2, nnodes2, nodes2,
                                      off2, feats2, &entries, &dist);
    if (rc != ORBX_OK) return rc;
    return resolve_bow_matches_kf((int)(entries.size() / 4), entries.data(), dist.data(), feats2, n1, n2, valid2, angle1, angle2, nnratio,
                                  check_orientation, matches12);
}

// --------------------------------------------------------------- vocabulary

struct orbx_vocabulary {
    VocHost host;
    int device = 0;
    int32_t *d_child_off = nullptr, *d_child_ids = nullptr, *d_word_id = nullptr;
    uint8_t *d_desc = nullptr;
};

static int voc_upload(orbx_handle *h, orbx_vocabulary *v)
{
    const VocHost &H = v->host;
    v->device = h->device;
    ON_DEVICE(h);
    CU(cudaMalloc(&v->d_child_off, H.child_off.size() * 4));
    CU(cudaMalloc(&v->d_child_ids, std::max<size_t>(H.child_ids.size(), 1) * 4));
    CU(cudaMalloc(&v->d_word_id, H.word_id.size() * 4));
    CU(cudaMalloc(&v->d_desc, H.desc.size()));
    CU(cudaMemcpyAsync(v->d_child_off, H.child_off.data(), H.child_off.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(v->d_child_ids, H.child_ids.data(), H.child_ids.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(v->d_word_id, H.word_id.data(), H.word_id.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(v->d_desc, H.desc.data(), H.desc.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" void orbx_voc_destroy(orbx_vocabulary *v)
{
    if (!v) return;
    DeviceGuard device_guard_(v->device);
    dfree(v->d_child_off); dfree(v->d_child_ids); dfree(v->d_word_id); dfree(v->d_desc);
    delete v;
}

static int voc_finish(orbx_handle *h, orbx_vocabulary *v, int rc, const std::string &err, orbx_vocabulary **out)
{
    if (rc == ORBX_OK) rc = voc_upload(h, v); else fail(h, rc, err);
    if (rc != ORBX_OK) { orbx_voc_destroy(v); return rc; }
    *out = v;
    return ORBX_OK;
}

extern "C" int orbx_voc_load_text(orbx_handle *h, const char *path, orbx_vocabulary **out)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!path || !out) return fail(h, ORBX_ERR_BAD_ARG, "orbx_voc_load_text: NULL argument");
    *out = nullptr;
    orbx_vocabulary *v = new (std::nothrow) orbx_vocabulary();
    if (!v) return fail(h, ORBX_ERR_OOM, "host allocation failed");
    std::string err;
    const int rc = voc_load_text(path, &v->host, &err);
    return voc_finish(h, v, rc, err, out);
}

extern "C" int orbx_voc_create(orbx_handle *h, int k, int L, int scoring, int weighting, int nnodes, const int32_t *parent,
                               const uint8_t *is_leaf, const uint8_t *desc, const double *weight, orbx_vocabulary **out)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!parent || !is_leaf || !desc || !weight || !out) return fail(h, ORBX_ERR_BAD_ARG, "orbx_voc_create: NULL argument");
    *out = nullptr;
    orbx_vocabulary *v = new (std::nothrow) orbx_vocabulary();
    if (!v) return fail(h, ORBX_ERR_OOM, "host allocation failed");
    std::string err;
    const int rc = voc_build(k, L, scoring, weighting, nnodes, parent, is_leaf, desc, weight, &v->host, &err);
    return voc_finish(h, v, rc, err, out);
}

extern "C" int orbx_voc_info(const orbx_vocabulary *v, int *k, int *L, int *nodes, int *words)
{
    if (!v) return ORBX_ERR_BAD_ARG;
    if (k) *k = v->host.k;
    if (L) *L = v->host.L;
    if (nodes) *nodes = v->host.nnodes;
    if (words) *words = v->host.nwords;
    return ORBX_OK;
}

extern "C" int orbx_voc_transform(orbx_handle *h, const orbx_vocabulary *v, const uint8_t *desc, int n, int levelsup,
                                  int32_t *word, int32_t *node, double *weight)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (!v || n < 0 || (n > 0 && (!desc || !word || !node || !weight))) return fail(h, ORBX_ERR_BAD_ARG, "orbx_voc_transform: bad argument");
    if (v->device != h->device) return fail(h, ORBX_ERR_BAD_ARG, "orbx_voc_transform: the vocabulary lives on another device");
    if (n == 0) return ORBX_OK;
    ON_DEVICE(h);
    cudaStream_t st = h->stream;
    uint8_t *d = nullptr;
    const size_t feat_bytes = ((size_t)n * 32 + 255) / 256 * 256;
    { const int rcs = get_scratch(h, feat_bytes + (size_t)n * 12, (void **)&d); if (rcs != ORBX_OK) return rcs; }
    int32_t *d_word = (int32_t *)(d + feat_bytes), *d_node = d_word + n, *d_final = d_node + n;
    std::vector<int32_t> final_id((size_t)n);
    auto body = [&]() -> int {
        CU(cudaMemcpyAsync(d, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
        CU(launch_bow_descent(d, n, v->d_child_off, v->d_child_ids, v->d_desc, v->d_word_id, v->host.L - levelsup, d_word, d_node, d_final, st, &h->stats));
        CU(cudaMemcpyAsync(word, d_word, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(node, d_node, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(final_id.data(), d_final, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return ORBX_OK;
    };
    const int rc = body();
    if (rc != ORBX_OK) return rc;
    for (int i = 0; i < n; ++i) weight[i] = v->host.weight[final_id[i]];          // the word's weight (a double), looked up on the host
    return ORBX_OK;
}

extern "C" int orbx_voc_bow(const orbx_vocabulary *v, int n, const int32_t *word, const int32_t *node, const double *weight,
                            int32_t *bow_ids, double *bow_vals, int *n_bow, int32_t *fv_nodes, int32_t *fv_off, int32_t *fv_feats, int *n_fv)
{
    if (!v || n < 0 || !n_bow || !n_fv || !fv_off || (n > 0 && (!word || !node || !weight || !bow_ids || !bow_vals || !fv_nodes || !fv_feats)))
        return ORBX_ERR_BAD_ARG;
    return voc_bow(v->host, n, word, node, weight, bow_ids, bow_vals, n_bow, fv_nodes, fv_off, fv_feats, n_fv);
}

// ------------------------------------------------- distinctive descriptors

extern "C" int orbx_distinctive_descriptors(orbx_handle *h, const uint8_t *desc, const int32_t *offsets, int npoints,
                                            int32_t *best_idx, int32_t *best_median)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (npoints < 0 || (npoints > 0 && (!offsets || !best_idx))) return fail(h, ORBX_ERR_BAD_ARG, "orbx_distinctive_descriptors: bad argument");
    if (npoints == 0) return ORBX_OK;
    int max_obs = 0;
    for (int p = 0; p < npoints; ++p) {
        const int n = offsets[p + 1] - offsets[p];
        if (n < 0 || offsets[p] < 0) return fail(h, ORBX_ERR_BAD_ARG, "orbx_distinctive_descriptors: offsets must be non-negative and ascending");
        max_obs = std::max(max_obs, n);
    }
    if (max_obs > kDistinctiveMaxObs) return fail(h, ORBX_ERR_UNSUPPORTED, "orbx_distinctive_descriptors: more than 1024 observations of one map point");
    const int total = offsets[npoints];
    if (total > 0 && !desc) return fail(h, ORBX_ERR_BAD_ARG, "orbx_distinctive_descriptors: NULL desc");
    ON_DEVICE(h);
    cudaStream_t st = h->stream;
    uint8_t *d = nullptr;
    const size_t desc_bytes = ((size_t)std::max(total, 1) * 32 + 255) / 256 * 256, off_bytes = ((size_t)(npoints + 1) * 4 + 255) / 256 * 256;
    { const int rcs = get_scratch(h, desc_bytes + off_bytes + (size_t)npoints * 8, (void **)&d); if (rcs != ORBX_OK) return rcs; }
    int32_t *d_off = (int32_t *)(d + desc_bytes), *d_best = (int32_t *)(d + desc_bytes + off_bytes), *d_med = d_best + npoints;
    auto body = [&]() -> int {
        if (total > 0) CU(cudaMemcpyAsync(d, desc, (size_t)total * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_off, offsets, (size_t)(npoints + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(launch_distinctive(d, d_off, npoints, std::max(max_obs, 1), d_best, d_med, st, &h->stats));
        CU(cudaMemcpyAsync(best_idx, d_best, (size_t)npoints * 4, cudaMemcpyDeviceToHost, st));
        if (best_median) CU(cudaMemcpyAsync(best_median, d_med, (size_t)npoints * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return ORBX_OK;
    };
    return body();
}

// ------------------------------------------------------------------- stereo

extern "C" int orbx_stereo_match(orbx_handle *L, orbx_handle *R, int frame_left, int frame_right, float bf,
                                 float *u_right, float *depth, int32_t *desc_index, int cap, int *n_left)
{
    orbx_handle *h = L;
    if (!L || !R) return ORBX_ERR_BAD_ARG;
    int rc = stage_ready(L, frame_left, 0);
    if (rc != ORBX_OK) return rc;
    rc = stage_ready(R, frame_right, 0);
    if (rc != ORBX_OK) return fail(L, rc, "right handle: " + R->err);
    if (L->pending || R->pending) return fail(L, ORBX_ERR_STATE, "orbx_stereo_match: collect the pending batches first");
    if (L->device != R->device) return fail(L, ORBX_ERR_BAD_ARG, "orbx_stereo_match: the two handles live on different devices");
    if (L->geo.width != R->geo.width || L->geo.height != R->geo.height || L->tab.nlevels != R->tab.nlevels ||
        L->tab.scale_factor_f != R->tab.scale_factor_f)
        return fail(L, ORBX_ERR_BAD_ARG, "orbx_stereo_match: the two extractors differ in image size / levels / scale factor");
    ON_DEVICE(L);
    const int capL = L->geo.kp_frame_cap;
    // scratch: u_right | depth | desc_index | sad (capL each) | kept
    uint32_t *d = nullptr;
    { const int rcs = get_scratch(L, ((size_t)capL * 4 + 4) * sizeof(uint32_t), (void **)&d); if (rcs != ORBX_OK) return rcs; }
    float *d_ur = (float *)d, *d_dp = d_ur + capL;
    int32_t *d_di = (int32_t *)(d_dp + capL), *d_sad = d_di + capL;
    int *d_kept = (int *)(d_sad + capL);
    StereoScales sc;
    for (int l = 0; l < kMaxLevels; ++l) { sc.scale[l] = l < L->tab.nlevels ? L->tab.scale[l] : 1.f; sc.inv_scale[l] = l < L->tab.nlevels ? L->tab.inv_scale[l] : 1.f; }
    int nL = 0, kept = 0;
    auto body = [&]() -> int {
        CU(cudaStreamSynchronize(R->stream));                          // the right extractor's results are complete
        cudaStream_t st = L->stream;
        CU(launch_stereo(L->d_params, L->last_src0, frame_left, R->d_params, R->last_src0, frame_right, sc, bf, capL,
                         d_ur, d_dp, d_di, d_sad, d_kept, st, &L->stats));
        CU(cudaMemcpyAsync(&nL, L->d_out_n + frame_left, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&kept, d_kept, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (nL > capL) nL = capL;
        if (n_left) *n_left = nL;
        if (nL > cap && (u_right || depth || desc_index)) return fail(h, ORBX_ERR_CAPACITY, "orbx_stereo_match: cap is smaller than the number of left keypoints");
        if (u_right && nL) CU(cudaMemcpyAsync(u_right, d_ur, (size_t)nL * 4, cudaMemcpyDeviceToHost, st));
        if (depth && nL) CU(cudaMemcpyAsync(depth, d_dp, (size_t)nL * 4, cudaMemcpyDeviceToHost, st));
        if (desc_index && nL) CU(cudaMemcpyAsync(desc_index, d_di, (size_t)nL * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return ORBX_OK;
    };
    rc = body();
    if (rc != ORBX_OK) return rc;
    return kept < 0 ? 0 : kept;
}

// ------------------------------------------------------------------ options

extern "C" int orbx_set_option(orbx_handle *h, int option, int value)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    if (h->pending) return fail(h, ORBX_ERR_STATE, "orbx_set_option: a batch is pending on this handle");
    ON_DEVICE(h);
    switch (option) {
    case ORBX_OPT_TMA_STAGING:
        if (h->opt_tma != (value != 0)) {
            CU(cudaStreamSynchronize(h->stream));
            h->opt_tma = value != 0;
            const int keep = h->batch_cap;
            if (h->geo_valid && keep > 0) { free_batch_buffers(h); h->have_batch = false; return ensure_batch(h, keep); }   // re-encodes / drops the tensor maps
        }
        return ORBX_OK;
    case ORBX_OPT_FAST_TMA: h->opt_fast_tma = value != 0; return ORBX_OK;
    case ORBX_OPT_COPY_INPUT: h->opt_copy_input = value != 0; return ORBX_OK;
    case ORBX_OPT_COMPACT_KEYPOINTS:
        if (h->opt_compact != (value != 0)) {
            CU(cudaStreamSynchronize(h->stream));
            h->opt_compact = value != 0;
            const int keep = h->batch_cap;
            if (h->geo_valid && keep > 0) { free_batch_buffers(h); h->have_batch = false; return ensure_batch(h, keep); }
        }
        return ORBX_OK;
    default: return fail(h, ORBX_ERR_BAD_ARG, "orbx_set_option: unknown option");
    }
}

// ---------------------------------------------------------------- profiling

static const char *kStageNames[ORBX_NUM_STAGES] = {"input", "pyramid", "blur", "fast", "octree", "orient_desc", "d2h"};

extern "C" const char *orbx_stage_name(int stage) { return stage >= 0 && stage < ORBX_NUM_STAGES ? kStageNames[stage] : ""; }

extern "C" int orbx_set_profiling(orbx_handle *h, int on)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    ON_DEVICE(h);
    if (on && !h->ev[0])
        for (int i = 0; i <= ORBX_NUM_STAGES; ++i) CU(cudaEventCreate(&h->ev[i]));
    h->profiling = on != 0;
    h->ev_valid = false;
    return ORBX_OK;
}

extern "C" int orbx_get_stage_ms(orbx_handle *h, float *ms, int cap)
{
    if (!h || !ms) return ORBX_ERR_BAD_ARG;
    if (!h->ev_valid || h->pending) return fail(h, ORBX_ERR_STATE, "no profiled batch has completed (enable profiling, submit, collect)");
    ON_DEVICE(h);
    for (int i = 0; i < ORBX_NUM_STAGES && i < cap; ++i) CU(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
    return ORBX_NUM_STAGES;
}

// --------------------------------------------------------------------- misc

extern "C" int orbx_sync(orbx_handle *h)
{
    if (!h) return ORBX_ERR_BAD_ARG;
    ON_DEVICE(h);
    CU(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" void *orbx_stream(orbx_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" long long orbx_launch_count(const orbx_handle *h) { return h ? h->stats.launches : 0; }
