// multimot_track_b200/csrc/pool.cpp -- the frame-sharded multi-GPU dispatcher of include/orbx.h (orbx_pool_*).
//
// frames are independent inside ORBextractor::operator() (the reference already runs two extractor instances on two
// threads for a stereo pair, src/Frame.cc:96-99), so a batch of F frames is split into G contiguous blocks, one per GPU, with no
// exchange between them.  Here: one worker thread + `depth` extractor handles per device.  A submit only queues work; the worker
// thread issues the copies and kernel launches of its shard on the next handle of its ring, so the caller's thread (Python, or
// Tracking's frame loop) never sits in the CUDA launch path, and `depth` submits overlap on every GPU (H2D of one, kernels of the
// next, D2H of the one before).  Results land in each handle's pinned buffers and are handed out as views per shard.
//
// Built on the single-GPU C ABI only (orbx_create / orbx_submit_* / orbx_collect_view): no CUDA call of its own.
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orbx.h"

namespace {

struct Job {
    long long ticket = -1;
    bool device_resident = false;
    std::vector<const uint8_t *> frames;      // host frames of this shard (device_resident: frames[0] = first device frame)
    int nframes = 0, first_frame = 0, width = 0, height = 0, stride = 0;
    size_t frame_stride = 0;
};

struct Slot {                                  // one handle of a worker's ring and the state of the submit it carries
    orbx_handle *h = nullptr;
    long long ticket = -1;                     // ticket whose shard sits on this handle (-1: free)
    bool submitted = false;                    // the worker has queued it on the handle's stream
    int rc = ORBX_OK, nframes = 0, first_frame = 0;
    std::string err;
};

struct Worker {
    int device = 0, index = 0;
    std::vector<Slot> ring;
    std::deque<Job> queue;
    std::thread thread;
};

} // namespace

struct orbx_pool {
    std::vector<Worker> workers;
    int depth = 6;
    bool compact = false;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    bool stop = false;
    long long next_ticket = 0;                 // next ticket to hand out
    long long released = 0;                    // tickets below this have been collected (their slots are free)
    std::vector<char> collected;               // ring of `depth` flags: ticket t collected (index t % depth)
    std::string err;
};

static thread_local std::string g_pool_create_error;

namespace {

int pfail(orbx_pool *p, int code, const std::string &msg)
{
    if (p) p->err = msg; else g_pool_create_error = msg;
    return code;
}

void worker_main(orbx_pool *p, Worker *w)
{
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_work.wait(lk, [&] { return p->stop || !w->queue.empty(); });
            if (w->queue.empty()) return;      // stop requested and nothing left to issue
            job = std::move(w->queue.front());
            w->queue.pop_front();
        }
        Slot &s = w->ring[(size_t)(job.ticket % p->depth)];
        int rc = ORBX_OK;
        if (job.nframes > 0) {
            if (job.device_resident)
                rc = orbx_submit_device(s.h, job.frames[0], job.nframes, job.width, job.height, job.stride, job.frame_stride);
            else
                rc = orbx_submit_host(s.h, job.frames.data(), job.nframes, job.width, job.height, job.stride);
        }
        {
            std::lock_guard<std::mutex> lk(p->mu);
            s.rc = rc;
            if (rc != ORBX_OK) s.err = orbx_last_error(s.h);
            s.submitted = true;
        }
        p->cv_done.notify_all();
    }
}

// Queues one shard per worker under a fresh ticket; blocks while `depth` tickets are uncollected.
long long enqueue(orbx_pool *p, std::vector<Job> &jobs)
{
    std::unique_lock<std::mutex> lk(p->mu);
    if (p->next_ticket - p->released >= p->depth) {
        p->err = "orbx_pool: `depth` submits are in flight and none has been collected; call orbx_pool_collect first";
        return ORBX_ERR_STATE;
    }
    const long long t = p->next_ticket++;
    p->collected[(size_t)(t % p->depth)] = 0;
    for (size_t g = 0; g < p->workers.size(); ++g) {
        Worker &w = p->workers[g];
        Slot &s = w.ring[(size_t)(t % p->depth)];
        s.ticket = t; s.submitted = false; s.rc = ORBX_OK; s.err.clear();
        s.nframes = jobs[g].nframes; s.first_frame = jobs[g].first_frame;
        jobs[g].ticket = t;
        w.queue.push_back(std::move(jobs[g]));
    }
    lk.unlock();
    p->cv_work.notify_all();
    return t;
}

} // namespace

extern "C" int orbx_pool_create(const orbx_pool_config *cfg, orbx_pool **out)
{
    if (!cfg || !out) return pfail(nullptr, ORBX_ERR_BAD_ARG, "orbx_pool_create: NULL argument");
    *out = nullptr;
    if (cfg->ndevices < 0 || (cfg->ndevices > 0 && !cfg->devices) || cfg->depth < 0 || cfg->depth > 64)
        return pfail(nullptr, ORBX_ERR_BAD_ARG, "orbx_pool_create: ndevices >= 0 with a devices array, depth in 0..64 required");
    orbx_pool *p = new (std::nothrow) orbx_pool();
    if (!p) return pfail(nullptr, ORBX_ERR_OOM, "host allocation failed");
    p->depth = cfg->depth > 0 ? cfg->depth : 6;
    p->compact = cfg->compact_keypoints != 0;
    p->collected.assign((size_t)p->depth, 1);
    const int G = cfg->ndevices > 0 ? cfg->ndevices : 1;
    p->workers.resize((size_t)G);
    for (int g = 0; g < G; ++g) {
        Worker &w = p->workers[(size_t)g];
        w.index = g;
        w.device = cfg->ndevices > 0 ? cfg->devices[g] : -1;
        w.ring.resize((size_t)p->depth);
        for (int k = 0; k < p->depth; ++k) {
            orbx_config c = cfg->extractor;
            c.device_id = w.device;
            int rc = orbx_create(&c, &w.ring[(size_t)k].h);
            if (rc == ORBX_OK && cfg->compact_keypoints) rc = orbx_set_option(w.ring[(size_t)k].h, ORBX_OPT_COMPACT_KEYPOINTS, 1);
            if (rc != ORBX_OK) {
                const std::string msg = std::string("orbx_pool_create: device ") + std::to_string(w.device) + ": " + orbx_last_error(nullptr);
                orbx_pool_destroy(p);
                return pfail(nullptr, rc, msg);
            }
        }
    }
    for (Worker &w : p->workers) w.thread = std::thread(worker_main, p, &w);
    *out = p;
    return ORBX_OK;
}

extern "C" void orbx_pool_destroy(orbx_pool *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_work.notify_all();
    for (Worker &w : p->workers) if (w.thread.joinable()) w.thread.join();
    for (Worker &w : p->workers)
        for (Slot &s : w.ring) if (s.h) orbx_destroy(s.h);       // waits for the handle's stream
    delete p;
}

extern "C" const char *orbx_pool_last_error(const orbx_pool *p) { return p ? p->err.c_str() : g_pool_create_error.c_str(); }

extern "C" int orbx_pool_devices(const orbx_pool *p) { return p ? (int)p->workers.size() : ORBX_ERR_BAD_ARG; }

extern "C" int orbx_pool_depth(const orbx_pool *p) { return p ? p->depth : ORBX_ERR_BAD_ARG; }

extern "C" orbx_handle *orbx_pool_handle(orbx_pool *p, int shard, int slot)
{
    if (!p || shard < 0 || shard >= (int)p->workers.size() || slot < 0 || slot >= p->depth) return nullptr;
    return p->workers[(size_t)shard].ring[(size_t)slot].h;
}

extern "C" int orbx_pool_set_option(orbx_pool *p, int option, int value)
{
    if (!p) return ORBX_ERR_BAD_ARG;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->released != p->next_ticket) return pfail(p, ORBX_ERR_STATE, "orbx_pool_set_option: collect every outstanding ticket first");
    }
    for (Worker &w : p->workers)
        for (Slot &s : w.ring) {
            const int rc = orbx_set_option(s.h, option, value);
            if (rc != ORBX_OK) return pfail(p, rc, orbx_last_error(s.h));
        }
    if (option == ORBX_OPT_COMPACT_KEYPOINTS) p->compact = value != 0;
    return ORBX_OK;
}

extern "C" void orbx_pool_shard_range(int nframes, int nshards, int shard, int *first, int *count)
{
    // contiguous blocks [g*F/G, (g+1)*F/G): consecutive frames (the pairs a frame-to-frame matcher wants) stay on one GPU
    const long long a = (long long)nframes * shard / nshards, b = (long long)nframes * (shard + 1) / nshards;
    if (first) *first = (int)a;
    if (count) *count = (int)(b - a);
}

extern "C" long long orbx_pool_submit_host(orbx_pool *p, const uint8_t *const *frames, int nframes, int width, int height, int stride_bytes)
{
    if (!p) return ORBX_ERR_BAD_ARG;
    if (!frames || nframes <= 0 || width <= 0 || height <= 0 || stride_bytes < width) return pfail(p, ORBX_ERR_BAD_ARG, "orbx_pool_submit_host: bad argument");
    for (int f = 0; f < nframes; ++f) if (!frames[f]) return pfail(p, ORBX_ERR_BAD_ARG, "orbx_pool_submit_host: NULL frame pointer");
    const int G = (int)p->workers.size();
    std::vector<Job> jobs((size_t)G);
    for (int g = 0; g < G; ++g) {
        Job &j = jobs[(size_t)g];
        orbx_pool_shard_range(nframes, G, g, &j.first_frame, &j.nframes);
        j.frames.assign(frames + j.first_frame, frames + j.first_frame + j.nframes);
        j.width = width; j.height = height; j.stride = stride_bytes;
    }
    return enqueue(p, jobs);
}

extern "C" long long orbx_pool_submit_device(orbx_pool *p, const uint8_t *const *d_frames, const int32_t *nframes, int width, int height,
                                             int stride_bytes, size_t frame_stride_bytes)
{
    if (!p) return ORBX_ERR_BAD_ARG;
    if (!d_frames || !nframes || width <= 0 || height <= 0 || stride_bytes < width) return pfail(p, ORBX_ERR_BAD_ARG, "orbx_pool_submit_device: bad argument");
    const int G = (int)p->workers.size();
    std::vector<Job> jobs((size_t)G);
    int first = 0;
    for (int g = 0; g < G; ++g) {
        if (nframes[g] < 0 || (nframes[g] > 0 && !d_frames[g])) return pfail(p, ORBX_ERR_BAD_ARG, "orbx_pool_submit_device: bad shard");
        Job &j = jobs[(size_t)g];
        j.device_resident = true;
        j.frames.assign(1, d_frames[g]);
        j.nframes = nframes[g]; j.first_frame = first; first += nframes[g];
        j.width = width; j.height = height; j.stride = stride_bytes; j.frame_stride = frame_stride_bytes;
    }
    return enqueue(p, jobs);
}

extern "C" int orbx_pool_collect(orbx_pool *p, long long ticket, orbx_shard_result *shards)
{
    if (!p) return ORBX_ERR_BAD_ARG;
    const int G = (int)p->workers.size();
    {
        std::unique_lock<std::mutex> lk(p->mu);
        if (ticket < p->released || ticket >= p->next_ticket || p->collected[(size_t)(ticket % p->depth)])
            return pfail(p, ORBX_ERR_STATE, "orbx_pool_collect: unknown ticket, or collected before");
        // the workers have issued every shard of this ticket
        p->cv_done.wait(lk, [&] {
            for (const Worker &w : p->workers) if (!w.ring[(size_t)(ticket % p->depth)].submitted) return false;
            return true;
        });
    }
    int status = ORBX_OK;
    for (int g = 0; g < G; ++g) {
        Slot &s = p->workers[(size_t)g].ring[(size_t)(ticket % p->depth)];
        orbx_shard_result r;
        std::memset(&r, 0, sizeof r);
        r.device = p->workers[(size_t)g].device; r.nframes = s.nframes; r.first_frame = s.first_frame;
        int rc = s.rc;
        if (rc != ORBX_OK) pfail(p, rc, "shard " + std::to_string(g) + ": " + s.err);
        else if (s.nframes > 0) {
            const int *n = nullptr;
            int cap = 0;
            rc = p->compact ? orbx_collect_view_compact(s.h, &r.ckps, &r.desc, &n, &cap)      // waits for the shard's stream; views into pinned memory
                            : orbx_collect_view(s.h, &r.kps, &r.desc, &n, &cap);
            r.n = n; r.cap_per_frame = cap;
            if (rc != ORBX_OK) pfail(p, rc, "shard " + std::to_string(g) + ": " + orbx_last_error(s.h));
        }
        if (rc != ORBX_OK && status == ORBX_OK) status = rc;
        if (shards) shards[g] = r;
    }
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->collected[(size_t)(ticket % p->depth)] = 1;
        while (p->released < p->next_ticket && p->collected[(size_t)(p->released % p->depth)]) ++p->released;
    }
    return status;
}
