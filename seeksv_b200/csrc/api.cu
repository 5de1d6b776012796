// C-ABI glue: contexts, error strings, per-kernel timers, BAM residency.
#include <sys/mman.h>
#include <unistd.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <future>
#include <thread>

#include "../host/bamfile.h"
#include "common.cuh"

static thread_local std::string g_create_err;

int svb_fail(svb_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_err = buf;
    return code;
}

extern "C" int svb_abi_version(void) { return SVB_ABI_VERSION; }

extern "C" int svb_ctx_create(int device, svb_ctx **out)
{
    svb_ctx *ctx = nullptr;  // for CK
    if (!out) return svb_fail(nullptr, SVB_ERR_ARG, "svb_ctx_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return svb_fail(nullptr, SVB_ERR_NO_DEVICE, "no CUDA device (%s); seeksv_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return svb_fail(nullptr, SVB_ERR_ARG, "device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return svb_fail(nullptr, SVB_ERR_NO_DEVICE, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    std::unique_ptr<svb_ctx> c(new svb_ctx());
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->fork_event, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->join_event, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->sw_event, cudaEventDisableTiming));
    CK(cudaHostAlloc((void **)&c->ctl_host, 4096, cudaHostAllocDefault));
    for (int i = 0; i < svb_ctx::N_AUX; ++i) CK(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    {
        // The full-pass kernels read ~40-100 bytes out of every ~300-byte record: with the default L2 fetch granularity a
        // touched 32-byte sector drags its whole 128-byte line out of HBM. 32 B keeps DRAM traffic at what is used.
        const char *e = getenv("SEEKSV_B200_L2_FETCH");
        size_t g = e ? (size_t)atoi(e) : 32;
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
    }
    // keep freed scratch in the pool: the commands allocate and free the same sizes repeatedly
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ull;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = c.release();
    return 0;
}

extern "C" void svb_ctx_destroy(svb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->prof_flush();
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    for (auto &e : ctx->big_free) cudaFree(e.first);
    ctx->big_free.clear();
    for (auto &e : ctx->dev_free) cudaFree(e.first);
    ctx->dev_free.clear();
    for (int w = 0; w < 2; ++w)
        if (ctx->ws[w]) cudaFree(ctx->ws[w]);
    if (ctx->inflate_scratch) cudaFree(ctx->inflate_scratch);
    if (ctx->ctl_host) cudaFreeHost(ctx->ctl_host);
    if (ctx->join_event) cudaEventDestroy(ctx->join_event);
    if (ctx->sw_event) cudaEventDestroy(ctx->sw_event);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (cudaStream_t a : ctx->aux)
        if (a) cudaStreamDestroy(a);
    delete ctx;
}

extern "C" const char *svb_last_error(const svb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" void *svb_ctx_stream(svb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

uint8_t *svb_ctx::big_get(uint64_t bytes, uint64_t *cap)
{
    int best = -1;
    for (size_t i = 0; i < big_free.size(); ++i)
        if (big_free[i].second >= bytes && big_free[i].second <= 2 * bytes + (64ull << 20) &&
            (best < 0 || big_free[i].second < big_free[best].second))
            best = (int)i;
    if (best >= 0) {
        uint8_t *p = big_free[best].first;
        *cap = big_free[best].second;
        big_free.erase(big_free.begin() + best);
        return p;
    }
    uint8_t *p = nullptr;
    uint64_t want = bytes + bytes / 16;  // a little head room: the next file of the same kind usually fits
    if (cudaMallocAsync((void **)&p, want, stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    *cap = want;
    return p;
}

void svb_ctx::big_put(uint8_t *p, uint64_t cap)
{
    if (!p) return;
    big_free.emplace_back(p, cap);
    while (big_free.size() > 2) {  // keep the two largest
        size_t smallest = 0;
        for (size_t i = 1; i < big_free.size(); ++i)
            if (big_free[i].second < big_free[smallest].second) smallest = i;
        cudaFreeAsync(big_free[smallest].first, stream);
        big_free.erase(big_free.begin() + smallest);
    }
}

// the workspace of a sync-free pipeline: grows, never shrinks (a command owns it from its first launch to its read-back)
int svb_ctx::ws_reserve(int which, uint64_t bytes)
{
    svb_ctx *ctx = this;
    if (ws_cap[which] >= bytes) return 0;
    CK(cudaStreamSynchronize(stream));
    for (cudaStream_t a : aux) CK(cudaStreamSynchronize(a));
    if (ws[which]) CK(cudaFree(ws[which]));
    ws[which] = nullptr, ws_cap[which] = 0;
    const uint64_t want = bytes + bytes / 8 + (1u << 20);
    CK(cudaMalloc((void **)&ws[which], want));
    ws_cap[which] = want;
    return 0;
}

uint8_t *svb_ctx::dev_get(uint64_t bytes, uint64_t *cap)
{
    int best = -1;
    for (size_t i = 0; i < dev_free.size(); ++i)
        if (dev_free[i].second >= bytes && dev_free[i].second <= 3 * bytes + (32ull << 20) &&
            (best < 0 || dev_free[i].second < dev_free[best].second))
            best = (int)i;
    if (best >= 0) {
        uint8_t *p = dev_free[best].first;
        *cap = dev_free[best].second;
        dev_free.erase(dev_free.begin() + best);
        return p;
    }
    uint8_t *p = nullptr;
    const uint64_t want = bytes + bytes / 8 + 4096;
    if (cudaMallocAsync((void **)&p, want, stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    *cap = want;
    return p;
}
void svb_ctx::dev_put(uint8_t *p, uint64_t cap)
{
    if (!p) return;
    dev_free.emplace_back(p, cap);
    while (dev_free.size() > 16) {  // drop the smallest
        size_t smallest = 0;
        for (size_t i = 1; i < dev_free.size(); ++i)
            if (dev_free[i].second < dev_free[smallest].second) smallest = i;
        cudaFreeAsync(dev_free[smallest].first, stream);
        dev_free.erase(dev_free.begin() + smallest);
    }
}

char *svb_ctx::pinned_get(uint64_t bytes, uint64_t *cap)
{
    if (bytes == 0) {
        *cap = 0;
        return nullptr;
    }
    int best = -1;
    for (size_t i = 0; i < pinned_free.size(); ++i)
        if (pinned_free[i].second >= bytes && (best < 0 || pinned_free[i].second < pinned_free[best].second)) best = (int)i;
    if (best >= 0) {
        char *p = pinned_free[best].first;
        *cap = pinned_free[best].second;
        pinned_free.erase(pinned_free.begin() + best);
        return p;
    }
    uint64_t want = bytes + bytes / 4 + 4096;
    char *p = nullptr;
    if (cudaHostAlloc((void **)&p, want, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    *cap = want;
    return p;
}
void svb_ctx::pinned_put(char *p, uint64_t cap)
{
    if (pinned_free.size() >= 16) {
        cudaFreeHost(p);
        return;
    }
    pinned_free.emplace_back(p, cap);
}
svb_ctx::~svb_ctx()
{
    if (read_buf) pinned_put(read_buf, read_cap), read_buf = nullptr;
    for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
    for (auto &b : pinned_free) cudaFreeHost(b.first);
}

void svb_ctx::prof_flush()
{
    if (prof_pending.empty()) return;
    cudaStreamSynchronize(stream);
    for (cudaStream_t a : aux) cudaStreamSynchronize(a);
    for (auto &p : prof_pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        ProfEntry &e = prof_acc[p.name];
        e.ms += ms, e.bytes += p.bytes, e.launches += 1;
        prof_events.push_back(p.a);
        prof_events.push_back(p.b);
    }
    prof_pending.clear();
}
extern "C" void svb_prof_enable(svb_ctx *ctx, int on)
{
    if (ctx) ctx->prof = on != 0;
}
extern "C" void svb_prof_reset(svb_ctx *ctx)
{
    if (!ctx) return;
    ctx->prof_flush();
    ctx->prof_acc.clear();
}
extern "C" int svb_prof_read(svb_ctx *ctx, int cap, const char **names, double *ms, int64_t *launches, double *bytes)
{
    if (!ctx) return 0;
    ctx->prof_flush();
    int i = 0;
    for (auto &kv : ctx->prof_acc) {
        if (i < cap) {
            if (names) names[i] = kv.first.c_str();
            if (ms) ms[i] = kv.second.ms;
            if (launches) launches[i] = kv.second.launches;
            if (bytes) bytes[i] = kv.second.bytes;
        }
        ++i;
    }
    return i;
}

struct WallScope {  // host wall-clock accounting next to the kernel timers (only when profiling is on)
    svb_ctx *ctx;
    const char *name;
    double bytes;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    WallScope(svb_ctx *c, const char *n, double b = 0) : ctx(c), name(n), bytes(b) {}
    ~WallScope()
    {
        if (!ctx->prof) return;
        ProfEntry &e = ctx->prof_acc[name];
        e.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        e.bytes += bytes, e.launches += 1;
    }
};

// ---- BAM residency ----------------------------------------------------------------------------------------------
static int finish_bam(svb_ctx *ctx, std::unique_ptr<svb_bam> &b, svb_bam **out)
{
    WallScope ws(ctx, "index_records(wall)");
    CKR(index_records(ctx, b.get()));
    *out = b.release();
    return 0;
}

extern "C" int svb_bam_from_device(svb_ctx *ctx, const void *d_stream, uint64_t nbytes, uint64_t first_record, int32_t n_ref,
                                   svb_bam **out)
{
    if (!ctx || !out || (!d_stream && nbytes)) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_from_device: null argument");
    CK(cudaSetDevice(ctx->device));
    std::unique_ptr<svb_bam> b(new svb_bam());
    b->ctx = ctx, b->d_data = (const uint8_t *)d_stream, b->nbytes = nbytes, b->first = first_record, b->n_ref = n_ref;
    return finish_bam(ctx, b, out);
}

static int upload(svb_ctx *ctx, svb_bam *b, const void *h, uint64_t nbytes)
{
    b->d_owned = ctx->big_get(nbytes + 256, &b->owned_cap);
    if (!b->d_owned) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory", (unsigned long long)nbytes);
    CK(cudaMemsetAsync(b->d_owned + nbytes, 0, 256, ctx->stream));
    {
        ProfScope ps(ctx, "h2d_stream", (double)nbytes);
        CK(cudaMemcpyAsync(b->d_owned, h, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    b->d_data = b->d_owned;
    b->nbytes = nbytes;
    return 0;
}

extern "C" int svb_bam_from_host(svb_ctx *ctx, const void *h_stream, uint64_t nbytes, uint64_t first_record, int32_t n_ref,
                                 svb_bam **out)
{
    if (!ctx || !out || (!h_stream && nbytes)) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_from_host: null argument");
    CK(cudaSetDevice(ctx->device));
    std::unique_ptr<svb_bam> b(new svb_bam());
    b->ctx = ctx, b->first = first_record, b->n_ref = n_ref;
    CKR(upload(ctx, b.get(), h_stream, nbytes));
    return finish_bam(ctx, b, out);
}

// The staging copies of one load (page cache -> pinned slab, 32 MiB at a time) on a pool of threads that lives as long as the load:
// spawning 15 threads per slab cost a quarter to a third of the copy itself (23 slabs for the C2 file). A job is cut into 1 MiB
// parts; the caller takes part in the work and returns when the slab is filled. pread first (no page faults on a fresh mapping),
// memcpy from the mapping for any part pread cannot deliver.
class StagePool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t gen = 0;
    bool stop = false;
    uint8_t *dst = nullptr;
    const uint8_t *src = nullptr;
    int fd = -1;
    uint64_t off = 0, n = 0, parts = 0;
    std::atomic<uint64_t> next{0}, done{0};
    std::atomic<int> active{0};  // workers inside work(): a new job is published only when none is left in the previous one
    static constexpr uint64_t PART = 1ull << 20;
    void work()
    {
        for (;;) {
            const uint64_t i = next.fetch_add(1);
            if (i >= parts) return;
            uint64_t a = i * PART;
            const uint64_t b = std::min(n, a + PART);
            bool by_read = fd >= 0;
            while (by_read && a < b) {
                const ssize_t r = pread(fd, dst + a, b - a, (off_t)(off + a));
                if (r <= 0) by_read = false;
                else a += (uint64_t)r;
            }
            if (a < b) memcpy(dst + a, src + a, b - a);
            if (done.fetch_add(1) + 1 == parts) {
                std::lock_guard<std::mutex> lock(m);
                cv_done.notify_all();
            }
        }
    }

public:
    explicit StagePool(int n_threads)
    {
        for (int t = 1; t < n_threads; ++t)
            th.emplace_back([this]() {
                uint64_t seen = 0;
                for (;;) {
                    {
                        std::unique_lock<std::mutex> lock(m);
                        cv_job.wait(lock, [&]() { return stop || gen != seen; });
                        if (stop) return;
                        seen = gen;
                        active.fetch_add(1);
                    }
                    work();
                    active.fetch_sub(1);
                }
            });
    }
    ~StagePool()
    {
        {
            std::lock_guard<std::mutex> lock(m);
            stop = true;
        }
        cv_job.notify_all();
        for (auto &t : th) t.join();
    }
    // fills dst[0, n) from the file (fd at file_off) or, failing that, from src
    void run(uint8_t *dst_, int fd_, uint64_t file_off, const uint8_t *src_, uint64_t n_)
    {
        if (!n_) return;
        while (active.load() != 0) std::this_thread::yield();  // (stragglers of the previous job making their last, failing, grab)
        {
            std::lock_guard<std::mutex> lock(m);
            dst = dst_, src = src_, fd = fd_, off = file_off, n = n_, parts = (n_ + PART - 1) / PART;
            done.store(0), next.store(0);
            ++gen;
        }
        cv_job.notify_all();
        work();
        std::unique_lock<std::mutex> lock(m);
        cv_done.wait(lock, [&]() { return done.load() >= parts; });
    }
};

static bool use_host_inflate()
{
    const char *e = getenv("SEEKSV_B200_HOST_INFLATE");
    return e && *e && *e != '0';
}

namespace {
struct BigBuf {  // a large device buffer that goes back to the context's cache on every way out
    svb_ctx *ctx;
    uint8_t *p = nullptr;
    uint64_t cap = 0;
    ~BigBuf()
    {
        if (!p) return;
        cudaStreamSynchronize(ctx->copy_stream);
        for (cudaStream_t a : ctx->aux) cudaStreamSynchronize(a);
        cudaStreamSynchronize(ctx->stream);
        ctx->big_put(p, cap);
    }
};
struct Slabs {  // two recycled pinned staging buffers, each with the event that says "the device has taken it"
    static constexpr uint64_t SLAB = 32ull << 20;
    svb_ctx *ctx;
    uint8_t *p[2] = {nullptr, nullptr};
    uint64_t cap[2] = {0, 0};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool init()
    {
        for (int i = 0; i < 2; ++i) {
            p[i] = (uint8_t *)ctx->pinned_get(SLAB + (64 << 10), &cap[i]);
            if (!p[i] || cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        return true;
    }
    ~Slabs()
    {
        for (int i = 0; i < 2; ++i) {
            if (done[i]) cudaEventSynchronize(done[i]), cudaEventDestroy(done[i]);
            if (p[i]) ctx->pinned_put((char *)p[i], cap[i]);
        }
    }
};

// north-star variant: host threads inflate the blocks into the pinned slabs (zlib), the uncompressed bytes are streamed
int load_bgzf_host_inflate(svb_ctx *ctx, svb_bam *b, const uint8_t *h_file, uint64_t file_bytes, int n_threads, std::vector<uint8_t> &head)
{
    std::vector<BgzfBlock> blocks;
    uint64_t total = 0;
    std::string err;
    {
        WallScope ws(ctx, "bgzf_scan(wall)");
        if (!bgzf_scan(h_file, file_bytes, blocks, total, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    }
    WallScope ws(ctx, "host_inflate+h2d(wall)", (double)total);
    b->d_owned = ctx->big_get(total + 256, &b->owned_cap);
    if (!b->d_owned) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory", (unsigned long long)total);
    b->d_data = b->d_owned, b->nbytes = total;
    CK(cudaMemsetAsync(b->d_owned + total, 0, 256, ctx->stream));
    Slabs sl{ctx};
    if (!sl.init()) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate pinned staging");
    int slab = 0;
    size_t bi = 0;
    while (bi < blocks.size()) {
        size_t bj = bi;
        uint64_t bytes = 0;
        while (bj < blocks.size() && bytes + blocks[bj].ulen <= Slabs::SLAB + (64 << 10) && bytes < Slabs::SLAB) bytes += blocks[bj++].ulen;
        CK(cudaEventSynchronize(sl.done[slab]));
        if (!bgzf_inflate_range(h_file, blocks, bi, bj, sl.p[slab], n_threads, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
        if (head.size() < (1u << 20)) head.insert(head.end(), sl.p[slab], sl.p[slab] + std::min<uint64_t>(bytes, (4u << 20)));
        CK(cudaMemcpyAsync(b->d_owned + blocks[bi].uoff, sl.p[slab], bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(sl.done[slab], ctx->stream));
        slab ^= 1;
        bi = bj;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Default: the COMPRESSED image goes up slab by slab on the copy stream (page cache -> pinned -> HBM) while a helper thread
// finds the BGZF block boundaries; from the moment the block table is known, the blocks that every landed slab completes
// are inflated on the side streams (inflate.cu). Upload, scan and inflate overlap: the load costs about max of the three.
int load_bgzf_device_inflate(svb_ctx *ctx, svb_bam *b, const uint8_t *h_file, uint64_t file_bytes, int n_threads,
                             std::vector<uint8_t> &head, uint32_t lead = 0,  // lead: bytes left free in front of the output
                             int fd = -1, uint64_t fd_base = 0)              // fd >= 0: h_file is the file at offset fd_base
{
    WallScope ws(ctx, "h2d_compressed+inflate(wall)", (double)file_bytes);
    static_assert(sizeof(BgzfBlock) == 24, "BgzfBlock must match the device-side block descriptor");
    std::vector<BgzfBlock> blocks;
    uint64_t total = 0;
    std::string scan_err;
    std::future<bool> scan = std::async(std::launch::async, [&]() { return bgzf_scan(h_file, file_bytes, blocks, total, scan_err); });
    struct Join {  // the helper thread writes into this frame: never leave while it runs
        std::future<bool> &f;
        ~Join()
        {
            if (f.valid()) f.wait();
        }
    } join{scan};
    BigBuf d_file{ctx};
    d_file.p = ctx->big_get(file_bytes + SVB_INFLATE_PAD, &d_file.cap);
    if (!d_file.p) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory", (unsigned long long)file_bytes);
    Slabs sl{ctx};
    if (!sl.init()) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate pinned staging");
    DevBuf<BgzfBlock> d_blocks;
    DevBuf<uint32_t> d_err;
    cudaEvent_t ready;
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    struct EventGuard {
        cudaEvent_t e;
        ~EventGuard() { cudaEventDestroy(e); }
    } ready_guard{ready};
    CK(cudaMemsetAsync(d_file.p + file_bytes, 0, SVB_INFLATE_PAD, ctx->stream));
    CK(cudaEventRecord(ready, ctx->stream));  // the image buffer is ordered before the copy stream
    CK(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
    bool have_blocks = false;
    auto start_inflate = [&]() -> int {  // the scan is in: allocate the output, ship the block table
        if (!scan.get()) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", scan_err.c_str());
        b->d_owned = ctx->big_get(lead + total + 256, &b->owned_cap);
        if (!b->d_owned) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory", (unsigned long long)total);
        b->d_data = b->d_owned + lead, b->nbytes = total;
        CK(cudaMemsetAsync(b->d_owned + lead + total, 0, 256, ctx->stream));
        CK(d_blocks.alloc(blocks.size(), ctx->stream));
        CK(d_err.alloc(1, ctx->stream));
        CK(cudaMemcpyAsync(d_blocks.p, blocks.data(), blocks.size() * sizeof(BgzfBlock), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(d_err.p, 0, 4, ctx->stream));
        CK(cudaEventRecord(ready, ctx->stream));  // output buffer and block table are ordered before the side streams
        for (cudaStream_t a : ctx->aux) CK(cudaStreamWaitEvent(a, ready, 0));
        have_blocks = true;
        return 0;
    };
    size_t next_block = 0;
    int lane = 0, slab = 0;
    StagePool stage(n_threads);
    for (uint64_t o = 0; o < file_bytes; o += Slabs::SLAB) {
        const uint64_t n = std::min(Slabs::SLAB, file_bytes - o);
        const bool last = o + n >= file_bytes;
        CK(cudaEventSynchronize(sl.done[slab]));
        {
            WallScope wc(ctx, "stage_copy(wall)", (double)n);
            stage.run(sl.p[slab], fd, fd_base + o, h_file + o, n);
        }
        CK(cudaMemcpyAsync(d_file.p + o, sl.p[slab], n, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(sl.done[slab], ctx->copy_stream));
        if (!have_blocks && (last || scan.wait_for(std::chrono::seconds(0)) == std::future_status::ready)) CKR(start_inflate());
        if (have_blocks) {
            size_t b1 = next_block;
            if (last) b1 = blocks.size();
            else
                while (b1 < blocks.size() && blocks[b1].coff + blocks[b1].clen <= o + n) ++b1;
            if (b1 > next_block) {
                cudaStream_t a = ctx->aux[lane];
                lane = (lane + 1) % svb_ctx::N_AUX;
                CK(cudaStreamWaitEvent(a, sl.done[slab], 0));  // (the copy stream is in order: this slab implies the earlier ones)
                CKR(inflate_launch(ctx, a, d_file.p, d_blocks.p + next_block, (uint32_t)(b1 - next_block), b->d_owned + lead, d_err.p));
                next_block = b1;
            }
        }
        slab ^= 1;
    }
    if (!have_blocks) CKR(start_inflate());  // (empty file)
    for (cudaStream_t a : ctx->aux) {  // join the side streams
        CK(cudaEventRecord(ready, a));
        CK(cudaStreamWaitEvent(ctx->stream, ready, 0));
    }
    uint32_t h_err = 0;
    CK(cudaMemcpyAsync(&h_err, d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    head.resize(std::min<uint64_t>(total, 1u << 20));
    CK(cudaMemcpyAsync(head.data(), b->d_data, head.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h_err) return svb_fail(ctx, SVB_ERR_FORMAT, "BGZF inflate failed (corrupt deflate stream)");
    return 0;
}
}  // namespace

// BGZF file image (host) -> uncompressed stream in HBM.
//  default : the COMPRESSED image is uploaded and inflated on the device, one warp per BGZF block (inflate.cu) - about a
//            third of the PCIe bytes and no host zlib;
//  SEEKSV_B200_HOST_INFLATE=1 : host threads inflate the blocks (zlib) and the uncompressed bytes are uploaded.
static int bam_from_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, int n_threads, svb_bam **out, int fd);
extern "C" int svb_bam_from_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, int n_threads, svb_bam **out)
{
    return bam_from_bgzf(ctx, h_file, file_bytes, n_threads, out, -1);
}
// fd >= 0: h_file is a mapping of that file (the slabs are then filled with pread instead of touching the mapping)
static int bam_from_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, int n_threads, svb_bam **out, int fd)
{
    if (!ctx || !out || !h_file) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_from_bgzf: null argument");
    CK(cudaSetDevice(ctx->device));
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::unique_ptr<svb_bam, void (*)(svb_bam *)> b(new svb_bam(), [](svb_bam *x) { svb_bam_free(x); });
    b->ctx = ctx;
    std::vector<uint8_t> head;  // first uncompressed bytes, for the header parse
    std::string err;
    if (use_host_inflate()) CKR(load_bgzf_host_inflate(ctx, b.get(), (const uint8_t *)h_file, file_bytes, n_threads, head));
    else CKR(load_bgzf_device_inflate(ctx, b.get(), (const uint8_t *)h_file, file_bytes, n_threads, head, 0, fd, 0));
    const uint64_t total = b->nbytes;
    BamHeader hdr;
    if (!parse_bam_header(head.data(), head.size(), hdr, err)) {
        // header longer than the captured prefix: fetch what is needed from the device copy
        std::vector<uint8_t> all(std::min<uint64_t>(total, 256ull << 20));
        CK(cudaMemcpy(all.data(), b->d_owned, all.size(), cudaMemcpyDeviceToHost));
        if (!parse_bam_header(all.data(), all.size(), hdr, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    }
    b->first = hdr.first_record, b->n_ref = (int32_t)hdr.names.size();
    b->names = hdr.names, b->lens = hdr.lengths;
    b->whole_file = true;  // the record chain has to end exactly at the end of the stream (checked when it is walked)
    svb_bam *raw = b.release();
    std::unique_ptr<svb_bam> plain(raw);
    return finish_bam(ctx, plain, out);
}

// Inflate an arbitrary BGZF image on the device and return the bytes (diagnostics / tests of the inflate kernel alone).
extern "C" int svb_inflate_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, void *h_out, uint64_t out_cap, uint64_t *out_len)
{
    if (!ctx || !h_file || !out_len) return svb_fail(ctx, SVB_ERR_ARG, "svb_inflate_bgzf: null argument");
    CK(cudaSetDevice(ctx->device));
    std::vector<BgzfBlock> blocks;
    uint64_t total = 0;
    std::string err;
    if (!bgzf_scan((const uint8_t *)h_file, file_bytes, blocks, total, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    *out_len = total;
    if (!h_out) return 0;  // size query
    if (out_cap < total) return svb_fail(ctx, SVB_ERR_ARG, "svb_inflate_bgzf: output buffer too small");
    DevBuf<uint8_t> d_file, d_out;
    DevBuf<BgzfBlock> d_blocks;
    CK(d_file.alloc(file_bytes + SVB_INFLATE_PAD, ctx->stream));
    CK(d_out.alloc(total + 256, ctx->stream));
    CK(d_blocks.alloc(blocks.size(), ctx->stream));
    CK(cudaMemsetAsync(d_file.p + file_bytes, 0, SVB_INFLATE_PAD, ctx->stream));
    CK(cudaMemcpyAsync(d_file.p, h_file, file_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_blocks.p, blocks.data(), blocks.size() * sizeof(BgzfBlock), cudaMemcpyHostToDevice, ctx->stream));
    CKR(inflate_on_device(ctx, d_file.p, d_blocks.p, (uint32_t)blocks.size(), d_out.p, (double)total));
    CK(cudaMemcpyAsync(h_out, d_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// A gzip text file written by this library's device gzip (gzip.cu: one member per 64 KiB of text, 'SV' size field) -> its text,
// inflated by the BGZF inflate kernel. SVB_ERR_FORMAT: not such a file (the caller reads it with svb_read_gz / zlib instead).
extern "C" int svb_read_gz_device(svb_ctx *ctx, const char *path, const char **data, uint64_t *n)
{
    if (!ctx || !path || !data || !n) return svb_fail(ctx, SVB_ERR_ARG, "svb_read_gz_device: null argument");
    CK(cudaSetDevice(ctx->device));
    *data = nullptr, *n = 0;
    MappedFile mf;
    std::string err;
    if (!mf.open(path, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
    const uint8_t *f = mf.data;
    const uint64_t size = mf.size;
    std::vector<BgzfBlock> blocks;
    uint64_t o = 0, total = 0;
    while (o < size) {
        if (o + 28 > size || f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || f[o + 3] != 4 || f[o + 10] != 8 || f[o + 11] != 0 ||
            f[o + 12] != 'S' || f[o + 13] != 'V' || f[o + 14] != 4 || f[o + 15] != 0)
            return svb_fail(ctx, SVB_ERR_FORMAT, "%s: not a gzip file of this library's device writer", path);
        const uint64_t sz = f[o + 16] | (f[o + 17] << 8) | (f[o + 18] << 16) | ((uint64_t)f[o + 19] << 24);
        if (sz < 28 || o + sz > size) return svb_fail(ctx, SVB_ERR_FORMAT, "%s: member size beyond the file", path);
        const uint8_t *t = f + o + sz - 4;
        const uint32_t ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
        if (ulen > 65536) return svb_fail(ctx, SVB_ERR_FORMAT, "%s: members of more than 64 KiB (a host-written file)", path);
        if (ulen) blocks.push_back(BgzfBlock{o + 20, total, (uint32_t)(sz - 28), ulen});
        total += ulen;
        o += sz;
    }
    if (blocks.size() >= (1ull << 32)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s: too many members", path);
    if (ctx->read_buf) ctx->pinned_put(ctx->read_buf, ctx->read_cap), ctx->read_buf = nullptr, ctx->read_cap = 0;
    ctx->read_buf = ctx->pinned_get(total + 1, &ctx->read_cap);
    if (!ctx->read_buf) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of pinned memory", (unsigned long long)total);
    if (total) {
        cudaStream_t s = ctx->stream;
        DevBuf<uint8_t> d_file, d_out;
        DevBuf<BgzfBlock> d_blocks;
        CK(d_file.alloc(size + SVB_INFLATE_PAD, s));
        CK(d_out.alloc(total + 256, s));
        CK(d_blocks.alloc(blocks.size(), s));
        CK(cudaMemsetAsync(d_file.p + size, 0, SVB_INFLATE_PAD, s));
        CK(cudaMemcpyAsync(d_file.p, f, size, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_blocks.p, blocks.data(), blocks.size() * sizeof(BgzfBlock), cudaMemcpyHostToDevice, s));
        CKR(inflate_on_device(ctx, d_file.p, d_blocks.p, (uint32_t)blocks.size(), d_out.p, (double)total));
        CK(cudaMemcpyAsync(ctx->read_buf, d_out.p, total, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    ctx->read_buf[total] = 0;
    *data = ctx->read_buf, *n = total;
    return 0;
}

extern "C" int svb_bam_copy_stream(const svb_bam *b, void *h_dst, uint64_t offset, uint64_t nbytes)
{
    if (!b || !h_dst || offset + nbytes > b->nbytes) return SVB_ERR_ARG;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    return cudaMemcpy(h_dst, b->d_data + offset, nbytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

extern "C" int svb_bam_open(svb_ctx *ctx, const char *path, int n_threads, svb_bam **out)
{
    if (!ctx || !path || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_open: null argument");
    std::string p(path), err;
    // the reference treats a name ending in ".bam" as BAM and anything else as SAM text (clip_reads.h:367-373)
    if (p.size() >= 4 && p.rfind(".bam") == p.size() - 4) {
        // map the file: the inflate threads read the compressed blocks straight from the page cache
        MappedFile mf;
        {
            WallScope ws(ctx, "file_map(wall)");
            if (!mf.open(p, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
        }
        int rc = bam_from_bgzf(ctx, mf.data, mf.size, n_threads, out, mf.fd);
        if (mf.data) {  // tearing down a mapping of this size takes milliseconds: not on the caller's time
            std::thread([d = mf.data, n = mf.size]() { munmap((void *)d, n); }).detach();
            mf.data = nullptr;
        }
        return rc;
    }
    std::vector<uint8_t> file;
    if (!read_file(p, file, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
    BamHeader hdr;
    std::vector<uint8_t> stream;
    if (!sam_to_bam_stream(file, hdr, stream, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    CKR(svb_bam_from_host(ctx, stream.data(), stream.size(), hdr.first_record, (int32_t)hdr.names.size(), out));
    (*out)->names = hdr.names, (*out)->lens = hdr.lengths;
    return 0;
}

// The records between two BGZF virtual offsets (both record boundaries, e.g. from the .bai): only the blocks in between are
// read, uploaded and inflated; the view starts on a 16-byte boundary. v0 == NONE: an empty view.
static int open_between(svb_ctx *ctx, const MappedFile &mf, const BamHeader &hdr, uint64_t v0, uint64_t v1, int n_threads, svb_bam **out)
{
    const uint64_t NONE = ~0ull;
    std::string err;
    std::unique_ptr<svb_bam, void (*)(svb_bam *)> b(new svb_bam(), [](svb_bam *x) { svb_bam_free(x); });
    b->ctx = ctx;
    b->first = 0, b->n_ref = (int32_t)hdr.names.size(), b->names = hdr.names, b->lens = hdr.lengths;
    b->whole_file = true;  // the view ends on a record boundary: the chain has to end exactly there
    if (v0 != NONE && v0 != v1) {
        const uint64_t c0 = v0 >> 16, u0 = v0 & 0xffff;
        uint64_t c_end = mf.size, u1 = 0;
        bool cut_last = false;
        if (v1 != NONE) {
            c_end = v1 >> 16, u1 = v1 & 0xffff;
            if (u1) {  // the block that holds the boundary is needed too
                uint32_t bs = c_end < mf.size ? bgzf_block_size(mf.data, mf.size, c_end) : 0;
                if (!bs) return svb_fail(ctx, SVB_ERR_FORMAT, "virtual offset does not point at a BGZF block");
                c_end += bs;
                cut_last = true;
            }
        }
        if (c0 >= c_end || c_end > mf.size) return svb_fail(ctx, SVB_ERR_FORMAT, "virtual offsets out of range");
        std::vector<uint8_t> head;
        const uint32_t lead = (uint32_t)((16 - (u0 & 15)) & 15);  // the view's first byte lands on a 16-byte boundary
        CKR(load_bgzf_device_inflate(ctx, b.get(), mf.data + c0, c_end - c0, n_threads, head, lead, mf.fd, c0));
        uint64_t total = b->nbytes, tail_cut = 0;
        if (cut_last) {
            const uint8_t *t = mf.data + c_end - 4;  // uncompressed size of the last block = its ISIZE field
            uint64_t last_ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint64_t)t[3] << 24);
            if (u1 > last_ulen) return svb_fail(ctx, SVB_ERR_FORMAT, "virtual offset beyond its block");
            tail_cut = last_ulen - u1;
        }
        if (u0 + tail_cut > total) return svb_fail(ctx, SVB_ERR_FORMAT, "virtual offsets out of range");
        b->d_data += u0, b->nbytes = total - u0 - tail_cut;
    } else {  // no records in the range: an empty stream (on a real allocation, so that every kernel has a valid pointer)
        b->d_owned = ctx->big_get(4096, &b->owned_cap);
        if (!b->d_owned) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate device memory");
        CK(cudaMemsetAsync(b->d_owned, 0, 4096, ctx->stream));
        b->d_data = b->d_owned, b->nbytes = 0;
    }
    svb_bam *raw = b.release();
    std::unique_ptr<svb_bam> plain(raw);
    return finish_bam(ctx, plain, out);
}

// Records of references [tid_begin, tid_end) of a coordinate-sorted, indexed BAM (the chromosome shard of one rank). The
// .bai gives the BGZF virtual offset of the first record of every reference (bam_index_build, sam/bam.h:498-536), so the
// shard is cut at exact record boundaries. The range that reaches the last reference with records also takes the
// unplaced reads at the end of the file.
extern "C" int svb_bam_open_refs(svb_ctx *ctx, const char *bam_path, const char *bai_path, int32_t tid_begin, int32_t tid_end,
                                 int n_threads, svb_bam **out)
{
    if (!ctx || !bam_path || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_open_refs: null argument");
    CK(cudaSetDevice(ctx->device));
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::string err, bai = bai_path ? bai_path : std::string(bam_path) + ".bai";
    MappedFile mf;
    if (!mf.open(bam_path, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
    BamHeader hdr;
    if (!read_bam_header(mf.data, mf.size, hdr, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    std::vector<uint64_t> first;
    if (!bai_first_offsets(bai, first, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
    const int32_t n_ref = (int32_t)hdr.names.size();
    if (first.size() != (size_t)n_ref) return svb_fail(ctx, SVB_ERR_FORMAT, "%s does not belong to %s (reference count)", bai.c_str(), bam_path);
    tid_begin = std::max(tid_begin, 0), tid_end = std::min(tid_end, n_ref);
    const uint64_t NONE = ~0ull;
    uint64_t v0 = NONE, v1 = NONE;  // virtual offsets: first record of the range, first record after it (NONE: end of file)
    for (int32_t t = tid_begin; t < tid_end && v0 == NONE; ++t) v0 = first[t];
    for (int32_t t = tid_end; t < n_ref && v1 == NONE; ++t) v1 = first[t];
    return open_between(ctx, mf, hdr, v0, v1, n_threads, out);
}

// Coordinate-range shard: the records between two virtual offsets taken from the .bai's linear index (record boundaries).
// v_begin == 0: from the first record of the file; v_end == ~0: to the end of the file.
extern "C" int svb_bam_open_voffsets(svb_ctx *ctx, const char *bam_path, uint64_t v_begin, uint64_t v_end, int n_threads, svb_bam **out)
{
    if (!ctx || !bam_path || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_open_voffsets: null argument");
    CK(cudaSetDevice(ctx->device));
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::string err;
    MappedFile mf;
    if (!mf.open(bam_path, err)) return svb_fail(ctx, SVB_ERR_IO, "%s", err.c_str());
    BamHeader hdr;
    if (!read_bam_header(mf.data, mf.size, hdr, err)) return svb_fail(ctx, SVB_ERR_FORMAT, "%s", err.c_str());
    if (v_begin == 0) {  // virtual offset of the byte after the header
        uint64_t c = 0, left = hdr.first_record;
        for (;;) {
            uint32_t bs = c < mf.size ? bgzf_block_size(mf.data, mf.size, c) : 0;
            if (!bs) return svb_fail(ctx, SVB_ERR_FORMAT, "not a BGZF file (bad block header)");
            const uint8_t *t = mf.data + c + bs - 4;
            uint64_t ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint64_t)t[3] << 24);
            if (left < ulen) break;
            left -= ulen, c += bs;
            if (c >= mf.size) break;
        }
        v_begin = c >= mf.size ? ~0ull : (c << 16 | left);
    }
    return open_between(ctx, mf, hdr, v_begin, v_end, n_threads, out);
}

extern "C" void svb_bam_free(svb_bam *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaStream_t s = b->ctx->stream;
    free_rows(b);
    void *cols[4] = {b->d_guess /* + base, exit, count: one allocation */, b->d_fkey, b->d_scal, b->d_ref_len};
    for (void *c : cols)
        if (c) cudaFreeAsync(c, s);
    if (b->d_owned) b->ctx->big_put(b->d_owned, b->owned_cap);  // (the stream was synchronised above)
    delete b;
}

extern "C" int svb_bam_device_stream(const svb_bam *b, const void **d, uint64_t *n, uint64_t *first)
{
    if (!b || !d || !n || !first) return SVB_ERR_ARG;
    *d = b->d_data, *n = b->nbytes, *first = b->first;
    return 0;
}
// counts are produced by the first pass that walks the chain (lazily here when nobody has walked it yet)
extern "C" uint64_t svb_bam_n_records(const svb_bam *b)
{
    if (!b) return 0;
    if (!b->counted && ensure_counts(b->ctx, const_cast<svb_bam *>(b)) != 0) return 0;
    return b->n_rec;
}
extern "C" uint64_t svb_bam_record_bytes(const svb_bam *b)
{
    if (!b) return 0;
    if (!b->counted && ensure_counts(b->ctx, const_cast<svb_bam *>(b)) != 0) return 0;
    return b->rec_bytes;
}
extern "C" int32_t svb_bam_n_ref(const svb_bam *b) { return b ? b->n_ref : 0; }
extern "C" const char *svb_bam_ref_name(const svb_bam *b, int32_t tid)
{
    return (b && tid >= 0 && (size_t)tid < b->names.size()) ? b->names[tid].c_str() : nullptr;
}
extern "C" uint32_t svb_bam_ref_len(const svb_bam *b, int32_t tid)
{
    return (b && tid >= 0 && (size_t)tid < b->lens.size()) ? b->lens[tid] : 0;
}
extern "C" int svb_bam_set_refs(svb_bam *b, int32_t n_ref, const char *const *names, const uint32_t *lengths)
{
    if (!b || n_ref != b->n_ref || !names || !lengths) return SVB_ERR_ARG;
    b->names.assign(names, names + n_ref);
    b->lens.assign(lengths, lengths + n_ref);
    return 0;
}
