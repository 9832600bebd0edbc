// context.cu — per-thread context, error plumbing, buffer pools, spine scan, library lifetime.
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace rsn {

static thread_local Ctx g_ctx;

Ctx &ctx() { return g_ctx; }

static std::atomic<uint64_t> g_launches{0};  // all threads (batch workers launch from their own)
void count_launch() {
    g_ctx.launches++;
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// ----------------------------------------------------------------------------- stream waits

namespace {
thread_local bool t_blocking_sync = false;
thread_local cudaEvent_t t_wait_ev = nullptr;
thread_local int t_wait_ev_dev = -1;
}  // namespace
void set_thread_blocking_sync(bool on) { t_blocking_sync = on; }
cudaError_t stream_wait(cudaStream_t s) {
    if (!t_blocking_sync) return cudaStreamSynchronize(s);
    int dev = -1;
    cudaGetDevice(&dev);
    if (!t_wait_ev || t_wait_ev_dev != dev) {
        if (t_wait_ev) cudaEventDestroy(t_wait_ev);
        t_wait_ev = nullptr;
        if (cudaEventCreateWithFlags(&t_wait_ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return cudaStreamSynchronize(s);
        }
        t_wait_ev_dev = dev;
    }
    cudaError_t e = cudaEventRecord(t_wait_ev, s);
    if (e != cudaSuccess) return e;
    // Most waits of a batch stage are for kernels of 10-150 us; waking a sleeping thread costs about
    // as much again (the single-group timeline showed 0.2-0.6 ms between dependent kernels).  Poll
    // for a short while first, sleep only behind the long kernels (match search, Huffman trees).
    static const long spin_us = [] {
        const char *v = getenv("RSN_WAIT_SPIN_US");
        return v ? atol(v) : 120L;
    }();
    if (spin_us > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            e = cudaEventQuery(t_wait_ev);
            if (e != cudaErrorNotReady) return e;
            if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us)
                break;
        }
    }
    return cudaEventSynchronize(t_wait_ev);
}

// ----------------------------------------------------------------------------- per-kernel timing

namespace {
struct KTimer {
    std::atomic<bool> on{false};
    std::mutex mu;
    struct Rec {
        const char *name;
        cudaEvent_t a, b;
        cudaStream_t s;
    };
    std::vector<Rec> recs;
};
KTimer &ktimer() {
    static KTimer *t = new KTimer();
    return *t;
}
}  // namespace

int ktime_begin(const char *name, cudaStream_t s) {
    KTimer &t = ktimer();
    if (!t.on.load(std::memory_order_relaxed)) return -1;
    KTimer::Rec r{name, nullptr, nullptr, s};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cudaEventRecord(r.a, s);
    std::lock_guard<std::mutex> g(t.mu);
    t.recs.push_back(r);
    return (int)t.recs.size() - 1;
}
void ktime_end(int idx, cudaStream_t s) {
    KTimer &t = ktimer();
    cudaEvent_t b;
    {
        std::lock_guard<std::mutex> g(t.mu);
        b = t.recs[(size_t)idx].b;
    }
    cudaEventRecord(b, s);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    Ctx &c = ctx();
    snprintf(c.cuda_err, sizeof(c.cuda_err), "%s: %s at %s:%d (%s)", cudaGetErrorName(e), cudaGetErrorString(e), file,
             line, what);
    cudaGetLastError();  // clear non-sticky errors
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return RSN_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return RSN_ERR_NOMEM;
    return RSN_ERR_CUDA;
}

static void release_thread_resources();
void touch_thread_exit();

static int init_ctx(int device) {
    Ctx &c = ctx();
    int count = 0;
    RSN_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) return RSN_ERR_NO_DEVICE;
    if (device < 0) {
        RSN_CUDA(cudaGetDevice(&device));
    }
    if (device >= count) return RSN_ERR_INVALID_ARG;
    if (c.ready && c.device == device) {
        RSN_CUDA(cudaSetDevice(device));
        return RSN_OK;
    }
    if (c.ready) release_thread_resources();  // streams, events and arena blocks belong to the old device
    RSN_CUDA(cudaSetDevice(device));
    c.device = device;
    if (!c.own_stream) {
        int least = 0, greatest = 0;
        RSN_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        RSN_CUDA(cudaStreamCreateWithPriority(&c.own_stream, cudaStreamNonBlocking, greatest));
        if (!c.low_stream) RSN_CUDA(cudaStreamCreateWithPriority(&c.low_stream, cudaStreamNonBlocking, least));
        if (!c.batch_stream) {
            static std::atomic<int> ordinal{0};
            const int levels = least - greatest;  // priorities greatest .. least-1 (least is the match search's)
            const int k = ordinal.fetch_add(1);
            const int prio = levels > 0 ? greatest + (k % levels) : greatest;
            RSN_CUDA(cudaStreamCreateWithPriority(&c.batch_stream, cudaStreamNonBlocking, prio));
        }
        if (!c.low_before) RSN_CUDA(cudaEventCreateWithFlags(&c.low_before, cudaEventDisableTiming));
        if (!c.low_after) RSN_CUDA(cudaEventCreateWithFlags(&c.low_after, cudaEventDisableTiming));
    }
    if (!c.copy_stream) RSN_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (auto &e : c.chunk_ev)
        if (!e) RSN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!c.h_scalars) RSN_CUDA(cudaHostAlloc((void **)&c.h_scalars, 64 * sizeof(uint64_t), cudaHostAllocDefault));
    c.ready = true;
    touch_thread_exit();
    return RSN_OK;
}

int ensure_ctx() {
    Ctx &c = ctx();
    if (c.ready) {
        int cur = -1;
        RSN_CUDA(cudaGetDevice(&cur));
        if (cur != c.device) RSN_CUDA(cudaSetDevice(c.device));
        return RSN_OK;
    }
    return init_ctx(-1);
}

// RSN_ALLOC_TRACE=1: one line on stderr for every allocation that had to go to the driver (arena
// growth, result-buffer or pinned-pool miss) with the time it took: these are the calls that stall
// every other thread of the process.
static bool alloc_trace() {
    static const bool on = [] {
        const char *v = getenv("RSN_ALLOC_TRACE");
        return v && v[0] == '1';
    }();
    return on;
}
struct AllocTimer {
    const char *what;
    size_t bytes;
    std::chrono::steady_clock::time_point t0;
    AllocTimer(const char *w, size_t n) : what(w), bytes(n), t0(std::chrono::steady_clock::now()) {}
    ~AllocTimer() {
        if (!alloc_trace()) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[rsn alloc] %-12s %10zu bytes %8.3f ms\n", what, bytes, ms);
    }
};

// ----------------------------------------------------------------------------- arena

namespace {
struct ArenaBlock {
    char *p;
    size_t cap, off;
};
// A call's temporaries are bump-allocated from blocks its thread holds for the duration of the
// outermost ArenaScope.  The blocks themselves belong to the PROCESS: a scope takes them from a pool
// (best fit) and hands them back when it ends, each with an event recorded on the scope's stream that
// the next taker's stream waits for.  Blocks used to belong to threads, and a batch worker whose
// share of a later pass was larger than anything it had seen before grew its arena in the middle of
// the pass: a 1.8 GB cudaMalloc (50-120 ms), then cudaFree of the old chain (which waits for every
// stream of the device) and a second cudaMalloc — one 4096-file pass in five took 270-2100 ms instead
// of 205.  With the pool whichever worker draws the large group finds the large block.
struct Arena {
    std::vector<ArenaBlock> blocks;
    size_t cur = 0;
    int depth = 0;
    cudaStream_t last_stream = nullptr;
    bool used = false;
};
thread_local Arena g_arena;

constexpr size_t kArenaAlign = 256;
constexpr size_t kArenaMinBlock = (size_t)128 << 20;

struct ArenaPool {
    std::mutex mu;
    struct Ent {
        char *p;
        size_t cap;
        int device;
        cudaStream_t last;
        cudaEvent_t ev;
    };
    std::vector<Ent> free_;
    std::unordered_map<int, std::vector<cudaEvent_t>> evs;

    // smallest free block of this device that holds `need` bytes; its previous user's work is
    // ordered before anything the taker queues on `s`
    bool take(int dev, size_t need, cudaStream_t s, ArenaBlock &out) {
        Ent e{};
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].device == dev && free_[i].cap >= need && (best == free_.size() || free_[i].cap < free_[best].cap))
                    best = i;
            if (best == free_.size()) return false;
            e = free_[best];
            free_[best] = free_.back();
            free_.pop_back();
        }
        if (e.ev) {
            if (e.last != s && cudaStreamWaitEvent(s, e.ev, 0) != cudaSuccess) {
                cudaGetLastError();
                cudaEventSynchronize(e.ev);
            }
            std::lock_guard<std::mutex> g(mu);
            evs[dev].push_back(e.ev);
        }
        out = ArenaBlock{e.p, e.cap, 0};
        return true;
    }
    void give(int dev, const ArenaBlock &b, cudaStream_t s) {
        cudaEvent_t ev = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            auto &pool = evs[dev];
            if (!pool.empty()) {
                ev = pool.back();
                pool.pop_back();
            }
        }
        if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            ev = nullptr;
        }
        if (ev && cudaEventRecord(ev, s) != cudaSuccess) {
            cudaGetLastError();
            cudaEventDestroy(ev);
            ev = nullptr;
        }
        if (!ev) cudaStreamSynchronize(s);  // no event: the block goes back only once its work is done
        std::lock_guard<std::mutex> g(mu);
        free_.push_back(Ent{b.p, b.cap, dev, s, ev});
    }
    void forget_stream(cudaStream_t s) {
        std::lock_guard<std::mutex> g(mu);
        for (auto &e : free_)
            if (e.last == s) e.last = nullptr;
    }
    void drain() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &e : free_) {
            if (e.ev) {
                cudaEventSynchronize(e.ev);
                cudaEventDestroy(e.ev);
            }
            cudaFree(e.p);
        }
        free_.clear();
        for (auto &kv : evs)
            for (cudaEvent_t e : kv.second) cudaEventDestroy(e);
        evs.clear();
    }
};
ArenaPool &arena_pool() {
    static ArenaPool *p = new ArenaPool();  // leaked on purpose: safe at process exit
    return *p;
}

// blocks a thread still holds (a scope that was never closed, or a thread that dies inside one)
void arena_free_all() {
    for (auto &b : g_arena.blocks) cudaFree(b.p);
    g_arena.blocks.clear();
    g_arena.cur = 0;
}
}  // namespace

// Everything the calling thread owns on its current device: streams, events, pinned scalars, arena.
// Runs on rsn_shutdown, when the thread is re-initialised for another device, and when the thread
// exits (cgo callers migrate over many short-lived OS threads).
void outs_forget_stream(cudaStream_t s);

static void release_thread_resources() {
    Ctx &c = g_ctx;
    if (!c.ready && !c.own_stream && g_arena.blocks.empty()) return;
    if (c.device >= 0) cudaSetDevice(c.device);
    // work that still reads arena memory may be queued on a caller's stream (or on own_stream, which
    // is destroyed below): wait for it first
    if (g_arena.used && g_arena.last_stream && g_arena.last_stream != c.own_stream)
        cudaStreamSynchronize(g_arena.last_stream);
    if (c.own_stream) {
        cudaStreamSynchronize(c.own_stream);
        outs_forget_stream(c.own_stream);
        cudaStreamDestroy(c.own_stream);
        c.own_stream = nullptr;
    }
    if (c.copy_stream) {
        cudaStreamSynchronize(c.copy_stream);
        outs_forget_stream(c.copy_stream);
        cudaStreamDestroy(c.copy_stream);
        c.copy_stream = nullptr;
    }
    if (c.low_stream) {
        cudaStreamSynchronize(c.low_stream);
        cudaStreamDestroy(c.low_stream);
        c.low_stream = nullptr;
    }
    if (c.batch_stream) {
        cudaStreamSynchronize(c.batch_stream);
        outs_forget_stream(c.batch_stream);
        cudaStreamDestroy(c.batch_stream);
        c.batch_stream = nullptr;
    }
    if (c.low_before) cudaEventDestroy(c.low_before);
    if (c.low_after) cudaEventDestroy(c.low_after);
    c.low_before = c.low_after = nullptr;
    for (auto &e : c.chunk_ev)
        if (e) {
            cudaEventDestroy(e);
            e = nullptr;
        }
    if (c.h_scalars) {
        cudaFreeHost(c.h_scalars);
        c.h_scalars = nullptr;
    }
    arena_free_all();
    g_arena.used = false;
    g_arena.last_stream = nullptr;
    c.ready = false;
    cudaGetLastError();
}
namespace {
struct ThreadExit {
    ~ThreadExit() { release_thread_resources(); }
};
thread_local ThreadExit g_thread_exit;
}  // namespace
void touch_thread_exit() { (void)&g_thread_exit; }

void *arena_alloc(size_t n) {
    Arena &a = g_arena;
    n = (n + kArenaAlign - 1) & ~(kArenaAlign - 1);
    if (n == 0) n = kArenaAlign;
    for (; a.cur < a.blocks.size(); a.cur++) {
        ArenaBlock &b = a.blocks[a.cur];
        if (b.off + n <= b.cap) {
            void *p = b.p + b.off;
            b.off += n;
            return p;
        }
        if (a.cur + 1 < a.blocks.size()) a.blocks[a.cur + 1].off = 0;
    }
    ArenaBlock nb{};
    if (!arena_pool().take(g_ctx.device, n, a.last_stream, nb)) {
        size_t cap = n + n / 4;  // some headroom: the same call on slightly larger data still fits
        if (cap < kArenaMinBlock) cap = kArenaMinBlock;
        char *p = nullptr;
        AllocTimer at("arena", cap);
        if (cudaMalloc((void **)&p, cap) != cudaSuccess) {
            cudaGetLastError();
            if (cap == n || cudaMalloc((void **)&p, n) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            cap = n;
        }
        nb = ArenaBlock{p, cap, 0};
    }
    nb.off = n;
    a.blocks.push_back(nb);
    a.cur = a.blocks.size() - 1;
    return nb.p;
}

ArenaScope::ArenaScope(cudaStream_t stream) : s(stream) {
    Arena &a = g_arena;
    if (a.depth == 0) {
        a.last_stream = stream;
        a.used = true;
        a.cur = 0;
    }
    saved_block = a.cur;
    saved_off = a.blocks.empty() ? 0 : a.blocks[a.cur].off;
    a.depth++;
}

ArenaScope::~ArenaScope() {
    Arena &a = g_arena;
    a.depth--;
    if (a.depth > 0) {
        // (a scope that opened on an empty arena saved block 0, offset 0: the start of the first block)
        a.cur = saved_block;
        if (!a.blocks.empty()) a.blocks[a.cur].off = saved_off;
        return;
    }
    // outermost scope: every block back to the pool, ordered behind this call's work
    for (auto &b : a.blocks) arena_pool().give(g_ctx.device, b, s);
    a.blocks.clear();
    a.cur = 0;
}

// ----------------------------------------------------------------------------- result buffers

namespace {
struct OutCache {
    std::mutex mu;
    struct Ent {
        size_t cap;
        cudaStream_t last;
        int device;
    };
    std::unordered_map<void *, Ent> live;
    // (device, size class) -> free buffers: a buffer only ever goes back to a caller on its own device.
    // A free buffer remembers the stream that used it last and an event recorded there when it was
    // freed: the next user, if it runs on another stream, makes ITS STREAM wait for that event.  (The
    // host used to wait for the whole other stream here — with eight group workers handing
    // intermediate buffers to one another that stalled a worker behind another worker's queue of
    // match-search kernels, and a batch pass took anything between 73 and 460 ms.)
    struct Free {
        void *first;
        cudaStream_t second;
        cudaEvent_t ev;
    };
    std::unordered_map<uint64_t, std::vector<Free>> free_;
    std::unordered_map<int, std::vector<cudaEvent_t>> ev_pool;  // per device
    static uint64_t key(int device, size_t c) { return ((uint64_t)(uint32_t)device << 48) ^ (uint64_t)c; }
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)16 << 30;

    static size_t size_class(size_t n) {
        if (n < 4096) return 4096;
        size_t p2 = (size_t)1 << (63 - __builtin_clzll((unsigned long long)n));
        size_t step = p2 / 8;
        return (n + step - 1) / step * step;
    }
    void *get(size_t n, cudaStream_t s) {
        const size_t c = size_class(n);
        const int dev = g_ctx.device;
        void *p = nullptr;
        cudaStream_t last = nullptr;
        cudaEvent_t ev = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_.find(key(dev, c));
            if (it != free_.end() && !it->second.empty()) {
                p = it->second.back().first;
                last = it->second.back().second;
                ev = it->second.back().ev;
                it->second.pop_back();
                cached -= c;
                live[p] = Ent{c, s, dev};
            }
        }
        if (p) {
            if (last && last != s) {  // its previous user may still be reading it
                if (!ev || cudaStreamWaitEvent(s, ev, 0) != cudaSuccess) {
                    cudaGetLastError();
                    cudaStreamSynchronize(last);
                }
            }
            if (ev) {
                std::lock_guard<std::mutex> g(mu);
                ev_pool[dev].push_back(ev);
            }
            return p;
        }
        AllocTimer at("result buf", c);
        if (cudaMalloc(&p, c) != cudaSuccess) {
            cudaGetLastError();
            drain();
            if (cudaMalloc(&p, c) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        std::lock_guard<std::mutex> g(mu);
        live[p] = Ent{c, s, dev};
        return p;
    }
    void put(void *p, cudaStream_t s) {
        if (!p) return;
        size_t c = 0;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = live.find(p);
            if (it == live.end()) return;  // not ours
            c = it->second.cap;
            const int dev = it->second.device;
            if (cached + c <= kMaxCached) {
                cudaEvent_t ev = nullptr;
                if (s) {
                    auto &pool = ev_pool[dev];
                    if (!pool.empty()) {
                        ev = pool.back();
                        pool.pop_back();
                    } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
                        cudaGetLastError();
                        ev = nullptr;
                    }
                    if (ev && cudaEventRecord(ev, s) != cudaSuccess) {
                        cudaGetLastError();
                        pool.push_back(ev);
                        ev = nullptr;
                    }
                }
                live.erase(it);
                free_[key(dev, c)].push_back(Free{p, s, ev});
                cached += c;
                return;
            }
            live.erase(it);
        }
        cudaStreamSynchronize(s);
        cudaFree(p);
    }
    void drain() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : free_)
            for (auto &e : kv.second) {
                cudaFree(e.first);
                if (e.ev) cudaEventDestroy(e.ev);
            }
        free_.clear();
        for (auto &kv : ev_pool)
            for (cudaEvent_t e : kv.second) cudaEventDestroy(e);
        ev_pool.clear();
        cached = 0;
    }
    // A stream is about to be destroyed (its owner has synchronised it): cached buffers last used
    // on it need no further ordering, and must not name it any more.
    void forget_stream(cudaStream_t s) {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : free_)
            for (auto &e : kv.second)
                if (e.second == s) e.second = nullptr;
        for (auto &kv : live)
            if (kv.second.last == s) kv.second.last = nullptr;
    }
};
OutCache &outs() {
    static OutCache *c = new OutCache();
    return *c;
}
}  // namespace

void outs_forget_stream(cudaStream_t s) {
    outs().forget_stream(s);
    arena_pool().forget_stream(s);
}
void arena_drain() { arena_pool().drain(); }
void *out_alloc(size_t n, cudaStream_t s) { return outs().get(n ? n : 1, s); }
void out_free(void *p, cudaStream_t s) { outs().put(p, s); }

int DevBuf::alloc(size_t n, cudaStream_t stream) {
    reset();
    s = stream;
    is_out = false;
    bytes = n ? n : 1;
    p = arena_alloc(bytes);
    if (!p) {
        snprintf(ctx().cuda_err, sizeof(ctx().cuda_err), "device arena: cannot allocate %zu bytes", bytes);
        return RSN_ERR_NOMEM;
    }
    return RSN_OK;
}

int DevBuf::alloc_out(size_t n, cudaStream_t stream) {
    reset();
    s = stream;
    is_out = true;
    bytes = n ? n : 1;
    p = out_alloc(bytes, stream);
    if (!p) {
        snprintf(ctx().cuda_err, sizeof(ctx().cuda_err), "device result buffer: cannot allocate %zu bytes", bytes);
        return RSN_ERR_NOMEM;
    }
    return RSN_OK;
}

void DevBuf::reset() {
    if (p && is_out) out_free(p, s);  // arena temporaries are rewound by the ArenaScope
    p = nullptr;
}

// ----------------------------------------------------------------------------- tracer

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
Trace::Trace(const char *w, cudaStream_t stream) : s(stream), what(w) {
    static const bool enabled = getenv("RSN_TRACE") != nullptr;
    on = enabled;
    if (on) {
        cudaStreamSynchronize(s);
        t0 = now_ms();
    }
}
void Trace::mark(const char *label) {
    if (!on) return;
    cudaStreamSynchronize(s);
    const double t = now_ms();
    fprintf(stderr, "[rsn trace] %-10s %-22s %8.3f ms\n", what, label, t - t0);
    t0 = t;
}

// ----------------------------------------------------------------------------- spine scan

// Exclusive scan of u64 values.  Small inputs: one CTA.  Large inputs: per-chunk totals, a
// recursive scan of those, then a per-chunk scan seeded with its offset.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) k_scan_single(const uint64_t *__restrict__ in,
                                                               uint64_t *__restrict__ out,
                                                               uint64_t *__restrict__ total, size_t count) {
    __shared__ uint64_t sm[33];
    uint64_t carry = 0;
    for (size_t base = 0; base < count; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        uint64_t v = i < count ? in[i] : 0;
        uint64_t tot;
        uint64_t ex = block_exclusive_sum<uint64_t>(v, sm, tot);
        if (i < count) out[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_chunk_totals(const uint64_t *__restrict__ in, size_t count,
                                                                     uint64_t *__restrict__ chunk_tot) {
    __shared__ uint64_t sm[33];
    const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < count) v += in[base + k];
    uint64_t tot;
    block_exclusive_sum<uint64_t>(v, sm, tot);
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_chunks(const uint64_t *__restrict__ in, size_t count,
                                                               const uint64_t *__restrict__ chunk_off,
                                                               uint64_t *__restrict__ out) {
    __shared__ uint64_t sm[33];
    const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
    uint64_t x[kScanItems];
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        x[k] = base + k < count ? in[base + k] : 0;
        v += x[k];
    }
    uint64_t tot;
    uint64_t run = chunk_off[blockIdx.x] + block_exclusive_sum<uint64_t>(v, sm, tot);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < count) out[base + k] = run;
        run += x[k];
    }
}

int spine_scan_u64(const uint64_t *d_in, uint64_t *d_out, uint64_t *d_total, size_t count, cudaStream_t s) {
    if (count <= (size_t)kScanChunk * 2) {
        RSN_LAUNCH(k_scan_single, 1, kScanThreads, 0, s, d_in, d_out, d_total, count);
        return RSN_OK;
    }
    const size_t chunks = div_up(count, kScanChunk);
    DevBuf tot, off;
    RSN_TRY(tot.alloc(chunks * 8, s));
    RSN_TRY(off.alloc(chunks * 8, s));
    RSN_LAUNCH(k_scan_chunk_totals, (unsigned)chunks, kScanThreads, 0, s, d_in, count, tot.as<uint64_t>());
    RSN_TRY(spine_scan_u64(tot.as<uint64_t>(), off.as<uint64_t>(), d_total, chunks, s));
    RSN_LAUNCH(k_scan_chunks, (unsigned)chunks, kScanThreads, 0, s, d_in, count, off.as<uint64_t>(), d_out);
    return RSN_OK;
}

cudaError_t copy_many(void *const *dsts, const void *const *srcs, const size_t *sizes, size_t count, cudaStream_t s) {
    std::vector<void *> d, sr;
    std::vector<size_t> sz;
    d.reserve(count);
    sr.reserve(count);
    sz.reserve(count);
    for (size_t i = 0; i < count; i++)
        if (sizes[i]) {
            d.push_back(dsts[i]);
            sr.push_back(const_cast<void *>(srcs[i]));
            sz.push_back(sizes[i]);
        }
    if (d.empty()) return cudaSuccess;
    static std::atomic<bool> batch_ok{true};
    if (d.size() > 1 && batch_ok.load(std::memory_order_relaxed)) {
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;  // the sources stay valid until the stream has passed the copies
        at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t first = 0, fail = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(d.data(), sr.data(), sz.data(), d.size(), &at, &first, 1, &fail, s);
        if (e == cudaSuccess) return e;
        cudaGetLastError();
        // an older driver: one by one from now on; anything else the batch call refuses (a kind of
        // memory, an attribute): this batch one by one, and the plain copies report what is wrong
        if (e == cudaErrorNotSupported || e == cudaErrorCallRequiresNewerDriver || e == cudaErrorInsufficientDriver)
            batch_ok.store(false);
    }
    for (size_t i = 0; i < d.size(); i++) {
        const cudaError_t e = cudaMemcpyAsync(d[i], sr[i], sz[i], cudaMemcpyDefault, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int read_u64(const uint64_t *d_src, uint64_t *h_dst, cudaStream_t s) {
    Ctx &c = ctx();
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars, d_src, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    *h_dst = c.h_scalars[0];
    return RSN_OK;
}

// ----------------------------------------------------------------------------- pinned host pool

namespace {
struct PinnedPool {
    std::mutex mu;
    std::unordered_map<void *, size_t> live;            // ptr -> capacity
    std::unordered_map<size_t, std::vector<void *>> free_;  // capacity -> buffers
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCached = (size_t)8 << 30;

    static size_t size_class(size_t n) {
        size_t c = 4096;
        while (c < n) c <<= 1;
        return c;
    }
    void *get(size_t n) {
        size_t c = size_class(n ? n : 1);
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_.find(c);
            if (it != free_.end() && !it->second.empty()) {
                void *p = it->second.back();
                it->second.pop_back();
                cached_bytes -= c;
                live[p] = c;
                return p;
            }
        }
        void *p = nullptr;
        AllocTimer at("pinned", c);
        if (cudaHostAlloc(&p, c, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        std::lock_guard<std::mutex> g(mu);
        live[p] = c;
        return p;
    }
    // Parts of one block handed out as separate results (the outputs of a batch group arrive with
    // one device-to-host copy): every part is freed on its own, the block goes back to the pool with
    // the last of them.
    std::unordered_map<void *, void *> part_of;    // part -> block
    std::unordered_map<void *, size_t> parts_left;  // block -> parts not yet freed
    void adopt_parts(void *block, void *const *parts, size_t count) {
        std::lock_guard<std::mutex> g(mu);
        size_t k = 0;
        for (size_t i = 0; i < count; i++)
            if (parts[i]) {
                part_of[parts[i]] = block;
                k++;
            }
        parts_left[block] = k;
    }
    void put(void *p) {
        if (!p) return;
        size_t c = 0;
        {
            std::lock_guard<std::mutex> g(mu);
            auto pt = part_of.find(p);
            if (pt != part_of.end()) {
                void *block = pt->second;
                part_of.erase(pt);
                auto left = parts_left.find(block);
                if (left != parts_left.end() && --left->second > 0) return;
                if (left != parts_left.end()) parts_left.erase(left);
                p = block;  // the last part: the block itself is released below
            }
            auto it = live.find(p);
            if (it == live.end()) return;  // not ours
            c = it->second;
            live.erase(it);
            if (cached_bytes + c <= kMaxCached) {
                free_[c].push_back(p);
                cached_bytes += c;
                return;
            }
        }
        cudaFreeHost(p);
    }
    void drain() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : free_)
            for (void *p : kv.second) cudaFreeHost(p);
        free_.clear();
        cached_bytes = 0;
    }
};
PinnedPool &pinned() {
    static PinnedPool *p = new PinnedPool();  // leaked on purpose: safe at process exit
    return *p;
}
}  // namespace

void *host_out_alloc(size_t n) { return pinned().get(n); }
void host_out_adopt_parts(void *block, void *const *parts, size_t count) { pinned().adopt_parts(block, parts, count); }

}  // namespace rsn

// ----------------------------------------------------------------------------- C ABI: lifetime

extern "C" {

int rsn_init(int device) { return rsn::init_ctx(device); }

void rsn_shutdown(void) {
    // the calling thread's own streams, events and arena ...
    rsn::release_thread_resources();
    // ... and the process-wide caches of free pinned and device result buffers (buffers still held by
    // a caller stay valid and are freed when they come back)
    rsn::pinned().drain();
    rsn::outs().drain();
    rsn::arena_drain();
}

const char *rsn_strerror(int rc) {
    switch (rc) {
        case RSN_OK: return "ok";
        case RSN_ERR_CUDA: return "CUDA runtime error";
        case RSN_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case RSN_ERR_NOMEM: return "out of memory";
        case RSN_ERR_INVALID_ARG: return "invalid argument";
        case RSN_ERR_UNSUPPORTED: return "input outside this build's supported limits";
        case RSN_ERR_EMPTY_INPUT: return "huffman: empty input (reference panics in heap.Pop)";
        case RSN_ERR_NO_SEPARATOR: return "huffman: header/payload separator missing (reference: index out of range)";
        case RSN_ERR_BAD_HEADER: return "huffman: malformed header (reference: index out of range)";
        case RSN_ERR_TRUNCATED: return "huffman: bit stream ends inside a code (reference: index out of range)";
        case RSN_ERR_GUARD: return "huffman: Max recursion depth";
        case RSN_ERR_BAD_REFERENCE: return "lzss: back-reference outside the decoded data (reference: slice bounds out of range)";
        case RSN_ERR_SINGLE_LEAF_LOOP: return "huffman: single-symbol tree with payload bits (reference: unbounded recursion)";
        default: return "unknown error";
    }
}

const char *rsn_last_cuda_error(void) { return rsn::ctx().cuda_err; }

void rsn_free(void *p) { rsn::pinned().put(p); }

void *rsn_host_alloc(size_t n) { return rsn::pinned().get(n); }
void rsn_host_free(void *p) { rsn::pinned().put(p); }

void rsn_dev_free(void *d_ptr, void *stream) {
    if (!d_ptr) return;
    cudaStream_t s = stream ? (cudaStream_t)stream : rsn::ctx().own_stream;
    rsn::out_free(d_ptr, s);
}

void rsn_free_many(void *const *ptrs, size_t count) {
    for (size_t i = 0; i < count; i++) rsn::pinned().put(ptrs[i]);
}
void rsn_dev_free_many(void *const *d_ptrs, size_t count, void *stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : rsn::ctx().own_stream;
    // a stream with nothing pending orders nothing: the buffers go back without an event each
    if (count > 1 && cudaStreamQuery(s) == cudaSuccess) s = nullptr;
    else cudaGetLastError();
    for (size_t i = 0; i < count; i++)
        if (d_ptrs[i]) rsn::out_free(d_ptrs[i], s);
}

void rsn_kernel_timing(int enable) {
    rsn::KTimer &t = rsn::ktimer();
    t.on.store(false);
    cudaDeviceSynchronize();
    {
        std::lock_guard<std::mutex> g(t.mu);
        for (auto &r : t.recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        t.recs.clear();
    }
    t.on.store(enable != 0);
}

// one line per kernel name: "<name> <launches> <total ms>\n", most expensive first
size_t rsn_kernel_timing_report(char *buf, size_t cap) {
    rsn::KTimer &t = rsn::ktimer();
    cudaDeviceSynchronize();
    if (const char *path = getenv("RSN_KTIME_TIMELINE")) {  // every launch: name, stream, start and end (ms since the first)
        std::lock_guard<std::mutex> g(t.mu);
        if (FILE *f = fopen(path, "w")) {
            for (auto &r : t.recs) {
                float a = 0.f, b = 0.f;
                if (cudaEventElapsedTime(&a, t.recs[0].a, r.a) != cudaSuccess ||
                    cudaEventElapsedTime(&b, t.recs[0].a, r.b) != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                fprintf(f, "%s %p %.4f %.4f\n", r.name, (void *)r.s, a, b);
            }
            fclose(f);
        }
    }
    std::vector<std::pair<std::string, std::pair<uint64_t, double>>> rows;
    {
        std::lock_guard<std::mutex> g(t.mu);
        std::unordered_map<std::string, std::pair<uint64_t, double>> agg;
        for (auto &r : t.recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            auto &e = agg[r.name];
            e.first++;
            e.second += ms;
        }
        rows.assign(agg.begin(), agg.end());
    }
    std::sort(rows.begin(), rows.end(), [](const auto &x, const auto &y) { return x.second.second > y.second.second; });
    std::string out;
    for (auto &r : rows) {
        char line[256];
        snprintf(line, sizeof(line), "%s %llu %.6f\n", r.first.c_str(), (unsigned long long)r.second.first, r.second.second);
        out += line;
    }
    if (buf && cap) {
        const size_t k = std::min(cap - 1, out.size());
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return out.size();
}

uint64_t rsn_kernel_launches(void) { return rsn::g_launches.load(); }
void rsn_reset_kernel_launches(void) { rsn::g_launches.store(0); }
const char *rsn_version(void) { return "raisin_b200 0.1 sm_100a"; }

}  // extern "C"
