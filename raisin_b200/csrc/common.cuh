// common.cuh — context, error plumbing and block-level primitives shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include "../../include/raisin_b200.h"

namespace rsn {

// ----------------------------------------------------------------------------- context

struct Ctx {
    int device = -1;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host->device chunks that overlap the first kernels
    // Batches: the match search of a group (5 ms with every SM taken) runs on a stream of the lowest
    // priority and everything else on the thread's own stream, which has the highest.  With one
    // priority for all, the short kernels of a group that is past its match search queue behind the
    // match searches of the other groups; the eight groups of a pass then move in lockstep, reach the
    // host part of the Huffman stage together and leave the GPU idle for 8 ms and nearly idle for the
    // 8 ms of their tree kernels, twice per pass.
    cudaStream_t low_stream = nullptr;
    cudaEvent_t low_before = nullptr, low_after = nullptr;
    // The stream a batch worker runs its groups on.  Workers get DIFFERENT priorities (all above
    // low_stream's): with equal priorities the GPU shares itself out evenly, the groups of a pass
    // advance in lockstep and so do their copies — every worker uploads at the start and downloads at
    // the end, with nothing running beside the copies.  Staggered, the first worker's group finishes
    // first and its download runs under the other groups' kernels.
    cudaStream_t batch_stream = nullptr;
    cudaEvent_t chunk_ev[64] = {nullptr};
    bool ready = false;
    uint64_t launches = 0;
    char cuda_err[256] = {0};
    // small pinned scratch for scalar read-backs (sizes, flags)
    uint64_t *h_scalars = nullptr;  // 64 x u64
};

Ctx &ctx();
int ensure_ctx();
void count_launch();  // this thread's counter and the process-wide one behind rsn_kernel_launches()
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define RSN_CUDA(expr)                                                       \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return ::rsn::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define RSN_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != RSN_OK) return _rc; \
    } while (0)

// Per-kernel device timing (rsn_kernel_timing): when enabled, every launch is bracketed by two CUDA
// events on its own stream; rsn_kernel_timing_report() sums the durations per kernel name.
int ktime_begin(const char *name, cudaStream_t s);  // -1 when timing is off
void ktime_end(int idx, cudaStream_t s);

// kernel launch with launch counting, optional event timing and error check
#define RSN_LAUNCH(kernel, grid, block, smem, stream, ...)                   \
    do {                                                                     \
        const int _kt = ::rsn::ktime_begin(#kernel, (stream));               \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
        if (_kt >= 0) ::rsn::ktime_end(_kt, (stream));                       \
        ::rsn::count_launch();                                               \
        RSN_CUDA(cudaGetLastError());                                        \
    } while (0)

// Device memory.  Temporaries of a call are bump-allocated from a per-thread arena that is
// rewound when the call's ArenaScope ends (no driver call on the hot path); results handed back
// to the caller come from a size-class cache of cudaMalloc'd buffers (rsn_dev_free returns them).
// (cudaMallocAsync pools showed multi-millisecond, occasionally >100 ms, stalls here.)
struct ArenaScope {
    cudaStream_t s;
    size_t saved_block, saved_off;
    explicit ArenaScope(cudaStream_t stream);
    ~ArenaScope();
    ArenaScope(const ArenaScope &) = delete;
    ArenaScope &operator=(const ArenaScope &) = delete;
};
void *arena_alloc(size_t n);
void *out_alloc(size_t n, cudaStream_t s);
void out_free(void *p, cudaStream_t s);

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    bool is_out = false;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { reset(); }
    int alloc(size_t n, cudaStream_t stream);      // arena temporary (valid until the ArenaScope ends)
    int alloc_out(size_t n, cudaStream_t stream);  // caller-visible result
    void reset();
    void *release() {
        void *q = p;
        p = nullptr;
        return q;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// Wait for a stream.  Batch worker threads sleep on an event (cudaEventBlockingSync) instead of
// spinning: a box runs up to 8 ranks x 8 workers, nearly all of them waiting at any moment, and
// spinning waits starve the threads that have host work to do.  Other threads spin (lowest latency).
cudaError_t stream_wait(cudaStream_t s);
void set_thread_blocking_sync(bool on);

// Pinned host memory from the library's pool.  Every host buffer that a cudaMemcpyAsync touches
// must be pinned: an "async" copy to or from pageable memory waits inside the driver until the
// stream has reached it — with the context lock held, so a device-to-host copy queued behind a 4 ms
// kernel stops every other host thread from launching anything for those 4 ms (that is how the
// batch path ran its groups' Huffman tree kernels strictly one after another).
// `count` independent copies (host <-> device, any mix) queued with ONE driver call
// (cudaMemcpyBatchAsync): 2048 copies of 256 KiB from separately pinned buffers cost 9.3 ms of host
// time and run at 29 GB/s as cudaMemcpyAsync calls (less from several threads at once: the calls
// serialise on the driver), 0.7 ms and 52 GB/s as one batch.  Entries of size 0 are skipped; falls
// back to a loop of cudaMemcpyAsync where the batch call is not available.
cudaError_t copy_many(void *const *dsts, const void *const *srcs, const size_t *sizes, size_t count, cudaStream_t s);
void *host_out_alloc(size_t n);
// `parts` (null entries skipped) point into `block` (from host_out_alloc): from now on each part is
// released with rsn_free on its own and the block follows the last one; at least one part is required
void host_out_adopt_parts(void *block, void *const *parts, size_t count);
template <typename T>
struct HostVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    HostVec() = default;
    explicit HostVec(size_t k) { resize(k); }
    HostVec(const HostVec &) = delete;
    HostVec &operator=(const HostVec &) = delete;
    ~HostVec() {
        if (p) rsn_free(p);
    }
    bool resize(size_t k) {  // new elements are zeroed
        if (k > cap) {
            T *q = static_cast<T *>(host_out_alloc((k ? k : 1) * sizeof(T)));
            if (!q) return false;
            if (p) {
                memcpy(q, p, n * sizeof(T));
                rsn_free(p);
            }
            p = q;
            cap = k;
        }
        if (k > n) memset(static_cast<void *>(p + n), 0, (k - n) * sizeof(T));
        n = k;
        return true;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// Stage tracer: with RSN_TRACE=1 in the environment, prints host wall time between marks
// (each mark synchronises the stream first, so it is for diagnosis only).
struct Trace {
    cudaStream_t s;
    bool on;
    double t0;
    const char *what;
    Trace(const char *w, cudaStream_t stream);
    void mark(const char *label);
};

// ----------------------------------------------------------------------------- geometry

constexpr int kTileThreads = 256;      // threads per CTA for byte-stream kernels
constexpr int kItems = 16;             // bytes per thread (one uint4 load)
constexpr int kTile = kTileThreads * kItems;  // 4096 bytes per CTA tile

static inline size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- device helpers

#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

// Load 16 consecutive bytes starting at base+idx (idx multiple of 16 when base is 16B-aligned);
// bytes beyond n read as `fill`.  Falls back to byte loads at the ragged end / unaligned base.
__device__ __forceinline__ void load16(const uint8_t *__restrict__ base, size_t idx, size_t n, uint8_t fill,
                                       uint8_t (&v)[16]) {
    if (idx + 16 <= n && ((reinterpret_cast<uintptr_t>(base + idx) & 15) == 0)) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(base + idx));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = (idx + k < n) ? __ldg(base + idx + k) : fill;
    }
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v += o;
    }
    return v;
}

// Exclusive sum over the CTA (blockDim.x multiple of 32, <= 1024).  Returns the exclusive
// prefix of `v` and writes the CTA total to `total`.  `smem` must hold 33 T's.
template <typename T>
__device__ __forceinline__ T block_exclusive_sum(T v, T *smem, T &total) {
    const unsigned lane = lane_id(), wid = warp_id();
    const unsigned nw = (blockDim.x + 31) >> 5;
    T inc = warp_inclusive_sum(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < nw ? smem[lane] : T(0);
        T winc = warp_inclusive_sum(w);
        smem[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[wid] + inc - v;
    total = smem[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ size_t div_up_dev(size_t a, size_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ int ndig_u32(uint32_t v) {
    return 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
           (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}

#endif  // __CUDACC__

// ----------------------------------------------------------------------------- shared kernels (scan.cu)

// Exclusive scan of per-tile u64 values by one CTA; out[t] = sum(in[0..t)), *total = sum of all.
int spine_scan_u64(const uint64_t *d_in, uint64_t *d_out, uint64_t *d_total, size_t count, cudaStream_t s);
// Read one u64 back to the host (synchronises the stream).
int read_u64(const uint64_t *d_src, uint64_t *h_dst, cudaStream_t s);

}  // namespace rsn
