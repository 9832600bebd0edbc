// batch.cuh — batches of independent small files (engine.BenchmarkSuite's per-file loop,
// engine.go:208-262, turned inside out): instead of running the whole pipeline file by file, every
// kernel of a stage runs once over a group of files (the file index is blockIdx.y, or .z where .y is
// taken), sizes that depend on the data stay on the device between the kernels of a stage, and the
// host synchronises once or twice per stage and group instead of several times per file.
#pragma once
#include "common.cuh"

#include <atomic>
#include <thread>
#include <utility>
#include <vector>

namespace rsn {

// G device-resident byte strings: the input or the output of one stage.
struct BatchIO {
    std::vector<const uint8_t *> ptr;  // device pointers, 16-byte aligned
    std::vector<uint64_t> n;
    std::vector<int> rc;               // RSN_OK, or the error that took the file out of the batch
    std::vector<void *> owned;         // result buffers (out_alloc) backing ptr[]
    // device ranges, each inside ONE allocation, in which the stage laid its results out back to back:
    // what lets the caller fetch a group's results with one copy per range instead of one per file
    std::vector<std::pair<const uint8_t *, size_t>> spans;
    size_t size() const { return ptr.size(); }
    void resize(size_t g) {
        ptr.assign(g, nullptr);
        n.assign(g, 0);
        rc.assign(g, RSN_OK);
    }
    void release(cudaStream_t s) {
        for (void *p : owned) out_free(p, s);
        owned.clear();
        spans.clear();
    }
};

// Largest file the batched kernels take (per-file scans and hierarchies are single-CTA / two-level).
constexpr size_t kBatchMaxFile = (size_t)4 << 20;

int lzss_compress_batch(const BatchIO &in, BatchIO &out, int64_t window, cudaStream_t s);
int lzss_decompress_batch(const BatchIO &in, BatchIO &out, cudaStream_t s);
int huff_compress_batch(const BatchIO &in, BatchIO &out, cudaStream_t s);
// h_in[f]: host copy of file f's stream (the header is parsed on the host); with `h_prefix_only` it
// holds the stream up to and including the byte after the first 5C 0A (or all of it, if there is none)
int huff_decompress_batch(const BatchIO &in, const uint8_t *const *h_in, BatchIO &out, cudaStream_t s,
                          bool h_prefix_only = false);

// host threads one group may use for its per-file host work (set per batch call from the core count
// and the number of group workers)
int batch_host_threads();
void set_batch_host_threads(int t);

// fn(i) for i in [0, count) on up to `threads` host threads (tree building, header parsing)
template <class F>
void parallel_for(size_t count, int threads, F fn);

#ifdef __CUDACC__
// Exclusive scan of `count` u64 values by one CTA; every thread gets the total.  sm: 33 u64.
__device__ __forceinline__ uint64_t cta_scan_u64(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
                                                 size_t count, uint64_t *sm) {
    uint64_t carry = 0;
    for (size_t base = 0; base < count; base += blockDim.x) {
        const size_t i = base + threadIdx.x;
        const uint64_t v = i < count ? in[i] : 0;
        uint64_t total;
        const uint64_t pre = block_exclusive_sum<uint64_t>(v, sm, total);
        if (i < count) out[i] = carry + pre;
        carry += total;
    }
    return carry;
}
#endif

template <class F>
void parallel_for(size_t count, int threads, F fn) {
    if (threads > (int)count) threads = (int)count;
    if (threads <= 1) {
        for (size_t i = 0; i < count; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    auto body = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= count) break;
            fn(i);
        }
    };
    for (int t = 1; t < threads; t++) pool.emplace_back(body);
    body();
    for (auto &th : pool) th.join();
}

}  // namespace rsn
