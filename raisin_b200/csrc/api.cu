// api.cu — the C ABI (include/raisin_b200.h): host-buffer entry points a cgo shim binds,
// device-buffer entry points, and the engine's layer loops.
#include "batch.cuh"
#include "common.cuh"
#include "huff.cuh"
#include "lzss.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <string>
#include <vector>

namespace rsn {

void *host_out_alloc(size_t n);

static cudaStream_t pick_stream(void *stream) { return stream ? (cudaStream_t)stream : ctx().own_stream; }

// device result -> library-owned pinned host buffer
static int to_host(uint8_t *d, size_t n, uint8_t **out, size_t *out_n, cudaStream_t s) {
    uint8_t *h = (uint8_t *)host_out_alloc(n ? n : 1);
    if (!h) {
        out_free(d, s);
        return RSN_ERR_NOMEM;
    }
    cudaError_t e = cudaSuccess;
    if (n) e = cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = stream_wait(s);
    out_free(d, s);
    if (e != cudaSuccess) {
        rsn_free(h);
        return cuda_fail(e, "to_host", __FILE__, __LINE__);
    }
    *out = h;
    *out_n = n;
    return RSN_OK;
}

static int to_device(const uint8_t *in, size_t n, DevBuf &d, cudaStream_t s) {
    RSN_TRY(d.alloc(n + 64, s));
    if (n) RSN_CUDA(cudaMemcpyAsync(d.p, in, n, cudaMemcpyHostToDevice, s));
    return RSN_OK;
}

enum Algo { ALGO_LZSS, ALGO_HUFFMAN };

static int parse_layers(const char *algorithms, std::vector<Algo> &out) {
    if (!algorithms) return RSN_ERR_INVALID_ARG;
    std::string s(algorithms);
    size_t a = 0;
    while (a <= s.size()) {
        size_t b = s.find(',', a);
        if (b == std::string::npos) b = s.size();
        std::string name = s.substr(a, b - a);
        if (name == "lzss") out.push_back(ALGO_LZSS);
        else if (name == "huffman") out.push_back(ALGO_HUFFMAN);
        else return RSN_ERR_INVALID_ARG;
        a = b + 1;
    }
    return out.empty() ? RSN_ERR_INVALID_ARG : RSN_OK;
}

// engine.compress (engine.go:443-452): each algorithm consumes the previous one's output;
// lz.NewWriter uses CompressAsync with DefaultWindowSize 4096 (lzss.go:37-40, 53-57).
static int layers_dev(const std::vector<Algo> &algos, bool compress, const uint8_t *d_in, size_t n, uint8_t **d_out,
                      size_t *out_n, cudaStream_t s) {
    const uint8_t *cur = d_in;
    size_t cur_n = n;
    uint8_t *owned = nullptr;
    const size_t k = algos.size();
    for (size_t step = 0; step < k; step++) {
        const Algo a = compress ? algos[step] : algos[k - 1 - step];
        uint8_t *next = nullptr;
        size_t next_n = 0;
        int rc;
        if (a == ALGO_LZSS)
            rc = compress ? lzss_compress_dev(cur, cur_n, 4096, RSN_LZSS_ASYNC, &next, &next_n, s)
                          : lzss_decompress_dev(cur, cur_n, &next, &next_n, s);
        else
            rc = compress ? huff_compress_dev(cur, cur_n, &next, &next_n, s)
                          : huff_decompress_dev(cur, cur_n, nullptr, 0, &next, &next_n, s);
        if (owned) out_free(owned, s);
        owned = nullptr;
        if (rc != RSN_OK) return rc;
        owned = next;
        cur = next;
        cur_n = next_n;
    }
    *d_out = owned;
    *out_n = cur_n;
    return RSN_OK;
}

}  // namespace rsn

using namespace rsn;

extern "C" {

// Large inputs: the host->device copy is cut into chunks on a second stream and the match search
// of a chunk starts as soon as the chunk after it has landed (it needs a window of look-ahead).
// The search runs on the raw bytes, which is only right when nothing needs escaping; that is known
// once the whole input is on the device, and otherwise the speculative arrays are dropped.
static int lzss_compress_overlapped(const uint8_t *in, size_t n, int64_t window, int variant, uint8_t **r, size_t *rn,
                                    cudaStream_t s) {
    Ctx &c = ctx();
    const size_t T = lzss_match_tile_size();
    const size_t chunk = (size_t)8 << 20;  // multiple of the tile size
    const size_t chunks = div_up(n, chunk);
    uint32_t W = 0;
    RSN_TRY(lzss_effective_window(window, n, &W));
    DevBuf d, packed;
    RSN_TRY(d.alloc(n + 64, s));
    RSN_TRY(packed.alloc(div_up(n, 4096) * 4096 * 4 + 64, s));
    // the arena memory may still be read by work queued on s: order the copies after it
    RSN_CUDA(cudaEventRecord(c.chunk_ev[63], s));
    RSN_CUDA(cudaStreamWaitEvent(c.copy_stream, c.chunk_ev[63], 0));
    auto copy_chunk = [&](size_t k) -> int {
        const size_t lo = k * chunk, len = std::min(chunk, n - lo);
        RSN_CUDA(cudaMemcpyAsync(d.as<uint8_t>() + lo, in + lo, len, cudaMemcpyHostToDevice, c.copy_stream));
        RSN_CUDA(cudaEventRecord(c.chunk_ev[k], c.copy_stream));
        return RSN_OK;
    };
    const size_t tiles = div_up(n, T), tiles_per_chunk = chunk / T;
    RSN_TRY(copy_chunk(0));
    for (size_t k = 0; k < chunks; k++) {
        if (k + 1 < chunks) RSN_TRY(copy_chunk(k + 1));  // interleaved so pageable sources overlap too
        // tiles of chunk k need bytes up to their end + W: wait for chunk k+1 (or the last one)
        RSN_CUDA(cudaStreamWaitEvent(s, c.chunk_ev[std::min(k + 1, chunks - 1)], 0));
        const size_t t_lo = k * tiles_per_chunk, t_hi = std::min(tiles, t_lo + tiles_per_chunk);
        RSN_TRY(lzss_match_tile_range(d.as<uint8_t>(), n, W, packed.as<uint32_t>(), t_lo, t_hi, k + 1 == chunks, s));
    }
    return lzss_compress_dev_ex(d.as<uint8_t>(), n, window, variant, packed.as<uint32_t>(), r, rn, s);
}

int rsn_lzss_compress(const uint8_t *in, size_t n, int64_t window, int variant, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    uint8_t *r = nullptr;
    size_t rn = 0;
    const bool tile_window = window > 0 && window <= 4096;
    if (tile_window && n >= ((size_t)16 << 20) && n <= ((size_t)8 << 20) * 60) {
        RSN_TRY(lzss_compress_overlapped(in, n, window, variant, &r, &rn, s));
        return to_host(r, rn, out, out_n, s);
    }
    DevBuf d;
    RSN_TRY(to_device(in, n, d, s));
    RSN_TRY(lzss_compress_dev(d.as<uint8_t>(), n, window, variant, &r, &rn, s));
    return to_host(r, rn, out, out_n, s);
}

// BASELINE configs[4]: lz.CompressAsync of one large stream with the match search sharded by
// position range over `ngpus` shards (shard g on device g mod the device count).  The iterative
// variant has no sharded form (its start-of-match test scans the whole history): it runs on one GPU.
int rsn_lzss_compress_sharded(const uint8_t *in, size_t n, int64_t window, int variant, int ngpus, uint8_t **out,
                              size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    if (variant != RSN_LZSS_ASYNC && variant != RSN_LZSS_ITER) return RSN_ERR_INVALID_ARG;
    if (variant == RSN_LZSS_ITER || ngpus <= 1) return rsn_lzss_compress(in, n, window, variant, out, out_n);
    RSN_TRY(ensure_ctx());
    return lzss_compress_sharded(in, n, window, ngpus, out, out_n);
}
uint64_t rsn_sharded_peer_bytes(void) { return lzss_sharded_last_peer_bytes(); }

int rsn_lzss_decompress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    DevBuf d;
    RSN_TRY(to_device(in, n, d, s));
    uint8_t *r = nullptr;
    size_t rn = 0;
    RSN_TRY(lzss_decompress_dev(d.as<uint8_t>(), n, &r, &rn, s));
    return to_host(r, rn, out, out_n, s);
}

int rsn_huff_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    DevBuf d;
    RSN_TRY(to_device(in, n, d, s));
    uint8_t *r = nullptr;
    size_t rn = 0;
    RSN_TRY(huff_compress_dev(d.as<uint8_t>(), n, &r, &rn, s));
    return to_host(r, rn, out, out_n, s);
}

int rsn_huff_decompress(const uint8_t *in, size_t n, int strict_limits, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    DevBuf d;
    RSN_TRY(to_device(in, n, d, s));
    uint8_t *r = nullptr;
    size_t rn = 0;
    RSN_TRY(huff_decompress_dev(d.as<uint8_t>(), n, in, strict_limits, &r, &rn, s));
    return to_host(r, rn, out, out_n, s);
}

static int layers_host(const char *algorithms, bool compress, const uint8_t *in, size_t n, uint8_t **out,
                       size_t *out_n) {
    if ((!in && n) || !out || !out_n) return RSN_ERR_INVALID_ARG;
    std::vector<Algo> algos;
    RSN_TRY(parse_layers(algorithms, algos));
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    DevBuf d;
    RSN_TRY(to_device(in, n, d, s));
    uint8_t *r = nullptr;
    size_t rn = 0;
    RSN_TRY(layers_dev(algos, compress, d.as<uint8_t>(), n, &r, &rn, s));
    return to_host(r, rn, out, out_n, s);
}

}  // extern "C"

namespace rsn {
namespace {
// engine.BenchmarkFile's side computations (engine.go:367-377, 408, 412-423) on the device: byte
// histogram (256 bins per CTA in shared memory, 16-byte loads, one flush per CTA) and the
// DeepEqual of input and round trip.
__global__ void __launch_bounds__(256) k_byte_hist(const uint8_t *__restrict__ in, size_t n,
                                                   unsigned long long *__restrict__ hist) {
    __shared__ uint32_t bins[256];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const size_t per = (size_t)64 << 10;  // bytes per CTA
    const size_t lo = (size_t)blockIdx.x * per, hi = min(n, lo + per);
    for (size_t i = lo + (size_t)threadIdx.x * 16; i < hi; i += 256 * 16) {
        uint8_t v[16];
        load16(in, i, hi, 0, v);
        const int valid = (int)min((size_t)16, hi - i);
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k < valid) atomicAdd(&bins[v[k]], 1u);
    }
    __syncthreads();
    if (bins[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)bins[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_bytes_differ(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                      size_t n, uint32_t *__restrict__ differ) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    uint8_t x[16], y[16];
    load16(a, i, n, 0, x);
    load16(b, i, n, 0, y);
    bool d = false;
#pragma unroll
    for (int k = 0; k < 16; k++) d |= x[k] != y[k];
    if (d) *differ = 1;
}
// natural-log Shannon entropy of counts / total (goent.Entropy with math.Log)
double entropy_nat(const unsigned long long *hist, double total) {
    double e = 0;
    for (int b = 0; b < 256; b++)
        if (hist[b]) {
            const double p = (double)hist[b] / total;
            e -= p * log(p);
        }
    return e;
}
}  // namespace
}  // namespace rsn

extern "C" {

// engine.BenchmarkFile (engine/engine.go:357-441) in one call: entropy of the input, compress and
// decompress through the layer list (the timed region, wall clock, host buffer in), lossless flag,
// ratio, and the "actual entropy" exactly as the reference computes it — the histogram of the
// DEcompressed bytes divided by the COMPRESSED length (engine.go:412-418; the README's tables carry
// those values).  A codec failure (the reference would panic and AsyncBenchmarkFile would report
// Failed, engine.go:315-328) sets res->failed and returns RSN_OK.
int rsn_benchmark_file(const char *algorithms, const uint8_t *in, size_t n, rsn_bench_result *res) {
    if ((!in && n) || !res) return RSN_ERR_INVALID_ARG;
    std::vector<Algo> algos;
    RSN_TRY(parse_layers(algorithms, algos));
    RSN_TRY(ensure_ctx());
    cudaStream_t s = ctx().own_stream;
    ArenaScope scope(s);
    memset(res, 0, sizeof(*res));
    timespec t0, t1;
    DevBuf d, dh;
    RSN_TRY(dh.alloc(2 * 256 * 8 + 16, s));
    unsigned long long *h_in = dh.as<unsigned long long>(), *h_back = h_in + 256;
    uint32_t *differ = reinterpret_cast<uint32_t *>(h_back + 256);
    RSN_CUDA(cudaMemsetAsync(dh.p, 0, 2 * 256 * 8 + 16, s));
    clock_gettime(CLOCK_MONOTONIC, &t0);  // start := time.Now() comes after the histogram in the reference;
    RSN_TRY(to_device(in, n, d, s));      // the upload is this implementation's cost and is timed
    if (n) RSN_LAUNCH(k_byte_hist, (unsigned)div_up(n, (size_t)64 << 10), 256, 0, s, d.as<uint8_t>(), n, h_in);
    uint8_t *comp = nullptr, *back = nullptr;
    size_t comp_n = 0, back_n = 0;
    int rc = layers_dev(algos, true, d.as<uint8_t>(), n, &comp, &comp_n, s);
    if (rc == RSN_OK) rc = layers_dev(algos, false, comp, comp_n, &back, &back_n, s);
    if (rc == RSN_OK) RSN_CUDA(stream_wait(s));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    HostVec<unsigned long long> hh(2 * 256 + 2);
    if (!hh.data()) rc = rc == RSN_OK ? RSN_ERR_NOMEM : rc;
    if (rc == RSN_OK) {
        if (back_n) RSN_LAUNCH(k_byte_hist, (unsigned)div_up(back_n, (size_t)64 << 10), 256, 0, s, back, back_n, h_back);
        if (back_n == n && n)
            RSN_LAUNCH(k_bytes_differ, (unsigned)div_up(div_up(n, 16), 256), 256, 0, s, d.as<uint8_t>(), back, n, differ);
        RSN_CUDA(cudaMemcpyAsync(hh.data(), dh.p, 2 * 256 * 8 + 16, cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
    }
    if (comp) out_free(comp, s);
    if (back) out_free(back, s);
    if (rc == RSN_ERR_CUDA || rc == RSN_ERR_NO_DEVICE || rc == RSN_ERR_NOMEM) return rc;
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (rc != RSN_OK) {
        res->failed = 1;
        res->error = rc;
        return RSN_OK;
    }
    res->compressed_n = comp_n;
    res->decompressed_n = back_n;
    res->entropy = n ? entropy_nat(hh.data(), (double)n) : 0.0;
    res->actual_entropy = (float)(comp_n ? entropy_nat(hh.data() + 256, (double)comp_n) : 0.0);
    res->ratio = (float)comp_n / (float)n * 100.0f;
    res->lossless = back_n == n && (uint32_t)hh[512] == 0;
    return RSN_OK;
}

int rsn_compress_layers(const char *algorithms, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    return layers_host(algorithms, true, in, n, out, out_n);
}
int rsn_decompress_layers(const char *algorithms, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    return layers_host(algorithms, false, in, n, out, out_n);
}

// ---- batches of independent files

}  // extern "C"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace rsn {
namespace {
// Persistent worker threads: each keeps its thread-local context (stream, arena) across batches.
class WorkerPool {
  public:
    // One batch at a time: job_/pending_/running_ describe a single run, and done_.wait releases
    // mu_, so concurrent callers (the header promises re-entrancy) queue here instead of
    // overwriting each other's job.
    void run(int workers, const std::function<void()> &job) {
        std::lock_guard<std::mutex> serial(run_mu_);
        std::unique_lock<std::mutex> lk(mu_);
        while ((int)threads_.size() < workers) threads_.emplace_back([this] { loop(); });
        job_ = job;
        pending_ = workers;
        running_ = workers;
        generation_++;
        cv_.notify_all();
        done_.wait(lk, [this] { return running_ == 0; });
    }

  private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen && pending_ > 0; });
                seen = generation_;
                pending_--;
                job = job_;
            }
            job();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--running_ == 0) done_.notify_all();
            }
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    std::function<void()> job_;
    int pending_ = 0, running_ = 0;
    uint64_t generation_ = 0;
};
WorkerPool &pool() {
    static WorkerPool *p = new WorkerPool();  // threads live for the process
    return *p;
}
}  // namespace
}  // namespace rsn

extern "C" {

}  // extern "C"

namespace rsn {
// Which files of a batch travel together.  groups[g] lists the files of group g (every kernel of a
// stage runs once per group), `singles` the files that go through the per-file path: empty files,
// files above kBatchMaxFile, null inputs, and device-resident files that are not 16-byte aligned.
// The kernels of a group size their per-file arrays for the group's LARGEST file, so a group only
// holds files of one size class (within a factor of two; everything below 4 KiB is one class):
// memory stays proportional to the bytes in the group whatever the mix of sizes.  `in` may be null
// (planning only).
void batch_plan(size_t count, const uint8_t *const *in, const size_t *in_n, int device,
                std::vector<std::vector<size_t>> &groups, std::vector<size_t> &singles, int workers = 8) {
    // Up to 64 MiB per group: the per-group costs that do not shrink with the group (the serial heap
    // replay of the Huffman tree kernel, 2-4 ms for a file of random bytes, and a handful of host
    // synchronisations per stage) are paid 16 times per GiB instead of 64 (2048 config-4 files:
    // 3.6 GB/s with 16 MiB groups, 4.4 GB/s with 64 MiB, 8 workers).  Small batches are cut finer so
    // that every worker still gets a group (a rank's share of a batch split over 8 GPUs), but not
    // below 16 MiB (512 files: 38 ms per pass with 8 MiB groups, 34 ms with 16 MiB).
    size_t total_bytes = 0;
    for (size_t i = 0; i < count; i++)
        if (in_n[i] <= kBatchMaxFile) total_bytes += in_n[i];
    size_t group_bytes = total_bytes / (size_t)(workers > 0 ? workers : 1);
    group_bytes = std::min<size_t>((size_t)64 << 20, std::max<size_t>((size_t)16 << 20, group_bytes));
    size_t kGroupFiles = 2048;
    if (const char *e = getenv("RSN_BATCH_GROUP_MIB")) {  // tuning knobs
        const long v = atol(e);
        if (v >= 1 && v <= 4096) group_bytes = (size_t)v << 20;
    }
    if (const char *e = getenv("RSN_BATCH_GROUP_FILES")) {
        const long v = atol(e);
        if (v >= 1 && v <= 32768) kGroupFiles = (size_t)v;
    }
    struct Open {
        long group = -1;
        size_t bytes = 0;
    };
    Open open[64];
    for (size_t i = 0; i < count; i++) {
        if (in_n[i] == 0 || in_n[i] > kBatchMaxFile || (in && !in[i]) ||
            (device && in && (reinterpret_cast<uintptr_t>(in[i]) & 15))) {  // the grouped kernels load 16 bytes at a time
            singles.push_back(i);
            continue;
        }
        const int cls = in_n[i] < 4096 ? 11 : 63 - __builtin_clzll((unsigned long long)in_n[i]);
        Open &o = open[cls];
        if (o.group < 0 || o.bytes + in_n[i] > group_bytes || groups[(size_t)o.group].size() >= kGroupFiles) {
            groups.emplace_back();
            o.group = (long)groups.size() - 1;
            o.bytes = 0;
        }
        groups[(size_t)o.group].push_back(i);
        o.bytes += in_n[i];
    }
}

static std::atomic<int> g_batch_host_threads{4};
int batch_host_threads() { return g_batch_host_threads.load(); }
void set_batch_host_threads(int t) { g_batch_host_threads.store(t < 1 ? 1 : t); }

namespace {

// One stage over a group, file by file (stages without a batched implementation, or groups the
// batched one declines).
int stage_per_file(Algo a, bool compress, const BatchIO &in, const uint8_t *const *h_in, BatchIO &out, cudaStream_t s,
                   bool h_prefix_only = false) {
    if (h_prefix_only) h_in = nullptr;  // the single-stream call wants the whole stream or none of it
    const size_t G = in.size();
    out.resize(G);
    out.rc = in.rc;
    for (size_t f = 0; f < G; f++) {
        if (in.rc[f] != RSN_OK) continue;
        uint8_t *next = nullptr;
        size_t next_n = 0;
        int rc;
        if (a == ALGO_LZSS)
            rc = compress ? lzss_compress_dev(in.ptr[f], in.n[f], 4096, RSN_LZSS_ASYNC, &next, &next_n, s)
                          : lzss_decompress_dev(in.ptr[f], in.n[f], &next, &next_n, s);
        else
            rc = compress ? huff_compress_dev(in.ptr[f], in.n[f], &next, &next_n, s)
                          : huff_decompress_dev(in.ptr[f], in.n[f], h_in ? h_in[f] : nullptr, 0, &next, &next_n, s);
        out.rc[f] = rc;
        if (rc != RSN_OK) continue;
        out.ptr[f] = next;
        out.n[f] = next_n;
        out.owned.push_back(next);
    }
    return RSN_OK;
}

int stage_batched(Algo a, bool compress, const BatchIO &in, const uint8_t *const *h_in, BatchIO &out, cudaStream_t s,
                  bool h_prefix_only = false) {
    if (a == ALGO_LZSS) return compress ? lzss_compress_batch(in, out, 4096, s) : lzss_decompress_batch(in, out, s);
    if (a == ALGO_HUFFMAN)
        return compress ? huff_compress_batch(in, out, s) : huff_decompress_batch(in, h_in, out, s, h_prefix_only);
    return RSN_ERR_UNSUPPORTED;
}

// results of a device-resident group: every file from the group's buffer into its own buffer
struct CopyJob {
    const uint8_t *src;
    uint8_t *dst;
    uint64_t n;
};
__global__ void __launch_bounds__(256) kb_copy_out(const CopyJob *__restrict__ jobs) {
    const CopyJob j = jobs[blockIdx.y];
    const size_t chunk = (size_t)256 * 16 * 4;  // 16 KiB per CTA
    const size_t lo = (size_t)blockIdx.x * chunk;
    if (lo >= j.n) return;
    const size_t hi = min((size_t)j.n, lo + chunk);
    if (((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(j.dst)) & 15) == 0) {
        const size_t v_lo = lo / 16, v_hi = hi / 16;
        const uint4 *sv = reinterpret_cast<const uint4 *>(j.src);
        uint4 *dv = reinterpret_cast<uint4 *>(j.dst);
        for (size_t v = v_lo + threadIdx.x; v < v_hi; v += blockDim.x) dv[v] = __ldg(sv + v);
        for (size_t b = v_hi * 16 + threadIdx.x; b < hi; b += blockDim.x) j.dst[b] = j.src[b];
    } else {
        for (size_t b = lo + threadIdx.x; b < hi; b += blockDim.x) j.dst[b] = j.src[b];
    }
}

// The files idx[0..G) of a batch through every layer.  Host mode: inputs are uploaded into one
// buffer and results land in library-owned pinned host buffers.  Device mode: inputs are used in
// place and every result gets its own device buffer (rsn_dev_free).
// (Tried and dropped: a worker's copies on streams of their own, one group ahead of / behind its
// kernels.  With the copies batched it bought nothing — 249 against 240-244 ms per 4096-file pass — and
// one pass in eight took 670 ms.)
int batch_group(const std::vector<Algo> &algos, bool compress, const std::vector<size_t> &idx,
                const uint8_t *const *in, const size_t *in_n, uint8_t **out, size_t *out_n, int *rcs, bool device,
                cudaStream_t s) {
    const size_t G = idx.size();
    ArenaScope scope(s);
    Trace tr("batch", s);
    BatchIO cur;
    cur.resize(G);
    size_t total = 0;
    for (size_t f = 0; f < G; f++) total += (in_n[idx[f]] + 64 + 255) & ~(size_t)255;
    DevBuf d;
    std::vector<const uint8_t *> h_in(G);
    uint8_t *h_stage = nullptr;  // device mode: host copy of the streams for the Huffman header parser
    struct StageFree {
        uint8_t *&p;
        ~StageFree() {
            if (p) rsn_free(p);
        }
    } stage_free{h_stage};
    const Algo first = compress ? algos.front() : algos.back();
    if (!device) {
        RSN_TRY(d.alloc(total + 256, s));
    } else if (!compress && first == ALGO_HUFFMAN) {
        h_stage = (uint8_t *)host_out_alloc(total + 256);
        if (!h_stage) return RSN_ERR_NOMEM;
    }
    size_t off = 0;
    {
        std::vector<void *> dsts;
        std::vector<const void *> srcs;
        std::vector<size_t> sizes;
        for (size_t f = 0; f < G; f++) {
            const size_t i = idx[f];
            cur.n[f] = in_n[i];
            if (!device) {
                cur.ptr[f] = d.as<uint8_t>() + off;
                h_in[f] = in[i];
                dsts.push_back(d.as<uint8_t>() + off);
                srcs.push_back(in[i]);
                sizes.push_back(in_n[i]);
            } else {
                cur.ptr[f] = in[i];
                h_in[f] = h_stage ? h_stage + off : nullptr;
            }
            off += (in_n[i] + 64 + 255) & ~(size_t)255;
        }
        RSN_CUDA(copy_many(dsts.data(), srcs.data(), sizes.data(), dsts.size(), s));
    }
    if (h_stage) {
        // The header parser only reads a stream up to the byte after its first 5C 0A: fetch growing
        // prefixes (4 KiB covers text-like alphabets, 64 KiB the ~2 800 records of a file of random
        // bytes) instead of the whole stream — all of config 4's compressed files are 528 MB, their
        // headers 35 MB, and nothing else runs while a pass waits for this copy.
        std::vector<size_t> have(G, 0);
        std::vector<char> done(G, 0);
        const size_t rounds[3] = {(size_t)4 << 10, (size_t)64 << 10, ~(size_t)0};
        for (int r = 0; r < 3; r++) {
            std::vector<void *> dsts;
            std::vector<const void *> srcs;
            std::vector<size_t> sizes;
            std::vector<size_t> upto(G, 0);
            for (size_t f = 0; f < G; f++) {
                const size_t n = in_n[idx[f]];
                if (done[f]) continue;
                upto[f] = std::min(n, rounds[r]);
                dsts.push_back(const_cast<uint8_t *>(h_in[f]) + have[f]);
                srcs.push_back(in[idx[f]] + have[f]);
                sizes.push_back(upto[f] - have[f]);
            }
            if (dsts.empty()) break;
            RSN_CUDA(copy_many(dsts.data(), srcs.data(), sizes.data(), dsts.size(), s));
            RSN_CUDA(stream_wait(s));
            for (size_t f = 0; f < G; f++) {
                if (done[f]) continue;
                const size_t n = in_n[idx[f]];
                const uint8_t *h = h_in[f];
                // first 5C 0A among the bytes present (rescan from one byte before the previous end)
                size_t i = have[f] ? have[f] - 1 : 0;
                bool found = false;
                while (i + 1 < upto[f]) {
                    const void *q = memchr(h + i, 0x5C, upto[f] - 1 - i);
                    if (!q) break;
                    i = (size_t)(static_cast<const uint8_t *>(q) - h);
                    if (h[i + 1] == 0x0A) {
                        found = true;
                        break;
                    }
                    i++;
                }
                have[f] = upto[f];
                if ((found && i + 3 <= have[f]) || have[f] == n) done[f] = 1;
            }
        }
    }
    const bool have_host = !device || h_stage != nullptr;
    tr.mark(device ? "stage" : "h2d");
    const size_t k = algos.size();
    for (size_t step = 0; step < k; step++) {
        const Algo a = compress ? algos[step] : algos[k - 1 - step];
        const uint8_t *const *hp = step == 0 && have_host ? h_in.data() : nullptr;  // host copies: first stage only
        BatchIO next;
        const bool prefix_only = hp && h_stage != nullptr;
        int rc = stage_batched(a, compress, cur, hp, next, s, prefix_only);
        if (rc == RSN_ERR_UNSUPPORTED) {
            next.release(s);
            rc = stage_per_file(a, compress, cur, hp, next, s, prefix_only);
        }
        cur.release(s);
        if (rc != RSN_OK) {
            next.release(s);
            return rc;
        }
        cur = std::move(next);
        tr.mark(a == ALGO_LZSS ? "lzss stage" : "huffman stage");
    }
    if (device) {  // every file into its own result buffer, one launch for the group
        HostVec<CopyJob> jobs(G);
        if (!jobs.data()) return RSN_ERR_NOMEM;
        size_t cap = 0;
        int rc = RSN_OK;
        for (size_t f = 0; f < G; f++) {
            const size_t i = idx[f];
            out[i] = nullptr;
            out_n[i] = 0;
            jobs[f] = CopyJob{nullptr, nullptr, 0};
            if (rcs) rcs[i] = cur.rc[f];
            if (cur.rc[f] != RSN_OK) continue;
            uint8_t *p = (uint8_t *)out_alloc(cur.n[f] + 16, s);
            if (!p) {
                if (rcs) rcs[i] = RSN_ERR_NOMEM;
                rc = RSN_ERR_NOMEM;
                continue;
            }
            jobs[f] = CopyJob{cur.ptr[f], p, cur.n[f]};
            cap = std::max<size_t>(cap, cur.n[f]);
            out[i] = p;
            out_n[i] = cur.n[f];
        }
        DevBuf dj;
        RSN_TRY(dj.alloc(G * sizeof(CopyJob), s));
        RSN_CUDA(cudaMemcpyAsync(dj.p, jobs.data(), G * sizeof(CopyJob), cudaMemcpyHostToDevice, s));
        if (cap) RSN_LAUNCH(kb_copy_out, dim3((unsigned)div_up(cap, 16384), (unsigned)G), 256, 0, s, dj.as<CopyJob>());
        cudaError_t e = stream_wait(s);  // `jobs` is read by the copy above
        if (e != cudaSuccess) rc = cuda_fail(e, "batch copy-out sync", __FILE__, __LINE__);
        cur.release(s);
        tr.mark("copy out");
        return rc;
    }
    // device -> host: all copies queued, one synchronisation.  Results that the last stage laid out
    // back to back in one device buffer come over with ONE copy into one pinned block, of which every
    // file gets its part (a 256 KiB copy costs the host as much driver time as the wire time it saves).
    int rc = RSN_OK;
    std::vector<char> fetched(G, 0);
    for (const auto &span : cur.spans) {
        const uint8_t *lo = nullptr, *hi = nullptr;
        size_t sum = 0, members = 0;
        for (size_t f = 0; f < G; f++) {
            if (cur.rc[f] != RSN_OK || fetched[f] || cur.n[f] == 0) continue;
            if (cur.ptr[f] < span.first || cur.ptr[f] + cur.n[f] > span.first + span.second) continue;
            if (!lo || cur.ptr[f] < lo) lo = cur.ptr[f];
            if (!hi || cur.ptr[f] + cur.n[f] > hi) hi = cur.ptr[f] + cur.n[f];
            sum += cur.n[f];
            members++;
        }
        if (members < 2 || (size_t)(hi - lo) > sum + sum / 4 + (members << 9)) continue;  // sparse: file by file
        uint8_t *block = (uint8_t *)host_out_alloc((size_t)(hi - lo));
        if (!block) continue;
        // in pieces of 2 MiB, queued with one call: the copy engine takes copies in order, and the
        // 8-byte size and flag read-backs of the OTHER workers' kernels would otherwise wait behind
        // whole 64 MiB results (the last group of a pass crawled through 22 + 9 ms of such waits)
        cudaError_t e;
        {
            const size_t piece = (size_t)2 << 20, span_n = (size_t)(hi - lo);
            std::vector<void *> dsts;
            std::vector<const void *> srcs;
            std::vector<size_t> sizes;
            for (size_t at = 0; at < span_n; at += piece) {
                dsts.push_back(block + at);
                srcs.push_back(lo + at);
                sizes.push_back(std::min(piece, span_n - at));
            }
            e = copy_many(dsts.data(), srcs.data(), sizes.data(), dsts.size(), s);
        }
        if (e != cudaSuccess) {
            rsn_free(block);
            rc = cuda_fail(e, "batch d2h", __FILE__, __LINE__);
            continue;
        }
        std::vector<void *> parts;
        parts.reserve(members);
        for (size_t f = 0; f < G; f++) {
            if (cur.rc[f] != RSN_OK || fetched[f] || cur.n[f] == 0) continue;
            if (cur.ptr[f] < span.first || cur.ptr[f] + cur.n[f] > span.first + span.second) continue;
            const size_t i = idx[f];
            out[i] = block + (cur.ptr[f] - lo);
            out_n[i] = cur.n[f];
            if (rcs) rcs[i] = RSN_OK;
            fetched[f] = 1;
            parts.push_back(out[i]);
        }
        host_out_adopt_parts(block, parts.data(), parts.size());
    }
    for (size_t f = 0; f < G; f++) {
        if (fetched[f]) continue;
        const size_t i = idx[f];
        out[i] = nullptr;
        out_n[i] = 0;
        if (rcs) rcs[i] = cur.rc[f];
        if (cur.rc[f] != RSN_OK) continue;
        uint8_t *h = (uint8_t *)host_out_alloc(cur.n[f] ? cur.n[f] : 1);
        if (!h) {
            if (rcs) rcs[i] = RSN_ERR_NOMEM;
            rc = RSN_ERR_NOMEM;
            continue;
        }
        if (cur.n[f]) {
            cudaError_t e = cudaMemcpyAsync(h, cur.ptr[f], cur.n[f], cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) rc = cuda_fail(e, "batch d2h", __FILE__, __LINE__);
        }
        out[i] = h;
        out_n[i] = cur.n[f];
    }
    cudaError_t e = stream_wait(s);
    if (e != cudaSuccess) rc = cuda_fail(e, "batch d2h sync", __FILE__, __LINE__);
    cur.release(s);
    tr.mark("d2h");
    return rc;
}

}  // namespace
}  // namespace rsn

extern "C" {

// The grouping rsn_batch_layers would apply to files of these sizes (host logic only, no device
// needed): group_of[i] = group index of file i, or -1 if it goes through the per-file path.
int rsn_batch_plan(size_t count, const size_t *in_n, int64_t *group_of, size_t *n_groups) {
    if ((!in_n || !group_of) && count) return RSN_ERR_INVALID_ARG;
    std::vector<std::vector<size_t>> groups;
    std::vector<size_t> singles;
    batch_plan(count, nullptr, in_n, 0, groups, singles);
    for (size_t i : singles) group_of[i] = -1;
    for (size_t g = 0; g < groups.size(); g++)
        for (size_t i : groups[g]) group_of[i] = (int64_t)g;
    if (n_groups) *n_groups = groups.size();
    return RSN_OK;
}

// Host buffers (device == 0): small files are cut into groups and every stage runs once per group
// (batch.cuh); files that are empty or larger than kBatchMaxFile, and device-resident batches
// (device != 0), go through the per-file path.  Groups and leftover files are spread over
// `workers` host threads, each with its own stream and arena.
int rsn_batch_layers(const char *algorithms, int compress, size_t count, const uint8_t *const *in, const size_t *in_n,
                     uint8_t **out, size_t *out_n, int *rcs, int workers, int device) {
    if (!in || !in_n || !out || !out_n) return RSN_ERR_INVALID_ARG;
    std::vector<Algo> algos;
    RSN_TRY(parse_layers(algorithms, algos));
    RSN_TRY(ensure_ctx());
    const int dev = ctx().device;
    if (workers <= 0) workers = 8;
    std::vector<int> rc_local;
    if (!rcs) {
        rc_local.assign(count ? count : 1, RSN_OK);
        rcs = rc_local.data();
    }
    std::vector<std::vector<size_t>> groups;
    std::vector<size_t> singles;
    batch_plan(count, in, in_n, device, groups, singles, workers);
    const size_t units = groups.size() + singles.size();
    if ((size_t)workers > units) workers = (int)(units ? units : 1);
    {
        // host cores this process may count on: all of them, unless RSN_HOST_CORES says otherwise (one
        // rank per GPU on a shared box: cores / ranks)
        int hw = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("RSN_HOST_CORES")) {
            const int v = atoi(e);
            if (v >= 1) hw = v;
        }
        set_batch_host_threads(std::max(1, (hw > 0 ? hw : 8) / workers));
    }
    std::atomic<size_t> next{0};
    std::atomic<int> first_err{RSN_OK};
    auto note = [&](int rc) {
        int expect = RSN_OK;
        if (rc != RSN_OK) first_err.compare_exchange_strong(expect, rc);
    };
    auto job = [&]() {
        if (rsn_init(dev) != RSN_OK) {
            note(RSN_ERR_CUDA);
            return;
        }
        static const bool stagger = [] {
            const char *v = getenv("RSN_BATCH_STAGGER");
            return !(v && v[0] == '0');
        }();
        cudaStream_t s = stagger && ctx().batch_stream ? ctx().batch_stream : ctx().own_stream;
        set_thread_blocking_sync(true);
        for (;;) {
            const size_t u = next.fetch_add(1);
            if (u >= units) break;
            if (u < groups.size()) {
                const std::vector<size_t> &idx = groups[u];
                const int rc = batch_group(algos, compress != 0, idx, in, in_n, out, out_n, rcs, device != 0, s);
                if (rc != RSN_OK) {  // the whole group failed (device error, out of memory)
                    for (size_t i : idx) {
                        if (out[i] && device) out_free(out[i], s);
                        else if (out[i]) rsn_free(out[i]);
                        out[i] = nullptr;
                        out_n[i] = 0;
                        if (rcs) rcs[i] = rc;
                    }
                    note(rc);
                } else {
                    for (size_t i : idx) note(rcs[i]);
                }
                continue;
            }
            const size_t i = singles[u - groups.size()];
            int rc;
            uint8_t *r = nullptr;
            size_t rn = 0;
            {
                ArenaScope scope(s);
                const uint8_t *d_in = in[i];
                DevBuf d;
                rc = RSN_OK;
                if (!device) {
                    rc = to_device(in[i], in_n[i], d, s);
                    d_in = d.as<uint8_t>();
                }
                if (rc == RSN_OK) rc = layers_dev(algos, compress != 0, d_in, in_n[i], &r, &rn, s);
                if (rc == RSN_OK) {
                    if (device) {
                        stream_wait(s);
                        out[i] = r;
                        out_n[i] = rn;
                    } else {
                        rc = to_host(r, rn, &out[i], &out_n[i], s);
                    }
                }
            }
            if (rc != RSN_OK) {
                out[i] = nullptr;
                out_n[i] = 0;
                note(rc);
            }
            if (rcs) rcs[i] = rc;
        }
    };
    pool().run(workers, job);
    return first_err.load();
}

// ---- device-buffer API

// stream == NULL: run on the context's own stream and synchronise before returning
static int finish(int rc, void *stream) {
    if (!stream) {
        cudaError_t e = cudaStreamSynchronize(ctx().own_stream);
        if (rc == RSN_OK && e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    return rc;
}

int rsn_dev_lzss_compress(const uint8_t *d_in, size_t n, int64_t window, int variant, uint8_t **d_out, size_t *out_n,
                          void *stream) {
    if ((!d_in && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(lzss_compress_dev(d_in, n, window, variant, d_out, out_n, pick_stream(stream)), stream);
}
int rsn_dev_lzss_decompress(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream) {
    if ((!d_in && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(lzss_decompress_dev(d_in, n, d_out, out_n, pick_stream(stream)), stream);
}
int rsn_dev_huff_compress(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream) {
    if ((!d_in && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(huff_compress_dev(d_in, n, d_out, out_n, pick_stream(stream)), stream);
}
int rsn_dev_huff_decompress(const uint8_t *d_in, size_t n, int strict_limits, uint8_t **d_out, size_t *out_n,
                            void *stream) {
    if ((!d_in && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(huff_decompress_dev(d_in, n, nullptr, strict_limits, d_out, out_n, pick_stream(stream)), stream);
}

int rsn_dev_lzss_emit(const uint8_t *d_enc, size_t n, int64_t window, int variant, const uint32_t *d_packed,
                      uint8_t **d_out, size_t *out_n, void *stream) {
    if ((!d_enc && n) || (!d_packed && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(lzss_emit_dev(d_enc, n, window, variant, d_packed, d_out, out_n, pick_stream(stream)), stream);
}
int rsn_dev_lzss_escape(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream) {
    if ((!d_in && n) || !d_out || !out_n) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    return finish(lzss_escape_dev(d_in, n, d_out, out_n, pick_stream(stream)), stream);
}

int rsn_dev_download(const void *d_src, size_t n, void *h_dst, void *stream) {
    RSN_TRY(ensure_ctx());
    cudaStream_t s = pick_stream(stream);
    if (n) RSN_CUDA(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    return RSN_OK;
}
int rsn_dev_upload(const void *h_src, size_t n, void *d_dst, void *stream) {
    RSN_TRY(ensure_ctx());
    cudaStream_t s = pick_stream(stream);
    if (n) RSN_CUDA(cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, s));
    RSN_CUDA(stream_wait(s));
    return RSN_OK;
}

int rsn_dev_lzss_match(const uint8_t *d_enc, size_t n, int64_t window, uint32_t *d_packed, void *stream) {
    if ((!d_enc && n) || (!d_packed && n)) return RSN_ERR_INVALID_ARG;
    RSN_TRY(ensure_ctx());
    if (n == 0) return RSN_OK;
    uint32_t W = 0;
    RSN_TRY(lzss_effective_window(window, n, &W));
    ArenaScope scope(pick_stream(stream));
    return finish(lzss_match(d_enc, n, W, d_packed, pick_stream(stream)), stream);
}

}  // extern "C"
