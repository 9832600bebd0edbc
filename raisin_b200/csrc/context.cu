// context.cu — per-thread context, error plumbing, buffer pools, spine scan, library lifetime.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace rsn {

static thread_local Ctx g_ctx;

Ctx &ctx() { return g_ctx; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    Ctx &c = ctx();
    snprintf(c.cuda_err, sizeof(c.cuda_err), "%s: %s at %s:%d (%s)", cudaGetErrorName(e), cudaGetErrorString(e), file,
             line, what);
    cudaGetLastError();  // clear non-sticky errors
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return RSN_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return RSN_ERR_NOMEM;
    return RSN_ERR_CUDA;
}

static int init_ctx(int device) {
    Ctx &c = ctx();
    int count = 0;
    RSN_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) return RSN_ERR_NO_DEVICE;
    if (device < 0) {
        RSN_CUDA(cudaGetDevice(&device));
    }
    if (device >= count) return RSN_ERR_INVALID_ARG;
    RSN_CUDA(cudaSetDevice(device));
    if (c.ready && c.device == device) return RSN_OK;
    c.device = device;
    if (!c.own_stream) RSN_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    if (!c.h_scalars) RSN_CUDA(cudaHostAlloc((void **)&c.h_scalars, 64 * sizeof(uint64_t), cudaHostAllocDefault));
    // keep stream-ordered allocations cached in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    RSN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    RSN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    c.ready = true;
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

int DevBuf::alloc(size_t n, cudaStream_t stream) {
    reset();
    s = stream;
    bytes = n ? n : 1;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess) {
        p = nullptr;
        return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    }
    return RSN_OK;
}

void DevBuf::reset() {
    if (p) {
        cudaFreeAsync(p, s);
        p = nullptr;
    }
}

// ----------------------------------------------------------------------------- spine scan

__global__ void __launch_bounds__(1024) k_spine_scan_u64(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
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

int spine_scan_u64(const uint64_t *d_in, uint64_t *d_out, uint64_t *d_total, size_t count, cudaStream_t s) {
    RSN_LAUNCH(k_spine_scan_u64, 1, 1024, 0, s, d_in, d_out, d_total, count);
    return RSN_OK;
}

int read_u64(const uint64_t *d_src, uint64_t *h_dst, cudaStream_t s) {
    Ctx &c = ctx();
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars, d_src, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(cudaStreamSynchronize(s));
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
        if (cudaHostAlloc(&p, c, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        std::lock_guard<std::mutex> g(mu);
        live[p] = c;
        return p;
    }
    void put(void *p) {
        if (!p) return;
        size_t c = 0;
        {
            std::lock_guard<std::mutex> g(mu);
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

}  // namespace rsn

// ----------------------------------------------------------------------------- C ABI: lifetime

extern "C" {

int rsn_init(int device) { return rsn::init_ctx(device); }

void rsn_shutdown(void) {
    rsn::Ctx &c = rsn::ctx();
    if (c.own_stream) {
        cudaStreamSynchronize(c.own_stream);
        cudaStreamDestroy(c.own_stream);
        c.own_stream = nullptr;
    }
    if (c.h_scalars) {
        cudaFreeHost(c.h_scalars);
        c.h_scalars = nullptr;
    }
    c.ready = false;
    rsn::pinned().drain();
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
    cudaFreeAsync(d_ptr, s);
}

uint64_t rsn_kernel_launches(void) { return rsn::ctx().launches; }
void rsn_reset_kernel_launches(void) { rsn::ctx().launches = 0; }
const char *rsn_version(void) { return "raisin_b200 0.1 sm_100a"; }

}  // extern "C"
