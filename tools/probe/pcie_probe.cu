// Copy-rate probe for the batch path's upload: N separately pinned 256 KiB host buffers into one
// device buffer by (a) N cudaMemcpyAsync calls, (b) one cudaMemcpyBatchAsync, (c) one gather kernel
// that reads the pinned buffers over PCIe itself.  nvcc -O2 -arch=sm_100a -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Job { const uint4 *src; uint4 *dst; size_t n16; };
__global__ void __launch_bounds__(256) k_gather(const Job *jobs) {
    const Job j = jobs[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < j.n16; i += (size_t)gridDim.x * 256) j.dst[i] = j.src[i];
}

int main() {
    const size_t N = 2048, SZ = 256 << 10;
    std::vector<void *> h(N);
    for (auto &p : h) CK(cudaHostAlloc(&p, SZ, cudaHostAllocDefault));
    uint8_t *d;
    CK(cudaMalloc(&d, N * SZ));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    const double gb = N * SZ / 1e9;
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = now();
        for (size_t i = 0; i < N; i++) CK(cudaMemcpyAsync(d + i * SZ, h[i], SZ, cudaMemcpyHostToDevice, s));
        auto t1 = now();
        CK(cudaStreamSynchronize(s));
        auto t2 = now();
        printf("memcpyAsync x%zu: issue %.2f ms, total %.2f ms, %.1f GB/s\n", N, secs(t0, t1) * 1e3, secs(t0, t2) * 1e3, gb / secs(t0, t2));
    }
    {
        std::vector<void *> dsts(N), srcs(N);
        std::vector<size_t> sizes(N, SZ);
        for (size_t i = 0; i < N; i++) { dsts[i] = d + i * SZ; srcs[i] = h[i]; }
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t idx0 = 0, fail = 0;
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = now();
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), N, &at, &idx0, 1, &fail, s);
            auto t1 = now();
            if (e != cudaSuccess) { printf("cudaMemcpyBatchAsync: %s (fail idx %zu)\n", cudaGetErrorString(e), fail); break; }
            CK(cudaStreamSynchronize(s));
            auto t2 = now();
            printf("memcpyBatchAsync: issue %.2f ms, total %.2f ms, %.1f GB/s\n", secs(t0, t1) * 1e3, secs(t0, t2) * 1e3, gb / secs(t0, t2));
        }
    }
    {
        std::vector<Job> jobs(N);
        for (size_t i = 0; i < N; i++) jobs[i] = Job{(const uint4 *)h[i], (uint4 *)(d + i * SZ), SZ / 16};
        Job *dj;
        CK(cudaMalloc(&dj, N * sizeof(Job)));
        CK(cudaMemcpy(dj, jobs.data(), N * sizeof(Job), cudaMemcpyHostToDevice));
        for (int bx : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; rep++) {
                auto t0 = now();
                k_gather<<<dim3(bx, N), 256, 0, s>>>(dj);
                CK(cudaStreamSynchronize(s));
                auto t2 = now();
                if (rep) printf("gather kernel (%d CTAs per file): %.2f ms, %.1f GB/s\n", bx, secs(t0, t2) * 1e3, gb / secs(t0, t2));
            }
        }
        // device -> pinned host by a kernel (scatter)
        for (size_t i = 0; i < N; i++) jobs[i] = Job{(const uint4 *)(d + i * SZ), (uint4 *)h[i], SZ / 16};
        CK(cudaMemcpy(dj, jobs.data(), N * sizeof(Job), cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 2; rep++) {
            auto t0 = now();
            k_gather<<<dim3(4, N), 256, 0, s>>>(dj);
            CK(cudaStreamSynchronize(s));
            auto t2 = now();
            if (rep) printf("scatter kernel to pinned host (4 CTAs per file): %.2f ms, %.1f GB/s\n", secs(t0, t2) * 1e3, gb / secs(t0, t2));
        }
    }
    // issue cost from 8 threads at once (each its own stream)
    {
        std::vector<std::thread> th;
        std::vector<cudaStream_t> ss(8);
        for (auto &x : ss) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
        auto t0 = now();
        for (int k = 0; k < 8; k++)
            th.emplace_back([&, k] {
                for (size_t i = k * N / 8; i < (k + 1) * N / 8; i++) cudaMemcpyAsync(d + i * SZ, h[i], SZ, cudaMemcpyHostToDevice, ss[k]);
                cudaStreamSynchronize(ss[k]);
            });
        for (auto &t : th) t.join();
        auto t2 = now();
        printf("memcpyAsync x%zu from 8 threads/streams: %.2f ms, %.1f GB/s\n", N, secs(t0, t2) * 1e3, gb / secs(t0, t2));
    }
    return 0;
}
