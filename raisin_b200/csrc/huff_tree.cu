// huff_tree.cu — buildTree + printCodes (huffman.go:58-103, 110-127) on the device, for batches.
//
// The shape of the tree under frequency ties is fixed by Go's container/heap sift rules
// (Less compares frequencies only, huffman.go:43-45), so the reference's heap has to be replayed
// step by step; that is a chain of dependent loads (~0.5 ms on a host core for the ~2 800 symbols
// Go's rune semantics find in 256 KiB of binary data) and in a batch it is the host that runs out
// first.  Here every file gets one warp: lane 0 replays the heap in shared memory (all files of the
// group at once, on different SMs), then the 32 lanes derive the codes leaf by leaf and fill the
// tables the encode / decode kernels read.  huff_host.cpp remains the single-stream implementation;
// both are checked against the oracle.
#include "common.cuh"
#include "huff.cuh"

namespace rsn {

namespace {

constexpr int kLutBitsTree = 12;  // == kLutBits in huff_decode.cu

// heap entry = (frequency sum << 32) | node: Less compares the high words only (huffman.go:43-45)
__device__ __forceinline__ bool heap_less(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }

// container/heap.down / up (Go 1.15) with the moving element carried in a register: the array ends
// up exactly as after Go's swaps.
// Two levels per round trip to shared memory: the four grandchildren are loaded together with the two
// children, so the second comparison does not wait for another load (the replay is a chain of
// dependent loads, and one lane per file runs it).
__device__ __forceinline__ void heap_down(uint64_t *h, int i, int n) {
    const uint64_t v = h[i];
    for (;;) {
        const int j1 = 2 * i + 1;
        if (j1 >= n) break;
        const int j2 = j1 + 1 < n ? j1 + 1 : j1;
        const int g = 2 * j1 + 1;  // grandchildren g .. g+3 (children of j1, then of j1+1)
        const uint64_t a = h[j1], b = h[j2];
        const uint64_t g0 = g < n ? h[g] : 0, g1 = g + 1 < n ? h[g + 1] : 0, g2 = g + 2 < n ? h[g + 2] : 0,
                       g3 = g + 3 < n ? h[g + 3] : 0;
        const bool right = heap_less(b, a);  // j2 == j1 gives false
        const uint64_t c = right ? b : a;
        if (!heap_less(c, v)) break;
        h[i] = c;
        i = right ? j2 : j1;
        // second level from the prefetched values
        const int k1 = 2 * i + 1;
        if (k1 >= n) break;
        const bool has2 = k1 + 1 < n;
        const uint64_t a2 = right ? g2 : g0, b2 = has2 ? (right ? g3 : g1) : a2;
        const bool right2 = heap_less(b2, a2);
        const uint64_t c2 = right2 ? b2 : a2;
        if (!heap_less(c2, v)) break;
        h[i] = c2;
        i = right2 ? k1 + 1 : k1;
    }
    h[i] = v;
}
__device__ __forceinline__ void heap_up(uint64_t *h, int j) {
    const uint64_t v = h[j];
    while (j > 0) {
        const int i = (j - 1) / 2;
        if (!heap_less(v, h[i])) break;
        h[j] = h[i];
        j = i;
    }
    h[j] = v;
}

__global__ void __launch_bounds__(32) kb_huff_tree(TreeJob *__restrict__ jobs) {
    extern __shared__ uint64_t heap[];
    TreeJob &job = jobs[blockIdx.x];
    const int k = (int)job.k;
    if (k == 0) return;
    const unsigned lane = threadIdx.x;
    HuffNodeDev *nodes = job.nodes;
    uint32_t *parent = job.parent;
    for (int i = lane; i < k; i += 32) {
        heap[i] = ((uint64_t)job.freq[i] << 32) | (uint64_t)i;
        nodes[i] = HuffNodeDev{-1, (int32_t)job.rune[i]};
    }
    __syncwarp();
    int root = 0;
    if (lane == 0) {
        int n = k;
        for (int i = n / 2 - 1; i >= 0; i--) heap_down(heap, i, n);  // heap.Init
        int next = k;
        while (n > 1) {
            // a = Pop(), b = Pop(): swap root and last, sift the new root over the shorter heap
            const uint64_t a = heap[0];
            heap[0] = heap[--n];
            heap_down(heap, 0, n);
            const uint64_t b = heap[0];
            heap[0] = heap[--n];
            if (n > 0) heap_down(heap, 0, n);
            const uint32_t an = (uint32_t)a, bn = (uint32_t)b;
            nodes[next] = HuffNodeDev{(int32_t)an, (int32_t)bn};  // left = first popped (huffman.go:96-99)
            parent[an] = (uint32_t)next;
            parent[bn] = (uint32_t)next | 0x80000000u;
            heap[n] = (((a >> 32) + (b >> 32)) << 32) | (uint64_t)next;  // Push (sums stay below 2^32: checked by the host)
            heap_up(heap, n);
            n++;
            next++;
        }
        root = (int)(uint32_t)heap[0];
        job.root = root;
    }
    root = __shfl_sync(0xffffffffu, root, 0);
    __threadfence_block();
    __syncwarp();
    // codes: left = 0, right = 1 from the root down (printCodes); walking up collects them LSB first
    uint64_t bits_sum = 0;
    uint32_t maxlen = 0, flags = 0;
    if (job.scode) {
        for (int i = lane; i < k; i += 32) {
            uint64_t code = 0;
            uint32_t len = 0;
            int node = i;
            while (node != root) {
                const uint32_t p = parent[node];
                if (len < 64) code |= (uint64_t)(p >> 31) << len;
                len++;
                node = (int)(p & 0x7FFFFFFFu);
            }
            if (len > 64) flags = 1;
            bits_sum += (uint64_t)len * (uint64_t)job.freq[i];
            maxlen = max(maxlen, len);
            const uint32_t r = job.rune[i];
            if (r < 256) {
                job.scode[r] = code;
                job.slen[r] = (uint8_t)len;
            } else {
                uint32_t h = ((r * 2654435761u) >> 12) & job.bmask;  // big_hash
                while (atomicCAS(&job.btab[h].rune, 0xFFFFFFFFu, r) != 0xFFFFFFFFu) h = (h + 1) & job.bmask;
                job.btab[h].len = len;
                job.btab[h].code = code;
            }
        }
        for (int d = 16; d; d >>= 1) {
            bits_sum += __shfl_down_sync(0xffffffffu, bits_sum, d);
            maxlen = max(maxlen, __shfl_down_sync(0xffffffffu, maxlen, d));
            flags |= __shfl_down_sync(0xffffffffu, flags, d);
        }
        if (lane == 0) {
            job.total_bits = bits_sum;
            job.maxlen = maxlen;
            job.flags = flags;
        }
    }
    if (job.lut) {
        for (uint32_t idx = lane; idx < (1u << kLutBitsTree); idx += 32) {
            int node = root;
            uint32_t len = 0;
            HuffNodeDev nd = nodes[node];
            while (nd.left >= 0 && len < (uint32_t)kLutBitsTree) {
                node = ((idx >> (kLutBitsTree - 1 - len)) & 1u) ? nd.right : nd.left;
                nd = nodes[node];
                len++;
            }
            job.lut[idx] = nd.left < 0 ? ((1u << 31) | (len << 21) | ((uint32_t)nd.right & 0x1FFFFFu)) : (uint32_t)node;
        }
    }
}

}  // namespace

int huff_tree_batch(TreeJob *d_jobs, size_t G, uint32_t kmax, cudaStream_t s) {
    if (G == 0) return RSN_OK;
    if (kmax > kTreeMaxLeaves) return RSN_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(kmax ? kmax : 1) * 8;
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        RSN_CUDA(cudaFuncSetAttribute(kb_huff_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTreeMaxLeaves * 8)));
        attr = (size_t)kTreeMaxLeaves * 8;
    }
    RSN_LAUNCH(kb_huff_tree, (unsigned)G, 32, smem, s, d_jobs);
    return RSN_OK;
}

}  // namespace rsn
