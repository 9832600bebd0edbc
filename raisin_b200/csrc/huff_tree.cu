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

#include <algorithm>
#include "huff.cuh"

namespace rsn {

namespace {

constexpr int kLutBitsTree = 12;  // == kLutBits in huff_decode.cu

// Heap entry = (frequency sum << SH) | node; Less compares the frequency sums only (huffman.go:43-45).
// Two layouts: 64-bit entries with SH = 32 for any file, and 32-bit entries with SH = 13 for files of
// fewer than 2^19 runes and fewer than 4096 leaves (every file of up to 512 KiB) — the replay is one
// lane following a chain of dependent shared-memory loads and integer operations, and 32-bit entries
// take about a third of the instructions out of that chain.
template <typename E, int SH>
struct HeapOps {
    static __device__ __forceinline__ bool less(E a, E b) { return (a >> SH) < (b >> SH); }
    // container/heap.down / up (Go 1.15) with the moving element carried in a register: the array
    // ends up exactly as after Go's swaps.
    static __device__ __forceinline__ void down(E *h, int i, int n) {
        const E v = h[i];
        for (;;) {
            const int j1 = 2 * i + 1;
            if (j1 >= n) break;
            const int j2 = j1 + 1 < n ? j1 + 1 : j1;
            const E a = h[j1], b = h[j2];
            const bool right = less(b, a);  // j2 == j1 gives false
            const E c = right ? b : a;
            if (!less(c, v)) break;
            h[i] = c;
            i = right ? j2 : j1;
        }
        h[i] = v;
    }
    static __device__ __forceinline__ void up(E *h, int j) {
        const E v = h[j];
        while (j > 0) {
            const int i = (j - 1) / 2;
            if (!less(v, h[i])) break;
            h[j] = h[i];
            j = i;
        }
        h[j] = v;
    }
    // heap.Init, then a = Pop(), b = Pop(), Push(a + b) until one node is left (huffman.go:88-102)
    static __device__ __forceinline__ int replay(E *heap, int k, HuffNodeDev *nodes, uint32_t *parent) {
        int n = k;
        for (int i = n / 2 - 1; i >= 0; i--) down(heap, i, n);
        int next = k;
        const E node_mask = ((E)1 << SH) - 1;
        while (n > 1) {
            // Pop: swap root and last, sift the new root over the shorter heap
            const E a = heap[0];
            heap[0] = heap[--n];
            down(heap, 0, n);
            const E b = heap[0];
            heap[0] = heap[--n];
            if (n > 0) down(heap, 0, n);
            const uint32_t an = (uint32_t)(a & node_mask), bn = (uint32_t)(b & node_mask);
            nodes[next] = HuffNodeDev{(int32_t)an, (int32_t)bn};  // left = first popped (huffman.go:96-99)
            parent[an] = (uint32_t)next;
            parent[bn] = (uint32_t)next | 0x80000000u;
            heap[n] = (((a >> SH) + (b >> SH)) << SH) | (E)next;  // Push (the sums fit: checked by the caller)
            up(heap, n);
            n++;
            next++;
        }
        return (int)(uint32_t)(heap[0] & node_mask);
    }
};

// The same replay for files of fewer than 2^19 runes and fewer than 2^DEPTH leaves, arranged for the
// one lane that runs it: 32-bit entries (frequency << 13 | node); the array starts one word into an
// 8-byte aligned buffer so that both children of a node come with ONE 8-byte load; every slot past
// the heap's end holds an all-ones sentinel (nothing is smaller), so a level of the sift is just
// address -> load -> compare -> select, with no bounds tests and no branch; where the sinking element
// stops is settled afterwards from the keys collected on the way.  The array ends up exactly as after
// container/heap's swaps.  hp must hold 2^(DEPTH+1) entries.
constexpr uint32_t kSent = 0xFFFFFFFFu;
__device__ __forceinline__ bool less13(uint32_t a, uint32_t b) { return (a >> 13) < (b >> 13); }

template <int DEPTH>
__device__ __forceinline__ void down_root_small(uint32_t *hp) {
    static_assert(DEPTH % 2 == 0, "two levels per step");
    const uint32_t v = hp[0];
    int idx[DEPTH + 1];
    uint32_t key[DEPTH];
    idx[0] = 0;
    // Two levels per step: the children of node i sit at 2i+1, 2i+2 (one 8-byte load) and its four
    // grandchildren at 4i+3 .. 4i+6 (one 16-byte load, aligned by the same one-word shift), and both
    // addresses follow from i alone — so the chain of dependent loads is half as long as with one
    // level per load (3.1 -> 2.5 ms for the ~2 900 leaves of a file of random bytes).
#pragma unroll
    for (int l = 0; l < DEPTH; l += 2) {
        const int i = idx[l];
        const uint2 c = *reinterpret_cast<const uint2 *>(hp + 2 * i + 1);
        const uint4 g = *reinterpret_cast<const uint4 *>(hp + 4 * i + 3);
        const bool r1 = less13(c.y, c.x);  // a missing right child is a sentinel: false
        key[l] = r1 ? c.y : c.x;
        const int i1 = 2 * i + 1 + (r1 ? 1 : 0);
        idx[l + 1] = i1;
        const uint32_t ga = r1 ? g.z : g.x, gb = r1 ? g.w : g.y;  // the children of i1
        const bool r2 = less13(gb, ga);
        key[l + 1] = r2 ? gb : ga;
        idx[l + 2] = 2 * i1 + 1 + (r2 ? 1 : 0);
    }
    bool going = true;
    int stop = 0;
#pragma unroll
    for (int l = 0; l < DEPTH; l++) {
        going = going && less13(key[l], v);  // Go: if !less(j, i) break
        if (going) {
            hp[idx[l]] = key[l];
            stop = idx[l + 1];
        }
    }
    hp[stop] = v;
}

template <int DEPTH>
__device__ __forceinline__ int replay_small(uint32_t *hp, int k, HuffNodeDev *nodes, uint32_t *parent) {
    int n = k;
    for (int i = n / 2 - 1; i >= 0; i--) HeapOps<uint32_t, 13>::down(hp, i, n);  // heap.Init
    int next = k;
    while (n > 1) {
        // Pop: the last element replaces the root and sinks over the shorter heap
        const uint32_t a = hp[0];
        --n;
        hp[0] = hp[n];
        hp[n] = kSent;
        down_root_small<DEPTH>(hp);
        const uint32_t b = hp[0];
        --n;
        hp[0] = hp[n];
        hp[n] = kSent;
        if (n > 0) down_root_small<DEPTH>(hp);
        const uint32_t an = a & 0x1FFFu, bn = b & 0x1FFFu;
        nodes[next] = HuffNodeDev{(int32_t)an, (int32_t)bn};  // left = first popped (huffman.go:96-99)
        parent[an] = (uint32_t)next;
        parent[bn] = (uint32_t)next | 0x80000000u;
        hp[n] = (((a >> 13) + (b >> 13)) << 13) | (uint32_t)next;  // Push
        HeapOps<uint32_t, 13>::up(hp, n);
        n++;
        next++;
    }
    return (int)(hp[0] & 0x1FFFu);
}

__global__ void __launch_bounds__(32) kb_huff_tree(TreeJob *__restrict__ jobs) {
    extern __shared__ __align__(16) uint64_t heap[];
    TreeJob &job = jobs[blockIdx.x];
    const int k = (int)job.k;
    if (k == 0) return;
    const unsigned lane = threadIdx.x;
    HuffNodeDev *nodes = job.nodes;
    uint32_t *parent = job.parent;
    // 32-bit entries when the sums and the node numbers fit 19 + 13 bits
    uint32_t fsum = 0;
    for (int i = lane; i < k; i += 32) fsum += job.freq[i] < (1u << 19) ? job.freq[i] : (1u << 19);
    for (int d = 16; d; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
    const bool small = k < 4096 && fsum < (1u << 19);
    uint32_t *hp = reinterpret_cast<uint32_t *>(heap) + 1;  // children pairs 8-byte aligned
    const int slots = small ? (k <= 255 ? 512 : 8192) : 0;
    for (int i = lane; i < max(k, slots); i += 32) {
        if (small) hp[i] = i < k ? ((job.freq[i] << 13) | (uint32_t)i) : kSent;
        else heap[i] = ((uint64_t)job.freq[i] << 32) | (uint64_t)i;
        if (i < k) nodes[i] = HuffNodeDev{-1, (int32_t)job.rune[i]};
    }
    __syncwarp();
    int root = 0;
    if (lane == 0) {
        if (!small) root = HeapOps<uint64_t, 32>::replay(heap, k, nodes, parent);
        else if (k <= 255) root = replay_small<8>(hp, k, nodes, parent);  // text-like alphabets
        else root = replay_small<12>(hp, k, nodes, parent);
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
    // 64-bit entries: kmax * 8 bytes; the 32-bit layout pads to 512 or 8192 entries (+ 1 word)
    const size_t smem = std::max<size_t>((size_t)(kmax ? kmax : 1) * 8, kmax <= 255 ? 513 * 4 + 12 : 8193 * 4 + 12);
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        RSN_CUDA(cudaFuncSetAttribute(kb_huff_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTreeMaxLeaves * 8)));
        attr = (size_t)kTreeMaxLeaves * 8;
    }
    RSN_LAUNCH(kb_huff_tree, (unsigned)G, 32, smem, s, d_jobs);
    return RSN_OK;
}

}  // namespace rsn
