// huff.cuh — internal interfaces of the Huffman kernels.
#pragma once
#include "common.cuh"

namespace rsn {

int huff_compress_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s);
// h_prefix: host copy of the first h_prefix_n bytes of the stream (header search), may be the
// whole stream.
int huff_decompress_dev(const uint8_t *d_in, size_t n, const uint8_t *h_in_or_null, int strict, uint8_t **d_out,
                        size_t *out_n, cudaStream_t s);

// ---- batches: tree, codes and tables built on the device, one warp per file (huff_tree.cu)

struct CodeEntry {
    uint32_t rune;
    uint32_t len;
    uint64_t code;
};
struct HuffNodeDev {
    int32_t left;   // child index, or -1 for a leaf
    int32_t right;  // child index, or the rune for a leaf
};

struct TreeJob {             // one file (device memory)
    const uint32_t *freq;    // k leaves in (freq asc, rune asc) order, as the reference sorts them
    const uint32_t *rune;
    uint32_t k;
    uint32_t bmask;          // compress: btab has bmask + 1 slots, all empty (rune 0xFFFFFFFF)
    HuffNodeDev *nodes;      // out: 2k - 1 nodes, leaves first
    uint32_t *parent;        // scratch: 2k - 1 words
    uint64_t *scode;         // compress: code / length of runes < 256 (256 entries, zeroed), else null
    uint8_t *slen;
    CodeEntry *btab;         // compress: codes of the other runes (open addressing)
    uint32_t *lut;           // decompress: 2^12-entry table over the first bits of a code, else null
    uint64_t total_bits;     // out: sum of freq * code length
    int32_t root;            // out
    uint32_t maxlen;         // out
    uint32_t flags;          // out: bit 0 = a code is longer than 64 bits
    uint32_t pad;
};
constexpr uint32_t kTreeMaxLeaves = 24576;  // the heap of one file lives in shared memory
int huff_tree_batch(TreeJob *d_jobs, size_t G, uint32_t kmax, cudaStream_t s);

}  // namespace rsn
