// huff_host.h — host side of the Huffman path: the tree is built on the host exactly as the
// reference builds it (north_star (3)); everything per-symbol runs on the device.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace rsn {

struct HuffLeaf {
    int64_t freq;  // Go int (sums wrap as in Go)
    int32_t rune;
};

struct HuffNode {
    int32_t left;   // child index, or -1 for a leaf
    int32_t right;  // child index, or the rune for a leaf
};

struct HuffTree {
    std::vector<HuffNode> nodes;  // leaves first (sorted by freq, rune), then internal nodes
    std::vector<int64_t> freq;
    int32_t root = -1;
    size_t n_leaves = 0;
};

// buildTree (huffman.go:58-103): leaves ordered (freq asc, rune asc), container/heap Init,
// then Pop,Pop,Push(a+b, left=a, right=b) until one node remains.  leaves must be non-empty.
void huff_build_tree(std::vector<HuffLeaf> leaves, HuffTree &t);
// the (freq asc, rune asc) leaf order alone (huffman.go:64-87)
void huff_sort_leaves(std::vector<HuffLeaf> &leaves);

struct HuffCode {
    int32_t rune;
    uint8_t len;
    uint64_t code;  // MSB-first in the low `len` bits
    int64_t freq;   // the leaf's frequency
};
// printCodes (huffman.go:110-127).  Returns false if a code is longer than 64 bits.
bool huff_codes(const HuffTree &t, std::vector<HuffCode> &codes);

// Header bytes (huffman.go:312-318) in canonical order: ascending rune, last two records
// swapped if the last would be rune 0x5C (DESIGN.md "header order").
void huff_header(const std::vector<HuffLeaf> &leaves, std::vector<uint8_t> &hdr);

// decodeTree's parse (huffman.go:196-227) of header bytes into the rune->freq map (as leaves,
// unique by rune, last assignment wins).  Returns false where the reference would panic.
bool huff_parse_header(const uint8_t *h, size_t hn, std::vector<HuffLeaf> &leaves);

}  // namespace rsn
