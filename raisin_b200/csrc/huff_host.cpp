// huff_host.cpp — see huff_host.h.  Host-side product code (not the oracle).
#include "huff_host.h"

#include <algorithm>
#include <unordered_map>

#include "utf8.cuh"

namespace rsn {

namespace {
// LSD byte-wise radix sort of 64-bit keys, skipping bytes that are equal in all keys.
void radix_sort_u64(std::vector<uint64_t> &keys) {
    const size_t n = keys.size();
    if (n < 2048) {
        std::sort(keys.begin(), keys.end());
        return;
    }
    uint64_t all_or = 0, all_and = ~0ull;
    for (uint64_t k : keys) {
        all_or |= k;
        all_and &= k;
    }
    const uint64_t varying = all_or ^ all_and;
    std::vector<uint64_t> tmp(n);
    uint64_t *src = keys.data(), *dst = tmp.data();
    for (int byte = 0; byte < 8; byte++) {
        if (((varying >> (8 * byte)) & 0xFF) == 0) continue;
        size_t cnt[257] = {0};
        for (size_t i = 0; i < n; i++) cnt[((src[i] >> (8 * byte)) & 0xFF) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; i++) dst[cnt[(src[i] >> (8 * byte)) & 0xFF]++] = src[i];
        std::swap(src, dst);
    }
    if (src != keys.data()) std::copy(src, src + n, keys.data());
}

// container/heap (Go 1.15); Less compares frequencies only (huffman.go:43-45), so the shape of
// the tree under ties is fixed by these exact sift rules.  Entries carry the frequency next to the
// node index so that sifting does not chase pointers.
struct HeapEnt {
    int64_t freq;
    int32_t node;
};
template <typename E, typename Less>
struct GoHeapT {
    std::vector<E> h;
    Less lt;
    bool less(int i, int j) const { return lt(h[i], h[j]); }
    // Go's up/down swap the moving element with a parent/child at every level; carrying it in a
    // register and writing it once at the end leaves the array in exactly the same state.
    void up(int j) {
        const E v = h[j];
        for (;;) {
            int i = (j - 1) / 2;  // parent; truncating division keeps j == 0 at 0
            if (i == j || !lt(v, h[i])) break;
            h[j] = h[i];
            j = i;
        }
        h[j] = v;
    }
    void down(int i, int n) {
        const E v = h[i];
        for (;;) {
            int j1 = 2 * i + 1;
            if (j1 >= n || j1 < 0) break;
            // which child is smaller is a coin flip: select without a branch
            const int j2 = j1 + 1 < n ? j1 + 1 : j1;
            const int j = lt(h[j2], h[j1]) ? j2 : j1;
            if (!lt(h[j], v)) break;
            h[i] = h[j];
            i = j;
        }
        h[i] = v;
    }
    void init() {
        int n = (int)h.size();
        for (int i = n / 2 - 1; i >= 0; i--) down(i, n);
    }
    E pop() {
        int n = (int)h.size() - 1;
        std::swap(h[0], h[n]);
        down(0, n);
        E v = h.back();
        h.pop_back();
        return v;
    }
    void push(E v) {
        h.push_back(v);
        up((int)h.size() - 1);
    }
};
struct LessEnt {
    bool operator()(const HeapEnt &a, const HeapEnt &b) const { return a.freq < b.freq; }
};
// frequency in the high 40 bits, node index in the low 24: half the bytes per sift step
struct LessPacked {
    bool operator()(uint64_t a, uint64_t b) const { return (a >> 24) < (b >> 24); }
};
}  // namespace

void huff_sort_leaves(std::vector<HuffLeaf> &leaves) {
    // leaves in (freq asc, rune asc) order (huffman.go:64-87)
    bool packable = true;
    for (const HuffLeaf &l : leaves)
        if (l.freq < 0 || l.freq >= ((int64_t)1 << 42)) packable = false;
    if (packable) {  // one integer key per leaf sorts much faster than a comparator on pairs
        std::vector<uint64_t> keys(leaves.size());
        for (size_t i = 0; i < leaves.size(); i++) keys[i] = ((uint64_t)leaves[i].freq << 21) | (uint32_t)leaves[i].rune;
        radix_sort_u64(keys);
        for (size_t i = 0; i < leaves.size(); i++) leaves[i] = HuffLeaf{(int64_t)(keys[i] >> 21), (int32_t)(keys[i] & 0x1FFFFFu)};
    } else {
        std::sort(leaves.begin(), leaves.end(), [](const HuffLeaf &a, const HuffLeaf &b) {
            return a.freq != b.freq ? a.freq < b.freq : a.rune < b.rune;
        });
    }
}

void huff_build_tree(std::vector<HuffLeaf> leaves, HuffTree &t) {
    huff_sort_leaves(leaves);
    const size_t k = leaves.size();
    t.nodes.clear();
    t.freq.clear();
    t.nodes.reserve(2 * k);
    t.freq.reserve(2 * k);
    t.n_leaves = k;
    for (size_t i = 0; i < k; i++) {
        t.nodes.push_back(HuffNode{-1, leaves[i].rune});
        t.freq.push_back(leaves[i].freq);
    }
    // all partial sums stay below 2^40 and node indices below 2^24: use the packed heap
    bool small = 2 * k < ((size_t)1 << 24);
    if (small) {
        uint64_t total = 0;
        for (const HuffLeaf &l : leaves) {
            if (l.freq < 0 || l.freq >= ((int64_t)1 << 39)) small = false;
            else total += (uint64_t)l.freq;
        }
        if (total >= ((uint64_t)1 << 39)) small = false;
    }
    if (small) {
        GoHeapT<uint64_t, LessPacked> hp;
        hp.h.resize(k);
        for (size_t i = 0; i < k; i++) hp.h[i] = ((uint64_t)leaves[i].freq << 24) | (uint64_t)i;
        hp.init();
        while (hp.h.size() > 1) {
            const uint64_t a = hp.pop();
            const uint64_t b = hp.pop();
            const uint64_t f = (a >> 24) + (b >> 24);
            t.nodes.push_back(HuffNode{(int32_t)(a & 0xFFFFFFu), (int32_t)(b & 0xFFFFFFu)});
            t.freq.push_back((int64_t)f);
            hp.push((f << 24) | (uint64_t)(t.nodes.size() - 1));
        }
        t.root = (int32_t)(hp.h[0] & 0xFFFFFFu);
        return;
    }
    GoHeapT<HeapEnt, LessEnt> hp;
    hp.h.resize(k);
    for (size_t i = 0; i < k; i++) hp.h[i] = HeapEnt{leaves[i].freq, (int32_t)i};
    hp.init();
    while (hp.h.size() > 1) {
        const HeapEnt a = hp.pop();
        const HeapEnt b = hp.pop();
        const int64_t f = (int64_t)((uint64_t)a.freq + (uint64_t)b.freq);  // wraps like Go's int
        t.nodes.push_back(HuffNode{a.node, b.node});
        t.freq.push_back(f);
        hp.push(HeapEnt{f, (int32_t)t.nodes.size() - 1});
    }
    t.root = hp.h[0].node;
}

bool huff_codes(const HuffTree &t, std::vector<HuffCode> &codes) {
    // printCodes (huffman.go:110-127): left appends '0', right appends '1'.  Every internal node was
    // created after its children (larger index), so one sweep from the root down assigns all paths.
    const size_t nn = t.nodes.size();
    std::vector<uint64_t> code(nn, 0);
    std::vector<uint32_t> depth(nn, 0);
    bool ok = true;
    for (size_t v = nn; v-- > t.n_leaves;) {
        const HuffNode &nd = t.nodes[v];
        code[nd.left] = code[v] << 1;
        code[nd.right] = (code[v] << 1) | 1;
        depth[nd.left] = depth[nd.right] = depth[v] + 1;
    }
    codes.clear();
    codes.reserve(t.n_leaves);
    for (size_t v = 0; v < t.n_leaves; v++) {
        if (depth[v] > 64) ok = false;
        codes.push_back(HuffCode{t.nodes[v].right, (uint8_t)std::min<uint32_t>(depth[v], 255), code[v], t.freq[v]});
    }
    return ok;
}

void huff_header(const std::vector<HuffLeaf> &leaves_in, std::vector<uint8_t> &hdr) {
    std::vector<HuffLeaf> lv(leaves_in);
    // the batch path hands the leaves over in ascending rune order already: nothing to sort then
    bool ascending = true;
    for (size_t i = 1; i < lv.size() && ascending; i++) ascending = lv[i - 1].rune < lv[i].rune;
    if (!ascending) {
        bool packable = true;
        for (const HuffLeaf &l : lv)
            if (l.freq < 0 || l.freq >= ((int64_t)1 << 42)) packable = false;
        if (packable) {
            std::vector<uint64_t> keys(lv.size());
            for (size_t i = 0; i < lv.size(); i++) keys[i] = ((uint64_t)(uint32_t)lv[i].rune << 42) | (uint64_t)lv[i].freq;
            radix_sort_u64(keys);
            for (size_t i = 0; i < lv.size(); i++)
                lv[i] = HuffLeaf{(int64_t)(keys[i] & (((uint64_t)1 << 42) - 1)), (int32_t)(keys[i] >> 42)};
        } else {
            std::sort(lv.begin(), lv.end(), [](const HuffLeaf &a, const HuffLeaf &b) { return a.rune < b.rune; });
        }
    }
    if (lv.size() >= 2 && lv.back().rune == 0x5C) std::swap(lv[lv.size() - 1], lv[lv.size() - 2]);
    // a record is at most 20 digits + '|' + 4 bytes: format into a buffer of that size and trim
    hdr.resize(lv.size() * 25);
    uint8_t *o = hdr.data();
    for (const HuffLeaf &l : lv) {
        uint64_t v = (uint64_t)l.freq;
        char tmp[24];
        int k = 0;
        do {
            tmp[k++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (k--) *o++ = (uint8_t)tmp[k];
        *o++ = '|';
        if (l.rune == 10) {  // huffman.go:315-317
            *o++ = '\\';
            *o++ = 'n';
        } else if (l.rune >= 0 && l.rune < 0x80) {
            *o++ = (uint8_t)l.rune;
        } else {
            o += utf8_encode(l.rune, o);
        }
    }
    hdr.resize((size_t)(o - hdr.data()));
}

// strconv.Atoi over a digit-only string, error dropped: empty => 0, overflow => MaxInt64.
static int64_t atoi_digits(const std::vector<uint8_t> &d) {
    if (d.empty()) return 0;
    const uint64_t maxv = ~0ull, cutoff = maxv / 10 + 1;
    uint64_t un = 0;
    for (uint8_t c : d) {
        if (un >= cutoff) {
            un = maxv;
            break;
        }
        un *= 10;
        uint64_t n1 = un + (uint64_t)(c - '0');
        if (n1 < un) {
            un = maxv;
            break;
        }
        un = n1;
    }
    if (un >= (1ull << 63)) return INT64_MAX;
    return (int64_t)un;
}

bool huff_parse_header(const uint8_t *h, size_t hn, std::vector<HuffLeaf> &leaves) {
    // rune -> freq with map semantics (the last assignment wins).  Records are collected in order
    // and de-duplicated afterwards: through a dense table over the rune space when the header is
    // huge (~1e5 records, binary input), else by sorting (rune, sequence number) keys — touching
    // 10 MB of table per call would dominate small files.
    std::vector<HuffLeaf> rec;
    std::vector<uint8_t> temp;
    rec.reserve(hn / 4 + 1);
    for (size_t i = 0; i < hn; i++) {
        if (h[i] != '|') {
            if (h[i] >= '0' && h[i] <= '9') temp.push_back(h[i]);
            continue;
        }
        const int64_t f = atoi_digits(temp);
        temp.clear();
        if (i + 1 >= hn) return false;  // tree[i+1] out of range
        int32_t sym;
        if (h[i + 1] == '\\') {
            if (i + 2 >= hn) return false;  // tree[i+2] out of range
            if (h[i + 2] == 'n') {
                sym = 10;
                i++;
            } else {
                sym = '\\';
            }
        } else {
            const size_t p = i + 1;
            const uint8_t b1 = p + 1 < hn ? h[p + 1] : 0, b2 = p + 2 < hn ? h[p + 2] : 0, b3 = p + 3 < hn ? h[p + 3] : 0;
            utf8_decode_at(h[p], b1, b2, b3, hn - p, &sym);
        }
        rec.push_back(HuffLeaf{f, sym});
        i++;
    }
    leaves.clear();
    const size_t k = rec.size();
    {
        // records in strictly ascending rune order (every header this library writes, bar the swap
        // that keeps a backslash off the end): no duplicates, and the order asked for
        bool ascending = true;
        for (size_t i = 1; i < k && ascending; i++) ascending = rec[i - 1].rune < rec[i].rune;
        if (ascending) {
            leaves.swap(rec);
            return true;
        }
    }
    if (k > ((size_t)1 << 16)) {
        std::vector<uint32_t> last(0x110000, 0);  // 1 + index of the last record of the rune
        for (size_t i = 0; i < k; i++) last[rec[i].rune] = (uint32_t)i + 1;
        for (size_t i = 0; i < k; i++)
            if (last[rec[i].rune] == (uint32_t)i + 1) leaves.push_back(rec[i]);
        return true;
    }
    std::vector<uint64_t> keys(k);
    for (size_t i = 0; i < k; i++) keys[i] = ((uint64_t)(uint32_t)rec[i].rune << 32) | (uint64_t)i;
    std::sort(keys.begin(), keys.end());
    leaves.reserve(k);
    for (size_t i = 0; i < k; i++)
        if (i + 1 == k || (keys[i + 1] >> 32) != (keys[i] >> 32)) leaves.push_back(rec[(size_t)(keys[i] & 0xFFFFFFFFu)]);
    return true;
}

}  // namespace rsn
