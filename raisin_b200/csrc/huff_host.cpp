// huff_host.cpp — see huff_host.h.  Host-side product code (not the oracle).
#include "huff_host.h"

#include <algorithm>
#include <unordered_map>

#include "utf8.cuh"

namespace rsn {

namespace {
// container/heap (Go 1.15) on an index array; Less compares frequencies only (huffman.go:43-45),
// so the shape of the tree under ties is fixed by these exact sift rules.
struct GoHeap {
    const std::vector<int64_t> &f;
    std::vector<int32_t> h;
    explicit GoHeap(const std::vector<int64_t> &freq) : f(freq) {}
    bool less(int i, int j) const { return f[h[i]] < f[h[j]]; }
    void up(int j) {
        for (;;) {
            int i = (j - 1) / 2;  // parent; truncating division keeps j == 0 at 0
            if (i == j || !less(j, i)) break;
            std::swap(h[i], h[j]);
            j = i;
        }
    }
    void down(int i, int n) {
        for (;;) {
            int j1 = 2 * i + 1;
            if (j1 >= n || j1 < 0) break;
            int j = j1;
            if (j1 + 1 < n && less(j1 + 1, j1)) j = j1 + 1;
            if (!less(j, i)) break;
            std::swap(h[i], h[j]);
            i = j;
        }
    }
    void init() {
        int n = (int)h.size();
        for (int i = n / 2 - 1; i >= 0; i--) down(i, n);
    }
    int32_t pop() {
        int n = (int)h.size() - 1;
        std::swap(h[0], h[n]);
        down(0, n);
        int32_t v = h.back();
        h.pop_back();
        return v;
    }
    void push(int32_t v) {
        h.push_back(v);
        up((int)h.size() - 1);
    }
};
}  // namespace

void huff_build_tree(std::vector<HuffLeaf> leaves, HuffTree &t) {
    std::sort(leaves.begin(), leaves.end(), [](const HuffLeaf &a, const HuffLeaf &b) {
        return a.freq != b.freq ? a.freq < b.freq : a.rune < b.rune;
    });
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
    GoHeap hp(t.freq);
    hp.h.resize(k);
    for (size_t i = 0; i < k; i++) hp.h[i] = (int32_t)i;
    hp.init();
    while (hp.h.size() > 1) {
        int32_t a = hp.pop();
        int32_t b = hp.pop();
        t.nodes.push_back(HuffNode{a, b});
        t.freq.push_back((int64_t)((uint64_t)t.freq[a] + (uint64_t)t.freq[b]));
        hp.push((int32_t)t.nodes.size() - 1);
    }
    t.root = hp.h[0];
}

bool huff_codes(const HuffTree &t, std::vector<HuffCode> &codes) {
    codes.clear();
    codes.reserve(t.n_leaves);
    struct Item {
        int32_t node;
        uint32_t depth;
        uint64_t code;
    };
    std::vector<Item> stack;
    stack.push_back({t.root, 0, 0});
    bool ok = true;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        const HuffNode &nd = t.nodes[it.node];
        if (nd.left < 0) {
            if (it.depth > 64) ok = false;
            codes.push_back(HuffCode{nd.right, (uint8_t)std::min<uint32_t>(it.depth, 255), it.code, t.freq[it.node]});
            continue;
        }
        stack.push_back({nd.right, it.depth + 1, (it.code << 1) | 1});  // right appends '1'
        stack.push_back({nd.left, it.depth + 1, it.code << 1});          // left appends '0'
    }
    return ok;
}

static void put_dec(std::vector<uint8_t> &o, uint64_t v) {
    char tmp[24];
    int k = 0;
    do {
        tmp[k++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (k--) o.push_back((uint8_t)tmp[k]);
}

void huff_header(const std::vector<HuffLeaf> &leaves_in, std::vector<uint8_t> &hdr) {
    std::vector<HuffLeaf> lv(leaves_in);
    std::sort(lv.begin(), lv.end(), [](const HuffLeaf &a, const HuffLeaf &b) { return a.rune < b.rune; });
    if (lv.size() >= 2 && lv.back().rune == 0x5C) std::swap(lv[lv.size() - 1], lv[lv.size() - 2]);
    hdr.clear();
    for (const HuffLeaf &l : lv) {
        put_dec(hdr, (uint64_t)l.freq);
        hdr.push_back('|');
        if (l.rune == 10) {  // huffman.go:315-317
            hdr.push_back('\\');
            hdr.push_back('n');
        } else {
            uint8_t u[4];
            int w = utf8_encode(l.rune, u);
            hdr.insert(hdr.end(), u, u + w);
        }
    }
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
    std::unordered_map<int32_t, int64_t> m;
    std::vector<int32_t> order;
    std::vector<uint8_t> temp;
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
        if (m.find(sym) == m.end()) order.push_back(sym);
        m[sym] = f;
        i++;
    }
    leaves.clear();
    for (int32_t r : order) leaves.push_back(HuffLeaf{m[r], r});
    return true;
}

}  // namespace rsn
