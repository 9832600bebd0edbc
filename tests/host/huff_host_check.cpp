// Test harness (not shipped): drives the product's host-side Huffman logic (huff_host.cpp) from
// stdin/stdout so that the CPU test suite can compare it with the oracle without a GPU.
//   codes:  "<k>\n" then k lines "<rune> <freq>"   ->  k lines "<rune> <len> <code hex>", then "header <hex>"
//   parse:  "<hex of header bytes>"                 ->  "bad" | "ok <k>" and k lines "<rune> <freq>" (ascending rune)
#include "huff_host.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

using namespace rsn;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "codes")) {
        size_t k;
        std::cin >> k;
        std::vector<HuffLeaf> leaves(k);
        for (auto &l : leaves) {
            long long f;
            int r;
            std::cin >> r >> f;
            l = HuffLeaf{(int64_t)f, (int32_t)r};
        }
        HuffTree t;
        huff_build_tree(leaves, t);
        std::vector<HuffCode> codes;
        if (!huff_codes(t, codes)) {
            printf("toolong\n");
            return 0;
        }
        std::sort(codes.begin(), codes.end(), [](const HuffCode &a, const HuffCode &b) { return a.rune < b.rune; });
        for (auto &c : codes) printf("%d %u %llx\n", c.rune, (unsigned)c.len, (unsigned long long)c.code);
        std::vector<uint8_t> hdr;
        huff_header(leaves, hdr);
        printf("header ");
        for (uint8_t b : hdr) printf("%02x", b);
        printf("\n");
        return 0;
    }
    if (!strcmp(argv[1], "parse")) {
        std::string hex;
        std::cin >> hex;
        std::vector<uint8_t> h;
        for (size_t i = 0; i + 1 < hex.size(); i += 2) h.push_back((uint8_t)std::stoi(hex.substr(i, 2), nullptr, 16));
        std::vector<HuffLeaf> leaves;
        if (!huff_parse_header(h.data(), h.size(), leaves)) {
            printf("bad\n");
            return 0;
        }
        std::sort(leaves.begin(), leaves.end(), [](const HuffLeaf &a, const HuffLeaf &b) { return a.rune < b.rune; });
        printf("ok %zu\n", leaves.size());
        for (auto &l : leaves) printf("%d %lld\n", l.rune, (long long)l.freq);
        return 0;
    }
    return 2;
}
