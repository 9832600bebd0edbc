// raisin_b200 — command line with the reference CLI's flags for this path (cmd/cli.go):
//   raisin_b200 [-algorithm=lzss,huffman] [-out=FILE] FILE...        compress to FILE.rsn
//   raisin_b200 -decompress [-algorithm=...] [-out=FILE] FILE.rsn    decompress
//   raisin_b200 -benchmark [-algorithm=lzss,huffman,[lzss,huffman]] FILE...
// In -benchmark, top-level commas separate independent runs and [a,b] is one layered run
// (cli.go:203-231).
#include <cstdio>
#include <cstring>

#include "raisin.hpp"

using namespace raisin;

static std::vector<std::vector<std::string>> parse_algorithms(const std::string &s) {  // cli.go:203-231
    std::vector<std::vector<std::string>> out;
    size_t i = 0;
    while (i < s.size()) {
        if (s[i] == '[') {
            const size_t j = s.find(']', i);
            out.push_back(engine::split(s.substr(i + 1, j - i - 1), ','));
            i = j == std::string::npos ? s.size() : j + 1;
            if (i < s.size() && s[i] == ',') i++;
        } else {
            size_t j = s.find(',', i);
            if (j == std::string::npos) j = s.size();
            out.push_back({s.substr(i, j - i)});
            i = j + 1;
        }
    }
    return out;
}

int main(int argc, char **argv) {
    std::string algorithm = "lzss,huffman", out;
    bool dec = false, bench = false;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a.rfind("-algorithm=", 0) == 0) algorithm = a.substr(11);
        else if (a.rfind("-out=", 0) == 0) out = a.substr(5);
        else if (a == "-decompress") dec = true;
        else if (a == "-compress") dec = false;
        else if (a == "-benchmark") bench = true;
        else files.push_back(a);
    }
    if (files.empty()) {
        fprintf(stderr, "usage: raisin_b200 [-compress|-decompress|-benchmark] [-algorithm=lzss,huffman] [-out=FILE] FILE...\n");
        return 2;
    }
    try {
        if (bench) {
            for (const auto &f : files) {
                printf("%-20s %-12s %-10s %-9s %-9s %s\n", "ENGINE", "TIME TAKEN", "RATIO", "ACT.ENT", "ENTROPY", "LOSSLESS");
                for (const auto &layers : parse_algorithms(algorithm)) {
                    const engine::Result r = engine::BenchmarkFile(layers, f);
                    if (r.Failed) printf("%-20s %-12s\n", r.CompressionEngine.c_str(), "DNF");
                    else printf("%-20s %-12s %-9.2f%% %-9.2f %-9.2f %s\n", r.CompressionEngine.c_str(), r.TimeTaken.c_str(),
                                r.Ratio, r.ActualEntropy, r.Entropy, r.Lossless ? "true" : "false");
                }
                printf("File %s\n", f.c_str());
            }
            return 0;
        }
        const auto layers = engine::split(algorithm, ',');
        for (const auto &f : files) {
            if (dec) engine::DecompressFile(layers, f, out.empty() ? f + ".decompressed" : out);
            else engine::CompressFile(layers, f, out.empty() ? f + ".rsn" : out);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "panic: %s\n", e.what());
        return 1;
    }
    return 0;
}
