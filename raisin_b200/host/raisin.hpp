// raisin.hpp — C++ host-side mirror of the reference's Go packages for the LZSS/Huffman path,
// over the C ABI (include/raisin_b200.h).  The reference is compiled Go and this image has no Go
// toolchain, so this header plays the role the cgo shim (go/) plays in a real raisin checkout:
// same names, argument meaning and failure behaviour (a Go panic becomes a C++ exception).
//
//   raisin::lz::CompressAsync / Compress / Decompress          compressor/lz/lzss.go:109, 224, 323
//   raisin::huffman::Compress / Decompress                      compressor/huffman/huffman.go:299, 327
//   raisin::engine::compress / decompress                       engine/engine.go:443-479
//   raisin::engine::CompressFile / DecompressFile (.rsn)        engine/engine.go:157-199
//   raisin::engine::BenchmarkFile                               engine/engine.go:357-441
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/raisin_b200.h"

namespace raisin {

using Bytes = std::vector<uint8_t>;

struct Panic : std::runtime_error {  // the reference signals every failure on this path by panicking
    int rc;
    Panic(int code, const std::string &what) : std::runtime_error(what), rc(code) {}
};

namespace detail {
inline Bytes take(int rc, uint8_t *out, size_t n, const char *who) {
    if (rc != RSN_OK) throw Panic(rc, std::string(who) + ": " + rsn_strerror(rc) + " " + rsn_last_cuda_error());
    Bytes b(out, out + n);
    rsn_free(out);
    return b;
}
}  // namespace detail

namespace lz {
constexpr int DefaultWindowSize = 4096;  // lzss.go:35
inline Bytes CompressAsync(const Bytes &fileContents, bool /*useProgressBar*/, int maxSearchBufferLength) {
    uint8_t *out = nullptr;
    size_t n = 0;
    int rc = rsn_lzss_compress(fileContents.data(), fileContents.size(), maxSearchBufferLength, RSN_LZSS_ASYNC, &out, &n);
    return detail::take(rc, out, n, "lzss");
}
inline Bytes Compress(const Bytes &fileContents, bool /*useProgressBar*/, int maxSearchBufferLength) {
    uint8_t *out = nullptr;
    size_t n = 0;
    int rc = rsn_lzss_compress(fileContents.data(), fileContents.size(), maxSearchBufferLength, RSN_LZSS_ITER, &out, &n);
    return detail::take(rc, out, n, "lzss");
}
inline Bytes Decompress(const Bytes &fileContents, bool /*useProgressBar*/) {
    uint8_t *out = nullptr;
    size_t n = 0;
    int rc = rsn_lzss_decompress(fileContents.data(), fileContents.size(), &out, &n);
    return detail::take(rc, out, n, "lzss");
}
}  // namespace lz

namespace huffman {
inline Bytes Compress(const Bytes &fileContents) {
    uint8_t *out = nullptr;
    size_t n = 0;
    int rc = rsn_huff_compress(fileContents.data(), fileContents.size(), &out, &n);
    return detail::take(rc, out, n, "huffman");
}
inline Bytes Decompress(const Bytes &fileContents, bool strict_limits = false) {
    uint8_t *out = nullptr;
    size_t n = 0;
    int rc = rsn_huff_decompress(fileContents.data(), fileContents.size(), strict_limits ? 1 : 0, &out, &n);
    return detail::take(rc, out, n, "huffman");
}
}  // namespace huffman

namespace engine {

inline std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, sep)) out.push_back(item);
    return out;
}

// engine.go:113-139 via the Writers registry: lz.NewWriter -> CompressAsync(window 4096), huffman.NewWriter -> Compress
inline Bytes write_layer(const std::string &algorithm, const Bytes &content) {
    if (algorithm == "lzss") return lz::CompressAsync(content, true, lz::DefaultWindowSize);
    if (algorithm == "huffman") return huffman::Compress(content);
    throw Panic(RSN_ERR_INVALID_ARG, "unknown algorithm " + algorithm);
}
inline Bytes read_layer(const std::string &algorithm, const Bytes &content) {
    if (algorithm == "lzss") return lz::Decompress(content, true);
    if (algorithm == "huffman") return huffman::Decompress(content);
    throw Panic(RSN_ERR_INVALID_ARG, "unknown algorithm " + algorithm);
}
inline Bytes compress(Bytes content, const std::vector<std::string> &algorithms) {  // engine.go:443-452
    for (const auto &a : algorithms) content = write_layer(a, content);
    return content;
}
inline Bytes decompress(Bytes content, const std::vector<std::string> &algorithms) {  // engine.go:454-479
    for (size_t i = algorithms.size(); i-- > 0;) content = read_layer(algorithms[i], content);
    return content;
}

inline Bytes read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    return Bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline void write_file(const std::string &path, const Bytes &b) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(b.data()), (std::streamsize)b.size());
}
inline void CompressFile(const std::vector<std::string> &algorithms, const std::string &path, const std::string &output) {
    write_file(output, compress(read_file(path), algorithms));  // engine.go:157-166, no framing
}
inline void DecompressFile(const std::vector<std::string> &algorithms, const std::string &path, const std::string &output) {
    write_file(output, decompress(read_file(path), algorithms));
}

struct Result {  // engine.go:201-209
    std::string CompressionEngine, TimeTaken;
    float Ratio = 0, ActualEntropy = 0;
    double Entropy = 0;
    bool Lossless = false, Failed = false;
    double Seconds = 0;
};

inline double entropy(const Bytes &b, size_t total) {
    size_t h[256] = {0};
    for (uint8_t c : b) h[c]++;
    double e = 0;
    for (size_t c : h)
        if (c) {
            const double p = (double)c / (double)total;
            e -= p * std::log(p);
        }
    return e;
}

// engine.go:357-441: timed region = compress + decompress (wall clock)
inline Result BenchmarkFile(const std::vector<std::string> &algorithms, const std::string &path) {
    Result r;
    for (size_t i = 0; i < algorithms.size(); i++) r.CompressionEngine += (i ? "," : "") + algorithms[i];
    try {
        const Bytes fileContents = read_file(path);
        r.Entropy = fileContents.empty() ? 0 : entropy(fileContents, fileContents.size());
        const auto t0 = std::chrono::steady_clock::now();
        const Bytes compressed = compress(fileContents, algorithms);
        const Bytes decompressed = decompress(compressed, algorithms);
        r.Seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        r.Lossless = decompressed == fileContents;
        r.Ratio = (float)compressed.size() / (float)fileContents.size() * 100.f;
        r.ActualEntropy = compressed.empty() ? 0.f : (float)entropy(decompressed, compressed.size());  // sic, engine.go:412-423
        char buf[64];
        snprintf(buf, sizeof buf, "%.2fms", r.Seconds * 1e3);
        r.TimeTaken = buf;
    } catch (const Panic &) {  // AsyncBenchmarkFile recovers panics into a Failed row (engine.go:315-328)
        r.Failed = true;
        r.TimeTaken = "DNF";
    }
    return r;
}

}  // namespace engine
}  // namespace raisin
