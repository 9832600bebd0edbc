// huff_encode.cu — Huffman compress (huffman.Compress, huffman.go:299-325).
//
//   K8/K9  rune classify + histogram   (range-over-string at huffman.go:309-311)
//   host   tree + codes + header        (buildTree 58-103, printCodes 110-127, header 312-318)
//   K10    code-length scan + bit pack  (encode 229-256, AsByteSlice 174-191: MSB-first,
//                                        pad zero bits in FRONT of the first payload byte)
#include "batch.cuh"
#include "common.cuh"
#include "huff.cuh"
#include "huff_host.h"
#include "utf8.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace rsn {

constexpr uint32_t kRuneSpace = 0x110000;
constexpr int kSmallBins = 256;

// bytes base-3 .. base+18 around a thread's 16-byte chunk
__device__ __forceinline__ void load_chunk_halo(const uint8_t *__restrict__ in, size_t base, size_t n, uint8_t *w) {
    uint8_t v[16];
    load16(in, base, n, 0, v);
#pragma unroll
    for (int k = 0; k < 16; k++) w[k + 3] = v[k];
#pragma unroll
    for (int q = 1; q <= 3; q++) w[3 - q] = base >= (size_t)q ? __ldg(in + base - q) : 0;
#pragma unroll
    for (int q = 0; q < 3; q++) w[19 + q] = base + 16 + q < n ? __ldg(in + base + 16 + q) : 0;
}

// The rune starts of a thread's 16 bytes.  Fast path: a chunk of pure ASCII bytes (< 0x80) needs no
// neighbours: a byte outside [80,BF] always starts a rune and an ASCII byte is that rune.
// Returns the mask of starts; runes[k] is set for each start.
__device__ __forceinline__ uint32_t classify16(const uint8_t *__restrict__ in, size_t base, size_t n, int valid,
                                               int32_t (&runes)[16]) {
    uint32_t startmask = 0;
    if (valid == 16 && ((reinterpret_cast<uintptr_t>(in + base) & 15) == 0)) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in + base));
        if (((q.x | q.y | q.z | q.w) & 0x80808080u) == 0) {
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 16; k++) runes[k] = (int32_t)((w[k >> 2] >> ((k & 3) * 8)) & 0xFFu);
            return 0xFFFFu;
        }
    }
    uint8_t w[22];
    load_chunk_halo(in, base, n, w);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        runes[k] = 0;
        if (k < valid) {
            int32_t r;
            if (utf8_start_at(w, k, base + k, n, &r)) {
                startmask |= 1u << k;
                runes[k] = r;
            }
        }
    }
    return startmask;
}

// ============================================================================= K8/K9 histogram

constexpr int kHistCopies = 8;  // replicated shared bins: lanes of a warp spread over copies

__global__ void __launch_bounds__(kTileThreads) k_rune_hist(const uint8_t *__restrict__ in, size_t n, size_t tiles,
                                                            unsigned long long *__restrict__ hist) {
    __shared__ uint32_t bins[kHistCopies][kSmallBins];
    __shared__ uint32_t fffd_sm;
    for (int i = threadIdx.x; i < kHistCopies * kSmallBins; i += blockDim.x) (&bins[0][0])[i] = 0;
    if (threadIdx.x == 0) fffd_sm = 0;
    __syncthreads();
    uint32_t *mybins = bins[threadIdx.x & (kHistCopies - 1)];
    uint32_t fffd = 0;
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const size_t base = tile * kTile + (size_t)threadIdx.x * kItems;
        if (base >= n) continue;
        const int valid = (int)min((size_t)16, n - base);
        int32_t runes[16];
        const uint32_t startmask = classify16(in, base, n, valid, runes);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (startmask & (1u << k)) {
                const int32_t r = runes[k];
                if (r < kSmallBins) atomicAdd(&mybins[r], 1u);
                else if (r == 0xFFFD) fffd++;
                else atomicAdd(&hist[r], 1ull);
            }
        }
    }
    // U+FFFD dominates on non-UTF-8 input: keep it out of the atomic paths
    for (int d = 16; d; d >>= 1) fffd += __shfl_down_sync(0xffffffffu, fffd, d);
    if (lane_id() == 0 && fffd) atomicAdd(&fffd_sm, fffd);
    __syncthreads();
    for (int i = threadIdx.x; i < kSmallBins; i += blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int c = 0; c < kHistCopies; c++) v += bins[c][i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
    if (threadIdx.x == 0 && fffd_sm) atomicAdd(&hist[0xFFFD], (unsigned long long)fffd_sm);
}

struct RuneFreq {
    uint32_t rune;
    uint32_t pad;
    uint64_t freq;
};

__global__ void k_hist_compact(const unsigned long long *__restrict__ hist, RuneFreq *__restrict__ list,
                               unsigned long long *__restrict__ count) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= kRuneSpace) return;
    const unsigned long long f = hist[r];
    if (f) {
        const unsigned long long slot = atomicAdd(count, 1ull);
        list[slot] = RuneFreq{r, 0, f};
    }
}

// ============================================================================= K10 encode

// Codes of runes >= 256: a dense table over the rune space (single stream), or a small
// open-addressing table per file (batches; every rune looked up is present).
struct DenseBig {
    const uint64_t *code;
    const uint8_t *len;
    __device__ __forceinline__ uint32_t len_of(int32_t r) const { return __ldg(len + r); }
    __device__ __forceinline__ uint64_t code_of(int32_t r) const { return __ldg(code + r); }
};
__host__ __device__ __forceinline__ uint32_t big_hash(uint32_t rune) { return (rune * 2654435761u) >> 12; }
struct HashBig {
    const CodeEntry *tab;
    uint32_t mask;
    __device__ __forceinline__ const CodeEntry &find(int32_t r) const {
        uint32_t h = big_hash((uint32_t)r) & mask;
        while (tab[h].rune != (uint32_t)r) h = (h + 1) & mask;
        return tab[h];
    }
    __device__ __forceinline__ uint32_t len_of(int32_t r) const { return find(r).len; }
    __device__ __forceinline__ uint64_t code_of(int32_t r) const { return find(r).code; }
};

__global__ void k_code_scatter(const CodeEntry *__restrict__ list, size_t k, uint64_t *__restrict__ code,
                               uint8_t *__restrict__ len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    code[list[i].rune] = list[i].code;
    len[list[i].rune] = (uint8_t)list[i].len;
}

// per-tile total code bits
template <class Big>
__device__ __forceinline__ void enc_count_body(const uint8_t *__restrict__ in, size_t n,
                                               const uint8_t *__restrict__ len_tab, const Big big,
                                               uint64_t *__restrict__ tile_bits) {
    __shared__ uint8_t slen[kSmallBins];
    __shared__ uint32_t sm[33];
    for (int i = threadIdx.x; i < kSmallBins; i += blockDim.x) slen[i] = len_tab[i];
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint32_t bits = 0;
    if (base < n) {
        const int valid = (int)min((size_t)16, n - base);
        int32_t runes[16];
        const uint32_t startmask = classify16(in, base, n, valid, runes);
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (startmask & (1u << k)) bits += runes[k] < kSmallBins ? slen[runes[k]] : big.len_of(runes[k]);
    }
    uint32_t total;
    block_exclusive_sum<uint32_t>(bits, sm, total);
    if (threadIdx.x == 0) tile_bits[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kTileThreads) k_enc_count(const uint8_t *__restrict__ in, size_t n,
                                                            const uint8_t *__restrict__ len_tab,
                                                            const uint64_t *__restrict__ code_tab,
                                                            uint64_t *__restrict__ tile_bits) {
    enc_count_body(in, n, len_tab, DenseBig{code_tab, len_tab}, tile_bits);
}

// MSB-first bit writer over a zero-initialised, 4-byte aligned output: partially covered words
// are merged with atomicOr, fully covered words are stored directly.
struct BitWriter {
    uint32_t *out;
    uint64_t word;   // index of the word being filled
    uint32_t acc;    // bits so far, left-aligned
    uint32_t fill;   // bits used in acc
    bool first;      // current word may be shared with the previous writer
    __device__ __forceinline__ void init(uint32_t *o, uint64_t bitpos) {
        out = o;
        word = bitpos >> 5;
        fill = (uint32_t)(bitpos & 31);
        acc = 0;
        first = fill != 0;
    }
    __device__ __forceinline__ void flush_full() {
        const uint32_t be = __byte_perm(acc, 0, 0x0123);  // big-endian bit order in memory
        if (first) atomicOr(out + word, be);
        else out[word] = be;
        first = false;
        word++;
        acc = 0;
        fill = 0;
    }
    __device__ __forceinline__ void put(uint64_t code, uint32_t len) {
        while (len) {
            const uint32_t space = 32 - fill;
            const uint32_t take = len < space ? len : space;
            const uint32_t bits = (uint32_t)((code >> (len - take)) & ((take == 32) ? 0xFFFFFFFFull : ((1ull << take) - 1)));
            acc |= bits << (space - take);
            fill += take;
            len -= take;
            if (fill == 32) flush_full();
        }
    }
    __device__ __forceinline__ void finish() {
        if (fill) atomicOr(out + word, __byte_perm(acc, 0, 0x0123));
    }
};

// PACKED: every code is at most 24 bits long; runes < 256 then come from a shared table of
// (len << 24 | code) words and the bits of a thread are assembled in one 64-bit register.
template <bool PACKED, class Big>
__device__ __forceinline__ void enc_write_body(const uint8_t *__restrict__ in, size_t n,
                                               const uint64_t *__restrict__ code_tab,
                                               const uint8_t *__restrict__ len_tab, const Big big,
                                               const uint64_t *__restrict__ tile_bitoff, uint64_t bit_base,
                                               uint32_t *__restrict__ out) {
    __shared__ uint8_t slen[PACKED ? 1 : kSmallBins];
    __shared__ uint64_t scode[PACKED ? 1 : kSmallBins];
    __shared__ uint32_t spack[PACKED ? kSmallBins : 1];
    __shared__ uint32_t sm[33];
    for (int i = threadIdx.x; i < kSmallBins; i += blockDim.x) {
        if (PACKED) {
            spack[i] = ((uint32_t)len_tab[i] << 24) | (uint32_t)code_tab[i];
        } else {
            slen[i] = len_tab[i];
            scode[i] = code_tab[i];
        }
    }
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint32_t bits = 0, startmask = 0;
    int32_t runes[16];
    uint32_t ent[PACKED ? 16 : 1];
    if (base < n) {
        const int valid = (int)min((size_t)16, n - base);
        startmask = classify16(in, base, n, valid, runes);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (startmask & (1u << k)) {
                const int32_t r = runes[k];
                if (PACKED) {
                    ent[k] = r < kSmallBins ? spack[r] : ((big.len_of(r) << 24) | (uint32_t)big.code_of(r));
                    bits += ent[k] >> 24;
                } else {
                    bits += r < kSmallBins ? slen[r] : big.len_of(r);
                }
            }
        }
    }
    uint32_t total;
    const uint32_t pre = block_exclusive_sum<uint32_t>(bits, sm, total);
    if (!startmask) return;
    const uint64_t bitpos = bit_base + tile_bitoff[blockIdx.x] + pre;
    if (PACKED) {
        // acc holds `nb` pending bits in its low end; the first word may be shared with the previous
        // writer (its bits before `bitpos` stay zero here and are merged with atomicOr)
        uint64_t word = bitpos >> 5;
        uint32_t nb = (uint32_t)(bitpos & 31);
        uint64_t acc = 0;
        bool first = nb != 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (startmask & (1u << k)) {
                const uint32_t len = ent[k] >> 24;
                acc = (acc << len) | (uint64_t)(ent[k] & 0xFFFFFFu);
                nb += len;
                if (nb >= 32) {
                    const uint32_t w32 = (uint32_t)(acc >> (nb - 32));
                    const uint32_t be = __byte_perm(w32, 0, 0x0123);  // big-endian bit order in memory
                    if (first) atomicOr(out + word, be);
                    else out[word] = be;
                    first = false;
                    word++;
                    nb -= 32;
                }
            }
        }
        if (nb) atomicOr(out + word, __byte_perm((uint32_t)(acc << (32 - nb)), 0, 0x0123));
        return;
    }
    BitWriter bw;
    bw.init(out, bitpos);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (startmask & (1u << k)) {
            const int32_t r = runes[k];
            if (r < kSmallBins) bw.put(scode[r], slen[r]);
            else bw.put(big.code_of(r), big.len_of(r));
        }
    }
    bw.finish();
}
template <bool PACKED>
__global__ void __launch_bounds__(kTileThreads) k_enc_write(const uint8_t *__restrict__ in, size_t n,
                                                            const uint64_t *__restrict__ code_tab,
                                                            const uint8_t *__restrict__ len_tab,
                                                            const uint64_t *__restrict__ tile_bitoff,
                                                            uint64_t bit_base, uint32_t *__restrict__ out) {
    enc_write_body<PACKED>(in, n, code_tab, len_tab, DenseBig{code_tab, len_tab}, tile_bitoff, bit_base, out);
}

// ============================================================================= host orchestration

int huff_compress_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    ArenaScope scope(s);
    if (n == 0) return RSN_ERR_EMPTY_INPUT;
    const size_t tiles = div_up(n, kTile);

    Trace tr("hc", s);
    // ---- histogram
    DevBuf hist, list, count;
    RSN_TRY(hist.alloc((size_t)kRuneSpace * 8, s));
    RSN_TRY(count.alloc(16, s));
    RSN_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)kRuneSpace * 8, s));
    RSN_CUDA(cudaMemsetAsync(count.p, 0, 16, s));
    const unsigned hgrid = (unsigned)min(tiles, (size_t)148 * 8);
    RSN_LAUNCH(k_rune_hist, hgrid, kTileThreads, 0, s, d_in, n, tiles, hist.as<unsigned long long>());
    RSN_TRY(list.alloc((size_t)kRuneSpace * sizeof(RuneFreq), s));
    RSN_LAUNCH(k_hist_compact, (unsigned)div_up(kRuneSpace, 256), 256, 0, s, hist.as<unsigned long long>(),
               list.as<RuneFreq>(), count.as<unsigned long long>());
    uint64_t k = 0;
    RSN_TRY(read_u64(count.as<uint64_t>(), &k, s));
    HostVec<RuneFreq> h_list(k);  // pinned (see HostVec)
    if (!h_list.data()) return RSN_ERR_NOMEM;
    RSN_CUDA(cudaMemcpyAsync(h_list.data(), list.p, k * sizeof(RuneFreq), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    list.reset();

    tr.mark("hist+d2h");
    // ---- host: tree, codes, header (exactly as the reference builds them)
    std::vector<HuffLeaf> leaves(k);
    for (size_t i = 0; i < k; i++) leaves[i] = HuffLeaf{(int64_t)h_list[i].freq, (int32_t)h_list[i].rune};
    HuffTree tree;
    huff_build_tree(leaves, tree);
    std::vector<HuffCode> codes;
    if (!huff_codes(tree, codes)) return RSN_ERR_UNSUPPORTED;  // a code longer than 64 bits
    std::vector<uint8_t> hdr;
    huff_header(leaves, hdr);
    HostVec<CodeEntry> h_codes(codes.size());
    if (!h_codes.data()) return RSN_ERR_NOMEM;
    uint64_t total_bits = 0;
    uint32_t maxlen = 0;
    for (size_t i = 0; i < codes.size(); i++) {
        h_codes[i] = CodeEntry{(uint32_t)codes[i].rune, codes[i].len, codes[i].code};
        total_bits += (uint64_t)codes[i].len * (uint64_t)codes[i].freq;
        maxlen = std::max<uint32_t>(maxlen, codes[i].len);
    }
    const uint32_t pad = (uint32_t)((8 - total_bits % 8) % 8);  // huffman.go:245-249
    const size_t payload = (size_t)((total_bits + pad) / 8);
    const size_t prefix = hdr.size() + 3;  // header, 5C 0A, pad byte
    const size_t total = prefix + payload;

    tr.mark("host tree");
    // ---- device: code tables
    DevBuf code_tab, len_tab, clist;
    RSN_TRY(code_tab.alloc((size_t)kRuneSpace * 8, s));
    RSN_TRY(len_tab.alloc((size_t)kRuneSpace, s));
    RSN_TRY(clist.alloc(h_codes.size() * sizeof(CodeEntry) + 16, s));
    RSN_CUDA(cudaMemcpyAsync(clist.p, h_codes.data(), h_codes.size() * sizeof(CodeEntry), cudaMemcpyHostToDevice, s));
    RSN_LAUNCH(k_code_scatter, (unsigned)div_up(h_codes.size(), 256), 256, 0, s, clist.as<CodeEntry>(), h_codes.size(),
               code_tab.as<uint64_t>(), len_tab.as<uint8_t>());

    // ---- output buffer: zero, then header | 5C 0A | pad, then the bits
    DevBuf out;
    RSN_TRY(out.alloc_out(total + 16, s));
    RSN_CUDA(cudaMemsetAsync(out.p, 0, total + 16, s));
    HostVec<uint8_t> pre(hdr.size() + 3);
    if (!pre.data()) return RSN_ERR_NOMEM;
    memcpy(pre.data(), hdr.data(), hdr.size());
    pre[hdr.size()] = 0x5C;
    pre[hdr.size() + 1] = 0x0A;
    pre[hdr.size() + 2] = (uint8_t)pad;
    RSN_CUDA(cudaMemcpyAsync(out.p, pre.data(), pre.size(), cudaMemcpyHostToDevice, s));
    if (total_bits) {
        DevBuf tb, tbo;
        RSN_TRY(tb.alloc(tiles * 8, s));
        RSN_TRY(tbo.alloc((tiles + 1) * 8, s));
        RSN_LAUNCH(k_enc_count, (unsigned)tiles, kTileThreads, 0, s, d_in, n, len_tab.as<uint8_t>(),
                   code_tab.as<uint64_t>(), tb.as<uint64_t>());
        RSN_TRY(spine_scan_u64(tb.as<uint64_t>(), tbo.as<uint64_t>(), tbo.as<uint64_t>() + tiles, tiles, s));
        if (maxlen <= 24)
            RSN_LAUNCH(k_enc_write<true>, (unsigned)tiles, kTileThreads, 0, s, d_in, n, code_tab.as<uint64_t>(),
                       len_tab.as<uint8_t>(), tbo.as<uint64_t>(), (uint64_t)prefix * 8 + pad, out.as<uint32_t>());
        else
            RSN_LAUNCH(k_enc_write<false>, (unsigned)tiles, kTileThreads, 0, s, d_in, n, code_tab.as<uint64_t>(),
                       len_tab.as<uint8_t>(), tbo.as<uint64_t>(), (uint64_t)prefix * 8 + pad, out.as<uint32_t>());
    }
    tr.mark("encode");
    // h_codes / pre are read by async copies: wait before they go out of scope
    RSN_CUDA(stream_wait(s));
    *d_out = (uint8_t *)out.release();
    *out_n = total;
    return RSN_OK;
}

// ============================================================================= batches of small files
//
// One launch per kernel for a group of files (batch.cuh).  A dense table over the rune space per
// file would cost far more than the file itself, so: runes below 256 are counted in 256 bins per
// file, U+FFFD in a counter, and the remaining runes (multi-byte UTF-8 in binary-looking data) are
// appended to a per-file list that the host sorts and counts.  Trees, codes and headers are built on
// the host for every file (in parallel), the code tables of all files go up in one copy, codes of
// runes >= 256 in a small hash table per file.

constexpr int kHistTilesPerCta = 8;
constexpr int kHistStride = 260;  // 256 bins, [256] = U+FFFD

struct HencFile {
    const uint8_t *in;
    uint64_t n;
    uint32_t big_n;       // runes appended to the file's list (phase 1)
    uint32_t packed;      // every code is at most 24 bits long
    const uint64_t *scode;  // codes / lengths of runes < 256
    const uint8_t *slen;
    const CodeEntry *btab;  // hash table of the other runes
    uint32_t bmask;
    uint32_t pad;         // zero bits in front of the first payload byte
    uint32_t *out_words;  // 4-byte aligned start of the file's output
    uint64_t bit_base;    // first code bit: 8 * (header + 5C 0A + pad byte) + pad
    const uint8_t *prefix;  // header | 5C 0A
    uint64_t prefix_n;
};

struct HencBatch {
    HencFile *files;
    uint32_t *hist;        // [G][kHistStride]
    uint32_t *big;         // [G][big_stride]
    size_t big_stride;
    uint64_t *tile_bits, *tile_off;  // [G][tc_stride]
    size_t tc_stride;
};

// the per-file lists of runes >= 256 (each at its own stride) back to back: one device-to-host copy
// for the group instead of one per file
__global__ void __launch_bounds__(256) kb_big_gather(HencBatch b, const uint64_t *__restrict__ at,
                                                     uint32_t *__restrict__ dst) {
    const uint32_t n = b.files[blockIdx.y].big_n;
    const uint32_t *src = b.big + (size_t)blockIdx.y * b.big_stride;
    uint32_t *d = dst + at[blockIdx.y];
    for (uint32_t i = blockIdx.x * 1024u + threadIdx.x; i < min(n, (blockIdx.x + 1) * 1024u); i += 256) d[i] = src[i];
}

__global__ void __launch_bounds__(kTileThreads) kb_rune_hist(HencBatch b) {
    __shared__ uint32_t bins[kHistCopies][kSmallBins];
    __shared__ uint32_t fffd_sm;
    HencFile &f = b.files[blockIdx.y];
    const size_t n = (size_t)f.n;
    const size_t tile0 = (size_t)blockIdx.x * kHistTilesPerCta;
    if (tile0 * kTile >= n) return;
    for (int i = threadIdx.x; i < kHistCopies * kSmallBins; i += blockDim.x) (&bins[0][0])[i] = 0;
    if (threadIdx.x == 0) fffd_sm = 0;
    __syncthreads();
    uint32_t *mybins = bins[threadIdx.x & (kHistCopies - 1)];
    uint32_t *biglist = b.big + (size_t)blockIdx.y * b.big_stride;
    uint32_t fffd = 0;
    for (int q = 0; q < kHistTilesPerCta; q++) {
        const size_t base = (tile0 + q) * kTile + (size_t)threadIdx.x * kItems;
        if (base >= n) break;
        const int valid = (int)min((size_t)16, n - base);
        int32_t runes[16];
        const uint32_t startmask = classify16(f.in, base, n, valid, runes);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (startmask & (1u << k)) {
                const int32_t r = runes[k];
                if (r < kSmallBins) atomicAdd(&mybins[r], 1u);
                else if (r == 0xFFFD) fffd++;
                else biglist[atomicAdd(&f.big_n, 1u)] = (uint32_t)r;
            }
        }
    }
    for (int d = 16; d; d >>= 1) fffd += __shfl_down_sync(0xffffffffu, fffd, d);
    if (lane_id() == 0 && fffd) atomicAdd(&fffd_sm, fffd);
    __syncthreads();
    uint32_t *hist = b.hist + (size_t)blockIdx.y * kHistStride;
    for (int i = threadIdx.x; i < kSmallBins; i += blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int c = 0; c < kHistCopies; c++) v += bins[c][i];
        if (v) atomicAdd(&hist[i], v);
    }
    if (threadIdx.x == 0 && fffd_sm) atomicAdd(&hist[256], fffd_sm);
}

__global__ void __launch_bounds__(256) kb_enc_prefix(HencBatch b) {
    const HencFile &f = b.files[blockIdx.x];
    uint8_t *dst = reinterpret_cast<uint8_t *>(f.out_words);
    for (uint64_t i = threadIdx.x; i < f.prefix_n; i += blockDim.x) dst[i] = f.prefix[i];
    if (threadIdx.x == 0 && f.prefix_n) dst[f.prefix_n] = (uint8_t)f.pad;  // huffman.go:250
}

__global__ void __launch_bounds__(kTileThreads) kb_enc_count(HencBatch b) {
    const HencFile &f = b.files[blockIdx.y];
    if ((size_t)blockIdx.x * kTile >= f.n) return;
    enc_count_body(f.in, (size_t)f.n, f.slen, HashBig{f.btab, f.bmask}, b.tile_bits + (size_t)blockIdx.y * b.tc_stride);
}

__global__ void __launch_bounds__(256) kb_enc_finish(HencBatch b) {
    __shared__ uint64_t sm[33];
    const HencFile &f = b.files[blockIdx.x];
    const size_t o = (size_t)blockIdx.x * b.tc_stride;
    cta_scan_u64(b.tile_bits + o, b.tile_off + o, div_up_dev((size_t)f.n, (size_t)kTile), sm);
}

template <bool PACKED>
__global__ void __launch_bounds__(kTileThreads) kb_enc_write(HencBatch b) {
    const HencFile &f = b.files[blockIdx.y];
    if ((f.packed != 0) != PACKED || (size_t)blockIdx.x * kTile >= f.n) return;
    enc_write_body<PACKED>(f.in, (size_t)f.n, f.scode, f.slen, HashBig{f.btab, f.bmask},
                           b.tile_off + (size_t)blockIdx.y * b.tc_stride, f.bit_base, f.out_words);
}

namespace {
struct HencHost {
    int rc = RSN_OK;
    bool per_file = false;            // alphabet too large for the device tree builder
    std::vector<uint8_t> prefix;      // header | 5C 0A (the pad byte follows once the bit count is known)
    std::vector<uint32_t> freq, rune; // leaves in (freq asc, rune asc) order
    size_t btab_cap = 2;              // power-of-two hash table size for the runes >= 256
};

// runes are below 2^21: three byte-wise counting passes
void sort_runes(uint32_t *v, size_t n) {
    if (n < 256) {
        std::sort(v, v + n);
        return;
    }
    std::vector<uint32_t> tmp(n);
    uint32_t *src = v, *dst = tmp.data();
    for (int byte = 0; byte < 3; byte++) {
        size_t cnt[257] = {0};
        for (size_t i = 0; i < n; i++) cnt[((src[i] >> (8 * byte)) & 0xFF) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; i++) dst[cnt[(src[i] >> (8 * byte)) & 0xFF]++] = src[i];
        std::swap(src, dst);
    }
    if (src != v) std::copy(src, src + n, v);
}

// histogram -> leaves in the reference's order, header bytes (huffman.go:312-318)
void henc_host_plan(const uint32_t *hist, uint32_t *big, size_t big_n, HencHost &pl) {
    std::vector<HuffLeaf> leaves;
    for (int r = 0; r < kSmallBins; r++)
        if (hist[r]) leaves.push_back(HuffLeaf{(int64_t)hist[r], r});
    sort_runes(big, big_n);
    size_t nbig = 0;
    bool fffd_pending = hist[256] != 0;  // U+FFFD is counted apart; it goes in at its place in rune order
    for (size_t i = 0; i < big_n;) {
        size_t j = i;
        while (j < big_n && big[j] == big[i]) j++;
        if (fffd_pending && big[i] > 0xFFFDu) {
            leaves.push_back(HuffLeaf{(int64_t)hist[256], 0xFFFD});
            nbig++;
            fffd_pending = false;
        }
        leaves.push_back(HuffLeaf{(int64_t)(j - i), (int32_t)big[i]});
        nbig++;
        i = j;
    }
    if (fffd_pending) {
        leaves.push_back(HuffLeaf{(int64_t)hist[256], 0xFFFD});
        nbig++;
    }
    if (leaves.empty()) {
        pl.rc = RSN_ERR_EMPTY_INPUT;
        return;
    }
    if (leaves.size() > kTreeMaxLeaves) {
        pl.per_file = true;
        return;
    }
    huff_header(leaves, pl.prefix);
    pl.prefix.push_back(0x5C);
    pl.prefix.push_back(0x0A);
    huff_sort_leaves(leaves);
    pl.freq.resize(leaves.size());
    pl.rune.resize(leaves.size());
    for (size_t i = 0; i < leaves.size(); i++) {
        pl.freq[i] = (uint32_t)leaves[i].freq;
        pl.rune[i] = (uint32_t)leaves[i].rune;
    }
    while (pl.btab_cap < 2 * nbig) pl.btab_cap <<= 1;
}
}  // namespace

int huff_compress_batch(const BatchIO &in, BatchIO &out, cudaStream_t s) {
    const size_t G = in.size();
    out.resize(G);
    out.rc = in.rc;
    if (G == 0) return RSN_OK;
    ArenaScope scope(s);
    Trace tr("hc batch", s);
    size_t cap = 1;
    for (size_t f = 0; f < G; f++) {
        if (in.rc[f] != RSN_OK) continue;
        if (in.n[f] == 0) out.rc[f] = RSN_ERR_EMPTY_INPUT;  // heap.Pop on an empty heap panics
        cap = std::max<size_t>(cap, in.n[f]);
    }
    if (cap > kBatchMaxFile) return RSN_ERR_UNSUPPORTED;
    const size_t tiles_cap = div_up(cap, kTile);
    HencBatch b{};
    b.big_stride = cap / 2 + 16;
    b.tc_stride = tiles_cap + 1;
    HostVec<HencFile> h(G);
    if (!h.data()) return RSN_ERR_NOMEM;
    for (size_t f = 0; f < G; f++) {
        h[f] = HencFile{};
        h[f].in = in.ptr[f];
        h[f].n = out.rc[f] == RSN_OK ? in.n[f] : 0;
    }
    DevBuf files, hist, big, tb, tbo;
    RSN_TRY(files.alloc(G * sizeof(HencFile), s));
    RSN_TRY(hist.alloc(G * kHistStride * 4, s));
    RSN_TRY(big.alloc(G * b.big_stride * 4, s));
    RSN_TRY(tb.alloc(G * b.tc_stride * 8, s));
    RSN_TRY(tbo.alloc(G * b.tc_stride * 8, s));
    b.files = files.as<HencFile>();
    b.hist = hist.as<uint32_t>();
    b.big = big.as<uint32_t>();
    b.tile_bits = tb.as<uint64_t>();
    b.tile_off = tbo.as<uint64_t>();
    const unsigned g = (unsigned)G;
    // ---- phase 1: histograms
    RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(HencFile), cudaMemcpyHostToDevice, s));
    RSN_CUDA(cudaMemsetAsync(hist.p, 0, G * kHistStride * 4, s));
    RSN_LAUNCH(kb_rune_hist, dim3((unsigned)div_up(tiles_cap, kHistTilesPerCta), g), kTileThreads, 0, s, b);
    HostVec<uint32_t> h_hist(G * kHistStride);
    if (!h_hist.data()) return RSN_ERR_NOMEM;
    RSN_CUDA(cudaMemcpyAsync(h_hist.data(), hist.p, G * kHistStride * 4, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(HencFile), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    tr.mark("hist");
    // the per-file lists of runes >= 256, back to back in one pinned buffer
    HostVec<uint64_t> big_at(G + 1);
    if (!big_at.data()) return RSN_ERR_NOMEM;
    uint32_t big_max = 0;
    big_at[0] = 0;
    for (size_t f = 0; f < G; f++) {
        big_at[f + 1] = big_at[f] + h[f].big_n;
        big_max = std::max(big_max, h[f].big_n);
    }
    HostVec<uint32_t> h_big(big_at[G] + 1);
    if (!h_big.data()) return RSN_ERR_NOMEM;
    if (big_at[G]) {
        DevBuf d_at, d_all;
        RSN_TRY(d_at.alloc((G + 1) * 8, s));
        RSN_TRY(d_all.alloc(big_at[G] * 4, s));
        RSN_CUDA(cudaMemcpyAsync(d_at.p, big_at.data(), (G + 1) * 8, cudaMemcpyHostToDevice, s));
        RSN_LAUNCH(kb_big_gather, dim3((unsigned)div_up(big_max, 1024), g), 256, 0, s, b, d_at.as<uint64_t>(),
                   d_all.as<uint32_t>());
        RSN_CUDA(cudaMemcpyAsync(h_big.data(), d_all.p, big_at[G] * 4, cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
    }
    tr.mark("big lists d2h");
    // ---- host: leaves in the reference's order and header bytes per file; trees and codes on the
    // device, one warp per file (huff_tree.cu)
    std::vector<HencHost> plan(G);
    parallel_for(G, batch_host_threads(), [&](size_t f) {
        if (out.rc[f] != RSN_OK) {
            plan[f].rc = out.rc[f];
            return;
        }
        henc_host_plan(h_hist.data() + f * kHistStride, h_big.data() + big_at[f], big_at[f + 1] - big_at[f], plan[f]);
    });
    tr.mark("host leaves+headers");
    static thread_local HostVec<uint8_t> tab;  // pinned staging: kept between groups
    std::vector<size_t> o_freq(G), o_rune(G), o_pre(G), o_small(G), o_btab(G), o_nodes(G), o_parent(G);
    size_t tab_n = 0, zero_n = 0, ff_n = 0, scr_n = 0;
    auto room = [](size_t &cursor, size_t bytes) {
        const size_t at = (cursor + 15) & ~(size_t)15;
        cursor = at + bytes;
        return at;
    };
    uint32_t kmax = 1;
    for (size_t f = 0; f < G; f++) {
        out.rc[f] = plan[f].rc;
        if (plan[f].rc != RSN_OK || plan[f].per_file) continue;
        const size_t k = plan[f].freq.size();
        kmax = std::max<uint32_t>(kmax, (uint32_t)k);
        o_freq[f] = room(tab_n, k * 4);
        o_rune[f] = room(tab_n, k * 4);
        o_pre[f] = room(tab_n, plan[f].prefix.size());
        o_small[f] = room(zero_n, kSmallBins * 8 + kSmallBins);
        o_btab[f] = room(ff_n, plan[f].btab_cap * sizeof(CodeEntry));
        o_nodes[f] = room(scr_n, (2 * k) * sizeof(HuffNodeDev));
        o_parent[f] = room(scr_n, (2 * k) * 4);
    }
    if (tab.size() < tab_n + 256 && !tab.resize(tab_n + 256)) return RSN_ERR_NOMEM;
    uint8_t *const tabp = tab.data();  // (a thread_local name inside the lambda would be the helper thread's own)
    parallel_for(G, batch_host_threads(), [&, tabp](size_t f) {
        if (plan[f].rc != RSN_OK || plan[f].per_file) return;
        memcpy(tabp + o_freq[f], plan[f].freq.data(), plan[f].freq.size() * 4);
        memcpy(tabp + o_rune[f], plan[f].rune.data(), plan[f].rune.size() * 4);
        memcpy(tabp + o_pre[f], plan[f].prefix.data(), plan[f].prefix.size());
    });
    DevBuf dtab, dzero, dff, dscr, djobs;
    RSN_TRY(dtab.alloc(tab_n + 256, s));
    RSN_TRY(dzero.alloc(zero_n + 256, s));
    RSN_TRY(dff.alloc(ff_n + 256, s));
    RSN_TRY(dscr.alloc(scr_n + 256, s));
    RSN_TRY(djobs.alloc(G * sizeof(TreeJob), s));
    HostVec<TreeJob> jobs(G);
    if (!jobs.data()) return RSN_ERR_NOMEM;
    for (size_t f = 0; f < G; f++) {
        TreeJob &j = jobs[f];
        j = TreeJob{};
        if (plan[f].rc != RSN_OK || plan[f].per_file) continue;
        j.freq = reinterpret_cast<const uint32_t *>(dtab.as<uint8_t>() + o_freq[f]);
        j.rune = reinterpret_cast<const uint32_t *>(dtab.as<uint8_t>() + o_rune[f]);
        j.k = (uint32_t)plan[f].freq.size();
        j.bmask = (uint32_t)(plan[f].btab_cap - 1);
        j.nodes = reinterpret_cast<HuffNodeDev *>(dscr.as<uint8_t>() + o_nodes[f]);
        j.parent = reinterpret_cast<uint32_t *>(dscr.as<uint8_t>() + o_parent[f]);
        j.scode = reinterpret_cast<uint64_t *>(dzero.as<uint8_t>() + o_small[f]);
        j.slen = dzero.as<uint8_t>() + o_small[f] + kSmallBins * 8;
        j.btab = reinterpret_cast<CodeEntry *>(dff.as<uint8_t>() + o_btab[f]);
    }
    if (tab_n) RSN_CUDA(cudaMemcpyAsync(dtab.p, tabp, tab_n, cudaMemcpyHostToDevice, s));
    RSN_CUDA(cudaMemcpyAsync(djobs.p, jobs.data(), G * sizeof(TreeJob), cudaMemcpyHostToDevice, s));
    RSN_CUDA(cudaMemsetAsync(dzero.p, 0, zero_n + 256, s));
    RSN_CUDA(cudaMemsetAsync(dff.p, 0xFF, ff_n + 256, s));
    RSN_TRY(huff_tree_batch(djobs.as<TreeJob>(), G, kmax, s));
    RSN_CUDA(cudaMemcpyAsync(jobs.data(), djobs.p, G * sizeof(TreeJob), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    tr.mark("device trees");
    // ---- sizes, one result buffer
    size_t total = 0;
    std::vector<size_t> out_base(G, 0), out_total(G, 0);
    bool any_packed = false, any_wide = false;
    for (size_t f = 0; f < G; f++) {
        HencFile &r = h[f];
        if (plan[f].rc != RSN_OK || plan[f].per_file) {
            r.n = 0;
            continue;
        }
        const TreeJob &j = jobs[f];
        if (j.flags & 1u) {  // a code longer than 64 bits
            out.rc[f] = RSN_ERR_UNSUPPORTED;
            r.n = 0;
            continue;
        }
        const uint32_t pad = (uint32_t)((8 - j.total_bits % 8) % 8);  // huffman.go:245-249
        out_total[f] = plan[f].prefix.size() + 1 + (size_t)((j.total_bits + pad) / 8);
        out_base[f] = total;
        total += (out_total[f] + 16 + 255) & ~(size_t)255;
        r.packed = j.maxlen <= 24;
        (r.packed ? any_packed : any_wide) = true;
        r.scode = j.scode;
        r.slen = j.slen;
        r.btab = j.btab;
        r.bmask = j.bmask;
        r.pad = pad;
        r.bit_base = (uint64_t)(plan[f].prefix.size() + 1) * 8 + pad;
        r.prefix = dtab.as<uint8_t>() + o_pre[f];
        r.prefix_n = plan[f].prefix.size();
    }
    DevBuf res;
    RSN_TRY(res.alloc_out(total + 256, s));
    for (size_t f = 0; f < G; f++)
        if (h[f].n) h[f].out_words = reinterpret_cast<uint32_t *>(res.as<uint8_t>() + out_base[f]);
    RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(HencFile), cudaMemcpyHostToDevice, s));
    RSN_CUDA(cudaMemsetAsync(res.p, 0, total + 256, s));
    tr.mark("tables");
    // ---- phase 2: code-length scan and bit pack
    const dim3 tgrid((unsigned)tiles_cap, g);
    RSN_LAUNCH(kb_enc_prefix, g, 256, 0, s, b);
    RSN_LAUNCH(kb_enc_count, tgrid, kTileThreads, 0, s, b);
    RSN_LAUNCH(kb_enc_finish, g, 256, 0, s, b);
    if (any_packed) RSN_LAUNCH(kb_enc_write<true>, tgrid, kTileThreads, 0, s, b);
    if (any_wide) RSN_LAUNCH(kb_enc_write<false>, tgrid, kTileThreads, 0, s, b);
    RSN_CUDA(stream_wait(s));  // h is read by the copy above
    tr.mark("encode");
    for (size_t f = 0; f < G; f++) {
        if (out.rc[f] != RSN_OK || plan[f].per_file) continue;
        out.ptr[f] = res.as<uint8_t>() + out_base[f];
        out.n[f] = out_total[f];
    }
    out.spans.push_back({res.as<uint8_t>(), res.bytes});
    out.owned.push_back(res.release());
    // alphabets beyond the device tree builder: the single-stream call
    for (size_t f = 0; f < G; f++) {
        if (out.rc[f] != RSN_OK || !plan[f].per_file) continue;
        uint8_t *r = nullptr;
        size_t rn = 0;
        out.rc[f] = huff_compress_dev(in.ptr[f], (size_t)in.n[f], &r, &rn, s);
        if (out.rc[f] != RSN_OK) continue;
        out.ptr[f] = r;
        out.n[f] = rn;
        out.owned.push_back(r);
    }
    return RSN_OK;
}

}  // namespace rsn
