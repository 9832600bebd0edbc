// huff_decode.cu — Huffman decompress (huffman.Decompress, huffman.go:327-330, 258-297).
//
// The reference stream is ONE bit string with no block structure, so parallel decode has to
// find code boundaries by itself: every thread decodes its own fixed-size subsequence from a
// guessed start and we iterate "start[t] = end[t-1]" to a fixed point (self-synchronisation).
// Subsequence 0 starts at bit 0, which is exact, so the fixed point is the sequential decode
// of findCodes (huffman.go:131-153): walk from the root, left on 0 / right on 1, emit at a
// leaf, restart at the root while bits remain, fail if the bits end inside a code.
#include "batch.cuh"
#include "common.cuh"
#include "huff.cuh"
#include "huff_host.h"
#include "utf8.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace rsn {

constexpr uint32_t kSubBits = 256;  // bits per subsequence (one thread)

constexpr int kLutBits = 12;

struct DecParams {
    const uint32_t *words;  // payload as 4-byte words from an aligned base
    uint64_t nwords;        // words that may be read (the rest reads as 0)
    uint64_t bit0;          // bit offset of code bit 0 from that base (alignment slack + pad bits)
    uint64_t max;           // number of code bits
    const HuffNode *nodes;
    const uint32_t *lut;    // 2^kLutBits entries: leaf -> 1<<31 | len<<21 | rune ; else node index after kLutBits bits
    int32_t root;
};

// Forward bit reader over the big-endian payload: `buf` holds the next `avail` stream bits in its
// high end (MSB first); refilled 32 bits at a time from aligned words.
struct BitReader {
    const uint32_t *words;
    uint64_t nwords, next;  // next word index to append
    uint64_t buf;
    uint32_t avail;
    __device__ __forceinline__ uint32_t word(uint64_t w) const {
        return w < nwords ? __byte_perm(__ldg(words + w), 0, 0x0123) : 0u;
    }
    __device__ __forceinline__ void seek(const DecParams &p, uint64_t pos);
    __device__ __forceinline__ void refill() {
        if (avail <= 32) {
            buf |= (uint64_t)word(next) << (32 - avail);
            avail += 32;
            next++;
        }
    }
    __device__ __forceinline__ uint32_t peek(int bits) const { return (uint32_t)(buf >> (64 - bits)); }
    __device__ __forceinline__ void skip(uint32_t bits) {
        buf <<= bits;
        avail -= bits;
    }
};

__device__ __forceinline__ void BitReader::seek(const DecParams &p, uint64_t pos) {
    words = p.words;
    nwords = p.nwords;
    const uint64_t b = p.bit0 + pos;
    const uint64_t w = b >> 5;
    const uint32_t sh = (uint32_t)(b & 31);
    buf = (((uint64_t)word(w) << 32) | (uint64_t)word(w + 1)) << sh;
    avail = 64 - sh;
    next = w + 2;
}

// UTF-8 bytes to a byte address with word stores once the address is 4-byte aligned.
struct ByteWriter {
    uint8_t *p;
    uint32_t acc;  // pending bytes, little-endian
    uint32_t k;    // number of pending bytes (only while p is 4-byte aligned)
    __device__ __forceinline__ void init(uint8_t *dst) {
        p = dst;
        acc = 0;
        k = 0;
    }
    __device__ __forceinline__ void put(uint8_t b) {
        if (k == 0 && (reinterpret_cast<uintptr_t>(p) & 3)) {
            *p++ = b;  // unaligned head
            return;
        }
        acc |= (uint32_t)b << (8 * k);
        if (++k == 4) {
            *reinterpret_cast<uint32_t *>(p) = acc;
            p += 4;
            acc = 0;
            k = 0;
        }
    }
    __device__ __forceinline__ void finish() {
        for (uint32_t i = 0; i < k; i++) p[i] = (uint8_t)(acc >> (8 * i));
    }
};

// Decode codes that START in [pos, limit).  Returns the position after the last complete code
// (>= limit unless the bits ran out).  `truncated` is set if the bits end inside a code.
// With WRITE the UTF-8 bytes are stored at out.
template <bool WRITE>
__device__ __forceinline__ uint64_t decode_span(const DecParams &p, uint64_t pos, uint64_t limit, uint64_t &bytes,
                                                bool &truncated, uint8_t *out) {
    bytes = 0;
    truncated = false;
    BitReader br;
    br.seek(p, pos);
    ByteWriter bw;
    if (WRITE) bw.init(out);
    while (pos < limit) {
        br.refill();
        const uint32_t ent = __ldg(p.lut + br.peek(kLutBits));
        int32_t rune;
        if (ent >> 31) {
            const uint32_t len = (ent >> 21) & 0x3FFu;
            if (pos + len > p.max) {  // the code would need bits past the end (huffman.go:145)
                truncated = true;
                pos = p.max;
                break;
            }
            br.skip(len);
            pos += len;
            rune = (int32_t)(ent & 0x1FFFFFu);
        } else {  // longer than the table: finish on the tree
            int32_t node = (int32_t)ent;
            if (pos + kLutBits > p.max) {
                truncated = true;
                pos = p.max;
                break;
            }
            br.skip(kLutBits);
            pos += kLutBits;
            HuffNode nd = p.nodes[node];
            while (nd.left >= 0) {
                if (pos >= p.max) {
                    truncated = true;
                    break;
                }
                br.refill();
                node = br.peek(1) ? nd.right : nd.left;
                br.skip(1);
                nd = p.nodes[node];
                pos++;
            }
            if (truncated) break;
            rune = nd.right;
        }
        if (WRITE) {
            uint8_t u[4];
            const int w = utf8_encode(rune, u);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < w) bw.put(u[i]);
            bytes += (uint64_t)w;
        } else {
            bytes += (uint64_t)utf8_width(rune);
        }
    }
    if (WRITE) bw.finish();
    return pos;
}

__global__ void k_hdec_init(DecParams p, size_t subs, uint64_t *__restrict__ start, uint64_t *__restrict__ end,
                            uint64_t *__restrict__ cnt) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= subs) return;
    const uint64_t s0 = (uint64_t)t * kSubBits;
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, p.max);
    uint64_t bytes;
    bool trunc;
    const uint64_t e = decode_span<false>(p, s0, limit, bytes, trunc, nullptr);
    start[t] = s0;
    end[t] = e;
    cnt[t] = bytes;
}

__global__ void k_hdec_sync(DecParams p, size_t subs, uint64_t *__restrict__ start,
                            const uint64_t *__restrict__ end_prev, uint64_t *__restrict__ end_next,
                            uint64_t *__restrict__ cnt, uint32_t *__restrict__ changed) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= subs) return;
    if (t == 0) {
        end_next[0] = end_prev[0];
        return;
    }
    const uint64_t s1 = end_prev[t - 1];
    if (s1 == start[t]) {
        end_next[t] = end_prev[t];
        return;
    }
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, p.max);
    uint64_t bytes = 0;
    bool trunc;
    uint64_t e = s1;
    if (s1 < limit) e = decode_span<false>(p, s1, limit, bytes, trunc, nullptr);
    start[t] = s1;
    end_next[t] = e;
    cnt[t] = bytes;
    *changed = 1;
}

__global__ void k_hdec_write(DecParams p, size_t subs, const uint64_t *__restrict__ start,
                             const uint64_t *__restrict__ off, uint8_t *__restrict__ out,
                             uint32_t *__restrict__ err) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= subs) return;
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, p.max);
    const uint64_t s0 = start[t];
    if (s0 >= limit) return;
    uint64_t bytes;
    bool trunc;
    decode_span<true>(p, s0, limit, bytes, trunc, out + off[t]);
    if (trunc) *err = 1;  // data[i] with i == len(data) (huffman.go:145)
}

// ---- exact fallback for streams that do not self-synchronise (a complete fixed-length code whose
// length does not divide the subsequence size never re-aligns: the fix-up above would move one
// subsequence per round).  A code that starts before a subsequence boundary ends fewer than M =
// (longest code) bits after it, so the true start of subsequence t is t*256 + o for some o < M.
// Decode every subsequence from all M candidate starts (M x the decode work, only ever paid by such
// streams), which gives its transfer function "entry offset -> entry offset of the next one";
// compose the functions block by block, walk the blocks, and hand every subsequence its true start.
constexpr int kXferBlock = 256;  // subsequences per composition block

__global__ void k_hdec_xfer(DecParams p, size_t subs, uint32_t M, uint8_t *__restrict__ E) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= subs * M) return;
    const size_t t = idx / M;
    const uint32_t o = (uint32_t)(idx % M);
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, p.max);
    const uint64_t s0 = (uint64_t)t * kSubBits + o;
    uint64_t e = s0, bytes;
    bool trunc;
    if (s0 < limit) e = decode_span<false>(p, s0, limit, bytes, trunc, nullptr);
    E[idx] = (uint8_t)(e >= limit ? min(e - limit, (uint64_t)M - 1) : 0);
}
// F[b][o]: entry offset after the block's subsequences when the block is entered at offset o
__global__ void k_hdec_xfer_blocks(const uint8_t *__restrict__ E, size_t subs, uint32_t M, uint8_t *__restrict__ F) {
    const uint32_t o = threadIdx.x;
    if (o >= M) return;
    const size_t lo = (size_t)blockIdx.x * kXferBlock, hi = min(subs, lo + kXferBlock);
    uint32_t e = o;
    for (size_t t = lo; t < hi; t++) e = E[t * M + e];
    F[(size_t)blockIdx.x * M + o] = (uint8_t)e;
}
__global__ void k_hdec_xfer_spine(const uint8_t *__restrict__ F, size_t blocks, uint32_t M, uint8_t *__restrict__ ent) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t e = 0;  // bit 0 starts a code
    for (size_t b = 0; b < blocks; b++) {
        ent[b] = (uint8_t)e;
        e = F[b * M + e];
    }
}
__global__ void k_hdec_xfer_starts(const uint8_t *__restrict__ E, const uint8_t *__restrict__ ent, size_t subs,
                                   uint32_t M, uint64_t *__restrict__ start) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = b * kXferBlock;
    if (lo >= subs) return;
    const size_t hi = min(subs, lo + kXferBlock);
    uint32_t e = ent[b];
    for (size_t t = lo; t < hi; t++) {
        start[t] = (uint64_t)t * kSubBits + e;
        e = E[t * M + e];
    }
}
__global__ void k_hdec_recount(DecParams p, size_t subs, const uint64_t *__restrict__ start, uint64_t *__restrict__ cnt) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= subs) return;
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, p.max);
    uint64_t bytes = 0;
    bool trunc;
    if (start[t] < limit) decode_span<false>(p, start[t], limit, bytes, trunc, nullptr);
    cnt[t] = bytes;
}

constexpr int kSyncBatch = 4;       // fix-up rounds per look at the flags
constexpr int kSyncMaxRounds = 16;  // then the stream is taken not to synchronise

int huff_decompress_dev(const uint8_t *d_in, size_t n, const uint8_t *h_in, int strict, uint8_t **d_out, size_t *out_n,
                        cudaStream_t s) {
    ArenaScope scope(s);
    Trace tr("hd", s);
    // ---- host: find the first 5C 0A (strings.SplitN, huffman.go:261) and parse the header
    HostVec<uint8_t> h_copy;  // pinned (see HostVec)
    const uint8_t *h = h_in;
    size_t hn = n;
    auto find_sep = [](const uint8_t *b, size_t len) -> ptrdiff_t {
        for (size_t i = 0; i + 1 < len; i++)
            if (b[i] == 0x5C && b[i + 1] == 0x0A) return (ptrdiff_t)i;
        return -1;
    };
    ptrdiff_t sp = -1;
    if (h) {
        sp = find_sep(h, n);
    } else {
        // device-resident input: pull a growing prefix until it holds the separator
        size_t take = n < ((size_t)64 << 10) ? n : ((size_t)64 << 10);
        size_t have = 0;
        for (;;) {
            if (!h_copy.resize(take)) return RSN_ERR_NOMEM;
            RSN_CUDA(cudaMemcpyAsync(h_copy.data() + have, d_in + have, take - have, cudaMemcpyDeviceToHost, s));
            RSN_CUDA(stream_wait(s));
            // the separator may straddle the previous chunk boundary: rescan from one byte before it
            const size_t from = have ? have - 1 : 0;
            const ptrdiff_t r = find_sep(h_copy.data() + from, take - from);
            if (r >= 0) {
                sp = r + (ptrdiff_t)from;
                break;
            }
            if (take == n) break;
            have = take;
            take = take * 8 < n ? take * 8 : n;
        }
        h = h_copy.data();
        hn = take;
    }
    (void)hn;
    if (sp < 0) return RSN_ERR_NO_SEPARATOR;
    tr.mark("find sep");
    std::vector<HuffLeaf> leaves;
    if (!huff_parse_header(h, (size_t)sp, leaves)) return RSN_ERR_BAD_HEADER;
    if (leaves.empty()) return RSN_ERR_BAD_HEADER;  // buildTree on an empty map panics
    HuffTree tree;
    huff_build_tree(leaves, tree);

    tr.mark("header+tree");
    const size_t pay_off = (size_t)sp + 2;
    const size_t pn = n - pay_off;
    uint64_t diff = 0;
    if (pn) {
        if (h_in) {
            diff = h_in[pay_off];
        } else if (pay_off < h_copy.size()) {
            diff = h_copy[pay_off];
        } else {
            uint8_t b = 0;
            RSN_CUDA(cudaMemcpyAsync(&b, d_in + pay_off, 1, cudaMemcpyDeviceToHost, s));
            RSN_CUDA(stream_wait(s));
            diff = b;
        }
    }
    const uint64_t nbits = pn ? (uint64_t)(pn - 1) * 8 : 0;
    if (diff > nbits) return RSN_ERR_TRUNCATED;  // contentString[int(diff):] out of range
    const uint64_t max = nbits - diff;
    if (strict && max > 900000) return RSN_ERR_GUARD;

    DevBuf out;
    const HuffNode &rootn = tree.nodes[tree.root];
    if (rootn.left < 0) {  // single-leaf tree: exactly one symbol when no bits remain
        if (max > 0) return RSN_ERR_SINGLE_LEAF_LOOP;
        uint8_t u[4];
        const int w = utf8_encode(rootn.right, u);
        RSN_TRY(out.alloc_out(16, s));
        RSN_CUDA(cudaMemcpyAsync(out.p, u, (size_t)w, cudaMemcpyHostToDevice, s));
        RSN_CUDA(stream_wait(s));
        *d_out = (uint8_t *)out.release();
        *out_n = (size_t)w;
        return RSN_OK;
    }
    if (max == 0) return RSN_ERR_TRUNCATED;  // data[0] on an empty bit string

    DevBuf nodes;
    RSN_TRY(nodes.alloc(tree.nodes.size() * sizeof(HuffNode), s));
    HostVec<HuffNode> h_nodes(tree.nodes.size());  // pinned (see HostVec)
    if (!h_nodes.data()) return RSN_ERR_NOMEM;
    memcpy(h_nodes.data(), tree.nodes.data(), tree.nodes.size() * sizeof(HuffNode));
    RSN_CUDA(cudaMemcpyAsync(nodes.p, h_nodes.data(), tree.nodes.size() * sizeof(HuffNode), cudaMemcpyHostToDevice, s));

    // lookup table over the first kLutBits bits of a code
    HostVec<uint32_t> h_lut((size_t)1 << kLutBits);
    if (!h_lut.data()) return RSN_ERR_NOMEM;
    for (uint32_t idx = 0; idx < h_lut.size(); idx++) {
        int32_t node = tree.root;
        uint32_t len = 0;
        while (tree.nodes[node].left >= 0 && len < (uint32_t)kLutBits) {
            const uint32_t bit = (idx >> (kLutBits - 1 - len)) & 1u;
            node = bit ? tree.nodes[node].right : tree.nodes[node].left;
            len++;
        }
        if (tree.nodes[node].left < 0) h_lut[idx] = (1u << 31) | (len << 21) | ((uint32_t)tree.nodes[node].right & 0x1FFFFFu);
        else h_lut[idx] = (uint32_t)node;
    }
    DevBuf lut;
    RSN_TRY(lut.alloc(h_lut.size() * 4, s));
    RSN_CUDA(cudaMemcpyAsync(lut.p, h_lut.data(), h_lut.size() * 4, cudaMemcpyHostToDevice, s));

    DecParams p;
    {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(d_in + pay_off + 1);
        p.words = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        p.bit0 = (uint64_t)(addr & 3) * 8 + diff;
        p.nwords = ((addr & 3) + (pn - 1)) / 4;  // whole words inside the buffer; a ragged tail word is
        // still inside the allocation when the base is 4-byte aligned, which device buffers are
        if (((addr & 3) + (pn - 1)) % 4) p.nwords += 1;
    }
    p.max = max;
    p.nodes = nodes.as<HuffNode>();
    p.lut = lut.as<uint32_t>();
    p.root = tree.root;

    const size_t subs = (size_t)div_up(max, kSubBits);
    DevBuf start, endA, endB, cnt, off, flag;
    RSN_TRY(start.alloc(subs * 8, s));
    RSN_TRY(endA.alloc(subs * 8, s));
    RSN_TRY(endB.alloc(subs * 8, s));
    RSN_TRY(cnt.alloc(subs * 8, s));
    RSN_TRY(off.alloc((subs + 1) * 8, s));
    RSN_TRY(flag.alloc(32, s));
    const unsigned grid = (unsigned)div_up(subs, 128);
    RSN_LAUNCH(k_hdec_init, grid, 128, 0, s, p, subs, start.as<uint64_t>(), endA.as<uint64_t>(), cnt.as<uint64_t>());
    uint64_t *e_prev = endA.as<uint64_t>(), *e_next = endB.as<uint64_t>();
    bool converged = false;
    Ctx &cx = ctx();
    for (int rounds = 0; !converged && rounds < kSyncMaxRounds; rounds += kSyncBatch) {
        // a few fix-up rounds back to back, one look at their "changed" flags
        RSN_CUDA(cudaMemsetAsync(flag.p, 0, 32, s));
        for (int r = 0; r < kSyncBatch; r++) {
            RSN_LAUNCH(k_hdec_sync, grid, 128, 0, s, p, subs, start.as<uint64_t>(), e_prev, e_next, cnt.as<uint64_t>(),
                       flag.as<uint32_t>() + r);
            std::swap(e_prev, e_next);
        }
        RSN_CUDA(cudaMemcpyAsync(cx.h_scalars, flag.p, 16, cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
        converged = reinterpret_cast<const uint32_t *>(cx.h_scalars)[kSyncBatch - 1] == 0;
    }
    if (!converged) {  // exact fallback: transfer functions over the M candidate starts
        uint32_t M = 1;
        {
            std::vector<std::pair<int32_t, uint32_t>> st{{tree.root, 0u}};
            while (!st.empty()) {
                auto [node, depth] = st.back();
                st.pop_back();
                const HuffNode &nd = tree.nodes[node];
                if (nd.left < 0) {
                    M = std::max(M, depth);
                    continue;
                }
                st.push_back({nd.left, depth + 1});
                st.push_back({nd.right, depth + 1});
            }
        }
        if (M > 255) return RSN_ERR_UNSUPPORTED;
        const size_t xb = div_up(subs, (size_t)kXferBlock);
        DevBuf E, F, ent;
        RSN_TRY(E.alloc(subs * M + 16, s));
        RSN_TRY(F.alloc(xb * M + 16, s));
        RSN_TRY(ent.alloc(xb + 16, s));
        RSN_LAUNCH(k_hdec_xfer, (unsigned)div_up(subs * M, 128), 128, 0, s, p, subs, M, E.as<uint8_t>());
        RSN_LAUNCH(k_hdec_xfer_blocks, (unsigned)xb, 256, 0, s, E.as<uint8_t>(), subs, M, F.as<uint8_t>());
        RSN_LAUNCH(k_hdec_xfer_spine, 1, 32, 0, s, F.as<uint8_t>(), xb, M, ent.as<uint8_t>());
        RSN_LAUNCH(k_hdec_xfer_starts, (unsigned)div_up(xb, 128), 128, 0, s, E.as<uint8_t>(), ent.as<uint8_t>(), subs, M,
                   start.as<uint64_t>());
        RSN_LAUNCH(k_hdec_recount, grid, 128, 0, s, p, subs, start.as<uint64_t>(), cnt.as<uint64_t>());
        tr.mark("exact fallback");
    }
    tr.mark("init+sync");
    RSN_TRY(spine_scan_u64(cnt.as<uint64_t>(), off.as<uint64_t>(), off.as<uint64_t>() + subs, subs, s));
    uint64_t total = 0;
    RSN_TRY(read_u64(off.as<uint64_t>() + subs, &total, s));
    RSN_TRY(out.alloc_out(total + 16, s));
    RSN_CUDA(cudaMemsetAsync(flag.p, 0, 8, s));
    RSN_LAUNCH(k_hdec_write, grid, 128, 0, s, p, subs, start.as<uint64_t>(), off.as<uint64_t>(), out.as<uint8_t>(),
               flag.as<uint32_t>());
    uint64_t err = 0;
    RSN_TRY(read_u64(flag.as<uint64_t>(), &err, s));
    tr.mark("write");
    if ((uint32_t)err) return RSN_ERR_TRUNCATED;
    *d_out = (uint8_t *)out.release();
    *out_n = (size_t)total;
    return RSN_OK;
}

// ============================================================================= batches of small files
//
// Headers are parsed and trees built on the host for every file of the group (in parallel), the
// tables of all files go up in one copy, and each decode kernel runs once over the group
// (blockIdx.y = file).  The self-synchronisation loop iterates until no file changes.

struct HdecFile {
    DecParams p;
    uint64_t sub_base;   // first subsequence of the file in the group's start/end/cnt arrays
    uint64_t subs;
    uint64_t out_base;   // offset of the file's bytes in the group's result buffer
    uint64_t total;      // decoded bytes
    uint32_t err, pad;
};

__global__ void kb_hdec_init(const HdecFile *__restrict__ files, uint64_t *__restrict__ start,
                             uint64_t *__restrict__ end, uint64_t *__restrict__ cnt) {
    const HdecFile &f = files[blockIdx.y];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.subs) return;
    const uint64_t s0 = (uint64_t)t * kSubBits;
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, f.p.max);
    uint64_t bytes;
    bool trunc;
    const uint64_t e = decode_span<false>(f.p, s0, limit, bytes, trunc, nullptr);
    start[f.sub_base + t] = s0;
    end[f.sub_base + t] = e;
    cnt[f.sub_base + t] = bytes;
}

// changed[blockIdx.y] is set when a subsequence of that file moved in this round
__global__ void kb_hdec_sync(const HdecFile *__restrict__ files, uint64_t *__restrict__ start,
                             const uint64_t *__restrict__ end_prev, uint64_t *__restrict__ end_next,
                             uint64_t *__restrict__ cnt, uint32_t *__restrict__ changed) {
    const HdecFile &f = files[blockIdx.y];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.subs) return;
    const size_t g = f.sub_base + t;
    if (t == 0) {
        end_next[g] = end_prev[g];
        return;
    }
    const uint64_t s1 = end_prev[g - 1];
    if (s1 == start[g]) {
        end_next[g] = end_prev[g];
        return;
    }
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, f.p.max);
    uint64_t bytes = 0;
    bool trunc;
    uint64_t e = s1;
    if (s1 < limit) e = decode_span<false>(f.p, s1, limit, bytes, trunc, nullptr);
    start[g] = s1;
    end_next[g] = e;
    cnt[g] = bytes;
    changed[blockIdx.y] = 1;
}

__global__ void __launch_bounds__(256) kb_hdec_finish(HdecFile *__restrict__ files, const uint64_t *__restrict__ cnt,
                                                      uint64_t *__restrict__ off) {
    __shared__ uint64_t sm[33];
    HdecFile &f = files[blockIdx.x];
    const uint64_t total = cta_scan_u64(cnt + f.sub_base, off + f.sub_base, (size_t)f.subs, sm);
    if (threadIdx.x == 0) f.total = total;
}

__global__ void kb_hdec_write(HdecFile *__restrict__ files, const uint64_t *__restrict__ start,
                              const uint64_t *__restrict__ off, uint8_t *__restrict__ out) {
    HdecFile &f = files[blockIdx.y];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.subs) return;
    const uint64_t limit = min((uint64_t)(t + 1) * kSubBits, f.p.max);
    const uint64_t s0 = start[f.sub_base + t];
    if (s0 >= limit) return;
    uint64_t bytes;
    bool trunc;
    decode_span<true>(f.p, s0, limit, bytes, trunc, out + f.out_base + off[f.sub_base + t]);
    if (trunc) f.err = 1;  // data[i] with i == len(data) (huffman.go:145)
}

namespace {
struct HdecHost {           // what the host learns about one file before any kernel runs
    int rc = RSN_OK;
    bool per_file = false;  // leave it to huff_decompress_dev (single-leaf trees, huge frequencies)
    size_t pay_off = 0;
    uint64_t diff = 0, max = 0;
    std::vector<uint32_t> freq, rune;  // leaves in (freq asc, rune asc) order
};

void hdec_host_plan(const uint8_t *h, size_t n, HdecHost &pl) {
    ptrdiff_t sp = -1;
    for (size_t i = 0; i + 1 < n; i++)
        if (h[i] == 0x5C && h[i + 1] == 0x0A) {  // strings.SplitN(..., 2), huffman.go:261
            sp = (ptrdiff_t)i;
            break;
        }
    if (sp < 0) {
        pl.rc = RSN_ERR_NO_SEPARATOR;
        return;
    }
    std::vector<HuffLeaf> leaves;
    if (!huff_parse_header(h, (size_t)sp, leaves) || leaves.empty()) {
        pl.rc = RSN_ERR_BAD_HEADER;
        return;
    }
    pl.pay_off = (size_t)sp + 2;
    const size_t pn = n - pl.pay_off;
    pl.diff = pn ? h[pl.pay_off] : 0;
    const uint64_t nbits = pn ? (uint64_t)(pn - 1) * 8 : 0;
    if (pl.diff > nbits) {
        pl.rc = RSN_ERR_TRUNCATED;
        return;
    }
    pl.max = nbits - pl.diff;
    // the device builder keeps (frequency sum, node) in one 64-bit word, 32 bits each: anything a
    // header could say beyond that, and the single-leaf special cases, go through the single-stream call
    uint64_t sum = 0;
    bool fits = leaves.size() >= 2 && leaves.size() <= kTreeMaxLeaves;
    for (const HuffLeaf &l : leaves) {
        if (l.freq < 0 || l.freq >= ((int64_t)1 << 32)) fits = false;
        else sum += (uint64_t)l.freq;
    }
    if (!fits || sum >= ((uint64_t)1 << 32)) {
        pl.per_file = true;
        return;
    }
    if (pl.max == 0) {
        pl.rc = RSN_ERR_TRUNCATED;  // data[0] on an empty bit string
        return;
    }
    huff_sort_leaves(leaves);
    pl.freq.resize(leaves.size());
    pl.rune.resize(leaves.size());
    for (size_t i = 0; i < leaves.size(); i++) {
        pl.freq[i] = (uint32_t)leaves[i].freq;
        pl.rune[i] = (uint32_t)leaves[i].rune;
    }
}
}  // namespace

int huff_decompress_batch(const BatchIO &in, const uint8_t *const *h_in, BatchIO &out, cudaStream_t s,
                          bool h_prefix_only) {
    const size_t G = in.size();
    out.resize(G);
    out.rc = in.rc;
    if (G == 0) return RSN_OK;
    if (!h_in) return RSN_ERR_UNSUPPORTED;  // headers are parsed from the host copy of the streams
    ArenaScope scope(s);
    Trace tr("hd batch", s);
    std::vector<HdecHost> plan(G);
    parallel_for(G, batch_host_threads(), [&](size_t f) {
        if (in.rc[f] != RSN_OK) {
            plan[f].rc = in.rc[f];
            return;
        }
        hdec_host_plan(h_in[f], (size_t)in.n[f], plan[f]);
    });
    tr.mark("host headers");
    // leaves of all files in one upload; trees and 12-bit tables on the device (huff_tree.cu)
    HostVec<HdecFile> h(G);
    if (!h.data()) return RSN_ERR_NOMEM;
    static thread_local HostVec<uint8_t> tab;  // pinned staging: kept between groups
    std::vector<size_t> o_freq(G, 0), o_rune(G, 0), o_nodes(G, 0), o_parent(G, 0), o_lut(G, 0);
    size_t tab_n = 0, scr_n = 0;
    auto room = [](size_t &cursor, size_t bytes) {
        const size_t at = (cursor + 15) & ~(size_t)15;
        cursor = at + bytes;
        return at;
    };
    size_t subs_total = 0, subs_cap = 1;
    uint32_t kmax = 1;
    for (size_t f = 0; f < G; f++) {
        h[f] = HdecFile{};
        out.rc[f] = plan[f].rc;
        if (plan[f].rc != RSN_OK || plan[f].per_file) continue;
        const size_t k = plan[f].freq.size();
        kmax = std::max<uint32_t>(kmax, (uint32_t)k);
        o_freq[f] = room(tab_n, k * 4);
        o_rune[f] = room(tab_n, k * 4);
        o_nodes[f] = room(scr_n, 2 * k * sizeof(HuffNodeDev));
        o_parent[f] = room(scr_n, 2 * k * 4);
        o_lut[f] = room(scr_n, ((size_t)1 << kLutBits) * 4);
        h[f].subs = div_up(plan[f].max, kSubBits);
        h[f].sub_base = subs_total;
        subs_total += h[f].subs;
        subs_cap = std::max<size_t>(subs_cap, h[f].subs);
    }
    if (tab.size() < tab_n + 256 && !tab.resize(tab_n + 256)) return RSN_ERR_NOMEM;
    uint8_t *const tabp = tab.data();
    for (size_t f = 0; f < G; f++) {
        if (!h[f].subs) continue;
        memcpy(tabp + o_freq[f], plan[f].freq.data(), plan[f].freq.size() * 4);
        memcpy(tabp + o_rune[f], plan[f].rune.data(), plan[f].rune.size() * 4);
    }
    DevBuf dtab, dscr, djobs, files, start, endA, endB, cnt, off, flag;
    RSN_TRY(dtab.alloc(tab_n + 256, s));
    RSN_TRY(dscr.alloc(scr_n + 256, s));
    RSN_TRY(djobs.alloc(G * sizeof(TreeJob), s));
    RSN_TRY(files.alloc(G * sizeof(HdecFile), s));
    RSN_TRY(start.alloc(subs_total * 8 + 8, s));
    RSN_TRY(endA.alloc(subs_total * 8 + 8, s));
    RSN_TRY(endB.alloc(subs_total * 8 + 8, s));
    RSN_TRY(cnt.alloc(subs_total * 8 + 8, s));
    RSN_TRY(off.alloc(subs_total * 8 + 8, s));
    RSN_TRY(flag.alloc(G * 4 + 16, s));
    HostVec<TreeJob> jobs(G);
    if (!jobs.data()) return RSN_ERR_NOMEM;
    for (size_t f = 0; f < G; f++) {
        jobs[f] = TreeJob{};
        if (!h[f].subs) continue;
        TreeJob &j = jobs[f];
        j.freq = reinterpret_cast<const uint32_t *>(dtab.as<uint8_t>() + o_freq[f]);
        j.rune = reinterpret_cast<const uint32_t *>(dtab.as<uint8_t>() + o_rune[f]);
        j.k = (uint32_t)plan[f].freq.size();
        j.nodes = reinterpret_cast<HuffNodeDev *>(dscr.as<uint8_t>() + o_nodes[f]);
        j.parent = reinterpret_cast<uint32_t *>(dscr.as<uint8_t>() + o_parent[f]);
        j.lut = reinterpret_cast<uint32_t *>(dscr.as<uint8_t>() + o_lut[f]);
        DecParams &p = h[f].p;
        const size_t pn = (size_t)in.n[f] - plan[f].pay_off;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(in.ptr[f] + plan[f].pay_off + 1);
        p.words = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        p.bit0 = (uint64_t)(addr & 3) * 8 + plan[f].diff;
        p.nwords = ((addr & 3) + (pn - 1) + 3) / 4;
        p.max = plan[f].max;
        p.nodes = reinterpret_cast<const HuffNode *>(j.nodes);
        p.lut = j.lut;
        p.root = 0;  // (the decode kernels start from the table, not from the root)
    }
    if (tab_n) RSN_CUDA(cudaMemcpyAsync(dtab.p, tabp, tab_n, cudaMemcpyHostToDevice, s));
    RSN_CUDA(cudaMemcpyAsync(djobs.p, jobs.data(), G * sizeof(TreeJob), cudaMemcpyHostToDevice, s));
    if (subs_total) RSN_TRY(huff_tree_batch(djobs.as<TreeJob>(), G, kmax, s));
    RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(HdecFile), cudaMemcpyHostToDevice, s));
    const dim3 grid((unsigned)div_up(subs_cap, 128), (unsigned)G);
    tr.mark("tables");
    if (subs_total) {
        RSN_LAUNCH(kb_hdec_init, grid, 128, 0, s, files.as<HdecFile>(), start.as<uint64_t>(), endA.as<uint64_t>(),
                   cnt.as<uint64_t>());
        uint64_t *e_prev = endA.as<uint64_t>(), *e_next = endB.as<uint64_t>();
        // fix-up rounds in batches of kSyncBatch, one look per batch at the per-file flags of its last
        // round; files still moving after kSyncMaxRounds do not synchronise and take the
        // single-stream call, which has the exact fallback
        HostVec<uint32_t> h_changed(G);
        if (!h_changed.data()) return RSN_ERR_NOMEM;
        bool moving = true;
        for (int rounds = 0; moving && rounds < kSyncMaxRounds; rounds += kSyncBatch) {
            for (int r = 0; r < kSyncBatch; r++) {
                if (r == kSyncBatch - 1) RSN_CUDA(cudaMemsetAsync(flag.p, 0, G * 4, s));
                RSN_LAUNCH(kb_hdec_sync, grid, 128, 0, s, files.as<HdecFile>(), start.as<uint64_t>(), e_prev, e_next,
                           cnt.as<uint64_t>(), flag.as<uint32_t>());
                std::swap(e_prev, e_next);
            }
            RSN_CUDA(cudaMemcpyAsync(h_changed.data(), flag.p, G * 4, cudaMemcpyDeviceToHost, s));
            RSN_CUDA(stream_wait(s));
            moving = false;
            for (size_t f = 0; f < G; f++) moving |= h_changed[f] != 0;
        }
        if (moving)
            for (size_t f = 0; f < G; f++)
                if (h_changed[f] && h[f].subs) plan[f].per_file = true;
        RSN_LAUNCH(kb_hdec_finish, (unsigned)G, 256, 0, s, files.as<HdecFile>(), cnt.as<uint64_t>(), off.as<uint64_t>());
    }
    RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(HdecFile), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    tr.mark("init+sync");
    size_t total = 0;
    for (size_t f = 0; f < G; f++) {
        h[f].out_base = total;
        total += (h[f].total + 16 + 255) & ~(uint64_t)255;
    }
    DevBuf res;
    RSN_TRY(res.alloc_out(total + 256, s));
    if (subs_total) {
        RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(HdecFile), cudaMemcpyHostToDevice, s));
        RSN_LAUNCH(kb_hdec_write, grid, 128, 0, s, files.as<HdecFile>(), start.as<uint64_t>(), off.as<uint64_t>(),
                   res.as<uint8_t>());
        RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(HdecFile), cudaMemcpyDeviceToHost, s));
    }
    RSN_CUDA(stream_wait(s));
    tr.mark("write");
    for (size_t f = 0; f < G; f++) {
        if (out.rc[f] != RSN_OK || plan[f].per_file) continue;
        if (h[f].err) {
            out.rc[f] = RSN_ERR_TRUNCATED;
            continue;
        }
        out.ptr[f] = res.as<uint8_t>() + h[f].out_base;
        out.n[f] = h[f].total;
    }
    out.spans.push_back({res.as<uint8_t>(), res.bytes});
    out.owned.push_back(res.release());
    // the few files the kernels do not take (single-leaf trees) go through the per-file call
    for (size_t f = 0; f < G; f++) {
        if (out.rc[f] != RSN_OK || !plan[f].per_file) continue;
        uint8_t *r = nullptr;
        size_t rn = 0;
        out.rc[f] = huff_decompress_dev(in.ptr[f], (size_t)in.n[f], h_prefix_only ? nullptr : h_in[f], 0, &r, &rn, s);
        if (out.rc[f] != RSN_OK) continue;
        out.ptr[f] = r;
        out.n[f] = rn;
        out.owned.push_back(r);
    }
    return RSN_OK;
}

}  // namespace rsn
