// lzss_match.cu — K2: every position's longest match, in parallel.
//
// Semantics (compressorWorker, lzss.go:166-184, over the window of lzss.go:123-129):
//   win(i) = enc[max(0,i-W) : i]   (W bytes once i > W; exactly i bytes before that)
//   L(i)   = largest k with i+k <= n and enc[i:i+k] occurring wholly inside win(i)
//   off(i) = len(win) - (leftmost start of enc[i:i+L] in win)
// In distance form: L(i) = max over d in [1, min(i,W)] of min(lcp(i-d, i), d, n-i) and off(i)
// is the LARGEST d attaining it (leftmost source).
//
// The result feeds only the parse/emit of lzss.go:134-151, which needs L(i) exactly for every
// i but off(i) only where a reference can be emitted (token shorter than L, i.e. L >= 6).
// Contract of this kernel: packed[i] = (L << 16) | off with off exact for L >= 4 and
// unspecified for L < 4.
//
// Method: exact, candidate-pruned search.
//   k_chain_build  one warp scans a chunk left to right (with a W-byte warm-up halo) keeping, per
//                  k in {2,3,4}, a 4096-entry last-position table of hashed k-grams in shared
//                  memory; it emits prev_k[i] = distance to the nearest earlier position with the
//                  same k-gram hash (0 = none within W) and a bit "byte enc[i] occurs in win(i)".
//                  Intra-warp predecessors are resolved with __match_any_sync.
//   k_match_walk   one thread per position follows the prev_4 chain (all candidates with the
//                  same 4-gram hash, nearest first), verifies bytes with 4-byte unaligned loads,
//                  keeps the longest match and, among equals, the farthest.  Positions without a
//                  4-byte match fall to the prev_3 / prev_2 chains and the 1-byte bit.
//   k_match_hard   positions whose chain walk exceeds a work budget (degenerate inputs: long runs,
//                  short periods) are redone by a bounded far-to-near scan that stops as soon as
//                  no nearer candidate can win.
#include "lzss.cuh"

namespace rsn {

constexpr int kHashBits = 12;
constexpr int kHashSize = 1 << kHashBits;
constexpr uint32_t kWorkBudget = 1024;
constexpr uint32_t kHardSentinel = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t hash4(uint32_t w) { return (w * 2654435761u) >> (32 - kHashBits); }
__device__ __forceinline__ uint32_t hash3(uint32_t w) { return ((w & 0xFFFFFFu) * 2654435761u) >> (32 - kHashBits); }
__device__ __forceinline__ uint32_t hash2(uint32_t w) { return ((w & 0xFFFFu) * 2654435761u) >> (32 - kHashBits); }

// 4 bytes at byte offset pos (little-endian: byte pos is the low byte) from a 4-byte aligned
// buffer of nwords words; bytes past the buffer read as 0.
__device__ __forceinline__ uint32_t load32(const uint32_t *__restrict__ words, size_t nwords, size_t pos) {
    const size_t a = pos >> 2;
    const uint32_t lo = a < nwords ? __ldg(words + a) : 0u;
    const uint32_t sh = (uint32_t)(pos & 3) * 8;
    if (sh == 0) return lo;
    const uint32_t hi = a + 1 < nwords ? __ldg(words + a + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// ============================================================================= chain build

struct ChainOut {
    uint16_t *prev2, *prev3, *prev4;  // distance to the previous position with the same k-gram hash
    uint32_t *has1;                   // bit i%32 of has1[i/32]: enc[i] occurs in win(i)
};

// One warp per chunk.  Shared memory: three hash tables + a byte table + a 1 KiB staging buffer.
__global__ void __launch_bounds__(32) k_chain_build(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                    size_t chunk, ChainOut out) {
    __shared__ uint16_t tab2[kHashSize], tab3[kHashSize], tab4[kHashSize];
    __shared__ uint16_t tab1[256];
    __shared__ __align__(16) uint8_t stage[1024 + 16];
    const unsigned lane = threadIdx.x;
    const size_t c_lo = (size_t)blockIdx.x * chunk;
    const size_t c_hi = min(n, c_lo + chunk);
    size_t s0 = c_lo > W ? c_lo - W : 0;
    s0 &= ~(size_t)1023;  // macro-steps of 1024 positions, aligned
    for (int i = lane; i < kHashSize; i += 32) {
        tab2[i] = 0;
        tab3[i] = 0;
        tab4[i] = 0;
    }
    for (int i = lane; i < 256; i += 32) tab1[i] = 0;
    __syncwarp();
    const unsigned lt_mask = (1u << lane) - 1;
    for (size_t m0 = s0; m0 < c_hi; m0 += 1024) {
        // stage 1024 + 3 bytes
        for (int k = lane; k < (1024 + 16) / 16; k += 32) {
            const size_t g = m0 + (size_t)k * 16;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (g + 16 <= n && ((reinterpret_cast<uintptr_t>(enc + g) & 15) == 0)) {
                q = __ldg(reinterpret_cast<const uint4 *>(enc + g));
            } else if (g < n) {
                uint32_t w[4] = {0, 0, 0, 0};
                for (int b = 0; b < 16; b++)
                    if (g + b < n) w[b >> 2] |= (uint32_t)__ldg(enc + g + b) << ((b & 3) * 8);
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4 *>(stage + k * 16) = q;
        }
        __syncwarp();
        for (int t = 0; t < 32; t++) {
            const size_t p0 = m0 + (size_t)t * 32;
            if (p0 >= c_hi) break;
            const size_t i = p0 + lane;
            const uint32_t rel = (uint32_t)(i - s0) + 1;  // table entry for position i (0 = empty)
            const uint8_t *sp = stage + t * 32 + lane;
            const uint32_t w = (uint32_t)sp[0] | ((uint32_t)sp[1] << 8) | ((uint32_t)sp[2] << 16) | ((uint32_t)sp[3] << 24);
            const bool emit = i >= c_lo && i < c_hi;
            // ---- k = 1: exact byte table
            {
                const bool valid = i < n;
                const uint32_t key = valid ? (w & 0xFF) : (0x100u | lane);
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const unsigned lower = peers & lt_mask;
                uint32_t cand = 0;
                if (valid) cand = lower ? (uint32_t)(p0 - s0) + (31 - __clz(lower)) + 1 : tab1[key];
                const bool hit = valid && cand && (rel - cand) <= W;
                const unsigned bits = __ballot_sync(0xffffffffu, hit);
                if (lane == 0 && p0 >= c_lo) out.has1[p0 >> 5] = bits;
                __syncwarp();
                if (valid && (peers >> lane) == 1u) tab1[key] = (uint16_t)rel;  // highest lane of the group
            }
            // ---- k = 2, 3, 4: hashed tables
#pragma unroll
            for (int k = 2; k <= 4; k++) {
                const bool valid = i + k <= n;
                const uint32_t h = k == 2 ? hash2(w) : k == 3 ? hash3(w) : hash4(w);
                uint16_t *tab = k == 2 ? tab2 : k == 3 ? tab3 : tab4;
                uint16_t *prev = k == 2 ? out.prev2 : k == 3 ? out.prev3 : out.prev4;
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? h : (0x10000u | lane));
                const unsigned lower = peers & lt_mask;
                uint32_t cand = 0;
                if (valid) cand = lower ? (uint32_t)(p0 - s0) + (31 - __clz(lower)) + 1 : tab[h];
                uint32_t d = (valid && cand) ? rel - cand : 0;
                if (d > W) d = 0;
                if (emit) prev[i] = (uint16_t)d;
                __syncwarp();
                if (valid && (peers >> lane) == 1u) tab[h] = (uint16_t)rel;
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// ============================================================================= candidate walk

struct WalkIn {
    const uint32_t *words;  // enc as 4-byte words
    size_t nwords;
    const uint16_t *prev2, *prev3, *prev4;
    const uint32_t *has1;
};

__global__ void __launch_bounds__(256) k_match_walk(WalkIn in, size_t n, uint32_t W, uint32_t *__restrict__ packed,
                                                    uint32_t *__restrict__ hard_list,
                                                    unsigned long long *__restrict__ hard_count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t room = (uint32_t)min((size_t)W, n - i);  // L <= min(d, n-i) <= W
    uint32_t best = 0, boff = 0, work = 0;
    const uint32_t w0 = load32(in.words, in.nwords, i);
    if (room >= 4) {
        uint32_t d = 0;
        size_t j = i;
        for (;;) {
            const uint32_t step = __ldg(in.prev4 + j);
            if (step == 0) break;
            d += step;
            if (d > W) break;
            j -= step;
            work++;
            if (d < 4) continue;  // min(lcp, d) < 4: covered by the shorter chains
            if (load32(in.words, in.nwords, j) != w0) continue;  // hash collision
            const uint32_t cap = min(d, room);
            if (cap < best) continue;
            if (best > 4 && load32(in.words, in.nwords, j + best - 4) != load32(in.words, in.nwords, i + best - 4))
                continue;  // cannot reach the current best
            uint32_t l = 4;
            while (l < cap) {
                const uint32_t x = load32(in.words, in.nwords, j + l) ^ load32(in.words, in.nwords, i + l);
                work++;
                if (x) {
                    l += (__ffs(x) - 1) >> 3;
                    break;
                }
                l += 4;
            }
            l = min(l, cap);
            if (l >= best) {  // nearest first: an equal length farther away wins (leftmost source)
                best = l;
                boff = d;
            }
            if (work > kWorkBudget) break;
        }
    }
    if (work > kWorkBudget) {
        packed[i] = kHardSentinel;
        const unsigned long long slot = atomicAdd(hard_count, 1ull);
        hard_list[slot] = (uint32_t)(i & 0xFFFFFFFFu);
        hard_list[slot + n] = (uint32_t)(i >> 32);  // high half (streams beyond 4 GiB)
        return;
    }
    if (best < 4) {
        best = 0;
        boff = 0;
        if (room >= 3) {  // any 3-byte match at distance >= 3
            uint32_t d = 0;
            size_t j = i;
            for (;;) {
                const uint32_t step = __ldg(in.prev3 + j);
                if (step == 0) break;
                d += step;
                if (d > W) break;
                j -= step;
                if (d >= 3 && ((load32(in.words, in.nwords, j) ^ w0) & 0xFFFFFFu) == 0) {
                    best = 3;
                    boff = d;
                    break;
                }
            }
        }
        if (best == 0 && room >= 2) {
            uint32_t d = 0;
            size_t j = i;
            for (;;) {
                const uint32_t step = __ldg(in.prev2 + j);
                if (step == 0) break;
                d += step;
                if (d > W) break;
                j -= step;
                if (d >= 2 && ((load32(in.words, in.nwords, j) ^ w0) & 0xFFFFu) == 0) {
                    best = 2;
                    boff = d;
                    break;
                }
            }
        }
        if (best == 0 && ((__ldg(in.has1 + (i >> 5)) >> (i & 31)) & 1u)) {
            best = 1;
            boff = 1;
        }
    }
    packed[i] = (best << 16) | boff;
}

// ============================================================================= bounded far-to-near scan

// Exact for any input; cost is bounded by W candidate probes plus one long compare per
// improvement.  Distances go from far to near so ties keep the larger distance and the scan
// stops once d <= best (a candidate at distance d yields at most d).
__device__ __forceinline__ uint32_t match_far_to_near(const uint8_t *__restrict__ enc,
                                                      const uint32_t *__restrict__ words, size_t nwords, size_t n,
                                                      uint32_t W, size_t i) {
    const uint32_t dmax = (uint32_t)min((size_t)W, i);
    const uint32_t room = (uint32_t)min((size_t)W, n - i);
    const uint8_t c0 = __ldg(enc + i);
    uint32_t best = 0, boff = 0;
    for (uint32_t d = dmax; d >= 1 && d > best; d--) {
        const size_t j = i - d;
        if (__ldg(enc + j) != c0) continue;
        const uint32_t cap = min(d, room);
        if (best >= cap) continue;
        if (best && __ldg(enc + j + best) != __ldg(enc + i + best)) continue;  // must beat `best`
        uint32_t l = 0;
        while (l < cap) {
            const uint32_t x = load32(words, nwords, j + l) ^ load32(words, nwords, i + l);
            if (x) {
                l += (__ffs(x) - 1) >> 3;
                break;
            }
            l += 4;
        }
        l = min(l, cap);
        if (l > best) {
            best = l;
            boff = d;
        }
    }
    return (best << 16) | boff;
}

__global__ void __launch_bounds__(128) k_match_hard(const uint8_t *__restrict__ enc,
                                                    const uint32_t *__restrict__ words, size_t nwords, size_t n,
                                                    uint32_t W, uint32_t *__restrict__ packed,
                                                    const uint32_t *__restrict__ hard_list,
                                                    const unsigned long long *__restrict__ hard_count) {
    const unsigned long long count = *hard_count;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < count;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const size_t i = (size_t)hard_list[k] | ((size_t)hard_list[k + n] << 32);
        packed[i] = match_far_to_near(enc, words, nwords, n, W, i);
    }
}

// Kept for very small inputs and as the reference point of the first profile (profiles/r1_v0_*).
__global__ void __launch_bounds__(256) k_match_v0(const uint8_t *__restrict__ enc, const uint32_t *__restrict__ words,
                                                  size_t nwords, size_t n, uint32_t W,
                                                  uint32_t *__restrict__ packed) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    packed[i] = match_far_to_near(enc, words, nwords, n, W, i);
}

int lzss_match(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, cudaStream_t s) {
    if (n == 0) return RSN_OK;
    if (W < 1 || W > kMaxWindow) return RSN_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(d_enc) & 3) return RSN_ERR_INVALID_ARG;  // word loads
    if (W <= 4096) return lzss_match_tile(d_enc, n, W, d_packed, s);  // shared-memory tile kernel
    // wider windows: hash-chain path (global memory)
    const uint32_t *words = reinterpret_cast<const uint32_t *>(d_enc);
    const size_t nwords = (n + 3) / 4;
    if (n < 2048) {
        RSN_LAUNCH(k_match_v0, (unsigned)div_up(n, 256), 256, 0, s, d_enc, words, nwords, n, W, d_packed);
        return RSN_OK;
    }
    // chunk: enough warps to fill the machine, but at least 2W so the warm-up halo stays cheap
    size_t chunk = div_up(n, (size_t)148 * 9 * 2);
    chunk = max(chunk, (size_t)2 * W);
    chunk = max(chunk, (size_t)8192);
    chunk = div_up(chunk, 1024) * 1024;
    if (chunk + W + 2048 > 65000) chunk = ((65000 - W - 2048) / 1024) * 1024;  // u16 table entries
    const size_t chunks = div_up(n, chunk);
    DevBuf p2, p3, p4, h1, hard, hcount;
    RSN_TRY(p2.alloc(n * 2 + 16, s));
    RSN_TRY(p3.alloc(n * 2 + 16, s));
    RSN_TRY(p4.alloc(n * 2 + 16, s));
    RSN_TRY(h1.alloc((n / 32 + 2) * 4, s));
    RSN_TRY(hard.alloc(n * 8 + 16, s));
    RSN_TRY(hcount.alloc(16, s));
    RSN_CUDA(cudaMemsetAsync(hcount.p, 0, 16, s));
    ChainOut co{p2.as<uint16_t>(), p3.as<uint16_t>(), p4.as<uint16_t>(), h1.as<uint32_t>()};
    RSN_LAUNCH(k_chain_build, (unsigned)chunks, 32, 0, s, d_enc, n, W, chunk, co);
    WalkIn wi{words, nwords, p2.as<uint16_t>(), p3.as<uint16_t>(), p4.as<uint16_t>(), h1.as<uint32_t>()};
    RSN_LAUNCH(k_match_walk, (unsigned)div_up(n, 256), 256, 0, s, wi, n, W, d_packed, hard.as<uint32_t>(),
               hcount.as<unsigned long long>());
    RSN_LAUNCH(k_match_hard, 148 * 8, 128, 0, s, d_enc, words, nwords, n, W, d_packed, hard.as<uint32_t>(),
               hcount.as<unsigned long long>());
    return RSN_OK;
}

}  // namespace rsn
