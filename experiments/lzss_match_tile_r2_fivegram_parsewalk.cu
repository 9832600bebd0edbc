// lzss_match_tile.cu — K2 for windows up to 4096 (the engine's window, lzss.go:35): every
// position's longest match (every compressorWorker of lzss.go:156-184) from a tile staged in
// shared memory.
//
//   L(i)   = max over d in [1, min(i, W)] of min(lcp(i-d, i), d, n-i)
//   off(i) = the LARGEST d attaining it (leftmost source, bytes.Index)
//
// One CTA owns T = 8192 consecutive positions plus a halo of W earlier bytes and W bytes of
// look-ahead, all staged in shared memory.  The positions e of halo+tile ("entries") are sorted
// stably by their 4-gram, one byte per stage, byte 0 first.  A stable sort by byte k-1 of entries
// already grouped by their first k-1 bytes leaves them grouped by their k-gram, in position order
// inside every group (which group comes first does not matter), so after stage k the predecessor
// of an entry in the sorted order is the nearest earlier position with the same k-gram:
//   stages 1..3  ->  "L(q) >= k" for k = 1, 2, 3 from a look at the predecessors (distance in [k, W])
//   stage 4      ->  the complete candidate list of every position for L >= 4: a contiguous,
//                    position-ordered slice of the sorted array
// Digits.  A counting pass costs the same per entry whatever the digit width, so the width decides
// the cost of a stage.  Text-like tiles (the 31 most frequent byte values of the tile cover 7/8 of a
// sample of it) are sorted with ONE 32-way pass per byte over a per-tile code table: code 1..31 for
// the frequent values, code 0 for all the others.  Groups are then exact unless a gram holds a
// code-0 byte; such (rare) grams share a group with grams that differ in those bytes, so wherever a
// gram has a code-0 byte the bytes themselves are compared as well.  Other tiles (binary data) take
// two 16-way passes per byte on the byte itself.
//
// Candidates are evaluated far to near, one sorted slot per lane, a warp on 32 consecutive slots
// (slots of one group have lists that differ by one entry each, so the lanes of a warp carry
// similar work and read the same shared-memory words).  Bytes 4 and 5 of every entry sit beside the
// sorted positions, so a candidate is settled from one 16-bit load unless it matches 6 bytes; the
// scan stops as soon as the remaining distances cannot beat the best length (d <= best never
// wins), which also bounds degenerate inputs (runs, short periods).
#include "lzss.cuh"

#include <atomic>
#include <cstdlib>

namespace rsn {

namespace tile {

constexpr int T = 8192;             // positions per CTA
constexpr int WMAX = 4096;          // largest window handled here
constexpr int EMAX = T + WMAX + 16; // entries: halo (W, start rounded down to 16 bytes) + tile
constexpr int ECAP = EMAX + 48;     // rounded for warp chunks
constexpr int SLEN = EMAX + WMAX + 32;
constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int D5 = 32;              // digit values of a 5-bit pass
constexpr int SR = 16;              // parse walk: positions per sub-range
constexpr int NSUB = T / SR;
constexpr uint32_t kNone = 0xFFFFu;

struct Smem {
    uint32_t s_words[SLEN / 4 + 8];   // staged bytes: [base, base + avail), zero padded
    alignas(16) uint16_t a[ECAP];     // ping: sorted positions
    uint16_t b[ECAP];                 // pong; after the sort a and b together hold position | next two bytes << 16
    alignas(16) uint16_t ctr[D5 * THREADS];  // radix counters [digit][thread]; before: byte histogram; after: slot list
    uint32_t heads[ECAP / 32 + 2];    // bit r: slot r starts a 4-gram group
    uint8_t lowL[T];                  // 0..3 from stages 1..3
    uint8_t lut[256];                 // byte -> 5-bit code (0 = "other")
    uint32_t other[ECAP / 32 + 2];    // bit p: staged byte p (an entry, or one of the K bytes after the last) has code 0
    // parse walk (k_match_parse): one record per 16-position sub-range of the tile
    uint32_t land[NSUB];              // speculation rounds: landing (pred << 16 | position); chase: two u16 link tables
    uint16_t ent[NSUB];               // assumed entry (tile-relative), kNone = none
    uint16_t ext[NSUB];               // where the path from that entry leaves the sub-range
    uint32_t mark[NSUB / 32];         // chase: sub-ranges on the orbit
    uint32_t scan[33];
    uint32_t misc[8];
    // Diagonal cache for long matches: entry = (d << 32) | (start << 16) | end records that
    // s[x] == s[x - d] for every staged x in [start, end).  Entries are only ever written after the
    // bytes were compared, the data never changes, so any entry read (even a racing one) is true.
    unsigned long long diag[16 * 8];  // 16 sets (d & 15) x 8 ways
};

static_assert(sizeof(Smem) <= 115712, "two CTAs per SM: 2 x (sizeof(Smem) + 1 KiB) must fit 228 KiB");
static_assert(NSUB == THREADS, "parse walk: one lane per sub-range");

constexpr int kVoteSteps = 8;  // candidates a lane may walk between two votes

__device__ __forceinline__ uint32_t lds32(const uint8_t *s, uint32_t pos) {
    const uint32_t a = pos & ~3u;
    const uint32_t lo = *reinterpret_cast<const uint32_t *>(s + a);
    const uint32_t hi = *reinterpret_cast<const uint32_t *>(s + a + 4);
    return __funnelshift_r(lo, hi, (pos & 3u) * 8);
}

// ---- 16-way pass on a nibble of the byte itself (binary tiles).  Thread t owns the contiguous
// slots [t*per, (t+1)*per) and a private column of 16 digit counters ctr[digit][t], counted in two
// packed registers; one block-wide exclusive scan over the counters in (digit, thread) order turns
// them into stable destinations.
__device__ __forceinline__ void radix_pass4(Smem &sm, const uint8_t *s, const uint16_t *src, uint16_t *dst,
                                            uint32_t ev, uint32_t byteoff, uint32_t shift) {
    const uint32_t t = threadIdx.x;
    const uint32_t per = (ev + THREADS - 1) / THREADS;
    const uint32_t lo = min(ev, t * per), hi = min(ev, lo + per);
    uint16_t *col = sm.ctr + t;
    uint64_t acc0 = 0, acc1 = 0;
#pragma unroll 5
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t d = (s[src[i] + byteoff] >> shift) & 15u;
        const uint64_t inc = 1ull << ((d & 7u) * 8);
        acc0 += d < 8 ? inc : 0ull;
        acc1 += d < 8 ? 0ull : inc;
    }
#pragma unroll
    for (int d = 0; d < 8; d++) {
        col[d * THREADS] = (uint16_t)((acc0 >> (8 * d)) & 0xFFu);
        col[(d + 8) * THREADS] = (uint16_t)((acc1 >> (8 * d)) & 0xFFu);
    }
    __syncthreads();
    {
        uint4 *p = reinterpret_cast<uint4 *>(sm.ctr + t * 16);
        uint4 q0 = p[0], q1 = p[1];
        uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += (w[k] & 0xFFFFu) + (w[k] >> 16);
        uint32_t total;
        uint32_t run = block_exclusive_sum<uint32_t>(sum, sm.scan, total);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t a = w[k] & 0xFFFFu, b2 = w[k] >> 16;
            w[k] = run | ((run + a) << 16);
            run += a + b2;
        }
        p[0] = make_uint4(w[0], w[1], w[2], w[3]);
        p[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    __syncthreads();
    acc0 = 0;
    acc1 = 0;
#pragma unroll 5
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t e = src[i];
        const uint32_t d = (s[e + byteoff] >> shift) & 15u;
        const uint32_t sh = (d & 7u) * 8;
        const uint32_t rank = (uint32_t)(((d < 8 ? acc0 : acc1) >> sh) & 0xFFu);
        const uint64_t inc = 1ull << sh;
        acc0 += d < 8 ? inc : 0ull;
        acc1 += d < 8 ? 0ull : inc;
        dst[col[d * THREADS] + rank] = (uint16_t)e;
    }
    __syncthreads();
}

// ---- 32-way pass on the code of the byte (text-like tiles): same ownership, the private counters
// live in shared memory (a thread's column is touched by that thread only).
__device__ __forceinline__ void radix_pass5(Smem &sm, const uint8_t *s, const uint16_t *src, uint16_t *dst,
                                            uint32_t ev, uint32_t byteoff) {
    const uint32_t t = threadIdx.x;
    const uint32_t per = (ev + THREADS - 1) / THREADS;
    const uint32_t lo = min(ev, t * per), hi = min(ev, lo + per);
    uint16_t *col = sm.ctr + t;
#pragma unroll
    for (int d = 0; d < D5; d++) col[d * THREADS] = 0;
#pragma unroll 4
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t d = sm.lut[s[src[i] + byteoff]];
        col[d * THREADS] += 1;
    }
    __syncthreads();
    {
        // thread t scans the 32 consecutive counters [32t, 32t+32) of the linear (digit, thread) order
        uint4 *p = reinterpret_cast<uint4 *>(sm.ctr + t * D5);
        uint4 q[4] = {p[0], p[1], p[2], p[3]};
        uint32_t *w = reinterpret_cast<uint32_t *>(q);
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) sum += (w[k] & 0xFFFFu) + (w[k] >> 16);
        uint32_t total;
        uint32_t run = block_exclusive_sum<uint32_t>(sum, sm.scan, total);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t a = w[k] & 0xFFFFu, b2 = w[k] >> 16;
            w[k] = run | ((run + a) << 16);
            run += a + b2;
        }
        p[0] = q[0];
        p[1] = q[1];
        p[2] = q[2];
        p[3] = q[3];
    }
    __syncthreads();
#pragma unroll 4
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t e = src[i];
        const uint32_t d = sm.lut[s[e + byteoff]];
        const uint32_t pos = col[d * THREADS];
        col[d * THREADS] = (uint16_t)(pos + 1);
        dst[pos] = (uint16_t)e;
    }
    __syncthreads();
}

// codes of the first k bytes at p, 5 bits each
__device__ __forceinline__ uint32_t code_gram(const Smem &sm, const uint8_t *s, uint32_t p, int k) {
    uint32_t g = 0;
    for (int i = 0; i < k; i++) g |= (uint32_t)sm.lut[s[p + i]] << (5 * i);
    return g;
}
// does one of the first k bytes at p carry code 0 (its group may hold other grams)?  One bit per
// staged byte in sm.other.
__device__ __forceinline__ bool gram_has_other(const Smem &sm, uint32_t p, int k) {
    const uint32_t wi = p >> 5;
    const uint32_t bits = __funnelshift_r(sm.other[wi], sm.other[wi + 1], p & 31u);
    return (bits & ((1u << k) - 1u)) != 0;
}
// the first k (<= 5) bytes at p as one comparable value
template <int K>
__device__ __forceinline__ uint64_t gram_at(const uint8_t *s, uint32_t p) {
    const uint32_t lo = lds32(s, p);
    if (K <= 4) return lo & (0xFFFFFFFFu >> (32 - 8 * (K < 4 ? K : 4)));
    return (uint64_t)lo | ((uint64_t)s[p + 4] << 32);
}

// Continuation of a match that is already 32+ bytes long: compare on, but consult and feed the
// diagonal cache so that the thousands of positions of a tile that sit on the same long diagonal
// run (highly repetitive data) do not each re-compare it.  Returns the (uncapped) match length.
__device__ __noinline__ uint32_t long_lcp(Smem &sm, const uint8_t *s, uint32_t e, uint32_t d, uint32_t l, uint32_t cap,
                                          uint32_t avail) {
    const uint32_t j = e - d;
    unsigned long long *set = sm.diag + (d & 15u) * 8;
    uint32_t words = 0;
    while (l < cap) {
        if ((words++ & 15u) == 0) {  // is the rest of this diagonal already known?
#pragma unroll
            for (int way = 0; way < 8; way++) {
                const unsigned long long ent = set[way];
                const uint32_t st = (uint32_t)(ent >> 16) & 0xFFFFu, en = (uint32_t)ent & 0xFFFFu;
                if ((uint32_t)(ent >> 32) == d && st <= e + l && e + l < en) l = en - e;
            }
            if (l >= cap) break;
        }
        const uint32_t x = lds32(s, j + l) ^ lds32(s, e + l);
        if (x) {
            l += (__ffs(x) - 1) >> 3;
            break;
        }
        l += 4;
    }
    if (words >= 8 || l >= cap) {  // publish what was verified: equal on [e, e + l), clipped to staged bytes
        uint32_t st = e, en = min(e + l, avail);
        int victim = (int)((e >> 5) & 7u);
#pragma unroll
        for (int way = 0; way < 8; way++) {
            const unsigned long long cur = set[way];
            const uint32_t cst = (uint32_t)(cur >> 16) & 0xFFFFu, cen = (uint32_t)cur & 0xFFFFu;
            if ((uint32_t)(cur >> 32) == d && cst <= en && st <= cen) {  // overlapping: keep the union
                st = min(st, cst);
                en = max(en, cen);
                victim = way;
            }
        }
        set[victim] = ((unsigned long long)d << 32) | ((unsigned long long)st << 16) | en;
    }
    return l;
}

// Candidates of the sorted slot r (entry e): the slots [lo, r) of its group whose positions lie
// inside the window, in position order (farthest first).  arr32: position in the low half.
__device__ __forceinline__ uint32_t slot_lo(const Smem &sm, const uint32_t *arr32, uint32_t r, uint32_t e, uint32_t W) {
    uint32_t wi = r >> 5;
    uint32_t bits = sm.heads[wi] & (0xFFFFFFFFu >> (31 - (r & 31)));
    while (bits == 0) bits = sm.heads[--wi];
    uint32_t lo = (wi << 5) + (31 - __clz(bits)), hi = r;
    if (e > W && lo < hi && (arr32[lo] & 0xFFFFu) < e - W) {  // first candidate inside the window
        const uint32_t minpos = e - W;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((arr32[mid] & 0xFFFFu) < minpos) lo = mid + 1;
            else hi = mid;
        }
    }
    return lo;
}

// After stage K (entries grouped by K-gram, position order inside a group): L(q) >= K iff an earlier
// position with the same K-gram lies at a distance in [K, W] (the source must end at or before q).
// The predecessor in the sorted order is the nearest earlier member of the group, so one look at
// the neighbouring lane's entry settles almost every position; sources that overlap (d < K: runs)
// and merged groups (a code-0 byte in the gram) walk further back.
template <int K>
__device__ __forceinline__ void stage_flags(Smem &sm, const uint8_t *s, const uint16_t *arr, uint32_t ev, uint32_t halo,
                                            uint32_t W, size_t base, size_t n, bool use5) {
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = (threadIdx.x >> 5) * 32; c < ev; c += THREADS) {
        const uint32_t r = c + lane;
        uint32_t e = 0;
        uint64_t g = 0;
        if (r < ev) {
            e = arr[r];
            g = gram_at<K>(s, e);
        }
        uint32_t pe = __shfl_up_sync(0xffffffffu, e, 1);
        uint64_t pg = __shfl_up_sync(0xffffffffu, g, 1);
        if (lane == 0 && r > 0 && r < ev) {
            pe = arr[r - 1];
            pg = gram_at<K>(s, pe);
        }
        if (r >= ev || r == 0 || e < halo || base + e + K > n) continue;
        const uint32_t d1 = e - pe;  // pe > e (another group) wraps to a huge distance
        bool hit = false, slow = false;
        if (pg == g) {
            if (d1 >= (uint32_t)K) hit = d1 <= W;
            else slow = true;
        } else if (use5 && d1 <= W) {
            slow = gram_has_other(sm, e, K);
        }
        if (slow) {
            const uint32_t ge = use5 ? code_gram(sm, s, e, K) : 0u;
            for (uint32_t rr = r; rr-- > 0;) {
                const uint32_t p = arr[rr];
                const uint32_t d = e - p;
                if (d > W) break;
                if (gram_at<K>(s, p) == g) {
                    if (d >= (uint32_t)K) {
                        hit = true;
                        break;
                    }
                    continue;  // overlapping source: a farther one may do
                }
                if (!use5 || code_gram(sm, s, p, K) != ge) break;  // another group
            }
        }
        if (hit) sm.lowL[e - halo] = (uint8_t)K;
    }
}

// ---- candidate lists: K = gram length of the final groups (5 for text-like tiles, 3 for binary
// tiles).  arr32[r] = position | bytes K, K+1 of the entry << 16.
template <int K>
__device__ __forceinline__ void candidates(Smem &sm, const uint8_t *s, const uint32_t *arr32, const uint16_t *order,
                                           uint32_t n_order, uint32_t halo, uint32_t W, size_t base, size_t n,
                                           uint32_t avail, bool use5, uint32_t *__restrict__ packed) {
    const unsigned lane = threadIdx.x & 31;
    for (;;) {
        uint32_t k0 = 0;
        if (lane == 0) k0 = atomicAdd(&sm.misc[3], 32u);  // chunks of 32 consecutive slots, first come first served
        k0 = __shfl_sync(0xffffffffu, k0, 0);
        if (k0 >= n_order) break;
        const uint32_t k = k0 + lane;
        const bool act = k < n_order;
        uint32_t r = 0, e = 0, room = 0, c = 0, rend = 0, best = K - 1, boff = 0, mynb = 0;
        uint64_t g0 = 0;
        bool approx = false;
        if (act) {
            r = order[k];
            const uint32_t v = arr32[r];
            e = v & 0xFFFFu;
            mynb = v >> 16;
            room = (uint32_t)min((size_t)W, n - (base + e));
            if (room >= (uint32_t)K) {
                c = slot_lo(sm, arr32, r, e, W);
                rend = r;
                if (use5 && c < r && gram_has_other(sm, e, K)) {
                    approx = true;
                    g0 = gram_at<K>(s, e);
                }
            }
        }
        // Far to near.  A candidate at distance d yields at most min(d, room), so only slots with
        // position < jlim = e - best can win (the list is in position order).
        uint32_t jlim = e - best, tgt = 0;
        const uint8_t *sb = s;
        // The loop runs in warp-wide rounds: every lane walks up to kVoteSteps candidates of its list
        // (most are settled from their two inline bytes) or until one needs a real comparison, then
        // all lanes that hold such a survivor compare together.
        for (;;) {
            bool have = false;
            uint32_t j = 0;
#pragma unroll 1
            for (int step = 0; step < kVoteSteps && c < rend; step++) {
                const uint32_t v = arr32[c];
                j = v & 0xFFFFu;
                if (j >= jlim) {  // nearer candidates yield even less
                    c = rend;
                    break;
                }
                c++;
                const uint32_t x = (v >> 16) ^ mynb;
                if (approx && gram_at<K>(s, j) != g0) continue;  // merged group: another K-gram
                if (best >= (uint32_t)K + 2) {
                    if (x == 0 && sb[j] == tgt) {
                        have = true;
                        break;
                    }
                    continue;
                }
                const uint32_t cap = min(e - j, room);
                if (x == 0 && cap > (uint32_t)K + 2) {
                    have = true;
                    break;
                }
                uint32_t l = (x & 0xFFu) ? (uint32_t)K : ((x & 0xFF00u) ? (uint32_t)K + 1 : (uint32_t)K + 2);
                l = min(l, cap);
                if (l > best) {
                    best = l;
                    boff = e - j;
                    if (room <= best) {
                        c = rend;
                    } else {
                        jlim = e - best;
                        if (best >= (uint32_t)K + 2) {
                            sb = s + best;
                            tgt = s[e + best];
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, have)) {
                if (!__any_sync(0xffffffffu, c < rend)) break;
                continue;
            }
            if (have) {
                const uint32_t d = e - j;
                const uint32_t cap = min(d, room);
                uint32_t l = K + 2;
                while (l < cap && l < 38) {
                    const uint32_t x = lds32(s, j + l) ^ lds32(s, e + l);
                    if (x) {
                        l += (__ffs(x) - 1) >> 3;
                        goto lcp_done_h;
                    }
                    l += 4;
                }
                if (l < cap) l = long_lcp(sm, s, e, d, l, cap, avail);
            lcp_done_h:
                l = min(l, cap);
                if (l > best) {
                    best = l;
                    boff = d;
                    if (room <= best) {
                        c = rend;
                    } else {
                        jlim = e - best;
                        sb = s + best;
                        tgt = s[e + best];
                    }
                }
            }
        }
        if (act) {
            uint32_t L = sm.lowL[e - halo], off = 0;
            if (best >= (uint32_t)K) {
                L = best;
                off = boff;
            }
            packed[base + e] = (L << 16) | off;
        }
    }
}

// After the last stage (sorted positions in sm.a): pack the two bytes that follow each K-gram beside
// the positions (arr32 takes the place of a and b), mark the group heads, list the tile's slots that
// have an earlier group member in sorted order, and write the result of all the others.
// WALK: no slot list; instead inv[p] = slot of tile position p and jl[p] = 0 (nothing known yet).
template <int K, bool WALK>
__device__ __forceinline__ uint32_t finish_sort(Smem &sm, const uint8_t *s, uint32_t ev, uint32_t halo, size_t base,
                                                bool use5, uint16_t *order, uint32_t *__restrict__ packed, uint32_t Wwin = 0) {
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t *arr32 = reinterpret_cast<uint32_t *>(sm.a);
    constexpr uint32_t HALF = ECAP / 2;
    // upper half: the words land in b, nothing that is still needed is overwritten
    for (uint32_t r = HALF + threadIdx.x; r < ev; r += THREADS) {
        const uint32_t e = sm.a[r];
        arr32[r] = e | ((lds32(s, e + K) & 0xFFFFu) << 16);
    }
    // lower half: read everything first, then overwrite a
    constexpr int PER = (HALF + THREADS - 1) / THREADS;
    uint32_t keep[PER];
    const uint32_t lim = min(ev, HALF);
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t r = threadIdx.x + i * THREADS;
        keep[i] = r < lim ? sm.a[r] : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t r = threadIdx.x + i * THREADS;
        if (r < lim) arr32[r] = keep[i] | ((lds32(s, keep[i] + K) & 0xFFFFu) << 16);
    }
    __syncthreads();
    // heads; per-warp counts of the slots that go to the candidate phase
    const uint32_t per = ((ev + WARPS - 1) / WARPS + 31) & ~31u;
    const uint32_t lo = min(ev, w * per), hi = min(ev, lo + per);
    uint32_t mine = 0;
    for (uint32_t c = lo; c < hi; c += 32) {
        const uint32_t r = c + lane;
        uint32_t e = 0;
        uint64_t g = 0;
        if (r < hi) {
            e = arr32[r] & 0xFFFFu;
            g = gram_at<K>(s, e);
        }
        uint32_t pe = __shfl_up_sync(0xffffffffu, e, 1);
        uint64_t pg = __shfl_up_sync(0xffffffffu, g, 1);
        if (lane == 0 && r > 0 && r < hi) {
            pe = arr32[r - 1] & 0xFFFFu;
            pg = gram_at<K>(s, pe);
        }
        bool head = false;
        if (r < hi) {
            head = r == 0 || pg != g;
            if (head && r > 0 && use5 && gram_has_other(sm, e, K)) head = code_gram(sm, s, e, K) != code_gram(sm, s, pe, K);
        }
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        if (lane == 0) sm.heads[c >> 5] = hm;
        if (WALK) {
            if (r < hi && e >= halo) {
                order[e - halo] = (uint16_t)r;       // inv
                order[T + (e - halo)] = 0;           // jl
            }
            continue;  // (candidate counts follow once all heads are known)
        }
        mine += __popc(__ballot_sync(0xffffffffu, r < hi && !head && e >= halo));
    }
    if (WALK) {
        __syncthreads();
        // size of every tile slot's candidate list (capped at 31) beside its low-length flags, so the walk
        // does not have to find the list start again
        const uint32_t *a32 = reinterpret_cast<const uint32_t *>(sm.a);
        for (uint32_t r = threadIdx.x; r < ev; r += THREADS) {
            const uint32_t e = a32[r] & 0xFFFFu;
            if (e < halo || ((sm.heads[r >> 5] >> (r & 31)) & 1u)) continue;
            const uint32_t cnt = r - slot_lo(sm, a32, r, e, Wwin);
            sm.lowL[e - halo] |= (uint8_t)(min(cnt, 31u) << 3);
        }
        __syncthreads();
        return 0;
    }
    if (lane == 0) sm.scan[w] = mine;
    __syncthreads();
    uint32_t at = 0, total = 0;
    for (int k = 0; k < WARPS; k++) {
        const uint32_t v = sm.scan[k];
        at += k < (int)w ? v : 0u;
        total += v;
    }
    for (uint32_t c = lo; c < hi; c += 32) {
        const uint32_t r = c + lane;
        bool cand = false, tile_slot = false;
        uint32_t e = 0;
        if (r < hi) {
            e = arr32[r] & 0xFFFFu;
            tile_slot = e >= halo;
            cand = tile_slot && !((sm.heads[c >> 5] >> lane) & 1u);
        }
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        if (cand) order[at + __popc(m & ((1u << lane) - 1))] = (uint16_t)r;
        at += __popc(m);
        if (tile_slot && !cand) packed[base + e] = (uint32_t)sm.lowL[e - halo] << 16;  // first of its group
    }
    __syncthreads();
    return total;
}

}  // namespace tile

namespace tile {
struct Geo {
    size_t tile_start, base;
    uint32_t tile_len, halo, avail, ev;
};

// Stage the tile, pick its alphabet and sort the entries (5 stages of one 32-way pass for text-like
// tiles, 3 stages of two 16-way passes for binary ones), reading the low-length flags off every
// stage but the last.  Returns true for a text-like tile.  Sorted positions end up in sm.a.
__device__ __forceinline__ bool tile_sort(Smem &sm, const uint8_t *__restrict__ enc, size_t n, uint32_t W, size_t tile,
                                          Geo &geo) {
    uint8_t *s = reinterpret_cast<uint8_t *>(sm.s_words);
    const unsigned lane = threadIdx.x & 31;

    const size_t tile_start = tile * T;
    const uint32_t tile_len = (uint32_t)min((size_t)T, n - tile_start);
    const size_t base = tile_start > (size_t)W ? ((tile_start - W) & ~(size_t)15) : 0;  // 16-byte aligned
    const uint32_t halo = (uint32_t)(tile_start - base);
    const uint32_t avail = (uint32_t)min(n - base, (size_t)(halo + T + W));  // bytes staged
    const uint32_t ev = halo + tile_len;  // entries: every position up to the end of the tile
    geo = Geo{tile_start, base, tile_len, halo, avail, ev};

    // ---- stage bytes (zero padded), clear flags
    {
        const uint32_t nwords = (avail + 3) / 4;
        const uint8_t *g = enc + base;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            const uint4 *gv = reinterpret_cast<const uint4 *>(g);
            uint4 *sv = reinterpret_cast<uint4 *>(sm.s_words);
            const uint32_t fullv = avail / 16;
            for (uint32_t i = threadIdx.x; i < fullv; i += THREADS) sv[i] = __ldg(gv + i);
            for (uint32_t i = fullv * 4 + threadIdx.x; i < nwords; i += THREADS) {
                uint32_t v = 0;
                for (uint32_t b = 0; b < 4; b++)
                    if (i * 4 + b < avail) v |= (uint32_t)__ldg(g + i * 4 + b) << (b * 8);
                sm.s_words[i] = v;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < nwords; i += THREADS) {
                uint32_t v = 0;
                for (uint32_t b = 0; b < 4; b++)
                    if (i * 4 + b < avail) v |= (uint32_t)__ldg(g + i * 4 + b) << (b * 8);
                sm.s_words[i] = v;
            }
        }
        for (uint32_t i = nwords + threadIdx.x; i < nwords + 8 && i < SLEN / 4 + 8; i += THREADS) sm.s_words[i] = 0;
        for (uint32_t i = threadIdx.x; i < T; i += THREADS) sm.lowL[i] = 0;
        if (threadIdx.x < 128) sm.diag[threadIdx.x] = 0;  // d = 0 never matches a real distance
        if (threadIdx.x < 8) sm.misc[threadIdx.x] = 0;
    }
    // ---- alphabet of the tile: a sampled byte histogram, the 31 most frequent values get codes 1..31
    uint32_t *hist = reinterpret_cast<uint32_t *>(sm.ctr);  // [256] counts, [256..512) present list
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x * 4 + 1; i < ev; i += THREADS * 4) atomicAdd(&hist[s[i]], 1u);
    __syncthreads();
    if (threadIdx.x < 256) {
        const uint32_t c = hist[threadIdx.x];
        const unsigned m = __ballot_sync(0xffffffffu, c != 0);
        uint32_t wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&sm.misc[0], (uint32_t)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (c) hist[256 + wbase + __popc(m & ((1u << lane) - 1))] = (c << 8) | threadIdx.x;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        const uint32_t c = hist[threadIdx.x];
        uint32_t code = 0;
        if (c) {
            const uint32_t np = sm.misc[0], key = (c << 8) | threadIdx.x;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < np; j++) rank += hist[256 + j] > key;
            if (rank < 31) code = rank + 1;
            else atomicAdd(&sm.misc[1], c);  // sampled bytes that fall to "other"
            atomicAdd(&sm.misc[2], c);
        }
        sm.lut[threadIdx.x] = (uint8_t)code;
    }
    __syncthreads();
    const bool use5 = sm.misc[1] * 8 <= sm.misc[2];  // "other" is at most 1/8 of the sample
    // one bit per staged byte: its code is 0 (only meaningful for text-like tiles)
    for (uint32_t c = (threadIdx.x >> 5) * 32; c < ((ev + 71u) & ~31u); c += THREADS) {
        const uint32_t p = c + lane;
        const unsigned m = __ballot_sync(0xffffffffu, use5 && p < avail && sm.lut[s[p]] == 0);
        if (lane == 0) sm.other[c >> 5] = m;
    }
    // identity order: text-like tiles take an odd number of passes, so they start in b to end in a
    {
        uint16_t *id = use5 ? sm.b : sm.a;
        for (uint32_t e = threadIdx.x; e < ev; e += THREADS) id[e] = (uint16_t)e;
    }
    __syncthreads();  // hist (in ctr) is dead from here

    if (use5) {
        // ---- five stages of one 32-way pass: sort by byte k-1, then read the k-gram flags off the neighbours
        radix_pass5(sm, s, sm.b, sm.a, ev, 0);
        stage_flags<1>(sm, s, sm.a, ev, halo, W, base, n, true);
        radix_pass5(sm, s, sm.a, sm.b, ev, 1);
        stage_flags<2>(sm, s, sm.b, ev, halo, W, base, n, true);
        radix_pass5(sm, s, sm.b, sm.a, ev, 2);
        stage_flags<3>(sm, s, sm.a, ev, halo, W, base, n, true);
        radix_pass5(sm, s, sm.a, sm.b, ev, 3);
        stage_flags<4>(sm, s, sm.b, ev, halo, W, base, n, true);
        radix_pass5(sm, s, sm.b, sm.a, ev, 4);
        // (stage_flags only reads the sorted array and writes lowL: the next pass's first barrier orders it)
    } else {
        // ---- binary tiles: three stages of two 16-way passes, candidate lists by 3-gram
        radix_pass4(sm, s, sm.a, sm.b, ev, 0, 0);
        radix_pass4(sm, s, sm.b, sm.a, ev, 0, 4);
        stage_flags<1>(sm, s, sm.a, ev, halo, W, base, n, false);
        radix_pass4(sm, s, sm.a, sm.b, ev, 1, 0);
        radix_pass4(sm, s, sm.b, sm.a, ev, 1, 4);
        stage_flags<2>(sm, s, sm.a, ev, halo, W, base, n, false);
        radix_pass4(sm, s, sm.a, sm.b, ev, 2, 0);
        radix_pass4(sm, s, sm.b, sm.a, ev, 2, 4);
    }
    return use5;
}
}  // namespace tile

// ============================================================================= parse walk
//
// The merge loop of lzss.go:134-151 only ever looks at the positions of one orbit: 0, then
// i + max(L(i), 1), ...  On text that is a quarter of the positions, on log-like data a fifteenth, and
// most of those have L < 5, which the sort stages have already settled.  k_match_parse therefore
// evaluates candidate lists only where the parse can actually land:
//   1. every 32-position sub-range of the tile walks the orbit from its own first position (256
//      lanes side by side); orbits from different starts merge within a few tokens, so
//   2. a few speculation rounds re-walk each sub-range from where its predecessor's orbit lands in it,
//      until it joins positions that are already known, and record the path and its exit;
//   3. one warp then chases the orbit through the tile from a given entry, hopping from sub-range to
//      sub-range over those recorded paths and evaluating only where speculation missed.
// The tile first chases from its own first position and publishes that exit; it then takes the
// exit published by the tile before it as its true entry and chases again until it meets its first
// chase.  If a tile's two chases leave at different positions the tile after it started from a
// wrong entry: the kernel reports it and the caller falls back to the all-positions pipeline
// (degenerate data only: runs and short periods, whose orbits never merge).
namespace tile {

struct ParseOut {
    uint32_t *packed;      // [n] (L << 16) | off, written at evaluated positions only
    uint16_t *visited;     // orbit bitmap, one u16 per 16 positions (whole 4096-position blocks)
    uint64_t *blk_bytes;   // output bytes of each 4096-position block
    uint32_t *spec_exit;   // [tiles] exit of the tile's own-start orbit, relative to the tile's end; kNotReady before
    uint32_t *ctl;         // [0] tile tickets, [1] != 0: some tile's chases disagree
};
constexpr uint32_t kNotReady = 0xFFFFFFFFu;

// jl[p]: bit 15 known, bits 13-14 digits of off - 1, bits 0-12 L
__device__ __forceinline__ uint32_t jl_jump(uint32_t v) {
    const uint32_t L = v & 0x1FFFu;
    return L ? L : 1u;
}
__device__ __forceinline__ uint32_t jl_bytes(uint32_t v) {  // lzss.go:141-150: literal, token iff strictly shorter, else raw
    const uint32_t L = v & 0x1FFFu;
    if (L == 0) return 1u;
    const uint32_t tl = 3u + (((v >> 13) & 3u) + 1u) + (uint32_t)ndig_u32(L);
    return tl < L ? tl : L;
}

struct WalkCtx {
    Smem &sm;
    const uint8_t *s;
    const uint32_t *arr32;
    const uint16_t *inv;
    uint16_t *jl;
    uint32_t halo, W, avail;
    size_t base, n;
    bool use5;
    uint32_t *packed;
};

// Every lane walks p -> p + jump(p) -> ... while p < pend, evaluating the positions it meets that are
// not known yet (lanes with p >= pend idle along).  STOP_KNOWN: a lane stops at the first known
// position instead of following it.  Returns where the lane stopped.  Warp-synchronous.
template <int K, bool STOP_KNOWN>
__device__ __forceinline__ uint32_t walk(const WalkCtx &w, uint32_t p, uint32_t pend) {
    Smem &sm = w.sm;
    const uint8_t *s = w.s;
    bool busy = false;
    uint32_t e = 0, room = 0, c = 0, rend = 0, best = 0, boff = 0, mynb = 0, jlim = 0, tgt = 0;
    uint64_t g0 = 0;
    bool approx = false;
    const uint8_t *sb = s;
#define RSN_WALK_FINALIZE(L_, off_)                                                               \
    do {                                                                                          \
        const uint32_t fl_ = (L_), fo_ = (off_);                                                  \
        w.jl[p] = (uint16_t)(0x8000u | ((uint32_t)(ndig_u32(fo_) - 1) << 13) | fl_);              \
        w.packed[w.base + w.halo + p] = (fl_ << 16) | fo_;                                        \
        p += fl_ ? fl_ : 1u;                                                                      \
    } while (0)
    for (;;) {
        if (!busy) {
            while (p < pend) {
                const uint32_t v16 = w.jl[p];
                if (v16 & 0x8000u) {
                    if (STOP_KNOWN) {
                        pend = 0;
                        break;
                    }
                    p += jl_jump(v16);
                    continue;
                }
                const uint32_t fl = sm.lowL[p];  // low-length flags | candidate count << 3
                e = w.halo + p;
                room = (uint32_t)min((size_t)w.W, w.n - (w.base + e));
                if ((fl >> 3) == 0 || room < (uint32_t)K) {
                    RSN_WALK_FINALIZE(fl & 7u, 0u);
                    continue;
                }
                const uint32_t r = w.inv[p];
                const uint32_t v = w.arr32[r];
                c = (fl >> 3) < 31u ? r - (fl >> 3) : slot_lo(sm, w.arr32, r, e, w.W);
                rend = r;
                mynb = v >> 16;
                best = K - 1;
                boff = 0;
                jlim = e - best;
                approx = w.use5 && gram_has_other(sm, e, K);
                if (approx) g0 = gram_at<K>(s, e);
                busy = true;
                break;
            }
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        bool have = false;
        uint32_t j = 0;
        if (busy) {
#pragma unroll 1
            for (int step = 0; step < kVoteSteps && c < rend; step++) {
                const uint32_t v = w.arr32[c];
                j = v & 0xFFFFu;
                if (j >= jlim) {  // nearer candidates yield even less
                    c = rend;
                    break;
                }
                c++;
                const uint32_t x = (v >> 16) ^ mynb;
                if (approx && gram_at<K>(s, j) != g0) continue;  // merged group: another K-gram
                if (best >= (uint32_t)K + 2) {
                    if (x == 0 && sb[j] == tgt) {
                        have = true;
                        break;
                    }
                    continue;
                }
                const uint32_t cap = min(e - j, room);
                if (x == 0 && cap > (uint32_t)K + 2) {
                    have = true;
                    break;
                }
                uint32_t l = (x & 0xFFu) ? (uint32_t)K : ((x & 0xFF00u) ? (uint32_t)K + 1 : (uint32_t)K + 2);
                l = min(l, cap);
                if (l > best) {
                    best = l;
                    boff = e - j;
                    if (room <= best) {
                        c = rend;
                    } else {
                        jlim = e - best;
                        if (best >= (uint32_t)K + 2) {
                            sb = s + best;
                            tgt = s[e + best];
                        }
                    }
                }
            }
        }
        if (__any_sync(0xffffffffu, have)) {
            if (have) {
                const uint32_t d = e - j;
                const uint32_t cap = min(d, room);
                uint32_t l = K + 2;
                while (l < cap && l < 38) {
                    const uint32_t x = lds32(s, j + l) ^ lds32(s, e + l);
                    if (x) {
                        l += (__ffs(x) - 1) >> 3;
                        goto lcp_done_w;
                    }
                    l += 4;
                }
                if (l < cap) l = long_lcp(sm, s, e, d, l, cap, w.avail);
            lcp_done_w:
                l = min(l, cap);
                if (l > best) {
                    best = l;
                    boff = d;
                    if (room <= best) {
                        c = rend;
                    } else {
                        jlim = e - best;
                        sb = s + best;
                        tgt = s[e + best];
                    }
                }
            }
        }
        if (busy && c >= rend) {
            if (best >= (uint32_t)K) RSN_WALK_FINALIZE(best, boff);
            else RSN_WALK_FINALIZE(sm.lowL[p] & 7u, 0u);
            busy = false;
        }
    }
    return p;
}

// path of known positions from a to the end of its sub-range: bits (relative to the sub-range) and exit
__device__ __forceinline__ void known_path(const uint16_t *jl, uint32_t a, uint32_t end, uint32_t &bits, uint32_t &exit_) {
    uint32_t q = a, b = 0;
    while (q < end) {
        b |= 1u << (q & (SR - 1));
        q += jl_jump(jl[q]);
    }
    bits = b;
    exit_ = q;
}

template <int K>
__device__ __forceinline__ void parse_walk(Smem &sm, const Geo &g, uint32_t W, size_t n, bool use5, size_t tile,
                                           const ParseOut &o) {
    const uint8_t *s = reinterpret_cast<const uint8_t *>(sm.s_words);
    uint16_t *inv = sm.ctr, *jl = sm.ctr + T;
    const WalkCtx w{sm, s, reinterpret_cast<const uint32_t *>(sm.a), inv, jl, g.halo, W, g.avail, g.base, n, use5, o.packed};
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile_len = g.tile_len;
    const uint32_t nsub = (tile_len + SR - 1) / SR;
    const uint32_t j = threadIdx.x;  // one lane per sub-range (NSUB == THREADS)
    const uint32_t sub_lo = j * SR, sub_end = min(sub_lo + SR, tile_len);
    const bool mine = j < nsub;
    sm.land[j] = 0;
    sm.ent[j] = (uint16_t)kNone;
    sm.ext[j] = 0;
    if (threadIdx.x < 8) sm.misc[threadIdx.x] = 0;
    __syncthreads();
    // ---- 1. every sub-range from its own first position
    {
        const uint32_t x = walk<K, false>(w, mine ? sub_lo : 0u, mine ? sub_end : 0u);
        if (mine && x < tile_len) atomicMax(&sm.land[x / SR], (j << 16) | x);
    }
    __syncthreads();
    // ---- 2. speculation rounds: from where the predecessor's orbit lands, until known positions
    for (int round = 0; round < 4; round++) {
        uint32_t a = kNone;
        if (mine) {
            const uint32_t l = sm.land[j];
            a = j == 0 ? 0u : (l ? (l & 0xFFFFu) : kNone);
        }
        const bool changed = mine && a != sm.ent[j];
        __syncthreads();
        sm.land[j] = 0;
        const bool go = changed && a != kNone;
        walk<K, true>(w, go ? a : 0u, go ? sub_end : 0u);
        if (changed) {
            uint32_t bits = 0, x = 0;
            if (a != kNone) known_path(jl, a, sub_end, bits, x);
            sm.ent[j] = (uint16_t)a;
            sm.ext[j] = (uint16_t)x;
        }
        __syncthreads();
        if (mine && sm.ent[j] != kNone) {
            const uint32_t x = sm.ext[j];
            if (x < tile_len) atomicMax(&sm.land[x / SR], (j << 16) | x);
        }
        __syncthreads();
    }
    // ---- 3. chase: the sub-ranges on the orbit from entry E.  A sub-range links to the one its path
    // lands in if that one's assumed entry is exactly the landing; marks spread along the links by
    // pointer doubling.  Where a link is missing the landing sub-range is walked from the true entry
    // (warp 0) and its record corrected, then the marks are spread again.
    uint16_t *nxa = reinterpret_cast<uint16_t *>(sm.land), *nxb = nxa + NSUB;
    uint32_t spec = 0, final_exit = 0;
    for (int run = 0; run < 2; run++) {
        uint32_t E = 0;
        if (run == 1) {
            // the true entry: where the orbit of the tile before this one leaves it
            if (tile > 0) {
                if (threadIdx.x == 0) {
                    volatile uint32_t *src = o.spec_exit + (tile - 1);
                    uint32_t spins = 0, v;
                    while ((v = *src) == kNotReady) {
                        __nanosleep(200);
                        if (++spins > (1u << 23)) {  // seconds: never in a healthy run; fail the call, do not hang
                            atomicOr(o.ctl + 1, 2u);
                            v = 0;
                            break;
                        }
                    }
                    sm.misc[0] = v;
                }
                __syncthreads();
                E = sm.misc[0];
            }
            if (E == 0) {  // same entry as the first chase: its marks stand
                final_exit = spec;
                break;
            }
        }
        uint32_t exit_pos = E;  // tile-relative position where the orbit leaves the tile
        for (int attempt = 0; attempt < 96 && exit_pos < tile_len; attempt++) {
            // the orbit stands at exit_pos inside the tile: its sub-range must start a chain there
            const uint32_t js = exit_pos / SR;
            if (sm.ent[js] != exit_pos) {  // repair: walk that sub-range from the true entry
                __syncthreads();
                if (warp == 0) {
                    const uint32_t end = min((js + 1) * SR, tile_len);
                    const uint32_t x = walk<K, false>(w, lane == 0 ? exit_pos : 0u, lane == 0 ? end : 0u);
                    if (lane == 0) {
                        sm.ent[js] = (uint16_t)exit_pos;
                        sm.ext[js] = (uint16_t)x;
                    }
                }
                __syncthreads();
            }
            // links
            uint32_t nx = NSUB;
            if (mine && sm.ent[j] != kNone) {
                const uint32_t x = sm.ext[j];
                if (x < tile_len && sm.ent[x / SR] == x) nx = x / SR;
            }
            nxa[j] = (uint16_t)nx;
            if (attempt == 0 && j < NSUB / 32) sm.mark[j] = 0;
            __syncthreads();
            if (j == js) atomicOr(&sm.mark[js >> 5], 1u << (js & 31));
            __syncthreads();
            uint16_t *cur = nxa, *nxt2 = nxb;
            for (int r = 0; r < 9; r++) {  // 2^9 sub-ranges
                const uint32_t t = cur[j];
                if (t < NSUB) {
                    if ((sm.mark[j >> 5] >> (j & 31)) & 1u) atomicOr(&sm.mark[t >> 5], 1u << (t & 31));
                    nxt2[j] = cur[t];
                } else {
                    nxt2[j] = (uint16_t)NSUB;
                }
                __syncthreads();
                uint16_t *tmp = cur;
                cur = nxt2;
                nxt2 = tmp;
            }
            // the marked sub-range without a link is where the chain ends
            if (mine && ((sm.mark[j >> 5] >> (j & 31)) & 1u) && sm.ent[j] != kNone) {
                const uint32_t x = sm.ext[j];
                if (!(x < tile_len && sm.ent[x / SR] == x)) sm.misc[1] = x;
            }
            __syncthreads();
            exit_pos = sm.misc[1];
            __syncthreads();
        }
        if (exit_pos < tile_len && threadIdx.x == 0) atomicOr(o.ctl + 1, 4u);  // too many repairs: give up (fallback)
        const uint32_t exit_rel = exit_pos >= tile_len ? exit_pos - tile_len : 0u;
        if (run == 0) {
            spec = exit_rel;
            if (threadIdx.x == 0) {
                *(volatile uint32_t *)(o.spec_exit + tile) = spec;
                __threadfence();
            }
        }
        final_exit = exit_rel;
    }
    if (threadIdx.x == 0 && final_exit != spec && g.tile_start + tile_len < n) atomicOr(o.ctl + 1, 1u);
    __syncthreads();
    // ---- results: orbit bitmap and output bytes of the tile's (up to) two 4096-position blocks
    const uint32_t nblk = (tile_len + 4095) / 4096;
    {
        uint32_t bits = 0, x = 0, bytes = 0;
        if (mine && ((sm.mark[j >> 5] >> (j & 31)) & 1u) && sm.ent[j] != kNone) {
            known_path(jl, sm.ent[j], sub_end, bits, x);
            for (uint32_t b = bits; b; b &= b - 1) bytes += jl_bytes(jl[sub_lo + (__ffs(b) - 1)]);
        }
        if (threadIdx.x < nblk * 256) o.visited[(g.tile_start >> 4) + threadIdx.x] = (uint16_t)bits;
        for (int d = 16; d; d >>= 1) bytes += __shfl_down_sync(0xffffffffu, bytes, d);
        if (lane == 0 && bytes) atomicAdd(&sm.misc[5 + (j >> 8)], bytes);
    }
    __syncthreads();
    if (threadIdx.x < nblk) o.blk_bytes[(g.tile_start >> 12) + threadIdx.x] = sm.misc[5 + threadIdx.x];
}

}  // namespace tile

__device__ __forceinline__ void match_parse_body(const uint8_t *__restrict__ enc, size_t n, uint32_t W, size_t tile,
                                                 const tile::ParseOut &o) {
    using namespace tile;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const uint8_t *s = reinterpret_cast<const uint8_t *>(sm.s_words);
    Geo g;
    const bool use5 = tile_sort(sm, enc, n, W, tile, g);
    if (use5) {
        finish_sort<5, true>(sm, s, g.ev, g.halo, g.base, true, sm.ctr, o.packed, W);
        parse_walk<5>(sm, g, W, n, true, tile, o);
    } else {
        finish_sort<3, true>(sm, s, g.ev, g.halo, g.base, false, sm.ctr, o.packed, W);
        parse_walk<3>(sm, g, W, n, false, tile, o);
    }
}
// Tiles are handed out in order (a ticket per CTA), so the tile a CTA waits for is always running
// or finished.
__global__ void __launch_bounds__(tile::THREADS, 2) k_match_parse(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                               size_t tiles, tile::ParseOut o) {
    __shared__ uint32_t ticket;
    if (threadIdx.x == 0) ticket = atomicAdd(o.ctl, 1u);
    __syncthreads();
    const size_t t = ticket;
    if (t >= tiles) return;
    match_parse_body(enc, n, W, t, o);
}

__device__ __forceinline__ void match_tile_body(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                uint32_t *__restrict__ packed, size_t first_tile) {
    using namespace tile;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const uint8_t *s = reinterpret_cast<const uint8_t *>(sm.s_words);
    Geo g;
    const bool use5 = tile_sort(sm, enc, n, W, first_tile + blockIdx.x, g);
    uint16_t *order = sm.ctr;  // the slots the candidate phase takes, in sorted order
    const uint32_t *arr32 = reinterpret_cast<const uint32_t *>(sm.a);
    if (use5) {
        const uint32_t n_order = finish_sort<5, false>(sm, s, g.ev, g.halo, g.base, true, order, packed);
        candidates<5>(sm, s, arr32, order, n_order, g.halo, W, g.base, n, g.avail, true, packed);
    } else {
        const uint32_t n_order = finish_sort<3, false>(sm, s, g.ev, g.halo, g.base, false, order, packed);
        candidates<3>(sm, s, arr32, order, n_order, g.halo, W, g.base, n, g.avail, false, packed);
    }
}
__global__ void __launch_bounds__(tile::THREADS) k_match_tile(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                              uint32_t *__restrict__ packed, size_t first_tile) {
    match_tile_body(enc, n, W, packed, first_tile);
}
// one tile of one file of a batch (blockIdx.y = file)
__global__ void __launch_bounds__(tile::THREADS) kb_match_tile(const LzFile *__restrict__ files,
                                                               uint32_t *__restrict__ packed, size_t packed_stride) {
    const LzFile &f = files[blockIdx.y];
    if ((size_t)blockIdx.x * tile::T >= f.en) return;
    match_tile_body(f.enc, (size_t)f.en, f.W, packed + (size_t)blockIdx.y * packed_stride, 0);
}

// cudaFuncSetAttribute is per device: remember which devices have seen it
static int tile_attr(const void *fn) {
    static std::atomic<uint64_t> done[4] = {{0}, {0}, {0}, {0}};  // bit = device, one word per kernel
    const int which = fn == (const void *)k_match_tile ? 0 : fn == (const void *)kb_match_tile ? 1 : 2;
    int dev = 0;
    RSN_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (done[which].load(std::memory_order_acquire) & bit) return RSN_OK;
    RSN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tile::Smem)));
    done[which].fetch_or(bit, std::memory_order_release);
    return RSN_OK;
}

int lzss_match_tile(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, cudaStream_t s) {
    return lzss_match_tile_range(d_enc, n, W, d_packed, 0, div_up(n, tile::T), true, s);
}

// Tiles [tile_lo, tile_hi) only (positions tile*T ...); the caller guarantees that the bytes up to
// min(n, tile_hi*T + W) are in place.
int lzss_match_tile_range(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, size_t tile_lo,
                          size_t tile_hi, bool, cudaStream_t s) {
    RSN_TRY(tile_attr((const void *)k_match_tile));
    if (tile_hi > tile_lo)
        RSN_LAUNCH(k_match_tile, (unsigned)(tile_hi - tile_lo), tile::THREADS, sizeof(tile::Smem), s, d_enc, n, W, d_packed,
                   tile_lo);
    return RSN_OK;
}
size_t lzss_match_tile_size() { return tile::T; }

// Match search fused with the parse (variant A): orbit bitmap, per-block output sizes and the match
// of every orbit point.  d_spec: one u32 per tile; d_ctl: 2 u32 (zeroed here).  After the kernel
// d_ctl[1] != 0 means the result must be discarded (see parse_walk).
int lzss_match_parse(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, uint16_t *d_visited,
                     uint64_t *d_blk_bytes, uint32_t *d_spec, uint32_t *d_ctl, cudaStream_t s) {
    RSN_TRY(tile_attr((const void *)k_match_parse));
    const size_t tiles = div_up(n, tile::T);
    RSN_CUDA(cudaMemsetAsync(d_spec, 0xFF, tiles * 4, s));
    RSN_CUDA(cudaMemsetAsync(d_ctl, 0, 32, s));
    if (getenv("RSN_PARSE_STATS")) RSN_CUDA(cudaMemsetAsync(d_ctl + 7, 1, 1, s));
    tile::ParseOut o{d_packed, d_visited, d_blk_bytes, d_spec, d_ctl};
    RSN_LAUNCH(k_match_parse, (unsigned)tiles, tile::THREADS, sizeof(tile::Smem), s, d_enc, n, W, tiles, o);
    return RSN_OK;
}

// every file of a batch: files[f].enc / en / W are device-resident (en <= ecap, W <= window <= 4096)
int lzss_match_tile_batch(const LzFile *d_files, size_t G, size_t ecap, uint32_t window, uint32_t *d_packed,
                          size_t packed_stride, cudaStream_t s) {
    (void)window;
    RSN_TRY(tile_attr((const void *)kb_match_tile));
    const dim3 grid((unsigned)div_up(ecap, tile::T), (unsigned)G);
    RSN_LAUNCH(kb_match_tile, grid, tile::THREADS, sizeof(tile::Smem), s, d_files, d_packed, packed_stride);
    return RSN_OK;
}

}  // namespace rsn
