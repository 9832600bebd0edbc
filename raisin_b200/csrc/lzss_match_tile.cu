// lzss_match_tile.cu — K2 for windows up to 4096 (the engine's window, lzss.go:35): every
// position's longest match from a tile staged in shared memory.
//
// One CTA owns T = 8192 consecutive positions plus a halo of W+2 earlier bytes and W bytes of
// look-ahead, all staged in shared memory.  The positions e of halo+tile ("entries") are sorted
// by their 3-gram (enc[e], enc[e+1], enc[e+2]) with a stable LSD radix sort, least significant
// byte = enc[e+2] first.  Stability plus that digit order give, for free:
//   after pass 1: entries grouped by enc[e+2], in position order  -> nearest earlier equal byte
//                 of position q = e+2                              -> "L(q) >= 1"
//   after pass 2: grouped by (enc[e+1], enc[e+2])                  -> "L(q) >= 2" for q = e+1
//   after pass 3: grouped by the whole 3-gram, in position order   -> "L(q) >= 3" for q = e and
//                 the complete, contiguous candidate list of every position for L >= 4.
// A radix pass ranks 32 entries at a time with 8 warp ballots (peers with the same digit), per-warp
// digit counters (no atomics) and one block scan.
//
// Candidates of a position are then evaluated far to near (the list is in position order), with
// 4-byte unaligned shared-memory compares; the scan stops as soon as the remaining distances
// cannot beat the best length, which also keeps degenerate inputs (runs, short periods) bounded:
// L(i) = max_d min(lcp(i-d,i), d, n-i), ties to the larger d (leftmost source, lzss.go:166-184).
#include "lzss.cuh"

#include <atomic>

namespace rsn {

namespace tile {

constexpr int T = 8192;             // positions per CTA
constexpr int WMAX = 4096;          // largest window handled here
constexpr int EMAX = T + WMAX + 24; // entries: halo (W+2, rounded down to 16 bytes) + tile
constexpr int ECAP = EMAX + 40;     // rounded for warp chunks
constexpr int SLEN = EMAX + WMAX + 32;
constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;

struct Smem {
    uint32_t s_words[SLEN / 4 + 4];   // staged bytes: [base, base + avail), zero padded
    uint16_t a[ECAP];                 // ping
    uint16_t b[ECAP];                 // pong
    alignas(16) uint16_t ctr[16 * THREADS];  // radix: [digit][thread]; later: work-class counters
    uint32_t heads[ECAP / 32 + 2];    // bit r: slot r starts a 3-gram group
    uint8_t info[ECAP];               // per slot: work class of its candidate list
    uint8_t lowL[T];                  // 0..3 from the 1/2/3-gram stages, 0xFF once the final result is written
    uint32_t scan[33];
    uint32_t wtot[WARPS + 1];
    // Diagonal cache for long matches: entry = (d << 32) | (start << 16) | end records that
    // s[x] == s[x - d] for every staged x in [start, end).  Entries are only ever written after the
    // bytes were compared, the data never changes, so any entry read (even a racing one) is true.
    unsigned long long diag[16 * 8];  // 16 sets (d & 15) x 8 ways
    unsigned long long mbar;          // mbarrier of the bulk copy that stages the tile
};

constexpr int kVoteSteps = 8;  // candidates a lane may walk between two votes (3: 5.16 ms, 8: 4.99 ms, 16: 5.29 ms on 64 MiB text)

// The tile's bytes come in with ONE bulk asynchronous copy (cp.async.bulk, the 1-D form of the TMA
// path): thread 0 arms an mbarrier with the byte count and issues the copy, the copy engine moves
// global -> shared without passing through registers, and every thread waits on the barrier's phase.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_tile(void *dst_smem, const void *src_global, uint32_t bytes,
                                               unsigned long long *mbar) {
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_global), "r"(bytes), "r"(bar)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
    }
}

__device__ __forceinline__ uint32_t lds32(const uint8_t *s, uint32_t pos) {
    const uint32_t a = pos & ~3u;
    const uint32_t lo = *reinterpret_cast<const uint32_t *>(s + a);
    const uint32_t hi = *reinterpret_cast<const uint32_t *>(s + a + 4);
    return __funnelshift_r(lo, hi, (pos & 3u) * 8);
}

// One stable counting pass on the 4-bit digit (s[e + byteoff] >> shift) & 15 from src to dst over
// slots [0, ev).  Thread t owns the contiguous slots [t*per, (t+1)*per) and a private column of
// 16 digit counters ctr[digit][t] (plain shared-memory increments, no atomics, no ballots); one
// block-wide exclusive scan over the counters in (digit, thread) order turns them into stable
// destinations.  Two passes (low nibble, high nibble) sort by one byte.
__device__ __forceinline__ void radix_pass4(Smem &sm, const uint8_t *s, const uint16_t *src, uint16_t *dst,
                                            uint32_t ev, uint32_t byteoff, uint32_t shift) {
    const uint32_t t = threadIdx.x;
    const uint32_t per = (ev + THREADS - 1) / THREADS;
    const uint32_t lo = min(ev, t * per), hi = min(ev, lo + per);
    uint16_t *col = sm.ctr + t;
    // digit counts of my slots in two packed registers (16 x 8-bit fields; a thread owns <= 255 slots),
    // so the loop is independent loads + register arithmetic with no shared-memory read-modify-write
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
        // thread t scans the 16 consecutive counters [16t, 16t+16) of the linear (digit, thread) order
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
        const uint32_t rank = (uint32_t)(((d < 8 ? acc0 : acc1) >> sh) & 0xFFu);  // earlier slots of mine, same digit
        const uint64_t inc = 1ull << sh;
        acc0 += d < 8 ? inc : 0ull;
        acc1 += d < 8 ? 0ull : inc;
        dst[col[d * THREADS] + rank] = (uint16_t)e;
    }
    __syncthreads();
}

// stable sort by the byte s[e + byteoff]: src -> (tmp) -> src; returns with the result in src
__device__ __forceinline__ void radix_byte(Smem &sm, const uint8_t *s, uint16_t *src, uint16_t *tmp, uint32_t ev,
                                           uint32_t byteoff) {
    radix_pass4(sm, s, src, tmp, ev, byteoff, 0);
    radix_pass4(sm, s, tmp, src, ev, byteoff, 4);
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

// Candidates of the sorted slot r (entry e): the slots [lo, r) of its 3-gram group whose positions
// lie inside the window, in position order (farthest first).
__device__ __forceinline__ void slot_range(const Smem &sm, const uint16_t *arr, uint32_t r, uint32_t e, uint32_t W,
                                           uint32_t &lo_out, uint32_t &cnt_out) {
    // start of the group: highest head bit at or below r
    uint32_t wi = r >> 5;
    uint32_t bits = sm.heads[wi] & (0xFFFFFFFFu >> (31 - (r & 31)));
    while (bits == 0) bits = sm.heads[--wi];
    uint32_t lo = (wi << 5) + (31 - __clz(bits)), hi = r;
    // first candidate inside the window: lowest slot in [group start, r) with position >= e - W
    if (e > W && lo < hi && arr[lo] < e - W) {
        const uint32_t minpos = e - W;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (arr[mid] < minpos) lo = mid + 1;
            else hi = mid;
        }
    }
    lo_out = lo;
    cnt_out = r - lo;
}

}  // namespace tile

__device__ __forceinline__ void match_tile_body(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                uint32_t *__restrict__ packed, size_t first_tile) {
    using namespace tile;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    uint8_t *s = reinterpret_cast<uint8_t *>(sm.s_words);
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    const size_t tile_start = (first_tile + blockIdx.x) * T;
    const uint32_t tile_len = (uint32_t)min((size_t)T, n - tile_start);
    const size_t base = tile_start > (size_t)W + 2 ? ((tile_start - W - 2) & ~(size_t)15) : 0;  // 16-byte aligned
    const uint32_t halo = (uint32_t)(tile_start - base);
    const uint32_t avail = (uint32_t)min(n - base, (size_t)(halo + T + W));  // bytes staged
    // entries: positions with a complete 3-gram, up to the end of the tile
    const uint32_t ev = avail >= 3 ? min(halo + tile_len, avail - 2) : 0;

    // ---- stage bytes (zero padded), identity order, clear flags
    {
        const uint32_t nwords = (avail + 3) / 4;
        const uint8_t *g = enc + base;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            const uint32_t fullv = avail / 16;
            if (fullv) bulk_load_tile(sm.s_words, g, fullv * 16, &sm.mbar);
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
        for (uint32_t i = nwords + threadIdx.x; i < nwords + 4 && i < SLEN / 4 + 4; i += THREADS) sm.s_words[i] = 0;
        for (uint32_t e = threadIdx.x; e < ev; e += THREADS) sm.a[e] = (uint16_t)e;
        for (uint32_t i = threadIdx.x; i < T; i += THREADS) sm.lowL[i] = 0;
    }
    __syncthreads();

    // ---- pass 1: by enc[e+2]  ->  1-byte matches of q = e+2
    radix_byte(sm, s, sm.a, sm.b, ev, 2);
    for (uint32_t r = threadIdx.x; r < ev; r += THREADS) {
        const uint32_t e = sm.a[r], q = e + 2;
        if (q >= halo && q < halo + tile_len && r > 0) {
            const uint32_t p = sm.a[r - 1];
            if (s[p + 2] == s[q] && e - p <= W) sm.lowL[q - halo] = 1;
        }
    }
    __syncthreads();
    // ---- pass 2: by (enc[e+1], enc[e+2])  ->  2-byte matches of q = e+1 at distance >= 2
    radix_byte(sm, s, sm.a, sm.b, ev, 1);
    for (uint32_t r = threadIdx.x; r < ev; r += THREADS) {
        const uint32_t e = sm.a[r], q = e + 1;
        if (q >= halo && q < halo + tile_len) {
            const uint32_t key = lds32(s, q) & 0xFFFFu;
            bool hit = false;
            for (uint32_t k = 1; k <= 2 && k <= r; k++) {
                const uint32_t p = sm.a[r - k];
                if ((lds32(s, p + 1) & 0xFFFFu) != key) break;
                const uint32_t d = e - p;
                if (d >= 2) {
                    hit = d <= W;
                    break;
                }
            }
            if (hit) sm.lowL[q - halo] = 2;
        }
    }
    __syncthreads();
    // ---- pass 3: by the 3-gram  ->  candidate lists in position order
    radix_byte(sm, s, sm.a, sm.b, ev, 0);
    const uint16_t *arr = sm.a;  // sorted by (3-gram, position)
    uint16_t *order = sm.b;      // spare buffer: tile slots that have candidates, grouped by work class
    // ---- group heads
    {
        const uint32_t per = ((ev + WARPS - 1) / WARPS + 31) & ~31u;
        const uint32_t lo = min(ev, w * per), hi = min(ev, lo + per);
        for (uint32_t c = lo; c < hi; c += 32) {
            const uint32_t r = c + lane;
            bool head = false;
            if (r < hi) head = r == 0 || ((lds32(s, arr[r]) ^ lds32(s, arr[r - 1])) & 0xFFFFFFu) != 0;
            const unsigned hm = __ballot_sync(0xffffffffu, head);
            if (lane == 0) sm.heads[c >> 5] = hm;
        }
    }
    for (int i = threadIdx.x; i < 16 * WARPS; i += THREADS) sm.ctr[i] = 0;
    if (threadIdx.x < 128) sm.diag[threadIdx.x] = 0;  // d = 0 never matches a real distance
    __syncthreads();
    // ---- per tile slot: candidate range, work class.  Slots without candidates are final here.
    // Lanes of a warp later take slots of ONE class (similar candidate counts), which is what keeps
    // the candidate loop from idling most lanes behind the few slots of very frequent 3-grams.
    {
        const unsigned lt = (1u << lane) - 1;
        const uint32_t per = ((ev + WARPS - 1) / WARPS + 31) & ~31u;
        const uint32_t lo_s = min(ev, w * per), hi_s = min(ev, lo_s + per);
        for (int phase = 0; phase < 2; phase++) {
            for (uint32_t c0 = lo_s; c0 < hi_s; c0 += 32) {
                const uint32_t r = c0 + lane;
                uint32_t cls = 0xFF;
                if (phase == 0) {
                    if (r < hi_s) {
                        const uint32_t e = arr[r];
                        if (e >= halo) {
                            uint32_t lo, cnt;
                            slot_range(sm, arr, r, e, W, lo, cnt);
                            if (cnt == 0) {  // no earlier occurrence of this 3-gram in the window
                                packed[base + e] = (uint32_t)sm.lowL[e - halo] << 16;
                                sm.lowL[e - halo] = 0xFF;
                            } else {
                                cls = 31 - __clz(cnt);  // 0..12: floor(log2(count))
                                cls = cls > 7 ? 7 : cls;
                            }
                        }
                        sm.info[r] = (uint8_t)cls;
                    }
                } else if (r < hi_s) {
                    cls = sm.info[r];
                }
                const bool act = cls != 0xFF;
                unsigned peers = __ballot_sync(0xffffffffu, act);
#pragma unroll
                for (int bit = 0; bit < 3; bit++) {
                    const unsigned m = __ballot_sync(0xffffffffu, (cls >> bit) & 1u);
                    peers &= ((cls >> bit) & 1u) ? m : ~m;
                }
                if (act) {
                    const uint32_t rank = __popc(peers & lt);
                    const bool last = (peers >> lane) == 1u;
                    if (phase == 0) {
                        if (last) sm.ctr[cls * WARPS + w] += (uint16_t)(rank + 1);
                    } else {
                        const uint32_t basev = sm.ctr[cls * WARPS + w];
                        order[basev + rank] = (uint16_t)r;
                        __syncwarp(peers);
                        if (last) sm.ctr[cls * WARPS + w] = (uint16_t)(basev + rank + 1);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
            if (phase == 0) {
                if (threadIdx.x == 0) {  // exclusive scan of the 8 x WARPS class counters, class-major
                    uint32_t run = 0;
                    for (int k = 0; k < 8 * WARPS; k++) {
                        const uint32_t v = sm.ctr[k];
                        sm.ctr[k] = (uint16_t)run;
                        run += v;
                    }
                    sm.wtot[0] = run;
                }
                __syncthreads();
            }
        }
    }
    const uint32_t n_order = sm.wtot[0];

    // ---- candidates, far to near; one thread per slot, slots of similar work side by side
    for (uint32_t k0 = w * 32; k0 < n_order; k0 += THREADS) {
        const uint32_t k = k0 + lane;
        const bool act = k < n_order;
        uint32_t r = 0, e = 0, room = 0, c = 0, best = 3, boff = 0;
        bool has3 = false;
        if (act) {
            r = order[k];
            e = arr[r];
            room = (uint32_t)min((size_t)W, n - (base + e));
            uint32_t lo, cnt;
            slot_range(sm, arr, r, e, W, lo, cnt);
            c = lo;
            // arr[lo] is the farthest candidate: a 3-byte match exists iff its distance is >= 3
            has3 = lo < r && e - arr[lo] >= 3;
        }
        // Far to near.  A candidate at distance d yields at most min(d, room), so only slots with
        // arr[c] < jlim = e - best can win (the list is in position order), and a winner must match
        // the byte at offset `best` (tgt).  Lanes of a warp hold slots of one work class.
        uint32_t jlim = 0, tgt = 0;
        const uint8_t *sb = s;
        if (act && room > best) {
            jlim = e - best;
            sb = s + best;
            tgt = s[e + best];
        } else {
            c = r;
        }
        // The loop runs in warp-wide rounds: every lane walks up to kVoteSteps candidates of its list
        // (most fail the byte filter) or until one survives, then all lanes that hold a survivor
        // compare together.  With the comparison inside a plain per-lane walk it ran with 3-4 of 32
        // lanes (text and logs alike); an unbounded walk per round left the walk itself at 7 lanes.
        for (;;) {
            bool have = false;
            uint32_t j = 0;
            // at most kVoteSteps candidates per lane and round: lanes that found a survivor wait
            // only that long for the others, and a round still gathers survivors from many lanes
#pragma unroll 1
            for (int step = 0; step < kVoteSteps && c < r; step++) {
                j = arr[c];
                if (j >= jlim) {  // nearer candidates yield even less
                    c = r;
                    break;
                }
                c++;
                if (sb[j] == tgt) {
                    have = true;
                    break;
                }
            }
            if (!__any_sync(0xffffffffu, have)) {
                if (!__any_sync(0xffffffffu, c < r)) break;
                continue;
            }
            if (have) {
                const uint32_t d = e - j;
                const uint32_t cap = min(d, room);
                uint32_t l = 3;
                while (l < cap && l < 35) {
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
                    boff = e - j;
                    if (room <= best) {
                        c = r;
                    } else {
                        jlim = e - best;
                        sb = s + best;
                        tgt = s[e + best];
                    }
                }
            }
        }
        if (act) {
            uint32_t L, off = 0;
            if (best >= 4) {
                L = best;
                off = boff;
            } else if (has3 && room >= 3) {
                L = 3;
            } else {
                L = sm.lowL[e - halo];
            }
            sm.lowL[e - halo] = 0xFF;  // final result written
            packed[base + e] = (L << 16) | off;
        }
    }
    __syncthreads();
    // positions of the tile that have no entry (no complete 3-gram: the last two of the stream)
    for (uint32_t x = threadIdx.x; x < tile_len; x += THREADS) {
        const uint8_t v = sm.lowL[x];
        if (v != 0xFF) packed[tile_start + x] = (uint32_t)v << 16;
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

// Tile 0 has no halo, so stream positions 0 and 1 never appear as the third byte of an entry and
// position 0 never as the second: add the 1- and 2-byte matches whose source starts there.
__device__ __forceinline__ void match_tile_fix0_body(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                     uint32_t *__restrict__ packed) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || q > (size_t)W + 1) return;
    uint32_t L = packed[q] >> 16;
    if (L >= 3) return;
    uint32_t want = L;
    // 2-gram at source 0 (distance q >= 2, q <= W)
    if (q >= 2 && q <= W && q + 2 <= n && enc[0] == enc[q] && enc[1] == enc[q + 1]) want = max(want, 2u);
    // single byte at source 0 (distance q) or 1 (distance q-1)
    if (q >= 1 && q <= W && enc[0] == enc[q]) want = max(want, 1u);
    if (q >= 2 && q - 1 <= W && enc[1] == enc[q]) want = max(want, 1u);
    if (want != L) packed[q] = want << 16;
}
__global__ void k_match_tile_fix0(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                  uint32_t *__restrict__ packed) {
    match_tile_fix0_body(enc, n, W, packed);
}
__global__ void kb_match_tile_fix0(const LzFile *__restrict__ files, uint32_t *__restrict__ packed,
                                   size_t packed_stride) {
    const LzFile &f = files[blockIdx.y];
    match_tile_fix0_body(f.enc, (size_t)f.en, f.W, packed + (size_t)blockIdx.y * packed_stride);
}

// cudaFuncSetAttribute is per device: remember which devices have seen it
static int tile_attr(const void *fn) {
    static std::atomic<uint64_t> done[2] = {{0}, {0}};  // bit = device; [0] k_match_tile, [1] kb_match_tile
    const int which = fn == (const void *)k_match_tile ? 0 : 1;
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
// min(n, tile_hi*T + W) are in place.  fix0 must be run once tile 0 and tile 1 are done.
int lzss_match_tile_range(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, size_t tile_lo,
                          size_t tile_hi, bool run_fix0, cudaStream_t s) {
    RSN_TRY(tile_attr((const void *)k_match_tile));
    const size_t smem = sizeof(tile::Smem);
    if (tile_hi > tile_lo)
        RSN_LAUNCH(k_match_tile, (unsigned)(tile_hi - tile_lo), tile::THREADS, smem, s, d_enc, n, W, d_packed, tile_lo);
    if (run_fix0)
        RSN_LAUNCH(k_match_tile_fix0, (unsigned)div_up(min(n, (size_t)W + 2), 128), 128, 0, s, d_enc, n, W, d_packed);
    return RSN_OK;
}
size_t lzss_match_tile_size() { return tile::T; }

// every file of a batch: files[f].enc / en / W are device-resident (en <= ecap, W <= window <= 4096)
int lzss_match_tile_batch(const LzFile *d_files, size_t G, size_t ecap, uint32_t window, uint32_t *d_packed,
                          size_t packed_stride, cudaStream_t s) {
    RSN_TRY(tile_attr((const void *)kb_match_tile));
    const size_t smem = sizeof(tile::Smem);
    const dim3 grid((unsigned)div_up(ecap, tile::T), (unsigned)G);
    RSN_LAUNCH(kb_match_tile, grid, tile::THREADS, smem, s, d_files, d_packed, packed_stride);
    const dim3 gfix((unsigned)div_up((size_t)window + 2, 128), (unsigned)G);
    RSN_LAUNCH(kb_match_tile_fix0, gfix, 128, 0, s, d_files, d_packed, packed_stride);
    return RSN_OK;
}

}  // namespace rsn
