// lzss_match_chain.cu — K2 for windows up to 4096 (the engine's window, lzss.go:35), round-2 form:
// every position's longest match (compressorWorker, lzss.go:156-184) from hash chains that are
// built in shared memory without a sort.
//
// One CTA owns T = 8192 consecutive positions plus a halo of W earlier bytes and W bytes of
// look-ahead, all staged in shared memory ("entries" = positions of halo + tile).
//
//   L(i) = max_{1<=d<=min(i,W)} min(lcp(i-d, i), d, n-i),   off(i) = the largest d attaining it.
//
// A chain of the k-gram starting at every entry needs the entries of one hash bucket in position
// order.  Entries are first split, stably, into 64 classes by a hash of their first two bytes;
// every k-gram (k >= 2) of a class lives in buckets that belong to that class alone, so ONE warp
// can take a class and append its entries to their buckets in position order, 32 per step, with
// plain shared-memory stores: lanes of a step that fall into the same bucket find each other with
// match.any and link up among themselves, the first of them links to the bucket's previous head.
// Every bucket also keeps a tail pointer (oldest entry still inside the window), advanced as the
// positions move on, so an entry learns its FARTHEST in-window candidate in amortised O(1).
// Three such passes over the same class lists:
//   k = 2, 3: far-to-near walk to the first verified k-gram at distance >= k  ->  "L >= k"
//   k = 4:    the same walk leaves the farthest verified 4-gram candidate of every position;
//             positions that have one go to a work list.
// L >= 1 needs no chain: a 256 x 32-bit table says which 512-byte blocks of the staged range
// contain a byte value; the blocks that lie wholly inside a position's window decide (the rare
// miss scans the two ragged ends).
//
// The work list is then evaluated far to near along the 4-gram chain (position order = chain
// order), lanes refilling themselves from the list as they finish, with a one-byte "must beat
// the best" filter, 4-byte unaligned shared-memory compares, the exact cut-off d <= best, and the
// diagonal cache for matches beyond 32 bytes.
#include "lzss.cuh"

#include <atomic>

namespace rsn {

namespace chain {

constexpr int T = 8192;             // positions per CTA
constexpr int WMAX = 4096;          // largest window handled here
constexpr int EMAX = T + WMAX + 16; // entries: halo (W, rounded down to 16 bytes) + tile
constexpr int SLEN = EMAX + WMAX + 32;
constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int NCLS = 64;            // classes (hash of the first two bytes)
constexpr int BPC = 64;             // buckets per class
constexpr int NB = NCLS * BPC;
constexpr uint32_t NONE = 0xFFFFu;

struct Smem {
    uint32_t s_words[SLEN / 4 + 4];   // staged bytes: [base, base + avail), zero padded
    uint16_t lists[EMAX + 32];        // entries grouped by class, position order inside a class; later the work list
    uint16_t link[EMAX + 32];         // next entry of the same bucket (ascending positions), NONE at the end
    alignas(16) uint16_t head[NB];    // newest entry of a bucket
    alignas(16) uint16_t tail[NB];    // oldest entry of a bucket that can still be inside a window
    uint16_t far[T];                  // per tile position: farthest verified 4-gram candidate
    uint8_t lowL[T];                  // 0..3 from the k-gram passes, 0xFF once the final result is written
    uint16_t cnt[NCLS * WARPS];       // partition counters, [class][warp]
    uint32_t cstart[NCLS + 1];
    uint32_t bm[256];                 // byte value -> set of 512-entry blocks that contain it
    unsigned long long diag[16 * 8];  // diagonal cache (see long_lcp)
    uint32_t scan[33];
    uint32_t claim;                   // next class to take in a chain pass
    uint32_t n_miss;
    uint32_t work_next;
};

__device__ __forceinline__ uint32_t lds32(const uint8_t *s, uint32_t pos) {
    const uint32_t a = pos & ~3u;
    const uint32_t lo = *reinterpret_cast<const uint32_t *>(s + a);
    const uint32_t hi = *reinterpret_cast<const uint32_t *>(s + a + 4);
    return __funnelshift_r(lo, hi, (pos & 3u) * 8);
}

__device__ __forceinline__ uint32_t class_of(uint32_t g) { return ((g & 0xFFFFu) * 0x9E3779B1u) >> 26; }

template <int K>
__device__ __forceinline__ uint32_t bucket_of(uint32_t cls, uint32_t g) {
    constexpr uint32_t mask = K == 2 ? 0xFFFFu : K == 3 ? 0xFFFFFFu : 0xFFFFFFFFu;
    constexpr uint32_t mul = K == 2 ? 0x85EBCA6Bu : K == 3 ? 0xC2B2AE35u : 0x27D4EB2Fu;
    return cls * BPC + (((g & mask) * mul) >> 26);
}

// Continuation of a match that is already 32+ bytes long: compare on, but consult and feed the
// diagonal cache so that the thousands of positions of a tile that sit on the same long diagonal
// run (highly repetitive data) do not each re-compare it.  entry = (d << 32) | (start << 16) | end
// records that s[x] == s[x - d] for every staged x in [start, end); entries are only ever written
// after the bytes were compared and the data never changes, so any entry read (even a racing one)
// is true.  Returns the (uncapped) match length.
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

// Chain pass for K = 2, 3 ("is there a K-gram match"): buckets hold the newest entry, every entry
// links to the previous one of its bucket, and a tile position walks near to far to the first
// verified K-gram at distance >= K inside its window (usually the first hop).  Raises lowL to K.
template <int K>
__device__ __forceinline__ void chain_pass_near(Smem &sm, const uint8_t *s, uint32_t halo, uint32_t W, uint32_t nrel) {
    constexpr uint32_t mask = K == 2 ? 0xFFFFu : 0xFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1, gt = ~lt & ~(1u << lane);
    const int16_t *plink = reinterpret_cast<const int16_t *>(sm.link);  // NONE reads as -1
    {
        uint4 *p = reinterpret_cast<uint4 *>(sm.head);
        const uint4 none4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t i = threadIdx.x; i < NB * 2 / 16; i += THREADS) p[i] = none4;
        if (threadIdx.x == 0) sm.claim = 0;
    }
    __syncthreads();
    for (;;) {
        uint32_t cls = 0;
        if (lane == 0) cls = atomicAdd(&sm.claim, 1u);
        cls = __shfl_sync(0xffffffffu, cls, 0);
        if (cls >= NCLS) break;
        const uint32_t lo = sm.cstart[cls], hi = sm.cstart[cls + 1];
        for (uint32_t s0 = lo; s0 < hi; s0 += 32) {
            const uint32_t slot = s0 + lane;
            const bool valid = slot < hi;
            const uint32_t e = valid ? sm.lists[slot] : 0u;
            const uint32_t g = lds32(s, e);
            const uint32_t b = valid ? bucket_of<K>(cls, g) : (0x10000u | lane);
            const unsigned m = __match_any_sync(0xffffffffu, b);
            const unsigned below = m & lt, above = m & gt;
            const uint32_t e_pred = __shfl_sync(0xffffffffu, e, below ? 31 - __clz(below) : lane);
            int c = -1;
            if (valid) {
                c = below ? (int)e_pred : (int)reinterpret_cast<const int16_t *>(sm.head)[b];
                sm.link[e] = (uint16_t)c;
            }
            __syncwarp();
            if (valid) {
                if (!above) sm.head[b] = (uint16_t)e;
                if (e >= halo) {
                    const uint32_t x = e - halo;
                    bool want = nrel - e >= (uint32_t)K && e >= (uint32_t)K;
                    if (K == 2) want = want && sm.lowL[x] < 2;
                    if (K > 2) want = want && sm.lowL[x] == K - 1;  // no (K-1)-gram match, no K-gram match
                    if (want) {
                        const int minpos = e > W ? (int)(e - W) : 0, limit = (int)e - K;
                        bool hit = false;
                        while (c >= minpos) {
                            if (c <= limit && ((lds32(s, (uint32_t)c) ^ g) & mask) == 0) {
                                hit = true;
                                break;
                            }
                            c = plink[c];
                        }
                        if (hit) sm.lowL[x] = (uint8_t)K;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

// Chain pass for K = 4: entries link to the NEXT one of their bucket, and every bucket keeps a tail
// (oldest entry that can still be inside a window), so a tile position gets the farthest in-window
// entry of its bucket in amortised O(1) and walks far to near to the first verified 4-gram at
// distance >= 4: far[x], the start of the candidate evaluation.
__device__ __forceinline__ void chain_pass_far(Smem &sm, const uint8_t *s, uint32_t halo, uint32_t W, uint32_t nrel) {
    constexpr int K = 4;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1, gt = ~lt & ~(1u << lane);
    {
        uint4 *p = reinterpret_cast<uint4 *>(sm.head);  // head and tail are adjacent
        const uint4 none4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t i = threadIdx.x; i < 2 * NB * 2 / 16; i += THREADS) p[i] = none4;
        if (threadIdx.x == 0) sm.claim = 0;
    }
    __syncthreads();
    for (;;) {
        uint32_t cls = 0;
        if (lane == 0) cls = atomicAdd(&sm.claim, 1u);
        cls = __shfl_sync(0xffffffffu, cls, 0);
        if (cls >= NCLS) break;
        const uint32_t lo = sm.cstart[cls], hi = sm.cstart[cls + 1];
        for (uint32_t s0 = lo; s0 < hi; s0 += 32) {
            const uint32_t slot = s0 + lane;
            const bool valid = slot < hi;
            const uint32_t e = valid ? sm.lists[slot] : 0u;
            const uint32_t g = lds32(s, e);
            const uint32_t b = valid ? bucket_of<K>(cls, g) : (0x10000u | lane);
            const unsigned m = __match_any_sync(0xffffffffu, b);
            const unsigned below = m & lt, above = m & gt;
            const uint32_t e_succ = __shfl_sync(0xffffffffu, e, above ? __ffs(above) - 1 : lane);
            uint32_t t = e;
            if (valid && !below) {  // first of its bucket in this step: link behind the old head
                const uint32_t old = sm.head[b];
                const uint32_t t0 = sm.tail[b];
                if (old != NONE) sm.link[old] = (uint16_t)e;
                if (t0 != NONE) t = t0;
            }
            if (valid) sm.link[e] = (uint16_t)(above ? e_succ : NONE);
            t = __shfl_sync(0xffffffffu, t, __ffs(m) - 1);
            __syncwarp();
            if (valid) {
                if (!above) sm.head[b] = (uint16_t)e;
                const uint32_t minpos = e > W ? e - W : 0u;
                while (t < minpos) t = sm.link[t];  // reaches e at the latest
                if (!above) sm.tail[b] = (uint16_t)t;
                if (e >= halo) {
                    const uint32_t x = e - halo;
                    const bool want = nrel - e >= (uint32_t)K && e >= (uint32_t)K && sm.lowL[x] == K - 1;
                    uint32_t c = t;
                    bool hit = false;
                    if (want) {
                        const uint32_t limit = e - K;  // source must end at or before e
                        while (c <= limit) {
                            if (lds32(s, c) == g) {
                                hit = true;
                                break;
                            }
                            c = sm.link[c];
                        }
                    }
                    sm.far[x] = (uint16_t)(hit ? c : NONE);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

}  // namespace chain

__device__ __forceinline__ void match_chain_body(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                 uint32_t *__restrict__ packed, size_t first_tile) {
    using namespace chain;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    uint8_t *s = reinterpret_cast<uint8_t *>(sm.s_words);
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1;

    const size_t tile_start = (first_tile + blockIdx.x) * T;
    const uint32_t tile_len = (uint32_t)min((size_t)T, n - tile_start);
    const size_t base = tile_start > (size_t)W ? ((tile_start - W) & ~(size_t)15) : 0;  // 16-byte aligned
    const uint32_t halo = (uint32_t)(tile_start - base);
    const uint32_t avail = (uint32_t)min(n - base, (size_t)(halo + T + W));  // bytes staged
    const uint32_t nrel = (uint32_t)min(n - base, (size_t)0x7FFFFFFFu);      // stream end, relative
    const uint32_t ev = halo + tile_len;                                     // entries

    // ---- stage bytes (zero padded), clear tables
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
        for (uint32_t i = nwords + threadIdx.x; i < nwords + 4 && i < SLEN / 4 + 4; i += THREADS) sm.s_words[i] = 0;
        for (uint32_t i = threadIdx.x; i < T / 4; i += THREADS) reinterpret_cast<uint32_t *>(sm.lowL)[i] = 0;
        for (uint32_t i = threadIdx.x; i < NCLS * WARPS / 2; i += THREADS) reinterpret_cast<uint32_t *>(sm.cnt)[i] = 0;
        if (threadIdx.x < 256) sm.bm[threadIdx.x] = 0;
        if (threadIdx.x < 128) sm.diag[threadIdx.x] = 0;  // d = 0 never matches a real distance
        if (threadIdx.x == 0) {
            sm.n_miss = 0;
            sm.work_next = 0;
        }
    }
    __syncthreads();

    // ---- split the entries into classes (stable): count, scan, write.  Warp w owns a contiguous
    // run of entries; lists[] of a class is the concatenation of the warps' runs, i.e. position order.
    const uint32_t per = ((ev + WARPS - 1) / WARPS + 31) & ~31u;
    const uint32_t lo_e = min(ev, w * per), hi_e = min(ev, lo_e + per);
    for (uint32_t c0 = lo_e; c0 < hi_e; c0 += 32) {
        const uint32_t e = c0 + lane;
        const bool valid = e < hi_e;
        const uint32_t cls = valid ? class_of(lds32(s, e)) : (0x100u | lane);
        const unsigned m = __match_any_sync(0xffffffffu, cls);
        if (valid) {
            if ((m & lt) == 0) sm.cnt[cls * WARPS + w] += (uint16_t)__popc(m);
            // which 512-entry blocks hold this byte value (L >= 1 test at the end)
            const uint32_t v = s[e], bit = 1u << (e >> 9);
            if (!(sm.bm[v] & bit)) atomicOr(&sm.bm[v], bit);
        }
        __syncwarp();
    }
    __syncthreads();
    {
        // exclusive scan of the counters in (class, warp) order: thread t owns counters 2t, 2t+1
        uint32_t *p = reinterpret_cast<uint32_t *>(sm.cnt) + threadIdx.x;
        const uint32_t v = *p;
        const uint32_t a = v & 0xFFFFu, b2 = v >> 16;
        uint32_t total;
        const uint32_t run = block_exclusive_sum<uint32_t>(a + b2, sm.scan, total);
        *p = run | ((run + a) << 16);
        if ((threadIdx.x & (WARPS / 2 - 1)) == 0) sm.cstart[threadIdx.x / (WARPS / 2)] = run;
        if (threadIdx.x == 0) sm.cstart[NCLS] = ev;
    }
    __syncthreads();
    for (uint32_t c0 = lo_e; c0 < hi_e; c0 += 32) {
        const uint32_t e = c0 + lane;
        const bool valid = e < hi_e;
        const uint32_t cls = valid ? class_of(lds32(s, e)) : (0x100u | lane);
        const unsigned m = __match_any_sync(0xffffffffu, cls);
        uint32_t basev = 0;
        if (valid) {
            basev = sm.cnt[cls * WARPS + w];
            sm.lists[basev + __popc(m & lt)] = (uint16_t)e;
        }
        __syncwarp();
        if (valid && (m >> lane) == 1u) sm.cnt[cls * WARPS + w] = (uint16_t)(basev + __popc(m));
        __syncwarp();
    }
    __syncthreads();

    // ---- chains
    chain_pass_near<2>(sm, s, halo, W, nrel);
    chain_pass_near<3>(sm, s, halo, W, nrel);
    chain_pass_far(sm, s, halo, W, nrel);

    // ---- candidates, far to near along the 4-gram chain.  A candidate at distance d yields at most
    // min(d, room), so only entries j < jlim = e - best can win, and a winner must match the byte at
    // offset `best` (tgt).  A lane takes a run of 8 consecutive tile positions and works through
    // those that have a candidate, in order: since L(x+1) >= L(x) - 1 (the same source, one byte on),
    // a position that follows a match of L starts its walk from best = L - 2, so that nearly every
    // candidate of a long repeat fails the one-byte filter.  Every turn of the loop a lane advances
    // up to 4 chain entries and parks up to two survivors of the filter; the survivors are compared
    // when enough lanes hold one (or nobody can walk), so both halves run on mostly full warps.
    {
        const uint32_t n_runs = (tile_len + 7) >> 3;
        bool active = false, drained = false, found = false;
        uint32_t xbase = 0, pend = 0, curx = 0, lastx = 0xFFFFFFF0u, lastL = 0;
        uint32_t e = 0, room = 0, j = NONE, jlim = 0, best = 3, boff = 0, tgt = 0, ew0 = 0, ew1 = 0;
        uint32_t q0 = 0, q1 = 0, nq = 0;
        for (;;) {
            // -- a new run for lanes that have none
            const bool want_run = !active && !drained && pend == 0;
            const unsigned need = __ballot_sync(0xffffffffu, want_run);
            if (need) {
                const int leader = __ffs(need) - 1;
                uint32_t r = 0;
                if ((int)lane == leader) r = atomicAdd(&sm.work_next, (uint32_t)__popc(need));
                r = __shfl_sync(0xffffffffu, r, leader) + __popc(need & lt);
                if (want_run) {
                    if (r < n_runs) {
                        xbase = r * 8;
                        const uint4 f = *reinterpret_cast<const uint4 *>(sm.far + xbase);
                        const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
                        uint32_t mk = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            mk |= ((fw[k] & 0xFFFFu) != 0xFFFFu ? 1u : 0u) << (2 * k);
                            mk |= ((fw[k] >> 16) != 0xFFFFu ? 1u : 0u) << (2 * k + 1);
                        }
                        const uint32_t nvalid = min(8u, tile_len - xbase);
                        pend = mk & ((1u << nvalid) - 1u);
                        lastx = 0xFFFFFFF0u;
                    } else {
                        drained = true;
                    }
                }
            }
            // -- next position of the run
            if (!active && pend) {
                const uint32_t k = __ffs(pend) - 1;
                pend &= pend - 1;
                curx = xbase + k;
                e = curx + halo;
                room = min(W, nrel - e);
                j = sm.far[curx];
                best = (lastx + 1 == curx && lastL >= 6) ? lastL - 2 : 3;
                found = false;
                boff = 0;
                jlim = e - best;
                tgt = s[e + best];
                ew0 = lds32(s, e);
                ew1 = lds32(s, e + 4);
                nq = 0;
                active = true;
            }
            if (!__any_sync(0xffffffffu, active)) {
                if (__all_sync(0xffffffffu, drained)) break;
                continue;
            }
            // -- walk
            if (__any_sync(0xffffffffu, active && nq < 2 && j < jlim)) {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    if (active && nq < 2 && j < jlim) {
                        const uint32_t cand = j;
                        j = sm.link[cand];
                        if (s[cand + best] == tgt) {
                            if (nq == 0) q0 = cand;
                            else q1 = cand;
                            nq++;
                        }
                    }
                }
            }
            // -- compare the parked survivors
            const unsigned pm = __ballot_sync(0xffffffffu, nq > 0);
            if (pm && (__popc(pm) >= 8 || !__any_sync(0xffffffffu, active && nq < 2 && j < jlim))) {
                if (nq > 0) {
                    const uint32_t cand = q0;
                    q0 = q1;
                    nq--;
                    const uint32_t d = e - cand;
                    const uint32_t cap = min(d, room);
                    uint32_t l;
                    uint32_t xw = lds32(s, cand) ^ ew0;
                    if (xw) {
                        l = (__ffs(xw) - 1) >> 3;
                    } else {
                        xw = lds32(s, cand + 4) ^ ew1;
                        if (xw) {
                            l = 4 + ((__ffs(xw) - 1) >> 3);
                        } else {
                            l = 8;
                            while (l < cap && l < 36) {
                                xw = lds32(s, cand + l) ^ lds32(s, e + l);
                                if (xw) {
                                    l += (__ffs(xw) - 1) >> 3;
                                    goto lcp_done;
                                }
                                l += 4;
                            }
                            if (l < cap) l = long_lcp(sm, s, e, d, l, cap, avail);
                        }
                    }
                lcp_done:
                    l = min(l, cap);
                    if (l > best) {
                        best = l;
                        boff = d;
                        found = true;
                        if (room <= best) {
                            j = NONE;
                            nq = 0;
                        } else {
                            jlim = e - best;
                            tgt = s[e + best];
                            if (nq && !(q0 < jlim && s[q0 + best] == tgt)) nq = 0;  // parked under the old filter
                        }
                    }
                }
            }
            // -- done with this position?
            if (active && nq == 0 && j >= jlim) {
                if (!found && best > 3) {  // the inherited bound was not reached (cannot happen): plain walk
                    best = 3;
                    jlim = e - 3;
                    tgt = s[e + 3];
                    j = sm.far[curx];
                } else {
                    if (found) {
                        packed[base + e] = (best << 16) | boff;
                        sm.lowL[curx] = 0xFF;
                        lastL = best;
                        lastx = curx;
                    }
                    active = false;
                }
            }
        }
    }
    __syncthreads();

    // ---- positions without a match of 4 or more: L from the k-gram passes, L >= 1 from the block
    // table; the rare position whose byte is in none of the whole blocks of its window goes to a list
    // and a warp scans the two ragged ends of its window.
    uint16_t *miss = sm.lists;
    for (uint32_t x = threadIdx.x; x < tile_len; x += THREADS) {
        uint32_t v = sm.lowL[x];
        if (v == 0xFF) continue;
        if (v == 0) {
            const uint32_t e = x + halo;
            const uint32_t lo = e > W ? e - W : 0u;  // window [lo, e)
            const uint32_t kfirst = (lo + 511) >> 9, kend = e >> 9;  // whole blocks [kfirst, kend)
            bool hit = false;
            if (kfirst < kend) hit = (sm.bm[s[e]] & ((1u << kend) - 1u) & ~((1u << kfirst) - 1u)) != 0;
            if (!hit) {
                miss[atomicAdd(&sm.n_miss, 1u)] = (uint16_t)x;
                continue;
            }
            v = 1;
        }
        packed[tile_start + x] = v << 16;
    }
    __syncthreads();
    const uint32_t n_miss = sm.n_miss;
    for (uint32_t k = w; k < n_miss; k += WARPS) {
        const uint32_t x = miss[k], e = x + halo;
        const uint32_t lo = e > W ? e - W : 0u;
        const uint32_t kfirst = (lo + 511) >> 9, kend = e >> 9;
        uint32_t a_end = e, b_start = e;  // ragged ends: [lo, a_end) and [b_start, e)
        if (kfirst < kend) {
            a_end = kfirst << 9;
            b_start = kend << 9;
        }
        const uint32_t byte = s[e];
        bool hit = false;
        for (uint32_t q = lo + lane; q < a_end; q += 32) hit |= s[q] == byte;
        for (uint32_t q = b_start + lane; q < e; q += 32) hit |= s[q] == byte;
        hit = __any_sync(0xffffffffu, hit);
        if (lane == 0) packed[tile_start + x] = hit ? (1u << 16) : 0u;
    }
}

__global__ void __launch_bounds__(chain::THREADS, 2) k_match_chain(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                                   uint32_t *__restrict__ packed, size_t first_tile) {
    match_chain_body(enc, n, W, packed, first_tile);
}
// one tile of one file of a batch (blockIdx.y = file)
__global__ void __launch_bounds__(chain::THREADS, 2) kb_match_chain(const LzFile *__restrict__ files,
                                                                    uint32_t *__restrict__ packed, size_t packed_stride) {
    const LzFile &f = files[blockIdx.y];
    if ((size_t)blockIdx.x * chain::T >= f.en) return;
    match_chain_body(f.enc, (size_t)f.en, f.W, packed + (size_t)blockIdx.y * packed_stride, 0);
}

// cudaFuncSetAttribute is per device: remember which devices have seen it
static int chain_attr(const void *fn) {
    static std::atomic<uint64_t> done[2] = {{0}, {0}};  // bit = device; [0] k_match_chain, [1] kb_match_chain
    const int which = fn == (const void *)k_match_chain ? 0 : 1;
    int dev = 0;
    RSN_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (done[which].load(std::memory_order_acquire) & bit) return RSN_OK;
    RSN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(chain::Smem)));
    done[which].fetch_or(bit, std::memory_order_release);
    return RSN_OK;
}

static_assert(chain::T == 8192, "the tile size is shared with lzss_match_tile.cu (lzss_match_tile_size)");

// Tiles [tile_lo, tile_hi) only (positions tile*T ...); the caller guarantees that the bytes up to
// min(n, tile_hi*T + W) are in place.
int lzss_match_chain_range(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, size_t tile_lo,
                           size_t tile_hi, cudaStream_t s) {
    RSN_TRY(chain_attr((const void *)k_match_chain));
    const size_t smem = sizeof(chain::Smem);
    if (tile_hi > tile_lo)
        RSN_LAUNCH(k_match_chain, (unsigned)(tile_hi - tile_lo), chain::THREADS, smem, s, d_enc, n, W, d_packed, tile_lo);
    return RSN_OK;
}

// every file of a batch: files[f].enc / en / W are device-resident (en <= ecap, W <= window <= 4096)
int lzss_match_chain_batch(const LzFile *d_files, size_t G, size_t ecap, uint32_t *d_packed, size_t packed_stride,
                           cudaStream_t s) {
    RSN_TRY(chain_attr((const void *)kb_match_chain));
    const size_t smem = sizeof(chain::Smem);
    const dim3 grid((unsigned)div_up(ecap, chain::T), (unsigned)G);
    RSN_LAUNCH(kb_match_chain, grid, chain::THREADS, smem, s, d_files, d_packed, packed_stride);
    return RSN_OK;
}

}  // namespace rsn
