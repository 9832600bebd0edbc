// lzss_decode.cu — LZSS decompress (lz.Decompress, lzss.go:323-364, then DecodeOpeningSymbols,
// lzss.go:391-406).
//
// The reference is a byte-at-a-time state machine {Open, Sep, Close} (lookingFor, lzss.go:331).
// Each input byte is a transition function on those three states, so the state in front of
// every byte is an exclusive scan under function composition — exact for ANY input, not only
// for streams our compressor produced.
//   K5a  token-state scan (tile maps -> spine -> per-byte state); maps composed by table look-up
//   K5b  per-token (ptr,cnt) parse with strconv.Atoi semantics, output sizes, offsets
//   K5c  scatter literals and per-byte source distances
//   K6   back-reference resolve by bounded pointer chasing with path compression
//   K7   un-escape: a 2-state transducer, same scan pattern, compacting
#include "batch.cuh"
#include "common.cuh"
#include "lzss.cuh"

#include <algorithm>
#include <vector>

namespace rsn {

// ============================================================================= generic block scan

// Inclusive scan of arbitrary T under associative (non-commutative) op, in thread order.
// buf must hold blockDim.x elements.
template <typename T, typename Op>
__device__ __forceinline__ T block_inclusive_scan_generic(T v, T *buf, Op op) {
    buf[threadIdx.x] = v;
    __syncthreads();
    for (unsigned d = 1; d < blockDim.x; d <<= 1) {
        T o = v;
        if (threadIdx.x >= d) o = op(buf[threadIdx.x - d], v);
        __syncthreads();
        v = o;
        buf[threadIdx.x] = v;
        __syncthreads();
    }
    return v;
}

// ============================================================================= K5a token states

enum : uint8_t { ST_OPEN = 0, ST_SEP = 1, ST_CLOSE = 2 };
constexpr uint8_t kMapId = 0x24;  // f(0)=0, f(1)=1, f(2)=2, two bits each

__host__ __device__ constexpr uint8_t tok_map_of(uint8_t b) {
    // lzss.go:333-360: '<' acts only in Open, ',' only in Sep, '>' only in Close
    return b == 0x3C ? 0x25 : b == 0x2C ? 0x28 : b == 0x3E ? 0x04 : kMapId;
}
__host__ __device__ constexpr uint8_t map_apply(uint8_t m, uint8_t st) { return (m >> (2 * st)) & 3; }
// first a, then b
__host__ __device__ constexpr uint8_t map_compose(uint8_t a, uint8_t b) {
    return (uint8_t)(map_apply(b, map_apply(a, 0)) | (map_apply(b, map_apply(a, 1)) << 2) |
                     (map_apply(b, map_apply(a, 2)) << 4));
}
// Composition costs ~20 integer operations and runs once per input byte: the kernels look it up
// instead.  lut[a * 64 + b] = map_compose(a, b) for all 6-bit maps, built at compile time, copied
// to shared memory by each CTA.
struct alignas(16) MapLut {
    uint8_t v[64 * 64];
};
constexpr MapLut make_map_lut() {
    MapLut l{};
    for (int a = 0; a < 64; a++)
        for (int b = 0; b < 64; b++) l.v[a * 64 + b] = map_compose((uint8_t)a, (uint8_t)b);
    return l;
}
__device__ const MapLut g_map_lut = make_map_lut();

__device__ __forceinline__ void load_map_lut(uint8_t *lut) {  // 4096 bytes, blockDim.x >= 256
    const uint4 *src = reinterpret_cast<const uint4 *>(g_map_lut.v);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<uint4 *>(lut)[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ uint8_t thread_map(const uint8_t (&v)[16], int valid, const uint8_t *lut) {
    uint32_t m = kMapId;
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < valid) m = lut[m * 64 + tok_map_of(v[k])];
    return (uint8_t)m;
}

// Exclusive scan of the threads' maps in thread order (blockDim.x a multiple of 32, <= 1024):
// returns the composition of all earlier threads' maps; *total (if not null) receives the
// composition of the whole CTA.  wt: one byte per warp + 1.
__device__ __forceinline__ uint8_t block_exclusive_map(uint8_t m, const uint8_t *lut, uint8_t *wt, uint8_t *total) {
    const unsigned lane = lane_id(), wid = warp_id(), nw = (blockDim.x + 31) >> 5;
    uint32_t v = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v = lut[o * 64 + v];
    }
    if (lane == 31) wt[wid] = (uint8_t)v;
    __syncthreads();
    uint32_t pre = kMapId;
    for (unsigned k = 0; k < wid; k++) pre = lut[pre * 64 + wt[k]];
    uint32_t exc = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) exc = kMapId;
    exc = lut[pre * 64 + exc];
    if (total && threadIdx.x == blockDim.x - 1) *total = lut[pre * 64 + v];
    (void)nw;
    return (uint8_t)exc;
}

__device__ __forceinline__ void tok_reduce_body(const uint8_t *__restrict__ in, size_t n,
                                                uint8_t *__restrict__ tile_map) {
    __shared__ __align__(16) uint8_t lut[64 * 64];
    __shared__ uint8_t wt[33];
    load_map_lut(lut);
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t m = kMapId;
    if (base < n) {
        uint8_t v[16];
        load16(in, base, n, 0, v);
        m = thread_map(v, (int)min((size_t)16, n - base), lut);
    }
    const uint8_t exc = block_exclusive_map(m, lut, wt, nullptr);
    if (threadIdx.x == blockDim.x - 1) tile_map[blockIdx.x] = lut[exc * 64 + m];
}
__global__ void __launch_bounds__(kTileThreads) k_tok_reduce(const uint8_t *__restrict__ in, size_t n,
                                                             uint8_t *__restrict__ tile_map) {
    tok_reduce_body(in, n, tile_map);
}

// state in front of each tile, starting from Open
__device__ __forceinline__ void tok_spine_body(const uint8_t *__restrict__ tile_map, size_t tiles,
                                               uint8_t *__restrict__ tile_state) {
    __shared__ __align__(16) uint8_t lut[64 * 64];
    __shared__ uint8_t wt[34];
    load_map_lut(lut);
    uint8_t carry = ST_OPEN;
    // 16 consecutive tiles per thread: one scan round covers 16 * blockDim.x tiles
    for (size_t base = 0; base < tiles; base += (size_t)blockDim.x * 16) {
        const size_t t0 = base + (size_t)threadIdx.x * 16;
        uint8_t mt[16];
        uint32_t m = kMapId;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            mt[k] = t0 + k < tiles ? tile_map[t0 + k] : kMapId;
            m = lut[m * 64 + mt[k]];
        }
        uint8_t *total = &wt[33];
        uint32_t pre = block_exclusive_map((uint8_t)m, lut, wt, total);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (t0 + k < tiles) tile_state[t0 + k] = map_apply((uint8_t)pre, carry);
            pre = lut[pre * 64 + mt[k]];
        }
        __syncthreads();
        carry = map_apply(*total, carry);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_tok_spine(const uint8_t *__restrict__ tile_map, size_t tiles,
                                                    uint8_t *__restrict__ tile_state) {
    tok_spine_body(tile_map, tiles, tile_state);
}

// ============================================================================= K5b/K5c tokens

// strconv.Atoi (Go 1.15) fed one byte at a time, error dropped as lzss.go:338,346 do: a syntax
// error yields 0, a range error the clamped int64 extreme; whichever comes first wins because
// Go's scan stops at the first offending byte.
struct GoAtoi {
    uint64_t un = 0;
    bool first = true, neg = false, any = false, bad = false, range = false;
    __device__ __forceinline__ void feed(uint8_t c) {
        if (first) {
            first = false;
            if (c == '+' || c == '-') {
                neg = c == '-';
                return;
            }
        }
        if (bad || range) return;
        if (c < '0' || c > '9') {
            bad = true;
            return;
        }
        any = true;
        const uint64_t maxv = ~0ull, cutoff = maxv / 10 + 1;
        if (un >= cutoff) {
            range = true;
            return;
        }
        un *= 10;
        const uint64_t n1 = un + (uint64_t)(c - '0');
        if (n1 < un) {
            range = true;
            return;
        }
        un = n1;
    }
    __device__ __forceinline__ int64_t value() const {
        if (bad || !any) return 0;
        const uint64_t icut = 1ull << 63;
        const uint64_t u = range ? ~0ull : un;
        if (!neg && u >= icut) return INT64_MAX;
        if (neg && u > icut) return INT64_MIN;
        return neg ? -(int64_t)u : (int64_t)u;
    }
};

// The token that opens at in[i] ('<' seen in state Open): scan forward to the first ',' then the
// first '>' exactly as the state machine of lzss.go:333-360 does.  Returns false if the input ends
// first (the reference silently drops an unterminated token).
__device__ __forceinline__ bool parse_token_fwd(const uint8_t *__restrict__ in, size_t n, size_t i, int64_t &ptr,
                                                int64_t &cnt) {
    GoAtoi a, b;
    size_t j = i + 1;
    for (;; j++) {
        if (j >= n) return false;
        const uint8_t c = __ldg(in + j);
        if (c == 0x2C) break;
        a.feed(c);
    }
    for (j++;; j++) {
        if (j >= n) return false;
        const uint8_t c = __ldg(in + j);
        if (c == 0x3E) break;
        b.feed(c);
    }
    ptr = a.value();
    cnt = b.value();
    return true;
}

// ERR_HUGE: a token whose pointer is 2^32 or more.  Such a pointer is only valid once 4 GiB have
// been decoded, which is beyond this build's limit anyway, so its count never enters the size sums
// (Atoi clamps to 2^63-1: two such counts would wrap a u64 sum and defeat every later bound check).
enum : uint32_t { ERR_BAD_REF = 1u, FLAG_NEEDS_UNESCAPE = 2u, ERR_HUGE = 4u };
constexpr uint64_t kTileOutCap = 1ull << 33;  // per-tile size sums saturate here: >= 2^32 is unsupported
constexpr int kMaxTok = kTile / 3 + 2;  // "<,>" is the shortest token

// One tile of the compressed stream.  Threads first classify their 16 bytes (state machine from
// the tile's incoming state), the token openers of the whole tile are compacted into a list, and
// the list is parsed with one token per thread (so token parsing runs on full warps instead of
// one lane at a time).  WRITE == false: output bytes of the tile.  WRITE == true: literals to
// sb, pointer distances of referenced bytes to dist (zero-filled beforehand).
template <bool WRITE>
__device__ __forceinline__ void tok_tile_body(const uint8_t *__restrict__ in, size_t n,
                                              const uint8_t *__restrict__ tile_state,
                                              const uint64_t *__restrict__ tile_off, uint64_t *__restrict__ tile_out,
                                              uint8_t *__restrict__ sb, uint32_t *__restrict__ dist,
                                              uint32_t *__restrict__ err) {
    __shared__ __align__(16) uint8_t lut[64 * 64];
    __shared__ uint8_t wt[33];
    __shared__ uint64_t sm64[33];
    __shared__ uint32_t sm32[33];
    __shared__ uint32_t n_long;
    __shared__ uint16_t long_list[WRITE ? kMaxTok : 1];
    __shared__ uint16_t list[kMaxTok];
    __shared__ uint64_t tcnt[kMaxTok];
    __shared__ uint64_t tptr[WRITE ? kMaxTok : 1];
    __shared__ uint64_t toff[WRITE ? kMaxTok : 1];
    const size_t tile_base = (size_t)blockIdx.x * kTile;
    const size_t base = tile_base + (size_t)threadIdx.x * kItems;
    uint8_t v[16];
    int valid = 0;
    uint8_t m = kMapId;
    if (WRITE && threadIdx.x == 0) n_long = 0;
    load_map_lut(lut);
    if (base < n) {
        load16(in, base, n, 0, v);
        valid = (int)min((size_t)16, n - base);
        m = thread_map(v, valid, lut);
    }
    const uint8_t exc = block_exclusive_map(m, lut, wt, nullptr);
    uint8_t st = map_apply(exc, tile_state[blockIdx.x]);
    uint32_t lit_mask = 0, open_mask = 0, flags = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < valid) {
            const uint8_t b = v[k];
            if (st == ST_OPEN) {
                if (b == 0x3C) open_mask |= 1u << k;
                else {
                    lit_mask |= 1u << k;
                    if (b == 0x5C || b == 0xFF) flags |= FLAG_NEEDS_UNESCAPE;
                }
            }
            st = map_apply(tok_map_of(b), st);
        }
    }
    // compact the token openers of the tile
    const uint32_t ntok = __popc(open_mask);
    uint32_t T;
    const uint32_t tbase = block_exclusive_sum<uint32_t>(ntok, sm32, T);
    {
        uint32_t j = tbase;
        for (uint32_t mk = open_mask; mk; mk &= mk - 1) list[j++] = (uint16_t)(threadIdx.x * kItems + (__ffs(mk) - 1));
    }
    __syncthreads();
    // one token per thread
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
        int64_t ptr = 0, cnt = 0;
        if (parse_token_fwd(in, n, tile_base + list[t], ptr, cnt)) {
            if (cnt < 0 || ptr < cnt) {  // a = len - ptr: need 0 <= a <= a + cnt <= len (lzss.go:349-350)
                flags |= ERR_BAD_REF;
                cnt = 0;
            } else if (ptr >= (int64_t)1 << 32) {
                flags |= ERR_HUGE;
                cnt = 0;
            }
        } else {
            cnt = 0;  // unterminated at end of input: dropped
        }
        tcnt[t] = (uint64_t)cnt;
        if (WRITE) tptr[t] = (uint64_t)ptr;
    }
    __syncthreads();
    uint64_t c = __popc(lit_mask);
    for (uint32_t j = 0; j < ntok; j++) c += tcnt[tbase + j];
    uint64_t total;
    const uint64_t pre = block_exclusive_sum<uint64_t>(c, sm64, total);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_out[blockIdx.x] = min(total, kTileOutCap);
        if (flags) atomicOr(err, flags);
        return;
    }
    // literals in place, token output offsets for the fill phase
    uint64_t o = tile_off[blockIdx.x] + pre;
    {
        uint32_t j = tbase;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (lit_mask & (1u << k)) sb[o++] = v[k];
            else if (open_mask & (1u << k)) {
                toff[j] = o;
                o += tcnt[j];
                j++;
            }
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
        const uint64_t cnt = tcnt[t], ptr = tptr[t], off = toff[t];
        if (ptr > off) {  // absolutePointer < 0 (lzss.go:349): a panic even for an empty slice
            atomicOr(err, ERR_BAD_REF);
            continue;
        }
        if (cnt == 0) continue;
        if (cnt > 256) {  // long runs are filled by the whole CTA below
            long_list[atomicAdd(&n_long, 1u)] = (uint16_t)t;
            continue;
        }
        for (uint64_t q = 0; q < cnt; q++) dist[off + q] = (uint32_t)ptr;
    }
    __syncthreads();
    const uint32_t nl = n_long;
    for (uint32_t i = 0; i < nl; i++) {
        const uint32_t t = long_list[i];
        const uint64_t cnt = tcnt[t], ptr = tptr[t], off = toff[t];
        for (uint64_t q = threadIdx.x; q < cnt; q += blockDim.x) dist[off + q] = (uint32_t)ptr;
    }
}
template <bool WRITE>
__global__ void __launch_bounds__(kTileThreads) k_tok_tile(const uint8_t *__restrict__ in, size_t n,
                                                           const uint8_t *__restrict__ tile_state,
                                                           const uint64_t *__restrict__ tile_off,
                                                           uint64_t *__restrict__ tile_out, uint8_t *__restrict__ sb,
                                                           uint32_t *__restrict__ dist, uint32_t *__restrict__ err) {
    tok_tile_body<WRITE>(in, n, tile_state, tile_off, tile_out, sb, dist, err);
}

// ============================================================================= K6 resolve

constexpr int kHops = 64;

// Every referenced byte follows its source chain to a literal.  Chains longer than kHops are
// shortened in place (dist[o] := distance to the furthest ancestor reached) and finished by a
// later round; both the old and the new distance name a true ancestor, so concurrent readers
// are safe.  Literal bytes (dist == 0) are never written here.  Bytes that could not be finished are
// appended to `work_out`; later rounds walk that list only.
__device__ __forceinline__ void resolve_one(uint8_t *__restrict__ sb, uint32_t *__restrict__ dist, size_t o,
                                            uint32_t d, uint32_t *__restrict__ work_out,
                                            uint32_t *__restrict__ work_count) {
    size_t p = o - d;
    int hops = 0;
    uint32_t dp;
    // plain (L1-cacheable) loads: a stale distance still names a true ancestor
    while ((dp = dist[p]) != 0 && hops < kHops) {
        p -= dp;
        hops++;
    }
    if (dp == 0) {
        sb[o] = sb[p];
    } else {
        work_out[atomicAdd(work_count, 1u)] = (uint32_t)o;  // rare: chains deeper than kHops
    }
    if (hops) dist[o] = (uint32_t)(o - p);
}

// First round.  A CTA walks a contiguous chunk of the output front to back, each warp 128 bytes per
// step (four per lane, one 16-byte load of their distances): sources lie a window back, i.e. in
// lines this same SM loaded a few steps earlier, so the chase mostly hits L1 instead of going to L2.
// Two thirds of the bytes of a text are literals with nothing to chase, so the referenced ones are
// first compacted into a per-warp queue (warp scan, no CTA barrier) and the chase then runs on
// full warps: more independent chains in flight per warp (9 of 32 lanes were active without it).
constexpr int kResolveChunk = 32768;

__global__ void __launch_bounds__(256) k_resolve4(uint8_t *__restrict__ sb, uint32_t *__restrict__ dist, size_t n,
                                                  uint32_t *__restrict__ work_out, uint32_t *__restrict__ work_count) {
    __shared__ uint32_t q_dist[8][128];
    __shared__ uint8_t q_pos[8][128];
    const unsigned lane = lane_id(), w = warp_id();
    const size_t c_lo = (size_t)blockIdx.x * kResolveChunk;
    const size_t c_hi = min(n, c_lo + kResolveChunk);
    for (size_t s0 = c_lo + (size_t)w * 128; s0 < c_hi; s0 += 8 * 128) {
        const size_t o0 = s0 + (size_t)lane * 4;
        uint32_t d[4];
        if (o0 + 4 <= c_hi) {
            const uint4 q = *reinterpret_cast<const uint4 *>(dist + o0);
            d[0] = q.x;
            d[1] = q.y;
            d[2] = q.z;
            d[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) d[k] = o0 + k < c_hi ? dist[o0 + k] : 0u;
        }
        const uint32_t cnt = (d[0] != 0) + (d[1] != 0) + (d[2] != 0) + (d[3] != 0);
        uint32_t inc = cnt;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, sh);
            if (lane >= (unsigned)sh) inc += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t at = inc - cnt;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (d[k]) {
                q_dist[w][at] = d[k];
                q_pos[w][at] = (uint8_t)(lane * 4 + k);
                at++;
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < total; i += 32) resolve_one(sb, dist, s0 + q_pos[w][i], q_dist[w][i], work_out, work_count);
        __syncwarp();
    }
}

// later rounds: the bytes of the previous round's work list
__global__ void __launch_bounds__(256) k_resolve(uint8_t *__restrict__ sb, uint32_t *__restrict__ dist, size_t n,
                                                 const uint32_t *__restrict__ work, uint32_t *__restrict__ work_out,
                                                 uint32_t *__restrict__ work_count) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t o = work[t];
    const uint32_t d = dist[o];
    if (d) resolve_one(sb, dist, o, d, work_out, work_count);
}

// ============================================================================= K7 un-escape

// State = "previous byte was an unescaped 5C".  Aggregate: next state and emitted byte count for
// both incoming states.
struct UnescAgg {
    uint32_t nxt;      // bit0: f(0), bit1: f(1)
    uint64_t cnt[2];
};
__device__ __forceinline__ UnescAgg unesc_identity() { return UnescAgg{2u, {0, 0}}; }
__device__ __forceinline__ UnescAgg unesc_compose(const UnescAgg &a, const UnescAgg &b) {
    UnescAgg r;
    const uint32_t a0 = a.nxt & 1, a1 = (a.nxt >> 1) & 1;
    r.nxt = ((b.nxt >> a0) & 1) | (((b.nxt >> a1) & 1) << 1);
    r.cnt[0] = a.cnt[0] + b.cnt[a0];
    r.cnt[1] = a.cnt[1] + b.cnt[a1];
    return r;
}
struct UnescCompose {
    __device__ UnescAgg operator()(const UnescAgg &a, const UnescAgg &b) const { return unesc_compose(a, b); }
};

__device__ __forceinline__ UnescAgg unesc_thread(const uint8_t (&v)[16], int valid) {
    // run the 16 bytes from both incoming states (lzss.go:394-404)
    UnescAgg r;
    uint32_t e0 = 0, e1 = 1;
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < valid) {
            const bool bs = v[k] == 0x5C;
            c0 += (bs && !e0) ? 0 : 1;
            c1 += (bs && !e1) ? 0 : 1;
            e0 = (bs && !e0) ? 1 : 0;
            e1 = (bs && !e1) ? 1 : 0;
        }
    }
    r.nxt = e0 | (e1 << 1);
    r.cnt[0] = c0;
    r.cnt[1] = c1;
    return r;
}

__device__ __forceinline__ void unesc_reduce_body(const uint8_t *__restrict__ in, size_t n,
                                                  UnescAgg *__restrict__ tile_agg) {
    __shared__ UnescAgg buf[kTileThreads];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    UnescAgg a = unesc_identity();
    if (base < n) {
        uint8_t v[16];
        load16(in, base, n, 0, v);
        a = unesc_thread(v, (int)min((size_t)16, n - base));
    }
    a = block_inclusive_scan_generic<UnescAgg>(a, buf, UnescCompose());
    if (threadIdx.x == blockDim.x - 1) tile_agg[blockIdx.x] = a;
}
__global__ void __launch_bounds__(kTileThreads) k_unesc_reduce(const uint8_t *__restrict__ in, size_t n,
                                                               UnescAgg *__restrict__ tile_agg) {
    unesc_reduce_body(in, n, tile_agg);
}

// per tile: incoming escape state and output offset; total output size
__device__ __forceinline__ void unesc_spine_body(const UnescAgg *__restrict__ tile_agg, size_t tiles,
                                                 uint8_t *__restrict__ tile_state, uint64_t *__restrict__ tile_off,
                                                 uint64_t *__restrict__ total) {
    __shared__ UnescAgg buf[1024];
    uint32_t carry_state = 0;
    uint64_t carry_off = 0;
    for (size_t base = 0; base < tiles; base += blockDim.x) {
        const size_t t = base + threadIdx.x;
        UnescAgg a = t < tiles ? tile_agg[t] : unesc_identity();
        block_inclusive_scan_generic<UnescAgg>(a, buf, UnescCompose());
        const UnescAgg exc = threadIdx.x ? buf[threadIdx.x - 1] : unesc_identity();
        if (t < tiles) {
            tile_state[t] = (uint8_t)((exc.nxt >> carry_state) & 1);
            tile_off[t] = carry_off + exc.cnt[carry_state];
        }
        const UnescAgg last = buf[blockDim.x - 1];
        __syncthreads();
        carry_off += last.cnt[carry_state];
        carry_state = (last.nxt >> carry_state) & 1;
    }
    if (threadIdx.x == 0) *total = carry_off;
}
__global__ void __launch_bounds__(1024) k_unesc_spine(const UnescAgg *__restrict__ tile_agg, size_t tiles,
                                                      uint8_t *__restrict__ tile_state,
                                                      uint64_t *__restrict__ tile_off, uint64_t *__restrict__ total) {
    unesc_spine_body(tile_agg, tiles, tile_state, tile_off, total);
}

__device__ __forceinline__ void unesc_apply_body(const uint8_t *__restrict__ in, size_t n,
                                                 const uint8_t *__restrict__ tile_state,
                                                 const uint64_t *__restrict__ tile_off, uint8_t *__restrict__ out) {
    __shared__ UnescAgg buf[kTileThreads];
    __shared__ uint8_t stage[kTile];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t v[16];
    int valid = 0;
    UnescAgg a = unesc_identity();
    if (base < n) {
        load16(in, base, n, 0, v);
        valid = (int)min((size_t)16, n - base);
        a = unesc_thread(v, valid);
    }
    block_inclusive_scan_generic<UnescAgg>(a, buf, UnescCompose());
    const UnescAgg exc = threadIdx.x ? buf[threadIdx.x - 1] : unesc_identity();
    const uint32_t s_in = tile_state[blockIdx.x];
    uint32_t esc = (exc.nxt >> s_in) & 1;
    uint32_t pos = (uint32_t)exc.cnt[s_in];
    const uint32_t tile_total = (uint32_t)buf[blockDim.x - 1].cnt[s_in];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < valid) {
            const uint8_t b = v[k];
            if (b == 0xFF && !esc) {
                stage[pos++] = 0x3C;
            } else if (b == 0x5C && !esc) {
                esc = 1;
            } else {
                esc = 0;
                stage[pos++] = b;
            }
        }
    }
    __syncthreads();
    uint8_t *dst = out + tile_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < tile_total; i += blockDim.x) dst[i] = stage[i];
}
__global__ void __launch_bounds__(kTileThreads) k_unesc_apply(const uint8_t *__restrict__ in, size_t n,
                                                              const uint8_t *__restrict__ tile_state,
                                                              const uint64_t *__restrict__ tile_off,
                                                              uint8_t *__restrict__ out) {
    unesc_apply_body(in, n, tile_state, tile_off, out);
}

static int lzss_unescape(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    DevBuf out;
    if (n == 0) {
        RSN_TRY(out.alloc_out(16, s));
        *d_out = (uint8_t *)out.release();
        *out_n = 0;
        return RSN_OK;
    }
    const size_t tiles = div_up(n, kTile);
    DevBuf agg, tstate, toff;
    RSN_TRY(agg.alloc(tiles * sizeof(UnescAgg), s));
    RSN_TRY(tstate.alloc(tiles, s));
    RSN_TRY(toff.alloc((tiles + 1) * 8, s));
    RSN_LAUNCH(k_unesc_reduce, (unsigned)tiles, kTileThreads, 0, s, d_in, n, agg.as<UnescAgg>());
    RSN_LAUNCH(k_unesc_spine, 1, 1024, 0, s, agg.as<UnescAgg>(), tiles, tstate.as<uint8_t>(), toff.as<uint64_t>(),
               toff.as<uint64_t>() + tiles);
    uint64_t total = 0;
    RSN_TRY(read_u64(toff.as<uint64_t>() + tiles, &total, s));
    RSN_TRY(out.alloc_out(total + 16, s));
    RSN_LAUNCH(k_unesc_apply, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tstate.as<uint8_t>(), toff.as<uint64_t>(),
               out.as<uint8_t>());
    *d_out = (uint8_t *)out.release();
    *out_n = (size_t)total;
    return RSN_OK;
}

// ============================================================================= host orchestration

int lzss_decompress_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    ArenaScope scope(s);
    if (n == 0) return lzss_unescape(d_in, 0, d_out, out_n, s);
    const size_t tiles = div_up(n, kTile);
    Trace tr("lzd", s);
    DevBuf tmap, tstate, tout, toff, err;
    RSN_TRY(tmap.alloc(tiles, s));
    RSN_TRY(tstate.alloc(tiles, s));
    RSN_TRY(tout.alloc(tiles * 8, s));
    RSN_TRY(toff.alloc((tiles + 1) * 8, s));
    RSN_TRY(err.alloc(16, s));
    RSN_CUDA(cudaMemsetAsync(err.p, 0, 16, s));
    RSN_LAUNCH(k_tok_reduce, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tmap.as<uint8_t>());
    RSN_LAUNCH(k_tok_spine, 1, 1024, 0, s, tmap.as<uint8_t>(), tiles, tstate.as<uint8_t>());
    RSN_LAUNCH(k_tok_tile<false>, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tstate.as<uint8_t>(),
               (const uint64_t *)nullptr, tout.as<uint64_t>(), (uint8_t *)nullptr, (uint32_t *)nullptr,
               err.as<uint32_t>());
    RSN_TRY(spine_scan_u64(tout.as<uint64_t>(), toff.as<uint64_t>(), toff.as<uint64_t>() + tiles, tiles, s));
    Ctx &c = ctx();
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars, toff.as<uint64_t>() + tiles, 8, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars + 1, err.p, 8, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    const uint64_t sbn = c.h_scalars[0];
    const uint32_t flags = (uint32_t)c.h_scalars[1];
    if (flags & ERR_BAD_REF) return RSN_ERR_BAD_REFERENCE;
    if (sbn >= (1ull << 32)) return RSN_ERR_UNSUPPORTED;  // u32 source distances (documented limit)
    // a pointer of 2^32 or more in a stream that decodes to less: before the start of the output
    if (flags & ERR_HUGE) return RSN_ERR_BAD_REFERENCE;
    tr.mark("plan");
    const bool needs_unescape = (flags & FLAG_NEEDS_UNESCAPE) != 0;
    DevBuf sb, dist;  // sb becomes the result itself when no literal needs un-escaping
    RSN_TRY(sb.alloc_out(sbn + 16, s));
    if (sbn) {
        RSN_TRY(dist.alloc(sbn * 4 + 16, s));
        RSN_CUDA(cudaMemsetAsync(dist.p, 0, sbn * 4 + 16, s));
        RSN_LAUNCH(k_tok_tile<true>, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tstate.as<uint8_t>(),
                   toff.as<uint64_t>(), (uint64_t *)nullptr, sb.as<uint8_t>(), dist.as<uint32_t>(), err.as<uint32_t>());
        tr.mark("scatter");
        // round 0 visits every byte; bytes whose chain is deeper than kHops go to a work list (their
        // distance already shortened), and later rounds only walk that list
        DevBuf wl[2];
        RSN_TRY(wl[0].alloc(sbn * 4 + 16, s));
        RSN_TRY(wl[1].alloc(sbn * 4 + 16, s));
        uint32_t *count = err.as<uint32_t>() + 1;
        size_t todo = (size_t)sbn;
        for (int round = 0; todo; round++) {
            RSN_CUDA(cudaMemsetAsync(count, 0, 4, s));
            if (round == 0)
                RSN_LAUNCH(k_resolve4, (unsigned)div_up(todo, kResolveChunk), 256, 0, s, sb.as<uint8_t>(),
                           dist.as<uint32_t>(), todo, wl[0].as<uint32_t>(), count);
            else
                RSN_LAUNCH(k_resolve, (unsigned)div_up(todo, 256), 256, 0, s, sb.as<uint8_t>(), dist.as<uint32_t>(), todo,
                           wl[(round + 1) & 1].as<uint32_t>(), wl[round & 1].as<uint32_t>(), count);
            RSN_CUDA(cudaMemcpyAsync(c.h_scalars, err.p, 8, cudaMemcpyDeviceToHost, s));
            RSN_CUDA(stream_wait(s));
            const uint32_t e = (uint32_t)c.h_scalars[0];
            if (e & ERR_BAD_REF) return RSN_ERR_BAD_REFERENCE;
            todo = (size_t)(uint32_t)(c.h_scalars[0] >> 32);
        }
        tr.mark("resolve");
    }
    if (!needs_unescape) {  // no 0x5C / 0xFF among the literals: the buffer is already the answer
        *d_out = (uint8_t *)sb.release();
        *out_n = (size_t)sbn;
        return RSN_OK;
    }
    const int rc = lzss_unescape(sb.as<uint8_t>(), (size_t)sbn, d_out, out_n, s);
    tr.mark("unescape");
    return rc;
}

// ============================================================================= batches of small files
//
// One launch per kernel for a whole group of files (batch.cuh).  The decoded files of a group sit
// back to back in ONE buffer, with one distance array beside it: a validated pointer never reaches
// below the start of its own file, so the resolve kernels run over the whole buffer as if it were
// a single stream, and only the token kernels know about files.

struct DecFile {
    const uint8_t *in;
    uint64_t n;        // 0: not part of this phase
    uint64_t gbase;    // offset of the file's bytes in the group buffers (phase B on)
    uint64_t sbn;      // decoded size before un-escaping
    uint64_t un_n;     // size after un-escaping
    uint32_t flags;    // ERR_BAD_REF | FLAG_NEEDS_UNESCAPE
    uint32_t unesc;    // phase C runs on this file
};

struct DecBatch {
    DecFile *files;
    uint8_t *tmap, *tstate;       // [G][tc_stride]
    uint64_t *tout, *toff;        // [G][tc_stride]
    size_t tc_stride;
    uint8_t *sb, *out2;           // group buffers
    uint32_t *dist;
    UnescAgg *agg;                // [G][uc_stride]
    uint8_t *ustate;
    uint64_t *uoff;
    size_t uc_stride;
};

__global__ void __launch_bounds__(kTileThreads) kb_tok_reduce(DecBatch b) {
    const DecFile &f = b.files[blockIdx.y];
    if ((size_t)blockIdx.x * kTile >= f.n) return;
    tok_reduce_body(f.in, (size_t)f.n, b.tmap + (size_t)blockIdx.y * b.tc_stride);
}
__global__ void __launch_bounds__(1024) kb_tok_spine(DecBatch b) {
    const DecFile &f = b.files[blockIdx.x];
    if (f.n == 0) return;
    tok_spine_body(b.tmap + (size_t)blockIdx.x * b.tc_stride, div_up_dev((size_t)f.n, (size_t)kTile),
                   b.tstate + (size_t)blockIdx.x * b.tc_stride);
}
template <bool WRITE>
__global__ void __launch_bounds__(kTileThreads) kb_tok_tile(DecBatch b) {
    DecFile &f = b.files[blockIdx.y];
    if ((size_t)blockIdx.x * kTile >= f.n) return;
    const size_t o = (size_t)blockIdx.y * b.tc_stride;
    tok_tile_body<WRITE>(f.in, (size_t)f.n, b.tstate + o, b.toff + o, b.tout + o, WRITE ? b.sb + f.gbase : nullptr,
                         WRITE ? b.dist + f.gbase : nullptr, &f.flags);
}
__global__ void __launch_bounds__(256) kb_tok_finish(DecBatch b) {
    __shared__ uint64_t sm[33];
    DecFile &f = b.files[blockIdx.x];
    const size_t o = (size_t)blockIdx.x * b.tc_stride;
    const uint64_t total = cta_scan_u64(b.tout + o, b.toff + o, div_up_dev((size_t)f.n, (size_t)kTile), sm);
    if (threadIdx.x == 0) f.sbn = total;
}

__global__ void __launch_bounds__(kTileThreads) kb_unesc_reduce(DecBatch b) {
    const DecFile &f = b.files[blockIdx.y];
    if (!f.unesc || (size_t)blockIdx.x * kTile >= f.sbn) return;
    unesc_reduce_body(b.sb + f.gbase, (size_t)f.sbn, b.agg + (size_t)blockIdx.y * b.uc_stride);
}
__global__ void __launch_bounds__(1024) kb_unesc_spine(DecBatch b) {
    DecFile &f = b.files[blockIdx.x];
    if (!f.unesc) return;
    const size_t o = (size_t)blockIdx.x * b.uc_stride;
    unesc_spine_body(b.agg + o, div_up_dev((size_t)f.sbn, (size_t)kTile), b.ustate + o, b.uoff + o, &f.un_n);
}
__global__ void __launch_bounds__(kTileThreads) kb_unesc_apply(DecBatch b) {
    const DecFile &f = b.files[blockIdx.y];
    if (!f.unesc || (size_t)blockIdx.x * kTile >= f.sbn) return;
    const size_t o = (size_t)blockIdx.y * b.uc_stride;
    unesc_apply_body(b.sb + f.gbase, (size_t)f.sbn, b.ustate + o, b.uoff + o, b.out2 + f.gbase);
}

// lz.Decompress (lzss.go:323-364 + 391-406) over every file of the group.
int lzss_decompress_batch(const BatchIO &in, BatchIO &out, cudaStream_t s) {
    const size_t G = in.size();
    out.resize(G);
    out.rc = in.rc;
    if (G == 0) return RSN_OK;
    ArenaScope scope(s);
    size_t cap = 1;
    for (size_t f = 0; f < G; f++)
        if (in.rc[f] == RSN_OK) cap = std::max<size_t>(cap, in.n[f]);
    if (cap > kBatchMaxFile) return RSN_ERR_UNSUPPORTED;
    const size_t tiles_cap = div_up(cap, kTile);
    DecBatch b{};
    b.tc_stride = tiles_cap + 1;
    HostVec<DecFile> h(G);
    if (!h.data()) return RSN_ERR_NOMEM;
    for (size_t f = 0; f < G; f++) {
        h[f] = DecFile{};
        h[f].in = in.ptr[f];
        h[f].n = in.rc[f] == RSN_OK ? in.n[f] : 0;
    }
    DevBuf files, tmap, tstate, tout, toff;
    RSN_TRY(files.alloc(G * sizeof(DecFile), s));
    RSN_TRY(tmap.alloc(G * b.tc_stride, s));
    RSN_TRY(tstate.alloc(G * b.tc_stride, s));
    RSN_TRY(tout.alloc(G * b.tc_stride * 8, s));
    RSN_TRY(toff.alloc(G * b.tc_stride * 8, s));
    b.files = files.as<DecFile>();
    b.tmap = tmap.as<uint8_t>();
    b.tstate = tstate.as<uint8_t>();
    b.tout = tout.as<uint64_t>();
    b.toff = toff.as<uint64_t>();
    const unsigned g = (unsigned)G;
    const dim3 tgrid((unsigned)tiles_cap, g);
    // ---- phase A: token states and decoded sizes
    RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(DecFile), cudaMemcpyHostToDevice, s));
    RSN_LAUNCH(kb_tok_reduce, tgrid, kTileThreads, 0, s, b);
    RSN_LAUNCH(kb_tok_spine, g, 1024, 0, s, b);
    RSN_LAUNCH(kb_tok_tile<false>, tgrid, kTileThreads, 0, s, b);
    RSN_LAUNCH(kb_tok_finish, g, 256, 0, s, b);
    RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(DecFile), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    size_t total = 0, sb_cap = 1;
    for (size_t f = 0; f < G; f++) {
        if (h[f].n && (h[f].flags & ERR_BAD_REF)) {
            out.rc[f] = RSN_ERR_BAD_REFERENCE;
            h[f].n = 0;
        } else if (h[f].n && h[f].sbn >= (1ull << 32)) {  // same limit as the single-stream call
            out.rc[f] = RSN_ERR_UNSUPPORTED;
            h[f].n = 0;
        } else if (h[f].n && (h[f].flags & ERR_HUGE)) {
            out.rc[f] = RSN_ERR_BAD_REFERENCE;
            h[f].n = 0;
        }
        if (out.rc[f] != RSN_OK) {
            h[f].n = 0;
            h[f].sbn = 0;
        }
        h[f].gbase = total;
        total += (h[f].sbn + 16 + 255) & ~(uint64_t)255;
        sb_cap = std::max<size_t>(sb_cap, h[f].sbn);
    }
    // distances and work lists take 12 bytes per decoded byte: a group that decodes to more than
    // 256 MiB is cut in two (compressed sizes say little about decoded ones: 1365 log files of config 4
    // are 50 MB compressed and 341 MiB decoded); a single file that large goes through the per-file call
    if (total > ((size_t)256 << 20)) {
        if (G == 1) return RSN_ERR_UNSUPPORTED;
        const size_t half = G / 2;
        for (int part = 0; part < 2; part++) {
            const size_t lo = part ? half : 0, hi = part ? G : half;
            BatchIO sub_in, sub_out;
            sub_in.resize(hi - lo);
            for (size_t f = lo; f < hi; f++) {
                sub_in.ptr[f - lo] = in.ptr[f];
                sub_in.n[f - lo] = in.n[f];
                sub_in.rc[f - lo] = out.rc[f];
            }
            int rc = lzss_decompress_batch(sub_in, sub_out, s);
            if (rc == RSN_ERR_UNSUPPORTED) {  // one huge file left: per-file call
                sub_out.resize(hi - lo);
                sub_out.rc = sub_in.rc;
                for (size_t f = 0; f < hi - lo; f++) {
                    if (sub_in.rc[f] != RSN_OK) continue;
                    uint8_t *r = nullptr;
                    size_t rn = 0;
                    sub_out.rc[f] = lzss_decompress_dev(sub_in.ptr[f], (size_t)sub_in.n[f], &r, &rn, s);
                    if (sub_out.rc[f] != RSN_OK) continue;
                    sub_out.ptr[f] = r;
                    sub_out.n[f] = rn;
                    sub_out.owned.push_back(r);
                }
                rc = RSN_OK;
            }
            if (rc != RSN_OK) {
                sub_out.release(s);
                out.release(s);
                return rc;
            }
            for (size_t f = lo; f < hi; f++) {
                out.ptr[f] = sub_out.ptr[f - lo];
                out.n[f] = sub_out.n[f - lo];
                out.rc[f] = sub_out.rc[f - lo];
            }
            out.owned.insert(out.owned.end(), sub_out.owned.begin(), sub_out.owned.end());
            out.spans.insert(out.spans.end(), sub_out.spans.begin(), sub_out.spans.end());
        }
        return RSN_OK;
    }
    // ---- phase B: literals and distances, then the chase
    DevBuf sb, dist, wl[2], cnt;
    RSN_TRY(sb.alloc_out(total + 256, s));
    RSN_TRY(dist.alloc(total * 4 + 64, s));
    RSN_TRY(cnt.alloc(16, s));
    RSN_CUDA(cudaMemsetAsync(dist.p, 0, total * 4 + 64, s));
    b.sb = sb.as<uint8_t>();
    b.dist = dist.as<uint32_t>();
    RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(DecFile), cudaMemcpyHostToDevice, s));
    RSN_LAUNCH(kb_tok_tile<true>, tgrid, kTileThreads, 0, s, b);
    RSN_TRY(wl[0].alloc(total * 4 + 16, s));
    RSN_TRY(wl[1].alloc(total * 4 + 16, s));
    Ctx &c = ctx();
    size_t todo = total;
    for (int round = 0; todo; round++) {
        RSN_CUDA(cudaMemsetAsync(cnt.p, 0, 4, s));
        if (round == 0)
            RSN_LAUNCH(k_resolve4, (unsigned)div_up(todo, kResolveChunk), 256, 0, s, b.sb, b.dist, todo,
                       wl[0].as<uint32_t>(), cnt.as<uint32_t>());
        else
            RSN_LAUNCH(k_resolve, (unsigned)div_up(todo, 256), 256, 0, s, b.sb, b.dist, todo,
                       wl[(round + 1) & 1].as<uint32_t>(), wl[round & 1].as<uint32_t>(), cnt.as<uint32_t>());
        RSN_CUDA(cudaMemcpyAsync(c.h_scalars, cnt.p, 4, cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
        todo = (size_t)(uint32_t)c.h_scalars[0];
    }
    RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(DecFile), cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    // ---- phase C: un-escape the files whose literals hold 5C / FF
    bool any_unesc = false;
    for (size_t f = 0; f < G; f++) {
        if (h[f].n && (h[f].flags & ERR_BAD_REF)) {  // a pointer before the start of the output (lzss.go:349)
            out.rc[f] = RSN_ERR_BAD_REFERENCE;
            h[f].n = 0;
        }
        h[f].unesc = (h[f].n && (h[f].flags & FLAG_NEEDS_UNESCAPE)) ? 1u : 0u;
        h[f].un_n = h[f].sbn;
        any_unesc |= h[f].unesc != 0;
    }
    DevBuf out2;
    if (any_unesc) {
        const size_t ut = div_up(sb_cap, kTile);
        b.uc_stride = ut + 1;
        DevBuf agg, ustate, uoff;
        RSN_TRY(agg.alloc(G * b.uc_stride * sizeof(UnescAgg), s));
        RSN_TRY(ustate.alloc(G * b.uc_stride, s));
        RSN_TRY(uoff.alloc(G * b.uc_stride * 8, s));
        RSN_TRY(out2.alloc_out(total + 256, s));
        b.agg = agg.as<UnescAgg>();
        b.ustate = ustate.as<uint8_t>();
        b.uoff = uoff.as<uint64_t>();
        b.out2 = out2.as<uint8_t>();
        RSN_CUDA(cudaMemcpyAsync(files.p, h.data(), G * sizeof(DecFile), cudaMemcpyHostToDevice, s));
        const dim3 ugrid((unsigned)ut, g);
        RSN_LAUNCH(kb_unesc_reduce, ugrid, kTileThreads, 0, s, b);
        RSN_LAUNCH(kb_unesc_spine, g, 1024, 0, s, b);
        RSN_LAUNCH(kb_unesc_apply, ugrid, kTileThreads, 0, s, b);
        RSN_CUDA(cudaMemcpyAsync(h.data(), files.p, G * sizeof(DecFile), cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
    }
    for (size_t f = 0; f < G; f++) {
        if (out.rc[f] != RSN_OK) continue;
        out.ptr[f] = (h[f].unesc ? out2.as<uint8_t>() : sb.as<uint8_t>()) + h[f].gbase;
        out.n[f] = h[f].unesc ? h[f].un_n : h[f].sbn;
    }
    out.spans.push_back({sb.as<uint8_t>(), sb.bytes});
    out.owned.push_back(sb.release());
    if (out2.p) {
        out.spans.push_back({out2.as<uint8_t>(), out2.bytes});
        out.owned.push_back(out2.release());
    }
    return RSN_OK;
}

}  // namespace rsn
