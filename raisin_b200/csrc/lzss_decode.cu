// lzss_decode.cu — LZSS decompress (lz.Decompress, lzss.go:323-364, then DecodeOpeningSymbols,
// lzss.go:391-406).
//
// The reference is a byte-at-a-time state machine {Open, Sep, Close} (lookingFor, lzss.go:331).
// Each input byte is a transition function on those three states, so the state in front of
// every byte is an exclusive scan under function composition — exact for ANY input, not only
// for streams our compressor produced.
//   K5a  token-state scan (tile maps -> spine -> per-byte state)
//   K5b  per-token (ptr,cnt) parse with strconv.Atoi semantics, output sizes, offsets
//   K5c  scatter literals and per-byte source distances
//   K6   back-reference resolve by bounded pointer chasing with path compression
//   K7   un-escape: a 2-state transducer, same scan pattern, compacting
#include "common.cuh"
#include "lzss.cuh"

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

__device__ __forceinline__ uint8_t tok_map_of(uint8_t b) {
    // lzss.go:333-360: '<' acts only in Open, ',' only in Sep, '>' only in Close
    return b == 0x3C ? 0x25 : b == 0x2C ? 0x28 : b == 0x3E ? 0x04 : kMapId;
}
__device__ __forceinline__ uint8_t map_apply(uint8_t m, uint8_t st) { return (m >> (2 * st)) & 3; }
// first a, then b
__device__ __forceinline__ uint8_t map_compose(uint8_t a, uint8_t b) {
    return (uint8_t)(map_apply(b, map_apply(a, 0)) | (map_apply(b, map_apply(a, 1)) << 2) |
                     (map_apply(b, map_apply(a, 2)) << 4));
}
struct MapCompose {
    __device__ uint8_t operator()(uint8_t a, uint8_t b) const { return map_compose(a, b); }
};

__device__ __forceinline__ uint8_t thread_map(const uint8_t (&v)[16], int valid) {
    uint8_t m = kMapId;
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < valid) m = map_compose(m, tok_map_of(v[k]));
    return m;
}

__global__ void __launch_bounds__(kTileThreads) k_tok_reduce(const uint8_t *__restrict__ in, size_t n,
                                                             uint8_t *__restrict__ tile_map) {
    __shared__ uint8_t buf[kTileThreads];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t m = kMapId;
    if (base < n) {
        uint8_t v[16];
        load16(in, base, n, 0, v);
        m = thread_map(v, (int)min((size_t)16, n - base));
    }
    m = block_inclusive_scan_generic<uint8_t>(m, buf, MapCompose());
    if (threadIdx.x == blockDim.x - 1) tile_map[blockIdx.x] = m;
}

// state in front of each tile, starting from Open
__global__ void __launch_bounds__(1024) k_tok_spine(const uint8_t *__restrict__ tile_map, size_t tiles,
                                                    uint8_t *__restrict__ tile_state) {
    __shared__ uint8_t buf[1024];
    uint8_t carry = ST_OPEN;
    for (size_t base = 0; base < tiles; base += blockDim.x) {
        const size_t t = base + threadIdx.x;
        uint8_t m = t < tiles ? tile_map[t] : kMapId;
        uint8_t inc = block_inclusive_scan_generic<uint8_t>(m, buf, MapCompose());
        // exclusive prefix = inclusive of the previous thread
        uint8_t exc = threadIdx.x ? buf[threadIdx.x - 1] : kMapId;
        if (t < tiles) tile_state[t] = map_apply(exc, carry);
        const uint8_t last = buf[blockDim.x - 1];
        __syncthreads();
        carry = map_apply(last, carry);
        (void)inc;
    }
}

// st[i] = state in front of byte i
__global__ void __launch_bounds__(kTileThreads) k_tok_states(const uint8_t *__restrict__ in, size_t n,
                                                             const uint8_t *__restrict__ tile_state,
                                                             uint8_t *__restrict__ st) {
    __shared__ uint8_t buf[kTileThreads];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t v[16];
    int valid = 0;
    uint8_t m = kMapId;
    if (base < n) {
        load16(in, base, n, 0, v);
        valid = (int)min((size_t)16, n - base);
        m = thread_map(v, valid);
    }
    block_inclusive_scan_generic<uint8_t>(m, buf, MapCompose());
    const uint8_t exc = threadIdx.x ? buf[threadIdx.x - 1] : kMapId;
    uint8_t s = map_apply(exc, tile_state[blockIdx.x]);
    if (valid == 16 && ((reinterpret_cast<uintptr_t>(st + base) & 15) == 0)) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            w[k >> 2] |= (uint32_t)s << ((k & 3) * 8);
            s = map_apply(tok_map_of(v[k]), s);
        }
        *reinterpret_cast<uint4 *>(st + base) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int k = 0; k < valid; k++) {
            st[base + k] = s;
            s = map_apply(tok_map_of(v[k]), s);
        }
    }
}

// ============================================================================= K5b token parse

// strconv.Atoi (Go 1.15) with the error dropped, over in[a, b): syntax error => 0, range error
// => clamped int64 extreme (lzss.go:338, 346 discard the error).
__device__ int64_t go_atoi_dev(const uint8_t *__restrict__ in, size_t a, size_t b) {
    if (a >= b) return 0;
    bool neg = false;
    uint8_t c0 = in[a];
    if (c0 == '+' || c0 == '-') {
        neg = c0 == '-';
        a++;
        if (a >= b) return 0;
    }
    const uint64_t maxv = ~0ull, cutoff = maxv / 10 + 1;
    uint64_t un = 0;
    for (size_t i = a; i < b; i++) {
        const uint8_t c = in[i];
        if (c < '0' || c > '9') return 0;
        if (un >= cutoff) {
            un = maxv;
            break;
        }
        un *= 10;
        const uint64_t n1 = un + (uint64_t)(c - '0');
        if (n1 < un) {
            un = maxv;
            break;
        }
        un = n1;
    }
    const uint64_t icut = 1ull << 63;
    if (!neg && un >= icut) return INT64_MAX;
    if (neg && un > icut) return INT64_MIN;
    const int64_t v = (int64_t)un;
    return neg ? -v : v;
}

// For a closing '>' at i (state Close): locate the token's separator and opening by walking the
// state array backwards, and evaluate pointer and count as the reference does.
__device__ __forceinline__ void parse_token(const uint8_t *__restrict__ in, const uint8_t *__restrict__ st, size_t i,
                                            int64_t &ptr, int64_t &cnt) {
    size_t j = i;  // bytes (j, i) carry state Close; j is the ',' (state Sep in front of it)
    while (st[j - 1] == ST_CLOSE) j--;
    j--;
    size_t k = j;  // bytes (k, j) carry state Sep; k is the '<'
    while (st[k - 1] == ST_SEP) k--;
    k--;
    ptr = go_atoi_dev(in, k + 1, j);
    cnt = go_atoi_dev(in, j + 1, i);
}

enum : uint32_t { ERR_BAD_REF = 1u, FLAG_NEEDS_UNESCAPE = 2u };

// out-size contribution of byte i (literal: 1, closing '>': cnt, else 0)
__device__ __forceinline__ uint64_t tok_contrib(const uint8_t *__restrict__ in, const uint8_t *__restrict__ st,
                                                size_t i, uint8_t b, uint8_t s, uint32_t *err) {
    if (s == ST_OPEN) {
        if (b == 0x3C) return 0;
        if (err && (b == 0x5C || b == 0xFF)) atomicOr(err, FLAG_NEEDS_UNESCAPE);  // rare bytes
        return 1;
    }
    if (s == ST_CLOSE && b == 0x3E) {
        int64_t ptr, cnt;
        parse_token(in, st, i, ptr, cnt);
        if (cnt < 0 || ptr < cnt) {  // a = len-ptr, need 0 <= a <= a+cnt <= len
            if (err) atomicOr(err, ERR_BAD_REF);
            return 0;
        }
        return (uint64_t)cnt;
    }
    return 0;
}

__global__ void __launch_bounds__(kTileThreads) k_tok_sizes(const uint8_t *__restrict__ in,
                                                            const uint8_t *__restrict__ st, size_t n,
                                                            uint64_t *__restrict__ tile_out,
                                                            uint32_t *__restrict__ err) {
    __shared__ uint64_t sm[33];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint64_t c = 0;
    if (base < n) {
        uint8_t v[16], sv[16];
        load16(in, base, n, 0, v);
        load16(st, base, n, 0, sv);
        const int valid = (int)min((size_t)16, n - base);
        for (int k = 0; k < valid; k++) c += tok_contrib(in, st, base + k, v[k], sv[k], err);
    }
    uint64_t total;
    block_exclusive_sum<uint64_t>(c, sm, total);
    if (threadIdx.x == 0) tile_out[blockIdx.x] = total;
}

// ============================================================================= K5c scatter

// sb[o] = literal bytes; dist[o] = 0 for literals, pointer distance for referenced bytes.
__global__ void __launch_bounds__(kTileThreads) k_tok_scatter(const uint8_t *__restrict__ in,
                                                              const uint8_t *__restrict__ st, size_t n,
                                                              const uint64_t *__restrict__ tile_off,
                                                              uint8_t *__restrict__ sb, uint32_t *__restrict__ dist,
                                                              uint32_t *__restrict__ err) {
    __shared__ uint64_t sm[33];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t v[16], sv[16];
    int valid = 0;
    uint64_t c = 0;
    if (base < n) {
        load16(in, base, n, 0, v);
        load16(st, base, n, 0, sv);
        valid = (int)min((size_t)16, n - base);
        for (int k = 0; k < valid; k++) c += tok_contrib(in, st, base + k, v[k], sv[k], nullptr);
    }
    uint64_t total;
    uint64_t o = tile_off[blockIdx.x] + block_exclusive_sum<uint64_t>(c, sm, total);
    for (int k = 0; k < valid; k++) {
        const uint8_t b = v[k], s = sv[k];
        if (s == ST_OPEN) {
            if (b != 0x3C) {
                sb[o] = b;
                dist[o] = 0;
                o++;
            }
        } else if (s == ST_CLOSE && b == 0x3E) {
            int64_t ptr, cnt;
            parse_token(in, st, base + k, ptr, cnt);
            if (cnt < 0 || ptr < cnt) continue;  // already flagged by k_tok_sizes
            if ((uint64_t)ptr > o) {             // absolutePointer < 0 (lzss.go:349)
                atomicOr(err, ERR_BAD_REF);
                // keep offsets consistent: mark the bytes as literals of value 0
                for (int64_t q = 0; q < cnt; q++) {
                    sb[o + q] = 0;
                    dist[o + q] = 0;
                }
            } else {
                for (int64_t q = 0; q < cnt; q++) dist[o + q] = (uint32_t)ptr;
            }
            o += (uint64_t)cnt;
        }
    }
}

// ============================================================================= K6 resolve

constexpr int kHops = 64;

// Every referenced byte follows its source chain to a literal.  Chains longer than kHops are
// shortened in place (dist[o] := distance to the furthest ancestor reached) and finished by a
// later round; both the old and the new distance name a true ancestor, so concurrent readers
// are safe.  Literal bytes (dist == 0) are never written here.
// `work` == nullptr: first round, one thread per output byte; unfinished bytes are appended to
// `work_out`.  Later rounds walk the previous round's list only.
__global__ void __launch_bounds__(256) k_resolve(uint8_t *__restrict__ sb, uint32_t *__restrict__ dist, size_t n,
                                                 const uint32_t *__restrict__ work, uint32_t *__restrict__ work_out,
                                                 uint32_t *__restrict__ work_count) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t o = work ? work[t] : t;
    uint32_t d = dist[o];
    if (d == 0) return;
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

__global__ void __launch_bounds__(kTileThreads) k_unesc_reduce(const uint8_t *__restrict__ in, size_t n,
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

// per tile: incoming escape state and output offset; total output size
__global__ void __launch_bounds__(1024) k_unesc_spine(const UnescAgg *__restrict__ tile_agg, size_t tiles,
                                                      uint8_t *__restrict__ tile_state,
                                                      uint64_t *__restrict__ tile_off, uint64_t *__restrict__ total) {
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

__global__ void __launch_bounds__(kTileThreads) k_unesc_apply(const uint8_t *__restrict__ in, size_t n,
                                                              const uint8_t *__restrict__ tile_state,
                                                              const uint64_t *__restrict__ tile_off,
                                                              uint8_t *__restrict__ out) {
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
    DevBuf tmap, tstate, st, tout, toff, err;
    RSN_TRY(tmap.alloc(tiles, s));
    RSN_TRY(tstate.alloc(tiles, s));
    RSN_TRY(st.alloc(n + 16, s));
    RSN_TRY(tout.alloc(tiles * 8, s));
    RSN_TRY(toff.alloc((tiles + 1) * 8, s));
    RSN_TRY(err.alloc(16, s));
    RSN_CUDA(cudaMemsetAsync(err.p, 0, 16, s));
    tr.mark("alloc");
    RSN_LAUNCH(k_tok_reduce, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tmap.as<uint8_t>());
    RSN_LAUNCH(k_tok_spine, 1, 1024, 0, s, tmap.as<uint8_t>(), tiles, tstate.as<uint8_t>());
    RSN_LAUNCH(k_tok_states, (unsigned)tiles, kTileThreads, 0, s, d_in, n, tstate.as<uint8_t>(), st.as<uint8_t>());
    RSN_LAUNCH(k_tok_sizes, (unsigned)tiles, kTileThreads, 0, s, d_in, st.as<uint8_t>(), n, tout.as<uint64_t>(),
               err.as<uint32_t>());
    RSN_TRY(spine_scan_u64(tout.as<uint64_t>(), toff.as<uint64_t>(), toff.as<uint64_t>() + tiles, tiles, s));
    uint64_t sbn = 0;
    RSN_TRY(read_u64(toff.as<uint64_t>() + tiles, &sbn, s));
    Ctx &c = ctx();
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars, err.p, 8, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(cudaStreamSynchronize(s));
    if ((uint32_t)c.h_scalars[0] & ERR_BAD_REF) return RSN_ERR_BAD_REFERENCE;
    if (sbn >= (1ull << 32)) return RSN_ERR_UNSUPPORTED;  // u32 source distances (documented limit)
    tr.mark("states+sizes");
    const bool needs_unescape = ((uint32_t)c.h_scalars[0] & FLAG_NEEDS_UNESCAPE) != 0;
    DevBuf sb, dist;  // sb becomes the result itself when no literal needs un-escaping
    RSN_TRY(sb.alloc_out(sbn + 16, s));
    RSN_TRY(dist.alloc(sbn * 4 + 16, s));
    tr.mark("alloc sb/dist");
    RSN_LAUNCH(k_tok_scatter, (unsigned)tiles, kTileThreads, 0, s, d_in, st.as<uint8_t>(), n, toff.as<uint64_t>(),
               sb.as<uint8_t>(), dist.as<uint32_t>(), err.as<uint32_t>());
    tr.mark("scatter");
    if (sbn) {
        // round 0 visits every byte; bytes whose chain is deeper than kHops go to a work list (their
        // distance already shortened), and later rounds only walk that list
        DevBuf wl[2];
        RSN_TRY(wl[0].alloc(sbn * 4 + 16, s));
        RSN_TRY(wl[1].alloc(sbn * 4 + 16, s));
        uint32_t *count = err.as<uint32_t>() + 1;
        size_t todo = (size_t)sbn;
        for (int round = 0; todo; round++) {
            RSN_CUDA(cudaMemsetAsync(count, 0, 4, s));
            RSN_LAUNCH(k_resolve, (unsigned)div_up(todo, 256), 256, 0, s, sb.as<uint8_t>(), dist.as<uint32_t>(), todo,
                       round ? wl[(round + 1) & 1].as<uint32_t>() : (const uint32_t *)nullptr, wl[round & 1].as<uint32_t>(),
                       count);
            RSN_CUDA(cudaMemcpyAsync(c.h_scalars, err.p, 8, cudaMemcpyDeviceToHost, s));
            RSN_CUDA(cudaStreamSynchronize(s));
            const uint32_t e = (uint32_t)c.h_scalars[0];
            if (e & ERR_BAD_REF) return RSN_ERR_BAD_REFERENCE;
            todo = (size_t)(uint32_t)(c.h_scalars[0] >> 32);
        }
    }
    tr.mark("resolve");
    if (!needs_unescape) {  // no 0x5C / 0xFF among the literals: the buffer is already the answer
        *d_out = (uint8_t *)sb.release();
        *out_n = (size_t)sbn;
        return RSN_OK;
    }
    const int rc = lzss_unescape(sb.as<uint8_t>(), (size_t)sbn, d_out, out_n, s);
    tr.mark("unescape");
    return rc;
}

}  // namespace rsn
