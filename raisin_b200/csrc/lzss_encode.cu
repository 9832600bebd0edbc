// lzss_encode.cu — LZSS compress (variant A = lz.CompressAsync, lzss.go:109-184).
//
// Pipeline (all on device):
//   K1  escape-expand            EncodeOpeningSymbols, lzss.go:369-389
//   K2  per-position longest match (every compressorWorker, lzss.go:156-184, in parallel)
//   K3  greedy-parse reconstruction: the reference's sequential merge loop (lzss.go:134-151)
//       visits i, i+max(L,1), ...; we rebuild that chain with per-block exit tables composed
//       through a 64-ary hierarchy, so no pass is sequential in n
//   K4  token sizing + emit      getEncoding, lzss.go:318-320 and the `<` rule at lzss.go:143
#include "batch.cuh"
#include "common.cuh"
#include "lzss.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace rsn {

// ============================================================================= K1 escape

__device__ __forceinline__ bool is_special(uint8_t v) { return v == 0x5C || v == 0xFF; }

// tile_cnt[t] = escaped size of tile t; *touched != 0 iff some byte changes ('<', 0x5C or 0xFF)
__device__ __forceinline__ void escape_count_body(const uint8_t *__restrict__ in, size_t n,
                                                  uint64_t *__restrict__ tile_cnt, uint32_t *__restrict__ touched) {
    __shared__ uint32_t sm[33];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint32_t cnt = 0, changes = 0;
    if (base < n) {
        uint8_t v[16];
        load16(in, base, n, 0, v);
        const int valid = (int)min((size_t)16, n - base);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            cnt += (k < valid) ? 1u + (is_special(v[k]) ? 1u : 0u) : 0u;
            changes += (k < valid && (is_special(v[k]) || v[k] == 0x3C)) ? 1u : 0u;
        }
    }
    uint32_t total;
    block_exclusive_sum<uint32_t>(cnt, sm, total);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
    if (__syncthreads_or(changes != 0) && threadIdx.x == 0) *touched = 1;
}
__global__ void __launch_bounds__(kTileThreads) k_escape_count(const uint8_t *__restrict__ in, size_t n,
                                                               uint64_t *__restrict__ tile_cnt,
                                                               uint32_t *__restrict__ touched) {
    escape_count_body(in, n, tile_cnt, touched);
}

__device__ __forceinline__ void escape_apply_body(const uint8_t *__restrict__ in, size_t n,
                                                  const uint64_t *__restrict__ tile_off, uint8_t *__restrict__ out) {
    __shared__ uint32_t sm[33];
    __shared__ uint8_t stage[2 * kTile];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint8_t v[16];
    int valid = 0;
    uint32_t cnt = 0;
    if (base < n) {
        load16(in, base, n, 0, v);
        valid = (int)min((size_t)16, n - base);
#pragma unroll
        for (int k = 0; k < 16; k++) cnt += (k < valid) ? 1u + (is_special(v[k]) ? 1u : 0u) : 0u;
    }
    uint32_t total;
    uint32_t pos = block_exclusive_sum<uint32_t>(cnt, sm, total);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < valid) {
            uint8_t b = v[k];
            if (b == 0x3C) {
                b = 0xFF;  // '<' is remapped, not escaped (lzss.go:373-377)
            } else if (is_special(b)) {
                stage[pos++] = 0x5C;
            }
            stage[pos++] = b;
        }
    }
    __syncthreads();
    uint8_t *dst = out + tile_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) dst[i] = stage[i];
}
__global__ void __launch_bounds__(kTileThreads) k_escape_apply(const uint8_t *__restrict__ in, size_t n,
                                                               const uint64_t *__restrict__ tile_off,
                                                               uint8_t *__restrict__ out) {
    escape_apply_body(in, n, tile_off, out);
}

// On return *enc_ptr is the escaped buffer: either `enc` (owned) or d_in itself when no byte of
// the input needs escaping or remapping (then nothing is copied).
int lzss_escape(const uint8_t *d_in, size_t n, DevBuf &enc, const uint8_t **enc_ptr, size_t *enc_n, cudaStream_t s) {
    *enc_ptr = d_in;
    if (n == 0) {
        *enc_n = 0;
        return RSN_OK;
    }
    const size_t tiles = div_up(n, kTile);
    DevBuf cnt, off;
    RSN_TRY(cnt.alloc(tiles * 8, s));
    RSN_TRY(off.alloc((tiles + 2) * 8, s));
    uint32_t *touched = reinterpret_cast<uint32_t *>(off.as<uint64_t>() + tiles + 1);
    RSN_CUDA(cudaMemsetAsync(touched, 0, 8, s));
    RSN_LAUNCH(k_escape_count, (unsigned)tiles, kTileThreads, 0, s, d_in, n, cnt.as<uint64_t>(), touched);
    RSN_TRY(spine_scan_u64(cnt.as<uint64_t>(), off.as<uint64_t>(), off.as<uint64_t>() + tiles, tiles, s));
    Ctx &c = ctx();
    RSN_CUDA(cudaMemcpyAsync(c.h_scalars, off.as<uint64_t>() + tiles, 16, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));
    const uint64_t total = c.h_scalars[0];
    if ((uint32_t)c.h_scalars[1] == 0 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0) {
        *enc_n = n;  // identity: use the input in place
        return RSN_OK;
    }
    RSN_TRY(enc.alloc(total + 64, s));
    *enc_ptr = enc.as<uint8_t>();
    RSN_LAUNCH(k_escape_apply, (unsigned)tiles, kTileThreads, 0, s, d_in, n, off.as<uint64_t>(), enc.as<uint8_t>());
    *enc_n = (size_t)total;
    return RSN_OK;
}

// ============================================================================= K3 parse
//
// The reference's merge loop (lzss.go:134-151) visits position 0, then i + max(L(i),1), ...; the
// iterative variant (lzss.go:240-296) visits p, then p + 1 (no match may start at p) or
// p + max(L,1) + 1 (the byte that ends a match is always a literal).  Either way the visited set
// is the orbit of 0 under a per-position jump that only looks forward by at most J bytes.
//   k_parse_exits   per 4096-position block, one warp, a right-to-left recurrence: for EVERY position
//                   of the block, the first orbit position at or beyond the block end (u16, relative)
//   k_parse_up/top/down   a 64-ary hierarchy of such exit tables gives every block's true entry
//   k_emit_plan     per block, one warp: the orbit from the block's entry, threaded through 32
//                   sub-ranges whose speculative orbits were walked in parallel (see emit_plan_warp);
//                   stores the visited bitmap and the block's output size
//   k_emit_write    per block: token sizes -> block scan -> tokens staged in shared memory ->
//                   coalesced copy to the output

constexpr int kPB = 4096;        // parse block (positions)
constexpr int kPT = 256;         // threads per CTA
constexpr int kPI = kPB / kPT;   // 16 consecutive positions per thread
constexpr int kFan = 16;         // hierarchy fan-out (single stream)

struct ParseCfg {
    uint32_t W;            // effective window
    uint32_t J;            // largest jump: W (variant A) or W + 1 (variant B)
    int variant;           // RSN_LZSS_ASYNC / RSN_LZSS_ITER
    const uint32_t *sbits; // variant B: bit p = "a match may start at p" (FindReverse, lzss.go:423-433)
};

__device__ __forceinline__ bool sbit(const ParseCfg &cfg, size_t g) {
    return (__ldg(cfg.sbits + (g >> 5)) >> (g & 31)) & 1u;
}

struct ParseLevels {
    int top;                 // highest level; level l regions have size kPB * kFan^l
    size_t regions[8];       // region count per level
    size_t rsize[8];         // region size per level
};

// one orbit step at level `lvl` from absolute position p (p < n, p inside region p / rsize)
__device__ __forceinline__ size_t level_step(size_t p, int lvl, size_t rsize, const uint16_t *__restrict__ E0,
                                             const uint16_t *__restrict__ T, uint32_t J, size_t n) {
    const size_t r = p / rsize;
    const size_t end = (r + 1) * rsize;
    if (lvl == 0) return end + __ldg(E0 + p);
    // a region cut short by the end of the array exits relative to that end (the array may be one
    // shard of a longer stream, whose parse goes on in the next shard)
    return min(end, n) + __ldg(T + r * (size_t)(J + 1) + (p - r * rsize));
}

// T_l[r][rel] for rel in [0, J]: follow level l-1 until leaving region r (or the input).
__device__ __forceinline__ void parse_up_body(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tprev,
                                              uint16_t *__restrict__ Tcur, int lvl, size_t rsize_prev,
                                              size_t rsize_cur, uint32_t J, size_t n, size_t r, uint32_t rel) {
    if (rel > J) return;
    const size_t start = r * rsize_cur, end = min(start + rsize_cur, n);
    size_t p = start + rel;
    while (p < end) p = level_step(p, lvl - 1, rsize_prev, E0, Tprev, J, n);
    Tcur[r * (size_t)(J + 1) + rel] = (uint16_t)(p - end);
}
__global__ void k_parse_up(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tprev,
                           uint16_t *__restrict__ Tcur, int lvl, size_t rsize_prev, size_t rsize_cur, uint32_t J,
                           size_t n) {
    // regions along x (a stream of 8 GiB has 131072 level-1 regions; y stops at 65535)
    parse_up_body(E0, Tprev, Tcur, lvl, rsize_prev, rsize_cur, J, n, blockIdx.x, blockIdx.y * blockDim.x + threadIdx.x);
}

// sequential walk over the (<= kFan) top-level regions
__device__ __forceinline__ void parse_top_body(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Ttop,
                                               int lvl, size_t rsize, size_t regions, uint32_t J, size_t n,
                                               uint64_t *__restrict__ entry, size_t first = 0) {
    size_t p = first;  // where the orbit enters the array (0 for a whole stream)
    for (size_t r = 0; r < regions; r++) {
        entry[r] = p;
        const size_t end = (r + 1) * rsize;
        if (p < end && p < n) p = level_step(p, lvl, rsize, E0, Ttop, J, n);
    }
}
__global__ void k_parse_top(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Ttop, int lvl,
                            size_t rsize, size_t regions, uint32_t J, size_t n, uint64_t *__restrict__ entry,
                            size_t first) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    parse_top_body(E0, Ttop, lvl, rsize, regions, J, n, entry, first);
}
// One shard of a longer stream: where the orbit leaves the array (relative to its end) for every
// possible entry offset 0..J.  Composed over the shards on the host, this gives every shard's true entry.
__global__ void k_shard_exits(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Ttop, int lvl,
                              size_t rsize, uint32_t J, size_t n, uint16_t *__restrict__ exits) {
    const uint32_t rel = blockIdx.x * blockDim.x + threadIdx.x;
    if (rel > J) return;
    size_t p = rel;
    while (p < n) p = level_step(p, lvl, rsize, E0, Ttop, J, n);
    exits[rel] = (uint16_t)(p - n);
}

// entries of the children (level lvl-1) of each level-lvl region
__device__ __forceinline__ void parse_down_body(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tchild,
                                                int child_lvl, size_t rsize_child, size_t regions_parent,
                                                size_t regions_child, uint32_t J, size_t n,
                                                const uint64_t *__restrict__ entry_parent,
                                                uint64_t *__restrict__ entry_child, int fan) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= regions_parent) return;
    size_t p = entry_parent[r];
    for (int c = 0; c < fan; c++) {
        const size_t cr = r * fan + c;
        if (cr >= regions_child) break;
        entry_child[cr] = p;
        const size_t end = (cr + 1) * rsize_child;
        if (p < end && p < n) p = level_step(p, child_lvl, rsize_child, E0, Tchild, J, n);
    }
}
__global__ void k_parse_down(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tchild, int child_lvl,
                             size_t rsize_child, size_t regions_parent, size_t regions_child, uint32_t J, size_t n,
                             const uint64_t *__restrict__ entry_parent, uint64_t *__restrict__ entry_child) {
    parse_down_body(E0, Tchild, child_lvl, rsize_child, regions_parent, regions_child, J, n, entry_parent, entry_child,
                    kFan);
}

// ============================================================================= K4 emit

__device__ __forceinline__ int ndig_u64(uint64_t v) {
    int d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}

// What the orbit point g emits.  Variant A (lzss.go:141-150): literal, or token iff strictly
// shorter than the match, else the raw match.  Variant B (lzss.go:268-289): the match (token iff
// not longer than the match) followed by the literal that ended it; a literal alone where no match
// may start.
struct Tok {
    uint32_t bytes;   // total bytes emitted at this orbit point
    uint32_t L;       // match length (0: single literal)
    uint64_t ptr;     // pointer written in the token
    bool as_token;    // "<ptr,L>" instead of the raw bytes
    bool tail;        // variant B: one literal byte follows the match
};

__device__ __forceinline__ Tok token_at(const ParseCfg &cfg, uint32_t packed, size_t g, size_t n, bool s_ok) {
    Tok t;
    const uint32_t L = packed >> 16, off = packed & 0xFFFFu;
    t.tail = false;
    if (cfg.variant == RSN_LZSS_ASYNC) {
        t.L = L;
        t.ptr = off;
        if (L == 0) {
            t.as_token = false;
            t.bytes = 1;
            return t;
        }
        const uint32_t tl = 3 + ndig_u32(off) + ndig_u32(L);
        t.as_token = tl < L;  // strict '<' (lzss.go:143)
        t.bytes = t.as_token ? tl : L;
        return t;
    }
    if (!s_ok) {  // FindReverse found nothing: plain literal (lzss.go:289)
        t.L = 0;
        t.ptr = 0;
        t.as_token = false;
        t.bytes = 1;
        return t;
    }
    const uint32_t Lb = L ? L : 1u;
    t.L = Lb;
    // lzss.go:256: pointer = len(searchBuffer) - (index inside the WINDOW slice)
    t.ptr = Lb >= 2 ? (uint64_t)off + (g > cfg.W ? g - cfg.W : 0) : 1;
    const uint32_t tl = 3 + ndig_u64(t.ptr) + ndig_u32(Lb);
    t.as_token = Lb >= 2 && tl <= Lb;  // '<=' (lzss.go:272); a 1-byte match is never shorter as a token
    t.tail = g + Lb < n;
    t.bytes = (t.as_token ? tl : Lb) + (t.tail ? 1u : 0u);
    return t;
}

__device__ __forceinline__ uint8_t *put_dec(uint8_t *o, uint32_t v) {
    const int d = ndig_u32(v);
    for (int k = d - 1; k >= 0; k--) {
        o[k] = (uint8_t)('0' + v % 10);
        v /= 10;
    }
    return o + d;
}

__device__ __forceinline__ uint8_t *put_dec64(uint8_t *o, uint64_t v) {
    const int d = ndig_u64(v);
    for (int k = d - 1; k >= 0; k--) {
        o[k] = (uint8_t)('0' + v % 10);
        v /= 10;
    }
    return o + d;
}

// visited bitmap (one u16 per 16 positions) and output bytes of every block.
//
// One WARP per block.  The orbit of the block's entry under p -> p + jump(p) is a chain of up to 4096
// dependent steps; two orbits that start a few bytes apart merge after a step or two (both take the
// same matches), which is what makes it parallel: every lane first walks the orbit of the START of
// its 128-position sub-range and records it as a bit mask; then the true orbit is threaded through
// the sub-ranges in order, each lane walking from the point where the true orbit enters its
// sub-range only until it meets its speculative orbit (from there on the two are the same).  About
// 37 + 3 steps per lane instead of 12 rounds of pointer doubling over all 4096 positions by 256
// threads (a sixth of the instructions).
constexpr int kPlanWarps = 4;                       // blocks per CTA
constexpr int kPlanSub = kPB / 32;                  // 128 positions per lane
constexpr int kPlanPitch = kPB + 2 * 32;            // one padding word per sub-range: lanes at the same
                                                    // offset of their sub-ranges hit different banks
__device__ __forceinline__ uint32_t plan_idx(uint32_t p) { return p + ((p / kPlanSub) << 1); }

// jump[] of one block for one warp (swizzled by plan_idx); positions >= n get 1
__device__ __forceinline__ void plan_load_jumps(const ParseCfg &cfg, const uint32_t *__restrict__ lo, size_t n,
                                                size_t start, uint16_t *jump) {
    const unsigned lane = threadIdx.x & 31;
    // ---- jumps of the block's positions (positions >= n: 1), 4 positions per lane and turn; eight
    // loads in flight per lane (one load per turn left the warp waiting for memory 32 times)
    constexpr int kTurns = kPB / 128, kBatch = 8;
    for (uint32_t it0 = 0; it0 < kTurns; it0 += kBatch) {
      uint4 vv[kBatch];
      const bool whole = start + (size_t)(it0 + kBatch) * 128 <= n;
      if (whole) {
#pragma unroll
          for (int q = 0; q < kBatch; q++)
              vv[q] = __ldg(reinterpret_cast<const uint4 *>(lo + start + (it0 + q) * 128 + lane * 4));
      }
#pragma unroll
      for (int q = 0; q < kBatch; q++) {
        const uint32_t it = it0 + q;
        const uint32_t p = it * 128 + lane * 4;
        uint32_t L[4];
        if (whole) {
            L[0] = vv[q].x >> 16;
            L[1] = vv[q].y >> 16;
            L[2] = vv[q].z >> 16;
            L[3] = vv[q].w >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) L[k] = start + p + k < n ? (__ldg(lo + start + p + k) >> 16) : 0u;
        }
        uint32_t sb = 0;
        if (cfg.variant == RSN_LZSS_ITER && start + p < n)  // 4 consecutive S bits (start + p is a multiple of 4)
            sb = (__ldg(cfg.sbits + ((start + p) >> 5)) >> ((start + p) & 31)) & 0xFu;
        uint32_t j[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            j[k] = L[k] ? L[k] : 1u;
            if (cfg.variant == RSN_LZSS_ITER) j[k] = ((sb >> k) & 1u) ? j[k] + 1 : 1u;
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(jump + plan_idx(p));  // even u16 index: 4-byte aligned
        dst[0] = j[0] | (j[1] << 16);
        dst[1] = j[2] | (j[3] << 16);
      }
    }
    __syncwarp();
}

// E0[p] for every position p of one block (one warp): the first orbit position at or beyond the block
// end, relative to it (0: the orbit leaves the input inside the block).  exit(p) = exit(p + jump(p)) is a
// recurrence from the right: every lane resolves its 128-position sub-range from its last position
// down (a pointer either leaves the sub-range or lands on an entry that is already resolved), then
// the sub-ranges are stitched from the last to the first, 128 positions at a time.  One pass over
// the positions instead of ~12 rounds of pointer doubling over all of them.
__device__ __forceinline__ void parse_exits_warp(const ParseCfg &cfg, const uint32_t *__restrict__ lo, size_t n,
                                                 size_t block, uint16_t *__restrict__ E0, uint16_t *jump) {
    const unsigned lane = threadIdx.x & 31;
    const size_t start = block * kPB;
    const uint32_t nrel = (uint32_t)min((size_t)kPB, n - start);
    plan_load_jumps(cfg, lo, n, start, jump);
    __syncwarp();
    const uint32_t sub0 = lane * kPlanSub, sub1 = min(sub0 + kPlanSub, nrel);
    // in place: entries above p hold exits from the sub-range, entries up to p still hold jumps
    for (uint32_t p = sub1; p-- > sub0;) {
        uint32_t t = p + jump[plan_idx(p)];
        if (t < sub1) t = jump[plan_idx(t)];
        jump[plan_idx(p)] = (uint16_t)t;
    }
    __syncwarp();
    // sub-range 31 (or the last one that has positions) is final; stitch the others from the right
    for (int i = 30; i >= 0; i--) {
        const uint32_t q0 = (uint32_t)i * kPlanSub + lane * 4;
        if (q0 < nrel) {
            uint32_t *w = reinterpret_cast<uint32_t *>(jump + plan_idx(q0));
            uint32_t v[4] = {w[0] & 0xFFFFu, w[0] >> 16, w[1] & 0xFFFFu, w[1] >> 16};
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (q0 + k < nrel && v[k] < nrel) v[k] = jump[plan_idx(v[k])];  // lies in a later, final sub-range
            w[0] = v[0] | (v[1] << 16);
            w[1] = v[2] | (v[3] << 16);
        }
        __syncwarp();
    }
    for (uint32_t it = 0; it < kPB / 128; it++) {
        const uint32_t q0 = it * 128 + lane * 4;
        if (q0 >= nrel) break;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(jump + plan_idx(q0));
        const uint32_t v[4] = {w[0] & 0xFFFFu, w[0] >> 16, w[1] & 0xFFFFu, w[1] >> 16};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = v[k] >= (uint32_t)kPB ? v[k] - kPB : 0u;
        if (q0 + 4 <= nrel) {
            *reinterpret_cast<uint2 *>(E0 + start + q0) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        } else {
            for (int k = 0; k < 4; k++)
                if (q0 + k < nrel) E0[start + q0 + k] = (uint16_t)o[k];
        }
    }
}
__global__ void __launch_bounds__(kPlanWarps * 32) k_parse_exits(ParseCfg cfg, const uint32_t *__restrict__ lo, size_t n,
                                                                 uint16_t *__restrict__ E0) {
    __shared__ __align__(16) uint16_t jump[kPlanWarps][kPlanPitch];
    const size_t block = (size_t)blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    if (block * kPB >= n) return;
    parse_exits_warp(cfg, lo, n, block, E0, jump[threadIdx.x >> 5]);
}

__device__ __forceinline__ void emit_plan_warp(const ParseCfg &cfg, const uint32_t *__restrict__ lo, size_t n,
                                               size_t block, const uint64_t *__restrict__ entry0,
                                               uint16_t *__restrict__ visited, uint64_t *__restrict__ blk_bytes,
                                               uint16_t *jump) {
    const unsigned lane = threadIdx.x & 31;
    const size_t start = block * kPB;
    const uint32_t nrel = (uint32_t)min((size_t)kPB, n - start);
    plan_load_jumps(cfg, lo, n, start, jump);
    // ---- speculative orbit of the start of my sub-range
    const uint32_t sub0 = lane * kPlanSub, sub1 = min(sub0 + kPlanSub, nrel);
    unsigned long long spec_lo = 0, spec_hi = 0;
    uint32_t spec_exit = sub0;
    {
        uint32_t p = sub0;
        while (p < sub1) {
            const uint32_t b = p - sub0;
            if (b < 64) spec_lo |= 1ull << b;
            else spec_hi |= 1ull << (b - 64);
            p += jump[plan_idx(p)];
        }
        spec_exit = p;
    }
    // ---- the true orbit, sub-range by sub-range
    unsigned long long fin_lo = 0, fin_hi = 0;
    uint32_t x = nrel;  // where the true orbit stands (relative); nrel: it has left the block
    {
        const uint64_t e = entry0[block];
        if (e >= start && e - start < nrel) x = (uint32_t)(e - start);
    }
    for (unsigned j = 0; j < 32 && x < nrel; j++) {
        if (x / kPlanSub != j) continue;  // the orbit jumps over this sub-range (x is the same in all lanes)
        uint32_t my_exit = 0;
        if (lane == j) {
            unsigned long long path_lo = 0, path_hi = 0;
            uint32_t p = x;
            bool merged = false;
            uint32_t b = 0;
            while (p < sub1) {
                b = p - sub0;
                const bool in_spec = b < 64 ? (spec_lo >> b) & 1ull : (spec_hi >> (b - 64)) & 1ull;
                if (in_spec) {
                    merged = true;
                    break;
                }
                if (b < 64) path_lo |= 1ull << b;
                else path_hi |= 1ull << (b - 64);
                p += jump[plan_idx(p)];
            }
            if (merged) {  // from b on the speculative orbit is the true one
                const unsigned long long keep_lo = b < 64 ? ~0ull << b : 0ull;
                const unsigned long long keep_hi = b < 64 ? ~0ull : ~0ull << (b - 64);
                fin_lo = path_lo | (spec_lo & keep_lo);
                fin_hi = path_hi | (spec_hi & keep_hi);
                my_exit = spec_exit;
            } else {
                fin_lo = path_lo;
                fin_hi = path_hi;
                my_exit = p;
            }
        }
        x = __shfl_sync(0xffffffffu, my_exit, j);
    }
    // ---- output bytes of my orbit points, visited bits
    uint32_t bytes = 0;
    for (int half = 0; half < 2; half++) {
        unsigned long long m = half ? fin_hi : fin_lo;
        while (m) {
            const int k = __ffsll((long long)m) - 1;
            m &= m - 1;
            const uint32_t p = sub0 + half * 64 + k;
            const uint32_t jp = jump[plan_idx(p)];
            if (cfg.variant == RSN_LZSS_ASYNC && jp <= 5) {
                bytes += jp;  // a literal, or a match too short for any token ("<d,d>" is 5 bytes): its own bytes
            } else {
                const size_t g = start + p;
                const bool s_ok = cfg.variant == RSN_LZSS_ITER ? sbit(cfg, g) : true;
                bytes += token_at(cfg, __ldg(lo + g), g, n, s_ok).bytes;
            }
        }
    }
    *reinterpret_cast<uint4 *>(visited + (start >> 4) + lane * (kPlanSub / 16)) =
        make_uint4((uint32_t)fin_lo, (uint32_t)(fin_lo >> 32), (uint32_t)fin_hi, (uint32_t)(fin_hi >> 32));
#pragma unroll
    for (int d = 16; d; d >>= 1) bytes += __shfl_down_sync(0xffffffffu, bytes, d);
    if (lane == 0) blk_bytes[block] = bytes;
}
__global__ void __launch_bounds__(kPlanWarps * 32) k_emit_plan(ParseCfg cfg, const uint32_t *__restrict__ lo, size_t n,
                                                               const uint64_t *__restrict__ entry0,
                                                               uint16_t *__restrict__ visited,
                                                               uint64_t *__restrict__ blk_bytes) {
    __shared__ __align__(16) uint16_t jump[kPlanWarps][kPlanPitch];
    const size_t block = (size_t)blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    if (block * kPB >= n) return;
    emit_plan_warp(cfg, lo, n, block, entry0, visited, blk_bytes, jump[threadIdx.x >> 5]);
}

constexpr int kStage = kPB + 64;  // output bytes of one block never exceed consumed + one token

__device__ __forceinline__ void emit_write_body(const ParseCfg &cfg, const uint8_t *__restrict__ enc,
                                                const uint32_t *__restrict__ lo, size_t n,
                                                const uint16_t *__restrict__ visited,
                                                const uint64_t *__restrict__ blk_off, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t stage[kStage];
    __shared__ uint16_t tokpos[kPB];  // the block's orbit points, in order
    __shared__ uint32_t sm[33];
    const size_t start = (size_t)blockIdx.x * kPB;
    const uint32_t p0 = threadIdx.x * kPI;
    const uint32_t bits = visited[(start >> 4) + threadIdx.x];
    // Orbit points are spread unevenly over the positions (16 in a run of literals, one or two
    // inside matches): with a thread formatting the points of ITS 16 positions a warp waited for its
    // busiest lane (11 of 32 lanes active, 6 warp instructions per position).  The points are
    // compacted first, and every thread takes an equal, contiguous share of them.
    uint32_t ntok;
    {
        uint32_t at = block_exclusive_sum<uint32_t>((uint32_t)__popc(bits), sm, ntok);
        for (uint32_t b = bits; b; b &= b - 1) tokpos[at++] = (uint16_t)(p0 + __ffs(b) - 1);
    }
    __syncthreads();
    const uint32_t per = (ntok + kPT - 1) / kPT;
    const uint32_t k_lo = min(ntok, threadIdx.x * per), k_hi = min(ntok, k_lo + per);
    uint32_t bytes = 0;
    for (uint32_t k = k_lo; k < k_hi; k++) {
        const size_t g = start + tokpos[k];
        const bool s_ok = cfg.variant == RSN_LZSS_ITER ? sbit(cfg, g) : true;
        bytes += token_at(cfg, __ldg(lo + g), g, n, s_ok).bytes;
    }
    uint32_t total;
    uint32_t pos = block_exclusive_sum<uint32_t>(bytes, sm, total);
    for (uint32_t k = k_lo; k < k_hi; k++) {
        const size_t g = start + tokpos[k];
        const bool s_ok = cfg.variant == RSN_LZSS_ITER ? sbit(cfg, g) : true;
        const Tok t = token_at(cfg, __ldg(lo + g), g, n, s_ok);
        uint8_t *o = stage + pos;
        if (t.L == 0) {
            *o++ = __ldg(enc + g);
        } else if (t.as_token) {
            *o++ = '<';
            o = put_dec64(o, t.ptr);
            *o++ = ',';
            o = put_dec(o, t.L);
            *o++ = '>';
        } else {
            for (uint32_t q = 0; q < t.L; q++) o[q] = __ldg(enc + g + q);
            o += t.L;
        }
        if (t.tail) *o++ = __ldg(enc + g + t.L);
        pos += t.bytes;
    }
    __syncthreads();
    uint8_t *dst = out + blk_off[blockIdx.x];
    // coalesced copy: byte head up to 16-byte alignment of dst, then 16-byte vectors, then tail
    const uint32_t mis = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    const uint32_t head = min(mis, total);
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = stage[i];
    const uint32_t nvec = (total - head) / 16;
    for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint8_t *sp = stage + head + v * 16;  // shared side may be unaligned: assemble words
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            w[q] = (uint32_t)sp[4 * q] | ((uint32_t)sp[4 * q + 1] << 8) | ((uint32_t)sp[4 * q + 2] << 16) |
                   ((uint32_t)sp[4 * q + 3] << 24);
        *reinterpret_cast<uint4 *>(dst + head + v * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (uint32_t i = head + nvec * 16 + threadIdx.x; i < total; i += blockDim.x) dst[i] = stage[i];
}
__global__ void __launch_bounds__(kPT) k_emit_write(ParseCfg cfg, const uint8_t *__restrict__ enc,
                                                    const uint32_t *__restrict__ lo, size_t n,
                                                    const uint16_t *__restrict__ visited,
                                                    const uint64_t *__restrict__ blk_off, uint8_t *__restrict__ out) {
    emit_write_body(cfg, enc, lo, n, visited, blk_off, out);
}

// ============================================================================= variant B: start bits

// firstpos[b][parity] = first position of byte b at that position parity (FindReverse scans the
// WHOLE history with stride 2, lzss.go:423-433, so a match may start at p iff byte enc[p] occurs
// earlier at a position of parity (p-1) mod 2).
__global__ void __launch_bounds__(256) k_first_pos(const uint8_t *__restrict__ enc, size_t n,
                                                   unsigned long long *__restrict__ firstpos) {
    __shared__ unsigned long long best[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) best[i] = ~0ull;
    __syncthreads();
    const size_t chunk = 8192;
    const size_t lo_p = (size_t)blockIdx.x * chunk, hi_p = min(n, lo_p + chunk);
    for (size_t p = lo_p + threadIdx.x; p < hi_p; p += blockDim.x)
        atomicMin(&best[(uint32_t)enc[p] * 2 + (p & 1)], (unsigned long long)p);
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x)
        if (best[i] != ~0ull) atomicMin(&firstpos[i], best[i]);
}

__global__ void __launch_bounds__(256) k_start_bits(const uint8_t *__restrict__ enc, size_t n,
                                                    const unsigned long long *__restrict__ firstpos,
                                                    uint32_t *__restrict__ sbits) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (p < n && p > 0) ok = firstpos[(uint32_t)enc[p] * 2 + ((p - 1) & 1)] < p;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && p < n + 32) sbits[p >> 5] = m;
}

// ============================================================================= host orchestration

// The parse of one array of match records, in three steps so that a shard of a longer stream can put
// an exchange between them: tables (no entry needed), resolution from the position where the orbit
// enters the array, emit.
struct ParsePlan {
    ParseLevels lv;
    size_t blocks = 0;
    DevBuf E0, T[8], entry[8], vis, bb, bo;
};

static int parse_build_tables(const ParseCfg &cfg, const uint32_t *d_lo, size_t n, ParsePlan &pp, cudaStream_t s) {
    pp.blocks = div_up(n, kPB);
    const uint32_t J = cfg.J;
    ParseLevels &lv = pp.lv;
    lv.top = 0;
    lv.rsize[0] = kPB;
    lv.regions[0] = pp.blocks;
    while (lv.regions[lv.top] > (size_t)kFan) {
        lv.rsize[lv.top + 1] = lv.rsize[lv.top] * kFan;
        lv.regions[lv.top + 1] = div_up(n, lv.rsize[lv.top + 1]);
        lv.top++;
    }
    RSN_TRY(pp.E0.alloc(pp.blocks * kPB * 2 + 64, s));
    RSN_LAUNCH(k_parse_exits, (unsigned)div_up(pp.blocks, kPlanWarps), kPlanWarps * 32, 0, s, cfg, d_lo, n, pp.E0.as<uint16_t>());
    for (int l = 1; l <= lv.top; l++) {
        RSN_TRY(pp.T[l].alloc(lv.regions[l] * (size_t)(J + 1) * 2, s));
        dim3 grid((unsigned)lv.regions[l], (unsigned)div_up((size_t)J + 1, 256));
        RSN_LAUNCH(k_parse_up, grid, 256, 0, s, pp.E0.as<uint16_t>(), pp.T[l - 1].as<uint16_t>(), pp.T[l].as<uint16_t>(), l,
                   lv.rsize[l - 1], lv.rsize[l], J, n);
    }
    return RSN_OK;
}

// entries of every block for an orbit that enters the array at `first`, orbit bitmap, output size
static int parse_resolve(const ParseCfg &cfg, const uint32_t *d_lo, size_t n, size_t first, ParsePlan &pp,
                         uint64_t *total, cudaStream_t s) {
    const uint32_t J = cfg.J;
    const ParseLevels &lv = pp.lv;
    const size_t blocks = pp.blocks;
    for (int l = 0; l <= lv.top; l++) RSN_TRY(pp.entry[l].alloc(lv.regions[l] * 8, s));
    RSN_LAUNCH(k_parse_top, 1, 32, 0, s, pp.E0.as<uint16_t>(), pp.T[lv.top].as<uint16_t>(), lv.top, lv.rsize[lv.top],
               lv.regions[lv.top], J, n, pp.entry[lv.top].as<uint64_t>(), first);
    for (int l = lv.top; l >= 1; l--) {
        RSN_LAUNCH(k_parse_down, (unsigned)div_up(lv.regions[l], 128), 128, 0, s, pp.E0.as<uint16_t>(),
                   pp.T[l - 1].as<uint16_t>(), l - 1, lv.rsize[l - 1], lv.regions[l], lv.regions[l - 1], J, n,
                   pp.entry[l].as<uint64_t>(), pp.entry[l - 1].as<uint64_t>());
    }
    RSN_TRY(pp.vis.alloc(blocks * (kPB / 16) * 2 + 16, s));
    RSN_TRY(pp.bb.alloc(blocks * 8, s));
    RSN_TRY(pp.bo.alloc((blocks + 1) * 8, s));
    RSN_LAUNCH(k_emit_plan, (unsigned)div_up(blocks, kPlanWarps), kPlanWarps * 32, 0, s, cfg, d_lo, n, pp.entry[0].as<uint64_t>(), pp.vis.as<uint16_t>(),
               pp.bb.as<uint64_t>());
    RSN_TRY(spine_scan_u64(pp.bb.as<uint64_t>(), pp.bo.as<uint64_t>(), pp.bo.as<uint64_t>() + blocks, blocks, s));
    return read_u64(pp.bo.as<uint64_t>() + blocks, total, s);
}

static int parse_emit(const ParseCfg &cfg, const uint8_t *d_enc, const uint32_t *d_lo, size_t n, ParsePlan &pp,
                      uint8_t *d_dst, cudaStream_t s) {
    RSN_LAUNCH(k_emit_write, (unsigned)pp.blocks, kPT, 0, s, cfg, d_enc, d_lo, n, pp.vis.as<uint16_t>(),
               pp.bo.as<uint64_t>(), d_dst);
    return RSN_OK;
}

static int parse_and_emit(const uint8_t *d_enc, size_t n, ParseCfg cfg, const uint32_t *d_lo, uint8_t **d_out,
                          size_t *out_n, cudaStream_t s) {
    ParsePlan pp;
    RSN_TRY(parse_build_tables(cfg, d_lo, n, pp, s));
    uint64_t total = 0;
    RSN_TRY(parse_resolve(cfg, d_lo, n, 0, pp, &total, s));
    DevBuf out;
    RSN_TRY(out.alloc_out(total + 16, s));
    RSN_TRY(parse_emit(cfg, d_enc, d_lo, n, pp, out.as<uint8_t>(), s));
    *d_out = (uint8_t *)out.release();
    *out_n = (size_t)total;
    return RSN_OK;
}

int lzss_effective_window(int64_t window, size_t enc_n, uint32_t *W) {
    // window <= 0: unbounded (lzss.go:125).  A window >= enc_n behaves exactly like enc_n.
    uint64_t w = window <= 0 ? (uint64_t)enc_n : (uint64_t)window;
    if (w > enc_n) w = enc_n;
    if (w < 1) w = 1;
    if (w > kMaxWindow) return RSN_ERR_UNSUPPORTED;
    *W = (uint32_t)w;
    return RSN_OK;
}

// parse + emit over caller-provided match arrays (enc must be the escaped buffer they refer to)
int lzss_emit_dev(const uint8_t *d_enc, size_t n, int64_t window, int variant, const uint32_t *d_packed,
                  uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    if (variant != RSN_LZSS_ASYNC && variant != RSN_LZSS_ITER) return RSN_ERR_INVALID_ARG;
    ArenaScope scope(s);
    if (n == 0) {
        DevBuf out;
        RSN_TRY(out.alloc_out(16, s));
        *d_out = (uint8_t *)out.release();
        *out_n = 0;
        return RSN_OK;
    }
    uint32_t W = 0;
    RSN_TRY(lzss_effective_window(window, n, &W));
    // the parse kernels read whole 4096-position blocks of the arrays: work on a padded copy
    DevBuf lo, sbits, fp;
    const size_t padded = div_up(n, kPB) * kPB;
    RSN_TRY(lo.alloc(padded * 4 + 64, s));
    RSN_CUDA(cudaMemcpyAsync(lo.p, d_packed, n * 4, cudaMemcpyDeviceToDevice, s));
    ParseCfg cfg{W, W, variant, nullptr};
    if (variant == RSN_LZSS_ITER) {
        cfg.J = W + 1;
        RSN_TRY(fp.alloc(512 * 8, s));
        const size_t sb_bytes = (padded / 32 + 4) * 4;
        RSN_TRY(sbits.alloc(sb_bytes, s));
        RSN_CUDA(cudaMemsetAsync(sbits.p, 0, sb_bytes, s));
        RSN_CUDA(cudaMemsetAsync(fp.p, 0xFF, 512 * 8, s));
        RSN_LAUNCH(k_first_pos, (unsigned)div_up(n, 8192), 256, 0, s, d_enc, n, fp.as<unsigned long long>());
        RSN_LAUNCH(k_start_bits, (unsigned)div_up(n + 31, 256), 256, 0, s, d_enc, n, fp.as<unsigned long long>(),
                   sbits.as<uint32_t>());
        cfg.sbits = sbits.as<uint32_t>();
    }
    return parse_and_emit(d_enc, n, cfg, lo.as<uint32_t>(), d_out, out_n, s);
}

int lzss_escape_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    ArenaScope scope(s);
    DevBuf enc_buf, out;
    const uint8_t *enc = nullptr;
    size_t en = 0;
    RSN_TRY(lzss_escape(d_in, n, enc_buf, &enc, &en, s));
    RSN_TRY(out.alloc_out(en + 64, s));
    if (en) RSN_CUDA(cudaMemcpyAsync(out.p, enc, en, cudaMemcpyDeviceToDevice, s));
    RSN_CUDA(stream_wait(s));
    *d_out = (uint8_t *)out.release();
    *out_n = en;
    return RSN_OK;
}

int lzss_compress_dev(const uint8_t *d_in, size_t n, int64_t window, int variant, uint8_t **d_out, size_t *out_n,
                      cudaStream_t s) {
    return lzss_compress_dev_ex(d_in, n, window, variant, nullptr, d_out, out_n, s);
}

int lzss_compress_dev_ex(const uint8_t *d_in, size_t n, int64_t window, int variant, uint32_t *spec_packed,
                         uint8_t **d_out, size_t *out_n, cudaStream_t s) {
    ArenaScope scope(s);
    if (variant != RSN_LZSS_ASYNC && variant != RSN_LZSS_ITER) return RSN_ERR_INVALID_ARG;
    DevBuf enc_buf;
    const uint8_t *enc = nullptr;
    size_t en = 0;
    RSN_TRY(lzss_escape(d_in, n, enc_buf, &enc, &en, s));
    if (en == 0) {
        DevBuf out;
        RSN_TRY(out.alloc_out(16, s));
        *d_out = (uint8_t *)out.release();
        *out_n = 0;
        return RSN_OK;
    }
    uint32_t W = 0;
    RSN_TRY(lzss_effective_window(window, en, &W));
    DevBuf lo, sbits, fp;
    uint32_t *packed = spec_packed;
    if (!(spec_packed && enc == d_in)) {  // no usable speculative arrays: search now
        RSN_TRY(lo.alloc(div_up(en, kPB) * kPB * 4 + 64, s));
        RSN_TRY(lzss_match(enc, en, W, lo.as<uint32_t>(), s));
        packed = lo.as<uint32_t>();
    }
    ParseCfg cfg{W, W, variant, nullptr};
    if (variant == RSN_LZSS_ITER) {
        cfg.J = W + 1;
        RSN_TRY(fp.alloc(512 * 8, s));
        const size_t sb_bytes = (div_up(en, kPB) * kPB / 32 + 4) * 4;  // whole blocks: load_jumps reads 16-bit groups
        RSN_TRY(sbits.alloc(sb_bytes, s));
        RSN_CUDA(cudaMemsetAsync(sbits.p, 0, sb_bytes, s));
        RSN_CUDA(cudaMemsetAsync(fp.p, 0xFF, 512 * 8, s));
        RSN_LAUNCH(k_first_pos, (unsigned)div_up(en, 8192), 256, 0, s, enc, en, fp.as<unsigned long long>());
        RSN_LAUNCH(k_start_bits, (unsigned)div_up(en + 31, 256), 256, 0, s, enc, en, fp.as<unsigned long long>(),
                   sbits.as<uint32_t>());
        cfg.sbits = sbits.as<uint32_t>();
    }
    return parse_and_emit(enc, en, cfg, packed, d_out, out_n, s);
}

// ============================================================================= batches of small files
//
// The same kernels, one launch per stage for a whole group of files (batch.cuh).  File sizes after
// escaping live in the LzFile records on the device; grids are sized for the largest possible file
// and CTAs beyond a file's real size leave at once.

constexpr int kBatchFan = 64;  // batches: at most two levels (files of <= 4 MiB)

struct LzBatch {
    LzFile *files;
    uint32_t window;
    uint64_t *tile_cnt, *tile_off;  // [G][tc_stride]
    size_t tc_stride;
    uint8_t *enc;                   // [G][enc_stride]
    size_t enc_stride;
    uint32_t *packed;               // [G][packed_stride]
    size_t packed_stride;
    uint16_t *E0;                   // [G][e0_stride]
    size_t e0_stride;
    uint16_t *T1;                   // [G][t1_stride]   level-1 exit tables (top == 1 only)
    size_t t1_stride;
    uint64_t *entry0, *entry1;      // [G][entry0_stride], [G][entry1_stride]
    size_t entry0_stride, entry1_stride;
    uint16_t *vis;                  // [G][vis_stride]
    size_t vis_stride;
    uint64_t *bb, *bo;              // [G][bb_stride]
    size_t bb_stride;
    uint64_t *out_n;                // [G]
    int top;                        // 0: one level of blocks; 1: blocks grouped kBatchFan at a time
};

__device__ __forceinline__ ParseCfg batch_cfg(const LzFile &f) {
    return ParseCfg{f.W, f.W, RSN_LZSS_ASYNC, nullptr};
}

__global__ void __launch_bounds__(kTileThreads) kb_escape_count(LzBatch b) {
    LzFile &f = b.files[blockIdx.y];
    if ((size_t)blockIdx.x * kTile >= f.n) return;
    escape_count_body(f.in, (size_t)f.n, b.tile_cnt + (size_t)blockIdx.y * b.tc_stride, &f.touched);
}

// per file: tile offsets, escaped size, effective window (lzss.go:125: the window never exceeds the data)
__global__ void __launch_bounds__(256) kb_escape_finish(LzBatch b) {
    __shared__ uint64_t sm[33];
    LzFile &f = b.files[blockIdx.x];
    const size_t tiles = div_up_dev((size_t)f.n, (size_t)kTile);
    const uint64_t total = cta_scan_u64(b.tile_cnt + (size_t)blockIdx.x * b.tc_stride,
                                        b.tile_off + (size_t)blockIdx.x * b.tc_stride, tiles, sm);
    if (threadIdx.x == 0) {
        const bool alias = f.touched == 0 && (reinterpret_cast<uintptr_t>(f.in) & 15) == 0;
        f.enc = alias ? f.in : b.enc + (size_t)blockIdx.x * b.enc_stride;
        f.touched = alias ? 0u : 1u;
        f.en = total;
        uint64_t w = b.window;
        if (w > total) w = total;
        if (w < 1) w = 1;
        f.W = (uint32_t)w;
    }
}

__global__ void __launch_bounds__(kTileThreads) kb_escape_apply(LzBatch b) {
    const LzFile &f = b.files[blockIdx.y];
    if (!f.touched || (size_t)blockIdx.x * kTile >= f.n) return;
    escape_apply_body(f.in, (size_t)f.n, b.tile_off + (size_t)blockIdx.y * b.tc_stride, const_cast<uint8_t *>(f.enc));
}

__global__ void __launch_bounds__(kPlanWarps * 32) kb_parse_exits(LzBatch b) {
    __shared__ __align__(16) uint16_t jump[kPlanWarps][kPlanPitch];
    const LzFile &f = b.files[blockIdx.y];
    const size_t block = (size_t)blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    if (block * kPB >= f.en) return;
    parse_exits_warp(batch_cfg(f), b.packed + (size_t)blockIdx.y * b.packed_stride, (size_t)f.en, block,
                     b.E0 + (size_t)blockIdx.y * b.e0_stride, jump[threadIdx.x >> 5]);
}

__global__ void kb_parse_up(LzBatch b) {  // grid: (rel chunks, level-1 regions, files)
    const LzFile &f = b.files[blockIdx.z];
    const size_t rsize1 = (size_t)kPB * kBatchFan;
    if ((size_t)blockIdx.y * rsize1 >= f.en) return;
    parse_up_body(b.E0 + (size_t)blockIdx.z * b.e0_stride, nullptr, b.T1 + (size_t)blockIdx.z * b.t1_stride, 1, kPB,
                  rsize1, f.W, (size_t)f.en, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void kb_parse_top(LzBatch b, size_t G) {  // one thread per file
    const size_t y = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= G) return;
    const LzFile &f = b.files[y];
    if (f.en == 0) return;
    const size_t rsize = b.top ? (size_t)kPB * kBatchFan : (size_t)kPB;
    uint64_t *entry = b.top ? b.entry1 + y * b.entry1_stride : b.entry0 + y * b.entry0_stride;
    parse_top_body(b.E0 + y * b.e0_stride, b.top ? b.T1 + y * b.t1_stride : nullptr, b.top, rsize,
                   div_up_dev((size_t)f.en, rsize), f.W, (size_t)f.en, entry);
}

__global__ void kb_parse_down(LzBatch b) {  // top == 1: level-1 regions -> blocks
    const LzFile &f = b.files[blockIdx.y];
    if (f.en == 0) return;
    parse_down_body(b.E0 + (size_t)blockIdx.y * b.e0_stride, nullptr, 0, kPB,
                    div_up_dev((size_t)f.en, (size_t)kPB * kBatchFan), div_up_dev((size_t)f.en, (size_t)kPB), f.W,
                    (size_t)f.en, b.entry1 + (size_t)blockIdx.y * b.entry1_stride,
                    b.entry0 + (size_t)blockIdx.y * b.entry0_stride, kBatchFan);
}

__global__ void __launch_bounds__(kPlanWarps * 32) kb_emit_plan(LzBatch b) {
    __shared__ __align__(16) uint16_t jump[kPlanWarps][kPlanPitch];
    const LzFile &f = b.files[blockIdx.y];
    const size_t block = (size_t)blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    if (block * kPB >= f.en) return;
    emit_plan_warp(batch_cfg(f), b.packed + (size_t)blockIdx.y * b.packed_stride, (size_t)f.en, block,
                   b.entry0 + (size_t)blockIdx.y * b.entry0_stride, b.vis + (size_t)blockIdx.y * b.vis_stride,
                   b.bb + (size_t)blockIdx.y * b.bb_stride, jump[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(256) kb_emit_finish(LzBatch b) {
    __shared__ uint64_t sm[33];
    LzFile &f = b.files[blockIdx.x];
    const size_t blocks = div_up_dev((size_t)f.en, (size_t)kPB);
    const uint64_t total = cta_scan_u64(b.bb + (size_t)blockIdx.x * b.bb_stride, b.bo + (size_t)blockIdx.x * b.bb_stride,
                                        blocks, sm);
    if (threadIdx.x == 0) {
        f.out_n = total;
        b.out_n[blockIdx.x] = total;
    }
}

__global__ void __launch_bounds__(kPT) kb_emit_write(LzBatch b, uint8_t *const *__restrict__ outp) {
    const LzFile &f = b.files[blockIdx.y];
    if ((size_t)blockIdx.x * kPB >= f.en) return;
    emit_write_body(batch_cfg(f), f.enc, b.packed + (size_t)blockIdx.y * b.packed_stride, (size_t)f.en,
                    b.vis + (size_t)blockIdx.y * b.vis_stride, b.bo + (size_t)blockIdx.y * b.bb_stride,
                    outp[blockIdx.y]);
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// lz.CompressAsync (lzss.go:109-154) over every file of the group; window in [1, 4096].
int lzss_compress_batch(const BatchIO &in, BatchIO &out, int64_t window, cudaStream_t s) {
    const size_t G = in.size();
    out.resize(G);
    out.rc = in.rc;
    if (G == 0) return RSN_OK;
    if (window < 1 || window > 4096) return RSN_ERR_UNSUPPORTED;
    ArenaScope scope(s);
    size_t cap = 1;
    for (size_t f = 0; f < G; f++)
        if (in.rc[f] == RSN_OK) cap = std::max<size_t>(cap, in.n[f]);
    if (cap > kBatchMaxFile) return RSN_ERR_UNSUPPORTED;
    const size_t ecap = round_up(2 * cap, 8192);  // escaping at most doubles a file
    const size_t tiles_cap = div_up(cap, kTile), blocks_cap = ecap / kPB;
    LzBatch b{};
    b.window = (uint32_t)window;
    b.top = blocks_cap > (size_t)kBatchFan ? 1 : 0;
    const size_t regions1 = div_up(blocks_cap, kBatchFan);
    if (regions1 > (size_t)kBatchFan) return RSN_ERR_UNSUPPORTED;
    b.tc_stride = tiles_cap + 1;
    b.enc_stride = ecap + 256;
    b.packed_stride = ecap + 64;
    b.e0_stride = ecap + 64;
    b.t1_stride = regions1 * ((size_t)window + 1);
    b.entry0_stride = blocks_cap;
    b.entry1_stride = regions1;
    b.vis_stride = blocks_cap * (kPB / 16);
    b.bb_stride = blocks_cap + 1;

    HostVec<LzFile> h_files(G);
    if (!h_files.data()) return RSN_ERR_NOMEM;
    for (size_t f = 0; f < G; f++) {
        LzFile r{};
        r.in = in.ptr[f];
        r.n = in.rc[f] == RSN_OK ? in.n[f] : 0;
        h_files[f] = r;
    }
    DevBuf files, tcnt, toff, enc, packed, E0, T1, e0, e1, vis, bb, bo, outn, outp;
    RSN_TRY(files.alloc(G * sizeof(LzFile), s));
    RSN_TRY(tcnt.alloc(G * b.tc_stride * 8, s));
    RSN_TRY(toff.alloc(G * b.tc_stride * 8, s));
    RSN_TRY(enc.alloc(G * b.enc_stride, s));
    RSN_TRY(packed.alloc(G * b.packed_stride * 4, s));
    RSN_TRY(E0.alloc(G * b.e0_stride * 2, s));
    if (b.top) RSN_TRY(T1.alloc(G * b.t1_stride * 2, s));
    RSN_TRY(e0.alloc(G * b.entry0_stride * 8, s));
    RSN_TRY(e1.alloc(G * b.entry1_stride * 8, s));
    RSN_TRY(vis.alloc(G * b.vis_stride * 2, s));
    RSN_TRY(bb.alloc(G * b.bb_stride * 8, s));
    RSN_TRY(bo.alloc(G * b.bb_stride * 8, s));
    RSN_TRY(outn.alloc(G * 8, s));
    RSN_TRY(outp.alloc(G * 8, s));
    b.files = files.as<LzFile>();
    b.tile_cnt = tcnt.as<uint64_t>();
    b.tile_off = toff.as<uint64_t>();
    b.enc = enc.as<uint8_t>();
    b.packed = packed.as<uint32_t>();
    b.E0 = E0.as<uint16_t>();
    b.T1 = T1.as<uint16_t>();
    b.entry0 = e0.as<uint64_t>();
    b.entry1 = e1.as<uint64_t>();
    b.vis = vis.as<uint16_t>();
    b.bb = bb.as<uint64_t>();
    b.bo = bo.as<uint64_t>();
    b.out_n = outn.as<uint64_t>();
    RSN_CUDA(cudaMemcpyAsync(files.p, h_files.data(), G * sizeof(LzFile), cudaMemcpyHostToDevice, s));

    const unsigned g = (unsigned)G;
    RSN_LAUNCH(kb_escape_count, dim3((unsigned)tiles_cap, g), kTileThreads, 0, s, b);
    RSN_LAUNCH(kb_escape_finish, g, 256, 0, s, b);
    RSN_LAUNCH(kb_escape_apply, dim3((unsigned)tiles_cap, g), kTileThreads, 0, s, b);
    {
        // the match search on the low-priority stream (see Ctx::low_stream), between two events
        Ctx &c = ctx();
        static const bool split = [] {
            const char *v = getenv("RSN_K2_LOW_PRIORITY");
            return !(v && v[0] == '0');
        }();
        if (split && c.low_stream && (s == c.own_stream || s == c.batch_stream)) {
            RSN_CUDA(cudaEventRecord(c.low_before, s));
            RSN_CUDA(cudaStreamWaitEvent(c.low_stream, c.low_before, 0));
            RSN_TRY(lzss_match_tile_batch(b.files, G, ecap, b.window, b.packed, b.packed_stride, c.low_stream));
            RSN_CUDA(cudaEventRecord(c.low_after, c.low_stream));
            RSN_CUDA(cudaStreamWaitEvent(s, c.low_after, 0));
        } else {
            RSN_TRY(lzss_match_tile_batch(b.files, G, ecap, b.window, b.packed, b.packed_stride, s));
        }
    }
    RSN_LAUNCH(kb_parse_exits, dim3((unsigned)div_up(blocks_cap, kPlanWarps), g), kPlanWarps * 32, 0, s, b);
    if (b.top)
        RSN_LAUNCH(kb_parse_up, dim3((unsigned)div_up((size_t)window + 1, 256), (unsigned)regions1, g), 256, 0, s, b);
    RSN_LAUNCH(kb_parse_top, (unsigned)div_up(G, 64), 64, 0, s, b, G);
    if (b.top) RSN_LAUNCH(kb_parse_down, dim3((unsigned)div_up(regions1, 128), g), 128, 0, s, b);
    RSN_LAUNCH(kb_emit_plan, dim3((unsigned)div_up(blocks_cap, kPlanWarps), g), kPlanWarps * 32, 0, s, b);
    RSN_LAUNCH(kb_emit_finish, g, 256, 0, s, b);
    HostVec<uint64_t> h_outn(G);
    if (!h_outn.data()) return RSN_ERR_NOMEM;
    RSN_CUDA(cudaMemcpyAsync(h_outn.data(), outn.p, G * 8, cudaMemcpyDeviceToHost, s));
    RSN_CUDA(stream_wait(s));

    // one result buffer for the group, files at 256-byte aligned offsets
    HostVec<uint8_t *> h_outp(G);
    if (!h_outp.data()) return RSN_ERR_NOMEM;
    size_t total = 0;
    for (size_t f = 0; f < G; f++) total += round_up(h_outn[f] + 16, 256);
    DevBuf res;
    RSN_TRY(res.alloc_out(total + 256, s));
    size_t off = 0;
    for (size_t f = 0; f < G; f++) {
        h_outp[f] = res.as<uint8_t>() + off;
        out.ptr[f] = h_outp[f];
        out.n[f] = h_outn[f];
        off += round_up(h_outn[f] + 16, 256);
    }
    RSN_CUDA(cudaMemcpyAsync(outp.p, h_outp.data(), G * 8, cudaMemcpyHostToDevice, s));
    RSN_LAUNCH(kb_emit_write, dim3((unsigned)blocks_cap, g), kPT, 0, s, b, outp.as<uint8_t *>());
    RSN_CUDA(stream_wait(s));  // h_outp is read by the copy above
    out.spans.push_back({res.as<uint8_t>(), res.bytes});
    out.owned.push_back(res.release());
    return RSN_OK;
}

// ============================================================================= one stream over several GPUs
//
// BASELINE configs[4]: a single large stream whose match search is sharded by position range.  One
// host thread and stream per shard (shard g runs on device g mod the device count, so the whole
// scheme can be exercised on one GPU):
//   1. every shard uploads and escapes its range of the INPUT; the escaped lengths give the global
//      escaped offsets;
//   2. the escaped stream is re-cut into ranges of whole 4096-position parse blocks, and every shard
//      assembles its range plus a window of bytes before it (the search buffer of lzss.go:123-129) and
//      a window after it (a match is at most W long) — peer copies from whichever shards hold them;
//      with no escapes that is just the two halos, 2 x W bytes per shard;
//   3. match search over the slice; exit tables of the owned range; the table of where the orbit of
//      lzss.go:134-151 leaves the shard for each of the J + 1 possible entry offsets (8 KB) goes to the
//      host, which composes them into every shard's true entry;
//   4. every shard resolves its orbit from its entry, emits locally and copies its part of the output
//      straight into the (pinned) result buffer at its offset.
// No match array leaves its GPU; GPU-to-GPU traffic is the slices' halos.
namespace {

class Barrier {
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const uint64_t gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }

  private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    uint64_t gen_ = 0;
};

struct Sharded {
    int shards = 0, ndev = 1;
    const uint8_t *in = nullptr;
    size_t n = 0;
    int64_t window = 0;
    std::vector<size_t> a;                  // input range of shard g: [a[g], a[g+1])
    std::vector<const uint8_t *> enc;       // its escaped bytes (device memory of device g mod ndev)
    std::vector<size_t> A;                  // global escaped offset of those bytes: [A[g], A[g+1])
    size_t en = 0;
    uint32_t W = 0;
    std::vector<size_t> S;                  // escaped range OWNED by shard g: [S[g], S[g+1]), whole parse blocks
    std::vector<HostVec<uint16_t>> exits;   // [g][0..J]
    std::vector<size_t> entry;              // where the orbit enters shard g, relative to S[g]
    std::vector<uint64_t> total, O;         // output bytes and output offset of shard g
    uint8_t *h_out = nullptr;
    std::atomic<int> rc{RSN_OK};
    std::atomic<uint64_t> peer_bytes{0};
    Barrier bar;
    explicit Sharded(int k) : bar(k) {}
    void fail(int code) {
        int expect = RSN_OK;
        if (code != RSN_OK) rc.compare_exchange_strong(expect, code);
    }
    bool ok() const { return rc.load() == RSN_OK; }
};

thread_local uint64_t g_last_peer_bytes = 0;

int shard_upload_escape(Sharded &sh, int g, DevBuf &d_in, DevBuf &enc_buf, cudaStream_t s) {
    const size_t len = sh.a[g + 1] - sh.a[g];
    RSN_TRY(d_in.alloc(len + 64, s));
    if (len) RSN_CUDA(cudaMemcpyAsync(d_in.p, sh.in + sh.a[g], len, cudaMemcpyHostToDevice, s));
    const uint8_t *e = nullptr;
    size_t en = 0;
    RSN_TRY(lzss_escape(d_in.as<uint8_t>(), len, enc_buf, &e, &en, s));
    RSN_CUDA(stream_wait(s));
    sh.enc[g] = e;
    sh.A[g + 1] = en;  // lengths for now; prefix-summed by shard 0
    return RSN_OK;
}

int shard_assemble(Sharded &sh, int g, size_t lo, size_t hi, DevBuf &slice, cudaStream_t s) {
    RSN_TRY(slice.alloc(hi - lo + 64, s));
    const int mydev = g % sh.ndev;
    for (int h = 0; h < sh.shards; h++) {
        const size_t x = std::max(lo, sh.A[h]), y = std::min(hi, sh.A[h + 1]);
        if (x >= y) continue;
        uint8_t *dst = slice.as<uint8_t>() + (x - lo);
        const uint8_t *src = sh.enc[h] + (x - sh.A[h]);
        const int hdev = h % sh.ndev;
        if (hdev == mydev) {
            RSN_CUDA(cudaMemcpyAsync(dst, src, y - x, cudaMemcpyDeviceToDevice, s));
        } else {
            RSN_CUDA(cudaMemcpyPeerAsync(dst, mydev, src, hdev, y - x, s));
            sh.peer_bytes.fetch_add(y - x);
        }
    }
    RSN_CUDA(stream_wait(s));  // the sources belong to other threads' arenas: done before the next barrier
    return RSN_OK;
}

void shard_thread(Sharded &sh, int g) {
    const int dev = g % sh.ndev;
    int rc = rsn_init(dev);
    sh.fail(rc);
    cudaStream_t s = rc == RSN_OK ? ctx().own_stream : nullptr;
    if (rc == RSN_OK && sh.ndev > 1) {
        for (int d = 0; d < sh.ndev; d++)
            if (d != dev && cudaDeviceEnablePeerAccess(d, 0) != cudaSuccess) cudaGetLastError();  // already on / unsupported: copies still work
    }
    ArenaScope scope(s);
    DevBuf d_in, enc_buf, slice, packed, dex, d_out;
    ParsePlan pp;
    // ---- 1. upload, escape, global escaped offsets
    if (sh.ok()) sh.fail(shard_upload_escape(sh, g, d_in, enc_buf, s));
    sh.bar.wait();
    if (g == 0 && sh.ok()) {
        for (int k = 0; k < sh.shards; k++) sh.A[k + 1] += sh.A[k];
        sh.en = sh.A[sh.shards];
        if (sh.en) sh.fail(lzss_effective_window(sh.window, sh.en, &sh.W));
        const size_t per = div_up(div_up(sh.en, (size_t)sh.shards), kPB) * kPB;
        for (int k = 0; k <= sh.shards; k++) sh.S[k] = std::min(sh.en, (size_t)k * per);
    }
    sh.bar.wait();
    // ---- 2. the shard's slice of the escaped stream: a window before, the owned range, a window after
    const uint32_t W = sh.W, J = W;
    const size_t n_own = sh.S[g + 1] - sh.S[g];
    const size_t W4 = ((size_t)W + 3) & ~(size_t)3;  // keeps the owned records 16-byte aligned
    const size_t lo = sh.S[g] > W4 ? sh.S[g] - W4 : 0, hi = std::min(sh.en, sh.S[g + 1] + W);
    const size_t own_at = sh.S[g] - lo;
    if (sh.ok() && n_own) sh.fail(shard_assemble(sh, g, lo, hi, slice, s));
    sh.bar.wait();
    // ---- 3. match search, exit tables, exits for every entry offset
    const ParseCfg cfg{W, J, RSN_LZSS_ASYNC, nullptr};
    const uint32_t *d_lo = nullptr;
    const uint8_t *d_enc = nullptr;
    auto step3 = [&]() -> int {
        HostVec<uint16_t> &ex = sh.exits[g];
        if (!ex.resize((size_t)J + 1)) return RSN_ERR_NOMEM;
        if (n_own == 0) {  // nothing owned: the orbit passes through
            for (uint32_t r = 0; r <= J; r++) ex[r] = (uint16_t)r;
            return RSN_OK;
        }
        const size_t sn = hi - lo;
        RSN_TRY(packed.alloc(div_up(sn, kPB) * kPB * 4 + 64, s));
        RSN_TRY(lzss_match(slice.as<uint8_t>(), sn, W, packed.as<uint32_t>(), s));
        d_lo = packed.as<uint32_t>() + own_at;
        d_enc = slice.as<uint8_t>() + own_at;
        RSN_TRY(parse_build_tables(cfg, d_lo, n_own, pp, s));
        RSN_TRY(dex.alloc(((size_t)J + 1) * 2, s));
        RSN_LAUNCH(k_shard_exits, (unsigned)div_up((size_t)J + 1, 256), 256, 0, s, pp.E0.as<uint16_t>(),
                   pp.T[pp.lv.top].as<uint16_t>(), pp.lv.top, pp.lv.rsize[pp.lv.top], J, n_own, dex.as<uint16_t>());
        RSN_CUDA(cudaMemcpyAsync(ex.data(), dex.p, ((size_t)J + 1) * 2, cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
        return RSN_OK;
    };
    if (sh.ok()) sh.fail(step3());
    sh.bar.wait();
    if (g == 0 && sh.ok()) {  // compose: the exit of one shard is the entry of the next
        size_t e = 0;
        for (int k = 0; k < sh.shards; k++) {
            sh.entry[k] = e;
            e = sh.exits[k][e];
        }
    }
    sh.bar.wait();
    // ---- 4. the orbit from the true entry, sizes, emit into the result at the shard's offset
    if (sh.ok() && n_own) sh.fail(parse_resolve(cfg, d_lo, n_own, sh.entry[g], pp, &sh.total[g], s));
    sh.bar.wait();
    if (g == 0 && sh.ok()) {
        uint64_t at = 0;
        for (int k = 0; k < sh.shards; k++) {
            sh.O[k] = at;
            at += sh.total[k];
        }
        sh.O[sh.shards] = at;
        sh.h_out = (uint8_t *)host_out_alloc(at ? at : 1);
        if (!sh.h_out) sh.fail(RSN_ERR_NOMEM);
    }
    sh.bar.wait();
    auto step4 = [&]() -> int {
        if (!n_own || !sh.total[g]) return RSN_OK;
        RSN_TRY(d_out.alloc(sh.total[g] + 16, s));
        RSN_TRY(parse_emit(cfg, d_enc, d_lo, n_own, pp, d_out.as<uint8_t>(), s));
        RSN_CUDA(cudaMemcpyAsync(sh.h_out + sh.O[g], d_out.p, sh.total[g], cudaMemcpyDeviceToHost, s));
        RSN_CUDA(stream_wait(s));
        return RSN_OK;
    };
    if (sh.ok()) sh.fail(step4());
    if (s) stream_wait(s);
    sh.bar.wait();  // nobody rewinds its arena while another shard may still read from it
}

}  // namespace

uint64_t lzss_sharded_last_peer_bytes() { return g_last_peer_bytes; }

// lz.CompressAsync (lzss.go:109-154) of one host buffer over `shards` shards (see above).
int lzss_compress_sharded(const uint8_t *in, size_t n, int64_t window, int shards, uint8_t **out, size_t *out_n) {
    int ndev = 0;
    RSN_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return RSN_ERR_NO_DEVICE;
    if (shards < 1 || shards > 64) return RSN_ERR_INVALID_ARG;
    Sharded sh(shards);
    sh.shards = shards;
    sh.ndev = ndev;
    sh.in = in;
    sh.n = n;
    sh.window = window;
    sh.a.resize(shards + 1);
    for (int g = 0; g <= shards; g++) sh.a[g] = (size_t)((unsigned __int128)n * g / shards);
    sh.enc.assign(shards, nullptr);
    sh.A.assign(shards + 1, 0);
    sh.S.assign(shards + 1, 0);
    sh.exits = std::vector<HostVec<uint16_t>>(shards);
    sh.entry.assign(shards, 0);
    sh.total.assign(shards, 0);
    sh.O.assign(shards + 1, 0);
    std::vector<std::thread> th;
    for (int g = 1; g < shards; g++) th.emplace_back(shard_thread, std::ref(sh), g);
    {
        // shard 0 runs on the calling thread; its context goes back to the device it had
        const int prev = ctx().ready ? ctx().device : -1;
        shard_thread(sh, 0);
        if (prev >= 0 && prev != ctx().device) rsn_init(prev);
    }
    for (auto &t : th) t.join();
    g_last_peer_bytes = sh.peer_bytes.load();
    if (!sh.ok()) {
        if (sh.h_out) rsn_free(sh.h_out);
        return sh.rc.load();
    }
    *out = sh.h_out;
    *out_n = (size_t)sh.O[shards];
    return RSN_OK;
}

}  // namespace rsn
