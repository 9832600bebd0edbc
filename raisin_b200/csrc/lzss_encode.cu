// lzss_encode.cu — LZSS compress (variant A = lz.CompressAsync, lzss.go:109-184).
//
// Pipeline (all on device):
//   K1  escape-expand            EncodeOpeningSymbols, lzss.go:369-389
//   K2  per-position longest match (every compressorWorker, lzss.go:156-184, in parallel)
//   K3  greedy-parse reconstruction: the reference's sequential merge loop (lzss.go:134-151)
//       visits i, i+max(L,1), ...; we rebuild that chain with per-block exit tables composed
//       through a 64-ary hierarchy, so no pass is sequential in n
//   K4  token sizing + emit      getEncoding, lzss.go:318-320 and the `<` rule at lzss.go:143
#include "common.cuh"
#include "lzss.cuh"

namespace rsn {

// ============================================================================= K1 escape

__device__ __forceinline__ bool is_special(uint8_t v) { return v == 0x5C || v == 0xFF; }

__global__ void __launch_bounds__(kTileThreads) k_escape_count(const uint8_t *__restrict__ in, size_t n,
                                                               uint64_t *__restrict__ tile_cnt) {
    __shared__ uint32_t sm[33];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    uint32_t cnt = 0;
    if (base < n) {
        uint8_t v[16];
        load16(in, base, n, 0, v);
        const int valid = (int)min((size_t)16, n - base);
#pragma unroll
        for (int k = 0; k < 16; k++) cnt += (k < valid) ? 1u + (is_special(v[k]) ? 1u : 0u) : 0u;
    }
    uint32_t total;
    block_exclusive_sum<uint32_t>(cnt, sm, total);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kTileThreads) k_escape_apply(const uint8_t *__restrict__ in, size_t n,
                                                               const uint64_t *__restrict__ tile_off,
                                                               uint8_t *__restrict__ out) {
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

int lzss_escape(const uint8_t *d_in, size_t n, DevBuf &enc, size_t *enc_n, cudaStream_t s) {
    if (n == 0) {
        RSN_TRY(enc.alloc(16, s));
        *enc_n = 0;
        return RSN_OK;
    }
    const size_t tiles = div_up(n, kTile);
    DevBuf cnt, off;
    RSN_TRY(cnt.alloc(tiles * 8, s));
    RSN_TRY(off.alloc((tiles + 1) * 8, s));
    RSN_LAUNCH(k_escape_count, (unsigned)tiles, kTileThreads, 0, s, d_in, n, cnt.as<uint64_t>());
    RSN_TRY(spine_scan_u64(cnt.as<uint64_t>(), off.as<uint64_t>(), off.as<uint64_t>() + tiles, tiles, s));
    uint64_t total = 0;
    RSN_TRY(read_u64(off.as<uint64_t>() + tiles, &total, s));
    RSN_TRY(enc.alloc(total + 64, s));
    RSN_LAUNCH(k_escape_apply, (unsigned)tiles, kTileThreads, 0, s, d_in, n, off.as<uint64_t>(), enc.as<uint8_t>());
    *enc_n = (size_t)total;
    return RSN_OK;
}

// ============================================================================= K3 parse

constexpr int kPB = 4096;   // parse block (positions)
constexpr int kPS = 64;     // sub-block handled by one thread
constexpr int kPT = kPB / kPS;  // 64 threads per CTA
constexpr int kFan = 64;    // hierarchy fan-out
static_assert(kPT == kPS, "k_parse_exits composes one sub-block per step with one thread per element");

// jump[p] = max(L,1) for the block's positions into shared memory; positions >= n get 1.
__device__ __forceinline__ void load_jumps(const uint32_t *__restrict__ lo, size_t start, size_t n, uint16_t *jump) {
    for (int p = threadIdx.x; p < kPB; p += blockDim.x) {
        const size_t g = start + p;
        uint32_t L = g < n ? (__ldg(lo + g) >> 16) : 1u;
        jump[p] = (uint16_t)(L ? L : 1u);
    }
}

// x[p] = first chain position (block-relative) at or beyond the end of p's sub-block, or beyond n.
__device__ __forceinline__ void sub_exits(const uint16_t *jump, uint16_t *x, uint32_t nrel) {
    const int s = threadIdx.x;  // one thread per sub-block
    const uint32_t lo_p = s * kPS, hi_p = lo_p + kPS;
    for (int p = (int)hi_p - 1; p >= (int)lo_p; p--) {
        uint32_t t = (uint32_t)p + jump[p];
        x[p] = (uint16_t)((t >= hi_p || t >= nrel) ? t : x[t]);
    }
}

// E0[g] = (first chain position at or beyond the block end) - block end, for the chain from g.
__global__ void __launch_bounds__(kPT) k_parse_exits(const uint32_t *__restrict__ lo, size_t n,
                                                     uint16_t *__restrict__ E0) {
    __shared__ uint16_t jump[kPB];
    __shared__ uint16_t x[kPB];
    const size_t start = (size_t)blockIdx.x * kPB;
    const uint32_t nrel = (uint32_t)min((size_t)kPB, n - start);
    load_jumps(lo, start, n, jump);
    __syncthreads();
    sub_exits(jump, x, nrel);
    __syncthreads();
    // compose sub-block exits right to left: later sub-blocks are already final
    for (int s = kPT - 1; s >= 0; s--) {
        const int p = s * kPS + threadIdx.x;
        uint32_t v = x[p];
        if (v < (uint32_t)kPB && v < nrel) v = x[v];  // v lies in a later, already final sub-block
        x[p] = (uint16_t)v;
        __syncthreads();
    }
    for (int p = threadIdx.x; p < (int)nrel; p += blockDim.x) {
        uint32_t v = x[p];
        E0[start + p] = (uint16_t)(v >= (uint32_t)kPB ? v - kPB : 0u);
    }
}

struct ParseLevels {
    int top;                 // highest level; level l regions have size kPB * kFan^l
    size_t regions[8];       // region count per level
    size_t rsize[8];         // region size per level
};

// one chain step at level `lvl` from absolute position p (p < n, p inside region p / rsize)
__device__ __forceinline__ size_t level_step(size_t p, int lvl, size_t rsize, const uint16_t *__restrict__ E0,
                                             const uint16_t *__restrict__ T, uint32_t W) {
    const size_t r = p / rsize;
    const size_t end = (r + 1) * rsize;
    if (lvl == 0) return end + __ldg(E0 + p);
    return end + __ldg(T + r * W + (p - r * rsize));
}

// T_l[r][rel] for rel in [0, W): follow level l-1 until leaving region r (or the input).
__global__ void k_parse_up(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tprev,
                           uint16_t *__restrict__ Tcur, int lvl, size_t rsize_prev, size_t rsize_cur, uint32_t W,
                           size_t n) {
    const size_t r = blockIdx.y;
    const uint32_t rel = blockIdx.x * blockDim.x + threadIdx.x;
    if (rel >= W) return;
    const size_t start = r * rsize_cur, end = start + rsize_cur;
    size_t p = start + rel;
    while (p < end && p < n) p = level_step(p, lvl - 1, rsize_prev, E0, Tprev, W);
    Tcur[r * W + rel] = (uint16_t)(p >= end ? p - end : 0);
}

// sequential walk over the (<= kFan) top-level regions
__global__ void k_parse_top(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Ttop, int lvl,
                            size_t rsize, size_t regions, uint32_t W, size_t n, uint64_t *__restrict__ entry) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    size_t p = 0;
    for (size_t r = 0; r < regions; r++) {
        entry[r] = p;
        const size_t end = (r + 1) * rsize;
        if (p < end && p < n) p = level_step(p, lvl, rsize, E0, Ttop, W);
    }
}

// entries of the children (level lvl-1) of each level-lvl region
__global__ void k_parse_down(const uint16_t *__restrict__ E0, const uint16_t *__restrict__ Tchild, int child_lvl,
                             size_t rsize_child, size_t regions_parent, size_t regions_child, uint32_t W, size_t n,
                             const uint64_t *__restrict__ entry_parent, uint64_t *__restrict__ entry_child) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= regions_parent) return;
    size_t p = entry_parent[r];
    for (int c = 0; c < kFan; c++) {
        const size_t cr = r * kFan + c;
        if (cr >= regions_child) break;
        entry_child[cr] = p;
        const size_t end = (cr + 1) * rsize_child;
        if (p < end && p < n) p = level_step(p, child_lvl, rsize_child, E0, Tchild, W);
    }
}

// ============================================================================= K4 emit

__device__ __forceinline__ uint32_t token_size(uint32_t packed) {
    const uint32_t L = packed >> 16, off = packed & 0xFFFFu;
    if (L == 0) return 1;
    const uint32_t tl = 3 + ndig_u32(off) + ndig_u32(L);
    return tl < L ? tl : L;  // strict '<' (lzss.go:143)
}

__device__ __forceinline__ uint8_t *put_dec(uint8_t *o, uint32_t v) {
    const int d = ndig_u32(v);
    for (int k = d - 1; k >= 0; k--) {
        o[k] = (uint8_t)('0' + v % 10);
        v /= 10;
    }
    return o + d;
}

template <bool WRITE>
__global__ void __launch_bounds__(kPT) k_emit(const uint8_t *__restrict__ enc, const uint32_t *__restrict__ lo,
                                              size_t n, const uint64_t *__restrict__ entry0,
                                              uint64_t *__restrict__ blk_bytes, const uint64_t *__restrict__ blk_off,
                                              uint8_t *__restrict__ out) {
    __shared__ uint16_t jump[kPB];
    __shared__ uint16_t x[kPB];
    __shared__ uint32_t sub_entry[kPT];
    __shared__ uint32_t sm[33];
    const size_t start = (size_t)blockIdx.x * kPB;
    const uint32_t nrel = (uint32_t)min((size_t)kPB, n - start);
    load_jumps(lo, start, n, jump);
    __syncthreads();
    sub_exits(jump, x, nrel);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t e0 = entry0[blockIdx.x];
        uint32_t e = (e0 - start) > 0xFFFFull ? 0xFFFFu : (uint32_t)(e0 - start);  // may lie beyond this block
        for (int s = 0; s < kPT; s++) {
            sub_entry[s] = e;
            if (e < (uint32_t)(s + 1) * kPS && e < nrel) e = x[e];
        }
    }
    __syncthreads();
    const uint32_t hi = (threadIdx.x + 1) * kPS;
    uint32_t bytes = 0;
    for (uint32_t p = sub_entry[threadIdx.x]; p < hi && p < nrel; p += jump[p]) bytes += token_size(__ldg(lo + start + p));
    uint32_t total;
    uint32_t pre = block_exclusive_sum<uint32_t>(bytes, sm, total);
    if (!WRITE) {
        if (threadIdx.x == 0) blk_bytes[blockIdx.x] = total;
        return;
    }
    uint8_t *o = out + blk_off[blockIdx.x] + pre;
    for (uint32_t p = sub_entry[threadIdx.x]; p < hi && p < nrel; p += jump[p]) {
        const size_t g = start + p;
        const uint32_t packed = __ldg(lo + g);
        const uint32_t L = packed >> 16, off = packed & 0xFFFFu;
        if (L == 0) {
            *o++ = enc[g];
        } else {
            const uint32_t tl = 3 + ndig_u32(off) + ndig_u32(L);
            if (tl < L) {
                *o++ = '<';
                o = put_dec(o, off);
                *o++ = ',';
                o = put_dec(o, L);
                *o++ = '>';
            } else {
                for (uint32_t k = 0; k < L; k++) o[k] = enc[g + k];
                o += L;
            }
        }
    }
}

// ============================================================================= host orchestration

static int parse_and_emit(const uint8_t *d_enc, size_t n, uint32_t W, const uint32_t *d_lo, uint8_t **d_out,
                          size_t *out_n, cudaStream_t s) {
    const size_t blocks = div_up(n, kPB);
    // level geometry
    ParseLevels lv;
    lv.top = 0;
    lv.rsize[0] = kPB;
    lv.regions[0] = blocks;
    while (lv.regions[lv.top] > (size_t)kFan) {
        lv.rsize[lv.top + 1] = lv.rsize[lv.top] * kFan;
        lv.regions[lv.top + 1] = div_up(n, lv.rsize[lv.top + 1]);
        lv.top++;
    }
    DevBuf E0, T[8], entry[8];
    RSN_TRY(E0.alloc(n * 2 + 16, s));
    RSN_LAUNCH(k_parse_exits, (unsigned)blocks, kPT, 0, s, d_lo, n, E0.as<uint16_t>());
    for (int l = 1; l <= lv.top; l++) {
        RSN_TRY(T[l].alloc(lv.regions[l] * W * 2, s));
        dim3 grid((unsigned)div_up(W, 256), (unsigned)lv.regions[l]);
        RSN_LAUNCH(k_parse_up, grid, 256, 0, s, E0.as<uint16_t>(), T[l - 1].as<uint16_t>(), T[l].as<uint16_t>(), l,
                   lv.rsize[l - 1], lv.rsize[l], W, n);
    }
    for (int l = 0; l <= lv.top; l++) RSN_TRY(entry[l].alloc(lv.regions[l] * 8, s));
    RSN_LAUNCH(k_parse_top, 1, 32, 0, s, E0.as<uint16_t>(), T[lv.top].as<uint16_t>(), lv.top, lv.rsize[lv.top],
               lv.regions[lv.top], W, n, entry[lv.top].as<uint64_t>());
    for (int l = lv.top; l >= 1; l--) {
        RSN_LAUNCH(k_parse_down, (unsigned)div_up(lv.regions[l], 128), 128, 0, s, E0.as<uint16_t>(),
                   T[l - 1].as<uint16_t>(), l - 1, lv.rsize[l - 1], lv.regions[l], lv.regions[l - 1], W, n,
                   entry[l].as<uint64_t>(), entry[l - 1].as<uint64_t>());
    }
    DevBuf bb, bo;
    RSN_TRY(bb.alloc(blocks * 8, s));
    RSN_TRY(bo.alloc((blocks + 1) * 8, s));
    RSN_LAUNCH(k_emit<false>, (unsigned)blocks, kPT, 0, s, d_enc, d_lo, n, entry[0].as<uint64_t>(), bb.as<uint64_t>(),
               (const uint64_t *)nullptr, (uint8_t *)nullptr);
    RSN_TRY(spine_scan_u64(bb.as<uint64_t>(), bo.as<uint64_t>(), bo.as<uint64_t>() + blocks, blocks, s));
    uint64_t total = 0;
    RSN_TRY(read_u64(bo.as<uint64_t>() + blocks, &total, s));
    DevBuf out;
    RSN_TRY(out.alloc(total + 16, s));
    RSN_LAUNCH(k_emit<true>, (unsigned)blocks, kPT, 0, s, d_enc, d_lo, n, entry[0].as<uint64_t>(),
               (uint64_t *)nullptr, bo.as<uint64_t>(), out.as<uint8_t>());
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

int lzss_compress_dev(const uint8_t *d_in, size_t n, int64_t window, int variant, uint8_t **d_out, size_t *out_n,
                      cudaStream_t s) {
    if (variant != RSN_LZSS_ASYNC) return RSN_ERR_UNSUPPORTED;  // variant B: planned
    DevBuf enc;
    size_t en = 0;
    RSN_TRY(lzss_escape(d_in, n, enc, &en, s));
    if (en == 0) {
        DevBuf out;
        RSN_TRY(out.alloc(16, s));
        *d_out = (uint8_t *)out.release();
        *out_n = 0;
        return RSN_OK;
    }
    uint32_t W = 0;
    RSN_TRY(lzss_effective_window(window, en, &W));
    DevBuf lo;
    RSN_TRY(lo.alloc(en * 4 + 16, s));
    RSN_TRY(lzss_match(enc.as<uint8_t>(), en, W, lo.as<uint32_t>(), s));
    return parse_and_emit(enc.as<uint8_t>(), en, W, lo.as<uint32_t>(), d_out, out_n, s);
}

}  // namespace rsn
