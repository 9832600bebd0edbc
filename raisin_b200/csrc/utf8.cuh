// utf8.cuh — Go's UTF-8 semantics (`for _, c := range string(b)`, huffman.go:309,235 and
// string(rune), huffman.go:138) as pure per-position functions.
//
// A byte outside [80,BF] always starts a rune.  A continuation byte at p is consumed iff the
// nearest non-continuation byte q in {p-1,p-2,p-3} (everything between being continuation
// bytes) leads a VALID sequence longer than p-q; otherwise p itself starts a rune (U+FFFD,
// width 1).  So start/rune/width at p depend only on bytes p-3 .. p+3 and the buffer end.
#pragma once
#include <stdint.h>

namespace rsn {

__host__ __device__ __forceinline__ bool utf8_is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }

// Valid width (2..4) of the sequence led by c with following bytes b1,b2,b3 and `avail` bytes
// available from the lead (inclusive); 1 if invalid/ASCII.  *rune receives the decoded value
// (U+FFFD when invalid and c >= 0x80).
__host__ __device__ __forceinline__ int utf8_decode_at(uint8_t c, uint8_t b1, uint8_t b2, uint8_t b3, uint64_t avail,
                                                       int32_t *rune) {
    if (c < 0x80) {
        *rune = c;
        return 1;
    }
    *rune = 0xFFFD;
    int need;
    uint8_t lo = 0x80, hi = 0xBF;
    if (c >= 0xC2 && c <= 0xDF) need = 2;
    else if (c == 0xE0) { need = 3; lo = 0xA0; }
    else if ((c >= 0xE1 && c <= 0xEC) || c == 0xEE || c == 0xEF) need = 3;
    else if (c == 0xED) { need = 3; hi = 0x9F; }
    else if (c == 0xF0) { need = 4; lo = 0x90; }
    else if (c >= 0xF1 && c <= 0xF3) need = 4;
    else if (c == 0xF4) { need = 4; hi = 0x8F; }
    else return 1;
    if (avail < (uint64_t)need) return 1;
    if (b1 < lo || b1 > hi) return 1;
    if (need == 2) {
        *rune = ((int32_t)(c & 0x1F) << 6) | (b1 & 0x3F);
        return 2;
    }
    if (!utf8_is_cont(b2)) return 1;
    if (need == 3) {
        *rune = ((int32_t)(c & 0x0F) << 12) | ((int32_t)(b1 & 0x3F) << 6) | (b2 & 0x3F);
        return 3;
    }
    if (!utf8_is_cont(b3)) return 1;
    *rune = ((int32_t)(c & 0x07) << 18) | ((int32_t)(b1 & 0x3F) << 12) | ((int32_t)(b2 & 0x3F) << 6) | (b3 & 0x3F);
    return 4;
}

// w[] holds bytes base-3 .. base+18 (w[k+3] = byte base+k).  Returns true iff byte base+k
// starts a rune; then *rune is its value.  `pos` = base+k (absolute), n = buffer length.
__host__ __device__ __forceinline__ bool utf8_start_at(const uint8_t *w, int k, uint64_t pos, uint64_t n,
                                                       int32_t *rune) {
    const uint8_t c = w[k + 3];
    if (utf8_is_cont(c)) {
#pragma unroll
        for (int q = 1; q <= 3; q++) {
            if (pos < (uint64_t)q) break;  // nothing before the buffer start
            const uint8_t lead = w[k + 3 - q];
            if (!utf8_is_cont(lead)) {
                int32_t r;
                const int wd = utf8_decode_at(lead, w[k + 4 - q], w[k + 5 - q], w[k + 6 - q], n - (pos - q), &r);
                if (wd > q) return false;  // consumed by that sequence
                break;
            }
        }
        *rune = 0xFFFD;
        return true;
    }
    utf8_decode_at(c, w[k + 4], w[k + 5], w[k + 6], n - pos, rune);
    return true;
}

__host__ __device__ __forceinline__ int utf8_width(int32_t r) {
    if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) return 3;  // U+FFFD
    return r < 0x80 ? 1 : r < 0x800 ? 2 : r < 0x10000 ? 3 : 4;
}

__host__ __device__ __forceinline__ int utf8_encode(int32_t r, uint8_t *o) {
    if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = 0xFFFD;
    if (r < 0x80) {
        o[0] = (uint8_t)r;
        return 1;
    }
    if (r < 0x800) {
        o[0] = (uint8_t)(0xC0 | (r >> 6));
        o[1] = (uint8_t)(0x80 | (r & 0x3F));
        return 2;
    }
    if (r < 0x10000) {
        o[0] = (uint8_t)(0xE0 | (r >> 12));
        o[1] = (uint8_t)(0x80 | ((r >> 6) & 0x3F));
        o[2] = (uint8_t)(0x80 | (r & 0x3F));
        return 3;
    }
    o[0] = (uint8_t)(0xF0 | (r >> 18));
    o[1] = (uint8_t)(0x80 | ((r >> 12) & 0x3F));
    o[2] = (uint8_t)(0x80 | ((r >> 6) & 0x3F));
    o[3] = (uint8_t)(0x80 | (r & 0x3F));
    return 4;
}

}  // namespace rsn
