// lzss_match.cu — K2: every position's longest match, in parallel.
//
// Semantics (compressorWorker, lzss.go:166-184, over the window of lzss.go:123-129):
//   win(i) = enc[max(0,i-W) : i]   (W bytes once i > W; exactly i bytes before that)
//   L(i)   = largest k with i+k <= n and enc[i:i+k] occurring wholly inside win(i)
//   off(i) = len(win) - (leftmost start of enc[i:i+L] in win)
// In distance form: L(i) = max over d in [1, min(i,W)] of min(lcp(i-d, i), d, n-i) and off(i)
// is the LARGEST d attaining it (leftmost source).
#include "lzss.cuh"

namespace rsn {

// v0: one thread per position, distances scanned from far to near so that ties keep the
// larger distance and the scan can stop once d <= best (a candidate at distance d yields <= d).
__global__ void __launch_bounds__(256) k_match_v0(const uint8_t *__restrict__ enc, size_t n, uint32_t W,
                                                  uint32_t *__restrict__ packed) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t dmax = (uint32_t)min((size_t)W, i);
    const uint32_t room = (uint32_t)min((size_t)W, n - i);  // L <= min(d, n-i) <= W
    const uint8_t c0 = enc[i];
    uint32_t best = 0, boff = 0;
    for (uint32_t d = dmax; d >= 1 && d > best; d--) {
        const size_t j = i - d;
        if (__ldg(enc + j) != c0) continue;
        const uint32_t cap = min(d, room);
        if (best >= cap) continue;                                    // cannot improve
        if (best && __ldg(enc + j + best) != __ldg(enc + i + best)) continue;  // must beat `best`
        uint32_t l = 1;
        while (l < cap && __ldg(enc + j + l) == __ldg(enc + i + l)) l++;
        if (l > best) {
            best = l;
            boff = d;
        }
    }
    packed[i] = (best << 16) | boff;
}

int lzss_match(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, cudaStream_t s) {
    if (n == 0) return RSN_OK;
    if (W < 1 || W > kMaxWindow) return RSN_ERR_INVALID_ARG;
    RSN_LAUNCH(k_match_v0, (unsigned)div_up(n, 256), 256, 0, s, d_enc, n, W, d_packed);
    return RSN_OK;
}

}  // namespace rsn
