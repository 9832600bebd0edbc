// lzss.cuh — internal interfaces of the LZSS kernels.
#pragma once
#include "common.cuh"

namespace rsn {

// Largest search window the device path supports (packed (len<<16)|off and u16 parse tables).
constexpr uint32_t kMaxWindow = 32768;

// K2: packed[i] = (len << 16) | off of compressorWorker (lzss.go:166-184) for every position
// of the escaped buffer.  1 <= W <= kMaxWindow.
int lzss_match(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, cudaStream_t s);

int lzss_match_tile(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, cudaStream_t s);
int lzss_match_tile_range(const uint8_t *d_enc, size_t n, uint32_t W, uint32_t *d_packed, size_t tile_lo,
                          size_t tile_hi, bool run_fix0, cudaStream_t s);
size_t lzss_match_tile_size();

// One file of a batched compress (device memory): in/n come from the host, the rest is filled in by
// the kernels of the stage.
struct LzFile {
    const uint8_t *in;
    uint64_t n;
    const uint8_t *enc;  // escaped bytes: `in` itself when nothing had to change
    uint64_t en;
    uint32_t W;          // effective window min(window, en)
    uint32_t touched;
    uint64_t out_n;
};
int lzss_match_tile_batch(const LzFile *d_files, size_t G, size_t ecap, uint32_t window, uint32_t *d_packed,
                          size_t packed_stride, cudaStream_t s);
int lzss_effective_window(int64_t window, size_t enc_n, uint32_t *W);
int lzss_escape(const uint8_t *d_in, size_t n, DevBuf &enc, const uint8_t **enc_ptr, size_t *enc_n, cudaStream_t s);
int lzss_compress_dev(const uint8_t *d_in, size_t n, int64_t window, int variant, uint8_t **d_out, size_t *out_n,
                      cudaStream_t s);
// `spec_packed`: match arrays already computed over d_in itself (valid iff nothing needed escaping)
int lzss_compress_dev_ex(const uint8_t *d_in, size_t n, int64_t window, int variant, uint32_t *spec_packed,
                         uint8_t **d_out, size_t *out_n, cudaStream_t s);
int lzss_emit_dev(const uint8_t *d_enc, size_t n, int64_t window, int variant, const uint32_t *d_packed,
                  uint8_t **d_out, size_t *out_n, cudaStream_t s);
int lzss_compress_sharded(const uint8_t *in, size_t n, int64_t window, int shards, uint8_t **out, size_t *out_n);
uint64_t lzss_sharded_last_peer_bytes();
int lzss_escape_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s);
int lzss_decompress_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s);

}  // namespace rsn
