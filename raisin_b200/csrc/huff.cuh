// huff.cuh — internal interfaces of the Huffman kernels.
#pragma once
#include "common.cuh"

namespace rsn {

int huff_compress_dev(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, cudaStream_t s);
// h_prefix: host copy of the first h_prefix_n bytes of the stream (header search), may be the
// whole stream.
int huff_decompress_dev(const uint8_t *d_in, size_t n, const uint8_t *h_in_or_null, int strict, uint8_t **d_out,
                        size_t *out_n, cudaStream_t s);

}  // namespace rsn
