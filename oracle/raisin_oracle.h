/*
 * raisin_oracle.h — CPU restatement of go-compression/raisin's LZSS + Huffman hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / reported CPU baseline.  The product (libraisin_b200.so) never
 * links or calls this code and has no CPU fallback.
 *
 * PARITY PIN STATUS: the reference is pure Go and no Go toolchain exists in this image, so
 * the reference itself cannot be run here.  The reference's own tests hold no golden bytes
 * (round-trip only).  The oracle is pinned by (1) every size the reference README publishes
 * (README.md:153,157,165,167), (2) the reference's own round-trip tests on its samIAm text
 * (compressor/lz/lzss_test.go:25-47), and (3) agreement with a second, independent,
 * statement-by-statement Python transcription (oracle/go_literal.py).  Exact payload bytes
 * beyond those are therefore "parity unpinned by reference-run outputs"; see DESIGN.md.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference repository root).
 */
#ifndef RAISIN_ORACLE_H
#define RAISIN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (negative) — mirror the reference's panics */
#define RSNO_OK 0
#define RSNO_ERR_NOMEM (-1)
#define RSNO_ERR_EMPTY_INPUT (-2)       /* huffman.Compress on empty input: heap.Pop on empty heap panics */
#define RSNO_ERR_NO_SEPARATOR (-3)      /* huffman decode: no "\\\n" => sections[1] index panic (huffman.go:261-264) */
#define RSNO_ERR_BAD_HEADER (-4)        /* decodeTree index out of range (huffman.go:210) or empty symbol table */
#define RSNO_ERR_TRUNCATED (-5)         /* findCodes reads data[max] (huffman.go:145) / pad > bit count (huffman.go:294) */
#define RSNO_ERR_GUARD (-6)             /* findCodes "Max recursion depth" (huffman.go:132-134), strict mode only */
#define RSNO_ERR_BAD_REFERENCE (-7)     /* lz.Decompress slice out of range (lzss.go:349-350) */
#define RSNO_ERR_SINGLE_LEAF_LOOP (-8)  /* single-leaf tree with bits left: unbounded recursion (huffman.go:139-140) */

void rsno_free(void *p);

/* lzss.go:369-389 EncodeOpeningSymbols / lzss.go:391-406 DecodeOpeningSymbols */
int rsno_escape(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);
int rsno_unescape(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);

/*
 * Per-position result of compressorWorker (lzss.go:166-184) over the ESCAPED buffer enc:
 * len_out[i] = size of the longest pattern enc[i:i+k] found by bytes.Index in the window
 * enc[max(0,i-W):i] (W<=0: unbounded), 0 if even one byte is not found;
 * off_out[i] = len(window) - leftmost index of that pattern (0 when len is 0).
 * mode 1: literal (the pattern grows one byte per level and every level re-runs the leftmost
 * search from the window start, as the reference's recursion does); mode 0: same result by
 * monotonicity (doubling + bisection on the length, then one leftmost search).
 * threads >= 1 splits positions over pthreads (the goroutine-per-byte analogue).
 */
int rsno_lzss_match_arrays(const uint8_t *enc, size_t n, int64_t window, int mode, int threads,
                           uint32_t *len_out, uint32_t *off_out);

/* lzss.go:109-154 CompressAsync (variant A, the engine path). */
int rsno_lzss_compress_async(const uint8_t *in, size_t n, int64_t window, int mode, int threads,
                             uint8_t **out, size_t *out_n);
/* lzss.go:224-316 Compress (variant B, exported iterative function). */
int rsno_lzss_compress_iter(const uint8_t *in, size_t n, int64_t window, uint8_t **out, size_t *out_n);
/* lzss.go:323-364 Decompress. */
int rsno_lzss_decompress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);

/* Go `for _, c := range string(b)` (huffman.go:309): runes_out needs room for n entries. */
size_t rsno_utf8_decode(const uint8_t *in, size_t n, int32_t *runes_out);

/*
 * huffman.go:299-325 Compress.  Header records are written in ascending rune order, and if
 * the last record would be the rune 0x5C and there are >= 2 records the last two are swapped
 * (Go's own order is map-iteration order, i.e. unspecified).
 */
int rsno_huff_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);
/* huffman.go:327-330, 258-297 Decompress.  strict != 0 reproduces the 900000-bit guard. */
int rsno_huff_decompress(const uint8_t *in, size_t n, int strict, uint8_t **out, size_t *out_n);

/*
 * Code table of the tree buildTree (huffman.go:58-103) makes for the given (rune,freq) set:
 * for leaf k (input order) code bits (MSB-first in the low `len` bits of code_out[k]) and len.
 * Returns RSNO_OK, or RSNO_ERR_EMPTY_INPUT for k == 0.  Codes longer than 64 bits => -1.
 */
int rsno_huff_code_table(const int32_t *runes, const uint64_t *freqs, size_t k,
                         uint64_t *code_out, uint8_t *len_out);

#ifdef __cplusplus
}
#endif
#endif
