/*
 * raisin_b200.h — C ABI of libraisin_b200.so: the B200 (sm_100a) implementation of
 * go-compression/raisin's LZSS + Huffman hot path.
 *
 * These entry points are what a cgo shim inside the reference's compressor/lz and
 * compressor/huffman packages binds (see INTEGRATION.md and go/).  Each one names the
 * reference function whose body it replaces.  Plain pointers and sizes only; no torch
 * types.  All functions are re-entrant from multiple OS threads (the reference's
 * BenchmarkSuite calls the codecs from several goroutines at once, engine/engine.go:235-244);
 * the library keeps no global mutable codec state (deliberately unlike
 * compressor/huffman/huffman.go:56,129).
 *
 * There is NO CPU fallback: every codec call runs on the CUDA device or fails with
 * RSN_ERR_CUDA / RSN_ERR_NO_DEVICE.
 *
 * Error convention: 0 on success, negative code on failure.  The reference signals the
 * same conditions by panicking; a shim turns rc != 0 into panic(rsn_strerror(rc)) so that
 * engine.AsyncBenchmarkFile's recover() (engine/engine.go:315-328) still reports "Failed".
 */
#ifndef RAISIN_B200_H
#define RAISIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSN_OK 0
#define RSN_ERR_CUDA (-1)              /* a CUDA runtime call failed; see rsn_last_cuda_error() */
#define RSN_ERR_NO_DEVICE (-2)         /* no usable CUDA device */
#define RSN_ERR_NOMEM (-3)
#define RSN_ERR_INVALID_ARG (-4)
#define RSN_ERR_UNSUPPORTED (-5)       /* valid for the reference, outside this build's limits (documented) */
#define RSN_ERR_EMPTY_INPUT (-10)      /* huffman.Compress(empty): heap.Pop on empty heap panics (huffman.go:102) */
#define RSN_ERR_NO_SEPARATOR (-11)     /* huffman decode: no 5C 0A => sections[1] panics (huffman.go:261-264) */
#define RSN_ERR_BAD_HEADER (-12)       /* decodeTree index out of range (huffman.go:210) / empty table */
#define RSN_ERR_TRUNCATED (-13)        /* findCodes reads data[max] (huffman.go:145) / pad > bits (huffman.go:294) */
#define RSN_ERR_GUARD (-14)            /* "Max recursion depth" (huffman.go:132-134); strict_limits only */
#define RSN_ERR_BAD_REFERENCE (-15)    /* lz.Decompress slice out of range (lzss.go:349-350) */
#define RSN_ERR_SINGLE_LEAF_LOOP (-16) /* single-leaf tree with bits left: unbounded recursion (huffman.go:139-140) */

/* LZSS variants of the reference */
#define RSN_LZSS_ASYNC 0 /* lz.CompressAsync (lzss.go:109) — what lz.NewWriter / the engine use */
#define RSN_LZSS_ITER 1  /* lz.Compress      (lzss.go:224) — exported iterative variant */

/* ---- library lifetime ------------------------------------------------------------------ */

/* Select the CUDA device for the calling thread's context (default: current device / 0). */
int rsn_init(int device);
/*
 * Release the calling thread's streams, events and scratch memory, and drain the process-wide
 * caches of free result buffers.  A thread that exits releases its own resources by itself.
 */
void rsn_shutdown(void);
const char *rsn_strerror(int rc);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char *rsn_last_cuda_error(void);
/* Release a buffer returned through an `out` parameter of the host-buffer API. */
void rsn_free(void *p);
/* The results of a batch call in one go (NULL entries are skipped). */
void rsn_free_many(void *const *ptrs, size_t count);
void rsn_dev_free_many(void *const *d_ptrs, size_t count, void *stream);
/* Pinned host allocation helpers (optional; any host pointer is accepted as input). */
void *rsn_host_alloc(size_t n);
void rsn_host_free(void *p);

/* ---- host-buffer API: what the cgo shim binds ------------------------------------------- */

/*
 * Replaces the bodies of lz.CompressAsync (lzss.go:109-154, variant RSN_LZSS_ASYNC) and
 * lz.Compress (lzss.go:224-316, variant RSN_LZSS_ITER).  `window` is maxSearchBufferLength
 * (lz.NewWriter passes DefaultWindowSize = 4096, lzss.go:35-40); window <= 0 means an
 * unbounded search buffer exactly as in the reference (lzss.go:125, 249).
 * `in` is only read during the call.  `*out` is library-owned; release with rsn_free().
 */
int rsn_lzss_compress(const uint8_t *in, size_t n, int64_t window, int variant, uint8_t **out, size_t *out_n);
/*
 * lz.CompressAsync of ONE large stream with the match search sharded by position range over
 * `ngpus` shards (BASELINE configs[4]); same bytes as rsn_lzss_compress.  One host thread and CUDA
 * stream per shard inside the call; shard g runs on device g mod the device count.  Each shard holds
 * only its range plus a window of bytes either side; what crosses between GPUs is those halos (peer
 * copies) and, through the host, an 8 KB table per shard.  RSN_LZSS_ITER runs on one GPU.
 */
int rsn_lzss_compress_sharded(const uint8_t *in, size_t n, int64_t window, int variant, int ngpus, uint8_t **out,
                              size_t *out_n);
/* Bytes copied GPU-to-GPU by the calling thread's last rsn_lzss_compress_sharded. */
uint64_t rsn_sharded_peer_bytes(void);
/* Replaces lz.Decompress (lzss.go:323-364). */
int rsn_lzss_decompress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);
/* Replaces huffman.Compress (huffman.go:299-325). */
int rsn_huff_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);
/*
 * Replaces huffman.Decompress (huffman.go:327-330).  strict_limits != 0 reproduces the
 * reference's 900000-bit recursion guard as RSN_ERR_GUARD; 0 lifts it.
 */
int rsn_huff_decompress(const uint8_t *in, size_t n, int strict_limits, uint8_t **out, size_t *out_n);

/*
 * engine.compress / engine.decompress (engine/engine.go:443-479) for a layer list such as
 * "lzss,huffman": compress applies the algorithms left to right, decompress right to left.
 * Intermediate buffers stay on the device.  Only "lzss" and "huffman" are known names.
 */
int rsn_compress_layers(const char *algorithms, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);
int rsn_decompress_layers(const char *algorithms, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);

/*
 * engine.BenchmarkFile (engine/engine.go:357-441) for one buffer: Shannon entropy (natural log) of
 * the input bytes, compress + decompress through the layer list (the timed region), lossless flag,
 * ratio in percent as float32, and ActualEntropy exactly as the reference computes it (histogram of
 * the DEcompressed bytes over the COMPRESSED length, engine.go:412-423).  Histograms and the
 * comparison run on the device.  A codec failure sets `failed` (and `error`) and returns RSN_OK, as
 * AsyncBenchmarkFile's recover() turns a panic into a Failed row (engine.go:315-328).
 */
typedef struct rsn_bench_result {
    double seconds;        /* compress + decompress, wall clock, host buffer in */
    double entropy;        /* Result.Entropy */
    float actual_entropy;  /* Result.ActualEntropy */
    float ratio;           /* Result.Ratio */
    int lossless;          /* Result.Lossless */
    int failed;            /* Result.Failed */
    int error;             /* the codec's code when failed */
    size_t compressed_n, decompressed_n;
} rsn_bench_result;
int rsn_benchmark_file(const char *algorithms, const uint8_t *in, size_t n, rsn_bench_result *res);

/*
 * Batches of independent files (BASELINE configs[3]; what engine.BenchmarkSuite's per-file loop does,
 * engine.go:208-262): file i is compressed (or decompressed) with the layer list exactly as
 * rsn_compress_layers would.  Host buffers (device == 0): files of up to 4 MiB are cut into groups
 * of up to 64 MiB and every kernel of a stage runs once per group (the file index is a grid
 * dimension), so a small file costs no kernel launches or synchronisations of its own; groups, and
 * the files that go one by one (empty, larger than 4 MiB), are spread over `workers` host threads
 * (0 = default), each with its own CUDA stream.  out[i]/out_n[i] receive library-owned buffers
 * (rsn_free each, in any order: the results of a group share one pinned block that goes back to the
 * pool with its last part); rcs[i] (optional) the per-file code: a file the reference would panic on
 * fails alone.  The per-file host work of a group (leaf order, header bytes) uses cores / workers helper
 * threads; RSN_HOST_CORES in the environment overrides the core count (set it to cores / ranks when
 * several ranks share a box).  Returns RSN_OK or the first failing file's code.  With device != 0 the in/out pointers
 * are device pointers on the calling thread's device: inputs are used in place (grouped like host
 * files when 16-byte aligned, otherwise one by one), every result is its own device buffer
 * (release with rsn_dev_free(p, NULL)); nothing crosses PCIe except sizes — and, when the first
 * layer to undo is "huffman", the header of every stream (up to the byte after its first 5C 0A) for
 * the header parser.
 */
int rsn_batch_layers(const char *algorithms, int compress, size_t count, const uint8_t *const *in, const size_t *in_n,
                     uint8_t **out, size_t *out_n, int *rcs, int workers, int device);
/*
 * The grouping rsn_batch_layers applies to host-buffer files of these sizes (host logic only; needs
 * no device): group_of[i] = index of the group file i travels in, or -1 for the per-file path
 * (empty files, files above 4 MiB).  A group holds files of one size class (within a factor of
 * two of each other; everything below 4 KiB is one class), at most 2048 files and at most 64 MiB
 * (less for small batches: total / workers, but at least 16 MiB, so that every worker gets a group).
 */
int rsn_batch_plan(size_t count, const size_t *in_n, int64_t *group_of, size_t *n_groups);

/* ---- device-buffer API ------------------------------------------------------------------- */
/*
 * Same operations with input and output resident in device memory of the context's device.
 * `stream` is a cudaStream_t (NULL = the context's own stream).  `*d_out` is allocated by the
 * library (stream-ordered); release with rsn_dev_free().  The calls synchronise the stream
 * where the algorithm needs a size on the host (documented in DESIGN.md).
 */
int rsn_dev_lzss_compress(const uint8_t *d_in, size_t n, int64_t window, int variant, uint8_t **d_out, size_t *out_n,
                          void *stream);
int rsn_dev_lzss_decompress(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream);
int rsn_dev_huff_compress(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream);
int rsn_dev_huff_decompress(const uint8_t *d_in, size_t n, int strict_limits, uint8_t **d_out, size_t *out_n,
                            void *stream);
void rsn_dev_free(void *d_ptr, void *stream);
/* Convenience copies between host and device buffers on `stream` (synchronous on return). */
int rsn_dev_download(const void *d_src, size_t n, void *h_dst, void *stream);
int rsn_dev_upload(const void *h_src, size_t n, void *d_dst, void *stream);

/*
 * Per-position match arrays of variant A over an already-escaped device buffer: for each i,
 * packed[i] = (len << 16) | off as computed by compressorWorker (lzss.go:166-184).
 * Exposed for parity tests and for the torch.distributed form of the position-range sharded path.
 * window in [1, 32768]; d_enc must be 4-byte aligned (RSN_ERR_INVALID_ARG otherwise).  off is exact
 * where a reference can be emitted (len >= 4; lzss.go:143 needs len >= 6) and unspecified below.
 */
int rsn_dev_lzss_match(const uint8_t *d_enc, size_t n, int64_t window, uint32_t *d_packed, void *stream);

/*
 * Second half of variant A/B compression from precomputed match arrays (the sequential merge
 * of lzss.go:134-151 / 240-311): d_packed[i] as produced by rsn_dev_lzss_match over the same
 * already-escaped buffer.  Used by the position-range sharded path after the arrays of all
 * shards have been gathered.
 */
int rsn_dev_lzss_emit(const uint8_t *d_enc, size_t n, int64_t window, int variant, const uint32_t *d_packed,
                      uint8_t **d_out, size_t *out_n, void *stream);
/* EncodeOpeningSymbols (lzss.go:369-389) alone: the escaped buffer the match arrays refer to. */
int rsn_dev_lzss_escape(const uint8_t *d_in, size_t n, uint8_t **d_out, size_t *out_n, void *stream);

/* ---- introspection ----------------------------------------------------------------------- */

/* Number of kernels this library has launched (from any thread) since the last reset. */
uint64_t rsn_kernel_launches(void);
void rsn_reset_kernel_launches(void);
/*
 * Per-kernel device timing for benchmarks: rsn_kernel_timing(1) makes every kernel launch of the
 * library record two CUDA events on its own stream (and clears earlier records); 0 turns it off.
 * rsn_kernel_timing_report writes one line per kernel name, "<name> <launches> <total ms>\n", most
 * expensive first, into buf (NUL terminated, truncated to cap) and returns the full length.
 */
void rsn_kernel_timing(int enable);
size_t rsn_kernel_timing_report(char *buf, size_t cap);
/* "raisin_b200 <version> sm_100a" */
const char *rsn_version(void);

#ifdef __cplusplus
}
#endif
#endif
