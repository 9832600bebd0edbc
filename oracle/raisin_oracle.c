/*
 * raisin_oracle.c — CPU restatement of go-compression/raisin's LZSS + Huffman hot path.
 * TEST INFRASTRUCTURE ONLY (see raisin_oracle.h for the rules and the parity-pin status).
 *
 * Reference files restated here (paths relative to the reference repository):
 *   compressor/lz/lzss.go          109-184 (CompressAsync + workers), 224-320 (Compress,
 *                                  getEncoding), 323-364 (Decompress), 366-406 (escapes),
 *                                  418-433 (FindReverseSlice / FindReverse)
 *   compressor/huffman/huffman.go  58-103 (buildTree), 110-127 (printCodes), 131-153
 *                                  (findCodes), 174-191 (AsByteSlice), 196-227 (decodeTree),
 *                                  229-256 (encode), 258-297 (decode), 299-325 (Compress)
 * Go standard-library behaviour relied on by those lines (Go 1.15, go.mod:3) is restated
 * where used: bytes.Index, container/heap, strconv.Atoi/Itoa, range-over-string UTF-8
 * decoding, string(rune) encoding, strings.SplitN.
 */
#define _GNU_SOURCE
#include "raisin_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void rsno_free(void *p) { free(p); }

/* ------------------------------------------------------------------ small helpers */

typedef struct {
    uint8_t *p;
    size_t n, cap;
} buf_t;

static int buf_reserve(buf_t *b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 64;
    while (nc < b->n + extra) nc *= 2;
    uint8_t *q = (uint8_t *)realloc(b->p, nc);
    if (!q) return -1;
    b->p = q;
    b->cap = nc;
    return 0;
}
static int buf_put(buf_t *b, const uint8_t *s, size_t n) {
    if (buf_reserve(b, n)) return -1;
    memcpy(b->p + b->n, s, n);
    b->n += n;
    return 0;
}
static int buf_putc(buf_t *b, uint8_t c) { return buf_put(b, &c, 1); }

static int ndig_u64(uint64_t v) {
    int d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}

/* strconv.Itoa for v >= 0 */
static int put_dec(buf_t *b, uint64_t v) {
    char tmp[24];
    int k = 0;
    do {
        tmp[k++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (k--)
        if (buf_putc(b, (uint8_t)tmp[k])) return -1;
    return 0;
}

/* lzss.go:318-320 getEncoding: "<" + Itoa(pointer) + "," + Itoa(size) + ">" */
static int put_token(buf_t *b, uint64_t ptr, uint64_t size) {
    if (buf_putc(b, '<') || put_dec(b, ptr) || buf_putc(b, ',') || put_dec(b, size) || buf_putc(b, '>')) return -1;
    return 0;
}
static int token_len(uint64_t ptr, uint64_t size) { return 3 + ndig_u64(ptr) + ndig_u64(size); }

/*
 * strconv.Atoi (Go 1.15 strconv/atoi.go) with the error discarded, as lzss.go:338,346 do:
 * syntax error => 0; range error => clamped int64 extreme.  The scan stops at the first
 * offending byte, so an overflow seen before a bad byte reports the range clamp.
 */
static int64_t go_atoi(const uint8_t *s, size_t n) {
    if (n == 0) return 0;
    int neg = 0;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        s++;
        n--;
        if (n == 0) return 0;
    }
    const uint64_t maxv = UINT64_MAX;
    const uint64_t cutoff = maxv / 10 + 1;
    uint64_t un = 0;
    int range = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        if (c < '0' || c > '9') return 0; /* ErrSyntax => (0, err) */
        if (un >= cutoff) {
            un = maxv;
            range = 1;
            break;
        }
        un *= 10;
        uint64_t n1 = un + (uint64_t)(c - '0');
        if (n1 < un) {
            un = maxv;
            range = 1;
            break;
        }
        un = n1;
    }
    (void)range;
    const uint64_t icut = (uint64_t)1 << 63;
    if (!neg && un >= icut) return INT64_MAX;
    if (neg && un > icut) return INT64_MIN;
    int64_t v = (int64_t)un;
    return neg ? -v : v;
}

/* ------------------------------------------------------------------ escape layer */

/* lzss.go:369-389 — the third branch (380-385) is unreachable and foundEscape is never set. */
int rsno_escape(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    buf_t b = {0};
    if (buf_reserve(&b, n + n / 8 + 16)) return RSNO_ERR_NOMEM;
    for (size_t i = 0; i < n; i++) {
        uint8_t v = in[i];
        if (v == 0x3C) {
            v = 0xFF;
        } else if (v == 0xFF || v == 0x5C) {
            if (buf_putc(&b, 0x5C)) goto oom;
        }
        if (buf_putc(&b, v)) goto oom;
    }
    *out = b.p;
    *out_n = b.n;
    return RSNO_OK;
oom:
    free(b.p);
    return RSNO_ERR_NOMEM;
}

/* lzss.go:391-406 */
int rsno_unescape(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    buf_t b = {0};
    if (buf_reserve(&b, n + 16)) return RSNO_ERR_NOMEM;
    int esc = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t v = in[i];
        if (v == 0xFF && !esc) {
            b.p[b.n++] = 0x3C;
        } else if (v == 0x5C && !esc) {
            esc = 1;
        } else {
            esc = 0;
            b.p[b.n++] = v;
        }
    }
    *out = b.p;
    *out_n = b.n;
    return RSNO_OK;
}

/* ------------------------------------------------------------------ LZSS variant A */

/* bytes.Index: smallest index of pat in hay, or -1 (lzss.go:418-421 FindReverseSlice). */
static ptrdiff_t bytes_index(const uint8_t *hay, size_t hn, const uint8_t *pat, size_t pn) {
    if (pn == 0) return 0;
    if (pn > hn) return -1;
    const uint8_t *r = (const uint8_t *)memmem(hay, hn, pat, pn);
    return r ? r - hay : -1;
}

/*
 * compressorWorker (lzss.go:166-184), unrolled from recursion to a loop: the pattern starts
 * as enc[i:i+1] and grows by one byte while bytes.Index still finds it in the window and
 * bytes remain (len(nextBytes) > 0); the deepest level that was found is returned, with
 * negativeOffset = len(searchBuffer) - index of THAT level's leftmost hit.
 */
static void match_at(const uint8_t *enc, size_t n, int64_t window, int mode, size_t i, uint32_t *len_out,
                     uint32_t *off_out) {
    size_t ws = 0;
    if (window > 0 && i > (size_t)window) ws = i - (size_t)window; /* lzss.go:123-127 */
    const uint8_t *win = enc + ws;
    size_t wn = i - ws;
    size_t best = 0, best_idx = 0;
    if (mode != 0) {
        /* literal: one more byte per level, each level a fresh leftmost search (bytes.Index) */
        for (size_t k = 1; i + k <= n; k++) {
            ptrdiff_t idx = bytes_index(win, wn, enc + i, k);
            if (idx < 0) break;
            best = k;
            best_idx = (size_t)idx;
        }
    } else {
        /*
         * Same result, fewer searches: "enc[i:i+k] occurs in win" is monotone in k (a prefix of
         * an occurring pattern occurs), so the deepest found level is the largest such k; find
         * it by doubling then bisection, then take the leftmost hit of that k.
         */
        size_t kmax = n - i < wn ? n - i : wn;
        size_t lo = 0, hi = 1; /* invariant: lo occurs (lo = 0 trivially), answer < hi+... */
        while (hi <= kmax && bytes_index(win, wn, enc + i, hi) >= 0) {
            lo = hi;
            hi *= 2;
        }
        if (hi > kmax) hi = kmax + 1;
        /* largest k in [lo, hi) that occurs */
        while (lo + 1 < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (bytes_index(win, wn, enc + i, mid) >= 0)
                lo = mid;
            else
                hi = mid;
        }
        best = lo;
        if (best) best_idx = (size_t)bytes_index(win, wn, enc + i, best);
    }
    *len_out = (uint32_t)best;
    *off_out = best ? (uint32_t)(wn - best_idx) : 0;
}

typedef struct {
    const uint8_t *enc;
    size_t n;
    int64_t window;
    int mode;
    uint32_t *len_out, *off_out;
    size_t *next; /* shared chunk cursor */
    pthread_mutex_t *mu;
} match_job_t;

#define MATCH_CHUNK 2048

static void *match_thread(void *arg) {
    match_job_t *j = (match_job_t *)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        size_t lo = *j->next;
        *j->next = lo + MATCH_CHUNK;
        pthread_mutex_unlock(j->mu);
        if (lo >= j->n) break;
        size_t hi = lo + MATCH_CHUNK < j->n ? lo + MATCH_CHUNK : j->n;
        for (size_t i = lo; i < hi; i++) match_at(j->enc, j->n, j->window, j->mode, i, j->len_out + i, j->off_out + i);
    }
    return NULL;
}

int rsno_lzss_match_arrays(const uint8_t *enc, size_t n, int64_t window, int mode, int threads, uint32_t *len_out,
                           uint32_t *off_out) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads == 1 || n < 4 * MATCH_CHUNK) {
        for (size_t i = 0; i < n; i++) match_at(enc, n, window, mode, i, len_out + i, off_out + i);
        return RSNO_OK;
    }
    pthread_t th[256];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    size_t next = 0;
    match_job_t job = {enc, n, window, mode, len_out, off_out, &next, &mu};
    int started = 0;
    for (int t = 0; t < threads; t++) {
        if (pthread_create(&th[t], NULL, match_thread, &job)) break;
        started++;
    }
    if (started == 0) match_thread(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return RSNO_OK;
}

/* lzss.go:109-154: escape, per-position workers, then the sequential merge (134-151). */
int rsno_lzss_compress_async(const uint8_t *in, size_t n, int64_t window, int mode, int threads, uint8_t **out,
                             size_t *out_n) {
    uint8_t *enc = NULL;
    size_t en = 0;
    int rc = rsno_escape(in, n, &enc, &en);
    if (rc) return rc;
    uint32_t *L = (uint32_t *)malloc((en + 1) * sizeof(uint32_t));
    uint32_t *O = (uint32_t *)malloc((en + 1) * sizeof(uint32_t));
    buf_t b = {0};
    if (!L || !O || buf_reserve(&b, en + 16)) goto oom;
    rsno_lzss_match_arrays(enc, en, window, mode, threads, L, O);
    size_t i = 0;
    while (i < en) {
        if (L[i] == 0) { /* !ref.isReference: value is the single byte (lzss.go:148-150) */
            if (buf_putc(&b, enc[i])) goto oom;
            i += 1;
        } else { /* ignoreNextChars = size-1; strict '<' at lzss.go:143 */
            if (token_len(O[i], L[i]) < (int)L[i]) {
                if (put_token(&b, O[i], L[i])) goto oom;
            } else {
                if (buf_put(&b, enc + i, L[i])) goto oom;
            }
            i += L[i];
        }
    }
    free(enc);
    free(L);
    free(O);
    *out = b.p;
    *out_n = b.n;
    return RSNO_OK;
oom:
    free(enc);
    free(L);
    free(O);
    free(b.p);
    return RSNO_ERR_NOMEM;
}

/* ------------------------------------------------------------------ LZSS variant B */

/* lzss.go:423-433 FindReverse: note the second decrement inside the loop body (stride 2). */
static ptrdiff_t find_reverse(const uint8_t *s, size_t n, uint8_t v) {
    for (ptrdiff_t i = (ptrdiff_t)n - 1; i >= 0; i--) {
        if (s[i] == v) return i;
        i--;
    }
    return -1;
}

/* lzss.go:224-316, statement for statement. */
int rsno_lzss_compress_iter(const uint8_t *in, size_t n, int64_t window, uint8_t **out, size_t *out_n) {
    uint8_t *enc = NULL;
    size_t en = 0;
    int rc = rsno_escape(in, n, &enc, &en);
    if (rc) return rc;
    buf_t sb = {0}, ob = {0}, m = {0}, pat = {0};
    if (buf_reserve(&sb, en + 16) || buf_reserve(&ob, en + 16) || buf_reserve(&m, 64) || buf_reserve(&pat, 64)) goto oom;
    int check = 0;
    uint64_t start_ptr = 0, check_off = 0;
    for (size_t p = 0; p < en; p++) {
        uint8_t v = enc[p];
        ptrdiff_t idx = -1;
        if (!check) {
            idx = find_reverse(sb.p, sb.n, v); /* lzss.go:246 */
        } else {
            size_t dr = 0;
            if (window > 0 && sb.n > (size_t)window) dr = sb.n - (size_t)window; /* lzss.go:248-251 */
            pat.n = 0;
            if (buf_put(&pat, m.p, m.n) || buf_putc(&pat, v)) goto oom;
            idx = bytes_index(sb.p + dr, sb.n - dr, pat.p, pat.n); /* lzss.go:252, window-relative index */
        }
        int found = idx >= 0;
        if (found && check) { /* lzss.go:255-259: pointer uses the window-relative index */
            start_ptr = (uint64_t)sb.n - (uint64_t)idx;
            check_off++;
            if (buf_putc(&m, v)) goto oom;
        } else if (found && !check) { /* lzss.go:260-266 */
            start_ptr = (uint64_t)sb.n - (uint64_t)idx;
            check_off = 1;
            check = 1;
            if (buf_putc(&m, v)) goto oom;
        } else {
            if (check) { /* lzss.go:268-288 */
                int should_add = 1;
                if ((size_t)token_len(start_ptr, check_off) > m.n) should_add = 0;
                if (should_add) {
                    if (put_token(&ob, start_ptr, check_off)) goto oom;
                } else {
                    if (buf_put(&ob, m.p, m.n)) goto oom;
                }
                start_ptr = 0;
                check_off = 0;
                check = 0;
                if (buf_put(&sb, m.p, m.n)) goto oom;
                m.n = 0;
            }
            if (buf_putc(&ob, v)) goto oom; /* lzss.go:289 */
        }
        if (!check) {
            if (buf_putc(&sb, v)) goto oom; /* lzss.go:292-295 */
        }
    }
    if (check) { /* lzss.go:297-311 */
        int should_add = 1;
        if ((size_t)token_len(start_ptr, check_off) > m.n) should_add = 0;
        if (should_add) {
            if (put_token(&ob, start_ptr, check_off)) goto oom;
        } else {
            if (buf_put(&ob, m.p, m.n)) goto oom;
        }
    }
    free(enc);
    free(sb.p);
    free(m.p);
    free(pat.p);
    *out = ob.p;
    *out_n = ob.n;
    return RSNO_OK;
oom:
    free(enc);
    free(sb.p);
    free(m.p);
    free(pat.p);
    free(ob.p);
    return RSNO_ERR_NOMEM;
}

/* ------------------------------------------------------------------ LZSS decode */

/*
 * lzss.go:323-364.  `output` and `searchBuffer` always hold the same bytes, so one buffer is
 * kept.  Go evaluates searchBuffer[a:a+cnt] before appending (no overlap semantics) and panics
 * when the slice is out of range; reading slack capacity (a+cnt > len but <= cap) is treated
 * as the same error here because the bytes Go would read are unspecified.
 */
int rsno_lzss_decompress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    buf_t sb = {0}, pb = {0}, cb = {0};
    if (buf_reserve(&sb, n + 16) || buf_reserve(&pb, 32) || buf_reserve(&cb, 32)) goto oom;
    enum { OPEN, SEP, CLOSE } st = OPEN;
    int64_t ptr = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t v = in[i];
        if (st == OPEN && v == '<') {
            st = SEP;
        } else if (st == SEP) {
            if (v == ',') {
                st = CLOSE;
                ptr = go_atoi(pb.p, pb.n);
                pb.n = 0;
            } else if (buf_putc(&pb, v))
                goto oom;
        } else if (st == CLOSE) {
            if (v == '>') {
                st = OPEN;
                int64_t cnt = go_atoi(cb.p, cb.n);
                cb.n = 0;
                /* a = len - ptr; need 0 <= a <= a+cnt <= len  <=>  0 <= cnt <= ptr <= len */
                if (cnt < 0 || ptr < cnt || (uint64_t)ptr > (uint64_t)sb.n) {
                    free(sb.p);
                    free(pb.p);
                    free(cb.p);
                    return RSNO_ERR_BAD_REFERENCE;
                }
                size_t a = sb.n - (size_t)ptr;
                if (buf_reserve(&sb, (size_t)cnt)) goto oom;
                memcpy(sb.p + sb.n, sb.p + a, (size_t)cnt); /* source lies wholly before sb.n */
                sb.n += (size_t)cnt;
            } else if (buf_putc(&cb, v))
                goto oom;
        } else {
            if (buf_putc(&sb, v)) goto oom;
        }
    }
    free(pb.p);
    free(cb.p);
    int rc = rsno_unescape(sb.p, sb.n, out, out_n); /* lzss.go:362 */
    free(sb.p);
    return rc;
oom:
    free(sb.p);
    free(pb.p);
    free(cb.p);
    return RSNO_ERR_NOMEM;
}

/* ------------------------------------------------------------------ Go UTF-8 */

/*
 * One step of `for _, c := range s` (Go spec + unicode/utf8.DecodeRuneInString): returns the
 * rune at s[0] and stores its width.  Any malformed or truncated sequence yields U+FFFD with
 * width 1.  Accept ranges follow utf8's `first`/`acceptRanges` tables.
 */
static int32_t go_decode_rune(const uint8_t *s, size_t n, int *width) {
    uint8_t c = s[0];
    *width = 1;
    if (c < 0x80) return c;
    int need;
    uint8_t lo = 0x80, hi = 0xBF;
    if (c >= 0xC2 && c <= 0xDF)
        need = 2;
    else if (c == 0xE0) {
        need = 3;
        lo = 0xA0;
    } else if ((c >= 0xE1 && c <= 0xEC) || c == 0xEE || c == 0xEF)
        need = 3;
    else if (c == 0xED) {
        need = 3;
        hi = 0x9F;
    } else if (c == 0xF0) {
        need = 4;
        lo = 0x90;
    } else if (c >= 0xF1 && c <= 0xF3)
        need = 4;
    else if (c == 0xF4) {
        need = 4;
        hi = 0x8F;
    } else
        return 0xFFFD;
    if (n < (size_t)need) return 0xFFFD;
    if (s[1] < lo || s[1] > hi) return 0xFFFD;
    if (need == 2) {
        *width = 2;
        return ((int32_t)(c & 0x1F) << 6) | (s[1] & 0x3F);
    }
    if (s[2] < 0x80 || s[2] > 0xBF) return 0xFFFD;
    if (need == 3) {
        *width = 3;
        return ((int32_t)(c & 0x0F) << 12) | ((int32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F);
    }
    if (s[3] < 0x80 || s[3] > 0xBF) return 0xFFFD;
    *width = 4;
    return ((int32_t)(c & 0x07) << 18) | ((int32_t)(s[1] & 0x3F) << 12) | ((int32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
}

/* string(rune): UTF-8 encoding; surrogates and out-of-range values become U+FFFD. */
static int go_encode_rune(int32_t r, uint8_t *o) {
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

size_t rsno_utf8_decode(const uint8_t *in, size_t n, int32_t *runes_out) {
    size_t m = 0, i = 0;
    while (i < n) {
        int w;
        runes_out[m++] = go_decode_rune(in + i, n - i, &w);
        i += (size_t)w;
    }
    return m;
}

/* ------------------------------------------------------------------ Huffman tree */

typedef struct {
    int64_t freq; /* Go int; sums wrap like Go's */
    int32_t rune; /* leaf value, -1 for internal nodes */
    int32_t left, right;
} hnode_t;

typedef struct {
    hnode_t *nodes;
    int32_t n_nodes;
    int32_t root;
} htree_t;

typedef struct {
    int64_t freq;
    int32_t rune;
} leaf_t;

static int leaf_cmp(const void *a, const void *b) {
    const leaf_t *x = (const leaf_t *)a, *y = (const leaf_t *)b;
    if (x->freq != y->freq) return x->freq < y->freq ? -1 : 1;
    if (x->rune != y->rune) return x->rune < y->rune ? -1 : 1;
    return 0;
}

/* container/heap (Go 1.15) over an array of node indices; Less is freq-only (huffman.go:43-45). */
static inline int h_less(const hnode_t *nd, const int32_t *h, int i, int j) { return nd[h[i]].freq < nd[h[j]].freq; }
static inline void h_swap(int32_t *h, int i, int j) {
    int32_t t = h[i];
    h[i] = h[j];
    h[j] = t;
}
static void h_up(const hnode_t *nd, int32_t *h, int j) {
    for (;;) {
        int i = (j - 1) / 2; /* Go integer division truncates toward zero: j=0 -> i=0 */
        if (i == j || !h_less(nd, h, j, i)) break;
        h_swap(h, i, j);
        j = i;
    }
}
static void h_down(const hnode_t *nd, int32_t *h, int i0, int n) {
    int i = i0;
    for (;;) {
        int j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int j = j1;
        int j2 = j1 + 1;
        if (j2 < n && h_less(nd, h, j2, j1)) j = j2;
        if (!h_less(nd, h, j, i)) break;
        h_swap(h, i, j);
        i = j;
    }
}

/*
 * buildTree (huffman.go:58-103).  The remove-and-resort loop at 64-87 yields the leaves in
 * (freq ascending, rune ascending) order; then heap.Init, and Pop,Pop,Push until one is left.
 * leaves[] need not be sorted on entry.  k must be >= 1.
 */
static int build_tree(leaf_t *leaves, size_t k, htree_t *t) {
    qsort(leaves, k, sizeof(leaf_t), leaf_cmp);
    t->nodes = (hnode_t *)malloc((2 * k) * sizeof(hnode_t));
    int32_t *heap = (int32_t *)malloc(k * sizeof(int32_t));
    if (!t->nodes || !heap) {
        free(t->nodes);
        free(heap);
        return RSNO_ERR_NOMEM;
    }
    for (size_t i = 0; i < k; i++) {
        t->nodes[i].freq = leaves[i].freq;
        t->nodes[i].rune = leaves[i].rune;
        t->nodes[i].left = t->nodes[i].right = -1;
        heap[i] = (int32_t)i;
    }
    int n = (int)k;
    int32_t nn = (int32_t)k;
    for (int i = n / 2 - 1; i >= 0; i--) h_down(t->nodes, heap, i, n); /* heap.Init */
    while (n > 1) {
        /* heap.Pop: Swap(0,n-1); down(0,n-1); remove last */
        h_swap(heap, 0, n - 1);
        h_down(t->nodes, heap, 0, n - 1);
        int32_t a = heap[--n];
        h_swap(heap, 0, n - 1);
        h_down(t->nodes, heap, 0, n - 1);
        int32_t b = heap[--n];
        t->nodes[nn].freq = (int64_t)((uint64_t)t->nodes[a].freq + (uint64_t)t->nodes[b].freq);
        t->nodes[nn].rune = -1;
        t->nodes[nn].left = a;
        t->nodes[nn].right = b;
        heap[n++] = nn; /* heap.Push: append; up(len-1) */
        h_up(t->nodes, heap, n - 1);
        nn++;
    }
    t->root = heap[0];
    t->n_nodes = nn;
    free(heap);
    return RSNO_OK;
}

/*
 * printCodes (huffman.go:110-127): DFS, left appends '0', right appends '1'.  Fills
 * code/len per NODE index (only leaves meaningful).  Iterative to survive deep trees.
 * Returns -1 if some leaf's code is longer than 64 bits.
 */
static int tree_codes(const htree_t *t, uint64_t *code, uint8_t *len) {
    int32_t *stack = (int32_t *)malloc((size_t)t->n_nodes * sizeof(int32_t));
    uint32_t *depth = (uint32_t *)calloc((size_t)t->n_nodes, sizeof(uint32_t));
    if (!stack || !depth) {
        free(stack);
        free(depth);
        return RSNO_ERR_NOMEM;
    }
    int sp = 0, rc = RSNO_OK;
    stack[sp++] = t->root;
    code[t->root] = 0;
    depth[t->root] = 0;
    while (sp) {
        int32_t v = stack[--sp];
        const hnode_t *nd = &t->nodes[v];
        if (nd->left < 0) {
            if (depth[v] > 64) rc = -1;
            len[v] = (uint8_t)(depth[v] > 255 ? 255 : depth[v]);
            continue;
        }
        code[nd->left] = code[v] << 1;
        code[nd->right] = (code[v] << 1) | 1;
        depth[nd->left] = depth[nd->right] = depth[v] + 1;
        stack[sp++] = nd->right;
        stack[sp++] = nd->left;
    }
    free(stack);
    free(depth);
    return rc;
}

int rsno_huff_code_table(const int32_t *runes, const uint64_t *freqs, size_t k, uint64_t *code_out, uint8_t *len_out) {
    if (k == 0) return RSNO_ERR_EMPTY_INPUT;
    leaf_t *lv = (leaf_t *)malloc(k * sizeof(leaf_t));
    if (!lv) return RSNO_ERR_NOMEM;
    for (size_t i = 0; i < k; i++) {
        lv[i].freq = (int64_t)freqs[i];
        lv[i].rune = runes[i];
    }
    htree_t t = {0};
    int rc = build_tree(lv, k, &t);
    if (rc) {
        free(lv);
        return rc;
    }
    uint64_t *code = (uint64_t *)calloc((size_t)t.n_nodes, sizeof(uint64_t));
    uint8_t *len = (uint8_t *)calloc((size_t)t.n_nodes, 1);
    rc = tree_codes(&t, code, len);
    /* leaves were sorted inside build_tree: map back by rune */
    for (size_t i = 0; i < k && rc == RSNO_OK; i++) {
        for (size_t j = 0; j < k; j++) {
            if (t.nodes[j].rune == runes[i]) {
                code_out[i] = code[j];
                len_out[i] = len[j];
                break;
            }
        }
    }
    free(code);
    free(len);
    free(t.nodes);
    free(lv);
    return rc;
}

/* ------------------------------------------------------------------ Huffman encode */

#define MAX_RUNE_TABLE 0x110000

/* huffman.go:299-325 Compress + 229-256 encode + 174-191 AsByteSlice. */
int rsno_huff_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    if (n == 0) return RSNO_ERR_EMPTY_INPUT; /* buildTree on an empty map: heap.Pop panics */
    uint64_t *freq = (uint64_t *)calloc(MAX_RUNE_TABLE, sizeof(uint64_t));
    if (!freq) return RSNO_ERR_NOMEM;
    size_t i = 0;
    while (i < n) { /* huffman.go:309-311 */
        int w;
        int32_t r = go_decode_rune(in + i, n - i, &w);
        freq[r]++;
        i += (size_t)w;
    }
    size_t k = 0;
    for (int32_t r = 0; r < MAX_RUNE_TABLE; r++)
        if (freq[r]) k++;
    leaf_t *lv = (leaf_t *)malloc(k * sizeof(leaf_t));
    int32_t *order = (int32_t *)malloc(k * sizeof(int32_t));
    if (!lv || !order) {
        free(freq);
        free(lv);
        free(order);
        return RSNO_ERR_NOMEM;
    }
    size_t q = 0;
    for (int32_t r = 0; r < MAX_RUNE_TABLE; r++)
        if (freq[r]) {
            lv[q].freq = (int64_t)freq[r];
            lv[q].rune = r;
            order[q] = r; /* header order: ascending rune ... */
            q++;
        }
    if (k >= 2 && order[k - 1] == 0x5C) { /* ... except never end on '\\' (decoder would panic) */
        int32_t t = order[k - 1];
        order[k - 1] = order[k - 2];
        order[k - 2] = t;
    }
    buf_t b = {0};
    htree_t t = {0};
    uint64_t *code = NULL;
    uint8_t *len = NULL;
    uint64_t *rcode = NULL;
    uint8_t *rlen = NULL;
    int rc = RSNO_ERR_NOMEM;
    for (size_t j = 0; j < k; j++) { /* huffman.go:312-318 */
        int32_t r = order[j];
        if (put_dec(&b, freq[r]) || buf_putc(&b, '|')) goto done;
        if (r == 10) {
            if (buf_putc(&b, '\\') || buf_putc(&b, 'n')) goto done;
        } else {
            uint8_t u[4];
            int w = go_encode_rune(r, u);
            if (buf_put(&b, u, (size_t)w)) goto done;
        }
    }
    rc = build_tree(lv, k, &t);
    if (rc) goto done;
    rc = RSNO_ERR_NOMEM;
    code = (uint64_t *)calloc((size_t)t.n_nodes, sizeof(uint64_t));
    len = (uint8_t *)calloc((size_t)t.n_nodes, 1);
    rcode = (uint64_t *)calloc(MAX_RUNE_TABLE, sizeof(uint64_t));
    rlen = (uint8_t *)calloc(MAX_RUNE_TABLE, 1);
    if (!code || !len || !rcode || !rlen) goto done;
    rc = tree_codes(&t, code, len);
    if (rc) goto done;
    rc = RSNO_ERR_NOMEM;
    uint64_t bits = 0;
    for (size_t j = 0; j < k; j++) {
        int32_t r = t.nodes[j].rune;
        rcode[r] = code[j];
        rlen[r] = len[j];
        bits += (uint64_t)len[j] * freq[r];
    }
    unsigned pad = (unsigned)((8 - bits % 8) % 8); /* huffman.go:245-249 */
    if (buf_putc(&b, '\\') || buf_putc(&b, '\n') || buf_putc(&b, (uint8_t)pad)) goto done;
    size_t payload = (size_t)((bits + pad) / 8);
    if (buf_reserve(&b, payload + 8)) goto done;
    uint8_t *pp = b.p + b.n;
    memset(pp, 0, payload);
    uint64_t bp = pad; /* right-aligned: the pad zeros sit at the front (huffman.go:178-183) */
    i = 0;
    while (i < n) { /* huffman.go:235-241 */
        int w;
        int32_t r = go_decode_rune(in + i, n - i, &w);
        i += (size_t)w;
        uint64_t c = rcode[r];
        for (int s = rlen[r] - 1; s >= 0; s--) {
            if ((c >> s) & 1) pp[bp >> 3] |= (uint8_t)(0x80 >> (bp & 7));
            bp++;
        }
    }
    b.n += payload;
    *out = b.p;
    *out_n = b.n;
    b.p = NULL;
    rc = RSNO_OK;
done:
    free(freq);
    free(lv);
    free(order);
    free(code);
    free(len);
    free(rcode);
    free(rlen);
    free(t.nodes);
    free(b.p);
    return rc;
}

/* ------------------------------------------------------------------ Huffman decode */

/*
 * decodeTree (huffman.go:196-227): returns the rune->freq map as a sorted-unique leaf list.
 * Digits accumulate into temp, other non-'|' bytes are ignored; at '|' freq = Atoi(temp)
 * (error => 0), the symbol is "\n" for the two bytes 5C 6E, else the rune decoded at i+1.
 */
static int parse_header(const uint8_t *h, size_t hn, leaf_t **leaves_out, size_t *k_out) {
    size_t cap = 64, k = 0;
    leaf_t *lv = (leaf_t *)malloc(cap * sizeof(leaf_t));
    buf_t temp = {0};
    if (!lv) return RSNO_ERR_NOMEM;
    for (size_t i = 0; i < hn; i++) {
        if (h[i] != '|') {
            if (h[i] >= '0' && h[i] <= '9')
                if (buf_putc(&temp, h[i])) goto oom;
        } else {
            int64_t f = go_atoi(temp.p, temp.n);
            temp.n = 0;
            int32_t sym;
            if (i + 1 >= hn) goto bad; /* tree[i+1] out of range */
            if (h[i + 1] == '\\') {
                if (i + 2 >= hn) goto bad; /* tree[i+2] out of range */
                if (h[i + 2] == 'n') {
                    sym = 10;
                    i++;
                } else {
                    sym = '\\';
                }
            } else {
                int w;
                sym = go_decode_rune(h + i + 1, hn - (i + 1), &w);
            }
            if (k == cap) {
                cap *= 2;
                leaf_t *q = (leaf_t *)realloc(lv, cap * sizeof(leaf_t));
                if (!q) goto oom;
                lv = q;
            }
            lv[k].rune = sym;
            lv[k].freq = f;
            k++;
            i++;
        }
    }
    free(temp.p);
    /* map semantics: the last assignment to a rune wins */
    {
        /* stable dedupe keeping the last: mark by scanning from the end with a seen-table */
        uint8_t *seen = (uint8_t *)calloc(MAX_RUNE_TABLE, 1);
        if (!seen) {
            free(lv);
            return RSNO_ERR_NOMEM;
        }
        size_t w = k;
        for (size_t j = k; j-- > 0;) {
            if (!seen[lv[j].rune]) {
                seen[lv[j].rune] = 1;
                lv[--w] = lv[j];
            }
        }
        memmove(lv, lv + w, (k - w) * sizeof(leaf_t));
        k -= w;
        free(seen);
    }
    *leaves_out = lv;
    *k_out = k;
    return RSNO_OK;
bad:
    free(temp.p);
    free(lv);
    return RSNO_ERR_BAD_HEADER;
oom:
    free(temp.p);
    free(lv);
    return RSNO_ERR_NOMEM;
}

int rsno_huff_decompress(const uint8_t *in, size_t n, int strict, uint8_t **out, size_t *out_n) {
    /* strings.SplitN(s, "\\\n", 2): split at the FIRST 5C 0A (huffman.go:261) */
    const uint8_t sep[2] = {0x5C, 0x0A};
    ptrdiff_t sp = bytes_index(in, n, sep, 2);
    if (sp < 0) return RSNO_ERR_NO_SEPARATOR;
    leaf_t *lv = NULL;
    size_t k = 0;
    int rc = parse_header(in, (size_t)sp, &lv, &k);
    if (rc) return rc;
    if (k == 0) {
        free(lv);
        return RSNO_ERR_BAD_HEADER; /* buildTree on empty map panics */
    }
    htree_t t = {0};
    rc = build_tree(lv, k, &t);
    free(lv);
    if (rc) return rc;
    const uint8_t *pay = in + sp + 2;
    size_t pn = n - (size_t)sp - 2;
    uint64_t diff = pn ? pay[0] : 0;                 /* huffman.go:274-278 */
    uint64_t nbits = pn ? (uint64_t)(pn - 1) * 8 : 0; /* bytes after the first */
    const uint8_t *bits = pay + 1;
    buf_t ob = {0};
    if (diff > nbits) { /* contentString.String()[int(diff):] out of range */
        rc = RSNO_ERR_TRUNCATED;
        goto done;
    }
    uint64_t max = nbits - diff;
    /* findCodes (huffman.go:131-153) as a loop over bit index i */
    if (strict && max > 900000) {
        rc = RSNO_ERR_GUARD;
        goto done;
    }
    if (buf_reserve(&ob, (size_t)(max / 2 + 16))) {
        rc = RSNO_ERR_NOMEM;
        goto done;
    }
    {
        uint64_t i = 0;
        int32_t node = t.root;
        for (;;) {
            const hnode_t *nd = &t.nodes[node];
            if (nd->left < 0) { /* HuffmanLeaf */
                uint8_t u[4];
                int w = go_encode_rune(nd->rune, u);
                if (buf_put(&ob, u, (size_t)w)) {
                    rc = RSNO_ERR_NOMEM;
                    goto done;
                }
                if (i < max) {
                    if (node == t.root) { /* single-leaf tree never consumes a bit */
                        rc = RSNO_ERR_SINGLE_LEAF_LOOP;
                        goto done;
                    }
                    node = t.root;
                    continue;
                }
                break;
            }
            if (i >= max) { /* data[i] with i == len(data) */
                rc = RSNO_ERR_TRUNCATED;
                goto done;
            }
            uint64_t bp = diff + i;
            int bit = (bits[bp >> 3] >> (7 - (bp & 7))) & 1;
            node = bit ? nd->right : nd->left;
            i++;
        }
    }
    *out = ob.p;
    *out_n = ob.n;
    ob.p = NULL;
    rc = RSNO_OK;
done:
    free(ob.p);
    free(t.nodes);
    return rc;
}
