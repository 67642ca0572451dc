/*
 * ssq_oracle.c -- CPU restatement of the ShortSeq hot path (pack / count /
 * Hamming / decode), in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it.  The product (shortseq_b200/) never does, and
 * fails loudly when its CUDA library is missing.
 *
 * Parity is PINNED: tests/test_oracle_pinned.py checks every function here
 * against (a) the known-answer vectors of SURVEY.md section 8c and (b) golden
 * fixtures generated from the unmodified reference built by
 * oracle/build_ref.py (tests/golden/, generator tests/golden/make_golden.py),
 * and -- when oracle/_ref is present -- against the live reference.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Nothing here is copied: the reference is
 * Cython over CPython objects; this is array-in / array-out C.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define SSQ_OK 0
#define SSQ_ERR_BAD_BASE 1   /* "Unsupported base character" */
#define SSQ_ERR_TOO_LONG 2   /* "Sequences longer than 1024 bases are not supported." */
#define SSQ_ERR_CLASS 3      /* length outside the requested container class */
#define SSQ_ERR_LEN_MISMATCH 5 /* Hamming: "requires sequences of equal length" */
#define SSQ_ERR_UB 100       /* reference behaviour undefined (bloom alias byte, SURVEY trap T1) */

/* shortseq/util.pyx:42 */
#define NT_PER_BLOCK 32
/* shortseq/short_seq_64.pyx:27-28, short_seq_192.pyx:21-22, short_seq_var.pyx:8-9 */
#define MAX_64_NT 32
#define MAX_192_NT 96
#define MAX_VAR_NT 1024

/* shortseq/util.pyx:75 */
static const uint64_t BLOOM = 0xFFFFFFFFFFEFFF75ULL;
/* shortseq/util.pyx:39 */
static const uint64_t PEXT_MASK_64 = 0x0606060606060606ULL;
/* shortseq/util.pyx:52 */
static const char CHARMAP[4] = {'A', 'C', 'T', 'G'};

/* shortseq/util.pyx:44-50 : table_91.  Only the entries the reference can
 * legally reach are restated: A->0 C->1 G->3 T->2 U->2, everything else 4. */
static int table_91(uint8_t c, int *ub)
{
    if (c >= 91) { *ub = 1; return 0; }     /* out-of-bounds read in the reference */
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 3;
    case 'T': return 2;
    case 'U': return 2;
    default:  return 4;
    }
}

/* shortseq/util.pxd:98-99 : is_base -- bit (c & 63) of bloom must be clear. */
int ssq_oracle_is_base(uint8_t c)
{
    return (BLOOM & (1ULL << (c & 63))) == 0;
}

/* shortseq/util.pxd:116-127 : _bloom_filter_64 over one 8-byte chunk. */
static int bloom_filter_64(uint64_t block)
{
    uint64_t shifts = block & 0x3F3F3F3F3F3F3F3FULL;
    uint64_t query = 0;
    for (int k = 0; k < 64; k += 8)
        query |= 1ULL << ((shifts >> k) & 0xFF);
    return (BLOOM & query) == 0;
}

/* BMI2 pext restated portably (util.pxd:54-61 binds _pext_u64). */
static uint64_t pext_u64(uint64_t x, uint64_t mask)
{
    uint64_t out = 0;
    int k = 0;
    for (int b = 0; b < 64; b++)
        if ((mask >> b) & 1) { out |= ((x >> b) & 1ULL) << k; k++; }
    return out;
}

/* Error detail for a rejected read, mirroring what the reference's exception
 * message carries (SURVEY trap T2). */
typedef struct {
    int32_t code;        /* SSQ_* */
    int32_t bad_len;     /* 1 (single char) or 8 (whole pext chunk) */
    int64_t bad_pos;     /* byte position inside the read of bad_chars[0] */
    uint8_t bad_chars[8];
} ssq_oracle_err;

/* shortseq/util.pyx:125-140 (_marshall_partial_block) and, identically,
 * shortseq/short_seq_64.pyx:96-108 (_marshall_bytes_64): reverse scan,
 * is_base() then (block<<2)|table_91[c]. */
static int marshall_partial(const uint8_t *seq, size_t length, size_t pos0,
                            uint64_t *out, ssq_oracle_err *err)
{
    uint64_t block = 0;
    int ub = 0;
    for (size_t i = length; i-- > 0;) {
        uint8_t c = seq[i];
        if (!ssq_oracle_is_base(c)) {
            err->code = SSQ_ERR_BAD_BASE; err->bad_len = 1;
            err->bad_pos = (int64_t)(pos0 + i); err->bad_chars[0] = c;
            return SSQ_ERR_BAD_BASE;
        }
        int v = table_91(c, &ub);
        if (ub || v == 4) {               /* bloom alias: reference result is garbage/UB */
            err->code = SSQ_ERR_UB; err->bad_len = 1;
            err->bad_pos = (int64_t)(pos0 + i); err->bad_chars[0] = c;
            return SSQ_ERR_UB;
        }
        block = (block << 2) | (uint64_t)v;
    }
    *out = block;
    return SSQ_OK;
}

/* shortseq/util.pyx:100-119 (_marshall_full_blocks): per 32-nt block, chunks
 * j = 3..0, bloom check on the 8-byte chunk, block = (block<<16)|pext(chunk). */
static int marshall_full_blocks(uint64_t *dst, const uint8_t *seq, size_t n_blocks,
                                ssq_oracle_err *err)
{
    for (size_t i = 0; i < n_blocks; i++) {
        uint64_t block = 0;
        for (int j = 3; j >= 0; j--) {
            uint64_t chunk;
            memcpy(&chunk, seq + 32 * i + 8 * (size_t)j, 8);
            if (!bloom_filter_64(chunk)) {
                err->code = SSQ_ERR_BAD_BASE; err->bad_len = 8;
                err->bad_pos = (int64_t)(32 * i + 8 * (size_t)j);
                memcpy(err->bad_chars, &chunk, 8);
                return SSQ_ERR_BAD_BASE;
            }
            /* alias bytes pass the bloom and silently encode as (c>>1)&3: flag as UB */
            for (int b = 0; b < 8; b++) {
                uint8_t c = (uint8_t)(chunk >> (8 * b));
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
                    err->code = SSQ_ERR_UB; err->bad_len = 1;
                    err->bad_pos = (int64_t)(32 * i + 8 * (size_t)j + (size_t)b);
                    err->bad_chars[0] = c;
                    return SSQ_ERR_UB;
                }
            }
            block = (block << 16) | pext_u64(chunk, PEXT_MASK_64);
        }
        dst[i] = block;
    }
    return SSQ_OK;
}

/* shortseq/util.pyx:78-94 (_marshall_bytes_array): full blocks first, then tail. */
static int marshall_bytes_array(uint64_t *dst, const uint8_t *src, size_t length,
                                ssq_oracle_err *err)
{
    size_t full = length / NT_PER_BLOCK, rem = length % NT_PER_BLOCK;
    int rc = marshall_full_blocks(dst, src, full, err);
    if (rc) return rc;
    if (rem) return marshall_partial(src + 32 * full, rem, 32 * full, &dst[full], err);
    return SSQ_OK;
}

/* shortseq/util.pyx:29-33 (_nt_len_to_block_num) == ceil(len/32). */
size_t ssq_oracle_nblocks(size_t length) { return (length + 31) / 32; }

/* shortseq/short_seq.pyx:54-74 (_new): class by length.
 * returns 0 (len<=32 -> ShortSeq64, incl. the empty singleton), 1 (ShortSeq192),
 * 2 (ShortSeqVar), -1 (too long). */
int ssq_oracle_class(size_t length)
{
    if (length <= MAX_64_NT) return 0;
    if (length <= MAX_192_NT) return 1;
    if (length <= MAX_VAR_NT) return 2;
    return -1;
}

/* Number of u64 words the container of this length holds:
 * ShortSeq64 1 (short_seq_64.pxd:11-14), ShortSeq192 always 3 with unused words
 * zero (short_seq_192.pxd:11-14), ShortSeqVar ceil(L/32)
 * (short_seq_var.pyx:123-132). */
size_t ssq_oracle_container_words(size_t length)
{
    int k = ssq_oracle_class(length);
    if (k == 0) return 1;
    if (k == 1) return 3;
    if (k == 2) return ssq_oracle_nblocks(length);
    return 0;
}

/* One read: shortseq/short_seq.pyx:54-74 dispatch + the marshallers above.
 * words must hold ssq_oracle_container_words(length) entries; it is zeroed
 * first (tp_alloc / PyObject_Calloc zero the reference's storage). */
int ssq_oracle_pack_one(const uint8_t *seq, size_t length, uint64_t *words,
                        ssq_oracle_err *err)
{
    ssq_oracle_err local;
    if (!err) err = &local;
    memset(err, 0, sizeof(*err));
    int k = ssq_oracle_class(length);
    if (k < 0) { err->code = SSQ_ERR_TOO_LONG; return SSQ_ERR_TOO_LONG; }
    size_t nw = ssq_oracle_container_words(length);
    memset(words, 0, nw * sizeof(uint64_t));
    if (length == 0) return SSQ_OK;                        /* short_seq.pyx:55-56 */
    if (k == 0)                                            /* short_seq.pyx:57-62 */
        return marshall_partial(seq, length, 0, &words[0], err);
    return marshall_bytes_array(words, seq, length, err);  /* :63-72 */
}

/*
 * Batch pack of a class-homogeneous batch.
 *   klass 0: words[n]      (ShortSeq64)
 *   klass 1: words[n][3]   (ShortSeq192)
 *   klass 2: words[word_off[i] .. word_off[i+1])  (ShortSeqVar, CSR; word_off
 *            has n+1 entries and is an OUTPUT)
 * lens[i] receives the length (int32 so it fits 1024).
 * Stops at the FIRST failing read in list order, like
 * ShortSeqCounter(list) / a Python loop over sq.pack (counter.pyx:23-29);
 * *first_bad receives its index (or -1).  A read whose length is outside the
 * requested class yields SSQ_ERR_CLASS (or SSQ_ERR_TOO_LONG beyond 1024).
 */
int ssq_oracle_pack_batch(int klass, const uint8_t *ascii, const int64_t *offsets,
                          int64_t n, uint64_t *words, int32_t *lens,
                          int64_t *word_off, int64_t *first_bad, ssq_oracle_err *err)
{
    ssq_oracle_err local;
    if (!err) err = &local;
    memset(err, 0, sizeof(*err));
    if (first_bad) *first_bad = -1;
    int64_t wo = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t len = offsets[i + 1] - offsets[i];
        if (len < 0 || len > MAX_VAR_NT) {
            err->code = SSQ_ERR_TOO_LONG; if (first_bad) *first_bad = i;
            return SSQ_ERR_TOO_LONG;
        }
        if (ssq_oracle_class((size_t)len) != klass) {
            err->code = SSQ_ERR_CLASS; if (first_bad) *first_bad = i;
            return SSQ_ERR_CLASS;
        }
        uint64_t *dst = klass == 0 ? words + i : klass == 1 ? words + 3 * i : words + wo;
        if (klass == 2) word_off[i] = wo;
        int rc = ssq_oracle_pack_one(ascii + offsets[i], (size_t)len, dst, err);
        if (rc) { if (first_bad) *first_bad = i; return rc; }
        lens[i] = (int32_t)len;
        wo += (int64_t)ssq_oracle_container_words((size_t)len);
    }
    if (klass == 2) word_off[n] = wo;
    return SSQ_OK;
}

/* shortseq/short_seq_64.pyx:35-36, short_seq_192.pyx:29-30, short_seq_var.pyx:16-17:
 * __hash__ returns the first block; CPython then maps -1 to -2. */
int64_t ssq_oracle_pyhash(uint64_t word0)
{
    int64_t h = (int64_t)word0;
    return h == -1 ? -2 : h;
}

/* shortseq/short_seq_64.pyx:77-84, short_seq_192.pyx:74-91, short_seq_var.pyx:64-81:
 * sum over ceil(L/32) blocks of popcount(((x>>1)|x) & 0x5555...) with x = a^b. */
int ssq_oracle_hamming(const uint64_t *a, size_t len_a, const uint64_t *b, size_t len_b,
                       int64_t *dist)
{
    if (len_a != len_b) return SSQ_ERR_LEN_MISMATCH;
    size_t nb = ssq_oracle_nblocks(len_a);
    int64_t d = 0;
    for (size_t i = 0; i < nb; i++) {
        uint64_t x = a[i] ^ b[i];
        x = ((x >> 1) | x) & 0x5555555555555555ULL;
        d += __builtin_popcountll(x);
    }
    *dist = d;
    return SSQ_OK;
}

/* shortseq/short_seq_64.pyx:114-121, short_seq_192.pyx:114-127,
 * short_seq_var.pyx:98-120: out[i] = charmap[(word >> 2(i%32)) & 3]. */
void ssq_oracle_decode(const uint64_t *words, size_t length, uint8_t *out)
{
    for (size_t i = 0; i < length; i++)
        out[i] = (uint8_t)CHARMAP[(words[i / 32] >> (2 * (i % 32))) & 3];
}

/* Batched wrappers over fixed-stride word arrays (stride = words per read). */
int ssq_oracle_hamming_batch(const uint64_t *a, const uint64_t *b, const int32_t *len_a,
                             const int32_t *len_b, int64_t n, size_t stride, int32_t *dist,
                             int64_t *first_bad)
{
    if (first_bad) *first_bad = -1;
    for (int64_t i = 0; i < n; i++) {
        int64_t d;
        int rc = ssq_oracle_hamming(a + stride * i, (size_t)len_a[i], b + stride * i,
                                    (size_t)len_b[i], &d);
        if (rc) { if (first_bad) *first_bad = i; return rc; }
        dist[i] = (int32_t)d;
    }
    return SSQ_OK;
}

void ssq_oracle_decode_batch(const uint64_t *words, const int64_t *word_off, size_t stride,
                             const int32_t *lens, int64_t n, uint8_t *out,
                             const int64_t *out_off)
{
    for (int64_t i = 0; i < n; i++) {
        const uint64_t *w = word_off ? words + word_off[i] : words + stride * i;
        ssq_oracle_decode(w, (size_t)lens[i], out + out_off[i]);
    }
}

/*
 * Dedup counting: shortseq/counter.pyx:23-29 (_count_py_bytes_list) +
 * counter.pyx:41-54 (_count_sequence).  The reference is a CPython dict keyed
 * by the ShortSeq object: equality = same class, same length, same first
 * ceil(L/32) words (short_seq_64.pyx:41-44, short_seq_192.pyx:35-41); the
 * dict's slot hash is word0 (util.pxd:68-70) which parity never observes, so
 * any hash may be used here.  Items come out in first-occurrence order (dict
 * insertion order).  Input is a class-homogeneous packed batch with a fixed
 * word stride (1 or 3); ShortSeqVar is never deduplicated by the reference
 * (SURVEY trap T3) and is not offered.
 *
 * Outputs (caller allocates n entries each): uniq_words[stride*u],
 * uniq_lens[u], counts[u], first_idx[u]; returns the number of uniques.
 */
static uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

int64_t ssq_oracle_count(const uint64_t *words, const int32_t *lens, int64_t n, size_t stride,
                         uint64_t *uniq_words, int32_t *uniq_lens, int64_t *counts,
                         int64_t *first_idx)
{
    size_t cap = 16;
    while (cap < (size_t)n * 2 + 16) cap <<= 1;
    int64_t *slots = (int64_t *)malloc(cap * sizeof(int64_t));
    if (!slots) return -1;
    for (size_t i = 0; i < cap; i++) slots[i] = -1;
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint64_t *w = words + stride * (size_t)i;
        uint64_t h = (uint64_t)lens[i] * 0x9E3779B97F4A7C15ULL;
        for (size_t k = 0; k < stride; k++) h = mix64(h ^ w[k]);
        size_t s = (size_t)h & (cap - 1);
        for (;;) {
            int64_t e = slots[s];
            if (e < 0) {
                slots[s] = u;
                memcpy(uniq_words + stride * (size_t)u, w, stride * sizeof(uint64_t));
                uniq_lens[u] = lens[i]; counts[u] = 1; first_idx[u] = i;
                u++;
                break;
            }
            if (uniq_lens[e] == lens[i] &&
                memcmp(uniq_words + stride * (size_t)e, w, stride * sizeof(uint64_t)) == 0) {
                counts[e]++;
                break;
            }
            s = (s + 1) & (cap - 1);
        }
    }
    free(slots);
    return u;
}

/*
 * Synthetic read generator shared by the oracle, the tests and bench.py
 * (SURVEY section 8d).  Measurement tooling, not reference behaviour; the CUDA
 * library has the same generator (ssq_synth_reads) so CPU and GPU see identical
 * bytes without a transfer.
 *   key_id(i)       = mix64(seed + i) mod n_keys
 *   base(key_id, j) = "ACGT"[(mix64(seed2 + key_id*32 + j/32) >> 2(j%32)) & 3]
 * Read length: len_lo + (mix64(seed3 + key_id) mod (len_hi-len_lo+1)) -- a
 * function of key_id so that equal keys give equal reads.
 */
static const char SYNTH_ALPHA[4] = {'A', 'C', 'G', 'T'};

uint64_t ssq_oracle_mix64(uint64_t x) { return mix64(x); }

int64_t ssq_oracle_synth_key(uint64_t seed, int64_t i, int64_t n_keys)
{
    return (int64_t)(mix64(seed + (uint64_t)i) % (uint64_t)n_keys);
}

int32_t ssq_oracle_synth_len(uint64_t seed, int64_t key, int32_t len_lo, int32_t len_hi)
{
    if (len_hi <= len_lo) return len_lo;
    return len_lo + (int32_t)(mix64(seed + 0x5EED0003ULL + (uint64_t)key) %
                              (uint64_t)(len_hi - len_lo + 1));
}

/* Fills offsets[0..n] (offsets[0] = 0) and, when ascii != NULL, the bases.
 * Returns the total number of bytes.  first_read lets a shard generate the
 * reads [first_read, first_read+n) of the global sequence. */
int64_t ssq_oracle_synth_reads(uint64_t seed, int64_t first_read, int64_t n, int64_t n_keys,
                               int32_t len_lo, int32_t len_hi, uint8_t *ascii,
                               int64_t *offsets)
{
    int64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t key = ssq_oracle_synth_key(seed, first_read + i, n_keys);
        int32_t len = ssq_oracle_synth_len(seed, key, len_lo, len_hi);
        offsets[i] = pos;
        if (ascii) {
            for (int32_t j = 0; j < len; j++) {
                uint64_t r = mix64(seed + 0x5EED0002ULL + (uint64_t)key * 32ULL + (uint64_t)(j / 32));
                ascii[pos + j] = (uint8_t)SYNTH_ALPHA[(r >> (2 * (j % 32))) & 3];
            }
        }
        pos += len;
    }
    offsets[n] = pos;
    return pos;
}
