/*
 * shortseq_b200.h -- C ABI of the B200-native ShortSeq hot path
 * (batched pack / dedup-count / Hamming / decode of short reads).
 *
 * The reference (AlexTate/ShortSeq) has no FFI layer: its boundary is the
 * CPython type protocol of its Cython classes.  Each entry point below names
 * the reference function(s) whose per-object work it performs for a whole
 * batch (paths relative to the reference repo, file:line).  INTEGRATION.md
 * shows the Cython `cdef extern` / ctypes stub a maintainer would add.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  Unless a name says `host`, every data
 *    pointer is a DEVICE pointer on the context's GPU and the call only
 *    ENQUEUES work on the context's stream; nothing is read back until
 *    ssq_ctx_sync().  The caller owns every buffer; the library owns only
 *    contexts and counters.
 *  - Return value: SSQ_OK or an SSQ_ERR_* for errors detectable at enqueue
 *    time (bad argument, CUDA launch failure).  DATA errors (non-ACGT base,
 *    over-long read, length mismatch, table overflow) are recorded on the
 *    device and reported by the next ssq_ctx_sync() for the LOWEST read index,
 *    which is the read the reference's serial loop would have raised on
 *    (counter.pyx:23-29).
 *  - Batch input format: one contiguous uint8 ASCII buffer holding the reads
 *    back to back plus int64 offsets[n+1]; read i is
 *    ascii[offsets[i] .. offsets[i+1]).
 *  - Packed layout (util.pyx:109-118,133-138): base j of a read lives in bits
 *    [2(j%32), 2(j%32)+1] of 64-bit word j/32, code = (c>>1)&3 (A0 C1 T2 G3),
 *    unused high bits zero.  ShortSeq64: 1 word/read (0..32 nt).  ShortSeq192:
 *    3 words/read, AoS words[n][3], unused words zero (33..96 nt).
 *    ShortSeqVar: CSR, words[word_off[i] .. word_off[i+1]) with
 *    ceil(len/32) words (97..1024 nt).
 *  - There is no CPU fallback: without a CUDA device every compute call fails
 *    with SSQ_ERR_CUDA.
 */
#ifndef SHORTSEQ_B200_H
#define SHORTSEQ_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSQ_ABI_VERSION 1

/* status codes */
#define SSQ_OK 0
#define SSQ_ERR_BAD_BASE 1      /* "Unsupported base character" (short_seq_64.pyx:105, util.pyx:115,137) */
#define SSQ_ERR_TOO_LONG 2      /* "Sequences longer than 1024 bases are not supported." (short_seq.pyx:74) */
#define SSQ_ERR_CLASS 3         /* read length outside the container class of the call (or negative) */
#define SSQ_ERR_CUDA 4          /* CUDA runtime error / no device; see ssq_last_error() */
#define SSQ_ERR_LEN_MISMATCH 5  /* "Hamming distance requires sequences of equal length" (short_seq_64.pyx:78-80) */
#define SSQ_ERR_TABLE_FULL 6    /* counter overflowed beyond recovery; the counter must be discarded */
#define SSQ_ERR_ARG 7           /* invalid argument */
#define SSQ_ERR_EXCHANGE 8      /* multi-GPU merge: another rank's share never arrived (ssq_counter_merge_alltoall) */

/* container classes (short_seq.pyx:54-74) */
#define SSQ_CLASS_64 0   /* 0..32 nt   */
#define SSQ_CLASS_192 1  /* 33..96 nt  */
#define SSQ_CLASS_VAR 2  /* 97..1024 nt */

typedef struct ssq_ctx ssq_ctx;         /* one GPU + one stream + error state */
typedef struct ssq_counter ssq_counter; /* device hash table of (len, words) -> count */

/* Result of ssq_ctx_sync(): the first data error in read order since the last sync. */
typedef struct ssq_report {
    int32_t code;          /* SSQ_OK or the SSQ_ERR_* of the lowest failing read */
    int32_t reserved;
    int64_t first_bad_read; /* index of that read within its call's batch, or -1 */
} ssq_report;

/* ---- library / context ------------------------------------------------- */
int ssq_abi_version(void);
const char *ssq_last_error(void);      /* thread-local message for the last SSQ_ERR_CUDA / SSQ_ERR_ARG */
uint64_t ssq_launch_count(void);      /* kernels this library has launched in this process (all contexts) */
int ssq_device_count(int *count);
int ssq_ctx_create(int device, ssq_ctx **out);
int ssq_ctx_destroy(ssq_ctx *ctx);
/* Enqueue on an externally owned cudaStream_t (e.g. the caller framework's current stream).  The handle
 * is used as given: NULL is CUDA's legacy default stream.  ssq_ctx_reset_stream() returns to the
 * context's own (non-blocking) stream, which is what a new context uses. */
int ssq_ctx_set_stream(ssq_ctx *ctx, void *cuda_stream);
int ssq_ctx_reset_stream(ssq_ctx *ctx);
void *ssq_ctx_stream(ssq_ctx *ctx);
/* Wait for all enqueued work, fetch and clear the device error record. */
int ssq_ctx_sync(ssq_ctx *ctx, ssq_report *report);

/* device / pinned-host memory helpers for callers that have no CUDA binding of their own */
int ssq_malloc(ssq_ctx *ctx, size_t bytes, void **dptr);
int ssq_free(ssq_ctx *ctx, void *dptr);
int ssq_host_alloc(size_t bytes, void **hptr);   /* pinned */
int ssq_host_free(void *hptr);
int ssq_memcpy_h2d(ssq_ctx *ctx, void *dst, const void *src, size_t bytes); /* async on ctx stream */
int ssq_memcpy_d2h(ssq_ctx *ctx, void *dst, const void *src, size_t bytes); /* async on ctx stream */
int ssq_memset(ssq_ctx *ctx, void *dst, int value, size_t bytes);

/* ---- packing ------------------------------------------------------------
 * Replaces, for a whole batch: short_seq.pyx:54-74 (_new) ->
 * short_seq_64.pyx:96-108 (_marshall_bytes_64) / short_seq_192.pyx:103-108 ->
 * util.pyx:78-140 (_marshall_bytes_array, _marshall_full_blocks,
 * _marshall_partial_block), with validation (util.pxd:98-127) fused into the
 * same pass.  Validation is EXACT {A,C,G,T}; the reference's bloom filter also
 * lets 16 alias byte values through with undefined results (SURVEY trap T1).
 * ascii_bytes = size of the ASCII buffer (so vector loads never leave it). */
int ssq_pack64(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets,
               int64_t n, uint64_t *words /*[n]*/, uint8_t *lens /*[n]*/);
int ssq_pack192(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets,
                int64_t n, uint64_t *words /*[n][3]*/, uint8_t *lens /*[n]*/);
/* ShortSeqVar (short_seq_var.pyx:123-132).  word_off[n+1] is an OUTPUT (exclusive scan of
 * ceil(len/32)); words must hold at least ssq_packvar_words_bound(ascii_bytes, n) entries. */
int64_t ssq_packvar_words_bound(int64_t ascii_bytes, int64_t n);
int ssq_packvar(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets,
                int64_t n, int64_t *word_off /*[n+1]*/, uint64_t *words, uint16_t *lens /*[n]*/);

/* ---- decoding -------------------------------------------------------------
 * Replaces short_seq_64.pyx:114-121, short_seq_192.pyx:114-127,
 * short_seq_var.pyx:98-120 (charmap util.pyx:52).  out_offsets[n+1] is the
 * exclusive scan of lens, produced by ssq_lens_to_offsets. */
int ssq_lens_to_offsets(ssq_ctx *ctx, const void *lens, int len_bytes /*1|2*/, int64_t n,
                        int64_t *out_offsets /*[n+1]*/);
int ssq_decode64(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n,
                 const int64_t *out_offsets, uint8_t *ascii_out);
int ssq_decode192(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n,
                  const int64_t *out_offsets, uint8_t *ascii_out);
/* Fused-offsets form for the fixed classes (no n-entry offsets scan): ssq_decode_tiles sums the lengths of every tile of
 * SSQ_DECODE_TILE reads and scans the tile totals into tile_base (the buffer holds 2 * ntiles + 2 int64, ntiles =
 * ceil(n / SSQ_DECODE_TILE); tile_base[ntiles] = total bases, what the caller sizes ascii_out with); ssq_decode64_fused /
 * ssq_decode192_fused then decode, deriving every read's offset inside the kernel and writing out_offsets[n + 1] as well. */
#define SSQ_DECODE_TILE 512
int ssq_decode_tiles(ssq_ctx *ctx, const uint8_t *lens, int64_t n, int max_len /*32 | 96*/, int64_t *tile_base);
int ssq_decode64_fused(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *tile_base,
                       int64_t *out_offsets, uint8_t *ascii_out);
int ssq_decode192_fused(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *tile_base,
                        int64_t *out_offsets, uint8_t *ascii_out);
int ssq_decodevar(ssq_ctx *ctx, const uint64_t *words, const int64_t *word_off, const uint16_t *lens,
                  int64_t n, const int64_t *out_offsets, uint8_t *ascii_out);

/* ---- Hamming distance (the `^` operator) ----------------------------------
 * Replaces short_seq_64.pyx:77-84, short_seq_192.pyx:74-91, short_seq_var.pyx:64-81:
 * sum over blocks of popcount(((x>>1)|x) & 0x5555...), x = a^b.  Pairs with different
 * lengths record SSQ_ERR_LEN_MISMATCH (their dist is written as the all-ones value). */
int ssq_hamming_pairs64(ssq_ctx *ctx, const uint64_t *a, const uint8_t *len_a, const uint64_t *b,
                        const uint8_t *len_b, int64_t n, uint8_t *dist);
int ssq_hamming_pairs192(ssq_ctx *ctx, const uint64_t *a, const uint8_t *len_a, const uint64_t *b,
                         const uint8_t *len_b, int64_t n, uint8_t *dist);
int ssq_hamming_pairsvar(ssq_ctx *ctx, const uint64_t *a, const int64_t *a_off, const uint16_t *len_a,
                         const uint64_t *b, const int64_t *b_off, const uint16_t *len_b, int64_t n,
                         uint16_t *dist);
/* Each query against a reference set of nr sequences (UMI-collapse style).  Only refs of the
 * query's length are comparable; min_dist = 255 / argmin = 0xFFFFFFFF when there is none.
 * Ties keep the lowest ref index.  n_within (optional, may be NULL) = number of refs at
 * distance <= thresh.  words_per_seq is 1 (ShortSeq64) or 3 (ShortSeq192). */
int ssq_hamming_refset(ssq_ctx *ctx, int words_per_seq, const uint64_t *q, const uint8_t *len_q,
                       int64_t nq, const uint64_t *refs, const uint8_t *len_r, int32_t nr,
                       int32_t thresh, uint8_t *min_dist, uint32_t *argmin, uint32_t *n_within);

/* ---- single objects (HOST pointers) -----------------------------------------
 * The reference's per-object calls -- sq.pack(x) (short_seq.pyx:13-74), str(s) (short_seq_64.pyx:114-121,
 * short_seq_192.pyx:114-127, short_seq_var.pyx:98-120) and a ^ b (short_seq_64.pyx:77-84 and twins) -- as
 * batches of one ON THE DEVICE: operands travel through a page of mapped pinned memory the context owns, the call
 * launches one small kernel on the context's stream and returns when its result has arrived (no device allocation,
 * no copy call, no stream synchronisation).  len is 1..1024; the empty sequence never reaches the library.
 * ssq_pack_one: words[] receives 1 (len <= 32), 3 (len <= 96) or ceil(len/32) canonical blocks, *klass the container
 *   class; returns SSQ_ERR_BAD_BASE (and the offending position in *first_bad, may be NULL) for a base outside
 *   {A,C,G,T} -- the reference's "Unsupported base character" -- in which case words[] must not be used.
 * ssq_decode_one: ascii_out receives len characters.   ssq_hamming_one: both operands hold ceil(len/32) (or 1 / 3)
 *   canonical blocks of sequences of the same length len. */
int ssq_pack_one(ssq_ctx *ctx, const uint8_t *ascii, int32_t len, uint64_t *words, int32_t *klass, int32_t *first_bad);
int ssq_decode_one(ssq_ctx *ctx, const uint64_t *words, int32_t len, uint8_t *ascii_out);
int ssq_hamming_one(ssq_ctx *ctx, const uint64_t *a, const uint64_t *b, int32_t len, int32_t *dist);

/* ---- dedup counting --------------------------------------------------------
 * Replaces ShortSeqCounter (counter.pyx:10-54): key = (length, words) -- the reference's
 * __eq__ (short_seq_64.pyx:41-44, short_seq_192.pyx:35-41); value = multiplicity.  The
 * reference's dict slot hash (word0, util.pxd:68-70) is not observable; the table mixes
 * its own.  klass is SSQ_CLASS_64 or SSQ_CLASS_192 (the reference never deduplicates
 * ShortSeqVar, SURVEY trap T3).  expected_unique > 0 is the caller's bound on the distinct keys:
 * the table gets at least twice that many slots and every batch is counted in one pass; the table
 * grows before a pass that could load it beyond 75 % and after one that left it above 60 %, and a
 * table that holds more keys than the bound continues as if expected_unique were 0.  Only a single
 * pass that alone brings far more distinct keys than the bound can fill a table region; that is
 * reported as SSQ_ERR_TABLE_FULL by ssq_ctx_sync (the counter must then be discarded).
 * expected_unique = 0: no bound -- gated sub-batches, the table grows x4 whenever it has to.
 * hash_rot (0..56) rotates the slot hash so that a table holding only the keys of one
 * hash partition (multi-GPU owner tables) still spreads over all slots. */
int ssq_counter_create(ssq_ctx *ctx, int klass, int64_t expected_unique, int hash_rot,
                       ssq_counter **out);
int ssq_counter_destroy(ssq_counter *c);
int ssq_counter_clear(ssq_counter *c);
/* count already packed reads (counter.pyx:31-33 _count_short_seq_vector) */
int ssq_counter_insert(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n);
/* add weighted keys: counts[i] occurrences of key i (merge step of the multi-GPU counter) */
int ssq_counter_merge(ssq_counter *c, const uint64_t *words, const uint8_t *lens,
                      const uint64_t *counts, int64_t n);
/* fused pack + count (counter.pyx:23-29 _count_py_bytes_list): also emits the packed batch */
int ssq_counter_pack_count(ssq_counter *c, const uint8_t *ascii, int64_t ascii_bytes,
                           const int64_t *offsets, int64_t n, uint64_t *words, uint8_t *lens);
/* record, per key, the lowest (base_index + i) at which it occurs in this packed batch, so
 * that exports can be put in the reference's dict order (first occurrence). */
int ssq_counter_first_index(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n,
                            int64_t base_index);
/* counts of the given keys (0 when absent): dict lookup / `in` */
int ssq_counter_lookup(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n,
                       uint64_t *counts);
/* number of distinct keys; synchronises the context (data errors stay pending) */
int ssq_counter_size(ssq_counter *c, int64_t *n_unique);
int ssq_counter_capacity(ssq_counter *c, int64_t *slots);
/* Device time (CUDA events on the context's stream) of the last single-pass ssq_counter_pack_count:
 * phase 1 = fused pack (+ scatter to hash partitions for tables larger than L2), phase 2 = routing the keys to
 * their table regions and counting them there (0 when the keys were inserted directly).  For profiling. */
int ssq_counter_last_pass_ms(ssq_counter *c, float *phase1_ms, float *phase2_ms);
/* The same pass split three ways: fused pack + level-1 scatter, level-2 (per table region) scatter, and the
 * shared-memory region count.  The last two are 0 / everything-after-phase-1 when the keys took another path. */
int ssq_counter_last_pass_detail(ssq_counter *c, float *pack_scatter_ms, float *region_scatter_ms, float *count_ms);
/* Export all (key, len, count[, first_idx]) tuples, grouped into n_parts hash partitions
 * (owner = top log2(n_parts) bits of the key hash; n_parts a power of two, 1 = no
 * grouping).  Buffers must hold ssq_counter_size() tuples; part_counts[n_parts] (device)
 * receives the tuples per partition; first_idx may be NULL. */
int ssq_counter_export(ssq_counter *c, int n_parts, uint64_t *words, uint8_t *lens,
                       uint64_t *counts, int64_t *first_idx, int64_t *part_counts);
/* Region-aligned ssq_counter_merge for ShortSeq64 owner tables of a multi-GPU merge.  The tuples are n_blocks consecutive
 * blocks (block_counts[] on the HOST), one per sending rank, each in the order ssq_counter_export_to wrote it; block b
 * comes from a table of block_regions[b]
 * regions (ssq_counter_regions of the sender, divided by the number of owners) and region_bases + b * rb_stride holds
 * the sender's exclusive scan of its region sizes for this owner (ssq_counter_export_region_bases; block_regions[b] + 1
 * entries).  Every owner region is counted in shared memory from one contiguous range of every block: one pass, no
 * global atomics.  Falls back to the plain weighted insert when the region grids do not nest. */
int ssq_counter_regions(ssq_counter *c, int64_t *n_regions);
int ssq_counter_merge_regions(ssq_counter *c, const uint64_t *words, const uint8_t *lens, const uint64_t *counts,
                              const int64_t *block_counts, const int64_t *block_regions, int n_blocks,
                              const int64_t *region_bases, int64_t rb_stride);
/* dst[p] (device table of n_parts DEVICE pointers, peer memory allowed) receives regions/n_parts + 1 offsets: where each of
 * partition p's regions starts inside the block ssq_counter_export_to sends to p. */
int ssq_counter_export_region_bases(ssq_counter *c, int n_parts, int64_t *const *dst);
/* Multi-GPU send side without an intermediate buffer (ShortSeq64 counters): part_counts[n_parts] (device) = tuples
 * per hash partition; then partition p's tuples are written to dst_words[p][0..], dst_lens[p][0..], dst_counts[p][0..]
 * -- device arrays of n_parts DEVICE pointers, each of which may point into another GPU's memory (opened with
 * ssq_ipc_open): the export kernel stores over NVLink, no collective call moves the data.  The kernel walks the
 * partitions cyclically starting at first_part: rank r of P passes (r + 1) % P so that at any moment every owner
 * receives from one sender instead of all senders converging on owner 0, then owner 1, ... */
int ssq_counter_export_counts(ssq_counter *c, int n_parts, int64_t *part_counts);
int ssq_counter_export_to(ssq_counter *c, int n_parts, int first_part, uint64_t *const *dst_words,
                          uint8_t *const *dst_lens, uint64_t *const *dst_counts);
/* ---- multi-GPU merge (the reference is single-process) -----------------------------------------
 * One process per GPU of one box.  Every rank counts its own shard of the reads into a LOCAL counter (no
 * communication); ssq_counter_merge_alltoall then moves the distinct keys -- not the reads -- to their OWNER rank
 * (top log2(world) bits of the key hash) and adds them into `owner` (a counter created with hash_rot = log2(world));
 * the global counter is the disjoint union of the owner tables.  world must be a power of two.
 *   ssq_comm_unique_id   rank 0 produces 128 opaque bytes (an ncclUniqueId) and passes them to the others out of band
 *   ssq_comm_init        collective; NCCL (libnccl.so.2) is loaded at run time
 *   ssq_counter_merge_alltoall   collective.  ShortSeq64: sizes by ncclAllGather, then the export kernel stores every
 *       owner's share into that owner's receive buffer over NVLink (CUDA IPC), arrival flags on the device, the owner
 *       counts region by region in shared memory.  ShortSeq192 / no IPC: grouped ncclSend / ncclRecv of a staged
 *       export.  exchange_ms / merge_ms (may be NULL): device time of this rank's send side and of its owner-side merge.
 *   ssq_comm_attach      collective; the STREAMED exchange.  Binds a ShortSeq64 `local` counter (and the owner it will be
 *       merged into) to the communicator: from then on the kernel that counts a table region during
 *       ssq_counter_pack_count also stores that region's final (key, count) pairs into a fixed place of the owning rank's
 *       receive buffer (over NVLink), so the transfer overlaps the count and ssq_counter_merge_alltoall has nothing left to
 *       export: it publishes the arrival flags and merges.  Every rank must attach tables of the same capacity; *streams
 *       (may be NULL) tells whether streaming is active (0: no CUDA IPC / regions do not nest -- merges take the path above).
 *       A pass that cannot stream (table grown, direct-insert path) silently falls back for that merge, on every rank alike;
 *       attach again after `local` has grown.  Every streamed pass sends the WHOLE table (earlier passes' keys included):
 *       attach before the one large pass, or the last one, not around a chunked ingest.
 *       local == NULL detaches and frees the buffers (collective).
 *   A rank whose peers never deliver gets SSQ_ERR_EXCHANGE from the next ssq_ctx_sync (bounded device-side wait). */
typedef struct ssq_comm ssq_comm;
int ssq_comm_unique_id(uint8_t *id128);
int ssq_comm_init(ssq_ctx *ctx, const uint8_t *id128, int rank, int world, ssq_comm **out);
int ssq_comm_destroy(ssq_comm *comm);
int ssq_comm_uses_peer_stores(ssq_comm *comm);
int ssq_comm_attach(ssq_comm *comm, ssq_counter *local, ssq_counter *owner, int *streams);
int ssq_comm_last_merge_streamed(ssq_comm *comm);   /* 1 when the last ssq_counter_merge_alltoall used the streamed regions */
int ssq_counter_merge_alltoall(ssq_comm *comm, ssq_counter *local, ssq_counter *owner, float *exchange_ms, float *merge_ms);

/* CUDA IPC for the peer exchange: handle64 = 64 opaque bytes to pass to the other single-GPU processes of the box. */
int ssq_ipc_get_handle(ssq_ctx *ctx, void *dptr, void *handle64);
int ssq_ipc_open(ssq_ctx *ctx, const void *handle64, void **dptr);
int ssq_ipc_close(ssq_ctx *ctx, void *dptr);

/* ---- host-buffer pipeline ---------------------------------------------------
 * The call a host-language binding makes with HOST memory: chunks the batch,
 * overlaps host->device copies, the fused pack+count kernel and the
 * device->host copy of the packed words/lens on separate streams, and
 * returns after everything completed (report filled as by ssq_ctx_sync).
 * h_words / h_lens may be NULL when only the counts are wanted.  Pinned host
 * buffers (ssq_host_alloc) give full PCIe bandwidth. */
int ssq_host_pack_count(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii,
                        const int64_t *h_offsets, int64_t n, uint64_t *h_words, uint8_t *h_lens,
                        int64_t chunk_reads, ssq_report *report);
/* Same pipeline with the read boundaries given as one uint8 length per read (what a list of bytes objects
 * carries; reads of the countable classes are <= 96 nt): 1 instead of 8 bytes per read cross PCIe, the offsets
 * are scanned on the device, and only the packed words come back (the caller already holds the lengths).
 * h_words may be NULL. */
int ssq_host_pack_count_lens(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii, const uint8_t *h_lens,
                             int64_t n, uint64_t *h_words, int64_t chunk_reads, ssq_report *report);

/* Reads per container class of a batch (short_seq.pyx:54-74 picks the class by length): counts is a DEVICE array of 5
 * int64 -- reads of 0..32, 33..96, 97..1024 and > 1024 nt, then the lowest read index longer than 96 nt (-1 if none). */
int ssq_classify(ssq_ctx *ctx, const int64_t *offsets, int64_t n, int64_t *counts);

/* ---- FASTQ ingest ----------------------------------------------------------
 * Replaces read_and_count_fastq (counter.pyx:57-71) and its getline loop (fast_read.pyx:3-20): h_text is the whole
 * FASTQ file in host memory; every line whose 1-based number is 2 mod 4 is a read, minus its last byte (the newline --
 * or the last base of an unterminated final line, which the reference drops too, short_seq.pyx:50-52).  The text goes
 * to the GPU in chunks of chunk_bytes (0 = 256 MB); newline scan, line selection, gather into ASCII + offsets and the
 * fused pack+count all run on the device.  Reads of 0..32 nt are counted into c64, 33..96 nt into c192 (either may be
 * NULL when the file holds no such read).  n_reads = reads seen; n_longer / first_longer = number and lowest read
 * number of reads longer than 96 nt (not counted: SSQ_CLASS_VAR keys cannot be deduplicated by the reference either).
 * report = lowest read number with a bad base (code SSQ_ERR_BAD_BASE); counting stops at the chunk that holds it.
 * track_first_index != 0 records first-occurrence read numbers for ssq_counter_export (dict order). */
int ssq_host_fastq_count(ssq_ctx *ctx, ssq_counter *c64, ssq_counter *c192, const uint8_t *h_text, int64_t nbytes,
                         int64_t chunk_bytes, int track_first_index, int64_t *n_reads, int64_t *n_longer,
                         int64_t *first_longer, ssq_report *report);

/* ---- batched slicing, k-mers, tolerant alphabet ------
 * Replace, for whole arrays, the reference's per-object slice constructors (short_seq.pyx:94-116 _slice,
 * :119-199 _slice_to_ShortSeq64/192/Var, :202-238 _shift_copy_trim: funnel-shift the blocks, trim the tail).
 * A packed array is (klass, words, word_off, lens): word_off only for SSQ_CLASS_VAR (lens uint16 then, else uint8).
 * ssq_slice: out[i] = in[i][start_i : stop_i], both clamped to the read like Python's slices (the caller resolves
 * negative indices) and to `width` bases.  starts / stops are per-read int64 arrays, or NULL for the scalars
 * start0 / stop0.  out_klass must hold `width` bases (SSQ_CLASS_64: <= 32, SSQ_CLASS_192: <= 96, else SSQ_CLASS_VAR:
 * its out_word_off[n+1] comes from ssq_slice_words, out_lens is uint16).  Results are tail-trimmed: bits >= 2 len are 0. */
int ssq_slice_words(ssq_ctx *ctx, int in_klass, const void *lens, int64_t n, const int64_t *starts, const int64_t *stops,
                    int64_t start0, int64_t stop0, int32_t width, int64_t *out_word_off);
int ssq_slice(ssq_ctx *ctx, int in_klass, const uint64_t *words, const int64_t *word_off, const void *lens, int64_t n,
              const int64_t *starts, const int64_t *stops, int64_t start0, int64_t stop0, int32_t width, int out_klass,
              uint64_t *out_words, const int64_t *out_word_off, void *out_lens);
/* Every k-mer (1 <= k <= 32, every `stride`-th start) of every read as a ShortSeq64 word, grouped by read:
 * ssq_kmers_count fills kmer_off[n+1] (exclusive scan of (len - k) / stride + 1 for reads of at least k bases),
 * ssq_kmers64 writes kmer_off[n] words and lengths (= k).  Counting them with ssq_counter_insert is k-mer counting. */
int ssq_kmers_count(ssq_ctx *ctx, int in_klass, const void *lens, int64_t n, int32_t k, int32_t stride, int64_t *kmer_off);
int ssq_kmers64(ssq_ctx *ctx, int in_klass, const uint64_t *words, const int64_t *word_off, const void *lens, int64_t n,
                int32_t k, int32_t stride, const int64_t *kmer_off, uint64_t *out_words, uint8_t *out_lens);
/* Opt-in tolerant alphabet: out = ascii with a c g t rewritten to upper case and u U to T (the reference's table_91
 * maps U to T's code, util.pyx:44-50, but its validators reject U and lower case; the default stays exact A C G T).
 * Pack `out` with the offsets of `ascii`. */
int ssq_normalize(ssq_ctx *ctx, const uint8_t *ascii, int64_t nbytes, uint8_t *out);

/* ---- UMI collapse ------------------------------------------------------------------------
 * Clusters the distinct UMIs (ShortSeq64 words / lens with their counts, e.g. a counter's export) of every group
 * [group_off[g], group_off[g+1]) by Hamming distance, with UMI-tools' published rules (network.py; the reference itself
 * only compares `a ^ b` with UMI-tools' edit_distance, README.md:82-88, and sketches packed UMI objects,
 * shortseq/umi/umi.pxd:31-55).  method 0 "directional": a claims b when hamming(a, b) <= threshold and
 * count[a] >= 2 count[b] - 1, UMIs visited by decreasing count (ties: input order); method 1 "cluster": connected
 * components of hamming <= threshold.  UMIs of different lengths are never adjacent.
 * rep[i] = index (into the input arrays) of the representative of UMI i; n_clusters[g] (may be NULL) = clusters of group g. */
int ssq_umi_cluster(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, const uint64_t *counts, int64_t n,
                    const int64_t *group_off, int64_t n_groups, int32_t threshold, int32_t method, int64_t *rep,
                    int64_t *n_clusters);

/* ---- synthetic reads (measurement tooling) ----------------
 * Deterministic counter-based generator, identical to oracle/ssq_oracle.c's
 * ssq_oracle_synth_reads.  offsets[n+1] and ascii are outputs; ascii must hold
 * n*len_hi bytes. */
int ssq_synth_reads(ssq_ctx *ctx, uint64_t seed, int64_t first_read, int64_t n, int64_t n_keys,
                    int32_t len_lo, int32_t len_hi, int64_t *offsets, uint8_t *ascii);

#ifdef __cplusplus
}
#endif
#endif /* SHORTSEQ_B200_H */
