// ssq_pack.cu -- batched 2-bit packing with fused validation, and the fused front half of counting.
//
// Replaces, for a whole batch, the reference's per-object encoders
// (short_seq.pyx:54-74 -> short_seq_64.pyx:96-108, short_seq_192.pyx:103-108,
// short_seq_var.pyx:123-132 -> util.pyx:78-140) and validators (util.pxd:98-127).
//
// Two stages per tile of reads, both inside one kernel:
//   1. STREAM ENCODE.  The tile's reads are one contiguous byte range of the ASCII buffer.
//      Threads sweep it with coalesced, 16-byte aligned, L1-bypassing vector loads (four in
//      flight per thread); each 16-byte chunk becomes 32 bits of 2-bit codes in shared memory
//      and an invalid-byte indicator is OR-accumulated (exact {A,C,G,T} test, 4 bytes per op).
//   2. EXTRACT.  A read's packed words are just the bits [2*start, 2*(start+len)) of that code
//      stream: a few shared-memory loads and funnel shifts per 64-bit word, masked to the read
//      length.  ShortSeq64/192: threads own reads; ShortSeqVar: one warp per read, lane j owns
//      word j.  Words and lengths are written coalesced.
// Validation is consulted per tile: only when some byte near the tile is invalid do reads
// re-check their own bytes (rare slow path) to find the lowest failing read index.
//
// Counting modes of the fixed-class kernel:
//   kModeDirect   insert each key straight into the table (tables that fit in L2),
//   kModeScatter  append each key's 64-bit table key to one of 256 hash partitions -- every
//                 persistent CTA owns a private segment of every partition and stages keys in
//                 shared-memory rings that are flushed as whole 128-byte lines, no global
//                 atomics -- so that ssq_counter.cu can route the keys on to their table
//                 region and count them in shared memory.
#include <stdlib.h>
#include "ssq_internal.h"
#include "ssq_table.cuh"

namespace ssq {

constexpr int kPackThreads = 256;
constexpr int kRPT = 2;                               // reads per thread per tile (ShortSeq64)
constexpr int kTileReads = kPackThreads * kRPT;
constexpr int kLoadUnroll = 4;                        // 16-byte loads in flight per thread

// Tile shape of pack_fixed_kernel per class.  ShortSeq64: 512 reads (<= 16 KB of ASCII = 4 prefetched 16-byte chunks per
// thread).  ShortSeq192: 256 reads (<= 24 KB = 6 chunks per thread, all of them prefetched: with 512-read tiles more than
// half of a tile's chunks were loaded on demand inside the encode loop, one dependent DRAM round trip after the other --
// ncu showed the scatter-mode kernel waiting on them for 8 of every 9 issue slots at 2 CTAs per SM).
#ifndef SSQ_LU192
#define SSQ_LU192 6
#endif
template <int KLASS> struct FixedCfg {
    static constexpr int RPT = KLASS == SSQ_CLASS_64 ? kRPT : 1;
    static constexpr int TILE = kPackThreads * RPT;
    static constexpr int LU = KLASS == SSQ_CLASS_64 ? kLoadUnroll : SSQ_LU192;
    static constexpr int PARTS = KLASS == SSQ_CLASS_64 ? kParts : kParts192;
    static constexpr int RING = KLASS == SSQ_CLASS_64 ? kRingKeys : kRingKeys192;   // staging ring per partition, in 64-bit words
};

enum { kModePack = 0, kModeDirect = 1, kModeScatter = 2 };

struct PackArgs {
    const uint8_t *ascii;     // base the offsets index into (may be a virtual base for a staged slice)
    int64_t lo, hi;           // valid byte index range [lo, hi) of `ascii`
    const int64_t *offsets;
    int64_t n;                // reads in this launch
    int64_t index_base;       // index of read 0 of this launch within the caller's batch
    u64 *words;
    void *lens;
    const int64_t *word_off;  // ShortSeqVar only
    DevReport *rep;
};

// Load the 16-byte chunk that starts at byte index `idx` and may stick out of [lo, hi).
__device__ __noinline__ uint4 load_chunk_guarded(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t idx) {
    u32 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int64_t j = idx + 4 * k + b;
            u32 c = (j >= lo && j < hi) ? ascii[j] : (u32)'A';
            x |= c << (8 * b);
        }
        w[k] = x;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Geometry of one tile's byte range [t0, t0 + tile_bytes): chunk c covers the 16 bytes from byte index
// a0 + 16c, a0 = t0 rounded down to a 16-byte ADDRESS boundary, lead = t0 - a0.  Bytes of neighbouring
// tiles that share the edge chunks are encoded and validated too: their codes are never extracted, and a
// stray invalid byte there merely sends this tile through the exact per-read re-check.
struct TileGeom {
    const uint8_t *src;   // ascii + a0 (16-byte aligned address)
    int64_t a0;
    int nchunks;
    int lead;
    bool interior;        // every chunk lies inside [lo, hi): loads need no guard
};

__device__ __forceinline__ TileGeom tile_geom(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t t0, int tile_bytes) {
    TileGeom g;
    const int mis = (int)((uintptr_t)ascii & 15);
    g.lead = (int)((t0 + mis) & 15);
    g.a0 = t0 - g.lead;
    g.nchunks = (tile_bytes + g.lead + 15) >> 4;
    g.src = ascii + g.a0;
    g.interior = g.a0 >= lo && g.a0 + 16 * (int64_t)g.nchunks <= hi;
    return g;
}

// Issue the first kLoadUnroll rounds of 16-byte loads of an interior tile (results land in v[]).
template <int THREADS, int LU = kLoadUnroll>
__device__ __forceinline__ void issue_tile_loads(const TileGeom &g, uint4 (&v)[LU]) {
#pragma unroll
    for (int j = 0; j < LU; j++) {
        const int c = threadIdx.x + j * THREADS;
        if (c < g.nchunks) v[j] = ld_stream_v4(g.src + 16 * c);
    }
}

// Stage 1: codes[c] = 2-bit codes of chunk c.  `v` holds the prefetched first rounds when `prefetched`.
// `pad` extra words after the last chunk are zeroed.  The caller must __syncthreads().
template <int THREADS, int LU = kLoadUnroll>
__device__ __forceinline__ void encode_tile(const uint8_t *ascii, int64_t lo, int64_t hi, const TileGeom &g, bool prefetched,
                                            const uint4 (&v)[LU], u32 *codes, int pad, u32 &bad) {
    if (prefetched) {
#pragma unroll
        for (int j = 0; j < LU; j++) {
            const int c = threadIdx.x + j * THREADS;
            if (c < g.nchunks) codes[c] = encode16(v[j], bad);
        }
        for (int c = threadIdx.x + LU * THREADS; c < g.nchunks; c += THREADS)   // tiles longer than the prefetch window
            codes[c] = encode16(ld_stream_v4(g.src + 16 * c), bad);
    } else {                                                     // first / last tile of the buffer
        for (int c = threadIdx.x; c < g.nchunks; c += THREADS) {
            const int64_t idx = g.a0 + 16 * (int64_t)c;
            uint4 x = (idx >= lo && idx + 16 <= hi) ? ld_stream_v4(ascii + idx) : load_chunk_guarded(ascii, lo, hi, idx);
            codes[c] = encode16(x, bad);
        }
    }
    if ((int)threadIdx.x < pad) codes[g.nchunks + threadIdx.x] = 0;
}

// 64 bits of the code stream starting at bit `bit` (even, >= 0) of codes[].
__device__ __forceinline__ u64 extract64(const u32 *codes, int bit) {
    int wi = bit >> 5;
    u32 sh = (u32)bit & 31;
    u32 c0 = codes[wi], c1 = codes[wi + 1], c2 = codes[wi + 2];
    u32 lo = __funnelshift_r(c0, c1, sh);
    u32 hi = __funnelshift_r(c1, c2, sh);
    return ((u64)hi << 32) | lo;
}

// keep the low `nbits` bits (any int: <= 0 gives 0, >= 64 keeps all)
__device__ __forceinline__ u64 keep_bits(u64 x, int nbits) {
    if (nbits >= 64) return x;
    if (nbits <= 0) return 0;
    return x & ((1ull << nbits) - 1);
}

// Slow path: exact re-check of one read's bytes.
__device__ __noinline__ bool read_has_bad_base(const uint8_t *p, int len) {
    for (int j = 0; j < len; j++)
        if (!is_acgt(p[j])) return true;
    return false;
}

__device__ __noinline__ void report_len(DevReport *rep, int64_t len, u64 idx) {
    if (len > 1024) atomicMin(&rep->first_too_long, idx);
    else atomicMin(&rep->first_bad_len, idx);
}

// Offsets of one tile, held in registers so that the next tile's can be in flight during this one.
template <int RPT>
struct TileOffsets {
    int64_t t0, t1;
    int64_t start[RPT];
    int nreads;
};

template <int RPT>
__device__ __forceinline__ TileOffsets<RPT> load_tile_offsets(const int64_t *offsets, int64_t n, int64_t tile) {
    TileOffsets<RPT> o;
    const int64_t first = tile * (kPackThreads * RPT);
    o.nreads = first < n ? (int)min((int64_t)(kPackThreads * RPT), n - first) : 0;
    o.t0 = o.t1 = 0;
#pragma unroll
    for (int k = 0; k < RPT; k++) o.start[k] = 0;
    if (o.nreads > 0) {
        o.t0 = offsets[first];
        o.t1 = offsets[first + o.nreads];
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int r = threadIdx.x + k * kPackThreads;
            if (r < o.nreads) o.start[k] = offsets[first + r];
        }
    }
    return o;
}

// This thread's part of the test "the tile is kTileReads reads of exactly 32 bytes, the first one on a 16-byte address
// boundary, every 16-byte chunk prefetched" (fast path of pack_fixed_kernel; all threads must agree).
__device__ __forceinline__ bool tile_is_uniform32(const TileOffsets<kRPT> &o, const TileGeom &g, bool prefetched) {
    bool u = prefetched && o.nreads == kTileReads && g.lead == 0 && (o.t1 - o.t0) == (int64_t)kTileReads * 32;
#pragma unroll
    for (int k = 0; k < kRPT; k++) u = u && o.start[k] == o.t0 + 32 * (int64_t)(threadIdx.x + k * kPackThreads);
    return u;
}

// ---- ShortSeq64 / ShortSeq192 ------------------------------------------------------------------
// Persistent CTAs walk the tiles with a grid stride.  The loop is software-pipelined: while a tile is being
// extracted, the 16-byte loads of the CTA's next tile and the offsets of the one after are already in flight.
// Three resident CTAs per SM (<= 85 registers) measured best on B200: two lose latency hiding, four (<= 64
// registers) spill the prefetch state.
#ifndef SSQ_PACK_MIN_BLOCKS
#define SSQ_PACK_MIN_BLOCKS 3
#endif
// ShortSeq192 with counting fused in needs ~115 registers; capped at 80 (three CTAs per SM) ptxas spills part of the
// prefetch window, i.e. waits for the loads right where they are issued, and the kernel runs at the speed of one DRAM
// round trip per tile (2e8 x 75 nt: 13.5 ms); uncapped at two CTAs per SM the same kernel takes 7.6 ms.
template <int KLASS, int MODE>
__global__ void __launch_bounds__(kPackThreads, (KLASS == SSQ_CLASS_192 && MODE != kModePack) ? 2 : SSQ_PACK_MIN_BLOCKS) pack_fixed_kernel(PackArgs a, TableView t, PartView pv, const u64 *stop) {
    constexpr int MAXLEN = KLASS == SSQ_CLASS_64 ? 32 : 96;
    constexpr int MINLEN = KLASS == SSQ_CLASS_64 ? 0 : 33;
    constexpr int W = KLASS == SSQ_CLASS_64 ? 1 : 3;
    constexpr int PAD = 2 * W + 1;
    using Cfg = FixedCfg<KLASS>;
    constexpr int kRPT = Cfg::RPT, kTileReads = Cfg::TILE, kLoadUnroll = Cfg::LU, kParts = Cfg::PARTS;   // this class's tile shape
    constexpr int MAX_CHUNKS = (kTileReads * MAXLEN + 30) / 16 + 1;
    constexpr int RW = KLASS == SSQ_CLASS_64 ? 1 : 4;     // 64-bit words per staged record (table key / {w0, w1, w2, meta})
    __shared__ u32 codes[MAX_CHUNKS + PAD];
    __shared__ u32 srel[kTileReads + 1];                // read starts relative to the tile start
    __shared__ u32 s_new[kPackThreads / 32];
    // scatter mode: per-partition staging rings of this CTA (see Stager in ssq_table.cuh) in dynamic shared memory
    extern __shared__ __align__(16) u64 dyn_ring[];
    __shared__ u32 s_head[MODE == kModeScatter ? kParts : 1];
    __shared__ u32 s_tail[MODE == kModeScatter ? kParts : 1];
    constexpr int kRing = Cfg::RING;
    __shared__ u32 s_list[MODE == kModeScatter ? (kPackThreads / 32) * 32 * (kRing / kLineKeys) : 1];
    __shared__ u32 s_unstaged_new, s_ovf_n;
    __shared__ u32 s_uni[3];                            // fast-path votes, see the tile loop
    const Stager stg = make_stager(dyn_ring, s_head, s_tail, s_list);
    u64 *const seg0 = MODE == kModeScatter ? pv.keys + (size_t)blockIdx.x * kParts * pv.seg_cap * RW : nullptr;

    if (MODE != kModePack && stop != nullptr && *stop != 0) return;
    if (MODE == kModeScatter) {                      // ordered by the first tile's barrier
        stager_init<kParts>(stg);
        if (threadIdx.x == 0) { s_unstaged_new = 0; s_ovf_n = 0; }
    }

    u32 my_new = 0;
    u32 since_flush = 0;
    const int64_t ntiles = (a.n + kTileReads - 1) / kTileReads;
    const int64_t stride = gridDim.x;
    int64_t tile = blockIdx.x;

    // prologue: this tile's offsets and loads, the next tile's offsets
    TileOffsets<kRPT> cur = load_tile_offsets<kRPT>(a.offsets, a.n, tile);
    bool cur_ok = cur.nreads > 0 && cur.t0 >= a.lo && cur.t1 >= cur.t0 && cur.t1 <= a.hi &&
                  (cur.t1 - cur.t0) <= (int64_t)kTileReads * MAXLEN;
    TileGeom geom = tile_geom(a.ascii, a.lo, a.hi, cur.t0, cur_ok ? (int)(cur.t1 - cur.t0) : 0);
    uint4 v[kLoadUnroll];
    bool prefetched = cur_ok && geom.interior;
    if (prefetched) issue_tile_loads<kPackThreads, kLoadUnroll>(geom, v);
    TileOffsets<kRPT> nxt = load_tile_offsets<kRPT>(a.offsets, a.n, tile + stride);
    // FAST PATH (ShortSeq64): a full tile of 32-nt reads whose first byte sits on a 16-byte address boundary.  Chunk c
    // of the tile is then half (c & 1) of read c >> 1, so the prefetched 16-byte chunks are encoded in registers and a
    // read's two halves meet through one shuffle between neighbouring lanes: no code stream in shared memory, no
    // per-read offsets, no extraction.  Whether a tile qualifies is voted one tile ahead (its offsets are already in
    // registers) at the barrier every tile has anyway.
    const bool lens_aligned = ((uintptr_t)a.lens & 15) == 0;
    bool cur_fast = false;
    u32 vote_slot = 0;
    if constexpr (KLASS == SSQ_CLASS_64) {
        if (threadIdx.x == 0) s_uni[0] = 1;
        cur_fast = __syncthreads_and(tile_is_uniform32(cur, geom, prefetched)) != 0 && lens_aligned;
    }

    for (; tile < ntiles; tile += stride) {
        const int64_t first = tile * kTileReads;
        const int nreads = cur.nreads;
        const int64_t t0 = cur.t0;
        const int tile_bytes = cur_ok ? (int)(cur.t1 - cur.t0) : 0;
        const int lead = geom.lead;
        u32 bad = 0;
        u32 fc[4];                                                // fast path: codes of this thread's four chunks
        if (cur_fast) {
#pragma unroll
            for (int j = 0; j < 4; j++) fc[j] = encode16(v[j], bad);
        } else if (cur_ok) {
#pragma unroll
            for (int k = 0; k < kRPT; k++) {
                const int r = threadIdx.x + k * kPackThreads;
                // out-of-tile starts become an impossible value that fails the checks below
                if (r < nreads) srel[r] = (cur.start[k] >= t0 && cur.start[k] <= cur.t1) ? (u32)(cur.start[k] - t0) : 0xFFFFFFFFu;
            }
            if (threadIdx.x == 0) srel[nreads] = (u32)tile_bytes;
            encode_tile<kPackThreads, kLoadUnroll>(a.ascii, a.lo, a.hi, geom, prefetched, v, codes, PAD, bad);
        } else {
            // a tile whose byte range is inconsistent or larger than the staging buffer holds a read of the wrong
            // class (or offsets outside the buffer): report per read, pack nothing
#pragma unroll
            for (int k = 0; k < kRPT; k++) {
                const int r = threadIdx.x + k * kPackThreads;
                if (r < nreads) {
                    const int64_t len = a.offsets[first + r + 1] - cur.start[k];
                    if (len < MINLEN || len > MAXLEN) report_len(a.rep, len, (u64)(a.index_base + first + r));
                }
            }
            if (threadIdx.x == 0 && (t0 < a.lo || cur.t1 > a.hi || cur.t1 < t0))
                atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
        }

        // ---- prefetch: the next tile's bytes, the offsets of the tile after it
        const bool nxt_ok = nxt.nreads > 0 && nxt.t0 >= a.lo && nxt.t1 >= nxt.t0 && nxt.t1 <= a.hi &&
                            (nxt.t1 - nxt.t0) <= (int64_t)kTileReads * MAXLEN;
        const TileGeom ngeom = tile_geom(a.ascii, a.lo, a.hi, nxt.t0, nxt_ok ? (int)(nxt.t1 - nxt.t0) : 0);
        const bool nprefetched = nxt_ok && ngeom.interior;
        if (nprefetched) issue_tile_loads<kPackThreads, kLoadUnroll>(ngeom, v);
        const TileOffsets<kRPT> nxt2 = load_tile_offsets<kRPT>(a.offsets, a.n, tile + 2 * stride);

        // The barrier's own vote carries "some byte was invalid"; the vote on the next tile's uniformity rides along in one
        // of three rotating shared flags: a warp that disagrees clears the flag before the barrier, thread 0 re-arms the
        // flag of the next iteration (last read two barriers ago, next cleared only after this barrier).
        bool nxt_fast = false;
        if constexpr (KLASS == SSQ_CLASS_64) {
            if (threadIdx.x == 0) s_uni[(vote_slot + 1) % 3] = 1;
            if (!__all_sync(0xFFFFFFFFu, tile_is_uniform32(nxt, ngeom, nprefetched)) && (threadIdx.x & 31) == 0) s_uni[vote_slot] = 0;
        }
        const int tile_bad = __syncthreads_or(bad != 0);
        if constexpr (KLASS == SSQ_CLASS_64) {
            nxt_fast = s_uni[vote_slot] != 0 && lens_aligned;
            vote_slot = (vote_slot + 1) % 3;
        }

        bool flushed = false;
        if (cur_fast) {
            if constexpr (KLASS == SSQ_CLASS_64) {
                // lanes 2m / 2m+1 hold the low / high half of read (t >> 1) + 128 j in fc[j]; the even lane assembles the
                // reads of j = 0, 1, the odd lane those of j = 2, 3
                const bool odd = threadIdx.x & 1;
                const u32 x0 = __shfl_xor_sync(0xFFFFFFFFu, odd ? fc[0] : fc[2], 1);
                const u32 x1 = __shfl_xor_sync(0xFFFFFFFFu, odd ? fc[1] : fc[3], 1);
                u64 w2[2];
                w2[0] = odd ? ((u64)fc[2] << 32) | x0 : ((u64)x0 << 32) | fc[0];
                w2[1] = odd ? ((u64)fc[3] << 32) | x1 : ((u64)x1 << 32) | fc[1];
                const int r0 = (int)(threadIdx.x >> 1) + (odd ? 2 * (kPackThreads / 2) : 0);   // reads r0 and r0 + 128
                bool ok2[2] = {true, true};
                if (__any_sync(0xFFFFFFFFu, bad != 0)) {              // some byte of this warp's chunks is invalid: exact re-check
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int r = r0 + q * (kPackThreads / 2);
                        if (read_has_bad_base(a.ascii + t0 + 32 * r, 32)) {
                            ok2[q] = false;
                            atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + first + r));
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; q++) a.words[(size_t)first + r0 + q * (kPackThreads / 2)] = w2[q];
                if (threadIdx.x < kTileReads / 16)
                    reinterpret_cast<uint4 *>((uint8_t *)a.lens + first)[threadIdx.x] = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    if (MODE == kModeDirect && ok2[q]) {
                        bool is_new = false;
                        insert64(t, w2[q], 32u, 1ull, is_new);
                        my_new += is_new ? 1u : 0u;
                    }
                    if constexpr (MODE == kModeScatter) {
                        if (ok2[q]) {
                            const u64 h2 = table_hash64(w2[q], t.rot);
                            const u64 key = key64_of(h2, 32u);
                            if (!stage_key(stg, (u32)(h2 >> 56), key)) insert64_slow(t, h2, key, &s_unstaged_new);
                        }
                    }
                }
            }
        } else if (cur_ok) {
#pragma unroll
            for (int k = 0; k < kRPT; k++) {
                const int r = threadIdx.x + k * kPackThreads;
                const int64_t i = first + r;
                u64 w[W];
                bool ok = false;
                int len = 0;
                if (r < nreads) {
                    const u32 r0 = srel[r], r1 = srel[r + 1];
                    len = (int)(r1 - r0);
                    const bool len_ok = r0 <= (u32)tile_bytes && r1 <= (u32)tile_bytes && len >= MINLEN && len <= MAXLEN;
                    if (len_ok) {
                        const int bit = 2 * ((int)r0 + lead);
#pragma unroll
                        for (int j = 0; j < W; j++) w[j] = keep_bits(extract64(codes, bit + 64 * j), 2 * len - 64 * j);
                        ok = !(tile_bad && read_has_bad_base(a.ascii + t0 + r0, len));
                        if (!ok) atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
                    } else {
#pragma unroll
                        for (int j = 0; j < W; j++) w[j] = 0;
                        report_len(a.rep, (int64_t)a.offsets[i + 1] - a.offsets[i], (u64)(a.index_base + i));
                    }
#pragma unroll
                    for (int j = 0; j < W; j++) a.words[(size_t)i * W + j] = w[j];
                    ((uint8_t *)a.lens)[i] = len_ok ? (uint8_t)len : 0;
                }
                if (MODE == kModeDirect && ok) {
                    bool is_new = false;
                    if constexpr (KLASS == SSQ_CLASS_64) insert64(t, w[0], (u32)len, 1ull, is_new);
                    else insert192(t, w[0], w[1], w[2], (u32)len, 1ull, is_new);
                    my_new += is_new ? 1u : 0u;
                }
                if constexpr (MODE == kModeScatter) {
                    if (ok) {
                        // stage the table key / record for its hash partition; a full staging ring: count it right away
                        if constexpr (KLASS == SSQ_CLASS_64) {
                            const u64 h2 = table_hash64(w[0], t.rot);
                            const u64 key = key64_of(h2, (u32)len);
                            if (!stage_key(stg, (u32)(h2 >> 56), key)) insert64_slow(t, h2, key, &s_unstaged_new);
                        } else {
#ifdef SSQ_X_NOHASH
                            const u64 h2 = (w[0] ^ w[1] ^ w[2]) * 0x9E3779B97F4A7C15ull;
#else
                            const u64 h2 = rotl64(hash192(w[0], w[1], w[2], (u32)len), t.rot);
#endif
                            const u64 meta = meta192_of(h2, (u32)len);
#ifdef SSQ_X_NOSTAGE
                            if (meta == 0x1234567ull) a.words[0] = meta;
                            if (false && !stage_rec192<kRing>(stg, (u32)(h2 >> (kParts192 == 128 ? 57 : 56)), w[0], w[1], w[2], meta)) {
                                const u32 pos = atomicAdd(&s_ovf_n, 1u);         // ring full: park the record in the overflow segment
                                if (pos < pv.ovf_cap) {
                                    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(pv.ovf + ((size_t)blockIdx.x * pv.ovf_cap + pos) * 4);
                                    dst[0] = make_ulonglong2(w[0], w[1]);
                                    dst[1] = make_ulonglong2(w[2], meta);
                                } else {
                                    insert192_slow(t, h2, w[0], w[1], w[2], (u32)len, &s_unstaged_new);
                                }
                            }
#else
                            if (!stage_rec192<kRing>(stg, (u32)(h2 >> (kParts192 == 128 ? 57 : 56)), w[0], w[1], w[2], meta)) {
                                const u32 pos = atomicAdd(&s_ovf_n, 1u);         // ring full: park the record in the overflow segment
                                if (pos < pv.ovf_cap) {
                                    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(pv.ovf + ((size_t)blockIdx.x * pv.ovf_cap + pos) * 4);
                                    dst[0] = make_ulonglong2(w[0], w[1]);
                                    dst[1] = make_ulonglong2(w[2], meta);
                                } else {
                                    insert192_slow(t, h2, w[0], w[1], w[2], (u32)len, &s_unstaged_new);
                                }
                            }
#endif
                        }
                    }
                }
            }
        }
        if constexpr (MODE == kModeScatter) {
            // every flush_every-th tile: move the complete 128-byte lines of the staging rings to global memory
            if ((cur_fast || cur_ok) && ++since_flush >= pv.flush_every) {
                since_flush = 0;
                flushed = true;
                __syncthreads();
                flush_lines<false, RW, kParts, kRing>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
            }
        }
        // codes[] / srel[] are rewritten by the next tile, and the rings are staged into again after a flush; a fast-path
        // tile touched neither codes[] nor srel[]
        if (!cur_fast || flushed) __syncthreads();

        cur = nxt; cur_ok = nxt_ok; geom = ngeom; prefetched = nprefetched; nxt = nxt2; cur_fast = nxt_fast;
    }
    if constexpr (MODE == kModeScatter) {
        __syncthreads();   // a fast-path tile ends without a barrier
        flush_lines<true, RW, kParts, kRing>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
        __syncthreads();
        for (int p = threadIdx.x; p < kParts; p += kPackThreads)
            pv.seg_count[(size_t)blockIdx.x * kParts + p] = stager_seg_count(stg, p, pv.seg_cap);
        if (threadIdx.x == 0) {
            my_new += s_unstaged_new;
            if (KLASS != SSQ_CLASS_64) pv.ovf_count[blockIdx.x] = min(s_ovf_n, pv.ovf_cap);
        }
    }
    if (MODE != kModePack) {
        // one size update per CTA: a single global counter cannot take one atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
        if ((threadIdx.x & 31) == 0) s_new[threadIdx.x >> 5] = my_new;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 tot = 0;
            for (int k = 0; k < kPackThreads / 32; k++) tot += s_new[k];
            if (tot) atomicAdd(t.size, tot);
        }
    }
}

// ---- ShortSeq64, uniform batches of 32-nt reads ----------------------------------------------------
// The C2 shape (1e9 x 32 nt): offsets[i] = offsets[0] + 32 i and the first base sits on a 16-byte address boundary.
// pack_fixed_kernel spends most of its instructions on what such a batch does not need (tile geometry, a code stream in
// shared memory, per-read extraction, 64-bit offset arithmetic).  Here a WARP owns tiles of 64 reads = 2048 contiguous
// bytes = four fully coalesced 512-byte loads (a lane's 16-byte chunk j is half (lane & 1) of read 16 j + lane / 2);
// the chunks are encoded in registers, neighbouring lanes swap halves with two shuffles and every lane ends up with two
// whole reads.  The tile's 65 offsets are checked against the arithmetic progression by the same warp (a vote, no
// barrier); a tile that fails the check, and the ragged last tile, take a plain lane-per-read path with the general
// kernel's semantics.  The CTA only synchronises to flush the staging rings.  The host picks this kernel when the
// batch holds exactly 32 n bytes (launch_fixed); everything else is decided here, tile by tile.
constexpr int kWarpTileReads = 64;
#ifndef SSQ_PACK32_FLUSH
#define SSQ_PACK32_FLUSH 4
#endif
constexpr int kPack32FlushEvery = SSQ_PACK32_FLUSH;   // warp tiles between two flushes: 8 warps x 4 x 64 = 2048 keys, as pack_fixed_kernel

// One read the slow way (any length, any alignment): returns false when the read must not be counted.
__device__ __noinline__ bool pack32_slow_read(const PackArgs &a, int64_t i, u64 &word, u32 &len_out) {
    const int64_t o0 = a.offsets[i], o1 = a.offsets[i + 1];
    const int64_t len = o1 - o0;
    word = 0;
    len_out = 0;
    if (o0 < a.lo || o1 > a.hi || len < 0) { atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + i)); a.words[i] = 0; ((uint8_t *)a.lens)[i] = 0; return false; }
    if (len > 32) { report_len(a.rep, len, (u64)(a.index_base + i)); a.words[i] = 0; ((uint8_t *)a.lens)[i] = 0; return false; }
    bool bad = false;
    for (int k = 0; k < (int)len; k++) {
        const u32 ch = a.ascii[o0 + k];
        bad |= !is_acgt((uint8_t)ch);
        word |= (u64)((ch >> 1) & 3u) << (2 * k);
    }
    a.words[i] = word;
    ((uint8_t *)a.lens)[i] = (uint8_t)len;
    len_out = (u32)len;
    if (bad) atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
    return !bad;
}

#ifndef SSQ_PACK32_MIN_BLOCKS
#define SSQ_PACK32_MIN_BLOCKS 3      /* development: 4 (<= 64 registers) needs 32 KB staging rings (SSQ_LINE_KEYS=8) */
#endif
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : SSQ_PACK32_MIN_BLOCKS) pack32_kernel(PackArgs a, TableView t, PartView pv, const u64 *stop) {
    extern __shared__ __align__(16) u64 dyn_ring[];
    __shared__ u32 s_head[MODE == kModeScatter ? kParts : 1];
    __shared__ u32 s_tail[MODE == kModeScatter ? kParts : 1];
    __shared__ u32 s_list[MODE == kModeScatter ? (THREADS / 32) * 64 : 1];
    __shared__ u32 s_new[THREADS / 32];
    __shared__ u32 s_unstaged_new;
    const Stager stg = make_stager(dyn_ring, s_head, s_tail, s_list);
    u64 *const seg0 = MODE == kModeScatter ? pv.keys + (size_t)blockIdx.x * kParts * pv.seg_cap : nullptr;
    if (MODE != kModePack && stop != nullptr && *stop != 0) return;
    if (MODE == kModeScatter) {
        stager_init(stg);
        if (threadIdx.x == 0) s_unstaged_new = 0;
        __syncthreads();
    }
    constexpr int kWarps = THREADS / 32;
    constexpr int kFlushEvery = kPack32FlushEvery * 8 / kWarps;          // 2048 keys between two flushes whatever the CTA size
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool odd = lane & 1;
    const int64_t ntiles = (a.n + kWarpTileReads - 1) / kWarpTileReads;
    const int64_t wstride = (int64_t)gridDim.x * kWarps;
    const int64_t cta_first = (int64_t)blockIdx.x * kWarps;
    const int iters = cta_first < ntiles ? (int)((ntiles - cta_first + wstride - 1) / wstride) : 0;   // trip count of warp 0: CTA-uniform
    const int64_t o0 = a.offsets[0];
    const uint8_t *const base = a.ascii + o0;                 // read i starts at base + 32 i in a uniform batch
    const bool aligned = ((uintptr_t)base & 15) == 0 && ((uintptr_t)a.lens & 15) == 0 && o0 >= a.lo;
    u32 my_new = 0;

    // prefetch state of the next tile: its four chunks and the offsets this lane checks
    uint4 v[4];
    int64_t chk[3];
    bool pre_full = false;
    auto prefetch = [&](int64_t wt) {
        const int64_t first = wt * kWarpTileReads;
        pre_full = wt < ntiles && aligned && first + kWarpTileReads <= a.n && o0 + 32 * (first + kWarpTileReads) <= a.hi;
        if (pre_full) {
            const uint8_t *src = base + 32 * first + 16 * lane;
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = ld_stream_v4(src + 512 * j);
            chk[0] = a.offsets[first + lane];
            chk[1] = a.offsets[first + 32 + lane];
            chk[2] = lane == 0 ? a.offsets[first + kWarpTileReads] : 0;
        }
    };
    int64_t wt = cta_first + warp;
    prefetch(wt);
    for (int it = 0; it < iters; it++, wt += wstride) {
        const int64_t first = wt * kWarpTileReads;
        bool fast = pre_full;
        u32 fc[4];
        u32 bad = 0;
        if (fast) {
            const int64_t e0 = o0 + 32 * (first + lane);
            bool uni = chk[0] == e0 && chk[1] == e0 + 32 * 32 && (lane != 0 || chk[2] == e0 + 32 * kWarpTileReads);
            fast = __all_sync(0xFFFFFFFFu, uni);
#pragma unroll
            for (int j = 0; j < 4; j++) fc[j] = encode16(v[j], bad);
        }
        prefetch(wt + wstride);
        if (fast) {
            // lanes 2m / 2m+1 hold the low / high half of read 16 j + m in fc[j]; the even lane assembles the reads of
            // j = 0, 1 (reads m, 16 + m), the odd lane those of j = 2, 3 (reads 32 + m, 48 + m)
            const u32 x0 = __shfl_xor_sync(0xFFFFFFFFu, odd ? fc[0] : fc[2], 1);
            const u32 x1 = __shfl_xor_sync(0xFFFFFFFFu, odd ? fc[1] : fc[3], 1);
            u64 w2[2];
            w2[0] = odd ? ((u64)fc[2] << 32) | x0 : ((u64)x0 << 32) | fc[0];
            w2[1] = odd ? ((u64)fc[3] << 32) | x1 : ((u64)x1 << 32) | fc[1];
            const u32 r0 = (lane >> 1) + (odd ? 32u : 0u);               // reads r0 and r0 + 16 of the tile
            bool ok2[2] = {true, true};
            if (__any_sync(0xFFFFFFFFu, bad != 0)) {                      // some byte of the tile is invalid: exact re-check
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int64_t i = first + r0 + 16 * q;
                    if (read_has_bad_base(base + 32 * i, 32)) {
                        ok2[q] = false;
                        atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
                    }
                }
            }
            u64 *wdst = a.words + first + r0;
            wdst[0] = w2[0];
            wdst[16] = w2[1];
            if (lane < kWarpTileReads / 16)
                reinterpret_cast<uint4 *>((uint8_t *)a.lens + first)[lane] = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
            if constexpr (MODE == kModeDirect) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    if (ok2[q]) {
                        bool is_new = false;
                        insert64(t, w2[q], 32u, 1ull, is_new);
                        my_new += is_new ? 1u : 0u;
                    }
                }
            }
            if constexpr (MODE == kModeScatter) {
                u64 h2[2], key[2];
                u32 part[2];
                bool staged[2];
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    h2[q] = table_hash64(w2[q], t.rot);
                    key[q] = key64_of(h2[q], 32u);
                    part[q] = (u32)(h2[q] >> 56);
                }
#ifdef SSQ_X_NOSTAGE      /* development: the keys leave coalesced, no staging at all (what the scatter itself costs) */
                pv.keys[first + r0] = key[0] + part[0];
                pv.keys[first + r0 + 16] = key[1] + part[1];
                staged[0] = staged[1] = true;
#else
                stage_keys<2>(stg, part, key, ok2, staged);
#pragma unroll
                for (int q = 0; q < 2; q++)
                    if (!staged[q]) insert64_slow(t, h2[q], key[q], &s_unstaged_new);
#endif
            }
        } else if (wt < ntiles) {
            // a tile that is not 64 aligned 32-nt reads (or the ragged end of the batch): lane-per-read
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                const int64_t i = first + lane + 32 * q;
                if (i >= a.n) continue;
                u64 word;
                u32 len;
                if (!pack32_slow_read(a, i, word, len)) continue;
                if (MODE == kModeDirect) {
                    bool is_new = false;
                    insert64(t, word, len, 1ull, is_new);
                    my_new += is_new ? 1u : 0u;
                }
                if constexpr (MODE == kModeScatter) {
                    const u64 h2 = table_hash64(word, t.rot);
                    const u64 key = key64_of(h2, len);
                    if (!stage_key(stg, (u32)(h2 >> 56), key)) insert64_slow(t, h2, key, &s_unstaged_new);
                }
            }
        }
        if constexpr (MODE == kModeScatter) {
            if ((it % kFlushEvery) == kFlushEvery - 1) {
                __syncthreads();
                flush_lines<false, 1>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
                __syncthreads();
            }
        }
    }
    if constexpr (MODE == kModeScatter) {
        __syncthreads();
        flush_lines<true, 1>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
        __syncthreads();
        for (int p = threadIdx.x; p < kParts; p += THREADS)
            pv.seg_count[(size_t)blockIdx.x * kParts + p] = stager_seg_count(stg, p, pv.seg_cap);
        if (threadIdx.x == 0) my_new += s_unstaged_new;
    }
    if (MODE != kModePack) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
        if (lane == 0) s_new[warp] = my_new;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 tot = 0;
            for (int k = 0; k < kWarps; k++) tot += s_new[k];
            if (tot) atomicAdd(t.size, tot);
        }
    }
}

// ---- ShortSeqVar: one warp per read, lane j owns word j ----------------------------------
constexpr int kVarTileReads = 32;
constexpr int kVarMaxChunks = (kVarTileReads * 1024 + 30) / 16 + 1;

__global__ void __launch_bounds__(kPackThreads) pack_var_kernel(PackArgs a) {
    constexpr int PAD = 3;
    __shared__ u32 codes[kVarMaxChunks + PAD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (a.n + kVarTileReads - 1) / kVarTileReads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kVarTileReads;
        const int nreads = (int)min((int64_t)kVarTileReads, a.n - first);
        const int64_t t0 = a.offsets[first];
        const int64_t t1 = a.offsets[first + nreads];
        // lane r of warp 0 checks read r's length
        if (warp == 0 && lane < nreads) {
            int64_t len = a.offsets[first + lane + 1] - a.offsets[first + lane];
            if (len < 97 || len > 1024) report_len(a.rep, len, (u64)(a.index_base + first + lane));
        }
        const bool tile_ok = t0 >= a.lo && t1 >= t0 && t1 <= a.hi && (t1 - t0) <= (int64_t)kVarTileReads * 1024;
        if (!tile_ok) {
            if (threadIdx.x == 0 && (t0 < a.lo || t1 > a.hi || t1 < t0))
                atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
            continue;
        }
        u32 bad = 0;
        const TileGeom geom = tile_geom(a.ascii, a.lo, a.hi, t0, (int)(t1 - t0));
        const int lead = geom.lead;
        uint4 v[kLoadUnroll];
        if (geom.interior) issue_tile_loads<kPackThreads>(geom, v);
        encode_tile<kPackThreads>(a.ascii, a.lo, a.hi, geom, geom.interior, v, codes, PAD, bad);
        const int tile_bad = __syncthreads_or(bad != 0);
        for (int r = warp; r < nreads; r += kPackThreads / 32) {
            const int64_t i = first + r;
            const int64_t o0 = a.offsets[i], o1 = a.offsets[i + 1];
            const int64_t len = o1 - o0;
            const bool len_ok = len >= 97 && len <= 1024 && o0 >= t0 && o1 <= t1;
            if (len_ok) {
                const int nwords = (int)((len + 31) >> 5);
                if (lane < nwords) {
                    u64 w = keep_bits(extract64(codes, 2 * ((int)(o0 - t0) + lead) + 64 * lane), 2 * (int)len - 64 * lane);
                    a.words[a.word_off[i] + lane] = w;
                }
                if (tile_bad) {
                    bool b = false;
                    for (int64_t j = lane; j < len; j += 32) b |= !is_acgt(a.ascii[o0 + j]);
                    if (__any_sync(0xFFFFFFFFu, b) && lane == 0) atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
                }
            }
            if (lane == 0) ((uint16_t *)a.lens)[i] = len_ok ? (uint16_t)len : 0;
        }
        __syncthreads();
    }
}

// ---- ShortSeqVar, bulk-copy pipeline --------------------------------------------------------------
// pack_var_kernel above gives a read to a warp and a word to a lane: a 150-nt read keeps 5 of 32 lanes busy, and every
// tile pays three dependent global round trips (tile bounds -> bytes -> per-read offsets) with nothing in flight
// behind them (ncu: 7 of 8 issue slots waiting on the long scoreboard, 1.8 TB/s on the 150/300/1000-nt mix).
// Here a persistent CTA keeps kVar2Stages tiles of 32 reads in flight as 1-D bulk copies (cp.async.bulk, SASS UBLKCP):
// one lane computes the tile's 16-byte aligned byte range from offsets it loaded one tile earlier, arms the stage's
// mbarrier with the byte count and issues the copy -- no registers, no issue slots while the bytes travel.  Consumers
// wait on the barrier, encode the raw bytes from shared memory into the 2-bit code stream (4:1), and then extract the
// tile's words as ONE flat list (thread k -> word k of the tile, its read found by a 5-step search in the tile's 33
// word offsets): every lane has work whatever the read lengths, and the word stores are fully coalesced.
#ifndef SSQ_VAR2_READS
#define SSQ_VAR2_READS 32
#endif
#ifndef SSQ_VAR2_STAGES
#define SSQ_VAR2_STAGES 2
#endif
#ifndef SSQ_VAR2_CTAS
#define SSQ_VAR2_CTAS 3
#endif
#ifndef SSQ_VAR2_CODES
#define SSQ_VAR2_CODES 1
#endif
constexpr int kVar2Codes = SSQ_VAR2_CODES;             // code-stream buffers: 2 = one barrier per tile, 1 = two barriers, 8 KB less
constexpr int kVar2Reads = SSQ_VAR2_READS;             // <= 32: lane r of warp 0 owns read r's offsets
constexpr int kVar2Log2Reads = kVar2Reads > 16 ? 5 : (kVar2Reads > 8 ? 4 : 3);
constexpr int kVar2Stages = SSQ_VAR2_STAGES;
constexpr int kVar2Meta = kVar2Stages + 1;             // metadata slots: a tile's slot is rewritten only after its extraction
constexpr int kVar2RawBytes = kVar2Reads * 1024 + 32;  // + lead (< 16) + round-up of the tail
constexpr int kVar2BadCap = 96;                        // invalid chunks remembered per tile (more: the reads are re-read)
constexpr int kVar2MaxChunks = kVar2RawBytes / 16;

struct Var2Meta {
    int64_t t0, wbase;
    u32 srel[kVar2Reads + 1];   // read starts relative to t0 (0xFFFFFFFF: outside the tile)
    u32 wrel[kVar2Reads + 1];   // first word of each read relative to wbase
    int nreads, nchunks, lead, mode;   // mode 0: inconsistent tile (skipped), 1: bytes arrive by bulk copy, 2: edge tile, guarded loads
    int64_t a0;
};

struct Var2Smem {
    uint8_t raw[kVar2Stages][kVar2RawBytes];
    u32 codes[kVar2Codes][kVar2MaxChunks + 4];
    Var2Meta meta[kVar2Meta];
    u64 full[kVar2Stages];       // producer -> consumers: tile published (metadata written, bytes landed)
    u64 enc_done[kVar2Stages];   // consumers -> producer: raw[s] has been encoded, the stage may be refilled
    u64 ext_done[kVar2Meta];     // consumers -> producer: the tile of this metadata slot has been extracted
    // chunks that hold an invalid byte, as (chunk << 16 | 16-bit byte mask); one list per tile parity.  With it the
    // failing READ is found from shared memory; the first version re-read every read of such a tile from global memory
    // byte by byte (1 % bad reads: 18.7 ms instead of 7.0 ms per 5e7 reads).
    u32 bad_n[2];
    u32 bad_list[2][kVar2BadCap];
};

// 8 consumer warps + 1 producer warp.  The producer owns everything that is per tile and serial -- the offsets (loaded two
// tiles ahead), the tile geometry, the lengths / length errors, the metadata slot, arming the mbarrier and issuing the
// bulk copy -- and runs up to kVar2Stages tiles ahead of the consumers, which only wait on the stage's `full` barrier.
// In the first bulk-copy version warp 0 did that work between the consumers' barriers: ~150 instructions on top of a
// warp's ~175 per tile, so every tile waited for warp 0 (ncu: 31 % of the stall samples at that barrier).
constexpr int kVar2Consumers = kPackThreads;               // 256 threads extract
constexpr int kVar2Threads = kVar2Consumers + 32;          // + the producer warp

// Rare path of the encode loop: chunk c holds at least one byte outside {A,C,G,T}; remember which bytes.
static __device__ __noinline__ void note_bad_chunk(uint4 x, int c, u32 *n, u32 *list) {
    const u32 w[4] = {x.x, x.y, x.z, x.w};
    u32 mask = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) mask |= (u32)(!is_acgt((uint8_t)(w[k >> 2] >> (8 * (k & 3))))) << k;
    const u32 pos = atomicAdd(n, 1u);
    if (pos < (u32)kVar2BadCap) list[pos] = ((u32)c << 16) | mask;
}

__global__ void __launch_bounds__(kVar2Threads, SSQ_VAR2_CTAS) pack_var2_kernel(PackArgs a) {
    extern __shared__ __align__(128) uint8_t var2_dyn[];
    Var2Smem &sm = *reinterpret_cast<Var2Smem *>(var2_dyn);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (a.n + kVar2Reads - 1) / kVar2Reads;
    const int64_t stride = gridDim.x;
    const int mytiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + stride - 1) / stride) : 0;
    if (threadIdx.x == 0) {
        sm.bad_n[0] = sm.bad_n[1] = 0;
#pragma unroll
        for (int s = 0; s < kVar2Stages; s++) { mbar_init(smem_addr(&sm.full[s]), 1); mbar_init(smem_addr(&sm.enc_done[s]), kVar2Consumers / 32); }
#pragma unroll
        for (int m = 0; m < kVar2Meta; m++) mbar_init(smem_addr(&sm.ext_done[m]), kVar2Consumers / 32);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kVar2Consumers / 32) {
        // =========================================== producer warp ===========================================
        const u64 drop = l2_policy_evict_first();
        struct MetaRegs { int64_t p_off = 0, p_off_last = 0, p_woff = 0, p_woff_last = 0; } regs_a, regs_b;
        auto fetch_meta = [&](int j, MetaRegs &R) {     // offsets / word offsets of this CTA's j-th tile into registers
            if (j >= mytiles) return;
            const int64_t first = ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads;
            const int nreads = (int)min((int64_t)kVar2Reads, a.n - first);
            if (lane <= nreads) { R.p_off = a.offsets[first + lane]; R.p_woff = a.word_off[first + lane]; }
            if (lane == 0) { R.p_off_last = a.offsets[first + nreads]; R.p_woff_last = a.word_off[first + nreads]; }
        };
        auto issue = [&](int j, const MetaRegs &R) {    // publish tile j (its offsets are in R) and start its copy
            const int64_t p_off = R.p_off, p_off_last = R.p_off_last, p_woff = R.p_woff, p_woff_last = R.p_woff_last;
            const int s = j % kVar2Stages, ms = j % kVar2Meta;
            // the stage was last filled for tile j - kVar2Stages, the metadata slot last used by tile j - kVar2Meta
            if (j >= kVar2Stages) mbar_wait(smem_addr(&sm.enc_done[s]), (u32)(j / kVar2Stages - 1) & 1u);
            if (j >= kVar2Meta) mbar_wait(smem_addr(&sm.ext_done[ms]), (u32)(j / kVar2Meta - 1) & 1u);
            const int64_t first = ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads;
            const int nreads = (int)min((int64_t)kVar2Reads, a.n - first);
            Var2Meta &m = sm.meta[ms];
            const int64_t t0 = __shfl_sync(0xFFFFFFFFu, p_off, 0), wbase = __shfl_sync(0xFFFFFFFFu, p_woff, 0);
            const int64_t t1 = __shfl_sync(0xFFFFFFFFu, p_off_last, 0), wend = __shfl_sync(0xFFFFFFFFu, p_woff_last, 0);
            int64_t nxt = __shfl_down_sync(0xFFFFFFFFu, p_off, 1);
            if (lane == nreads - 1) nxt = t1;
            const bool tile_ok = t0 >= a.lo && t1 >= t0 && t1 <= a.hi && (t1 - t0) <= (int64_t)kVar2Reads * 1024 &&
                                 wend >= wbase && wend - wbase <= (int64_t)kVar2Reads * 32;
            if (lane < nreads) {
                const int64_t len = nxt - p_off;
                const bool len_ok = len >= 97 && len <= 1024 && p_off >= t0 && nxt <= t1;
                if (len < 97 || len > 1024) report_len(a.rep, len, (u64)(a.index_base + first + lane));
                ((uint16_t *)a.lens)[first + lane] = (tile_ok && len_ok) ? (uint16_t)len : 0;
            }
            if (lane <= nreads) {
                const int64_t o = lane == nreads ? t1 : p_off, w = lane == nreads ? wend : p_woff;
                m.srel[lane] = (o >= t0 && o <= t1) ? (u32)(o - t0) : 0xFFFFFFFFu;
                m.wrel[lane] = (u32)min(max(w - wbase, (int64_t)0), (int64_t)kVar2Reads * 32);
            }
            if (lane == 0 && nreads == kVar2Reads) {    // entry 32 belongs to lane 0 (a warp has 32 lanes, a tile 33 boundaries)
                m.srel[nreads] = (u32)(t1 - t0);
                m.wrel[nreads] = (u32)min(max(wend - wbase, (int64_t)0), (int64_t)kVar2Reads * 32);
            }
            __syncwarp();                               // every lane's metadata stores precede lane 0's arrive (release)
            if (lane == 0) {
                int mode = 0;
                TileGeom g = tile_geom(a.ascii, a.lo, a.hi, t0, tile_ok ? (int)(t1 - t0) : 0);
                if (tile_ok) mode = (g.interior && g.nchunks > 0) ? 1 : 2;
                else if (t0 < a.lo || t1 > a.hi || t1 < t0) atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
                m.t0 = t0; m.wbase = wbase; m.nreads = nreads; m.nchunks = g.nchunks; m.lead = g.lead; m.mode = mode; m.a0 = g.a0;
                const u32 bar = smem_addr(&sm.full[s]);
                if (mode == 1) {
                    mbar_expect_tx(bar, 16u * (u32)g.nchunks);
                    bulk_g2s(smem_addr(sm.raw[s]), g.src, 16u * (u32)g.nchunks, bar, drop);
                } else {
                    mbar_arrive(bar);                   // nothing travels: the tile is complete as published
                }
            }
            __syncwarp();
        };
        fetch_meta(0, regs_a);
        fetch_meta(1, regs_b);
        for (int j = 0; j < mytiles; j += 2) {
            issue(j, regs_a);
            fetch_meta(j + 2, regs_a);
            if (j + 1 < mytiles) {
                issue(j + 1, regs_b);
                fetch_meta(j + 3, regs_b);
            }
        }
        return;
    }

    // ================================================ consumer warps ================================================
    for (int j = 0; j < mytiles; j++) {
        const int s = j % kVar2Stages, ms = j % kVar2Meta;
        u32 *const codes = sm.codes[j % kVar2Codes];
        mbar_wait(smem_addr(&sm.full[s]), (u32)(j / kVar2Stages) & 1u);
        const Var2Meta &m = sm.meta[ms];
        const int mode = m.mode, nchunks = m.nchunks, lead = m.lead, nreads = m.nreads;
        const int64_t t0 = m.t0;
        u32 bad = 0;
        if (mode == 1) {
            const u32 raw = smem_addr(sm.raw[s]);
            for (int c = threadIdx.x; c < nchunks; c += kVar2Consumers) codes[c] = encode16(lds_v4(raw + 16 * c), bad);
        } else if (mode == 2) {
            for (int c = threadIdx.x; c < nchunks; c += kVar2Consumers) {
                const int64_t idx = m.a0 + 16 * (int64_t)c;
                const uint4 x = (idx >= a.lo && idx + 16 <= a.hi) ? ld_stream_v4(a.ascii + idx) : load_chunk_guarded(a.ascii, a.lo, a.hi, idx);
                codes[c] = encode16(x, bad);
            }
        }
        if (bad) {
            // rare: one of this thread's chunks holds an invalid byte -- look at them again (raw[s] is still ours: this warp
            // has not signalled enc_done yet) and put the offending chunks on the tile's list
            for (int c = threadIdx.x; c < nchunks; c += kVar2Consumers) {
                const int64_t idx = m.a0 + 16 * (int64_t)c;
                const uint4 x = mode == 1 ? lds_v4(smem_addr(sm.raw[s]) + 16 * c)
                                          : ((idx >= a.lo && idx + 16 <= a.hi) ? ld_stream_v4(a.ascii + idx) : load_chunk_guarded(a.ascii, a.lo, a.hi, idx));
                u32 badc = 0;
                (void)encode16(x, badc);
                if (badc) note_bad_chunk(x, c, &sm.bad_n[j & 1], sm.bad_list[j & 1]);
            }
        }
        if (threadIdx.x < 4) codes[nchunks + threadIdx.x] = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&sm.enc_done[s]));        // this warp no longer reads raw[s]
        const bool tile_bad = named_bar_or(1, kVar2Consumers, bad != 0);    // codes complete
        if (threadIdx.x == 0) sm.bad_n[(j + 1) & 1] = 0;                    // the next tile's list (last read two barriers ago)
        if (mode != 0) {
            const u32 tw = m.wrel[nreads];
            u64 *wdst = a.words + m.wbase;
            for (u32 k = threadIdx.x; k < tw; k += kVar2Consumers) {
                int lo_ = 0, hi_ = nreads;                    // wrel[lo_] <= k < wrel[hi_]
#pragma unroll
                for (int it = 0; it < kVar2Log2Reads; it++) {
                    const int mid = (lo_ + hi_) >> 1;
                    if (m.wrel[mid] <= k) lo_ = mid; else hi_ = mid;
                }
                const u32 r0 = m.srel[lo_], r1 = m.srel[lo_ + 1];
                const int len = (int)(r1 - r0), jw = (int)(k - m.wrel[lo_]);
                const bool ok = r0 != 0xFFFFFFFFu && r1 != 0xFFFFFFFFu && len >= 97 && len <= 1024;
                wdst[k] = ok ? keep_bits(extract64(codes, 2 * ((int)r0 + lead) + 64 * jw), 2 * len - 64 * jw) : 0ull;
            }
            const u32 nbad = tile_bad ? sm.bad_n[j & 1] : 0u;
            if (tile_bad && nbad <= (u32)kVar2BadCap) {
                // every invalid byte of the tile is on the list: map each to its read (bytes outside the tile's own range belong
                // to the neighbouring tiles, which see them too)
                const int64_t first = ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads;
                for (u32 e = threadIdx.x; e < nbad; e += kVar2Consumers) {
                    const u32 ent = sm.bad_list[j & 1][e];
                    const int cbase = 16 * (int)(ent >> 16) - lead;
                    for (u32 bits = ent & 0xFFFFu; bits; bits &= bits - 1) {
                        const int pos = cbase + (__ffs(bits) - 1);
                        if (pos < 0 || (u32)pos >= m.srel[nreads] || m.srel[nreads] == 0xFFFFFFFFu) continue;
                        int lo_ = 0, hi_ = nreads;                // srel[lo_] <= pos < srel[hi_]
#pragma unroll
                        for (int it = 0; it < kVar2Log2Reads; it++) {
                            const int mid = (lo_ + hi_) >> 1;
                            if (m.srel[mid] != 0xFFFFFFFFu && m.srel[mid] <= (u32)pos) lo_ = mid; else hi_ = mid;
                        }
                        const u32 r0 = m.srel[lo_], r1 = m.srel[lo_ + 1];
                        if (r0 == 0xFFFFFFFFu || r1 == 0xFFFFFFFFu || (u32)pos < r0 || (u32)pos >= r1 || r1 - r0 > 1024 || r1 - r0 < 97) continue;
                        atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + first + lo_));
                    }
                }
            } else if (tile_bad) {                            // too many invalid chunks for the list: exact re-check, warp per read
                for (int r = warp; r < nreads; r += kVar2Consumers / 32) {
                    const u32 r0 = m.srel[r], r1 = m.srel[r + 1];
                    if (r0 == 0xFFFFFFFFu || r1 == 0xFFFFFFFFu || r1 < r0 || r1 - r0 > 1024 || r1 - r0 < 97) continue;
                    bool b = false;
                    for (u32 q = r0 + lane; q < r1; q += 32) b |= !is_acgt(a.ascii[t0 + q]);
                    if (__any_sync(0xFFFFFFFFu, b) && lane == 0)
                        atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads + r));
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&sm.ext_done[ms]));       // this warp no longer reads the tile's metadata
        if (kVar2Codes == 1) named_bar_sync(1, kVar2Consumers);        // single code buffer: the next tile's encode rewrites it
    }
}

// Persistent grid: exactly as many CTAs as are resident at once (one wave), capped by the number of tiles.
template <int MODE, int KLASS = SSQ_CLASS_64>
constexpr size_t pack_dyn_smem() { return MODE == kModeScatter ? (size_t)FixedCfg<KLASS>::PARTS * FixedCfg<KLASS>::RING * sizeof(u64) : 0; }

// The scatter mode needs 64 KB of dynamic shared memory per CTA (three CTAs per SM = most of the 227 KB).
template <int KLASS, int MODE>
static int prepare_fixed_kernel() {
    if (MODE == kModeScatter) {
        static bool done = false;      // per process; attributes are per device but identical on every B200
        static int done_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!done || done_dev != dev) {
            SSQ_CUDA(cudaFuncSetAttribute(pack_fixed_kernel<KLASS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_dyn_smem<MODE, KLASS>()));
            SSQ_CUDA(cudaFuncSetAttribute(pack_fixed_kernel<KLASS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            done = true;
            done_dev = dev;
        }
    }
    return SSQ_OK;
}

template <int KLASS, int MODE>
static int fixed_grid(ssq_ctx *ctx, int64_t n) {
    int per_sm = 0;
    if (prepare_fixed_kernel<KLASS, MODE>() != SSQ_OK) return 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pack_fixed_kernel<KLASS, MODE>, kPackThreads, pack_dyn_smem<MODE, KLASS>()) != cudaSuccess || per_sm < 1)
        per_sm = MODE == kModeScatter ? 3 : 4;
    return grid_for(ctx, (n + FixedCfg<KLASS>::TILE - 1) / FixedCfg<KLASS>::TILE, per_sm);
}

// A batch of exactly 32 n bytes whose first byte is 16-byte aligned is (almost certainly) n reads of 32 nt: such
// batches go to pack32_kernel, which verifies the offsets tile by tile and handles anything else it meets.
static bool looks_uniform32(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t n) {
    static const bool off = getenv("SSQ_NO_PACK32") != nullptr;        // development: force the general kernel
    return !off && n >= 4096 && hi - lo == 32 * n && (((uintptr_t)ascii + (uintptr_t)lo) & 15) == 0;
}

static int pack32_threads() {
    static int v = 0;
    if (v == 0) { const char *e = getenv("SSQ_PACK32_THREADS"); v = e && atoi(e) == 512 ? 512 : 256; }
    return v;
}

template <int MODE, int THREADS>
static int pack32_grid(ssq_ctx *ctx, int64_t n) {
    static bool done = false;
    static int done_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (MODE == kModeScatter && (!done || done_dev != dev)) {
        if (cudaFuncSetAttribute(pack32_kernel<MODE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_dyn_smem<MODE>()) != cudaSuccess) return 0;
        if (cudaFuncSetAttribute(pack32_kernel<MODE, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return 0;
        done = true;
        done_dev = dev;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pack32_kernel<MODE, THREADS>, THREADS, pack_dyn_smem<MODE>()) != cudaSuccess || per_sm < 1)
        per_sm = THREADS == 512 ? 2 : 3;
    return grid_for(ctx, (n * 32 + (int64_t)THREADS * 64 - 1) / ((int64_t)THREADS * 64), per_sm);
}

template <int MODE>
static int launch_pack32(ssq_ctx *ctx, const PackArgs &a, const TableView &t, const PartView &pv, const u64 *stop, int grid) {
    if (pack32_threads() == 512) {
        const int g = pack32_grid<MODE, 512>(ctx, a.n);
        if (g <= 0) { set_error("pack32_kernel: cannot configure shared memory"); return SSQ_ERR_CUDA; }
        pack32_kernel<MODE, 512><<<grid > 0 ? grid : g, 512, pack_dyn_smem<MODE>(), ctx->stream>>>(a, t, pv, stop);
    } else {
        const int g = pack32_grid<MODE, 256>(ctx, a.n);
        if (g <= 0) { set_error("pack32_kernel: cannot configure shared memory"); return SSQ_ERR_CUDA; }
        pack32_kernel<MODE, 256><<<grid > 0 ? grid : g, 256, pack_dyn_smem<MODE>(), ctx->stream>>>(a, t, pv, stop);
    }
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

template <int KLASS, int MODE>
static int launch_fixed(ssq_ctx *ctx, const PackArgs &a, const TableView &t, const PartView &pv, const u64 *stop, int grid = 0) {
    if (a.n <= 0) return SSQ_OK;
    if constexpr (KLASS == SSQ_CLASS_64) {
        if (looks_uniform32(a.ascii, a.lo, a.hi, a.n)) return launch_pack32<MODE>(ctx, a, t, pv, stop, grid);
    }
    if (grid <= 0) grid = fixed_grid<KLASS, MODE>(ctx, a.n);
    int rc = prepare_fixed_kernel<KLASS, MODE>();
    if (rc) return rc;
    pack_fixed_kernel<KLASS, MODE><<<grid, kPackThreads, pack_dyn_smem<MODE, KLASS>(), ctx->stream>>>(a, t, pv, stop);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

// grid of the scatter-mode launch for the n reads in ascii[lo, hi): it fixes the segment layout of the partition buffers
int scatter_grid(ssq_ctx *ctx, int klass, const uint8_t *ascii, int64_t lo, int64_t hi, int64_t n) {
    if (klass == SSQ_CLASS_64 && looks_uniform32(ascii, lo, hi, n)) {
        const int g = pack32_threads() == 512 ? pack32_grid<kModeScatter, 512>(ctx, n) : pack32_grid<kModeScatter, 256>(ctx, n);
        if (g > 0) return g;
    }
    return klass == SSQ_CLASS_64 ? fixed_grid<SSQ_CLASS_64, kModeScatter>(ctx, n) : fixed_grid<SSQ_CLASS_192, kModeScatter>(ctx, n);
}

// used by ssq_counter.cu: fused pack + (direct insert | partition scatter)
int launch_pack_count(ssq_ctx *ctx, int klass, bool scatter, const uint8_t *ascii, int64_t lo, int64_t hi,
                      const int64_t *offsets, int64_t n, int64_t index_base, u64 *words, uint8_t *lens,
                      const TableView &t, const PartView &pv, const u64 *stop) {
    PackArgs a{ascii, lo, hi, offsets, n, index_base, words, lens, nullptr, ctx->d_report};
    if (klass == SSQ_CLASS_64)
        return scatter ? launch_fixed<SSQ_CLASS_64, kModeScatter>(ctx, a, t, pv, stop, (int)pv.num_ctas)
                       : launch_fixed<SSQ_CLASS_64, kModeDirect>(ctx, a, t, pv, stop);
    return scatter ? launch_fixed<SSQ_CLASS_192, kModeScatter>(ctx, a, t, pv, stop, (int)pv.num_ctas)
                   : launch_fixed<SSQ_CLASS_192, kModeDirect>(ctx, a, t, pv, stop);
}

}  // namespace ssq

using namespace ssq;

static int check_pack_args(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets, int64_t n,
                           const void *words, const void *lens) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    SSQ_ARG(n >= 0 && ascii_bytes >= 0, "negative size");
    SSQ_ARG(n == 0 || (offsets != nullptr && words != nullptr && lens != nullptr), "NULL buffer");
    SSQ_ARG(ascii_bytes == 0 || ascii != nullptr, "ascii is NULL");
    return SSQ_OK;
}

extern "C" {

int ssq_pack64(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets, int64_t n,
               uint64_t *words, uint8_t *lens) {
    int rc = check_pack_args(ctx, ascii, ascii_bytes, offsets, n, words, lens);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    PackArgs a{ascii, 0, ascii_bytes, offsets, n, 0, (u64 *)words, lens, nullptr, ctx->d_report};
    return launch_fixed<SSQ_CLASS_64, kModePack>(ctx, a, TableView{}, PartView{}, nullptr);
}

int ssq_pack192(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets, int64_t n,
                uint64_t *words, uint8_t *lens) {
    int rc = check_pack_args(ctx, ascii, ascii_bytes, offsets, n, words, lens);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    PackArgs a{ascii, 0, ascii_bytes, offsets, n, 0, (u64 *)words, lens, nullptr, ctx->d_report};
    return launch_fixed<SSQ_CLASS_192, kModePack>(ctx, a, TableView{}, PartView{}, nullptr);
}

int64_t ssq_packvar_words_bound(int64_t ascii_bytes, int64_t n) {
    // sum ceil(len_i/32) <= (sum len_i + 31 n) / 32
    return (ascii_bytes + 31 * n) / 32 + 1;
}

int ssq_packvar(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets, int64_t n,
                int64_t *word_off, uint64_t *words, uint16_t *lens) {
    int rc = check_pack_args(ctx, ascii, ascii_bytes, offsets, n, words, lens);
    if (rc) return rc;
    SSQ_ARG(word_off != nullptr, "word_off is NULL");
    DeviceGuard g(ctx->device);
    rc = scan_var_words(ctx, offsets, n, word_off);
    if (rc || n == 0) return rc;
    PackArgs a{ascii, 0, ascii_bytes, offsets, n, 0, (u64 *)words, lens, word_off, ctx->d_report};
    static const bool v1 = getenv("SSQ_VAR_V1") != nullptr;      // development: the first version of the kernel
    if (v1) {
        int64_t ntiles = (n + kVarTileReads - 1) / kVarTileReads;
        int grid = grid_for(ctx, ntiles, 8);
        pack_var_kernel<<<grid, kPackThreads, 0, ctx->stream>>>(a);
        SSQ_LAUNCH_CHECK();
        return SSQ_OK;
    }
    SSQ_CUDA(cudaFuncSetAttribute(pack_var2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Var2Smem)));
    SSQ_CUDA(cudaFuncSetAttribute(pack_var2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pack_var2_kernel, kVar2Threads, sizeof(Var2Smem)) != cudaSuccess || per_sm < 1) per_sm = 2;
    pack_var2_kernel<<<grid_for(ctx, (n + kVar2Reads - 1) / kVar2Reads, per_sm), kVar2Threads, sizeof(Var2Smem), ctx->stream>>>(a);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

}  // extern "C"
