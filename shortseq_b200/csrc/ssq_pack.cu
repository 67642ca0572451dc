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
//   kModeScatter  append each key's 64-bit table key to one of 256 hash partitions
//                 (block-level histogram in shared memory -> one global cursor bump per
//                 partition per tile), so that ssq_counter.cu can insert partition by
//                 partition with the table region resident in L2.
#include "ssq_internal.h"
#include "ssq_table.cuh"

namespace ssq {

constexpr int kPackThreads = 256;
constexpr int kRPT = 2;                               // reads per thread per tile (fixed classes)
constexpr int kTileReads = kPackThreads * kRPT;
constexpr int kLoadUnroll = 4;                        // 16-byte loads in flight per thread

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

// Stage 1.  Encodes the 16-byte chunks covering bytes [t0, t1) into codes[]; chunk c covers the 16
// bytes from byte index a0 + 16c, a0 = t0 rounded down to a 16-byte ADDRESS boundary.  `lead`
// receives t0 - a0.  Bytes of neighbouring tiles that share the edge chunks are encoded and
// validated too: their codes are never extracted, and a stray invalid byte there merely sends this
// tile through the exact per-read re-check.  `pad` extra words after the last chunk are zeroed.
// The caller must __syncthreads().
template <int THREADS>
__device__ __forceinline__ void encode_tile(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t t0, int tile_bytes,
                                            u32 *codes, int pad, u32 &bad, int &lead) {
    const int mis = (int)((uintptr_t)ascii & 15);
    lead = (int)((t0 + mis) & 15);
    const int64_t a0 = t0 - lead;
    const int nchunks = (tile_bytes + lead + 15) >> 4;
    const uint8_t *src = ascii + a0;
    if (a0 >= lo && a0 + 16 * (int64_t)nchunks <= hi) {          // whole sweep inside the buffer: no guards
        for (int c0 = threadIdx.x; c0 < nchunks; c0 += THREADS * kLoadUnroll) {
            uint4 v[kLoadUnroll];
#pragma unroll
            for (int j = 0; j < kLoadUnroll; j++) {
                int c = c0 + j * THREADS;
                if (c < nchunks) v[j] = ld_stream_v4(src + 16 * c);
            }
#pragma unroll
            for (int j = 0; j < kLoadUnroll; j++) {
                int c = c0 + j * THREADS;
                if (c < nchunks) codes[c] = encode16(v[j], bad);
            }
        }
    } else {                                                     // first / last tile of the buffer
        for (int c = threadIdx.x; c < nchunks; c += THREADS) {
            int64_t idx = a0 + 16 * (int64_t)c;
            uint4 v = (idx >= lo && idx + 16 <= hi) ? ld_stream_v4(ascii + idx) : load_chunk_guarded(ascii, lo, hi, idx);
            codes[c] = encode16(v, bad);
        }
    }
    if ((int)threadIdx.x < pad) codes[nchunks + threadIdx.x] = 0;
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

// ---- ShortSeq64 / ShortSeq192 ------------------------------------------------------------------
template <int KLASS, int MODE>
__global__ void __launch_bounds__(kPackThreads) pack_fixed_kernel(PackArgs a, TableView t, PartView pv, const u64 *stop) {
    constexpr int MAXLEN = KLASS == SSQ_CLASS_64 ? 32 : 96;
    constexpr int MINLEN = KLASS == SSQ_CLASS_64 ? 0 : 33;
    constexpr int W = KLASS == SSQ_CLASS_64 ? 1 : 3;
    constexpr int PAD = 2 * W + 1;
    constexpr int MAX_CHUNKS = (kTileReads * MAXLEN + 30) / 16 + 1;
    static_assert(MODE != kModeScatter || KLASS == SSQ_CLASS_64, "partition scatter is implemented for ShortSeq64 keys");
    __shared__ u32 codes[MAX_CHUNKS + PAD];
    __shared__ u32 srel[kTileReads + 1];                // read starts relative to the tile start
    __shared__ u32 s_new[kPackThreads / 32];
    __shared__ u32 hist[MODE == kModeScatter ? 2 * kParts : 1];
    __shared__ u32 pbase[MODE == kModeScatter ? kParts : 1];

    if (MODE != kModePack && stop != nullptr && *stop != 0) return;
    if (MODE == kModeScatter) {
        for (int p = threadIdx.x; p < 2 * kParts; p += kPackThreads) hist[p] = 0;
        __syncthreads();
    }

    u32 my_new = 0;
    int flip = 0;
    const int64_t ntiles = (a.n + kTileReads - 1) / kTileReads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kTileReads;
        const int nreads = (int)min((int64_t)kTileReads, a.n - first);
        const int64_t t0 = a.offsets[first];
        const int64_t t1 = a.offsets[first + nreads];
        // a tile whose byte range is inconsistent or larger than the staging buffer holds a read of the
        // wrong class (or offsets outside the buffer): report per read, pack nothing
        const bool tile_ok = t0 >= a.lo && t1 >= t0 && t1 <= a.hi && (t1 - t0) <= (int64_t)kTileReads * MAXLEN;
#pragma unroll
        for (int k = 0; k < kRPT; k++) {
            const int r = threadIdx.x + k * kPackThreads;
            if (r < nreads) {
                const int64_t o = a.offsets[first + r];
                if (tile_ok) {
                    // out-of-tile starts become an impossible value that fails the checks below
                    srel[r] = (o >= t0 && o <= t1) ? (u32)(o - t0) : 0xFFFFFFFFu;
                } else {
                    const int64_t len = a.offsets[first + r + 1] - o;
                    if (len < MINLEN || len > MAXLEN) report_len(a.rep, len, (u64)(a.index_base + first + r));
                }
            }
        }
        if (!tile_ok) {
            if (threadIdx.x == 0 && (t0 < a.lo || t1 > a.hi || t1 < t0))
                atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
            continue;
        }
        const int tile_bytes = (int)(t1 - t0);
        if (threadIdx.x == 0) srel[nreads] = (u32)tile_bytes;
        u32 bad = 0;
        int lead;
        encode_tile<kPackThreads>(a.ascii, a.lo, a.hi, t0, tile_bytes, codes, PAD, bad, lead);
        const int tile_bad = __syncthreads_or(bad != 0);

        u64 key[kRPT];
        u32 rank[kRPT], part[kRPT];
        bool ok[kRPT];
        u32 *h = hist + flip * kParts;
#pragma unroll
        for (int k = 0; k < kRPT; k++) {
            const int r = threadIdx.x + k * kPackThreads;
            ok[k] = false;
            if (r >= nreads) continue;
            const int64_t i = first + r;
            const u32 r0 = srel[r], r1 = srel[r + 1];
            const int len = (int)(r1 - r0);
            const bool len_ok = r0 <= (u32)tile_bytes && r1 <= (u32)tile_bytes && len >= MINLEN && len <= MAXLEN;
            u64 w[W];
            if (len_ok) {
                const int bit = 2 * ((int)r0 + lead);
#pragma unroll
                for (int j = 0; j < W; j++) w[j] = keep_bits(extract64(codes, bit + 64 * j), 2 * len - 64 * j);
                ok[k] = !(tile_bad && read_has_bad_base(a.ascii + t0 + r0, len));
                if (!ok[k]) atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
            } else {
#pragma unroll
                for (int j = 0; j < W; j++) w[j] = 0;
                report_len(a.rep, (int64_t)a.offsets[i + 1] - a.offsets[i], (u64)(a.index_base + i));
            }
#pragma unroll
            for (int j = 0; j < W; j++) a.words[(size_t)i * W + j] = w[j];
            ((uint8_t *)a.lens)[i] = len_ok ? (uint8_t)len : 0;
            if (MODE == kModeDirect && ok[k]) {
                bool is_new = false;
                if constexpr (KLASS == SSQ_CLASS_64) insert64(t, w[0], (u32)len, 1ull, is_new);
                else insert192(t, w[0], w[1], w[2], (u32)len, 1ull, is_new);
                my_new += is_new ? 1u : 0u;
            }
            if (MODE == kModeScatter && ok[k]) {
                const u64 h2 = rotl64(mix64(w[0]), t.rot);
                key[k] = key64_of(h2, (u32)len);
                part[k] = (u32)(h2 >> 56);
                rank[k] = atomicAdd(&h[part[k]], 1u);
            }
        }
        if (MODE == kModeScatter) {
            __syncthreads();
            // one cursor bump per non-empty partition per tile; the other histogram is cleared for the next tile
            for (int p = threadIdx.x; p < kParts; p += kPackThreads) {
                const u32 cnt = h[p];
                pbase[p] = cnt ? atomicAdd(&pv.cursor[p], cnt) : 0u;
                hist[(flip ^ 1) * kParts + p] = 0;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRPT; k++) {
                if (!ok[k]) continue;
                const u32 pos = pbase[part[k]] + rank[k];
                if (pos < pv.cap_per_part) {
                    pv.keys[(size_t)part[k] * pv.cap_per_part + pos] = key[k];
                } else {                                   // partition buffer full: count it right away
                    bool is_new = false;
                    const u64 h2 = ((u64)part[k] << 56) | (key[k] & kMask56);
                    insert64_hashed(t, h2, key[k], 1ull, is_new);
                    my_new += is_new ? 1u : 0u;
                }
            }
            flip ^= 1;
        } else {
            __syncthreads();   // codes[] / srel[] are rewritten by the next tile
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
        int lead;
        encode_tile<kPackThreads>(a.ascii, a.lo, a.hi, t0, (int)(t1 - t0), codes, PAD, bad, lead);
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

template <int KLASS, int MODE>
static int launch_fixed(ssq_ctx *ctx, const PackArgs &a, const TableView &t, const PartView &pv, const u64 *stop) {
    if (a.n <= 0) return SSQ_OK;
    int64_t ntiles = (a.n + kTileReads - 1) / kTileReads;
    int grid = grid_for(ctx, ntiles, 6);
    pack_fixed_kernel<KLASS, MODE><<<grid, kPackThreads, 0, ctx->stream>>>(a, t, pv, stop);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

// used by ssq_counter.cu: fused pack + (direct insert | partition scatter)
int launch_pack_count(ssq_ctx *ctx, int klass, bool scatter, const uint8_t *ascii, int64_t lo, int64_t hi,
                      const int64_t *offsets, int64_t n, int64_t index_base, u64 *words, uint8_t *lens,
                      const TableView &t, const PartView &pv, const u64 *stop) {
    PackArgs a{ascii, lo, hi, offsets, n, index_base, words, lens, nullptr, ctx->d_report};
    if (klass == SSQ_CLASS_64)
        return scatter ? launch_fixed<SSQ_CLASS_64, kModeScatter>(ctx, a, t, pv, stop)
                       : launch_fixed<SSQ_CLASS_64, kModeDirect>(ctx, a, t, pv, stop);
    return launch_fixed<SSQ_CLASS_192, kModeDirect>(ctx, a, t, pv, stop);
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
    int64_t ntiles = (n + kVarTileReads - 1) / kVarTileReads;
    int grid = grid_for(ctx, ntiles, 4);
    pack_var_kernel<<<grid, kPackThreads, 0, ctx->stream>>>(a);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

}  // extern "C"
