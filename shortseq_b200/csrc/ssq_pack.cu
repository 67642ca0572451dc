// ssq_pack.cu -- batched 2-bit packing with fused validation (and optional fused counting).
//
// Replaces, for a whole batch, the reference's per-object encoders
// (short_seq.pyx:54-74 -> short_seq_64.pyx:96-108, short_seq_192.pyx:103-108,
// short_seq_var.pyx:123-132 -> util.pyx:78-140) and validators (util.pxd:98-127).
//
// Two stages per tile of reads, both inside one kernel:
//   1. STREAM ENCODE.  The tile's reads are one contiguous byte range of the ASCII
//      buffer.  Threads sweep it with coalesced, 16-byte aligned, L1-bypassing vector
//      loads; each 16-byte chunk becomes 32 bits of 2-bit codes in shared memory and an
//      invalid-byte indicator is OR-accumulated (exact {A,C,G,T} test, 4 bytes per op).
//   2. EXTRACT.  A read's packed words are just the bits [2*start, 2*(start+len)) of that
//      code stream: a few shared-memory loads and funnel shifts per 64-bit word, masked
//      to the read length.  ShortSeq64/192: one thread per read; ShortSeqVar: one warp
//      per read, lane j owns word j.  Words and lengths are written coalesced.
// The per-byte validation result is consulted per tile: only when some byte of the tile
// is invalid do reads re-check their own bytes (rare slow path) to find the lowest
// failing read index.  With COUNT the extracted key goes straight into the dedup table.
#include "ssq_internal.h"
#include "ssq_table.cuh"

namespace ssq {

constexpr int kPackThreads = 256;

struct PackArgs {
    const uint8_t *ascii;     // base the offsets index into (may be a virtual base for a staged slice)
    int64_t lo, hi;           // valid byte index range [lo, hi) of `ascii`
    const int64_t *offsets;
    int64_t n;            // reads in this launch
    int64_t index_base;   // index of read 0 of this launch within the caller's batch
    u64 *words;
    void *lens;
    const int64_t *word_off;  // ShortSeqVar only
    DevReport *rep;
};

// Load the 16-byte chunk that starts at byte index `idx` (may stick out of the buffer).
__device__ __forceinline__ uint4 load_chunk(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t idx) {
    if (idx >= lo && idx + 16 <= hi) return ld_stream_v4(ascii + idx);
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

// Replace the bytes of a chunk that lie outside [lo, hi) (byte indices) by 'A' so that
// bytes of neighbouring tiles are neither validated nor encoded here.
__device__ __forceinline__ uint4 clip_chunk(uint4 v, int64_t idx, int64_t lo, int64_t hi) {
    u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 keep = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int64_t j = idx + 4 * k + b;
            if (j >= lo && j < hi) keep |= 0xFFu << (8 * b);
        }
        w[k] = (w[k] & keep) | (0x41414141u & ~keep);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Stage 1.  Encodes the bytes [t0, t1) into codes[]; codes[c] covers the 16 bytes starting at
// byte index a0 + 16c where a0 is t0 rounded down to a 16-byte ADDRESS boundary.  Returns a0.
// `pad` extra words after the last chunk are zeroed.  The caller must __syncthreads().
template <int THREADS>
__device__ __forceinline__ int64_t encode_tile(const uint8_t *ascii, int64_t lo, int64_t hi, int64_t t0, int64_t t1,
                                               u32 *codes, int pad, u32 &bad) {
    const int64_t mis = (int64_t)((uintptr_t)ascii & 15);
    const int64_t a0 = ((t0 + mis) & ~(int64_t)15) - mis;
    const int nchunks = (int)((t1 - a0 + 15) >> 4);
    for (int c = threadIdx.x; c < nchunks; c += THREADS) {
        int64_t idx = a0 + 16 * (int64_t)c;
        uint4 v = load_chunk(ascii, lo, hi, idx);
        if (c == 0 || c == nchunks - 1) v = clip_chunk(v, idx, t0, t1);
        codes[c] = encode16(v, bad);
    }
    if ((int)threadIdx.x < pad) codes[nchunks + threadIdx.x] = 0;
    return a0;
}

// 64 bits of the code stream starting at bit `bit` (even) of codes[].
__device__ __forceinline__ u64 extract64(const u32 *codes, int64_t bit) {
    int wi = (int)(bit >> 5);
    u32 sh = (u32)bit & 31;
    u32 c0 = codes[wi], c1 = codes[wi + 1], c2 = codes[wi + 2];
    u32 lo = __funnelshift_r(c0, c1, sh);
    u32 hi = __funnelshift_r(c1, c2, sh);
    return ((u64)hi << 32) | lo;
}

// keep the low `nbits` (0..64) bits
__device__ __forceinline__ u64 keep_bits(u64 x, int nbits) {
    if (nbits >= 64) return x;
    if (nbits <= 0) return 0;
    return x & ((1ull << nbits) - 1);
}

// Slow path: exact re-check of one read's bytes.
__device__ __noinline__ bool read_has_bad_base(const uint8_t *ascii, int64_t o0, int64_t len) {
    for (int64_t j = 0; j < len; j++)
        if (!is_acgt(ascii[o0 + j])) return true;
    return false;
}

__device__ __forceinline__ void report_len(DevReport *rep, int64_t len, u64 idx) {
    if (len > 1024) atomicMin(&rep->first_too_long, idx);
    else atomicMin(&rep->first_bad_len, idx);
}

// ---- ShortSeq64 / ShortSeq192: one thread per read --------------------------------------
template <int KLASS, bool COUNT>
__global__ void __launch_bounds__(kPackThreads) pack_fixed_kernel(PackArgs a, TableView t, const u64 *stop) {
    constexpr int MAXLEN = KLASS == SSQ_CLASS_64 ? 32 : 96;
    constexpr int MINLEN = KLASS == SSQ_CLASS_64 ? 0 : 33;
    constexpr int W = KLASS == SSQ_CLASS_64 ? 1 : 3;
    constexpr int PAD = 2 * W + 1;
    constexpr int MAX_CHUNKS = (kPackThreads * MAXLEN + 30) / 16 + 1;
    __shared__ u32 codes[MAX_CHUNKS + PAD];
    __shared__ u32 s_new[kPackThreads / 32];

    if (COUNT && stop != nullptr && *stop != 0) return;

    u32 my_new = 0;
    const int64_t ntiles = (a.n + kPackThreads - 1) / kPackThreads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kPackThreads;
        const int nreads = (int)min((int64_t)kPackThreads, a.n - first);
        const bool mine = (int)threadIdx.x < nreads;
        const int64_t i = first + threadIdx.x;
        int64_t o0 = 0, o1 = 0;
        if (mine) { o0 = a.offsets[i]; o1 = a.offsets[i + 1]; }
        const int64_t t0 = a.offsets[first];
        const int64_t t1 = a.offsets[first + nreads];
        const int64_t len = o1 - o0;
        const bool len_ok = mine && len >= MINLEN && len <= MAXLEN;
        if (mine && !len_ok) report_len(a.rep, len, (u64)(a.index_base + i));
        // A tile whose byte range is inconsistent or does not fit the staging buffer contains a
        // read that was reported above (or offsets outside the buffer): skip it.
        const bool tile_ok = t0 >= a.lo && t1 >= t0 && t1 <= a.hi && (t1 - t0) <= (int64_t)kPackThreads * MAXLEN;
        if (!tile_ok) {
            if (threadIdx.x == 0 && (t0 < a.lo || t1 > a.hi)) atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
            continue;
        }
        u32 bad = 0;
        const int64_t a0 = encode_tile<kPackThreads>(a.ascii, a.lo, a.hi, t0, t1, codes, PAD, bad);
        const int tile_bad = __syncthreads_or(bad != 0);

        const bool in_tile = len_ok && o0 >= t0 && o1 <= t1;   // non-monotonic offsets were reported via len
        bool ok = in_tile;
        u64 w[W];
        if (in_tile) {
            const int64_t bit = 2 * (o0 - a0);
#pragma unroll
            for (int k = 0; k < W; k++) w[k] = keep_bits(extract64(codes, bit + 64 * k), 2 * (int)len - 64 * k);
            if (tile_bad && read_has_bad_base(a.ascii, o0, len)) {
                atomicMin(&a.rep->first_bad_base, (u64)(a.index_base + i));
                ok = false;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) w[k] = 0;
        }
        if (mine) {
#pragma unroll
            for (int k = 0; k < W; k++) a.words[(size_t)i * W + k] = w[k];
            ((uint8_t *)a.lens)[i] = len_ok ? (uint8_t)len : 0;
        }
        if (COUNT) {
            bool is_new = false;
            if (ok) {
                if constexpr (KLASS == SSQ_CLASS_64) insert64(t, w[0], (u32)len, 1ull, is_new);
                else insert192(t, w[0], w[1], w[2], (u32)len, 1ull, is_new);
            }
            my_new += is_new ? 1u : 0u;
        }
        __syncthreads();   // codes[] is rewritten by the next tile
    }
    if (COUNT) {
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
        // every lane r of warp 0 checks read r's length
        if (warp == 0 && lane < nreads) {
            int64_t len = a.offsets[first + lane + 1] - a.offsets[first + lane];
            if (len < 97 || len > 1024) report_len(a.rep, len, (u64)(a.index_base + first + lane));
        }
        const bool tile_ok = t0 >= a.lo && t1 >= t0 && t1 <= a.hi && (t1 - t0) <= (int64_t)kVarTileReads * 1024;
        if (!tile_ok) {
            if (threadIdx.x == 0 && (t0 < a.lo || t1 > a.hi)) atomicMin(&a.rep->first_bad_len, (u64)(a.index_base + first));
            continue;
        }
        u32 bad = 0;
        const int64_t a0 = encode_tile<kPackThreads>(a.ascii, a.lo, a.hi, t0, t1, codes, PAD, bad);
        const int tile_bad = __syncthreads_or(bad != 0);
        for (int r = warp; r < nreads; r += kPackThreads / 32) {
            const int64_t i = first + r;
            const int64_t o0 = a.offsets[i], o1 = a.offsets[i + 1];
            const int64_t len = o1 - o0;
            const bool len_ok = len >= 97 && len <= 1024 && o0 >= t0 && o1 <= t1;
            if (len_ok) {
                const int nwords = (int)((len + 31) >> 5);
                if (lane < nwords) {
                    u64 w = keep_bits(extract64(codes, 2 * (o0 - a0) + 64 * lane), 2 * (int)len - 64 * lane);
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

template <int KLASS, bool COUNT>
static int launch_fixed(ssq_ctx *ctx, const PackArgs &a, const TableView &t, const u64 *stop) {
    if (a.n <= 0) return SSQ_OK;
    int64_t ntiles = (a.n + kPackThreads - 1) / kPackThreads;
    int grid = grid_for(ctx, ntiles, 8);
    pack_fixed_kernel<KLASS, COUNT><<<grid, kPackThreads, 0, ctx->stream>>>(a, t, stop);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

// used by ssq_counter.cu for the fused pack+count launches
int launch_pack_count(ssq_ctx *ctx, int klass, const uint8_t *ascii, int64_t lo, int64_t hi, const int64_t *offsets,
                      int64_t n, int64_t index_base, u64 *words, uint8_t *lens, const TableView &t, const u64 *stop) {
    PackArgs a{ascii, lo, hi, offsets, n, index_base, words, lens, nullptr, ctx->d_report};
    if (klass == SSQ_CLASS_64) return launch_fixed<SSQ_CLASS_64, true>(ctx, a, t, stop);
    return launch_fixed<SSQ_CLASS_192, true>(ctx, a, t, stop);
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
    return launch_fixed<SSQ_CLASS_64, false>(ctx, a, TableView{}, nullptr);
}

int ssq_pack192(ssq_ctx *ctx, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets, int64_t n,
                uint64_t *words, uint8_t *lens) {
    int rc = check_pack_args(ctx, ascii, ascii_bytes, offsets, n, words, lens);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    PackArgs a{ascii, 0, ascii_bytes, offsets, n, 0, (u64 *)words, lens, nullptr, ctx->d_report};
    return launch_fixed<SSQ_CLASS_192, false>(ctx, a, TableView{}, nullptr);
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
