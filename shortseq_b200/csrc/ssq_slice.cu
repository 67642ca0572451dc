// ssq_slice.cu -- batched slicing and k-mer extraction over packed arrays, and the tolerant (U / lower case) alphabet.
//
// "Next" row N4 of SURVEY section 8f.  The reference slices one object at a time on the host
// (short_seq.pyx:94-116 _slice, :119-199 _slice_to_ShortSeq64/192/Var, :202-238 _shift_copy_trim): the blocks of the
// source are funnel-shifted so that base `start` lands in bits 0..1 of block 0, and the tail is trimmed (bzhi) so that
// whole-block compares and popcounts of the result stay valid.  Here the same is done for a whole array per launch:
// a sub-sequence is just the bits [2 start, 2 (start + len)) of the read's packed words.
//   ssq_slice      out[i] = in[i][start_i : start_i + width]  (Python slice clamping; the output class is chosen by the
//                  caller from `width`, like the reference's result class follows the slice length)
//   ssq_kmers64    every k-mer (k <= 32, step `stride`) of every read as ShortSeq64 words, grouped by read (CSR offsets
//                  from ssq_kmers_count); feeding them to ssq_counter_insert is k-mer counting
//   ssq_normalize  opt-in alphabet: a c g t u U are rewritten to A C G T T T before packing (the reference's table_91
//                  maps U to T's code, util.pyx:44-50, but its validator rejects it; lower case is always rejected)
#include "ssq_internal.h"

namespace ssq {

int scan_kmer_counts(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, int k, int stride, int64_t *out);
int scan_slice_words(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, const int64_t *starts, const int64_t *stops, int64_t start0,
                     int64_t stop0, int32_t width, int64_t *out);

constexpr int kSliceThreads = 256;

// A packed array of any class as the kernels see it.
struct PackedView {
    const u64 *words;
    const int64_t *word_off;   // ShortSeqVar only
    const void *lens;          // uint8 (ShortSeq64 / ShortSeq192) or uint16 (ShortSeqVar)
    int klass;
};

__device__ __forceinline__ u32 view_len(const PackedView &v, int64_t i) {
    return v.klass == SSQ_CLASS_VAR ? ((const uint16_t *)v.lens)[i] : ((const uint8_t *)v.lens)[i];
}
// first word and number of words of read i
__device__ __forceinline__ const u64 *view_words(const PackedView &v, int64_t i, u32 len, u32 &nwords) {
    if (v.klass == SSQ_CLASS_64) { nwords = 1; return v.words + i; }
    if (v.klass == SSQ_CLASS_192) { nwords = 3; return v.words + 3 * i; }
    nwords = (len + 31) >> 5;
    return v.words + v.word_off[i];
}
// 64 bits of a read's code stream starting at bit `bit` (even); words past the end read as 0
__device__ __forceinline__ u64 bits_at(const u64 *w, u32 nwords, u32 bit) {
    const u32 wi = bit >> 6, sh = bit & 63;
    const u64 lo = wi < nwords ? w[wi] : 0ull;
    if (sh == 0) return lo;
    const u64 hi = wi + 1 < nwords ? w[wi + 1] : 0ull;
    return (lo >> sh) | (hi << (64 - sh));
}
__device__ __forceinline__ u64 keep_low_bits(u64 x, int nbits) {
    if (nbits >= 64) return x;
    if (nbits <= 0) return 0;
    return x & ((1ull << nbits) - 1);
}
// [start, stop) clamped to a read of `len` bases (both already non-negative: the caller resolves Python's negative
// indices) and to at most `width` bases -> first base, number of bases
__device__ __forceinline__ void clamp_slice(int64_t start, int64_t stop, int32_t width, u32 len, u32 &s, u32 &n) {
    int64_t a = start < 0 ? 0 : (start > (int64_t)len ? (int64_t)len : start);
    int64_t b = stop > (int64_t)len ? (int64_t)len : stop;
    if (b < a) b = a;
    if (b - a > width) b = a + width;
    s = (u32)a;
    n = (u32)(b - a);
}

// fixed-class output: one thread per read, W_OUT words each
template <int W_OUT>
__global__ void __launch_bounds__(kSliceThreads) slice_fixed_kernel(PackedView in, int64_t n, const int64_t *starts, const int64_t *stops, int64_t start0, int64_t stop0, int32_t width,
                                                                    u64 *out_words, uint8_t *out_lens) {
    for (int64_t i = (int64_t)blockIdx.x * kSliceThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSliceThreads) {
        const u32 len = view_len(in, i);
        u32 nwords, s, m;
        const u64 *w = view_words(in, i, len, nwords);
        clamp_slice(starts ? starts[i] : start0, stops ? stops[i] : stop0, width, len, s, m);
#pragma unroll
        for (int j = 0; j < W_OUT; j++)
            out_words[i * W_OUT + j] = keep_low_bits(bits_at(w, nwords, 2 * s + 64 * j), 2 * (int)m - 64 * j);
        out_lens[i] = (uint8_t)m;
    }
}

// ShortSeqVar output: one warp per read, lane j owns word j
__global__ void __launch_bounds__(kSliceThreads) slice_var_kernel(PackedView in, int64_t n, const int64_t *starts, const int64_t *stops, int64_t start0, int64_t stop0, int32_t width,
                                                                  const int64_t *out_word_off, u64 *out_words, uint16_t *out_lens) {
    const u32 lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kSliceThreads / 32);
    for (int64_t i = (int64_t)blockIdx.x * (kSliceThreads / 32) + (threadIdx.x >> 5); i < n; i += warps) {
        const u32 len = view_len(in, i);
        u32 nwords, s, m;
        const u64 *w = view_words(in, i, len, nwords);
        clamp_slice(starts ? starts[i] : start0, stops ? stops[i] : stop0, width, len, s, m);
        const u32 ow = (m + 31) >> 5;
        if (lane < ow) out_words[out_word_off[i] + lane] = keep_low_bits(bits_at(w, nwords, 2 * s + 64 * lane), 2 * (int)m - 64 * (int)lane);
        if (lane == 0) out_lens[i] = (uint16_t)m;
    }
}

// k-mers: one warp per read, lanes write consecutive k-mers
__global__ void __launch_bounds__(kSliceThreads) kmers64_kernel(PackedView in, int64_t n, int k, int stride, const int64_t *kmer_off,
                                                                u64 *out_words, uint8_t *out_lens) {
    const u32 lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kSliceThreads / 32);
    for (int64_t i = (int64_t)blockIdx.x * (kSliceThreads / 32) + (threadIdx.x >> 5); i < n; i += warps) {
        const u32 len = view_len(in, i);
        u32 nwords;
        const u64 *w = view_words(in, i, len, nwords);
        const int64_t o = kmer_off[i];
        const u32 cnt = (u32)(kmer_off[i + 1] - o);
        for (u32 m = lane; m < cnt; m += 32) {
            out_words[o + m] = keep_low_bits(bits_at(w, nwords, 2 * m * (u32)stride), 2 * k);
            out_lens[o + m] = (uint8_t)k;
        }
    }
}

// a c g t u U -> A C G T T T, everything else unchanged (16 bytes per thread; the tail byte by byte)
__device__ __forceinline__ u32 normalize4(u32 w) {
    // letters only: bytes whose upper-cased value is in 'A'..'Z' have bit 6 set and bit 7 clear; clearing bit 5 of any
    // other byte could turn an invalid byte into a valid one ('!' & 0xDF = 0x01 is harmless, but keep the rule simple):
    // only the ten accepted characters are rewritten, byte by byte through their exact values.
    u32 out = w;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const u32 c = (w >> (8 * b)) & 0xFFu;
        u32 r = c;
        if (c == 'a' || c == 'c' || c == 'g' || c == 't') r = c - 32;
        else if (c == 'u' || c == 'U') r = 'T';
        out = (out & ~(0xFFu << (8 * b))) | (r << (8 * b));
    }
    return out;
}
__global__ void __launch_bounds__(kSliceThreads) normalize_kernel(const uint8_t *in, uint8_t *out, int64_t nbytes) {
    const int64_t nvec = ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 ? nbytes >> 4 : 0;
    for (int64_t v = (int64_t)blockIdx.x * kSliceThreads + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kSliceThreads) {
        uint4 x = reinterpret_cast<const uint4 *>(in)[v];
        x.x = normalize4(x.x); x.y = normalize4(x.y); x.z = normalize4(x.z); x.w = normalize4(x.w);
        reinterpret_cast<uint4 *>(out)[v] = x;
    }
    for (int64_t i = (nvec << 4) + (int64_t)blockIdx.x * kSliceThreads + threadIdx.x; i < nbytes; i += (int64_t)gridDim.x * kSliceThreads) {
        const u32 c = in[i];
        out[i] = (uint8_t)(normalize4(c) & 0xFFu);
    }
}

}  // namespace ssq

using namespace ssq;

static int check_view(ssq_ctx *ctx, int klass, const void *words, const void *word_off, const void *lens, int64_t n) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    SSQ_ARG(klass == SSQ_CLASS_64 || klass == SSQ_CLASS_192 || klass == SSQ_CLASS_VAR, "bad class");
    SSQ_ARG(n >= 0, "negative size");
    SSQ_ARG(n == 0 || (words != nullptr && lens != nullptr), "NULL buffer");
    SSQ_ARG(n == 0 || klass != SSQ_CLASS_VAR || word_off != nullptr, "word_off is NULL");
    return SSQ_OK;
}

extern "C" {

int ssq_slice_words(ssq_ctx *ctx, int in_klass, const void *lens, int64_t n, const int64_t *starts, const int64_t *stops,
                    int64_t start0, int64_t stop0, int32_t width, int64_t *out_word_off) {
    SSQ_ARG(ctx != nullptr && out_word_off != nullptr && (n == 0 || lens != nullptr) && n >= 0 && width >= 0, "bad arguments");
    DeviceGuard g(ctx->device);
    return scan_slice_words(ctx, lens, in_klass == SSQ_CLASS_VAR ? 2 : 1, n, starts, stops, start0, stop0, width, out_word_off);
}

int ssq_slice(ssq_ctx *ctx, int in_klass, const uint64_t *words, const int64_t *word_off, const void *lens, int64_t n,
              const int64_t *starts, const int64_t *stops, int64_t start0, int64_t stop0, int32_t width, int out_klass,
              uint64_t *out_words, const int64_t *out_word_off, void *out_lens) {
    int rc = check_view(ctx, in_klass, words, word_off, lens, n);
    if (rc) return rc;
    SSQ_ARG(width >= 0 && width <= 1024, "width outside 0..1024");
    SSQ_ARG(out_klass == SSQ_CLASS_64 || out_klass == SSQ_CLASS_192 || out_klass == SSQ_CLASS_VAR, "bad output class");
    SSQ_ARG(out_klass != SSQ_CLASS_64 || width <= 32, "width > 32 needs ShortSeq192 or ShortSeqVar output");
    SSQ_ARG(out_klass != SSQ_CLASS_192 || width <= 96, "width > 96 needs ShortSeqVar output");
    SSQ_ARG(n == 0 || (out_words != nullptr && out_lens != nullptr), "NULL output");
    SSQ_ARG(n == 0 || out_klass != SSQ_CLASS_VAR || out_word_off != nullptr, "out_word_off is NULL (ssq_slice_words fills it)");
    if (n == 0) return SSQ_OK;
    DeviceGuard g(ctx->device);
    const PackedView in{(const u64 *)words, word_off, lens, in_klass};
    if (out_klass == SSQ_CLASS_VAR) {
        const int grid = grid_for(ctx, (n + kSliceThreads / 32 - 1) / (kSliceThreads / 32), 8);
        slice_var_kernel<<<grid, kSliceThreads, 0, ctx->stream>>>(in, n, starts, stops, start0, stop0, width, out_word_off, (u64 *)out_words,
                                                                  (uint16_t *)out_lens);
    } else {
        const int grid = grid_for(ctx, (n + kSliceThreads - 1) / kSliceThreads, 8);
        if (out_klass == SSQ_CLASS_64)
            slice_fixed_kernel<1><<<grid, kSliceThreads, 0, ctx->stream>>>(in, n, starts, stops, start0, stop0, width, (u64 *)out_words, (uint8_t *)out_lens);
        else
            slice_fixed_kernel<3><<<grid, kSliceThreads, 0, ctx->stream>>>(in, n, starts, stops, start0, stop0, width, (u64 *)out_words, (uint8_t *)out_lens);
    }
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_kmers_count(ssq_ctx *ctx, int in_klass, const void *lens, int64_t n, int32_t k, int32_t stride, int64_t *kmer_off) {
    SSQ_ARG(ctx != nullptr && kmer_off != nullptr && (n == 0 || lens != nullptr) && n >= 0, "bad arguments");
    SSQ_ARG(k >= 1 && k <= 32 && stride >= 1, "k must be 1..32 and stride >= 1");
    DeviceGuard g(ctx->device);
    return scan_kmer_counts(ctx, lens, in_klass == SSQ_CLASS_VAR ? 2 : 1, n, k, stride, kmer_off);
}

int ssq_kmers64(ssq_ctx *ctx, int in_klass, const uint64_t *words, const int64_t *word_off, const void *lens, int64_t n,
                int32_t k, int32_t stride, const int64_t *kmer_off, uint64_t *out_words, uint8_t *out_lens) {
    int rc = check_view(ctx, in_klass, words, word_off, lens, n);
    if (rc) return rc;
    SSQ_ARG(k >= 1 && k <= 32 && stride >= 1, "k must be 1..32 and stride >= 1");
    SSQ_ARG(n == 0 || (kmer_off != nullptr && out_words != nullptr && out_lens != nullptr), "NULL buffer");
    if (n == 0) return SSQ_OK;
    DeviceGuard g(ctx->device);
    const PackedView in{(const u64 *)words, word_off, lens, in_klass};
    const int grid = grid_for(ctx, (n + kSliceThreads / 32 - 1) / (kSliceThreads / 32), 8);
    kmers64_kernel<<<grid, kSliceThreads, 0, ctx->stream>>>(in, n, k, stride, kmer_off, (u64 *)out_words, out_lens);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_normalize(ssq_ctx *ctx, const uint8_t *ascii, int64_t nbytes, uint8_t *out) {
    SSQ_ARG(ctx != nullptr && nbytes >= 0 && (nbytes == 0 || (ascii != nullptr && out != nullptr)), "bad arguments");
    if (nbytes == 0) return SSQ_OK;
    DeviceGuard g(ctx->device);
    const int grid = grid_for(ctx, (nbytes / 16 + kSliceThreads) / kSliceThreads, 8);
    normalize_kernel<<<grid, kSliceThreads, 0, ctx->stream>>>(ascii, out, nbytes);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

}  // extern "C"

namespace ssq {

// ---- UMI collapse ---------------------------------------------------------------------------
// The consumer of config C5's Hamming kernel: within every group of distinct UMIs (one group per mapping position, say),
// find the UMIs within `threshold` mismatches of each other and merge them into clusters, UMI-tools style.  The
// reference only has an unfinished sketch of packed UMI objects (shortseq/umi/umi.pxd:31-55) and its README measures
// `a ^ b` against UMI-tools' edit_distance (README.md:82-88, tests/benchmark.py:125-165); the clustering rules are the
// published ones of UMI-tools (network.py):
//   directional   edge a -> b when hamming(a, b) <= threshold and count[a] >= 2 count[b] - 1; UMIs are visited in order of
//                 decreasing count (ties: input order) and every unvisited UMI claims all it can reach.  The
//                 representative of a UMI is therefore the best-ranked UMI it is reachable from, which is what the kernel
//                 computes: labels start as ranks and the minimum flows along the edges until nothing changes.
//   cluster       the same with undirected edges hamming(a, b) <= threshold: connected components, represented by their
//                 most frequent member.
// One CTA per group; all-pairs distances are recomputed in every sweep (xor + popc per pair, the C5 arithmetic) instead of
// storing an n x n adjacency.  Groups of up to 4096 UMIs are staged in shared memory; larger ones work out of global
// scratch.
constexpr int kUmiThreads = 256;
constexpr int kUmiSmemGroup = 4096;
struct UmiScratch { u32 *label, *by_rank; };

// Groups of at most 32 UMIs -- the usual case, one group per mapping position -- are clustered by ONE WARP each: a lane
// holds one UMI in registers, ranks and labels travel by shuffles, no shared memory and no block barrier (a 256-thread CTA
// per 30-UMI group left seven of its eight warps idle at every barrier).
__global__ void __launch_bounds__(kUmiThreads) umi_cluster_small_kernel(const u64 *words, const uint8_t *lens, const u64 *counts, const int64_t *group_off,
                                                                        int64_t n_groups, int threshold, int method, int64_t *rep, int64_t *n_clusters) {
    const u32 lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * kUmiThreads + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kUmiThreads) >> 5;
    for (int64_t g = warp0; g < n_groups; g += nwarps) {
        const int64_t base = group_off[g];
        const int64_t n64 = group_off[g + 1] - base;
        if (n64 > 32) continue;                                   // umi_cluster_kernel's
        const u32 n = (u32)n64;
        if (n == 0) { if (lane == 0 && n_clusters) n_clusters[g] = 0; continue; }
        const bool mine = lane < n;
        const u64 wv = mine ? words[base + lane] : 0ull;
        const u64 cv = mine ? counts[base + lane] : 0ull;
        const u32 lv = mine ? lens[base + lane] : 0xFFu;
        // rank: decreasing count, ties in input order
        u32 label = 0;
        for (u32 u = 0; u < n; u++) {
            const u64 cu = __shfl_sync(0xFFFFFFFFu, cv, u);
            label += (cu > cv || (cu == cv && u < lane)) ? 1u : 0u;
        }
        const u32 rank = label;
        // the minimum label flows along the edges until a sweep changes nothing
        for (;;) {
            u32 best = label;
            for (u32 u = 0; u < n; u++) {
                const u64 wu = __shfl_sync(0xFFFFFFFFu, wv, u);
                const u64 cu = __shfl_sync(0xFFFFFFFFu, cv, u);
                const u32 lu = __shfl_sync(0xFFFFFFFFu, lv, u);
                const u32 bu = __shfl_sync(0xFFFFFFFFu, label, u);
                if (!mine || bu >= best || lu != lv) continue;
                if (method == 0 && cu + 1 < 2 * cv) continue;            // directional: count[u] >= 2 count[v] - 1
                if (diff_bases(wu, wv) <= threshold) best = bu;
            }
            const bool changed = best < label;
            label = best;
            if (!__any_sync(0xFFFFFFFFu, changed)) break;
        }
        // representative = the UMI whose rank is this UMI's final label
        u32 rep_lane = 0;
        for (u32 u = 0; u < n; u++) {
            const u32 ru = __shfl_sync(0xFFFFFFFFu, rank, u);
            if (ru == label) rep_lane = u;
        }
        if (mine) rep[base + lane] = base + rep_lane;
        const u32 heads = __ballot_sync(0xFFFFFFFFu, mine && rep_lane == lane);
        if (lane == 0 && n_clusters) n_clusters[g] = __popc(heads);
    }
}

template <bool SKIP_SMALL>
__global__ void __launch_bounds__(kUmiThreads) umi_cluster_kernel(const u64 *words, const uint8_t *lens, const u64 *counts, const int64_t *group_off,
                                                                  int64_t n_groups, int threshold, int method, int64_t *rep, int64_t *n_clusters,
                                                                  UmiScratch sc) {
    extern __shared__ __align__(16) u64 umi_smem[];
    __shared__ int s_changed;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int64_t base = group_off[g];
        const u32 n = (u32)(group_off[g + 1] - base);
        if (SKIP_SMALL && group_off[g + 1] - base <= 32) continue;     // umi_cluster_small_kernel's (CTA-uniform: no barrier is skipped by a part of the CTA)
        if (n == 0) { if (threadIdx.x == 0 && n_clusters) n_clusters[g] = 0; continue; }
        const bool staged = n <= (u32)kUmiSmemGroup;
        // word (8) | count (8) | len (1, padded) | label (4) | by_rank (4)
        const u64 *w = words + base;
        const u64 *c = counts + base;
        const uint8_t *l = lens + base;
        u32 *label = sc.label + base, *by_rank = sc.by_rank + base;
        if (staged) {
            u64 *sw = umi_smem, *scn = umi_smem + n;
            u32 *sl = reinterpret_cast<u32 *>(umi_smem + 2 * (size_t)n);
            uint8_t *sll = reinterpret_cast<uint8_t *>(sl + 2 * (size_t)n);
            for (u32 i = threadIdx.x; i < n; i += kUmiThreads) { sw[i] = w[i]; scn[i] = c[i]; sll[i] = l[i]; }
            w = sw; c = scn; l = sll; label = sl; by_rank = sl + n;
        }
        __syncthreads();
        // rank: decreasing count, ties in input order
        for (u32 v = threadIdx.x; v < n; v += kUmiThreads) {
            const u64 cv = c[v];
            u32 r = 0;
            for (u32 u = 0; u < n; u++) r += (c[u] > cv || (c[u] == cv && u < v)) ? 1u : 0u;
            label[v] = r;
            by_rank[r] = v;
        }
        __syncthreads();
        // the minimum label flows along the edges until a sweep changes nothing
        for (;;) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            for (u32 v = threadIdx.x; v < n; v += kUmiThreads) {
                const u64 wv = w[v], cv = c[v];
                const u32 lv = l[v];
                u32 best = label[v];
                for (u32 u = 0; u < n; u++) {
                    const u32 lu = label[u];
                    if (lu >= best || l[u] != lv) continue;
                    if (method == 0 && c[u] + 1 < 2 * cv) continue;          // directional: count[u] >= 2 count[v] - 1
                    if (diff_bases(w[u], wv) <= threshold) best = lu;
                }
                if (best < label[v]) { label[v] = best; s_changed = 1; }       // labels only decrease: a stale read costs a sweep, not correctness
            }
            __syncthreads();
            const int again = s_changed;
            __syncthreads();
            if (!again) break;
        }
        u32 mine = 0;
        for (u32 v = threadIdx.x; v < n; v += kUmiThreads) {
            rep[base + v] = base + by_rank[label[v]];
            mine += by_rank[label[v]] == v ? 1u : 0u;             // v represents itself: one cluster
        }
        if (n_clusters) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            if (mine) atomicAdd(&s_changed, (int)mine);
            __syncthreads();
            if (threadIdx.x == 0) n_clusters[g] = s_changed;
        }
        __syncthreads();
    }
}

}  // namespace ssq

extern "C" {

int ssq_umi_cluster(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, const uint64_t *counts, int64_t n,
                    const int64_t *group_off, int64_t n_groups, int32_t threshold, int32_t method, int64_t *rep,
                    int64_t *n_clusters) {
    SSQ_ARG(ctx != nullptr && n >= 0 && n_groups >= 0, "bad arguments");
    SSQ_ARG(n == 0 || (words != nullptr && lens != nullptr && counts != nullptr && group_off != nullptr && rep != nullptr), "NULL buffer");
    SSQ_ARG(threshold >= 0 && threshold <= 32 && (method == 0 || method == 1), "threshold must be 0..32, method 0 (directional) or 1 (cluster)");
    if (n == 0 || n_groups == 0) return SSQ_OK;
    ssq::DeviceGuard g(ctx->device);
    void *scratch = nullptr;
    int rc = ssq::ctx_scratch(ctx, sizeof(ssq::u32) * 2 * (size_t)n, &scratch);
    if (rc) return rc;
    ssq::UmiScratch sc{(ssq::u32 *)scratch, (ssq::u32 *)scratch + n};
    const size_t smem = (size_t)ssq::kUmiSmemGroup * (8 + 8 + 4 + 4 + 1) + 16;
    static bool configured = false;
    if (!configured) {
        SSQ_CUDA(cudaFuncSetAttribute(ssq::umi_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // small groups: a warp each; everything else: a CTA each (that kernel skips the small ones)
    const int sgrid = ssq::grid_for(ctx, (n_groups + ssq::kUmiThreads / 32 - 1) / (ssq::kUmiThreads / 32), 8);
    ssq::umi_cluster_small_kernel<<<sgrid, ssq::kUmiThreads, 0, ctx->stream>>>((const ssq::u64 *)words, lens, (const ssq::u64 *)counts, group_off, n_groups,
                                                                             threshold, method, rep, n_clusters);
    SSQ_LAUNCH_CHECK();
    const int grid = ssq::grid_for(ctx, n_groups, 2);
    ssq::umi_cluster_kernel<true><<<grid, ssq::kUmiThreads, smem, ctx->stream>>>((const ssq::u64 *)words, lens, (const ssq::u64 *)counts, group_off, n_groups,
                                                                               threshold, method, rep, n_clusters, sc);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

}  // extern "C"
