// ssq_scan.cu -- exclusive prefix sums over per-read sizes (int64 results).
//
// Three launches per level: per-block totals, scan of the totals (recursive),
// then the per-element exclusive scan.  The value of element i comes from a
// functor, so the same code turns lengths into byte offsets (decode output),
// read lengths into ShortSeqVar word offsets, and synthetic lengths into
// offsets.
#include "ssq_internal.h"

namespace ssq {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

struct LenU8 { const uint8_t *p; __device__ int64_t operator()(int64_t i) const { return p[i]; } };
struct LenU16 { const uint16_t *p; __device__ int64_t operator()(int64_t i) const { return p[i]; } };
struct I64 { const int64_t *p; __device__ int64_t operator()(int64_t i) const { return p[i]; } };
struct CountU32 { const u32 *p; __device__ int64_t operator()(int64_t i) const { return p[i]; } };
// ceil(len/32) words of a ShortSeqVar read (reference util.pyx:29-33); lengths outside
// 0..1024 contribute 0 words (the pack kernel reports them).
struct VarWords {
    const int64_t *off;
    __device__ int64_t operator()(int64_t i) const {
        int64_t len = off[i + 1] - off[i];
        return (len < 0 || len > 1024) ? 0 : (len + 31) >> 5;
    }
};
// FASTQ ingest (ssq_fastq.cu): 1 for the reads of a length class; length of the j-th selected read
struct FastqClassFlag {
    const int64_t *starts, *ends; int klass;
    __device__ int64_t operator()(int64_t k) const {
        const int64_t len = ends[k] - starts[k];
        return (klass == SSQ_CLASS_64 ? (len >= 0 && len <= 32) : (len >= 33 && len <= 96)) ? 1 : 0;
    }
};
struct FastqSelLen {
    const int64_t *starts, *ends, *sel;
    __device__ int64_t operator()(int64_t j) const { const int64_t k = sel ? sel[j] : j; return ends[k] - starts[k]; }
};
struct SynthLen {
    uint64_t seed; int64_t first, n_keys; int32_t lo, hi;
    __device__ int64_t operator()(int64_t i) const {
        if (hi <= lo) return lo;
        u64 key = mix64(seed + (u64)(first + i)) % (u64)n_keys;
        return lo + (int64_t)(mix64(seed + 0x5EED0003ull + key) % (u64)(hi - lo + 1));
    }
};

// k-mers of a read: (len - k) / stride + 1 when len >= k (ssq_slice.cu)
struct KmerCount {
    const void *lens; int len_bytes; int k, stride;
    __device__ int64_t operator()(int64_t i) const {
        const int64_t len = len_bytes == 2 ? ((const uint16_t *)lens)[i] : ((const uint8_t *)lens)[i];
        return len >= k ? (len - k) / stride + 1 : 0;
    }
};
// 64-bit words of the slice [start, stop) of a read, clamped like ssq_slice does
struct SliceWords {
    const void *lens; int len_bytes; const int64_t *starts, *stops; int64_t start0, stop0; int32_t width;
    __device__ int64_t operator()(int64_t i) const {
        const int64_t len = len_bytes == 2 ? ((const uint16_t *)lens)[i] : ((const uint8_t *)lens)[i];
        int64_t a = starts ? starts[i] : start0, b = stops ? stops[i] : stop0;
        a = a < 0 ? 0 : (a > len ? len : a);
        b = b > len ? len : b;
        if (b < a) b = a;
        if (b - a > width) b = a + width;
        return (b - a + 31) >> 5;
    }
};

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *smem /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    int64_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        int64_t s = smem[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_totals_kernel(F f, int64_t n, int64_t *totals) {
    __shared__ int64_t smem[8];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += f(base + k);
    int64_t tot;
    block_exclusive_scan(s, &tot, smem);
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

// out[i] = block_base + exclusive scan; the last block also writes out[n] = grand total.
template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(F f, int64_t n, const int64_t *block_base, int64_t *out,
                                                                  bool write_total) {
    __shared__ int64_t smem[8];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = base + k < n ? f(base + k) : 0;
        s += v[k];
    }
    int64_t tot;
    int64_t run = block_exclusive_scan(s, &tot, smem) + (block_base ? block_base[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = run;
}

// scratch entries (int64) a scan of n elements needs for its block totals, all levels
static int64_t scan_scratch_entries(int64_t n) {
    int64_t total = 0;
    while (n > kScanTile) {
        const int64_t nblocks = (n + kScanTile - 1) / kScanTile;
        total += 2 * nblocks + 2;
        n = nblocks;
    }
    return total;
}

template <class F>
static int scan_level(ssq_ctx *ctx, F f, int64_t n, int64_t *out /*[n+1]*/, int64_t *scratch) {
    cudaStream_t st = ctx->stream;
    int64_t nblocks = (n + kScanTile - 1) / kScanTile;
    if (nblocks == 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, st>>>(f, n, (const int64_t *)nullptr, out, true);
        SSQ_LAUNCH_CHECK();
        return SSQ_OK;
    }
    int64_t *totals = scratch;            // [nblocks] totals, then [nblocks+1] scanned (+1 pad)
    int64_t *scanned = totals + nblocks;
    scan_totals_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>>(f, n, totals);
    SSQ_LAUNCH_CHECK();
    int rc = scan_level(ctx, I64{totals}, nblocks, scanned, scratch + 2 * nblocks + 2);
    if (rc) return rc;
    scan_apply_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>>(f, n, scanned, out, true);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

template <class F>
static int scan_exclusive(ssq_ctx *ctx, F f, int64_t n, int64_t *out /*[n+1]*/) {
    if (n <= 0) {
        SSQ_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), ctx->stream));
        return SSQ_OK;
    }
    void *scratch = nullptr;
    int rc = ctx_scratch(ctx, sizeof(int64_t) * (size_t)(scan_scratch_entries(n) + 2), &scratch);
    if (rc) return rc;
    return scan_level(ctx, f, n, out, (int64_t *)scratch);
}

int scan_lens_to_offsets(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, int64_t *out) {
    if (len_bytes == 1) return scan_exclusive(ctx, LenU8{(const uint8_t *)lens}, n, out);
    return scan_exclusive(ctx, LenU16{(const uint16_t *)lens}, n, out);
}

int scan_i64(ssq_ctx *ctx, const int64_t *values, int64_t n, int64_t *out) {
    return scan_exclusive(ctx, I64{values}, n, out);
}

int scan_u32_counts(ssq_ctx *ctx, const u32 *counts, int64_t n, int64_t *out) {
    return scan_exclusive(ctx, CountU32{counts}, n, out);
}

int scan_fastq_flags(ssq_ctx *ctx, const int64_t *starts, const int64_t *ends, int klass, int64_t n, int64_t *out) {
    return scan_exclusive(ctx, FastqClassFlag{starts, ends, klass}, n, out);
}

int scan_fastq_lens(ssq_ctx *ctx, const int64_t *starts, const int64_t *ends, const int64_t *sel, int64_t n, int64_t *out) {
    return scan_exclusive(ctx, FastqSelLen{starts, ends, sel}, n, out);
}

int scan_kmer_counts(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, int k, int stride, int64_t *out) {
    return scan_exclusive(ctx, KmerCount{lens, len_bytes, k, stride}, n, out);
}

int scan_slice_words(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, const int64_t *starts, const int64_t *stops, int64_t start0,
                     int64_t stop0, int32_t width, int64_t *out) {
    return scan_exclusive(ctx, SliceWords{lens, len_bytes, starts, stops, start0, stop0, width}, n, out);
}

int scan_var_words(ssq_ctx *ctx, const int64_t *offsets, int64_t n, int64_t *word_off) {
    return scan_exclusive(ctx, VarWords{offsets}, n, word_off);
}

int scan_synth_lens(ssq_ctx *ctx, uint64_t seed, int64_t first_read, int64_t n, int64_t n_keys, int32_t len_lo,
                    int32_t len_hi, int64_t *offsets) {
    return scan_exclusive(ctx, SynthLen{seed, first_read, n_keys, len_lo, len_hi}, n, offsets);
}

}  // namespace ssq

using namespace ssq;

extern "C" int ssq_lens_to_offsets(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, int64_t *out_offsets) {
    SSQ_ARG(ctx != nullptr && out_offsets != nullptr, "NULL argument");
    SSQ_ARG(len_bytes == 1 || len_bytes == 2, "len_bytes must be 1 or 2");
    SSQ_ARG(n >= 0 && (n == 0 || lens != nullptr), "bad n / lens");
    DeviceGuard g(ctx->device);
    return scan_lens_to_offsets(ctx, lens, len_bytes, n, out_offsets);
}
