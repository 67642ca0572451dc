// ssq_codec.cu -- batched decode (2-bit -> ASCII), Hamming distance, synthetic reads.
#include <stdlib.h>
#include "ssq_internal.h"

namespace ssq {

constexpr int kThreads = 256;

// =============================================================================================
// Decode.  Replaces the reference's per-object __str__ (short_seq_64.pyx:114-121,
// short_seq_192.pyx:114-127, short_seq_var.pyx:98-120; charmap "ACTG", util.pyx:52).
//
// Mirror image of the pack kernel: the tile's output bytes form one contiguous range.
//   A. every read ORs its 2*len code bits into a shared-memory bit stream at the bit position
//      of its first output byte (reads abut at arbitrary 2-bit positions, hence atomicOr);
//   B. the stream is expanded 32 bits -> 16 ASCII bytes per thread with byte-permute lookups
//      and written with aligned 16-byte stores (edge chunks byte-wise: they are shared with
//      the neighbouring tiles).
// =============================================================================================

// 8 bits of codes (4 bases) -> 4 ASCII bytes.
__device__ __forceinline__ u32 expand4(u32 p) {
    u32 x = (p | (p << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;           // code k in nibble k
    return __byte_perm(0x47544341u, 0u, x); // "ACTG"[code]
}

__device__ __forceinline__ void deposit64(u32 *stream, int64_t bit, u64 w) {
    int wi = (int)(bit >> 5);
    u32 sh = (u32)bit & 31;
    u32 lo = (u32)w, hi = (u32)(w >> 32);
    u32 a = lo << sh;
    u32 b = __funnelshift_l(lo, hi, sh);
    u32 c = sh ? hi >> (32 - sh) : 0u;
    if (a) atomicOr(&stream[wi], a);
    if (b) atomicOr(&stream[wi + 1], b);
    if (c) atomicOr(&stream[wi + 2], c);
}

// Expand stream[] (covering output bytes from index a0, chunk c = 16 bytes) into out[t0, t1) and leave the words it
// read (plus the `pad` words deposits may have spilled into) zeroed for the next tile.
// Only the first and the last chunk of a tile can be partial (they are shared with the neighbouring tiles).  They are
// written by 32 lanes, one byte each: the first version let the thread that owned such a chunk run a 16-step loop of
// predicated byte stores while the other 255 threads waited at the barrier behind it (ncu on decode_var2: 38 % of all
// stall samples at that barrier, 8 % of the instructions in the byte loop).
template <int THREADS>
__device__ __forceinline__ void store_tile(uint8_t *out, int64_t a0, int64_t t0, int64_t t1, u32 *stream, int pad = 0) {
    const int nchunks = (int)((t1 - a0 + 15) >> 4);
    for (int c = threadIdx.x; c < nchunks; c += THREADS) {
        const int64_t idx = a0 + 16 * (int64_t)c;
        if (idx >= t0 && idx + 16 <= t1) {
            const u32 s = stream[c];
            stream[c] = 0;
            *reinterpret_cast<uint4 *>(out + idx) = make_uint4(expand4(s & 0xFF), expand4((s >> 8) & 0xFF), expand4((s >> 16) & 0xFF), expand4(s >> 24));
        }
    }
    if (threadIdx.x < 32 && nchunks > 0) {
        const int lane = threadIdx.x;
        const int c = lane < 16 ? 0 : nchunks - 1;
        const int64_t idx = a0 + 16 * (int64_t)c;
        const bool partial = !(idx >= t0 && idx + 16 <= t1) && (lane < 16 || nchunks > 1);
        const u32 s = partial ? stream[c] : 0u;
        const int64_t j = idx + (lane & 15);
        if (partial && j >= t0 && j < t1) out[j] = (uint8_t)(0x47544341u >> (8 * ((s >> (2 * (lane & 15))) & 3u)));
        __syncwarp();
        if (partial && (lane & 15) == 0) stream[c] = 0;
    }
    if ((int)threadIdx.x < pad) stream[nchunks + threadIdx.x] = 0;
}

// Persistent CTAs, 2 reads per thread per tile, software-pipelined like the pack kernel: the words, lengths and
// offsets of the CTA's next tile are loaded while the current tile is deposited and stored.  Two barriers per tile:
// the store phase clears the stream words it consumed.
constexpr int kDecRPT = 2;
constexpr int kDecTile = kThreads * kDecRPT;

template <int W>
struct DecTile {
    int64_t t0, t1;
    int64_t o0[kDecRPT];
    u64 w[kDecRPT][W];
    int len[kDecRPT];
    int nreads;
};

template <int W>
__device__ __forceinline__ DecTile<W> load_dec_tile(const u64 *words, const uint8_t *lens, const int64_t *out_off, int64_t n, int64_t tile) {
    constexpr int MAXLEN = 32 * W;
    DecTile<W> d;
    const int64_t first = tile * kDecTile;
    d.nreads = first < n ? (int)min((int64_t)kDecTile, n - first) : 0;
    d.t0 = d.t1 = 0;
    if (d.nreads > 0) { d.t0 = out_off[first]; d.t1 = out_off[first + d.nreads]; }
#pragma unroll
    for (int k = 0; k < kDecRPT; k++) {
        const int r = threadIdx.x + k * kThreads;
        d.len[k] = -1;
        d.o0[k] = 0;
        if (r < d.nreads) {
            const int64_t i = first + r;
            d.len[k] = min((int)lens[i], MAXLEN);
            d.o0[k] = out_off[i];
#pragma unroll
            for (int j = 0; j < W; j++) d.w[k][j] = words[(size_t)i * W + j];
        }
    }
    return d;
}

template <int W>
__global__ void __launch_bounds__(kThreads) decode_fixed_kernel(const u64 *words, const uint8_t *lens, int64_t n,
                                                                const int64_t *out_off, uint8_t *out) {
    constexpr int MAXLEN = 32 * W;
    constexpr int MAX_CHUNKS = (kDecTile * MAXLEN + 30) / 16 + 1;
    __shared__ u32 stream[MAX_CHUNKS + 3];
    const int64_t mis = (int64_t)((uintptr_t)out & 15);
    const int64_t ntiles = (n + kDecTile - 1) / kDecTile;
    for (int c = threadIdx.x; c < MAX_CHUNKS + 3; c += kThreads) stream[c] = 0;
    DecTile<W> cur = load_dec_tile<W>(words, lens, out_off, n, blockIdx.x);
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = cur.t0, t1 = cur.t1;
        const bool sane = t1 >= t0 && t1 - t0 <= (int64_t)kDecTile * MAXLEN;    // inconsistent offsets: nothing sane to write
        const int64_t a0 = ((t0 + mis) & ~(int64_t)15) - mis;
        if (sane) {
#pragma unroll
            for (int k = 0; k < kDecRPT; k++) {
                const int len = cur.len[k];
                const int64_t o0 = cur.o0[k];
                if (len >= 0 && o0 >= t0 && o0 + len <= t1) {
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        const int nb = 2 * len - 64 * j;
                        if (nb > 0) {
                            u64 w = cur.w[k][j];
                            if (nb < 64) w &= (1ull << nb) - 1;
                            deposit64(stream, 2 * (o0 - a0) + 64 * j, w);
                        }
                    }
                }
            }
        }
        const DecTile<W> nxt = load_dec_tile<W>(words, lens, out_off, n, tile + gridDim.x);   // in flight during the store phase
        __syncthreads();
        if (sane) store_tile<kThreads>(out, a0, t0, t1, stream, 3);
        __syncthreads();
        cur = nxt;
    }
}

// ---- fused offsets: decode without a full offsets scan -------------------------------------------
// decode_batch used to scan all n lengths into n+1 offsets (4 GB written, then read back by the decode kernel, for
// 5e8 reads).  Instead: one pass sums the lengths of every 512-read tile, the ~n/512 tile totals are scanned, and the
// decode kernel derives each read's offset with a block scan of the tile's lengths -- writing out_offsets (a required
// output) as a by-product.
__global__ void __launch_bounds__(kThreads) tile_totals_kernel(const uint8_t *lens, int64_t n, int max_len, int64_t *totals) {
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (n + kDecTile - 1) / kDecTile;
    const int64_t warp0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
    const bool aligned = ((uintptr_t)lens & 15) == 0;
    for (int64_t tile = warp0; tile < ntiles; tile += nwarps) {
        const int64_t at = tile * kDecTile + lane * 16;
        u32 sum = 0;
        if (aligned && at + 16 <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(lens + at);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++) sum += min((w[k] >> (8 * b)) & 0xFFu, (u32)max_len);
        } else {
            for (int b = 0; b < 16; b++)
                if (at + b < n) sum += min((u32)lens[at + b], (u32)max_len);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) totals[tile] = (int64_t)sum;
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads) decode_fixed_fused_kernel(const u64 *words, const uint8_t *lens, int64_t n,
                                                                      const int64_t *tile_base, int64_t *out_off, uint8_t *out) {
    constexpr int MAXLEN = 32 * W;
    constexpr int MAX_CHUNKS = (kDecTile * MAXLEN + 30) / 16 + 1;
    __shared__ u32 stream[MAX_CHUNKS + 3];
    __shared__ u32 s_w[kDecRPT][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t mis = (int64_t)((uintptr_t)out & 15);
    const int64_t ntiles = (n + kDecTile - 1) / kDecTile;
    for (int c = threadIdx.x; c < MAX_CHUNKS + 3; c += kThreads) stream[c] = 0;
    // per-tile state: this thread's reads' lengths and words (the next tile's are loaded a tile ahead)
    auto load = [&](int64_t tile, int (&len)[kDecRPT], u64 (&w)[kDecRPT][W], int64_t &t0, int64_t &t1) {
        const int64_t first = tile * kDecTile;
        t0 = t1 = 0;
        if (tile < ntiles) { t0 = tile_base[tile]; t1 = tile_base[tile + 1]; }
#pragma unroll
        for (int k = 0; k < kDecRPT; k++) {
            const int64_t i = first + threadIdx.x + k * kThreads;
            len[k] = -1;
            if (tile < ntiles && i < n) {
                len[k] = min((int)lens[i], MAXLEN);
#pragma unroll
                for (int j = 0; j < W; j++) w[k][j] = words[(size_t)i * W + j];
            }
        }
    };
    int clen[kDecRPT]; u64 cw[kDecRPT][W]; int64_t t0, t1;
    load(blockIdx.x, clen, cw, t0, t1);
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kDecTile;
        // offsets of this tile's reads: block scan of their lengths, reads [0, 256) then [256, 512)
        u32 incl[kDecRPT];
#pragma unroll
        for (int k = 0; k < kDecRPT; k++) {
            u32 v = clen[k] > 0 ? (u32)clen[k] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += x; }
            incl[k] = v;
            if (lane == 31) s_w[k][warp] = v;
        }
        __syncthreads();
        int64_t o0[kDecRPT];
        {
            u32 run = 0;
#pragma unroll
            for (int k = 0; k < kDecRPT; k++) {
                u32 before = 0, total = 0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; w++) { const u32 x = s_w[k][w]; if (w < warp) before += x; total += x; }
                o0[k] = t0 + run + before + incl[k] - (clen[k] > 0 ? (u32)clen[k] : 0u);
                run += total;
            }
        }
        const int64_t a0 = ((t0 + mis) & ~(int64_t)15) - mis;
#pragma unroll
        for (int k = 0; k < kDecRPT; k++) {
            const int len = clen[k];
            if (len < 0) continue;
            const int64_t i = first + threadIdx.x + k * kThreads;
            out_off[i] = o0[k];
            if (i == n - 1) out_off[n] = o0[k] + len;
#pragma unroll
            for (int j = 0; j < W; j++) {
                const int nb = 2 * len - 64 * j;
                if (nb > 0) {
                    u64 w = cw[k][j];
                    if (nb < 64) w &= (1ull << nb) - 1;
                    deposit64(stream, 2 * (o0[k] - a0) + 64 * j, w);
                }
            }
        }
        int nlen[kDecRPT]; u64 nw[kDecRPT][W]; int64_t nt0, nt1;
        load(tile + gridDim.x, nlen, nw, nt0, nt1);                      // in flight during the store phase
        __syncthreads();
        store_tile<kThreads>(out, a0, t0, t1, stream, 3);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kDecRPT; k++) {
            clen[k] = nlen[k];
#pragma unroll
            for (int j = 0; j < W; j++) cw[k][j] = nw[k][j];
        }
        t0 = nt0; t1 = nt1;
    }
}

constexpr int kVarTileReads = 32;
constexpr int kVarMaxChunks = (kVarTileReads * 1024 + 30) / 16 + 1;

__global__ void __launch_bounds__(kThreads) decode_var_kernel(const u64 *words, const int64_t *word_off,
                                                              const uint16_t *lens, int64_t n, const int64_t *out_off,
                                                              uint8_t *out) {
    __shared__ u32 stream[kVarMaxChunks + 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t mis = (int64_t)((uintptr_t)out & 15);
    const int64_t ntiles = (n + kVarTileReads - 1) / kVarTileReads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kVarTileReads;
        const int nreads = (int)min((int64_t)kVarTileReads, n - first);
        const int64_t t0 = out_off[first], t1 = out_off[first + nreads];
        if (t1 < t0 || t1 - t0 > (int64_t)kVarTileReads * 1024) continue;
        const int64_t a0 = ((t0 + mis) & ~(int64_t)15) - mis;
        const int nchunks = (int)((t1 - a0 + 15) >> 4);
        for (int c = threadIdx.x; c < nchunks + 3; c += kThreads) stream[c] = 0;
        __syncthreads();
        for (int r = warp; r < nreads; r += kThreads / 32) {
            const int64_t i = first + r;
            const int len = min((int)lens[i], 1024);
            const int64_t o0 = out_off[i];
            const int nb = 2 * len - 64 * lane;
            if (nb > 0 && o0 >= t0 && o0 + len <= t1) {
                u64 w = words[word_off[i] + lane];
                if (nb < 64) w &= (1ull << nb) - 1;
                deposit64(stream, 2 * (o0 - a0) + 64 * lane, w);
            }
        }
        __syncthreads();
        store_tile<kThreads>(out, a0, t0, t1, stream);
        __syncthreads();
    }
}

// ShortSeqVar decode, second version.  decode_var_kernel gives a read to a warp and a word to a lane (5 of 32 lanes
// busy on a 150-nt read) and chains four dependent global loads per read (length -> output offset -> word offset ->
// word).  Here a persistent CTA treats the tile's words as ONE flat list -- they are contiguous in the CSR array -- that
// is loaded coalesced one tile ahead; thread k finds the read of word k with a 5-step search in the tile's 33 word
// offsets and deposits it.  A ninth PRODUCER warp owns the per-tile serial work (offsets loaded two tiles ahead,
// relative offsets, sanity checks) and publishes each tile's metadata in shared memory behind an mbarrier, running up to
// kVar2Meta - 1 tiles ahead; the eight consumer warps synchronise among themselves with a named barrier.  (With warp 0
// publishing between the CTA barriers every tile waited for it: 7 of every 8 stall cycles were barrier waits.)
constexpr int kVar2Reads = 32;
constexpr int kVar2Meta = 4;
constexpr int kVar2WPT = (kVar2Reads * 32) / kThreads;      // words per thread per tile (4)
constexpr int kVar2Threads = kThreads + 32;                 // consumers + the producer warp
struct DecVar2Meta {
    int64_t t0, t1, wbase;
    u32 orel[kVar2Reads + 1];   // first output byte of each read relative to t0 (0xFFFFFFFF: outside the tile)
    u32 wrel[kVar2Reads + 1];   // first word of each read relative to wbase
    u32 len[kVar2Reads];
    int nreads, sane;
};

__global__ void __launch_bounds__(kVar2Threads) decode_var2_kernel(const u64 *words, const int64_t *word_off,
                                                                   const uint16_t *lens, int64_t n, const int64_t *out_off,
                                                                   uint8_t *out) {
    __shared__ u32 stream[kVarMaxChunks + 3];
    __shared__ DecVar2Meta meta[kVar2Meta];
    __shared__ __align__(8) u64 full[kVar2Meta], freed[kVar2Meta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t mis = (int64_t)((uintptr_t)out & 15);
    const int64_t ntiles = (n + kVar2Reads - 1) / kVar2Reads;
    const int64_t stride = gridDim.x;
    const int mytiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + stride - 1) / stride) : 0;
    for (int c = threadIdx.x; c < kVarMaxChunks + 3; c += kVar2Threads) stream[c] = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int m = 0; m < kVar2Meta; m++) { mbar_init(smem_addr(&full[m]), 1); mbar_init(smem_addr(&freed[m]), kThreads / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kThreads / 32) {
        // ---- producer warp
        struct MetaRegs { int64_t p_o = 0, p_w = 0, p_ol = 0, p_wl = 0; u32 p_len = 0; } regs_a, regs_b;
        auto fetch_meta = [&](int j, MetaRegs &R) {
            if (j >= mytiles) return;
            const int64_t first = ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads;
            const int nreads = (int)min((int64_t)kVar2Reads, n - first);
            if (lane < nreads) { R.p_o = out_off[first + lane]; R.p_w = word_off[first + lane]; R.p_len = min((u32)lens[first + lane], 1024u); }
            if (lane == 0) { R.p_ol = out_off[first + nreads]; R.p_wl = word_off[first + nreads]; }
        };
        auto publish = [&](int j, const MetaRegs &R) {
            const int ms = j % kVar2Meta;
            if (j >= kVar2Meta) mbar_wait(smem_addr(&freed[ms]), (u32)(j / kVar2Meta - 1) & 1u);     // tile j - kVar2Meta is done with the slot
            const int64_t first = ((int64_t)blockIdx.x + (int64_t)j * stride) * kVar2Reads;
            const int nreads = (int)min((int64_t)kVar2Reads, n - first);
            DecVar2Meta &m = meta[ms];
            const int64_t t0 = __shfl_sync(0xFFFFFFFFu, R.p_o, 0), wbase = __shfl_sync(0xFFFFFFFFu, R.p_w, 0);
            const int64_t t1 = __shfl_sync(0xFFFFFFFFu, R.p_ol, 0), wend = __shfl_sync(0xFFFFFFFFu, R.p_wl, 0);
            if (lane < nreads) {
                m.orel[lane] = (R.p_o >= t0 && R.p_o + R.p_len <= t1) ? (u32)(R.p_o - t0) : 0xFFFFFFFFu;
                m.wrel[lane] = (u32)min(max(R.p_w - wbase, (int64_t)0), (int64_t)kVar2Reads * 32);
                m.len[lane] = R.p_len;
            }
            if (lane == 0) {
                m.wrel[nreads] = (u32)min(max(wend - wbase, (int64_t)0), (int64_t)kVar2Reads * 32);
                m.t0 = t0; m.t1 = t1; m.wbase = wbase; m.nreads = nreads;
                m.sane = t1 >= t0 && t1 - t0 <= (int64_t)kVar2Reads * 1024;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&full[ms]));
        };
        fetch_meta(0, regs_a);
        fetch_meta(1, regs_b);
        for (int j = 0; j < mytiles; j += 2) {
            publish(j, regs_a);
            fetch_meta(j + 2, regs_a);
            if (j + 1 < mytiles) {
                publish(j + 1, regs_b);
                fetch_meta(j + 3, regs_b);
            }
        }
        return;
    }

    // ---- consumer warps
    auto load_words = [&](int j, u64 (&w)[kVar2WPT]) {   // the words of tile j, once its metadata is published
#pragma unroll
        for (int i = 0; i < kVar2WPT; i++) w[i] = 0;
        if (j >= mytiles) return;
        mbar_wait(smem_addr(&full[j % kVar2Meta]), (u32)(j / kVar2Meta) & 1u);
        const DecVar2Meta &m = meta[j % kVar2Meta];
        const u32 tw = m.wrel[m.nreads];
        const u64 *src = words + m.wbase;
#pragma unroll
        for (int i = 0; i < kVar2WPT; i++) {
            const u32 k = threadIdx.x + i * kThreads;
            if (k < tw) w[i] = src[k];
        }
    };
    u64 cw[kVar2WPT];
    load_words(0, cw);
    for (int j = 0; j < mytiles; j++) {
        const DecVar2Meta &m = meta[j % kVar2Meta];
        const int64_t t0 = m.t0, t1 = m.t1;
        const int nreads = m.nreads;
        const bool sane = m.sane != 0;
        const int64_t a0 = ((t0 + mis) & ~(int64_t)15) - mis;
        if (sane) {
            const u32 tw = m.wrel[nreads];
#pragma unroll
            for (int i = 0; i < kVar2WPT; i++) {
                const u32 k = threadIdx.x + i * kThreads;
                if (k >= tw) continue;
                int lo_ = 0, hi_ = nreads;                    // wrel[lo_] <= k < wrel[hi_]
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const int mid = (lo_ + hi_) >> 1;
                    if (m.wrel[mid] <= k) lo_ = mid; else hi_ = mid;
                }
                const int jw = (int)(k - m.wrel[lo_]);
                const int nb = 2 * (int)m.len[lo_] - 64 * jw;
                const u32 o = m.orel[lo_];
                if (nb > 0 && o != 0xFFFFFFFFu) {
                    u64 w = cw[i];
                    if (nb < 64) w &= (1ull << nb) - 1;
                    deposit64(stream, 2 * ((int64_t)o + (t0 - a0)) + 64 * jw, w);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&freed[j % kVar2Meta]));     // t0 / t1 / a0 are in registers: the slot is free
        named_bar_sync(1, kThreads);                                       // every deposit of the tile has landed
        u64 nw[kVar2WPT];
        load_words(j + 1, nw);                                             // in flight during the store phase
        if (sane) store_tile<kThreads>(out, a0, t0, t1, stream, 3);
        named_bar_sync(1, kThreads);                                       // the stream is clean again
#pragma unroll
        for (int i = 0; i < kVar2WPT; i++) cw[i] = nw[i];
    }
}

// =============================================================================================
// Hamming distance.  Replaces __xor__ (short_seq_64.pyx:77-84, short_seq_192.pyx:74-91,
// short_seq_var.pyx:64-81).  Canonical packed words keep bits beyond 2*len zero,
// so whole-word popcounts equal the reference's sum over ceil(len/32) blocks.
// =============================================================================================
__global__ void __launch_bounds__(kThreads) hamming64_kernel(const u64 *a, const uint8_t *la, const u64 *b,
                                                             const uint8_t *lb, int64_t n, uint8_t *dist, DevReport *rep) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        int d = diff_bases(a[i], b[i]);
        if (la[i] != lb[i]) { atomicMin(&rep->first_len_mismatch, (u64)i); d = 0xFF; }
        dist[i] = (uint8_t)d;
    }
}

// 3 words per read: the flat word arrays are swept coalesced, per-word differences go through
// shared memory and one thread per read adds its three.
__global__ void __launch_bounds__(kThreads) hamming192_kernel(const u64 *a, const uint8_t *la, const u64 *b,
                                                              const uint8_t *lb, int64_t n, uint8_t *dist, DevReport *rep) {
    __shared__ uint8_t d[3 * kThreads];
    const int64_t ntiles = (n + kThreads - 1) / kThreads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kThreads;
        const int nreads = (int)min((int64_t)kThreads, n - first);
        for (int k = threadIdx.x; k < 3 * nreads; k += kThreads)
            d[k] = (uint8_t)diff_bases(a[3 * first + k], b[3 * first + k]);
        __syncthreads();
        if ((int)threadIdx.x < nreads) {
            const int64_t i = first + threadIdx.x;
            int s = d[3 * threadIdx.x] + d[3 * threadIdx.x + 1] + d[3 * threadIdx.x + 2];
            if (la[i] != lb[i]) { atomicMin(&rep->first_len_mismatch, (u64)i); s = 0xFF; }
            dist[i] = (uint8_t)s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) hammingvar_kernel(const u64 *a, const int64_t *aoff, const uint16_t *la,
                                                              const u64 *b, const int64_t *boff, const uint16_t *lb,
                                                              int64_t n, uint16_t *dist, DevReport *rep) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int len = la[i];
        int d = 0;
        if (len != lb[i]) {
            if (lane == 0) { atomicMin(&rep->first_len_mismatch, (u64)i); dist[i] = 0xFFFF; }
            continue;
        }
        const int nw = (len + 31) >> 5;
        if (lane < nw) d = diff_bases(a[aoff[i] + lane], b[boff[i] + lane]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, s);
        if (lane == 0) dist[i] = (uint16_t)d;
    }
}

// Query x reference set.  The refs (words + lengths) are staged in shared memory in chunks;
// every thread owns one query and scans the chunk.
constexpr int kRefChunk = 1024;       // the (distance << 10 | index) packing below relies on it
static_assert(kRefChunk <= 1024, "chunk index must fit in 10 bits");
template <int W>
__global__ void __launch_bounds__(kThreads) refset_kernel(const u64 *q, const uint8_t *lq, int64_t nq, const u64 *refs,
                                                          const uint8_t *lr, int nr, int thresh, uint8_t *min_dist,
                                                          u32 *argmin, u32 *n_within) {
    __shared__ u64 s_ref[kRefChunk * W];
    __shared__ uint8_t s_len[kRefChunk];
    const int64_t ntiles = (nq + kThreads - 1) / kThreads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i = tile * kThreads + threadIdx.x;
        const bool mine = i < nq;
        u64 qw[W];
        int qlen = -1;
        if (mine) {
#pragma unroll
            for (int k = 0; k < W; k++) qw[k] = q[(size_t)i * W + k];
            qlen = lq[i];
        }
        int best = 255, within = 0;
        u32 best_j = 0xFFFFFFFFu;
        for (int r0 = 0; r0 < nr; r0 += kRefChunk) {
            const int cnt = min(kRefChunk, nr - r0);
            __syncthreads();
            for (int k = threadIdx.x; k < cnt * W; k += kThreads) s_ref[k] = refs[(size_t)r0 * W + k];
            int same = 1;
            const int len0 = lr[r0];
            for (int k = threadIdx.x; k < cnt; k += kThreads) { const uint8_t l = lr[r0 + k]; s_len[k] = l; same &= l == len0; }
            const int uniform = __syncthreads_and(same);      // the usual case: every reference of the chunk has one length
            if (mine && uniform) {
                if (qlen == len0) {
                    // (distance, index in chunk) travel as one number, so that min + argmin are one IMNMX per reference
                    // (the smaller index wins a tie, as `d < best` in index order does)
                    u32 bestk = 0xFFFFFFFFu;
                    if (W == 1 && len0 <= 16) {
                        // UMIs: <= 16 bases live in the low 32 bits (bits >= 2 len are zero): half the XOR / collapse work
                        const u32 q32 = (u32)qw[0];
                        const u32 *r32 = reinterpret_cast<const u32 *>(s_ref);
#pragma unroll 8
                        for (int j = 0; j < cnt; j++) {
                            const u32 x = q32 ^ r32[2 * j];
                            const u32 d = __popc(((x >> 1) | x) & 0x55555555u);
                            bestk = min(bestk, (d << 10) | (u32)j);
                            within += (int)d <= thresh ? 1 : 0;
                        }
                    } else {
#pragma unroll 4
                        for (int j = 0; j < cnt; j++) {
                            u32 d = 0;
#pragma unroll
                            for (int k = 0; k < W; k++) d += (u32)diff_bases(qw[k], s_ref[j * W + k]);
                            bestk = min(bestk, (d << 10) | (u32)j);
                            within += (int)d <= thresh ? 1 : 0;
                        }
                    }
                    if (cnt > 0 && (int)(bestk >> 10) < best) { best = (int)(bestk >> 10); best_j = (u32)r0 + (bestk & 1023u); }
                }
            } else if (mine) {
                for (int j = 0; j < cnt; j++) {
                    if (s_len[j] != qlen) continue;
                    int d = 0;
#pragma unroll
                    for (int k = 0; k < W; k++) d += diff_bases(qw[k], s_ref[j * W + k]);
                    if (d < best) { best = d; best_j = (u32)(r0 + j); }
                    within += d <= thresh ? 1 : 0;
                }
            }
        }
        if (mine) {
            min_dist[i] = (uint8_t)best;
            argmin[i] = best_j;
            if (n_within) n_within[i] = (u32)within;
        }
    }
}

// =============================================================================================
// Synthetic reads (measurement tooling; same generator as oracle/ssq_oracle.c).
// =============================================================================================
__device__ __forceinline__ u64 synth_key(u64 seed, int64_t i, int64_t n_keys) { return mix64(seed + (u64)i) % (u64)n_keys; }

// One thread per (read, 32-base block).
__global__ void __launch_bounds__(kThreads) synth_fill_kernel(u64 seed, int64_t first_read, int64_t n, int64_t n_keys,
                                                              int blocks_per_read, const int64_t *offsets, uint8_t *ascii) {
    const int64_t total = n * blocks_per_read;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        const int64_t i = t / blocks_per_read;
        const int blk = (int)(t - i * blocks_per_read);
        const int64_t o0 = offsets[i];
        const int len = (int)(offsets[i + 1] - o0);
        const int nb = min(32, len - 32 * blk);
        if (nb <= 0) continue;
        const u64 key = synth_key(seed, first_read + i, n_keys);
        u64 r = mix64(seed + 0x5EED0002ull + key * 32ull + (u64)blk);
        uint8_t *dst = ascii + o0 + 32 * blk;
        if (nb == 32 && (((uintptr_t)dst) & 15) == 0) {
            u32 w[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                u32 p = (u32)(r >> (8 * k)) & 0xFF;
                u32 x = (p | (p << 4)) & 0x0F0Fu;
                x = (x | (x << 2)) & 0x3333u;
                w[k] = __byte_perm(0x54474341u, 0u, x);   // "ACGT"[code]
            }
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
            for (int j = 0; j < nb; j++) dst[j] = (uint8_t)"ACGT"[(r >> (2 * j)) & 3];
        }
    }
}

__global__ void synth_fixed_offsets_kernel(int64_t n, int64_t len, int64_t *offsets) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        offsets[i] = i * len;
}

}  // namespace ssq

using namespace ssq;

extern "C" {

int ssq_decode64(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *out_offsets,
                 uint8_t *ascii_out) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(words && lens && out_offsets && ascii_out, "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n + kDecTile - 1) / kDecTile, 8);
    decode_fixed_kernel<1><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)words, lens, n, out_offsets, ascii_out);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_decode192(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *out_offsets,
                  uint8_t *ascii_out) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(words && lens && out_offsets && ascii_out, "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n + kDecTile - 1) / kDecTile, 8);
    decode_fixed_kernel<3><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)words, lens, n, out_offsets, ascii_out);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_decode_tiles(ssq_ctx *ctx, const uint8_t *lens, int64_t n, int max_len, int64_t *tile_base) {
    SSQ_ARG(ctx != nullptr && n >= 0 && tile_base != nullptr, "bad ctx / n / tile_base");
    SSQ_ARG(max_len == 32 || max_len == 96, "max_len must be 32 (ShortSeq64) or 96 (ShortSeq192)");
    SSQ_ARG(n == 0 || lens != nullptr, "lens is NULL");
    DeviceGuard g(ctx->device);
    const int64_t ntiles = (n + kDecTile - 1) / kDecTile;
    if (ntiles == 0) { SSQ_CUDA(cudaMemsetAsync(tile_base, 0, sizeof(int64_t), ctx->stream)); return SSQ_OK; }
    int64_t *totals = tile_base + ntiles + 1;          // the caller's buffer holds 2 * ntiles + 2 entries: [bases | totals]
    tile_totals_kernel<<<grid_for(ctx, (ntiles + kThreads / 32 - 1) / (kThreads / 32), 8), kThreads, 0, ctx->stream>>>(lens, n, max_len, totals);
    SSQ_LAUNCH_CHECK();
    return scan_i64(ctx, totals, ntiles, tile_base);
}

static int check_fused(ssq_ctx *ctx, const void *words, const void *lens, int64_t n, const void *tile_base, const void *out_offsets,
                       const void *ascii_out) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    SSQ_ARG(n == 0 || (words && lens && tile_base && out_offsets && ascii_out), "NULL buffer");
    return SSQ_OK;
}

int ssq_decode64_fused(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *tile_base,
                       int64_t *out_offsets, uint8_t *ascii_out) {
    int rc = check_fused(ctx, words, lens, n, tile_base, out_offsets, ascii_out);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    if (n == 0) { SSQ_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t), ctx->stream)); return SSQ_OK; }
    int grid = grid_for(ctx, (n + kDecTile - 1) / kDecTile, 8);
    decode_fixed_fused_kernel<1><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)words, lens, n, tile_base, out_offsets, ascii_out);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_decode192_fused(ssq_ctx *ctx, const uint64_t *words, const uint8_t *lens, int64_t n, const int64_t *tile_base,
                        int64_t *out_offsets, uint8_t *ascii_out) {
    int rc = check_fused(ctx, words, lens, n, tile_base, out_offsets, ascii_out);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    if (n == 0) { SSQ_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t), ctx->stream)); return SSQ_OK; }
    int grid = grid_for(ctx, (n + kDecTile - 1) / kDecTile, 8);
    decode_fixed_fused_kernel<3><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)words, lens, n, tile_base, out_offsets, ascii_out);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_decodevar(ssq_ctx *ctx, const uint64_t *words, const int64_t *word_off, const uint16_t *lens, int64_t n,
                  const int64_t *out_offsets, uint8_t *ascii_out) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(words && word_off && lens && out_offsets && ascii_out, "NULL buffer");
    DeviceGuard g(ctx->device);
    static const bool v1 = getenv("SSQ_VAR_V1") != nullptr;      // development: the first version of the kernel
    int per_sm = 8;
    if (!v1 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_var2_kernel, kVar2Threads, 0) != cudaSuccess || per_sm < 1)) per_sm = 4;
    int grid = grid_for(ctx, (n + kVarTileReads - 1) / kVarTileReads, per_sm);      // the second version is persistent: exactly one wave
    if (v1) decode_var_kernel<<<grid, kThreads, 0, ctx->stream>>>((const u64 *)words, word_off, lens, n, out_offsets, ascii_out);
    else decode_var2_kernel<<<grid, kVar2Threads, 0, ctx->stream>>>((const u64 *)words, word_off, lens, n, out_offsets, ascii_out);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_hamming_pairs64(ssq_ctx *ctx, const uint64_t *a, const uint8_t *len_a, const uint64_t *b, const uint8_t *len_b,
                        int64_t n, uint8_t *dist) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(a && b && len_a && len_b && dist, "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
    hamming64_kernel<<<grid, kThreads, 0, ctx->stream>>>((const u64 *)a, len_a, (const u64 *)b, len_b, n, dist, ctx->d_report);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_hamming_pairs192(ssq_ctx *ctx, const uint64_t *a, const uint8_t *len_a, const uint64_t *b, const uint8_t *len_b,
                         int64_t n, uint8_t *dist) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(a && b && len_a && len_b && dist, "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
    hamming192_kernel<<<grid, kThreads, 0, ctx->stream>>>((const u64 *)a, len_a, (const u64 *)b, len_b, n, dist, ctx->d_report);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_hamming_pairsvar(ssq_ctx *ctx, const uint64_t *a, const int64_t *a_off, const uint16_t *len_a, const uint64_t *b,
                         const int64_t *b_off, const uint16_t *len_b, int64_t n, uint16_t *dist) {
    SSQ_ARG(ctx != nullptr && n >= 0, "bad ctx / n");
    if (n == 0) return SSQ_OK;
    SSQ_ARG(a && b && a_off && b_off && len_a && len_b && dist, "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n * 32 + kThreads - 1) / kThreads, 8);
    hammingvar_kernel<<<grid, kThreads, 0, ctx->stream>>>((const u64 *)a, a_off, len_a, (const u64 *)b, b_off, len_b, n, dist,
                                                          ctx->d_report);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_hamming_refset(ssq_ctx *ctx, int words_per_seq, const uint64_t *q, const uint8_t *len_q, int64_t nq,
                       const uint64_t *refs, const uint8_t *len_r, int32_t nr, int32_t thresh, uint8_t *min_dist,
                       uint32_t *argmin, uint32_t *n_within) {
    SSQ_ARG(ctx != nullptr && nq >= 0 && nr >= 0, "bad ctx / sizes");
    SSQ_ARG(words_per_seq == 1 || words_per_seq == 3, "words_per_seq must be 1 or 3");
    if (nq == 0) return SSQ_OK;
    SSQ_ARG(q && len_q && min_dist && argmin && (nr == 0 || (refs && len_r)), "NULL buffer");
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (nq + kThreads - 1) / kThreads, 4);
    if (words_per_seq == 1)
        refset_kernel<1><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)q, len_q, nq, (const u64 *)refs, len_r, nr, thresh,
                                                             min_dist, argmin, n_within);
    else
        refset_kernel<3><<<grid, kThreads, 0, ctx->stream>>>((const u64 *)q, len_q, nq, (const u64 *)refs, len_r, nr, thresh,
                                                             min_dist, argmin, n_within);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_synth_reads(ssq_ctx *ctx, uint64_t seed, int64_t first_read, int64_t n, int64_t n_keys, int32_t len_lo,
                    int32_t len_hi, int64_t *offsets, uint8_t *ascii) {
    SSQ_ARG(ctx != nullptr && n >= 0 && n_keys > 0, "bad ctx / sizes");
    SSQ_ARG(len_lo >= 0 && len_hi >= len_lo && len_hi <= 1024, "bad length range");
    SSQ_ARG(offsets != nullptr && (ascii != nullptr || n == 0 || len_hi == 0), "NULL buffer");
    DeviceGuard g(ctx->device);
    if (len_lo == len_hi) {
        synth_fixed_offsets_kernel<<<grid_for(ctx, (n + 1 + 255) / 256, 8), 256, 0, ctx->stream>>>(n, len_lo, offsets);
        SSQ_LAUNCH_CHECK();
    } else {
        int rc = scan_synth_lens(ctx, seed, first_read, n, n_keys, len_lo, len_hi, offsets);
        if (rc) return rc;
    }
    if (n == 0 || len_hi == 0) return SSQ_OK;
    const int bpr = (len_hi + 31) / 32;
    int grid = grid_for(ctx, (n * bpr + kThreads - 1) / kThreads, 8);
    synth_fill_kernel<<<grid, kThreads, 0, ctx->stream>>>(seed, first_read, n, n_keys, bpr, offsets, ascii);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

}  // extern "C"
