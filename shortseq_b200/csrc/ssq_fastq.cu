// ssq_fastq.cu -- FASTQ ingest on the GPU: the step immediately before the hot path.
//
// Replaces the reference's read_and_count_fastq (counter.pyx:57-71) and its C getline loop
// (fast_read.pyx:3-20): every line whose 1-based number is 2 mod 4 is a read, and the last byte of
// that line -- the newline, or the last base when the file's final line is unterminated: the
// reference's _from_chars drops it unconditionally (short_seq.pyx:50-52) -- is not part of it.
//
// The file travels to the GPU in chunks.  Per chunk:
//   1. newline_count_kernel   newlines per 4 KB block, then an exclusive scan -> every block's first line number;
//   2. mark_lines_kernel      for every newline: line 4k ends -> read k starts at the next byte, line 4k+1 ends ->
//                             read k ends here; the newline that closes the chunk's last complete record gives the
//                             number of bytes consumed (the rest is re-read with the next chunk);
//   3. per sequence class (ShortSeq64 / ShortSeq192): select the reads of that class (scan of flags), scan their
//      lengths into offsets, gather their bases into one contiguous ASCII buffer (gather_reads_kernel) and hand that
//      to the fused pack+count pass -- the same kernels as ssq_counter_pack_count.
// Errors keep the reference's order: the lowest read number with a bad base wins.
#include <string.h>
#include "ssq_internal.h"

namespace ssq {

constexpr int kFqThreads = 256;
constexpr int kFqBlockBytes = kFqThreads * 16;

// bit i of the result is set iff byte i of the 16 bytes is '\n'
__device__ __forceinline__ u32 newline_bits(uint4 v) {
    u32 w[4] = {v.x, v.y, v.z, v.w};
    u32 m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 x = w[k] ^ 0x0A0A0A0Au;
        const u32 z = (x - 0x01010101u) & ~x & 0x80808080u;      // 0x80 in every byte that was '\n'
        // exact for this use: a borrow can only mis-flag a byte ABOVE a true zero byte when that byte is 0x01 --
        // byte value 0x0B ('\v'); the slow exact path below repairs those rare chunks
        m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    return m;
}

__device__ __forceinline__ u32 chunk_newlines(const uint8_t *text, int64_t nbytes, int64_t at) {
    if (at + 16 <= nbytes) {
        const uint4 v = *reinterpret_cast<const uint4 *>(text + at);
        const u32 m = newline_bits(v);
        // vertical tab (0x0B) right above a newline would be mis-flagged by the borrow: recheck such chunks exactly
        const u32 x0 = v.x ^ 0x0B0B0B0Bu, x1 = v.y ^ 0x0B0B0B0Bu, x2 = v.z ^ 0x0B0B0B0Bu, x3 = v.w ^ 0x0B0B0B0Bu;
        const u32 vt = ((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2) | ((x3 - 0x01010101u) & ~x3);
        if ((vt & 0x80808080u) == 0) return m;
    }
    u32 m = 0;
    for (int b = 0; b < 16; b++)
        if (at + b < nbytes && text[at + b] == '\n') m |= 1u << b;
    return m;
}

__global__ void __launch_bounds__(kFqThreads) newline_count_kernel(const uint8_t *text, int64_t nbytes, u32 *block_counts) {
    __shared__ u32 s_warp[kFqThreads / 32];
    const int64_t at = ((int64_t)blockIdx.x * kFqThreads + threadIdx.x) * 16;
    u32 n = at < nbytes ? (u32)__popc(chunk_newlines(text, nbytes, at)) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int k = 0; k < kFqThreads / 32; k++) tot += s_warp[k];
        block_counts[blockIdx.x] = tot;
    }
}

// starts[k] / ends[k]: byte range of read k (k < nreads); *consumed = byte after the newline with index last_nl.
__global__ void __launch_bounds__(kFqThreads) mark_lines_kernel(const uint8_t *text, int64_t nbytes, const int64_t *block_base,
                                                                int64_t nreads, int64_t last_nl, int64_t *starts, int64_t *ends,
                                                                int64_t *consumed) {
    __shared__ u32 s_warp[kFqThreads / 32];
    const int64_t at = ((int64_t)blockIdx.x * kFqThreads + threadIdx.x) * 16;
    u32 m = at < nbytes ? chunk_newlines(text, nbytes, at) : 0u;
    const u32 n = (u32)__popc(m);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += x; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < warp; w++) before += s_warp[w];
    int64_t line = block_base[blockIdx.x] + before + incl - n;     // number of the line the first newline of this thread ends
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int64_t p = at + b;
        const int64_t k = line >> 2;
        const int which = (int)(line & 3);
        if (k < nreads) {
            if (which == 0) starts[k] = p + 1;
            else if (which == 1) ends[k] = p;
        }
        if (line == last_nl) *consumed = p + 1;
        ++line;
    }
}

// per-class read counts: out[0] = reads of 0..32 nt, out[1] = 33..96, out[2] = 97..1024, out[3] = longer;
// out[4] = lowest read number longer than 96 nt (atomicMin)
__global__ void __launch_bounds__(kFqThreads) class_count_kernel(const int64_t *starts, const int64_t *ends, int64_t nreads, u64 *out) {
    u32 c[4] = {0, 0, 0, 0};
    u64 first_long = kNoIndex;
    for (int64_t k = (int64_t)blockIdx.x * kFqThreads + threadIdx.x; k < nreads; k += (int64_t)gridDim.x * kFqThreads) {
        const int64_t len = ends[k] - starts[k];
        const int cls = len <= 32 ? 0 : (len <= 96 ? 1 : (len <= 1024 ? 2 : 3));
        c[cls]++;
        if (cls >= 2 && (u64)k < first_long) first_long = (u64)k;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c[j] += __shfl_xor_sync(0xFFFFFFFFu, c[j], d);
        if ((threadIdx.x & 31) == 0 && c[j]) atomicAdd(&out[j], (u64)c[j]);
    }
    if (first_long != kNoIndex) atomicMin(&out[4], first_long);
}

__device__ __forceinline__ bool in_class(int64_t len, int klass) { return klass == SSQ_CLASS_64 ? (len >= 0 && len <= 32) : (len >= 33 && len <= 96); }

// sel[pos[k]] = k for the reads of the class (pos = exclusive scan of the class flags)
__global__ void __launch_bounds__(kFqThreads) select_kernel(const int64_t *starts, const int64_t *ends, const int64_t *pos, int64_t nreads,
                                                            int klass, int64_t *sel) {
    for (int64_t k = (int64_t)blockIdx.x * kFqThreads + threadIdx.x; k < nreads; k += (int64_t)gridDim.x * kFqThreads)
        if (in_class(ends[k] - starts[k], klass)) sel[pos[k]] = k;
}

// One warp per selected read: copy its bases to ascii[offsets[j] ...).  sel == nullptr: read j is record j.
__global__ void __launch_bounds__(kFqThreads) gather_reads_kernel(const uint8_t *text, const int64_t *starts, const int64_t *sel,
                                                                  const int64_t *offsets, int64_t n, uint8_t *ascii) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * kFqThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kFqThreads) >> 5;
    for (int64_t j = warp0; j < n; j += nwarps) {
        const int64_t k = sel ? sel[j] : j;
        const int64_t s = starts[k], o = offsets[j];
        const int len = (int)(offsets[j + 1] - o);
        for (int b = lane; b < len; b += 32) ascii[o + b] = text[s + b];
    }
}

// exclusive scans over per-read values (ssq_scan.cu)
int scan_fastq_flags(ssq_ctx *ctx, const int64_t *starts, const int64_t *ends, int klass, int64_t n, int64_t *out);
int scan_fastq_lens(ssq_ctx *ctx, const int64_t *starts, const int64_t *ends, const int64_t *sel, int64_t n, int64_t *out);

}  // namespace ssq

using namespace ssq;

namespace {

struct Carve {
    uint8_t *p;
    template <class T> T *take(size_t count) {
        T *r = reinterpret_cast<T *>(p);
        p += (count * sizeof(T) + 255) & ~(size_t)255;
        return r;
    }
};

}  // namespace

// Reads per container class of a batch given as offsets (ssq_classify; short_seq.pyx:54-74 decides the class
// from the length).  counts (device, 5 entries): reads of 0..32, 33..96, 97..1024 and > 1024 nt, then the lowest read
// index longer than 96 nt (-1 if none).
extern "C" int ssq_classify(ssq_ctx *ctx, const int64_t *offsets, int64_t n, int64_t *counts) {
    SSQ_ARG(ctx != nullptr && counts != nullptr && n >= 0, "bad arguments");
    SSQ_ARG(n == 0 || offsets != nullptr, "offsets is NULL");
    DeviceGuard g(ctx->device);
    SSQ_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int64_t), ctx->stream));
    SSQ_CUDA(cudaMemsetAsync(counts + 4, 0xFF, sizeof(int64_t), ctx->stream));
    if (n == 0) return SSQ_OK;
    class_count_kernel<<<grid_for(ctx, (n + kFqThreads - 1) / kFqThreads, 8), kFqThreads, 0, ctx->stream>>>(offsets, offsets + 1, n, (u64 *)counts);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

extern "C" int ssq_host_fastq_count(ssq_ctx *ctx, ssq_counter *c64, ssq_counter *c192, const uint8_t *h_text, int64_t nbytes,
                                    int64_t chunk_bytes, int track_first_index, int64_t *n_reads, int64_t *n_longer,
                                    int64_t *first_longer, ssq_report *report) {
    SSQ_ARG(ctx != nullptr && n_reads != nullptr && n_longer != nullptr && first_longer != nullptr, "NULL argument");
    SSQ_ARG(nbytes >= 0 && (nbytes == 0 || h_text != nullptr), "bad text buffer");
    SSQ_ARG(c64 == nullptr || (c64->ctx == ctx && c64->klass == SSQ_CLASS_64), "c64 must be a ShortSeq64 counter of this context");
    SSQ_ARG(c192 == nullptr || (c192->ctx == ctx && c192->klass == SSQ_CLASS_192), "c192 must be a ShortSeq192 counter of this context");
    *n_reads = 0; *n_longer = 0; *first_longer = -1;
    if (report) { report->code = SSQ_OK; report->reserved = 0; report->first_bad_read = -1; }
    if (nbytes == 0) return SSQ_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (chunk_bytes <= 0) chunk_bytes = (int64_t)256 << 20;
    if (chunk_bytes > nbytes) chunk_bytes = nbytes;
    if (chunk_bytes < 4096) chunk_bytes = nbytes < 4096 ? nbytes : 4096;

    const int64_t C = chunk_bytes;
    const int64_t max_reads = C / 8 + 1024;                        // records of a chunk handled at once (real FASTQ: >= ~40 bytes each)
    const int64_t nblocks_max = (C + kFqBlockBytes - 1) / kFqBlockBytes;
    // one scratch block, carved
    size_t need = 0;
    auto add = [&](size_t bytes) { need += (bytes + 255) & ~(size_t)255; };
    add(C + 16); add(C + 16); add(C + 16); add(4 * nblocks_max); add(8 * (nblocks_max + 1)); add(8 * max_reads); add(8 * max_reads);
    add(8 * (max_reads + 1)); add(8 * max_reads); add(8 * (max_reads + 1)); add(24 * max_reads); add(max_reads); add(64);
    void *base = nullptr;
    SSQ_CUDA(cudaMalloc(&base, need));
    struct Free { void *p; ~Free() { cudaFree(p); } } freer{base};
    Carve cv{(uint8_t *)base};
    uint8_t *d_text_buf[2] = {cv.take<uint8_t>(C + 16), cv.take<uint8_t>(C + 16)};   // chunk k+1 is copied in while chunk k is processed
    uint8_t *d_ascii = cv.take<uint8_t>(C + 16);
    u32 *d_bcount = cv.take<u32>(nblocks_max);
    int64_t *d_bbase = cv.take<int64_t>(nblocks_max + 1);
    int64_t *d_starts = cv.take<int64_t>(max_reads);
    int64_t *d_ends = cv.take<int64_t>(max_reads);
    int64_t *d_pos = cv.take<int64_t>(max_reads + 1);
    int64_t *d_sel = cv.take<int64_t>(max_reads);
    int64_t *d_offsets = cv.take<int64_t>(max_reads + 1);
    u64 *d_words = cv.take<u64>(3 * max_reads);
    uint8_t *d_lens = cv.take<uint8_t>(max_reads);
    u64 *d_small = cv.take<u64>(8);                                 // [0..3] class counts, [4] first long read, [5] consumed
    int64_t h_vals[8];

    cudaStream_t s_copy = ctx->copy_streams[0];
    cudaEvent_t ev_in[2];
    for (int b = 0; b < 2; b++) SSQ_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
    struct EvFree { cudaEvent_t *e; ~EvFree() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } evfree{ev_in};
    auto copy_in = [&](int b, int64_t at) -> int {
        const int64_t l = nbytes - at < C ? nbytes - at : C;
        SSQ_CUDA(cudaMemcpyAsync(d_text_buf[b], h_text + at, (size_t)l, cudaMemcpyHostToDevice, s_copy));
        SSQ_CUDA(cudaEventRecord(ev_in[b], s_copy));
        return SSQ_OK;
    };
    int64_t pos = 0, reads_before = 0;
    int64_t best_bad = -1; int best_code = SSQ_OK;
    int buf = 0;
    {
        int rc0 = copy_in(0, 0);
        if (rc0) return rc0;
    }
    while (pos < nbytes) {
        const int64_t len = nbytes - pos < C ? nbytes - pos : C;
        const bool last = pos + len == nbytes;
        uint8_t *d_text = d_text_buf[buf];
        SSQ_CUDA(cudaStreamWaitEvent(st, ev_in[buf], 0));
        const int64_t nb = (len + kFqBlockBytes - 1) / kFqBlockBytes;
        newline_count_kernel<<<(unsigned)nb, kFqThreads, 0, st>>>(d_text, len, d_bcount);
        SSQ_LAUNCH_CHECK();
        int rc = scan_u32_counts(ctx, d_bcount, nb, d_bbase);
        if (rc) return rc;
        int64_t total_nl = 0;
        SSQ_CUDA(cudaMemcpyAsync(&total_nl, d_bbase + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SSQ_CUDA(cudaStreamSynchronize(st));
        const bool unterminated = last && h_text[nbytes - 1] != '\n';
        int64_t nreads, last_nl = -1;
        bool whole = last;                                         // the chunk is consumed to its end
        if (last) {
            const int64_t total_lines = total_nl + (unterminated ? 1 : 0);
            nreads = (total_lines + 2) / 4;
        } else {
            nreads = total_nl / 4;
        }
        if (nreads > max_reads) { nreads = max_reads; whole = false; }
        if (!whole) {
            if (nreads == 0) { set_error("FASTQ record larger than the chunk size (%lld bytes)", (long long)C); return SSQ_ERR_ARG; }
            last_nl = 4 * nreads - 1;
        }
        SSQ_CUDA(cudaMemsetAsync(d_small, 0, 8 * sizeof(u64), st));
        SSQ_CUDA(cudaMemsetAsync(d_small + 4, 0xFF, sizeof(u64), st));
        mark_lines_kernel<<<(unsigned)nb, kFqThreads, 0, st>>>(d_text, len, d_bbase, nreads, last_nl, d_starts, d_ends, (int64_t *)(d_small + 5));
        SSQ_LAUNCH_CHECK();
        if (nreads > 0 && unterminated && (total_nl & 3) == 1 && (total_nl >> 2) < nreads) {
            const int64_t e = len - 1;                              // the unterminated last line is a read: its last byte is dropped
            SSQ_CUDA(cudaMemcpyAsync(d_ends + (total_nl >> 2), &e, sizeof(int64_t), cudaMemcpyHostToDevice, st));
            SSQ_CUDA(cudaStreamSynchronize(st));                    // `e` is a stack variable
        }
        if (nreads > 0) {
            class_count_kernel<<<grid_for(ctx, (nreads + kFqThreads - 1) / kFqThreads, 8), kFqThreads, 0, st>>>(d_starts, d_ends, nreads, d_small);
            SSQ_LAUNCH_CHECK();
        }
        SSQ_CUDA(cudaMemcpyAsync(h_vals, d_small, 8 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        SSQ_CUDA(cudaStreamSynchronize(st));
        const int64_t n_by_class[2] = {h_vals[0], h_vals[1]};
        const int64_t n_long = h_vals[2] + h_vals[3];
        if (n_long > 0) {
            if (*first_longer < 0) *first_longer = reads_before + h_vals[4];
            *n_longer += n_long;
        }
        const int64_t consumed = whole ? len : h_vals[5];
        if (consumed > 0 && pos + consumed < nbytes) {             // the next chunk travels while this one is counted
            rc = copy_in(buf ^ 1, pos + consumed);
            if (rc) return rc;
        }
        // ---- per class: select, offsets, gather, fused pack + count
        for (int k = 0; k < 2 && nreads > 0; k++) {
            ssq_counter *c = k == 0 ? c64 : c192;
            const int klass = k == 0 ? SSQ_CLASS_64 : SSQ_CLASS_192;
            const int64_t n = n_by_class[k];
            if (n == 0) continue;
            if (c == nullptr) {                                     // no counter for this class: report like a class error
                if (best_bad < 0) { best_code = SSQ_ERR_CLASS; best_bad = reads_before; }
                continue;
            }
            const int64_t *sel = nullptr;
            if (n != nreads) {                                      // mixed classes: compact the read numbers of this class
                rc = scan_fastq_flags(ctx, d_starts, d_ends, klass, nreads, d_pos);
                if (rc) return rc;
                select_kernel<<<grid_for(ctx, (nreads + kFqThreads - 1) / kFqThreads, 8), kFqThreads, 0, st>>>(d_starts, d_ends, d_pos, nreads, klass, d_sel);
                SSQ_LAUNCH_CHECK();
                sel = d_sel;
            }
            rc = scan_fastq_lens(ctx, d_starts, d_ends, sel, n, d_offsets);
            if (rc) return rc;
            int64_t total_bases = 0;
            SSQ_CUDA(cudaMemcpyAsync(&total_bases, d_offsets + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            gather_reads_kernel<<<grid_for(ctx, (n + kFqThreads / 32 - 1) / (kFqThreads / 32), 8), kFqThreads, 0, st>>>(d_text, d_starts, sel, d_offsets, n, d_ascii);
            SSQ_LAUNCH_CHECK();
            SSQ_CUDA(cudaStreamSynchronize(st));
            rc = pack_count_impl(c, d_ascii, 0, total_bases, d_offsets, n, 0, d_words, d_lens);
            if (rc) return rc;
            if (track_first_index) {
                rc = counter_first_index_indexed(c, d_words, d_lens, n, sel, reads_before);
                if (rc) return rc;
            }
            ssq_report rep;
            rc = ssq_ctx_sync(ctx, &rep);
            if (rc) return rc;
            if (rep.code != SSQ_OK) {
                int64_t rec = rep.first_bad_read;
                if (rec >= 0 && sel != nullptr) {
                    SSQ_CUDA(cudaMemcpy(&rec, d_sel + rep.first_bad_read, sizeof(int64_t), cudaMemcpyDeviceToHost));
                }
                const int64_t global = rec >= 0 ? reads_before + rec : -1;
                if (rep.code == SSQ_ERR_TABLE_FULL) { if (report) { report->code = rep.code; report->first_bad_read = -1; } return SSQ_OK; }
                if (best_bad < 0 || (global >= 0 && global < best_bad)) { best_bad = global; best_code = rep.code; }
            }
        }
        reads_before += nreads;
        pos += consumed;
        buf ^= 1;
        if (best_bad >= 0) { cudaStreamSynchronize(s_copy); break; }   // the reference stops at the first bad read
        if (consumed <= 0) { set_error("FASTQ ingest made no progress"); return SSQ_ERR_ARG; }
    }
    *n_reads = reads_before;
    if (report && best_bad >= 0) { report->code = best_code; report->first_bad_read = best_bad; }
    return SSQ_OK;
}
