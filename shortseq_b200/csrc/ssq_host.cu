// ssq_host.cu -- the host-buffer entry points: chunked, double-buffered pack+count.
//
// These are the calls a host-language binding makes with HOST memory (the reference's
// ShortSeqCounter(list_of_bytes), counter.pyx:11-29, after the list has been gathered into one
// buffer).  Chunk k+1 is copied host->device on one copy stream while chunk k runs the fused
// kernel on the compute stream and chunk k-1's packed words travel device->host on the other
// copy stream.  Two input encodings of the read boundaries:
//   ssq_host_pack_count       int64 offsets[n+1]           (8 bytes per read over PCIe)
//   ssq_host_pack_count_lens  uint8 lens[n] (reads <= 96)  (1 byte per read; offsets are scanned on the device)
#include "ssq_internal.h"

using namespace ssq;

namespace {

// Make the context's staging buffers large enough for chunks of `reads` reads / `ascii_bytes` bytes / `words` words.
int ensure_staging(ssq_ctx *ctx, size_t ascii_bytes, size_t reads, size_t word_entries, bool by_lens, int nbuf) {
    ssq_host_staging &st = ctx->staging;
    if (!st.events) {
        for (int b = 0; b < 2; b++) {
            SSQ_CUDA(cudaEventCreateWithFlags(&st.ev_in[b], cudaEventDisableTiming));
            SSQ_CUDA(cudaEventCreateWithFlags(&st.ev_out[b], cudaEventDisableTiming));
        }
        st.events = true;
    }
    const bool grow_ascii = ascii_bytes > st.ascii_bytes, grow_reads = reads > st.reads, grow_words = word_entries > st.word_entries;
    if (grow_ascii || grow_reads || grow_words) SSQ_CUDA(cudaDeviceSynchronize());    // nothing may still be using the old buffers
    for (int b = 0; b < 2; b++) {
        if (grow_ascii) {
            if (st.ascii[b]) SSQ_CUDA(cudaFree(st.ascii[b]));
            st.ascii[b] = nullptr;
            SSQ_CUDA(cudaMalloc(&st.ascii[b], ascii_bytes + 16));
        }
        if (grow_reads) {
            if (st.offsets[b]) SSQ_CUDA(cudaFree(st.offsets[b]));
            if (st.lens_in[b]) SSQ_CUDA(cudaFree(st.lens_in[b]));
            if (st.lens[b]) SSQ_CUDA(cudaFree(st.lens[b]));
            st.offsets[b] = nullptr; st.lens_in[b] = nullptr; st.lens[b] = nullptr;
            SSQ_CUDA(cudaMalloc(&st.offsets[b], sizeof(int64_t) * (reads + 1)));
            SSQ_CUDA(cudaMalloc(&st.lens_in[b], reads));
            SSQ_CUDA(cudaMalloc(&st.lens[b], reads));
        }
        if (grow_words) {
            if (st.words[b]) SSQ_CUDA(cudaFree(st.words[b]));
            st.words[b] = nullptr;
            SSQ_CUDA(cudaMalloc(&st.words[b], sizeof(u64) * word_entries));
        }
    }
    if (grow_ascii) st.ascii_bytes = ascii_bytes;
    if (grow_reads) st.reads = reads;
    if (grow_words) st.word_entries = word_entries;
    (void)by_lens; (void)nbuf;
    return SSQ_OK;
}

int64_t sum_u8(const uint8_t *p, int64_t n) {
    int64_t s = 0;
    for (int64_t i = 0; i < n; i++) s += p[i];
    return s;
}

// h_offsets != NULL: boundaries given as offsets; else h_lens_in (uint8 per read).
int host_pipeline(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii, const int64_t *h_offsets, const uint8_t *h_lens_in,
                  int64_t n, uint64_t *h_words, uint8_t *h_lens, int64_t chunk_reads, ssq_report *report) {
    if (report) { report->code = SSQ_OK; report->reserved = 0; report->first_bad_read = -1; }
    if (n == 0) return ssq_ctx_sync(ctx, report);
    DeviceGuard g(ctx->device);
    if (chunk_reads <= 0) chunk_reads = (int64_t)1 << 22;
    if (chunk_reads > n) chunk_reads = n;
    const int W = c->klass == SSQ_CLASS_64 ? 1 : 3;
    const int maxlen = c->klass == SSQ_CLASS_64 ? 32 : 96;
    const int64_t nchunks = (n + chunk_reads - 1) / chunk_reads;
    const bool by_lens = h_offsets == nullptr;

    int64_t max_bytes = 0;
    if (by_lens) {
        max_bytes = chunk_reads * 255;                       // any uint8 lengths fit; wrong-class reads are reported by the kernel
        if (max_bytes > chunk_reads * (int64_t)maxlen * 2) max_bytes = chunk_reads * (int64_t)maxlen * 2;
    } else {
        // host-side sanity of the offsets bounding each chunk (the kernel checks every read)
        for (int64_t k = 0; k < nchunks; k++) {
            int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
            int64_t b = h_offsets[e] - h_offsets[s];
            SSQ_ARG(b >= 0 && h_offsets[s] >= 0, "offsets must be non-decreasing and non-negative");
            if (b > max_bytes) max_bytes = b;
        }
    }
    SSQ_ARG(max_bytes == 0 || h_ascii != nullptr, "h_ascii is NULL");

    const bool want_words = h_words != nullptr, want_lens = h_lens != nullptr;
    {
        int rc0 = ensure_staging(ctx, (size_t)max_bytes, (size_t)chunk_reads, (size_t)chunk_reads * W, by_lens, nchunks > 1 ? 2 : 1);
        if (rc0) return rc0;
    }
    ssq_host_staging &st = ctx->staging;
    cudaStream_t s_in = ctx->copy_streams[0], s_out = ctx->copy_streams[1], s_run = ctx->stream;

    int64_t byte_pos = 0;          // by_lens: running byte position of the next chunk to copy in
    int64_t chunk_byte0[2] = {0, 0}, chunk_bytes[2] = {0, 0};
    auto copy_in = [&](int64_t k) -> int {
        const int b = (int)(k & 1);
        const int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
        int64_t bytes;
        if (by_lens) {
            bytes = sum_u8(h_lens_in + s, e - s);
            if (bytes > max_bytes) { set_error("chunk %lld holds reads longer than the counter's class allows", (long long)k); return SSQ_ERR_ARG; }
            chunk_byte0[b] = byte_pos;
            SSQ_CUDA(cudaMemcpyAsync(st.lens_in[b], h_lens_in + s, (size_t)(e - s), cudaMemcpyHostToDevice, s_in));
        } else {
            bytes = h_offsets[e] - h_offsets[s];
            chunk_byte0[b] = h_offsets[s];
            SSQ_CUDA(cudaMemcpyAsync(st.offsets[b], h_offsets + s, sizeof(int64_t) * (size_t)(e - s + 1), cudaMemcpyHostToDevice, s_in));
        }
        chunk_bytes[b] = bytes;
        if (bytes) SSQ_CUDA(cudaMemcpyAsync(st.ascii[b], h_ascii + chunk_byte0[b], (size_t)bytes, cudaMemcpyHostToDevice, s_in));
        byte_pos = chunk_byte0[b] + bytes;
        SSQ_CUDA(cudaEventRecord(st.ev_in[b], s_in));
        return SSQ_OK;
    };

    int rc = copy_in(0);
    for (int64_t k = 0; k < nchunks && rc == SSQ_OK; k++) {
        const int b = (int)(k & 1);
        const int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
        const int64_t byte0 = chunk_byte0[b], bytes = chunk_bytes[b];   // read before copy_in(k+1) touches the other slot
        if (k + 1 < nchunks) { rc = copy_in(k + 1); if (rc) break; }   // overlaps with this chunk's kernel
        SSQ_CUDA(cudaStreamWaitEvent(s_run, st.ev_in[b], 0));
        if (k >= 2 && (want_words || want_lens)) SSQ_CUDA(cudaStreamWaitEvent(s_run, st.ev_out[b], 0));
        if (by_lens) {
            // chunk-local offsets from the lengths (exclusive scan on the device); the staged bytes start at 0
            rc = scan_lens_to_offsets(ctx, st.lens_in[b], 1, e - s, st.offsets[b]);
            if (rc) break;
            rc = pack_count_impl(c, st.ascii[b], 0, bytes, st.offsets[b], e - s, s, st.words[b], st.lens[b]);
        } else {
            // offsets are absolute positions in h_ascii: give the kernel a virtual base so that base + byte0 is the
            // first staged byte
            rc = pack_count_impl(c, st.ascii[b] - byte0, byte0, byte0 + bytes, st.offsets[b], e - s, s, st.words[b], st.lens[b]);
        }
        if (rc) break;
        if (want_words || want_lens) {   // the kernel has completed (pack_count_impl synchronises the compute stream)
            if (want_words)
                SSQ_CUDA(cudaMemcpyAsync(h_words + (size_t)s * W, st.words[b], sizeof(u64) * (size_t)(e - s) * W, cudaMemcpyDeviceToHost, s_out));
            if (want_lens) SSQ_CUDA(cudaMemcpyAsync(h_lens + s, st.lens[b], (size_t)(e - s), cudaMemcpyDeviceToHost, s_out));
            SSQ_CUDA(cudaEventRecord(st.ev_out[b], s_out));
        }
    }
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_out);
    if (rc) { cudaStreamSynchronize(s_run); return rc; }
    return ssq_ctx_sync(ctx, report);
}

}  // namespace

extern "C" int ssq_host_pack_count(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii, const int64_t *h_offsets,
                                   int64_t n, uint64_t *h_words, uint8_t *h_lens, int64_t chunk_reads,
                                   ssq_report *report) {
    SSQ_ARG(ctx != nullptr && c != nullptr && c->ctx == ctx, "ctx / counter mismatch");
    SSQ_ARG(n >= 0 && (n == 0 || h_offsets != nullptr), "bad batch");
    SSQ_ARG((h_words == nullptr) == (h_lens == nullptr), "h_words and h_lens must both be given or both be NULL");
    return host_pipeline(ctx, c, h_ascii, h_offsets, nullptr, n, h_words, h_lens, chunk_reads, report);
}

extern "C" int ssq_host_pack_count_lens(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii, const uint8_t *h_lens,
                                        int64_t n, uint64_t *h_words, int64_t chunk_reads, ssq_report *report) {
    SSQ_ARG(ctx != nullptr && c != nullptr && c->ctx == ctx, "ctx / counter mismatch");
    SSQ_ARG(n >= 0 && (n == 0 || h_lens != nullptr), "bad batch");
    return host_pipeline(ctx, c, h_ascii, nullptr, h_lens, n, h_words, nullptr, chunk_reads, report);
}
