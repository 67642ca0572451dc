// ssq_host.cu -- the host-buffer entry point: chunked, double-buffered pack+count.
//
// This is the call a host-language binding makes with HOST memory (the reference's
// ShortSeqCounter(list_of_bytes), counter.pyx:11-29, after the list has been gathered
// into one buffer + offsets).  Chunk k+1 is copied host->device on one copy stream while
// chunk k runs the fused kernel on the compute stream and chunk k-1's packed words/lens
// travel device->host on the other copy stream.
#include "ssq_internal.h"

using namespace ssq;

namespace {

struct Staging {
    uint8_t *ascii[2] = {nullptr, nullptr};
    int64_t *offsets[2] = {nullptr, nullptr};
    u64 *words[2] = {nullptr, nullptr};
    uint8_t *lens[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    ~Staging() {
        for (int b = 0; b < 2; b++) {
            cudaFree(ascii[b]); cudaFree(offsets[b]); cudaFree(words[b]); cudaFree(lens[b]);
            if (ev_in[b]) cudaEventDestroy(ev_in[b]);
            if (ev_out[b]) cudaEventDestroy(ev_out[b]);
        }
    }
};

}  // namespace

extern "C" int ssq_host_pack_count(ssq_ctx *ctx, ssq_counter *c, const uint8_t *h_ascii, const int64_t *h_offsets,
                                   int64_t n, uint64_t *h_words, uint8_t *h_lens, int64_t chunk_reads,
                                   ssq_report *report) {
    SSQ_ARG(ctx != nullptr && c != nullptr && c->ctx == ctx, "ctx / counter mismatch");
    SSQ_ARG(n >= 0 && (n == 0 || h_offsets != nullptr), "bad batch");
    SSQ_ARG((h_words == nullptr) == (h_lens == nullptr), "h_words and h_lens must both be given or both be NULL");
    if (report) { report->code = SSQ_OK; report->reserved = 0; report->first_bad_read = -1; }
    if (n == 0) return ssq_ctx_sync(ctx, report);
    DeviceGuard g(ctx->device);
    if (chunk_reads <= 0) chunk_reads = (int64_t)1 << 22;
    if (chunk_reads > n) chunk_reads = n;
    const int W = c->klass == SSQ_CLASS_64 ? 1 : 3;
    const int64_t nchunks = (n + chunk_reads - 1) / chunk_reads;

    // host-side sanity of the offsets bounding each chunk (the kernel checks every read)
    int64_t max_bytes = 0;
    for (int64_t k = 0; k < nchunks; k++) {
        int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
        int64_t b = h_offsets[e] - h_offsets[s];
        SSQ_ARG(b >= 0 && h_offsets[s] >= 0, "offsets must be non-decreasing and non-negative");
        if (b > max_bytes) max_bytes = b;
    }
    SSQ_ARG(max_bytes == 0 || h_ascii != nullptr, "h_ascii is NULL");

    Staging st;
    const bool want_out = h_words != nullptr;
    for (int b = 0; b < (nchunks > 1 ? 2 : 1); b++) {
        SSQ_CUDA(cudaMalloc(&st.ascii[b], (size_t)max_bytes + 16));
        SSQ_CUDA(cudaMalloc(&st.offsets[b], sizeof(int64_t) * (size_t)(chunk_reads + 1)));
        SSQ_CUDA(cudaMalloc(&st.words[b], sizeof(u64) * (size_t)chunk_reads * W));
        SSQ_CUDA(cudaMalloc(&st.lens[b], (size_t)chunk_reads));
        SSQ_CUDA(cudaEventCreateWithFlags(&st.ev_in[b], cudaEventDisableTiming));
        SSQ_CUDA(cudaEventCreateWithFlags(&st.ev_out[b], cudaEventDisableTiming));
    }
    cudaStream_t s_in = ctx->copy_streams[0], s_out = ctx->copy_streams[1], s_run = ctx->stream;

    auto copy_in = [&](int64_t k) -> int {
        const int b = (int)(k & 1);
        const int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
        const int64_t bytes = h_offsets[e] - h_offsets[s];
        if (bytes) SSQ_CUDA(cudaMemcpyAsync(st.ascii[b], h_ascii + h_offsets[s], (size_t)bytes, cudaMemcpyHostToDevice, s_in));
        SSQ_CUDA(cudaMemcpyAsync(st.offsets[b], h_offsets + s, sizeof(int64_t) * (size_t)(e - s + 1), cudaMemcpyHostToDevice, s_in));
        SSQ_CUDA(cudaEventRecord(st.ev_in[b], s_in));
        return SSQ_OK;
    };

    int rc = copy_in(0);
    for (int64_t k = 0; k < nchunks && rc == SSQ_OK; k++) {
        const int b = (int)(k & 1);
        const int64_t s = k * chunk_reads, e = s + chunk_reads < n ? s + chunk_reads : n;
        if (k + 1 < nchunks) { rc = copy_in(k + 1); if (rc) break; }   // overlaps with this chunk's kernel
        SSQ_CUDA(cudaStreamWaitEvent(s_run, st.ev_in[b], 0));
        if (k >= 2 && want_out) SSQ_CUDA(cudaStreamWaitEvent(s_run, st.ev_out[b], 0));
        // offsets are absolute positions in h_ascii: give the kernel a virtual base so that
        // base + h_offsets[s] is the first staged byte
        const uint8_t *vbase = st.ascii[b] - h_offsets[s];
        rc = pack_count_impl(c, vbase, h_offsets[s], h_offsets[e], st.offsets[b], e - s, s, st.words[b], st.lens[b]);
        if (rc) break;
        if (want_out) {   // the kernel has completed (pack_count_impl synchronises the compute stream)
            SSQ_CUDA(cudaMemcpyAsync(h_words + (size_t)s * W, st.words[b], sizeof(u64) * (size_t)(e - s) * W, cudaMemcpyDeviceToHost, s_out));
            SSQ_CUDA(cudaMemcpyAsync(h_lens + s, st.lens[b], (size_t)(e - s), cudaMemcpyDeviceToHost, s_out));
            SSQ_CUDA(cudaEventRecord(st.ev_out[b], s_out));
        }
    }
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_out);
    if (rc) { cudaStreamSynchronize(s_run); return rc; }
    return ssq_ctx_sync(ctx, report);
}
