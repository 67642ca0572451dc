// ssq_ctx.cu -- contexts, error plumbing and memory helpers of the C ABI.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "ssq_internal.h"

namespace ssq {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear the sticky-less error state
    return SSQ_ERR_CUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool debug_sync() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SSQ_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int ctx_scratch(ssq_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->scratch_bytes) {
        SSQ_CUDA(cudaDeviceSynchronize());
        if (ctx->scratch) SSQ_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes < ((size_t)1 << 20) ? ((size_t)1 << 20) : bytes + bytes / 2;
        SSQ_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return SSQ_OK;
}

__global__ void reset_report_kernel(DevReport *r) {
    r->first_bad_base = kNoIndex;
    r->first_bad_len = kNoIndex;
    r->first_too_long = kNoIndex;
    r->first_len_mismatch = kNoIndex;
    r->table_overflow = 0;
}

}  // namespace ssq

using namespace ssq;

extern "C" {

int ssq_abi_version(void) { return SSQ_ABI_VERSION; }

const char *ssq_last_error(void) { return g_err; }

uint64_t ssq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ssq_device_count(int *count) {
    SSQ_ARG(count != nullptr, "count is NULL");
    *count = 0;
    SSQ_CUDA(cudaGetDeviceCount(count));
    return SSQ_OK;
}

int ssq_ctx_create(int device, ssq_ctx **out) {
    SSQ_ARG(out != nullptr, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    SSQ_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) { set_error("no CUDA device (there is no CPU fallback)"); return SSQ_ERR_CUDA; }
    SSQ_ARG(device >= 0 && device < ndev, "device index out of range");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    SSQ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return SSQ_ERR_CUDA;
    }
    ssq_ctx *ctx = new ssq_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    SSQ_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 2; i++) SSQ_CUDA(cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking));
    SSQ_CUDA(cudaMalloc(&ctx->d_report, sizeof(DevReport)));
    SSQ_CUDA(cudaHostAlloc(&ctx->h_report, sizeof(DevReport), cudaHostAllocDefault));
    reset_report_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_report);
    SSQ_LAUNCH_CHECK();
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return SSQ_OK;
}

int ssq_ctx_destroy(ssq_ctx *ctx) {
    if (!ctx) return SSQ_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < 2; b++) {
        ssq_host_staging &st = ctx->staging;
        cudaFree(st.ascii[b]); cudaFree(st.offsets[b]); cudaFree(st.lens_in[b]); cudaFree(st.words[b]); cudaFree(st.lens[b]);
        if (st.events) { cudaEventDestroy(st.ev_in[b]); cudaEventDestroy(st.ev_out[b]); }
    }
    cudaFree(ctx->scratch);
    cudaFree(ctx->d_report);
    cudaFreeHost(ctx->h_report);
    if (ctx->one_host) cudaFreeHost(ctx->one_host);
    for (int i = 0; i < 2; i++) cudaStreamDestroy(ctx->copy_streams[i]);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return SSQ_OK;
}

int ssq_ctx_set_stream(ssq_ctx *ctx, void *cuda_stream) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    ctx->stream = (cudaStream_t)cuda_stream;
    return SSQ_OK;
}

int ssq_ctx_reset_stream(ssq_ctx *ctx) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    ctx->stream = ctx->own_stream;
    return SSQ_OK;
}

void *ssq_ctx_stream(ssq_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int ssq_ctx_sync(ssq_ctx *ctx, ssq_report *report) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (report) { report->code = SSQ_OK; report->reserved = 0; report->first_bad_read = -1; }
    SSQ_CUDA(cudaMemcpyAsync(ctx->h_report, ctx->d_report, sizeof(DevReport), cudaMemcpyDeviceToHost, ctx->stream));
    reset_report_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_report);
    SSQ_LAUNCH_CHECK();
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    const DevReport &r = *ctx->h_report;
    // lowest failing read wins, like the reference's serial loop (counter.pyx:23-29)
    u64 best = kNoIndex;
    int code = SSQ_OK;
    if (r.first_too_long < best) { best = r.first_too_long; code = SSQ_ERR_TOO_LONG; }
    if (r.first_bad_len < best) { best = r.first_bad_len; code = SSQ_ERR_CLASS; }
    if (r.first_bad_base < best) { best = r.first_bad_base; code = SSQ_ERR_BAD_BASE; }
    if (r.first_len_mismatch < best) { best = r.first_len_mismatch; code = SSQ_ERR_LEN_MISMATCH; }
    if (code == SSQ_OK && r.table_overflow != 0) { code = SSQ_ERR_TABLE_FULL; best = kNoIndex; }
    if (r.exchange_timeout != 0) {          // a peer's arrival flag never showed up: the merged counts are incomplete
        set_error("multi-GPU merge: %llu arrival flag(s) of other ranks did not show up in time", (unsigned long long)r.exchange_timeout);
        code = SSQ_ERR_EXCHANGE;
        best = kNoIndex;
    }
    if (report) {
        report->code = code;
        report->first_bad_read = best == kNoIndex ? -1 : (int64_t)best;
    }
    return SSQ_OK;
}

int ssq_malloc(ssq_ctx *ctx, size_t bytes, void **dptr) {
    SSQ_ARG(ctx != nullptr && dptr != nullptr, "NULL argument");
    DeviceGuard g(ctx->device);
    *dptr = nullptr;
    SSQ_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return SSQ_OK;
}

int ssq_free(ssq_ctx *ctx, void *dptr) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    SSQ_CUDA(cudaFree(dptr));
    return SSQ_OK;
}

int ssq_host_alloc(size_t bytes, void **hptr) {
    SSQ_ARG(hptr != nullptr, "hptr is NULL");
    *hptr = nullptr;
    SSQ_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SSQ_OK;
}

int ssq_host_free(void *hptr) {
    SSQ_CUDA(cudaFreeHost(hptr));
    return SSQ_OK;
}

int ssq_memcpy_h2d(ssq_ctx *ctx, void *dst, const void *src, size_t bytes) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (bytes) SSQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SSQ_OK;
}

int ssq_memcpy_d2h(ssq_ctx *ctx, void *dst, const void *src, size_t bytes) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (bytes) SSQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SSQ_OK;
}

int ssq_memset(ssq_ctx *ctx, void *dst, int value, size_t bytes) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (bytes) SSQ_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return SSQ_OK;
}

// CUDA IPC: share a cudaMalloc'ed buffer with the other single-GPU processes of the box (multi-GPU exchange over
// NVLink peer memory).  handle = the 64 bytes of a cudaIpcMemHandle_t.
int ssq_ipc_get_handle(ssq_ctx *ctx, void *dptr, void *handle64) {
    SSQ_ARG(ctx != nullptr && dptr != nullptr && handle64 != nullptr, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    SSQ_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, sizeof(h));
    return SSQ_OK;
}

int ssq_ipc_open(ssq_ctx *ctx, const void *handle64, void **dptr) {
    SSQ_ARG(ctx != nullptr && dptr != nullptr && handle64 != nullptr, "NULL argument");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *dptr = nullptr;
    SSQ_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SSQ_OK;
}

int ssq_ipc_close(ssq_ctx *ctx, void *dptr) {
    SSQ_ARG(ctx != nullptr, "ctx is NULL");
    if (!dptr) return SSQ_OK;
    DeviceGuard g(ctx->device);
    SSQ_CUDA(cudaIpcCloseMemHandle(dptr));
    return SSQ_OK;
}

}  // extern "C"
