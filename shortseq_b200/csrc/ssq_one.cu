// ssq_one.cu -- the per-object calls of the reference API as batches of ONE on the device.
//
// sq.pack(x), str(s) and a ^ b of the reference (short_seq.pyx:13-74, short_seq_64.pyx:77-121 and the 192 / Var
// twins) are sub-microsecond C calls.  The drop-in keeps them on the GPU (no CPU fallback anywhere in this library),
// so a call costs a kernel launch; everything else is taken out of the way here: the context owns one page of
// MAPPED pinned host memory, the host writes the operand into it, a one-CTA kernel reads it over PCIe, writes the
// result back into the same page followed by a sequence number (st.release.sys), and the host polls that word instead
// of paying a stream synchronisation.  No device allocation, no tensor, no copy call: ~10 us per call, against
// ~115 us for the batch-of-one path through pack_batch / decode_batch / hamming_batch.
#include <string.h>
#include "ssq_internal.h"

namespace ssq {

constexpr int kOneThreads = 256;
constexpr int kOneMaxLen = 1024;
constexpr int kOneMaxWords = kOneMaxLen / 32;

// layout of the mapped page
struct OnePage {
    uint8_t ascii[kOneMaxLen];        // pack: in; decode: out
    u64 a[kOneMaxWords];              // pack: out; decode / hamming: in
    u64 b[kOneMaxWords];              // hamming: in
    volatile u32 seq;                 // written last by the kernel: the call's sequence number
    u32 status;                       // pack: 0 ok, 1 = a base outside {A,C,G,T}
    u32 result;                       // hamming: distance
    u32 first_bad;                    // pack: index of the first invalid base
};

__device__ __forceinline__ void publish(OnePage *pg, u32 seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&pg->seq), "r"(seq) : "memory");
    }
}

// pack: thread t encodes bases 4t .. 4t+3 into one byte of 2-bit codes (util.pyx:100-119: code = (c >> 1) & 3)
__global__ void __launch_bounds__(kOneThreads) pack_one_kernel(OnePage *pg, int len, u32 seq) {
    __shared__ __align__(8) uint8_t codes[kOneMaxLen / 4];
    __shared__ u32 s_bad;
    if (threadIdx.x == 0) s_bad = 0xFFFFFFFFu;
    __syncthreads();
    const int t = threadIdx.x;
    u32 code = 0;
    u32 bad_at = 0xFFFFFFFFu;
    if (4 * t < len) {
        const u32 w = reinterpret_cast<const u32 *>(pg->ascii)[t];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (4 * t + k < len) {
                const uint8_t c = (uint8_t)(w >> (8 * k));
                if (!is_acgt(c)) bad_at = min(bad_at, (u32)(4 * t + k));
                code |= (u32)((c >> 1) & 3u) << (2 * k);
            }
        }
    }
    codes[t] = (uint8_t)code;
    if (bad_at != 0xFFFFFFFFu) atomicMin(&s_bad, bad_at);
    __syncthreads();
    if (t < kOneMaxWords) pg->a[t] = reinterpret_cast<const u64 *>(codes)[t];    // bases beyond len encoded as 0: canonical
    if (t == 0) { pg->status = s_bad != 0xFFFFFFFFu; pg->first_bad = s_bad; }
    publish(pg, seq);
}

// decode: thread t writes bases 4t .. 4t+3 ("ACTG"[code], util.pyx:52)
__global__ void __launch_bounds__(kOneThreads) decode_one_kernel(OnePage *pg, int len, u32 seq) {
    const int t = threadIdx.x;
    if (4 * t < len) {
        const u32 byte = (u32)(pg->a[t >> 3] >> (8 * (t & 7))) & 0xFFu;
        u32 out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= ((0x47544341u >> (8 * ((byte >> (2 * k)) & 3u))) & 0xFFu) << (8 * k);
        reinterpret_cast<u32 *>(pg->ascii)[t] = out;
    }
    publish(pg, seq);
}

// Hamming distance of two canonical sequences of equal length: lane w counts block w
__global__ void __launch_bounds__(32) hamming_one_kernel(OnePage *pg, int nwords, u32 seq) {
    const int w = threadIdx.x;
    u32 d = w < nwords ? (u32)diff_bases(pg->a[w], pg->b[w]) : 0u;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, s);
    if (w == 0) pg->result = d;
    publish(pg, seq);
}

static int one_page(ssq_ctx *ctx, OnePage **host, OnePage **dev) {
    if (ctx->one_host == nullptr) {
        void *h = nullptr, *d = nullptr;
        SSQ_CUDA(cudaHostAlloc(&h, sizeof(OnePage), cudaHostAllocMapped));
        memset(h, 0, sizeof(OnePage));
        SSQ_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        ctx->one_host = h;
        ctx->one_dev = d;
        ctx->one_seq = 0;
    }
    *host = (OnePage *)ctx->one_host;
    *dev = (OnePage *)ctx->one_dev;
    return SSQ_OK;
}

// Wait for the kernel's sequence number.  The flag is polled for the first ~200 us (a launch completes in ~5 us);
// after that the stream is synchronised, which also surfaces launch failures.
static int one_wait(ssq_ctx *ctx, OnePage *host, u32 seq) {
    for (int spin = 0; spin < 200000; spin++) {
        if (host->seq == seq) { __sync_synchronize(); return SSQ_OK; }
        __builtin_ia32_pause();
    }
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (host->seq != seq) { set_error("single-object kernel did not complete"); return SSQ_ERR_CUDA; }
    __sync_synchronize();
    return SSQ_OK;
}

}  // namespace ssq

using namespace ssq;

extern "C" {

int ssq_pack_one(ssq_ctx *ctx, const uint8_t *ascii, int32_t len, uint64_t *words, int32_t *klass, int32_t *first_bad) {
    SSQ_ARG(ctx != nullptr && ascii != nullptr && words != nullptr && klass != nullptr, "NULL argument");
    SSQ_ARG(len >= 1 && len <= kOneMaxLen, "length out of range");
    DeviceGuard g(ctx->device);
    OnePage *host, *dev;
    int rc = one_page(ctx, &host, &dev);
    if (rc) return rc;
    memcpy(host->ascii, ascii, (size_t)len);
    const u32 seq = ++ctx->one_seq;
    __sync_synchronize();
    pack_one_kernel<<<1, kOneThreads, 0, ctx->stream>>>(dev, len, seq);
    SSQ_LAUNCH_CHECK();
    rc = one_wait(ctx, host, seq);
    if (rc) return rc;
    const int k = len <= 32 ? SSQ_CLASS_64 : (len <= 96 ? SSQ_CLASS_192 : SSQ_CLASS_VAR);
    const int nw = k == SSQ_CLASS_64 ? 1 : (k == SSQ_CLASS_192 ? 3 : (len + 31) / 32);
    for (int i = 0; i < nw; i++) words[i] = host->a[i];
    *klass = k;
    if (first_bad) *first_bad = host->status ? (int32_t)host->first_bad : -1;
    return host->status ? SSQ_ERR_BAD_BASE : SSQ_OK;
}

int ssq_decode_one(ssq_ctx *ctx, const uint64_t *words, int32_t len, uint8_t *ascii_out) {
    SSQ_ARG(ctx != nullptr && words != nullptr && ascii_out != nullptr, "NULL argument");
    SSQ_ARG(len >= 1 && len <= kOneMaxLen, "length out of range");
    DeviceGuard g(ctx->device);
    OnePage *host, *dev;
    int rc = one_page(ctx, &host, &dev);
    if (rc) return rc;
    const int nw = (len + 31) / 32;
    for (int i = 0; i < nw; i++) host->a[i] = words[i];
    const u32 seq = ++ctx->one_seq;
    __sync_synchronize();
    decode_one_kernel<<<1, kOneThreads, 0, ctx->stream>>>(dev, len, seq);
    SSQ_LAUNCH_CHECK();
    rc = one_wait(ctx, host, seq);
    if (rc) return rc;
    memcpy(ascii_out, host->ascii, (size_t)len);
    return SSQ_OK;
}

int ssq_hamming_one(ssq_ctx *ctx, const uint64_t *a, const uint64_t *b, int32_t len, int32_t *dist) {
    SSQ_ARG(ctx != nullptr && a != nullptr && b != nullptr && dist != nullptr, "NULL argument");
    SSQ_ARG(len >= 1 && len <= kOneMaxLen, "length out of range");
    DeviceGuard g(ctx->device);
    OnePage *host, *dev;
    int rc = one_page(ctx, &host, &dev);
    if (rc) return rc;
    const int nw = (len + 31) / 32;
    for (int i = 0; i < nw; i++) { host->a[i] = a[i]; host->b[i] = b[i]; }
    const u32 seq = ++ctx->one_seq;
    __sync_synchronize();
    hamming_one_kernel<<<1, 32, 0, ctx->stream>>>(dev, nw, seq);
    SSQ_LAUNCH_CHECK();
    rc = one_wait(ctx, host, seq);
    if (rc) return rc;
    *dist = (int32_t)host->result;
    return SSQ_OK;
}

}  // extern "C"
