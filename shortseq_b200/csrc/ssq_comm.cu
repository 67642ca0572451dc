// ssq_comm.cu -- the multi-GPU merge of the dedup counter behind the C ABI.
//
// The reference has no multi-process mode; this is the "one exchange step" of the sharded path:
// every rank (one process per GPU) holds a LOCAL counter of its shard of the reads; the global counter is the
// disjoint union of per-rank OWNER tables, owner = top log2(P) bits of the key hash.  ssq_counter_merge_alltoall moves
// the distinct keys -- not the reads -- to their owners and adds them up there:
//
//   1. sizes: tuples per owner from the table's region occupancy (no pass over the table), ncclAllGather of P + 1
//      numbers per rank, one short host read-back of the P x (P + 1) matrix (it sizes the receive buffers);
//   2. ShortSeq64 (peer path): the export kernel compacts the table region by region and STORES every owner's share
//      straight into that owner's receive buffer over NVLink (buffers shared with CUDA IPC) -- compute step and exchange
//      are one kernel, no collective moves the payload; a one-thread kernel behind it publishes an ARRIVAL FLAG
//      (st.release.sys) in every owner's memory;
//   3. the owner's merge kernel waits for the P flags on the device (ld.acquire.sys, bounded spin) and counts the
//      received, region-ordered tuples region by region in shared memory.  No cudaDeviceSynchronize, no host barrier:
//      the next merge's size all-gather orders the reuse of the buffers (a rank joins it only after its own merge).
//   ShortSeq192 counters, and boxes without CUDA IPC, take the staged path: export into a local buffer grouped by
//   owner, grouped ncclSend / ncclRecv (all-to-all-v), weighted insert on the owner.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch already loaded, or the system's), so the
// library has no link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ssq_internal.h"
#include "ssq_table.cuh"

namespace ssq {

int counter_merge_regions_impl(ssq_counter *c, const u64 *words, const uint8_t *lens, const u64 *counts, const int64_t *block_counts,
                               const int64_t *block_regions, int n_blocks, const int64_t *region_bases, int64_t rb_stride,
                               const u64 *flags, u64 epoch);
int counter_export_count_pass(ssq_counter *c, int n_parts, int64_t *part_counts);
int counter_export_scatter_to(ssq_counter *c, int n_parts, u64 *const *dst_words, uint8_t *const *dst_lens, u64 *const *dst_counts);
bool counter_stream_nests(const ssq_counter *owner, int n_blocks, int64_t block_regions, int log2_sregion);
int counter_merge_streamed_impl(ssq_counter *c, const void *entries, const uint32_t *cnts, int n_blocks, int64_t block_regions,
                                int log2_sregion, const u64 *flags, u64 epoch);

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclAllGather) AllGather;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclGetErrorString) GetErrorString;
    bool ok;
};

static NcclApi *nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        memset(&api, 0, sizeof(api));
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
#define SSQ_SYM(field, name) api.field = (decltype(api.field))dlsym(h, name)
            SSQ_SYM(GetUniqueId, "ncclGetUniqueId");
            SSQ_SYM(CommInitRank, "ncclCommInitRank");
            SSQ_SYM(CommDestroy, "ncclCommDestroy");
            SSQ_SYM(AllGather, "ncclAllGather");
            SSQ_SYM(Send, "ncclSend");
            SSQ_SYM(Recv, "ncclRecv");
            SSQ_SYM(GroupStart, "ncclGroupStart");
            SSQ_SYM(GroupEnd, "ncclGroupEnd");
            SSQ_SYM(GetErrorString, "ncclGetErrorString");
#undef SSQ_SYM
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv &&
                     api.GroupStart && api.GroupEnd && api.GetErrorString;
        }
    }
    return &api;
}

#define SSQ_NCCL(call)                                                                                    \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            set_error("NCCL error %d (%s) in %s at %s:%d", (int)r__, nccl()->GetErrorString(r__), #call, __FILE__, __LINE__); \
            return SSQ_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

constexpr int kMaxRanks = 32;
// kBufXE / kBufXC (+ parity): fixed-place receive buffers of the streamed exchange (ssq_comm_attach)
enum { kBufWords = 0, kBufLens = 1, kBufCounts = 2, kBufBases = 3, kBufFlags = 4, kBufXE = 5, kBufXC = 7, kNumBufs = 9 };

}  // namespace ssq

struct ssq_comm {
    ssq_ctx *ctx;
    ncclComm_t comm;
    int rank, world;
    bool peer_ok;                         // every rank could map every other rank's buffers (CUDA IPC)
    void *mine[ssq::kNumBufs];            // this rank's receive buffers (words / lens / counts / region bases / arrival flags)
    void *peer[ssq::kNumBufs][ssq::kMaxRanks];   // the same buffers of every rank as this process sees them
    int64_t cap;                          // tuples the receive buffers hold
    int64_t rb_cap;                       // sender regions per owner the region-base buffer holds
    int64_t words_per_tuple;              // 1 or 3: what the buffers were sized for
    uint64_t epoch;                       // merges done; arrival flags carry it
    int64_t *d_mine, *d_matrix, *h_matrix;   // [world + 1] / [world][world + 1] size exchange (device, device, pinned)
    int64_t *d_table;                     // [4][world] destination pointers of the export kernels (device)
    int64_t *h_table;                     // pinned staging of the same
    void *stage[3];                       // staged path: local export buffers (words / lens / counts)
    int64_t stage_cap;
    cudaEvent_t ev[3];
    // streamed exchange: a local counter attached with ssq_comm_attach sends its regions from inside the count kernel
    ssq_stream_out so;
    ssq_counter *attached;                // cleared by ssq_counter_destroy of that counter
    bool x_ready;                         // the kBufXE / kBufXC buffers exist (collective state: the same on every rank)
    int64_t x_block_regions;              // sender regions per (sender, owner) block
    int x_log2_sregion;                   // log2(slots per sender region)
    int64_t *d_xtable;                    // [5][world]: entry / count block pointers per parity, flag pointers
    bool last_streamed;                   // the last merge took the streamed path
};

namespace ssq {

// epoch flags: flags[src] in the owner's memory; written by rank src's signal kernel after its export kernel
__global__ void signal_kernel(u64 *const *flag_ptrs, int world, int rank, u64 epoch) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        for (int d = 0; d < world; d++) {
            u64 *f = flag_ptrs[d] + rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(epoch) : "memory");
        }
    }
}

static int comm_barrier(ssq_comm *cm) {
    // a one-number all-gather on the stream, then a stream sync: every rank has got here
    SSQ_NCCL(nccl()->AllGather(cm->d_mine, cm->d_matrix, 1, ncclInt64, cm->comm, cm->ctx->stream));
    SSQ_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    return SSQ_OK;
}

// Collective: unmap the peers' buffers [first, last), barrier, free this rank's (CUDA requires importers to close before the
// exporter frees).
static int release_range(ssq_comm *cm, int first, int last) {
    for (int k = first; k < last; k++)
        for (int r = 0; r < cm->world; r++) {
            if (r != cm->rank && cm->peer[k][r]) cudaIpcCloseMemHandle(cm->peer[k][r]);
            cm->peer[k][r] = nullptr;
        }
    int rc = comm_barrier(cm);
    if (rc) return rc;
    for (int k = first; k < last; k++) {
        if (cm->mine[k]) cudaFree(cm->mine[k]);
        cm->mine[k] = nullptr;
    }
    return SSQ_OK;
}

static int release_buffers(ssq_comm *cm, bool with_flags) {
    int rc = release_range(cm, 0, with_flags ? kBufFlags + 1 : kBufFlags);
    if (!with_flags) { cm->cap = 0; cm->rb_cap = 0; }
    return rc;
}

static void detach_counter(ssq_comm *cm) {
    if (cm->attached && cm->attached->stream_out == &cm->so) cm->attached->stream_out = nullptr;
    cm->attached = nullptr;
}

static int release_stream_buffers(ssq_comm *cm) {
    detach_counter(cm);
    if (!cm->x_ready) return SSQ_OK;
    cm->x_ready = false;
    SSQ_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    return release_range(cm, kBufXE, kNumBufs);
}

// Collective: allocate buffers k in [first, last) with the given sizes, exchange their IPC handles, map the peers'.
// Sets cm->peer_ok = false (on EVERY rank alike) when some rank cannot export or map a buffer.
static int share_buffers(ssq_comm *cm, int first, int last, const size_t *bytes) {
    ssq_ctx *ctx = cm->ctx;
    const int nb = last - first;
    std::vector<unsigned char> handles((size_t)nb * 64 + 8, 0);
    int64_t ok = 1;
    for (int k = first; k < last; k++) {
        SSQ_CUDA(cudaMalloc(&cm->mine[k], bytes[k - first] ? bytes[k - first] : 1));
        SSQ_CUDA(cudaMemsetAsync(cm->mine[k], 0, bytes[k - first] ? bytes[k - first] : 1, ctx->stream));
        cudaIpcMemHandle_t h;
        if (cm->peer_ok && cudaIpcGetMemHandle(&h, cm->mine[k]) == cudaSuccess) memcpy(&handles[(size_t)(k - first) * 64], &h, 64);
        else { ok = 0; cudaGetLastError(); }
    }
    memcpy(&handles[(size_t)nb * 64], &ok, 8);
    const size_t per = (size_t)nb * 64 + 8;
    unsigned char *d_send = nullptr, *d_recv = nullptr;
    SSQ_CUDA(cudaMalloc(&d_send, per));
    SSQ_CUDA(cudaMalloc(&d_recv, per * cm->world));
    SSQ_CUDA(cudaMemcpyAsync(d_send, handles.data(), per, cudaMemcpyHostToDevice, ctx->stream));
    SSQ_NCCL(nccl()->AllGather(d_send, d_recv, per, ncclUint8, cm->comm, ctx->stream));
    std::vector<unsigned char> all(per * cm->world);
    SSQ_CUDA(cudaMemcpyAsync(all.data(), d_recv, per * cm->world, cudaMemcpyDeviceToHost, ctx->stream));
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_send);
    cudaFree(d_recv);
    bool everyone = cm->peer_ok;
    for (int r = 0; r < cm->world; r++) {
        int64_t rok;
        memcpy(&rok, &all[per * r + (size_t)nb * 64], 8);
        everyone = everyone && rok != 0;
    }
    int64_t mapped = 1;
    for (int k = first; k < last; k++)
        for (int r = 0; r < cm->world; r++) {
            cm->peer[k][r] = nullptr;
            if (r == cm->rank) { cm->peer[k][r] = cm->mine[k]; continue; }
            if (!everyone) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, &all[per * r + (size_t)(k - first) * 64], 64);
            if (cudaIpcOpenMemHandle(&cm->peer[k][r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cm->peer[k][r] = nullptr;
                mapped = 0;
                cudaGetLastError();
            }
        }
    // every rank must take the same road: agree on whether all mappings succeeded
    SSQ_CUDA(cudaMemcpyAsync(cm->d_mine, &mapped, 8, cudaMemcpyHostToDevice, ctx->stream));
    SSQ_NCCL(nccl()->AllGather(cm->d_mine, cm->d_matrix, 1, ncclInt64, cm->comm, ctx->stream));
    SSQ_CUDA(cudaMemcpyAsync(cm->h_matrix, cm->d_matrix, 8 * cm->world, cudaMemcpyDeviceToHost, ctx->stream));
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < cm->world; r++) everyone = everyone && cm->h_matrix[r] != 0;
    cm->peer_ok = everyone;
    return SSQ_OK;
}

static int ensure_buffers(ssq_comm *cm, int64_t need, int64_t need_regions, int64_t wpt) {
    if (need <= cm->cap && need_regions <= cm->rb_cap && wpt <= cm->words_per_tuple) return SSQ_OK;
    int rc = release_buffers(cm, false);
    if (rc) return rc;
    const int64_t cap = need + need / 4 + 1024;
    const int64_t rb_cap = need_regions > 1 ? need_regions : 1;
    const size_t bytes[4] = {(size_t)8 * wpt * cap, (size_t)cap, (size_t)8 * cap, (size_t)8 * cm->world * (rb_cap + 1)};
    rc = share_buffers(cm, kBufWords, kBufFlags, bytes);
    if (rc) return rc;
    cm->cap = cap;
    cm->rb_cap = rb_cap;
    cm->words_per_tuple = wpt;
    return SSQ_OK;
}

static int ensure_stage(ssq_comm *cm, int64_t tuples, int64_t wpt) {
    if (tuples <= cm->stage_cap) return SSQ_OK;
    SSQ_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    for (int k = 0; k < 3; k++) { if (cm->stage[k]) cudaFree(cm->stage[k]); cm->stage[k] = nullptr; }
    const int64_t cap = tuples + tuples / 4 + 1024;
    SSQ_CUDA(cudaMalloc(&cm->stage[0], (size_t)8 * 3 * cap));      // sized for the widest key
    SSQ_CUDA(cudaMalloc(&cm->stage[1], (size_t)cap));
    SSQ_CUDA(cudaMalloc(&cm->stage[2], (size_t)8 * cap));
    cm->stage_cap = cap;
    (void)wpt;
    return SSQ_OK;
}

}  // namespace ssq

using namespace ssq;

extern "C" {

int ssq_comm_unique_id(uint8_t *id128) {
    SSQ_ARG(id128 != nullptr, "id128 is NULL");
    if (!nccl()->ok) { set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols"); return SSQ_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    SSQ_NCCL(nccl()->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SSQ_OK;
}

int ssq_comm_init(ssq_ctx *ctx, const uint8_t *id128, int rank, int world, ssq_comm **out) {
    SSQ_ARG(ctx != nullptr && id128 != nullptr && out != nullptr, "NULL argument");
    SSQ_ARG(world >= 1 && world <= kMaxRanks && (world & (world - 1)) == 0 && rank >= 0 && rank < world, "world must be a power of two <= 32");
    *out = nullptr;
    if (!nccl()->ok) { set_error("NCCL (libnccl.so.2) could not be loaded"); return SSQ_ERR_CUDA; }
    DeviceGuard g(ctx->device);
    ssq_comm *cm = new ssq_comm();
    memset(cm, 0, sizeof(*cm));
    cm->ctx = ctx;
    cm->rank = rank;
    cm->world = world;
    cm->peer_ok = getenv("SSQ_NO_PEER_EXCHANGE") == nullptr;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    SSQ_NCCL(nccl()->CommInitRank(&cm->comm, world, id, rank));
    SSQ_CUDA(cudaMalloc(&cm->d_mine, 8 * (world + 1)));
    SSQ_CUDA(cudaMalloc(&cm->d_matrix, 8 * (size_t)world * (world + 1)));
    SSQ_CUDA(cudaHostAlloc(&cm->h_matrix, 8 * (size_t)world * (world + 1), cudaHostAllocDefault));
    SSQ_CUDA(cudaMalloc(&cm->d_table, 8 * 5 * (size_t)world));
    SSQ_CUDA(cudaHostAlloc(&cm->h_table, 8 * 5 * (size_t)world, cudaHostAllocDefault));
    SSQ_CUDA(cudaMemsetAsync(cm->d_mine, 0, 8 * (world + 1), ctx->stream));
    for (int i = 0; i < 3; i++) SSQ_CUDA(cudaEventCreate(&cm->ev[i]));
    SSQ_CUDA(cudaMalloc(&cm->d_xtable, 8 * 5 * (size_t)world));
    const size_t fbytes[1] = {(size_t)8 * world};
    int rc = share_buffers(cm, kBufFlags, kBufFlags + 1, fbytes);      // arrival flags: fixed size, shared once
    if (rc) return rc;
    *out = cm;
    return SSQ_OK;
}

int ssq_comm_destroy(ssq_comm *cm) {
    if (!cm) return SSQ_OK;
    DeviceGuard g(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    release_stream_buffers(cm);
    release_buffers(cm, true);
    cudaFree(cm->d_xtable);
    for (int k = 0; k < 3; k++) if (cm->stage[k]) cudaFree(cm->stage[k]);
    cudaFree(cm->d_mine);
    cudaFree(cm->d_matrix);
    cudaFreeHost(cm->h_matrix);
    cudaFree(cm->d_table);
    cudaFreeHost(cm->h_table);
    for (int i = 0; i < 3; i++) cudaEventDestroy(cm->ev[i]);
    nccl()->CommDestroy(cm->comm);
    delete cm;
    return SSQ_OK;
}

int ssq_comm_uses_peer_stores(ssq_comm *cm) { return cm && cm->peer_ok ? 1 : 0; }

int ssq_comm_last_merge_streamed(ssq_comm *cm) { return cm && cm->last_streamed ? 1 : 0; }

int ssq_comm_attach(ssq_comm *cm, ssq_counter *local, ssq_counter *owner, int *streams) {
    SSQ_ARG(cm != nullptr, "communicator is NULL");
    ssq_ctx *ctx = cm->ctx;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    const int P = cm->world, me = cm->rank;
    if (streams) *streams = 0;
    detach_counter(cm);
    // what this rank proposes: the local table's capacity and the owner's hash rotation (-1: nothing to stream)
    int64_t tag[2] = {-1, -1};
    int lr = 0;
    int64_t block_regions = 0;
    if (local != nullptr && owner != nullptr && cm->peer_ok && local->klass == SSQ_CLASS_64 && owner->klass == SSQ_CLASS_64 &&
        local->ctx == ctx && owner->ctx == ctx) {
        lr = region_bits_for(local->log2_cap);
        const int log2_regions = local->log2_cap - lr;
        int log2_parts = 0;
        while ((1 << log2_parts) < P) log2_parts++;
        block_regions = ((int64_t)1 << log2_regions) >> log2_parts;
        if (local->log2_cap >= 22 && lr <= 14 && log2_regions >= 6 && log2_regions >= log2_parts &&
            counter_stream_nests(owner, P, block_regions, lr)) {
            tag[0] = local->log2_cap;
            tag[1] = owner->hash_rot;
        }
    }
    SSQ_CUDA(cudaMemcpyAsync(cm->d_mine, tag, 16, cudaMemcpyHostToDevice, st));
    SSQ_NCCL(nccl()->AllGather(cm->d_mine, cm->d_matrix, 2, ncclInt64, cm->comm, st));
    SSQ_CUDA(cudaMemcpyAsync(cm->h_matrix, cm->d_matrix, 16 * (size_t)P, cudaMemcpyDeviceToHost, st));
    SSQ_CUDA(cudaStreamSynchronize(st));
    bool same = tag[0] >= 0;
    for (int r = 0; r < P; r++) same = same && cm->h_matrix[2 * r] == tag[0] && cm->h_matrix[2 * r + 1] == tag[1];
    bool any = false;
    for (int r = 0; r < P; r++) any = any || cm->h_matrix[2 * r] >= 0;
    if (!same) {                                   // every rank sees the same matrix and takes the same road
        int rc = release_stream_buffers(cm);
        if (rc) return rc;
        if (any && local != nullptr) { set_error("ssq_comm_attach: every rank must attach ShortSeq64 tables of the same capacity whose regions nest in the owner's"); return SSQ_ERR_ARG; }
        return SSQ_OK;
    }
    if (!cm->x_ready || cm->so.log2_cap != local->log2_cap) {
        int rc = release_stream_buffers(cm);
        if (rc) return rc;
        const size_t eb = (size_t)16 << local->log2_cap, cb = (size_t)4 * P * block_regions;
        const size_t bytes[4] = {eb, eb, cb, cb};
        rc = share_buffers(cm, kBufXE, kNumBufs, bytes);
        if (rc) return rc;
        cm->x_ready = true;
        if (!cm->peer_ok) return release_stream_buffers(cm);      // a rank could not map: nobody streams
    }
    cm->x_block_regions = block_regions;
    cm->x_log2_sregion = lr;
    // this sender's block inside every owner's buffers, per parity; the owners' flag words
    const size_t block_entries = (size_t)block_regions << lr;
    for (int par = 0; par < 2; par++)
        for (int d = 0; d < P; d++) {
            cm->h_table[(0 + par) * P + d] = (int64_t)((ulonglong2 *)cm->peer[kBufXE + par][d] + (size_t)me * block_entries);
            cm->h_table[(2 + par) * P + d] = (int64_t)((uint32_t *)cm->peer[kBufXC + par][d] + (size_t)me * block_regions);
        }
    for (int d = 0; d < P; d++) cm->h_table[4 * P + d] = (int64_t)cm->peer[kBufFlags][d];
    SSQ_CUDA(cudaMemcpyAsync(cm->d_xtable, cm->h_table, 8 * 5 * (size_t)P, cudaMemcpyHostToDevice, st));
    SSQ_CUDA(cudaStreamSynchronize(st));
    cm->so.d_dst[0] = cm->d_xtable;
    cm->so.d_dst[1] = cm->d_xtable + P;
    cm->so.d_cnt[0] = cm->d_xtable + 2 * P;
    cm->so.d_cnt[1] = cm->d_xtable + 3 * P;
    cm->so.log2_parts = 0;
    while ((1 << cm->so.log2_parts) < P) cm->so.log2_parts++;
    cm->so.rot_owner = owner->hash_rot;
    cm->so.rank = me;
    cm->so.log2_cap = local->log2_cap;
    cm->so.epoch = &cm->epoch;
    cm->so.attached_slot = &cm->attached;
    cm->attached = local;
    local->stream_out = &cm->so;
    local->streamed_seq = 0;
    if (streams) *streams = 1;
    return SSQ_OK;
}

int ssq_counter_merge_alltoall(ssq_comm *cm, ssq_counter *local, ssq_counter *owner, float *exchange_ms, float *merge_ms) {
    SSQ_ARG(cm != nullptr && local != nullptr && owner != nullptr, "NULL argument");
    SSQ_ARG(local->klass == owner->klass, "local and owner counters differ in class");
    SSQ_ARG(local->ctx == cm->ctx && owner->ctx == cm->ctx, "counters and communicator must share a context");
    ssq_ctx *ctx = cm->ctx;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    const int P = cm->world, me = cm->rank;
    const int W = local->klass == SSQ_CLASS_64 ? 1 : 3;
    if (exchange_ms) *exchange_ms = 0.0f;
    if (merge_ms) *merge_ms = 0.0f;
    cm->last_streamed = false;
    if (cm->x_ready) {
        // ---- streamed exchange: the regions already sit in the owners' buffers if every rank's last pass streamed them --------
        const int par = (int)((cm->epoch + 1) & 1);
        int64_t tag = -1;
        if (cm->attached == local && local->stream_out == &cm->so && local->streamed_seq != 0 && local->streamed_seq == local->mod_seq &&
            local->streamed_parity == par && local->log2_cap == cm->so.log2_cap && owner->hash_rot == cm->so.rot_owner &&
            counter_stream_nests(owner, P, cm->x_block_regions, cm->x_log2_sregion))
            tag = local->log2_cap;
        SSQ_CUDA(cudaMemcpyAsync(cm->d_mine, &tag, 8, cudaMemcpyHostToDevice, st));
        SSQ_NCCL(nccl()->AllGather(cm->d_mine, cm->d_matrix, 1, ncclInt64, cm->comm, st));
        SSQ_CUDA(cudaMemcpyAsync(cm->h_matrix, cm->d_matrix, 8 * (size_t)P, cudaMemcpyDeviceToHost, st));
        SSQ_CUDA(cudaStreamSynchronize(st));
        bool all = tag >= 0;
        for (int r = 0; r < P; r++) all = all && cm->h_matrix[r] == tag;
        if (all) {
            cm->last_streamed = true;
            cm->epoch++;
            SSQ_CUDA(cudaEventRecord(cm->ev[0], st));
            signal_kernel<<<1, 32, 0, st>>>((u64 *const *)(cm->d_xtable + 4 * P), P, me, cm->epoch);
            SSQ_LAUNCH_CHECK();
            SSQ_CUDA(cudaEventRecord(cm->ev[1], st));
            int rc2 = counter_merge_streamed_impl(owner, cm->mine[kBufXE + par], (const uint32_t *)cm->mine[kBufXC + par], P,
                                                  cm->x_block_regions, cm->x_log2_sregion, (const u64 *)cm->mine[kBufFlags], cm->epoch);
            if (rc2) return rc2;
            SSQ_CUDA(cudaEventRecord(cm->ev[2], st));
            SSQ_CUDA(cudaEventSynchronize(cm->ev[2]));
            if (exchange_ms) SSQ_CUDA(cudaEventElapsedTime(exchange_ms, cm->ev[0], cm->ev[1]));
            if (merge_ms) SSQ_CUDA(cudaEventElapsedTime(merge_ms, cm->ev[1], cm->ev[2]));
            return SSQ_OK;
        }
    }
    int64_t local_regions = 0;
    ssq_counter_regions(local, &local_regions);
    const bool region_export = local->klass == SSQ_CLASS_64 && local_regions >= P;
    bool peer = cm->peer_ok && region_export;
    // ShortSeq192 tables have no regions: two passes over the table (count, then scatter), but the scatter pass stores
    // straight into the owners' buffers too -- same arrival flags, weighted insert on the owner
    const bool peer2 = cm->peer_ok && !region_export;
    int rc;

    // ---- 1. sizes ---------------------------------------------------------------------------------------------
    int64_t n_local = 0;
    if (region_export) {
        rc = ssq_counter_export_counts(local, P, cm->d_mine);                 // from the region occupancy, no table pass
        if (rc) return rc;
    } else if (peer2) {
        rc = counter_export_count_pass(local, P, cm->d_mine);
        if (rc) return rc;
    } else {
        rc = ssq_counter_size(local, &n_local);                                // staged path: the export counts while it groups
        if (rc) return rc;
        rc = ensure_stage(cm, n_local, W);
        if (rc) return rc;
        rc = ssq_counter_export(local, P, (uint64_t *)cm->stage[0], (uint8_t *)cm->stage[1], (uint64_t *)cm->stage[2], nullptr, cm->d_mine);
        if (rc) return rc;
    }
    SSQ_CUDA(cudaMemcpyAsync(cm->d_mine + P, &local_regions, 8, cudaMemcpyHostToDevice, st));
    SSQ_NCCL(nccl()->AllGather(cm->d_mine, cm->d_matrix, P + 1, ncclInt64, cm->comm, st));
    SSQ_CUDA(cudaMemcpyAsync(cm->h_matrix, cm->d_matrix, 8 * (size_t)P * (P + 1), cudaMemcpyDeviceToHost, st));
    SSQ_CUDA(cudaStreamSynchronize(st));          // the only host wait before the end: P (P + 1) numbers
    // m[src * (P + 1) + dst] = tuples src -> dst; m[src * (P + 1) + P] = src's regions (a copy: growing the buffers reuses the pinned matrix)
    const std::vector<int64_t> mat(cm->h_matrix, cm->h_matrix + (size_t)P * (P + 1));
    const int64_t *m = mat.data();
    int64_t max_in = 0, max_regions = 0, my_in = 0, before_me[kMaxRanks], block_counts[kMaxRanks], block_regions[kMaxRanks];
    for (int d = 0; d < P; d++) {
        int64_t in = 0;
        for (int s = 0; s < P; s++) {
            if (s == me) before_me[d] = in;
            in += m[s * (P + 1) + d];
        }
        if (in > max_in) max_in = in;
        if (d == me) my_in = in;
    }
    for (int s = 0; s < P; s++) {
        block_counts[s] = m[s * (P + 1) + me];
        block_regions[s] = m[s * (P + 1) + P] / P;
        if (block_regions[s] > max_regions) max_regions = block_regions[s];
        if (m[s * (P + 1) + P] < P) peer = false;          // every rank decides alike: the matrix is the same everywhere
    }
    rc = ensure_buffers(cm, max_in, max_regions, W);     // collective, but every rank sees the same numbers
    if (rc) return rc;
    peer = peer && cm->peer_ok;
    cm->epoch++;
    SSQ_CUDA(cudaEventRecord(cm->ev[0], st));

    if (!peer && peer2 && cm->peer_ok) {
        // ---- ShortSeq192 over peer memory: scatter pass of the export = exchange, flags, weighted insert on the owner ----
        const int64_t rb_stride = cm->rb_cap + 1;
        for (int d = 0; d < P; d++) {
            cm->h_table[0 * P + d] = (int64_t)((uint64_t *)cm->peer[kBufWords][d] + before_me[d] * W);
            cm->h_table[1 * P + d] = (int64_t)((uint8_t *)cm->peer[kBufLens][d] + before_me[d]);
            cm->h_table[2 * P + d] = (int64_t)((uint64_t *)cm->peer[kBufCounts][d] + before_me[d]);
            cm->h_table[4 * P + d] = (int64_t)cm->peer[kBufFlags][d];
        }
        SSQ_CUDA(cudaMemcpyAsync(cm->d_table, cm->h_table, 8 * 5 * (size_t)P, cudaMemcpyHostToDevice, st));
        rc = counter_export_scatter_to(local, P, (u64 *const *)(cm->d_table), (uint8_t *const *)(cm->d_table + P), (u64 *const *)(cm->d_table + 2 * P));
        if (rc) return rc;
        signal_kernel<<<1, 32, 0, st>>>((u64 *const *)(cm->d_table + 4 * P), P, me, cm->epoch);
        SSQ_LAUNCH_CHECK();
        SSQ_CUDA(cudaEventRecord(cm->ev[1], st));
        for (int s = 0; s < P; s++) block_regions[s] = 0;     // no region grid: counter_merge_regions_impl takes its plain path
        rc = counter_merge_regions_impl(owner, (const u64 *)cm->mine[kBufWords], (const uint8_t *)cm->mine[kBufLens],
                                        (const u64 *)cm->mine[kBufCounts], block_counts, block_regions, P,
                                        (const int64_t *)cm->mine[kBufBases], rb_stride, (const u64 *)cm->mine[kBufFlags], cm->epoch);
        if (rc) return rc;
    } else

    if (peer) {
        // ---- 2. export = exchange: peer stores, then the arrival flags -------------------------------------------
        const int64_t rb_stride = cm->rb_cap + 1;
        for (int d = 0; d < P; d++) {
            cm->h_table[0 * P + d] = (int64_t)((uint64_t *)cm->peer[kBufWords][d] + before_me[d]);
            cm->h_table[1 * P + d] = (int64_t)((uint8_t *)cm->peer[kBufLens][d] + before_me[d]);
            cm->h_table[2 * P + d] = (int64_t)((uint64_t *)cm->peer[kBufCounts][d] + before_me[d]);
            cm->h_table[3 * P + d] = (int64_t)((int64_t *)cm->peer[kBufBases][d] + (int64_t)me * rb_stride);
            cm->h_table[4 * P + d] = (int64_t)cm->peer[kBufFlags][d];
        }
        SSQ_CUDA(cudaMemcpyAsync(cm->d_table, cm->h_table, 8 * 5 * (size_t)P, cudaMemcpyHostToDevice, st));
        // rank r starts with owner r + 1: at any moment every owner receives from one sender
        rc = ssq_counter_export_to(local, P, (me + 1) % P, (uint64_t *const *)(cm->d_table), (uint8_t *const *)(cm->d_table + P),
                                   (uint64_t *const *)(cm->d_table + 2 * P));
        if (rc) return rc;
        rc = ssq_counter_export_region_bases(local, P, (int64_t *const *)(cm->d_table + 3 * P));
        if (rc) return rc;
        signal_kernel<<<1, 32, 0, st>>>((u64 *const *)(cm->d_table + 4 * P), P, me, cm->epoch);
        SSQ_LAUNCH_CHECK();
        SSQ_CUDA(cudaEventRecord(cm->ev[1], st));
        // ---- 3. the owner side: wait for the P flags on the device, count region by region ----------------------
        rc = counter_merge_regions_impl(owner, (const u64 *)cm->mine[kBufWords], (const uint8_t *)cm->mine[kBufLens],
                                        (const u64 *)cm->mine[kBufCounts], block_counts, block_regions, P,
                                        (const int64_t *)cm->mine[kBufBases], rb_stride, (const u64 *)cm->mine[kBufFlags], cm->epoch);
        if (rc) return rc;
    } else {
        // ---- staged path: all-to-all-v of the grouped export with grouped ncclSend / ncclRecv --------------------
        if (region_export || peer2) {   // the sizes came from the region occupancy / the count pass: export now
            rc = ssq_counter_size(local, &n_local);
            if (rc) return rc;
            rc = ensure_stage(cm, n_local, W);
            if (rc) return rc;
            rc = ssq_counter_export(local, P, (uint64_t *)cm->stage[0], (uint8_t *)cm->stage[1], (uint64_t *)cm->stage[2], nullptr, cm->d_mine);
            if (rc) return rc;
        }
        SSQ_NCCL(nccl()->GroupStart());
        int64_t soff = 0, roff = 0;
        for (int p = 0; p < P; p++) {
            const int64_t sc = m[me * (P + 1) + p], rcnt = m[p * (P + 1) + me];
            if (sc) {
                SSQ_NCCL(nccl()->Send((const uint64_t *)cm->stage[0] + soff * W, (size_t)sc * W, ncclUint64, p, cm->comm, st));
                SSQ_NCCL(nccl()->Send((const uint8_t *)cm->stage[1] + soff, (size_t)sc, ncclUint8, p, cm->comm, st));
                SSQ_NCCL(nccl()->Send((const uint64_t *)cm->stage[2] + soff, (size_t)sc, ncclUint64, p, cm->comm, st));
            }
            if (rcnt) {
                SSQ_NCCL(nccl()->Recv((uint64_t *)cm->mine[kBufWords] + roff * W, (size_t)rcnt * W, ncclUint64, p, cm->comm, st));
                SSQ_NCCL(nccl()->Recv((uint8_t *)cm->mine[kBufLens] + roff, (size_t)rcnt, ncclUint8, p, cm->comm, st));
                SSQ_NCCL(nccl()->Recv((uint64_t *)cm->mine[kBufCounts] + roff, (size_t)rcnt, ncclUint64, p, cm->comm, st));
            }
            soff += sc;
            roff += rcnt;
        }
        SSQ_NCCL(nccl()->GroupEnd());
        SSQ_CUDA(cudaEventRecord(cm->ev[1], st));
        if (my_in) {
            rc = ssq_counter_merge(owner, (const uint64_t *)cm->mine[kBufWords], (const uint8_t *)cm->mine[kBufLens],
                                   (const uint64_t *)cm->mine[kBufCounts], my_in);
            if (rc) return rc;
        }
    }
    SSQ_CUDA(cudaEventRecord(cm->ev[2], st));
    SSQ_CUDA(cudaEventSynchronize(cm->ev[2]));
    if (exchange_ms) SSQ_CUDA(cudaEventElapsedTime(exchange_ms, cm->ev[0], cm->ev[1]));
    if (merge_ms) SSQ_CUDA(cudaEventElapsedTime(merge_ms, cm->ev[1], cm->ev[2]));
    return SSQ_OK;
}

}  // extern "C"
