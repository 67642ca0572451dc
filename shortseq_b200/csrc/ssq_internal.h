// ssq_internal.h -- host-side structures behind the opaque C-ABI handles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/shortseq_b200.h"
#include "ssq_device.cuh"

// Device staging buffers of the host-buffer pipeline (ssq_host.cu): allocated on first use, grown when a call
// needs more, released with the context -- cudaMalloc/cudaFree per call cost more than the copies they stage.
struct ssq_host_staging {
    uint8_t *ascii[2];
    int64_t *offsets[2];
    uint8_t *lens_in[2];
    ssq::u64 *words[2];
    uint8_t *lens[2];
    size_t ascii_bytes, reads, word_entries;   // current capacities (per buffer)
    cudaEvent_t ev_in[2], ev_out[2];
    bool events;
};

struct ssq_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;          // the stream work is enqueued on (own or external)
    cudaStream_t copy_streams[2]; // host pipeline: H2D / D2H
    ssq::DevReport *d_report;     // device error record
    ssq::DevReport *h_report;     // pinned mirror
    ssq_host_staging staging;     // zero-initialised
    void *one_host, *one_dev;     // mapped pinned page of the single-object calls (ssq_one.cu), allocated on first use
    uint32_t one_seq;
    void *scratch;                // small device scratch (scan block totals, export cursors); grown on demand
    size_t scratch_bytes;
};

// Streamed exchange (ssq_comm_attach): owned by an ssq_comm; a counter attached to it sends every region of its
// deferred passes to the owners' receive buffers from inside the count kernel (ssq_counter.cu: StreamOut).
struct ssq_stream_out {
    void *d_dst[2];               // device arrays [P] of entry-block pointers (this sender's block in every owner), per parity
    void *d_cnt[2];               // device arrays [P] of count-block pointers, per parity
    int log2_parts, rot_owner, rank;
    int log2_cap;                 // the local table geometry the receive buffers are sized for
    const uint64_t *epoch;        // the communicator's merge counter: a pass streams into parity (*epoch + 1) & 1
    struct ssq_counter **attached_slot;   // the communicator's pointer to the attached counter (cleared when that counter dies)
};

struct ssq_counter {
    ssq_ctx *ctx;
    int klass;            // SSQ_CLASS_64 | SSQ_CLASS_192
    int hash_rot;
    int log2_cap;         // capacity = 1 << log2_cap slots
    void *slots;          // class 64: {key, count}[cap] (16 B); class 192: {meta, w0, w1, w2}[cap] (32 B)
    ssq::u64 *first_idx;  // optional side array [cap] (lazily allocated)
    ssq::u32 *region_count;  // class 64: occupied slots per table region [cap >> region bits]
    int64_t *region_base;    // class 64: [regions + 1] exclusive scan of region_count (filled by the export)
    ssq::u64 *d_size;     // number of occupied slots (device)
    ssq::u64 *h_size;     // pinned
    ssq::u64 *d_gate;     // {stop flag, first stopped sub-batch} (device, see run_gated)
    ssq::u64 *h_gate;     // pinned
    int64_t expected_unique;   // caller's bound on the distinct keys (0 = unknown: conservative gated inserts)
    int64_t known_size;        // occupied slots as of the last host read-back (finish_pass / run_gated / clear / grow)
    ssq::u64 *part_keys;       // partition buffers of the deferred-insert path (lazily sized)
    ssq::u32 *part_cursor;     // [part_ctas][kParts] segment fill counts
    int64_t part_cap;          // key entries currently allocated
    int part_ctas;             // scatter CTAs the count array is sized for
    ssq::u64 *region_keys;     // level-2 (per table region) buffers of the deferred path
    ssq::u32 *region_cursor;   // [kParts][slices][kParts] segment fill counts
    int64_t region_cap;        // key entries currently allocated
    int64_t region_segs;       // segments the count array is sized for
    cudaEvent_t ev[4];         // last bounded pass: start / after pack+scatter / after the region scatter / end
    int last_pass_phases;      // 0 none, 1 direct single kernel, 2 deferred (scatter + count)
    ssq_stream_out *stream_out;   // non-null while attached to a communicator (ssq_comm_attach)
    uint64_t mod_seq;             // bumped by everything that changes keys or counts
    uint64_t streamed_seq;        // mod_seq of the pass whose regions were streamed (valid while == mod_seq)
    int streamed_parity;          // receive-buffer parity that pass wrote
};

namespace ssq {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SSQ_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ssq::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// SSQ_DEBUG_SYNC=1 in the environment makes every launch synchronous so that a device fault is
// attributed to the kernel that caused it (compute-sanitizer is not available on the GPU pool).
bool debug_sync();
void count_launch();   // every kernel launch of this library is counted (ssq_launch_count)
#define SSQ_LAUNCH_CHECK()                                        \
    do {                                                          \
        ssq::count_launch();                                      \
        SSQ_CUDA(cudaGetLastError());                             \
        if (ssq::debug_sync()) SSQ_CUDA(cudaDeviceSynchronize()); \
    } while (0)

#define SSQ_ARG(cond, msg)                                   \
    do {                                                     \
        if (!(cond)) { ssq::set_error("invalid argument: %s", msg); return SSQ_ERR_ARG; } \
    } while (0)

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// grid size for a grid-stride kernel: enough CTAs to fill the GPU, capped by the work
inline int grid_for(const ssq_ctx *ctx, int64_t work_items, int ctas_per_sm) {
    int64_t g = (int64_t)ctx->sm_count * ctas_per_sm;
    if (work_items < g) g = work_items;
    if (g < 1) g = 1;
    return (int)g;
}

// Device scratch of at least `bytes` bytes, valid until the next call that asks for more (work using it must be
// enqueued on ctx->stream).  Growing synchronises the device; cudaMallocAsync is avoided on purpose: its pool is
// trimmed at every stream synchronisation, which made per-chunk allocations cost milliseconds.
int ctx_scratch(ssq_ctx *ctx, size_t bytes, void **out);

// fused pack+count over a (possibly staged) slice (ssq_counter.cu)
int pack_count_impl(ssq_counter *c, const uint8_t *ascii, int64_t lo, int64_t hi, const int64_t *offsets, int64_t n,
                    int64_t index_base, u64 *words, uint8_t *lens);

// first-occurrence tracking with explicit read numbers (ssq_counter.cu): key i occurred at base_index + indices[i]
// (indices == nullptr: base_index + i)
int counter_first_index_indexed(ssq_counter *c, const u64 *words, const uint8_t *lens, int64_t n, const int64_t *indices,
                                int64_t base_index);

// exclusive scan helpers (ssq_scan.cu)
int scan_lens_to_offsets(ssq_ctx *ctx, const void *lens, int len_bytes, int64_t n, int64_t *out);
int scan_var_words(ssq_ctx *ctx, const int64_t *offsets, int64_t n, int64_t *word_off);
int scan_u32_counts(ssq_ctx *ctx, const u32 *counts, int64_t n, int64_t *out /*[n+1]*/);
int scan_i64(ssq_ctx *ctx, const int64_t *values, int64_t n, int64_t *out /*[n+1]*/);
int scan_synth_lens(ssq_ctx *ctx, uint64_t seed, int64_t first_read, int64_t n, int64_t n_keys,
                    int32_t len_lo, int32_t len_hi, int64_t *offsets);

}  // namespace ssq
