// ssq_counter.cu -- the dedup counter: a device hash table of (length, words) -> count.
//
// Replaces ShortSeqCounter (reference counter.pyx:10-54).  See ssq_table.cuh for the slot
// formats.  Host-side policy in this file:
//   * capacity is a power of two >= 2 x expected_unique (min 2^16 slots);
//   * a batch is enqueued as sub-batches of cap/8 reads, each preceded by a one-thread
//     "gate" kernel that sets a stop flag when the table could exceed 75 % load during the
//     sub-batch; kernels behind a raised flag do nothing.  The host synchronises once per
//     batch, and only if the flag was raised grows the table (x4, rehash on the device) and
//     resumes from the stopped sub-batch.  This conservative mode is used when the caller gave no
//     bound on the distinct keys (expected_unique = 0);
//   * with a bound (expected_unique > 0, capacity >= 2x the bound) a batch is processed in ONE pass.
//     Tables that fit in L2 take the keys directly.  Larger ShortSeq64 tables use the two-phase
//     DEFERRED path: phase 1 (fused into the pack kernel, or scatter_packed_kernel for packed input)
//     appends each key to one of 256 hash partitions; phase 2 (count_parts_kernel) walks the
//     partitions in order, so all SMs insert into the same 1/256 of the table at a time and that
//     region stays resident in L2 -- random DRAM sector traffic becomes two streaming passes over
//     8 bytes per read.  Before a pass the table grows if size + min(n, bound - size) keys would load it beyond 75 %
//     (make_room); after the pass it grows (x2) once it is more than 60 % full; a table that holds more keys than the
//     bound has proven the bound wrong and the counter continues in the conservative mode.  Only a single pass that
//     alone brings far more distinct keys than the bound can still fill a region: that is reported as
//     SSQ_ERR_TABLE_FULL, never silently.
#include <math.h>
#include <stdlib.h>
#include "ssq_internal.h"
#include "ssq_table.cuh"

namespace ssq {

int launch_pack_count(ssq_ctx *ctx, int klass, bool scatter, const uint8_t *ascii, int64_t lo, int64_t hi,
                      const int64_t *offsets, int64_t n, int64_t index_base, u64 *words, uint8_t *lens,
                      const TableView &t, const PartView &pv, const u64 *stop);

constexpr int kThreads = 256;

static TableView view_of(const ssq_counter *c);
int counter_merge_regions_impl(ssq_counter *c, const u64 *words, const uint8_t *lens, const u64 *counts, const int64_t *block_counts,
                               const int64_t *block_regions, int n_blocks, const int64_t *region_bases, int64_t rb_stride,
                               const u64 *flags, u64 epoch);

// gate[0] = stop flag, gate[1] = index of the first stopped sub-batch
__global__ void gate_kernel(u64 *gate, const u64 *size, u64 incoming, u64 limit, u64 batch_index) {
    if (gate[0] == 0 && *size + incoming > limit) { gate[0] = 1; gate[1] = batch_index; }
}

__device__ __forceinline__ void block_add_new(const TableView &t, u32 my_new, u32 *s_new) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
    if ((threadIdx.x & 31) == 0) s_new[threadIdx.x >> 5] = my_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 tot = 0;
        for (int k = 0; k < (int)(blockDim.x / 32); k++) tot += s_new[k];
        if (tot) atomicAdd(t.size, tot);
    }
}

__device__ __forceinline__ bool len_in_class(int klass, u32 len) {
    return klass == SSQ_CLASS_64 ? len <= 32 : (len >= 33 && len <= 96);
}

// Insert packed keys; counts == nullptr means weight 1.
template <int KLASS>
__global__ void __launch_bounds__(kThreads) insert_kernel(TableView t, const u64 *words, const uint8_t *lens,
                                                          const u64 *counts, int64_t n, int64_t index_base,
                                                          const u64 *stop) {
    __shared__ u32 s_new[kThreads / 32];
    if (stop != nullptr && *stop != 0) return;
    u32 my_new = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        u32 len = lens[i];
        u64 add = counts ? counts[i] : 1ull;
        bool is_new = false;
        u64 slot = 0;
        if (!len_in_class(KLASS, len)) {
            atomicMin(&t.rep->first_bad_len, (u64)(index_base + i));
        } else if (add != 0) {
            if constexpr (KLASS == SSQ_CLASS_64) slot = insert64(t, words[i], len, add, is_new, true);
            else insert192(t, words[3 * i], words[3 * i + 1], words[3 * i + 2], len, add, is_new);
        }
        if constexpr (KLASS == SSQ_CLASS_64) add_region_counts(t, is_new, slot);
        my_new += is_new ? 1u : 0u;
    }
    block_add_new(t, my_new, s_new);
}

// ShortSeq192 direct inserts with the home slots of a thread's keys loaded TOGETHER (one 256-bit load each, L2 evict_last)
// before the first probe, as count_parts192_kernel does.  The weighted insert of a multi-GPU merge receives its tuples in
// the senders' table order -- nearly this table's order -- so the slots stream through L2 while the probes run.
constexpr int kIns192 = 2;
__global__ void __launch_bounds__(kThreads) insert192_pre_kernel(TableView t, const u64 *words, const uint8_t *lens, const u64 *counts,
                                                                 int64_t n, int64_t index_base, const u64 *stop) {
    __shared__ u32 s_new[kThreads / 32];
    if (stop != nullptr && *stop != 0) return;
    const u64 keep = l2_policy_evict_last();
    const int sh = 64 - t.log2_cap;
    u32 my_new = 0;
    for (int64_t i0 = (int64_t)blockIdx.x * (kThreads * kIns192); i0 < n; i0 += (int64_t)gridDim.x * (kThreads * kIns192)) {
        u64 w0[kIns192], w1[kIns192], w2[kIns192], add[kIns192], h2[kIns192];
        u32 len[kIns192];
        bool ok[kIns192];
#pragma unroll
        for (int j = 0; j < kIns192; j++) {
            const int64_t i = i0 + j * kThreads + threadIdx.x;
            ok[j] = i < n;
            if (ok[j]) {
                w0[j] = words[3 * i]; w1[j] = words[3 * i + 1]; w2[j] = words[3 * i + 2];
                len[j] = lens[i];
                add[j] = counts ? counts[i] : 1ull;
                if (!len_in_class(SSQ_CLASS_192, len[j])) { atomicMin(&t.rep->first_bad_len, (u64)(index_base + i)); ok[j] = false; }
                else if (add[j] == 0) ok[j] = false;
                else h2[j] = rotl64(hash192(w0[j], w1[j], w2[j], len[j]), t.rot);
            }
        }
        u64 sm[kIns192], s0[kIns192], s1[kIns192], s2[kIns192];
#pragma unroll
        for (int j = 0; j < kIns192; j++)
            if (ok[j]) ld_relaxed_v4u64_hint(t.slots + 4 * (h2[j] >> sh), keep, sm[j], s0[j], s1[j], s2[j]);
#pragma unroll
        for (int j = 0; j < kIns192; j++) {
            if (!ok[j]) continue;
            bool is_new = false;
            insert192_impl<true>(t, h2[j], w0[j], w1[j], w2[j], len[j], add[j], is_new, sm[j], s0[j], s1[j], s2[j]);
            my_new += is_new ? 1u : 0u;
        }
    }
    block_add_new(t, my_new, s_new);
}

constexpr int kMaxMergeBlocks = 32;

// Region-aligned weighted merge (ShortSeq64 owner tables of a multi-GPU merge).  Every sender's block is ordered by
// ITS table regions, and it also sent the exclusive scan of its region sizes (region_bases).  An owner region is a
// whole number of consecutive sender regions, so the tuples that belong to one owner region are one contiguous range
// of every block: a CTA loads the owner region into shared memory, adds those ranges with shared-memory atomics and
// writes the region back -- one pass, no global atomics (the weighted global insert it replaces is bound by L2
// atomic throughput).
struct MergeRegions {
    int64_t off[kMaxMergeBlocks + 1];      // first tuple of every block
    int ratio_log2[kMaxMergeBlocks];       // log2(sender regions per owner region)
    int n;
    int64_t rb_stride;                     // int64 entries between the region_bases of consecutive blocks
};

// Multi-GPU merge: flags[b] (this GPU's memory) is set to the merge's epoch by sender b once its block has landed
// (st.release.sys behind its export kernel, ssq_comm.cu).  Bounded spin: a sender that never shows up surfaces as
// DevReport::exchange_timeout instead of a hung GPU.  Returns false on a timeout.
__device__ __forceinline__ bool wait_arrival(const u64 *flag, u64 epoch, DevReport *rep) {
    u64 v = 0;
    for (u64 it = 0;; ++it) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= epoch) return true;
        if (it > (1ull << 24)) { atomicAdd(&rep->exchange_timeout, 1ull); return false; }
        __nanosleep(200);
    }
}
__global__ void wait_flags_kernel(const u64 *flags, int n, u64 epoch, DevReport *rep) {
    if (blockIdx.x == 0 && (int)threadIdx.x < n) wait_arrival(flags + threadIdx.x, epoch, rep);
}

// The probing scheme is count_regions2_kernel's (below): a lane makes ONE probe per tuple straight from registers; tuples
// that did not settle are parked -- ballot-compacted -- in a warp-private queue and probed again 32 at a time, so no lane
// waits for another lane's probe sequence (the first version ran every tuple's probe loop to its end, warp-wide: 7.7 warp
// instructions per tuple, 2.4 ms per 1e8 tuples).  The region's counts are 64-bit DELTAS kept as two 32-bit halves: one
// native 32-bit shared-memory atomic on the low half, the carry (and a sender count >= 2^32) added to the high half.
constexpr int kMerge2Threads = 512;
constexpr int kMerge2KPT = 2;
constexpr int kMerge2Chunk = 32 * kMerge2KPT;
constexpr int kMerge2Queue = 32 + kMerge2Chunk;          // a chunk is only started with fewer than 32 entries parked

static size_t merge_regions2_smem(int log2_region) {
    return ((size_t)16 << log2_region) + (size_t)20 * kMerge2Queue * (kMerge2Threads / 32);
}

// One probe of key (klo, khi) with weight `add` at slot `off`; true: found or inserted, delta added.
__device__ __forceinline__ bool probe_once_weighted(u32 ks_a, u32 cs_a, u32 klo, u32 khi, u32 off, u64 add, u32 &my_new) {
    u32 clo, chi;
    lds_v2(ks_a + off * 8, clo, chi);
    if (chi == 0) {                          // empty slot (a key's high word holds len + 1 in bits 58..63: never 0): claim it
        const u64 old = atoms_cas_u64(ks_a + off * 8, 0ull, ((u64)khi << 32) | klo);
        clo = (u32)old;
        chi = (u32)(old >> 32);
        if (chi == 0) { ++my_new; clo = klo; chi = khi; }
    }
    if (clo == klo && chi == khi) {
        const u32 alo = (u32)add;
        const u32 old = atoms_add_u32(cs_a + off * 8, alo);
        const u32 hi_add = (u32)(add >> 32) + ((u32)(old + alo) < alo ? 1u : 0u);
        if (hi_add) reds_add_u32(cs_a + off * 8 + 4, hi_add);
        return true;
    }
    return false;
}

// STREAMED: the blocks are the fixed-place receive buffers of the streamed exchange (count_regions2_kernel<., true>):
// block b = sender b's entries[regions per owner][2^log2_sregion] of {key in THIS table's format, count} and
// cnts[regions per owner]; an owner region is 2^ratio_log2 consecutive sender regions of every block.
struct MergeStream {
    const ulonglong2 *entries;
    const u32 *cnts;
    int64_t block_regions;        // sender regions per block
    int log2_sregion;
    int ratio_log2;
};

template <bool STREAMED>
__global__ void __launch_bounds__(kMerge2Threads, 2) merge_regions_kernel(TableView t, const u64 *words, const uint8_t *lens, const u64 *counts,
                                                                         const int64_t *region_bases, MergeRegions mr, MergeStream ms,
                                                                         const u64 *flags, u64 epoch) {
    extern __shared__ __align__(16) u64 dyn_region[];
    constexpr u32 W = kMerge2Threads / 32;
    __shared__ u32 s_new[W];
    __shared__ int s_arrived;
    __shared__ int64_t s_lo[kMaxMergeBlocks];
    __shared__ u32 s_pre[kMaxMergeBlocks + 1];           // exclusive scan of this region's tuple counts per block
    if (flags != nullptr) {                                   // every sender's block must have landed before anything is read
        if (threadIdx.x == 0) s_arrived = 1;
        __syncthreads();
        if ((int)threadIdx.x < mr.n && !wait_arrival(flags + threadIdx.x, epoch, t.rep)) s_arrived = 0;
        __syncthreads();
        if (!s_arrived) return;
    }
    const u32 R = 1u << t.log2_region, rmask = R - 1;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *ks = dyn_region, *cs = dyn_region + R;
    u64 *qk = dyn_region + 2 * R + warp * kMerge2Queue;                                           // parked keys of this warp
    u64 *qc = dyn_region + 2 * R + (W + warp) * kMerge2Queue;                                     // their weights
    u32 *qo = reinterpret_cast<u32 *>(dyn_region + 2 * R + 2 * W * kMerge2Queue) + warp * kMerge2Queue;   // slot | probes << 16
    const u32 region = blockIdx.x;
    ulonglong2 *gslots = reinterpret_cast<ulonglong2 *>(t.slots) + ((size_t)region << t.log2_region);
    const int nseg = STREAMED ? mr.n << ms.ratio_log2 : mr.n;       // <= 32: one lane each
    if (warp == 0) {                                          // this region's tuple range in every block (loads overlap the region load)
        int64_t lo = 0, hi = 0;
        if ((int)lane < nseg) {
            if constexpr (STREAMED) {                         // segment = (sender block, sender region inside this region)
                const int64_t sreg = ((int64_t)region << ms.ratio_log2) + (lane & ((1u << ms.ratio_log2) - 1));
                const int64_t at = (int64_t)(lane >> ms.ratio_log2) * ms.block_regions + sreg;
                lo = at << ms.log2_sregion;
                hi = lo + min(ms.cnts[at], 1u << ms.log2_sregion);
            } else {
                const int64_t *rb = region_bases + (size_t)lane * mr.rb_stride;
                lo = mr.off[lane] + rb[(size_t)region << mr.ratio_log2[lane]];
                hi = mr.off[lane] + rb[((size_t)region + 1) << mr.ratio_log2[lane]];
            }
            s_lo[lane] = lo;
        }
        u32 incl = (u32)(hi - lo);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
        if ((int)lane < nseg) s_pre[lane + 1] = incl;
        if (lane == 0) s_pre[0] = 0;
    }
    const bool was_empty = t.region_count[region] == 0;
    for (u32 i = threadIdx.x; i < R; i += kMerge2Threads) {
        ks[i] = was_empty ? 0ull : gslots[i].x;
        cs[i] = 0;
    }
    __syncthreads();
    const u32 total = s_pre[nseg];
    const int off_shift = 64 - t.log2_cap;
    const u32 ks_a = smem_addr(ks), cs_a = smem_addr(cs), qk_a = smem_addr(qk), qc_a = smem_addr(qc), qo_a = smem_addr(qo);
    const u32 lt_mask = (1u << lane) - 1;
    u32 my_new = 0, overflow = 0, bad = 0, seg = 0;
    u32 qn = 0;                                          // warp-uniform: parked entries
    auto park = [&](bool want, u32 klo, u32 khi, u64 add, u32 slot_probes) {
        const u32 m = __ballot_sync(0xFFFFFFFFu, want);
        if (want) {
            const u32 at = qn + __popc(m & lt_mask);
            sts_u64(qk_a + at * 8, ((u64)khi << 32) | klo);
            sts_u64(qc_a + at * 8, add);
            sts_u32(qo_a + at * 4, slot_probes);
        }
        qn += __popc(m);
    };
    auto serve = [&]() {                                 // pop up to 32 entries, one more probe each, re-park what is still unsettled
        __syncwarp();
        const u32 take = min(qn, 32u);
        const u32 idx = qn - take + lane;
        u32 klo = 0, khi = 0, sp = 0;
        u64 add = 0;
        const bool mine = lane < take;
        if (mine) { lds_v2(qk_a + idx * 8, klo, khi); add = lds_u64(qc_a + idx * 8); sp = lds_u32(qo_a + idx * 4); }
        __syncwarp();
        qn -= take;
        bool again = false;
        if (mine) {
            const u32 off = sp & 0xFFFFu, probes = sp >> 16;
            if (probes > rmask) ++overflow;              // went around: the region is full
            else if (!probe_once_weighted(ks_a, cs_a, klo, khi, off, add, my_new)) { again = true; sp = ((off + 1) & rmask) | ((probes + 1) << 16); }
        }
        park(again, klo, khi, add, sp);
    };
    u64 nw[kMerge2KPT], na[kMerge2KPT];
    u32 nl[kMerge2KPT];
    auto load_chunk = [&](u32 c) {
        const u32 i0 = c * kMerge2Chunk;
        while (i0 >= s_pre[seg + 1]) ++seg;                     // warp-uniform: the block the chunk starts in
        u32 sg = seg;
#pragma unroll
        for (int r = 0; r < kMerge2KPT; r++) {
            const u32 i = i0 + r * 32 + lane;
            nl[r] = 0xFFFFFFFFu;
            if (i < total) {
                while (i >= s_pre[sg + 1]) ++sg;
                const int64_t src = s_lo[sg] + (i - s_pre[sg]);
                if constexpr (STREAMED) {
                    const ulonglong2 e = ms.entries[src];
                    nw[r] = e.x;
                    na[r] = e.y;
                    nl[r] = 0;
                } else {
                    nw[r] = words[src];
                    na[r] = counts[src];
                    nl[r] = lens[src];
                }
            }
        }
    };
    u32 c = warp;
    if (c * kMerge2Chunk < total) load_chunk(c);
    while (c * kMerge2Chunk < total) {
        u64 wv[kMerge2KPT], av[kMerge2KPT];
        u32 lv[kMerge2KPT];
#pragma unroll
        for (int j = 0; j < kMerge2KPT; j++) { wv[j] = nw[j]; av[j] = na[j]; lv[j] = nl[j]; }
        c += W;
        if (c * kMerge2Chunk < total) load_chunk(c);          // in flight during the probes
#pragma unroll
        for (int j = 0; j < kMerge2KPT; j++) {
            bool pending = false;
            u32 klo = 0, khi = 0, off = 0;
            if (STREAMED && lv[j] != 0xFFFFFFFFu && av[j] != 0) {
                // the entry is a table key already: its slot bits below the top six must name this region
                const u32 slot_low = (u32)((wv[j] & kMask58) >> off_shift);
                const u32 len1 = (u32)(wv[j] >> 58);
                if (len1 == 0 || len1 > 33 || (slot_low >> t.log2_region) != (region & ((1u << (t.log2_cap - t.log2_region - 6)) - 1))) ++bad;
                else {
                    klo = (u32)wv[j];
                    khi = (u32)(wv[j] >> 32);
                    off = slot_low & rmask;
                    pending = !probe_once_weighted(ks_a, cs_a, klo, khi, off, av[j], my_new);
                }
            } else if (!STREAMED && lv[j] != 0xFFFFFFFFu && av[j] != 0) {
                const u64 h2 = table_hash64(wv[j], t.rot);
                if (lv[j] > 32 || (u32)((h2 >> off_shift) >> t.log2_region) != region) ++bad;     // not this region's tuple
                else {
                    const u64 key = key64_of(h2, lv[j]);
                    klo = (u32)key;
                    khi = (u32)(key >> 32);
                    off = (u32)(h2 >> off_shift) & rmask;
                    pending = !probe_once_weighted(ks_a, cs_a, klo, khi, off, av[j], my_new);
                }
            }
            park(pending, klo, khi, av[j], ((off + 1) & rmask) | (1u << 16));
        }
        while (qn >= 32) serve();
    }
    while (qn > 0) serve();
    __syncthreads();
    for (u32 i = threadIdx.x; i < R; i += kMerge2Threads) {
        const u64 d = cs[i];
        if (was_empty) gslots[i] = make_ulonglong2(ks[i], d);
        else if (d) gslots[i] = make_ulonglong2(ks[i], gslots[i].y + d);
    }
    if (overflow) atomicAdd(&t.rep->table_overflow, (u64)overflow);
    if (bad) atomicMin(&t.rep->first_bad_len, 0ull);        // inconsistent blocks: surfaces as an error, never as wrong counts
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
    if (lane == 0) s_new[warp] = my_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (u32 k = 0; k < W; k++) tot += s_new[k];
        if (tot) { atomicAdd(t.size, (u64)tot); red_add_u32(t.region_count + region, tot); }
    }
}

// dst[p][j] = region_base[p * R + j] - region_base[p * R], j = 0..R (R = regions per partition): what owner p needs to find
// its regions' tuples inside the block this rank sends it
__global__ void export_region_bases_kernel(const int64_t *region_base, int log2_regions, int log2_parts, int64_t *const *dst) {
    const int64_t per = (int64_t)1 << (log2_regions - log2_parts);
    const int64_t total = (per + 1) << log2_parts;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / (per + 1), j = i - p * (per + 1);
        dst[p][j] = region_base[p * per + j] - region_base[p * per];
    }
}

template <int KLASS>
__global__ void __launch_bounds__(kThreads) lookup_kernel(TableView t, const u64 *words, const uint8_t *lens, int64_t n,
                                                          u64 *counts) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        u32 len = lens[i];
        u64 c = 0;
        if (len_in_class(KLASS, len)) {
            if constexpr (KLASS == SSQ_CLASS_64) {
                u64 s = find64(t, words[i], len);
                if (s != kNoIndex) c = ld_relaxed_u64(t.slots + 2 * s + 1);
            } else {
                u64 s = find192(t, words[3 * i], words[3 * i + 1], words[3 * i + 2], len);
                if (s != kNoIndex) c = ld_relaxed_u64(t.slots + 4 * s) >> 8;
            }
        }
        counts[i] = c;
    }
}

// indices == nullptr: key i occurred at read base_index + i; else at read base_index + indices[i]
template <int KLASS>
__global__ void __launch_bounds__(kThreads) first_index_kernel(TableView t, const u64 *words, const uint8_t *lens,
                                                               int64_t n, int64_t base_index, const int64_t *indices) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        u32 len = lens[i];
        if (!len_in_class(KLASS, len)) continue;
        u64 s;
        if constexpr (KLASS == SSQ_CLASS_64) s = find64(t, words[i], len);
        else s = find192(t, words[3 * i], words[3 * i + 1], words[3 * i + 2], len);
        if (s != kNoIndex) red_min_u64(t.first_idx + s, (u64)(base_index + (indices ? indices[i] : i)));
    }
}

// Re-insert every occupied slot of `src` into `dst` (growth).
template <int KLASS>
__global__ void __launch_bounds__(kThreads) rehash_kernel(TableView src, TableView dst) {
    __shared__ u32 s_new[kThreads / 32];
    const u64 cap = 1ull << src.log2_cap;
    u32 my_new = 0;
    for (u64 s = (u64)blockIdx.x * kThreads + threadIdx.x; s < cap; s += (u64)gridDim.x * kThreads) {
        bool is_new = false;
        u64 ds = kNoIndex;
        if constexpr (KLASS == SSQ_CLASS_64) {
            u64 key = src.slots[2 * s];
            if (key != 0) {
                u32 len;
                u64 h2 = slot64_h2(src, s, key, len);
                ds = insert64_hashed(dst, h2, key, src.slots[2 * s + 1], is_new);
            }
        } else {
            u64 meta = src.slots[4 * s];
            if ((meta & 0xFF) != 0)
                ds = insert192(dst, src.slots[4 * s + 1], src.slots[4 * s + 2], src.slots[4 * s + 3],
                               (u32)(meta & 0xFF) + 32, meta >> 8, is_new);
        }
        if (ds != kNoIndex && src.first_idx != nullptr) dst.first_idx[ds] = src.first_idx[s];
        my_new += is_new ? 1u : 0u;
    }
    block_add_new(dst, my_new, s_new);
}

// ---- deferred (hash-partitioned) inserts, ShortSeq64 ------------------------------------------
// Three streaming passes replace one random DRAM access per read:
//   1. level-1 scatter (fused into the pack kernel, or scatter_packed_kernel for packed input): every key goes
//      to one of 256 partitions (top 8 hash bits);
//   2. region_scatter_kernel: the keys of one level-1 partition go on to the <= 256 table regions inside it;
//   3. count_regions_kernel: one CTA per region loads the region's slots into shared memory, counts the
//      region's keys with shared-memory atomics and writes the slots back -- the table itself is read and
//      written exactly once, coalesced, and no global atomic is issued.
// Tables whose regions do not fit in shared memory (>= 2^30 slots) fall back to count_parts_kernel, which inserts
// level-1 partitions in order so that the table range being hit (cap/256 slots) stays resident in L2.
constexpr int kCountKeysPerThread = 8;
constexpr int kScatterRounds = kLineKeys / 4;    // keys per thread between flushes (mean 4 per 32-key ring)

// Level 1 for already packed input.
__global__ void __launch_bounds__(kThreads, 3) scatter_packed_kernel(TableView t, PartView pv, const u64 *words,
                                                                     const uint8_t *lens, int64_t n, int64_t index_base) {
    extern __shared__ __align__(16) u64 dyn_ring[];
    __shared__ u32 s_head[kParts], s_tail[kParts];
    __shared__ u32 s_new[kThreads / 32];
    __shared__ u32 s_list[(kThreads / 32) * 64];
    __shared__ u32 s_unstaged_new;
    const Stager stg = make_stager(dyn_ring, s_head, s_tail, s_list);
    u64 *const seg0 = pv.keys + (size_t)blockIdx.x * kParts * pv.seg_cap;
    stager_init(stg);
    if (threadIdx.x == 0) s_unstaged_new = 0;
    __syncthreads();
    u32 my_new = 0;
    const int64_t tile_keys = (int64_t)kThreads * kScatterRounds;
    const int64_t ntiles = (n + tile_keys - 1) / tile_keys;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        u64 word[kScatterRounds];
        u32 len[kScatterRounds];
#pragma unroll
        for (int k = 0; k < kScatterRounds; k++) {
            const int64_t i = tile * tile_keys + k * kThreads + threadIdx.x;
            len[k] = 0xFFFFFFFFu;
            if (i < n) { word[k] = words[i]; len[k] = lens[i]; }
        }
#pragma unroll
        for (int k = 0; k < kScatterRounds; k++) {
            if (len[k] == 0xFFFFFFFFu) continue;
            if (len[k] > 32) {
                atomicMin(&t.rep->first_bad_len, (u64)(index_base + tile * tile_keys + k * kThreads + threadIdx.x));
                continue;
            }
            const u64 h2 = table_hash64(word[k], t.rot);
            const u64 key = key64_of(h2, len[k]);
            if (!stage_key(stg, (u32)(h2 >> 56), key)) {
                bool is_new = false;
                insert64_hashed(t, h2, key, 1ull, is_new);
                my_new += is_new ? 1u : 0u;
            }
        }
        __syncthreads();
        flush_lines<false>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
        __syncthreads();
    }
    flush_lines<true>(stg, seg0, pv.seg_cap, t, -1, &s_unstaged_new);
    __syncthreads();
    pv.seg_count[(size_t)blockIdx.x * kParts + threadIdx.x] = stager_seg_count(stg, threadIdx.x, pv.seg_cap);
    if (threadIdx.x == 0) my_new += s_unstaged_new;
    block_add_new(t, my_new, s_new);
}

// A CTA's input of the level-2 scatter / the region count: a few segments read as one concatenated stream.
// pre[k] = number of keys in the segments before segment k (pre[nseg] = total); each thread walks the
// segments monotonically, so the segment of an index is found by advancing a private cursor.
constexpr int kMaxStreamSegs = 640;       // scatter CTAs of the fused pack pass (148 SMs x up to 4)

// Level 2: CTA (p, s) reads slice s of the level-1 segments of partition p (scatter CTAs [c0, c1)) and appends
// every key to the segment of its level-2 partition (hash bits 55..48).  The loads of the next tile are in
// flight while the current tile is staged and flushed.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 3) region_scatter_kernel(TableView t, PartView pv, RegionParts rp, u32 flush_keys) {
    extern __shared__ __align__(16) u64 dyn_ring[];
    __shared__ u32 s_head[kParts], s_tail[kParts];
    __shared__ u32 pre[kMaxStreamSegs + 1];
    __shared__ u32 s_list[(kParts / 32) * 64];
    __shared__ u32 s_unstaged_new;
    const Stager stg = make_stager(dyn_ring, s_head, s_tail, s_list);
    const u32 p = blockIdx.x / rp.slices, s = blockIdx.x - p * rp.slices;
    const u32 c0 = (u32)((u64)pv.num_ctas * s / rp.slices), c1 = (u32)((u64)pv.num_ctas * (s + 1) / rp.slices);
    const u32 nseg = c1 - c0;
    u64 *const seg0 = rp.keys + (size_t)blockIdx.x * kParts * rp.seg_cap;
    stager_init(stg);
    if (threadIdx.x == 0) s_unstaged_new = 0;
    if (threadIdx.x < 32) {                                   // pre[] = exclusive scan of the segment sizes
        const u32 lane = threadIdx.x;
        u32 carry = 0;
        for (u32 b = 0; b < nseg; b += 32) {
            const u32 k = b + lane;
            const u32 v = k < nseg ? pv.seg_count[(size_t)(c0 + k) * kParts + p] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
            if (k < nseg) pre[k + 1] = carry + incl;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) pre[0] = 0;
    }
    __syncthreads();
    const u64 drop = l2_policy_evict_first();
    const u64 *const part0 = pv.keys + ((size_t)c0 * kParts + p) * pv.seg_cap;     // segment k starts at part0 + k * seg_stride
    const size_t seg_stride = (size_t)kParts * pv.seg_cap;
    constexpr u32 kTile = THREADS * kScatterRounds;
    // Tiles never straddle segments, so a tile's keys sit at consecutive addresses.  (seg, pos) = where the next tile
    // to LOAD starts; both are CTA-uniform.
    u32 seg = 0, pos = 0;
    u32 seg_len = nseg ? pre[1] : 0u;           // keys of segment `seg`: re-read only when the segment changes
    u64 nk[kScatterRounds];
    auto load_tile = [&]() -> u32 {             // returns the number of keys of the tile (0 = stream exhausted)
        while (seg < nseg && pos >= seg_len) { ++seg; pos = 0; seg_len = seg < nseg ? pre[seg + 1] - pre[seg] : 0u; }
        if (seg >= nseg) return 0u;
        const u32 here = min(seg_len - pos, kTile);
        const u64 *src = part0 + seg * seg_stride + pos + threadIdx.x;
#pragma unroll
        for (int j = 0; j < kScatterRounds; j++) nk[j] = (u32)(j * THREADS) + threadIdx.x < here ? ld_stream_u64(src + j * THREADS, drop) : 0ull;
        pos += here;
        return here;
    };
    u32 have = load_tile(), unflushed = 0;
    while (have) {
        u64 k[kScatterRounds];
#pragma unroll
        for (int j = 0; j < kScatterRounds; j++) k[j] = nk[j];
        unflushed += have;
        have = load_tile();                      // in flight while this tile is staged and flushed
        {
            u32 part[kScatterRounds];
            bool valid[kScatterRounds], ok[kScatterRounds];
#pragma unroll
            for (int j = 0; j < kScatterRounds; j++) { part[j] = (u32)(k[j] >> 48) & 0xFFu; valid[j] = k[j] != 0; }
            stage_keys<kScatterRounds>(stg, part, k, valid, ok);
#pragma unroll
            for (int j = 0; j < kScatterRounds; j++)
                if (!ok[j]) insert_unstaged(t, k[j], p, &s_unstaged_new);
        }
        if (unflushed >= flush_keys || !have) {   // flush once enough keys are staged (a short tail tile waits for the next one)
            unflushed = 0;
            __syncthreads();
            flush_lines<false>(stg, seg0, rp.seg_cap, t, (int)p, &s_unstaged_new);
            __syncthreads();
        }
    }
    flush_lines<true>(stg, seg0, rp.seg_cap, t, (int)p, &s_unstaged_new);
    __syncthreads();
    if (threadIdx.x < (u32)kParts) rp.seg_count[(size_t)blockIdx.x * kParts + threadIdx.x] = stager_seg_count(stg, threadIdx.x, rp.seg_cap);
    if (threadIdx.x == 0 && s_unstaged_new) atomicAdd(t.size, (u64)s_unstaged_new);
}

// Level 3: one CTA per table region.  Dynamic shared memory: keys u64[R] | deltas u32[R] | per-warp key queues
// u64[warps][32 * KPT] (R = 2^log2_region slots).  The region's keys are loaded into shared memory; every warp then
// walks its share of the region's key stream in chunks of 32 x KPT keys, loaded coalesced (the next chunk is in
// flight meanwhile) and parked in the warp's queue.  A rolled loop consumes the queue: in every iteration each lane
// makes ONE probe of its current key, and lanes that finished a key take the next queue entries (ballot-ranked, no
// atomics) -- no lane waits for another lane's probe sequence, and the queue is refilled as soon as it runs dry.
// Counts are accumulated as 32-bit deltas; at the end the touched slots are written back (old count re-read from
// L2, where the region load asked it to stay).
constexpr int kCountKPT = 8;
constexpr int kCountChunk = 32 * kCountKPT;
constexpr int kMaxRegionSegs = 384;      // slices (<= 6 counted here; more only with 1:1 regions) x level-2 partitions per region (<= 64)

template <int THREADS>
static size_t count_regions_smem(int log2_region, unsigned nseg) {
    return ((size_t)12 << log2_region) + sizeof(u64) * kCountKPT * THREADS + sizeof(u32) * (2 * nseg + 2);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 3)
count_regions_kernel(TableView t, RegionParts rp) {
    extern __shared__ __align__(16) u64 dyn_region[];
    __shared__ u32 s_new[THREADS / 32];
    constexpr u32 W = THREADS / 32;
    const u32 R = 1u << t.log2_region, rmask = R - 1;
    u64 *ks = dyn_region;
    u32 *ds = reinterpret_cast<u32 *>(dyn_region + R);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *wq = dyn_region + R + R / 2 + warp * kCountChunk;     // this warp's key queue
    u32 *pre = reinterpret_cast<u32 *>(dyn_region + R + R / 2 + W * kCountChunk);   // [nseg + 1] exclusive scan of the segment sizes
    u32 *segoff = pre + (rp.slices << (8 - rp.qbits)) + 1;                          // [nseg] first entry of stream segment k in rp.keys
    const u32 region = blockIdx.x;
    const u32 sub_bits = 8 - rp.qbits, sub_mask = (1u << sub_bits) - 1;
    const u32 p = region >> rp.qbits, q = region & ((1u << rp.qbits) - 1);
    const u32 nseg = rp.slices << sub_bits;                   // stream segment k = (slice k >> sub_bits, level-2 partition (q << sub_bits) | (k & sub_mask))
    auto seg_index = [&](u32 k) -> size_t { return ((size_t)p * rp.slices + (k >> sub_bits)) * kParts + ((q << sub_bits) | (k & sub_mask)); };
    if (warp == 0) {                                          // pre[] = exclusive scan of the segment sizes
        u32 carry = 0;
        for (u32 b = 0; b < nseg; b += 32) {
            const u32 k = b + lane;
            const u32 v = k < nseg ? rp.seg_count[seg_index(k)] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
            if (k < nseg) {
                pre[k + 1] = carry + incl;
                segoff[k] = (u32)(seg_index(k) * rp.seg_cap);      // the host checked that the buffer has < 2^32 entries
            }
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) pre[0] = 0;
    }
    const u64 keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    ulonglong2 *gslots = reinterpret_cast<ulonglong2 *>(t.slots) + ((size_t)region << t.log2_region);
    // a region nobody has inserted into yet (first pass over a cleared table) is known to be all zero: nothing to read
    const bool was_empty = t.region_count[region] == 0;
    for (u32 i = threadIdx.x; i < R; i += THREADS) {
        ks[i] = was_empty ? 0ull : ld_hint_v2u64(gslots + i, keep).x;
        ds[i] = 0;
    }
    __syncthreads();
    const u32 total = pre[nseg];
    const int off_shift = 64 - t.log2_cap;      // home slot = h2 >> off_shift; its low log2_region bits lie inside the key's hash bits
    const u32 lt_mask = (1u << lane) - 1;
    u32 my_new = 0, overflow = 0, seg = 0;
    u64 nk[kCountKPT];
    auto load_chunk = [&](u32 c) {
        const u32 i0 = c * kCountChunk;
        while (i0 >= pre[seg + 1]) ++seg;                       // warp-uniform: the segment the chunk starts in
        if (i0 + kCountChunk <= pre[seg + 1]) {                 // the usual case: the whole chunk lies in one segment
            const u64 *src = rp.keys + (segoff[seg] + (i0 - pre[seg])) + lane;
#pragma unroll
            for (int r = 0; r < kCountKPT; r++) nk[r] = ld_stream_u64(src + r * 32, drop);
        } else {                                                // it straddles segments or the end of the stream
            u32 sg = seg;
#pragma unroll
            for (int r = 0; r < kCountKPT; r++) {
                const u32 i = i0 + r * 32 + lane;
                nk[r] = 0;
                if (i < total) {
                    while (i >= pre[sg + 1]) ++sg;
                    nk[r] = ld_stream_u64(rp.keys + (segoff[sg] + (i - pre[sg])), drop);
                }
            }
        }
    };
    u32 c = warp;
    bool more = c * kCountChunk < total;         // warp-uniform: nk[] holds a chunk that is not yet in the queue
    if (more) load_chunk(c);
    u32 qpos = kCountChunk;                      // warp-uniform: next unread queue entry (>= kCountChunk: queue empty)
    // The loop works on 32-bit shared-space addresses and on key halves.  A key's high word is never 0 (it holds
    // len + 1 in bits 58..63), so `khi == 0` means "this lane is between keys" / "empty slot".  The home slot's
    // region offset lies in the high word too: off_shift >= 32 for every table this kernel is launched on.
    const u32 ks_a = smem_addr(ks), ds_a = smem_addr(ds), wq_a = smem_addr(wq);
    const u32 hi_shift = (u32)off_shift - 32;
    u32 klo = 0, khi = 0, off = 0, home = 0;
    for (;;) {
        if (qpos >= (u32)kCountChunk && more) {  // refill the queue, start the loads of the chunk after it
#pragma unroll
            for (int r = 0; r < kCountKPT; r++) sts_u64(wq_a + (r * 32 + lane) * 8, nk[r]);
            __syncwarp();
            qpos = 0;
            c += W;
            more = c * kCountChunk < total;
            if (more) load_chunk(c);
        }
        const u32 need = __ballot_sync(0xFFFFFFFFu, khi == 0);
        if (khi == 0) {
            const u32 idx = qpos + __popc(need & lt_mask);
            if (idx < (u32)kCountChunk) {
                lds_v2(wq_a + idx * 8, klo, khi);
                home = off = (khi >> hi_shift) & rmask;
            }
        }
        qpos += __popc(need);
        if (!__any_sync(0xFFFFFFFFu, khi != 0)) {
            if (qpos >= (u32)kCountChunk && !more) break;
            continue;                            // only padding was fetched, or a refill is due
        }
        if (khi != 0) {
            u32 clo, chi;
            lds_v2(ks_a + off * 8, clo, chi);
            if (chi == 0) {                      // empty slot: claim it
                const u64 old = atoms_cas_u64(ks_a + off * 8, 0ull, ((u64)khi << 32) | klo);
                clo = (u32)old;
                chi = (u32)(old >> 32);
                if (chi == 0) { ++my_new; clo = klo; chi = khi; }
            }
            if (clo == klo && chi == khi) {
                reds_add_u32(ds_a + off * 4, 1u);
                khi = 0;
            } else {
                off = (off + 1) & rmask;
                if (off == home) { ++overflow; khi = 0; }          // went around: the region is full
            }
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < R; i += THREADS) {
        const u32 d = ds[i];
        // a region that was empty is written whole (full, coalesced sectors; its slots are known to be zero otherwise);
        // elsewhere only the touched slots, with the old count re-read from L2
        if (was_empty) gslots[i] = make_ulonglong2(ks[i], (u64)d);
        else if (d) gslots[i] = make_ulonglong2(ks[i], gslots[i].y + d);
    }
    if (overflow) atomicAdd(&t.rep->table_overflow, (u64)overflow);
    // keys created in this region: table size and the region's occupancy
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
    if (lane == 0) s_new[warp] = my_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (u32 k = 0; k < W; k++) tot += s_new[k];
        if (tot) {
            atomicAdd(t.size, (u64)tot);
            red_add_u32(t.region_count + region, tot);
        }
    }
}

// Level 3, second version: the same CTA-per-region layout, but the queue of count_regions_kernel only sees the keys
// that need it.  A warp streams its share of the region's keys in chunks of 32 x KPT; every lane makes ONE probe of
// each of its KPT keys straight from registers (independent shared-memory loads, 32-bit key halves, a handful of
// instructions per key): at the loads the table policy allows, three keys in four are settled by it.  The others are
// parked -- ballot-compacted, no atomics -- in a small warp-private queue as (key, next slot, probes so far); whenever
// the queue holds a warp's worth, 32 entries are popped, each lane makes one more probe of its entry and re-parks it
// if it is still unsettled, so no lane ever waits for another lane's probe sequence.
constexpr int kCount2KPT = 4;
constexpr int kCount2Chunk = 32 * kCount2KPT;
constexpr int kCount2Queue = 32 + kCount2Chunk;          // a chunk is only started with fewer than 32 entries parked

template <int THREADS>
static size_t count_regions2_smem(int log2_region, unsigned nseg) {
    return ((size_t)12 << log2_region) + (size_t)12 * kCount2Queue * (THREADS / 32) + sizeof(u32) * (2 * nseg + 2);
}

// One probe of key (klo, khi) at slot `off` of the shared-memory region (ks_a: keys, ds_a: count deltas).
// true: the key was found or inserted and its delta incremented; false: the slot belongs to another key.
__device__ __forceinline__ bool probe_once(u32 ks_a, u32 ds_a, u32 klo, u32 khi, u32 off, u32 &my_new) {
    u32 clo, chi;
    lds_v2(ks_a + off * 8, clo, chi);
    if (chi == 0) {                          // empty slot (a key's high word holds len + 1 in bits 58..63: never 0): claim it
        const u64 old = atoms_cas_u64(ks_a + off * 8, 0ull, ((u64)khi << 32) | klo);
        clo = (u32)old;
        chi = (u32)(old >> 32);
        if (chi == 0) { ++my_new; clo = klo; chi = khi; }
    }
    if (clo == klo && chi == khi) { reds_add_u32(ds_a + off * 4, 1u); return true; }
    return false;
}

// Streamed exchange (multi-GPU, ssq_comm_attach): the CTA that has just counted a region also compacts the region's final
// (key, count) pairs -- still in shared memory -- and stores them into the receive buffer of the rank that owns the
// region's hash range, over NVLink when that is another GPU.  Every (sender, sender region) has a FIXED place there
// (2^log2_region entries + one count), so nothing has to be sized or scanned first: the export pass over the table, the
// size matrix and the region-offset arrays of the unstreamed exchange disappear, and the transfer overlaps the count.
// Keys travel in the OWNER table's format (its hash rotation applied), so the owner's merge does not hash at all.
struct StreamOut {
    ulonglong2 *const *dst;      // [1 << log2_parts] this sender's entry block in every owner's receive buffer
    u32 *const *dst_cnt;         // [1 << log2_parts] this sender's count block there
    int log2_parts;
    int rot_owner;
    u32 first_region;            // CTA b counts region (b + first_region) mod regions: rank r starts with owner r + 1, so
                                 // that at any moment every owner receives from one sender
};

template <int THREADS, bool STREAM>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 3)
count_regions2_kernel(TableView t, RegionParts rp, StreamOut so) {
    extern __shared__ __align__(16) u64 dyn_region[];
    __shared__ u32 s_new[THREADS / 32];
    __shared__ u32 s_out;
    constexpr u32 W = THREADS / 32;
    const u32 R = 1u << t.log2_region, rmask = R - 1;
    u64 *ks = dyn_region;
    u32 *ds = reinterpret_cast<u32 *>(dyn_region + R);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *qk = dyn_region + R + R / 2 + warp * kCount2Queue;                                     // parked keys of this warp
    u32 *qo = reinterpret_cast<u32 *>(dyn_region + R + R / 2 + W * kCount2Queue) + warp * kCount2Queue;   // slot | probes << 16
    u32 *pre = reinterpret_cast<u32 *>(dyn_region + R + R / 2 + W * kCount2Queue) + W * kCount2Queue;    // [nseg + 1] exclusive scan of the segment sizes
    u32 *segoff = pre + (rp.slices << (8 - rp.qbits)) + 1;                        // [nseg] first entry of stream segment k in rp.keys
    const u32 region = STREAM ? (blockIdx.x + so.first_region) & (gridDim.x - 1) : blockIdx.x;
    const u32 sub_bits = 8 - rp.qbits, sub_mask = (1u << sub_bits) - 1;
    const u32 p = region >> rp.qbits, q = region & ((1u << rp.qbits) - 1);
    const u32 nseg = rp.slices << sub_bits;
    auto seg_index = [&](u32 k) -> size_t { return ((size_t)p * rp.slices + (k >> sub_bits)) * kParts + ((q << sub_bits) | (k & sub_mask)); };
    if (warp == 0) {                                          // pre[] = exclusive scan of the segment sizes
        u32 carry = 0;
        for (u32 b = 0; b < nseg; b += 32) {
            const u32 k = b + lane;
            const u32 v = k < nseg ? rp.seg_count[seg_index(k)] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
            if (k < nseg) {
                pre[k + 1] = carry + incl;
                segoff[k] = (u32)(seg_index(k) * rp.seg_cap);
            }
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) pre[0] = 0;
    }
    const u64 keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    ulonglong2 *gslots = reinterpret_cast<ulonglong2 *>(t.slots) + ((size_t)region << t.log2_region);
    const bool was_empty = t.region_count[region] == 0;
    for (u32 i = threadIdx.x; i < R; i += THREADS) {
        ks[i] = was_empty ? 0ull : ld_hint_v2u64(gslots + i, keep).x;
        ds[i] = 0;
    }
    if (STREAM && threadIdx.x == 0) s_out = 0;
    __syncthreads();
    const u32 total = pre[nseg];
    const u32 hi_shift = (u32)(64 - t.log2_cap) - 32;   // the home slot's region offset lies in the key's high word
    const u32 ks_a = smem_addr(ks), ds_a = smem_addr(ds), qk_a = smem_addr(qk), qo_a = smem_addr(qo);
    const u32 lt_mask = (1u << lane) - 1;
    u32 my_new = 0, overflow = 0, seg = 0;
    u32 qn = 0;                                          // warp-uniform: parked entries
    // park one entry per lane with `want` set (ballot-compacted append)
    auto park = [&](bool want, u32 klo, u32 khi, u32 slot_probes) {
        const u32 m = __ballot_sync(0xFFFFFFFFu, want);
        if (want) {
            const u32 at = qn + __popc(m & lt_mask);
            sts_u64(qk_a + at * 8, ((u64)khi << 32) | klo);
            sts_u32(qo_a + at * 4, slot_probes);
        }
        qn += __popc(m);
    };
    // pop up to 32 entries, one more probe each, re-park what is still unsettled
    auto serve = [&]() {
        __syncwarp();
        const u32 take = min(qn, 32u);
        const u32 idx = qn - take + lane;
        u32 klo = 0, khi = 0, sp = 0;
        const bool mine = lane < take;
        if (mine) { lds_v2(qk_a + idx * 8, klo, khi); sp = lds_u32(qo_a + idx * 4); }
        __syncwarp();
        qn -= take;
        bool again = false;
        if (mine) {
            u32 off = sp & 0xFFFFu;
            const u32 probes = sp >> 16;
            if (probes > rmask) ++overflow;              // went around: the region is full
            else if (!probe_once(ks_a, ds_a, klo, khi, off, my_new)) { again = true; sp = ((off + 1) & rmask) | ((probes + 1) << 16); }
        }
        park(again, klo, khi, sp);
    };
    u64 nk[kCount2KPT];
    bool full_chunk = false;
    auto load_chunk = [&](u32 c) {
        const u32 i0 = c * kCount2Chunk;
        while (i0 >= pre[seg + 1]) ++seg;                       // warp-uniform: the segment the chunk starts in
        full_chunk = i0 + kCount2Chunk <= pre[seg + 1];
        if (full_chunk) {                                       // the usual case: the whole chunk lies in one segment
            const u64 *src = rp.keys + (segoff[seg] + (i0 - pre[seg])) + lane;
#pragma unroll
            for (int r = 0; r < kCount2KPT; r++) nk[r] = ld_stream_u64(src + r * 32, drop);
        } else {                                                // it straddles segments or the end of the stream
            u32 sg = seg;
#pragma unroll
            for (int r = 0; r < kCount2KPT; r++) {
                const u32 i = i0 + r * 32 + lane;
                nk[r] = 0;
                if (i < total) {
                    while (i >= pre[sg + 1]) ++sg;
                    nk[r] = ld_stream_u64(rp.keys + (segoff[sg] + (i - pre[sg])), drop);
                }
            }
        }
    };
    u32 c = warp;
    if (c * kCount2Chunk < total) load_chunk(c);
    while (c * kCount2Chunk < total) {
        u32 klo[kCount2KPT], khi[kCount2KPT];
#pragma unroll
        for (int j = 0; j < kCount2KPT; j++) { klo[j] = (u32)nk[j]; khi[j] = (u32)(nk[j] >> 32); }
        const bool all_valid = full_chunk;
        c += W;
        if (c * kCount2Chunk < total) load_chunk(c);          // in flight during the probes
        u32 off[kCount2KPT];
        bool hit[kCount2KPT];
#pragma unroll
        for (int j = 0; j < kCount2KPT; j++) off[j] = (khi[j] >> hi_shift) & rmask;
        if (all_valid) {
#pragma unroll
            for (int j = 0; j < kCount2KPT; j++) hit[j] = probe_once(ks_a, ds_a, klo[j], khi[j], off[j], my_new);
        } else {
#pragma unroll
            for (int j = 0; j < kCount2KPT; j++) hit[j] = khi[j] == 0 || probe_once(ks_a, ds_a, klo[j], khi[j], off[j], my_new);
        }
#pragma unroll
        for (int j = 0; j < kCount2KPT; j++) park(!hit[j], klo[j], khi[j], ((off[j] + 1) & rmask) | (1u << 16));
        while (qn >= 32) serve();
    }
    while (qn > 0) serve();
    __syncthreads();
    if constexpr (!STREAM) {
        for (u32 i = threadIdx.x; i < R; i += THREADS) {
            const u32 d = ds[i];
            if (was_empty) gslots[i] = make_ulonglong2(ks[i], (u64)d);
            else if (d) gslots[i] = make_ulonglong2(ks[i], gslots[i].y + d);
        }
    } else {
        // write-back + the region's final tuples to their owner (R and THREADS are multiples of 32: a warp is in or out as one)
        const int log2_regions = t.log2_cap - t.log2_region;
        const u32 owner = region >> (log2_regions - so.log2_parts);
        const u32 in_owner = region & ((1u << (log2_regions - so.log2_parts)) - 1);
        ulonglong2 *out = so.dst[owner] + ((size_t)in_owner << t.log2_region);
        const u64 top6 = (u64)(region >> (log2_regions - 6)) << 58;
        for (u32 i = threadIdx.x; i < R; i += THREADS) {
            const u32 d = ds[i];
            const u64 k = ks[i];
            u64 cnt = d;
            if (was_empty) gslots[i] = make_ulonglong2(k, cnt);
            else if (k != 0) {                                   // keys of earlier passes travel too, with their whole count
                cnt += gslots[i].y;
                if (d) gslots[i] = make_ulonglong2(k, cnt);
            }
            const bool occ = k != 0;
            const u32 m = __ballot_sync(0xFFFFFFFFu, occ);
            u32 base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_out, (u32)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (occ) {
                const u64 h = rotr64(top6 | (k & kMask58), t.rot);          // hash64(word)
                out[base + __popc(m & lt_mask)] = make_ulonglong2((rotl64(h, so.rot_owner) & kMask58) | (k & ~kMask58), cnt);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) so.dst_cnt[owner][in_owner] = s_out;
    }
    if (overflow) atomicAdd(&t.rep->table_overflow, (u64)overflow);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
    if (lane == 0) s_new[warp] = my_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (u32 k = 0; k < W; k++) tot += s_new[k];
        if (tot) {
            atomicAdd(t.size, (u64)tot);
            red_add_u32(t.region_count + region, tot);
        }
    }
}

// Fallback for tables whose regions exceed shared memory: insert the level-1 partitions in order.  Block b handles
// the segment that scatter CTA (b % num_ctas) filled for partition (b / num_ctas); blocks are scheduled in index
// order, so at any time the whole GPU works on one or two neighbouring partitions whose table range (cap/256
// slots) stays in L2.
__global__ void __launch_bounds__(kThreads) count_parts_kernel(TableView t, PartView pv) {
    __shared__ u32 s_new[kThreads / 32];
    const u32 p = blockIdx.x / pv.num_ctas;
    const u32 c = blockIdx.x - p * pv.num_ctas;
    const size_t seg = (size_t)c * kParts + p;
    const u32 cnt = pv.seg_count[seg];
    const u64 *keys = pv.keys + seg * pv.seg_cap;
    const u64 top = (u64)p << 56;
    u32 my_new = 0;
    for (u32 i0 = 0; i0 < cnt; i0 += kThreads * kCountKeysPerThread) {
        u64 k[kCountKeysPerThread];
#pragma unroll
        for (int j = 0; j < kCountKeysPerThread; j++) {
            const u32 i = i0 + j * kThreads + threadIdx.x;
            k[j] = i < cnt ? keys[i] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kCountKeysPerThread; j++) {
            if (k[j] == 0) continue;
            bool is_new = false;
            insert64_hashed(t, top | (k[j] & kMask56), k[j], 1ull, is_new);
            my_new += is_new ? 1u : 0u;
        }
    }
    block_add_new(t, my_new, s_new);
}

// ShortSeq192: the records of the level-1 partitions are inserted in partition order (table range of cap / kParts192
// slots stays in L2); block b handles the segment scatter CTA (b % num_ctas) filled for partition (b / num_ctas).
// A record is {w0, w1, w2, meta}, see meta192_of.  The last num_ctas blocks insert the scatter CTAs' overflow segments
// (unordered, a fraction of a percent of the reads).
// A thread takes kRecs192 records per round: the record loads are issued together, then the 32-byte loads of the home
// slots (L2::evict_last, so that the streaming records do not push the table range out), then the compares / REDs.
// Measured (2e8 records, 1.25e7 distinct): 1 record per thread at 8 CTAs/SM 5.65 ms, 2 at 80 registers 6.4 ms, 4 at 78
// registers 9.9 ms -- the kernel is bound by L2 random-sector operations (one slot load + one RED per record, ~70 G/s),
// which more independent work per thread does not raise while the lost occupancy costs.
#ifndef SSQ_RECS192
#define SSQ_RECS192 1
#endif
constexpr int kRecs192 = SSQ_RECS192;
__global__ void __launch_bounds__(kThreads) count_parts192_kernel(TableView t, PartView pv) {
    const u32 p = blockIdx.x / pv.num_ctas;
    const u32 c = blockIdx.x - p * pv.num_ctas;
    const size_t seg = (size_t)c * kParts192 + p;
    const bool overflow_seg = p >= (u32)kParts192;
    const u32 cnt = overflow_seg ? pv.ovf_count[c] : pv.seg_count[seg];
    const u64 *recs = overflow_seg ? pv.ovf + (size_t)c * pv.ovf_cap * 4 : pv.keys + seg * pv.seg_cap * 4;
    const u64 drop = l2_policy_evict_first(), keep = l2_policy_evict_last();
    const int sh = 64 - t.log2_cap;
    u32 my_new = 0;
    for (u32 i0 = 0; i0 < cnt; i0 += kThreads * kRecs192) {
        u64 w0[kRecs192], w1[kRecs192], w2[kRecs192], mt[kRecs192];
#pragma unroll
        for (int j = 0; j < kRecs192; j++) {
            const u32 i = i0 + j * kThreads + threadIdx.x;
            mt[j] = 0;
            if (i < cnt) ld_stream_v4u64(recs + 4 * (size_t)i, drop, w0[j], w1[j], w2[j], mt[j]);
        }
        u64 sm[kRecs192], s0[kRecs192], s1[kRecs192], s2[kRecs192];
#pragma unroll
        for (int j = 0; j < kRecs192; j++) {
            sm[j] = 0;
            if (i0 + j * kThreads + threadIdx.x < cnt)
                ld_relaxed_v4u64_hint(t.slots + 4 * ((mt[j] & ~0xFFull) >> sh), keep, sm[j], s0[j], s1[j], s2[j]);
        }
#pragma unroll
        for (int j = 0; j < kRecs192; j++) {
            if (i0 + j * kThreads + threadIdx.x >= cnt) continue;
            bool is_new = false;
            insert192_impl<true>(t, mt[j] & ~0xFFull, w0[j], w1[j], w2[j], (u32)(mt[j] & 0xFF) + 32, 1ull, is_new, sm[j], s0[j], s1[j], s2[j]);
            my_new += is_new ? 1u : 0u;
        }
    }
    // one size update per warp (a block handles ~1e4 records; no barrier at the end of a block)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, d);
    if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(t.size, (u64)my_new);
}

// ---- export ---------------------------------------------------------------------------------
constexpr int kExportItems = 8;    // slots per thread
constexpr int kMaxParts = 256;

struct Tuple {
    u64 w0, w1, w2, count, first;
    u32 len, part;
    bool used;
};

template <int KLASS>
__device__ __forceinline__ Tuple read_slot(const TableView &t, u64 s, int log2_parts) {
    Tuple r;
    r.used = false;
    r.w1 = r.w2 = 0;
    u64 h;
    if constexpr (KLASS == SSQ_CLASS_64) {
        u64 key = t.slots[2 * s];
        if (key == 0) return r;
        u64 h2 = slot64_h2(t, s, key, r.len);
        h = rotr64(h2, t.rot);
        r.w0 = unhash64(h);
        r.count = t.slots[2 * s + 1];
    } else {
        u64 meta = t.slots[4 * s];
        if ((meta & 0xFF) == 0) return r;
        r.len = (u32)(meta & 0xFF) + 32;
        r.count = meta >> 8;
        r.w0 = t.slots[4 * s + 1]; r.w1 = t.slots[4 * s + 2]; r.w2 = t.slots[4 * s + 3];
        h = hash192(r.w0, r.w1, r.w2, r.len);
    }
    r.first = t.first_idx ? t.first_idx[s] : kNoIndex;
    r.part = log2_parts ? (u32)(h >> (64 - log2_parts)) : 0u;
    r.used = true;
    return r;
}

template <int KLASS>
__global__ void __launch_bounds__(kThreads) export_count_kernel(TableView t, int log2_parts, u64 *part_counts) {
    __shared__ u32 cnt[kMaxParts];
    for (int p = threadIdx.x; p < kMaxParts; p += kThreads) cnt[p] = 0;
    __syncthreads();
    const u64 cap = 1ull << t.log2_cap;
    // the table is hash-ordered, so a thread sees long runs of one partition: count runs locally
    u32 run_part = 0, run = 0;
    for (u64 s = (u64)blockIdx.x * kThreads + threadIdx.x; s < cap; s += (u64)gridDim.x * kThreads) {
        Tuple r = read_slot<KLASS>(t, s, log2_parts);
        if (!r.used) continue;
        if (r.part != run_part) {
            if (run) atomicAdd(&cnt[run_part], run);
            run_part = r.part;
            run = 0;
        }
        run++;
    }
    if (run) atomicAdd(&cnt[run_part], run);
    __syncthreads();
    for (int p = threadIdx.x; p < (1 << log2_parts); p += kThreads)
        if (cnt[p]) atomicAdd(&part_counts[p], (u64)cnt[p]);
}

// cursors[p] = exclusive scan of part_counts: where partition p starts in the output
__global__ void export_bases_kernel(const u64 *part_counts, u64 *cursors, int nparts) {
    u64 run = 0;
    for (int p = 0; p < nparts; p++) { cursors[p] = run; run += part_counts[p]; }
}

// Coalesced copy of n 64-bit words / bytes from shared memory to global memory (possibly another GPU's): the bulk leaves as
// aligned 16-byte stores by consecutive threads -- fine-grained 8-byte stores cost one NVLink packet each.
__device__ __forceinline__ void copy_out_u64(u64 *dst, const u64 *sm, u32 n) {
    const u32 head = (((uintptr_t)dst & 15) != 0 && n > 0) ? 1u : 0u;
    if (head && threadIdx.x == 0) dst[0] = sm[0];
    const u32 pairs = (n - head) / 2;
    for (u32 i = threadIdx.x; i < pairs; i += blockDim.x)
        *reinterpret_cast<ulonglong2 *>(dst + head + 2 * i) = make_ulonglong2(sm[head + 2 * i], sm[head + 2 * i + 1]);
    if (((n - head) & 1) && threadIdx.x == 32) dst[n - 1] = sm[n - 1];
}
__device__ __forceinline__ void copy_out_u8(uint8_t *dst, const uint8_t *sm, u32 n) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    if (threadIdx.x < head) dst[threadIdx.x] = sm[threadIdx.x];
    const u32 vecs = (n - head) / 16;
    for (u32 i = threadIdx.x; i < vecs; i += blockDim.x) {
        const uint8_t *p = sm + head + 16 * i;
        u32 w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = (u32)p[4 * k] | ((u32)p[4 * k + 1] << 8) | ((u32)p[4 * k + 2] << 16) | ((u32)p[4 * k + 3] << 24);
        *reinterpret_cast<uint4 *>(dst + head + 16 * i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const u32 done = head + 16 * vecs;
    if (threadIdx.x < n - done) dst[done + threadIdx.x] = sm[done + threadIdx.x];
}

// pw / pl / pc (all or none): per-partition destination arrays -- partition p's tuples go to pw[p] / pl[p] / pc[p], which
// may be another GPU's memory (the ShortSeq192 send side of ssq_counter_merge_alltoall); the cursors then start at 0.
template <int KLASS>
__global__ void __launch_bounds__(kThreads) export_scatter_kernel(TableView t, int log2_parts,
                                                                  u64 *cursors, u64 *words, uint8_t *lens, u64 *counts,
                                                                  int64_t *first_idx, u64 *const *pw = nullptr,
                                                                  uint8_t *const *pl = nullptr, u64 *const *pc = nullptr) {
    // dynamic shared memory (optional): staging of a one-partition tile's tuples in output order -- words | counts | lens
    extern __shared__ __align__(16) u64 dyn_stage[];
    constexpr int kTileSlots = kThreads * kExportItems;
    constexpr int EW = KLASS == SSQ_CLASS_64 ? 1 : 3;
    const bool staged_out = first_idx == nullptr && pw != nullptr;       // the launch provided the staging space
    u64 *const st_words = dyn_stage, *const st_counts = dyn_stage + (size_t)EW * kTileSlots;
    uint8_t *const st_lens = reinterpret_cast<uint8_t *>(dyn_stage + (size_t)(EW + 1) * kTileSlots);
    __shared__ u32 s_total;
    __shared__ u32 cnt[kMaxParts];
    __shared__ u64 base[kMaxParts];
    __shared__ u32 s_part[2];
    __shared__ u32 s_warp[kThreads / 32];
    const int nparts = 1 << log2_parts;
    const u64 cap = 1ull << t.log2_cap;
    const u64 tile = (u64)kThreads * kExportItems;
    for (u64 tile0 = (u64)blockIdx.x * tile; tile0 < cap; tile0 += (u64)gridDim.x * tile) {
        for (int p = threadIdx.x; p < nparts; p += kThreads) cnt[p] = 0;
        __syncthreads();
        Tuple r[kExportItems];
        u32 rank[kExportItems];
        u32 n_used = 0, pmin = 0xFFFFFFFFu, pmax = 0;
#pragma unroll
        for (int k = 0; k < kExportItems; k++) {
            u64 s = tile0 + (u64)k * kThreads + threadIdx.x;
            r[k].used = false;
            if (s < cap) r[k] = read_slot<KLASS>(t, s, log2_parts);
            if (r[k].used) { n_used++; pmin = min(pmin, r[k].part); pmax = max(pmax, r[k].part); }
        }
        // The table is hash-ordered: almost every tile lies inside ONE partition.  Then ranks come from a block scan of
        // the per-thread counts; only a tile straddling a partition boundary ranks with shared-memory atomics.
        const bool uniform = __syncthreads_and(n_used == 0 || pmin == pmax) != 0;
        bool one_part = false;
        if (uniform) {
            // all threads that hold entries agree with their own pmin; agree across threads via shared memory
            if (threadIdx.x == 0) { s_part[0] = 0xFFFFFFFFu; s_part[1] = 0; }
            __syncthreads();
            if (n_used) { atomicMin(&s_part[0], pmin); atomicMax(&s_part[1], pmin); }
            __syncthreads();
            one_part = s_part[0] == 0xFFFFFFFFu || s_part[0] == s_part[1];
        }
        if (one_part) {
            // exclusive scan of n_used over the block
            const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            u32 incl = n_used;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            u32 before = 0, total = 0;
            for (int w = 0; w < kThreads / 32; w++) { if (w < (int)warp) before += s_warp[w]; total += s_warp[w]; }
            if (threadIdx.x == 0) { base[0] = total ? atomicAdd(&cursors[s_part[0]], (u64)total) : 0; s_total = total; }
            __syncthreads();
            u32 run = before + incl - n_used;
#pragma unroll
            for (int k = 0; k < kExportItems; k++)
                if (r[k].used) { rank[k] = run++; r[k].part = 0; }      // base[0] holds this tile's start
        } else {
#pragma unroll
            for (int k = 0; k < kExportItems; k++)
                if (r[k].used) rank[k] = atomicAdd(&cnt[r[k].part], 1u);
            __syncthreads();
            for (int p = threadIdx.x; p < nparts; p += kThreads)
                base[p] = cnt[p] ? atomicAdd(&cursors[p], (u64)cnt[p]) : 0;   // cursors start at the partition bases
            __syncthreads();
        }
        const u32 tile_part = one_part ? s_part[0] : 0u;     // the partition every entry of a one-partition tile belongs to
        if (staged_out && one_part) {
            // the tile's tuples are one contiguous output range: order them in shared memory, copy out coalesced
#pragma unroll
            for (int k = 0; k < kExportItems; k++) {
                if (!r[k].used) continue;
                const u32 o = rank[k];
                if constexpr (KLASS == SSQ_CLASS_64) st_words[o] = r[k].w0;
                else { st_words[3 * o] = r[k].w0; st_words[3 * o + 1] = r[k].w1; st_words[3 * o + 2] = r[k].w2; }
                st_counts[o] = r[k].count;
                st_lens[o] = (uint8_t)r[k].len;
            }
            __syncthreads();
            const u32 total = s_total;
            if (total) {
                const u64 b0 = base[0];
                copy_out_u64(pw[tile_part] + (size_t)EW * b0, st_words, EW * total);
                copy_out_u64(pc[tile_part] + b0, st_counts, total);
                copy_out_u8(pl[tile_part] + b0, st_lens, total);
            }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int k = 0; k < kExportItems; k++) {
            if (!r[k].used) continue;
            u64 o = base[r[k].part] + rank[k];
            u64 *wd = words; uint8_t *ld = lens; u64 *cd = counts;
            if (pw != nullptr) { const u32 dp = one_part ? tile_part : r[k].part; wd = pw[dp]; ld = pl[dp]; cd = pc[dp]; }
            if constexpr (KLASS == SSQ_CLASS_64) wd[o] = r[k].w0;
            else { wd[3 * o] = r[k].w0; wd[3 * o + 1] = r[k].w1; wd[3 * o + 2] = r[k].w2; }
            ld[o] = (uint8_t)r[k].len;
            cd[o] = r[k].count;
            if (first_idx) first_idx[o] = (int64_t)r[k].first;
        }
        __syncthreads();
    }
}

// ---- single-pass export of ShortSeq64 tables ---------------------------------------------------
// The table is hash-ordered and probing never leaves a region, so hash partition p is exactly the slot range
// [p * cap / P, (p + 1) * cap / P) and the occupied slots of every region are known (TableView::region_count).
// An exclusive scan of those counts gives every region its place in the output; one pass over the table then
// compacts the regions in order -- no counting pass, no atomics, a deterministic order.  With per-partition
// destination pointers (ssq_counter_export_to) partition p's tuples go straight to dst[p], which may be memory of
// another GPU (NVLink peer stores): the export IS the send side of the multi-GPU exchange.
struct ExportDst {
    u64 *words; uint8_t *lens; u64 *counts; int64_t *first_idx;       // one output array each, or ...
    u64 *const *pw; uint8_t *const *pl; u64 *const *pc;               // ... device tables of per-partition destinations
};

__global__ void __launch_bounds__(kThreads) export_regions_kernel(TableView t, const int64_t *region_base, int log2_parts, u32 first_region,
                                                                  ExportDst d) {
    constexpr int kTile = kThreads * kExportItems;
    __shared__ u32 s_warp[kThreads / 32];
    // The tile's tuples are staged in output order, shifted so that a staged element and its destination have the same
    // 16-byte phase: the bulk of a tile then leaves as coalesced 16-byte stores (what NVLink peer stores need).
    __shared__ __align__(16) u64 s_word[kTile + 2], s_count[kTile + 2];
    __shared__ __align__(16) uint8_t s_len[kTile + 16];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int log2_regions = t.log2_cap - t.log2_region;
    const u32 nregions = 1u << log2_regions;
    const u32 R = 1u << t.log2_region;
    for (u32 ri = blockIdx.x; ri < nregions; ri += gridDim.x) {
        const u32 region = (ri + first_region) & (nregions - 1);       // the walk starts at first_region and wraps around
        const int64_t base = region_base[region];
        if (region_base[region + 1] == base) continue;                  // empty region (CTA-uniform)
        u64 *W; uint8_t *L; u64 *C;
        int64_t o = base;
        if (d.pw != nullptr) {
            const u32 part = region >> (log2_regions - log2_parts);
            o = base - region_base[(size_t)part << (log2_regions - log2_parts)];
            W = d.pw[part]; L = d.pl[part]; C = d.pc[part];
        } else {
            W = d.words; L = d.lens; C = d.counts;
        }
        const u64 slot0 = (u64)region << t.log2_region;
        for (u32 tile0 = 0; tile0 < R; tile0 += kTile) {
            u64 key[kExportItems], cnt[kExportItems];
            u32 n_used = 0;
#pragma unroll
            for (int k = 0; k < kExportItems; k++) {
                const u32 s = tile0 + k * kThreads + threadIdx.x;
                key[k] = 0;
                if (s < R) {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(t.slots)[slot0 + s];
                    key[k] = v.x; cnt[k] = v.y;
                }
                n_used += key[k] != 0;
            }
            u32 incl = n_used;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, dd); if ((int)lane >= dd) incl += x; }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            u32 before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) { const u32 x = s_warp[w]; if (w < (int)warp) before += x; total += x; }
            // element i of the tile goes to index o + i: stage it at i + phase so that both share alignment
            const u32 ph8 = (u32)((reinterpret_cast<uintptr_t>(W + o) >> 3) & 1);        // u64 arrays: phase within 16 bytes
            const u32 phc = (u32)((reinterpret_cast<uintptr_t>(C + o) >> 3) & 1);
            const u32 ph1 = (u32)(reinterpret_cast<uintptr_t>(L + o) & 15);              // byte array
            u32 at = before + incl - n_used;
#pragma unroll
            for (int k = 0; k < kExportItems; k++) {
                if (key[k] == 0) continue;
                const u64 s = slot0 + tile0 + k * kThreads + threadIdx.x;
                u32 len;
                const u64 h2 = slot64_h2(t, s, key[k], len);
                s_word[at + ph8] = unhash64(rotr64(h2, t.rot));
                s_count[at + phc] = cnt[k];
                s_len[at + ph1] = (uint8_t)len;
                if (d.first_idx) d.first_idx[o + at] = (int64_t)(t.first_idx ? t.first_idx[s] : kNoIndex);
                ++at;
            }
            __syncthreads();
            {   // u64 arrays: staged span [ph, ph + total); whole 16-byte pairs in the middle, single elements at the ends
                for (int a = 0; a < 2; a++) {
                    const u32 ph = a ? phc : ph8;
                    u64 *dst = (a ? C : W) + o - ph;                    // staged index j <-> dst[j]; dst is 16-byte aligned
                    const u64 *src = a ? s_count : s_word;
                    const u32 lo = ph, hi = ph + total;
                    const u32 mid_lo = (lo + 1) & ~1u, mid_hi = hi & ~1u;
                    if (threadIdx.x == 0 && lo < mid_lo && lo < hi) dst[lo] = src[lo];
                    if (threadIdx.x == 1 && mid_hi < hi && mid_hi >= mid_lo) dst[mid_hi] = src[mid_hi];
                    for (u32 j = mid_lo + 2 * threadIdx.x; j + 2 <= mid_hi; j += 2 * kThreads)
                        *reinterpret_cast<ulonglong2 *>(dst + j) = *reinterpret_cast<const ulonglong2 *>(src + j);
                }
                uint8_t *Lb = L + o - ph1;                              // 16-byte aligned
                const u32 blo = ph1, bhi = ph1 + total;
                const u32 bmid_lo = min((blo + 15) & ~15u, bhi), bmid_hi = max(bhi & ~15u, bmid_lo);
                for (u32 j = blo + threadIdx.x; j < bmid_lo; j += kThreads) Lb[j] = s_len[j];
                for (u32 j = bmid_hi + threadIdx.x; j < bhi; j += kThreads) Lb[j] = s_len[j];
                for (u32 j = bmid_lo + 16 * threadIdx.x; j + 16 <= bmid_hi; j += 16 * kThreads)
                    *reinterpret_cast<uint4 *>(Lb + j) = *reinterpret_cast<const uint4 *>(s_len + j);
            }
            o += total;
            __syncthreads();
        }
    }
}

__global__ void export_part_counts_kernel(const int64_t *region_base, int log2_regions, int log2_parts, int64_t *part_counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < (1 << log2_parts)) {
        const int sh = log2_regions - log2_parts;
        part_counts[p] = region_base[(size_t)(p + 1) << sh] - region_base[(size_t)p << sh];
    }
}

// region_base = exclusive scan of the region counts (enqueued on the context's stream)
static int scan_regions(ssq_counter *c) {
    const int64_t nregions = (int64_t)1 << (c->log2_cap - region_bits_for(c->log2_cap));
    if (!c->region_base) SSQ_CUDA(cudaMalloc(&c->region_base, sizeof(int64_t) * (size_t)(nregions + 1)));
    return scan_u32_counts(c->ctx, c->region_count, nregions, c->region_base);
}

static bool region_export_ok(const ssq_counter *c, int log2_parts) {
    return c->klass == SSQ_CLASS_64 && log2_parts <= c->log2_cap - region_bits_for(c->log2_cap);
}

static int launch_export_regions(ssq_counter *c, int log2_parts, int first_part, const ExportDst &d) {
    const int64_t nregions = (int64_t)1 << (c->log2_cap - region_bits_for(c->log2_cap));
    const int grid = grid_for(c->ctx, nregions, 8);
    const u32 first_region = (u32)(((int64_t)first_part * nregions) >> log2_parts);
    export_regions_kernel<<<grid, kThreads, 0, c->ctx->stream>>>(view_of(c), c->region_base, log2_parts, first_region, d);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

__global__ void fill_u64_kernel(u64 *p, u64 v, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = v;
}

static size_t slot_bytes(int klass) { return klass == SSQ_CLASS_64 ? 16 : 32; }

static TableView view_of(const ssq_counter *c) {
    TableView t;
    t.slots = (u64 *)c->slots;
    t.first_idx = c->first_idx;
    t.size = c->d_size;
    t.rep = c->ctx->d_report;
    t.log2_cap = c->log2_cap;
    t.rot = c->hash_rot;
    t.log2_region = region_bits_for(c->log2_cap);
    t.region_count = c->region_count;
    return t;
}

static int log2_cap_for(int64_t expected_unique) {
    int l = kMinLog2Cap;
    while (l < 40 && (1ll << l) < 2 * expected_unique) l++;
    return l;
}

// Grow to 2^new_log2 slots, rehashing on the device.  The stream is idle on entry and on exit.
static int grow(ssq_counter *c, int new_log2) {
    ssq_ctx *ctx = c->ctx;
    c->mod_seq++;
    cudaStream_t st = ctx->stream;
    const size_t nslots = (size_t)1 << new_log2;
    void *nslots_p = nullptr;
    u64 *nfirst = nullptr;
    SSQ_CUDA(cudaMalloc(&nslots_p, nslots * slot_bytes(c->klass)));
    SSQ_CUDA(cudaMemsetAsync(nslots_p, 0, nslots * slot_bytes(c->klass), st));
    if (c->first_idx) {
        SSQ_CUDA(cudaMalloc(&nfirst, nslots * sizeof(u64)));
        SSQ_CUDA(cudaMemsetAsync(nfirst, 0xFF, nslots * sizeof(u64), st));
    }
    u32 *nregion = nullptr;
    const size_t nregions = nslots >> region_bits_for(new_log2);
    if (c->klass == SSQ_CLASS_64) {
        SSQ_CUDA(cudaMalloc(&nregion, nregions * sizeof(u32)));
        SSQ_CUDA(cudaMemsetAsync(nregion, 0, nregions * sizeof(u32), st));
    }
    TableView src = view_of(c);
    TableView dst = src;
    dst.slots = (u64 *)nslots_p;
    dst.first_idx = nfirst;
    dst.log2_cap = new_log2;
    dst.log2_region = region_bits_for(new_log2);
    dst.region_count = nregion;
    SSQ_CUDA(cudaMemsetAsync(c->d_size, 0, sizeof(u64), st));
    int grid = grid_for(ctx, ((int64_t)1 << c->log2_cap) / kThreads, 8);
    if (c->klass == SSQ_CLASS_64) rehash_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, st>>>(src, dst);
    else rehash_kernel<SSQ_CLASS_192><<<grid, kThreads, 0, st>>>(src, dst);
    SSQ_LAUNCH_CHECK();
    SSQ_CUDA(cudaStreamSynchronize(st));
    SSQ_CUDA(cudaFree(c->slots));
    if (c->first_idx) SSQ_CUDA(cudaFree(c->first_idx));
    if (c->region_count) SSQ_CUDA(cudaFree(c->region_count));
    if (c->region_base) SSQ_CUDA(cudaFree(c->region_base));
    c->region_base = nullptr;
    c->slots = nslots_p;
    c->first_idx = nfirst;
    c->region_count = nregion;
    c->log2_cap = new_log2;
    return SSQ_OK;
}

// Runs launch(start, count, stop_flag) over [0, n) in gated sub-batches; grows and resumes when stopped.
template <class Launch>
static int run_gated(ssq_counter *c, int64_t n, Launch launch) {
    ssq_ctx *ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    int64_t pos = 0;
    while (pos < n) {
        const int64_t cap = (int64_t)1 << c->log2_cap;
        const int64_t sub = cap / 8;
        const u64 limit = (u64)(cap - cap / 4);
        SSQ_CUDA(cudaMemsetAsync(c->d_gate, 0, 2 * sizeof(u64), st));
        int64_t k = 0;
        for (int64_t p = pos; p < n; p += sub, k++) {
            int64_t cnt = n - p < sub ? n - p : sub;
            gate_kernel<<<1, 1, 0, st>>>(c->d_gate, c->d_size, (u64)cnt, limit, (u64)k);
            SSQ_LAUNCH_CHECK();
            int rc = launch(p, cnt, c->d_gate);
            if (rc) return rc;
        }
        SSQ_CUDA(cudaMemcpyAsync(c->h_gate, c->d_gate, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        SSQ_CUDA(cudaMemcpyAsync(c->h_size, c->d_size, sizeof(u64), cudaMemcpyDeviceToHost, st));
        SSQ_CUDA(cudaStreamSynchronize(st));
        c->known_size = (int64_t)*c->h_size;
        if (c->h_gate[0] == 0) break;
        pos += (int64_t)c->h_gate[1] * sub;
        // room for what is there plus the rest of this batch, at least x4
        int want = log2_cap_for((int64_t)*c->h_size + (n - pos < 4 * cap ? n - pos : 4 * cap));
        if (want < c->log2_cap + 2) want = c->log2_cap + 2;
        int rc = grow(c, want);
        if (rc) return rc;
    }
    return SSQ_OK;
}

// Tables up to this size are assumed to stay resident in the 126 MB L2 under random inserts.
constexpr size_t kDirectTableBytes = (size_t)48 << 20;

static bool use_deferred(const ssq_counter *c, int64_t n) {
    if (c->expected_unique <= 0) return false;
    static const bool off = getenv("SSQ_NO_DEFERRED") != nullptr;      // development: direct inserts whatever the table size
    if (off) return false;
    const size_t table_bytes = ((size_t)1 << c->log2_cap) * slot_bytes(c->klass);
    // worthwhile when the table cannot live in L2 and the pass brings at least ~1 key per 4 slots
    return table_bytes > kDirectTableBytes && n >= ((int64_t)1 << c->log2_cap) / 4;
}

int scatter_grid(ssq_ctx *ctx, int klass, const uint8_t *ascii, int64_t lo, int64_t hi, int64_t n);   // ssq_pack.cu: grid of the fused pack+scatter launch

// development tunables (read once per process)
static int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    int v = e ? atoi(e) : dflt;
    return v < lo ? lo : (v > hi ? hi : v);
}

// Size the level-1 partition buffers for a pass of n keys scattered by `grid` persistent CTAs.
static int prepare_parts(ssq_counter *c, int64_t n, int grid, PartView *pv) {
    ssq_ctx *ctx = c->ctx;
    const int rw = c->klass == SSQ_CLASS_64 ? 1 : 4;        // 64-bit words per record
    const int line = kLineKeys / rw;                        // records per 128-byte line
    const int kParts = c->klass == SSQ_CLASS_64 ? ssq::kParts : kParts192;   // level-1 partitions of this class
    // a CTA sees ~n/grid keys, 1/kParts of them per partition: mean + 6 % + slack (overflow is handled, not fatal)
    int64_t per = n / ((int64_t)grid * kParts);
    per = per + per * env_int("SSQ_SEG_SLACK_PCT", 6, 0, 100) / 100 + env_int("SSQ_SEG_SLACK_ABS", 64, 0, 1 << 20);   // tests shrink the slack
    per = (per + line - 1) & ~(int64_t)(line - 1);
    if (per < line) per = line;
    // ShortSeq192 overflow segments: 2 % of a CTA's reads + slack
    const int64_t ovf_cap = c->klass == SSQ_CLASS_64 ? 0 : n / grid / 50 + 256;
    const int64_t main_entries = per * grid * kParts * rw;
    const int64_t need = main_entries + ovf_cap * grid * rw;
    if (need > c->part_cap) {
        SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (c->part_keys) SSQ_CUDA(cudaFree(c->part_keys));
        c->part_keys = nullptr;
        c->part_cap = 0;
        SSQ_CUDA(cudaMalloc(&c->part_keys, sizeof(u64) * (size_t)need));
        c->part_cap = need;
    }
    if (grid > c->part_ctas) {
        SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (c->part_cursor) SSQ_CUDA(cudaFree(c->part_cursor));
        c->part_cursor = nullptr;
        c->part_ctas = 0;
        SSQ_CUDA(cudaMalloc(&c->part_cursor, sizeof(u32) * (size_t)grid * (ssq::kParts + 1)));
        c->part_ctas = grid;
    }
    pv->keys = c->part_keys;
    pv->seg_count = c->part_cursor;
    pv->seg_cap = (u32)per;
    pv->num_ctas = (u32)grid;
    pv->ovf = c->part_keys + main_entries;
    pv->ovf_count = c->part_cursor + (size_t)grid * kParts;
    pv->ovf_cap = (u32)ovf_cap;
    // a tile brings 512 keys = 2 per partition on average and a ring holds 32: flushing every 4th tile keeps ring
    // overflow (handled, but slow) below one key in a thousand tiles
    // (ShortSeq192: a 256-read tile brings 2 records per partition and a ring holds 16: flush every second tile)
    pv->flush_every = c->klass == SSQ_CLASS_64 ? (u32)env_int("SSQ_FLUSH_EVERY", 4, 1, 8) : (u32)env_int("SSQ_FLUSH_EVERY192", kRingKeys192 >= 64 ? 2 : 1, 1, 8);
    return SSQ_OK;
}

// Regions of up to 2^14 slots (192 KB of keys + count deltas: one CTA per SM) are counted in shared memory.
static bool regions_fit_smem(const ssq_counter *c) {
    const int lr = region_bits_for(c->log2_cap);
    return c->log2_cap >= 22 && lr <= 14;       // >= 2^22 slots: at most 64 level-2 partitions per region
}

static int region_slices() {
    static int v = 0;
    if (v == 0) v = env_int("SSQ_REGION_SLICES", 6, 1, 6);
    return v;
}

// Size the level-2 (per region) buffers for n keys.
static int prepare_regions(ssq_counter *c, int64_t n, RegionParts *rp) {
    ssq_ctx *ctx = c->ctx;
    const int lr = region_bits_for(c->log2_cap);
    const int qbits = c->log2_cap - 8 - lr;
    const int slices = region_slices();
    if ((slices << (8 - qbits)) > kMaxRegionSegs) { set_error("too many level-2 segments per region"); return SSQ_ERR_ARG; }
    int64_t per = n / ((int64_t)kParts * slices * kParts);
    // A level-2 partition holds ~U / 65536 DISTINCT keys, each with all of its copies: with few distinct keys the
    // partition sizes are lumpy (relative deviation ~ sqrt(65536 / U)), and a segment that overflows sends its keys down
    // the slow direct-insert path (1e9 reads of 1e6 keys: 8.3 ms instead of 4.2).  Slack = 4 deviations, at least 6 %.
    int slack_pct = env_int("SSQ_SEG_SLACK_PCT", 6, 0, 100);
    if (getenv("SSQ_SEG_SLACK_PCT") == nullptr && c->expected_unique > 0) {
        const double dev = 400.0 * sqrt(65536.0 / (double)c->expected_unique);
        if (dev > slack_pct) slack_pct = dev > 300.0 ? 300 : (int)dev;
    }
    per = per + per * slack_pct / 100 + env_int("SSQ_SEG_SLACK_ABS", 64, 0, 1 << 20);
    per = (per + kLineKeys - 1) & ~(int64_t)(kLineKeys - 1);
    if (per < kLineKeys) per = kLineKeys;
    const int64_t nseg = (int64_t)kParts * slices * kParts;
    // the count kernel addresses this buffer with 32-bit entry offsets: less slack rather than more than 2^32 entries
    if (per * nseg >= 4000000000ll) per = (4000000000ll / nseg) & ~(int64_t)(kLineKeys - 1);
    const int64_t need = per * nseg;
    if (need > c->region_cap) {
        SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (c->region_keys) SSQ_CUDA(cudaFree(c->region_keys));
        c->region_keys = nullptr;
        c->region_cap = 0;
        SSQ_CUDA(cudaMalloc(&c->region_keys, sizeof(u64) * (size_t)need));
        c->region_cap = need;
    }
    if (nseg > c->region_segs) {
        SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (c->region_cursor) SSQ_CUDA(cudaFree(c->region_cursor));
        c->region_cursor = nullptr;
        c->region_segs = 0;
        SSQ_CUDA(cudaMalloc(&c->region_cursor, sizeof(u32) * (size_t)nseg));
        c->region_segs = nseg;
    }
    rp->keys = c->region_keys;
    rp->seg_count = c->region_cursor;
    rp->seg_cap = (u32)per;
    rp->slices = (u32)slices;
    rp->qbits = (u32)qbits;
    return SSQ_OK;
}

static int set_max_smem(const void *kernel, size_t bytes) {
    SSQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    SSQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return SSQ_OK;
}

template <int THREADS>
static int launch_count_regions2(ssq_counter *c, const TableView &t, const RegionParts &rp, unsigned nregions, size_t bytes, bool stream) {
    cudaStream_t st = c->ctx->stream;
    StreamOut so{};
    int rc;
    if (stream) {
        const ssq_stream_out *x = c->stream_out;
        const int par = (int)((*x->epoch + 1) & 1);
        so.dst = (ulonglong2 *const *)x->d_dst[par];
        so.dst_cnt = (u32 *const *)x->d_cnt[par];
        so.log2_parts = x->log2_parts;
        so.rot_owner = x->rot_owner;
        so.first_region = (u32)((unsigned)((x->rank + 1) & ((1 << x->log2_parts) - 1)) * (nregions >> x->log2_parts));
        rc = set_max_smem((const void *)count_regions2_kernel<THREADS, true>, bytes);
        if (rc) return rc;
        count_regions2_kernel<THREADS, true><<<nregions, THREADS, bytes, st>>>(t, rp, so);
        c->streamed_seq = c->mod_seq;
        c->streamed_parity = par;
    } else {
        rc = set_max_smem((const void *)count_regions2_kernel<THREADS, false>, bytes);
        if (rc) return rc;
        count_regions2_kernel<THREADS, false><<<nregions, THREADS, bytes, st>>>(t, rp, so);
    }
    return SSQ_OK;
}

// Phase 2 of the deferred path: count the keys of the level-1 partitions `pv` into the table.
// ev_mid (may be null) is recorded between the level-2 scatter and the region count.
static int launch_count_parts(ssq_counter *c, int64_t n, const PartView &pv, cudaEvent_t ev_mid) {
    ssq_ctx *ctx = c->ctx;
    const TableView t = view_of(c);
    // the count kernel addresses the level-2 buffer with 32-bit entry offsets
    const bool small_enough = (double)n * 1.07 + 64.0 * kParts * kParts * region_slices() < 4.0e9;
    if (!regions_fit_smem(c) || pv.num_ctas > (u32)kMaxStreamSegs || !small_enough || env_int("SSQ_FORCE_L2_COUNT", 0, 0, 1)) {
        if (ev_mid) SSQ_CUDA(cudaEventRecord(ev_mid, ctx->stream));
        count_parts_kernel<<<kParts * pv.num_ctas, kThreads, 0, ctx->stream>>>(t, pv);
        SSQ_LAUNCH_CHECK();
        return SSQ_OK;
    }
    RegionParts rp;
    int rc = prepare_regions(c, n, &rp);
    if (rc) return rc;
    if (env_int("SSQ_SCATTER_THREADS", 384, 256, 384) == 384) {
        rc = set_max_smem((const void *)region_scatter_kernel<384>, kStagerRingBytes);
        if (rc) return rc;
        region_scatter_kernel<384><<<kParts * rp.slices, 384, kStagerRingBytes, ctx->stream>>>(t, pv, rp, (u32)env_int("SSQ_SCATTER_FLUSH_KEYS", 768, 1, 4096));
    } else {
        rc = set_max_smem((const void *)region_scatter_kernel<256>, kStagerRingBytes);
        if (rc) return rc;
        region_scatter_kernel<256><<<kParts * rp.slices, 256, kStagerRingBytes, ctx->stream>>>(t, pv, rp, (u32)env_int("SSQ_SCATTER_FLUSH_KEYS", 512, 1, 4096));
    }
    SSQ_LAUNCH_CHECK();
    if (ev_mid) SSQ_CUDA(cudaEventRecord(ev_mid, ctx->stream));
    const unsigned nregions = 1u << (t.log2_cap - t.log2_region);
    const unsigned nseg = rp.slices << (8 - rp.qbits);
    const int cthreads = env_int("SSQ_COUNT_THREADS", t.log2_region >= 14 ? 512 : 384, 256, 512);   // one CTA per SM at 2^14-slot regions: make it a big one
    // attached to a communicator whose receive buffers match this table: the count kernel also sends every region to its owner
    const bool stream = c->stream_out != nullptr && c->stream_out->log2_cap == c->log2_cap &&
                        t.log2_cap - t.log2_region >= 6 && t.log2_cap - t.log2_region >= c->stream_out->log2_parts;
    if (env_int("SSQ_COUNT_V2", 1, 0, 1)) {
        if (cthreads == 256) {
            const size_t bytes = count_regions2_smem<256>(t.log2_region, nseg);
            rc = launch_count_regions2<256>(c, t, rp, nregions, bytes, stream);
            if (rc) return rc;
        } else if (cthreads == 512) {
            const size_t bytes = count_regions2_smem<512>(t.log2_region, nseg);
            rc = launch_count_regions2<512>(c, t, rp, nregions, bytes, stream);
            if (rc) return rc;
        } else {
            const size_t bytes = count_regions2_smem<384>(t.log2_region, nseg);
            rc = launch_count_regions2<384>(c, t, rp, nregions, bytes, stream);
            if (rc) return rc;
        }
    } else if (cthreads == 256) {
        const size_t bytes = count_regions_smem<256>(t.log2_region, nseg);
        rc = set_max_smem((const void *)count_regions_kernel<256>, bytes);
        if (rc) return rc;
        count_regions_kernel<256><<<nregions, 256, bytes, ctx->stream>>>(t, rp);
    } else if (cthreads == 512) {
        const size_t bytes = count_regions_smem<512>(t.log2_region, nseg);
        rc = set_max_smem((const void *)count_regions_kernel<512>, bytes);
        if (rc) return rc;
        count_regions_kernel<512><<<nregions, 512, bytes, ctx->stream>>>(t, rp);
    } else {
        const size_t bytes = count_regions_smem<384>(t.log2_region, nseg);
        rc = set_max_smem((const void *)count_regions_kernel<384>, bytes);
        if (rc) return rc;
        count_regions_kernel<384><<<nregions, 384, bytes, ctx->stream>>>(t, rp);
    }
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

// After a single-pass (bounded) insert: wait, read the size, grow once the table is past 60 % load.
static int finish_pass(ssq_counter *c) {
    ssq_ctx *ctx = c->ctx;
    SSQ_CUDA(cudaMemcpyAsync(c->h_size, c->d_size, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
    c->known_size = (int64_t)*c->h_size;
    const int64_t cap = (int64_t)1 << c->log2_cap;
    if ((int64_t)*c->h_size > cap - cap / 4 - cap / 8 - cap / 40) return grow(c, c->log2_cap + 1);
    return SSQ_OK;
}

// Before a bounded single pass of n keys: make room.  The pass can create at most n keys and, if the caller's bound
// holds, at most expected_unique - size of them; when that many would load the table beyond 75 % the table grows FIRST
// (a pass cannot grow the table while it runs, and a full region drops keys).  A table that already holds MORE keys
// than the bound has proven the bound wrong: the counter switches to the conservative gated mode for good
// (expected_unique = 0), which grows whenever it has to.  A bound that holds never triggers either.
static int make_room(ssq_counter *c, int64_t n) {
    if (c->known_size > c->expected_unique) { c->expected_unique = 0; return SSQ_OK; }
    const int64_t cap = (int64_t)1 << c->log2_cap;
    int64_t ub = c->expected_unique - c->known_size;
    if (n < ub) ub = n;
    if (c->known_size + ub <= cap - cap / 4) return SSQ_OK;
    SSQ_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return grow(c, log2_cap_for(c->known_size + ub));
}

// Fused pack+count of reads whose bytes are ascii[lo, hi) (ascii may be a virtual base pointer);
// index_base is added to the read indices that data errors are reported with.
int pack_count_impl(ssq_counter *c, const uint8_t *ascii, int64_t lo, int64_t hi, const int64_t *offsets, int64_t n,
                    int64_t index_base, u64 *words, uint8_t *lens) {
    ssq_ctx *ctx = c->ctx;
    const int W = c->klass == SSQ_CLASS_64 ? 1 : 3;
    int rc = SSQ_OK;
    c->mod_seq++;
    if (c->expected_unique > 0 && (rc = make_room(c, n)) != SSQ_OK) return rc;
    if (c->expected_unique <= 0) {
        c->last_pass_phases = 0;
        return run_gated(c, n, [&](int64_t p, int64_t cnt, const u64 *stop) -> int {
            return launch_pack_count(ctx, c->klass, false, ascii, lo, hi, offsets + p, cnt, index_base + p,
                                     words + (size_t)p * W, lens + p, view_of(c), PartView{}, stop);
        });
    }
    // A table small enough for L2 takes direct inserts, but at ~1 read per 24 ps the L2 atomic units are the limit
    // (1e9 reads into 2^21 slots: 23.8 ms against 20 ms for the deferred path): a pass this large gets at least the
    // smallest table the deferred path works with.
    if (c->klass == SSQ_CLASS_64 && n >= ((int64_t)1 << 27) && c->log2_cap < 22) {
        SSQ_CUDA(cudaStreamSynchronize(ctx->stream));
        rc = grow(c, 22);
        if (rc) return rc;
    }
    SSQ_CUDA(cudaEventRecord(c->ev[0], ctx->stream));
    if (use_deferred(c, n)) {
        c->last_pass_phases = 2;
        PartView pv;
        int sgrid = scatter_grid(ctx, c->klass, ascii, lo, hi, n);
        if ((int64_t)sgrid * 65536 > n) sgrid = (int)(n / 65536 > 0 ? n / 65536 : 1);   // keep the segments from being mostly padding
        rc = prepare_parts(c, n, sgrid, &pv);
        if (rc) return rc;
        rc = launch_pack_count(ctx, c->klass, true, ascii, lo, hi, offsets, n, index_base, words, lens, view_of(c), pv, nullptr);
        if (rc) return rc;
        SSQ_CUDA(cudaEventRecord(c->ev[1], ctx->stream));
        if (c->klass == SSQ_CLASS_64) {
            rc = launch_count_parts(c, n, pv, c->ev[2]);
        } else {
            SSQ_CUDA(cudaEventRecord(c->ev[2], ctx->stream));
            count_parts192_kernel<<<(kParts192 + 1) * pv.num_ctas, kThreads, 0, ctx->stream>>>(view_of(c), pv);
            SSQ_LAUNCH_CHECK();
        }
    } else {
        c->last_pass_phases = 1;
        rc = launch_pack_count(ctx, c->klass, false, ascii, lo, hi, offsets, n, index_base, words, lens, view_of(c),
                               PartView{}, nullptr);
        SSQ_CUDA(cudaEventRecord(c->ev[1], ctx->stream));
        SSQ_CUDA(cudaEventRecord(c->ev[2], ctx->stream));
    }
    if (rc) return rc;
    SSQ_CUDA(cudaEventRecord(c->ev[3], ctx->stream));
    return finish_pass(c);
}

}  // namespace ssq

using namespace ssq;

extern "C" {

int ssq_counter_create(ssq_ctx *ctx, int klass, int64_t expected_unique, int hash_rot, ssq_counter **out) {
    SSQ_ARG(ctx != nullptr && out != nullptr, "NULL argument");
    SSQ_ARG(klass == SSQ_CLASS_64 || klass == SSQ_CLASS_192, "klass must be SSQ_CLASS_64 or SSQ_CLASS_192");
    SSQ_ARG(hash_rot >= 0 && hash_rot <= 56, "hash_rot out of range");
    SSQ_ARG(expected_unique >= 0, "expected_unique < 0");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    ssq_counter *c = new ssq_counter();
    c->ctx = ctx;
    c->klass = klass;
    c->hash_rot = hash_rot;
    c->log2_cap = log2_cap_for(expected_unique);
    c->slots = nullptr;
    c->first_idx = nullptr;
    c->region_count = nullptr;
    c->region_base = nullptr;
    c->expected_unique = expected_unique;
    c->known_size = 0;
    c->part_keys = nullptr;
    c->part_cursor = nullptr;
    c->part_cap = 0;
    c->part_ctas = 0;
    c->region_keys = nullptr;
    c->region_cursor = nullptr;
    c->region_cap = 0;
    c->region_segs = 0;
    c->last_pass_phases = 0;
    for (int i = 0; i < 4; i++) SSQ_CUDA(cudaEventCreate(&c->ev[i]));
    const size_t bytes = ((size_t)1 << c->log2_cap) * slot_bytes(klass);
    SSQ_CUDA(cudaMalloc(&c->slots, bytes));
    SSQ_CUDA(cudaMalloc(&c->d_size, 4 * sizeof(u64)));
    c->d_gate = c->d_size + 1;
    SSQ_CUDA(cudaHostAlloc(&c->h_size, 4 * sizeof(u64), cudaHostAllocDefault));
    c->h_gate = c->h_size + 1;
    SSQ_CUDA(cudaMemsetAsync(c->slots, 0, bytes, ctx->stream));
    SSQ_CUDA(cudaMemsetAsync(c->d_size, 0, 4 * sizeof(u64), ctx->stream));
    if (klass == SSQ_CLASS_64) {
        const size_t nregions = ((size_t)1 << c->log2_cap) >> region_bits_for(c->log2_cap);
        SSQ_CUDA(cudaMalloc(&c->region_count, nregions * sizeof(u32)));
        SSQ_CUDA(cudaMemsetAsync(c->region_count, 0, nregions * sizeof(u32), ctx->stream));
    }
    *out = c;
    return SSQ_OK;
}

int ssq_counter_destroy(ssq_counter *c) {
    if (!c) return SSQ_OK;
    DeviceGuard g(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->stream_out && c->stream_out->attached_slot) *c->stream_out->attached_slot = nullptr;   // the communicator forgets this counter
    cudaFree(c->slots);
    cudaFree(c->first_idx);
    cudaFree(c->region_count);
    cudaFree(c->region_base);
    cudaFree(c->part_keys);
    cudaFree(c->part_cursor);
    cudaFree(c->region_keys);
    cudaFree(c->region_cursor);
    for (int i = 0; i < 4; i++) cudaEventDestroy(c->ev[i]);
    cudaFree(c->d_size);
    cudaFreeHost(c->h_size);
    delete c;
    return SSQ_OK;
}

int ssq_counter_clear(ssq_counter *c) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    DeviceGuard g(c->ctx->device);
    const size_t nslots = (size_t)1 << c->log2_cap;
    c->mod_seq++;
    SSQ_CUDA(cudaMemsetAsync(c->slots, 0, nslots * slot_bytes(c->klass), c->ctx->stream));
    SSQ_CUDA(cudaMemsetAsync(c->d_size, 0, sizeof(u64), c->ctx->stream));
    c->known_size = 0;
    if (c->first_idx) SSQ_CUDA(cudaMemsetAsync(c->first_idx, 0xFF, nslots * sizeof(u64), c->ctx->stream));
    if (c->region_count)
        SSQ_CUDA(cudaMemsetAsync(c->region_count, 0, (nslots >> region_bits_for(c->log2_cap)) * sizeof(u32), c->ctx->stream));
    return SSQ_OK;
}

static int insert_common(ssq_counter *c, const uint64_t *words, const uint8_t *lens, const uint64_t *counts, int64_t n) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    SSQ_ARG(n >= 0 && (n == 0 || (words != nullptr && lens != nullptr)), "bad batch");
    if (n == 0) return SSQ_OK;
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    c->mod_seq++;
    const int W = c->klass == SSQ_CLASS_64 ? 1 : 3;
    if (c->expected_unique > 0) {
        int rc0 = make_room(c, n);
        if (rc0) return rc0;
    }
    if (c->expected_unique > 0) {
        int rc = SSQ_OK;
        if (counts == nullptr && c->klass == SSQ_CLASS_64 && use_deferred(c, n)) {
            PartView pv;
            int grid = grid_for(ctx, (n + 65535) / 65536, 3);
            rc = prepare_parts(c, n, grid, &pv);
            if (rc) return rc;
            rc = set_max_smem((const void *)scatter_packed_kernel, kStagerRingBytes);
            if (rc) return rc;
            scatter_packed_kernel<<<grid, kThreads, kStagerRingBytes, ctx->stream>>>(view_of(c), pv, (const u64 *)words, lens, n, 0);
            SSQ_LAUNCH_CHECK();
            rc = launch_count_parts(c, n, pv, nullptr);
        } else {
            int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
            if (c->klass == SSQ_CLASS_64)
                insert_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, ctx->stream>>>(view_of(c), (const u64 *)words, lens,
                                                                                (const u64 *)counts, n, 0, nullptr);
            else
                insert192_pre_kernel<<<grid_for(ctx, (n + kThreads * kIns192 - 1) / (kThreads * kIns192), 8), kThreads, 0, ctx->stream>>>(
                    view_of(c), (const u64 *)words, lens, (const u64 *)counts, n, 0, nullptr);
            SSQ_LAUNCH_CHECK();
        }
        if (rc) return rc;
        return finish_pass(c);
    }
    return run_gated(c, n, [&](int64_t p, int64_t cnt, const u64 *stop) -> int {
        TableView t = view_of(c);
        int grid = grid_for(ctx, (cnt + kThreads - 1) / kThreads, 8);
        const u64 *w = (const u64 *)words + (size_t)p * W;
        const u64 *cn = counts ? (const u64 *)counts + p : nullptr;
        if (c->klass == SSQ_CLASS_64)
            insert_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, ctx->stream>>>(t, w, lens + p, cn, cnt, p, stop);
        else
            insert192_pre_kernel<<<grid_for(ctx, (cnt + kThreads * kIns192 - 1) / (kThreads * kIns192), 8), kThreads, 0, ctx->stream>>>(t, w, lens + p, cn, cnt, p, stop);
        SSQ_LAUNCH_CHECK();
        return SSQ_OK;
    });
}

int ssq_counter_insert(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n) {
    return insert_common(c, words, lens, nullptr, n);
}

int ssq_counter_merge(ssq_counter *c, const uint64_t *words, const uint8_t *lens, const uint64_t *counts, int64_t n) {
    SSQ_ARG(n == 0 || counts != nullptr, "counts is NULL");
    return insert_common(c, words, lens, counts, n);
}

int ssq_counter_regions(ssq_counter *c, int64_t *n_regions) {
    SSQ_ARG(c != nullptr && n_regions != nullptr, "NULL argument");
    *n_regions = c->klass == SSQ_CLASS_64 ? (int64_t)1 << (c->log2_cap - region_bits_for(c->log2_cap)) : 0;
    return SSQ_OK;
}

int ssq_counter_merge_regions(ssq_counter *c, const uint64_t *words, const uint8_t *lens, const uint64_t *counts,
                              const int64_t *block_counts, const int64_t *block_regions, int n_blocks,
                              const int64_t *region_bases, int64_t rb_stride) {
    return counter_merge_regions_impl(c, (const u64 *)words, lens, (const u64 *)counts, block_counts, block_regions, n_blocks, region_bases,
                                      rb_stride, nullptr, 0);
}

int ssq_counter_export_region_bases(ssq_counter *c, int n_parts, int64_t *const *dst) {
    SSQ_ARG(c != nullptr && dst != nullptr, "NULL argument");
    SSQ_ARG(n_parts >= 1 && n_parts <= kMaxParts && (n_parts & (n_parts - 1)) == 0, "n_parts must be a power of two <= 256");
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    if (!region_export_ok(c, log2_parts)) { set_error("ssq_counter_export_region_bases needs a ShortSeq64 counter with at least n_parts regions"); return SSQ_ERR_ARG; }
    DeviceGuard g(c->ctx->device);
    int rc = scan_regions(c);
    if (rc) return rc;
    const int log2_regions = c->log2_cap - region_bits_for(c->log2_cap);
    export_region_bases_kernel<<<grid_for(c->ctx, (((int64_t)1 << log2_regions) + n_parts + 255) / 256, 4), 256, 0, c->ctx->stream>>>(
        c->region_base, log2_regions, log2_parts, dst);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_counter_pack_count(ssq_counter *c, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *offsets,
                           int64_t n, uint64_t *words, uint8_t *lens) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    SSQ_ARG(n >= 0 && ascii_bytes >= 0, "negative size");
    SSQ_ARG(n == 0 || (offsets != nullptr && words != nullptr && lens != nullptr), "NULL buffer");
    if (n == 0) return SSQ_OK;
    DeviceGuard g(c->ctx->device);
    return pack_count_impl(c, ascii, 0, ascii_bytes, offsets, n, 0, (u64 *)words, lens);
}

int ssq_counter_first_index(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n, int64_t base_index) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    SSQ_ARG(n >= 0 && (n == 0 || (words != nullptr && lens != nullptr)), "bad batch");
    return counter_first_index_indexed(c, (const u64 *)words, lens, n, nullptr, base_index);
}

}  // extern "C"

namespace ssq {
int counter_first_index_indexed(ssq_counter *c, const u64 *words, const uint8_t *lens, int64_t n, const int64_t *indices,
                                int64_t base_index) {
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    if (!c->first_idx) {
        const size_t nslots = (size_t)1 << c->log2_cap;
        SSQ_CUDA(cudaMalloc(&c->first_idx, nslots * sizeof(u64)));
        SSQ_CUDA(cudaMemsetAsync(c->first_idx, 0xFF, nslots * sizeof(u64), ctx->stream));
    }
    if (n == 0) return SSQ_OK;
    int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
    TableView t = view_of(c);
    if (c->klass == SSQ_CLASS_64)
        first_index_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, ctx->stream>>>(t, (const u64 *)words, lens, n, base_index, indices);
    else
        first_index_kernel<SSQ_CLASS_192><<<grid, kThreads, 0, ctx->stream>>>(t, (const u64 *)words, lens, n, base_index, indices);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}
}  // namespace ssq

extern "C" {

int ssq_counter_lookup(ssq_counter *c, const uint64_t *words, const uint8_t *lens, int64_t n, uint64_t *counts) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    SSQ_ARG(n >= 0 && (n == 0 || (words != nullptr && lens != nullptr && counts != nullptr)), "bad batch");
    if (n == 0) return SSQ_OK;
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
    TableView t = view_of(c);
    if (c->klass == SSQ_CLASS_64)
        lookup_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, ctx->stream>>>(t, (const u64 *)words, lens, n, (u64 *)counts);
    else
        lookup_kernel<SSQ_CLASS_192><<<grid, kThreads, 0, ctx->stream>>>(t, (const u64 *)words, lens, n, (u64 *)counts);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_counter_size(ssq_counter *c, int64_t *n_unique) {
    SSQ_ARG(c != nullptr && n_unique != nullptr, "NULL argument");
    DeviceGuard g(c->ctx->device);
    SSQ_CUDA(cudaMemcpyAsync(c->h_size, c->d_size, sizeof(u64), cudaMemcpyDeviceToHost, c->ctx->stream));
    SSQ_CUDA(cudaStreamSynchronize(c->ctx->stream));
    *n_unique = (int64_t)*c->h_size;
    return SSQ_OK;
}

int ssq_counter_last_pass_ms(ssq_counter *c, float *phase1_ms, float *phase2_ms) {
    SSQ_ARG(c != nullptr && phase1_ms != nullptr && phase2_ms != nullptr, "NULL argument");
    *phase1_ms = *phase2_ms = 0.0f;
    if (c->last_pass_phases == 0) return SSQ_OK;
    DeviceGuard g(c->ctx->device);
    SSQ_CUDA(cudaEventSynchronize(c->ev[3]));
    SSQ_CUDA(cudaEventElapsedTime(phase1_ms, c->ev[0], c->ev[1]));
    SSQ_CUDA(cudaEventElapsedTime(phase2_ms, c->ev[1], c->ev[3]));
    return SSQ_OK;
}

int ssq_counter_last_pass_detail(ssq_counter *c, float *pack_scatter_ms, float *region_scatter_ms, float *count_ms) {
    SSQ_ARG(c != nullptr && pack_scatter_ms != nullptr && region_scatter_ms != nullptr && count_ms != nullptr, "NULL argument");
    *pack_scatter_ms = *region_scatter_ms = *count_ms = 0.0f;
    if (c->last_pass_phases == 0) return SSQ_OK;
    DeviceGuard g(c->ctx->device);
    SSQ_CUDA(cudaEventSynchronize(c->ev[3]));
    SSQ_CUDA(cudaEventElapsedTime(pack_scatter_ms, c->ev[0], c->ev[1]));
    SSQ_CUDA(cudaEventElapsedTime(region_scatter_ms, c->ev[1], c->ev[2]));
    SSQ_CUDA(cudaEventElapsedTime(count_ms, c->ev[2], c->ev[3]));
    return SSQ_OK;
}

int ssq_counter_capacity(ssq_counter *c, int64_t *slots) {
    SSQ_ARG(c != nullptr && slots != nullptr, "NULL argument");
    *slots = (int64_t)1 << c->log2_cap;
    return SSQ_OK;
}

int ssq_counter_export(ssq_counter *c, int n_parts, uint64_t *words, uint8_t *lens, uint64_t *counts,
                       int64_t *first_idx, int64_t *part_counts) {
    SSQ_ARG(c != nullptr, "counter is NULL");
    SSQ_ARG(n_parts >= 1 && n_parts <= kMaxParts && (n_parts & (n_parts - 1)) == 0, "n_parts must be a power of two <= 256");
    SSQ_ARG(words != nullptr && lens != nullptr && counts != nullptr && part_counts != nullptr, "NULL buffer");
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    if (region_export_ok(c, log2_parts)) {                 // ShortSeq64: one pass, region by region
        int rc = scan_regions(c);
        if (rc) return rc;
        export_part_counts_kernel<<<1, kMaxParts, 0, st>>>(c->region_base, c->log2_cap - region_bits_for(c->log2_cap), log2_parts, part_counts);
        SSQ_LAUNCH_CHECK();
        ExportDst d{(u64 *)words, lens, (u64 *)counts, first_idx, nullptr, nullptr, nullptr};
        return launch_export_regions(c, log2_parts, 0, d);
    }
    u64 *cursors = nullptr;
    {
        void *scratch = nullptr;
        int rc0 = ctx_scratch(ctx, kMaxParts * sizeof(u64), &scratch);
        if (rc0) return rc0;
        cursors = (u64 *)scratch;
    }
    SSQ_CUDA(cudaMemsetAsync(part_counts, 0, n_parts * sizeof(u64), st));
    TableView t = view_of(c);
    const int64_t cap = (int64_t)1 << c->log2_cap;
    int grid1 = grid_for(ctx, cap / kThreads, 8);
    int grid2 = grid_for(ctx, cap / (kThreads * kExportItems), 4);
    if (c->klass == SSQ_CLASS_64) {
        export_count_kernel<SSQ_CLASS_64><<<grid1, kThreads, 0, st>>>(t, log2_parts, (u64 *)part_counts);
        export_bases_kernel<<<1, 1, 0, st>>>((const u64 *)part_counts, cursors, n_parts);
        export_scatter_kernel<SSQ_CLASS_64><<<grid2, kThreads, 0, st>>>(t, log2_parts, cursors, (u64 *)words, lens,
                                                                        (u64 *)counts, first_idx);
    } else {
        export_count_kernel<SSQ_CLASS_192><<<grid1, kThreads, 0, st>>>(t, log2_parts, (u64 *)part_counts);
        export_bases_kernel<<<1, 1, 0, st>>>((const u64 *)part_counts, cursors, n_parts);
        export_scatter_kernel<SSQ_CLASS_192><<<grid2, kThreads, 0, st>>>(t, log2_parts, cursors, (u64 *)words, lens,
                                                                         (u64 *)counts, first_idx);
    }
    count_launch(); count_launch(); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "export kernels", __FILE__, __LINE__);
    return SSQ_OK;
}

}  // extern "C"

namespace ssq {
// Two-pass export of a table without regions (ShortSeq192), split for the multi-GPU exchange (ssq_comm.cu):
// counter_export_count_pass fills part_counts[n_parts] (pass 1 over the table); counter_export_scatter_to then stores
// partition p's tuples at dst_*[p] (pass 2), which may be peer memory.  Nothing may change the table in between.
int counter_export_count_pass(ssq_counter *c, int n_parts, int64_t *part_counts) {
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    ssq_ctx *ctx = c->ctx;
    const int64_t cap = (int64_t)1 << c->log2_cap;
    SSQ_CUDA(cudaMemsetAsync(part_counts, 0, n_parts * sizeof(u64), ctx->stream));
    const int grid = grid_for(ctx, cap / kThreads, 8);
    if (c->klass == SSQ_CLASS_64) export_count_kernel<SSQ_CLASS_64><<<grid, kThreads, 0, ctx->stream>>>(view_of(c), log2_parts, (u64 *)part_counts);
    else export_count_kernel<SSQ_CLASS_192><<<grid, kThreads, 0, ctx->stream>>>(view_of(c), log2_parts, (u64 *)part_counts);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}
int counter_export_scatter_to(ssq_counter *c, int n_parts, u64 *const *dst_words, uint8_t *const *dst_lens, u64 *const *dst_counts) {
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    ssq_ctx *ctx = c->ctx;
    void *scratch = nullptr;
    int rc = ctx_scratch(ctx, kMaxParts * sizeof(u64), &scratch);
    if (rc) return rc;
    u64 *cursors = (u64 *)scratch;
    SSQ_CUDA(cudaMemsetAsync(cursors, 0, kMaxParts * sizeof(u64), ctx->stream));      // positions relative to each destination
    const int64_t cap = (int64_t)1 << c->log2_cap;
    const int W = c->klass == SSQ_CLASS_64 ? 1 : 3;
    const size_t stage_bytes = (size_t)kThreads * kExportItems * (8 * (W + 1) + 1);
    const int grid = grid_for(ctx, cap / (kThreads * kExportItems), 3);
    if (c->klass == SSQ_CLASS_64) {
        rc = set_max_smem((const void *)export_scatter_kernel<SSQ_CLASS_64>, stage_bytes);
        if (rc) return rc;
        export_scatter_kernel<SSQ_CLASS_64><<<grid, kThreads, stage_bytes, ctx->stream>>>(view_of(c), log2_parts, cursors, nullptr, nullptr, nullptr, nullptr, dst_words, dst_lens, dst_counts);
    } else {
        rc = set_max_smem((const void *)export_scatter_kernel<SSQ_CLASS_192>, stage_bytes);
        if (rc) return rc;
        export_scatter_kernel<SSQ_CLASS_192><<<grid, kThreads, stage_bytes, ctx->stream>>>(view_of(c), log2_parts, cursors, nullptr, nullptr, nullptr, nullptr, dst_words, dst_lens, dst_counts);
    }
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}
}  // namespace ssq

extern "C" {

int ssq_counter_export_counts(ssq_counter *c, int n_parts, int64_t *part_counts) {
    SSQ_ARG(c != nullptr && part_counts != nullptr, "NULL argument");
    SSQ_ARG(n_parts >= 1 && n_parts <= kMaxParts && (n_parts & (n_parts - 1)) == 0, "n_parts must be a power of two <= 256");
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    if (!region_export_ok(c, log2_parts)) { set_error("ssq_counter_export_counts needs a ShortSeq64 counter with at least n_parts regions"); return SSQ_ERR_ARG; }
    DeviceGuard g(c->ctx->device);
    int rc = scan_regions(c);
    if (rc) return rc;
    export_part_counts_kernel<<<1, kMaxParts, 0, c->ctx->stream>>>(c->region_base, c->log2_cap - region_bits_for(c->log2_cap), log2_parts, part_counts);
    SSQ_LAUNCH_CHECK();
    return SSQ_OK;
}

int ssq_counter_export_to(ssq_counter *c, int n_parts, int first_part, uint64_t *const *dst_words, uint8_t *const *dst_lens,
                          uint64_t *const *dst_counts) {
    SSQ_ARG(c != nullptr && dst_words != nullptr && dst_lens != nullptr && dst_counts != nullptr, "NULL argument");
    SSQ_ARG(n_parts >= 1 && n_parts <= kMaxParts && (n_parts & (n_parts - 1)) == 0, "n_parts must be a power of two <= 256");
    SSQ_ARG(first_part >= 0 && first_part < n_parts, "first_part out of range");
    int log2_parts = 0;
    while ((1 << log2_parts) < n_parts) log2_parts++;
    if (!region_export_ok(c, log2_parts)) { set_error("ssq_counter_export_to needs a ShortSeq64 counter with at least n_parts regions"); return SSQ_ERR_ARG; }
    DeviceGuard g(c->ctx->device);
    int rc = scan_regions(c);
    if (rc) return rc;
    ExportDst d{nullptr, nullptr, nullptr, nullptr, (u64 *const *)dst_words, dst_lens, (u64 *const *)dst_counts};
    return launch_export_regions(c, log2_parts, first_part, d);
}

}  // extern "C"

namespace ssq {

// ssq_counter_merge_regions; flags != nullptr (ssq_comm.cu): the blocks are being written by other GPUs and block b is
// complete once flags[b] >= epoch -- the kernels wait on the device.
int counter_merge_regions_impl(ssq_counter *c, const u64 *words, const uint8_t *lens, const u64 *counts, const int64_t *block_counts,
                               const int64_t *block_regions, int n_blocks, const int64_t *region_bases, int64_t rb_stride,
                               const u64 *flags, u64 epoch) {
    SSQ_ARG(c != nullptr && block_counts != nullptr && block_regions != nullptr && n_blocks >= 1, "bad arguments");
    int64_t n = 0;
    for (int b = 0; b < n_blocks; b++) { SSQ_ARG(block_counts[b] >= 0, "negative block size"); n += block_counts[b]; }
    SSQ_ARG(n == 0 || (words != nullptr && lens != nullptr && counts != nullptr && region_bases != nullptr), "NULL buffer");
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    auto plain = [&]() -> int {                              // weighted global insert of everything received
        if (flags) {
            wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(flags, n_blocks, epoch, ctx->d_report);
            SSQ_LAUNCH_CHECK();
        }
        if (n == 0) return SSQ_OK;
        return ::insert_common(c, (const uint64_t *)words, lens, (const uint64_t *)counts, n);
    };
    if (n == 0) return plain();
    const int lr = region_bits_for(c->log2_cap);
    const int64_t my_regions = (int64_t)1 << (c->log2_cap - lr);
    bool ok = c->klass == SSQ_CLASS_64 && c->expected_unique > 0 && n_blocks <= kMaxMergeBlocks && lr <= 12;
    MergeRegions mr;
    mr.n = n_blocks;
    mr.rb_stride = rb_stride;
    mr.off[0] = 0;
    for (int b = 0; b < n_blocks && ok; b++) {
        mr.off[b + 1] = mr.off[b] + block_counts[b];
        int r = 0;
        while (((int64_t)my_regions << r) < block_regions[b]) r++;
        ok = ((int64_t)my_regions << r) == block_regions[b] && block_regions[b] + 1 <= rb_stride;   // a whole number of sender regions per owner region
        mr.ratio_log2[b] = r;
    }
    if (!ok) return plain();                                 // the region grids do not nest: plain weighted insert
    {
        const int before = c->log2_cap;
        int rc0 = make_room(c, n);
        if (rc0) return rc0;
        if (c->log2_cap != before || c->expected_unique <= 0) return plain();   // the region grid changed under the blocks / bound given up
    }
    const size_t bytes = merge_regions2_smem(lr);
    int rc = set_max_smem((const void *)merge_regions_kernel<false>, bytes);
    if (rc) return rc;
    merge_regions_kernel<false><<<(unsigned)my_regions, kMerge2Threads, bytes, ctx->stream>>>(view_of(c), words, lens, counts, region_bases, mr,
                                                                                             MergeStream{}, flags, epoch);
    SSQ_LAUNCH_CHECK();
    return finish_pass(c);
}

// Owner side of the streamed exchange: entries / cnts are this rank's receive buffers (n_blocks sender blocks of
// block_regions sender regions of 2^log2_sregion entries).  SSQ_ERR_ARG when the region grids do not nest -- the caller
// checked that before anything was streamed (counter_stream_nests).
bool counter_stream_nests(const ssq_counter *owner, int n_blocks, int64_t block_regions, int log2_sregion) {
    const int lr = region_bits_for(owner->log2_cap);
    const int64_t my_regions = (int64_t)1 << (owner->log2_cap - lr);
    int r = 0;
    while ((my_regions << r) < block_regions) r++;
    return owner->klass == SSQ_CLASS_64 && owner->expected_unique > 0 && lr <= 12 && owner->log2_cap - lr >= 6 &&
           (my_regions << r) == block_regions && ((int64_t)n_blocks << r) <= kMaxMergeBlocks && log2_sregion >= 1 && log2_sregion <= 16;
}

int counter_merge_streamed_impl(ssq_counter *c, const void *entries, const uint32_t *cnts, int n_blocks, int64_t block_regions,
                                int log2_sregion, const u64 *flags, u64 epoch) {
    SSQ_ARG(c != nullptr && entries != nullptr && cnts != nullptr && n_blocks >= 1, "bad arguments");
    ssq_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    if (!counter_stream_nests(c, n_blocks, block_regions, log2_sregion)) { set_error("streamed merge: the region grids do not nest"); return SSQ_ERR_ARG; }
    const int before = c->log2_cap;
    int rc = make_room(c, INT64_MAX / 4);                    // the number of incoming tuples is not known: the caller's bound decides
    if (rc) return rc;
    if (c->log2_cap != before || c->expected_unique <= 0) { set_error("streamed merge: the owner table had to grow"); return SSQ_ERR_ARG; }
    c->mod_seq++;
    const int lr = region_bits_for(c->log2_cap);
    const int64_t my_regions = (int64_t)1 << (c->log2_cap - lr);
    MergeRegions mr{};
    mr.n = n_blocks;
    MergeStream ms;
    ms.entries = (const ulonglong2 *)entries;
    ms.cnts = cnts;
    ms.block_regions = block_regions;
    ms.log2_sregion = log2_sregion;
    ms.ratio_log2 = 0;
    while ((my_regions << ms.ratio_log2) < block_regions) ms.ratio_log2++;
    const size_t bytes = merge_regions2_smem(lr);
    rc = set_max_smem((const void *)merge_regions_kernel<true>, bytes);
    if (rc) return rc;
    merge_regions_kernel<true><<<(unsigned)my_regions, kMerge2Threads, bytes, ctx->stream>>>(view_of(c), nullptr, nullptr, nullptr, nullptr, mr, ms, flags, epoch);
    SSQ_LAUNCH_CHECK();
    return finish_pass(c);
}

}  // namespace ssq
