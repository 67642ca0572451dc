// ssq_table.cuh -- device-side open-addressing tables of the dedup counter.
//
// Replaces the CPython dict under ShortSeqCounter (reference counter.pyx:41-54):
// key = (length, words), value = multiplicity.
//
// ShortSeq64 table (16-byte slots {key, count}):
//   h2  = rotl(hash64(word), rot)       hash64 (ssq_device.cuh) is a bijection
//   home slot = top log2(cap) bits of h2 -> the table is ordered by hash, so a
//               hash partition is a contiguous slot range (multi-GPU export,
//               L2-sized insertion passes)
//   key = low 58 bits of h2 | (len+1) << 58
//   The table is cut into REGIONS of 2^lr slots (lr = region_bits_for(log2 cap): 12 for tables of
//   2^20..2^28 slots, so there are 256..65536 regions); linear probing wraps around inside the
//   key's region.  The top 6 bits of h2 are therefore not stored: a region lies inside one 1/64 of
//   the table, so they are the top 6 bits of the slot index.  A region is also the unit the deferred counting pass loads into shared memory
//   (64 KB at lr = 12): keys are routed to their region by two 256-way scatters.
//   len+1 >= 1 makes every key non-zero, so 0 marks an empty slot and one 64-bit
//   atomicCAS claims a slot AND publishes the whole key.
//
// ShortSeq192 table (32-byte slots {meta, w0, w1, w2}; one 32-byte sector):
//   meta = count << 8 | state, state 0 empty, 0xFF being written, else len-32.
//   A writer claims with atomicCAS(meta, 0, 0xFF), stores the words, then
//   releases meta; readers that see 0xFF re-read.
#pragma once
#include "ssq_device.cuh"

namespace ssq {

constexpr u64 kMask56 = (1ull << 56) - 1;
constexpr u64 kMask58 = (1ull << 58) - 1;   // hash bits kept in a ShortSeq64 table key
constexpr int kMinLog2Cap = 16;

// Hash partitions of the deferred-insert path.  Level 1: partition = top 8 bits of h2, i.e. 1/256 of the slot
// range.  Level 2 (inside one level-1 partition): the next bits of h2 down to one table region.
// Every CTA of a scatter pass owns a private segment of every partition, so appending needs no global atomics.
// Isolated 8-byte stores to 256 different places cost one memory transaction each (scripts/bench_scatter.cu);
// keys are therefore staged per partition in shared memory and written as whole 128-byte lines (see Stager).
constexpr int kParts = 256;
#ifndef SSQ_PARTS192
#define SSQ_PARTS192 128
#endif
constexpr int kParts192 = SSQ_PARTS192;    // ShortSeq192 records are 32 bytes: half the partitions keep the staging rings at 32 KB
#ifndef SSQ_LINE_KEYS
#define SSQ_LINE_KEYS 16
#endif
constexpr int kLineKeys = SSQ_LINE_KEYS;   // keys per flushed line (16 keys = 128 bytes)
constexpr int kRingKeys = 2 * kLineKeys;   // staging ring per partition per CTA
#ifndef SSQ_RING192
#define SSQ_RING192 64
#endif
constexpr int kRingKeys192 = SSQ_RING192;  // ShortSeq192: ring of kRingKeys192 / 4 records (16 records = 4 lines) per partition
constexpr size_t kStagerRingBytes = (size_t)kParts * kRingKeys * sizeof(u64);   // 64 KB of dynamic shared memory
struct PartView {
    u64 *keys;          // [num_ctas][kParts][seg_cap] table keys (key64_of) awaiting insertion
    u32 *seg_count;     // [num_ctas][kParts] entries per segment (written when the scatter CTA finishes)
    u32 seg_cap;        // entries per segment (multiple of kLineKeys); keys beyond it are inserted directly by the scatter pass
    u32 num_ctas;       // grid size of the scatter pass
    u32 flush_every;    // fused pack+scatter: tiles between two flushes of the staging rings
    // ShortSeq192: records that find their staging ring full are appended to the CTA's overflow segment (one scattered
    // 32-byte store instead of a latency-bound global insert inside the pack loop) and inserted after the partitions
    u64 *ovf;           // [num_ctas][ovf_cap][4]
    u32 *ovf_count;     // [num_ctas]
    u32 ovf_cap;        // records per CTA
};
// Level-2 buffers: CTA (p, s) of the second scatter -- slice s of level-1 partition p -- owns the segments
// keys[((p * slices + s) * kParts + r) * seg_cap ...] of the level-2 partitions r (hash bits 55..48) of partition p.
// A table region of partition p is 2^(8 - qbits) consecutive level-2 partitions (exactly one for 2^28 slots).
struct RegionParts {
    u64 *keys;
    u32 *seg_count;     // [kParts][slices][kParts]
    u32 seg_cap;
    u32 slices;         // CTAs per level-1 partition
    u32 qbits;          // log2(regions per level-1 partition) <= 8
};

struct TableView {
    u64 *slots;
    u64 *first_idx;   // may be null
    u64 *size;        // occupied-slot counter
    DevReport *rep;
    int log2_cap;
    int rot;
    int log2_region;  // ShortSeq64 tables: probing wraps inside regions of 2^log2_region slots
    u32 *region_count;  // ShortSeq64 tables: occupied slots per region (the export derives its offsets from them)
};

// Regions of 2^12 slots for tables of 2^18 .. 2^28 slots (64 .. 65536 regions); 1/64 of a smaller table; larger
// regions beyond 2^28 slots so that a level-1 partition never holds more than 256 of them.
#ifndef SSQ_REGION_SHIFT
#define SSQ_REGION_SHIFT 16      /* development: 15 / 14 -> regions of 2^13 / 2^14 slots at 2^28 slots */
#endif
__host__ __device__ __forceinline__ int region_bits_for(int log2_cap) {
    return log2_cap < 18 ? log2_cap - 6 : (log2_cap <= 12 + SSQ_REGION_SHIFT ? 12 : log2_cap - SSQ_REGION_SHIFT);
}

// ---- ShortSeq64 ------------------------------------------------------------
__device__ __forceinline__ u64 key64_of(u64 h2, u32 len) {
    return (h2 & kMask58) | ((u64)(len + 1) << 58);
}

// Returns the slot index (or kNoIndex on overflow); is_new is set when this call created the key.
// defer_region: the caller adds the new key to region_count itself (warp-aggregated, see add_region_counts)
__device__ __forceinline__ u64 insert64_hashed(const TableView &t, u64 h2, u64 key, u64 add, bool &is_new, bool defer_region = false) {
    const u32 rmask = (1u << t.log2_region) - 1;
    const u64 home = h2 >> (64 - t.log2_cap);
    const u64 base = home & ~(u64)rmask;
    u32 off = (u32)home & rmask;
    is_new = false;
    for (u32 probe = 0; probe <= rmask; ++probe) {
        const u64 slot = base | off;
        u64 *p = t.slots + 2 * slot;
        u64 k = ld_relaxed_u64(p);
        if (k == 0) {
            k = atomicCAS(p, 0ull, key);
            if (k == 0) {
                red_add_u64(p + 1, add);
                if (!defer_region) red_add_u32(t.region_count + (slot >> t.log2_region), 1u);
                is_new = true;
                return slot;
            }
        }
        if (k == key) { red_add_u64(p + 1, add); return slot; }
        off = (off + 1) & rmask;
    }
    atomicAdd(&t.rep->table_overflow, 1ull);   // the key's region is full
    return kNoIndex;
}

__device__ __forceinline__ u64 insert64(const TableView &t, u64 word, u32 len, u64 add, bool &is_new, bool defer_region = false) {
    u64 h2 = table_hash64(word, t.rot);
    return insert64_hashed(t, h2, key64_of(h2, len), add, is_new, defer_region);
}

// Region occupancy update for the lanes of a warp that just created keys (slot = their slots): lanes whose keys fell
// into the same region send ONE add.  Hash-ordered input puts a whole warp into one or two regions, and same-address
// atomics serialise in L2.  Call with the warp's active lanes converged.
__device__ __forceinline__ void add_region_counts(const TableView &t, bool is_new, u64 slot) {
    const unsigned act = __activemask();
    const u32 region = is_new ? (u32)(slot >> t.log2_region) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(act, region);
    if (is_new && (threadIdx.x & 31) == (u32)(__ffs(peers) - 1)) red_add_u32(t.region_count + region, (u32)__popc(peers));
}

__device__ __forceinline__ u64 find64(const TableView &t, u64 word, u32 len) {
    const u32 rmask = (1u << t.log2_region) - 1;
    u64 h2 = table_hash64(word, t.rot);
    u64 key = key64_of(h2, len);
    const u64 home = h2 >> (64 - t.log2_cap);
    const u64 base = home & ~(u64)rmask;
    u32 off = (u32)home & rmask;
    for (u32 probe = 0; probe <= rmask; ++probe) {
        const u64 slot = base | off;
        u64 k = ld_relaxed_u64(t.slots + 2 * slot);
        if (k == key) return slot;
        if (k == 0) return kNoIndex;
        off = (off + 1) & rmask;
    }
    return kNoIndex;
}

// Recover (h2, len) of an occupied ShortSeq64 slot: probing never leaves the region, and a region lies inside
// one 1/64 of the table, so the top 6 bits of h2 are the top 6 bits of the slot index.
__device__ __forceinline__ u64 slot64_h2(const TableView &t, u64 slot, u64 key, u32 &len) {
    const u64 g = slot >> (t.log2_cap - 6);
    len = (u32)(key >> 58) - 1;
    return (g << 58) | (key & kMask58);
}

// ---- ShortSeq192 -----------------------------------------------------------
constexpr u64 kLocked = 0xFFull;

// h2 = rotl(hash192(w0, w1, w2, len), rot); only its top log2(cap) bits are used.
// PRE: (meta, a0, a1, a2) already hold a relaxed 32-byte view of the home slot (the caller issued the loads of several
// keys' home slots together so that their latencies overlap).
template <bool PRE>
__device__ __forceinline__ u64 insert192_impl(const TableView &t, u64 h2, u64 w0, u64 w1, u64 w2, u32 len, u64 add, bool &is_new,
                                              u64 meta, u64 a0, u64 a1, u64 a2) {
    const u64 mask = (1ull << t.log2_cap) - 1;
    const u64 state = (u64)(len - 32);
    u64 slot = h2 >> (64 - t.log2_cap);
    is_new = false;
    u64 probes = 0;
    const u64 limit = 1ull << t.log2_cap;
    bool have = PRE;
    while (probes < limit) {
        u64 *p = t.slots + 4 * slot;
        // The whole slot in one 32-byte load.  A full match is conclusive (words are written once, before the state is
        // released); anything else that could be a torn view is confirmed with ordered loads below.
        if (!have) ld_relaxed_v4u64(p, meta, a0, a1, a2);
        have = false;
        u64 st = meta & 0xFF;
        if (st == state && a0 == w0 && a1 == w1 && a2 == w2) {
            red_add_u64(p, add << 8);
            return slot;
        }
        if (st == 0) {
            if (atomicCAS(p, 0ull, kLocked) == 0) {
                st_relaxed_u64(p + 1, w0);
                st_relaxed_u64(p + 2, w1);
                st_relaxed_u64(p + 3, w2);
                st_release_u64(p, (add << 8) | state);   // nobody else touches meta while it is locked
                is_new = true;
                return slot;
            }
            continue;                                      // somebody else took the slot: look again
        }
        if (st == kLocked) continue;                       // being written: look again
        // Same length but the words differed.  Words change exactly once, from 0 to their final value, so a word that
        // is non-zero and different settles it; only a view whose differing words are all still 0 could be torn.
        const bool settled = (a0 != w0 && a0 != 0) || (a1 != w1 && a1 != 0) || (a2 != w2 && a2 != 0);
        if (st == state && !settled) {
            const u64 m2 = ld_acquire_u64(p);
            if ((m2 & 0xFF) == state && ld_relaxed_u64(p + 1) == w0 && ld_relaxed_u64(p + 2) == w1 &&
                ld_relaxed_u64(p + 3) == w2) {
                red_add_u64(p, add << 8);
                return slot;
            }
        }
        slot = (slot + 1) & mask;
        ++probes;
    }
    atomicAdd(&t.rep->table_overflow, 1ull);
    return kNoIndex;
}

__device__ __forceinline__ u64 insert192_hashed(const TableView &t, u64 h2, u64 w0, u64 w1, u64 w2, u32 len, u64 add, bool &is_new) {
    return insert192_impl<false>(t, h2, w0, w1, w2, len, add, is_new, 0, 0, 0, 0);
}

__device__ __forceinline__ u64 insert192(const TableView &t, u64 w0, u64 w1, u64 w2, u32 len, u64 add, bool &is_new) {
    return insert192_hashed(t, rotl64(hash192(w0, w1, w2, len), t.rot), w0, w1, w2, len, add, is_new);
}

// A ShortSeq192 key on its way through the hash partitions: {w0, w1, w2, meta}, meta = h2 with its low byte replaced
// by len - 32 (the home slot only needs the top bits of h2).
__device__ __forceinline__ u64 meta192_of(u64 h2, u32 len) { return (h2 & ~0xFFull) | (u64)(len - 32); }

__device__ __forceinline__ u64 find192(const TableView &t, u64 w0, u64 w1, u64 w2, u32 len) {
    const u64 mask = (1ull << t.log2_cap) - 1;
    const u64 state = (u64)(len - 32);
    u64 slot = rotl64(hash192(w0, w1, w2, len), t.rot) >> (64 - t.log2_cap);
    const u64 limit = 1ull << t.log2_cap;
    for (u64 probes = 0; probes < limit; ++probes) {
        const u64 *p = t.slots + 4 * slot;
        u64 st = ld_acquire_u64(p) & 0xFF;
        if (st == 0) return kNoIndex;
        if (st == state && ld_relaxed_u64(p + 1) == w0 && ld_relaxed_u64(p + 2) == w1 &&
            ld_relaxed_u64(p + 3) == w2)
            return slot;
        slot = (slot + 1) & mask;
    }
    return kNoIndex;
}

// ---- staged scatter into hash partitions (ShortSeq64) ------------------------------------------
// Shared-memory staging of one scatter CTA: a ring of kRingKeys keys per partition plus two monotonic
// cursors (head = keys staged so far, tail = keys already flushed).  Keys leave the SM only as whole,
// aligned LINES of kLineKeys keys (128 bytes), copied by kLineKeys/2 neighbouring lanes with one 16-byte
// store each -- one full-line memory transaction per 16 keys instead of one per key.
struct Stager {      // 32-bit shared-space addresses (see smem_addr)
    u32 ring;      // u64 [kParts][kRingKeys]
    u32 head;      // u32 [kParts]
    u32 tail;      // u32 [kParts]
    u32 list;      // u32 [warps][32 * lines per ring] scratch of flush_lines
};

__device__ __forceinline__ Stager make_stager(const u64 *ring, const u32 *head, const u32 *tail, const u32 *list) {
    return Stager{smem_addr(ring), smem_addr(head), smem_addr(tail), smem_addr(list)};
}

template <int PARTS = kParts>
__device__ __forceinline__ void stager_init(const Stager &s) {
    for (u32 p = threadIdx.x; p < (u32)PARTS; p += blockDim.x) { sts_u32(s.head + 4 * p, 0); sts_u32(s.tail + 4 * p, 0); }
}

// Append one key to partition `part`.  Returns false when the ring is full (more than kRingKeys keys of one
// partition between two flushes: heavily skewed input); the caller then inserts the key directly.
__device__ __forceinline__ bool stage_key(const Stager &s, u32 part, u64 key) {
    const u32 pos = atoms_add_u32(s.head + 4 * part, 1u);
    if (pos - lds_u32(s.tail + 4 * part) >= (u32)kRingKeys) return false;     // flush_lines clamps head back
    sts_u64(s.ring + 8 * (part * kRingKeys + (pos & (kRingKeys - 1))), key);
    return true;
}

// stage_key for N keys at once: the N shared-memory atomics are issued back to back, then the N tail reads, then the N
// stores, so their latencies overlap instead of adding up (each step of stage_key waits for the one before it).
// valid[j] = false skips key j; ok[j] = false on return: ring full, the caller inserts key j directly.
template <int N>
__device__ __forceinline__ void stage_keys(const Stager &s, const u32 (&part)[N], const u64 (&key)[N], const bool (&valid)[N], bool (&ok)[N]) {
    u32 pos[N], tl[N];
#pragma unroll
    for (int j = 0; j < N; j++) pos[j] = valid[j] ? atoms_add_u32(s.head + 4 * part[j], 1u) : 0u;
#pragma unroll
    for (int j = 0; j < N; j++) tl[j] = valid[j] ? lds_u32(s.tail + 4 * part[j]) : 0u;
#pragma unroll
    for (int j = 0; j < N; j++) {
        ok[j] = !valid[j] || pos[j] - tl[j] < (u32)kRingKeys;
        if (valid[j] && ok[j]) sts_u64(s.ring + 8 * (part[j] * kRingKeys + (pos[j] & (kRingKeys - 1))), key[j]);
    }
}

// Keys that find no room in their segment are counted right away.  `fixed_top` >= 0: the top 8 hash bits of
// every key of this CTA (second-level scatter); < 0: the partition index is the top 8 bits (first level).
static __device__ __noinline__ void insert_unstaged(const TableView &t, u64 key, u32 top, u32 *my_new) {
    bool is_new = false;
    insert64_hashed(t, ((u64)top << 56) | (key & kMask56), key, 1ull, is_new);
    if (is_new) atomicAdd(my_new, 1u);
}
// the same for a staged ShortSeq192 record at shared-space address `rec`
static __device__ __noinline__ void insert_unstaged192(const TableView &t, u32 rec, u32 *my_new) {
    const ulonglong2 a = lds_v2u64(rec), b = lds_v2u64(rec + 16);
    bool is_new = false;
    insert192_hashed(t, b.y & ~0xFFull, a.x, a.y, b.x, (u32)(b.y & 0xFF) + 32, 1ull, is_new);
    if (is_new) atomicAdd(my_new, 1u);
}

// slow paths of the scatter kernels, kept out of line so that they do not cost the hot loop registers
static __device__ __noinline__ void insert192_slow(const TableView &t, u64 h2, u64 w0, u64 w1, u64 w2, u32 len, u32 *my_new) {
    bool is_new = false;
    insert192_hashed(t, h2, w0, w1, w2, len, 1ull, is_new);
    if (is_new) atomicAdd(my_new, 1u);
}
static __device__ __noinline__ void insert64_slow(const TableView &t, u64 h2, u64 key, u32 *my_new) {
    bool is_new = false;
    insert64_hashed(t, h2, key, 1ull, is_new);
    if (is_new) atomicAdd(my_new, 1u);
}

// Append one ShortSeq192 record to partition `part` (ring of RING / 4 records).
template <int RING = kRingKeys192>
__device__ __forceinline__ bool stage_rec192(const Stager &s, u32 part, u64 w0, u64 w1, u64 w2, u64 meta) {
    constexpr u32 kRingRecs = RING / 4;
    const u32 pos = atoms_add_u32(s.head + 4 * part, 1u);
    if (pos - lds_u32(s.tail + 4 * part) >= kRingRecs) return false;
    const u32 a = s.ring + 8 * (part * RING + (pos & (kRingRecs - 1)) * 4);
    asm volatile("st.shared.v2.u64 [%0], {%1, %2};" :: "r"(a), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.shared.v2.u64 [%0], {%1, %2};" :: "r"(a + 16), "l"(w2), "l"(meta) : "memory");
    return true;
}

// Move every complete line (FINAL: everything) of the staging rings to this CTA's segments.  A record is RW
// 64-bit words (1: a ShortSeq64 table key, 4: a ShortSeq192 record); head / tail / seg_cap count records, a line is
// kLineKeys / RW records; partition q's segment starts at seg0 + q * seg_cap * RW.  All threads of the CTA call this
// between two barriers.  s_new is a shared-memory counter of keys created by the overflow path.
template <bool FINAL, int RW = 1, int PARTS = kParts, int RING = kRingKeys>
__device__ __forceinline__ void flush_lines(const Stager &s, u64 *seg0, u32 seg_cap, const TableView &t, int fixed_top,
                                            u32 *s_new) {
    constexpr u32 kLaneGroup = kLineKeys / 2;           // lanes that copy one line (16 bytes each)
    constexpr u32 kGroups = 32 / kLaneGroup;            // lines per warp-wide store
    constexpr u32 kRingRecs = RING / RW, kLineRecs = kLineKeys / RW;
    constexpr u32 kRingKeys = RING;                     // this instantiation's ring size shadows the default
    const u32 lane = threadIdx.x & 31;
    const u32 g = lane / kLaneGroup, sub = lane % kLaneGroup;
    const u32 lt_mask = (1u << lane) - 1;
    constexpr u32 kMaxLines = RING / kLineKeys;         // complete lines a ring can hold (2; 4 for ShortSeq192)
    const u32 list = s.list + (threadIdx.x >> 5) * (32 * kMaxLines * 4);  // this warp's work list: ready lines as (line << 5 | lane)
    for (u32 pbase = (threadIdx.x >> 5) * 32; pbase < (u32)PARTS; pbase += blockDim.x) {
        const u32 p = pbase + lane;
        const u32 tl = lds_u32(s.tail + 4 * p);
        const u32 hd = min(lds_u32(s.head + 4 * p), tl + kRingRecs);
        const u32 avail = hd - tl;
        const u32 nl = avail / kLineRecs;               // 0 .. kMaxLines complete lines
        u32 nready = 0;
#pragma unroll
        for (u32 l = 0; l < kMaxLines; l++) {
            const u32 m = __ballot_sync(0xFFFFFFFFu, nl > l);
            if (nl > l) sts_u32(list + 4 * (nready + __popc(m & lt_mask)), (l << 5) | lane);
            nready += __popc(m);
        }
        __syncwarp();
        for (u32 it = g; it < nready; it += kGroups) {   // group g copies the lines it, it + kGroups, ...
            const u32 e = lds_u32(list + 4 * it);
            const u32 q = pbase + (e & 31u);
            const u32 tq = lds_u32(s.tail + 4 * q) + (e >> 5) * kLineRecs;
            const u32 src = s.ring + 8 * (q * kRingKeys + (tq & (kRingRecs - 1)) * RW + 2 * sub);
            if (tq + kLineRecs <= seg_cap) {
#ifndef SSQ_X_NOSTORE
                *reinterpret_cast<ulonglong2 *>(seg0 + ((size_t)q * seg_cap + tq) * RW + 2 * sub) = lds_v2u64(src);
#endif
            } else if (RW == 1) {                         // segment full
                const u32 top = fixed_top >= 0 ? (u32)fixed_top : q;
                const ulonglong2 v = lds_v2u64(src);
                insert_unstaged(t, v.x, top, s_new);
                insert_unstaged(t, v.y, top, s_new);
            } else if ((sub & 1) == 0) {                  // every second lane takes one whole 32-byte record
                insert_unstaged192(t, src, s_new);
            }
        }
        __syncwarp();
        u32 new_tail = tl + nl * kLineRecs;
        if (FINAL) {                                      // the partial last line, record by record (once per CTA)
            const bool fits = tl + avail <= seg_cap;
            for (u32 j = nl * kLineRecs; j < avail; j++) {
                const u32 rec = s.ring + 8 * (p * kRingKeys + ((tl + j) & (kRingRecs - 1)) * RW);
                if (fits) {
#pragma unroll
                    for (int w = 0; w < RW; w++) seg0[((size_t)p * seg_cap + tl + j) * RW + w] = lds_u64(rec + 8 * w);
                } else if (RW == 1) {
                    insert_unstaged(t, lds_u64(rec), fixed_top >= 0 ? (u32)fixed_top : p, s_new);
                } else {
                    insert_unstaged192(t, rec, s_new);
                }
            }
            // after the final flush the tail is what stager_seg_count reports: the entries that really are in the
            // segment (whole lines up to seg_cap, plus the partial line only if it fitted)
            new_tail = fits ? tl + avail : min(new_tail, seg_cap);
        }
        sts_u32(s.tail + 4 * p, new_tail);
        sts_u32(s.head + 4 * p, hd);
    }
}

// entries of partition p's segment that were written to global memory (call after the FINAL flush + barrier)
__device__ __forceinline__ u32 stager_seg_count(const Stager &s, u32 p, u32 seg_cap) { return min(lds_u32(s.tail + 4 * p), seg_cap); }

// Add the number of keys this warp created to the table's size counter with one atomic per warp.
__device__ __forceinline__ void add_new_keys(const TableView &t, bool is_new) {
    unsigned m = __ballot_sync(__activemask(), is_new);
    if (m != 0 && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(t.size, (u64)__popc(m));
}

}  // namespace ssq
