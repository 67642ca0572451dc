// ssq_table.cuh -- device-side open-addressing tables of the dedup counter.
//
// Replaces the CPython dict under ShortSeqCounter (reference counter.pyx:41-54):
// key = (length, words), value = multiplicity.
//
// ShortSeq64 table (16-byte slots {key, count}):
//   h2  = rotl(mix64(word), rot)        mix64 is a bijection
//   home slot = top log2(cap) bits of h2 -> the table is ordered by hash, so a
//               hash partition is a contiguous slot range (multi-GPU export,
//               L2-sized insertion passes)
//   key = low 56 bits of h2 | (len+1) << 56 | (bit 56 of h2) << 62
//   The top 8 bits of h2 are not stored: with linear probing bounded to less than
//   cap/256 slots they follow from the slot position and the stored parity bit.
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
constexpr int kMinLog2Cap = 16;

// Hash partitions of the deferred-insert path: partition = top 8 bits of h2, i.e. 1/256 of the slot range.
constexpr int kParts = 256;
// Every CTA of the scatter pass owns a private segment of every partition.  Keys are first staged per partition
// in shared memory (one shared-memory atomicAdd each) and leave the SM only as whole, aligned 32-byte sectors
// (4 keys, written with two 16-byte stores by the thread that owns the partition): isolated 8-byte stores to
// 256 different places cost one memory transaction each, ~4x the time of this scheme (scripts/bench_scatter.cu).
constexpr int kStageCap = 12;        // staging slots per partition per CTA
struct PartView {
    u64 *keys;          // [num_ctas][kParts][seg_cap] table keys (key64_of) awaiting insertion
    u32 *seg_count;     // [num_ctas][kParts] entries per segment (written when the scatter CTA finishes)
    u32 seg_cap;        // entries per segment (multiple of 4); keys beyond it are inserted directly by the scatter pass
    u32 num_ctas;       // grid size of the scatter pass
};

struct TableView {
    u64 *slots;
    u64 *first_idx;   // may be null
    u64 *size;        // occupied-slot counter
    DevReport *rep;
    int log2_cap;
    int rot;
};

// ---- ShortSeq64 ------------------------------------------------------------
__device__ __forceinline__ u64 key64_of(u64 h2, u32 len) {
    return (h2 & kMask56) | ((u64)(len + 1) << 56) | (((h2 >> 56) & 1ull) << 62);
}

// Returns the slot index (or kNoIndex on overflow); is_new is set when this call created the key.
__device__ __forceinline__ u64 insert64_hashed(const TableView &t, u64 h2, u64 key, u64 add, bool &is_new) {
    const u64 mask = (1ull << t.log2_cap) - 1;
    const u32 limit = 1u << (t.log2_cap - 8);
    u64 slot = h2 >> (64 - t.log2_cap);
    is_new = false;
    for (u32 probe = 0; probe < limit; ++probe) {
        u64 *p = t.slots + 2 * slot;
        u64 k = ld_relaxed_u64(p);
        if (k == 0) {
            k = atomicCAS(p, 0ull, key);
            if (k == 0) { red_add_u64(p + 1, add); is_new = true; return slot; }
        }
        if (k == key) { red_add_u64(p + 1, add); return slot; }
        slot = (slot + 1) & mask;
    }
    atomicAdd(&t.rep->table_overflow, 1ull);
    return kNoIndex;
}

__device__ __forceinline__ u64 insert64(const TableView &t, u64 word, u32 len, u64 add, bool &is_new) {
    u64 h2 = rotl64(mix64(word), t.rot);
    return insert64_hashed(t, h2, key64_of(h2, len), add, is_new);
}

__device__ __forceinline__ u64 find64(const TableView &t, u64 word, u32 len) {
    const u64 mask = (1ull << t.log2_cap) - 1;
    const u32 limit = 1u << (t.log2_cap - 8);
    u64 h2 = rotl64(mix64(word), t.rot);
    u64 key = key64_of(h2, len);
    u64 slot = h2 >> (64 - t.log2_cap);
    for (u32 probe = 0; probe < limit; ++probe) {
        u64 k = ld_relaxed_u64(t.slots + 2 * slot);
        if (k == key) return slot;
        if (k == 0) return kNoIndex;
        slot = (slot + 1) & mask;
    }
    return kNoIndex;
}

// Recover (h2, len) of an occupied ShortSeq64 slot.
__device__ __forceinline__ u64 slot64_h2(const TableView &t, u64 slot, u64 key, u32 &len) {
    u64 g = slot >> (t.log2_cap - 8);
    if ((g & 1ull) != ((key >> 62) & 1ull)) g = (g - 1) & 0xFFull;
    len = (u32)((key >> 56) & 0x3F) - 1;
    return (g << 56) | (key & kMask56);
}

// ---- ShortSeq192 -----------------------------------------------------------
constexpr u64 kLocked = 0xFFull;

__device__ __forceinline__ u64 insert192(const TableView &t, u64 w0, u64 w1, u64 w2, u32 len, u64 add, bool &is_new) {
    const u64 mask = (1ull << t.log2_cap) - 1;
    const u64 state = (u64)(len - 32);
    u64 slot = rotl64(hash192(w0, w1, w2, len), t.rot) >> (64 - t.log2_cap);
    is_new = false;
    u64 probes = 0;
    const u64 limit = 1ull << t.log2_cap;
    while (probes < limit) {
        u64 *p = t.slots + 4 * slot;
        u64 meta = ld_acquire_u64(p);
        u64 st = meta & 0xFF;
        if (st == 0) {
            u64 old = atomicCAS(p, 0ull, kLocked);
            if (old == 0) {
                st_relaxed_u64(p + 1, w0);
                st_relaxed_u64(p + 2, w1);
                st_relaxed_u64(p + 3, w2);
                st_release_u64(p, (add << 8) | state);   // nobody else touches meta while it is locked
                is_new = true;
                return slot;
            }
            st = old & 0xFF;
        }
        if (st == kLocked) continue;                       // being written: look again
        if (st == state && ld_relaxed_u64(p + 1) == w0 && ld_relaxed_u64(p + 2) == w1 &&
            ld_relaxed_u64(p + 3) == w2) {
            red_add_u64(p, add << 8);
            return slot;
        }
        slot = (slot + 1) & mask;
        ++probes;
    }
    atomicAdd(&t.rep->table_overflow, 1ull);
    return kNoIndex;
}

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
// Append one table key to the CTA's staging area; returns false when the partition's staging is full (the
// caller then inserts the key directly).
__device__ __forceinline__ bool stage_key(u64 *stage, u32 *scnt, u32 part, u64 key) {
    const u32 pos = atomicAdd(&scnt[part], 1u);
    if (pos >= (u32)kStageCap) return false;
    stage[part * kStageCap + pos] = key;
    return true;
}

// Thread `p` moves partition p's staged keys to the CTA's segment in global memory: whole sectors only, the
// remainder (< 4 keys) stays staged -- unless `final`, which empties the staging.  Call between two barriers.
__device__ __forceinline__ void flush_staged(u64 *stage, u32 *scnt, u32 *sgcur, u32 p, const PartView &pv, const TableView &t,
                                             bool final, u32 &my_new) {
    const u32 c = min(scnt[p], (u32)kStageCap);
    const u32 nfl = final ? c : (c & ~3u);
    u64 *src = stage + p * kStageCap;
    if (nfl) {
        const u32 g = sgcur[p];
        if (g + nfl <= pv.seg_cap) {
            u64 *dst = pv.keys + ((size_t)blockIdx.x * kParts + p) * pv.seg_cap + g;
            u32 j = 0;
            for (; j + 4 <= nfl; j += 4) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(src + j);
                const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(src + j + 2);
                *reinterpret_cast<ulonglong2 *>(dst + j) = a;
                *reinterpret_cast<ulonglong2 *>(dst + j + 2) = b;
            }
            for (; j < nfl; j++) dst[j] = src[j];
            sgcur[p] = g + nfl;
        } else {                                   // segment full: count these keys right away
            for (u32 j = 0; j < nfl; j++) {
                bool is_new = false;
                insert64_hashed(t, ((u64)p << 56) | (src[j] & kMask56), src[j], 1ull, is_new);
                my_new += is_new ? 1u : 0u;
            }
        }
        for (u32 j = nfl; j < c; j++) src[j - nfl] = src[j];
    }
    scnt[p] = c - nfl;
}

// Add the number of keys this warp created to the table's size counter with one atomic per warp.
__device__ __forceinline__ void add_new_keys(const TableView &t, bool is_new) {
    unsigned m = __ballot_sync(__activemask(), is_new);
    if (m != 0 && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(t.size, (u64)__popc(m));
}

}  // namespace ssq
