// ssq_device.cuh -- device-side building blocks shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssq {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 kNoIndex = 0xFFFFFFFFFFFFFFFFull;

// Device-resident error record of a context.  Every field is an atomicMin /
// atomicOr target; ssq_ctx_sync() copies it back and resets it.
struct DevReport {
    u64 first_bad_base;      // lowest read index with a non-ACGT byte
    u64 first_bad_len;       // lowest read index whose length is outside the class (<= 1024)
    u64 first_too_long;      // lowest read index longer than 1024
    u64 first_len_mismatch;  // lowest pair index with len_a != len_b (Hamming)
    u64 table_overflow;      // number of inserts that found no slot (counter unusable if != 0)
    u64 exchange_timeout;    // multi-GPU merge: arrival flags that never showed up (ssq_comm.cu)
    u64 pad[2];
};

// ---- hashing ---------------------------------------------------------------
// splitmix64 finaliser: a bijection on 64-bit words, so a ShortSeq64 key can be
// stored in the table as its hash and recovered on export.
__host__ __device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ u64 unmix64(u64 x) {
    x ^= (x >> 31) ^ (x >> 62); x *= 0x319642B2D24D8EC3ull;
    x ^= (x >> 27) ^ (x >> 54); x *= 0x96DE1B173F119089ull;
    x ^= (x >> 30) ^ (x >> 60);
    return x;
}
__host__ __device__ __forceinline__ u64 rotl64(u64 x, int r) { return r ? (x << r) | (x >> (64 - r)) : x; }
__host__ __device__ __forceinline__ u64 rotr64(u64 x, int r) { return r ? (x >> r) | (x << (64 - r)) : x; }

// Table hash of a ShortSeq64 key: a BIJECTION on 64-bit words (the table stores the hash and recovers the word on export)
// whose top bits -- partition, region and home slot are the top bits of the hash -- depend on every input bit: the high
// word is folded into the low word (so that reads differing only in their last bases do not differ only in the top hash
// bits), then one multiply by an odd constant.  Five instructions; the splitmix64 finaliser it replaces (two 64-bit
// multiplies, three xor-shifts) was 9 % of the fused pack kernel's instructions, and the variable rotate behind it
// (rot = 0 on one GPU) another 4 %: table_hash64 takes the rotation as two funnel shifts.
constexpr u64 kHashMul = 0x9E3779B97F4A7C15ull, kHashMulInv = 0xF1DE83E19937733Dull;   // kHashMul * kHashMulInv == 1 (mod 2^64)
__host__ __device__ __forceinline__ u64 hash64(u64 x) { return (x ^ (x >> 32)) * kHashMul; }
__host__ __device__ __forceinline__ u64 unhash64(u64 h) { const u64 y = h * kHashMulInv; return y ^ (y >> 32); }
// rotl(hash64(word), rot), rot in 0..63 and uniform over the launch
__device__ __forceinline__ u64 table_hash64(u64 word, int rot) {
    const u64 h = hash64(word);
    u32 lo = (u32)h, hi = (u32)(h >> 32);
    if (rot & 32) { const u32 t = lo; lo = hi; hi = t; }
    const u32 r = (u32)rot & 31u;
    return ((u64)__funnelshift_l(lo, hi, r) << 32) | __funnelshift_l(hi, lo, r);
}

// Slot hash of a 3-word key (ShortSeq192).  Not a bijection; the key is stored verbatim.  The three words are folded
// with odd rotations (a 2-bit code never lines up with itself) and one multiply before a single splitmix round: the
// first version ran three rounds (six 64-bit multiplies) and was 12 % of the fused pack+scatter kernel's instructions.
__host__ __device__ __forceinline__ u64 hash192(u64 w0, u64 w1, u64 w2, u32 len) {
    return mix64(w0 ^ rotl64(w1, 21) ^ (rotl64(w2, 43) * 0x9E3779B97F4A7C15ull) ^ ((u64)len << 56));
}

// ---- memory intrinsics ---------------------------------------------------------
// 16-byte streaming load that does not allocate in L1 (every input byte is read once).
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// L2 cache policies: keep a line that will be touched again soon / drop a line that is read exactly once.
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ulonglong2 ld_hint_v2u64(const void *p, u64 policy) {
    ulonglong2 r;
    asm volatile("ld.global.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(policy) : "memory");
    return r;
}
// 8-byte streaming load: no L1 allocation, L2 line dropped first
__device__ __forceinline__ u64 ld_stream_u64(const void *p, u64 policy) {
    u64 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one 32-byte load (sm_100 LDG.256) of a 32-byte aligned slot; not an atomic snapshot as far as the memory model goes
__device__ __forceinline__ void ld_relaxed_v4u64(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.relaxed.gpu.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
// the same with an L2 cache policy (a table range that should stay resident while a stream of records passes through)
__device__ __forceinline__ void ld_relaxed_v4u64_hint(const u64 *p, u64 policy, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(policy) : "memory");
}
// 32-byte streaming load of a record that is read exactly once
__device__ __forceinline__ void ld_stream_v4u64(const void *p, u64 policy, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(policy));
}
__device__ __forceinline__ u64 ld_acquire_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(u64 *p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// fire-and-forget add (RED): the result is never needed
__device__ __forceinline__ void red_add_u64(u64 *p, u64 v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(u32 *p, u32 v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_min_u64(u64 *p, u64 v) {
    asm volatile("red.relaxed.gpu.global.min.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// ---- bulk asynchronous copies (TMA, 1-D) and the mbarrier they complete on ---------------------------
// One thread arms the barrier with the byte count and issues cp.async.bulk (SASS UBLKCP): the copy engine moves the
// whole range global -> shared without occupying registers or issue slots; consumers wait on the barrier's phase.
// Source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar, u64 policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// named barriers for a subset of the CTA's warps (the consumer warps of a warp-specialised kernel)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"(id), "r"(nthreads), "r"((int)pred) : "memory");
    return r != 0;
}
__device__ __forceinline__ uint4 lds_v4(u32 addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---- shared memory by 32-bit shared-space address ---------------------------------------------
// Hot loops address shared memory through a 32-bit shared-space address computed once: with generic pointers the
// compiler re-derives the shared window base (S2UR SR_CgaCtaId + ULEA ...) at every use inside the loop.
// The address passes through an opaque asm so that it lives in a register instead of being rematerialised.
__device__ __forceinline__ u32 smem_addr(const void *p) {
    u32 a = (u32)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}
__device__ __forceinline__ void lds_v2(u32 addr, u32 &lo, u32 &hi) {          // volatile: re-read every time
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
}
__device__ __forceinline__ u32 lds_u32(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(u32 addr, u64 v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 lds_u64(u32 addr) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 lds_v2u64(u32 addr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ u64 atoms_cas_u64(u32 addr, u64 cmp, u64 val) {
    u64 old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ u32 atoms_add_u32(u32 addr, u32 v) {
    u32 old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add_u32(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

// ---- 2-bit encoding -------------------------------------------------------------
// Four ASCII bases in one 32-bit word -> 8 bits of 2-bit codes in the TOP byte
// (code = (c>>1)&3: A0 C1 T2 G3, README.md:105-110 / util.pyx:39 of the reference).
// (w & 0x06060606) holds 2*code per byte; the multiply gathers the four 2-bit fields
// into bits 24..31 with no carries between partial products.
__device__ __forceinline__ u32 gather4(u32 w) { return (w & 0x06060606u) * 0x00820820u; }

// Exact {A,C,G,T} test of four bytes at once: returns a word that is non-zero iff
// some byte is not one of 0x41 0x43 0x47 0x54.  With m = [code == 2] (bit2 & ~bit1) per
// byte, a valid byte equals 0x41 | (code<<1) with 0x11 flipped when m: A,C,G keep
// bits 4,0 = 0,1 and T has them 1,0.
__device__ __forceinline__ u32 invalid4(u32 w) {
    u32 s1 = w >> 1;
    u32 m = (s1 >> 1) & ~s1 & 0x01010101u;
    u32 x = (w & 0xF9F9F9F9u) ^ (m * 0x11u);
    return x ^ 0x41414141u;
}

// 16 ASCII bytes (one uint4) -> 32 bits of codes (base k in bits 2k..2k+1); `bad`
// accumulates the invalid-byte indicator.
__device__ __forceinline__ u32 encode16(uint4 v, u32 &bad) {
    bad |= invalid4(v.x) | invalid4(v.y) | invalid4(v.z) | invalid4(v.w);
    u32 a = gather4(v.x), b = gather4(v.y), c = gather4(v.z), d = gather4(v.w);
    // pick the top byte of each: result = a.b3 | b.b3<<8 | c.b3<<16 | d.b3<<24
    u32 ab = __byte_perm(a, b, 0x0073);   // byte0 = a.b3, byte1 = b.b3
    u32 cd = __byte_perm(c, d, 0x0073);
    return __byte_perm(ab, cd, 0x5410);
}

// Exact per-byte test used on the rare slow path and for error positions.
__host__ __device__ __forceinline__ bool is_acgt(uint8_t c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

// popcount(((x>>1)|x) & 0x5555...) -- number of differing bases of one block
// (short_seq_64.pyx:82-84 of the reference).  The indicator bits sit at even positions only, so the high half is
// slid into the odd positions of the low half and ONE 32-bit POPC counts all 32 bases (POPC is a quarter-rate pipe).
__device__ __forceinline__ int diff_bases(u64 a, u64 b) {
    const u64 x = a ^ b;
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    const u32 ylo = ((lo >> 1) | lo) & 0x55555555u;
    const u32 yhi = ((hi >> 1) | hi) & 0x55555555u;
    return __popc(ylo | (yhi << 1));
}

}  // namespace ssq
