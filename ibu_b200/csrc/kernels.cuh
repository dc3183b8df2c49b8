// kernels.cuh — device-side building blocks shared by the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ibu {

constexpr int kWarp = 32;
constexpr int kBlockThreads = 256;
constexpr int kWarpsPerBlock = kBlockThreads / kWarp;
constexpr int kTileRecords = 128;                       // records per warp tile (unpack / reduce)
constexpr int kTileBytes = kTileRecords * 24;           // 3072
constexpr int kTileU4 = kTileBytes / 16;                // 192 x 16 B
constexpr int kTileU8 = kTileBytes / 32;                // 96 x 32 B
constexpr uint32_t kAcgt = 0x54474341u;                 // "ACGT" as PRMT lookup table

struct alignas(32) u64x4 {
    uint64_t x, y, z, w;
};

// ---- streaming global accesses: read-once / write-once data must not displace anything ----
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
        : "l"(p));
    return r;
}
// 256-bit load (sm_100+: LDG.E.256)
__device__ __forceinline__ u64x4 ldg_stream256(const void *p) {
    u64x4 r;
    asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];"
        : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w)
        : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t ldg_stream64(const uint64_t *p) {
    uint64_t r;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// (plain write-back stores: measured 1.4 % faster on K2 than L1::no_allocate / .cs, A/B in one
// process, 80 launches each)
__device__ __forceinline__ void stg_stream(uint4 *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream256(void *p, uint4 lo, uint4 hi) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                 "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z),
                 "r"(hi.w)
                 : "memory");
}

// ---- 2-bit -> ASCII, branch-free, LUT in a register (record.rs:19-27; base i at bits 2i..2i+1) ----
// 8 bases (16 bits: HALF picks the low or high half of w32) -> two u32 of ASCII.
template <int HALF>
__device__ __forceinline__ void decode8(uint32_t w32, uint32_t &a0, uint32_t &a1) {
    // spread the two source bytes to bytes 0 and 2, then nibbles, then 2-bit fields -> PRMT selectors
    uint32_t t = __byte_perm(w32, 0, HALF ? 0x4342 : 0x4140);
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;
    a0 = __byte_perm(kAcgt, 0, t);
    a1 = __byte_perm(kAcgt, 0, t >> 16);
}

// Decode the first 4*NG bases of w into asc[0..NG).
template <int NG>
__device__ __forceinline__ void decode_word(uint64_t w, uint32_t (&asc)[8]) {
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    if (NG > 0) decode8<0>(lo, asc[0], asc[1]);
    if (NG > 2) decode8<1>(lo, asc[2], asc[3]);
    if (NG > 4) decode8<0>(hi, asc[4], asc[5]);
    if (NG > 6) decode8<1>(hi, asc[6], asc[7]);
}

// ---- ASCII -> 2-bit, SWAR.  16 input bytes -> 32 bits; `bad` accumulates mismatch bits ----
// code = c' ^ (c' >> 1), c' = (c >> 1) & 3 maps A/a C/c G/g T/t to 0 1 2 3.  Validity without a
// table: bits 1-2 of a byte are its code by construction and bit 5 is the case, so only bits
// {7,6,4,3,0} remain to be checked; they must read 0x41, or 0x50 for T (c' = 2: bit 2 set, bit 1
// clear): ((c & 0xD9) ^ (T ? 0x11 : 0)) == 0x41  <=>  (c & 0xDF) in {A, C, G, T}  (exhaustive
// over all 256 byte values: tests/test_oracle_cross.py).  The four codes of a 32-bit group
// gather into the top byte of one multiply (no carries); pack16 collects four top bytes with
// three PRMTs.
__device__ __forceinline__ uint32_t pack4_top(uint32_t w, uint32_t &bad) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t code = (s1 & 0x03030303u) ^ (s2 & 0x01010101u);
    const uint32_t t = s2 & ~s1 & 0x01010101u;  // 1 in every byte that claims to be T
    bad |= ((w & 0xD9D9D9D9u) ^ (t * 0x11u)) ^ 0x41414141u;
    return code * 0x01041040u;  // top byte = b0 | b1<<2 | b2<<4 | b3<<6
}
__device__ __forceinline__ uint32_t pack16(uint4 v, uint32_t &bad) {
    const uint32_t p0 = pack4_top(v.x, bad), p1 = pack4_top(v.y, bad), p2 = pack4_top(v.z, bad),
                   p3 = pack4_top(v.w, bad);
    // bytes 3 of p0..p3 -> bytes 0..3
    return __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410);
}

// ---- warp reductions ----
// Wrapping 64-bit sum with three REDUX instead of a 5-level shuffle tree: the value is cut into
// 22-bit limbs; 32 lanes x (2^22 - 1) < 2^27, so every limb sum is exact in 32 bits, and
// s0 + s1 2^22 + s2 2^44 (mod 2^64) is the wrapping sum.
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
    const uint32_t l0 = (uint32_t)v & 0x3FFFFFu, l1 = (uint32_t)(v >> 22) & 0x3FFFFFu, l2 = (uint32_t)(v >> 44);
    const uint64_t s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1),
                   s2 = __reduce_add_sync(0xffffffffu, l2);
    return s0 + (s1 << 22) + (s2 << 44);
}
__device__ __forceinline__ uint64_t warp_xor64(uint64_t v) {
    const uint32_t lo = __reduce_xor_sync(0xffffffffu, (uint32_t)v), hi = __reduce_xor_sync(0xffffffffu, (uint32_t)(v >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// counter-based generator shared with the oracle (DESIGN.md §Synthetic data)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace ibu
