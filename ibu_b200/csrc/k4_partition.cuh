// k4_partition.cuh — device code of the K4 partition-then-aggregate path (see barcode_agg.cu for the
// host side and the overview).  A header so that tools/k4lab3.cu times exactly the shipped kernels.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace ibu {
namespace k4p {

constexpr uint64_t kEmpty = ~0ull;

// murmur3's 64-bit finaliser: a bijection on u64 (xor-shifts by >= 32 bits are involutions, the
// multipliers are odd), so unmix64(mix64(x)) == x.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
__host__ __device__ __forceinline__ uint64_t unmix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0x9cb4b2f8129337dbull;  // inverse of 0xc4ceb9fe1a85ec53 mod 2^64
    h ^= h >> 33;
    h *= 0x4f74430c22a54005ull;  // inverse of 0xff51afd7ed558ccd mod 2^64
    h ^= h >> 33;
    return h;
}

// counters shared by the kernels of one call
enum {
    kCtrWide = 0,      // records on the wide list
    kCtrFlags = 1,     // overflow flags, below
    kCtrSpecial = 2,   // weight of the one key whose mixed value equals the empty marker
    kCtrClaimed = 3,   // slots claimed in the barcode table = rows
    kCtrPairs = 4,     // distinct pairs seen so far
    kCtrCursor = 5,    // append cursor (pair output / row output)
    kCtrOnesRec = 6,   // n_records of barcode 0xFFFF'FFFF'FFFF'FFFF (collides with the table's empty marker)
    kCtrOnesDist = 7,  // n_distinct_umi of that barcode
    kCtrTail = 8,      // ordered path: wide rows appended after the last bucket
    kCtrWords = 16
};
enum { kFlagBucket = 1, kFlagWide = 2, kFlagTable = 4, kFlagSmem = 8, kFlagPairsOut = 16, kFlagLevel = 32 };

// ------------------------------------------------------------------------------------ sample
struct SampleArgs {
    const uint64_t *recs;
    uint64_t n, m;
    uint64_t *ptab;  // fingerprints of sampled pairs, memset to 0xFF
    uint64_t *btab;  // fingerprints of sampled barcodes
    uint32_t *pcnt, *bcnt;  // occurrences per slot, zeroed
    uint64_t mask;   // slots - 1 of both
    unsigned long long *out;  // kSmp* words
    uint32_t *hist;           // [2][65]: bit width of barcode / umi words
};
enum { kSmpPairs = 0, kSmpBarcodes = 1, kSmpUnordered = 2, kSmpPairColl = 3, kSmpPairF1 = 4, kSmpPairF2 = 5,
       kSmpBcF1 = 6, kSmpBcF2 = 7, kSmpBcMax = 8 /* most sampled records under one barcode */, kSmpWords = 9 };

// Inserts a fingerprint and returns how often it had been seen before (0 = new).
__device__ __forceinline__ uint32_t fp_insert(uint64_t *tab, uint32_t *cnt, uint64_t mask, uint64_t fp) {
    if (fp == kEmpty) fp = 0;
    uint64_t slot = fp & mask;
    for (uint32_t probe = 0; probe < 4096; probe++, slot = (slot + 1) & mask) {
        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long *>(tab + slot), kEmpty, fp);
        if (old == kEmpty || old == fp) return atomicAdd(cnt + slot, 1u);
    }
    return 0;
}

// Seen-once / seen-twice bookkeeping for the Chao1 estimate: `before` occurrences existed.
__device__ __forceinline__ void tally(uint32_t before, uint32_t &distinct, int32_t &f1, int32_t &f2) {
    if (before == 0) { distinct++; f1++; }
    else if (before == 1) { f1--; f2++; }
    else if (before == 2) { f2--; }
}

__global__ void __launch_bounds__(kBlockThreads) k_sample(const SampleArgs a) {
    __shared__ uint32_t h[2][65];
    for (uint32_t i = threadIdx.x; i < 130; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t np = 0, nb = 0, bad = 0, coll = 0, bc_max = 0;
    int32_t pf1 = 0, pf2 = 0, bf1 = 0, bf2 = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.m; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = a.m >= a.n ? j : mix64(j ^ 0x5bd1e995u) % a.n;  // hashed positions: no aliasing with periodic data
        const uint64_t bc = a.recs[3 * i], um = a.recs[3 * i + 1];
        atomicAdd(&h[0][bc ? 64 - __clzll((long long)bc) : 0], 1u);
        atomicAdd(&h[1][um ? 64 - __clzll((long long)um) : 0], 1u);
        if (i + 1 < a.n) {
            const uint64_t b2 = a.recs[3 * i + 3], u2 = a.recs[3 * i + 4];
            bad += (b2 < bc) | ((b2 == bc) & (u2 < um));
        }
        const uint32_t before = fp_insert(a.ptab, a.pcnt, a.mask, mix64(bc ^ mix64(um + 0x9E3779B97F4A7C15ull)));
        coll += before;
        tally(before, np, pf1, pf2);
        const uint32_t bc_before = fp_insert(a.btab, a.bcnt, a.mask, mix64(bc));
        tally(bc_before, nb, bf1, bf2);
        bc_max = max(bc_max, bc_before + 1u);
    }
    bc_max = __reduce_max_sync(0xffffffffu, bc_max);
    if ((threadIdx.x & 31u) == 0 && bc_max > 1u) atomicMax(a.out + kSmpBcMax, (unsigned long long)bc_max);
    const uint32_t vals[8] = {np, nb, bad, coll, (uint32_t)pf1, (uint32_t)pf2, (uint32_t)bf1, (uint32_t)bf2};
#pragma unroll
    for (int k = 0; k < 8; k++) {  // (f1 / f2 deltas may be negative: two's-complement sums are exact)
        const uint32_t v = __reduce_add_sync(0xffffffffu, vals[k]);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(a.out + k, (unsigned long long)(long long)(int32_t)v);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 130; i += blockDim.x)
        if ((&h[0][0])[i]) atomicAdd(a.hist + i, (&h[0][0])[i]);
}

// ------------------------------------------------------------------------------------ layout
// bases[b] = sum of counts[0..b) for b in 0..n (one CTA; n <= 2^21 buckets).  4096 counts per round:
// a thread takes four consecutive ones (coalesced 16-byte loads; a thread walking its own 128-count
// stretch cost 0.23 ms for 2^17 buckets), a shuffle scan per warp, eight warp totals through shared memory.
__global__ void __launch_bounds__(1024) k_bucket_bases(const uint32_t *__restrict__ counts, uint32_t n,
                                                       uint64_t *__restrict__ bases) {
    __shared__ uint64_t wsum[2][32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint64_t carry = 0;
    uint32_t par = 0;
    for (uint32_t at = 0; at < n; at += 4096, par ^= 1u) {
        const uint32_t i = at + 4 * tid;
        uint32_t c[4];
#pragma unroll
        for (int q = 0; q < 4; q++) c[q] = i + q < n ? counts[i + q] : 0u;  // (n is a power of two >= 2: whole uint4s in practice)
        const uint64_t mine = (uint64_t)c[0] + c[1] + c[2] + c[3];
        uint64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) wsum[par][warp] = inc;
        __syncthreads();  // (alternating scratch: one barrier per round)
        uint64_t before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            const uint64_t t = wsum[par][w];
            if ((uint32_t)w < warp) before += t;
            all += t;
        }
        uint64_t run = carry + before + inc - mine;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (i + q < n) bases[i + q] = run;
            run += c[q];
        }
        carry += all;
    }
    if (tid == 0) bases[n] = carry;
}

// ------------------------------------------------------------------------------ barcode table
// Global open-addressing table keyed by barcode, memset to 0xFF; a slot is claimed with a 64-bit
// CAS on the barcode and the counters take REDs.
//   packed (fewer than 2^28 unweighted records): 16-byte slots {barcode, n_distinct << 36 | n_records},
//     ONE RED per distinct pair (the word starts at -1: low 36 bits end at n_records - 1);
//   wide: 32-byte slots {barcode, n_records - 1, n_distinct - 1, unused}, two REDs.
struct TableRef {
    uint64_t *slots;
    uint64_t mask;
    unsigned long long *ctr;
    uint32_t packed;
};
constexpr uint32_t kPackShift = 36;

// the counters of a slot that is known to hold the barcode
__device__ __forceinline__ void table_hit(const TableRef &t, uint64_t slot, uint64_t n_rec, uint64_t n_dist) {
    uint64_t *s = t.slots + (t.packed ? 2 : 4) * slot;
    if (t.packed) {
        atomicAdd(reinterpret_cast<unsigned long long *>(s + 1), (unsigned long long)(n_rec + (n_dist << kPackShift)));
    } else {
        atomicAdd(reinterpret_cast<unsigned long long *>(s + 1), (unsigned long long)n_rec);
        atomicAdd(reinterpret_cast<unsigned long long *>(s + 2), (unsigned long long)n_dist);
    }
}

__device__ __forceinline__ void table_add(const TableRef &t, uint64_t bc, uint64_t n_rec, uint64_t n_dist) {
    if (bc == kEmpty) {  // only a wide record can carry it
        atomicAdd(t.ctr + kCtrOnesRec, (unsigned long long)n_rec);
        atomicAdd(t.ctr + kCtrOnesDist, (unsigned long long)n_dist);
        return;
    }
    const uint32_t words = t.packed ? 2 : 4;
    uint64_t slot = mix64(bc) & t.mask;
    for (uint32_t probe = 0; probe < 96; probe++, slot = (slot + 1) & t.mask) {  // a crowded table fails fast
        uint64_t *s = t.slots + words * slot;
        uint64_t cur = *reinterpret_cast<volatile uint64_t *>(s);
        if (cur != bc) {
            if (cur != kEmpty) continue;  // another barcode lives here (a slot never changes once claimed)
            cur = atomicCAS(reinterpret_cast<unsigned long long *>(s), kEmpty, bc);
            if (cur == kEmpty) atomicAdd(t.ctr + kCtrClaimed, 1ull);
            else if (cur != bc) continue;
        }
        if (t.packed) {
            atomicAdd(reinterpret_cast<unsigned long long *>(s + 1), (unsigned long long)(n_rec + (n_dist << kPackShift)));
        } else {
            atomicAdd(reinterpret_cast<unsigned long long *>(s + 1), (unsigned long long)n_rec);
            atomicAdd(reinterpret_cast<unsigned long long *>(s + 2), (unsigned long long)n_dist);
        }
        return;
    }
    atomicOr(t.ctr + kCtrFlags, (unsigned long long)kFlagTable);
}

// ------------------------------------------------------------------------------------- dedup
struct DedupArgs {
    const uint32_t *cursors;
    const uint64_t *bases;  // nullable (uniform layout: bucket b starts at b * cap)
    const uint64_t *keys;
    const uint64_t *wts;
    uint32_t n_buckets, cap, pb, ub;
    uint32_t s_bits;  // log2(slots of the shared-memory table)
    TableRef table;   // table mode
    uint64_t *pairs_out;  // pair mode: rows {barcode, umi, multiplicity}
    uint64_t pairs_cap;
};

// (barcode, umi, multiplicity) rows from outside the buckets — the de-duplicated wide list and the
// one special key — folded into the same table / pair output.
struct ExtraArgs {
    const uint64_t *rows;  // nullable
    uint64_t n;
    uint32_t ub;
    TableRef table;
    uint64_t *pairs_out;  // nullable: table mode
    uint64_t pairs_cap;
};

__global__ void __launch_bounds__(kBlockThreads) k_extra_pairs(const ExtraArgs a) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = gid; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t bc = a.rows[3 * i], um = a.rows[3 * i + 1], c = a.rows[3 * i + 2];
        if (a.pairs_out) {
            const uint64_t pos = atomicAdd(a.table.ctr + kCtrCursor, 1ull);
            if (pos < a.pairs_cap) {
                a.pairs_out[3 * pos] = bc;
                a.pairs_out[3 * pos + 1] = um;
                a.pairs_out[3 * pos + 2] = c;
            } else {
                atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagPairsOut);
            }
        } else {
            table_add(a.table, bc, c, 1ull);
        }
    }
    if (gid == 0) {
        const uint64_t w = a.table.ctr[kCtrSpecial];
        if (w) {  // the key whose mixed value is the empty marker
            const uint64_t comp = unmix64(kEmpty);
            const uint64_t bc = comp >> a.ub, um = comp & ((1ull << a.ub) - 1ull);
            atomicAdd(a.table.ctr + kCtrPairs, 1ull);
            if (a.pairs_out) {
                const uint64_t pos = atomicAdd(a.table.ctr + kCtrCursor, 1ull);
                if (pos < a.pairs_cap) {
                    a.pairs_out[3 * pos] = bc;
                    a.pairs_out[3 * pos + 1] = um;
                    a.pairs_out[3 * pos + 2] = w;
                } else {
                    atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagPairsOut);
                }
            } else {
                table_add(a.table, bc, w, 1ull);
            }
        }
    }
}

}  // namespace k4p
}  // namespace ibu

#include "k4_staged.cuh"
#include "k4_ordered.cuh"
