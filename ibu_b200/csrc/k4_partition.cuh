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
       kSmpBcF1 = 6, kSmpBcF2 = 7, kSmpWords = 8 };

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
    uint32_t np = 0, nb = 0, bad = 0, coll = 0;
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
        tally(fp_insert(a.btab, a.bcnt, a.mask, mix64(bc)), nb, bf1, bf2);
    }
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

// ------------------------------------------------------------------------------------ scatter
struct ScatterArgs {
    const uint64_t *recs;
    uint64_t n;
    uint32_t bb, ub, pb;  // barcode bits, umi bits (bb + ub <= 64, both in 1..63), log2(#buckets)
    uint32_t cap;         // keys per bucket (uniform layout)
    const uint64_t *bases;  // nullable: exact layout, bucket b owns keys[bases[b] .. bases[b + 1])
    uint32_t *cursors;    // [2^pb], zeroed
    uint64_t *keys;
    uint64_t *wts;        // same shape (WEIGHTED only)
    uint64_t *wide;       // records that do not fit the key layout
    uint64_t wide_cap;
    unsigned long long *ctr;
};

// COUNT_ONLY: the histogram pass of the exact layout (cursors[b] = keys of bucket b, nothing stored).
template <bool WEIGHTED, bool COUNT_ONLY>
__device__ __forceinline__ void scatter_one(const ScatterArgs &a, uint64_t bc, uint64_t um, uint64_t w) {
    if (((bc >> a.bb) | (um >> a.ub)) == 0ull) {
        const uint64_t k = mix64((bc << a.ub) | um);
        if (k == kEmpty) {  // the one key that looks like an empty slot
            if (!COUNT_ONLY) atomicAdd(a.ctr + kCtrSpecial, (unsigned long long)(WEIGHTED ? w : 1ull));
            return;
        }
        const uint32_t b = (uint32_t)(k >> (64 - a.pb));
        const uint32_t pos = atomicAdd(a.cursors + b, 1u);
        if (COUNT_ONLY) return;
        uint64_t base = (uint64_t)b * a.cap, room = a.cap;
        if (a.bases) {
            base = a.bases[b];
            room = a.bases[b + 1] - base;
        }
        if (pos < room) {
            a.keys[base + pos] = k;
            if (WEIGHTED) a.wts[base + pos] = w;
        } else {
            atomicOr(a.ctr + kCtrFlags, (unsigned long long)kFlagBucket);
        }
    } else if (!COUNT_ONLY) {
        const uint64_t pos = atomicAdd(a.ctr + kCtrWide, 1ull);
        if (pos < a.wide_cap) {
            a.wide[3 * pos] = bc;
            a.wide[3 * pos + 1] = um;
            a.wide[3 * pos + 2] = WEIGHTED ? w : 1ull;
        } else {
            atomicOr(a.ctr + kCtrFlags, (unsigned long long)kFlagWide);
        }
    }
}

// One 128-record tile per warp (lane l owns records 4l..4l+3: three LDG.E.256), block-scheduled
// like K1-K3.  Four independent atomics + stores per lane are in flight at a time.  The kernel is
// bound by its 8-byte scattered stores (one L2 write transaction each: 10^8 of them take 2.0 ms on
// B200 whatever the bucket count, tools/k4lab.cu), not by the atomics (1.1 ms at 2^17 cursors).
template <bool WEIGHTED, bool COUNT_ONLY>
__global__ void __launch_bounds__(kBlockThreads) k_scatter_keys(const ScatterArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t t = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint64_t n_tiles = a.n / 128;
    if (t < n_tiles) {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(a.recs) + t * (128 * 24) + lane * 96;
        const u64x4 v0 = ldg_stream256(p), v1 = ldg_stream256(p + 32), v2 = ldg_stream256(p + 64);
        scatter_one<WEIGHTED, COUNT_ONLY>(a, v0.x, v0.y, v0.z);
        scatter_one<WEIGHTED, COUNT_ONLY>(a, v0.w, v1.x, v1.y);
        scatter_one<WEIGHTED, COUNT_ONLY>(a, v1.z, v1.w, v2.x);
        scatter_one<WEIGHTED, COUNT_ONLY>(a, v2.y, v2.z, v2.w);
    } else if (t == n_tiles) {  // ragged tail (< 128 records)
        for (uint64_t i = n_tiles * 128 + lane; i < a.n; i += 32)
            scatter_one<WEIGHTED, COUNT_ONLY>(a, a.recs[3 * i], a.recs[3 * i + 1], a.recs[3 * i + 2]);
    }
}

// bases[b] = sum of counts[0..b) for b in 0..n (one CTA; n <= 2^21 buckets)
__global__ void __launch_bounds__(1024) k_bucket_bases(const uint32_t *__restrict__ counts, uint32_t n,
                                                       uint64_t *__restrict__ bases) {
    __shared__ uint64_t part[1024];
    const uint32_t tid = threadIdx.x, per = (n + 1023) / 1024;
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    uint64_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += counts[i];
    part[tid] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint64_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint64_t run = part[tid] - sum;
    for (uint32_t i = lo; i < hi; i++) {
        bases[i] = run;
        run += counts[i];
    }
    if (tid == 1023) bases[n] = part[1023];
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

// One CTA per bucket (block-strided over the buckets).  The table holds the bucket's DISTINCT keys
// (load <= ~0.5 by construction); its slot index comes from the key bits just below the bucket
// bits, which are as uniform as the bucket bits.
template <bool WEIGHTED, bool PAIRS>
__global__ void __launch_bounds__(kBlockThreads, 5) k_bucket_dedup(const DedupArgs a) {
    using Cnt = typename std::conditional<WEIGHTED, unsigned long long, uint32_t>::type;
    extern __shared__ __align__(16) unsigned long long smem[];
    const uint32_t S = 1u << a.s_bits, smask = S - 1u;
    unsigned long long *tkey = smem;
    Cnt *tcnt = reinterpret_cast<Cnt *>(smem + S);
    __shared__ uint32_t s_distinct, s_full;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t umask = (1ull << a.ub) - 1ull;
    const uint32_t hshift = 64 - a.pb - a.s_bits;
    constexpr uint32_t kMaxProbe = 128;  // at the design load (<= 0.6) a probe sequence this long does not occur
    constexpr uint32_t kBatch = 4 * kBlockThreads;

    // the first kBatch keys of a bucket, one batch of loads per thread; issued one bucket ahead so
    // that their latency (and the cursor's) hides behind the bucket being folded
    auto fetch = [&](uint32_t b, uint32_t &cnt, uint64_t &first, uint64_t (&k)[4], uint64_t (&w)[4]) {
        cnt = 0;
        first = 0;
        if (b < a.n_buckets) {
            first = a.bases ? a.bases[b] : (uint64_t)b * a.cap;
            cnt = min(a.cursors[b], a.bases ? (uint32_t)(a.bases[b + 1] - first) : a.cap);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + threadIdx.x;
            k[q] = i < cnt ? ldg_stream64(a.keys + first + i) : kEmpty;
            w[q] = (WEIGHTED && i < cnt) ? ldg_stream64(a.wts + first + i) : 1ull;
        }
    };
    uint32_t fresh = 0;
    auto insert = [&](uint64_t k, uint64_t w) {
        if (k == kEmpty) return;
        uint32_t slot = (uint32_t)(k >> hshift) & smask;
        uint32_t probe = 0;
        for (; probe < kMaxProbe; probe++, slot = (slot + 1) & smask) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tkey + slot);
            if (cur != k) {
                if (cur != kEmpty) continue;  // a slot never changes once claimed
                cur = atomicCAS(tkey + slot, kEmpty, (unsigned long long)k);
                if (cur == kEmpty) {
                    fresh++;
                    // unweighted: the counter holds the occurrences AFTER the first, so claiming a
                    // slot is the only atomic of a new key and a repeat costs one add: one shared-
                    // memory atomic per record (the unit's rate, ~0.5 per clock per SM, is what
                    // bounds this kernel)
                    if (!WEIGHTED) return;
                } else if (cur != k) {
                    continue;
                }
            }
            atomicAdd(tcnt + slot, (Cnt)w);
            return;
        }
        s_full = 1;  // (many) more distinct keys than the table was sized for
    };

    uint32_t cnt_n;
    uint64_t first_n, kn[4], wn[4];
    fetch(blockIdx.x, cnt_n, first_n, kn, wn);
    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x) {
        const uint32_t cnt = cnt_n;
        const uint64_t first = first_n;
        uint64_t k[4], w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { k[q] = kn[q]; w[q] = wn[q]; }
        fetch(b + gridDim.x, cnt_n, first_n, kn, wn);
        for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
            tkey[i] = kEmpty;
            tcnt[i] = 0;
        }
        if (threadIdx.x == 0) s_distinct = 0, s_full = 0;
        __syncthreads();
        fresh = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) insert(k[q], w[q]);
        for (uint32_t base = kBatch; base < cnt; base += kBatch) {  // long buckets (duplicate-heavy data)
            if (*reinterpret_cast<volatile uint32_t *>(&s_full)) break;  // the call is void anyway: do not crawl a full table
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t i = base + q * kBlockThreads + threadIdx.x;
                k[q] = i < cnt ? ldg_stream64(a.keys + first + i) : kEmpty;
                w[q] = (WEIGHTED && i < cnt) ? ldg_stream64(a.wts + first + i) : 1ull;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) insert(k[q], w[q]);
        }
        fresh = __reduce_add_sync(0xffffffffu, fresh);
        if (lane == 0 && fresh) atomicAdd(&s_distinct, fresh);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_distinct) atomicAdd(a.table.ctr + kCtrPairs, (unsigned long long)s_distinct);
            if (s_full) atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagSmem);
        }
        // every distinct pair of the bucket: one row (pair tables) or one add to its barcode's row
        if (PAIRS) {
            for (uint32_t base = 0; base < S; base += blockDim.x) {  // warp-uniform trip count
                const uint32_t i = base + threadIdx.x;
                const unsigned long long key = tkey[i];
                const bool live = key != kEmpty;
                const uint32_t m = __ballot_sync(0xffffffffu, live);
                if (!m) continue;
                unsigned long long pos = 0;
                if (lane == 0) pos = atomicAdd(a.table.ctr + kCtrCursor, (unsigned long long)__popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
                if (live) {
                    const uint64_t comp = unmix64(key);
                    if (pos < a.pairs_cap) {
                        a.pairs_out[3 * pos] = comp >> a.ub;
                        a.pairs_out[3 * pos + 1] = comp & umask;
                        a.pairs_out[3 * pos + 2] = (uint64_t)tcnt[i] + (WEIGHTED ? 0ull : 1ull);
                    } else {
                        atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagPairsOut);
                    }
                }
            }
        } else {
            // Four slots per thread at a time: the home slots of their barcodes are read from the
            // global table together (independent loads in flight), then resolved — a dependent
            // load per pair would expose its latency once per pair.
            const uint32_t words = a.table.packed ? 2 : 4;
#ifdef K4LAB_SKIP_TABLE  // tools/k4lab3.cu: the kernel without its global-table traffic
            if (a.table.mask == 0)
                for (uint32_t i = threadIdx.x; i < S; i += blockDim.x)
                    if (tkey[i] != kEmpty && tcnt[i] == 0x7fffffff) atomicAdd(a.table.ctr + kCtrOnesRec, 1ull);
            if (a.table.mask != 0)
#endif
            for (uint32_t base = 0; base < S; base += 4 * blockDim.x) {
                uint64_t bc[4], seen[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t i = base + q * blockDim.x + threadIdx.x;
                    const unsigned long long key = i < S ? tkey[i] : kEmpty;
                    bc[q] = kEmpty;
                    if (key != kEmpty) bc[q] = unmix64(key) >> a.ub;  // (a narrow barcode is never all ones)
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    seen[q] = bc[q] != kEmpty
                                  ? *reinterpret_cast<volatile uint64_t *>(a.table.slots + words * (mix64(bc[q]) & a.table.mask))
                                  : 0ull;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (bc[q] == kEmpty) continue;
                    const uint32_t i = base + q * blockDim.x + threadIdx.x;
                    const uint64_t c = (uint64_t)tcnt[i] + (WEIGHTED ? 0ull : 1ull);
                    if (seen[q] == bc[q]) table_hit(a.table, mix64(bc[q]) & a.table.mask, c, 1ull);  // the common case
                    else table_add(a.table, bc[q], c, 1ull);
                }
            }
        }
        __syncthreads();
    }
}

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

// occupied slots -> rows {barcode, n_records, n_distinct_umi}, order unspecified
__global__ void __launch_bounds__(kBlockThreads)
k_table_rows(const uint64_t *__restrict__ slots, uint64_t n_slots, uint32_t packed, uint64_t *__restrict__ rows,
             unsigned long long *__restrict__ ctr) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n_slots; base += step) {  // warp-uniform trips
        const uint64_t i = base + threadIdx.x;
        uint64_t bc = kEmpty, nr = 0, nd = 0;
        if (i < n_slots) {
            if (packed) {
                const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(slots)[i];
                bc = v.x;
                nr = (v.y & ((1ull << kPackShift) - 1)) + 1ull;
                nd = v.y >> kPackShift;
            } else {
                const u64x4 v = ldg_stream256(slots + 4 * i);
                bc = v.x; nr = v.y + 1ull; nd = v.z + 1ull;
            }
        }
        const bool live = bc != kEmpty;
        const uint32_t m = __ballot_sync(0xffffffffu, live);
        if (!m) continue;
        unsigned long long pos = 0;
        if (lane == 0) pos = atomicAdd(ctr + kCtrCursor, (unsigned long long)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
        if (live) {
            rows[3 * pos] = bc;
            rows[3 * pos + 1] = nr;
            rows[3 * pos + 2] = nd;
        }
    }
}

}  // namespace k4p
}  // namespace ibu

#include "k4_staged.cuh"
