// barcode_agg.cu — K4 for unsorted inputs: partition, then aggregate on chip.
//
// The per-barcode table (src/parallel.rs:79-98 HashMap<barcode, count>, extended with the number
// of distinct UMI words) needs every (barcode, umi) pair de-duplicated.  A global hash table is
// one random DRAM access per record once it outgrows L2, and a full sort moves every record
// once per digit.  This path touches a record twice:
//
//   k_sample        ~1 M records at hashed positions: bit widths of the two words, distinct pairs
//                   and distinct barcodes among them (sizes every table below; tiny)
//   k_scatter_keys  24 B read + 8 B written per record.  A record whose words fit the key layout
//                   (barcode < 2^bb, umi < 2^ub, bb + ub <= 64) becomes ONE u64:
//                   key = mix64(barcode << ub | umi), mix64 a bijection, so the key's top bits
//                   are a uniform hash and the pair can be recovered from it.  The key is
//                   appended to bucket (key >> (64 - pb)) with one atomicAdd on the bucket's
//                   cursor; the 8-byte stores of a bucket fill whole lines in the write-back L2
//                   (2^pb open lines: a few MB) before they reach HBM.  Records that do not fit
//                   (`wide`: the reference's own generator writes them, examples/random.rs:46)
//                   are copied to a side list.
//   k_bucket_dedup  8 B read per record.  One CTA per bucket folds the bucket's keys into a
//                   shared-memory hash table (distinct keys per bucket are balanced by the hash,
//                   however skewed the barcodes are; duplicates only make a bucket longer) and
//                   adds every distinct pair to the per-barcode table — a global open-addressing
//                   table sized from the sample, L2 resident for realistic barcode counts — or,
//                   for pair tables, appends (barcode, umi, multiplicity).
//   wide list       de-duplicated by the legacy path, then added to the same table.
//   k_table_rows    occupied slots -> rows, radix sorted by barcode (#barcodes rows, not #records).
//
// Inputs this does not suit fall back to the legacy path (barcode_count.cu): key layouts wider
// than 64 bits, almost no distinct keys (a tiny global table is L2 resident anyway), nearly as
// many barcodes as records (the rows would need a full-size sort at the end).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <set>
#include <utility>

#include "k4.h"
#include "kernels.cuh"

namespace ibu {

namespace {

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
enum { kFlagBucket = 1, kFlagWide = 2, kFlagTable = 4, kFlagSmem = 8, kFlagPairsOut = 16 };

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

// ---------------------------------------------------------------------------------------- host

uint64_t pow2_ceil(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
uint32_t log2_of(uint64_t pow2) {
    uint32_t l = 0;
    while ((1ull << l) < pow2) l++;
    return l;
}

// Chao1 lower bound on the number of distinct values behind a sample that showed `seen` of them,
// f1 exactly once and f2 exactly twice.  A sample without repeats says nothing: `unknown`.
double chao1(double seen, double f1, double f2, double unknown) {
    if (f2 < 1.0) return f1 >= seen && seen > 64 ? unknown : seen + f1 * (f1 - 1.0) / 2.0;
    return seen + f1 * f1 / (2.0 * f2);
}

// smallest width (in bits, rounded up to even) that holds all but `slack` words of the sample
uint32_t width_of(const uint32_t *hist, uint64_t slack) {
    uint64_t above = 0;
    uint32_t w = 64;
    for (; w > 0; w--) {
        above += hist[w];
        if (above > slack) break;
    }
    w = std::max(w, 1u);
    return (w + 1u) & ~1u;
}

template <class K>
int set_max_smem(K kern, int device, size_t bytes, ibu_error_t *err) {
    // once per kernel and device (the attribute is per function, not per launch)
    // (all instantiations share this function's statics: the key is the kernel's address)
    static std::mutex m;
    static std::set<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lock(m);
    const auto key = std::make_pair((const void *)kern, device);
    if (done.count(key)) return IBU_OK;
    IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done.insert(key);
    return IBU_OK;
}

#define IBU_LAUNCHED(name)                                                   \
    do {                                                                     \
        g_launches.fetch_add(1, std::memory_order_relaxed);                  \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) return ibu::cuda_fail(err, e__, name);       \
    } while (0)

// IBU_B200_TRACE=1: device time of each stage (synchronises after every stage: tuning only)
struct StageTimer {
    bool on;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(bool enabled, cudaStream_t stream) : on(enabled), s(stream) {
        if (on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, s);
        }
    }
    ~StageTimer() {
        if (on) {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    }
    void lap(const char *what) {
        if (!on) return;
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        fprintf(stderr, "[ibu trace]   %-26s %8.3f ms (device)\n", what, ms);
        cudaEventRecord(a, s);
    }
};

constexpr size_t kDedupMaxSmem = 8192 * 16;  // 8 Ki slots x (key + 64-bit count)

template <bool W, bool P>
int launch_dedup(ibu_gpu_ctx *ctx, const DedupArgs &a, cudaStream_t s, ibu_error_t *err) {
    const size_t smem = ((size_t)1 << a.s_bits) * (8 + (W ? 8 : 4));
    if (int rc = set_max_smem(k_bucket_dedup<W, P>, ctx->device, kDedupMaxSmem, err)) return rc;
    const int per_sm = std::max<int>(1, std::min<size_t>(5, (200u << 10) / (smem + 1024)));  // 48 registers: 5 CTAs by registers
    const uint32_t grid = (uint32_t)std::min<uint64_t>(a.n_buckets, (uint64_t)ctx->sm_count * per_sm);
    k_bucket_dedup<W, P><<<grid, kBlockThreads, smem, s>>>(a);
    IBU_LAUNCHED("k_bucket_dedup");
    return IBU_OK;
}

template <bool COUNT_ONLY>
int launch_scatter(const ScatterArgs &a, bool weighted, cudaStream_t s, ibu_error_t *err) {
    const uint64_t tiles = a.n / 128 + 1;
    const uint32_t grid = (uint32_t)((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (weighted) k_scatter_keys<true, COUNT_ONLY><<<grid, kBlockThreads, 0, s>>>(a);
    else k_scatter_keys<false, COUNT_ONLY><<<grid, kBlockThreads, 0, s>>>(a);
    IBU_LAUNCHED("k_scatter_keys");
    return IBU_OK;
}

}  // namespace

int k4_sample(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, K4Sample *out, ibu_error_t *err) {
    *out = K4Sample{};
    if (n == 0) return IBU_OK;
    PoolScratch sc(s);
    const uint64_t m = std::min<uint64_t>(n, 1u << 18);
    const uint64_t slots = pow2_ceil(2 * m);
    // [pair fingerprints | barcode fingerprints] 0xFF, then [pair counts | barcode counts | outputs] zero
    uint8_t *buf;
    const size_t fp_bytes = slots * 16, zero_bytes = slots * 8 + 1024;
    IBU_CUDA(sc.alloc(&buf, fp_bytes + zero_bytes));
    IBU_CUDA(cudaMemsetAsync(buf, 0xFF, fp_bytes, s));
    IBU_CUDA(cudaMemsetAsync(buf + fp_bytes, 0, zero_bytes, s));
    uint64_t *ptab = reinterpret_cast<uint64_t *>(buf);
    uint32_t *pcnt = reinterpret_cast<uint32_t *>(buf + fp_bytes);
    unsigned long long *sout = reinterpret_cast<unsigned long long *>(buf + fp_bytes + slots * 8);
    SampleArgs a{recs, n, m, ptab, ptab + slots, pcnt, pcnt + slots, slots - 1, sout, reinterpret_cast<uint32_t *>(sout + kSmpWords)};
    const uint32_t grid = (uint32_t)std::min<uint64_t>((m + kBlockThreads - 1) / kBlockThreads, (uint64_t)ctx->sm_count * 8);
    k_sample<<<grid, kBlockThreads, 0, s>>>(a);
    IBU_LAUNCHED("k_sample");
    unsigned long long *mail = ctx->h_mail;
    IBU_CUDA(cudaMemcpyAsync(mail, sout, 1024, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    out->valid = true;
    out->m = m;
    out->pairs = (double)mail[kSmpPairs];
    out->barcodes = (double)mail[kSmpBarcodes];
    out->unordered = mail[kSmpUnordered];
    out->pair_coll = (double)mail[kSmpPairColl];
    out->pair_f1 = (double)(long long)mail[kSmpPairF1];
    out->pair_f2 = (double)(long long)mail[kSmpPairF2];
    out->bc_f1 = (double)(long long)mail[kSmpBcF1];
    out->bc_f2 = (double)(long long)mail[kSmpBcF2];
    memcpy(out->hist, mail + kSmpWords, sizeof(out->hist));
    return IBU_OK;
}

// One table build of the partition path, in three steps so that an ingest pipeline can feed it chunk by
// chunk: begin (sizes from the sample, scratch), add (scatter a chunk's keys, stream ordered, any
// stream that waited for ready()), finish (de-duplicate, rows).
struct K4Job {
    ibu_gpu_ctx *ctx;
    cudaStream_t s0;
    PoolScratch sc;
    StageTimer timer;
    bool trace;
    uint64_t n = 0;       // records the job was sized for
    uint64_t added = 0;
    uint32_t bb = 0, ub = 0, pb = 0, s_bits = 11;
    uint64_t P = 0, cap = 0, wide_cap = 0, t_slots = 0;
    double r_est = 0;
    bool exact = false, weighted = false, packed = false;
    uint32_t *cursors = nullptr;
    uint64_t *keys = nullptr, *wts = nullptr, *wide = nullptr, *bases = nullptr;
    unsigned long long *ctr = nullptr;
    cudaEvent_t ready_ev = nullptr;
    K4Job(ibu_gpu_ctx *c, cudaStream_t s, bool tr) : ctx(c), s0(s), sc(s), timer(tr, s), trace(tr) {}
    ~K4Job() {
        if (ready_ev) cudaEventDestroy(ready_ev);
    }
};

void k4_job_destroy(K4Job *job) { delete job; }
cudaEvent_t k4_job_ready(K4Job *job) { return job->ready_ev; }

int k4_job_begin(ibu_gpu_ctx *ctx, uint64_t n, const K4Hints &hints, const K4Sample &smp, bool pair_mode,
                 bool weighted, cudaStream_t s, K4Job **out, ibu_error_t *err) {
    *out = nullptr;
    const bool forced = hints.force_path == kPathPartition;
    if (!smp.valid || n == 0 || n >= (1ull << 40)) return IBU_OK;
    static const bool trace = getenv("IBU_B200_TRACE") != nullptr;

    // ---- what the sample says ----
    const double m = (double)smp.m;
    uint32_t bb, ub;
    if (hints.bc_len && hints.umi_len) {
        bb = 2 * hints.bc_len;
        ub = 2 * hints.umi_len;
    } else {  // widths that hold 97 % of the sample; what does not fit goes to the wide list
        bb = width_of(smp.hist, smp.m / 32);
        ub = width_of(smp.hist + 65, smp.m / 32);
    }
    if (bb + ub > 64 || bb >= 64 || ub >= 64) return IBU_OK;  // no single-word key: legacy path
    const bool whole = smp.m >= n;  // the sample is the input
    const double d_est = whole ? smp.pairs : std::min((double)n, chao1(smp.pairs, smp.pair_f1, smp.pair_f2, (double)n));
    const double r_est = whole ? smp.barcodes : std::min((double)n, chao1(smp.barcodes, smp.bc_f1, smp.bc_f2, (double)n));
    // sum over keys of (occurrences)^2, from the colliding pairs of the sample: spread of the bucket loads
    const double scale = (double)n / m;
    const double sum_sq = (double)n + scale * scale * 2.0 * smp.pair_coll;
    if (trace)
        fprintf(stderr, "[ibu trace] sample: m=%llu pairs %.0f (f1 %.0f f2 %.0f coll %.0f) barcodes %.0f (f1 %.0f f2 %.0f) -> D~%.3g R~%.3g bb=%u ub=%u\n",
                (unsigned long long)smp.m, smp.pairs, smp.pair_f1, smp.pair_f2, smp.pair_coll, smp.barcodes, smp.bc_f1,
                smp.bc_f2, d_est, r_est, bb, ub);
    if (!forced) {
        if (d_est < 65536.0) return IBU_OK;  // a tiny global table is L2 resident: legacy hash path
        // about as many barcodes as records: the rows would need a full-size sort (and a table far
        // outside L2); the sort-based path handles that shape
        if (!pair_mode && r_est > 8.0e6 && r_est > 0.05 * (double)n) return IBU_OK;
    }

    // ---- sizes ----
    // ~512-1024 records per bucket, whatever the data: a bucket's distinct keys then always fit the
    // shared-memory table (duplicates only lengthen a bucket)
    std::unique_ptr<K4Job> job(new K4Job(ctx, s, trace));
    job->n = n;
    job->bb = bb;
    job->ub = ub;
    job->r_est = r_est;
    job->weighted = weighted;
    job->P = std::min<uint64_t>(std::max<uint64_t>(pow2_ceil((n + 1023) / 1024), 2), 1u << 21);
    job->pb = log2_of(job->P);
    const double mean = (double)n / (double)job->P;
    const double sigma = std::sqrt(sum_sq / (double)job->P);
    // uniform layout with mean + 6 sigma room per bucket; when duplicates make the loads too uneven
    // for that (or the first attempt overflows) the buckets are laid out exactly from a histogram
    job->cap = ((uint64_t)(mean + 6.0 * sigma) + 64 + 15) & ~15ull;
    job->exact = (double)job->cap > 3.0 * mean + 256.0;
    // shared-memory table: 1.6 slots per key of the fullest bucket the uniform layout admits (all of
    // them distinct at worst); duplicate-heavy data (exact layout) has far fewer distinct keys than
    // records per bucket.  Smaller tables = more resident CTAs = more buckets in flight per SM.
    while (job->s_bits < 13 &&
           (double)(1u << job->s_bits) < 1.6 * (job->exact ? mean + 6.0 * std::sqrt(mean) : (double)job->cap))
        job->s_bits++;
    if (job->pb + job->s_bits > 60) return IBU_OK;
    job->wide_cap = n / 8 + 4096;
    job->packed = !weighted && n < (1ull << 28);
    if (!job->exact && job->P * job->cap * (weighted ? 16 : 8) > (64ull << 30)) return IBU_OK;

    IBU_CUDA(job->sc.alloc(&job->cursors, job->P * 4));
    IBU_CUDA(job->sc.alloc(&job->wide, job->wide_cap * 24));
    IBU_CUDA(job->sc.alloc(&job->ctr, kCtrWords * 8));
    IBU_CUDA(cudaMemsetAsync(job->cursors, 0, job->P * 4, s));
    IBU_CUDA(cudaMemsetAsync(job->ctr, 0, kCtrWords * 8, s));
    if (!job->exact) {
        IBU_CUDA(job->sc.alloc(&job->keys, job->P * job->cap * 8));
        if (weighted) IBU_CUDA(job->sc.alloc(&job->wts, job->P * job->cap * 8));
    }
    IBU_CUDA(cudaEventCreateWithFlags(&job->ready_ev, cudaEventDisableTiming));
    IBU_CUDA(cudaEventRecord(job->ready_ev, s));
    *out = job.release();
    return IBU_OK;
}

// Scatter (uniform layout) or count (exact layout) the keys of `cnt` records on stream s, which
// must be the job's stream or have waited for k4_job_ready().  recs must be 32-byte aligned.
int k4_job_add(K4Job *job, const uint64_t *recs, uint64_t cnt, cudaStream_t s, ibu_error_t *err) {
    if (cnt == 0) return IBU_OK;
    job->added += cnt;
    ScatterArgs a{recs, cnt, job->bb, job->ub, job->pb, (uint32_t)job->cap, nullptr, job->cursors, job->keys, job->wts,
                  job->wide, job->wide_cap, job->ctr};
    return job->exact ? launch_scatter<true>(a, job->weighted, s, err) : launch_scatter<false>(a, job->weighted, s, err);
}

// Everything after the scatter, on the job's stream (which must have waited for every stream that
// added).  all_recs (nullable): the records the job saw, contiguous — needed when the buckets have
// to be laid out exactly (duplicate-heavy data, or an overflow of the uniform layout); without
// them such a job ends unhandled.  *handled = false: nothing produced, take another path.
int k4_job_finish(K4Job *job, const uint64_t *all_recs, bool pair_mode, bool pairs_sorted, uint64_t **rows_out,
                  uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err) {
    *handled = false;
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    ibu_gpu_ctx *ctx = job->ctx;
    cudaStream_t s = job->s0;
    PoolScratch &sc = job->sc;
    StageTimer &timer = job->timer;
    const bool trace = job->trace, weighted = job->weighted, packed = job->packed;
    unsigned long long *mail = ctx->h_mail;  // pinned: device -> host read-backs without a staged copy
    const uint64_t n = job->added, P = job->P, cap = job->cap;
    const uint32_t pb = job->pb, bb = job->bb, ub = job->ub, s_bits = job->s_bits;
    uint32_t *cursors = job->cursors;
    unsigned long long *ctr = job->ctr;
    uint64_t *wide = job->wide, *pairs = nullptr;
    if (n > job->n) return IBU_OK;  // more records than the job was sized for
    timer.lap(job->exact ? "histogram" : "k_scatter_keys");

    // ---- exact layout: the cursors hold the histogram; lay the buckets out and scatter everything ----
    if (!job->exact) {
        IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        if (mail[kCtrFlags] & kFlagWide) return IBU_OK;
        if (mail[kCtrFlags] & kFlagBucket) {
            if (trace) fprintf(stderr, "[ibu trace] a bucket overflowed (cap %llu): exact layout\n", (unsigned long long)cap);
            sc.free_now(job->keys);
            if (job->wts) sc.free_now(job->wts);
            job->keys = job->wts = nullptr;
            job->exact = true;  // the cursors counted every key, stored or not
        }
    }
    if (job->exact && !job->bases) {
        if (!all_recs) return IBU_OK;
        if (n * (weighted ? 16 : 8) > (64ull << 30)) return IBU_OK;
        IBU_CUDA(sc.alloc(&job->bases, (P + 1) * 8));
        k_bucket_bases<<<1, 1024, 0, s>>>(cursors, (uint32_t)P, job->bases);
        IBU_LAUNCHED("k_bucket_bases");
        IBU_CUDA(cudaMemsetAsync(cursors, 0, P * 4, s));
        IBU_CUDA(cudaMemsetAsync(ctr, 0, kCtrWords * 8, s));
        IBU_CUDA(sc.alloc(&job->keys, n * 8));
        if (weighted) IBU_CUDA(sc.alloc(&job->wts, n * 8));
        ScatterArgs a{all_recs, n, bb, ub, pb, (uint32_t)cap, job->bases, cursors, job->keys, job->wts, wide, job->wide_cap, ctr};
        if (int rc = launch_scatter<false>(a, weighted, s, err)) return rc;
        timer.lap("k_scatter_keys (exact)");
    }
    uint64_t *keys = job->keys, *wts = job->wts, *bases = job->bases;

    // ---- per-bucket de-duplication into the barcode table (grown if the estimate was short) ----
    // barcode table: generous (skewed barcode frequencies make the estimate a lower bound); only the
    // slots that are hit occupy L2
    uint64_t t_slots = pair_mode ? 0 : std::max<uint64_t>(1u << 14, pow2_ceil((uint64_t)(8.0 * std::min(job->r_est, (double)n))));
    const uint64_t pairs_cap = pair_mode ? n : 0;
    const uint32_t slot_words = packed ? 2 : 4;
    if (pair_mode) IBU_CUDA(sc.alloc(&pairs, pairs_cap * 24));
    uint64_t *slots = nullptr;
    for (int attempt = 0;; attempt++) {
        if (!pair_mode) {
            IBU_CUDA(sc.alloc(&slots, t_slots * slot_words * 8));
            IBU_CUDA(cudaMemsetAsync(slots, 0xFF, t_slots * slot_words * 8, s));
        }
        DedupArgs a{cursors, bases, keys, wts, (uint32_t)P, (uint32_t)cap, pb, ub, s_bits,
                    TableRef{slots, t_slots - 1, ctr, packed ? 1u : 0u}, pairs, pairs_cap};
        int rc = weighted ? (pair_mode ? launch_dedup<true, true>(ctx, a, s, err) : launch_dedup<true, false>(ctx, a, s, err))
                          : (pair_mode ? launch_dedup<false, true>(ctx, a, s, err) : launch_dedup<false, false>(ctx, a, s, err));
        if (rc) return rc;
        timer.lap("k_bucket_dedup");
        IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        const uint64_t flags = mail[kCtrFlags];
        if (trace)
            fprintf(stderr, "[ibu trace] partition: P=2^%u %s cap=%llu S=2^%u table=%llu x %u B -> wide %llu pairs %llu rows %llu flags %llx\n",
                    pb, job->exact ? "exact" : "uniform", (unsigned long long)cap, s_bits, (unsigned long long)t_slots,
                    slot_words * 8, mail[kCtrWide], mail[kCtrPairs], mail[kCtrClaimed], (unsigned long long)flags);
        if (flags & (kFlagBucket | kFlagWide | kFlagSmem | kFlagPairsOut)) return IBU_OK;  // legacy path
        const bool crowded = !pair_mode && mail[kCtrClaimed] > t_slots / 10 * 6;
        if (!(flags & kFlagTable) && !crowded) break;
        if (attempt == 3 || t_slots >= (1ull << 31)) return IBU_OK;
        // more barcodes than estimated: the buckets are intact, only this stage is repeated
        sc.free_now(slots);
        t_slots *= 8;
        const unsigned long long keep[3] = {mail[kCtrWide], 0ull, mail[kCtrSpecial]};
        IBU_CUDA(cudaMemsetAsync(ctr, 0, kCtrWords * 8, s));
        IBU_CUDA(cudaMemcpyAsync(ctr, keep, sizeof(keep), cudaMemcpyHostToDevice, s));
    }
    const uint64_t n_wide = mail[kCtrWide];
    sc.free_now(keys);
    if (wts) sc.free_now(wts);
    job->keys = job->wts = nullptr;

    // ---- 5. wide records and the special key ----
    uint64_t wide_pairs = 0;
    if (n_wide || mail[kCtrSpecial]) {
        uint64_t *wrows = nullptr, wn = 0, wp = 0;
        if (n_wide)
            if (int rc = k4_legacy_unsorted(ctx, wide, n_wide, s, true, true, &wrows, &wn, &wp, err)) return rc;
        ExtraArgs a{wrows, wn, ub, TableRef{slots, t_slots - 1, ctr, packed ? 1u : 0u}, pairs, pairs_cap};
        const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((wn + kBlockThreads - 1) / kBlockThreads, (uint64_t)ctx->sm_count * 8));
        k_extra_pairs<<<grid, kBlockThreads, 0, s>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (wrows && cudaFreeAsync(wrows, s) != cudaSuccess) cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(err, e, "k_extra_pairs");
        if (mail[kCtrFlags] & (kFlagTable | kFlagPairsOut)) return IBU_OK;  // (rows already added are discarded with the table)
        wide_pairs = wn;
        timer.lap("wide list");
    }
    sc.free_now(wide);
    const uint64_t total_pairs = mail[kCtrPairs] + wide_pairs;

    // ---- 6. rows ----
    uint64_t *out = nullptr;
    if (pair_mode) {
        const uint64_t k = mail[kCtrCursor];
        IBU_CUDA(cudaMallocAsync((void **)&out, k ? k * 24 : 256, s));
        cudaError_t e = cudaSuccess;
        int rc = IBU_OK;
        if (k && pairs_sorted) {
            const uint64_t vary[3] = {n_wide ? ~0ull : (1ull << bb) - 1, n_wide ? ~0ull : (1ull << ub) - 1, 0};
            static const int order[2] = {1, 0};
            rc = k4_sort_rows(ctx, pairs, k, vary, order, 2, s, out, err);
        } else if (k) {
            e = cudaMemcpyAsync(out, pairs, k * 24, cudaMemcpyDeviceToDevice, s);
        }
        if (rc == IBU_OK && e != cudaSuccess) rc = cuda_fail(err, e, "pair rows");
        if (rc == IBU_OK && (e = cudaStreamSynchronize(s)) != cudaSuccess) rc = cuda_fail(err, e, "pair rows");
        if (rc != IBU_OK) {
            cudaFreeAsync(out, s);
            return rc;
        }
        *rows_out = out;
        *n_rows = k;
        *n_pairs = k;
        *handled = true;
        return IBU_OK;
    }
    const uint64_t R = mail[kCtrClaimed], ones = mail[kCtrOnesRec] ? 1 : 0;
    uint64_t *unsorted;
    IBU_CUDA(sc.alloc(&unsorted, R ? R * 24 : 256));
    IBU_CUDA(cudaMemsetAsync(ctr + kCtrCursor, 0, 8, s));
    {
        const uint64_t blocks = (t_slots + kBlockThreads - 1) / kBlockThreads;
        k_table_rows<<<(uint32_t)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 16), kBlockThreads, 0, s>>>(
            slots, t_slots, packed ? 1u : 0u, unsorted, ctr);
        IBU_LAUNCHED("k_table_rows");
        timer.lap("k_table_rows");
    }
    IBU_CUDA(cudaMallocAsync((void **)&out, (R + ones) ? (R + ones) * 24 : 256, s));
    int rc = IBU_OK;
    if (R) {
        const uint64_t vary[3] = {n_wide ? ~0ull : (1ull << bb) - 1, 0, 0};
        static const int order[1] = {0};
        rc = k4_sort_rows(ctx, unsorted, R, vary, order, 1, s, out, err);
        timer.lap("sort rows by barcode");
    }
    if (rc == IBU_OK && ones) {  // barcode 2^64 - 1 sorts last
        const unsigned long long row[3] = {kEmpty, mail[kCtrOnesRec], mail[kCtrOnesDist]};
        cudaError_t e = cudaMemcpyAsync(out + 3 * R, row, sizeof(row), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "cudaMemcpyAsync");
    }
    if (rc == IBU_OK) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "barcode rows");
    }
    if (rc != IBU_OK) {
        cudaFreeAsync(out, s);
        return rc;
    }
    *rows_out = out;
    *n_rows = R + ones;
    *n_pairs = total_pairs;
    *handled = true;
    return IBU_OK;
}

int k4_partition_table(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, const K4Hints &hints, const K4Sample &smp,
                       bool pair_mode, bool pairs_sorted, bool weighted, cudaStream_t s, uint64_t **rows_out,
                       uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err) {
    *handled = false;
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    K4Job *job = nullptr;
    if (int rc = k4_job_begin(ctx, n, hints, smp, pair_mode, weighted, s, &job, err)) return rc;
    if (!job) return IBU_OK;
    int rc = k4_job_add(job, recs, n, s, err);
    if (rc == IBU_OK) rc = k4_job_finish(job, recs, pair_mode, pairs_sorted, rows_out, n_rows, n_pairs, handled, err);
    k4_job_destroy(job);
    return rc;
}

}  // namespace ibu
