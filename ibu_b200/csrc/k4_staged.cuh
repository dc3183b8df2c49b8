// k4_staged.cuh — the staged two-level partition of the K4 unsorted path (included by
// k4_partition.cuh, after the counters / mix64 it uses).
//
// k_scatter_keys pays one L2 write transaction per 8-byte key (10^8 scattered 8-byte stores take
// 2.0 ms on B200 whatever the bucket count, tools/k4lab.cu).  The staged form groups a CTA tile's
// keys by bucket in shared memory first, so that a bucket's keys leave as one run (tile / fan-out
// keys: 16 at 4096 / 256 = a full 128-byte line), and reaches 2^17 buckets in two levels of
// fan-out 2^pb1 x 2^pb2.  Ranks inside the tile come from shared-memory atomicAdd(+1)
// (ATOMS.POPC.INC: 0.12 clocks per key and SM measured, tools/k4lab2.cu — 8x cheaper than ballot
// matching); one global atomicAdd per (tile, bucket) reserves the run.
#pragma once

namespace ibu {
namespace k4p {

constexpr int kPartTile = 4096;                      // keys per CTA tile
constexpr int kPartKpt = kPartTile / kBlockThreads;  // 16 keys per thread
constexpr int kPartMaxFan = 512;
#ifndef K4_PART_MINB
#define K4_PART_MINB 3  // resident CTAs per SM the partition kernels are compiled for (tools/k4lab3.cu sweeps it)
#endif
#ifndef K4_PART_MINB_W
#define K4_PART_MINB_W 2  // the same for the weighted instantiations (keys and a second word: 64 KB of staging per CTA);
                          // 3 (80 registers, 140 - 280 bytes spilled) measured: k_part1 1.03 -> 1.17 ms, k_part2 0.89 -> 1.09 per 10^8
#endif

struct Part1Args {
    const uint64_t *recs;
    uint64_t n;
    uint32_t bb, ub, pb1;  // level-1 fan-out 2^pb1 <= 256
    uint64_t cap1;         // keys per level-1 bucket (uniform layout; < 2^32)
    uint32_t *cursors1;    // [2^pb1], zeroed
    uint64_t *keys1;       // [2^pb1][cap1]
    uint64_t *wts1;        // same shape (WEIGHTED only)
    uint64_t *wide;        // records that do not fit the key layout
    uint64_t wide_cap;
    unsigned long long *ctr;
    uint32_t ordered;      // 1: order-preserving keys, (barcode << ub | umi) left-aligned (k4_ordered.cuh), instead of mixed ones
};

// block-wide exclusive scan of one value per thread (256 threads); `tmp` holds 8 words
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *tmp, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; w++) {
        const uint32_t t = tmp[w];
        if ((uint32_t)w < warp) base += t;
        all += t;
    }
    total = all;
    return base + inc - v;
}

// Level 1: records -> keys grouped by the top pb1 bits of the mixed key.
template <bool WEIGHTED>
__global__ void __launch_bounds__(kBlockThreads, WEIGHTED ? K4_PART_MINB_W : K4_PART_MINB) k_part1(const Part1Args a) {
    extern __shared__ __align__(16) unsigned long long stage[];  // [kPartTile] keys (+ [kPartTile] weights)
    __shared__ uint32_t hist[256], start[256], delta[256], tmp[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t F = 1u << a.pb1, shift = 63 - a.pb1;  // bucket = (key >> 1) >> shift: pb1 = 0 is one bucket
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t tile0 = (uint64_t)blockIdx.x * kPartTile;
    uint64_t key[kPartKpt], wt[WEIGHTED ? kPartKpt : 1];
    uint32_t rank[kPartKpt];
#pragma unroll
    for (int r = 0; r < kPartKpt / 4; r++) {
        const uint64_t first = tile0 + warp * (kPartTile / kWarpsPerBlock) + r * 128 + lane * 4;
        uint64_t bc[4], um[4], w[4];
        if (first + 4 <= a.n) {
            const uint8_t *p = reinterpret_cast<const uint8_t *>(a.recs) + first * 24;
            const u64x4 v0 = ldg_stream256(p), v1 = ldg_stream256(p + 32), v2 = ldg_stream256(p + 64);
            bc[0] = v0.x; um[0] = v0.y; w[0] = v0.z;
            bc[1] = v0.w; um[1] = v1.x; w[1] = v1.y;
            bc[2] = v1.z; um[2] = v1.w; w[2] = v2.x;
            bc[3] = v2.y; um[3] = v2.z; w[3] = v2.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool in = first + q < a.n;
                bc[q] = in ? a.recs[3 * (first + q)] : 0;
                um[q] = in ? a.recs[3 * (first + q) + 1] : 0;
                w[q] = in ? a.recs[3 * (first + q) + 2] : 0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = r * 4 + q;
            const bool valid = first + q < a.n;
            const bool fits = valid && ((bc[q] >> a.bb) | (um[q] >> a.ub)) == 0ull;
            uint64_t k = kEmpty;  // kEmpty = not staged
            if (fits) {
                const uint64_t raw = (bc[q] << a.ub) | um[q];
                k = a.ordered ? raw << (64u - a.bb - a.ub) : mix64(raw);
                if (k == kEmpty) atomicAdd(a.ctr + kCtrSpecial, (unsigned long long)(WEIGHTED ? w[q] : 1ull));
            }
            // records that do not fit: one reservation per warp (10^6 of them in 10^8 records are 10^6 atomics
            // on one address otherwise: 0.3 ms)
            const uint32_t wmask = __ballot_sync(0xffffffffu, valid && !fits);
            if (wmask) {
                const uint32_t leader = (uint32_t)__ffs((int)wmask) - 1u;
                unsigned long long at = 0;
                if (lane == leader) at = atomicAdd(a.ctr + kCtrWide, (unsigned long long)__popc(wmask));
                at = __shfl_sync(0xffffffffu, at, leader);
                if (valid && !fits) {
                    const uint64_t pos = at + __popc(wmask & ((1u << lane) - 1u));
                    if (pos < a.wide_cap) {
                        a.wide[3 * pos] = bc[q];
                        a.wide[3 * pos + 1] = um[q];
                        a.wide[3 * pos + 2] = WEIGHTED ? w[q] : 1ull;
                    } else {
                        atomicOr(a.ctr + kCtrFlags, (unsigned long long)kFlagWide);
                    }
                }
            }
            key[j] = k;
            if (WEIGHTED) wt[j] = w[q];
            rank[j] = k != kEmpty ? atomicAdd(&hist[(uint32_t)((k >> 1) >> shift)], 1u) : 0u;
        }
    }
    __syncthreads();
    const uint32_t c = threadIdx.x < F ? hist[threadIdx.x] : 0u;
    uint32_t total;
    const uint32_t st = block_excl_scan(c, tmp, total);
    start[threadIdx.x] = st;
    uint32_t g = 0;
    if (c) {
        g = atomicAdd(a.cursors1 + threadIdx.x, c);
        if ((uint64_t)g + c > a.cap1) atomicOr(a.ctr + kCtrFlags, (unsigned long long)kFlagLevel);
    }
    delta[threadIdx.x] = g - st;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartKpt; j++) {
        if (key[j] != kEmpty) {
            const uint32_t p = start[(uint32_t)((key[j] >> 1) >> shift)] + rank[j];
            stage[p] = key[j];
            if (WEIGHTED) stage[kPartTile + p] = wt[j];
        }
    }
    __syncthreads();
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < total; i += kBlockThreads) {
        const uint64_t k = stage[i];
        const uint32_t b = (uint32_t)((k >> 1) >> shift);
        const uint32_t p = delta[b] + i;
        if (p < a.cap1) {
            a.keys1[(uint64_t)b * a.cap1 + p] = k;
            if (WEIGHTED) a.wts1[(uint64_t)b * a.cap1 + p] = stage[kPartTile + i];
        }
    }
}

struct Part2Args {
    const uint32_t *cursors_in;  // keys in each input bucket
    const uint64_t *keys_in, *wts_in;
    uint64_t cap_in;             // input layout: bucket b at keys_in[b * cap_in]
    uint32_t tiles_per_bucket;   // ceil(cap_in / kPartTile): the grid is [input buckets x tiles], tiles fastest
    uint32_t pb_in, pb;          // bits consumed so far, bits of this level (2^pb <= 512)
    uint64_t cap_out;            // keys per output bucket (uniform layout)
    const uint64_t *bases_out;   // nullable: exact layout, bucket b owns keys_out[bases_out[b] .. bases_out[b + 1])
    uint32_t *cursors_out;       // [2^(pb_in + pb)], zeroed
    uint64_t *keys_out, *wts_out;
    unsigned long long *ctr;
    uint32_t over_flag;          // what an overflowing output bucket raises (kFlagBucket / kFlagLevel)
};

// A further level: the keys of one input bucket -> 2^pb output buckets each.  COUNT_ONLY: the
// histogram of the exact layout (cursors_out[b] = keys of bucket b; nothing stored).
template <bool WEIGHTED, bool COUNT_ONLY>
__global__ void __launch_bounds__(kBlockThreads, WEIGHTED ? K4_PART_MINB_W : K4_PART_MINB) k_part2(const Part2Args a) {
    extern __shared__ __align__(16) unsigned long long stage[];
    __shared__ uint32_t hist[kPartMaxFan], start[kPartMaxFan], delta[kPartMaxFan], tmp[8];
    const uint32_t b1 = blockIdx.x / a.tiles_per_bucket;
    const uint64_t cnt = min((uint64_t)a.cursors_in[b1], a.cap_in);
    const uint64_t tile0 = (uint64_t)(blockIdx.x % a.tiles_per_bucket) * kPartTile;
    if (tile0 >= cnt) return;
    const uint32_t F = 1u << a.pb, shift = 64 - a.pb_in - a.pb, fmask = F - 1u;
    hist[threadIdx.x] = 0;
    hist[threadIdx.x + kBlockThreads] = 0;
    __syncthreads();
    const uint64_t *src = a.keys_in + (uint64_t)b1 * a.cap_in + tile0;
    const uint64_t *wsrc = WEIGHTED ? a.wts_in + (uint64_t)b1 * a.cap_in + tile0 : nullptr;
    const uint32_t here = (uint32_t)min((uint64_t)kPartTile, cnt - tile0);
    uint64_t key[kPartKpt], wt[WEIGHTED ? kPartKpt : 1];
    uint32_t rank[kPartKpt];
    // two consecutive keys per thread and load (cap_in is a multiple of 16 keys: 16-byte aligned)
#pragma unroll
    for (int j = 0; j < kPartKpt; j += 2) {
        const uint32_t i = (j / 2) * (2 * kBlockThreads) + 2 * threadIdx.x;
        ulonglong2 v = make_ulonglong2(kEmpty, kEmpty), w = make_ulonglong2(0, 0);
        if (i + 1 < here) {
            asm("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(src + i));
            if (WEIGHTED) asm("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(wsrc + i));
        } else if (i < here) {
            v.x = ldg_stream64(src + i);
            if (WEIGHTED) w.x = ldg_stream64(wsrc + i);
        }
        key[j] = v.x;
        key[j + 1] = v.y;
        if (WEIGHTED) { wt[j] = w.x; wt[j + 1] = w.y; }
    }
#pragma unroll
    for (int j = 0; j < kPartKpt; j++)
        rank[j] = key[j] != kEmpty ? atomicAdd(&hist[(uint32_t)(key[j] >> shift) & fmask], 1u) : 0u;
    __syncthreads();
    const uint32_t c0 = hist[2 * threadIdx.x], c1 = hist[2 * threadIdx.x + 1];  // (bins >= F stay 0)
    uint32_t total;
    const uint32_t st = block_excl_scan(c0 + c1, tmp, total);
    const uint64_t gb = ((uint64_t)b1 << a.pb) + 2 * threadIdx.x;
    uint32_t g0 = 0, g1 = 0;
    if (c0) g0 = atomicAdd(a.cursors_out + gb, c0);
    if (c1) g1 = atomicAdd(a.cursors_out + gb + 1, c1);
    if (COUNT_ONLY) return;
    start[2 * threadIdx.x] = st;
    start[2 * threadIdx.x + 1] = st + c0;
    delta[2 * threadIdx.x] = g0 - st;
    delta[2 * threadIdx.x + 1] = g1 - (st + c0);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartKpt; j++) {
        if (key[j] != kEmpty) {
            const uint32_t p = start[(uint32_t)(key[j] >> shift) & fmask] + rank[j];
            stage[p] = key[j];
            if (WEIGHTED) stage[kPartTile + p] = wt[j];
        }
    }
    __syncthreads();
    bool over = false;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < total; i += kBlockThreads) {
        const uint64_t k = stage[i];
        const uint32_t b2 = (uint32_t)(k >> shift) & fmask;
        const uint32_t p = delta[b2] + i;
        const uint64_t b = ((uint64_t)b1 << a.pb) + b2;
        uint64_t base = b * a.cap_out, room = a.cap_out;
        if (a.bases_out) {
            base = a.bases_out[b];
            room = a.bases_out[b + 1] - base;
        }
        if (p < room) {
            a.keys_out[base + p] = k;
            if (WEIGHTED) a.wts_out[base + p] = stage[kPartTile + i];
        } else {
            over = true;
        }
    }
    if (over) atomicOr(a.ctr + kCtrFlags, (unsigned long long)a.over_flag);
}

// The slow path of a table update whose home slot (already read: `seen`) does not hold the barcode.
__device__ __forceinline__ void table_add_from(const TableRef &t, uint64_t bc, uint64_t home, uint64_t seen, uint64_t n_rec,
                                               uint64_t n_dist) {
    const uint32_t words = t.packed ? 2 : 4;
    uint64_t slot = home, cur = seen;
    for (uint32_t probe = 0; probe < 96; probe++) {  // a crowded table fails fast
        uint64_t *s = t.slots + words * slot;
        if (probe) cur = *reinterpret_cast<volatile uint64_t *>(s);
        if (cur == kEmpty) {
            cur = atomicCAS(reinterpret_cast<unsigned long long *>(s), kEmpty, bc);
            if (cur == kEmpty) {
                atomicAdd(t.ctr + kCtrClaimed, 1ull);
                cur = bc;
            }
        }
        if (cur == bc) {
            table_hit(t, slot, n_rec, n_dist);
            return;
        }
        slot = (slot + 1) & t.mask;  // another barcode lives here (a slot never changes once claimed)
    }
    atomicOr(t.ctr + kCtrFlags, (unsigned long long)kFlagTable);
}

// k_bucket_dedup, second form.  ncu on the first form (profiles/r2_k4_dedup_v1_ncu_summary.json): 13.5 warp
// instructions per key, almost half of them in the scan of the bucket's table for live slots — at
// 5x duplication 7 % of the slots are live, so every warp instruction of the scan (two 64-bit
// multiplies to unmix a key, two more to hash its barcode) ran for one or two lanes.  Here a thread
// that claims a slot appends the slot number to a list (shared-memory atomicAdd(+1), the cheap
// kind), the bucket's distinct keys are then processed from that list with full warps, and the
// thread that reads an entry also resets its slot, so the table is cleared once per CTA, not once
// per bucket.  The per-bucket counters stay in registers until the CTA ends (131 072 same-address
// global atomics otherwise).
template <bool WEIGHTED, bool PAIRS>
__global__ void __launch_bounds__(kBlockThreads, 5) k_bucket_dedup2(const DedupArgs a) {
    using Cnt = typename std::conditional<WEIGHTED, unsigned long long, uint32_t>::type;
    extern __shared__ __align__(16) unsigned long long smem[];
    const uint32_t S = 1u << a.s_bits, smask = S - 1u;
    unsigned long long *tkey = smem;
    Cnt *tcnt = reinterpret_cast<Cnt *>(smem + S);
    uint16_t *list = reinterpret_cast<uint16_t *>(tcnt + S);  // slots claimed for the current bucket
    __shared__ uint32_t s_n[2], s_full;
    __shared__ unsigned long long s_base;
    const uint64_t umask = (1ull << a.ub) - 1ull;
    const uint32_t hshift = 64 - a.pb - a.s_bits;
    constexpr uint32_t kMaxProbe = 128;  // at the design load (<= 0.6) a probe sequence this long does not occur
    constexpr uint32_t kBatch = 4 * kBlockThreads;

    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
        tkey[i] = kEmpty;
        tcnt[i] = 0;
    }
    if (threadIdx.x == 0) s_n[0] = s_n[1] = 0, s_full = 0;

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
    uint32_t par = 0;
    auto insert = [&](uint64_t k, uint64_t w) {
        if (k == kEmpty) return;
        uint32_t slot = (uint32_t)(k >> hshift) & smask;
        for (uint32_t probe = 0; probe < kMaxProbe; probe++, slot = (slot + 1) & smask) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tkey + slot);
            if (cur != k) {
                if (cur != kEmpty) continue;  // a slot changes only when its bucket is over
                cur = atomicCAS(tkey + slot, kEmpty, (unsigned long long)k);
                if (cur == kEmpty) {
                    list[atomicAdd(&s_n[par], 1u)] = (uint16_t)slot;
                    // unweighted: the counter holds the occurrences AFTER the first, so claiming a
                    // slot is the only table atomic of a new key and a repeat costs one add
                    if (!WEIGHTED) return;
                } else if (cur != k) {
                    continue;
                }
            }
            if (WEIGHTED) atomicAdd(tcnt + slot, (Cnt)w);
            else atomicAdd(tcnt + slot, (Cnt)1);  // ATOMS.POPC.INC
            return;
        }
        s_full = 1;  // (many) more distinct keys than the table was sized for
    };

    uint64_t my_pairs = 0;  // thread 0: distinct keys of all the CTA's buckets
    uint32_t cnt_n;
    uint64_t first_n, kn[4], wn[4];
    fetch(blockIdx.x, cnt_n, first_n, kn, wn);
    __syncthreads();
    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x, par ^= 1u) {
        const uint32_t cnt = cnt_n;
        const uint64_t first = first_n;
        uint64_t k[4], w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { k[q] = kn[q]; w[q] = wn[q]; }
        fetch(b + gridDim.x, cnt_n, first_n, kn, wn);
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
        __syncthreads();
        const uint32_t live = s_n[par];
        if (threadIdx.x == 0) {
            s_n[par ^ 1u] = 0;  // (its readers finished before the barrier that ended the previous bucket)
            my_pairs += live;
            if (PAIRS) s_base = atomicAdd(a.table.ctr + kCtrCursor, (unsigned long long)live);
        }
        if (PAIRS) __syncthreads();
        // every distinct pair of the bucket, full warps: one row (pair tables) or one add to its barcode's row
        const uint32_t words = a.table.packed ? 2 : 4;
        for (uint32_t base = 0; base < live; base += 2 * kBlockThreads) {
            uint64_t bc[2], um[2], c[2], home[2], seen[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t i = base + q * kBlockThreads + threadIdx.x;
                bc[q] = kEmpty;
                if (i < live) {
                    const uint32_t slot = list[i];
                    const uint64_t comp = unmix64(tkey[slot]);
                    c[q] = (uint64_t)tcnt[slot] + (WEIGHTED ? 0ull : 1ull);
                    tkey[slot] = kEmpty;
                    tcnt[slot] = 0;
                    bc[q] = comp >> a.ub;  // (a narrow barcode is never all ones)
                    um[q] = comp & umask;
                    if (!PAIRS) {
                        home[q] = mix64(bc[q]) & a.table.mask;
                        seen[q] = *reinterpret_cast<volatile uint64_t *>(a.table.slots + words * home[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (bc[q] == kEmpty) continue;
                if (PAIRS) {
                    const uint64_t pos = s_base + base + q * kBlockThreads + threadIdx.x;
                    if (pos < a.pairs_cap) {
                        a.pairs_out[3 * pos] = bc[q];
                        a.pairs_out[3 * pos + 1] = um[q];
                        a.pairs_out[3 * pos + 2] = c[q];
                    } else {
                        atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagPairsOut);
                    }
                } else if (seen[q] == bc[q]) {
                    table_hit(a.table, home[q], c[q], 1ull);  // the common case
                } else {
                    table_add_from(a.table, bc[q], home[q], seen[q], c[q], 1ull);
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (my_pairs) atomicAdd(a.table.ctr + kCtrPairs, (unsigned long long)my_pairs);
        if (s_full) atomicOr(a.table.ctr + kCtrFlags, (unsigned long long)kFlagSmem);
    }
}

// occupied slots -> rows {barcode, n_records, n_distinct_umi}, order unspecified.  A CTA takes 1024
// slots at a time and reserves its rows with ONE atomicAdd (a warp-level reservation is 262 144
// same-address atomics for an 8 Mi-slot table: 0.2 ms of serialised atomics around 0.03 ms of reads).
__global__ void __launch_bounds__(kBlockThreads)
k_table_rows(const uint64_t *__restrict__ slots, uint64_t n_slots, uint32_t packed, uint64_t *__restrict__ rows,
             unsigned long long *__restrict__ ctr) {
    __shared__ uint32_t tmp[2][8];
    __shared__ unsigned long long s_pos[2];
    uint32_t par = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * 1024; base < n_slots; base += (uint64_t)gridDim.x * 1024, par ^= 1u) {
        uint64_t bc[4], nr[4], nd[4];
        uint32_t live = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint64_t i = base + q * kBlockThreads + threadIdx.x;
            bc[q] = kEmpty;
            if (i < n_slots) {
                if (packed) {
                    ulonglong2 v;
                    asm("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(slots + 2 * i));
                    bc[q] = v.x;
                    nr[q] = (v.y & ((1ull << kPackShift) - 1)) + 1ull;
                    nd[q] = v.y >> kPackShift;
                } else {
                    const u64x4 v = ldg_stream256(slots + 4 * i);
                    bc[q] = v.x; nr[q] = v.y + 1ull; nd[q] = v.z + 1ull;
                }
            }
            live += bc[q] != kEmpty;
        }
        uint32_t total;
        uint32_t at = block_excl_scan(live, tmp[par], total);  // (alternating scratch: one barrier per round suffices)
        if (threadIdx.x == 0 && total) s_pos[par] = atomicAdd(ctr + kCtrCursor, (unsigned long long)total);
        if (total == 0) continue;
        __syncthreads();
        uint64_t pos = s_pos[par] + at;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (bc[q] == kEmpty) continue;
            rows[3 * pos] = bc[q];
            rows[3 * pos + 1] = nr[q];
            rows[3 * pos + 2] = nd[q];
            pos++;
        }
    }
}

}  // namespace k4p
}  // namespace ibu
