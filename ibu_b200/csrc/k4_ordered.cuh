// k4_ordered.cuh — the last stage of the K4 partition path when there are about as many barcodes as
// records (included by k4_partition.cuh).
//
// A per-barcode hash table with ~10^8 live slots is one random DRAM access per pair and its rows
// need a full-size sort by barcode afterwards; the sort fallback moves every key once per digit.
// Here k_part1 / k_part2 run on ORDER-PRESERVING keys, (barcode << ub | umi) left-aligned in 64
// bits, so the final buckets (the top pb bits of the barcode) are already in barcode order and one
// barcode never spans two buckets.  Two kernels finish a bucket (~763 keys) per CTA round in
// shared memory:
//   k_bucket_sort  the bucket's keys are grouped by their next 10 bits (1024 bins, ~0.75 keys
//                  each: shared-memory histogram -> scan -> scatter) and a key's exact place is
//                  its bin's start plus the number of smaller keys in the bin — a full sort at
//                  about one comparison per key.  The sorted keys go back in place; the bucket's
//                  row count (distinct barcodes) is left in rows_of[b].
//   k_bucket_bases rows_of -> first row number of every bucket (the scan the exact layout uses).
//   k_bucket_emit  reads a sorted bucket, head flags (new pair, new barcode) and one block scan
//                  give its rows {barcode, n_records, n_distinct_umi} (src/parallel.rs:79-98 plus
//                  the distinct-UMI count), written straight to their final place, in barcode
//                  order: no table, no row sort.
// One kernel with a decoupled look-back for the row numbers was the first form: 8 bytes per key
// less traffic, but every CTA waits for the buckets before it to publish their counts, and 740
// resident CTAs that run the same short phases in step spend more time waiting than working
// (2.8 ms per 10^8 keys, whatever the look-back: flat, two-level, tickets early or late).
// Records that do not fit the key layout (`wide`) are few; the host counts them with the legacy
// path first and hands their rows in: a bucket merges the ones in its barcode range, the ones with
// barcodes beyond the layout are appended after the last bucket.
// 16 bytes read and 8 written per key, 24 written per row.
#pragma once

namespace ibu {
namespace k4p {

constexpr uint32_t kOrdBinBits = 10, kOrdBins = 1u << kOrdBinBits;
constexpr uint32_t kOrdMaxWide = kBlockThreads;  // wide rows one bucket merges
constexpr uint32_t kOrdMaxBin = 256;             // a bin this full (heavily repeated barcodes) voids the call: O(bin^2)

struct OrdArgs {
    const uint64_t *bases;  // exact layout: bucket b owns keys[bases[b] .. bases[b + 1]); NULL: the uniform layout,
    const uint32_t *cursors;  //   cursors[b] keys at keys[b * lcap]
    uint64_t lcap;
    uint64_t *keys;         // ((barcode << ub) | umi) << (64 - bb - ub); sorted in place, bucket by bucket
    uint32_t n_buckets, pb, bb;
    uint32_t cap;           // keys of one bucket that fit the shared-memory arrays (<= 4096)
    const uint64_t *wrows;  // nullable: rows {barcode, n_records, n_distinct} of the wide list, by barcode
    const uint32_t *wstart; // [n_buckets + 1]: first wide row whose barcode is >= the bucket's first barcode
    uint32_t *rows_of;      // [n_buckets], zeroed: rows of each bucket (k_bucket_sort)
    const uint64_t *row_base;  // [n_buckets + 1]: first row number of each bucket (k_bucket_emit)
    unsigned long long *ctr;   // kCtrPairs += distinct pairs
    uint64_t *rows;
    uint64_t rows_cap;
};

// wstart[b] = number of wide rows with barcode < (b << shift), b in 0..n_buckets (the last one: barcodes
// that still fit the layout)
__global__ void __launch_bounds__(kBlockThreads)
k_wide_starts(const uint64_t *__restrict__ wrows, uint32_t wn, uint32_t n_buckets, uint32_t shift, uint32_t *__restrict__ wstart) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_buckets) return;
    const uint64_t bound = (uint64_t)b << shift;
    uint32_t lo = 0, hi = wn;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (wrows[3 * (uint64_t)mid] < bound) lo = mid + 1;
        else hi = mid;
    }
    wstart[b] = lo;
}

// wide rows whose barcode lies beyond the key layout sort after every bucket's rows
__global__ void __launch_bounds__(kBlockThreads)
k_wide_tail(const uint64_t *__restrict__ wrows, uint32_t wn, const uint32_t *__restrict__ wstart, uint32_t n_buckets,
            const uint64_t *__restrict__ row_base, uint64_t *__restrict__ rows, uint64_t rows_cap, unsigned long long *ctr) {
    const uint32_t from = wstart[n_buckets];
    const uint64_t at = row_base[n_buckets], words = 3ull * (wn - from);
    if (at + (wn - from) > rows_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(ctr + kCtrFlags, (unsigned long long)kFlagPairsOut);
        return;
    }
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x)
        rows[3 * at + w] = wrows[3ull * from + w];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr[kCtrTail] = wn - from;
}

__global__ void __launch_bounds__(kBlockThreads, 5) k_bucket_sort(const OrdArgs a) {
    extern __shared__ __align__(16) unsigned long long smem[];
    unsigned long long *kA = smem, *kB = smem + a.cap;
    __shared__ __align__(16) uint32_t hist[kOrdBins], off[kOrdBins];
    __shared__ uint32_t tmp[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t bshift = 64 - a.pb - kOrdBinBits, sh = 64 - a.bb;
    uint32_t my_pairs = 0, my_flags = 0;  // (a thread sees fewer than 2^32 keys)

    // the first 1024 keys of the next bucket travel while the current one is finished
    uint64_t nfirst = 0, nk[4];
    uint32_t ncnt = 0;
    auto fetch = [&](uint32_t b) {
        ncnt = 0;
        if (b < a.n_buckets) {
            nfirst = a.bases ? a.bases[b] : (uint64_t)b * a.lcap;
            ncnt = a.bases ? (uint32_t)min(a.bases[b + 1] - nfirst, (uint64_t)0xffffffffu) : a.cursors[b];
            if (ncnt > a.cap || (!a.bases && ncnt > a.lcap)) {  // does not fit: the call is void
                my_flags |= kFlagSmem;
                ncnt = 0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            nk[q] = i < ncnt ? ldg_stream64(a.keys + nfirst + i) : 0ull;
        }
    };
    fetch(blockIdx.x);
    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x) {
        const uint64_t first = nfirst;
        uint32_t cnt = ncnt;
        reinterpret_cast<uint4 *>(hist)[tid] = make_uint4(0, 0, 0, 0);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            if (i < cnt) {
                kA[i] = nk[q];
                atomicAdd(&hist[(uint32_t)(nk[q] >> bshift) & (kOrdBins - 1u)], 1u);
            }
        }
        for (uint32_t i = tid + 4 * kBlockThreads; i < cnt; i += kBlockThreads) {
            const uint64_t k = ldg_stream64(a.keys + first + i);
            kA[i] = k;
            atomicAdd(&hist[(uint32_t)(k >> bshift) & (kOrdBins - 1u)], 1u);
        }
        __syncthreads();
        const uint4 c = reinterpret_cast<const uint4 *>(hist)[tid];
        uint32_t total;
        const uint32_t st = block_excl_scan(c.x + c.y + c.z + c.w, tmp, total);
        reinterpret_cast<uint4 *>(off)[tid] = make_uint4(st, st + c.x, st + c.x + c.y, st + c.x + c.y + c.z);
        const bool lumpy = __syncthreads_or(max(max(c.x, c.y), max(c.z, c.w)) > kOrdMaxBin) != 0;
        if (lumpy) {
            my_flags |= kFlagSmem;
            cnt = 0;
        }
        for (uint32_t i = tid; i < cnt; i += kBlockThreads) {
            const uint64_t k = kA[i];
            kB[atomicAdd(&off[(uint32_t)(k >> bshift) & (kOrdBins - 1u)], 1u)] = k;
        }
        __syncthreads();  // off[bin] is now the bin's end
        for (uint32_t p = tid; p < cnt; p += kBlockThreads) {
            const uint64_t k = kB[p];
            const uint32_t bin = (uint32_t)(k >> bshift) & (kOrdBins - 1u);
            const uint32_t en = off[bin], s0 = en - hist[bin];
            uint32_t less = 0;
            for (uint32_t q = s0; q < en; q++) {
                const uint64_t kq = kB[q];
                less += (kq < k) | ((kq == k) & (q < p));
            }
            kA[s0 + less] = k;
        }
        __syncthreads();
        fetch(b + gridDim.x);
        // the sorted keys back in place; rows (distinct barcodes) and distinct pairs counted on the way
        uint32_t rows = 0;
        for (uint32_t i = tid; i < cnt; i += kBlockThreads) {
            const uint64_t k = kA[i], prev = i ? kA[i - 1] : ~k;
            a.keys[first + i] = k;
            my_pairs += k != prev;
            rows += (k >> sh) != (prev >> sh) || i == 0;
        }
        if (a.wrows) {  // wide rows of this bucket's barcode range whose barcode has no key here become rows
            const uint32_t w0 = a.wstart[b], wcnt = a.wstart[b + 1] - w0;
            for (uint32_t t = tid; t < wcnt; t += kBlockThreads) {
                const uint64_t wbc = a.wrows[3ull * (w0 + t)];
                uint32_t l = 0, h = cnt;
                while (l < h) {
                    const uint32_t mid = (l + h) >> 1;
                    if ((kA[mid] >> sh) < wbc) l = mid + 1;
                    else h = mid;
                }
                rows += !(l < cnt && (kA[l] >> sh) == wbc);
            }
        }
        rows = __reduce_add_sync(0xffffffffu, rows);
        if (lane == 0 && rows) atomicAdd(a.rows_of + b, rows);
        // (the next round's first barrier comes before kA is written again)
    }
    if (my_flags) atomicOr(a.ctr + kCtrFlags, (unsigned long long)my_flags);
    my_pairs = __reduce_add_sync(0xffffffffu, my_pairs);
    if (lane == 0 && my_pairs) atomicAdd(a.ctr + kCtrPairs, (unsigned long long)my_pairs);
}

// ---- the same buckets as the last stage of a record SORT (ibu_gpu_sort_records, barcode_agg.cu: k4_sort_records_msd) ----
// k_part1<true> / k_part2<true> carry a second word per key (the multi-GPU merge's multiplicity); for a sort
// that word is the record's index, and a final bucket holds the (key, index) pairs of one key range.  A CTA
// round sorts one bucket by (key, index) — bins by the next 10 key bits, a pair's place = its bin's start + the
// number of smaller pairs in the bin — and writes the RECORDS {barcode, umi, index} at their final place:
// out_base[b] is the number of records in the buckets before b.  24 R + 16 W, 16 R + 16 W per level, 16 R +
// 24 W here: 112 bytes per record against 24 + 48 per 8-bit digit of the LSD sort (360 for bc16/umi12).
struct SortRecArgs {
    const uint64_t *bases;    // exact layout (bucket b owns [bases[b], bases[b + 1])) or NULL: cursors[b] pairs at b * lcap
    const uint32_t *cursors;
    uint64_t lcap;
    const uint64_t *keys, *wts;
    uint32_t n_buckets, pb, bb, ub;
    uint32_t cap;             // pairs of one bucket that fit the shared-memory arrays
    const uint64_t *out_base; // [n_buckets + 1]
    uint64_t *out;            // records
    unsigned long long *ctr;
};

__global__ void __launch_bounds__(kBlockThreads, 5) k_bucket_sort_records(const SortRecArgs a) {
    extern __shared__ __align__(16) unsigned long long smem[];
    unsigned long long *kA = smem, *wA = smem + a.cap;               // as loaded
    uint16_t *iB = reinterpret_cast<uint16_t *>(smem + 2 * a.cap);   // slots grouped by bin
    uint16_t *ord = iB + a.cap;                                      // slot of the j-th smallest pair
    __shared__ __align__(16) uint32_t hist[kOrdBins], off[kOrdBins];
    __shared__ uint32_t tmp[8];
    const uint32_t tid = threadIdx.x;
    const uint32_t bshift = 64 - a.pb - kOrdBinBits, sh_bc = 64 - a.bb, sh_um = 64 - a.bb - a.ub;
    const uint64_t umask = (1ull << a.ub) - 1ull;
    uint32_t my_flags = 0;

    uint64_t nfirst = 0, nout = 0, nk[4], nw[4];
    uint32_t ncnt = 0;
    auto fetch = [&](uint32_t b) {
        ncnt = 0;
        if (b < a.n_buckets) {
            nout = a.out_base[b];  // (needed a bucket later: its latency stays out of the output loop)
            nfirst = a.bases ? a.bases[b] : (uint64_t)b * a.lcap;
            ncnt = a.bases ? (uint32_t)min(a.bases[b + 1] - nfirst, (uint64_t)0xffffffffu) : a.cursors[b];
            if (ncnt > a.cap || (!a.bases && ncnt > a.lcap)) {  // does not fit: the call is void
                my_flags |= kFlagSmem;
                ncnt = 0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            nk[q] = i < ncnt ? ldg_stream64(a.keys + nfirst + i) : 0ull;
            nw[q] = i < ncnt ? ldg_stream64(a.wts + nfirst + i) : 0ull;
        }
    };
    fetch(blockIdx.x);
    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x) {
        const uint64_t first = nfirst;
        uint64_t *dst = a.out + 3 * nout;
        uint32_t cnt = ncnt;
        reinterpret_cast<uint4 *>(hist)[tid] = make_uint4(0, 0, 0, 0);
        __syncthreads();  // (also: every thread has finished the previous bucket's output)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            if (i < cnt) {
                kA[i] = nk[q];
                wA[i] = nw[q];
                atomicAdd(&hist[(uint32_t)(nk[q] >> bshift) & (kOrdBins - 1u)], 1u);
            }
        }
        for (uint32_t i = tid + 4 * kBlockThreads; i < cnt; i += kBlockThreads) {
            const uint64_t k = ldg_stream64(a.keys + first + i);
            kA[i] = k;
            wA[i] = ldg_stream64(a.wts + first + i);
            atomicAdd(&hist[(uint32_t)(k >> bshift) & (kOrdBins - 1u)], 1u);
        }
        __syncthreads();
        const uint4 c = reinterpret_cast<const uint4 *>(hist)[tid];
        uint32_t total;
        const uint32_t st = block_excl_scan(c.x + c.y + c.z + c.w, tmp, total);
        reinterpret_cast<uint4 *>(off)[tid] = make_uint4(st, st + c.x, st + c.x + c.y, st + c.x + c.y + c.z);
        const bool lumpy = __syncthreads_or(max(max(c.x, c.y), max(c.z, c.w)) > kOrdMaxBin) != 0;
        if (lumpy) {  // hundreds of records under a few keys: O(bin^2) here, the LSD sort takes the input
            my_flags |= kFlagSmem;
            cnt = 0;
        }
        for (uint32_t i = tid; i < cnt; i += kBlockThreads)
            iB[atomicAdd(&off[(uint32_t)(kA[i] >> bshift) & (kOrdBins - 1u)], 1u)] = (uint16_t)i;
        __syncthreads();  // off[bin] is now the bin's end
        for (uint32_t p = tid; p < cnt; p += kBlockThreads) {
            const uint32_t i = iB[p];
            const uint64_t k = kA[i], w = wA[i];
            const uint32_t bin = (uint32_t)(k >> bshift) & (kOrdBins - 1u);
            const uint32_t en = off[bin], s0 = en - hist[bin];
            uint32_t less = 0;
            for (uint32_t q = s0; q < en; q++) {
                const uint32_t iq = iB[q];
                const uint64_t kq = kA[iq];
                if (kq < k) {
                    less++;
                } else if (kq == k) {
                    const uint64_t wq = wA[iq];
                    less += (wq < w) | ((wq == w) & (q < p));
                }
            }
            ord[s0 + less] = (uint16_t)i;
        }
        __syncthreads();
        fetch(b + gridDim.x);
        // the records, word by word in sorted order: neighbouring threads store neighbouring words
        // (branch-free: both words of the pair are read, shift and mask picked by the field)
        for (uint32_t x = tid; x < 3 * cnt; x += kBlockThreads) {
            const uint32_t j = x / 3u, f = x - 3u * j, i = ord[j];
            const uint64_t k = kA[i], w = wA[i];
            const uint64_t v = (k >> (f == 0 ? sh_bc : sh_um)) & (f == 0 ? ~0ull : umask);
            dst[x] = f == 2 ? w : v;
        }
    }
    if (my_flags) atomicOr(a.ctr + kCtrFlags, (unsigned long long)my_flags);
}

__global__ void __launch_bounds__(kBlockThreads, 6) k_bucket_emit(const OrdArgs a) {
    extern __shared__ __align__(16) unsigned long long smem[];
    unsigned long long *kA = smem;
    uint16_t *rstart = reinterpret_cast<uint16_t *>(smem + a.cap);
    uint16_t *rpairs = rstart + a.cap + 2;
    __shared__ uint32_t tmp[2][8];
    __shared__ unsigned long long wbc[kOrdMaxWide];
    __shared__ uint16_t wU[kOrdMaxWide + 1];
    const uint32_t tid = threadIdx.x;
    const uint32_t sh = 64 - a.bb;
    uint32_t my_flags = 0;

    uint64_t nfirst = 0, nk[4];
    uint32_t ncnt = 0;
    auto fetch = [&](uint32_t b) {  // the first 1024 keys of the next bucket travel while the current one is written
        ncnt = 0;
        if (b < a.n_buckets) {
            nfirst = a.bases ? a.bases[b] : (uint64_t)b * a.lcap;
            ncnt = a.bases ? (uint32_t)min(a.bases[b + 1] - nfirst, (uint64_t)0xffffffffu) : a.cursors[b];
            if (ncnt > a.cap || (!a.bases && ncnt > a.lcap)) ncnt = 0;  // (k_bucket_sort raised the flag)
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            nk[q] = i < ncnt ? ldg_stream64(a.keys + nfirst + i) : 0ull;
        }
    };
    fetch(blockIdx.x);
    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x) {
        const uint64_t first = nfirst, base = a.row_base[b];
        const uint32_t n_out = (uint32_t)(a.row_base[b + 1] - base);
        const uint32_t cnt = ncnt;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = q * kBlockThreads + tid;
            if (i < cnt) kA[i] = nk[q];
        }
        for (uint32_t i = tid + 4 * kBlockThreads; i < cnt; i += kBlockThreads) kA[i] = ldg_stream64(a.keys + first + i);
        __syncthreads();
        // a new barcode starts a row, a new pair adds to its distinct count
        const uint32_t per = (cnt + kBlockThreads - 1) / kBlockThreads;
        const uint32_t lo = min(cnt, tid * per), hi = min(cnt, lo + per);
        uint32_t heads = 0;  // rows | pairs << 16 that start in [lo, hi)
        {
            uint64_t prev = lo ? kA[lo - 1] : 0;
            for (uint32_t i = lo; i < hi; i++) {
                const uint64_t k = kA[i];
                const uint32_t hp = (i == 0) | (k != prev), hb = (i == 0) | ((k >> sh) != (prev >> sh));
                heads += hb + (hp << 16);
                prev = k;
            }
        }
        uint32_t all;
        const uint32_t before = block_excl_scan(heads, tmp[0], all);
        const uint32_t nrows = all & 0xffffu, npairs = all >> 16;
        {
            uint32_t r = before & 0xffffu, pp = before >> 16;
            uint64_t prev = lo ? kA[lo - 1] : 0;
            for (uint32_t i = lo; i < hi; i++) {
                const uint64_t k = kA[i];
                const uint32_t hp = (i == 0) | (k != prev), hb = (i == 0) | ((k >> sh) != (prev >> sh));
                if (hb) {
                    rstart[r] = (uint16_t)i;
                    rpairs[r] = (uint16_t)pp;
                    r++;
                }
                pp += hp;
                prev = k;
            }
        }
        if (tid == 0) {
            rstart[nrows] = (uint16_t)cnt;
            rpairs[nrows] = (uint16_t)npairs;
        }
        __syncthreads();
        fetch(b + gridDim.x);
        uint32_t w0 = 0, wcnt = 0;
        if (a.wrows) {
            w0 = a.wstart[b];
            wcnt = a.wstart[b + 1] - w0;
        }
        if (wcnt > kOrdMaxWide) {  // more wide rows in one bucket's range than a CTA merges: the call is void
            my_flags |= kFlagSmem;
            wcnt = 0;
        }
        if (base + n_out > a.rows_cap) {
            my_flags |= kFlagPairsOut;
        } else if (wcnt == 0) {  // word-wise: consecutive threads write consecutive words
            for (uint32_t w = tid; w < 3 * nrows; w += kBlockThreads) {
                const uint32_t row = w / 3, f = w - 3 * row, i0 = rstart[row];
                const uint64_t v = f == 0 ? kA[i0] >> sh : f == 1 ? (uint64_t)(rstart[row + 1] - i0) : (uint64_t)(rpairs[row + 1] - rpairs[row]);
                a.rows[3 * base + w] = v;
            }
        } else {
            // wide rows of this bucket's barcode range: matched ones add to a row, the others become rows
            uint64_t my_wbc = 0;
            uint32_t my_ins = 0;
            bool my_unmatched = false;
            if (tid < wcnt) {
                my_wbc = a.wrows[3ull * (w0 + tid)];
                wbc[tid] = my_wbc;
                uint32_t l = 0, h = nrows;
                while (l < h) {
                    const uint32_t mid = (l + h) >> 1;
                    if ((kA[rstart[mid]] >> sh) < my_wbc) l = mid + 1;
                    else h = mid;
                }
                my_ins = l;
                my_unmatched = !(l < nrows && (kA[rstart[l]] >> sh) == my_wbc);
            }
            uint32_t n_un;
            const uint32_t my_u = block_excl_scan(my_unmatched ? 1u : 0u, tmp[1], n_un);
            if (tid < wcnt) wU[tid] = (uint16_t)my_u;
            if (tid == 0) wU[wcnt] = (uint16_t)n_un;
            __syncthreads();
            for (uint32_t r = tid; r < nrows; r += kBlockThreads) {
                const uint32_t i0 = rstart[r];
                const uint64_t bc = kA[i0] >> sh;
                uint64_t nr = rstart[r + 1] - i0, nd = rpairs[r + 1] - rpairs[r];
                uint32_t l = 0, h = wcnt;
                while (l < h) {
                    const uint32_t mid = (l + h) >> 1;
                    if (wbc[mid] < bc) l = mid + 1;
                    else h = mid;
                }
                if (l < wcnt && wbc[l] == bc) {
                    nr += a.wrows[3ull * (w0 + l) + 1];
                    nd += a.wrows[3ull * (w0 + l) + 2];
                }
                uint64_t *dst = a.rows + 3 * (base + r + wU[l]);
                dst[0] = bc;
                dst[1] = nr;
                dst[2] = nd;
            }
            if (my_unmatched) {
                uint64_t *dst = a.rows + 3 * (base + my_ins + my_u);
                dst[0] = my_wbc;
                dst[1] = a.wrows[3ull * (w0 + tid) + 1];
                dst[2] = a.wrows[3ull * (w0 + tid) + 2];
            }
        }
        __syncthreads();  // kA, rstart, wbc are rewritten by the next round
    }
    if (my_flags) atomicOr(a.ctr + kCtrFlags, (unsigned long long)my_flags);
}

}  // namespace k4p
}  // namespace ibu
