// barcode_count.cu — K4: per-barcode record / distinct-UMI table.
//
// The device form of the reference's HashMap<barcode, count> processor
// (src/parallel.rs:79-98), extended with the number of distinct UMI words per barcode, rows
// emitted in barcode order (Record's Ord, src/constructs/record.rs:58).
//
// Sorted input (by barcode, then umi) — the streaming path, 24 B/record, ONE pass over HBM:
//   k_segments   tiles of 2048 records, tile ids handed out in order by an atomic counter.
//                Each record is compared with its predecessor: barcode change = a new table
//                row (head), (barcode, umi) change = a new distinct pair.  Tile-local ranks
//                come from warp ballots; the global row number of the tile's first head
//                comes from a decoupled look-back over per-tile descriptors (status | count
//                in one 64-bit word), so no second pass over the records is needed.  A head
//                writes {barcode, start position, distinct pairs before it in its tile}.
//                The same pass verifies the order; a violation raises a flag and the host
//                falls back to the unsorted path.
//   k_scan_tiles exclusive scan of the per-tile pair counts (n/2048 values).
//   k_finalize   row r: n_records = start[r+1] - start[r]; n_distinct = pairs[r+1] - pairs[r].
// Unsorted input: (barcode, umi) pairs are extracted (16 B), LSD radix sorted 8 bits at a time
// over the bits that actually vary, and fed to the same segment kernel (stride 2 instead of 3).
#include <algorithm>

#include "ctx.h"
#include "kernels.cuh"

namespace ibu {

constexpr int kSegTile = 2048;                       // records per tile
constexpr int kSegPerThread = kSegTile / kBlockThreads;  // 8
#define kStatusAgg (1ull << 62)
#define kStatusPrefix (2ull << 62)
#define kValueMask ((1ull << 62) - 1)

struct SegArgs {
    const uint64_t *src;     // records (stride 3 u64) or sorted pairs (stride 2 u64)
    uint64_t n;
    uint64_t n_tiles;
    uint64_t *desc;          // [n_tiles] look-back descriptors, zeroed
    uint32_t *tile_pairs;    // [n_tiles] distinct pairs that start in the tile
    uint64_t *tmp_rows;      // [capacity][3]: barcode, start position, pairs before it in its tile
    uint64_t capacity;
    unsigned long long *counters;  // [0] next tile id, [1] total heads, [2] unsorted flag
};

// STRIDE = u64 words per element (3: Record, 2: (barcode, umi) pair).
template <int STRIDE>
__global__ void __launch_bounds__(kBlockThreads) k_segments(const SegArgs a) {
    extern __shared__ __align__(16) uint64_t tile[];  // kSegTile * STRIDE words
    __shared__ uint32_t cnt_b[kSegPerThread][kWarpsPerBlock], cnt_p[kSegPerThread][kWarpsPerBlock];
    __shared__ uint64_t s_prev[2];
    __shared__ uint64_t s_tile, s_head_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    volatile unsigned long long *v_flag = a.counters + 2;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(a.counters, 1ull);
        __syncthreads();
        const uint64_t t = s_tile;
        if (t >= a.n_tiles) break;
        const uint64_t first = t * kSegTile;
        const uint32_t count = (uint32_t)min((uint64_t)kSegTile, a.n - first);

        // ---- stage the tile (coalesced 16-byte loads; the tile start is 16-byte aligned) ----
        {
            const uint32_t words = count * STRIDE;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(a.src + first * STRIDE);
            uint4 *s4 = reinterpret_cast<uint4 *>(tile);
            const uint32_t n16 = words / 2;
            for (uint32_t i = tid; i < n16; i += kBlockThreads) s4[i] = ldg_stream(g4 + i);
            if ((words & 1u) && tid == 0) tile[words - 1] = ldg_stream64(a.src + first * STRIDE + words - 1);
            if (tid == 0 && first > 0) {
                s_prev[0] = ldg_stream64(a.src + (first - 1) * STRIDE);
                s_prev[1] = ldg_stream64(a.src + (first - 1) * STRIDE + 1);
            }
        }
        __syncthreads();

        // ---- flags: record i = tid + 256 q (24/16-byte stride: conflict-free 64-bit loads) ----
        uint32_t hb = 0, hp = 0;  // bit q: record q of this thread is a head / starts a new pair
        uint32_t bad = 0;
#pragma unroll
        for (int q = 0; q < kSegPerThread; q++) {
            const uint32_t i = tid + kBlockThreads * q;
            uint32_t is_b = 0, is_p = 0;
            if (i < count) {
                const uint64_t bc = tile[i * STRIDE], um = tile[i * STRIDE + 1];
                if (i == 0 && first == 0) {
                    is_b = is_p = 1;
                } else {
                    const uint64_t pb = i ? tile[(i - 1) * STRIDE] : s_prev[0];
                    const uint64_t pu = i ? tile[(i - 1) * STRIDE + 1] : s_prev[1];
                    is_b = bc != pb;
                    is_p = is_b | (um != pu);
                    bad |= (pb > bc) | ((pb == bc) & (pu > um));
                }
            }
            const uint32_t mb = __ballot_sync(0xffffffffu, is_b), mp = __ballot_sync(0xffffffffu, is_p);
            if (lane == 0) {
                cnt_b[q][warp] = __popc(mb);
                cnt_p[q][warp] = __popc(mp);
            }
            hb |= is_b << q;
            hp |= is_p << q;
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.counters + 2, 1ull);
        __syncthreads();

        // ---- warp 0: scan the 64 (q, warp) counts in record order, then the look-back ----
        if (warp == 0) {
            uint32_t *fb = &cnt_b[0][0], *fp = &cnt_p[0][0];
            uint32_t b0 = fb[2 * lane], b1 = fb[2 * lane + 1], p0 = fp[2 * lane], p1 = fp[2 * lane + 1];
            uint32_t sb = b0 + b1, sp = p0 + p1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t vb = __shfl_up_sync(0xffffffffu, sb, o), vp = __shfl_up_sync(0xffffffffu, sp, o);
                if (lane >= o) { sb += vb; sp += vp; }
            }
            const uint32_t tot_b = __shfl_sync(0xffffffffu, sb, 31), tot_p = __shfl_sync(0xffffffffu, sp, 31);
            fb[2 * lane] = sb - b0 - b1; fb[2 * lane + 1] = sb - b1;  // exclusive prefixes
            fp[2 * lane] = sp - p0 - p1; fp[2 * lane + 1] = sp - p1;

            volatile uint64_t *desc = a.desc;
            if (lane == 0) {
                a.tile_pairs[t] = tot_p;
                desc[t] = (t == 0 ? kStatusPrefix : kStatusAgg) | tot_b;
            }
            uint64_t exclusive = 0;
            if (t > 0) {
                int64_t j = (int64_t)t - 1;  // nearest predecessor; lane l inspects tile j - l
                for (;;) {
                    const int64_t mine = j - lane;
                    uint64_t d;
                    uint32_t need, spins = 0;
                    do {
                        d = mine >= 0 ? desc[mine] : kStatusPrefix;
                        const uint32_t inval = __ballot_sync(0xffffffffu, (d >> 62) == 0);
                        const uint32_t pref = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                        // lanes up to and including the first inclusive prefix must be published
                        const uint32_t upto = pref ? ((2u << (__ffs(pref) - 1)) - 1u) : 0xffffffffu;
                        need = inval & upto;  // warp-uniform
                        if (need) {
                            // stop waiting when the order check failed elsewhere (results are void
                            // then) or, as a watchdog, after ~4 M polls; the vote keeps the exit
                            // decision uniform even if lanes observe the flag at different times
                            const bool timeout = ++spins > (1u << 22);
                            if (__any_sync(0xffffffffu, *v_flag != 0ull) || timeout) {
                                if (timeout && lane == 0) atomicOr(a.counters + 2, 2ull);
                                need = 0;
                            }
                        }
                    } while (need);
                    const uint32_t pref = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                    const uint32_t upto = pref ? ((2u << (__ffs(pref) - 1)) - 1u) : 0xffffffffu;
                    uint64_t v = ((upto >> lane) & 1u) ? (d & kValueMask) : 0ull;
                    exclusive += warp_sum64(v);
                    if (pref) break;
                    j -= 32;
                }
                if (lane == 0) desc[t] = kStatusPrefix | (exclusive + tot_b);
            }
            if (lane == 0) {
                s_head_base = exclusive;
                if (t == a.n_tiles - 1) a.counters[1] = exclusive + tot_b;  // total rows
            }
        }
        __syncthreads();

        // ---- heads write their row stub ----
        const uint64_t head_base = s_head_base;
#pragma unroll
        for (int q = 0; q < kSegPerThread; q++) {
            const uint32_t is_b = (hb >> q) & 1u, is_p = (hp >> q) & 1u;
            const uint32_t mb = __ballot_sync(0xffffffffu, is_b), mp = __ballot_sync(0xffffffffu, is_p);
            if (is_b) {
                const uint32_t lt = (1u << lane) - 1u;
                const uint64_t row = head_base + cnt_b[q][warp] + __popc(mb & lt);
                if (row < a.capacity) {
                    const uint32_t i = tid + kBlockThreads * q;
                    uint64_t *dst = a.tmp_rows + 3 * row;
                    dst[0] = tile[i * STRIDE];
                    dst[1] = first + i;
                    dst[2] = cnt_p[q][warp] + __popc(mp & lt);
                }
            }
        }
        __syncthreads();  // the tile and the count arrays are reused by the next iteration
    }
}

// exclusive scan of tile_pairs -> tile_prefix (u64); totals[3] = number of distinct pairs
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *__restrict__ tile_pairs,
                                                     uint64_t *__restrict__ tile_prefix, uint64_t n_tiles,
                                                     unsigned long long *counters) {
    __shared__ uint64_t part[1024];
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n_tiles + 1023) / 1024;
    const uint64_t lo = min(n_tiles, tid * per), hi = min(n_tiles, lo + per);
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; i++) s += tile_pairs[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        uint64_t v = tid >= (uint32_t)o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint64_t run = part[tid] - s;
    for (uint64_t i = lo; i < hi; i++) {
        tile_prefix[i] = run;
        run += tile_pairs[i];
    }
    if (tid == 1023) counters[3] = part[1023];
}

__global__ void __launch_bounds__(kBlockThreads)
k_finalize_rows(const uint64_t *__restrict__ tmp_rows, const uint64_t *__restrict__ tile_prefix,
                uint64_t n_rows, uint64_t n, const unsigned long long *__restrict__ counters,
                ibu_barcode_row_t *__restrict__ rows) {
    const uint64_t total_pairs = counters[3];
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t bc = tmp_rows[3 * r], start = tmp_rows[3 * r + 1], lp = tmp_rows[3 * r + 2];
        const uint64_t g = tile_prefix[start / kSegTile] + lp;
        uint64_t next_start = n, next_g = total_pairs;
        if (r + 1 < n_rows) {
            next_start = tmp_rows[3 * r + 4];
            next_g = tile_prefix[next_start / kSegTile] + tmp_rows[3 * r + 5];
        }
        rows[r].barcode = bc;
        rows[r].n_records = next_start - start;
        rows[r].n_distinct_umi = next_g - g;
    }
}

// ============================================================ unsorted path: LSD radix sort
constexpr int kSortTile = 2048;                              // pairs per CTA tile
constexpr int kSortItems = kSortTile / kBlockThreads;        // 8 per thread

// records -> (barcode, umi) pairs, plus OR / AND of every word (which bits vary at all)
__global__ void __launch_bounds__(kBlockThreads)
k_extract_pairs(const uint64_t *__restrict__ recs, uint64_t n, uint64_t *__restrict__ pairs,
                unsigned long long *__restrict__ masks /* or_b, and_b, or_u, and_u */) {
    uint64_t ob = 0, ab = ~0ull, ou = 0, au = ~0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = ldg_stream64(recs + 3 * i), u = ldg_stream64(recs + 3 * i + 1);
        ob |= b; ab &= b; ou |= u; au &= u;
        reinterpret_cast<ulonglong2 *>(pairs)[i] = make_ulonglong2(b, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ob |= __shfl_xor_sync(0xffffffffu, ob, o); ab &= __shfl_xor_sync(0xffffffffu, ab, o);
        ou |= __shfl_xor_sync(0xffffffffu, ou, o); au &= __shfl_xor_sync(0xffffffffu, au, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicOr(masks + 0, ob); atomicAnd(masks + 1, ab);
        atomicOr(masks + 2, ou); atomicAnd(masks + 3, au);
    }
}

// per-tile digit histogram, stored digit-major: hist[d * n_tiles + tile]
__global__ void __launch_bounds__(kBlockThreads)
k_radix_hist(const uint64_t *__restrict__ pairs, uint64_t n, uint32_t word, uint32_t shift,
             uint32_t *__restrict__ hist, uint64_t n_tiles) {
    __shared__ uint32_t h[256];
    const uint64_t tile_id = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t first = tile_id * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, n - first);
    for (uint32_t i = threadIdx.x; i < count; i += kBlockThreads) {
        const uint64_t key = pairs[2 * (first + i) + (word ? 0 : 1)];  // (barcode, umi): word 0 = umi
        atomicAdd(&h[(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_tiles + tile_id] = h[threadIdx.x];
}

// one CTA per digit: exclusive scan of that digit's counts across tiles, and the digit total
__global__ void __launch_bounds__(kBlockThreads)
k_radix_scan(uint32_t *__restrict__ hist, uint64_t n_tiles, uint64_t *__restrict__ digit_total) {
    __shared__ uint64_t part[kBlockThreads];
    uint32_t *row = hist + (uint64_t)blockIdx.x * n_tiles;
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n_tiles + kBlockThreads - 1) / kBlockThreads;
    const uint64_t lo = min(n_tiles, tid * per), hi = min(n_tiles, lo + per);
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; i++) s += row[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < kBlockThreads; o <<= 1) {
        uint64_t v = tid >= (uint32_t)o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    // a tile's offset within one digit never exceeds n < 2^32 * tiles; keep 32 bits per entry by
    // storing offsets relative to the digit (the digit base is added from digit_total)
    uint64_t run = part[tid] - s;
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t c = row[i];
        row[i] = (uint32_t)run;
        run += c;
    }
    if (tid == kBlockThreads - 1) digit_total[blockIdx.x] = part[tid];
}

// stable scatter of one tile by one 8-bit digit
__global__ void __launch_bounds__(kBlockThreads)
k_radix_scatter(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n, uint32_t word,
                uint32_t shift, const uint32_t *__restrict__ hist, const uint64_t *__restrict__ digit_total,
                uint64_t n_tiles) {
    __shared__ uint32_t warp_cnt[kWarpsPerBlock][256];
    __shared__ uint64_t base[256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t tile_id = blockIdx.x;
    for (int w = 0; w < kWarpsPerBlock; w++) warp_cnt[w][tid] = 0;
    {   // exclusive scan of the 256 digit totals -> global base of each digit (+ this tile's offset)
        uint64_t v = digit_total[tid];
        base[tid] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            uint64_t x = tid >= (uint32_t)o ? base[tid - o] : 0;
            __syncthreads();
            base[tid] += x;
            __syncthreads();
        }
        const uint64_t excl = base[tid] - v;
        __syncthreads();
        base[tid] = excl + hist[(uint64_t)tid * n_tiles + tile_id];
    }
    __syncthreads();

    const uint64_t first = tile_id * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, n - first);
    // warp w owns elements [w*256, w*256+256) of the tile; item k of lane l is element w*256 + 32k + l
    ulonglong2 el[kSortItems];
    uint32_t rank[kSortItems];
    uint32_t dig[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const uint32_t i = warp * (kSortTile / kWarpsPerBlock) + 32 * k + lane;
        const bool live = i < count;
        if (live) el[k] = reinterpret_cast<const ulonglong2 *>(in)[first + i];
        const uint64_t key = live ? (word ? el[k].x : el[k].y) : 0;
        // pairs are stored (barcode, umi): word 1 = barcode = .x, word 0 = umi = .y
        const uint32_t d = live ? (uint32_t)((key >> shift) & 0xFFu) : 0x100u;  // 0x100: no element
        dig[k] = d;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (live && lane == leader) {
            old = warp_cnt[warp][d];
            warp_cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: exclusive scan over the 8 warps (thread = digit)
        uint32_t run = 0;
        for (int w = 0; w < kWarpsPerBlock; w++) {
            const uint32_t c = warp_cnt[w][tid];
            warp_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        if (dig[k] < 0x100u) {
            const uint64_t pos = base[dig[k]] + warp_cnt[warp][dig[k]] + rank[k];
            reinterpret_cast<ulonglong2 *>(out)[pos] = el[k];
        }
    }
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T>
    cudaError_t alloc(T **out, size_t bytes) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T *)p;
        return e;
    }
    void release(void *p) {  // hand ownership to the caller
        ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end());
    }
};

// Runs the segment pass over `src` (stride 3 or 2).  On success *rows_out (device, owned by the
// caller) holds *n_rows rows.  *unsorted is set when the order check failed (no rows then).
static int segment_pass(ibu_gpu_ctx *ctx, const uint64_t *src, int stride, uint64_t n, cudaStream_t s,
                        ibu_barcode_row_t **rows_out, uint64_t *n_rows, uint64_t *n_pairs, bool *unsorted,
                        ibu_error_t *err) {
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    *unsorted = false;
    const uint64_t n_tiles = (n + kSegTile - 1) / kSegTile;
    Scratch sc;
    uint64_t *desc, *tile_prefix, *tmp_rows;
    uint32_t *tile_pairs;
    unsigned long long *counters;
    IBU_CUDA(sc.alloc(&desc, n_tiles * 8));
    IBU_CUDA(sc.alloc(&tile_prefix, n_tiles * 8));
    IBU_CUDA(sc.alloc(&tile_pairs, n_tiles * 4));
    IBU_CUDA(sc.alloc(&counters, 4 * 8));
    const size_t smem = (size_t)kSegTile * stride * 8;
    auto kern = stride == 3 ? k_segments<3> : k_segments<2>;
    IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    IBU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, smem));
    // every CTA must be resident: a waiting tile spins on descriptors of earlier tiles
    const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * std::max(per_sm, 1), n_tiles);

    uint64_t capacity = std::min<uint64_t>(n, 8ull << 20);  // optimistic: <= 8 Mi distinct barcodes
    for (int attempt = 0; attempt < 2; attempt++) {
        IBU_CUDA(sc.alloc(&tmp_rows, capacity * 24));
        IBU_CUDA(cudaMemsetAsync(desc, 0, n_tiles * 8, s));
        IBU_CUDA(cudaMemsetAsync(counters, 0, 4 * 8, s));
        SegArgs a{src, n, n_tiles, desc, tile_pairs, tmp_rows, capacity, counters};
        kern<<<grid, kBlockThreads, smem, s>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        IBU_CUDA(cudaGetLastError());
        k_scan_tiles<<<1, 1024, 0, s>>>(tile_pairs, tile_prefix, n_tiles, counters);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        unsigned long long h[4];
        IBU_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        if (h[2] & 2ull)
            return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "barcode_count: look-back watchdog expired");
        if (h[2]) {
            *unsorted = true;
            return IBU_OK;
        }
        if (h[1] > capacity) {  // more distinct barcodes than the optimistic table: exact re-run
            capacity = h[1];
            continue;
        }
        ibu_barcode_row_t *rows;
        IBU_CUDA(sc.alloc(&rows, h[1] * sizeof(ibu_barcode_row_t)));
        if (h[1]) {
            const uint64_t blocks = (h[1] + kBlockThreads - 1) / kBlockThreads;
            k_finalize_rows<<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8), kBlockThreads, 0, s>>>(
                tmp_rows, tile_prefix, h[1], n, counters, rows);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            IBU_CUDA(cudaGetLastError());
            IBU_CUDA(cudaStreamSynchronize(s));
        }
        sc.release(rows);
        *rows_out = rows;
        *n_rows = h[1];
        *n_pairs = h[3];
        return IBU_OK;
    }
    return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "barcode table capacity retry failed");
}

// LSD radix sort of (barcode, umi) pairs over the bits that vary; returns the sorted buffer.
static int sort_pairs(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, Scratch &sc,
                      uint64_t **sorted, ibu_error_t *err) {
    if (n >= (1ull << 32))  // per-digit tile offsets are kept in 32 bits
        return set_error(err, IBU_ERR_ARG, 0, n, 0, "unsorted barcode_count supports fewer than 2^32 records per call");
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    uint64_t *buf[2], *digit_total;
    uint32_t *hist;
    unsigned long long *masks;
    IBU_CUDA(sc.alloc(&buf[0], n * 16));
    IBU_CUDA(sc.alloc(&buf[1], n * 16));
    IBU_CUDA(sc.alloc(&hist, 256 * n_tiles * 4));
    IBU_CUDA(sc.alloc(&digit_total, 256 * 8));
    IBU_CUDA(sc.alloc(&masks, 4 * 8));
    const unsigned long long init[4] = {0ull, ~0ull, 0ull, ~0ull};
    IBU_CUDA(cudaMemcpyAsync(masks, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
    k_extract_pairs<<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8), kBlockThreads, 0, s>>>(
        recs, n, buf[0], masks);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    unsigned long long m[4];
    IBU_CUDA(cudaMemcpyAsync(m, masks, sizeof(m), cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    const uint64_t vary[2] = {m[2] ^ m[3], m[0] ^ m[1]};  // word 0 = umi (minor key), word 1 = barcode
    int cur = 0;
    for (uint32_t word = 0; word < 2; word++) {
        for (uint32_t shift = 0; shift < 64; shift += 8) {
            if (((vary[word] >> shift) & 0xFFull) == 0) continue;  // every key agrees on this digit
            k_radix_hist<<<(int)n_tiles, kBlockThreads, 0, s>>>(buf[cur], n, word, shift, hist, n_tiles);
            k_radix_scan<<<256, kBlockThreads, 0, s>>>(hist, n_tiles, digit_total);
            k_radix_scatter<<<(int)n_tiles, kBlockThreads, 0, s>>>(buf[cur], buf[cur ^ 1], n, word, shift, hist,
                                                                    digit_total, n_tiles);
            g_launches.fetch_add(3, std::memory_order_relaxed);
            IBU_CUDA(cudaGetLastError());
            cur ^= 1;
        }
    }
    *sorted = buf[cur];
    return IBU_OK;
}

}  // namespace ibu

using namespace ibu;

extern "C" {

int ibu_gpu_barcode_count(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, int mode,
                          ibu_barcode_table_t *table, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !table || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (mode < 0 || mode > 2) return set_error(err, IBU_ERR_ARG, 0, mode, 0, "mode must be 0, 1 or 2");
    if (((uintptr_t)d_records & 15u) != 0)
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "d_records must be 16-byte aligned");
    memset(table, 0, sizeof(*table));
    table->n_records = n;
    if (n == 0) {
        table->input_was_sorted = 1;
        return IBU_OK;
    }
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    const uint64_t *src = reinterpret_cast<const uint64_t *>(d_records);
    ibu_barcode_row_t *rows = nullptr;
    uint64_t n_rows = 0, n_pairs = 0;
    bool unsorted = mode == 2;
    if (mode != 2) {
        if (int rc = segment_pass(ctx, src, 3, n, s, &rows, &n_rows, &n_pairs, &unsorted, err)) return rc;
        if (!unsorted) table->input_was_sorted = 1;
    }
    if (unsorted) {
        if (mode == 1) return IBU_OK;  // caller required sorted input: input_was_sorted = 0, no rows
        Scratch sc;
        uint64_t *sorted = nullptr;
        if (int rc = sort_pairs(ctx, src, n, s, sc, &sorted, err)) return rc;
        bool still_unsorted = false;
        if (int rc = segment_pass(ctx, sorted, 2, n, s, &rows, &n_rows, &n_pairs, &still_unsorted, err)) return rc;
        if (still_unsorted) return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "internal error: radix sort left the pairs unsorted");
    }
    table->d_rows = rows;
    table->n_rows = n_rows;
    table->n_distinct_pairs = n_pairs;
    return IBU_OK;
}

void ibu_gpu_table_free(ibu_gpu_ctx_t *ctx, ibu_barcode_table_t *table) {
    if (!ctx || !table) return;
    if (table->d_rows) ibu_gpu_free(ctx, table->d_rows);
    table->d_rows = nullptr;
    table->n_rows = 0;
}

}  // extern "C"
