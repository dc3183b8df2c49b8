// barcode_count.cu — K4: per-barcode record / distinct-UMI table.
//
// The device form of the reference's HashMap<barcode, count> processor
// (src/parallel.rs:79-98), extended with the number of distinct UMI words per barcode, rows
// emitted in barcode order (Record's Ord, src/constructs/record.rs:58).
//
// Sorted input (by barcode, then umi) — the streaming path, 24 B/record, ONE pass over HBM:
//   k_segments   every warp claims tiles of 2048 records in order from an atomic counter.
//                A lane owns 4 consecutive records (three LDG.E.256, whole sectors, static
//                register layout), so each record is compared with its predecessor register
//                to register: barcode change = a new table row (head), (barcode, umi) change
//                = a new distinct pair.  Ranks inside the tile come from ballots; the global
//                number of heads / pairs before the tile
//                comes from a decoupled look-back over per-tile descriptors (two 64-bit
//                words, each status | count), so no second pass over the records is needed.
//                A head writes {barcode, start position, distinct pairs before it}.
//                The same pass verifies the order; a violation raises a flag and the host
//                falls back to the unsorted path.
//   k_finalize   row r: n_records = start[r+1] - start[r]; n_distinct = pairs[r+1] - pairs[r].
// Unsorted input: (barcode, umi) pairs are extracted (16 B), LSD radix sorted 8 bits at a time
// over the bits that actually vary, and fed to the same segment kernel (stride 2 instead of 3).
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "ctx.h"
#include "k4.h"
#include "kernels.cuh"

namespace ibu {

struct Trace {  // IBU_B200_TRACE=1: host-side phase timing on stderr (tuning only)
    bool on = getenv("IBU_B200_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[ibu trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

constexpr int kSegSub = 128;                     // records per sub-tile (4 per lane)
// sub-tiles per (warp) tile: 16 = 2048 records.  Measured on B200 at 10^8 records: 4 -> 1.20 ms,
// 8 -> 0.66 ms, 16 -> 0.44 ms (short tiles leave the deferred look-back too little slack).
constexpr int kSegSubsDefault = 16;
#define kStatusAgg (1ull << 62)
#define kStatusPrefix (2ull << 62)
#define kValueMask ((1ull << 62) - 1)

struct SegArgs {
    const uint64_t *src;     // records (stride 3 u64) or sorted pairs (stride 2 u64)
    uint64_t n;
    uint64_t n_tiles;
    ulonglong2 *desc;        // [n_tiles] look-back descriptors {status|heads, status|pairs}, zeroed
    uint64_t *tmp_rows;      // [capacity][3]: barcode, start position, distinct pairs before it
    uint64_t capacity;
    unsigned long long *counters;  // [0] next tile id, [1] total heads, [2] flags, [3] total pairs
};

// The (barcode, umi) keys of the 4 consecutive elements a lane owns in a 128-element sub-tile.
// Records (STRIDE 3): 96 contiguous bytes = three LDG.E.256; pairs (STRIDE 2): 64 bytes = two.
// Every access is a whole 32-byte sector and the words land in registers in a static layout,
// so consecutive elements are compared register to register (no shared-memory transpose).
template <int STRIDE>
struct Keys4 {
    uint64_t bc[4], um[4];
    __device__ __forceinline__ void load(const uint64_t *sub, uint32_t lane) {
        if constexpr (STRIDE == 3) {
            const uint8_t *p = reinterpret_cast<const uint8_t *>(sub) + lane * 96;
            const u64x4 v0 = ldg_stream256(p), v1 = ldg_stream256(p + 32), v2 = ldg_stream256(p + 64);
            bc[0] = v0.x; um[0] = v0.y; bc[1] = v0.w; um[1] = v1.x;
            bc[2] = v1.z; um[2] = v1.w; bc[3] = v2.y; um[3] = v2.z;
        } else {
            const uint8_t *p = reinterpret_cast<const uint8_t *>(sub) + lane * 64;
            const u64x4 v0 = ldg_stream256(p), v1 = ldg_stream256(p + 32);
            bc[0] = v0.x; um[0] = v0.y; bc[1] = v0.z; um[1] = v0.w;
            bc[2] = v1.x; um[2] = v1.y; bc[3] = v1.z; um[3] = v1.w;
        }
    }
    // ragged sub-tile: element-wise; absent elements repeat the key (pb, pu) of the last element
    // of the input, so the padding never starts a run and never breaks the order
    __device__ __forceinline__ void load_partial(const uint64_t *sub, uint32_t lane, uint32_t count,
                                                 uint64_t pb, uint64_t pu) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = 4 * lane + q;
            if (i < count) {
                pb = ldg_stream64(sub + (uint64_t)i * STRIDE);
                pu = ldg_stream64(sub + (uint64_t)i * STRIDE + 1);
            }
            bc[q] = pb;
            um[q] = pu;
        }
    }
};

// STRIDE = u64 words per element (3: Record, 2: (barcode, umi) pair).
//
// Every WARP is an independent worker: it claims the next tile (ordered ids from one atomic
// counter, so a tile only ever waits on tiles whose warps are already running), streams its
// 16 sub-tiles with the next one in flight and publishes the tile's head / pair counts.  The
// look-back for the tile's global offsets does not stall the warp: it is polled, one
// non-blocking step per sub-tile, while the warp already streams its NEXT tile (the masks of
// two tiles are parked in shared memory), and only forced to completion before a third tile
// would start.  No block barrier anywhere.
// PAIRS: rows are distinct (barcode, umi) pairs instead of barcodes (the stub's third word then
// carries the umi): the de-duplicated pair table that shards exchange for an exact merge.
template <int STRIDE, int kSegSubs, bool PAIRS>
__global__ void __launch_bounds__(kBlockThreads, 4) k_segments(const SegArgs a) {
    constexpr int kSegTile = kSegSub * kSegSubs;
    // per (warp, buffer, sub-tile): each lane's head/pair masks and the packed counts before the
    // sub-tile; parked so the sub-tile loop stays rolled and a finished tile can wait for its
    // offsets while the next one streams
    __shared__ uint32_t s_mask[kWarpsPerBlock][2][kSegSubs][32], s_run[kWarpsPerBlock][2][kSegSubs];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    volatile unsigned long long *v_flag = a.counters + 2;
    volatile unsigned long long *desc = reinterpret_cast<volatile unsigned long long *>(a.desc);
    // key of the last element of the input: the padding of the ragged final sub-tile
    const uint64_t endb = a.src[(a.n - 1) * STRIDE], endu = a.src[(a.n - 1) * STRIDE + 1];

    // the tile whose look-back is still open (all fields warp-uniform)
    bool p_active = false;
    uint64_t p_t = 0, p_exb = 0, p_exp = 0;
    int64_t p_j = 0;
    uint32_t p_tot = 0, p_buf = 0;

    // One look-back attempt for the open tile: consume published windows of 32 predecessors until
    // an inclusive prefix is met.  Non-blocking mode gives up at the first window that is not
    // fully published yet.  Returns true when the tile's exclusive offsets are final.
    auto lookback = [&](bool blocking) -> bool {
        uint32_t spins = 0;
        for (;;) {
            const int64_t mine_t = p_j - lane;  // lane l inspects tile j - l
            const uint64_t d0 = mine_t >= 0 ? desc[2 * mine_t] : kStatusPrefix;
            const uint64_t d1 = mine_t >= 0 ? desc[2 * mine_t + 1] : kStatusPrefix;
            // a descriptor is usable when both words carry the same status
            const uint32_t s0 = (uint32_t)(d0 >> 62), s1 = (uint32_t)(d1 >> 62);
            const uint32_t st = s0 == s1 ? s0 : 0u;
            const uint32_t inval = __ballot_sync(0xffffffffu, st == 0);
            const uint32_t pref = __ballot_sync(0xffffffffu, st == 2);
            // lanes up to and including the first inclusive prefix must be published
            const uint32_t upto = pref ? ((2u << (__ffs(pref) - 1)) - 1u) : 0xffffffffu;
            if (inval & upto) {  // warp-uniform
                if (!blocking) return false;
                // stop waiting when the order check failed elsewhere (results are void then) or,
                // as a watchdog, after ~4 M polls; the vote keeps the decision uniform even if
                // lanes observe the flag at different times
                const bool timeout = ++spins > (1u << 22);
                if (__any_sync(0xffffffffu, *v_flag != 0ull) || timeout) {
                    if (timeout && lane == 0) atomicOr(a.counters + 2, 2ull);
                    return true;
                }
                __nanosleep(64);  // leave the issue slots to the streaming warps
                continue;
            }
            const bool take = (upto >> lane) & 1u;
            p_exb += warp_sum64(take ? (d0 & kValueMask) : 0ull);
            p_exp += warp_sum64(take ? (d1 & kValueMask) : 0ull);
            if (pref) return true;
            p_j -= 32;
        }
    };

    // Close the open tile: publish its inclusive prefix and write the row stubs of its heads
    // {barcode, start, distinct pairs before it}.  Ranks inside a sub-tile are rebuilt from the
    // parked masks with ballots, only for sub-tiles that contain a head (heads are rare).
    auto finish = [&]() {
        const uint64_t tot_b = p_tot & 0xFFFFu, tot_p = p_tot >> 16;
        if (lane == 0) {
            if (p_t > 0) {
                desc[2 * p_t] = kStatusPrefix | (p_exb + tot_b);
                desc[2 * p_t + 1] = kStatusPrefix | (p_exp + tot_p);
            }
            if (p_t == a.n_tiles - 1) {
                a.counters[1] = p_exb + tot_b;  // total rows
                a.counters[3] = p_exp + tot_p;  // total distinct pairs
            }
        }
        if (tot_b) {
            const uint64_t wfirst = p_t * kSegTile;
#pragma unroll 1
            for (int s = 0; s < kSegSubs; s++) {
                const uint32_t m = s_mask[warp][p_buf][s][lane], hb = m & 0xFu, hp = m >> 4;
                if (!__any_sync(0xffffffffu, hb)) continue;
                uint32_t below = s_run[warp][p_buf][s];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t vb = __ballot_sync(0xffffffffu, (hb >> q) & 1u), vp = __ballot_sync(0xffffffffu, (hp >> q) & 1u);
                    below += __popc(vb & lt) | (__popc(vp & lt) << 16);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if ((hb >> q) & 1u) {
                        const uint32_t lower = (1u << q) - 1u;
                        const uint64_t row = p_exb + (below & 0xFFFFu) + __popc(hb & lower);
                        if (row < a.capacity) {
                            // the barcode is re-read (an L2 hit) rather than kept in registers
                            const uint64_t pos = wfirst + (uint64_t)s * kSegSub + 4 * lane + q;
                            uint64_t *dst = a.tmp_rows + 3 * row;
                            dst[0] = a.src[pos * STRIDE];
                            dst[1] = pos;
                            dst[2] = PAIRS ? a.src[pos * STRIDE + 1] : p_exp + (below >> 16) + __popc(hp & lower);
                        }
                    }
                }
            }
        }
        p_active = false;
    };

    auto claim = [&]() -> uint64_t {  // tile ids are claimed in order
        uint64_t t = 0;
        if (lane == 0) t = atomicAdd(a.counters, 1ull);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    auto load_sub = [&](Keys4<STRIDE> &k, uint64_t tile, int s) {
        const uint64_t sfirst = tile * kSegTile + (uint64_t)s * kSegSub;
        if (sfirst + kSegSub <= a.n) {
            k.load(a.src + sfirst * STRIDE, lane);
        } else {  // ragged end of the input
            const uint32_t count = sfirst < a.n ? (uint32_t)(a.n - sfirst) : 0u;
            k.load_partial(a.src + sfirst * STRIDE, lane, count, endb, endu);
        }
    };
    // key of the element in front of a tile (what lane 0's first element is compared with)
    auto load_front = [&](uint64_t tile, uint64_t &fb, uint64_t &fu) {
        fb = fu = 0;
        if (tile > 0) {
            fb = ldg_stream64(a.src + (tile * kSegTile - 1) * STRIDE);
            fu = ldg_stream64(a.src + (tile * kSegTile - 1) * STRIDE + 1);
        }
    };

    uint64_t t = claim();
    if (t >= a.n_tiles) return;
    // (once any warp has seen an order violation the pass is void: everybody stops claiming)
    // The sub-tile stream is continuous across tiles: the next tile is claimed two sub-tiles
    // before the current one ends and its first sub-tile (and front key) are prefetched during
    // the last one, so a tile switch costs no exposed latency.
    Keys4<STRIDE> nxt;
    uint64_t nfb, nfu;
    load_sub(nxt, t, 0);
    load_front(t, nfb, nfu);
    uint32_t buf = 0;
    for (;;) {
        const uint64_t wfirst = t * kSegTile;  // first element of the tile
        // lb/lu: every lane's copy of "the last key of the previous sub-tile of lane 31"; only
        // lane 31's copy is ever consumed (by lane 0, through the rotate below).  Before the
        // first sub-tile it is the key of the element in front of the tile.
        uint64_t lb = nfb, lu = nfu;
        uint64_t t_next = a.n_tiles;
        // flags: element q of lane l in sub-tile s is tile element 128 s + 4 l + q
        uint32_t run = 0, bad = 0;  // run: heads | pairs << 16 so far in the tile
#pragma unroll 1
        for (int s = 0; s < kSegSubs; s++) {
            const Keys4<STRIDE> k = nxt;
            if (s + 1 < kSegSubs) {
                load_sub(nxt, t, s + 1);
            } else if (t_next < a.n_tiles) {
                load_sub(nxt, t_next, 0);
                load_front(t_next, nfb, nfu);
            }
            if (s == kSegSubs - 2) t_next = claim();
            // predecessor of the lane's first element = last element of lane l-1; lane 0 wraps
            // to lane 31, which offers the previous sub-tile's last key instead of its own
            const uint64_t ob = lane == 31 ? lb : k.bc[3], ou = lane == 31 ? lu : k.um[3];
            uint64_t qb = __shfl_sync(0xffffffffu, ob, (lane + 31) & 31u), qu = __shfl_sync(0xffffffffu, ou, (lane + 31) & 31u);
            lb = k.bc[3];
            lu = k.um[3];
            uint32_t mb = 0, mp = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint64_t cb = k.bc[q], cu = k.um[q];
                const uint32_t is_b = cb != qb, is_p = is_b | (cu != qu);
                bad |= (qb > cb) | ((qb == cb) & (qu > cu));
                mb |= is_b << q;
                mp |= is_p << q;
                qb = cb; qu = cu;
            }
            if (s == 0 && lane == 0 && wfirst == 0) { mb |= 1u; mp |= 1u; }  // the very first element
            if (PAIRS) mb = mp;
            s_mask[warp][buf][s][lane] = mb | (mp << 4);
            if (lane == 0) s_run[warp][buf][s] = run;  // heads | pairs << 16 before this sub-tile
            run += __reduce_add_sync(0xffffffffu, __popc(mb) | (__popc(mp) << 16));
            // the previous tile's offsets, if still open: one non-blocking look-back step
            if (p_active && lookback(false)) finish();
        }
        // (padding repeats a key and a missing predecessor reads as (0, 0): neither can look
        // like an order violation)
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.counters + 2, 1ull);

        // ---- publish this tile's counts first (successors only need the aggregate) ----
        if (lane == 0) {
            const uint64_t st = t == 0 ? kStatusPrefix : kStatusAgg;
            desc[2 * t] = st | (uint64_t)(run & 0xFFFFu);
            desc[2 * t + 1] = st | (uint64_t)(run >> 16);
        }
        // at most one tile stays open per warp: force the older one before taking its place
        if (p_active) {
            lookback(true);
            finish();
        }
        __syncwarp();
        p_active = true;
        p_t = t; p_tot = run; p_buf = buf; p_exb = 0; p_exp = 0; p_j = (int64_t)t - 1;
        if (t == 0) finish();  // no predecessors: offsets are zero
        buf ^= 1u;
        if (t_next >= a.n_tiles || *v_flag != 0ull) break;
        t = t_next;
    }
    if (p_active) {
        lookback(true);
        finish();
    }
}

// row r: n_records = start[r+1] - start[r], n_distinct_umi = pairs_before[r+1] - pairs_before[r];
// pair mode: row r = {barcode, umi, multiplicity}
__global__ void __launch_bounds__(kBlockThreads)
k_finalize_rows(const uint64_t *__restrict__ tmp_rows, uint64_t n_rows, uint64_t n,
                const unsigned long long *__restrict__ counters, uint64_t *__restrict__ rows, int pair_mode) {
    const uint64_t total_pairs = counters[3];
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t bc = tmp_rows[3 * r], start = tmp_rows[3 * r + 1], g = tmp_rows[3 * r + 2];
        uint64_t next_start = n, next_g = total_pairs;
        if (r + 1 < n_rows) {
            next_start = tmp_rows[3 * r + 4];
            next_g = tmp_rows[3 * r + 5];
        }
        rows[3 * r] = bc;
        rows[3 * r + 1] = pair_mode ? g : next_start - start;
        rows[3 * r + 2] = pair_mode ? next_start - start : next_g - g;
    }
}

// ============================================================ LSD radix sort (unsorted inputs)
// Elements are STRIDE u64 words (2: (barcode, umi) pair, 3: Record); one pass sorts stably by
// one 8-bit digit of one key word.  Digits on which every key agrees are skipped, so clean
// bc16/umi12 data needs 4 + 3 passes, not 16.
constexpr int kSortTile = 2048;                              // elements per CTA tile
constexpr int kSortItems = kSortTile / kBlockThreads;        // 8 per thread

template <int STRIDE>
struct Elem {
    uint64_t w[STRIDE];
};
template <int STRIDE>
__device__ __forceinline__ Elem<STRIDE> load_elem(const uint64_t *base, uint64_t i) {
    Elem<STRIDE> e;
    if constexpr (STRIDE == 2) {
        const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(base)[i];
        e.w[0] = v.x; e.w[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < STRIDE; k++) e.w[k] = base[i * STRIDE + k];
    }
    return e;
}
template <int STRIDE>
__device__ __forceinline__ void store_elem(uint64_t *base, uint64_t i, const Elem<STRIDE> &e) {
    if constexpr (STRIDE == 2) {
        reinterpret_cast<ulonglong2 *>(base)[i] = make_ulonglong2(e.w[0], e.w[1]);
    } else {
#pragma unroll
        for (int k = 0; k < STRIDE; k++) base[i * STRIDE + k] = e.w[k];
    }
}

// records -> (barcode, umi) pairs (optional), plus OR / AND of every key word: masks[2k] |= w_k,
// masks[2k+1] &= w_k (which bits vary at all)
template <int WORDS>
__global__ void __launch_bounds__(kBlockThreads)
k_key_masks(const uint64_t *__restrict__ recs, uint64_t n, uint64_t *__restrict__ pairs,
            unsigned long long *__restrict__ masks) {
    uint64_t o[WORDS], a[WORDS];
#pragma unroll
    for (int k = 0; k < WORDS; k++) { o[k] = 0; a[k] = ~0ull; }
    uint32_t descents = 0;  // WORDS == 3: records whose third word is smaller than their predecessor's
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t w[WORDS];
#pragma unroll
        for (int k = 0; k < WORDS; k++) {
            w[k] = ldg_stream64(recs + 3 * i + k);
            o[k] |= w[k];
            a[k] &= w[k];
        }
        if (WORDS == 3 && i) descents += w[WORDS - 1] < recs[3 * (i - 1) + 2];
        if (pairs) reinterpret_cast<ulonglong2 *>(pairs)[i] = make_ulonglong2(w[0], w[1]);
    }
    if (WORDS == 3) {
        descents = __reduce_add_sync(0xffffffffu, descents);
        if ((threadIdx.x & 31u) == 0 && descents) atomicAdd(masks + 6, (unsigned long long)descents);
    }
#pragma unroll
    for (int k = 0; k < WORDS; k++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            o[k] |= __shfl_xor_sync(0xffffffffu, o[k], off);
            a[k] &= __shfl_xor_sync(0xffffffffu, a[k], off);
        }
        if ((threadIdx.x & 31u) == 0) {
            atomicOr(masks + 2 * k, o[k]);
            atomicAnd(masks + 2 * k + 1, a[k]);
        }
    }
}

// (barcode, umi) as ONE word, barcode above umi: sorting it is sorting the pair, at 8 bytes per
// element and pass instead of 16
__global__ void __launch_bounds__(kBlockThreads)
k_make_keys(const uint64_t *__restrict__ recs, uint64_t n, uint32_t ub, uint64_t *__restrict__ keys) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        keys[i] = (ldg_stream64(recs + 3 * i) << ub) | ldg_stream64(recs + 3 * i + 1);
}
__global__ void __launch_bounds__(kBlockThreads)
k_keys_to_pairs(const uint64_t *__restrict__ keys, uint64_t n, uint32_t ub, uint64_t *__restrict__ pairs) {
    const uint64_t umask = (1ull << ub) - 1ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = ldg_stream64(keys + i);
        reinterpret_cast<ulonglong2 *>(pairs)[i] = make_ulonglong2(k >> ub, k & umask);
    }
}

// per-tile digit histogram, stored digit-major: hist[d * n_tiles + tile]
template <int STRIDE>
__global__ void __launch_bounds__(kBlockThreads)
k_radix_hist(const uint64_t *__restrict__ in, uint64_t n, uint32_t word, uint32_t shift,
             uint32_t *__restrict__ hist, uint64_t n_tiles) {
    __shared__ uint32_t h[256];
    const uint64_t tile_id = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t first = tile_id * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, n - first);
    for (uint32_t i = threadIdx.x; i < count; i += kBlockThreads) {
        const uint64_t key = in[(first + i) * STRIDE + word];
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
    // offsets are kept relative to the digit (32 bits suffice for n < 2^32); the digit base is
    // added from digit_total by the scatter kernel
    uint64_t run = part[tid] - s;
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t c = row[i];
        row[i] = (uint32_t)run;
        run += c;
    }
    if (tid == kBlockThreads - 1) digit_total[blockIdx.x] = part[tid];
}

// exclusive scan of one u64 per thread over the 256 threads of a CTA: warp shuffles + one barrier
// (the Hillis-Steele form in shared memory cost 16 barriers; a 488-tile pass over 10^6 table rows ran
// 36 us, all of it this latency)
__device__ __forceinline__ uint64_t block_excl_scan64(uint64_t v, uint64_t *tmp) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    uint64_t before = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; w++)
        if ((uint32_t)w < warp) before += tmp[w];
    return before + inc - v;
}

// stable scatter of one tile by one 8-bit digit.  The tile is first reordered in shared memory
// (its elements sorted by the digit), then written out in that order: neighbouring threads
// store neighbouring elements of the same digit run, i.e. to consecutive global addresses,
// instead of 2048 scattered 16/24-byte stores.
template <int STRIDE>
__global__ void __launch_bounds__(kBlockThreads)
k_radix_scatter(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n, uint32_t word,
                uint32_t shift, const uint32_t *__restrict__ hist, const uint64_t *__restrict__ digit_total,
                uint64_t n_tiles) {
    extern __shared__ __align__(16) uint64_t stage[];  // kSortTile elements, reordered
    __shared__ uint32_t warp_cnt[kWarpsPerBlock][256];
    __shared__ uint32_t tile_off[256];                 // first tile-local position of each digit
    __shared__ uint64_t base[256];                     // global position of this tile's first element of each digit
    __shared__ uint64_t scan_tmp[kWarpsPerBlock];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t tile_id = blockIdx.x;
    for (int w = 0; w < kWarpsPerBlock; w++) warp_cnt[w][tid] = 0;
    {   // exclusive scan of the 256 digit totals -> global base of each digit (+ this tile's offset)
        const uint64_t v = digit_total[tid];
        const uint64_t excl = block_excl_scan64(v, scan_tmp);
        base[tid] = excl + hist[(uint64_t)tid * n_tiles + tile_id];
    }
    __syncthreads();

    const uint64_t first = tile_id * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, n - first);
    // warp w owns elements [w*256, w*256+256) of the tile; item k of lane l is element w*256 + 32k + l
    Elem<STRIDE> el[kSortItems];
    uint32_t rank[kSortItems];
    uint32_t dig[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const uint32_t i = warp * (kSortTile / kWarpsPerBlock) + 32 * k + lane;
        const bool live = i < count;
        if (live) el[k] = load_elem<STRIDE>(in, first + i);
        const uint32_t d = live ? (uint32_t)((el[k].w[word] >> shift) & 0xFFu) : 0x100u;  // 0x100: no element
        dig[k] = d;
        // lanes with the same digit: nine ballots (0.93 clocks per key and SM against 1.84 for
        // match.any, tools/k4lab2.cu)
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int bit = 0; bit < 9; bit++) {
            const uint32_t m = __ballot_sync(0xffffffffu, (d >> bit) & 1u);
            peers &= ((d >> bit) & 1u) ? m : ~m;
        }
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
    {   // per digit: exclusive scan over the 8 warps (thread = digit), then over the digits
        uint32_t run = 0;
        for (int w = 0; w < kWarpsPerBlock; w++) {
            const uint32_t c = warp_cnt[w][tid];
            warp_cnt[w][tid] = run;
            run += c;
        }
        tile_off[tid] = (uint32_t)block_excl_scan64(run, scan_tmp);  // first tile-local position of digit `tid`
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        if (dig[k] < 0x100u) {
            const uint32_t pos = tile_off[dig[k]] + warp_cnt[warp][dig[k]] + rank[k];
            store_elem<STRIDE>(stage, pos, el[k]);
        }
    }
    __syncthreads();
    if constexpr (STRIDE == 3) {
        // 24-byte elements leave word by word: neighbouring threads store neighbouring 8-byte words of a
        // digit run (256 contiguous bytes per warp instruction) instead of three 8-byte stores per
        // thread at a 24-byte stride
        for (uint32_t x = tid; x < 3 * count; x += kBlockThreads) {
            const uint32_t j = x / 3u, k = x - 3u * j;
            const uint32_t d = (uint32_t)((stage[3 * j + word] >> shift) & 0xFFu);
            out[3 * (base[d] + (j - tile_off[d])) + k] = stage[x];
        }
    } else {
        for (uint32_t j = tid; j < count; j += kBlockThreads) {
            const Elem<STRIDE> e = load_elem<STRIDE>(stage, j);
            const uint32_t d = (uint32_t)((e.w[word] >> shift) & 0xFFu);
            store_elem<STRIDE>(out, base[d] + (j - tile_off[d]), e);
        }
    }
}

// ------------------------------------------------------------------ one-sweep form of a digit pass
// The three-kernel pass reads every element twice (histogram, scatter) and round-trips a
// [256][tiles] offset matrix through a scan kernel: 72 B per 24-byte element.  One sweep (Adinets &
// Merrill): ALL digit histograms of the sort come from one pass over the keys (k_hist_all), and a
// digit pass is ONE kernel that ranks its tile, publishes the tile's 256 digit counts and obtains
// its global offsets by decoupled look-back over the tiles before it — 48 B per element.  Tiles take
// their number from an atomic ticket, so a tile only ever waits on tiles that are already running.
// The tile is loaded with coalesced 16-byte loads and stays in shared memory; only digits and ranks
// live in registers (the three-kernel scatter held the elements in registers: 80 of them, 3 CTAs).
constexpr uint32_t kSweepAgg = 1u << 30, kSweepPrefix = 2u << 30, kSweepMask = (1u << 30) - 1u;
constexpr int kSweepMaxPasses = 24;

struct HistAllArgs {
    const uint64_t *in;
    uint64_t n;
    uint32_t n_passes;
    uint32_t word[kSweepMaxPasses], shift[kSweepMaxPasses];
    uint32_t *hist;  // [n_passes][256], zeroed
    unsigned long long *masks;  // MASKS: k_key_masks' seven words (OR / AND per word, descents of the last word)
};

// MASKS: the same pass also produces what k_key_masks produces (which bits vary at all, whether the last
// word ever descends).  ibu_gpu_sort_records guesses the digits to histogram from a sample, and this
// kernel's exact masks confirm the guess: one pass over the records instead of two.
template <int STRIDE, bool MASKS>
__global__ void __launch_bounds__(kBlockThreads) k_hist_all(const HistAllArgs a) {
    extern __shared__ uint32_t h[];  // [n_passes][256]
    for (uint32_t i = threadIdx.x; i < a.n_passes * 256; i += kBlockThreads) h[i] = 0;
    __syncthreads();
    uint64_t o[STRIDE], n[STRIDE];
#pragma unroll
    for (int k = 0; k < STRIDE; k++) { o[k] = 0; n[k] = ~0ull; }
    uint32_t descents = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * kBlockThreads + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * kBlockThreads) {
        uint64_t w[STRIDE];
#pragma unroll
        for (int k = 0; k < STRIDE; k++) w[k] = a.in[i * STRIDE + k];
        if (MASKS) {
#pragma unroll
            for (int k = 0; k < STRIDE; k++) { o[k] |= w[k]; n[k] &= w[k]; }
            if (i) descents += w[STRIDE - 1] < a.in[(i - 1) * STRIDE + STRIDE - 1];
        }
        for (uint32_t p = 0; p < a.n_passes; p++) {
            uint64_t key = w[0];
#pragma unroll
            for (int k = 1; k < STRIDE; k++)
                if (a.word[p] == (uint32_t)k) key = w[k];
            atomicAdd(&h[p * 256 + (uint32_t)((key >> a.shift[p]) & 0xFFu)], 1u);
        }
    }
    if (MASKS) {
        descents = __reduce_add_sync(0xffffffffu, descents);
        if ((threadIdx.x & 31u) == 0 && descents) atomicAdd(a.masks + 6, (unsigned long long)descents);
#pragma unroll
        for (int k = 0; k < STRIDE; k++) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                o[k] |= __shfl_xor_sync(0xffffffffu, o[k], off);
                n[k] &= __shfl_xor_sync(0xffffffffu, n[k], off);
            }
            if ((threadIdx.x & 31u) == 0) {
                atomicOr(a.masks + 2 * k, o[k]);
                atomicAnd(a.masks + 2 * k + 1, n[k]);
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < a.n_passes * 256; i += kBlockThreads)
        if (h[i]) atomicAdd(a.hist + i, h[i]);
}

// OR / AND of every word and descents of the third over m records spread evenly over the input (the
// guess that k_hist_all<3, true> then confirms)
__global__ void __launch_bounds__(kBlockThreads)
k_sample_masks(const uint64_t *__restrict__ recs, uint64_t n, uint64_t m, unsigned long long *__restrict__ masks) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t o[3] = {0, 0, 0}, a[3] = {~0ull, ~0ull, ~0ull};
    uint32_t desc = 0;
    if (j < m) {
        const uint64_t i = j * (n / m);
#pragma unroll
        for (int k = 0; k < 3; k++) o[k] = a[k] = recs[3 * i + k];
        if (i + 1 < n) {
            uint64_t w[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                w[k] = recs[3 * i + 3 + k];
                o[k] |= w[k];
                a[k] &= w[k];
            }
            desc = w[2] < recs[3 * i + 2];
        }
    }
    desc = __reduce_add_sync(0xffffffffu, desc);
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            o[k] |= __shfl_xor_sync(0xffffffffu, o[k], off);
            a[k] &= __shfl_xor_sync(0xffffffffu, a[k], off);
        }
    }
    if ((threadIdx.x & 31u) == 0) {
        if (desc) atomicAdd(masks + 6, (unsigned long long)desc);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicOr(masks + 2 * k, o[k]);
            atomicAnd(masks + 2 * k + 1, a[k]);
        }
    }
}

struct SweepArgs {
    const uint64_t *in;
    uint64_t *out;
    uint64_t n;
    uint32_t word, shift;
    const uint32_t *digit_total;  // [256] of this pass
    uint32_t *state;              // [tiles][256], zeroed: 2 flag bits | 30-bit count
    uint32_t *ticket;             // zeroed
    uint32_t *fail;               // set when a look-back gives up
};

// ncu on the first form of this kernel (profiles/r2_sort_onesweep_ncu_summary.json): issue slots 42 % busy,
// long-scoreboard 6 warps per issue — a CTA's phases run strictly one after the other (ticket, three
// batches of four tile loads, ranking, a look-back that walked ONE predecessor per L2 round trip, output)
// with only 3 CTAs per SM to overlap them.  Now the whole tile is requested at once with cp.async (12 x 16
// bytes per thread in flight, no registers), the digit totals are read before anything waits, full tiles
// rank with eight ballots instead of nine, and the look-back reads four predecessors per round trip:
// 1.54 -> 1.42 ms per pass over 10^8 records, 10^9 records 119-121 -> 112-114 ms.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int STRIDE, bool FULL>
__device__ __forceinline__ void sweep_rank(const uint64_t *tile, uint32_t count, uint32_t word, uint32_t shift,
                                           uint32_t (*warp_cnt)[256], uint32_t *rank, uint32_t *dig) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const uint32_t i = warp * (kSortTile / kWarpsPerBlock) + 32 * k + lane;
        const bool live = FULL || i < count;
        const uint32_t d = live ? (uint32_t)((tile[i * STRIDE + word] >> shift) & 0xFFu) : 0x100u;
        dig[k] = d;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int bit = 0; bit < (FULL ? 8 : 9); bit++) {
            const uint32_t m = __ballot_sync(0xffffffffu, (d >> bit) & 1u);
            peers &= ((d >> bit) & 1u) ? m : ~m;
        }
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
}

template <int STRIDE, int LOOK>
__global__ void __launch_bounds__(kBlockThreads, 3) k_onesweep(const SweepArgs a) {
    extern __shared__ __align__(16) uint64_t tile[];  // kSortTile elements as loaded
    __shared__ uint32_t warp_cnt[kWarpsPerBlock][256];
    __shared__ uint32_t tile_off[256], base[256], gdst[kSortTile];
    __shared__ uint16_t srcidx[kSortTile];
    __shared__ uint64_t scan_tmp[kWarpsPerBlock];
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    const uint32_t total = a.digit_total[tid];  // arrives while the ticket and the tile do
    for (int w = 0; w < kWarpsPerBlock; w++) warp_cnt[w][tid] = 0;
    __syncthreads();
    const uint64_t tile_id = s_tile;
    const uint64_t first = tile_id * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, a.n - first);
    {   // the tile, as it lies in memory (its start is 16-byte aligned: kSortTile * STRIDE * 8 bytes per tile)
        const uint32_t words = count * STRIDE;
        const uint4 *src16 = reinterpret_cast<const uint4 *>(a.in + first * STRIDE);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
        for (uint32_t x = tid; x < words / 2; x += kBlockThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * x), "l"(src16 + x) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if ((words & 1u) && tid == 0) tile[words - 1] = a.in[first * STRIDE + words - 1];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // warp w ranks elements [w*256, w*256+256); item k of lane l is element w*256 + 32k + l (stable order)
    uint32_t rank[kSortItems], dig[kSortItems];
    if (count == kSortTile) sweep_rank<STRIDE, true>(tile, count, a.word, a.shift, warp_cnt, rank, dig);
    else sweep_rank<STRIDE, false>(tile, count, a.word, a.shift, warp_cnt, rank, dig);
    __syncthreads();
    // thread = digit: this tile's count, published; offsets of the warps; look-back for the tiles before
    uint32_t c = 0;
    for (int w = 0; w < kWarpsPerBlock; w++) {
        const uint32_t x = warp_cnt[w][tid];
        warp_cnt[w][tid] = c;
        c += x;
    }
    uint32_t before = 0;  // elements with this digit in earlier tiles
    if (tile_id == 0) {
        st_relaxed_u32(a.state + tid, c | kSweepPrefix);
    } else {
        st_relaxed_u32(a.state + tile_id * 256 + tid, c | kSweepAgg);
        int64_t look = (int64_t)tile_id - 1;
        uint32_t spins = 0;
        bool done = false;
        while (!done) {
            // kLook predecessors per L2 round trip (measured: 4 beats 1, 8 and 16 — past 4 the loads that find
            // nothing published yet cost more than the round trips they save)
            constexpr int kLook = LOOK;
            uint32_t v[kLook];
#pragma unroll
            for (int i = 0; i < kLook; i++) v[i] = look - i >= 0 ? ld_relaxed_u32(a.state + (look - i) * 256 + tid) : 0u;
            bool stop = false;
            uint32_t used = 0;
#pragma unroll
            for (int i = 0; i < kLook; i++) {
                const uint32_t flag = v[i] & ~kSweepMask;
                if (!stop) {
                    if (flag == 0) {
                        stop = true;
                    } else {
                        before += v[i] & kSweepMask;
                        used++;
                        if (flag == kSweepPrefix) done = stop = true;  // (tile 0 always publishes a prefix)
                    }
                }
            }
            look -= used;
            if (!done && used == 0) {
                if (++spins > (1u << 24)) {  // watchdog: never hang the GPU (the host then reports an error)
                    *a.fail = 1;
                    break;
                }
                __nanosleep(20);
            }
        }
        st_relaxed_u32(a.state + tile_id * 256 + tid, (before + c) | kSweepPrefix);
    }
    const uint32_t off = (uint32_t)block_excl_scan64(c, scan_tmp);  // first tile-local position of the digit
    __syncthreads();
    const uint32_t gexcl = (uint32_t)block_excl_scan64(total, scan_tmp);  // digits below, whole input
    tile_off[tid] = off;
    base[tid] = gexcl + before - off;  // global position of sorted tile element j with this digit: base + j
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        if (dig[k] < 0x100u) {
            const uint32_t i = warp * (kSortTile / kWarpsPerBlock) + 32 * k + lane;
            const uint32_t pos = tile_off[dig[k]] + warp_cnt[warp][dig[k]] + rank[k];
            srcidx[pos] = (uint16_t)i;
            gdst[pos] = base[dig[k]] + pos;
        }
    }
    __syncthreads();
    // out, word by word in sorted order: neighbouring threads store neighbouring words of a digit run
    for (uint32_t x = tid; x < STRIDE * count; x += kBlockThreads) {
        const uint32_t j = x / STRIDE, k = x - STRIDE * j;
        a.out[(uint64_t)STRIDE * gdst[j] + k] = tile[STRIDE * (uint32_t)srcidx[j] + k];
    }
}

// Measured and dropped: the look-back run UNDER the ranking (the tile's counts taken first from a plain
// shared-memory histogram and published a ranking phase earlier; one look-back round trip per ranked item):
// 11.08 against 10.97 ms per 10^8 records, 109.0 against 107.9 ms per 10^9 — ncu's samples sit between the two
// barriers around the look-back (38 %), but the walk is not what the tile waits for.

// ============================================================ hash aggregation (unsorted inputs)
// Unsorted records are first folded into distinct (barcode, umi, multiplicity) pairs with an
// open-addressing table in HBM/L2 (linear probing, 32-byte slots {barcode, umi, count, pad},
// 128-bit CAS to claim a slot).  Duplicate-heavy data (the common case: PCR duplicates, the
// reference's own example pattern) shrinks by the duplication factor before anything is
// sorted; the few distinct pairs are then sorted and counted weighted.  All-distinct data is
// detected on a sample and goes straight to the radix sort instead.
#define kHashEmpty 0xFFFFFFFFFFFFFFFFull
constexpr uint32_t kHashMaxProbe = 256;
// start size / growth factor: 1-8 Mi slots and x2 / x4 all measured within 10 % of each other on
// 10^9 records (the insert rate is bound by L2 transactions: one 16-byte read + one RED per record)
constexpr uint64_t kHashStartSlots = 8ull << 20;
constexpr uint64_t kHashGrowth = 4;

struct HashArgs {
    const uint64_t *recs;
    uint64_t n;
    uint64_t *table;   // (mask + 1) slots of 4 u64, memset to 0xFF (count is stored minus one)
    uint64_t mask;
    unsigned long long *ctr;  // [0] slots claimed, [1] overflow flag, [2] weight of the all-ones key
    int weighted;
};

__device__ __forceinline__ void cas128(uint64_t *addr, uint64_t c0, uint64_t c1, uint64_t s0, uint64_t s1,
                                       uint64_t &o0, uint64_t &o1) {
    asm volatile(
        "{\n\t.reg .b128 c, s, d;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 s, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 d, [%6], c, s;\n\t"
        "mov.b128 {%0, %1}, d;\n\t}"
        : "=l"(o0), "=l"(o1)
        : "l"(c0), "l"(c1), "l"(s0), "l"(s1), "l"(addr)
        : "memory");
}

__device__ __forceinline__ void hash_insert(const HashArgs &a, uint64_t bc, uint64_t um, uint64_t w, uint32_t &fresh) {
    if (bc == kHashEmpty && um == kHashEmpty) {  // the one key that collides with the empty marker
        atomicAdd(a.ctr + 2, (unsigned long long)w);
        return;
    }
    uint64_t slot = splitmix64(bc ^ splitmix64(um)) & a.mask;
    for (uint32_t probe = 0; probe < kHashMaxProbe; probe++, slot = (slot + 1) & a.mask) {
        uint64_t *s = a.table + 4 * slot;
        uint64_t k0, k1;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(s));
        if (k0 != bc || k1 != um) {
            // a half that reads as "empty" is either an empty slot, a torn view of a slot being
            // claimed, or a key with an all-ones word: the CAS is the authoritative read
            if (k0 != kHashEmpty && k1 != kHashEmpty) continue;  // another key lives here
            cas128(s, kHashEmpty, kHashEmpty, bc, um, k0, k1);
            if (k0 == kHashEmpty && k1 == kHashEmpty) fresh++;   // claimed
            else if (k0 != bc || k1 != um) continue;             // lost the race to another key
        }
        atomicAdd(reinterpret_cast<unsigned long long *>(s + 2), (unsigned long long)w);
        return;
    }
    atomicOr(a.ctr + 1, 1ull);  // table too full
}

__global__ void __launch_bounds__(kBlockThreads) k_hash_insert(const HashArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t total_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_tiles = a.n / kSegSub;
    uint32_t fresh = 0;
    for (uint64_t t = gwarp; t < n_tiles; t += total_warps) {  // lane l owns records 4l..4l+3 of the tile
        const uint8_t *p = reinterpret_cast<const uint8_t *>(a.recs) + t * (kSegSub * 24) + lane * 96;
        const u64x4 v0 = ldg_stream256(p), v1 = ldg_stream256(p + 32), v2 = ldg_stream256(p + 64);
        hash_insert(a, v0.x, v0.y, a.weighted ? v0.z : 1ull, fresh);
        hash_insert(a, v0.w, v1.x, a.weighted ? v1.y : 1ull, fresh);
        hash_insert(a, v1.z, v1.w, a.weighted ? v2.x : 1ull, fresh);
        hash_insert(a, v2.y, v2.z, a.weighted ? v2.w : 1ull, fresh);
    }
    if (gwarp == 0)  // ragged tail
        for (uint64_t i = n_tiles * kSegSub + lane; i < a.n; i += 32)
            hash_insert(a, a.recs[3 * i], a.recs[3 * i + 1], a.weighted ? a.recs[3 * i + 2] : 1ull, fresh);
    fresh = __reduce_add_sync(0xffffffffu, fresh);
    if (lane == 0 && fresh) atomicAdd(a.ctr, (unsigned long long)fresh);
}

// occupied slots -> dense (barcode, umi, count) records, order unspecified
__global__ void __launch_bounds__(kBlockThreads)
k_hash_compact(const uint64_t *__restrict__ table, uint64_t slots, uint64_t *__restrict__ out,
               unsigned long long *__restrict__ cursor) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < slots; base += step) {  // warp-uniform trips
        const uint64_t i = base + threadIdx.x;
        u64x4 v{kHashEmpty, kHashEmpty, 0, 0};
        if (i < slots) v = ldg_stream256(table + 4 * i);
        const bool live = !(v.x == kHashEmpty && v.y == kHashEmpty);
        const uint32_t m = __ballot_sync(0xffffffffu, live);
        if (!m) continue;
        unsigned long long pos = 0;
        if (lane == 0) pos = atomicAdd(cursor, (unsigned long long)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
        if (live) { out[3 * pos] = v.x; out[3 * pos + 1] = v.y; out[3 * pos + 2] = v.z + 1ull; }
    }
}

// growth: every pair of the old table is re-inserted, with its count as the weight
__global__ void __launch_bounds__(kBlockThreads)
k_hash_rehash(const uint64_t *__restrict__ old_table, uint64_t old_slots, const HashArgs a) {
    uint32_t fresh = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_slots;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const u64x4 v = ldg_stream256(old_table + 4 * i);
        if (!(v.x == kHashEmpty && v.y == kHashEmpty)) hash_insert(a, v.x, v.y, v.z + 1ull, fresh);
    }
    fresh = __reduce_add_sync(0xffffffffu, fresh);
    if ((threadIdx.x & 31u) == 0 && fresh) atomicAdd(a.ctr, (unsigned long long)fresh);
}

// ---- owner partition of a pair table (the send side of the multi-GPU exchange) ----
// owner(barcode) = splitmix64(barcode) % world: every barcode's pairs meet on one rank.
__device__ __forceinline__ uint32_t owner_of(uint64_t barcode, uint32_t world) {
    return (uint32_t)(splitmix64(barcode) % world);
}

__global__ void __launch_bounds__(kBlockThreads)
k_owner_count(const uint64_t *__restrict__ recs, uint64_t n, uint32_t world, unsigned long long *__restrict__ counts) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&h[owner_of(recs[3 * i], world)], 1u);
    __syncthreads();
    if (threadIdx.x < world && h[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

__global__ void __launch_bounds__(kBlockThreads)
k_owner_scatter(const uint64_t *__restrict__ recs, uint64_t n, uint32_t world,
                unsigned long long *__restrict__ cursor /* preset to the bucket offsets */, uint64_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += step) {  // warp-uniform trip count
        const uint64_t i = base + threadIdx.x;
        const bool live = i < n;
        uint64_t b = 0, u = 0, c = 0;
        if (live) { b = recs[3 * i]; u = recs[3 * i + 1]; c = recs[3 * i + 2]; }
        const uint32_t o = live ? owner_of(b, world) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, o);  // one atomic per owner per warp
        const uint32_t leader = __ffs(peers) - 1;
        unsigned long long pos = 0;
        if (live && lane == leader) pos = atomicAdd(cursor + o, (unsigned long long)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(peers & ((1u << lane) - 1u));
        if (live) { out[3 * pos] = b; out[3 * pos + 1] = u; out[3 * pos + 2] = c; }
    }
}

// weighted tables (merging per-shard pair tables): a row's count is the sum of the `index`
// words (multiplicities) of the records in its run instead of the run length.  One warp per row.
__global__ void __launch_bounds__(kBlockThreads)
k_sum_weights(const uint64_t *__restrict__ tmp_rows, uint64_t n_rows, uint64_t n,
              const uint64_t *__restrict__ src, uint64_t *__restrict__ rows, int count_word) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * kWarpsPerBlock;
    for (uint64_t r = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); r < n_rows; r += warps) {
        const uint64_t start = tmp_rows[3 * r + 1], end = r + 1 < n_rows ? tmp_rows[3 * r + 4] : n;
        uint64_t s = 0;
        for (uint64_t i = start + lane; i < end; i += 32) s += src[3 * i + 2];
        s = warp_sum64(s);
        if (lane == 0) rows[3 * r + count_word] = s;
    }
}

// Scratch for one call: carved from the context's grow-only arena, falling back to cudaMalloc
// (freed on scope exit) when the arena is too small, e.g. for the exact-capacity retry.
struct Scratch {
    ibu_gpu_ctx *ctx;
    std::vector<void *> owned;
    explicit Scratch(ibu_gpu_ctx *c) : ctx(c) {}
    ~Scratch() {
        for (void *p : owned) cudaFree(p);
    }
    template <class T>
    cudaError_t alloc(T **out, size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (bytes == 0) bytes = 256;
        if (ctx->arena_off + bytes <= ctx->arena_cap) {
            *out = (T *)((uint8_t *)ctx->arena_base + ctx->arena_off);
            ctx->arena_off += bytes;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (getenv("IBU_B200_TRACE_ALLOC")) fprintf(stderr, "[ibu trace] arena overflow: cudaMalloc of %zu bytes\n", bytes);
        if (e == cudaSuccess) owned.push_back(p);
        *out = (T *)p;
        return e;
    }
};

// Make the arena at least `bytes` large and empty.  Only called while nothing carved from it
// is live (start of a call / between the sorted probe and the sort path).
static cudaError_t arena_reset(ibu_gpu_ctx *ctx, size_t bytes) {
    ctx->arena_off = 0;
    if (ctx->arena_cap >= bytes) return cudaSuccess;
    if (ctx->arena_base) cudaFree(ctx->arena_base);
    ctx->arena_base = nullptr;
    ctx->arena_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->arena_base, bytes);
    if (getenv("IBU_B200_TRACE_ALLOC")) fprintf(stderr, "[ibu trace] arena grows to %.3f GB\n", bytes / 1e9);
    if (e == cudaSuccess) ctx->arena_cap = bytes;
    return e;
}

static uint64_t seg_capacity(uint64_t n) { return std::min<uint64_t>(n, 8ull << 20); }
static size_t seg_scratch_bytes(uint64_t n) {
    const uint64_t n_tiles = (n + kSegSubsDefault * kSegSub - 1) / (kSegSubsDefault * kSegSub);
    return n_tiles * 16 + seg_capacity(n) * 24 + 8 * 256;
}

// Runs the segment pass over `src` (stride 3 or 2).  On success *rows_out (device, owned by the
// caller) holds *n_rows rows.  *unsorted is set when the order check failed (no rows then).

// Runs the segment pass over `src` (stride 3 or 2).  On success *rows_out (device, cudaMalloc'ed,
// owned by the caller) holds *n_rows rows of 3 u64: barcode rows {barcode, n_records,
// n_distinct_umi} or, in pair mode, {barcode, umi, multiplicity}.  `weighted` (stride 3 only):
// counts are sums of the records' index words.  *unsorted is set when the order check failed.
static int segment_pass(ibu_gpu_ctx *ctx, const uint64_t *src, int stride, uint64_t n, cudaStream_t s,
                        bool pair_mode, bool weighted, uint64_t **rows_out, uint64_t *n_rows, uint64_t *n_pairs,
                        bool *unsorted, ibu_error_t *err, uint64_t rows_hint = 0) {
    Trace tr;
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    *unsorted = false;
    const uint64_t tile = (uint64_t)kSegSub * kSegSubsDefault;
    const uint64_t n_tiles = (n + tile - 1) / tile;
    Scratch sc(ctx);
    ulonglong2 *desc;
    uint64_t *tmp_rows;
    unsigned long long *counters;
    IBU_CUDA(sc.alloc(&desc, n_tiles * 16));
    IBU_CUDA(sc.alloc(&counters, 4 * 8));
    void (*kern)(const SegArgs) =
        stride == 3 ? (pair_mode ? k_segments<3, kSegSubsDefault, true> : k_segments<3, kSegSubsDefault, false>)
                    : (pair_mode ? k_segments<2, kSegSubsDefault, true> : k_segments<2, kSegSubsDefault, false>);
    static const int carve = getenv("IBU_K4_CARVEOUT") ? atoi(getenv("IBU_K4_CARVEOUT")) : -1;  // tuning hook
    if (carve >= 0) IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    int per_sm = 0;
    IBU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, 0));
    // tile ids are claimed in order at run time, so a waiting tile only ever waits on tiles
    // whose warps are already running; the grid is one resident wave
    const uint64_t tiles_per_cta = kWarpsPerBlock;
    const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * std::max(per_sm, 1),
                                             (n_tiles + tiles_per_cta - 1) / tiles_per_cta);

    // optimistic: <= 8 Mi rows, unless the caller knows better (a sample said: about this many rows)
    uint64_t capacity = std::max(seg_capacity(n), std::min<uint64_t>(n, rows_hint + rows_hint / 4));
    for (int attempt = 0; attempt < 2; attempt++) {
        IBU_CUDA(sc.alloc(&tmp_rows, capacity * 24));
        IBU_CUDA(cudaMemsetAsync(desc, 0, n_tiles * 16, s));
        IBU_CUDA(cudaMemsetAsync(counters, 0, 4 * 8, s));
        SegArgs a{src, n, n_tiles, desc, tmp_rows, capacity, counters};
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (tr.on) {
            cudaEventCreate(&ev0);
            cudaEventCreate(&ev1);
            cudaEventRecord(ev0, s);
        }
        kern<<<grid, kBlockThreads, 0, s>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_segments");
        if (tr.on) {
            cudaEventRecord(ev1, s);
            cudaEventSynchronize(ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, ev0, ev1);
            fprintf(stderr, "[ibu trace] k_segments<%d> grid %d: %.3f ms = %.0f GB/s\n", stride, grid, ms,
                    (double)n * stride * 8 / ms / 1e6);
            cudaEventDestroy(ev0);
            cudaEventDestroy(ev1);
        }
        unsigned long long h[4];
        tr.mark("seg: setup+launch");
        IBU_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        tr.mark("seg: k_segments sync");
        if (h[2] & 2ull)
            return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "barcode_count: look-back watchdog expired");
        if (h[2]) {
            *unsorted = true;
            return IBU_OK;
        }
        if (h[1] > capacity) {  // more rows than the optimistic table: exact re-run
            capacity = h[1];
            continue;
        }
        // owned by the caller (ibu_gpu_table_free / ibu_gpu_free); from the stream-ordered pool,
        // whose cached blocks make this allocation cheap after the first call
        uint64_t *rows = nullptr;
        IBU_CUDA(alloc_result_rows(ctx, &rows, h[1] * 24, s));
        if (h[1]) {
            const uint64_t blocks = (h[1] + kBlockThreads - 1) / kBlockThreads;
            const int fgrid = (int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8);
            k_finalize_rows<<<fgrid, kBlockThreads, 0, s>>>(tmp_rows, h[1], n, counters, rows, pair_mode ? 1 : 0);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (weighted) {
                const uint64_t wblocks = (h[1] + kWarpsPerBlock - 1) / kWarpsPerBlock;
                k_sum_weights<<<(int)std::min<uint64_t>(wblocks, (uint64_t)ctx->sm_count * 8), kBlockThreads, 0, s>>>(
                    tmp_rows, h[1], n, src, rows, pair_mode ? 2 : 1);
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                cudaFree(rows);
                return cuda_fail(err, e, "k_finalize_rows");
            }
        }
        tr.mark("seg: rows malloc+finalize");
        *rows_out = rows;
        *n_rows = h[1];
        *n_pairs = pair_mode ? h[1] : h[3];
        return IBU_OK;
    }
    return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "barcode table capacity retry failed");
}

// OR/AND masks of the first WORDS words of every record (and, optionally, the (barcode, umi)
// pairs extracted to `pairs`): vary[k] = bits of word k on which the records disagree.
template <int WORDS>
static int key_masks(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, uint64_t *pairs, cudaStream_t s,
                     Scratch &sc, uint64_t vary[3], ibu_error_t *err, uint64_t *third_word_descents = nullptr,
                     uint64_t *or_masks = nullptr) {
    unsigned long long *masks;
    IBU_CUDA(sc.alloc(&masks, 7 * 8));
    const unsigned long long init[7] = {0ull, ~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
    IBU_CUDA(cudaMemcpyAsync(masks, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
    k_key_masks<WORDS><<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8), kBlockThreads, 0, s>>>(
        recs, n, pairs, masks);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_key_masks");
    unsigned long long m[7];
    IBU_CUDA(cudaMemcpyAsync(m, masks, sizeof(m), cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; k++) vary[k] = k < WORDS ? (m[2 * k] ^ m[2 * k + 1]) : 0;
    if (third_word_descents) *third_word_descents = m[6];
    if (or_masks)
        for (int k = 0; k < 3; k++) or_masks[k] = k < WORDS ? m[2 * k] : 0;
    return IBU_OK;
}

static int count_passes(const uint64_t vary[3], const int *key_order, int n_keys) {
    int p = 0;
    for (int k = 0; k < n_keys; k++)
        for (uint32_t shift = 0; shift < 64; shift += 8)
            if ((vary[key_order[k]] >> shift) & 0xFFull) p++;
    return p;
}

static bool sweep_eligible(uint64_t n, const void *in, const void *a, const void *b) {
    static const bool sweep_off = getenv("IBU_B200_ONESWEEP") && getenv("IBU_B200_ONESWEEP")[0] == '0';  // tuning
    return !sweep_off && n < (1ull << 30) && (((uintptr_t)in | (uintptr_t)a | (uintptr_t)b) & 15u) == 0;
}

// LSD radix sort of n elements of STRIDE words.  key_order lists the key words from the least
// to the most significant.  `in` is only read; passes ping-pong between `first_dst` and `other`
// (the first pass writes first_dst).  *result is where the sorted elements end up (`in` itself
// when no digit varies).
template <int STRIDE>
static int radix_sort(ibu_gpu_ctx *ctx, const uint64_t *in, uint64_t *first_dst, uint64_t *other, uint64_t n,
                      const uint64_t vary[3], const int *key_order, int n_keys, cudaStream_t s, Scratch &sc,
                      const uint64_t **result, ibu_error_t *err, const HistAllArgs *pre = nullptr) {
    if (n >= (1ull << 32))  // per-digit tile offsets are kept in 32 bits
        return set_error(err, IBU_ERR_ARG, 0, n, 0, "device sort supports fewer than 2^32 elements per call");
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    if (sweep_eligible(n, in, first_dst, other)) {
        // ---- one sweep: all histograms first, then one kernel per digit pass ----
        HistAllArgs h{};
        h.in = in;
        h.n = n;
        for (int k = 0; k < n_keys; k++)
            for (uint32_t shift = 0; shift < 64; shift += 8)
                if ((vary[key_order[k]] >> shift) & 0xFFull) {
                    h.word[h.n_passes] = (uint32_t)key_order[k];
                    h.shift[h.n_passes] = shift;
                    h.n_passes++;
                }
        if (h.n_passes == 0) {
            *result = in;
            return IBU_OK;
        }
        uint32_t *state, *ctl;
        const uint32_t *totals[kSweepMaxPasses];  // the 256 digit totals of each pass
        IBU_CUDA(sc.alloc(&state, n_tiles * 1024));
        IBU_CUDA(sc.alloc(&ctl, 256));  // [0] ticket, [1] fail
        IBU_CUDA(cudaMemsetAsync(ctl, 0, 256, s));
        if (pre) {  // the caller's histograms cover these passes (it checked) and were counted over `in`
            for (uint32_t p = 0; p < h.n_passes; p++) {
                totals[p] = nullptr;
                for (uint32_t q = 0; q < pre->n_passes; q++)
                    if (pre->word[q] == h.word[p] && pre->shift[q] == h.shift[p]) totals[p] = pre->hist + q * 256;
                if (!totals[p]) return set_error(err, IBU_ERR_ARG, 0, p, 0, "device sort: a pass without a histogram");
            }
        } else {
            IBU_CUDA(sc.alloc(&h.hist, (size_t)h.n_passes * 1024));
            IBU_CUDA(cudaMemsetAsync(h.hist, 0, (size_t)h.n_passes * 1024, s));
            const uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
            k_hist_all<STRIDE, false><<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8), kBlockThreads,
                                        h.n_passes * 1024, s>>>(h);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_hist_all");
            for (uint32_t p = 0; p < h.n_passes; p++) totals[p] = h.hist + p * 256;
        }
        // predecessors per look-back round trip: 4 (10^9 records: 108.0 ms; 8: 108.7; 16: 113.9; one at a time, the first form: 119)
        auto sweep = k_onesweep<STRIDE, 4>;
        IBU_CUDA(cudaFuncSetAttribute(sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortTile * STRIDE * 8));
        const uint64_t *src = in;
        uint64_t *dst = first_dst, *spare = other;
        for (uint32_t p = 0; p < h.n_passes; p++) {
            IBU_CUDA(cudaMemsetAsync(state, 0, n_tiles * 1024, s));
            IBU_CUDA(cudaMemsetAsync(ctl, 0, 4, s));
            SweepArgs a{src, dst, n, h.word[p], h.shift[p], totals[p], state, ctl, ctl + 1};
            sweep<<<(int)n_tiles, kBlockThreads, kSortTile * STRIDE * 8, s>>>(a);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_onesweep");
            src = dst;
            std::swap(dst, spare);
        }
        uint32_t failed = 0;
        IBU_CUDA(cudaMemcpyAsync(&failed, ctl + 1, 4, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        if (failed) return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "device sort: look-back watchdog expired");
        *result = src;
        return IBU_OK;
    }
    uint64_t *digit_total;
    uint32_t *hist;
    IBU_CUDA(sc.alloc(&hist, 256 * n_tiles * 4));
    IBU_CUDA(sc.alloc(&digit_total, 256 * 8));
    IBU_CUDA(cudaFuncSetAttribute(k_radix_scatter<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSortTile * STRIDE * 8));
    const uint64_t *src = in;
    uint64_t *dst = first_dst, *spare = other;
    for (int k = 0; k < n_keys; k++) {
        const uint32_t word = (uint32_t)key_order[k];
        for (uint32_t shift = 0; shift < 64; shift += 8) {
            if (((vary[word] >> shift) & 0xFFull) == 0) continue;  // every key agrees on this digit
            k_radix_hist<STRIDE><<<(int)n_tiles, kBlockThreads, 0, s>>>(src, n, word, shift, hist, n_tiles);
            k_radix_scan<<<256, kBlockThreads, 0, s>>>(hist, n_tiles, digit_total);
            k_radix_scatter<STRIDE><<<(int)n_tiles, kBlockThreads, kSortTile * STRIDE * 8, s>>>(
                src, dst, n, word, shift, hist, digit_total, n_tiles);
            g_launches.fetch_add(3, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "radix pass");
            src = dst;
            std::swap(dst, spare);
        }
    }
    *result = src;
    return IBU_OK;
}

// Table of an unsorted input: sort by (barcode, umi), then the streaming segment pass.
static int unsorted_table(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, bool pair_mode,
                          bool weighted, uint64_t **rows, uint64_t *n_rows, uint64_t *n_pairs, ibu_error_t *err,
                          uint64_t rows_hint = 0) {
    Scratch sc(ctx);
    static const int order[2] = {1, 0};  // umi is the minor key, barcode the major one
    uint64_t vary[3];
    const uint64_t *sorted = nullptr;
    int stride;
    if (!weighted) {  // only the keys are needed
        uint64_t *a, *b, ors[3];
        IBU_CUDA(sc.alloc(&a, n * 16));
        IBU_CUDA(sc.alloc(&b, n * 16));
        if (int rc = key_masks<2>(ctx, recs, n, nullptr, s, sc, vary, err, nullptr, ors)) return rc;
        const uint32_t bb = ors[0] ? 64 - (uint32_t)__builtin_clzll(ors[0]) : 1, ub = ors[1] ? 64 - (uint32_t)__builtin_clzll(ors[1]) : 1;
        const int grid = (int)std::min<uint64_t>((n + kBlockThreads - 1) / kBlockThreads, (uint64_t)ctx->sm_count * 16);
        if (bb + ub <= 64 && ub < 64) {
            // the pair fits one word (bc16/umi12: 56 bits): 8-byte keys through the digit passes
            // (24 instead of 48 bytes per element and pass), pairs again for the segment pass
            uint64_t *k0 = a, *k1 = a + n;  // the two halves of `a` ping-pong; `b` receives the pairs
            k_make_keys<<<grid, kBlockThreads, 0, s>>>(recs, n, ub, k0);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_make_keys");
            const uint64_t vary1[3] = {(vary[0] << ub) | vary[1], 0, 0};
            static const int order1[1] = {0};
            const uint64_t *skeys = nullptr;
            if (int rc = radix_sort<1>(ctx, k0, k1, k0, n, vary1, order1, 1, s, sc, &skeys, err)) return rc;
            k_keys_to_pairs<<<grid, kBlockThreads, 0, s>>>(skeys, n, ub, b);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_keys_to_pairs");
            sorted = b;
        } else {  // 16-byte (barcode, umi) pairs
            if (int rc = key_masks<2>(ctx, recs, n, a, s, sc, vary, err)) return rc;
            if (int rc = radix_sort<2>(ctx, a, b, a, n, vary, order, 2, s, sc, &sorted, err)) return rc;
        }
        stride = 2;
    } else {  // the multiplicity (index word) travels with its key: sort whole records
        uint64_t *a, *b;
        IBU_CUDA(sc.alloc(&a, n * 24));
        // by partition where that suits the keys (ibu_gpu_sort_records' first choice: the multiplicity is sorted
        // as a third key, which does no harm) — the weighted near-distinct input of a multi-GPU owner count
        bool by_partition = false;
        const char *msd_env = getenv("IBU_B200_SORT_MSD");
        if (n >= (1ull << 20) && !(msd_env && msd_env[0] == '0') && (((uintptr_t)recs | (uintptr_t)a) & 31u) == 0) {
            K4Sample smp;
            ibu_error_t attempt{};
            int rc = k4_sample(ctx, recs, n, s, &smp, &attempt);
            if (rc == IBU_OK) rc = k4_sort_records_msd(ctx, recs, n, smp, a, s, &by_partition, &attempt);
            if (rc != IBU_OK) {
                if (rc != IBU_ERR_CUDA || attempt.sys != (int)cudaErrorMemoryAllocation) {
                    if (err) *err = attempt;
                    return rc;
                }
                cudaGetLastError();
                by_partition = false;
            }
        }
        if (by_partition) {
            sorted = a;
        } else {
            IBU_CUDA(sc.alloc(&b, n * 24));
            if (int rc = key_masks<2>(ctx, recs, n, nullptr, s, sc, vary, err)) return rc;
            if (int rc = radix_sort<3>(ctx, recs, a, b, n, vary, order, 2, s, sc, &sorted, err)) return rc;
        }
        stride = 3;
    }
    bool still_unsorted = false;
    if (int rc = segment_pass(ctx, sorted, stride, n, s, pair_mode, weighted, rows, n_rows, n_pairs,
                              &still_unsorted, err, rows_hint))
        return rc;
    if (still_unsorted)
        return set_error(err, IBU_ERR_CUDA, 0, 0, 0, "internal error: radix sort left the keys unsorted");
    return IBU_OK;
}


static uint64_t pow2_ceil(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Distinct (barcode, umi, multiplicity) pairs of an unsorted input by hash aggregation.
// Records are inserted a chunk at a time; the table starts at 8 Mi slots (256 MiB) and is
// re-hashed into one four times larger whenever it passes 0.7 load, so duplicate-heavy data
// never pays for a table sized for the worst case.  *use_sort is set (and nothing else
// produced) when the first 4 Mi records are close to all-distinct — aggregation would only add a pass —
// or a table would not fit: the caller then sorts the input directly.
static int hash_aggregate(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, bool weighted, cudaStream_t s,
                          Scratch &sc, uint64_t **pairs, uint64_t *n_pairs, bool *use_sort, ibu_error_t *err) {
    Trace tr;
    *use_sort = false;
    *pairs = nullptr;
    *n_pairs = 0;
    unsigned long long *ctr;  // [0] slots claimed, [1] overflow flag, [2] weight of the all-ones key
    IBU_CUDA(sc.alloc(&ctr, 4 * 8));
    // 8 Mi slots = 256 MiB to start with: the hot part of a low-cardinality table then stays
    // inside L2 and TLB reach (a 1 GiB start measured 40 % slower on the reference pattern)
    uint64_t slots = std::max<uint64_t>(1024, std::min<uint64_t>(pow2_ceil(2 * n), kHashStartSlots));
    uint64_t *table;
    IBU_CUDA(sc.alloc(&table, slots * 32));
    IBU_CUDA(cudaMemsetAsync(table, 0xFF, slots * 32, s));
    IBU_CUDA(cudaMemsetAsync(ctr, 0, 4 * 8, s));
    const int max_grid = ctx->sm_count * 8;
    const uint64_t decide_at = 4ull << 20;  // records after which hash-vs-sort is decided
    uint64_t chunk = 2ull << 20;            // records per launch; grows while few keys are new
    unsigned long long h[3] = {0, 0, 0}, claimed_before = 0;
    bool decided = false;
    for (uint64_t pos = 0; pos < n;) {
        const uint64_t cnt = std::min(chunk, n - pos);
        HashArgs a{recs + 3 * pos, cnt, table, slots - 1, ctr, weighted ? 1 : 0};
        const uint64_t blocks = (cnt / kSegSub + kWarpsPerBlock) / kWarpsPerBlock;
        k_hash_insert<<<(int)std::min<uint64_t>(blocks, max_grid), kBlockThreads, 0, s>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_hash_insert");
        IBU_CUDA(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        tr.mark("hash: insert chunk");
        pos += cnt;
        const uint64_t fresh = h[0] - claimed_before;
        claimed_before = h[0];
        if (h[1]) {  // probe limit hit: the table could not keep up
            *use_sort = true;
            return IBU_OK;
        }
        if (!decided && pos >= decide_at) {
            decided = true;
            if (pos < n && h[0] > pos / 10 * 4) {  // (nearly) all-distinct so far
                *use_sort = true;
                return IBU_OK;
            }
        }
        if (pos >= n) break;
        if (h[0] > slots / 10 * 7) {  // past 0.7 load: grow x4 and re-hash what is there
            const uint64_t bigger = slots * kHashGrowth;
            if (bigger > (1ull << 31)) {
                *use_sort = true;
                return IBU_OK;
            }
            uint64_t *next;
            IBU_CUDA(sc.alloc(&next, bigger * 32));
            IBU_CUDA(cudaMemsetAsync(next, 0xFF, bigger * 32, s));
            IBU_CUDA(cudaMemsetAsync(ctr, 0, 2 * 8, s));  // claimed + overflow; the all-ones weight stays
            HashArgs b{nullptr, 0, next, bigger - 1, ctr, 1};
            k_hash_rehash<<<max_grid, kBlockThreads, 0, s>>>(table, slots, b);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_hash_rehash");
            table = next;
            slots = bigger;
        }
        // a chunk may add at most a quarter of the table while keys are still mostly new; once
        // they are mostly repeats the launches get longer (fewer host round trips)
        if (fresh < cnt / 8) chunk = std::min<uint64_t>(chunk * 2, 64ull << 20);
        else if (fresh > cnt / 2) chunk = std::max<uint64_t>(1ull << 20, std::min<uint64_t>(chunk, slots / 8));
        chunk &= ~3ull;  // chunk starts stay 32-byte aligned
    }
    if (h[0] > slots / 10 * 9) {  // the last chunk overfilled the table: probes were long, not wrong
        // (correct but slow; nothing to do)
    }
    // compact the occupied slots (plus the all-ones key, if it occurred) into pair records
    const uint64_t k = h[0] + (h[2] ? 1 : 0);
    uint64_t *out;
    IBU_CUDA(sc.alloc(&out, (k ? k : 1) * 24));
    IBU_CUDA(cudaMemsetAsync(ctr, 0, 8, s));
    const uint64_t blocks = (slots + kBlockThreads - 1) / kBlockThreads;
    k_hash_compact<<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 16), kBlockThreads, 0, s>>>(
        table, slots, out, ctr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_hash_compact");
    if (h[2]) {
        const unsigned long long special[3] = {kHashEmpty, kHashEmpty, h[2]};
        IBU_CUDA(cudaMemcpyAsync(out + 3 * h[0], special, sizeof(special), cudaMemcpyHostToDevice, s));
    }
    IBU_CUDA(cudaStreamSynchronize(s));
    tr.mark("hash: compact");
    *pairs = out;
    *n_pairs = k;
    return IBU_OK;
}

static size_t sort_scratch_bytes(uint64_t n, int elem_bytes) {
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    return n * 2 * elem_bytes + n_tiles * 1024 + 16 * 256 + kSweepMaxPasses * 1024 + 1024;
}

// The pre-partition unsorted path: hash-aggregate into a global table (or, when nearly every pair is
// distinct, radix sort 16-byte pairs), then the segment pass.  Caller holds ctx->arena_mutex.
int k4_legacy_unsorted(ibu_gpu_ctx *ctx, const uint64_t *src, uint64_t n, cudaStream_t s, bool pair_mode,
                       bool weighted, uint64_t **rows, uint64_t *n_rows, uint64_t *n_pairs, ibu_error_t *err,
                       double d_est, double r_est) {
    // duplicate-heavy inputs: fold to distinct pairs first, then sort/count only those
    // room for the first table (256 MiB), the compacted pairs and their sort; a larger
    // full-size table falls back to a one-off allocation inside Scratch
    IBU_CUDA(arena_reset(ctx, (768ull << 20) + seg_scratch_bytes(std::min<uint64_t>(n, 8ull << 20))));
    Trace tr;
    tr.mark("table: arena");
    // (a sample that says "nearly every pair is distinct" skips the hash attempt: it would only find that out again)
    bool use_sort = getenv("IBU_B200_NO_HASH") != nullptr || d_est > 0.4 * (double)n;
    const uint64_t rows_hint = (uint64_t)(pair_mode ? d_est : r_est);
    if (!use_sort) {
        Scratch sc(ctx);
        uint64_t *pairs = nullptr, k = 0;
        if (int rc = hash_aggregate(ctx, src, n, weighted, s, sc, &pairs, &k, &use_sort, err)) return rc;
        tr.mark("table: hash aggregate total");
        if (!use_sort) {  // the pairs live in `sc` (arena or fallback allocations) until we return
            int rc = unsorted_table(ctx, pairs, k, s, pair_mode, true, rows, n_rows, n_pairs, err);
            tr.mark("table: sort+count of pairs");
            return rc;
        }
    }
    IBU_CUDA(arena_reset(ctx, sort_scratch_bytes(n, weighted ? 24 : 16) + seg_scratch_bytes(n) +
                              std::min<uint64_t>(n, rows_hint + rows_hint / 4) * 24));
    return unsorted_table(ctx, src, n, s, pair_mode, weighted, rows, n_rows, n_pairs, err, rows_hint);
}

int k4_sort_rows(ibu_gpu_ctx *ctx, const uint64_t *rows, uint64_t n, const uint64_t vary[3], const int *key_order,
                 int n_keys, cudaStream_t s, uint64_t *dst, ibu_error_t *err) {
    if (n == 0) return IBU_OK;
    IBU_CUDA(arena_reset(ctx, sort_scratch_bytes(n, 24)));
    Scratch sc(ctx);
    uint64_t *spare;
    IBU_CUDA(sc.alloc(&spare, n * 24));
    // an odd number of passes must start into dst to end there
    const int passes = count_passes(vary, key_order, n_keys);
    const uint64_t *result = nullptr;
    if (int rc = radix_sort<3>(ctx, rows, (passes & 1) ? dst : spare, (passes & 1) ? spare : dst, n, vary, key_order,
                               n_keys, s, sc, &result, err))
        return rc;
    if (result != dst) IBU_CUDA(cudaMemcpyAsync(dst, result, n * 24, cudaMemcpyDeviceToDevice, s));
    return IBU_OK;
}

// Shared driver of ibu_gpu_barcode_count / ibu_gpu_pair_table.
int k4_build_table(ibu_gpu_ctx *ctx, const ibu_record_t *d_records, uint64_t n, int mode, const K4Hints &hints,
                   bool pair_mode, bool pairs_sorted, bool weighted, cudaStream_t s, uint64_t **rows,
                   uint64_t *n_rows, uint64_t *n_pairs, bool *was_sorted, ibu_error_t *err) {
    *rows = nullptr;
    *n_rows = *n_pairs = 0;
    *was_sorted = false;
    const uint64_t *src = reinterpret_cast<const uint64_t *>(d_records);
    std::lock_guard<std::mutex> lock(ctx->arena_mutex);  // one table build per context at a time
    if (s != ctx->stream) {
        // The build runs on the context's own stream, ordered after the caller's (the call is blocking, so
        // the caller's stream needs nothing back).  Its scratch and the rows it returns
        // (ibu_gpu_table_free releases them on ctx->stream) then always come from and go back to the
        // pool on ONE stream: blocks released on one stream and requested on another are not reused
        // until the driver has looked, and gigabyte blocks then cost fresh mappings (calls of 7 - 200 ms
        // measured for 10^8 near-distinct records, against 4.5 ms on one stream).
        if (!ctx->rows_ev) IBU_CUDA(cudaEventCreateWithFlags(&ctx->rows_ev, cudaEventDisableTiming));
        IBU_CUDA(cudaEventRecord(ctx->rows_ev, s));
        IBU_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->rows_ev, 0));
        s = ctx->stream;
    }
    struct PoolTrace {  // IBU_B200_TRACE_ALLOC: what the device's pool holds when a build starts and ends
        int dev;
        void say(const char *when) const {
            cudaMemPool_t pool;
            unsigned long long reserved = 0, used = 0, thr = 0;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            fprintf(stderr, "[ibu trace] pool %s: reserved %.3f GB used %.3f GB release threshold %llx\n", when, reserved / 1e9, used / 1e9, thr);
        }
        explicit PoolTrace(int d) : dev(d) { say("at entry"); }
        ~PoolTrace() { say("at exit"); }
    };
    std::unique_ptr<PoolTrace> pool_trace(getenv("IBU_B200_TRACE_ALLOC") ? new PoolTrace(ctx->device) : nullptr);
    bool unsorted = mode == 2;
    // Below 64 Ki records everything is launch latency and the round-1 flow is as good.  Above, a
    // sample decides first: unsorted input (the common case: the header's flag is advisory) never
    // pays for the streaming attempt, which cannot stop early enough to be free.
    K4Sample smp;
    const bool use_new = hints.force_path == kPathPartition ||
                         (n >= (1u << 16) && hints.force_path != kPathLegacy && hints.force_path != kPathSort);
    if (use_new && mode != 1) {
        Trace tr;
        if (int rc = k4_sample(ctx, src, n, s, &smp, err)) return rc;
        tr.mark("table: sample");
        if (smp.unordered) unsorted = true;
    }
    if (!unsorted) {
        Trace tr;
        IBU_CUDA(arena_reset(ctx, seg_scratch_bytes(n)));
        if (int rc = segment_pass(ctx, src, 3, n, s, pair_mode, weighted, rows, n_rows, n_pairs, &unsorted, err))
            return rc;
        *was_sorted = !unsorted;
        tr.mark("table: sorted probe total");
    }
    if (unsorted) {
        if (mode == 1) return IBU_OK;  // caller required sorted input: was_sorted = false, no rows
        if (use_new) {
            Trace tr;
            bool handled = false;
            if (int rc = k4_partition_table(ctx, src, n, hints, smp, pair_mode, pairs_sorted, weighted, s, rows, n_rows,
                                            n_pairs, &handled, err))
                return rc;
            tr.mark("table: partition path total");
            if (handled) return IBU_OK;
        }
        double d_est = 0, r_est = 0;
        k4_estimates(smp, n, &d_est, &r_est);
        return k4_legacy_unsorted(ctx, src, n, s, pair_mode, weighted, rows, n_rows, n_pairs, err, d_est, r_est);
    }
    return IBU_OK;
}

}  // namespace ibu

using namespace ibu;

// key-layout hints and forced path carried in the upper bits of `mode` / `flags`
K4Hints ibu::k4_hints_of(int mode) {
    K4Hints h;
    h.bc_len = ((unsigned)mode >> 8) & 0x3Fu;
    h.umi_len = ((unsigned)mode >> 16) & 0x3Fu;
    if (h.bc_len > 32 || h.umi_len > 32) h.bc_len = h.umi_len = 0;
    h.force_path = ((unsigned)mode >> 4) & 3u;
    return h;
}

extern "C" {

int ibu_gpu_barcode_count(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, int mode,
                          ibu_barcode_table_t *table, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !table || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    const bool weighted = (mode & IBU_COUNT_WEIGHTED) != 0;
    const K4Hints hints = k4_hints_of(mode);
    mode &= 7;
    if (mode < 0 || mode > 2) return set_error(err, IBU_ERR_ARG, 0, mode, 0, "mode must be 0, 1 or 2");
    if (((uintptr_t)d_records & 31u) != 0)
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "d_records must be 32-byte aligned");
    memset(table, 0, sizeof(*table));
    table->n_records = n;
    if (n == 0) {
        table->input_was_sorted = 1;
        return IBU_OK;
    }
    DeviceGuard guard(ctx->device);
    uint64_t *rows = nullptr, n_rows = 0, n_pairs = 0;
    bool was_sorted = false;
    if (int rc = k4_build_table(ctx, d_records, n, mode, hints, false, false, weighted, pick_stream(ctx, stream), &rows,
                             &n_rows, &n_pairs, &was_sorted, err))
        return rc;
    table->d_rows = reinterpret_cast<ibu_barcode_row_t *>(rows);
    table->n_rows = n_rows;
    table->n_distinct_pairs = n_pairs;
    table->input_was_sorted = was_sorted ? 1 : 0;
    return IBU_OK;
}

void ibu_gpu_table_free(ibu_gpu_ctx_t *ctx, ibu_barcode_table_t *table) {
    if (!ctx || !table) return;
    if (table->d_rows) {
        DeviceGuard guard(ctx->device);
        if (cudaFreeAsync(table->d_rows, ctx->stream) != cudaSuccess) cudaGetLastError();
    }
    table->d_rows = nullptr;
    table->n_rows = 0;
}

int ibu_gpu_pair_table(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, int weighted,
                       ibu_record_t **d_pairs, uint64_t *n_pairs, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !d_pairs || !n_pairs || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (((uintptr_t)d_records & 31u) != 0)
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "d_records must be 32-byte aligned");
    *d_pairs = nullptr;
    *n_pairs = 0;
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    uint64_t *rows = nullptr, n_rows = 0, np = 0;
    bool was_sorted = false;
    if (int rc = k4_build_table(ctx, d_records, n, 0, k4_hints_of(weighted), true, (weighted & IBU_PAIRS_UNORDERED) == 0,
                             (weighted & IBU_PAIRS_WEIGHTED) != 0, pick_stream(ctx, stream), &rows, &n_rows, &np,
                             &was_sorted, err))
        return rc;
    *d_pairs = reinterpret_cast<ibu_record_t *>(rows);
    *n_pairs = n_rows;
    return IBU_OK;
}

int ibu_gpu_sort_records(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, ibu_record_t *d_sorted,
                         void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (n && (!d_records || !d_sorted))) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (n == 0) return IBU_OK;
    if (d_records == d_sorted) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "d_sorted must not alias d_records");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    std::lock_guard<std::mutex> lock(ctx->arena_mutex);
    const uint64_t *src = reinterpret_cast<const uint64_t *>(d_records);
    uint64_t *dst = reinterpret_cast<uint64_t *>(d_sorted);
    // From 2^20 records on, by partition first (k4_sort_records_msd: 112 bytes moved per record instead of 24 + 48
    // per 8-bit digit); the LSD sort below takes what does not suit it.  IBU_B200_SORT_MSD=0 turns the attempt
    // off, =2 makes an input it gives up on an error (how the tests prove which path ran).
    const char *msd_env = getenv("IBU_B200_SORT_MSD");
    if (n >= (1ull << 20) && !(msd_env && msd_env[0] == '0') && (((uintptr_t)src | (uintptr_t)dst) & 31u) == 0) {
        if (s != ctx->stream) {  // scratch from the stream-ordered pool: always on the context's stream (k4_build_table)
            if (!ctx->rows_ev) IBU_CUDA(cudaEventCreateWithFlags(&ctx->rows_ev, cudaEventDisableTiming));
            IBU_CUDA(cudaEventRecord(ctx->rows_ev, s));
            IBU_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->rows_ev, 0));
            s = ctx->stream;
        }
        K4Sample smp;
        if (int rc = k4_sample(ctx, src, n, s, &smp, err)) return rc;
        bool handled = false;
        ibu_error_t attempt{};
        if (int rc = k4_sort_records_msd(ctx, src, n, smp, dst, s, &handled, &attempt)) {
            // no room for the partition's scratch (about 40 bytes per record): the LSD sort needs 24
            if (rc != IBU_ERR_CUDA || attempt.sys != (int)cudaErrorMemoryAllocation) {
                if (err) *err = attempt;
                return rc;
            }
            cudaGetLastError();
            handled = false;
        }
        if (handled) return IBU_OK;  // (synchronised)
        if (msd_env && msd_env[0] == '2') return set_error(err, IBU_ERR_ARG, 0, n, 0, "sort by partition gave up");
    }
    IBU_CUDA(arena_reset(ctx, sort_scratch_bytes(n, 12)));  // spare + look-back state + histograms: no allocation in steady state
    Scratch sc(ctx);
    uint64_t *spare;
    IBU_CUDA(sc.alloc(&spare, n * 24));
    uint64_t vary[3];
    uint64_t descents = 0;
    static const int order[3] = {2, 1, 0};  // Record's Ord: barcode, then umi, then index (record.rs:58)
    // Which digits vary (and whether the index ever descends) decides the passes, and the passes decide which
    // histograms to count: two scans of the records.  Instead a sample guesses the digits, ONE scan counts
    // their histograms together with the exact masks, and the guess stands when it covers what the masks say
    // (a guessed digit that does not vary after all is just not sorted).  A wrong guess — a bit that is set in
    // a handful of records the sample missed — costs the second scan the two-step form always paid.
    HistAllArgs pre{};
    bool have_pre = false;
    if (n >= (1ull << 20) && sweep_eligible(n, src, dst, spare)) {
        unsigned long long *masks;
        IBU_CUDA(sc.alloc(&masks, 2 * 7 * 8));
        const unsigned long long init[14] = {0ull, ~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
        IBU_CUDA(cudaMemcpyAsync(masks, init, sizeof(init), cudaMemcpyHostToDevice, s));
        const uint64_t m = 1ull << 16;
        k_sample_masks<<<(int)(m / kBlockThreads), kBlockThreads, 0, s>>>(src, n, m, masks);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        unsigned long long g[7];
        IBU_CUDA(cudaMemcpyAsync(g, masks, sizeof(g), cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        uint64_t guess[3];
        for (int k = 0; k < 3; k++) {  // every bit below the highest varying one is taken to vary too
            uint64_t v = g[2 * k] ^ g[2 * k + 1];
            for (int sh = 1; sh < 64; sh <<= 1) v |= v >> sh;
            guess[k] = v;
        }
        const bool guess_index = g[6] != 0;
        pre.in = src;
        pre.n = n;
        pre.masks = masks + 7;
        for (int k = guess_index ? 0 : 1; k < 3; k++)
            for (uint32_t shift = 0; shift < 64; shift += 8)
                if (((guess[order[k]] >> shift) & 0xFFull) && pre.n_passes < (uint32_t)kSweepMaxPasses) {
                    pre.word[pre.n_passes] = (uint32_t)order[k];
                    pre.shift[pre.n_passes] = shift;
                    pre.n_passes++;
                }
        if (pre.n_passes) {
            IBU_CUDA(sc.alloc(&pre.hist, (size_t)pre.n_passes * 1024));
            IBU_CUDA(cudaMemsetAsync(pre.hist, 0, (size_t)pre.n_passes * 1024, s));
            const uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
            k_hist_all<3, true><<<(int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8), kBlockThreads,
                                  pre.n_passes * 1024, s>>>(pre);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaError_t e__ = cudaGetLastError()) return cuda_fail(err, e__, "k_hist_all");
            unsigned long long x[7];
            IBU_CUDA(cudaMemcpyAsync(x, pre.masks, sizeof(x), cudaMemcpyDeviceToHost, s));
            IBU_CUDA(cudaStreamSynchronize(s));
            for (int k = 0; k < 3; k++) vary[k] = x[2 * k] ^ x[2 * k + 1];
            descents = x[6];
            have_pre = true;  // the exact masks, whatever the guess was worth
            bool covered = !descents || guess_index;
            for (int k = descents ? 0 : 1; k < 3 && covered; k++)
                for (uint32_t shift = 0; shift < 64; shift += 8)
                    if ((vary[order[k]] >> shift) & 0xFFull) {
                        bool found = false;
                        for (uint32_t q = 0; q < pre.n_passes; q++)
                            found |= pre.word[q] == (uint32_t)order[k] && pre.shift[q] == shift;
                        covered &= found;
                    }
            if (!covered) pre.n_passes = 0;  // radix_sort counts its own histograms from the exact masks
        }
    }
    if (!have_pre)
        if (int rc = key_masks<3>(ctx, src, n, nullptr, s, sc, vary, err, &descents)) return rc;
    // Records that already come in index order (what a writer that numbers reads as it goes produces)
    // need no index passes: the sort is stable, so ties on (barcode, umi) keep their input order.
    const int *keys = descents ? order : order + 1;
    const int n_keys = descents ? 3 : 2;
    const int passes = count_passes(vary, keys, n_keys);
    const uint64_t *result = nullptr;
    // an odd number of passes must start into d_sorted to end there
    if (int rc = radix_sort<3>(ctx, src, (passes & 1) ? dst : spare, (passes & 1) ? spare : dst, n, vary, keys, n_keys,
                               s, sc, &result, err, pre.n_passes ? &pre : nullptr))
        return rc;
    if (result != dst) IBU_CUDA(cudaMemcpyAsync(dst, result, n * 24, cudaMemcpyDeviceToDevice, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    return IBU_OK;
}

int ibu_gpu_partition_by_owner(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_pairs, uint64_t n, uint32_t world,
                               ibu_record_t *d_out, uint64_t *h_counts, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !h_counts || (n && (!d_pairs || !d_out))) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (world == 0 || world > 256) return set_error(err, IBU_ERR_ARG, 0, world, 0, "world must be 1..256");
    for (uint32_t r = 0; r < world; r++) h_counts[r] = 0;
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    unsigned long long *d_counts = nullptr;
    IBU_CUDA(cudaMallocAsync((void **)&d_counts, 256 * 8, s));  // (cudaMalloc / cudaFree would synchronise the device)
    int rc = IBU_OK;
    do {
        cudaError_t e = cudaMemsetAsync(d_counts, 0, 256 * 8, s);
        const uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
        const int grid = (int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 8);
        const uint64_t *src = reinterpret_cast<const uint64_t *>(d_pairs);
        k_owner_count<<<grid, kBlockThreads, 0, s>>>(src, n, world, d_counts);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        unsigned long long h[256];
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_counts, world * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { rc = cuda_fail(err, e, "k_owner_count"); break; }
        unsigned long long off[256], run = 0;
        for (uint32_t r = 0; r < world; r++) { h_counts[r] = h[r]; off[r] = run; run += h[r]; }
        e = cudaMemcpyAsync(d_counts, off, world * 8, cudaMemcpyHostToDevice, s);
        k_owner_scatter<<<grid, kBlockThreads, 0, s>>>(src, n, world, d_counts, reinterpret_cast<uint64_t *>(d_out));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "k_owner_scatter");
    } while (0);
    if (cudaFreeAsync(d_counts, s) != cudaSuccess) cudaGetLastError();
    return rc;
}

}  // extern "C"
