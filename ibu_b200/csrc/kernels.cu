// kernels.cu — sm_100a kernels of the bulk record path and their stream-ordered launchers.
//
//   K1 k_validate_reduce   24 B/record read            built-in reductions of process_parallel
//   K2 k_unpack            24 + bc_len + umi_len B     2-bit -> ASCII + validation + K1's reductions
//   K3 k_pack              bc_len + umi_len + 24 B     ASCII -> Record + validation
//   generators             synthetic inputs shared with the oracle
//
// All three are HBM-bound streamers: every global access is a 128/256-bit fully coalesced
// vector access of read-once / write-once data (loads: L1 no-allocate, L2 evict-first on the
// 256-bit ones; stores: plain write-back, measured fastest), the 24-byte
// AoS records are transposed to per-lane records through shared memory (K2) or consumed in
// place with a lane-static field rotation (K1), decode/encode is branch-free SWAR with the
// ACGT table held in a register (PRMT).  No tensor cores: nothing here is a contraction.
#include <algorithm>
#include <cstdlib>

#include "ctx.h"
#include "kernels.cuh"

namespace ibu {

std::atomic<uint64_t> g_launches{0};

// ============================================================================ K1
// One warp consumes tiles of 128 records = 96 x 32 B.  Lane l loads 32-byte units
// l, l+32, l+64 of the tile (three LDG.256, perfectly coalesced).  u64 number e of the
// tile is field e % 3 of a record; for lane l, unit k, element j: e = 4(l+32k)+j, so the
// field is (l + 2k + j) % 3 = (c + g) % 3 with c = l % 3 (lane constant) and
// g = (2k + j) % 3 (compile-time constant).  Elements are therefore accumulated into three
// groups g and rotated to fields once, after the loop.
struct ReduceAcc {
    uint64_t a0 = 0, a1 = 0, a2 = 0, x = 0;
    uint32_t b0 = 0, b1 = 0, b2 = 0;
};

#define IBU_ACC(G, V, SLOT)                           \
    do {                                              \
        acc.a##G += (V);                              \
        acc.x ^= (V);                                 \
        uint32_t bit__ = ((V) & m##G) != 0ull;        \
        acc.b##G += bit__;                            \
        bm |= bit__ << (SLOT);                        \
    } while (0)

// ---- result blocks ---------------------------------------------------------------------------
// A warp adds its partial result to one of 256 zeroed result blocks (one RED instruction whose
// lanes 1-7 carry the words of ibu_reduce_result_t, a second one for the xor word) and exits: no
// block barrier, no CTA-level tail.  k_fold_result folds the blocks into the caller's result and
// re-zeroes them.  Measured on unpack bc16/umi12, 10^8 records (tools/k2lab.cu): CTA reduction +
// one RED per CTA into the result itself 0.797 ms, this 0.757 ms, no result at all 0.754 ms — and
// 195 k CTAs RED-ing into ONE 64-byte block take 1.30 ms (same-address atomics serialise at
// ~3 ns each), which is why the blocks are spread.
__device__ __forceinline__ void red_spread(unsigned long long *blocks, uint32_t slot, uint32_t lane,
                                           uint64_t s_bc, uint64_t s_umi, uint64_t s_idx, uint64_t x_all,
                                           uint64_t n_bb, uint64_t n_bu, uint64_t n_br) {
    const uint64_t v = lane == 1 ? s_bc : lane == 2 ? s_umi : lane == 3 ? s_idx : lane == 4 ? x_all
                       : lane == 5 ? n_bb : lane == 6 ? n_bu : n_br;
    unsigned long long *blk = blocks + 8 * (slot & (kResultBlocks - 1));
    if (lane >= 1 && lane < 8 && v) {
        if (lane == 4) atomicXor(blk + 4, (unsigned long long)v);
        else atomicAdd(blk + lane, (unsigned long long)v);
    }
}

// <<<1, 256>>>: thread t owns result block t
__global__ void __launch_bounds__(kResultBlocks)
k_fold_result(unsigned long long *__restrict__ blocks, ibu_reduce_result_t *__restrict__ res, uint64_t n_records) {
    __shared__ uint64_t red[kResultBlocks / 32][8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    ulonglong2 *mine = reinterpret_cast<ulonglong2 *>(blocks + 8 * threadIdx.x);
    uint64_t w[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const ulonglong2 v = mine[k];
        w[2 * k] = v.x; w[2 * k + 1] = v.y;
        mine[k] = make_ulonglong2(0ull, 0ull);
    }
#pragma unroll
    for (int k = 1; k < 8; k++) {
        const uint64_t r = k == 4 ? warp_xor64(w[k]) : warp_sum64(w[k]);
        if (lane == 0) red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        uint64_t v = n_records;  // word 0
        if (threadIdx.x > 0) {
            v = 0;
            for (int q = 0; q < kResultBlocks / 32; q++) {
                if (threadIdx.x == 4) v ^= red[q][4]; else v += red[q][threadIdx.x];
            }
        }
        reinterpret_cast<uint64_t *>(res)[threadIdx.x] = v;
    }
}

// `head` records in front of the 32-byte aligned body (a record pointer is only 8-byte aligned
// in general, e.g. a slice of a larger array) are handled with the ragged tail.
// No grid-stride loop: CTA c owns the 8 TPW consecutive tiles from 8 TPW c, warp w takes tiles
// w, w + 8, ... of them.  TPW = 4 amortises the CTA's reduction (10^8 records: TPW 1 0.41 ms,
// 2 0.346, 4 and 8 0.333; a register prefetch of the next tile costs occupancy: 0.36-0.49).
template <int TPW>
__global__ void __launch_bounds__(kBlockThreads)
k_validate_reduce(const uint8_t *__restrict__ recs_all, uint64_t n_all, uint32_t head, uint64_t bc_hi,
                  uint64_t umi_hi, unsigned long long *__restrict__ blocks) {
    const uint8_t *recs = recs_all + (uint64_t)head * 24;
    const uint64_t n = n_all - head;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t c = lane % 3u;
    // mask of group g = mask of field (c + g) % 3; field 2 (index) is never invalid
    const uint64_t m0 = c == 0 ? bc_hi : c == 1 ? umi_hi : 0ull;
    const uint64_t m1 = c == 0 ? umi_hi : c == 1 ? 0ull : bc_hi;
    const uint64_t m2 = c == 0 ? 0ull : c == 1 ? bc_hi : umi_hi;

    ReduceAcc acc;
    uint32_t both = 0;  // records whose barcode AND umi are invalid
    const uint64_t n_tiles = n / kTileRecords;
    uint64_t t = (uint64_t)blockIdx.x * (kWarpsPerBlock * TPW) + warp;
#pragma unroll 1
    for (int j = 0; j < TPW && t < n_tiles; j++, t += kWarpsPerBlock) {
        const u64x4 *base = reinterpret_cast<const u64x4 *>(recs) + t * kTileU8 + lane;
        const u64x4 v0 = ldg_stream256(base), v1 = ldg_stream256(base + 32), v2 = ldg_stream256(base + 64);
        uint32_t bm = 0;  // bit 4k+j: element (k, j) failed its mask
        IBU_ACC(0, v0.x, 0); IBU_ACC(1, v0.y, 1); IBU_ACC(2, v0.z, 2);  IBU_ACC(0, v0.w, 3);
        IBU_ACC(2, v1.x, 4); IBU_ACC(0, v1.y, 5); IBU_ACC(1, v1.z, 6);  IBU_ACC(2, v1.w, 7);
        IBU_ACC(1, v2.x, 8); IBU_ACC(2, v2.y, 9); IBU_ACC(0, v2.z, 10); IBU_ACC(1, v2.w, 11);
        if (__any_sync(0xffffffffu, bm != 0)) {
            // rare path: pair each invalid barcode with the word that follows it (its umi).
            // Two adjacent failing words are always (barcode, umi): the index never fails.
            // The word after element (k,3) of lane l is element (k,0) of lane l+1, or
            // element (k+1,0) of lane 0 when l == 31.
            uint32_t firsts = (bm & 1u) | ((bm >> 3) & 2u) | ((bm >> 6) & 4u);
            uint32_t nxt = __shfl_sync(0xffffffffu, firsts, (lane + 1) & 31u);
            if (lane == 31) nxt >>= 1;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                uint32_t nib = (bm >> (4 * k)) & 0xFu;
                uint32_t ext = nib | (((nxt >> k) & 1u) << 4);
                both += __popc(nib & (ext >> 1));
            }
        }
    }
    // rotate groups to fields: field f lives in group (f - c) mod 3
    uint64_t s_bc = c == 0 ? acc.a0 : c == 1 ? acc.a2 : acc.a1;
    uint64_t s_umi = c == 0 ? acc.a1 : c == 1 ? acc.a0 : acc.a2;
    uint64_t s_idx = c == 0 ? acc.a2 : c == 1 ? acc.a1 : acc.a0;
    uint64_t n_bb = c == 0 ? acc.b0 : c == 1 ? acc.b2 : acc.b1;
    uint64_t n_bu = c == 0 ? acc.b1 : c == 1 ? acc.b0 : acc.b2;
    uint64_t x = acc.x, n_both = both;

    if (blockIdx.x == 0 && warp == 0) {  // head + ragged tail (< 4 + 128 records), one record per lane per step
        const uint64_t *r64 = reinterpret_cast<const uint64_t *>(recs_all);
        const uint64_t tail0 = head + n_tiles * kTileRecords;
        for (uint64_t k = lane; k < head + (n_all - tail0); k += 32) {
            const uint64_t r = k < head ? k : tail0 + (k - head);
            uint64_t b = ldg_stream64(r64 + 3 * r), u = ldg_stream64(r64 + 3 * r + 1),
                     i = ldg_stream64(r64 + 3 * r + 2);
            s_bc += b; s_umi += u; s_idx += i; x ^= b ^ u ^ i;
            bool bb = (b & bc_hi) != 0, bu = (u & umi_hi) != 0;
            n_bb += bb; n_bu += bu; n_both += (bb && bu);
        }
    }

    s_bc = warp_sum64(s_bc); s_umi = warp_sum64(s_umi); s_idx = warp_sum64(s_idx);
    x = warp_xor64(x);
    n_bb = warp_sum64(n_bb); n_bu = warp_sum64(n_bu); n_both = warp_sum64(n_both);
    red_spread(blocks, blockIdx.x * kWarpsPerBlock + warp, lane, s_bc, s_umi, s_idx, x, n_bb, n_bu,
               n_bb + n_bu - n_both);
}

// ============================================================================ K2
struct UnpackArgs {
    const uint8_t *recs;
    uint64_t n;
    uint8_t *bc_out, *umi_out, *flags;
    uint64_t bc_hi, umi_hi;
    unsigned long long *res_blocks;  // 256 spread result blocks (SUMS) or nullptr
    uint32_t bc_len, umi_len;
    uint32_t warp_smem_bytes;  // per-warp shared memory: input tile + staged outputs
    uint32_t bc_stage_off, umi_stage_off;
};

template <int L>
struct LenOf {
    static __device__ __forceinline__ uint32_t get(uint32_t) { return L; }
};
template <>
struct LenOf<0> {
    static __device__ __forceinline__ uint32_t get(uint32_t rt) { return rt; }
};

// Stage one decoded row of compile-time length L (any 1..32) at byte offset r * L.
//   L % 4 == 0: word stores;  L even: halfword stores;  L odd: rows start at any byte, so the row is
//   shifted into place with funnel shifts and OR-ed into the (zeroed) stage word by word with
//   shared-memory atomics - neighbouring rows share boundary words.
template <int L>
__device__ __forceinline__ void stage_row(uint64_t w, uint8_t *stage, uint32_t r) {
    constexpr int NG = (L + 3) / 4;
    uint32_t asc[8];
    decode_word<NG>(w, asc);
    if constexpr (L % 4 == 0) {
        uint32_t *s32 = reinterpret_cast<uint32_t *>(stage) + r * (L / 4);
#pragma unroll
        for (int g = 0; g < NG; g++) s32[g] = asc[g];
    } else if constexpr (L % 2 == 0) {
        uint16_t *s16 = reinterpret_cast<uint16_t *>(stage) + r * (L / 2);
#pragma unroll
        for (int h = 0; h < L / 2; h++) s16[h] = (uint16_t)(asc[h / 2] >> (16 * (h & 1)));
    } else {
        const uint32_t o = r * L, sh = (o & 3u) * 8u;
        uint32_t *s32 = reinterpret_cast<uint32_t *>(stage) + (o >> 2);
        asc[NG - 1] &= (1u << (8 * (L & 3))) - 1u;  // bytes past the row in its last word
        uint32_t prev = 0;
#pragma unroll
        for (int g = 0; g <= NG; g++) {
            const uint32_t cur = g < NG ? asc[g < NG ? g : 0] : 0u;
            if (4u * g < (o & 3u) + L) atomicOr(s32 + g, __funnelshift_l(prev, cur, sh));
            prev = cur;
        }
    }
}

// Emit one decoded row.  L = 32 / 16: straight from registers with a 256 / 128-bit store
// (a warp covers 1024 / 512 contiguous bytes).  Otherwise the row is staged in shared memory
// and the whole tile (128 x len bytes, 16-byte aligned in the output) is copied out afterwards.
template <int L>
__device__ __forceinline__ void emit_row(uint64_t w, uint32_t len, uint8_t *gout, uint64_t rec,
                                         uint8_t *stage, uint32_t r) {
    uint32_t asc[8];
    if (L == 32) {
        decode_word<8>(w, asc);
        stg_stream256(gout + rec * 32, make_uint4(asc[0], asc[1], asc[2], asc[3]),
                      make_uint4(asc[4], asc[5], asc[6], asc[7]));
    } else if (L == 16) {
        decode_word<4>(w, asc);
        stg_stream(reinterpret_cast<uint4 *>(gout) + rec, make_uint4(asc[0], asc[1], asc[2], asc[3]));
    } else if (L > 0) {  // compile-time length, staged
        stage_row<L>(w, stage, r);
    }
    // L == 0 (runtime length): staged per tile by stage_tile_rt, not per row
}

// All four rows of a lane for one runtime-length output.  The length is warp-uniform, so one
// switch per tile selects a fully specialised body: generic-length shapes then execute about the
// same instructions as the compile-time shapes instead of ~3x as many predicated ones.
template <int L>
__device__ __forceinline__ void stage_tile(const uint64_t (&w)[4], uint8_t *stage, uint32_t lane) {
    if constexpr (L % 2 == 1) {  // OR-ed rows need a zeroed stage
        uint4 *s4 = reinterpret_cast<uint4 *>(stage);
        for (uint32_t i = lane; i < 8 * L; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 4; q++) stage_row<L>(w[q], stage, lane + 32 * q);
}
__device__ __forceinline__ void stage_tile_rt(uint32_t len, const uint64_t (&w)[4], uint8_t *stage, uint32_t lane) {
    switch (len) {
#define IBU_CASE(L) case L: stage_tile<L>(w, stage, lane); break;
        IBU_CASE(1) IBU_CASE(2) IBU_CASE(3) IBU_CASE(4) IBU_CASE(5) IBU_CASE(6) IBU_CASE(7) IBU_CASE(8)
        IBU_CASE(9) IBU_CASE(10) IBU_CASE(11) IBU_CASE(12) IBU_CASE(13) IBU_CASE(14) IBU_CASE(15) IBU_CASE(16)
        IBU_CASE(17) IBU_CASE(18) IBU_CASE(19) IBU_CASE(20) IBU_CASE(21) IBU_CASE(22) IBU_CASE(23) IBU_CASE(24)
        IBU_CASE(25) IBU_CASE(26) IBU_CASE(27) IBU_CASE(28) IBU_CASE(29) IBU_CASE(30) IBU_CASE(31) IBU_CASE(32)
#undef IBU_CASE
    }
}

template <int L>
__device__ __forceinline__ void copy_out_stage(uint32_t len, uint8_t *gout, uint64_t tile,
                                               const uint8_t *stage, uint32_t lane) {
    if (L == 32 || L == 16) return;  // rows were stored directly
    const uint32_t n16 = 8 * len;    // 128 rows x len bytes / 16
    uint4 *dst = reinterpret_cast<uint4 *>(gout) + tile * n16;
    const uint4 *src = reinterpret_cast<const uint4 *>(stage);
    if constexpr (L > 0) {  // compile-time length: all shared loads first, then the stores
        constexpr int kIters = (8 * L + 31) / 32;
        uint4 v[kIters];
#pragma unroll
        for (int k = 0; k < kIters; k++)
            if (lane + 32 * k < 8 * L) v[k] = src[lane + 32 * k];
#pragma unroll
        for (int k = 0; k < kIters; k++)
            if (lane + 32 * k < 8 * L) stg_stream(dst + lane + 32 * k, v[k]);
    } else {
        for (uint32_t i = lane; i < n16; i += 32) stg_stream(dst + i, src[i]);
    }
}

// One 128-record tile per warp and no grid-stride loop: the grid is ceil(tiles / 8) CTAs.
// Measured on B200 (tools/k2lab.cu, 10^8 records): a persistent grid marching through memory
// in lock step moves this byte mix at 5.95-6.05 TB/s whatever the kernel does (a decode-free
// traffic kernel of the same shape is no faster), the same tiles handed out by the block
// scheduler at 6.8 TB/s (traffic-only: 6.8); every extra tile per warp costs bandwidth.
template <int BC, int UMI, bool SUMS>
__global__ void __launch_bounds__(kBlockThreads) k_unpack(const UnpackArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + warp * a.warp_smem_bytes;
    uint4 *in4 = reinterpret_cast<uint4 *>(wsm);
    const uint64_t *in64 = reinterpret_cast<const uint64_t *>(wsm);
    uint8_t *bc_stage = wsm + a.bc_stage_off, *umi_stage = wsm + a.umi_stage_off;
    const uint32_t bc_len = LenOf<BC>::get(a.bc_len), umi_len = LenOf<UMI>::get(a.umi_len);

    uint32_t n_bb = 0, n_bu = 0, n_br = 0;
    uint64_t s_bc = 0, s_umi = 0, s_idx = 0, x_all = 0;
    const uint64_t n_tiles = a.n / kTileRecords;
    const uint64_t t = (uint64_t)blockIdx.x * kWarpsPerBlock + warp;  // this warp's tile

    if (t < n_tiles) {
        const uint4 *g4 = reinterpret_cast<const uint4 *>(a.recs) + t * kTileU4;
#pragma unroll
        for (int k = 0; k < 6; k++) in4[lane + 32 * k] = ldg_stream(g4 + lane + 32 * k);
        __syncwarp();
        // Kernels with a runtime-length output take their four records into registers first.  When
        // BOTH outputs are staged the tile's shared memory then becomes the stage (kReuse: 55 ->
        // 31 KB per CTA at bc20/umi10), which is what lets them run in the large-L1 configurations.
        constexpr bool kRegs = (BC == 0 || UMI == 0);
        constexpr bool kReuse = (BC == 0) && !(UMI == 16 || UMI == 32);
        uint64_t rb[4], ru[4], ri[4];
        if constexpr (kRegs) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t r = lane + 32 * q;
                rb[q] = in64[3 * r]; ru[q] = in64[3 * r + 1];
                if (SUMS) ri[q] = in64[3 * r + 2];
            }
            if constexpr (kReuse) __syncwarp();  // every lane holds its records: the tile may be overwritten
            if constexpr (BC == 0) stage_tile_rt(bc_len, rb, bc_stage, lane);
            if constexpr (UMI == 0) stage_tile_rt(umi_len, ru, umi_stage, lane);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t r = lane + 32 * q;  // record of the tile handled by this lane
            // 64-bit shared loads at a 24-byte stride: conflict-free per half-warp
            uint64_t bc, umi;
            if constexpr (kRegs) { bc = rb[q]; umi = ru[q]; } else { bc = in64[3 * r]; umi = in64[3 * r + 1]; }
            const uint64_t rec = t * kTileRecords + r;
            emit_row<BC>(bc, bc_len, a.bc_out, rec, bc_stage, r);
            emit_row<UMI>(umi, umi_len, a.umi_out, rec, umi_stage, r);
            const uint32_t bb = (bc & a.bc_hi) != 0ull, bu = (umi & a.umi_hi) != 0ull;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            if (a.flags) a.flags[rec] = (uint8_t)(bb | (bu << 1));
            if (SUMS) {  // the reference processors' sums / checksum ride along (the tile is on chip)
                uint64_t idx;
                if constexpr (kRegs) idx = ri[q]; else idx = in64[3 * r + 2];
                s_bc += bc; s_umi += umi; s_idx += idx; x_all ^= bc ^ umi ^ idx;
            }
        }
        __syncwarp();
        copy_out_stage<BC>(bc_len, a.bc_out, t, bc_stage, lane);
        copy_out_stage<UMI>(umi_len, a.umi_out, t, umi_stage, lane);
    } else if (t == n_tiles) {  // ragged tail (< 128 records): plain per-record code
        const uint64_t *r64 = reinterpret_cast<const uint64_t *>(a.recs);
        for (uint64_t rec = n_tiles * kTileRecords + lane; rec < a.n; rec += 32) {
            const uint64_t bc = ldg_stream64(r64 + 3 * rec), umi = ldg_stream64(r64 + 3 * rec + 1);
            for (uint32_t i = 0; i < bc_len; i++)
                a.bc_out[rec * bc_len + i] = (uint8_t)(kAcgt >> (8 * ((bc >> (2 * i)) & 3u)));
            for (uint32_t i = 0; i < umi_len; i++)
                a.umi_out[rec * umi_len + i] = (uint8_t)(kAcgt >> (8 * ((umi >> (2 * i)) & 3u)));
            const uint32_t bb = (bc & a.bc_hi) != 0ull, bu = (umi & a.umi_hi) != 0ull;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            if (a.flags) a.flags[rec] = (uint8_t)(bb | (bu << 1));
            if (SUMS) {
                const uint64_t idx = ldg_stream64(r64 + 3 * rec + 2);
                s_bc += bc; s_umi += umi; s_idx += idx; x_all ^= bc ^ umi ^ idx;
            }
        }
    }

    if (SUMS) {  // a result was asked for: warp-level reduction, then straight into a spread block
        const uint64_t c_bb = __reduce_add_sync(0xffffffffu, n_bb), c_bu = __reduce_add_sync(0xffffffffu, n_bu),
                       c_br = __reduce_add_sync(0xffffffffu, n_br);
        s_bc = warp_sum64(s_bc); s_umi = warp_sum64(s_umi); s_idx = warp_sum64(s_idx);
        x_all = warp_xor64(x_all);
        red_spread(a.res_blocks, blockIdx.x * kWarpsPerBlock + warp, lane, s_bc, s_umi, s_idx, x_all, c_bb, c_bu, c_br);
    }
}

// ============================================================================ K3
struct PackArgs {
    const uint8_t *bc_in, *umi_in;
    const uint64_t *index;
    uint64_t index_base, n;
    uint8_t *recs_out, *flags;
    unsigned long long *res_blocks;  // 256 spread result blocks or nullptr
    uint32_t bc_len, umi_len;
    uint32_t warp_smem_bytes, bc_stage_off, umi_stage_off;
};

// One ASCII row as two 16-byte halves.  L = 32 / 16: the lane loads its own row with one
// 256 / 128-bit load (a warp covers 1024 / 512 contiguous bytes) and the next tile's rows are
// prefetched into registers.  L = 0: runtime length, rows reach the lane through shared memory.
// Q = rows per lane (a warp tile is 32 Q rows): 2 for 32-byte rows, 4 when every row is <= 16
// bytes, so that a lane keeps >= 64 bytes per input in flight either way.
template <int L, int Q>
struct RowRegs;
template <int Q>
struct RowRegs<32, Q> {
    u64x4 v[Q];
    __device__ __forceinline__ void load(const uint8_t *g, uint64_t tile, uint32_t lane, uint32_t) {
#pragma unroll
        for (int q = 0; q < Q; q++) v[q] = ldg_stream256(g + (tile * (32 * Q) + lane + 32 * q) * 32);
    }
    __device__ __forceinline__ void park(uint8_t *, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void get(int q, uint4 &lo, uint4 &hi) const {
        lo = make_uint4((uint32_t)v[q].x, (uint32_t)(v[q].x >> 32), (uint32_t)v[q].y, (uint32_t)(v[q].y >> 32));
        hi = make_uint4((uint32_t)v[q].z, (uint32_t)(v[q].z >> 32), (uint32_t)v[q].w, (uint32_t)(v[q].w >> 32));
    }
};
template <int Q>
struct RowRegs<16, Q> {
    uint4 v[Q];
    __device__ __forceinline__ void load(const uint8_t *g, uint64_t tile, uint32_t lane, uint32_t) {
#pragma unroll
        for (int q = 0; q < Q; q++)
            v[q] = ldg_stream(reinterpret_cast<const uint4 *>(g) + tile * (32 * Q) + lane + 32 * q);
    }
    __device__ __forceinline__ void park(uint8_t *, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void get(int q, uint4 &lo, uint4 &hi) const {
        lo = v[q];
        hi = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
    }
};
// Rows of any other length travel through shared memory: the 32 Q-row tile (32 Q x len bytes,
// 16-byte aligned in the input, at most 2 Q 16-byte pieces per lane) is prefetched into
// registers one tile ahead like the direct rows, then parked in the warp's stage.
template <int Q>
struct RowRegs<0, Q> {
    uint4 v[2 * Q];
    __device__ __forceinline__ void load(const uint8_t *g, uint64_t tile, uint32_t lane, uint32_t len) {
        const uint32_t n16 = 2 * Q * len;
        const uint4 *src = reinterpret_cast<const uint4 *>(g) + tile * n16;
#pragma unroll
        for (int k = 0; k < 2 * Q; k++)
            if (lane + 32 * k < n16) v[k] = ldg_stream(src + lane + 32 * k);
    }
    __device__ __forceinline__ void park(uint8_t *stage, uint32_t lane, uint32_t len) const {
        const uint32_t n16 = 2 * Q * len;
        uint4 *dst = reinterpret_cast<uint4 *>(stage);
#pragma unroll
        for (int k = 0; k < 2 * Q; k++)
            if (lane + 32 * k < n16) dst[lane + 32 * k] = v[k];
    }
};
template <int Q>
struct RowRegs<12, Q> : RowRegs<0, Q> {};  // 12- and 10-byte rows are staged too, with a compile-time length
template <int Q>
struct RowRegs<10, Q> : RowRegs<0, Q> {};

// Gather the staged row r of compile-time length L (any 1..32) into two 16-byte halves, padded
// with 'A' (code 0, valid).  L % 4 == 0: word loads (odd word strides are conflict-free);
// L even: halfword loads; L odd: rows start at any byte - aligned words + funnel shift (the stage
// has 16 spare bytes behind the tile for the one-word over-read).
template <int L>
__device__ __forceinline__ void load_row(const uint8_t *stage, uint32_t r, uint4 &lo, uint4 &hi) {
    constexpr int NG = (L + 3) / 4;
    uint32_t w[8];
#pragma unroll
    for (int g = 0; g < 8; g++) w[g] = 0x41414141u;
    if constexpr (L % 4 == 0) {
        const uint32_t *r32 = reinterpret_cast<const uint32_t *>(stage) + r * (L / 4);
#pragma unroll
        for (int g = 0; g < NG; g++) w[g] = r32[g];
    } else if constexpr (L % 2 == 0) {
        const uint16_t *r16 = reinterpret_cast<const uint16_t *>(stage) + r * (L / 2);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint32_t lo16 = r16[2 * g];
            const uint32_t hi16 = (4 * g + 2 < L) ? (uint32_t)r16[2 * g + 1] : 0x4141u;
            w[g] = lo16 | (hi16 << 16);
        }
    } else {
        const uint32_t o = r * L, sh = (o & 3u) * 8u;
        const uint32_t *r32 = reinterpret_cast<const uint32_t *>(stage) + (o >> 2);
        uint32_t cur = r32[0];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint32_t nxt = r32[g + 1];
            uint32_t v = __funnelshift_r(cur, nxt, sh);
            if (g == NG - 1) {
                constexpr uint32_t rem = L & 3;  // 1 or 3 bytes of the row in its last word
                v = (v & ((1u << (8 * rem)) - 1u)) | (0x41414141u << (8 * rem));
            }
            w[g] = v;
            cur = nxt;
        }
    }
    lo = make_uint4(w[0], w[1], w[2], w[3]);
    hi = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int L>
__device__ __forceinline__ uint64_t pack_row(uint4 lo, uint4 hi, uint32_t &badflag) {
    uint32_t bad = 0;
    uint64_t w = pack16(lo, bad);
    if constexpr (L > 16) w |= (uint64_t)pack16(hi, bad) << 32;
    badflag = bad != 0;
    return w;
}

// All Q rows of a lane for one runtime-length input: packed words and a bad-row mask.  The length
// is warp-uniform: one switch per tile selects a fully specialised body (see stage_tile_rt).
template <int L, int Q>
__device__ __forceinline__ void pack_tile(const uint8_t *stage, uint32_t lane, uint64_t (&w)[Q], uint32_t &badmask) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
        uint4 lo, hi;
        uint32_t bad;
        load_row<L>(stage, lane + 32 * q, lo, hi);
        w[q] = pack_row<L>(lo, hi, bad);
        badmask |= bad << q;
    }
}
template <int Q>
__device__ __forceinline__ void pack_tile_rt(uint32_t len, const uint8_t *stage, uint32_t lane, uint64_t (&w)[Q],
                                          uint32_t &badmask) {
    switch (len) {
#define IBU_CASE(L) case L: pack_tile<L, Q>(stage, lane, w, badmask); break;
        IBU_CASE(1) IBU_CASE(2) IBU_CASE(3) IBU_CASE(4) IBU_CASE(5) IBU_CASE(6) IBU_CASE(7) IBU_CASE(8)
        IBU_CASE(9) IBU_CASE(10) IBU_CASE(11) IBU_CASE(12) IBU_CASE(13) IBU_CASE(14) IBU_CASE(15) IBU_CASE(16)
        IBU_CASE(17) IBU_CASE(18) IBU_CASE(19) IBU_CASE(20) IBU_CASE(21) IBU_CASE(22) IBU_CASE(23) IBU_CASE(24)
        IBU_CASE(25) IBU_CASE(26) IBU_CASE(27) IBU_CASE(28) IBU_CASE(29) IBU_CASE(30) IBU_CASE(31) IBU_CASE(32)
#undef IBU_CASE
    }
}

// One tile of 32 Q rows per warp, no grid-stride loop (see k_unpack for the measurement).
template <int BC, int UMI, int Q, int MINB = 1>
__global__ void __launch_bounds__(kBlockThreads, MINB) k_pack(const PackArgs a) {
    constexpr int kRows = 32 * Q;  // rows per warp tile
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + warp * a.warp_smem_bytes;
    uint64_t *out64 = reinterpret_cast<uint64_t *>(wsm);  // kRows records x 24 B staged output
    const uint4 *out4 = reinterpret_cast<const uint4 *>(wsm);
    uint8_t *bc_stage = wsm + a.bc_stage_off, *umi_stage = wsm + a.umi_stage_off;
    const uint32_t bc_len = LenOf<BC>::get(a.bc_len), umi_len = LenOf<UMI>::get(a.umi_len);

    uint32_t n_bb = 0, n_bu = 0, n_br = 0;
    const uint64_t n_tiles = a.n / kRows;
    const uint64_t t = (uint64_t)blockIdx.x * kWarpsPerBlock + warp;  // this warp's tile
    if (t < n_tiles) {
        constexpr bool kStageBc = (BC != 16 && BC != 32), kStageUmi = (UMI != 16 && UMI != 32);
        RowRegs<BC, Q> bc_rows;
        RowRegs<UMI, Q> umi_rows;
        bc_rows.load(a.bc_in, t, lane, bc_len);
        umi_rows.load(a.umi_in, t, lane, umi_len);
        bc_rows.park(bc_stage, lane, bc_len);
        umi_rows.park(umi_stage, lane, umi_len);
        if (kStageBc || kStageUmi) __syncwarp();
        // Packed words of the lane's Q rows.  When both inputs are staged (kReuse) every row is
        // packed before the first record is written, so the record tile can overlay the input
        // stages (55 -> 31 KB of shared memory per CTA at bc20/umi10: the large-L1 configurations).
        constexpr bool kReuse = kStageBc && kStageUmi;
        uint64_t bcw[Q], umw[Q];
        uint32_t bcbad = 0, umbad = 0;
        if constexpr (BC == 0) pack_tile_rt<Q>(bc_len, bc_stage, lane, bcw, bcbad);
        if constexpr (UMI == 0) pack_tile_rt<Q>(umi_len, umi_stage, lane, umw, umbad);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const uint32_t r = lane + 32 * q;
            uint32_t bad;
            if constexpr (BC != 0) {
                uint4 lo, hi;
                if constexpr (!kStageBc) bc_rows.get(q, lo, hi); else load_row<BC>(bc_stage, r, lo, hi);
                bcw[q] = pack_row<BC>(lo, hi, bad);
                bcbad |= bad << q;
            }
            if constexpr (UMI != 0) {
                uint4 lo, hi;
                if constexpr (!kStageUmi) umi_rows.get(q, lo, hi); else load_row<UMI>(umi_stage, r, lo, hi);
                umw[q] = pack_row<UMI>(lo, hi, bad);
                umbad |= bad << q;
            }
        }
        if constexpr (kReuse) __syncwarp();  // every staged row has been read
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const uint32_t r = lane + 32 * q;
            const uint64_t row = t * kRows + r;
            const uint32_t bb = (bcbad >> q) & 1u, bu = (umbad >> q) & 1u;
            const uint64_t idx = a.index ? ldg_stream64(a.index + row) : a.index_base + row;
            out64[3 * r] = bcw[q]; out64[3 * r + 1] = umw[q]; out64[3 * r + 2] = idx;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            if (a.flags) a.flags[row] = (uint8_t)(bb | (bu << 1));
        }
        __syncwarp();
        uint4 *dst = reinterpret_cast<uint4 *>(a.recs_out) + t * (kRows * 24 / 16);
#pragma unroll
        for (int k = 0; k < 3 * Q / 2; k++) stg_stream(dst + lane + 32 * k, out4[lane + 32 * k]);
    } else if (t == n_tiles) {  // ragged tail (< kRows rows): plain per-row code
        uint64_t *o64 = reinterpret_cast<uint64_t *>(a.recs_out);
        for (uint64_t row = n_tiles * kRows + lane; row < a.n; row += 32) {
            uint64_t w[2];
            uint32_t bad[2];
            for (int s = 0; s < 2; s++) {
                const uint8_t *p = s ? a.umi_in + row * umi_len : a.bc_in + row * bc_len;
                const uint32_t len = s ? umi_len : bc_len;
                uint64_t acc = 0;
                uint32_t b = 0;
                for (uint32_t i = 0; i < len; i++) {
                    const uint32_t ch = p[i], c1 = (ch >> 1) & 3u, up = ch & 0xDFu;
                    acc |= (uint64_t)(c1 ^ (c1 >> 1)) << (2 * i);
                    b |= !(up == 0x41u || up == 0x43u || up == 0x47u || up == 0x54u);
                }
                w[s] = acc;
                bad[s] = b;
            }
            o64[3 * row] = w[0]; o64[3 * row + 1] = w[1];
            o64[3 * row + 2] = a.index ? a.index[row] : a.index_base + row;
            n_bb += bad[0]; n_bu += bad[1]; n_br += (bad[0] | bad[1]);
            if (a.flags) a.flags[row] = (uint8_t)(bad[0] | (bad[1] << 1));
        }
    }

    if (a.res_blocks) {  // counters: one RED per warp that saw a bad row, into a spread result block
        const uint64_t c_bb = __reduce_add_sync(0xffffffffu, n_bb), c_bu = __reduce_add_sync(0xffffffffu, n_bu),
                       c_br = __reduce_add_sync(0xffffffffu, n_br);
        if (c_br) red_spread(a.res_blocks, blockIdx.x * kWarpsPerBlock + warp, lane, 0, 0, 0, 0, c_bb, c_bu, c_br);
    }
}

// Runtime lengths on both inputs (no compile-time shape): the packing itself is ~1000 warp
// instructions per 128-row tile, as long as the tile's loads take to arrive, and with ONE tile per
// warp nothing is in flight while a warp packs (ncu: issue slots 47 % busy, 7.5 warps per issue
// waiting on the long scoreboard).  Here a warp walks kTpw consecutive tiles and keeps the next
// tile's rows arriving in a second stage (cp.async: global -> shared without registers) while it
// packs the current one; the record tile overlays the stage it was packed from.
constexpr int kPackRtTpw = 4;
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <int Q, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) k_pack_rt(const PackArgs a) {
    constexpr int kRows = 32 * Q;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + warp * (2 * a.warp_smem_bytes);  // two stages per warp
    const uint32_t bc_len = a.bc_len, umi_len = a.umi_len;
    const uint32_t n16_bc = 2 * Q * bc_len, n16_umi = 2 * Q * umi_len;  // 16-byte pieces per tile
    uint32_t n_bb = 0, n_bu = 0, n_br = 0;
    const uint64_t n_tiles = a.n / kRows;
    const uint64_t t0 = ((uint64_t)blockIdx.x * kWarpsPerBlock + warp) * kPackRtTpw;
    auto prefetch = [&](uint64_t t, uint32_t buf) {
        if (t < n_tiles) {
            uint8_t *st = wsm + buf * a.warp_smem_bytes;
            const uint4 *sb = reinterpret_cast<const uint4 *>(a.bc_in) + t * n16_bc;
            const uint4 *su = reinterpret_cast<const uint4 *>(a.umi_in) + t * n16_umi;
#pragma unroll
            for (int k = 0; k < 2 * Q; k++) {
                const uint32_t i = lane + 32 * k;
                if (i < n16_bc) cp_async16(st + a.bc_stage_off + 16 * i, sb + i);
                if (i < n16_umi) cp_async16(st + a.umi_stage_off + 16 * i, su + i);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(t0, 0);
#pragma unroll 1
    for (int it = 0; it < kPackRtTpw; it++) {
        const uint64_t t = t0 + it;
        const uint32_t buf = it & 1;
        prefetch(t + 1 < t0 + kPackRtTpw ? t + 1 : n_tiles, buf ^ 1u);  // (an empty group past the last tile)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        if (t >= n_tiles) break;
        uint8_t *st = wsm + buf * a.warp_smem_bytes;
        uint64_t bcw[Q], umw[Q];
        uint32_t bcbad = 0, umbad = 0;
        pack_tile_rt<Q>(bc_len, st + a.bc_stage_off, lane, bcw, bcbad);
        pack_tile_rt<Q>(umi_len, st + a.umi_stage_off, lane, umw, umbad);
        __syncwarp();  // every staged row has been read: the record tile may overlay the stage
        uint64_t *out64 = reinterpret_cast<uint64_t *>(st);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const uint32_t r = lane + 32 * q;
            const uint64_t row = t * kRows + r;
            const uint32_t bb = (bcbad >> q) & 1u, bu = (umbad >> q) & 1u;
            const uint64_t idx = a.index ? ldg_stream64(a.index + row) : a.index_base + row;
            out64[3 * r] = bcw[q]; out64[3 * r + 1] = umw[q]; out64[3 * r + 2] = idx;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            if (a.flags) a.flags[row] = (uint8_t)(bb | (bu << 1));
        }
        __syncwarp();
        const uint4 *out4 = reinterpret_cast<const uint4 *>(st);
        uint4 *dst = reinterpret_cast<uint4 *>(a.recs_out) + t * (kRows * 24 / 16);
#pragma unroll
        for (int k = 0; k < 3 * Q / 2; k++) stg_stream(dst + lane + 32 * k, out4[lane + 32 * k]);
        __syncwarp();  // the stores have read the stage before the next prefetch lands in it
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ragged tail (< kRows rows): the warp that would own tile n_tiles
    if (t0 <= n_tiles && n_tiles < t0 + kPackRtTpw) {
        uint64_t *o64 = reinterpret_cast<uint64_t *>(a.recs_out);
        for (uint64_t row = n_tiles * kRows + lane; row < a.n; row += 32) {
            uint64_t w[2];
            uint32_t bad[2];
            for (int s = 0; s < 2; s++) {
                const uint8_t *p = s ? a.umi_in + row * umi_len : a.bc_in + row * bc_len;
                const uint32_t len = s ? umi_len : bc_len;
                uint64_t acc = 0;
                uint32_t b = 0;
                for (uint32_t i = 0; i < len; i++) {
                    const uint32_t ch = p[i], c1 = (ch >> 1) & 3u, up = ch & 0xDFu;
                    acc |= (uint64_t)(c1 ^ (c1 >> 1)) << (2 * i);
                    b |= !(up == 0x41u || up == 0x43u || up == 0x47u || up == 0x54u);
                }
                w[s] = acc;
                bad[s] = b;
            }
            o64[3 * row] = w[0]; o64[3 * row + 1] = w[1];
            o64[3 * row + 2] = a.index ? a.index[row] : a.index_base + row;
            n_bb += bad[0]; n_bu += bad[1]; n_br += (bad[0] | bad[1]);
            if (a.flags) a.flags[row] = (uint8_t)(bad[0] | (bad[1] << 1));
        }
    }
    if (a.res_blocks) {
        const uint64_t c_bb = __reduce_add_sync(0xffffffffu, n_bb), c_bu = __reduce_add_sync(0xffffffffu, n_bu),
                       c_br = __reduce_add_sync(0xffffffffu, n_br);
        if (c_br) red_spread(a.res_blocks, blockIdx.x * kWarpsPerBlock + warp, lane, 0, 0, 0, 0, c_bb, c_bu, c_br);
    }
}

// ============================================================================ generators
__device__ __forceinline__ uint64_t low_mask_dev(uint32_t len) {
    return len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1);
}

__global__ void __launch_bounds__(kBlockThreads)
k_generate_records(uint64_t *__restrict__ out, uint64_t first, uint64_t n, uint32_t bc_len,
                   uint32_t umi_len, int mode, uint64_t param, uint64_t seed) {
    const uint64_t mb = low_mask_dev(bc_len), mu = low_mask_dev(umi_len);
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + k;
        const uint64_t key = splitmix64(seed ^ splitmix64(i));
        const uint64_t rb = splitmix64(key ^ 1), ru = splitmix64(key ^ 2);
        uint64_t b, u;
        if (mode == IBU_GEN_PATTERN) {
            b = i % 1000000ull;
            u = (i * 31ull) % 1000000ull;
        } else if (mode == IBU_GEN_SORTED) {
            uint64_t rpb = param & 0xFFFFFFFFull, dup = param >> 32;
            if (rpb == 0) rpb = 1000;
            if (dup == 0) dup = 1;
            b = (i / rpb) & mb;
            u = ((i % rpb) / dup) & mu;
        } else if (mode == IBU_GEN_WHITELIST) {
            uint64_t nb = param & 0xFFFFFFFFull, us = param >> 32;
            if (nb == 0) nb = 1000;
            b = splitmix64((rb % nb) ^ seed ^ 0xB) & mb;
            u = (us ? ru % us : ru) & mu;
        } else if (mode == IBU_GEN_ZIPF) {
            uint64_t nb = param & 0xFFFFFFFFull, us = param >> 32;
            if (nb == 0) nb = 1000;
            const uint32_t e = (uint32_t)(rb % (uint64_t)(64 - __clzll((long long)nb)));  // 0 .. floor(log2 nb)
            const uint64_t r = (((1ull << e) - 1) + (splitmix64(key ^ 7) & ((1ull << e) - 1))) % nb;
            b = splitmix64(r ^ seed ^ 0xB) & mb;
            u = (us ? ru % us : ru) & mu;
        } else {
            b = rb & mb;
            u = ru & mu;
            if (mode == IBU_GEN_DIRTY) {
                const uint64_t rd = splitmix64(key ^ 3);
                if (rd % 1000000ull < param) {
                    if (rd >> 63) b = rb; else u = ru;
                }
            }
        }
        out[3 * k] = b; out[3 * k + 1] = u; out[3 * k + 2] = i;
    }
}

__global__ void __launch_bounds__(kBlockThreads)
k_generate_ascii(uint8_t *__restrict__ out, uint64_t first_row, uint64_t n_rows, uint32_t len,
                 uint64_t dirty_ppm, uint64_t lower_ppm, uint64_t seed) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_rows;
         k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = splitmix64(seed ^ splitmix64(first_row + k));
        const uint64_t w = splitmix64(key ^ 4), rd = splitmix64(key ^ 5), rl = splitmix64(key ^ 6);
        const uint32_t lower = (rl % 1000000ull < lower_ppm) ? 0x20u : 0u;
        const uint32_t npos = (rd % 1000000ull < dirty_ppm) ? (uint32_t)((rd >> 40) % len) : 0xFFFFFFFFu;
        uint8_t *row = out + k * len;
        for (uint32_t i = 0; i < len; i++) {
            uint32_t ch = ((kAcgt >> (8 * ((w >> (2 * i)) & 3u))) & 0xFFu) | lower;
            row[i] = (uint8_t)(i == npos ? 'N' : ch);
        }
    }
}

// ============================================================================ launch helpers
// one tile per warp, plus the warp that takes the ragged tail
static unsigned tile_grid(uint64_t n_tiles) { return (unsigned)((n_tiles + 1 + kWarpsPerBlock - 1) / kWarpsPerBlock); }

// Shared-memory carve-out (percent of 228 KB, -1 = driver default) for a streaming kernel.  It
// fixes both the number of resident CTAs and what is left as L1, and these kernels are sensitive
// to the pair: unpack bc16/umi12 on 10^8 records runs in 0.864 ms at the default (6 CTAs, 28 KB
// L1), 0.766 at 75 % (5 CTAs), 0.763 at 64 % (4 CTAs, 92 KB L1), 0.800 at 50 % (3 CTAs); padding the
// request to get 4 CTAs WITHOUT enlarging L1 gave 0.876.  profiles/r1_carveout_sweep*.txt.
// MaxDynamicSharedMemorySize is an attribute of the function (per device), not of a launch: setting
// it to each launch's own request let two host threads launching one runtime-length instantiation
// with different lengths interleave set(large), set(small), launch(large) -> invalid value.  It is
// set once per kernel and device to the most any length can ask for.
static int set_max_dyn_smem(const void *kern, int device, size_t bytes, ibu_error_t *err) {
    static std::mutex m;
    static std::vector<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lock(m);
    for (const auto &d : done)
        if (d.first == kern && d.second == device) return IBU_OK;
    IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done.emplace_back(kern, device);
    return IBU_OK;
}
constexpr size_t kUnpackMaxSmem = (size_t)(kTileBytes + 2 * kTileRecords * 32) * kWarpsPerBlock;  // any lengths: 90 KB
constexpr size_t kPackMaxSmem = (size_t)(128 * 24 + 2 * (128 * 32 + 16)) * kWarpsPerBlock;      // Q = 4 rows per lane: 89 KB

static int set_carveout(const void *kern, int carve, ibu_error_t *err) {
    static const int env = getenv("IBU_CARVEOUT") ? atoi(getenv("IBU_CARVEOUT")) : -2;  // tuning hook
    if (env != -2) carve = env;
    if (carve >= 0) IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    return IBU_OK;
}

// One launch's hold on an entry of the context's result-scratch ring.  acquire(): take the next
// entry and make stream s wait for the fold of its previous user; fold(): fold the blocks into
// d_result behind the kernel and record the entry's event.  The entry's mutex is held in between,
// so a second host thread that wraps around the ring cannot read a stale event.
struct ResultLease {
    ibu_result_scratch *r = nullptr;
    ~ResultLease() {
        if (r) r->in_use.unlock();
    }
    int acquire(ibu_gpu_ctx *ctx, cudaStream_t s, ibu_error_t *err) {
        r = &ctx->result_ring[ctx->result_next.fetch_add(1, std::memory_order_relaxed) % kResultRing];
        r->in_use.lock();
        IBU_CUDA(cudaStreamWaitEvent(s, r->folded, 0));
        return IBU_OK;
    }
    unsigned long long *blocks() const { return r ? r->blocks : nullptr; }
    int fold(ibu_reduce_result_t *d_result, uint64_t n, cudaStream_t s, ibu_error_t *err) {
        k_fold_result<<<1, kResultBlocks, 0, s>>>(r->blocks, d_result, n);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        IBU_CUDA(cudaGetLastError());
        IBU_CUDA(cudaEventRecord(r->folded, s));
        return IBU_OK;
    }
};

static bool aligned(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

template <int BC, int UMI>
static int launch_unpack(ibu_gpu_ctx *ctx, UnpackArgs &a, cudaStream_t s, ibu_error_t *err) {
    // per-warp shared memory: input tile, then the staged outputs that are not stored directly;
    // when both outputs are staged (runtime-length barcode) the stage overlays the input tile
    constexpr bool kReuse = (BC == 0) && !(UMI == 16 || UMI == 32);
    uint32_t off = kReuse ? 0u : (uint32_t)kTileBytes;
    a.bc_stage_off = off;
    if (!(BC == 32 || BC == 16)) off += (kTileRecords * a.bc_len + 15u) & ~15u;
    a.umi_stage_off = off;
    if (!(UMI == 32 || UMI == 16)) off += (kTileRecords * a.umi_len + 15u) & ~15u;
    if (off < (uint32_t)kTileBytes) off = kTileBytes;
    a.warp_smem_bytes = off;
    const size_t smem = (size_t)off * kWarpsPerBlock;
    // with a result block the pass also carries K1's sums / checksum
    auto kern = a.res_blocks ? k_unpack<BC, UMI, true> : k_unpack<BC, UMI, false>;
    constexpr bool kStaged = !(BC == 32 || BC == 16) || !(UMI == 32 || UMI == 16);
    // staged kernels: the smallest configuration that holds 5 CTAs (bc16/umi12: 37 KB each -> 196 KB,
    // 60 KB of L1).  With the barrier-free result tail 4 and 5 CTAs are equal for bc16/umi12
    // (0.763 / 0.766 ms) and 5 is better for bc16/umi10 (0.776 vs 0.799); the default (6 CTAs,
    // 28 KB L1) costs 13 % and 3 CTAs 5 %.
    const int five_ctas = (int)std::min<size_t>(100, (5 * (smem + 1536) * 100 + 233471) / 233472);
    if (int rc = set_carveout((const void *)kern, kStaged ? five_ctas : -1, err)) return rc;
    if (int rc = set_max_dyn_smem((const void *)kern, ctx->device, kUnpackMaxSmem, err)) return rc;
    const unsigned grid = tile_grid(a.n / kTileRecords);
    kern<<<grid, kBlockThreads, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    return IBU_OK;
}

template <int BC, int UMI, int Q, int MINB>
static int launch_pack_q(ibu_gpu_ctx *ctx, PackArgs &a, cudaStream_t s, ibu_error_t *err) {
    constexpr uint32_t kRows = 32 * Q;  // rows per warp tile, Q per lane
    // per-warp shared memory: the record tile, then the staged inputs; with both inputs staged
    // the record tile overlays them (it is written after every row has been read)
    constexpr bool kReuse = (BC != 16 && BC != 32) && (UMI != 16 && UMI != 32);
    uint32_t off = kReuse ? 0u : kRows * 24;
    a.bc_stage_off = off;
    if (BC != 16 && BC != 32) off += ((kRows * a.bc_len + 15u) & ~15u) + 16u;  // +16: funnel-shift over-read
    a.umi_stage_off = off;
    if (UMI != 16 && UMI != 32) off += ((kRows * a.umi_len + 15u) & ~15u) + 16u;
    if (off < kRows * 24) off = kRows * 24;
    a.warp_smem_bytes = off;
    const size_t smem = (size_t)off * kWarpsPerBlock;
    auto kern = k_pack<BC, UMI, Q, MINB>;
    if (int rc = set_carveout((const void *)kern, -1, err)) return rc;
    if (int rc = set_max_dyn_smem((const void *)kern, ctx->device, kPackMaxSmem, err)) return rc;
    const unsigned grid = tile_grid(a.n / kRows);
    kern<<<grid, kBlockThreads, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    return IBU_OK;
}

// 4 rows per lane (128-row warp tiles) and at most 64 registers (4 resident CTAs): measured best
// or equal for every shape on 10^8 rows — bc32/umi32 7.07 TB/s, bc16/umi12 6.93, bc16/umi16 6.91,
// bc16/umi10 6.92 (2 rows per lane with the compiler's default allocation: 6.91 / 6.39 / 6.75 /
// 6.19; profiles/r1_k3_variants.txt).
// both lengths at run time: the prefetching kernel when two stages per warp still leave room for 4 CTAs
static int launch_pack_rt(ibu_gpu_ctx *ctx, PackArgs &a, cudaStream_t s, bool *done, ibu_error_t *err) {
    constexpr int Q = 4;
    constexpr uint32_t kRows = 32 * Q;
    *done = false;
    static const char *env = getenv("IBU_B200_K3_RT");  // 0 = the one-tile-per-warp kernel (tuning)
    if (env && env[0] == '0') return IBU_OK;
    uint32_t off = 0;
    a.bc_stage_off = off;
    off += ((kRows * a.bc_len + 15u) & ~15u) + 16u;  // +16: funnel-shift over-read
    a.umi_stage_off = off;
    off += ((kRows * a.umi_len + 15u) & ~15u) + 16u;
    if (off < kRows * 24) off = kRows * 24;
    const size_t smem = (size_t)2 * off * kWarpsPerBlock;
    if (smem > (56u << 10)) return IBU_OK;
    a.warp_smem_bytes = off;
    auto kern = k_pack_rt<Q, 4>;
    if (int rc = set_carveout((const void *)kern, -1, err)) return rc;
    if (int rc = set_max_dyn_smem((const void *)kern, ctx->device, 56u << 10, err)) return rc;
    const uint64_t tiles = a.n / kRows + 1;  // (+1: the warp that owns the ragged tail)
    const uint64_t warps = (tiles + kPackRtTpw - 1) / kPackRtTpw;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, (warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    kern<<<grid, kBlockThreads, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    *done = true;
    return IBU_OK;
}

template <int BC, int UMI>
static int launch_pack(ibu_gpu_ctx *ctx, PackArgs &a, cudaStream_t s, ibu_error_t *err) {
    if constexpr (BC == 0 && UMI == 0) {
        bool done = false;
        if (int rc = launch_pack_rt(ctx, a, s, &done, err)) return rc;
        if (done) return IBU_OK;
    }
    // (a 32-byte row next to a runtime-length one needs a few more registers: 3 resident CTAs)
    constexpr int kMinB = ((BC == 0 && UMI == 32) || (BC == 32 && UMI == 0)) ? 3 : 4;
    return launch_pack_q<BC, UMI, 4, kMinB>(ctx, a, s, err);
}

static int check_lens(uint32_t bc_len, uint32_t umi_len, ibu_error_t *err) {
    // same bounds as Header::validate (header.rs:179-184)
    if (bc_len == 0 || bc_len > 32)
        return set_error(err, IBU_ERR_INVALID_BARCODE_LENGTH, 0, bc_len, 0,
                         "Invalid barcode length: %u (must be 1-32)", bc_len);
    if (umi_len == 0 || umi_len > 32)
        return set_error(err, IBU_ERR_INVALID_UMI_LENGTH, 0, umi_len, 0,
                         "Invalid UMI length: %u (must be 1-32)", umi_len);
    return IBU_OK;
}

}  // namespace ibu

using namespace ibu;

extern "C" {

int ibu_gpu_validate_reduce_async(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n,
                                  uint32_t bc_len, uint32_t umi_len, ibu_reduce_result_t *d_result,
                                  void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !d_result || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (int rc = check_lens(bc_len, umi_len, err)) return rc;
    if (!aligned(d_records, 8)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "d_records must be 8-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    ResultLease lease;
    if (int rc = lease.acquire(ctx, s, err)) return rc;
    // records until the next 32-byte boundary: 24 h = -p (mod 32)  <=>  h = (p / 8) mod 4
    const uint32_t head = (uint32_t)std::min<uint64_t>(n, ((uintptr_t)d_records >> 3) & 3u);
    const uint64_t n_tiles = (n - head) / kTileRecords;
    constexpr int kTpw = 4;  // tiles per warp: a CTA covers 4096 records
    const unsigned grid = (unsigned)std::max<uint64_t>(1, (n_tiles + kWarpsPerBlock * kTpw - 1) / (kWarpsPerBlock * kTpw));
    k_validate_reduce<kTpw><<<grid, kBlockThreads, 0, s>>>((const uint8_t *)d_records, n, head, high_mask(bc_len),
                                                           high_mask(umi_len), lease.blocks());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    return lease.fold(d_result, n, s, err);
}

int ibu_gpu_unpack_async(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n,
                         uint32_t bc_len, uint32_t umi_len, uint8_t *d_bc_ascii, uint8_t *d_umi_ascii,
                         uint8_t *d_flags, ibu_reduce_result_t *d_result, void *stream,
                         ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (n && (!d_records || !d_bc_ascii || !d_umi_ascii)))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (int rc = check_lens(bc_len, umi_len, err)) return rc;
    if (!aligned(d_records, 16) || !aligned(d_bc_ascii, 16) || !aligned(d_umi_ascii, 16))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "device pointers must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    ResultLease lease;
    if (d_result)
        if (int rc = lease.acquire(ctx, s, err)) return rc;
    UnpackArgs a{};
    a.recs = (const uint8_t *)d_records;
    a.n = n;
    a.bc_out = d_bc_ascii;
    a.umi_out = d_umi_ascii;
    a.flags = d_flags;
    a.bc_hi = high_mask(bc_len);
    a.umi_hi = high_mask(umi_len);
    a.res_blocks = lease.blocks();
    a.bc_len = bc_len;
    a.umi_len = umi_len;
    // store mode per output: 32 / 16 = direct vector store (needs that alignment), 12 =
    // compile-time staged, 0 = runtime-length staged
    const int bm = (bc_len == 32 && aligned(d_bc_ascii, 32)) ? 32 : bc_len == 16 ? 16 : 0;
    const int um = (umi_len == 32 && aligned(d_umi_ascii, 32)) ? 32
                   : umi_len == 16 ? 16 : umi_len == 12 ? 12 : umi_len == 10 ? 10 : 0;
    int rc = -1;
#define IBU_UNPACK_CASE(B, U) \
    if (bm == B && um == U) rc = launch_unpack<B, U>(ctx, a, s, err);
    IBU_UNPACK_CASE(16, 12) IBU_UNPACK_CASE(16, 16) IBU_UNPACK_CASE(16, 32) IBU_UNPACK_CASE(16, 0)
    IBU_UNPACK_CASE(16, 10) IBU_UNPACK_CASE(32, 10) IBU_UNPACK_CASE(0, 10)
    IBU_UNPACK_CASE(32, 12) IBU_UNPACK_CASE(32, 16) IBU_UNPACK_CASE(32, 32) IBU_UNPACK_CASE(32, 0)
    IBU_UNPACK_CASE(0, 12) IBU_UNPACK_CASE(0, 16) IBU_UNPACK_CASE(0, 32) IBU_UNPACK_CASE(0, 0)
#undef IBU_UNPACK_CASE
    if (rc < 0) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "no unpack kernel for this shape");
    if (rc != IBU_OK || !lease.r) return rc;
    return lease.fold(d_result, n, s, err);
}

int ibu_gpu_pack_async(ibu_gpu_ctx_t *ctx, const uint8_t *d_bc_ascii, const uint8_t *d_umi_ascii,
                       const uint64_t *d_index, uint64_t index_base, uint64_t n, uint32_t bc_len,
                       uint32_t umi_len, ibu_record_t *d_records, uint8_t *d_flags,
                       ibu_reduce_result_t *d_result, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (n && (!d_records || !d_bc_ascii || !d_umi_ascii)))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (int rc = check_lens(bc_len, umi_len, err)) return rc;
    if (!aligned(d_records, 16) || !aligned(d_bc_ascii, 16) || !aligned(d_umi_ascii, 16) ||
        !aligned(d_index, 8))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "device pointers must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    ResultLease lease;
    if (d_result)
        if (int rc = lease.acquire(ctx, s, err)) return rc;
    PackArgs a{};
    a.bc_in = d_bc_ascii;
    a.umi_in = d_umi_ascii;
    a.index = d_index;
    a.index_base = index_base;
    a.n = n;
    a.recs_out = (uint8_t *)d_records;
    a.flags = d_flags;
    a.res_blocks = lease.blocks();
    a.bc_len = bc_len;
    a.umi_len = umi_len;
    const int bm = (bc_len == 32 && aligned(d_bc_ascii, 32)) ? 32 : bc_len == 16 ? 16 : 0;
    const int um = (umi_len == 32 && aligned(d_umi_ascii, 32)) ? 32
                   : umi_len == 16 ? 16 : umi_len == 12 ? 12 : umi_len == 10 ? 10 : 0;
    int rc = -1;
#define IBU_PACK_CASE(B, U) \
    if (bm == B && um == U) rc = launch_pack<B, U>(ctx, a, s, err);
    IBU_PACK_CASE(32, 32) IBU_PACK_CASE(32, 16) IBU_PACK_CASE(32, 12) IBU_PACK_CASE(32, 0)
    IBU_PACK_CASE(16, 32) IBU_PACK_CASE(16, 16) IBU_PACK_CASE(16, 12) IBU_PACK_CASE(16, 0)
    IBU_PACK_CASE(0, 12) IBU_PACK_CASE(16, 10) IBU_PACK_CASE(32, 10) IBU_PACK_CASE(0, 10)
    IBU_PACK_CASE(0, 32) IBU_PACK_CASE(0, 16) IBU_PACK_CASE(0, 0)
#undef IBU_PACK_CASE
    if (rc < 0) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "no pack kernel for this shape");
    if (rc != IBU_OK || !lease.r) return rc;
    return lease.fold(d_result, n, s, err);
}

int ibu_gpu_generate_records_async(ibu_gpu_ctx_t *ctx, ibu_record_t *d_records, uint64_t first,
                                   uint64_t n, uint32_t bc_len, uint32_t umi_len, int mode,
                                   uint64_t param, uint64_t seed, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (int rc = check_lens(bc_len, umi_len, err)) return rc;
    if (mode < IBU_GEN_CLEAN || mode > IBU_GEN_ZIPF) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "bad mode");
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    uint64_t blocks = (n + kBlockThreads - 1) / kBlockThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * 16;
    k_generate_records<<<(int)(blocks < cap ? blocks : cap), kBlockThreads, 0, s>>>(
        (uint64_t *)d_records, first, n, bc_len, umi_len, mode, param, seed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    return IBU_OK;
}

int ibu_gpu_generate_ascii_async(ibu_gpu_ctx_t *ctx, uint8_t *d_ascii, uint64_t first_row,
                                 uint64_t n_rows, uint32_t len, uint64_t dirty_ppm, uint64_t lower_ppm,
                                 uint64_t seed, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (!d_ascii && n_rows)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (len == 0 || len > 32) return set_error(err, IBU_ERR_ARG, 0, len, 0, "row length must be 1-32");
    if (n_rows == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = pick_stream(ctx, stream);
    uint64_t blocks = (n_rows + kBlockThreads - 1) / kBlockThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * 16;
    k_generate_ascii<<<(int)(blocks < cap ? blocks : cap), kBlockThreads, 0, s>>>(
        d_ascii, first_row, n_rows, len, dirty_ppm, lower_ppm, seed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    IBU_CUDA(cudaGetLastError());
    return IBU_OK;
}

}  // extern "C"
