// k2lab.cu — tuning lab for K2 (unpack bc16/umi12): A/B timing of kernel structures in one
// process, next to the shipped kernel (called through the C ABI) and next to pure-traffic
// kernels that move K2's byte mix (24 B read : 28 B written per record) with no decode at all.
// The traffic kernels measure what the memory system gives this mix; K2 cannot beat them.
// Build: see tools/Makefile (links ../ibu_b200/libibu_b200.so).  Prints JSON lines.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../ibu_b200/csrc/kernels.cuh"
#include "../include/ibu_b200.h"

using namespace ibu;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e__ = (x);                                                             \
        if (e__ != cudaSuccess) {                                                          \
            printf("{\"error\":\"%s at %s:%d\"}\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------ pure traffic, K2's mix
// Per warp tile: read 3072 B, write 2048 B to stream A and 1536 B to stream B.  Register data
// only (xor so the loads cannot be dropped); one tile prefetched ahead like K2.
template <int RU, int AU, int BU>
__global__ void __launch_bounds__(256) k_mix(const uint4 *__restrict__ in, uint4 *__restrict__ a,
                                             uint4 *__restrict__ b, uint64_t n_tiles) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gwarp = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5), total = (uint64_t)gridDim.x * 8;
    uint4 pre[RU];
    uint32_t sink = 0;
    uint64_t t = gwarp;
    if (t < n_tiles) {
#pragma unroll
        for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + t * (32 * RU) + lane + 32 * k);
    }
    while (t < n_tiles) {
        uint4 v[RU];
#pragma unroll
        for (int k = 0; k < RU; k++) v[k] = pre[k];
        const uint64_t tn = t + total;
        if (tn < n_tiles) {
#pragma unroll
            for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + tn * (32 * RU) + lane + 32 * k);
        }
#pragma unroll
        for (int k = 0; k < RU; k++) sink ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
#pragma unroll
        for (int k = 0; k < AU; k++) stg_stream(a + t * (32 * AU) + lane + 32 * k, v[k % RU]);
#pragma unroll
        for (int k = 0; k < BU; k++) stg_stream(b + t * (32 * BU) + lane + 32 * k, v[(k + 3) % RU]);
        t = tn;
    }
    if (sink == 0x9e3779b9u) b[0].x = sink;  // keeps every load alive
}

// CTA-level tiles: a 256-thread CTA owns 8 consecutive warp tiles and every instruction of the
// CTA covers 4 KB contiguous (warp w takes the w-th 512 B), unlike k_mix where a warp owns 3 KB.
template <int RU, int AU, int BU>
__global__ void __launch_bounds__(256) k_mix_cta(const uint4 *__restrict__ in, uint4 *__restrict__ a,
                                                 uint4 *__restrict__ b, uint64_t n_tiles) {
    const uint64_t n_ct = n_tiles / 8;
    uint4 pre[RU];
    uint32_t sink = 0;
    uint64_t t = blockIdx.x;
    if (t < n_ct) {
#pragma unroll
        for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + t * (256 * RU) + threadIdx.x + 256 * k);
    }
    while (t < n_ct) {
        uint4 v[RU];
#pragma unroll
        for (int k = 0; k < RU; k++) v[k] = pre[k];
        const uint64_t tn = t + gridDim.x;
        if (tn < n_ct) {
#pragma unroll
            for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + tn * (256 * RU) + threadIdx.x + 256 * k);
        }
#pragma unroll
        for (int k = 0; k < RU; k++) sink ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
#pragma unroll
        for (int k = 0; k < AU; k++) stg_stream(a + t * (256 * AU) + threadIdx.x + 256 * k, v[k % RU]);
#pragma unroll
        for (int k = 0; k < BU; k++) stg_stream(b + t * (256 * BU) + threadIdx.x + 256 * k, v[(k + 3) % RU]);
        t = tn;
    }
    if (sink == 0x9e3779b9u) b[0].x = sink;
}

// Non-persistent form: CTA c owns the contiguous run of 8*ROUNDS warp tiles starting at
// 8*ROUNDS*c; round j gives warp w tile 8*(ROUNDS*c + j) + w.  Grid = n_tiles / (8 ROUNDS):
// block scheduling (not a lock-step grid stride) decides which addresses are live together.
// HINT: 0 plain stores, 1 st.cs, 2 st.L1::no_allocate.
template <int HINT>
__device__ __forceinline__ void stg_h(uint4 *p, uint4 v) {
    if (HINT == 0) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (HINT == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (HINT == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int RU, int AU, int BU, int ROUNDS, int HINT>
__global__ void __launch_bounds__(256) k_mix_np(const uint4 *__restrict__ in, uint4 *__restrict__ a,
                                                uint4 *__restrict__ b, uint64_t n_tiles) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t sink = 0;
    uint4 pre[RU];
    uint64_t t = 8ull * ROUNDS * blockIdx.x + warp;
    if (t < n_tiles) {
#pragma unroll
        for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + t * (32 * RU) + lane + 32 * k);
    }
#pragma unroll 1
    for (int j = 0; j < ROUNDS && t < n_tiles; j++) {
        uint4 v[RU];
#pragma unroll
        for (int k = 0; k < RU; k++) v[k] = pre[k];
        const uint64_t tn = t + 8;
        if (j + 1 < ROUNDS && tn < n_tiles) {
#pragma unroll
            for (int k = 0; k < RU; k++) pre[k] = ldg_stream(in + tn * (32 * RU) + lane + 32 * k);
        }
#pragma unroll
        for (int k = 0; k < RU; k++) sink ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
#pragma unroll
        for (int k = 0; k < AU; k++) stg_h<HINT>(a + t * (32 * AU) + lane + 32 * k, v[k % RU]);
#pragma unroll
        for (int k = 0; k < BU; k++) stg_h<HINT>(b + t * (32 * BU) + lane + 32 * k, v[(k + 3) % RU]);
        t = tn;
    }
    if (sink == 0x9e3779b9u) b[0].x = sink;
}

// torch-style elementwise copy: non-persistent, 128 threads, 4 x 16 B per thread, block-contiguous 8 KB
__global__ void __launch_bounds__(128) k_copy_np(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint64_t n16) {
    const uint64_t base = (uint64_t)blockIdx.x * 512 + threadIdx.x;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + 128 * k < n16) v[k] = in[base + 128 * k];
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + 128 * k < n16) out[base + 128 * k] = v[k];
}

// The same bytes with the simplest possible mapping: thread i of the grid owns 16-byte unit i of
// every "super row" (grid-stride over the three streams independently, torch-copy style).
__global__ void __launch_bounds__(256) k_mix_flat(const uint4 *__restrict__ in, uint4 *__restrict__ a,
                                                  uint4 *__restrict__ b, uint64_t n_tiles) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    // 13 units of 16 B per 4 records: 6 in, 4 out A, 3 out B.  Thread handles groups of 4 records.
    const uint64_t groups = n_tiles * 32;
    for (uint64_t g = tid; g < groups; g += nth) {
        uint4 v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = ldg_stream(in + g * 6 + k);
        v[0].x ^= v[5].w;
#pragma unroll
        for (int k = 0; k < 4; k++) stg_stream(a + g * 4 + k, v[k]);
#pragma unroll
        for (int k = 0; k < 3; k++) stg_stream(b + g * 3 + k, v[(k + 3) % 6]);
    }
}

// plain copy of nbytes (torch copy_ analogue), 4 x 16 B per thread per step
__global__ void __launch_bounds__(256) k_copy(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint64_t n16) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i + 3 * nth < n16; i += 4 * nth) {
        uint4 v0 = ldg_stream(in + i), v1 = ldg_stream(in + i + nth), v2 = ldg_stream(in + i + 2 * nth),
              v3 = ldg_stream(in + i + 3 * nth);
        stg_stream(out + i, v0); stg_stream(out + i + nth, v1); stg_stream(out + i + 2 * nth, v2);
        stg_stream(out + i + 3 * nth, v3);
    }
}

// ------------------------------------------------------------------ K2 with bulk async copies
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One warp = one independent pipeline: NS input stages (3 KB tiles landed by cp.async.bulk on an
// mbarrier), OS output stages (2 KB + 1.5 KB, leaving by cp.async.bulk S2G).  TQ = tiles of 128
// records per stage (bigger bulk transfers per instruction).
template <int NS, int OS, int WARPS, int TQ>
__global__ void __launch_bounds__(WARPS * 32)
k_unpack_tma(const uint8_t *__restrict__ recs, uint64_t n_tiles128, uint8_t *__restrict__ bc_out,
             uint8_t *__restrict__ umi_out, uint64_t bc_hi, uint64_t umi_hi, unsigned long long *res) {
    constexpr uint32_t IN_B = 3072 * TQ, BC_B = 2048 * TQ, UMI_B = 1536 * TQ;
    constexpr uint32_t WARP_B = NS * IN_B + OS * (BC_B + UMI_B) + 8 * NS + ((8 * NS) % 16 ? 8 : 0);
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t gwarp = (uint64_t)blockIdx.x * WARPS + warp, total = (uint64_t)gridDim.x * WARPS;
    uint8_t *wsm = smem + warp * WARP_B;
    uint8_t *in_s = wsm, *bc_s = wsm + NS * IN_B, *umi_s = bc_s + OS * BC_B;
    const uint32_t bar0 = smem_u32(umi_s + OS * UMI_B);
    const uint64_t n_tiles = n_tiles128 / TQ;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const uint64_t t = gwarp + (uint64_t)s * total;
            if (t < n_tiles) {
                mbar_expect_tx(bar0 + 8 * s, IN_B);
                bulk_g2s(smem_u32(in_s + s * IN_B), recs + t * IN_B, IN_B, bar0 + 8 * s);
            }
        }
    }
    __syncwarp();

    uint32_t n_bb = 0, n_bu = 0, n_br = 0;
    uint64_t s_bc = 0, s_umi = 0, s_idx = 0, x_all = 0;
    uint32_t k = 0;
    for (uint64_t t = gwarp; t < n_tiles; t += total, k++) {
        const uint32_t s = k % NS, o = k % OS;
        mbar_wait(bar0 + 8 * s, (k / NS) & 1u);
        if (k >= OS) {  // the bulk store that last read output stage o must have drained it
            if (lane == 0) bulk_wait_read<OS - 1>();
            __syncwarp();
        }
        const uint64_t *in64 = reinterpret_cast<const uint64_t *>(in_s + s * IN_B);
        uint4 *bc4 = reinterpret_cast<uint4 *>(bc_s + o * BC_B);
        uint32_t *umi32 = reinterpret_cast<uint32_t *>(umi_s + o * UMI_B);
#pragma unroll
        for (int q = 0; q < 4 * TQ; q++) {
            const uint32_t r = lane + 32 * q;
            const uint64_t bc = in64[3 * r], umi = in64[3 * r + 1], idx = in64[3 * r + 2];
            uint32_t asc[8];
            decode_word<4>(bc, asc);
            bc4[r] = make_uint4(asc[0], asc[1], asc[2], asc[3]);
            decode_word<3>(umi, asc);
            umi32[3 * r] = asc[0]; umi32[3 * r + 1] = asc[1]; umi32[3 * r + 2] = asc[2];
            const uint32_t bb = (bc & bc_hi) != 0ull, bu = (umi & umi_hi) != 0ull;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            s_bc += bc; s_umi += umi; s_idx += idx; x_all ^= bc ^ umi ^ idx;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(bc_out + t * BC_B, smem_u32(bc4), BC_B);
            bulk_s2g(umi_out + t * UMI_B, smem_u32(umi32), UMI_B);
            bulk_commit();
            const uint64_t tn = t + (uint64_t)NS * total;  // refill the input stage just consumed
            if (tn < n_tiles) {
                mbar_expect_tx(bar0 + 8 * s, IN_B);
                bulk_g2s(smem_u32(in_s + s * IN_B), recs + tn * IN_B, IN_B, bar0 + 8 * s);
            }
        }
    }
    if (lane == 0) bulk_wait_all();
    n_bb = __reduce_add_sync(0xffffffffu, n_bb);
    n_bu = __reduce_add_sync(0xffffffffu, n_bu);
    n_br = __reduce_add_sync(0xffffffffu, n_br);
    s_bc = warp_sum64(s_bc); s_umi = warp_sum64(s_umi); s_idx = warp_sum64(s_idx); x_all = warp_xor64(x_all);
    if (lane == 0) {
        atomicAdd(res + 1, (unsigned long long)s_bc); atomicAdd(res + 2, (unsigned long long)s_umi);
        atomicAdd(res + 3, (unsigned long long)s_idx); atomicXor(res + 4, (unsigned long long)x_all);
        if (n_bb) atomicAdd(res + 5, (unsigned long long)n_bb);
        if (n_bu) atomicAdd(res + 6, (unsigned long long)n_bu);
        if (n_br) atomicAdd(res + 7, (unsigned long long)n_br);
        if (gwarp == 0) atomicAdd(res, (unsigned long long)n_tiles128 * 128ull);
    }
}

// Non-persistent K2: CTA c owns 8*TPW consecutive 128-record tiles; a warp takes one tile per
// round (prefetching the next round's tile when TPW > 1).  No grid-stride loop.
template <int TPW, int HINT, int WARPS = 8, int TAIL = 0>
__global__ void __launch_bounds__(WARPS * 32)
k_unpack_np(const uint8_t *__restrict__ recs, uint64_t n_tiles, uint8_t *__restrict__ bc_out,
            uint8_t *__restrict__ umi_out, uint64_t bc_hi, uint64_t umi_hi, unsigned long long *res) {
    __shared__ __align__(16) uint8_t smem[WARPS * (3072 + 1536)];
    __shared__ uint64_t red[WARPS][8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + warp * (3072 + 1536);
    uint4 *in4 = reinterpret_cast<uint4 *>(wsm);
    const uint64_t *in64 = reinterpret_cast<const uint64_t *>(wsm);
    uint32_t *umi_stage = reinterpret_cast<uint32_t *>(wsm + 3072);
    const uint4 *g4 = reinterpret_cast<const uint4 *>(recs);
    uint32_t n_bb = 0, n_bu = 0, n_br = 0;
    uint64_t s_bc = 0, s_umi = 0, s_idx = 0, x_all = 0;
    uint64_t t = (uint64_t)WARPS * TPW * blockIdx.x + warp;
    uint4 pre[6];
    if (t < n_tiles) {
#pragma unroll
        for (int k = 0; k < 6; k++) pre[k] = ldg_stream(g4 + t * 192 + lane + 32 * k);
    }
#pragma unroll 1
    for (int j = 0; j < TPW && t < n_tiles; j++) {
#pragma unroll
        for (int k = 0; k < 6; k++) in4[lane + 32 * k] = pre[k];
        __syncwarp();
        const uint64_t tn = t + WARPS;
        if (TPW > 1 && j + 1 < TPW && tn < n_tiles) {
#pragma unroll
            for (int k = 0; k < 6; k++) pre[k] = ldg_stream(g4 + tn * 192 + lane + 32 * k);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t r = lane + 32 * q;
            const uint64_t bc = in64[3 * r], umi = in64[3 * r + 1], idx = in64[3 * r + 2];
            uint32_t asc[8];
            decode_word<4>(bc, asc);
            stg_h<HINT>(reinterpret_cast<uint4 *>(bc_out) + t * 128 + r, make_uint4(asc[0], asc[1], asc[2], asc[3]));
            decode_word<3>(umi, asc);
            umi_stage[3 * r] = asc[0]; umi_stage[3 * r + 1] = asc[1]; umi_stage[3 * r + 2] = asc[2];
            const uint32_t bb = (bc & bc_hi) != 0ull, bu = (umi & umi_hi) != 0ull;
            n_bb += bb; n_bu += bu; n_br += (bb | bu);
            s_bc += bc; s_umi += umi; s_idx += idx; x_all ^= bc ^ umi ^ idx;
        }
        __syncwarp();
        const uint4 *st4 = reinterpret_cast<const uint4 *>(umi_stage);
#pragma unroll
        for (int k = 0; k < 3; k++) stg_h<HINT>(reinterpret_cast<uint4 *>(umi_out) + t * 96 + lane + 32 * k, st4[lane + 32 * k]);
        if (TPW > 1) __syncwarp();
        t = tn;
    }
    if (TAIL == 2) return;
    s_bc = warp_sum64(s_bc); s_umi = warp_sum64(s_umi); s_idx = warp_sum64(s_idx); x_all = warp_xor64(x_all);
    n_bb = __reduce_add_sync(0xffffffffu, n_bb);
    n_bu = __reduce_add_sync(0xffffffffu, n_bu);
    n_br = __reduce_add_sync(0xffffffffu, n_br);
    if (TAIL == 1) {  // no barrier: every warp REDs into one of 256 spread result blocks (res[256][8])
        const uint64_t v = lane == 1 ? s_bc : lane == 2 ? s_umi : lane == 3 ? s_idx : lane == 4 ? x_all
                           : lane == 5 ? n_bb : lane == 6 ? n_bu : n_br;
        unsigned long long *blk = res + 8 * ((blockIdx.x * WARPS + warp) & 255u);
        if (lane == 4) atomicXor(blk + 4, (unsigned long long)v);
        else if (lane >= 1 && lane < 8 && v) atomicAdd(blk + lane, (unsigned long long)v);
        return;
    }
    if (lane == 0) {
        red[warp][1] = s_bc; red[warp][2] = s_umi; red[warp][3] = s_idx; red[warp][4] = x_all;
        red[warp][5] = n_bb; red[warp][6] = n_bu; red[warp][7] = n_br;
    }
    __syncthreads();
    if (threadIdx.x >= 1 && threadIdx.x < 8) {
        uint64_t v = 0;
        for (int w = 0; w < WARPS; w++) { if (threadIdx.x == 4) v ^= red[w][4]; else v += red[w][threadIdx.x]; }
        if (threadIdx.x == 4) atomicXor(res + 4, (unsigned long long)v);
        else if (v) atomicAdd(res + threadIdx.x, (unsigned long long)v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(res, (unsigned long long)n_tiles * 128ull);
}

__global__ void k_diff(const uint4 *a, const uint4 *b, uint64_t n16, unsigned long long *count) {
    uint64_t bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 x = a[i], y = b[i];
        bad += (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (bad) atomicAdd(count, (unsigned long long)bad);
}

// ------------------------------------------------------------------ harness
struct Timer {
    cudaStream_t s;
    int iters;
    template <class F>
    void run(const char *name, double bytes, F f, const char *extra = "") {
        for (int i = 0; i < 3; i++) f();
        CK(cudaStreamSynchronize(s));
        std::vector<cudaEvent_t> a(iters), b(iters);
        for (int i = 0; i < iters; i++) { cudaEventCreate(&a[i]); cudaEventCreate(&b[i]); }
        for (int i = 0; i < iters; i++) { cudaEventRecord(a[i], s); f(); cudaEventRecord(b[i], s); }
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        std::vector<float> ms(iters);
        double sum = 0;
        for (int i = 0; i < iters; i++) { cudaEventElapsedTime(&ms[i], a[i], b[i]); sum += ms[i]; }
        std::sort(ms.begin(), ms.end());
        printf("{\"kernel\":\"%s\",\"ms_mean\":%.4f,\"ms_best\":%.4f,\"ms_median\":%.4f,\"gbs_mean\":%.1f,\"gbs_best\":%.1f%s}\n",
               name, sum / iters, ms[0], ms[iters / 2], bytes / (sum / iters) / 1e6, bytes / ms[0] / 1e6, extra);
        fflush(stdout);
        for (int i = 0; i < iters; i++) { cudaEventDestroy(a[i]); cudaEventDestroy(b[i]); }
    }
};

template <int NS, int OS, int WARPS, int TQ>
static void run_tma(Timer &T, int sms, const uint8_t *recs, uint64_t n, uint8_t *bc, uint8_t *umi,
                    unsigned long long *res, const uint8_t *ref_bc, const uint8_t *ref_umi,
                    const unsigned long long *ref_res_host, unsigned long long *d_count) {
    constexpr uint32_t IN_B = 3072 * TQ, BC_B = 2048 * TQ, UMI_B = 1536 * TQ;
    constexpr uint32_t WARP_B = NS * IN_B + OS * (BC_B + UMI_B) + 8 * NS + ((8 * NS) % 16 ? 8 : 0);
    const size_t smem = (size_t)WARP_B * WARPS;
    auto kern = k_unpack_tma<NS, OS, WARPS, TQ>;
    if (smem > 227 * 1024) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) return;
    const int grid = sms * per_sm;
    const uint64_t hi_bc = ~0ull << 32, hi_umi = ~0ull << 24;
    CK(cudaMemsetAsync(bc, 0, n * 16, T.s));
    CK(cudaMemsetAsync(umi, 0, n * 12, T.s));
    CK(cudaMemsetAsync(res, 0, 64, T.s));
    kern<<<grid, WARPS * 32, smem, T.s>>>(recs, n / 128, bc, umi, hi_bc, hi_umi, res);
    CK(cudaMemsetAsync(d_count, 0, 8, T.s));
    k_diff<<<sms * 8, 256, 0, T.s>>>((const uint4 *)bc, (const uint4 *)ref_bc, n, d_count);
    k_diff<<<sms * 8, 256, 0, T.s>>>((const uint4 *)umi, (const uint4 *)ref_umi, n * 12 / 16, d_count);
    unsigned long long bad = 0, got[8];
    CK(cudaMemcpyAsync(&bad, d_count, 8, cudaMemcpyDeviceToHost, T.s));
    CK(cudaMemcpyAsync(got, res, 64, cudaMemcpyDeviceToHost, T.s));
    CK(cudaStreamSynchronize(T.s));
    const bool res_ok = memcmp(got, ref_res_host, 64) == 0;
    char name[96], extra[160];
    snprintf(name, sizeof name, "k_unpack_tma<NS=%d,OS=%d,W=%d,TQ=%d>", NS, OS, WARPS, TQ);
    snprintf(extra, sizeof extra, ",\"ctas_per_sm\":%d,\"smem_per_cta\":%zu,\"mismatch_u4\":%llu,\"result_block_ok\":%s",
             per_sm, smem, bad, res_ok ? "true" : "false");
    T.run(name, 52.0 * n, [&] {
        cudaMemsetAsync(res, 0, 64, T.s);
        kern<<<grid, WARPS * 32, smem, T.s>>>(recs, n / 128, bc, umi, hi_bc, hi_umi, res);
    }, extra);
}

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;  // multiple of 256
    const int iters = argc > 2 ? atoi(argv[2]) : 20;
    if (n % 256) { printf("{\"error\":\"n must be a multiple of 256\"}\n"); return 1; }
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    ibu_error_t err;
    ibu_gpu_ctx_t *ctx = nullptr;
    if (ibu_gpu_ctx_create(0, nullptr, &ctx, &err)) { printf("{\"error\":\"%s\"}\n", err.msg); return 1; }
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    Timer T{s, iters};
    uint8_t *recs, *bc, *umi, *bc2, *umi2;
    unsigned long long *res, *res2, *d_count;
    CK(cudaMalloc(&recs, n * 24)); CK(cudaMalloc(&bc, n * 16)); CK(cudaMalloc(&umi, n * 12));
    CK(cudaMalloc(&bc2, n * 16)); CK(cudaMalloc(&umi2, n * 12));
    CK(cudaMalloc(&res, 64)); CK(cudaMalloc(&res2, 64)); CK(cudaMalloc(&d_count, 8));
    if (ibu_gpu_generate_records_async(ctx, (ibu_record_t *)recs, 0, n, 16, 12, IBU_GEN_DIRTY, 10000, 2024, s, &err)) {
        printf("{\"error\":\"%s\"}\n", err.msg); return 1;
    }
    // the shipped kernel: reference outputs + same-run baseline
    auto shipped = [&] {
        ibu_gpu_unpack_async(ctx, (const ibu_record_t *)recs, n, 16, 12, bc, umi, nullptr,
                             (ibu_reduce_result_t *)res, s, &err);
    };
    shipped();
    unsigned long long ref_res[8];
    CK(cudaMemcpyAsync(ref_res, res, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    T.run("shipped k_unpack<16,12> (C ABI)", 52.0 * n, shipped);

    // pure traffic: K2's mix and 50/50 mixes on the same structures (buffers big enough for all)
    const uint64_t n_tiles = n / 128;
    const bool only_k2 = argc > 3;
    if (!only_k2) {
    uint4 *tin, *ta, *tb;
    CK(cudaMalloc(&tin, n_tiles * 512 * 8)); CK(cudaMalloc(&ta, n_tiles * 512 * 8)); CK(cudaMalloc(&tb, n_tiles * 512 * 4));
    CK(cudaMemsetAsync(tin, 1, n_tiles * 512 * 8, s));
#define MIX(K, R, A, B, PER_SM, LABEL)                                                                  \
    {                                                                                                   \
        char nm[96];                                                                                    \
        snprintf(nm, sizeof nm, "traffic %s r%d:a%d:b%d x512B/tile, %d CTAs/SM", LABEL, R, A, B, PER_SM); \
        T.run(nm, 512.0 * (R + A + B) * n_tiles, [&] { K<R, A, B><<<sms * PER_SM, 256, 0, s>>>(tin, ta, tb, n_tiles); }); \
    }
    MIX(k_mix, 6, 4, 3, 2, "warp-tiles") MIX(k_mix, 6, 4, 3, 4, "warp-tiles")
    MIX(k_mix_cta, 6, 4, 3, 2, "cta-tiles") MIX(k_mix_cta, 6, 4, 3, 4, "cta-tiles") MIX(k_mix_cta, 6, 4, 3, 8, "cta-tiles")
    MIX(k_mix, 6, 6, 0, 4, "warp-tiles") MIX(k_mix_cta, 6, 6, 0, 4, "cta-tiles")
    MIX(k_mix, 8, 3, 0, 4, "warp-tiles (K3 32/32 mix)") MIX(k_mix, 7, 6, 0, 4, "warp-tiles (K3 16/12 mix)")
    MIX(k_mix, 8, 4, 4, 4, "warp-tiles") MIX(k_mix, 4, 4, 0, 8, "warp-tiles")
    MIX(k_mix, 6, 0, 0, 4, "warp-tiles (read only)") MIX(k_mix, 1, 6, 0, 4, "warp-tiles (write heavy)")
#define MIXNP(R, A, B, ROUNDS, HINT)                                                                   \
    {                                                                                                  \
        char nm[112];                                                                                  \
        snprintf(nm, sizeof nm, "traffic non-persistent r%d:a%d:b%d, %d tiles/warp, store hint %d", R, A, B, ROUNDS, HINT); \
        const unsigned grid = (unsigned)((n_tiles + 8 * ROUNDS - 1) / (8 * ROUNDS));                   \
        T.run(nm, 512.0 * (R + A + B) * n_tiles, [&] { k_mix_np<R, A, B, ROUNDS, HINT><<<grid, 256, 0, s>>>(tin, ta, tb, n_tiles); }); \
    }
    MIXNP(6, 4, 3, 1, 0) MIXNP(6, 4, 3, 2, 0) MIXNP(6, 4, 3, 4, 0) MIXNP(6, 4, 3, 16, 0) MIXNP(6, 4, 3, 64, 0)
    MIXNP(6, 4, 3, 1, 1) MIXNP(6, 4, 3, 1, 2) MIXNP(6, 4, 3, 16, 1)
    MIXNP(6, 0, 0, 1, 0) MIXNP(6, 0, 0, 4, 0) MIXNP(8, 3, 0, 1, 0) MIXNP(8, 3, 0, 1, 1) MIXNP(7, 6, 0, 1, 1)
    MIXNP(6, 6, 0, 1, 0) MIXNP(6, 6, 0, 4, 0) MIXNP(6, 6, 0, 16, 0) MIXNP(4, 4, 0, 1, 0) MIXNP(8, 8, 0, 1, 0)
    const uint64_t copy16 = n_tiles * 512 * 8 / 16 / 4096 * 4096;
    T.run("plain copy (LDG.128/STG.128 grid-stride x4, 8 CTAs/SM)", 2.0 * copy16 * 16,
          [&] { k_copy<<<sms * 8, 256, 0, s>>>(tin, ta, copy16); });
    T.run("torch-style copy (non-persistent, 128 thr x 4 x 16 B)", 2.0 * copy16 * 16,
          [&] { k_copy_np<<<(unsigned)(copy16 / 512), 128, 0, s>>>(tin, ta, copy16); });
    T.run("cudaMemsetAsync (write only)", 1.0 * copy16 * 16, [&] { cudaMemsetAsync(ta, 0, copy16 * 16, s); });
    T.run("cudaMemcpyAsync D2D", 2.0 * copy16 * 16, [&] { cudaMemcpyAsync(ta, tin, copy16 * 16, cudaMemcpyDeviceToDevice, s); });
    CK(cudaFree(tin)); CK(cudaFree(ta)); CK(cudaFree(tb));
    }

    // non-persistent K2
    unsigned long long *res_spread;
    CK(cudaMalloc(&res_spread, 256 * 64));
    auto run_np = [&](auto kern, const char *name, int tpw, int warps, int tail, int carve) {
        const uint64_t hi_bc = ~0ull << 32, hi_umi = ~0ull << 24;
        const unsigned grid = (unsigned)((n / 128 + warps * tpw - 1) / (warps * tpw));
        if (carve >= 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        unsigned long long *r = tail == 1 ? res_spread : res2;
        CK(cudaMemsetAsync(bc2, 0, n * 16, s)); CK(cudaMemsetAsync(umi2, 0, n * 12, s)); CK(cudaMemsetAsync(r, 0, tail == 1 ? 256 * 64 : 64, s));
        kern<<<grid, warps * 32, 0, s>>>(recs, n / 128, bc2, umi2, hi_bc, hi_umi, r);
        CK(cudaMemsetAsync(d_count, 0, 8, s));
        k_diff<<<sms * 8, 256, 0, s>>>((const uint4 *)bc2, (const uint4 *)bc, n, d_count);
        k_diff<<<sms * 8, 256, 0, s>>>((const uint4 *)umi2, (const uint4 *)umi, n * 12 / 16, d_count);
        unsigned long long bad = 0;
        std::vector<unsigned long long> got(tail == 1 ? 2048 : 8);
        CK(cudaMemcpyAsync(&bad, d_count, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(got.data(), r, got.size() * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        unsigned long long fold[8] = {0};
        for (size_t i = 0; i < got.size(); i++) { if (i % 8 == 4) fold[4] ^= got[i]; else fold[i % 8] += got[i]; }
        if (tail == 1) fold[0] = n;
        char extra[200];
        snprintf(extra, sizeof extra, ",\"grid\":%u,\"mismatch_u4\":%llu,\"result_block_ok\":%s", grid, bad,
                 tail == 2 ? "null" : memcmp(fold, ref_res, 64) == 0 ? "true" : "false");
        T.run(name, 52.0 * n, [&] {
            if (tail != 2) cudaMemsetAsync(r, 0, tail == 1 ? 256 * 64 : 64, s);
            kern<<<grid, warps * 32, 0, s>>>(recs, n / 128, bc2, umi2, hi_bc, hi_umi, r);
        }, extra);
    };
    run_np(k_unpack_np<1, 0, 8, 0>, "np W=8 tail=CTA-reduce, default carve-out", 1, 8, 0, -1);
    run_np(k_unpack_np<1, 0, 8, 0>, "np W=8 tail=CTA-reduce, carve-out 64", 1, 8, 0, 64);
    run_np(k_unpack_np<1, 0, 8, 1>, "np W=8 tail=per-warp RED spread, carve-out 64", 1, 8, 1, 64);
    run_np(k_unpack_np<1, 0, 8, 2>, "np W=8 no result, carve-out 64", 1, 8, 2, 64);
    run_np(k_unpack_np<1, 0, 4, 0>, "np W=4 tail=CTA-reduce, carve-out 64", 1, 4, 0, 64);
    run_np(k_unpack_np<1, 0, 4, 1>, "np W=4 tail=per-warp RED spread, carve-out 64", 1, 4, 1, 64);
    run_np(k_unpack_np<1, 0, 4, 0>, "np W=4 tail=CTA-reduce, carve-out 50", 1, 4, 0, 50);
    run_np(k_unpack_np<1, 0, 4, 0>, "np W=4 tail=CTA-reduce, carve-out 75", 1, 4, 0, 75);
    run_np(k_unpack_np<1, 0, 8, 1>, "np W=8 tail=per-warp RED spread, carve-out 75", 1, 8, 1, 75);
    run_np(k_unpack_np<2, 0, 8, 0>, "np W=8 TPW=2 tail=CTA-reduce, carve-out 64", 2, 8, 0, 64);

    // bulk-async K2 variants
    run_tma<2, 2, 8, 1>(T, sms, recs, n, bc2, umi2, res2, bc, umi, ref_res, d_count);
    run_tma<2, 1, 8, 1>(T, sms, recs, n, bc2, umi2, res2, bc, umi, ref_res, d_count);
    run_tma<2, 2, 4, 1>(T, sms, recs, n, bc2, umi2, res2, bc, umi, ref_res, d_count);
    run_tma<2, 2, 4, 2>(T, sms, recs, n, bc2, umi2, res2, bc, umi, ref_res, d_count);
    run_tma<2, 2, 2, 2>(T, sms, recs, n, bc2, umi2, res2, bc, umi, ref_res, d_count);
    T.run("shipped k_unpack<16,12> (C ABI), again", 52.0 * n, shipped);
    ibu_gpu_ctx_destroy(ctx);
    return 0;
}
