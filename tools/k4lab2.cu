// k4lab2.cu — what one SM gives the ranking / insertion steps of the K4 partition path.
//   T1  shared-memory atomicAdd on a 256-bin histogram (spread addresses)
//   T2  ballot-matched ranking (8 ballots) + one plain increment by the group's leader, warp-private bins
//   T3  hardware match.any ranking, warp-private bins
//   T4  64-bit open-addressing insert (LDS + CAS / ADD) of register-resident keys
//   T5  plain STS.64 + LDS.64 at random addresses (the floor of any staged scheme)
// Every thread works on values it computes in registers, so nothing but the SM is measured.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/k4lab2 tools/k4lab2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__host__ __device__ inline uint64_t mix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
__device__ inline uint32_t xs32(uint32_t x) {  // cheap generator: the measured loop should be the op, not the hash
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    return x;
}

constexpr int kIters = 2048;

template <int MODE, int BITS>
__global__ void __launch_bounds__(256) k_rank(unsigned long long *sink) {
    constexpr uint32_t BINS = 1u << BITS;
    __shared__ uint32_t bins[8 * BINS];
    __shared__ unsigned long long tab[4096];
    for (uint32_t i = threadIdx.x; i < 8 * BINS; i += 256) bins[i] = 0;
    for (uint32_t i = threadIdx.x; i < 4096; i += 256) tab[i] = ~0ull;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    uint32_t *mine = bins + warp * BINS;
    for (int it = 0; it < kIters; it++) {
        x = xs32(x);
        const uint32_t b = x >> (32 - BITS);
        if (MODE == 1) {
            acc += atomicAdd(bins + b, 1u);
        } else if (MODE == 2) {
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < BITS; k++) {
                const uint32_t m = __ballot_sync(0xffffffffu, (b >> k) & 1u);
                peers &= ((b >> k) & 1u) ? m : ~m;
            }
            const uint32_t leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) {
                old = mine[b];
                mine[b] = old + __popc(peers);
            }
            __syncwarp();
            acc += __shfl_sync(0xffffffffu, old, leader) + __popc(peers & ((1u << lane) - 1u));
        } else if (MODE == 3) {
            const uint32_t peers = __match_any_sync(0xffffffffu, b);
            const uint32_t leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) {
                old = mine[b];
                mine[b] = old + __popc(peers);
            }
            __syncwarp();
            acc += __shfl_sync(0xffffffffu, old, leader) + __popc(peers & ((1u << lane) - 1u));
        } else if (MODE == 4) {
            // distinct-heavy: 2/3 of the keys come from a small set (repeats), the table never fills
            const unsigned long long key = (x & 3u) ? (x >> 2) % 1200u : (x >> 2) % 100000u + 5000u;
            uint32_t slot = (uint32_t)(key * 2654435761u >> 20) & 4095u;
            if (it == kIters / 2) {  // keep the load bounded: clear once
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < 4096; i += 256) tab[i] = ~0ull;
                __syncthreads();
            }
            for (int probe = 0; probe < 64; probe++, slot = (slot + 1) & 4095u) {
                unsigned long long cur = *(volatile unsigned long long *)(tab + slot);
                if (cur != key) {
                    if (cur != ~0ull) continue;
                    cur = atomicCAS(tab + slot, ~0ull, key);
                    if (cur == ~0ull) { acc++; break; }
                    if (cur != key) continue;
                }
                atomicAdd(bins + (slot & (8 * BINS - 1)), 1u);
                break;
            }
        } else if (MODE == 5) {
            const uint32_t s = x >> 20;
            tab[s] = x;
            __syncwarp();
            acc += (uint32_t)tab[(s * 7u + lane) & 4095u];
        } else if (MODE == 6) {  // 8 ballots only (no shared memory): the ALU part of mode 2
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < BITS; k++) {
                const uint32_t m = __ballot_sync(0xffffffffu, (b >> k) & 1u);
                peers &= ((b >> k) & 1u) ? m : ~m;
            }
            acc += __popc(peers & ((1u << lane) - 1u));
        } else if (MODE == 7) {  // match.any only
            acc += __popc(__match_any_sync(0xffffffffu, b) & ((1u << lane) - 1u));
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE, int BITS>
void run(const char *name, int ctas_per_sm, unsigned long long *sink) {
    int sms;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk_khz;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e9;
    for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(a));
        k_rank<MODE, BITS><<<sms * ctas_per_sm, 256>>>(sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    const double ops_per_sm = (double)ctas_per_sm * 256 * kIters;
    const double ns_per_key = best * 1e6 / ops_per_sm;
    printf("{\"test\": \"%s\", \"bits\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"ns_per_key_per_sm\": %.4f, "
           "\"cyc_per_key_at_1.9GHz\": %.3f, \"ms_per_1e8_keys_chip\": %.3f}\n",
           name, BITS, ctas_per_sm, best, ns_per_key, ns_per_key * 1.9, ns_per_key * 1e8 / sms * 1e-6);
    fflush(stdout);
}

int main() {
    unsigned long long *sink;
    CK(cudaMalloc(&sink, 64));
    for (int c : {2, 4, 8}) {
        run<1, 8>("atoms_add_hist", c, sink);
        run<1, 9>("atoms_add_hist", c, sink);
        run<2, 8>("ballot_rank_private", c, sink);
        run<2, 9>("ballot_rank_private", c, sink);
        run<3, 8>("match_any_rank_private", c, sink);
        run<6, 8>("ballots_only", c, sink);
        run<7, 8>("match_any_only", c, sink);
        run<4, 8>("cas64_insert", c, sink);
        run<5, 8>("sts_lds_random", c, sink);
    }
    return 0;
}
