// k4lab.cu — what the memory system gives the scatter step of the K4 partition path.
// n uniformly hashed 8-byte keys are appended to P buckets; variants isolate the cost of the
// returning atomic, of the scattered 8-byte store and of a histogram with REDs only.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/k4lab tools/k4lab.cu
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

// mode 0: atomic (returning) + store      1: RED only (histogram)      2: returning atomic, no store
// mode 3: store only (position from the element number)  4: match_any-aggregated atomic + store
template <int MODE>
__global__ void __launch_bounds__(256) k_scatter(const uint64_t *__restrict__ recs, uint64_t n, uint32_t pb, uint32_t cap,
                                                 uint32_t *cursors, uint64_t *keys, unsigned long long *sink) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    uint64_t acc = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint64_t i = (t >> 5) * 128 + q * 32 + (t & 31);
        if (i >= n) continue;
        const uint64_t k = mix64(recs[3 * i] * 0x10000ull + recs[3 * i + 1]);
        const uint32_t b = (uint32_t)(k >> (64 - pb));
        if (MODE == 1) {
            atomicAdd(cursors + b, 1u);
        } else if (MODE == 3) {
            keys[(uint64_t)b * cap + (uint32_t)(i >> pb) % cap] = k;
        } else if (MODE == 4) {
            const uint32_t peers = __match_any_sync(__activemask(), b);
            const uint32_t leader = __ffs(peers) - 1, lane = threadIdx.x & 31u;
            uint32_t pos = 0;
            if (lane == leader) pos = atomicAdd(cursors + b, (uint32_t)__popc(peers));
            pos = __shfl_sync(peers, pos, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pos < cap) keys[(uint64_t)b * cap + pos] = k;
        } else {
            const uint32_t pos = atomicAdd(cursors + b, 1u);
            if (MODE == 0) {
                if (pos < cap) keys[(uint64_t)b * cap + pos] = k;
            } else {
                acc += pos;
            }
        }
    }
    if (MODE == 2 && acc == 0x123456789ull) *sink = acc;
}

// shared-memory table insert rate: every CTA folds `per` keys of its bucket into a 4096-slot table
__global__ void __launch_bounds__(256) k_dedup(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ cursors,
                                               uint32_t n_buckets, uint32_t cap, uint32_t pb, unsigned long long *out) {
    __shared__ unsigned long long tkey[4096];
    __shared__ uint32_t tcnt[4096];
    uint32_t fresh = 0;
    for (uint32_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < 4096; i += 256) { tkey[i] = ~0ull; tcnt[i] = 0; }
        __syncthreads();
        const uint32_t cnt = min(cursors[b], cap);
        for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
            const uint64_t k = keys[(uint64_t)b * cap + i];
            uint32_t slot = (uint32_t)(k >> (64 - pb - 12)) & 4095u;
            for (;; slot = (slot + 1) & 4095u) {
                unsigned long long cur = *(volatile unsigned long long *)(tkey + slot);
                if (cur != k) {
                    if (cur != ~0ull) continue;
                    cur = atomicCAS(tkey + slot, ~0ull, (unsigned long long)k);
                    if (cur == ~0ull) fresh++;
                    else if (cur != k) continue;
                }
                atomicAdd(tcnt + slot, 1u);
                break;
            }
        }
        __syncthreads();
    }
    if (fresh) atomicAdd(out, (unsigned long long)fresh);
}

__global__ void k_gen(uint64_t *recs, uint64_t n, uint64_t nb, uint64_t us) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = mix64(i + 1);
        recs[3 * i] = mix64(r % nb) & 0xFFFFFFFFull;
        recs[3 * i + 1] = mix64(r ^ 77) % us;
        recs[3 * i + 2] = i;
    }
}

template <int MODE>
float run(const uint64_t *recs, uint64_t n, uint32_t pb, uint32_t cap, uint32_t *cursors, uint64_t *keys, unsigned long long *sink) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e9;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemsetAsync(cursors, 0, ((size_t)4 << pb)));
        CK(cudaEventRecord(a));
        const uint64_t threads = (n + 3) / 4;
        k_scatter<MODE><<<(unsigned)((threads + 255) / 256), 256>>>(recs, n, pb, cap, cursors, keys, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it) best = ms < best ? ms : best;
    }
    return best;
}

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    uint64_t *recs, *keys;
    uint32_t *cursors;
    unsigned long long *sink;
    CK(cudaMalloc(&recs, n * 24));
    CK(cudaMalloc(&keys, n * 8 * 3));
    CK(cudaMalloc(&cursors, (size_t)4 << 22));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(sink, 0, 64));
    k_gen<<<148 * 8, 256>>>(recs, n, 1000000, 20);
    CK(cudaDeviceSynchronize());
    for (uint32_t pb : {10u, 12u, 15u, 17u, 19u, 21u}) {
        const uint32_t cap = (uint32_t)(((n >> pb) * 2 + 256 + 15) & ~15ull);
        const float t0 = run<0>(recs, n, pb, cap, cursors, keys, sink), t1 = run<1>(recs, n, pb, cap, cursors, keys, sink),
                    t2 = run<2>(recs, n, pb, cap, cursors, keys, sink), t3 = run<3>(recs, n, pb, cap, cursors, keys, sink),
                    t4 = run<4>(recs, n, pb, cap, cursors, keys, sink);
        float td = -1;
        if (pb >= 15 && cap <= 8192) {
            run<0>(recs, n, pb, cap, cursors, keys, sink);  // leaves the buckets filled
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            CK(cudaEventRecord(a));
            k_dedup<<<148 * 4, 256>>>(keys, cursors, 1u << pb, cap, pb, sink);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&td, a, b));
        }
        printf("{\"records\": %llu, \"log2_buckets\": %u, \"cap\": %u, \"atomic_store_ms\": %.3f, \"red_hist_ms\": %.3f, "
               "\"atomic_only_ms\": %.3f, \"store_only_ms\": %.3f, \"match_any_atomic_store_ms\": %.3f, \"smem_dedup_ms\": %.3f}\n",
               (unsigned long long)n, pb, cap, t0, t1, t2, t3, t4, td);
        fflush(stdout);
    }
    unsigned long long h;
    CK(cudaMemcpy(&h, sink, 8, cudaMemcpyDeviceToHost));
    fprintf(stderr, "distinct seen by the last dedup: %llu\n", h);
    return 0;
}
