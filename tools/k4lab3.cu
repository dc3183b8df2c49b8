// k4lab3.cu — the K4 partition path stage by stage, on the shipped kernels (k4_partition.cuh):
// the staged two-level partition (k_part1 + k_part2) checked against a naive one-level scatter
// (same keys in every bucket), then k_bucket_dedup2.  K4_PART_MINB / K4LAB_CARVEOUT sweep occupancy.
// (The first forms of the path — one-level k_scatter_keys, table-scanning k_bucket_dedup — were
// measured with this tool before they were removed: profiles/r2_k4lab3_*.jsonl.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I ibu_b200/csrc -o tools/k4lab3 tools/k4lab3.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "k4_partition.cuh"

using namespace ibu;
using namespace ibu::k4p;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__global__ void k_gen(uint64_t *recs, uint64_t n, uint64_t nb, uint64_t us, int pattern) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = mix64(i + 1);
        if (pattern) {
            recs[3 * i] = i % 1000000;
            recs[3 * i + 1] = (31 * i) % 1000000;
        } else {
            recs[3 * i] = mix64(r % nb) & 0xFFFFFFFFull;
            recs[3 * i + 1] = mix64((r >> 32) ^ 77) % us;
        }
        recs[3 * i + 2] = i;
    }
}

// naive one-level scatter: the reference layout for the check (and the cost of 8-byte scattered stores)
__global__ void k_naive_scatter(const uint64_t *recs, uint64_t n, uint32_t ub, uint32_t pb, uint32_t cap, uint32_t *cursors, uint64_t *keys) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = mix64((recs[3 * i] << ub) | recs[3 * i + 1]);
        const uint32_t b = (uint32_t)(k >> (64 - pb));
        const uint32_t pos = atomicAdd(cursors + b, 1u);
        if (pos < cap) keys[(uint64_t)b * cap + pos] = k;
    }
}

// per-bucket (count, sum, xor) of the keys: layout-independent fingerprint
__global__ void k_check(const uint32_t *cursors, const uint64_t *keys, uint32_t cap, uint32_t nb, uint64_t *out) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        const uint32_t c = min(cursors[b], cap);
        uint64_t s = 0, x = 0;
        for (uint32_t i = 0; i < c; i++) {
            const uint64_t k = keys[(uint64_t)b * cap + i];
            s += k;
            x ^= k * 0x9E3779B97F4A7C15ull;
        }
        out[3 * b] = c; out[3 * b + 1] = s; out[3 * b + 2] = x;
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    void start() { CK(cudaEventRecord(a)); }
    float stop() {
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
};

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    const int pattern = argc > 2 ? atoi(argv[2]) : 0;
    const uint64_t umi_space = argc > 3 ? strtoull(argv[3], 0, 10) : 20;
    const uint32_t pb = 17, pb1 = 8, pb2 = pb - pb1, bb = 32, ub = 24;
    const uint64_t P = 1ull << pb;
    const uint32_t cap = pattern ? 2496 : 1264;
    const uint64_t cap1 = ((n >> pb1) * (pattern ? 5 : 21) / (pattern ? 4 : 20) + 4096) & ~15ull;
    uint64_t *recs, *keysA, *keysB, *keys1, *wide, *chkA, *chkB, *slots;
    uint32_t *curA, *curB, *cur1;
    unsigned long long *ctr;
    CK(cudaMalloc(&recs, n * 24));
    CK(cudaMalloc(&keysA, P * cap * 8));
    CK(cudaMalloc(&keysB, P * cap * 8));
    CK(cudaMalloc(&keys1, (cap1 << pb1) * 8));
    CK(cudaMalloc(&wide, 1 << 20));
    CK(cudaMalloc(&chkA, P * 24));
    CK(cudaMalloc(&chkB, P * 24));
    CK(cudaMalloc(&curA, P * 4));
    CK(cudaMalloc(&curB, P * 4));
    CK(cudaMalloc(&cur1, 4 << pb1));
    CK(cudaMalloc(&ctr, kCtrWords * 8));
    const uint64_t t_slots = 1ull << (argc > 4 ? atoi(argv[4]) : 23);
    const int skip_table = argc > 5 ? atoi(argv[5]) : 0;
    const int dd_ctas = argc > 6 ? atoi(argv[6]) : 5;
    CK(cudaMalloc(&slots, t_slots * 16));
    k_gen<<<148 * 8, 256>>>(recs, n, 1000000, umi_space, pattern);
    CK(cudaDeviceSynchronize());
    Timer tm;
    CK(cudaFuncSetAttribute(k_part1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPartTile * 8));
    CK(cudaFuncSetAttribute(k_part2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPartTile * 8));
    CK(cudaFuncSetAttribute(k_bucket_dedup2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 18));
    if (getenv("K4LAB_CARVEOUT")) {
        CK(cudaFuncSetAttribute(k_part1<false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("K4LAB_CARVEOUT"))));
        CK(cudaFuncSetAttribute(k_part2<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("K4LAB_CARVEOUT"))));
    }
    float t_old = 1e9, t_p1 = 1e9, t_p2 = 1e9;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(curA, 0, P * 4));
        CK(cudaMemset(ctr, 0, kCtrWords * 8));
        tm.start();
        k_naive_scatter<<<148 * 16, 256>>>(recs, n, ub, pb, cap, curA, keysA);
        t_old = std::min(t_old, tm.stop());
        CK(cudaMemset(curB, 0, P * 4));
        CK(cudaMemset(cur1, 0, 4 << pb1));
        Part1Args a1{recs, n, bb, ub, pb1, cap1, cur1, keys1, nullptr, wide, 1000, ctr};
        tm.start();
        k_part1<false><<<(unsigned)((n + kPartTile - 1) / kPartTile), 256, kPartTile * 8>>>(a1);
        t_p1 = std::min(t_p1, tm.stop());
        Part2Args a2{cur1, keys1, nullptr, cap1, (uint32_t)((cap1 + kPartTile - 1) / kPartTile), pb1, pb2, cap, nullptr, curB, keysB, nullptr, ctr, (uint32_t)kFlagBucket};
        tm.start();
        k_part2<false, false><<<a2.tiles_per_bucket << pb1, 256, kPartTile * 8>>>(a2);
        t_p2 = std::min(t_p2, tm.stop());
    }
    CK(cudaGetLastError());
    unsigned long long h_ctr[kCtrWords];
    CK(cudaMemcpy(h_ctr, ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost));
    k_check<<<512, 256>>>(curA, keysA, cap, (uint32_t)P, chkA);
    k_check<<<512, 256>>>(curB, keysB, cap, (uint32_t)P, chkB);
    std::vector<uint64_t> ha(P * 3), hb(P * 3);
    CK(cudaMemcpy(ha.data(), chkA, P * 24, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), chkB, P * 24, cudaMemcpyDeviceToHost));
    uint64_t total = 0, maxc = 0;
    for (uint64_t b = 0; b < P; b++) { total += ha[3 * b]; maxc = std::max(maxc, ha[3 * b]); }
    const bool same = memcmp(ha.data(), hb.data(), P * 24) == 0;
    // dedup on both layouts
    float t_dd[2] = {1e9, 1e9};
    unsigned long long pairs[2] = {0, 0}, rows[2] = {0, 0};
    for (int which = 0; which < 2; which++)
        for (int it = 0; it < 3; it++) {
            CK(cudaMemset(slots, 0xFF, t_slots * 16));
            CK(cudaMemset(ctr, 0, kCtrWords * 8));
            DedupArgs d{curB, nullptr, keysB, nullptr, (uint32_t)P, cap, pb, ub, pattern ? 12u : 11u,
                        TableRef{slots, skip_table ? 0 : t_slots - 1, ctr, 1u}, nullptr, 0};
            const size_t smem = ((size_t)1 << d.s_bits) * 14;
            tm.start();
            k_bucket_dedup2<false, false><<<148 * dd_ctas, 256, smem>>>(d);
            t_dd[which] = std::min(t_dd[which], tm.stop());
            CK(cudaMemcpy(h_ctr, ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost));
            pairs[which] = h_ctr[kCtrPairs];
            rows[which] = h_ctr[kCtrClaimed];
        }
    printf("{\"records\": %llu, \"pattern\": %d, \"umi_space\": %llu, \"naive_scatter_ms\": %.3f, \"part1_ms\": %.3f, \"part2_ms\": %.3f, "
           "\"keys\": %llu, \"max_bucket\": %llu, \"same_buckets\": %s, \"flags\": %llu, \"dedup2_ms\": [%.3f, %.3f], "
           "\"pairs\": [%llu, %llu], \"rows\": [%llu, %llu], \"table_slots\": %llu, \"skip_table\": %d, \"dedup_ctas\": %d}\n",
           (unsigned long long)n, pattern, (unsigned long long)umi_space, t_old, t_p1, t_p2, (unsigned long long)total,
           (unsigned long long)maxc, same ? "true" : "false", h_ctr[kCtrFlags], t_dd[0], t_dd[1], pairs[0], pairs[1], rows[0], rows[1], (unsigned long long)t_slots, skip_table, dd_ctas);
    return 0;
}
