// poolprobe.cu — what cudaMallocAsync costs for gigabyte blocks that were released before (the K4
// scratch pattern): host time of every allocation over a few rounds of the same sequence.
//   nvcc -O2 -o poolprobe poolprobe.cu && ./poolprobe [scale]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    const double scale = argc > 1 ? atof(argv[1]) : 1.0;
    const int other_stream_free = argc > 2 ? atoi(argv[2]) : 0;
    cudaSetDevice(0);
    cudaMemPool_t pool;
    cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaStream_t s, t;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&t, cudaStreamNonBlocking);
    const double gb[] = {0.004, 1.0, 0.3, 1.0, 0.8, 2.4};  // sample, l0 keys, wide, l1 keys, exact keys, rows (per 1e8 records)
    for (int round = 0; round < 5; round++) {
        std::vector<void *> p(6, nullptr);
        printf("round %d:", round);
        for (int i = 0; i < 6; i++) {
            const size_t bytes = (size_t)(gb[i] * scale * 1e9);
            const double t0 = now();
            cudaError_t e = cudaMallocAsync(&p[i], bytes, s);
            cudaMemsetAsync(p[i], 0, 4096, s);
            const double t1 = now();
            printf(" %.2fGB %.3fms%s", bytes / 1e9, t1 - t0, e ? "(!)" : "");
            if (i == 3) { cudaFreeAsync(p[1], s); p[1] = nullptr; }
        }
        cudaStreamSynchronize(s);
        for (int i = 0; i < 5; i++) if (p[i]) cudaFreeAsync(p[i], s);
        cudaFreeAsync(p[5], other_stream_free ? t : s);
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(t);
        unsigned long long reserved = 0, used = 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        printf("  | reserved %.2f GB used %.2f GB\n", reserved / 1e9, used / 1e9);
    }
    return 0;
}
