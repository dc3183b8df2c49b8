// probe_pin.cu — what the GPU box allows for ingest from a memory-mapped .ibu file.
//   (1) cudaHostRegister of file-backed mappings (tmpfs / disk) x protections x flags;
//   (2) H2D rate from whatever registered; (3) rate of the pageable -> pinned staging copy
//   (glibc memcpy, non-temporal memcpy, pread) while an H2D stream is running.
// Build: nvcc -O3 -std=c++17 -Xcompiler -mavx2 -o tools/probe_pin tools/probe_pin.cu
// Prints one JSON object per line.  A measurement tool, not part of the library.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void nt_copy(void *dst, const void *src, size_t bytes) {  // 32-byte aligned dst, multiple of 128
    const __m256i *s = (const __m256i *)src;
    __m256i *d = (__m256i *)dst;
    size_t n = bytes / 32;
    for (size_t i = 0; i + 4 <= n; i += 4) {
        __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1),
                c = _mm256_loadu_si256(s + i + 2), e = _mm256_loadu_si256(s + i + 3);
        _mm256_stream_si256(d + i, a);
        _mm256_stream_si256(d + i + 1, b);
        _mm256_stream_si256(d + i + 2, c);
        _mm256_stream_si256(d + i + 3, e);
    }
    _mm_sfence();
}

template <class F>
static void fan(unsigned threads, size_t bytes, F f) {
    size_t per = ((bytes + threads - 1) / threads + 4095) / 4096 * 4096;
    std::vector<std::thread> pool;
    for (unsigned i = 0; i < threads; i++) {
        size_t off = (size_t)i * per;
        if (off >= bytes) break;
        size_t len = std::min(per, bytes - off);
        pool.emplace_back([=] { f(off, len); });
    }
    for (auto &t : pool) t.join();
}

static bool make_file(const std::string &path, size_t bytes) {
    int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (fd < 0) return false;
    std::vector<uint8_t> buf(16 << 20);
    for (size_t i = 0; i < buf.size(); i++) buf[i] = (uint8_t)(i * 131 + 7);
    size_t done = 0;
    while (done < bytes) {
        size_t n = std::min(buf.size(), bytes - done);
        if (write(fd, buf.data(), n) != (ssize_t)n) { close(fd); return false; }
        done += n;
    }
    close(fd);
    return true;
}

static double h2d_rate(void *d, const void *h, size_t bytes, cudaStream_t s) {
    double best = 0;
    for (int r = 0; r < 3; r++) {
        cudaStreamSynchronize(s);
        double t0 = now();
        if (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return -1; }
        cudaStreamSynchronize(s);
        best = std::max(best, bytes / (now() - t0) / 1e9);
    }
    return best;
}

int main(int argc, char **argv) {
    const size_t bytes = (argc > 1 ? strtoull(argv[1], nullptr, 10) : 2048ull) << 20;  // MiB
    int dev = 0;
    cudaSetDevice(dev);
    cudaFree(0);
    int a_reg = 0, a_ro = 0, a_page = 0, a_pt = 0, a_hostptr = 0;
    cudaDeviceGetAttribute(&a_reg, cudaDevAttrHostRegisterSupported, dev);
    cudaDeviceGetAttribute(&a_ro, cudaDevAttrHostRegisterReadOnlySupported, dev);
    cudaDeviceGetAttribute(&a_page, cudaDevAttrPageableMemoryAccess, dev);
    cudaDeviceGetAttribute(&a_pt, cudaDevAttrPageableMemoryAccessUsesHostPageTables, dev);
    cudaDeviceGetAttribute(&a_hostptr, cudaDevAttrCanUseHostPointerForRegisteredMem, dev);
    printf("{\"probe\":\"attrs\",\"host_register\":%d,\"host_register_read_only\":%d,\"pageable_access\":%d,"
           "\"uses_host_page_tables\":%d,\"host_ptr_for_registered\":%d,\"hw_threads\":%u,\"page\":%ld}\n",
           a_reg, a_ro, a_page, a_pt, a_hostptr, std::thread::hardware_concurrency(), sysconf(_SC_PAGESIZE));
    fflush(stdout);

    void *d = nullptr;
    cudaMalloc(&d, bytes);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    void *pinned = nullptr;
    cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault);
    memset(pinned, 1, bytes);
    printf("{\"probe\":\"h2d cudaHostAlloc\",\"gbs\":%.2f}\n", h2d_rate(d, pinned, bytes, s));
    fflush(stdout);

    // ---- (4) fresh mappings: page-fault cost included; pread through a cache-resident bounce ----
    if (argc > 2) {
        std::string path = "/dev/shm/ibu_probe_pin2.bin";
        make_file(path, bytes + 32);
        const size_t chunk = 96ull << 20;
        const int kSlots = 3;
        void *hs[kSlots];
        cudaEvent_t ev[kSlots];
        for (int i = 0; i < kSlots; i++) {
            cudaHostAlloc(&hs[i], chunk, cudaHostAllocDefault);
            memset(hs[i], 0, chunk);
            cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        }
        const char *modes[] = {"nt_copy fresh map", "nt_copy fresh map+populate", "pread->256K bounce->nt", "pread->64K bounce->nt", "pread"};
        for (unsigned threads : {8u, 12u, 16u}) {
            for (int mode = 0; mode < 5; mode++) {
                double best = 0, t_map = 0;
                for (int rep = 0; rep < 2; rep++) {
                    int fd = open(path.c_str(), O_RDONLY);
                    double t0 = now();
                    void *m = mmap(nullptr, bytes + 32, PROT_READ, MAP_PRIVATE | (mode == 1 ? MAP_POPULATE : 0), fd, 0);
                    t_map = now() - t0;
                    cudaStreamSynchronize(s);
                    t0 = now();
                    size_t n_chunks = bytes / chunk;
                    for (size_t c = 0; c < n_chunks; c++) {
                        int sl = c % kSlots;
                        if (c >= (size_t)kSlots) cudaEventSynchronize(ev[sl]);
                        const uint8_t *src = (const uint8_t *)m + 32 + c * chunk;
                        uint8_t *dst = (uint8_t *)hs[sl];
                        if (mode <= 1) fan(threads, chunk, [=](size_t o, size_t l) { nt_copy(dst + o, src + o, l); });
                        if (mode == 2 || mode == 3)
                            fan(threads, chunk, [=](size_t o, size_t l) {
                                const size_t bb = mode == 2 ? (256u << 10) : (64u << 10);
                                uint8_t *bounce = (uint8_t *)aligned_alloc(4096, bb);
                                size_t done = 0;
                                while (done < l) {
                                    size_t want = std::min(bb, l - done);
                                    ssize_t k = pread(fd, bounce, want, 32 + c * chunk + o + done);
                                    if (k <= 0) break;
                                    size_t k128 = (size_t)k / 128 * 128;
                                    nt_copy(dst + o + done, bounce, k128);
                                    if (k128 < (size_t)k) memcpy(dst + o + done + k128, bounce + k128, k - k128);
                                    done += k;
                                }
                                free(bounce);
                            });
                        if (mode == 4)
                            fan(threads, chunk, [=](size_t o, size_t l) {
                                size_t got = 0;
                                while (got < l) {
                                    ssize_t k = pread(fd, dst + o + got, l - got, 32 + c * chunk + o + got);
                                    if (k <= 0) break;
                                    got += k;
                                }
                            });
                        cudaMemcpyAsync((uint8_t *)d + c * chunk, dst, chunk, cudaMemcpyHostToDevice, s);
                        cudaEventRecord(ev[sl], s);
                    }
                    cudaStreamSynchronize(s);
                    best = std::max(best, (bytes / chunk) * chunk / (now() - t0) / 1e9);
                    munmap(m, bytes + 32);
                    close(fd);
                }
                printf("{\"probe\":\"staging2\",\"mode\":\"%s\",\"threads\":%u,\"with_h2d\":1,\"gbs\":%.2f,\"mmap_s\":%.3f}\n",
                       modes[mode], threads, best, t_map);
                fflush(stdout);
            }
        }
        // ---- (5) does registration parallelise?  K threads register K disjoint slices of one mapping ----
        for (unsigned k : {1u, 2u, 4u, 8u}) {
            int fd = open(path.c_str(), O_RDONLY);
            void *m = mmap(nullptr, bytes + 32, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
            size_t per = bytes / k / 4096 * 4096;
            std::vector<int> ok(k, 0);
            double t0 = now();
            {
                std::vector<std::thread> pool;
                for (unsigned i = 0; i < k; i++)
                    pool.emplace_back([&, i] {
                        cudaSetDevice(0);
                        ok[i] = cudaHostRegister((uint8_t *)m + i * per, per, cudaHostRegisterReadOnly) == cudaSuccess;
                    });
                for (auto &t : pool) t.join();
            }
            double t_reg = now() - t0;
            int good = 0;
            for (unsigned i = 0; i < k; i++) good += ok[i];
            double gbs = good == (int)k ? h2d_rate(d, m, per * k, s) : -1;
            t0 = now();
            for (unsigned i = 0; i < k; i++) if (ok[i]) cudaHostUnregister((uint8_t *)m + i * per);
            double t_unreg = now() - t0;
            cudaGetLastError();
            printf("{\"probe\":\"parallel register\",\"threads\":%u,\"ok\":%d,\"register_s\":%.3f,\"register_gbs\":%.2f,"
                   "\"unregister_s\":%.3f,\"h2d_gbs\":%.2f}\n", k, good, t_reg, per * k / t_reg / 1e9, t_unreg, gbs);
            fflush(stdout);
            munmap(m, bytes + 32);
            close(fd);
        }
        // mprotect an existing read-only private mapping to RW, then register read-only
        {
            int fd = open(path.c_str(), O_RDONLY);
            void *m = mmap(nullptr, bytes + 32, PROT_READ, MAP_PRIVATE, fd, 0);
            int mp = mprotect(m, bytes + 32, PROT_READ | PROT_WRITE);
            double t0 = now();
            cudaError_t e = cudaHostRegister(m, bytes + 32, cudaHostRegisterReadOnly | cudaHostRegisterPortable);
            double t_reg = now() - t0;
            int mp2 = e == cudaSuccess ? mprotect(m, bytes + 32, PROT_READ) : -1;
            double gbs = e == cudaSuccess ? h2d_rate(d, (uint8_t *)m + 32, bytes, s) : -1;
            printf("{\"probe\":\"mprotect rw + register read_only + mprotect r\",\"mprotect\":%d,\"result\":\"%s\",\"register_s\":%.3f,"
                   "\"mprotect_back\":%d,\"h2d_gbs\":%.2f}\n", mp, cudaGetErrorName(e), t_reg, mp2, gbs);
            if (e == cudaSuccess) cudaHostUnregister(m); else cudaGetLastError();
            munmap(m, bytes + 32);
            close(fd);
        }
        for (int i = 0; i < kSlots; i++) { cudaFreeHost(hs[i]); cudaEventDestroy(ev[i]); }
        unlink(path.c_str());
        return 0;
    }
    // ---- (1)/(2) registration of file-backed mappings ----
    const char *dirs[] = {"/dev/shm", "/tmp"};
    struct Prot { const char *name; int oflag, prot, mflag; };
    const Prot prots[] = {{"r_private", O_RDONLY, PROT_READ, MAP_PRIVATE},
                          {"r_shared", O_RDONLY, PROT_READ, MAP_SHARED},
                          {"rw_shared", O_RDWR, PROT_READ | PROT_WRITE, MAP_SHARED},
                          {"rw_private", O_RDONLY, PROT_READ | PROT_WRITE, MAP_PRIVATE}};
    struct Flag { const char *name; unsigned f; };
    const Flag flags[] = {{"default", cudaHostRegisterDefault},
                          {"read_only", cudaHostRegisterReadOnly},
                          {"portable|read_only", cudaHostRegisterPortable | cudaHostRegisterReadOnly},
                          {"mapped|read_only", cudaHostRegisterMapped | cudaHostRegisterReadOnly}};
    for (const char *dir : dirs) {
        std::string path = std::string(dir) + "/ibu_probe_pin.bin";
        if (!make_file(path, bytes + 32)) {  // 32 + payload: not a page multiple, like a real .ibu file
            printf("{\"probe\":\"file\",\"dir\":\"%s\",\"error\":\"cannot create\"}\n", dir);
            continue;
        }
        for (const Prot &p : prots) {
            for (int populate = 0; populate < 2; populate++) {
                int fd = open(path.c_str(), p.oflag);
                if (fd < 0) continue;
                void *m = mmap(nullptr, bytes + 32, p.prot, p.mflag | (populate ? MAP_POPULATE : 0), fd, 0);
                if (m == MAP_FAILED) { close(fd); continue; }
                for (const Flag &f : flags) {
                    if ((p.prot & PROT_WRITE) && (p.mflag & MAP_PRIVATE) && populate == 0 && f.f != cudaHostRegisterDefault)
                        continue;
                    for (int round_len = 0; round_len < 2; round_len++) {
                        size_t len = round_len ? (bytes + 32 + 4095) / 4096 * 4096 : bytes + 32;
                        double t0 = now();
                        cudaError_t e = cudaHostRegister(m, len, f.f);
                        double t_reg = now() - t0;
                        double gbs = -1;
                        if (e == cudaSuccess) {
                            gbs = h2d_rate(d, (uint8_t *)m + 32, bytes, s);
                            cudaHostUnregister(m);
                        } else {
                            cudaGetLastError();
                        }
                        printf("{\"probe\":\"register\",\"dir\":\"%s\",\"map\":\"%s\",\"populate\":%d,\"flags\":\"%s\","
                               "\"len_page_rounded\":%d,\"result\":\"%s\",\"register_s\":%.3f,\"h2d_gbs\":%.2f}\n",
                               dir, p.name, populate, f.name, round_len, cudaGetErrorName(e), t_reg, gbs);
                        fflush(stdout);
                        if (e == cudaSuccess) break;
                    }
                }
                munmap(m, bytes + 32);
                close(fd);
            }
        }
        // ---- direct cudaMemcpyAsync from the pageable mapping (the driver stages) ----
        {
            int fd = open(path.c_str(), O_RDONLY);
            void *m = mmap(nullptr, bytes + 32, PROT_READ, MAP_PRIVATE, fd, 0);
            printf("{\"probe\":\"h2d from pageable mmap (driver staging)\",\"dir\":\"%s\",\"gbs\":%.2f}\n", dir,
                   h2d_rate(d, (uint8_t *)m + 32, bytes, s));
            fflush(stdout);
            // ---- (3) staging copies, pipelined with the H2D of the previous chunk ----
            const size_t chunk = 96ull << 20;
            const int kSlots = 3;
            void *hs[kSlots];
            cudaEvent_t ev[kSlots];
            for (int i = 0; i < kSlots; i++) {
                cudaHostAlloc(&hs[i], chunk, cudaHostAllocDefault);
                memset(hs[i], 0, chunk);
                cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            }
            const char *modes[] = {"memcpy", "nt_copy", "pread"};
            for (unsigned threads : {4u, 8u, 16u, 32u}) {
                if (threads > 2 * std::thread::hardware_concurrency()) continue;
                for (int mode = 0; mode < 3; mode++) {
                    for (int with_dma = 0; with_dma < 2; with_dma++) {
                        double best = 0;
                        for (int rep = 0; rep < 2; rep++) {
                            cudaStreamSynchronize(s);
                            double t0 = now();
                            size_t n_chunks = bytes / chunk;
                            for (size_t c = 0; c < n_chunks; c++) {
                                int sl = c % kSlots;
                                if (c >= (size_t)kSlots) cudaEventSynchronize(ev[sl]);
                                const uint8_t *src = (const uint8_t *)m + 32 + c * chunk;
                                uint8_t *dst = (uint8_t *)hs[sl];
                                if (mode == 0) fan(threads, chunk, [=](size_t o, size_t l) { memcpy(dst + o, src + o, l); });
                                if (mode == 1) fan(threads, chunk, [=](size_t o, size_t l) { nt_copy(dst + o, src + o, l); });
                                if (mode == 2)
                                    fan(threads, chunk, [=](size_t o, size_t l) {
                                        size_t got = 0;
                                        while (got < l) {
                                            ssize_t k = pread(fd, dst + o + got, l - got, 32 + c * chunk + o + got);
                                            if (k <= 0) break;
                                            got += k;
                                        }
                                    });
                                if (with_dma) {
                                    cudaMemcpyAsync((uint8_t *)d + c * chunk, dst, chunk, cudaMemcpyHostToDevice, s);
                                    cudaEventRecord(ev[sl], s);
                                }
                            }
                            cudaStreamSynchronize(s);
                            best = std::max(best, (bytes / chunk) * chunk / (now() - t0) / 1e9);
                        }
                        printf("{\"probe\":\"staging\",\"dir\":\"%s\",\"mode\":\"%s\",\"threads\":%u,\"with_h2d\":%d,\"gbs\":%.2f}\n",
                               dir, modes[mode], threads, with_dma, best);
                        fflush(stdout);
                    }
                }
            }
            for (int i = 0; i < kSlots; i++) { cudaFreeHost(hs[i]); cudaEventDestroy(ev[i]); }
            munmap(m, bytes + 32);
            close(fd);
        }
        unlink(path.c_str());
    }
    // ---- control: anonymous memory registered after the fact ----
    {
        void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        memset(m, 3, bytes);
        double t0 = now();
        cudaError_t e = cudaHostRegister(m, bytes, cudaHostRegisterDefault);
        double t_reg = now() - t0;
        double gbs = e == cudaSuccess ? h2d_rate(d, m, bytes, s) : -1;
        printf("{\"probe\":\"register anonymous\",\"result\":\"%s\",\"register_s\":%.3f,\"h2d_gbs\":%.2f}\n",
               cudaGetErrorName(e), t_reg, gbs);
        if (e == cudaSuccess) cudaHostUnregister(m);
        munmap(m, bytes);
    }
    return 0;
}
