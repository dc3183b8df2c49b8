// roundtrip.cpp — the shape of the reference's examples/roundtrip.rs and examples/parallel.rs
// against the C ABI of libibu_b200 (include/ibu_b200.h), from compiled host code with no Python:
//
//   write N records of the example pattern (i % 10^6, 31 i % 10^6, i) through the Writer
//   (roundtrip.rs:24-50) -> reopen with the MmapReader -> checksum on the CPU the way roundtrip.rs
//   does (XOR of all fields, roundtrip.rs:84-87) -> the GPU counterpart of process_parallel
//   (ibu_gpu_process_mmap: count, field sums, XOR, validation) must agree -> load the file into HBM
//   (device path of load_to_vec), build the per-barcode record / distinct-UMI table, unpack to
//   ASCII and pack back, compare with the file.
//
// Build:  make -C examples          Run:  examples/roundtrip [records] [dir]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/ibu_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__ != IBU_OK) {                                                         \
            fprintf(stderr, "%s failed: %s (%s)\n", #call, ibu_strerror(rc__), err.msg); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 10000000ull;
    const std::string path = std::string(argc > 2 ? argv[2] : "/tmp") + "/ibu_b200_roundtrip.ibu";
    ibu_error_t err;
    printf("libibu_b200 %s — %llu records (%.2f GB)\n", ibu_version(), (unsigned long long)n, 24.0 * n / 1e9);

    // ---- write (Writer::new + write_batch + finish) ----
    ibu_header_t header;
    ibu_header_init(&header, 16, 12);
    CHECK(ibu_header_validate(&header, &err));
    double t0 = now();
    {
        ibu_writer_t *w = nullptr;
        CHECK(ibu_writer_open(path.c_str(), &header, &w, &err));
        std::vector<ibu_record_t> batch(1 << 20);
        for (uint64_t first = 0; first < n; first += batch.size()) {
            const size_t cnt = (size_t)std::min<uint64_t>(batch.size(), n - first);
            for (size_t k = 0; k < cnt; k++) {
                const uint64_t i = first + k;
                batch[k] = ibu_record_t{i % 1000000ull, (i * 31ull) % 1000000ull, i};
            }
            CHECK(ibu_writer_write_batch(w, batch.data(), cnt, &err));
        }
        CHECK(ibu_writer_finish(w, &err));
        ibu_writer_close(w);
    }
    printf("  write      %.3f s\n", now() - t0);

    // ---- read back on the CPU: mmap + XOR checksum ----
    ibu_mmap_reader_t *reader = nullptr;
    CHECK(ibu_mmap_open(path.c_str(), &reader, &err));
    if (ibu_mmap_len(reader) != n) { fprintf(stderr, "length mismatch\n"); return 1; }
    t0 = now();
    uint64_t checksum = 0;
    const ibu_record_t *all = nullptr;
    size_t n_all = 0;
    CHECK(ibu_mmap_slice(reader, 0, n, &all, &n_all, &err));
    for (uint64_t i = 0; i < n; i++) checksum ^= all[i].barcode ^ all[i].umi ^ all[i].index;
    printf("  cpu xor    %.3f s  checksum %016llx\n", now() - t0, (unsigned long long)checksum);

    // ---- GPU counterpart of process_parallel ----
    ibu_gpu_ctx_t *ctx = nullptr;
    CHECK(ibu_gpu_ctx_create(0, nullptr, &ctx, &err));
    ibu_reduce_result_t res;
    t0 = now();
    CHECK(ibu_gpu_process_mmap(ctx, reader, 0, UINT64_MAX, &res, nullptr, nullptr, &err));
    const double t_gpu = now() - t0;
    printf("  gpu ingest %.3f s  (%.1f GB/s)  n=%llu xor=%016llx bad=%llu\n", t_gpu, 24.0 * n / t_gpu / 1e9,
           (unsigned long long)res.n_records, (unsigned long long)res.xor_all, (unsigned long long)res.n_bad_records);
    if (res.n_records != n || res.xor_all != checksum || res.sum_index != (n * (n - 1) / 2)) {
        fprintf(stderr, "GPU result differs from the CPU checksum\n");
        return 1;
    }

    // ---- device path of load_to_vec + per-barcode table ----
    ibu_header_t h2;
    ibu_record_t *d_recs = nullptr;
    uint64_t n_dev = 0;
    CHECK(ibu_gpu_load_to_device(ctx, path.c_str(), 0, UINT64_MAX, &h2, &d_recs, &n_dev, &err));
    ibu_barcode_table_t table;
    t0 = now();
    CHECK(ibu_gpu_barcode_count(ctx, d_recs, n_dev, 0, &table, nullptr, &err));
    printf("  table      %.3f s  %llu barcodes, %llu distinct (barcode, umi) pairs, sorted input: %u\n", now() - t0,
           (unsigned long long)table.n_rows, (unsigned long long)table.n_distinct_pairs, table.input_was_sorted);
    const uint64_t want_rows = std::min<uint64_t>(n, 1000000ull);
    if (table.n_rows != want_rows || table.n_records != n || table.n_distinct_pairs != want_rows) {
        fprintf(stderr, "barcode table differs from the closed form\n");
        return 1;
    }
    ibu_gpu_table_free(ctx, &table);
    ibu_gpu_free(ctx, d_recs);

    // ---- 2-bit unpack to ASCII and pack back (host buffers) ----
    const uint64_t m = std::min<uint64_t>(n, 4000000ull);
    std::vector<uint8_t> bc(m * 16), umi(m * 12);
    std::vector<ibu_record_t> back(m);
    std::vector<uint64_t> index(m);
    for (uint64_t i = 0; i < m; i++) index[i] = all[i].index;
    ibu_reduce_result_t r2, r3;
    CHECK(ibu_gpu_unpack_host(ctx, all, m, 16, 12, bc.data(), umi.data(), nullptr, &r2, &err));
    CHECK(ibu_gpu_pack_host(ctx, bc.data(), umi.data(), index.data(), 0, m, 16, 12, back.data(), nullptr, &r3, &err));
    if (memcmp(back.data(), all, m * sizeof(ibu_record_t)) != 0 || r3.n_bad_records != 0) {
        fprintf(stderr, "unpack -> pack does not reproduce the records\n");
        return 1;
    }
    printf("  codec      record 1 = %.16s / %.12s; %llu records unpacked and packed back identically\n",
           (const char *)bc.data() + 16, (const char *)umi.data() + 12, (unsigned long long)m);

    ibu_gpu_ctx_destroy(ctx);
    ibu_mmap_close(reader);
    remove(path.c_str());
    printf("roundtrip ok (%llu kernel launches)\n", (unsigned long long)ibu_gpu_launch_count());
    return 0;
}
