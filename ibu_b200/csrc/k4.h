// k4.h — internal interfaces between the K4 (per-barcode table) translation units.
//   barcode_count.cu  streaming sorted path (k_segments), radix sort, legacy global-hash path, C ABI
//   barcode_agg.cu    partition-then-aggregate path for unsorted inputs
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ctx.h"

namespace ibu {

// Stream-ordered scratch from the device's memory pool (cudaMallocAsync; the context keeps freed
// blocks cached, so after the first call an allocation costs about a microsecond).  Everything
// still listed is freed on scope exit, in stream order.
struct PoolScratch {
    cudaStream_t s;
    std::vector<void *> ptrs;
    explicit PoolScratch(cudaStream_t stream) : s(stream) {}
    PoolScratch(const PoolScratch &) = delete;
    PoolScratch &operator=(const PoolScratch &) = delete;
    ~PoolScratch() {
        for (void *p : ptrs)
            if (cudaFreeAsync(p, s) != cudaSuccess) cudaGetLastError();
    }
    template <class T>
    cudaError_t alloc(T **out, size_t bytes) {
        void *p = nullptr;
        static const bool trace = getenv("IBU_B200_TRACE_ALLOC") != nullptr;  // tuning: allocations that took long
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 256, s);
        if (trace) {
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms > 0.2) fprintf(stderr, "[ibu trace] cudaMallocAsync of %.3f GB took %.2f ms\n", bytes / 1e9, ms);
        }
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T *)p;
        return e;
    }
    void free_now(void *p) {  // release one block early (stream ordered)
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i] == p) {
                ptrs.erase(ptrs.begin() + i);
                if (cudaFreeAsync(p, s) != cudaSuccess) cudaGetLastError();
                return;
            }
    }
    void *keep(void *p) {  // ownership passes to the caller
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i] == p) {
                ptrs.erase(ptrs.begin() + i);
                break;
            }
        return p;
    }
};

// What the caller tells the table builder about the key layout (0 = find out from a sample).
struct K4Hints {
    uint32_t bc_len = 0, umi_len = 0;  // header lengths in bases
    int force_path = 0;                // 0 auto, 1 partition, 2 composite sort, 3 legacy (tests / tuning)
    bool sort_records = false;         // the job partitions (key, index) pairs for k4_sort_records_msd, not for a table
};
enum { kPathAuto = 0, kPathPartition = 1, kPathSort = 2, kPathLegacy = 3 };
K4Hints k4_hints_of(int mode);  // from the upper bits of `mode` / `flags` (IBU_COUNT_LENS, IBU_COUNT_PATH_*)

// ---- barcode_count.cu (all of these expect ctx->arena_mutex to be held by the caller) ----

// Legacy unsorted path: global hash aggregation or 16-byte pair radix sort, then the segment pass.
// *rows comes from cudaMallocAsync on `s` (3 u64 per row).
// d_est / r_est (0 = unknown): the sample's estimates of distinct pairs / barcodes — a near-distinct
// input skips the hash-aggregation attempt, and the segment pass starts with room for its rows.
int k4_legacy_unsorted(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, bool pair_mode,
                       bool weighted, uint64_t **rows, uint64_t *n_rows, uint64_t *n_pairs, ibu_error_t *err,
                       double d_est = 0, double r_est = 0);

// Stable LSD radix sort of n 3-word rows by the key words listed in key_order (least significant
// first), restricted to the bits set in vary[word]; the sorted rows are written to dst (n * 24 B).
int k4_sort_rows(ibu_gpu_ctx *ctx, const uint64_t *rows, uint64_t n, const uint64_t vary[3], const int *key_order,
                 int n_keys, cudaStream_t s, uint64_t *dst, ibu_error_t *err);

// ---- barcode_agg.cu ----

// What a look at ~256 Ki records at hashed positions shows (k_sample): enough to tell sorted
// from unsorted input, to lay out the (barcode, umi) key and to size the tables.
struct K4Sample {
    bool valid = false;
    uint64_t m = 0;          // records sampled
    double pairs = 0, barcodes = 0;          // distinct among them
    uint64_t unordered = 0;                  // sampled records whose successor is smaller: 0 for sorted input
    double pair_coll = 0;                    // sum over pairs of C(occurrences, 2)
    double pair_f1 = 0, pair_f2 = 0, bc_f1 = 0, bc_f2 = 0;  // seen exactly once / twice
    uint64_t bc_max = 0;                     // most sampled records under one barcode (1 if none repeats)
    uint32_t hist[130] = {};                 // [2][65] bit widths of the barcode / umi words
};
int k4_sample(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, K4Sample *out, ibu_error_t *err);
void k4_estimates(const K4Sample &smp, uint64_t n, double *d_est, double *r_est);  // Chao1, capped at n

// Partition-then-aggregate table of an unsorted input.  *handled = false (and nothing produced)
// when the input does not suit this path (keys wider than 64 bits, too few distinct keys, too
// many distinct barcodes, a capacity overflow): the caller then takes the legacy path.
int k4_partition_table(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, const K4Hints &hints, const K4Sample &smp,
                       bool pair_mode, bool pairs_sorted, bool weighted, cudaStream_t s, uint64_t **rows,
                       uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err);


// ibu_gpu_sort_records by partition (MSD) instead of LSD digit passes: the partition levels above on
// order-preserving (barcode, umi) keys that carry the record's index, then every final bucket sorted in
// shared memory and written as records at its final place (k4_ordered.cuh: k_bucket_sort_records).
// *handled = false (nothing written that matters) when the input does not suit it: words wider than the
// sample saw, keys beyond 64 bits, key ranges that hold far more records than a bucket (a few barcodes
// with most of the records) — the LSD sort then takes the input.  `out` must not alias `recs`.
int k4_sort_records_msd(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, const K4Sample &smp, uint64_t *out,
                        cudaStream_t s, bool *handled, ibu_error_t *err);

// The same in three steps, for an ingest pipeline that feeds the table chunk by chunk while the
// next chunk is still on the link (pipeline.cu).  begin leaves *job NULL when the input does not
// suit the path.  Caller holds ctx->arena_mutex from begin to destroy.
struct K4Job;
struct K4Chunking {
    uint64_t chunk_records = 0;  // records per piece of add() (0 = default); an ingest pipeline passes its chunk size
    uint32_t n_streams = 1;      // streams that will call add() concurrently
    bool force_exact = false;    // whole levels and an exact final layout (after an overflow of the uniform one)
};
int k4_job_begin(ibu_gpu_ctx *ctx, uint64_t n, const K4Hints &hints, const K4Sample &smp, bool pair_mode,
                 bool weighted, const K4Chunking &chunking, cudaStream_t s, K4Job **job, ibu_error_t *err);
bool k4_job_overflowed(const K4Job *job);  // a final bucket overflowed its uniform layout
cudaEvent_t k4_job_ready(K4Job *job);  // recorded on the job's stream once its scratch is initialised
int k4_job_add(K4Job *job, const uint64_t *recs, uint64_t cnt, cudaStream_t s, ibu_error_t *err);
int k4_job_finish(K4Job *job, const uint64_t *all_recs, bool pair_mode, bool pairs_sorted, uint64_t **rows,
                  uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err);
void k4_job_destroy(K4Job *job);

// barcode_count.cu: the whole decision tree (sorted streaming pass, partition path, legacy) over
// device-resident records; locks ctx->arena_mutex itself.
int k4_build_table(ibu_gpu_ctx *ctx, const ibu_record_t *d_records, uint64_t n, int mode, const K4Hints &hints,
                   bool pair_mode, bool pairs_sorted, bool weighted, cudaStream_t s, uint64_t **rows, uint64_t *n_rows,
                   uint64_t *n_pairs, bool *was_sorted, ibu_error_t *err);


// pipeline.cu: one pass over host records (or a file range, fd >= 0) with an operation mask; the
// engine behind ibu_gpu_process_{mmap,host}_ops.  `extra` (group mode): IBU_OP_TABLE produces the
// range's de-duplicated (barcode, umi, multiplicity) rows (unordered, cudaMallocAsync on the
// context's stream) instead of its table — what the ranks of a multi-GPU job exchange.
struct OpsExtra {
    bool pairs = false;
    uint64_t **pair_rows = nullptr;
    uint64_t *n_pair_rows = nullptr;
};
int process_records_ops(ibu_gpu_ctx *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len, uint32_t umi_len,
                        uint64_t first_record, const ibu_process_request_t *req, ibu_reduce_result_t *h_result,
                        ibu_chunk_cb on_chunk, void *user, ibu_error_t *err, int fd, uint64_t file_off,
                        const OpsExtra *extra);

}  // namespace ibu
