// group.cu — several GPUs of one box behind ONE handle: the multi-GPU form of the bulk record path.
//
// Records shard across the GPUs by contiguous range with the reference's own partition rule
// (src/io/mmap.rs:297-307, ranks in place of threads, ibu_shard_range).  One host thread per GPU
// drives that GPU's context; the staging copies of all of them run on the ONE process-wide worker
// pool (pipeline.cu), reading ONE mapping.  What is merged:
//   * the 8-word reduction results: on the host, exactly like the reference processors'
//     on_batch_complete merge (mmap.rs:365-372, examples/parallel.rs:28-35);
//   * the per-barcode table: n_records is additive by barcode, n_distinct_umi is not (the same
//     (barcode, umi) pair may occur in two shards), so the shards exchange their DE-DUPLICATED
//     pairs: every rank groups its (barcode, umi, multiplicity) rows by owner(barcode) =
//     splitmix64(barcode) % size and the owners pull their groups over NVLink
//     (cudaMemcpyPeerAsync), through pinned host memory, or with ncclSend / ncclRecv — the path's one
//     exchange step — count them weighted, and the owners' disjoint row sets are gathered on rank 0,
//     put in barcode order and returned in host memory.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <dlfcn.h>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "k4.h"

using namespace ibu;

namespace {

// ---- NCCL, loaded at run time (the library has no link-time dependency on it) ----
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok() const { return lib && CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv; }
};
constexpr int kNcclUint64 = 5;  // ncclDataType_t: ncclUint64

NcclApi load_nccl() {
    NcclApi a;
    // An NCCL that the process already holds (e.g. the one bundled with torch) is reused: loading a
    // second libnccl.so.2 next to it would make later loads resolve against the wrong one.
    const char *env = getenv("IBU_B200_NCCL_LIB");
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char *names[] = {env, "libnccl.so.2", "libnccl.so", "/usr/local/cuda/lib64/libnccl.so.2"};
    for (const char *name : names) {
        if (a.lib) break;
        if (name) a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!a.lib) return a;
    a.CommInitAll = (int (*)(ncclComm_t *, int, const int *))dlsym(a.lib, "ncclCommInitAll");
    a.CommDestroy = (int (*)(ncclComm_t))dlsym(a.lib, "ncclCommDestroy");
    a.GroupStart = (int (*)())dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
    a.Send = (int (*)(const void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(a.lib, "ncclSend");
    a.Recv = (int (*)(void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(a.lib, "ncclRecv");
    a.GetErrorString = (const char *(*)(int))dlsym(a.lib, "ncclGetErrorString");
    return a;
}

class Barrier {
public:
    explicit Barrier(unsigned n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lock(m_);
        const unsigned gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(lock, [&] { return gen_ != gen; });
        }
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    unsigned n_, count_ = 0, gen_ = 0;
};

double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

struct ibu_gpu_group {
    std::vector<ibu_gpu_ctx *> ctxs;
    std::vector<int> devices;
    bool distinct_devices = true;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    bool nccl_tried = false;
    std::mutex call_mutex;  // one group call at a time
    // pinned landing area of the merged rows (grow-only): a D2H into freshly malloc'ed pageable memory
    // is a staged copy through page faults (11 ms per 10^6 rows measured); from here the rows go to the
    // caller's allocation with the multi-threaded non-temporal copy
    void *h_land = nullptr;
    size_t h_land_bytes = 0;
};

namespace {

// device rows -> a malloc'ed host array (released with ibu_free), through the group's pinned landing area
int rows_to_host(ibu_gpu_group *g, const uint64_t *d_rows, uint64_t n_rows, cudaStream_t s, ibu_barcode_row_t **out,
                 ibu_error_t *err) {
    *out = nullptr;
    if (!n_rows) return IBU_OK;
    const size_t bytes = n_rows * sizeof(ibu_barcode_row_t);
    if (g->h_land_bytes < bytes) {
        if (g->h_land) cudaFreeHost(g->h_land);
        g->h_land = nullptr;
        g->h_land_bytes = 0;
        IBU_CUDA(cudaHostAlloc(&g->h_land, bytes + bytes / 4, cudaHostAllocPortable));
        g->h_land_bytes = bytes + bytes / 4;
    }
    ibu_barcode_row_t *h = (ibu_barcode_row_t *)malloc(bytes);
    if (!h) return set_error(err, IBU_ERR_NOMEM, 0, 0, 0, "out of memory");
    cudaError_t e = cudaMemcpyAsync(g->h_land, d_rows, bytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        free(h);
        return cuda_fail(err, e, "row gather");
    }
    ibu_host_stream_copy(h, g->h_land, bytes, 0);
    *out = h;
    return IBU_OK;
}

// What one rank contributes to / takes from a table merge.
struct RankState {
    // inputs
    const ibu_record_t *d_shard = nullptr;  // resident records (nullable when pairs are given)
    uint64_t n_shard = 0;
    uint64_t *pairs = nullptr;  // de-duplicated (barcode, umi, multiplicity) rows on this rank's device
    uint64_t n_pairs = 0;
    // exchange
    ibu_record_t *send = nullptr;       // pairs grouped by owner
    std::vector<uint64_t> send_counts;  // rows for each owner
    ibu_record_t *recv = nullptr;
    uint64_t n_recv = 0;
    void *h_stage = nullptr;  // pinned staging (host exchange)
    // owner's rows
    uint64_t *rows = nullptr;
    uint64_t n_rows = 0, n_row_pairs = 0;
    int rc = IBU_OK;
    ibu_error_t err{};
    double t_local = 0, t_exchange = 0, t_owner = 0, t_gather = 0;
};

int ensure_nccl(ibu_gpu_group *g, ibu_error_t *err) {
    if (!g->nccl_tried) {
        g->nccl_tried = true;
        if (!g->distinct_devices) return set_error(err, IBU_ERR_NCCL, 0, 0, 0, "NCCL needs one distinct GPU per rank");
        g->nccl = load_nccl();
        if (!g->nccl.ok()) return set_error(err, IBU_ERR_NCCL, 0, 0, 0, "NCCL error: libnccl.so.2 could not be loaded");
        g->comms.assign(g->devices.size(), nullptr);
        const int rc = g->nccl.CommInitAll(g->comms.data(), (int)g->devices.size(), g->devices.data());
        if (rc != 0) {
            g->comms.clear();
            return set_error(err, IBU_ERR_NCCL, rc, 0, 0, "NCCL error in ncclCommInitAll: %s",
                             g->nccl.GetErrorString ? g->nccl.GetErrorString(rc) : "?");
        }
    }
    if (g->comms.empty()) return set_error(err, IBU_ERR_NCCL, 0, 0, 0, "NCCL error: communicators are not available");
    return IBU_OK;
}

// The exact whole-job table from per-rank shards (resident records or their pair rows).
// Every rank's state is consumed: pairs / send / recv / rows are released here.
int group_table(ibu_gpu_group *g, std::vector<RankState> &st, const K4Hints &hints, int mode, uint32_t exchange,
                ibu_host_table_t *table, ibu_group_timing_t *timing, ibu_error_t *err) {
    const uint32_t G = (uint32_t)g->ctxs.size();
    if (exchange == IBU_EXCHANGE_AUTO) exchange = IBU_EXCHANGE_P2P;
    if (exchange == IBU_EXCHANGE_NCCL)
        if (int rc = ensure_nccl(g, err)) return rc;
    Barrier bar(G);
    std::atomic<bool> failed{false};
    std::vector<uint64_t> all_rows(G, 0), all_pairs(G, 0);
    uint64_t *d_gather = nullptr;  // on rank 0's device
    ibu_barcode_row_t *h_rows = nullptr;
    uint64_t total_rows = 0;
    const auto t_begin = std::chrono::steady_clock::now();

    auto worker = [&](uint32_t r) {
        RankState &me = st[r];
        ibu_gpu_ctx *ctx = g->ctxs[r];
        cudaSetDevice(ctx->device);
        cudaStream_t s = ctx->stream;
        ibu_error_t *e = &me.err;
        auto fail = [&](int rc) {
            if (rc != IBU_OK && me.rc == IBU_OK) me.rc = rc;
            if (rc != IBU_OK) failed = true;
        };
        auto cuda_ok = [&](cudaError_t ce, const char *what) {
            if (ce != cudaSuccess) fail(cuda_fail(e, ce, what));
            return ce == cudaSuccess;
        };
        auto t0 = std::chrono::steady_clock::now();

        // ---- 1. the shard's de-duplicated pairs (unless the ingest already produced them) ----
        if (!me.pairs && me.n_shard) {
            uint64_t np = 0;
            bool was_sorted = false;
            fail(k4_build_table(ctx, me.d_shard, me.n_shard, mode, hints, true, false, false, s, &me.pairs, &me.n_pairs, &np,
                                &was_sorted, e));
        }
        me.t_local = ms_since(t0);
        t0 = std::chrono::steady_clock::now();

        // ---- 2. group the rows by owner ----
        me.send_counts.assign(G, 0);
        if (!failed && me.n_pairs) {
            if (cuda_ok(cudaMallocAsync((void **)&me.send, me.n_pairs * 24, s), "cudaMallocAsync"))
                fail(ibu_gpu_partition_by_owner(ctx, reinterpret_cast<const ibu_record_t *>(me.pairs), me.n_pairs, G, me.send,
                                                me.send_counts.data(), s, e));
        }
        if (me.pairs) {
            cudaFreeAsync(me.pairs, s);
            me.pairs = nullptr;
        }
        if (exchange == IBU_EXCHANGE_HOST && !failed && me.n_pairs) {  // my whole send buffer, staged in pinned memory
            if (cuda_ok(cudaHostAlloc(&me.h_stage, me.n_pairs * 24, cudaHostAllocPortable), "cudaHostAlloc"))
                cuda_ok(cudaMemcpyAsync(me.h_stage, me.send, me.n_pairs * 24, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
        }
        cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
        bar.wait();  // every send buffer and its counts are complete

        // ---- 3. pull what I own ----
        std::vector<uint64_t> src_off(G, 0);  // where my group starts in each source's send buffer
        for (uint32_t q = 0; q < G; q++) {
            for (uint32_t o = 0; o < r; o++) src_off[q] += st[q].send_counts[o];
            me.n_recv += st[q].send_counts[r];
        }
        if (!failed && me.n_recv) {
            if (cuda_ok(cudaMallocAsync((void **)&me.recv, me.n_recv * 24, s), "cudaMallocAsync")) {
                if (exchange == IBU_EXCHANGE_NCCL) {
                    NcclApi &nc = g->nccl;
                    int nrc = nc.GroupStart();
                    uint64_t off = 0, my_off = 0;
                    for (uint32_t q = 0; q < G && nrc == 0; q++) {
                        const uint64_t out = me.send_counts[q], in = st[q].send_counts[r];
                        if (out) nrc = nc.Send(me.send + my_off, out * 3, kNcclUint64, (int)q, g->comms[r], s);
                        if (in && nrc == 0) nrc = nc.Recv(me.recv + off, in * 3, kNcclUint64, (int)q, g->comms[r], s);
                        my_off += out;
                        off += in;
                    }
                    const int end_rc = nc.GroupEnd();
                    if (nrc == 0) nrc = end_rc;
                    if (nrc != 0)
                        fail(set_error(e, IBU_ERR_NCCL, nrc, 0, 0, "NCCL error in the pair exchange: %s",
                                       nc.GetErrorString ? nc.GetErrorString(nrc) : "?"));
                } else {
                    uint64_t off = 0;
                    for (uint32_t q = 0; q < G; q++) {
                        const uint64_t cnt = st[q].send_counts[r];
                        if (!cnt) continue;
                        cudaError_t ce;
                        if (exchange == IBU_EXCHANGE_HOST)
                            ce = cudaMemcpyAsync(me.recv + off, (const ibu_record_t *)st[q].h_stage + src_off[q], cnt * 24,
                                                 cudaMemcpyHostToDevice, s);
                        else if (g->ctxs[q]->device == ctx->device)
                            ce = cudaMemcpyAsync(me.recv + off, st[q].send + src_off[q], cnt * 24, cudaMemcpyDeviceToDevice, s);
                        else
                            ce = cudaMemcpyPeerAsync(me.recv + off, ctx->device, st[q].send + src_off[q], g->ctxs[q]->device,
                                                     cnt * 24, s);
                        if (!cuda_ok(ce, "pair exchange copy")) break;
                        off += cnt;
                    }
                }
            }
        } else if (!failed && exchange == IBU_EXCHANGE_NCCL) {  // nothing to receive, but peers may expect my sends
            NcclApi &nc = g->nccl;
            int nrc = nc.GroupStart();
            uint64_t my_off = 0;
            for (uint32_t q = 0; q < G && nrc == 0; q++) {
                if (me.send_counts[q]) nrc = nc.Send(me.send + my_off, me.send_counts[q] * 3, kNcclUint64, (int)q, g->comms[r], s);
                my_off += me.send_counts[q];
            }
            const int end_rc = nc.GroupEnd();
            if ((nrc ? nrc : end_rc) != 0) fail(set_error(e, IBU_ERR_NCCL, nrc ? nrc : end_rc, 0, 0, "NCCL error in the pair exchange"));
        }
        cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
        bar.wait();  // everybody has pulled: send buffers may go
        if (me.send) cudaFreeAsync(me.send, s);
        if (me.h_stage) cudaFreeHost(me.h_stage);
        me.send = nullptr;
        me.h_stage = nullptr;
        me.t_exchange = ms_since(t0);
        t0 = std::chrono::steady_clock::now();

        // ---- 4. the owner counts its barcodes (weighted: a row stands for `multiplicity` records) ----
        if (!failed && me.n_recv) {
            bool was_sorted = false;
            fail(k4_build_table(ctx, me.recv, me.n_recv, 2, hints, false, false, true, s, &me.rows, &me.n_rows, &me.n_row_pairs,
                                &was_sorted, e));
        }
        if (me.recv) cudaFreeAsync(me.recv, s);
        me.recv = nullptr;
        all_rows[r] = me.n_rows;
        all_pairs[r] = me.n_row_pairs;
        me.t_owner = ms_since(t0);
        t0 = std::chrono::steady_clock::now();
        bar.wait();

        // ---- 5. gather on rank 0, barcode order, host ----
        uint64_t my_off = 0;
        for (uint32_t q = 0; q < r; q++) my_off += all_rows[q];
        if (r == 0) {
            for (uint32_t q = 0; q < G; q++) total_rows += all_rows[q];
            if (!failed && total_rows) {
                cuda_ok(cudaMallocAsync((void **)&d_gather, total_rows * 24, s), "cudaMallocAsync");
                cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
            }
        }
        bar.wait();
        if (!failed && me.n_rows) {
            cudaError_t ce = g->ctxs[0]->device == ctx->device
                                 ? cudaMemcpyAsync(d_gather + 3 * my_off, me.rows, me.n_rows * 24, cudaMemcpyDeviceToDevice, s)
                                 : cudaMemcpyPeerAsync(d_gather + 3 * my_off, g->ctxs[0]->device, me.rows, ctx->device, me.n_rows * 24, s);
            cuda_ok(ce, "row gather copy");
        }
        if (me.rows) cudaFreeAsync(me.rows, s);
        me.rows = nullptr;
        cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
        bar.wait();
        if (r == 0 && !failed && total_rows) {
            uint64_t *d_sorted = nullptr;
            if (cuda_ok(cudaMallocAsync((void **)&d_sorted, total_rows * 24, s), "cudaMallocAsync")) {
                {   // the owners' barcodes are disjoint: ordering the rows by barcode finishes the table
                    std::lock_guard<std::mutex> lock(ctx->arena_mutex);
                    const uint64_t vary[3] = {~0ull, 0, 0};
                    static const int order[1] = {0};
                    fail(k4_sort_rows(ctx, d_gather, total_rows, vary, order, 1, s, d_sorted, e));
                }
                if (!failed) fail(rows_to_host(g, d_sorted, total_rows, s, &h_rows, e));
                cudaFreeAsync(d_sorted, s);
            }
        }
        if (r == 0 && d_gather) cudaFreeAsync(d_gather, s);
        me.t_gather = ms_since(t0);
    };

    std::vector<std::thread> threads;
    for (uint32_t r = 1; r < G; r++) threads.emplace_back(worker, r);
    worker(0);
    for (auto &t : threads) t.join();

    int rc = IBU_OK;
    for (uint32_t r = 0; r < G && rc == IBU_OK; r++)
        if (st[r].rc != IBU_OK) {
            rc = st[r].rc;
            if (err) *err = st[r].err;
        }
    if (rc != IBU_OK) {
        free(h_rows);
        return rc;
    }
    table->h_rows = h_rows;
    table->n_rows = total_rows;
    table->n_distinct_pairs = 0;
    for (uint32_t r = 0; r < G; r++) table->n_distinct_pairs += all_pairs[r];
    if (timing) {
        uint64_t sent_max = 0;
        for (uint32_t r = 0; r < G; r++) {
            timing->local_ms = std::max(timing->local_ms, st[r].t_local);
            timing->exchange_ms = std::max(timing->exchange_ms, st[r].t_exchange);
            timing->owner_ms = std::max(timing->owner_ms, st[r].t_owner);
            timing->gather_ms = std::max(timing->gather_ms, st[r].t_gather);
            uint64_t sent = 0;
            for (uint32_t o = 0; o < G; o++) {
                timing->pairs_local += st[r].send_counts[o];
                if (o != r) sent += st[r].send_counts[o] * 24;
            }
            sent_max = std::max(sent_max, sent);
        }
        timing->bytes_sent = sent_max;
        timing->exchange = exchange;
        timing->table_ms = ms_since(t_begin);
    }
    return IBU_OK;
}

void merge_result(ibu_reduce_result_t &t, const ibu_reduce_result_t &c) {
    // the on_batch_complete merge of the reference processors (mmap.rs:365-372): wrapping adds, xor
    t.n_records += c.n_records;
    t.sum_barcode += c.sum_barcode;
    t.sum_umi += c.sum_umi;
    t.sum_index += c.sum_index;
    t.xor_all ^= c.xor_all;
    t.n_bad_barcode += c.n_bad_barcode;
    t.n_bad_umi += c.n_bad_umi;
    t.n_bad_records += c.n_bad_records;
}

// process_parallel across the GPUs of the group: rank r ingests ibu_shard_range(n, r, size).
int group_process(ibu_gpu_group *g, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len, uint32_t umi_len,
                  uint64_t first_record, int fd, uint64_t file_off, const ibu_group_request_t *req,
                  ibu_reduce_result_t *h_result, ibu_error_t *err) {
    const uint32_t G = (uint32_t)g->ctxs.size();
    const bool want_table = (req->ops & IBU_OP_TABLE) != 0, keep = (req->ops & IBU_OP_KEEP) != 0;
    if (want_table && !req->table) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_TABLE needs request.table");
    if (keep && !req->d_records) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_KEEP needs request.d_records");
    if (req->ops & IBU_OP_UNPACK) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_UNPACK is a single-context operation");
    std::lock_guard<std::mutex> lock(g->call_mutex);
    memset(h_result, 0, sizeof(*h_result));
    if (want_table) memset(req->table, 0, sizeof(*req->table));
    if (req->timing) memset(req->timing, 0, sizeof(*req->timing));
    K4Hints hints = k4_hints_of(req->table_mode);
    if (!hints.bc_len || !hints.umi_len) {
        hints.bc_len = bc_len;
        hints.umi_len = umi_len;
    }
    const auto t_begin = std::chrono::steady_clock::now();
    if (G == 1 && want_table) {
        // one GPU: nothing to exchange — the context's own ingest -> table pass, rows to the host
        ibu_gpu_ctx *ctx = g->ctxs[0];
        DeviceGuard guard(ctx->device);
        ibu_barcode_table_t dt{};
        ibu_record_t *kept1 = nullptr;
        ibu_process_request_t one{};
        one.ops = IBU_OP_REDUCE | IBU_OP_TABLE | (keep ? IBU_OP_KEEP : 0u);
        one.table_mode = req->table_mode;
        one.table = &dt;
        one.d_records = &kept1;
        int rc = process_records_ops(ctx, h_records, n, bc_len, umi_len, first_record, &one, h_result, nullptr, nullptr, err, fd,
                                     file_off, nullptr);
        if (req->timing) req->timing->ingest_ms = ms_since(t_begin);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc == IBU_OK) rc = rows_to_host(g, reinterpret_cast<const uint64_t *>(dt.d_rows), dt.n_rows, ctx->stream, &req->table->h_rows, err);
        const uint64_t n_rows1 = dt.n_rows;
        if (dt.d_rows) ibu_gpu_table_free(ctx, &dt);
        if (rc == IBU_OK) {
            req->table->n_rows = n_rows1;
            req->table->n_distinct_pairs = dt.n_distinct_pairs;
            req->table->n_records = n;
            if (keep) req->d_records[0] = kept1;
            if (req->shard_records) req->shard_records[0] = n;
        } else if (kept1) {
            cudaFree(kept1);
        }
        if (req->timing) {
            req->timing->gather_ms = req->timing->table_ms = ms_since(t1);
            req->timing->pairs_local = dt.n_distinct_pairs;
            req->timing->exchange = req->exchange == IBU_EXCHANGE_AUTO ? IBU_EXCHANGE_P2P : req->exchange;
            req->timing->total_ms = ms_since(t_begin);
        }
        return rc;
    }
    std::vector<RankState> st(G);
    std::vector<ibu_reduce_result_t> results(G);
    std::vector<ibu_record_t *> kept(G, nullptr);
    std::vector<double> t_ingest(G, 0);
    auto ingest = [&](uint32_t r) {
        ibu_gpu_ctx *ctx = g->ctxs[r];
        cudaSetDevice(ctx->device);
        uint64_t s = 0, e = 0;
        ibu_shard_range(n, r, G, &s, &e);
        st[r].n_shard = e - s;
        ibu_process_request_t one{};
        one.ops = IBU_OP_REDUCE | (want_table ? IBU_OP_TABLE : 0u) | (keep ? IBU_OP_KEEP : 0u);
        one.table_mode = req->table_mode;
        one.d_records = &kept[r];
        OpsExtra extra;
        extra.pairs = true;
        extra.pair_rows = &st[r].pairs;
        extra.n_pair_rows = &st[r].n_pairs;
        const auto t0 = std::chrono::steady_clock::now();
        st[r].rc = process_records_ops(ctx, h_records + s, e - s, bc_len, umi_len, first_record + s, &one, &results[r], nullptr,
                                       nullptr, &st[r].err, fd, file_off + s * IBU_RECORD_SIZE, want_table ? &extra : nullptr);
        t_ingest[r] = ms_since(t0);
    };
    {
        std::vector<std::thread> threads;
        for (uint32_t r = 1; r < G; r++) threads.emplace_back(ingest, r);
        ingest(0);
        for (auto &t : threads) t.join();
    }
    int rc = IBU_OK;
    for (uint32_t r = 0; r < G; r++) {
        if (st[r].rc != IBU_OK && rc == IBU_OK) {
            rc = st[r].rc;
            if (err) *err = st[r].err;
        }
        merge_result(*h_result, results[r]);
        if (req->shard_records) req->shard_records[r] = st[r].n_shard;
    }
    if (req->timing)
        for (uint32_t r = 0; r < G; r++) req->timing->ingest_ms = std::max(req->timing->ingest_ms, t_ingest[r]);
    if (rc == IBU_OK && want_table) {
        for (uint32_t r = 0; r < G; r++) st[r].n_shard = 0;  // the pairs are there already
        rc = group_table(g, st, hints, 2, req->exchange, req->table, req->timing, err);
        if (rc == IBU_OK) req->table->n_records = n;
    }
    for (uint32_t r = 0; r < G; r++) {
        if (st[r].pairs) {  // (only after a failure)
            cudaSetDevice(g->ctxs[r]->device);
            cudaFreeAsync(st[r].pairs, g->ctxs[r]->stream);
        }
        if (rc == IBU_OK && keep) {
            req->d_records[r] = kept[r];
        } else if (kept[r]) {
            cudaSetDevice(g->ctxs[r]->device);
            cudaFree(kept[r]);
        }
    }
    if (req->timing) req->timing->total_ms = ms_since(t_begin);
    return rc;
}

}  // namespace

extern "C" {

int ibu_gpu_group_create(const int *devices, uint32_t n_devices, const ibu_gpu_config_t *cfg, ibu_gpu_group_t **out,
                         ibu_error_t *err) {
    clear_error(err);
    if (!out || !devices || n_devices == 0 || n_devices > 64) return set_error(err, IBU_ERR_ARG, 0, n_devices, 0, "bad device list");
    *out = nullptr;
    std::unique_ptr<ibu_gpu_group> g(new (std::nothrow) ibu_gpu_group);
    if (!g) return set_error(err, IBU_ERR_NOMEM, 0, 0, 0, "out of memory");
    ibu_gpu_config_t c{};
    if (cfg) c = *cfg;
    // the staging copies of every rank run on the one process-wide worker pool: a chunk is cut into
    // as many pieces as there are cores, so whichever rank is staging gets all of them
    if (!c.copy_threads) c.copy_threads = std::max(2u, std::min(32u, std::thread::hardware_concurrency()));
    for (uint32_t i = 0; i < n_devices; i++) {
        for (uint32_t j = 0; j < i; j++)
            if (devices[j] == devices[i]) g->distinct_devices = false;
        ibu_gpu_ctx *ctx = nullptr;
        if (int rc = ibu_gpu_ctx_create(devices[i], &c, &ctx, err)) {
            for (auto *x : g->ctxs) ibu_gpu_ctx_destroy(x);
            return rc;
        }
        g->ctxs.push_back(ctx);
        g->devices.push_back(devices[i]);
    }
    // direct loads / copies between the GPUs (NVLink); without it peer copies are staged by the driver
    for (uint32_t i = 0; i < n_devices; i++) {
        cudaSetDevice(devices[i]);
        for (uint32_t j = 0; j < n_devices; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess) cudaGetLastError();  // (already enabled is fine)
                // The exchange buffers come from the stream-ordered pool (cudaMallocAsync), which peer
                // access enabled above does NOT cover: without this grant a peer copy out of pool memory
                // is staged through the host (240 MB per rank took 9.5 ms = PCIe rate, not NVLink).
                cudaMemPool_t pool = nullptr;
                if (cudaDeviceGetDefaultMemPool(&pool, devices[j]) == cudaSuccess && pool) {
                    cudaMemAccessDesc desc{};
                    desc.location.type = cudaMemLocationTypeDevice;
                    desc.location.id = devices[i];
                    desc.flags = cudaMemAccessFlagsProtReadWrite;
                    if (cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) cudaGetLastError();
                }
            }
        }
    }
    *out = g.release();
    return IBU_OK;
}

void ibu_gpu_group_destroy(ibu_gpu_group_t *g) {
    if (!g) return;
    for (auto c : g->comms)
        if (c && g->nccl.CommDestroy) g->nccl.CommDestroy(c);
    if (g->h_land) cudaFreeHost(g->h_land);
    for (auto *ctx : g->ctxs) ibu_gpu_ctx_destroy(ctx);
    delete g;
}

uint32_t ibu_gpu_group_size(const ibu_gpu_group_t *g) { return g ? (uint32_t)g->ctxs.size() : 0; }

ibu_gpu_ctx_t *ibu_gpu_group_ctx(ibu_gpu_group_t *g, uint32_t rank) {
    return (g && rank < g->ctxs.size()) ? g->ctxs[rank] : nullptr;
}

int ibu_gpu_group_process_mmap(ibu_gpu_group_t *g, const ibu_mmap_reader_t *reader, uint64_t start, uint64_t end,
                               const ibu_group_request_t *req, ibu_reduce_result_t *h_result, ibu_error_t *err) {
    clear_error(err);
    if (!g || !reader || !req || !h_result) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (end == UINT64_MAX) end = reader->len;
    if (start > end || end > reader->len)
        return set_error(err, IBU_ERR_INVALID_INDEX, 0, end, reader->len,
                         "Invalid index (%llu) - Must be less than %zu", (unsigned long long)end, reader->len);
    const ibu_record_t *recs = (const ibu_record_t *)(ibu_mmap_base(reader) + IBU_HEADER_SIZE) + start;
    const char *mode = getenv("IBU_B200_STAGE");
    const int fd = (mode && !strcmp(mode, "mmap")) ? -1 : ibu_mmap_fd(reader);
    return group_process(g, recs, end - start, reader->header.bc_len, reader->header.umi_len, start, fd,
                         IBU_HEADER_SIZE + start * IBU_RECORD_SIZE, req, h_result, err);
}

int ibu_gpu_group_process_host(ibu_gpu_group_t *g, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len,
                               uint32_t umi_len, const ibu_group_request_t *req, ibu_reduce_result_t *h_result,
                               ibu_error_t *err) {
    clear_error(err);
    if (!g || !req || !h_result || (!h_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (bc_len < 1 || bc_len > 32 || umi_len < 1 || umi_len > 32)
        return set_error(err, IBU_ERR_ARG, 0, bc_len, umi_len, "bc_len and umi_len must be in 1..32");
    return group_process(g, h_records, n, bc_len, umi_len, 0, -1, 0, req, h_result, err);
}

int ibu_gpu_group_barcode_count(ibu_gpu_group_t *g, const ibu_record_t *const *d_shards, const uint64_t *shard_records,
                                int mode, uint32_t exchange, ibu_host_table_t *table, ibu_group_timing_t *timing,
                                ibu_error_t *err) {
    clear_error(err);
    if (!g || !d_shards || !shard_records || !table) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (mode & IBU_COUNT_WEIGHTED) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "weighted shards are not supported");
    std::lock_guard<std::mutex> lock(g->call_mutex);
    memset(table, 0, sizeof(*table));
    if (timing) memset(timing, 0, sizeof(*timing));
    const uint32_t G = (uint32_t)g->ctxs.size();
    std::vector<RankState> st(G);
    uint64_t total = 0;
    for (uint32_t r = 0; r < G; r++) {
        if (shard_records[r] && (!d_shards[r] || ((uintptr_t)d_shards[r] & 31u)))
            return set_error(err, IBU_ERR_ARG, 0, r, 0, "shard %u: device pointer must be non-null and 32-byte aligned", r);
        st[r].d_shard = d_shards[r];
        st[r].n_shard = shard_records[r];
        total += shard_records[r];
    }
    const auto t0 = std::chrono::steady_clock::now();
    int rc = group_table(g, st, k4_hints_of(mode), (mode & 7) == 2 ? 2 : 0, exchange, table, timing, err);
    if (rc == IBU_OK) table->n_records = total;
    if (timing) timing->total_ms = ms_since(t0);
    return rc;
}

}  // extern "C"
