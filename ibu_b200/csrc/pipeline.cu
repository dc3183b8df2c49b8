// pipeline.cu — GPU context, memory helpers and the host-buffer (end-to-end) paths:
// the device counterparts of MmapReader::process_parallel (src/io/mmap.rs:286-332) and
// load_to_vec (src/io/reader.rs:510-535), plus host->host unpack/pack through the GPU.
//
// Staging model: the record range is cut into chunks (a multiple of BATCH_SIZE, mmap.rs:284).
// Each chunk owns a slot = {stream, event, pinned staging, device buffers}.  For chunk c the
// host thread (1) waits for slot c % n_slots to drain and consumes its result, (2) brings the
// chunk's bytes into pinned memory if the source is pageable (parallel memcpy from the
// mmap / user buffer), (3) enqueues H2D -> kernel -> D2H on the slot's stream.  With >= 2 slots
// the copy engines, the SMs and the host memcpy of the next chunk all overlap.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <fcntl.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>

#include "ctx.h"
#include "k4.h"

using namespace ibu;

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }

bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Copy with non-temporal stores.  A staging buffer is written once and next read by the DMA
// engine: ordinary stores would first read each destination line for ownership, i.e. three
// DRAM transfers per byte instead of two, and the staging copy is bound by host DRAM
// bandwidth (B200 box, 16 threads, H2D running: memcpy / pread 38-39 GB/s, this 46-48 GB/s;
// profiles/r1_probe_pin*.jsonl).
void stream_copy(void *dst, const void *src, size_t bytes) {
#if defined(__SSE2__)
    uint8_t *d = (uint8_t *)dst;
    const uint8_t *s = (const uint8_t *)src;
    const size_t head = std::min(bytes, (size_t)(-(uintptr_t)d & 15u));
    if (head) {
        memcpy(d, s, head);
        d += head; s += head; bytes -= head;
    }
    const size_t n64 = bytes / 64;
    for (size_t i = 0; i < n64; i++, d += 64, s += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16)),
                      c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
        _mm_stream_si128((__m128i *)d, a);
        _mm_stream_si128((__m128i *)(d + 16), b);
        _mm_stream_si128((__m128i *)(d + 32), c);
        _mm_stream_si128((__m128i *)(d + 48), e);
    }
    _mm_sfence();
    if (bytes % 64) memcpy(d, s, bytes % 64);
#else
    memcpy(dst, src, bytes);
#endif
}

// Persistent host workers for the staging copies.  A chunk is staged in ~2 ms; creating and
// joining 16 threads per chunk cost a tenth of that.  One process-wide pool (hardware threads - 1
// workers, the caller works too); jobs from several contexts interleave safely.
class WorkPool {
public:
    static WorkPool &instance() {
        static WorkPool pool;
        return pool;
    }
    // fn(0) .. fn(pieces - 1), each exactly once, on the caller and the workers; returns when all are done
    template <class F>
    void run(unsigned pieces, F &&fn) {
        if (pieces == 0) return;
        if (pieces == 1 || workers_.empty()) {
            for (unsigned i = 0; i < pieces; i++) fn(i);
            return;
        }
        auto job = std::make_shared<Job>();
        job->fn = [&fn](unsigned i) { fn(i); };
        job->pieces = pieces;
        {
            std::lock_guard<std::mutex> lock(m_);
            jobs_.push_back(job);
        }
        cv_work_.notify_all();
        work_on(*job);
        std::unique_lock<std::mutex> lock(job->m);
        job->cv.wait(lock, [&] { return job->done == job->pieces; });
    }

private:
    struct Job {
        std::function<void(unsigned)> fn;
        unsigned pieces = 0;
        std::atomic<unsigned> next{0};
        unsigned done = 0;  // guarded by m
        std::mutex m;
        std::condition_variable cv;
    };
    static void work_on(Job &job) {
        unsigned mine = 0;
        for (unsigned i; (i = job.next.fetch_add(1, std::memory_order_relaxed)) < job.pieces;) {
            job.fn(i);
            mine++;
        }
        if (mine) {
            std::lock_guard<std::mutex> lock(job.m);
            job.done += mine;
            if (job.done == job.pieces) job.cv.notify_all();
        }
    }
    WorkPool() : owner_(getpid()) {
        // this process's share of the cores when a launcher runs one process per GPU (torchrun sets
        // LOCAL_WORLD_SIZE): eight ranks on a 32-thread host get 3 workers each, not 31 idle ones
        const char *e = getenv("LOCAL_WORLD_SIZE");
        const unsigned ranks = (unsigned)std::max(1, e ? atoi(e) : 1);
        unsigned hc = std::max(1u, std::thread::hardware_concurrency() / ranks);
        unsigned n = hc > 1 ? std::min(31u, hc - 1) : 0;
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
    }
    ~WorkPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        // (in a forked child the workers do not exist: run() then does every piece itself)
        for (auto &t : workers_) {
            if (getpid() == owner_) t.join(); else t.detach();
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_work_.wait(lock, [&] {
                    while (!jobs_.empty() && jobs_.front()->next.load(std::memory_order_relaxed) >= jobs_.front()->pieces)
                        jobs_.pop_front();  // fully handed out
                    return stop_ || !jobs_.empty();
                });
                if (stop_) return;
                job = jobs_.front();
            }
            work_on(*job);
        }
    }
    std::mutex m_;
    std::condition_variable cv_work_;
    std::deque<std::shared_ptr<Job>> jobs_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
    pid_t owner_;
};

// [0, bytes) cut into `threads` page-aligned pieces, piece i handed to fn(offset, length)
template <class F>
void for_pieces(size_t bytes, unsigned threads, size_t granule, F &&fn) {
    if (threads < 1) threads = 1;
    const size_t per = align_up((bytes + threads - 1) / threads, granule);
    const unsigned pieces = (unsigned)((bytes + per - 1) / per);
    WorkPool::instance().run(pieces, [&](unsigned i) {
        const size_t off = (size_t)i * per;
        fn(off, std::min(per, bytes - off));
    });
}

// How a pageable source is written into a pinned chunk slot.  Non-temporal stores (stream_copy) leave
// the chunk in DRAM for the DMA engine: 3 DRAM transfers per byte (source read, chunk write, DMA read).
// When the whole ring of slots is small enough to stay in the last-level cache, ordinary stores are
// better: the DMA engine reads the staged bytes from the cache and the next chunk overwrites them
// there (tools/stage_sweep.py on the 16-vCPU box, 60 MB L3, 10^8 records: 3 x 24 MB slots 51.7 GB/s
// plain against 49.4 non-temporal; 3 x 96 MB slots 45.6 against 48.8).  So: plain stores when this is
// the only context staging on the host and the ring is at most 96 MB.  IBU_B200_STAGE_PLAIN=0/1 forces.
std::atomic<int> g_live_contexts{0};
bool stage_plain(const ibu_gpu_ctx *ctx, size_t chunk_bytes) {
    if (const char *e = getenv("IBU_B200_STAGE_PLAIN")) return e[0] == '1';
    static const int ranks = [] {
        const char *e = getenv("LOCAL_WORLD_SIZE");
        return std::max(1, e ? atoi(e) : 1);
    }();
    return ranks == 1 && g_live_contexts.load(std::memory_order_relaxed) <= 1 &&
           chunk_bytes * ctx->slots.size() <= (96u << 20);
}

void parallel_memcpy(void *dst, const void *src, size_t bytes, unsigned threads, bool plain = false) {
    if (threads <= 1 || bytes < (1u << 20)) {
        if (plain) memcpy(dst, src, bytes); else stream_copy(dst, src, bytes);
        return;
    }
    for_pieces(bytes, threads, 4096, [&](size_t off, size_t len) {
        if (plain) memcpy((uint8_t *)dst + off, (const uint8_t *)src + off, len);
        else stream_copy((uint8_t *)dst + off, (const uint8_t *)src + off, len);
    });
}

// The same fan-out with pread(2): no page tables are populated for the mapping and an I/O error
// is a return code, not a SIGBUS.  Each thread reads 64 KB pieces into a cache-resident bounce
// buffer and streams them on (pread straight into the pinned buffer is the 38 GB/s case above).
// Returns false on failure with errno set to the failing worker's errno (EIO for an unexpected end of
// file, which sets none): errno is thread-local, so the worker's value is carried back explicitly.
bool parallel_pread(void *dst, int fd, uint64_t file_off, size_t bytes, unsigned threads, bool plain = false) {
    std::atomic<bool> ok{true};
    std::atomic<int> sys{0};
    for_pieces(bytes, threads, bytes >= (32u << 20) ? (1u << 20) : (64u << 10), [&](size_t off, size_t len) {
        constexpr size_t kBounce = 64u << 10;
        alignas(64) uint8_t bounce[kBounce];
        uint8_t *p = (uint8_t *)dst + off;
        uint64_t fo = file_off + off;
        while (len) {
            ssize_t got = ::pread(fd, plain ? p : bounce, std::min(len, plain ? len : kBounce), (off_t)fo);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) {
                sys = got < 0 ? errno : EIO;
                ok = false;
                return;
            }
            if (!plain) stream_copy(p, bounce, (size_t)got);
            p += got;
            fo += (uint64_t)got;
            len -= (size_t)got;
        }
    });
    if (!ok) errno = sys.load();
    return ok;
}

int ensure(void **p, size_t *have, size_t want, bool host, ibu_error_t *err) {
    if (*have >= want) return IBU_OK;
    if (*p) {
        if (host) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    cudaError_t e = host ? cudaHostAlloc(p, want, cudaHostAllocDefault) : cudaMalloc(p, want);
    if (e != cudaSuccess) return cuda_fail(err, e, host ? "cudaHostAlloc" : "cudaMalloc");
    *have = want;
    return IBU_OK;
}

void merge(ibu_reduce_result_t &t, const ibu_reduce_result_t &c) {
    // the on_batch_complete merge of the reference processors (mmap.rs:365-372,
    // examples/parallel.rs:28-35): wrapping adds; xor for the checksum
    t.n_records += c.n_records;
    t.sum_barcode += c.sum_barcode;
    t.sum_umi += c.sum_umi;
    t.sum_index += c.sum_index;
    t.xor_all ^= c.xor_all;
    t.n_bad_barcode += c.n_bad_barcode;
    t.n_bad_umi += c.n_bad_umi;
    t.n_bad_records += c.n_bad_records;
}

unsigned copy_threads(const ibu_gpu_ctx *ctx) {
    if (ctx->cfg.copy_threads) return ctx->cfg.copy_threads;
    // staging is a DRAM-bandwidth job: use the cores — this process's share of them when a launcher
    // says how many ranks run on the node (torchrun sets LOCAL_WORLD_SIZE)
    static const unsigned ranks = [] {
        const char *e = getenv("LOCAL_WORLD_SIZE");
        const int v = e ? atoi(e) : 1;
        return (unsigned)std::max(1, v);
    }();
    const unsigned hc = std::thread::hardware_concurrency();
    return std::max(2u, std::min(32u, hc / ranks));
}

// The context's chunk slots serve one host-buffer call at a time.  A chunk callback (on_chunk) runs
// while the lock is held: calling a host-buffer entry point of the same context from it is refused
// with IBU_ERR_ARG instead of deadlocking on the non-recursive mutex.
struct PipeOwner {
    ibu_gpu_ctx *ctx;
    explicit PipeOwner(ibu_gpu_ctx *c) : ctx(c) { ctx->pipe_owner.store(std::this_thread::get_id()); }
    ~PipeOwner() { ctx->pipe_owner.store(std::thread::id()); }
};
#define IBU_PIPE_LOCK(ctx)                                                                                        \
    if ((ctx)->pipe_owner.load() == std::this_thread::get_id())                                                   \
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "re-entered from a chunk callback: the context is in use");   \
    std::lock_guard<std::mutex> lock((ctx)->pipe_mutex);                                                          \
    PipeOwner pipe_owner_guard(ctx);                                                                              \
    if ((ctx)->open_stream) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "context busy: a streaming ingest is open on it")

// What one chunk moves: inputs (host -> device) and outputs (device -> host) as byte spans
// relative to the slot's device buffers.
struct Span {
    const void *h_src = nullptr;  // host source (inputs)
    void *h_dst = nullptr;        // host destination (outputs)
    size_t bytes = 0;
    size_t d_off = 0;             // offset in slot.d_in / slot.d_out
    bool pinned = false;          // host side is pinned: DMA straight from / to it
    int fd = -1;                  // inputs only: stage with pread(fd, file_off) instead of memcpy
    uint64_t file_off = 0;
    void *d_dst = nullptr;        // inputs only: device destination outside the slot (records that stay resident)
};

struct ChunkPlan {
    Span in[3];
    Span out[3];
    int n_in = 0, n_out = 0;
    size_t d_in_bytes = 0, d_out_bytes = 0;
    size_t h_in_bytes = 0;        // pinned staging for pageable inputs (0 = d_in_bytes)
};

// Drives chunks through the slots.  `enqueue(slot, chunk_idx)` launches the kernel(s) for a
// chunk whose inputs are already in flight on slot.stream; `consume(chunk_idx, result)` runs
// on the host, in chunk order, once the chunk's outputs have landed.
template <class PlanFn, class EnqueueFn, class ConsumeFn>
int run_chunks_impl(ibu_gpu_ctx *ctx, uint64_t n_chunks, PlanFn plan_of, EnqueueFn enqueue,
                    ConsumeFn consume, ibu_error_t *err) {
    const unsigned threads = copy_threads(ctx);
    const size_t n_slots = ctx->slots.size();
    std::vector<ChunkPlan> inflight(n_slots);
    std::vector<int64_t> owner(n_slots, -1);

    auto drain = [&](size_t s) -> int {
        if (owner[s] < 0) return IBU_OK;
        ibu_chunk_slot &slot = ctx->slots[s];
        IBU_CUDA(cudaEventSynchronize(slot.done));
        ChunkPlan &p = inflight[s];
        for (int k = 0; k < p.n_out; k++)
            if (!p.out[k].pinned && p.out[k].bytes)
                parallel_memcpy(p.out[k].h_dst, (uint8_t *)slot.h_out + p.out[k].d_off, p.out[k].bytes, threads);
        int rc = consume((uint64_t)owner[s], *slot.h_result);
        owner[s] = -1;
        if (rc) return set_error(err, IBU_ERR_PROCESS, rc, 0, 0, "Processing error: chunk callback returned %d", rc);
        return IBU_OK;
    };

    for (uint64_t c = 0; c < n_chunks; c++) {
        const size_t s = c % n_slots;
        if (int rc = drain(s)) return rc;
        ibu_chunk_slot &slot = ctx->slots[s];
        ChunkPlan p = plan_of(c);
        if (int rc = ensure(&slot.d_in, &slot.d_in_bytes, p.d_in_bytes, false, err)) return rc;
        if (int rc = ensure(&slot.d_out, &slot.d_out_bytes, p.d_out_bytes, false, err)) return rc;
        bool need_h_in = false, need_h_out = false;
        for (int k = 0; k < p.n_in; k++) need_h_in |= !p.in[k].pinned;
        for (int k = 0; k < p.n_out; k++) need_h_out |= !p.out[k].pinned;
        if (need_h_in)
            if (int rc = ensure(&slot.h_in, &slot.h_in_bytes, p.h_in_bytes ? p.h_in_bytes : p.d_in_bytes, true, err)) return rc;
        if (need_h_out)
            if (int rc = ensure(&slot.h_out, &slot.h_out_bytes, p.d_out_bytes, true, err)) return rc;
        for (int k = 0; k < p.n_in; k++) {
            const Span &sp = p.in[k];
            if (!sp.bytes) continue;
            const void *src = sp.h_src;
            if (!sp.pinned) {
                const bool plain = stage_plain(ctx, p.h_in_bytes ? p.h_in_bytes : p.d_in_bytes);
                if (sp.fd >= 0) {
                    if (!parallel_pread((uint8_t *)slot.h_in + sp.d_off, sp.fd, sp.file_off, sp.bytes, threads, plain))
                        return set_error(err, IBU_ERR_IO, errno, 0, 0, "I/O error: pread failed while staging");
                } else {
                    parallel_memcpy((uint8_t *)slot.h_in + sp.d_off, sp.h_src, sp.bytes, threads, plain);
                }
                src = (uint8_t *)slot.h_in + sp.d_off;
            }
            IBU_CUDA(cudaMemcpyAsync(sp.d_dst ? sp.d_dst : (void *)((uint8_t *)slot.d_in + sp.d_off), src, sp.bytes,
                                     cudaMemcpyHostToDevice, slot.stream));
        }
        if (int rc = enqueue(slot, c)) return rc;
        for (int k = 0; k < p.n_out; k++) {
            const Span &sp = p.out[k];
            if (!sp.bytes) continue;
            void *dst = sp.pinned ? sp.h_dst : (void *)((uint8_t *)slot.h_out + sp.d_off);
            IBU_CUDA(cudaMemcpyAsync(dst, (uint8_t *)slot.d_out + sp.d_off, sp.bytes,
                                     cudaMemcpyDeviceToHost, slot.stream));
        }
        IBU_CUDA(cudaMemcpyAsync(slot.h_result, slot.d_result, sizeof(ibu_reduce_result_t),
                                 cudaMemcpyDeviceToHost, slot.stream));
        IBU_CUDA(cudaEventRecord(slot.done, slot.stream));
        inflight[s] = p;
        owner[s] = (int64_t)c;
    }
    // drain in chunk order
    for (uint64_t c = n_chunks > n_slots ? n_chunks - n_slots : 0; c < n_chunks; c++)
        if (int rc = drain(c % n_slots)) return rc;
    return IBU_OK;
}

template <class PlanFn, class EnqueueFn, class ConsumeFn>
int run_chunks(ibu_gpu_ctx *ctx, uint64_t n_chunks, PlanFn plan_of, EnqueueFn enqueue,
               ConsumeFn consume, ibu_error_t *err) {
    int rc = run_chunks_impl(ctx, n_chunks, plan_of, enqueue, consume, err);
    if (rc != IBU_OK)  // nothing may still be reading or writing caller memory once we return
        for (auto &slot : ctx->slots)
            if (cudaStreamSynchronize(slot.stream) != cudaSuccess) cudaGetLastError();
    return rc;
}

uint64_t chunk_records(const ibu_gpu_ctx *ctx) {
    // default: BATCH_SIZE, the granularity at which the reference calls on_batch_complete (mmap.rs:284,
    // 322-326) — and three 24 MB slots are a ring the last-level cache can hold (stage_plain)
    return ctx->cfg.chunk_records ? ctx->cfg.chunk_records : (uint64_t)IBU_BATCH_SIZE;
}

int process_host_records(ibu_gpu_ctx *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len,
                         uint32_t umi_len, uint64_t first_record, ibu_reduce_result_t *h_result,
                         ibu_chunk_cb on_chunk, void *user, ibu_error_t *err, int fd = -1,
                         uint64_t file_off = 0) {
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    memset(h_result, 0, sizeof(*h_result));
    if (n == 0) return IBU_OK;  // an empty range never calls on_batch_complete (mmap.rs:502-519)
    const uint64_t chunk = chunk_records(ctx);
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const bool pinned = is_pinned(h_records);
    auto span = [&](uint64_t c, uint64_t &start, uint64_t &cnt) {
        start = c * chunk;
        cnt = std::min(chunk, n - start);
    };
    return run_chunks(
        ctx, n_chunks,
        [&](uint64_t c) {
            uint64_t start, cnt;
            span(c, start, cnt);
            ChunkPlan p;
            p.n_in = 1;
            p.in[0].h_src = h_records + start;
            p.in[0].bytes = cnt * IBU_RECORD_SIZE;
            p.in[0].pinned = pinned;
            p.in[0].fd = fd;
            p.in[0].file_off = file_off + start * IBU_RECORD_SIZE;
            p.d_in_bytes = align_up(chunk * IBU_RECORD_SIZE);
            p.d_out_bytes = kAlign;
            return p;
        },
        [&](ibu_chunk_slot &slot, uint64_t c) {
            uint64_t start, cnt;
            span(c, start, cnt);
            return ibu_gpu_validate_reduce_async(ctx, (const ibu_record_t *)slot.d_in, cnt, bc_len, umi_len,
                                                 slot.d_result, slot.stream, err);
        },
        [&](uint64_t c, const ibu_reduce_result_t &r) {
            uint64_t start, cnt;
            span(c, start, cnt);
            merge(*h_result, r);
            return on_chunk ? on_chunk(user, first_record + start, cnt, &r) : 0;
        },
        err);
}

}  // namespace

// Ingest with more than the reduction: the records land in ONE device allocation (chunk by chunk,
// same slots and streams), every chunk is validated / reduced (K1) or unpacked (K2, which carries
// K1's reductions) as it lands, and for IBU_OP_TABLE its (barcode, umi) keys are scattered into the
// table's buckets on the same stream — while the next chunk is still on the link.  After the last
// chunk only the bucket de-duplication and the rows remain.  This is process_parallel
// (mmap.rs:286-332) running the HashMap<barcode, count> processor of parallel.rs:79-98 (+ distinct
// UMIs) next to the count / sum processors, in one pass over the file.
int ibu::process_records_ops(ibu_gpu_ctx *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len, uint32_t umi_len,
                             uint64_t first_record, const ibu_process_request_t *req, ibu_reduce_result_t *h_result,
                             ibu_chunk_cb on_chunk, void *user, ibu_error_t *err, int fd, uint64_t file_off,
                             const OpsExtra *extra) {
    const bool pair_rows = extra && extra->pairs;  // group mode: the shard's de-duplicated pairs instead of its table
    const uint32_t ops = req->ops;
    const bool want_table = (ops & IBU_OP_TABLE) != 0, keep = (ops & IBU_OP_KEEP) != 0, unpack = (ops & IBU_OP_UNPACK) != 0;
    if (want_table && !req->table && !pair_rows) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_TABLE needs request.table");
    if (pair_rows) {
        *extra->pair_rows = nullptr;
        *extra->n_pair_rows = 0;
    }
    if (keep && !req->d_records) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_KEEP needs request.d_records");
    if (unpack && n && (!req->h_bc_ascii || !req->h_umi_ascii))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "IBU_OP_UNPACK needs request.h_bc_ascii and h_umi_ascii");
    if (want_table && req->table) memset(req->table, 0, sizeof(*req->table));
    if (keep) *req->d_records = nullptr;
    if (!want_table && !keep && !unpack)
        return process_host_records(ctx, h_records, n, bc_len, umi_len, first_record, h_result, on_chunk, user, err, fd, file_off);
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    memset(h_result, 0, sizeof(*h_result));
    if (want_table && req->table) {
        req->table->n_records = n;
        req->table->input_was_sorted = n == 0;
    }
    if (n == 0) return IBU_OK;
    const uint64_t chunk = chunk_records(ctx);
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const bool pinned = is_pinned(h_records);
    const bool resident = want_table || keep;
    // IBU_B200_TRACE=1: host-side phase times on stderr (tuning only)
    static const bool trace = getenv("IBU_B200_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ibu trace] ops: %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    // the resident copy comes from the stream-ordered pool: after the first call of a size it is
    // a cached block, not a fresh cudaMalloc (and cudaFree of GBs is a device-wide synchronisation)
    ibu_record_t *d_all = nullptr;
    if (resident) {
        IBU_CUDA(cudaMallocAsync((void **)&d_all, n * IBU_RECORD_SIZE, ctx->stream));
        IBU_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    mark("resident allocation");
    uint8_t *h_bc = req->h_bc_ascii, *h_umi = req->h_umi_ascii, *h_fl = req->h_flags;
    const bool pin_bc = unpack && is_pinned(h_bc), pin_umi = unpack && is_pinned(h_umi), pin_fl = unpack && h_fl && is_pinned(h_fl);
    const size_t off_umi = align_up(chunk * bc_len), off_fl = off_umi + align_up(chunk * umi_len);
    const size_t out_bytes = unpack ? off_fl + (h_fl ? align_up(chunk) : 0) : kAlign;

    K4Hints hints = k4_hints_of(req->table_mode);
    if (!hints.bc_len || !hints.umi_len) {
        hints.bc_len = bc_len;
        hints.umi_len = umi_len;
    }
    K4Job *job = nullptr;
    bool job_tried = false, saw_unordered = false;
    std::unique_lock<std::mutex> arena_lock(ctx->arena_mutex, std::defer_lock);
    auto span = [&](uint64_t c, uint64_t &start, uint64_t &cnt) {
        start = c * chunk;
        cnt = std::min(chunk, n - start);
    };
    // What the records look like decides how the table is built (and sizes its buckets); `d` holds
    // `cnt` records on the device once stream `s` gets there.
    auto decide = [&](const uint64_t *d, uint64_t cnt, cudaStream_t s) -> int {
        job_tried = true;
        arena_lock.lock();
        K4Sample smp;
        if (int r = k4_sample(ctx, d, cnt, s, &smp, err)) return r;
        saw_unordered = smp.unordered != 0;
        const int mode = req->table_mode & 7;
        // (sorted-looking input: the streaming pass over the resident records at the end is the fast path)
        if ((saw_unordered || mode == 2) && mode != 1 && hints.force_path != kPathLegacy && hints.force_path != kPathSort &&
            (n >= (1u << 16) || hints.force_path == kPathPartition)) {
            K4Chunking ch;
            ch.chunk_records = chunk;  // a staged chunk is one piece
            ch.n_streams = (uint32_t)ctx->slots.size();
            if (int r = k4_job_begin(ctx, n, hints, smp, pair_rows, false, ch, ctx->stream, &job, err)) return r;
        }
        return IBU_OK;
    };
    if (want_table && n_chunks > 1) {
        // The head of the range goes ahead of the pipeline (6 MB) and is sampled, so that no chunk has
        // to wait on the host for the decision (sampling the first chunk blocked the staging thread
        // for that chunk's whole H2D, ~2 ms).  Chunk 0 writes the same bytes again.
        const uint64_t head = std::min<uint64_t>(n, 1u << 18);
        IBU_CUDA(cudaMemcpyAsync(d_all, h_records, head * IBU_RECORD_SIZE, cudaMemcpyHostToDevice, ctx->stream));
        if (int r = decide(reinterpret_cast<const uint64_t *>(d_all), head, ctx->stream)) {
            if (arena_lock.owns_lock()) arena_lock.unlock();
            if (cudaFreeAsync(d_all, ctx->stream) != cudaSuccess) cudaGetLastError();
            return r;
        }
        mark("head sample");
    }
    int rc = run_chunks(
        ctx, n_chunks,
        [&](uint64_t c) {
            uint64_t start, cnt;
            span(c, start, cnt);
            ChunkPlan p;
            p.n_in = 1;
            p.in[0].h_src = h_records + start;
            p.in[0].bytes = cnt * IBU_RECORD_SIZE;
            p.in[0].pinned = pinned;
            p.in[0].fd = fd;
            p.in[0].file_off = file_off + start * IBU_RECORD_SIZE;
            p.in[0].d_dst = resident ? (void *)(d_all + start) : nullptr;
            p.h_in_bytes = align_up(chunk * IBU_RECORD_SIZE);
            p.d_in_bytes = resident ? kAlign : p.h_in_bytes;
            if (unpack) {
                p.out[0] = {nullptr, h_bc + start * bc_len, cnt * bc_len, 0, pin_bc};
                p.out[1] = {nullptr, h_umi + start * umi_len, cnt * umi_len, off_umi, pin_umi};
                p.n_out = 2;
                if (h_fl) p.out[p.n_out++] = {nullptr, h_fl + start, cnt, off_fl, pin_fl};
            }
            p.d_out_bytes = out_bytes;
            return p;
        },
        [&](ibu_chunk_slot &slot, uint64_t c) -> int {
            uint64_t start, cnt;
            span(c, start, cnt);
            const ibu_record_t *d_chunk = resident ? d_all + start : (const ibu_record_t *)slot.d_in;
            if (unpack) {
                uint8_t *d_out = (uint8_t *)slot.d_out;
                if (int r = ibu_gpu_unpack_async(ctx, d_chunk, cnt, bc_len, umi_len, d_out, d_out + off_umi,
                                                 h_fl ? d_out + off_fl : nullptr, slot.d_result, slot.stream, err))
                    return r;
            } else if (int r = ibu_gpu_validate_reduce_async(ctx, d_chunk, cnt, bc_len, umi_len, slot.d_result, slot.stream, err)) {
                return r;
            }
            if (!want_table) return IBU_OK;
            if (!job_tried)  // a single chunk: it is the sample
                if (int r = decide(reinterpret_cast<const uint64_t *>(d_chunk), cnt, slot.stream)) return r;
            if (job) {
                IBU_CUDA(cudaStreamWaitEvent(slot.stream, k4_job_ready(job), 0));
                return k4_job_add(job, reinterpret_cast<const uint64_t *>(d_chunk), cnt, slot.stream, err);
            }
            return IBU_OK;
        },
        [&](uint64_t c, const ibu_reduce_result_t &r) {
            uint64_t start, cnt;
            span(c, start, cnt);
            merge(*h_result, r);
            return on_chunk ? on_chunk(user, first_record + start, cnt, &r) : 0;
        },
        err);
    mark("chunks (ingest + per-chunk kernels)");
    if (rc == IBU_OK && want_table) {
        // every slot has drained (run_chunks waited for each chunk's event), so the context's stream
        // may read what the slot streams wrote
        uint64_t *rows = nullptr, n_rows = 0, n_pairs = 0;
        bool handled = false, was_sorted = false;
        if (job) rc = k4_job_finish(job, reinterpret_cast<const uint64_t *>(d_all), pair_rows, false, &rows, &n_rows, &n_pairs, &handled, err);
        if (job) k4_job_destroy(job);
        job = nullptr;
        if (arena_lock.owns_lock()) arena_lock.unlock();
        if (rc == IBU_OK && !handled) {
            int mode = req->table_mode & 7;
            if (mode == 0 && saw_unordered) mode = 2;  // already known not to be sorted
            rc = k4_build_table(ctx, d_all, n, mode, hints, pair_rows, false, false, ctx->stream, &rows, &n_rows, &n_pairs,
                                &was_sorted, err);
        }
        if (rc == IBU_OK && pair_rows) {
            *extra->pair_rows = rows;
            *extra->n_pair_rows = n_rows;
        } else if (rc == IBU_OK) {
            req->table->d_rows = reinterpret_cast<ibu_barcode_row_t *>(rows);
            req->table->n_rows = n_rows;
            req->table->n_distinct_pairs = n_pairs;
            req->table->input_was_sorted = was_sorted ? 1 : 0;
        }
    }
    if (job) k4_job_destroy(job);
    if (arena_lock.owns_lock()) arena_lock.unlock();
    mark("table finish");
    if (rc == IBU_OK && keep) {
        *req->d_records = d_all;
    } else if (d_all) {
        if (cudaFreeAsync(d_all, ctx->stream) != cudaSuccess) cudaGetLastError();
    }
    mark("release");
    return rc;
}

extern "C" {

int ibu_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ibu_gpu_ctx_create(int device, const ibu_gpu_config_t *cfg, ibu_gpu_ctx_t **out, ibu_error_t *err) {
    clear_error(err);
    if (!out) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    *out = nullptr;
    int count = 0;
    IBU_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return set_error(err, IBU_ERR_ARG, 0, device, count, "device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    IBU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)  // the only code in this library is sm_100a SASS: no fallback
        return set_error(err, IBU_ERR_CUDA, (int)cudaErrorNoKernelImageForDevice, prop.major, prop.minor,
                         "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                         prop.minor);
    DeviceGuard guard(device);
    auto *ctx = new (std::nothrow) ibu_gpu_ctx;
    if (!ctx) return set_error(err, IBU_ERR_NOMEM, 0, 0, 0, "out of memory");
    g_live_contexts.fetch_add(1, std::memory_order_relaxed);  // (ibu_gpu_ctx_destroy takes it back, also on the error path)
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cfg) ctx->cfg = *cfg;
    {   // keep freed blocks of the stream-ordered pool cached (result tables come from it)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    const uint32_t n_slots = ctx->cfg.n_slots ? ctx->cfg.n_slots : 3;
    ctx->slots.resize(n_slots);
    for (auto &s : ctx->slots) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_result, sizeof(ibu_reduce_result_t));
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&s.h_result, sizeof(ibu_reduce_result_t), cudaHostAllocDefault);
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->h_mail, kMailBytes, cudaHostAllocDefault);
    for (auto &r : ctx->result_ring) {
        const size_t bytes = (size_t)kResultBlocks * sizeof(ibu_reduce_result_t);
        if (e == cudaSuccess) e = cudaMalloc((void **)&r.blocks, bytes);
        if (e == cudaSuccess) e = cudaMemset(r.blocks, 0, bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.folded, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        ibu_gpu_ctx_destroy(ctx);
        return cuda_fail(err, e, "context creation");
    }
    *out = ctx;
    return IBU_OK;
}

__attribute__((visibility("hidden"))) void ibu_stream_detach(struct ibu_gpu_stream *st);  // (defined with the stream, below)

void ibu_gpu_ctx_destroy(ibu_gpu_ctx_t *ctx) {
    if (!ctx) return;
    g_live_contexts.fetch_sub(1, std::memory_order_relaxed);
    DeviceGuard guard(ctx->device);
    {   // a stream that is still open is detached: its later calls fail, its close only frees it
        std::lock_guard<std::mutex> lock(ctx->pipe_mutex);
        if (ctx->open_stream) ibu_stream_detach(ctx->open_stream);
        ctx->open_stream = nullptr;
    }
    for (auto &s : ctx->slots) {
        if (s.stream) {
            cudaStreamSynchronize(s.stream);
            cudaStreamDestroy(s.stream);
        }
        if (s.done) cudaEventDestroy(s.done);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_result) cudaFree(s.d_result);
        if (s.h_result) cudaFreeHost(s.h_result);
    }
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    for (auto &r : ctx->result_ring) {
        if (r.folded) {
            cudaEventSynchronize(r.folded);
            cudaEventDestroy(r.folded);
        }
        if (r.blocks) cudaFree(r.blocks);
    }
    if (ctx->arena_base) cudaFree(ctx->arena_base);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->rows_ev) cudaEventDestroy(ctx->rows_ev);
    delete ctx;
}

int ibu_gpu_ctx_device(const ibu_gpu_ctx_t *ctx) { return ctx ? ctx->device : -1; }
int ibu_gpu_ctx_sm_count(const ibu_gpu_ctx_t *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t ibu_gpu_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ibu_gpu_synchronize(ibu_gpu_ctx_t *ctx, void *stream, ibu_error_t *err) {
    clear_error(err);
    if (!ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null context");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaStreamSynchronize(pick_stream(ctx, stream)));
    return IBU_OK;
}

int ibu_gpu_malloc(ibu_gpu_ctx_t *ctx, size_t bytes, void **d_out, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !d_out) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaMalloc(d_out, bytes ? bytes : 1));
    return IBU_OK;
}

void ibu_gpu_free(ibu_gpu_ctx_t *ctx, void *d_ptr) {
    if (!ctx || !d_ptr) return;
    DeviceGuard guard(ctx->device);
    cudaFree(d_ptr);
}

int ibu_gpu_memcpy_h2d(ibu_gpu_ctx_t *ctx, void *d_dst, const void *h_src, size_t bytes, ibu_error_t *err) {
    clear_error(err);
    if (!ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null context");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    IBU_CUDA(cudaStreamSynchronize(ctx->stream));
    return IBU_OK;
}

int ibu_gpu_memcpy_d2h(ibu_gpu_ctx_t *ctx, void *h_dst, const void *d_src, size_t bytes, ibu_error_t *err) {
    clear_error(err);
    if (!ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null context");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    IBU_CUDA(cudaStreamSynchronize(ctx->stream));
    return IBU_OK;
}

int ibu_gpu_memcpy(ibu_gpu_ctx_t *ctx, void *dst, const void *src, size_t bytes, ibu_error_t *err) {
    clear_error(err);
    if (!ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null context");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    IBU_CUDA(cudaStreamSynchronize(ctx->stream));
    return IBU_OK;
}

int ibu_gpu_memset(ibu_gpu_ctx_t *ctx, void *d_dst, int value, size_t bytes, ibu_error_t *err) {
    clear_error(err);
    if (!ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null context");
    DeviceGuard guard(ctx->device);
    IBU_CUDA(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    IBU_CUDA(cudaStreamSynchronize(ctx->stream));
    return IBU_OK;
}

int ibu_host_alloc(size_t bytes, void **h_out, ibu_error_t *err) {
    clear_error(err);
    if (!h_out) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    IBU_CUDA(cudaHostAlloc(h_out, bytes ? bytes : 1, cudaHostAllocPortable));
    return IBU_OK;
}

void ibu_host_free(void *h_ptr) {
    if (h_ptr) cudaFreeHost(h_ptr);
}

int ibu_host_register(void *h_ptr, size_t bytes, int read_only, ibu_error_t *err) {
    clear_error(err);
    unsigned flags = cudaHostRegisterPortable | (read_only ? cudaHostRegisterReadOnly : 0u);
    IBU_CUDA(cudaHostRegister(h_ptr, bytes, flags));
    return IBU_OK;
}

void ibu_host_unregister(void *h_ptr) {
    if (h_ptr && cudaHostUnregister(h_ptr) != cudaSuccess) cudaGetLastError();
}

void ibu_host_stream_copy(void *dst, const void *src, size_t bytes, unsigned threads) {
    if (!bytes || !dst || !src) return;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    parallel_memcpy(dst, src, bytes, threads);
}

// Page-locks [p, p + bytes) of the mapping (page aligned), once per distinct range however many
// clones ask.  The driver refuses to page-lock a mapping without write permission
// (cudaErrorInvalidValue, read-only flag or not) but takes a private file mapping that has it, and
// with the read-only flag it pins the page-cache pages themselves, no copy-on-write
// (tools/probe_pin.cu).  The mapping is MAP_PRIVATE, so the permission never reaches the file; it
// is dropped again once the pages are locked.
static int pin_bytes(ibu_mmap_reader_t *reader, uint8_t *p, size_t bytes, ibu_error_t *err) {
    ibu_mmap_shared *sh = reader->shared;
    std::lock_guard<std::mutex> lock(sh->pin_mutex);
    for (auto &pr : sh->pinned) {
        if (pr.p == p && pr.bytes == bytes) {
            pr.count++;
            return IBU_OK;
        }
        if (p < pr.p + pr.bytes && pr.p < p + bytes)
            return set_error(err, IBU_ERR_ARG, 0, 0, 0, "range overlaps a differently pinned part of the mapping");
    }
    if (mprotect(p, bytes, PROT_READ | PROT_WRITE) != 0)
        return set_error(err, IBU_ERR_IO, errno, 0, 0, "I/O error: mprotect of the mapping failed");
    // fault the pages in on all cores first: registration itself walks them on one thread
    static const bool populate = getenv("IBU_B200_NO_POPULATE") == nullptr;
    if (populate && bytes >= (64u << 20)) {
        for_pieces(bytes, std::max(2u, std::thread::hardware_concurrency()), 2u << 20, [&](size_t off, size_t len) {
#ifdef MADV_POPULATE_READ
            if (madvise(p + off, len, MADV_POPULATE_READ) == 0) return;
#endif
            volatile uint8_t sink = 0;
            for (size_t i = 0; i < len; i += 4096) sink += p[off + i];
            (void)sink;
        });
    }
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    mprotect(p, bytes, PROT_READ);
    if (e != cudaSuccess) return cuda_fail(err, e, "cudaHostRegister(mmap)");
    sh->pinned.push_back({p, bytes, 1});
    return IBU_OK;
}

static void unpin_bytes(ibu_mmap_reader_t *reader, uint8_t *p, size_t bytes) {
    ibu_mmap_shared *sh = reader->shared;
    std::lock_guard<std::mutex> lock(sh->pin_mutex);
    for (size_t i = 0; i < sh->pinned.size(); i++) {
        auto &pr = sh->pinned[i];
        if (pr.p == p && (bytes == 0 || pr.bytes == bytes)) {
            if (--pr.count == 0) {
                if (cudaHostUnregister(pr.p) != cudaSuccess) cudaGetLastError();
                sh->pinned.erase(sh->pinned.begin() + i);
            }
            return;
        }
    }
}

// byte range of records [start, end) widened to whole pages
static void page_range(const ibu_mmap_reader_t *reader, uint64_t start, uint64_t end, uint8_t **p, size_t *bytes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const uint8_t *base = ibu_mmap_base(reader);
    size_t lo = IBU_HEADER_SIZE + start * IBU_RECORD_SIZE, hi = IBU_HEADER_SIZE + end * IBU_RECORD_SIZE;
    lo = lo / page * page;
    hi = std::min(ibu_mmap_bytes(reader), (hi + page - 1) / page * page);
    *p = const_cast<uint8_t *>(base) + lo;
    *bytes = hi - lo;
}

int ibu_mmap_pin(ibu_mmap_reader_t *reader, ibu_error_t *err) {
    clear_error(err);
    if (!reader) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null reader");
    return pin_bytes(reader, (uint8_t *)ibu_mmap_base(reader), ibu_mmap_bytes(reader), err);
}

void ibu_mmap_unpin(ibu_mmap_reader_t *reader) {
    if (reader) unpin_bytes(reader, (uint8_t *)ibu_mmap_base(reader), ibu_mmap_bytes(reader));
}

int ibu_mmap_pin_range(ibu_mmap_reader_t *reader, uint64_t start, uint64_t end, ibu_error_t *err) {
    clear_error(err);
    if (!reader) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null reader");
    if (start > end || end > reader->len)
        return set_error(err, IBU_ERR_INVALID_INDEX, 0, end, reader->len,
                         "Invalid index (%llu) - Must be less than %zu", (unsigned long long)end, reader->len);
    if (start == end) return IBU_OK;
    uint8_t *p;
    size_t bytes;
    page_range(reader, start, end, &p, &bytes);
    return pin_bytes(reader, p, bytes, err);
}

void ibu_mmap_unpin_range(ibu_mmap_reader_t *reader, uint64_t start, uint64_t end) {
    if (!reader || start >= end || end > reader->len) return;
    uint8_t *p;
    size_t bytes;
    page_range(reader, start, end, &p, &bytes);
    unpin_bytes(reader, p, bytes);
}

// ---- GPU counterpart of process_parallel -------------------------------------------------

int ibu_gpu_process_host(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len,
                         uint32_t umi_len, ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk,
                         void *user, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !h_result || (!h_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    return process_host_records(ctx, h_records, n, bc_len, umi_len, 0, h_result, on_chunk, user, err);
}

int ibu_gpu_process_mmap(ibu_gpu_ctx_t *ctx, const ibu_mmap_reader_t *reader, uint64_t start, uint64_t end,
                         ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk, void *user,
                         ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !reader || !h_result) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (end == UINT64_MAX) end = reader->len;
    if (start > end || end > reader->len)
        return set_error(err, IBU_ERR_INVALID_INDEX, 0, end, reader->len,
                         "Invalid index (%llu) - Must be less than %zu", (unsigned long long)end, reader->len);
    const ibu_record_t *recs =
        (const ibu_record_t *)(ibu_mmap_base(reader) + IBU_HEADER_SIZE) + start;
    // pageable source: stage with pread from the reader's descriptor (IBU_B200_STAGE=mmap forces
    // memcpy out of the mapping instead; a pinned mapping is DMA'd directly either way)
    const char *mode = getenv("IBU_B200_STAGE");
    const int fd = (mode && !strcmp(mode, "mmap")) ? -1 : ibu_mmap_fd(reader);
    return process_host_records(ctx, recs, end - start, reader->header.bc_len, reader->header.umi_len,
                                start, h_result, on_chunk, user, err, fd,
                                IBU_HEADER_SIZE + start * IBU_RECORD_SIZE);
}

int ibu_gpu_process_host_ops(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len,
                             uint32_t umi_len, const ibu_process_request_t *req, ibu_reduce_result_t *h_result,
                             ibu_chunk_cb on_chunk, void *user, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !req || !h_result || (!h_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (bc_len < 1 || bc_len > 32 || umi_len < 1 || umi_len > 32)
        return set_error(err, IBU_ERR_ARG, 0, bc_len, umi_len, "bc_len and umi_len must be in 1..32");
    return process_records_ops(ctx, h_records, n, bc_len, umi_len, 0, req, h_result, on_chunk, user, err, -1, 0, nullptr);
}

int ibu_gpu_process_mmap_ops(ibu_gpu_ctx_t *ctx, const ibu_mmap_reader_t *reader, uint64_t start, uint64_t end,
                             const ibu_process_request_t *req, ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk,
                             void *user, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !reader || !req || !h_result) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (end == UINT64_MAX) end = reader->len;
    if (start > end || end > reader->len)
        return set_error(err, IBU_ERR_INVALID_INDEX, 0, end, reader->len,
                         "Invalid index (%llu) - Must be less than %zu", (unsigned long long)end, reader->len);
    const ibu_record_t *recs = (const ibu_record_t *)(ibu_mmap_base(reader) + IBU_HEADER_SIZE) + start;
    const char *mode = getenv("IBU_B200_STAGE");
    const int fd = (mode && !strcmp(mode, "mmap")) ? -1 : ibu_mmap_fd(reader);
    return process_records_ops(ctx, recs, end - start, reader->header.bc_len, reader->header.umi_len, start, req,
                               h_result, on_chunk, user, err, fd, IBU_HEADER_SIZE + start * IBU_RECORD_SIZE, nullptr);
}

// ---- streaming ingest ----------------------------------------------------------------------

struct ibu_gpu_stream {
    ibu_gpu_ctx *ctx = nullptr;
    uint8_t header_bytes[IBU_HEADER_SIZE];
    size_t header_have = 0;
    ibu_header_t header{};
    uint64_t chunk_bytes = 0;
    size_t slot = 0, fill = 0;        // slot being filled, bytes already in its pinned buffer
    std::vector<char> busy;           // slot has a chunk in flight
    uint64_t records_submitted = 0;
    ibu_reduce_result_t total{};
    bool failed = false;
};

void ibu_stream_detach(ibu_gpu_stream *st) { st->ctx = nullptr; }

namespace {

int stream_drain(ibu_gpu_stream *st, size_t s, ibu_error_t *err) {
    if (!st->busy[s]) return IBU_OK;
    ibu_chunk_slot &slot = st->ctx->slots[s];
    IBU_CUDA(cudaEventSynchronize(slot.done));
    merge(st->total, *slot.h_result);
    st->busy[s] = 0;
    return IBU_OK;
}

// enqueue the records gathered in the current slot and move on to the next one
int stream_submit(ibu_gpu_stream *st, ibu_error_t *err) {
    ibu_gpu_ctx *ctx = st->ctx;
    ibu_chunk_slot &slot = ctx->slots[st->slot];
    const uint64_t n = st->fill / IBU_RECORD_SIZE;
    if (n) {
        if (int rc = ensure(&slot.d_in, &slot.d_in_bytes, align_up(st->chunk_bytes), false, err)) return rc;
        IBU_CUDA(cudaMemcpyAsync(slot.d_in, slot.h_in, n * IBU_RECORD_SIZE, cudaMemcpyHostToDevice, slot.stream));
        if (int rc = ibu_gpu_validate_reduce_async(ctx, (const ibu_record_t *)slot.d_in, n, st->header.bc_len,
                                                   st->header.umi_len, slot.d_result, slot.stream, err))
            return rc;
        IBU_CUDA(cudaMemcpyAsync(slot.h_result, slot.d_result, sizeof(ibu_reduce_result_t), cudaMemcpyDeviceToHost,
                                 slot.stream));
        IBU_CUDA(cudaEventRecord(slot.done, slot.stream));
        st->busy[st->slot] = 1;
        st->records_submitted += n;
    }
    // the bytes of an incomplete record move to the front of the next slot's buffer
    const size_t rest = st->fill - n * IBU_RECORD_SIZE;
    const size_t next = (st->slot + 1) % ctx->slots.size();
    if (int rc = stream_drain(st, next, err)) return rc;
    ibu_chunk_slot &ns = ctx->slots[next];
    if (int rc = ensure(&ns.h_in, &ns.h_in_bytes, align_up(st->chunk_bytes), true, err)) return rc;
    if (rest) memcpy(ns.h_in, (uint8_t *)slot.h_in + n * IBU_RECORD_SIZE, rest);
    st->slot = next;
    st->fill = rest;
    return IBU_OK;
}

}  // namespace

int ibu_gpu_stream_open(ibu_gpu_ctx_t *ctx, ibu_gpu_stream_t **out, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !out) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    *out = nullptr;
    std::unique_lock<std::mutex> lock(ctx->pipe_mutex, std::try_to_lock);
    if (!lock.owns_lock() || ctx->open_stream)
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "context busy: another stream or host-buffer call is active");
    auto *st = new (std::nothrow) ibu_gpu_stream;
    if (!st) return set_error(err, IBU_ERR_NOMEM, 0, 0, 0, "out of memory");
    st->ctx = ctx;
    st->chunk_bytes = chunk_records(ctx) * IBU_RECORD_SIZE;
    st->busy.assign(ctx->slots.size(), 0);
    DeviceGuard guard(ctx->device);
    ibu_chunk_slot &slot = ctx->slots[0];
    if (int rc = ensure(&slot.h_in, &slot.h_in_bytes, align_up(st->chunk_bytes), true, err)) {
        delete st;
        return rc;
    }
    ctx->open_stream = st;  // the slots are this stream's until close (or until the context goes)
    *out = st;
    return IBU_OK;
}

int ibu_gpu_stream_push(ibu_gpu_stream_t *st, const void *bytes, size_t len, ibu_error_t *err) {
    clear_error(err);
    if (!st || (!bytes && len)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (!st->ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "the stream's context has been destroyed");
    if (st->failed) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "stream already failed");
    DeviceGuard guard(st->ctx->device);
    const uint8_t *p = (const uint8_t *)bytes;
    if (st->header_have < IBU_HEADER_SIZE) {  // Reader::new: read_exact(32), then validate
        const size_t take = std::min(len, (size_t)IBU_HEADER_SIZE - st->header_have);
        memcpy(st->header_bytes + st->header_have, p, take);
        st->header_have += take;
        p += take;
        len -= take;
        if (st->header_have < IBU_HEADER_SIZE) return IBU_OK;
        memcpy(&st->header, st->header_bytes, IBU_HEADER_SIZE);
        if (int rc = ibu_header_validate(&st->header, err)) {
            st->failed = true;
            return rc;
        }
    }
    const unsigned threads = copy_threads(st->ctx);
    while (len) {
        ibu_chunk_slot &slot = st->ctx->slots[st->slot];
        const size_t take = std::min<size_t>(len, st->chunk_bytes - st->fill);
        parallel_memcpy((uint8_t *)slot.h_in + st->fill, p, take, threads);
        st->fill += take;
        p += take;
        len -= take;
        if (st->fill == st->chunk_bytes) {
            if (int rc = stream_submit(st, err)) {
                st->failed = true;
                return rc;
            }
        }
    }
    return IBU_OK;
}

int ibu_gpu_stream_header(const ibu_gpu_stream_t *st, ibu_header_t *header, ibu_error_t *err) {
    clear_error(err);
    if (!st || !header) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (st->header_have < IBU_HEADER_SIZE || st->failed)
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "no valid header yet");
    *header = st->header;
    return IBU_OK;
}

int ibu_gpu_stream_finish(ibu_gpu_stream_t *st, ibu_reduce_result_t *h_result, ibu_error_t *err) {
    clear_error(err);
    if (!st || !h_result) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    memset(h_result, 0, sizeof(*h_result));
    if (!st->ctx) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "the stream's context has been destroyed");
    if (st->failed) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "stream already failed");
    if (st->header_have < IBU_HEADER_SIZE)  // read_exact on a short stream: UnexpectedEof
        return set_error(err, IBU_ERR_IO, EIO, st->header_have, 0, "I/O error: stream ended inside the 32-byte header");
    DeviceGuard guard(st->ctx->device);
    const size_t rest = st->fill % IBU_RECORD_SIZE;
    int rc = stream_submit(st, err);
    for (size_t s = 0; s < st->busy.size() && rc == IBU_OK; s++) rc = stream_drain(st, s, err);
    st->failed = true;  // finished: no more pushes
    if (rc != IBU_OK) return rc;
    *h_result = st->total;
    if (rest) {
        const uint64_t pos = IBU_HEADER_SIZE + st->records_submitted * IBU_RECORD_SIZE;
        return set_error(err, IBU_ERR_TRUNCATED_RECORD, 0, pos, 0, "Truncated record at position %llu",
                         (unsigned long long)pos);
    }
    return IBU_OK;
}

void ibu_gpu_stream_close(ibu_gpu_stream_t *st) {
    if (!st) return;
    if (ibu_gpu_ctx *ctx = st->ctx) {  // (null: the context went first and detached us)
        DeviceGuard guard(ctx->device);
        for (auto &slot : ctx->slots)
            if (cudaStreamSynchronize(slot.stream) != cudaSuccess) cudaGetLastError();
        std::lock_guard<std::mutex> lock(ctx->pipe_mutex);
        if (ctx->open_stream == st) ctx->open_stream = nullptr;
    }
    delete st;
}

int ibu_gpu_table_to_host(ibu_gpu_ctx_t *ctx, const ibu_barcode_table_t *table, ibu_barcode_row_t *h_rows,
                          ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !table || (table->n_rows && (!table->d_rows || !h_rows))) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (!table->n_rows) return IBU_OK;
    DeviceGuard guard(ctx->device);
    const size_t bytes = (size_t)table->n_rows * sizeof(ibu_barcode_row_t);
    if (is_pinned(h_rows)) {
        IBU_CUDA(cudaMemcpyAsync(h_rows, table->d_rows, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        IBU_CUDA(cudaStreamSynchronize(ctx->stream));
        return IBU_OK;
    }
    std::lock_guard<std::mutex> lock(ctx->pipe_mutex);  // the landing area is slot 0's pinned output buffer
    if (ctx->open_stream) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "context busy: a streaming ingest is open on it");
    ibu_chunk_slot &slot = ctx->slots[0];
    const size_t piece_max = std::max<size_t>(slot.h_out_bytes, 32u << 20);
    if (int rc = ensure(&slot.h_out, &slot.h_out_bytes, std::min(align_up(bytes), piece_max), true, err)) return rc;
    const unsigned threads = copy_threads(ctx);
    for (size_t off = 0; off < bytes; off += slot.h_out_bytes) {
        const size_t len = std::min(slot.h_out_bytes, bytes - off);
        IBU_CUDA(cudaMemcpyAsync(slot.h_out, (const uint8_t *)table->d_rows + off, len, cudaMemcpyDeviceToHost, ctx->stream));
        IBU_CUDA(cudaStreamSynchronize(ctx->stream));
        parallel_memcpy((uint8_t *)h_rows + off, slot.h_out, len, threads);
    }
    return IBU_OK;
}

// ---- device path of load_to_vec -----------------------------------------------------------

int ibu_gpu_load_to_device(ibu_gpu_ctx_t *ctx, const char *path, uint64_t start, uint64_t end,
                           ibu_header_t *header, ibu_record_t **d_records, uint64_t *n, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !path || !header || !d_records || !n) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    *d_records = nullptr;
    *n = 0;
    ibu_mmap_reader_t *reader = nullptr;
    // same checks, same order as load_to_vec: header read + validate, then the size rule
    if (int rc = ibu_mmap_open(path, &reader, err)) return rc;
    *header = reader->header;
    if (end == UINT64_MAX) end = reader->len;
    if (start > end || end > reader->len) {
        int rc = set_error(err, IBU_ERR_INVALID_INDEX, 0, end, reader->len,
                           "Invalid index (%llu) - Must be less than %zu", (unsigned long long)end, reader->len);
        ibu_mmap_close(reader);
        return rc;
    }
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    const uint64_t count = end - start;
    void *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, count ? count * IBU_RECORD_SIZE : kAlign);
    if (e != cudaSuccess) {
        ibu_mmap_close(reader);
        return cuda_fail(err, e, "cudaMalloc");
    }
    int rc = IBU_OK;
    if (count) {
        const uint64_t chunk = chunk_records(ctx);
        const uint64_t n_chunks = (count + chunk - 1) / chunk;
        const unsigned threads = copy_threads(ctx);
        const size_t n_slots = ctx->slots.size();
        // pageable mmap -> pinned slot buffer (host threads) -> DMA into the final allocation
        for (uint64_t c = 0; c < n_chunks && rc == IBU_OK; c++) {
            ibu_chunk_slot &slot = ctx->slots[c % n_slots];
            const uint64_t off = c * chunk * IBU_RECORD_SIZE;
            const uint64_t bytes = std::min(chunk, count - c * chunk) * IBU_RECORD_SIZE;
            if (c >= n_slots && (e = cudaEventSynchronize(slot.done)) != cudaSuccess) {
                rc = cuda_fail(err, e, "cudaEventSynchronize");
                break;
            }
            if ((rc = ensure(&slot.h_in, &slot.h_in_bytes, align_up(chunk * IBU_RECORD_SIZE), true, err))) break;
            if (!parallel_pread(slot.h_in, ibu_mmap_fd(reader), IBU_HEADER_SIZE + start * IBU_RECORD_SIZE + off,
                                bytes, threads)) {
                rc = set_error(err, IBU_ERR_IO, errno, 0, 0, "I/O error: pread failed while loading");
                break;
            }
            e = cudaMemcpyAsync((uint8_t *)dev + off, slot.h_in, bytes, cudaMemcpyHostToDevice, slot.stream);
            if (e == cudaSuccess) e = cudaEventRecord(slot.done, slot.stream);
            if (e != cudaSuccess) rc = cuda_fail(err, e, "cudaMemcpyAsync");
        }
        for (auto &slot : ctx->slots) {
            e = cudaStreamSynchronize(slot.stream);
            if (e != cudaSuccess && rc == IBU_OK) rc = cuda_fail(err, e, "cudaStreamSynchronize");
        }
    }
    ibu_mmap_close(reader);
    if (rc != IBU_OK) {
        cudaFree(dev);
        return rc;
    }
    *d_records = (ibu_record_t *)dev;
    *n = count;
    return IBU_OK;
}

// ---- Writer device path ----------------------------------------------------------------------

int ibu_gpu_write_records(ibu_gpu_ctx_t *ctx, ibu_writer_t *writer, const ibu_record_t *d_records, uint64_t n,
                          ibu_error_t *err) {
    clear_error(err);
    if (!ctx || !writer || (!d_records && n)) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    const uint64_t chunk = chunk_records(ctx);
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const size_t n_slots = ctx->slots.size();
    int rc = IBU_OK;
    auto flush = [&](uint64_t c) -> int {  // chunk c has landed in its slot: append it to the file
        ibu_chunk_slot &slot = ctx->slots[c % n_slots];
        cudaError_t e = cudaEventSynchronize(slot.done);
        if (e != cudaSuccess) return cuda_fail(err, e, "cudaEventSynchronize");
        const uint64_t cnt = std::min(chunk, n - c * chunk);
        return ibu_writer_write_batch(writer, (const ibu_record_t *)slot.h_out, cnt, err);
    };
    for (uint64_t c = 0; c < n_chunks && rc == IBU_OK; c++) {
        ibu_chunk_slot &slot = ctx->slots[c % n_slots];
        if (c >= n_slots && (rc = flush(c - n_slots)) != IBU_OK) break;  // file order = chunk order
        if ((rc = ensure(&slot.h_out, &slot.h_out_bytes, align_up(chunk * IBU_RECORD_SIZE), true, err))) break;
        const uint64_t cnt = std::min(chunk, n - c * chunk);
        cudaError_t e = cudaMemcpyAsync(slot.h_out, d_records + c * chunk, cnt * IBU_RECORD_SIZE,
                                        cudaMemcpyDeviceToHost, slot.stream);
        if (e == cudaSuccess) e = cudaEventRecord(slot.done, slot.stream);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "cudaMemcpyAsync");
    }
    for (uint64_t c = n_chunks > n_slots ? n_chunks - n_slots : 0; c < n_chunks && rc == IBU_OK; c++) rc = flush(c);
    if (rc != IBU_OK)
        for (auto &slot : ctx->slots)
            if (cudaStreamSynchronize(slot.stream) != cudaSuccess) cudaGetLastError();
    return rc;
}

// ---- host -> host unpack / pack through the GPU --------------------------------------------

int ibu_gpu_unpack_host(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n, uint32_t bc_len,
                        uint32_t umi_len, uint8_t *h_bc_ascii, uint8_t *h_umi_ascii, uint8_t *h_flags,
                        ibu_reduce_result_t *h_result, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (n && (!h_records || !h_bc_ascii || !h_umi_ascii)))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    ibu_reduce_result_t total{};
    if (h_result) *h_result = total;
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    const uint64_t chunk = chunk_records(ctx);
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const bool pin_in = is_pinned(h_records), pin_bc = is_pinned(h_bc_ascii), pin_umi = is_pinned(h_umi_ascii),
               pin_fl = h_flags && is_pinned(h_flags);
    const size_t off_umi = align_up(chunk * bc_len), off_fl = off_umi + align_up(chunk * umi_len);
    const size_t out_bytes = off_fl + (h_flags ? align_up(chunk) : 0);
    int rc = run_chunks(
        ctx, n_chunks,
        [&](uint64_t c) {
            const uint64_t start = c * chunk, cnt = std::min(chunk, n - start);
            ChunkPlan p;
            p.n_in = 1;
            p.in[0] = {h_records + start, nullptr, cnt * IBU_RECORD_SIZE, 0, pin_in};
            p.out[0] = {nullptr, h_bc_ascii + start * bc_len, cnt * bc_len, 0, pin_bc};
            p.out[1] = {nullptr, h_umi_ascii + start * umi_len, cnt * umi_len, off_umi, pin_umi};
            p.n_out = 2;
            if (h_flags) p.out[p.n_out++] = {nullptr, h_flags + start, cnt, off_fl, pin_fl};
            p.d_in_bytes = align_up(chunk * IBU_RECORD_SIZE);
            p.d_out_bytes = out_bytes;
            return p;
        },
        [&](ibu_chunk_slot &slot, uint64_t c) {
            const uint64_t start = c * chunk, cnt = std::min(chunk, n - start);
            uint8_t *d_out = (uint8_t *)slot.d_out;
            return ibu_gpu_unpack_async(ctx, (const ibu_record_t *)slot.d_in, cnt, bc_len, umi_len, d_out,
                                        d_out + off_umi, h_flags ? d_out + off_fl : nullptr, slot.d_result,
                                        slot.stream, err);
        },
        [&](uint64_t, const ibu_reduce_result_t &r) {
            merge(total, r);
            return 0;
        },
        err);
    if (h_result) *h_result = total;
    return rc;
}

int ibu_gpu_pack_host(ibu_gpu_ctx_t *ctx, const uint8_t *h_bc_ascii, const uint8_t *h_umi_ascii,
                      const uint64_t *h_index, uint64_t index_base, uint64_t n, uint32_t bc_len,
                      uint32_t umi_len, ibu_record_t *h_records, uint8_t *h_flags,
                      ibu_reduce_result_t *h_result, ibu_error_t *err) {
    clear_error(err);
    if (!ctx || (n && (!h_records || !h_bc_ascii || !h_umi_ascii)))
        return set_error(err, IBU_ERR_ARG, 0, 0, 0, "null argument");
    ibu_reduce_result_t total{};
    if (h_result) *h_result = total;
    if (n == 0) return IBU_OK;
    DeviceGuard guard(ctx->device);
    IBU_PIPE_LOCK(ctx);
    const uint64_t chunk = chunk_records(ctx);
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const bool pin_bc = is_pinned(h_bc_ascii), pin_umi = is_pinned(h_umi_ascii),
               pin_idx = h_index && is_pinned(h_index), pin_out = is_pinned(h_records),
               pin_fl = h_flags && is_pinned(h_flags);
    const size_t in_umi = align_up(chunk * bc_len), in_idx = in_umi + align_up(chunk * umi_len);
    const size_t in_bytes = in_idx + (h_index ? align_up(chunk * 8) : 0);
    const size_t off_fl = align_up(chunk * IBU_RECORD_SIZE);
    const size_t out_bytes = off_fl + (h_flags ? align_up(chunk) : 0);
    int rc = run_chunks(
        ctx, n_chunks,
        [&](uint64_t c) {
            const uint64_t start = c * chunk, cnt = std::min(chunk, n - start);
            ChunkPlan p;
            p.in[0] = {h_bc_ascii + start * bc_len, nullptr, cnt * bc_len, 0, pin_bc};
            p.in[1] = {h_umi_ascii + start * umi_len, nullptr, cnt * umi_len, in_umi, pin_umi};
            p.n_in = 2;
            if (h_index) p.in[p.n_in++] = {h_index + start, nullptr, cnt * 8, in_idx, pin_idx};
            p.out[0] = {nullptr, h_records + start, cnt * IBU_RECORD_SIZE, 0, pin_out};
            p.n_out = 1;
            if (h_flags) p.out[p.n_out++] = {nullptr, h_flags + start, cnt, off_fl, pin_fl};
            p.d_in_bytes = in_bytes;
            p.d_out_bytes = out_bytes;
            return p;
        },
        [&](ibu_chunk_slot &slot, uint64_t c) {
            const uint64_t start = c * chunk, cnt = std::min(chunk, n - start);
            const uint8_t *d_in = (const uint8_t *)slot.d_in;
            uint8_t *d_out = (uint8_t *)slot.d_out;
            return ibu_gpu_pack_async(ctx, d_in, d_in + in_umi, h_index ? (const uint64_t *)(d_in + in_idx) : nullptr,
                                      index_base + start, cnt, bc_len, umi_len, (ibu_record_t *)d_out,
                                      h_flags ? d_out + off_fl : nullptr, slot.d_result, slot.stream, err);
        },
        [&](uint64_t, const ibu_reduce_result_t &r) {
            merge(total, r);
            return 0;
        },
        err);
    if (h_result) *h_result = total;
    return rc;
}

}  // extern "C"
