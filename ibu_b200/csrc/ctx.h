// ctx.h — the GPU context behind ibu_gpu_ctx_t (internal).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <atomic>
#include <cuda_runtime.h>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"

struct ibu_chunk_slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void *h_in = nullptr;    // pinned staging for pageable sources
    void *h_out = nullptr;   // pinned staging for pageable destinations
    void *d_in = nullptr;
    void *d_out = nullptr;
    size_t h_in_bytes = 0, h_out_bytes = 0, d_in_bytes = 0, d_out_bytes = 0;
    ibu_reduce_result_t *d_result = nullptr;  // device
    ibu_reduce_result_t *h_result = nullptr;  // pinned
};

// One entry of the result-scratch ring: 256 spread result blocks (256 x 8 u64, kept zeroed between
// uses) that the warps of a K1/K2 launch RED into, and the event recorded after the launch's fold
// kernel, which the next user of the entry waits on.
struct ibu_result_scratch {
    unsigned long long *blocks = nullptr;
    cudaEvent_t folded = nullptr;
    std::mutex in_use;  // held from the wait on `folded` to its next record (see ResultLease)
};
constexpr int kResultBlocks = 256;
constexpr int kResultRing = 16;

struct ibu_gpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // default stream of the context (non-blocking)
    ibu_gpu_config_t cfg{};
    std::vector<ibu_chunk_slot> slots;
    // grow-only device scratch of the blocking table builder (K4): cudaMalloc/cudaFree per call
    // would cost more than the streaming pass itself
    std::array<ibu_result_scratch, kResultRing> result_ring;
    std::atomic<uint32_t> result_next{0};
    std::mutex pipe_mutex;  // the chunk slots serve one host-buffer call at a time
    std::atomic<std::thread::id> pipe_owner{};  // who holds it: a chunk callback that calls back in is refused, not deadlocked
    // the streaming ingest that owns the slots between its open and close (guarded by pipe_mutex, which
    // is NOT held across calls: a stream may be closed from another thread, or after the context)
    struct ibu_gpu_stream *open_stream = nullptr;
    std::mutex arena_mutex;
    void *arena_base = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // pinned mailbox for the small device -> host read-backs of the table builder (guarded by
    // arena_mutex): a copy into it is a DMA, not a staged pageable copy
    unsigned long long *h_mail = nullptr;
    cudaEvent_t rows_ev = nullptr;  // orders a result block allocated on `stream` before its use on a caller's stream
};
constexpr size_t kMailBytes = 4096;

namespace ibu {

extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(ibu_error_t *err, cudaError_t e, const char *what) {
    return set_error(err, IBU_ERR_CUDA, (int)e, 0, 0, "CUDA error in %s: %s", what,
                     cudaGetErrorString(e));
}

#define IBU_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return ibu::cuda_fail(err, e__, #call);  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// A result block the caller will own (table rows: ibu_gpu_table_free releases them on ctx->stream).
// Allocated on the context's stream whatever stream the call runs on: a block freed on one stream and
// next requested on another is not reused by the pool until the driver has looked, and a 2.4 GB block
// (10^8 rows) then costs a fresh mapping — 7 to 200 ms per call measured, against 4.5 ms when
// allocation and release share a stream.  Caller holds ctx->arena_mutex (rows_ev).
inline cudaError_t alloc_result_rows(ibu_gpu_ctx *ctx, uint64_t **out, size_t bytes, cudaStream_t user) {
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMallocAsync((void **)out, bytes ? bytes : 256, ctx->stream);
    if (getenv("IBU_B200_TRACE_ALLOC")) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 0.2) fprintf(stderr, "[ibu trace] cudaMallocAsync (result rows) of %.3f GB took %.2f ms\n", bytes / 1e9, ms);
    }
    if (e != cudaSuccess || user == ctx->stream) return e;
    if (!ctx->rows_ev) e = cudaEventCreateWithFlags(&ctx->rows_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->rows_ev, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(user, ctx->rows_ev, 0);
    if (e != cudaSuccess) {
        cudaFreeAsync(*out, ctx->stream);
        *out = nullptr;
    }
    return e;
}

inline cudaStream_t pick_stream(ibu_gpu_ctx *ctx, void *stream) {
    return stream ? (cudaStream_t)stream : ctx->stream;
}

}  // namespace ibu
