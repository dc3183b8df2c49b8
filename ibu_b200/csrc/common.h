// common.h — internal declarations shared by the host and device halves of libibu_b200.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ibu_b200.h"

static_assert(sizeof(ibu_header_t) == IBU_HEADER_SIZE, "Header must stay bytemuck-identical");
static_assert(sizeof(ibu_record_t) == IBU_RECORD_SIZE, "Record must stay bytemuck-identical");
static_assert(sizeof(ibu_reduce_result_t) == 64, "result block is 8 u64");

namespace ibu {

inline int set_error(ibu_error_t *err, int code, int sys, uint64_t a, uint64_t b, const char *fmt,
                     ...) {
    if (err) {
        err->code = code;
        err->sys = sys;
        err->a = a;
        err->b = b;
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err->msg, sizeof(err->msg), fmt, ap);
        va_end(ap);
    }
    return code;
}

inline void clear_error(ibu_error_t *err) {
    if (err) {
        err->code = IBU_OK;
        err->sys = 0;
        err->a = err->b = 0;
        err->msg[0] = 0;
    }
}

// bits of a 2-bit word that must be zero for a sequence of `len` bases
inline uint64_t high_mask(uint32_t len) { return len >= 32 ? 0ull : ~((1ull << (2 * len)) - 1); }

}  // namespace ibu

// the mmap reader is shared between host_format.cpp and the staging pipeline
#include <atomic>
#include <mutex>
#include <vector>
struct ibu_mmap_shared {  // Arc<Mmap> (src/io/mmap.rs:99-107)
    const uint8_t *base;
    size_t bytes;
    int fd;  // kept open: the staging pipeline may pread() instead of touching the mapping
    std::atomic<long> refs;
    // page-locked parts of the mapping (ibu_mmap_pin / ibu_mmap_pin_range): registered once however
    // many clones ask, released by the matching unpin or, at the latest, before the mapping goes away
    struct PinnedRange {
        uint8_t *p;
        size_t bytes;
        long count;
    };
    std::mutex pin_mutex;
    std::vector<PinnedRange> pinned;
};
struct ibu_mmap_reader {
    ibu_mmap_shared *shared;  // Arc<Mmap>
    ibu_header_t header;
    size_t len;
};
const uint8_t *ibu_mmap_base(const ibu_mmap_reader *r);  // start of the mapping (header included)
size_t ibu_mmap_bytes(const ibu_mmap_reader *r);        // size of the mapping
int ibu_mmap_fd(const ibu_mmap_reader *r);              // descriptor of the mapped file (for pread staging)
