// ibu_oracle.cpp — TEST INFRASTRUCTURE ONLY (see ibu_oracle.h for the parity status).
//
// CPU restatement of the reference's bulk record path.  Every function cites the
// reference lines it follows (paths relative to the reference crate root).  The
// reference is Rust and cannot be compiled in this environment (no cargo/rustc), so
// this file restates its algorithm in C++17; threads are std::thread exactly where
// the reference uses std::thread::spawn, processors are template parameters exactly
// where the reference monomorphises a generic.
#include "ibu_oracle.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <unordered_set>
#include <vector>

static_assert(sizeof(orc_header_t) == 32, "header.rs:7 HEADER_SIZE");
static_assert(sizeof(orc_record_t) == 24, "record.rs:3 RECORD_SIZE");

namespace {

constexpr uint32_t MAGIC = 0x21554249u;       // header.rs:5
constexpr uint32_t VERSION = 2;               // header.rs:6
constexpr size_t HEADER_SIZE = 32;            // header.rs:7
constexpr size_t RECORD_SIZE = 24;            // record.rs:3
constexpr size_t BATCH_SIZE = 1024 * 1024;    // mmap.rs:284
constexpr size_t DEFAULT_BUFFER_SIZE = 48 * 1024 * RECORD_SIZE;  // reader.rs:14, writer.rs:10

int fail(orc_error_t *err, int code, uint64_t a = 0, uint64_t b = 0, int sys = 0) {
    if (err) {
        err->code = code;
        err->sys = sys;
        err->a = a;
        err->b = b;
    }
    return code;
}

}  // namespace

struct orc_mmap {
    // MmapReader { map: Arc<Mmap>, header, len } (mmap.rs:99-107)
    const uint8_t *map = nullptr;
    size_t map_len = 0;
    orc_header_t header{};
    size_t len = 0;
};

extern "C" {

// Header::new (header.rs:84-93)
void orc_header_new(orc_header_t *h, uint32_t bc_len, uint32_t umi_len) {
    std::memset(h, 0, sizeof(*h));
    h->magic = MAGIC;
    h->version = VERSION;
    h->bc_len = bc_len;
    h->umi_len = umi_len;
}

// Header::validate (header.rs:167-187): four checks, in this order.
int orc_header_validate(const orc_header_t *h, orc_error_t *err) {
    if (h->magic != MAGIC) return fail(err, ORC_INVALID_MAGIC, MAGIC, h->magic);
    if (h->version != VERSION) return fail(err, ORC_INVALID_VERSION, VERSION, h->version);
    if (h->bc_len == 0 || h->bc_len > 32) return fail(err, ORC_INVALID_BC_LEN, h->bc_len);
    if (h->umi_len == 0 || h->umi_len > 32) return fail(err, ORC_INVALID_UMI_LEN, h->umi_len);
    return ORC_OK;
}

// MmapReader::new (mmap.rs:143-161)
int orc_mmap_open(const char *path, orc_mmap_t **out, orc_error_t *err) {
    *out = nullptr;
    int fd = ::open(path, O_RDONLY);  // File::open
    if (fd < 0) return fail(err, ORC_IO, 0, 0, errno);
    struct stat st;
    if (fstat(fd, &st) != 0) {
        int e = errno;
        ::close(fd);
        return fail(err, ORC_IO, 0, 0, e);
    }
    size_t flen = (size_t)st.st_size;
    // Mmap::map fails on a zero-length file (Io); &map[0..32] would panic below 32 bytes
    // (mmap.rs:149) — both reported as Io here.
    if (flen < HEADER_SIZE) {
        ::close(fd);
        return fail(err, ORC_IO, 0, 0, EINVAL);
    }
    void *p = mmap(nullptr, flen, PROT_READ, MAP_PRIVATE, fd, 0);
    int e = errno;
    ::close(fd);
    if (p == MAP_FAILED) return fail(err, ORC_IO, 0, 0, e);
    orc_header_t h;
    std::memcpy(&h, p, HEADER_SIZE);  // Header::from_bytes(&map[0..HEADER_SIZE])
    int rc = orc_header_validate(&h, err);
    if (rc) {
        munmap(p, flen);
        return rc;
    }
    size_t record_bytes = flen - HEADER_SIZE;
    if (record_bytes % RECORD_SIZE != 0) {  // mmap.rs:154-157
        munmap(p, flen);
        return fail(err, ORC_INVALID_MAP_SIZE);
    }
    auto *m = new orc_mmap;
    m->map = (const uint8_t *)p;
    m->map_len = flen;
    m->header = h;
    m->len = record_bytes / RECORD_SIZE;
    *out = m;
    return ORC_OK;
}

void orc_mmap_close(orc_mmap_t *m) {
    if (!m) return;
    munmap((void *)m->map, m->map_len);
    delete m;
}
size_t orc_mmap_len(const orc_mmap_t *m) { return m->len; }                 // mmap.rs:178-180
orc_header_t orc_mmap_header(const orc_mmap_t *m) { return m->header; }     // mmap.rs:201-203

// MmapReader::slice (mmap.rs:253-270)
int orc_mmap_slice(const orc_mmap_t *m, size_t start, size_t end, const orc_record_t **out,
                   size_t *n, orc_error_t *err) {
    if (start >= m->len || end > m->len) return fail(err, ORC_INVALID_INDEX, end, m->len);
    if (end <= start) return fail(err, ORC_INVALID_INDEX, end, m->len);
    *out = (const orc_record_t *)(m->map + HEADER_SIZE + start * RECORD_SIZE);
    *n = end - start;
    return ORC_OK;
}

// load_to_vec (reader.rs:510-535)
int orc_load_to_vec(const char *path, orc_header_t *h, orc_record_t **records, size_t *n,
                    orc_error_t *err) {
    *records = nullptr;
    *n = 0;
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(err, ORC_IO, 0, 0, errno);
    uint8_t hb[HEADER_SIZE];
    if (std::fread(hb, 1, HEADER_SIZE, f) != HEADER_SIZE) {  // read_exact
        std::fclose(f);
        return fail(err, ORC_IO, 0, 0, EIO);
    }
    std::memcpy(h, hb, HEADER_SIZE);
    int rc = orc_header_validate(h, err);
    if (rc) {
        std::fclose(f);
        return rc;
    }
    struct stat st;
    fstat(fileno(f), &st);
    size_t data_size = (size_t)st.st_size - HEADER_SIZE;
    if (data_size % RECORD_SIZE != 0) {
        std::fclose(f);
        return fail(err, ORC_INVALID_MAP_SIZE);
    }
    size_t num = data_size / RECORD_SIZE;
    // vec![Record::default(); n]: zero-filled, then one read_exact over the byte view
    auto *recs = (orc_record_t *)std::calloc(num ? num : 1, sizeof(orc_record_t));
    if (num && std::fread(recs, RECORD_SIZE, num, f) != num) {
        std::free(recs);
        std::fclose(f);
        return fail(err, ORC_IO, 0, 0, EIO);
    }
    std::fclose(f);
    *records = recs;
    *n = num;
    return ORC_OK;
}

void orc_free(void *p) { std::free(p); }

// Reader::new + first Iterator::next over an in-memory stream (reader.rs:152-176,218-242,282-305)
int orc_stream_first(const uint8_t *bytes, size_t len, orc_record_t *rec, orc_error_t *err) {
    if (len < HEADER_SIZE) return fail(err, ORC_IO, 0, 0, EIO);
    orc_header_t h;
    std::memcpy(&h, bytes, HEADER_SIZE);  // pod_read_unaligned
    int rc = orc_header_validate(&h, err);
    if (rc) return rc;
    size_t bytes_read = HEADER_SIZE;
    size_t read = std::min(len - HEADER_SIZE, DEFAULT_BUFFER_SIZE);  // read_batch fills the buffer
    if (read % RECORD_SIZE != 0) {
        size_t non_rem = read - read % RECORD_SIZE;
        return fail(err, ORC_TRUNCATED, bytes_read + non_rem);
    }
    if (read == 0) return -1;  // iterator exhausted (None)
    std::memcpy(rec, bytes + HEADER_SIZE, RECORD_SIZE);
    return ORC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Writer (writer.rs:82-95 state; 129-143 new; 220-226 flush_buffer; 260-273 write_record;
// 315-351 write_batch/write_slice; 429-433 finish)
namespace {
struct Writer {
    FILE *inner;
    std::vector<uint8_t> buffer;
    size_t pos = 0;
    uint64_t records_written = 0;
    explicit Writer(FILE *f) : inner(f), buffer(DEFAULT_BUFFER_SIZE, 0) {}
    bool flush_buffer() {
        if (pos > 0) {
            if (std::fwrite(buffer.data(), 1, pos, inner) != pos) return false;
            pos = 0;
        }
        return true;
    }
    bool write_record(const orc_record_t &r) {
        if (pos + RECORD_SIZE > buffer.size() && !flush_buffer()) return false;
        std::memcpy(buffer.data() + pos, &r, RECORD_SIZE);
        pos += RECORD_SIZE;
        records_written += 1;
        return true;
    }
    bool write_slice(const uint8_t *buf, size_t len) {
        size_t num = len / RECORD_SIZE;
        if (len > buffer.size()) {  // direct write for batches larger than the buffer
            if (!flush_buffer()) return false;
            if (std::fwrite(buf, 1, len, inner) != len) return false;
            records_written += num;
            return true;
        }
        while (len) {
            size_t avail = buffer.size() - pos;
            size_t w = std::min(len, avail);
            std::memcpy(buffer.data() + pos, buf, w);
            pos += w;
            buf += w;
            len -= w;
            if (pos >= buffer.size() && !flush_buffer()) return false;
        }
        records_written += num;
        return true;
    }
    bool finish() { return flush_buffer() && std::fflush(inner) == 0; }
};
}  // namespace

extern "C" int orc_write_file(const char *path, const orc_header_t *h, const orc_record_t *recs,
                              size_t n, int mode, orc_error_t *err) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return fail(err, ORC_IO, 0, 0, errno);
    std::setvbuf(f, nullptr, _IONBF, 0);  // the Writer's own buffer is the only buffering
    // Writer::new writes the header bytes immediately, without validating (writer.rs:129-133)
    bool ok = std::fwrite(h, 1, HEADER_SIZE, f) == HEADER_SIZE;
    Writer w(f);
    if (ok) {
        if (mode == 0) {
            for (size_t i = 0; ok && i < n; i++) ok = w.write_record(recs[i]);
        } else {
            ok = w.write_slice((const uint8_t *)recs, n * RECORD_SIZE);
        }
    }
    ok = ok && w.finish();
    int e = errno;
    std::fclose(f);
    return ok ? ORC_OK : fail(err, ORC_IO, 0, 0, e);
}

// ---------------------------------------------------------------------------------------
// process_parallel (mmap.rs:286-332)

extern "C" int orc_num_cpus(void) {
    // num_cpus::get(): logical CPUs available to the process (affinity-aware on Linux)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

namespace {

// A source of records addressed like MmapReader::slice.
struct MmapSource {
    const orc_mmap_t *m;
    size_t len() const { return m->len; }
    int slice(size_t s, size_t e, const orc_record_t **out, size_t *n, orc_error_t *err) const {
        return orc_mmap_slice(m, s, e, out, n, err);
    }
};
struct ArraySource {
    const orc_record_t *recs;
    size_t n_total;
    size_t len() const { return n_total; }
    int slice(size_t s, size_t e, const orc_record_t **out, size_t *n, orc_error_t *err) const {
        if (s >= n_total || e > n_total || e <= s) return fail(err, ORC_INVALID_INDEX, e, n_total);
        *out = recs + s;
        *n = e - s;
        return ORC_OK;
    }
};

// The driver.  P needs: P(const P&) [Clone], int process_record(const orc_record_t&),
// int on_batch_complete().  `first` carries slice-relative record numbers for processors
// that write positional output (the reference passes records by value only; the position
// is a property of this restatement's batch processors, not of the reference trait).
template <class Source, class P>
int process_parallel(const Source &src, const P &processor, size_t num_threads, orc_error_t *err,
                     orc_thread_trace_t *trace = nullptr, size_t max_threads = 0,
                     size_t *n_threads_out = nullptr) {
    size_t ncpu = (size_t)orc_num_cpus();
    num_threads = num_threads == 0 ? ncpu : std::min(num_threads, ncpu);  // mmap.rs:292-296
    size_t len = src.len();
    size_t records_per_thread = len / num_threads;  // mmap.rs:297
    size_t remainder = len % num_threads;           // mmap.rs:298 (for the last thread)
    if (n_threads_out) *n_threads_out = num_threads;

    std::vector<std::thread> handles;
    std::vector<orc_error_t> errs(num_threads);
    std::vector<int> rcs(num_threads, ORC_OK);
    handles.reserve(num_threads);
    for (size_t i = 0; i < num_threads; i++) {
        size_t start = i * records_per_thread;
        size_t end = (i == num_threads - 1) ? start + records_per_thread + remainder
                                            : start + records_per_thread;  // mmap.rs:301-307
        P thread_processor(processor);  // processor.clone() (mmap.rs:309)
        orc_thread_trace_t *tr = (trace && i < max_threads) ? &trace[i] : nullptr;
        if (tr) *tr = {start, end, 0, 0};
        handles.emplace_back([&src, start, end, tr, i, &errs, &rcs,
                              tp = std::move(thread_processor)]() mutable {
            size_t batch_start = start;
            while (batch_start < end) {  // mmap.rs:312
                size_t batch_end = std::min(batch_start + BATCH_SIZE, end);
                const orc_record_t *slice;
                size_t n;
                int rc = src.slice(batch_start, batch_end, &slice, &n, &errs[i]);
                if (rc) {
                    rcs[i] = rc;
                    return;
                }
                tp.begin_slice(batch_start);
                for (size_t k = 0; k < n; k++) {  // the hot loop, mmap.rs:315-317
                    rc = tp.process_record(slice[k]);
                    if (rc) {
                        rcs[i] = fail(&errs[i], ORC_PROCESS, slice[k].index);
                        return;
                    }
                }
                rc = tp.on_batch_complete();  // mmap.rs:318
                if (rc) {
                    rcs[i] = fail(&errs[i], ORC_PROCESS);
                    return;
                }
                if (tr) {
                    tr->records += n;
                    tr->batches += 1;
                }
                batch_start += BATCH_SIZE;  // mmap.rs:319
            }
        });
    }
    // join in spawn order; the first Err in spawn order is returned (mmap.rs:326-328).
    // (The reference returns early and detaches the rest; joining them all is
    // observationally the same for the result.)
    for (auto &h : handles) h.join();
    for (size_t i = 0; i < num_threads; i++)
        if (rcs[i]) {
            if (err) *err = errs[i];
            return rcs[i];
        }
    return ORC_OK;
}

inline uint64_t low_mask(uint32_t len) { return len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1); }

// Shared accumulator behind Arc<Mutex<..>> (examples/parallel.rs:8-36) / Arc<AtomicU64>
// (mmap.rs:350-373): merged in on_batch_complete.
struct SharedReduce {
    orc_reduce_t total{};
    std::mutex mu;
};

// One processor computing every built-in reduction of SURVEY §8a a12 at once:
//   count + sum of fields      mmap.rs:359-363 (local_count, local_sum = bc+umi+idx)
//   three field sums           examples/parallel.rs:22-27
//   xor checksum               examples/roundtrip.rs:84-87
//   bad-word counters          new semantics: valid_word(w, L) = L==32 || (w >> 2L) == 0
struct ReduceProcessor {
    orc_reduce_t local{};
    SharedReduce *global;
    uint64_t bc_hi, umi_hi;  // bits that must be zero
    ReduceProcessor(SharedReduce *g, uint32_t bc_len, uint32_t umi_len)
        : global(g), bc_hi(~low_mask(bc_len)), umi_hi(~low_mask(umi_len)) {}
    void begin_slice(size_t) {}
    int process_record(const orc_record_t &r) {
        local.n_records += 1;
        local.sum_barcode += r.barcode;  // wrapping (release-mode +=)
        local.sum_umi += r.umi;
        local.sum_index += r.index;
        local.xor_all ^= r.barcode ^ r.umi ^ r.index;
        bool bb = (r.barcode & bc_hi) != 0, bu = (r.umi & umi_hi) != 0;
        local.n_bad_barcode += bb;
        local.n_bad_umi += bu;
        local.n_bad_records += (bb || bu);
        return 0;
    }
    int on_batch_complete() {
        std::lock_guard<std::mutex> g(global->mu);
        orc_reduce_t &t = global->total;
        t.n_records += local.n_records;
        t.sum_barcode += local.sum_barcode;
        t.sum_umi += local.sum_umi;
        t.sum_index += local.sum_index;
        t.xor_all ^= local.xor_all;
        t.n_bad_barcode += local.n_bad_barcode;
        t.n_bad_umi += local.n_bad_umi;
        t.n_bad_records += local.n_bad_records;
        local = orc_reduce_t{};
        return 0;
    }
};

// ErrorProcessor (parallel.rs:338-352)
struct FailProcessor {
    uint64_t fail_on;
    void begin_slice(size_t) {}
    int process_record(const orc_record_t &r) { return r.index == fail_on ? 1 : 0; }
    int on_batch_complete() { return 0; }
};

// Generic vtable-backed processor: the ParallelProcessor trait (parallel.rs:100-190)
struct VtableProcessor {
    const orc_processor_vtable_t *vt;
    void *self;
    bool owned;
    VtableProcessor(const orc_processor_vtable_t *v, void *s) : vt(v), self(s), owned(false) {}
    VtableProcessor(const VtableProcessor &o)
        : vt(o.vt), self(o.vt->clone ? o.vt->clone(o.self) : o.self), owned(o.vt->clone != nullptr) {}
    VtableProcessor(VtableProcessor &&o) noexcept : vt(o.vt), self(o.self), owned(o.owned) {
        o.owned = false;
    }
    ~VtableProcessor() {
        if (owned && vt->drop) vt->drop(self);
    }
    void begin_slice(size_t) {}
    int process_record(const orc_record_t &r) { return vt->process_record ? vt->process_record(self, &r) : 0; }
    int on_batch_complete() { return vt->on_batch_complete ? vt->on_batch_complete(self) : 0; }  // default: Ok(())
};

// Barcode histogram: HashMap<u64,u64> barcode -> count merged under a mutex in
// on_batch_complete (parallel.rs:79-98), extended with the set of UMIs per barcode.
struct PairHash {
    size_t operator()(const std::pair<uint64_t, uint64_t> &p) const {
        return (size_t)orc_splitmix64(p.first ^ orc_splitmix64(p.second));
    }
};
struct SharedBarcodes {
    std::unordered_map<uint64_t, uint64_t> counts;
    std::unordered_set<std::pair<uint64_t, uint64_t>, PairHash> pairs;
    std::mutex mu;
};
struct BarcodeProcessor {
    std::unordered_map<uint64_t, uint64_t> local_counts;
    std::unordered_set<std::pair<uint64_t, uint64_t>, PairHash> local_pairs;
    SharedBarcodes *global;
    explicit BarcodeProcessor(SharedBarcodes *g) : global(g) {}
    void begin_slice(size_t) {}
    int process_record(const orc_record_t &r) {
        local_counts[r.barcode] += 1;  // *entry(barcode).or_insert(0) += 1
        local_pairs.emplace(r.barcode, r.umi);
        return 0;
    }
    int on_batch_complete() {
        std::lock_guard<std::mutex> g(global->mu);
        for (auto &kv : local_counts) global->counts[kv.first] += kv.second;
        for (auto &p : local_pairs) global->pairs.insert(p);
        local_counts.clear();
        local_pairs.clear();
        return 0;
    }
};

int barcode_rows(SharedBarcodes &sb, orc_barcode_row_t **rows, size_t *n_rows, uint64_t *n_pairs) {
    std::unordered_map<uint64_t, uint64_t> distinct;
    for (auto &p : sb.pairs) distinct[p.first] += 1;
    size_t n = sb.counts.size();
    auto *out = (orc_barcode_row_t *)std::malloc(sizeof(orc_barcode_row_t) * (n ? n : 1));
    size_t k = 0;
    for (auto &kv : sb.counts) out[k++] = {kv.first, kv.second, distinct[kv.first]};
    std::sort(out, out + n, [](const orc_barcode_row_t &a, const orc_barcode_row_t &b) {
        return a.barcode < b.barcode;  // Record's Ord leads with barcode (record.rs:58)
    });
    *rows = out;
    *n_rows = n;
    if (n_pairs) *n_pairs = sb.pairs.size();
    return ORC_OK;
}

// 2-bit codec — record.rs:19-27 (A=00 C=01 G=10 T=11), bitnuc order: base i at bits [2i, 2i+1].
inline void unpack_word(uint64_t w, uint32_t len, uint8_t *out) {
    static const char LUT[4] = {'A', 'C', 'G', 'T'};
    for (uint32_t i = 0; i < len; i++) out[i] = (uint8_t)LUT[(w >> (2 * i)) & 3];
}
inline int pack_word(const uint8_t *s, uint32_t len, uint64_t *w) {
    uint64_t acc = 0;
    int bad = 0;
    for (uint32_t i = 0; i < len; i++) {
        uint8_t c = s[i];
        uint8_t up = c & 0xDF;  // case-insensitive
        bad |= !(up == 'A' || up == 'C' || up == 'G' || up == 'T');
        uint64_t c1 = (c >> 1) & 3;
        acc |= (c1 ^ (c1 >> 1)) << (2 * i);
    }
    *w = acc;
    return bad;
}

// Processor that decodes each record into positional ASCII rows — what a user's
// process_record would do with bitnuc::from_2bit — plus validation counters.
struct UnpackProcessor {
    uint32_t bc_len, umi_len;
    uint64_t bc_hi, umi_hi;
    uint8_t *bc_out, *umi_out, *flags;
    size_t pos = 0;
    orc_reduce_t local{};
    SharedReduce *global;
    void begin_slice(size_t first) { pos = first; }
    int process_record(const orc_record_t &r) {
        unpack_word(r.barcode, bc_len, bc_out + pos * bc_len);
        unpack_word(r.umi, umi_len, umi_out + pos * umi_len);
        bool bb = (r.barcode & bc_hi) != 0, bu = (r.umi & umi_hi) != 0;
        if (flags) flags[pos] = (uint8_t)(bb | (bu << 1));
        local.n_records += 1;
        local.n_bad_barcode += bb;
        local.n_bad_umi += bu;
        local.n_bad_records += (bb || bu);
        pos += 1;
        return 0;
    }
    int on_batch_complete() {
        std::lock_guard<std::mutex> g(global->mu);
        global->total.n_records += local.n_records;
        global->total.n_bad_barcode += local.n_bad_barcode;
        global->total.n_bad_umi += local.n_bad_umi;
        global->total.n_bad_records += local.n_bad_records;
        local = orc_reduce_t{};
        return 0;
    }
};

// Run f(start, end) over the process_parallel partition of [0, n) (mmap.rs:292-307).
template <class F>
void partitioned(size_t n, size_t num_threads, F f) {
    size_t ncpu = (size_t)orc_num_cpus();
    num_threads = num_threads == 0 ? ncpu : std::min(num_threads, ncpu);
    size_t per = n / num_threads, rem = n % num_threads;
    std::vector<std::thread> hs;
    for (size_t i = 0; i < num_threads; i++) {
        size_t s = i * per, e = (i == num_threads - 1) ? s + per + rem : s + per;
        hs.emplace_back([=] { f(s, e); });
    }
    for (auto &h : hs) h.join();
}

}  // namespace

extern "C" {

int orc_process_parallel_reduce(const orc_mmap_t *m, size_t num_threads, orc_reduce_t *out,
                                orc_thread_trace_t *trace, size_t max_threads,
                                size_t *n_threads_out, orc_error_t *err) {
    SharedReduce shared;
    ReduceProcessor proc(&shared, m->header.bc_len, m->header.umi_len);
    int rc = process_parallel(MmapSource{m}, proc, num_threads, err, trace, max_threads,
                              n_threads_out);
    *out = shared.total;
    return rc;
}

int orc_process_parallel_fail(const orc_mmap_t *m, size_t num_threads, uint64_t fail_index,
                              orc_error_t *err) {
    return process_parallel(MmapSource{m}, FailProcessor{fail_index}, num_threads, err);
}

int orc_process_parallel(const orc_mmap_t *m, const orc_processor_vtable_t *vt, void *proc,
                         size_t num_threads, orc_error_t *err) {
    VtableProcessor p(vt, proc);
    return process_parallel(MmapSource{m}, p, num_threads, err);
}

int orc_process_parallel_barcodes(const orc_mmap_t *m, size_t num_threads,
                                  orc_barcode_row_t **rows, size_t *n_rows, orc_error_t *err) {
    SharedBarcodes shared;
    int rc = process_parallel(MmapSource{m}, BarcodeProcessor(&shared), num_threads, err);
    if (rc) return rc;
    return barcode_rows(shared, rows, n_rows, nullptr);
}

void orc_reduce_records(const orc_record_t *recs, size_t n, uint32_t bc_len, uint32_t umi_len,
                        size_t num_threads, orc_reduce_t *out) {
    SharedReduce shared;
    ReduceProcessor proc(&shared, bc_len, umi_len);
    process_parallel(ArraySource{recs, n}, proc, num_threads, nullptr);
    *out = shared.total;
}

int orc_barcode_table(const orc_record_t *recs, size_t n, orc_barcode_row_t **rows,
                      size_t *n_rows, uint64_t *n_pairs) {
    SharedBarcodes shared;
    process_parallel(ArraySource{recs, n}, BarcodeProcessor(&shared), 0, nullptr);
    return barcode_rows(shared, rows, n_rows, n_pairs);
}

int orc_valid_word(uint64_t w, uint32_t len) { return (w & ~low_mask(len)) == 0; }
void orc_unpack_word(uint64_t w, uint32_t len, uint8_t *out) { unpack_word(w, len, out); }
int orc_pack_word(const uint8_t *s, uint32_t len, uint64_t *w) { return pack_word(s, len, w); }

void orc_unpack_records(const orc_record_t *recs, size_t n, uint32_t bc_len, uint32_t umi_len,
                        uint8_t *bc_ascii, uint8_t *umi_ascii, uint8_t *flags,
                        size_t num_threads, orc_reduce_t *out) {
    SharedReduce shared;
    UnpackProcessor proc{bc_len, umi_len, ~low_mask(bc_len), ~low_mask(umi_len),
                         bc_ascii, umi_ascii, flags, 0, {}, &shared};
    process_parallel(ArraySource{recs, n}, proc, num_threads, nullptr);
    if (out) *out = shared.total;
}

void orc_pack_records(const uint8_t *bc_ascii, const uint8_t *umi_ascii, const uint64_t *index,
                      uint64_t index_base, size_t n, uint32_t bc_len, uint32_t umi_len,
                      orc_record_t *recs, uint8_t *flags, size_t num_threads, orc_reduce_t *out) {
    std::mutex mu;
    orc_reduce_t total{};
    partitioned(n, num_threads, [&](size_t s, size_t e) {
        orc_reduce_t loc{};
        for (size_t i = s; i < e; i++) {
            uint64_t b, u;
            int bb = pack_word(bc_ascii + i * bc_len, bc_len, &b);
            int bu = pack_word(umi_ascii + i * umi_len, umi_len, &u);
            recs[i] = {b, u, index ? index[i] : index_base + i};
            if (flags) flags[i] = (uint8_t)(bb | (bu << 1));
            loc.n_records += 1;
            loc.n_bad_barcode += bb;
            loc.n_bad_umi += bu;
            loc.n_bad_records += (bb || bu);
        }
        std::lock_guard<std::mutex> g(mu);
        total.n_records += loc.n_records;
        total.n_bad_barcode += loc.n_bad_barcode;
        total.n_bad_umi += loc.n_bad_umi;
        total.n_bad_records += loc.n_bad_records;
    });
    if (out) *out = total;
}

// ---- synthetic data ---------------------------------------------------------------------
uint64_t orc_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline orc_record_t gen_record(uint64_t i, uint32_t bc_len, uint32_t umi_len, int mode,
                                      uint64_t param, uint64_t seed) {
    uint64_t key = orc_splitmix64(seed ^ orc_splitmix64(i));
    uint64_t rb = orc_splitmix64(key ^ 1), ru = orc_splitmix64(key ^ 2);
    uint64_t mb = low_mask(bc_len), mu = low_mask(umi_len);
    switch (mode) {
        default:
        case 0:  // CLEAN
            return {rb & mb, ru & mu, i};
        case 1: {  // DIRTY: `param` ppm of records keep one raw (unmasked) word, as the
                   // reference generator does for every UMI (examples/random.rs:46)
            uint64_t rd = orc_splitmix64(key ^ 3);
            orc_record_t r{rb & mb, ru & mu, i};
            if (rd % 1000000ull < param) {
                if (rd >> 63) r.barcode = rb; else r.umi = ru;
            }
            return r;
        }
        case 2:  // PATTERN (examples/parallel.rs:65-69, examples/roundtrip.rs:34-38)
            return {i % 1000000ull, (i * 31ull) % 1000000ull, i};
        case 4: {  // SORTED: low 32 bits of param = records per barcode, high 32 = records per umi
            uint64_t rpb = param & 0xFFFFFFFFull, dup = param >> 32;
            if (rpb == 0) rpb = 1000;
            if (dup == 0) dup = 1;
            return {(i / rpb) & mb, ((i % rpb) / dup) & mu, i};
        }
        case 3: {  // WHITELIST: low 32 bits of param = #barcodes, high 32 bits = umi space
            uint64_t nb = param & 0xFFFFFFFFull, us = param >> 32;
            if (nb == 0) nb = 1000;  // examples/random.rs default --barcodes
            uint64_t b = orc_splitmix64((rb % nb) ^ seed ^ 0xB) & mb;
            uint64_t u = (us ? ru % us : ru) & mu;
            return {b, u, i};
        }
        case 5: {  // ZIPF: as WHITELIST with log-uniform barcode ranks: the exponent e is uniform in
                   // 0..floor(log2 nb) and the rank uniform in [2^e - 1, 2^(e+1) - 2] (mod nb)
            uint64_t nb = param & 0xFFFFFFFFull, us = param >> 32;
            if (nb == 0) nb = 1000;
            uint32_t levels = 64 - (uint32_t)__builtin_clzll(nb);
            uint32_t e = (uint32_t)(rb % levels);
            uint64_t r = (((1ull << e) - 1) + (orc_splitmix64(key ^ 7) & ((1ull << e) - 1))) % nb;
            uint64_t b = orc_splitmix64(r ^ seed ^ 0xB) & mb;
            uint64_t u = (us ? ru % us : ru) & mu;
            return {b, u, i};
        }
    }
}

void orc_generate_records(orc_record_t *recs, uint64_t first, uint64_t n, uint32_t bc_len,
                          uint32_t umi_len, int mode, uint64_t param, uint64_t seed,
                          size_t num_threads) {
    partitioned(n, num_threads, [=](size_t s, size_t e) {
        for (size_t k = s; k < e; k++) recs[k] = gen_record(first + k, bc_len, umi_len, mode, param, seed);
    });
}

void orc_generate_ascii(uint8_t *ascii, uint64_t first_row, uint64_t n_rows, uint32_t len,
                        uint64_t dirty_ppm, uint64_t lower_ppm, uint64_t seed,
                        size_t num_threads) {
    partitioned(n_rows, num_threads, [=](size_t s, size_t e) {
        for (size_t k = s; k < e; k++) {
            uint64_t key = orc_splitmix64(seed ^ orc_splitmix64(first_row + k));
            uint64_t w = orc_splitmix64(key ^ 4);
            uint64_t rd = orc_splitmix64(key ^ 5), rl = orc_splitmix64(key ^ 6);
            uint8_t *row = ascii + k * len;
            unpack_word(w, len, row);
            if (rl % 1000000ull < lower_ppm)
                for (uint32_t j = 0; j < len; j++) row[j] |= 0x20;
            if (rd % 1000000ull < dirty_ppm) row[(rd >> 40) % len] = 'N';
        }
    });
}

}  // extern "C"
