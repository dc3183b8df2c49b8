/*
 * ibu_b200.h — C ABI of the B200-native bulk record path of the `ibu` format.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * A Rust `ibu` fork binds these symbols with an `extern "C"` block (see
 * INTEGRATION.md).  Each entry point names the reference interface it replaces
 * or extends (paths relative to the reference crate root).
 *
 * Conventions
 *   - every function returns an ibu_status (0 = ok); when `err` is non-NULL it
 *     receives the payload of the corresponding IbuError variant
 *     (src/error.rs:56-128).  Nothing aborts or throws across this boundary.
 *   - record-level invalidity (barcode/UMI word wider than bc_len/umi_len,
 *     non-ACGT base on pack) is DATA (counters, flags), never an error:
 *     the reference's own generator writes such records (examples/random.rs:46).
 *   - `d_*` pointers are device pointers on the context's GPU, `h_*` host.
 *     `stream` is a cudaStream_t passed as void* (NULL = the context's stream).
 *     Functions ending in `_async` only enqueue work; the others block.
 *   - there is no CPU fallback: GPU entry points return IBU_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef IBU_B200_H
#define IBU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ format */

#define IBU_MAGIC 0x21554249u /* "IBU!" little-endian; src/constructs/header.rs:5 */
#define IBU_VERSION 2u        /* src/constructs/header.rs:6 */
#define IBU_HEADER_SIZE 32u   /* src/constructs/header.rs:7 */
#define IBU_RECORD_SIZE 24u   /* src/constructs/record.rs:3 */
#define IBU_BATCH_SIZE (1024u * 1024u) /* src/io/mmap.rs:284 */
#define IBU_FLAG_SORTED 1ull  /* src/constructs/header.rs:111-113 */

/* byte-identical to `#[repr(C)] struct Header` (src/constructs/header.rs:44-61) */
typedef struct ibu_header {
    uint32_t magic;
    uint32_t version;
    uint32_t bc_len;
    uint32_t umi_len;
    uint64_t flags;
    uint8_t reserved[8];
} ibu_header_t;

/* byte-identical to `#[repr(C)] struct Record` (src/constructs/record.rs:58-66) */
typedef struct ibu_record {
    uint64_t barcode;
    uint64_t umi;
    uint64_t index;
} ibu_record_t;

/* ------------------------------------------------------------------ errors */

/* 1:1 with the IbuError variants (src/error.rs:56-128) + device-side codes */
typedef enum ibu_status {
    IBU_OK = 0,
    IBU_ERR_IO = 1,                     /* Io(std::io::Error): sys = errno */
    IBU_ERR_NIFFLER = 2,                /* kept for numbering; never produced */
    IBU_ERR_INVALID_MAGIC = 3,          /* a = expected, b = actual */
    IBU_ERR_TRUNCATED_RECORD = 4,       /* a = pos */
    IBU_ERR_INVALID_VERSION = 5,        /* a = expected, b = actual */
    IBU_ERR_INVALID_BARCODE_LENGTH = 6, /* a = bc_len */
    IBU_ERR_INVALID_UMI_LENGTH = 7,     /* a = umi_len */
    IBU_ERR_INVALID_MAP_SIZE = 8,
    IBU_ERR_INVALID_INDEX = 9,          /* a = idx, b = max */
    IBU_ERR_PROCESS = 10,               /* Process(Box<dyn Error>): callback failure */
    IBU_ERR_CUDA = 11,                  /* sys = cudaError_t */
    IBU_ERR_NCCL = 12,                  /* sys = ncclResult_t (IBU_EXCHANGE_NCCL only) */
    IBU_ERR_ARG = 13,                   /* NULL / misaligned / out-of-range argument */
    IBU_ERR_NOMEM = 14
} ibu_status;

typedef struct ibu_error {
    int32_t code; /* ibu_status */
    int32_t sys;  /* errno or cudaError_t */
    uint64_t a;
    uint64_t b;
    char msg[232];
} ibu_error_t;

/* Display string of the variant, formatted like src/error.rs:56-128. */
const char *ibu_strerror(int code);
const char *ibu_version(void);

/* ------------------------------------------------------------------ header */

/* Header::new (src/constructs/header.rs:84-93) */
void ibu_header_init(ibu_header_t *h, uint32_t bc_len, uint32_t umi_len);
/* Header::set_sorted / Header::sorted (header.rs:111-113, 130-132) */
void ibu_header_set_sorted(ibu_header_t *h);
int ibu_header_sorted(const ibu_header_t *h);
/* Header::validate (header.rs:167-187): magic, version, bc_len, umi_len — first failure wins */
int ibu_header_validate(const ibu_header_t *h, ibu_error_t *err);

/* ------------------------------------------------------------- mmap reader */

typedef struct ibu_mmap_reader ibu_mmap_reader_t;

/* MmapReader::new (src/io/mmap.rs:143-161).  A file shorter than 32 bytes
 * (a panic in the reference) is reported as IBU_ERR_IO. */
int ibu_mmap_open(const char *path, ibu_mmap_reader_t **out, ibu_error_t *err);
/* Clone = Arc bump (mmap.rs:99-107); each clone is closed separately. */
ibu_mmap_reader_t *ibu_mmap_clone(ibu_mmap_reader_t *r);
void ibu_mmap_close(ibu_mmap_reader_t *r);
/* MmapReader::len / header (mmap.rs:178-180, 201-203) */
size_t ibu_mmap_len(const ibu_mmap_reader_t *r);
ibu_header_t ibu_mmap_header(const ibu_mmap_reader_t *r);
/* MmapReader::slice (mmap.rs:253-270): zero-copy view, valid until the last clone closes. */
int ibu_mmap_slice(const ibu_mmap_reader_t *r, size_t start, size_t end,
                   const ibu_record_t **out, size_t *n_out, ibu_error_t *err);

/* load_to_vec (src/io/reader.rs:510-535).  *records is released with ibu_free. */
int ibu_load_to_vec(const char *path, ibu_header_t *header, ibu_record_t **records,
                    size_t *n, ibu_error_t *err);
void ibu_free(void *p);

/* Contiguous range of shard `rank` of `world` over `len` records, the
 * partition rule of process_parallel (mmap.rs:297-307): len/world each, the
 * last shard takes the remainder. */
void ibu_shard_range(uint64_t len, uint32_t rank, uint32_t world, uint64_t *start,
                     uint64_t *end);

/* ------------------------------------------------------------------ writer */

typedef struct ibu_writer ibu_writer_t;

/* Writer::from_path / Writer::new (src/io/writer.rs:129-143, 525-536): header
 * bytes are written immediately and NOT validated (as in the reference). */
int ibu_writer_open(const char *path, const ibu_header_t *header, ibu_writer_t **out,
                    ibu_error_t *err);
/* Writer::new_headless (writer.rs:169-179) onto a file (append = 1 opens O_APPEND). */
int ibu_writer_open_headless(const char *path, int append, ibu_writer_t **out,
                             ibu_error_t *err);
/* Writer::write_record (writer.rs:260-273) */
int ibu_writer_write_record(ibu_writer_t *w, const ibu_record_t *rec, ibu_error_t *err);
/* Writer::write_batch (writer.rs:315-351): batches larger than the 48 Ki-record buffer bypass it */
int ibu_writer_write_batch(ibu_writer_t *w, const ibu_record_t *recs, size_t n,
                           ibu_error_t *err);
/* Writer::records_written (writer.rs:207-209) */
uint64_t ibu_writer_records_written(const ibu_writer_t *w);
/* Writer::finish (writer.rs:429-433) */
int ibu_writer_finish(ibu_writer_t *w, ibu_error_t *err);
/* Drop (writer.rs:519-523): finish().ok() then release */
void ibu_writer_close(ibu_writer_t *w);

/* ------------------------------------------------------------- GPU context */

typedef struct ibu_gpu_ctx ibu_gpu_ctx_t;

typedef struct ibu_gpu_config {
    uint32_t chunk_records; /* records per staged chunk (and per on_chunk call); 0 = IBU_BATCH_SIZE */
    uint32_t n_slots;       /* chunk slots (streams) in flight; 0 = 3 */
    uint32_t copy_threads;  /* host threads for pageable->pinned staging; 0 = auto: all cores, divided
                             by LOCAL_WORLD_SIZE when a launcher (torchrun) sets it */
    uint32_t reserved;
} ibu_gpu_config_t;

int ibu_gpu_device_count(void);
/* One context per GPU (one process per GPU under torchrun; several contexts
 * per process also work).  cfg may be NULL. */
int ibu_gpu_ctx_create(int device, const ibu_gpu_config_t *cfg, ibu_gpu_ctx_t **out,
                       ibu_error_t *err);
void ibu_gpu_ctx_destroy(ibu_gpu_ctx_t *ctx);
int ibu_gpu_ctx_device(const ibu_gpu_ctx_t *ctx);
int ibu_gpu_ctx_sm_count(const ibu_gpu_ctx_t *ctx);
/* number of kernels this library has launched in this process (all contexts) */
uint64_t ibu_gpu_launch_count(void);
int ibu_gpu_synchronize(ibu_gpu_ctx_t *ctx, void *stream, ibu_error_t *err);

/* device / pinned-host memory helpers for callers without a CUDA runtime of their own */
int ibu_gpu_malloc(ibu_gpu_ctx_t *ctx, size_t bytes, void **d_out, ibu_error_t *err);
void ibu_gpu_free(ibu_gpu_ctx_t *ctx, void *d_ptr);
int ibu_gpu_memcpy_h2d(ibu_gpu_ctx_t *ctx, void *d_dst, const void *h_src, size_t bytes,
                       ibu_error_t *err);
int ibu_gpu_memcpy_d2h(ibu_gpu_ctx_t *ctx, void *h_dst, const void *d_src, size_t bytes,
                       ibu_error_t *err);
int ibu_gpu_memset(ibu_gpu_ctx_t *ctx, void *d_dst, int value, size_t bytes, ibu_error_t *err);
/* any direction (device<->device included), blocking */
int ibu_gpu_memcpy(ibu_gpu_ctx_t *ctx, void *dst, const void *src, size_t bytes, ibu_error_t *err);
int ibu_host_alloc(size_t bytes, void **h_out, ibu_error_t *err); /* pinned */
void ibu_host_free(void *h_ptr);
int ibu_host_register(void *h_ptr, size_t bytes, int read_only, ibu_error_t *err);
void ibu_host_unregister(void *h_ptr);
/* Host utility: copy with non-temporal stores on `threads` threads (0 = all cores) — what the
 * pipelines use to fill pinned staging buffers (no read-for-ownership of the destination).
 * Any alignment, any size. */
void ibu_host_stream_copy(void *dst, const void *src, size_t bytes, unsigned threads);

/* ---------------------------------------------------------- device kernels */

/* Result of one pass of the built-in reductions — the device counterpart of
 * the reference processors run through process_parallel (a12 in SURVEY §8a):
 *   n_records                         mmap.rs:350-373 (local_count)
 *   sum_barcode/sum_umi/sum_index     examples/parallel.rs:21-27 (wrapping u64)
 *   xor_all                           examples/roundtrip.rs:84-87
 *   n_bad_barcode / n_bad_umi         word >> 2*len != 0 (len < 32); new semantics
 *   n_bad_records                     records with either word bad (unpack) or
 *                                     any non-ACGT base (pack)
 * mmap.rs's count+sum processor value is sum_barcode+sum_umi+sum_index. */
typedef struct ibu_reduce_result {
    uint64_t n_records;
    uint64_t sum_barcode;
    uint64_t sum_umi;
    uint64_t sum_index;
    uint64_t xor_all;
    uint64_t n_bad_barcode;
    uint64_t n_bad_umi;
    uint64_t n_bad_records;
} ibu_reduce_result_t;

/* K1: validate + reduce over device-resident records (any 8-byte aligned record
 * pointer, e.g. a slice).  *d_result (device) is OVERWRITTEN with this pass's values. */
int ibu_gpu_validate_reduce_async(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records,
                                  uint64_t n, uint32_t bc_len, uint32_t umi_len,
                                  ibu_reduce_result_t *d_result, void *stream,
                                  ibu_error_t *err);

/* K2: 2-bit unpack to ASCII (bitnuc from_2bit convention: base i at bits
 * [2i,2i+1], A=0 C=1 G=2 T=3; src/constructs/record.rs:19-27) fused with
 * validation.  d_bc_ascii is [n][bc_len], d_umi_ascii is [n][umi_len], dense,
 * no terminators, 16-byte aligned bases.  d_flags (nullable) gets one byte per
 * record: bit0 = bad barcode, bit1 = bad umi.  d_result (nullable) is overwritten with
 * the full reduction of K1 (count, sums, xor, invalid-word counters): decode + validate +
 * count is ONE pass over the records. */
int ibu_gpu_unpack_async(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n,
                         uint32_t bc_len, uint32_t umi_len, uint8_t *d_bc_ascii,
                         uint8_t *d_umi_ascii, uint8_t *d_flags,
                         ibu_reduce_result_t *d_result, void *stream, ibu_error_t *err);

/* K3: ASCII -> Record pack (bitnuc as_2bit convention, case-insensitive).
 * d_index nullable: index = index_base + i.  Non-ACGT bytes: the record is
 * flagged (d_flags bit0 = barcode, bit1 = umi; nullable) and counted in
 * d_result->n_bad_*; its code is the branch-free formula's output
 * (c' = (c>>1)&3; code = c' ^ (c'>>1)), which is deterministic. */
int ibu_gpu_pack_async(ibu_gpu_ctx_t *ctx, const uint8_t *d_bc_ascii,
                       const uint8_t *d_umi_ascii, const uint64_t *d_index,
                       uint64_t index_base, uint64_t n, uint32_t bc_len, uint32_t umi_len,
                       ibu_record_t *d_records, uint8_t *d_flags,
                       ibu_reduce_result_t *d_result, void *stream, ibu_error_t *err);

/* K4: per-barcode table — the device form of the HashMap<barcode,count>
 * processor of src/parallel.rs:79-98, extended with distinct-UMI counts.
 * Rows are emitted sorted by barcode (Record's Ord, record.rs:58). */
typedef struct ibu_barcode_row {
    uint64_t barcode;
    uint64_t n_records;
    uint64_t n_distinct_umi;
} ibu_barcode_row_t;

typedef struct ibu_barcode_table {
    ibu_barcode_row_t *d_rows; /* device, owned by the library: ibu_gpu_table_free */
    uint64_t n_rows;
    uint64_t n_records;
    uint64_t n_distinct_pairs; /* distinct (barcode, umi) */
    uint32_t input_was_sorted; /* 1 = the streaming sorted path ran */
    uint32_t reserved;
} ibu_barcode_table_t;

/* Blocking (the table size is data dependent).  d_records must be 32-byte aligned.
 * mode 0 = auto: one streaming pass that also verifies the (barcode, umi) order
 *          (the header's `sorted` flag is advisory: examples/parallel.rs:52-53 sets it on
 *          unsorted data); if the order does not hold, the records are partitioned by a hash of
 *          their (barcode, umi) key and de-duplicated bucket by bucket in shared memory; with
 *          about as many barcodes as records they are partitioned by the barcode's own top bits
 *          instead and every bucket is sorted in shared memory, its rows written in barcode order
 *          (inputs that suit neither — keys wider than 64 bits, a handful of distinct keys, a
 *          few barcodes that hold most of many — are hash-aggregated in a global table or radix
 *          sorted);
 * mode 1 = streaming pass only: unsorted input is not an error, it returns
 *          input_was_sorted = 0 and no rows;
 * mode 2 = skip the streaming attempt.
 * OR-able into mode: IBU_COUNT_WEIGHTED, IBU_COUNT_LENS(bc_len, umi_len), IBU_COUNT_PATH_*. */
int ibu_gpu_barcode_count(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n,
                          int mode, ibu_barcode_table_t *table, void *stream,
                          ibu_error_t *err);
void ibu_gpu_table_free(ibu_gpu_ctx_t *ctx, ibu_barcode_table_t *table);
/* The rows of a device table into caller memory (h_rows: table->n_rows rows; the HashMap the
 * reference's processor ends with, parallel.rs:79-98, as a sorted array).  Pageable destinations are
 * filled through the context's pinned landing area with the multi-threaded copy: a plain D2H into
 * freshly allocated pageable memory is a staged copy through page faults (5-11 ms per 10^6 rows
 * against 1-2 ms).  Blocking; waits for the table's stream work first. */
int ibu_gpu_table_to_host(ibu_gpu_ctx_t *ctx, const ibu_barcode_table_t *table, ibu_barcode_row_t *h_rows,
                          ibu_error_t *err);

/* OR into `mode` of ibu_gpu_barcode_count: the records are a (barcode, umi, multiplicity) pair
 * table (index = how many records the pair stands for), e.g. the concatenation of several
 * shards' ibu_gpu_pair_table outputs; n_records sums the multiplicities. */
#define IBU_COUNT_WEIGHTED 8
/* The header's lengths (Header.bc_len / umi_len, header.rs:48-61), so that the (barcode, umi)
 * key layout need not be detected from a sample.  Words wider than that (invalid per the
 * header) are still counted exactly, on a side path. */
#define IBU_COUNT_LENS(bc_len, umi_len) ((((bc_len) & 0x3F) << 8) | (((umi_len) & 0x3F) << 16))
/* Tuning / tests: force one implementation of the unsorted path (default: chosen from a sample). */
#define IBU_COUNT_PATH_PARTITION 0x10 /* hash partition + shared-memory de-duplication */
#define IBU_COUNT_PATH_SORT 0x20      /* radix sort, then the streaming pass */
#define IBU_COUNT_PATH_LEGACY 0x30    /* global hash table (round-1 path) */

/* De-duplicated (barcode, umi) pairs of a record range with their multiplicities: rows are
 * ibu_record_t {barcode, umi, index = count}.  This is what shards exchange to merge
 * distinct-UMI counts exactly (they are not additive across shards).  `flags`:
 * IBU_PAIRS_WEIGHTED = the input's index words are multiplicities already; IBU_PAIRS_UNORDERED =
 * rows in no particular order (default: sorted by (barcode, umi)); IBU_COUNT_LENS / IBU_COUNT_PATH_*
 * as above.  *d_pairs is released with ibu_gpu_free.  Blocking. */
#define IBU_PAIRS_WEIGHTED 1
#define IBU_PAIRS_UNORDERED 2
int ibu_gpu_pair_table(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, int flags,
                       ibu_record_t **d_pairs, uint64_t *n_pairs, void *stream, ibu_error_t *err);

/* Send side of the multi-GPU pair exchange: groups the rows of a pair table by
 * owner(barcode) = splitmix64(barcode) % world into d_out (bucket r = rows owned by rank r,
 * buckets in rank order, order inside a bucket unspecified) and returns the bucket sizes in
 * h_counts[world] — the split sizes of the all-to-all.  Blocking. */
int ibu_gpu_partition_by_owner(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_pairs, uint64_t n,
                               uint32_t world, ibu_record_t *d_out, uint64_t *h_counts, void *stream,
                               ibu_error_t *err);

/* Device sort by Record's Ord (src/constructs/record.rs:29-32,58): barcode, then umi, then
 * index.  d_sorted (n records) must not alias d_records.  Blocking.  After it, a header with
 * set_sorted (header.rs:111-113) is truthful and ibu_gpu_barcode_count streams in one pass. */
int ibu_gpu_sort_records(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n,
                         ibu_record_t *d_sorted, void *stream, ibu_error_t *err);

/* ------------------------------------------------ synthetic data (benches) */

/* Counter-based generators (splitmix64 of seed and record number) shared by the
 * oracle so that CPU and GPU regenerate identical inputs without files. */
enum {
    IBU_GEN_CLEAN = 0,     /* barcode/umi masked to bc_len/umi_len, index = i */
    IBU_GEN_DIRTY = 1,     /* as CLEAN, `param` ppm of records carry an unmasked word */
    IBU_GEN_PATTERN = 2,   /* (i % 1e6, 31 i % 1e6, i): examples/parallel.rs:65-69 */
    IBU_GEN_WHITELIST = 3, /* param: low 32 bits = #distinct barcodes, high 32 = umi space; unsorted */
    IBU_GEN_SORTED = 4,    /* param: low 32 bits = records per barcode, high 32 = records per umi;
                              sorted by Record's Ord: (i / rpb, (i % rpb) / dup, i) */
    IBU_GEN_ZIPF = 5       /* as WHITELIST, barcode ranks log-uniform (rank r about as likely as 1/r):
                              a few cells hold most records, like a real library; unsorted */
};
int ibu_gpu_generate_records_async(ibu_gpu_ctx_t *ctx, ibu_record_t *d_records,
                                   uint64_t first, uint64_t n, uint32_t bc_len,
                                   uint32_t umi_len, int mode, uint64_t param, uint64_t seed,
                                   void *stream, ibu_error_t *err);
/* ASCII rows uniform over ACGT; `dirty_ppm` of bytes become 'N', `lower_ppm` lower-case. */
int ibu_gpu_generate_ascii_async(ibu_gpu_ctx_t *ctx, uint8_t *d_ascii, uint64_t first_row,
                                 uint64_t n_rows, uint32_t len, uint64_t dirty_ppm,
                                 uint64_t lower_ppm, uint64_t seed, void *stream,
                                 ibu_error_t *err);

/* ---------------------------------------------- host-buffer (end-to-end) path */

/* GPU counterpart of MmapReader::process_parallel (src/io/mmap.rs:286-332) for
 * the built-in reductions: records [start,end) of the reader are staged chunk
 * by chunk through pinned buffers on double-buffered streams, validated and
 * reduced on device.  `on_chunk` (nullable) is the on_batch_complete analogue
 * (parallel.rs:141-151): called on the calling thread after each chunk's result
 * has landed, in chunk order, with that chunk's result; a non-zero return
 * aborts with IBU_ERR_PROCESS.  A chunk is IBU_BATCH_SIZE records unless the context
 * was configured otherwise — the reference's own batch (mmap.rs:284, 322-326).
 * The callback runs while the context's chunk slots are in use: host-buffer entry
 * points of the SAME context called from it (process_*, unpack_host, pack_host,
 * load_to_device, write_records, table_to_host, stream_open) return IBU_ERR_ARG;
 * device-pointer (_async) calls and other contexts are fine. */
typedef int (*ibu_chunk_cb)(void *user, uint64_t chunk_start, uint64_t chunk_n,
                            const ibu_reduce_result_t *chunk_result);
int ibu_gpu_process_mmap(ibu_gpu_ctx_t *ctx, const ibu_mmap_reader_t *reader, uint64_t start,
                         uint64_t end, ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk,
                         void *user, ibu_error_t *err);
/* Optional: page-lock the reader's mapping (cudaHostRegister, read-only) so that
 * ibu_gpu_process_mmap DMAs straight out of the page cache at the link rate instead of staging
 * through pinned bounce buffers.  Locking costs about 0.1-0.5 s per GB (it faults every page
 * in), so it pays when the same reader is processed more than once; a single pass is faster
 * staged.  Fails with IBU_ERR_CUDA / IBU_ERR_IO where the driver or the kernel refuses; the
 * staged path then still applies. */
int ibu_mmap_pin(ibu_mmap_reader_t *reader, ibu_error_t *err);
void ibu_mmap_unpin(ibu_mmap_reader_t *reader);
/* The same for the pages that hold records [start, end) only — what one rank of a range-sharded
 * job locks (mmap.rs:297-307).  Pins are counted per range and shared by the clones of a reader;
 * whatever is still pinned is released when the last clone closes. */
int ibu_mmap_pin_range(ibu_mmap_reader_t *reader, uint64_t start, uint64_t end, ibu_error_t *err);
void ibu_mmap_unpin_range(ibu_mmap_reader_t *reader, uint64_t start, uint64_t end);
/* Same over a host array (pinned: copied directly; pageable: staged). */
int ibu_gpu_process_host(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n,
                         uint32_t bc_len, uint32_t umi_len, ibu_reduce_result_t *h_result,
                         ibu_chunk_cb on_chunk, void *user, ibu_error_t *err);

/* The same with an operation mask — ONE pass over the file for everything the north star lists:
 *   IBU_OP_REDUCE  validate + the built-in reductions (always on; the result of every call)
 *   IBU_OP_UNPACK  2-bit unpack of every record to ASCII in host memory (K2 instead of K1 per chunk)
 *   IBU_OP_TABLE   the per-barcode record / distinct-UMI table of the whole range: the
 *                  HashMap<barcode, count> processor of src/parallel.rs:79-98 driven by
 *                  process_parallel (mmap.rs:312-320), + distinct UMIs.  Each chunk's keys are
 *                  inserted on the chunk's stream while the next chunk is on the link; only the
 *                  de-duplication and the rows remain after the last chunk.
 *   IBU_OP_KEEP    the records stay on the device (*d_records, release with ibu_gpu_free)
 * table_mode: OR-able bits of ibu_gpu_barcode_count's mode (the header's lengths are filled in). */
#define IBU_OP_REDUCE 1u
#define IBU_OP_TABLE 2u
#define IBU_OP_KEEP 4u
#define IBU_OP_UNPACK 8u
typedef struct ibu_process_request {
    uint32_t ops;
    int32_t table_mode;
    ibu_barcode_table_t *table;  /* out: IBU_OP_TABLE (rows on the device: ibu_gpu_table_free) */
    ibu_record_t **d_records;    /* out: IBU_OP_KEEP */
    uint8_t *h_bc_ascii;         /* out: IBU_OP_UNPACK, [n][bc_len] host bytes */
    uint8_t *h_umi_ascii;        /* out: IBU_OP_UNPACK, [n][umi_len] */
    uint8_t *h_flags;            /* out: IBU_OP_UNPACK, nullable, [n] */
} ibu_process_request_t;
int ibu_gpu_process_mmap_ops(ibu_gpu_ctx_t *ctx, const ibu_mmap_reader_t *reader, uint64_t start,
                             uint64_t end, const ibu_process_request_t *req,
                             ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk, void *user,
                             ibu_error_t *err);
int ibu_gpu_process_host_ops(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n,
                             uint32_t bc_len, uint32_t umi_len, const ibu_process_request_t *req,
                             ibu_reduce_result_t *h_result, ibu_chunk_cb on_chunk, void *user,
                             ibu_error_t *err);

/* ------------------------------------------------------------- several GPUs */

/* The GPUs of one box behind one handle (SURVEY §8b `ibu_gpu_ctx_create(devices[], n, cfg)`, §8e).
 * Records shard across them by contiguous range with the partition rule of process_parallel
 * (src/io/mmap.rs:297-307: len / n each, the last rank takes the remainder — ibu_shard_range);
 * one host thread per GPU drives its context, all staging copies share the process-wide worker
 * pool and the one mapping.  The 8-word results are merged on the host exactly like the reference
 * processors' on_batch_complete merge (mmap.rs:365-372); the per-barcode tables are merged
 * exactly by exchanging the shards' de-duplicated (barcode, umi) pairs by owner(barcode) — the
 * path's one exchange step — see `exchange`.  The same device may be listed more than once
 * (ranks then share it; useful for tests on a single GPU; not with IBU_EXCHANGE_NCCL). */
typedef struct ibu_gpu_group ibu_gpu_group_t;
int ibu_gpu_group_create(const int *devices, uint32_t n_devices, const ibu_gpu_config_t *cfg,
                         ibu_gpu_group_t **out, ibu_error_t *err);
void ibu_gpu_group_destroy(ibu_gpu_group_t *g);
uint32_t ibu_gpu_group_size(const ibu_gpu_group_t *g);
/* the context of rank `rank` (owned by the group), e.g. for ibu_gpu_free of kept shards */
ibu_gpu_ctx_t *ibu_gpu_group_ctx(ibu_gpu_group_t *g, uint32_t rank);

#define IBU_EXCHANGE_AUTO 0u /* the fastest measured on NVLink boxes: P2P */
#define IBU_EXCHANGE_P2P 1u  /* owners pull with cudaMemcpyPeerAsync (NVLink / NVSwitch) */
#define IBU_EXCHANGE_HOST 2u /* through pinned host memory (D2H, then H2D by the owner) */
#define IBU_EXCHANGE_NCCL 3u /* ncclSend / ncclRecv; libnccl.so.2 is loaded at run time, IBU_ERR_NCCL if absent */

/* merged table in host memory: h_rows is released with ibu_free */
typedef struct ibu_host_table {
    ibu_barcode_row_t *h_rows; /* sorted by barcode */
    uint64_t n_rows;
    uint64_t n_records;
    uint64_t n_distinct_pairs;
} ibu_host_table_t;

/* where the time of a group call went (milliseconds, the slowest rank of each phase) */
typedef struct ibu_group_timing {
    double ingest_ms;   /* staging + H2D + per-chunk kernels (+ the shard's pair de-duplication) */
    double local_ms;    /* pair table of a resident shard (ibu_gpu_group_barcode_count only) */
    double exchange_ms; /* grouping by owner + the pair exchange */
    double owner_ms;    /* weighted count on the owners */
    double gather_ms;   /* rows to rank 0, barcode order, device -> host */
    double table_ms;    /* local + exchange + owner + gather as one span */
    double total_ms;
    uint64_t pairs_local; /* de-duplicated pairs before the exchange, all ranks */
    uint64_t bytes_sent;  /* most bytes one rank sent to other ranks */
    uint32_t exchange;    /* the exchange that ran */
    uint32_t reserved;
} ibu_group_timing_t;

typedef struct ibu_group_request {
    uint32_t ops;               /* IBU_OP_REDUCE | IBU_OP_TABLE | IBU_OP_KEEP */
    int32_t table_mode;         /* as ibu_process_request_t */
    uint32_t exchange;          /* IBU_EXCHANGE_* */
    uint32_t reserved;
    ibu_host_table_t *table;    /* out: IBU_OP_TABLE */
    ibu_record_t **d_records;   /* out[size]: IBU_OP_KEEP, shard r on the device of rank r */
    uint64_t *shard_records;    /* out[size], nullable: records of each shard */
    ibu_group_timing_t *timing; /* out, nullable */
} ibu_group_request_t;

/* GPU counterpart of process_parallel across the group: rank r processes
 * ibu_shard_range(end - start, r, size) of the range. */
int ibu_gpu_group_process_mmap(ibu_gpu_group_t *g, const ibu_mmap_reader_t *reader, uint64_t start,
                               uint64_t end, const ibu_group_request_t *req,
                               ibu_reduce_result_t *h_result, ibu_error_t *err);
int ibu_gpu_group_process_host(ibu_gpu_group_t *g, const ibu_record_t *h_records, uint64_t n,
                               uint32_t bc_len, uint32_t umi_len, const ibu_group_request_t *req,
                               ibu_reduce_result_t *h_result, ibu_error_t *err);
/* The exact table of device-resident shards (d_shards[r] on the device of rank r, 32-byte aligned;
 * mode as ibu_gpu_barcode_count without IBU_COUNT_WEIGHTED). */
int ibu_gpu_group_barcode_count(ibu_gpu_group_t *g, const ibu_record_t *const *d_shards,
                                const uint64_t *shard_records, int mode, uint32_t exchange,
                                ibu_host_table_t *table, ibu_group_timing_t *timing,
                                ibu_error_t *err);

/* Streaming ingest — the Reader<R> of src/io/reader.rs feeding the GPU.
 * The bytes of an .ibu stream (header first) are pushed in pieces of any size, from any source
 * (stdin, a decompressor, a socket): the first 32 bytes are parsed and validated like
 * Reader::new (reader.rs:152-176); record bytes are gathered into pinned chunks and each full
 * chunk goes H2D -> K1 on the next slot's stream while the caller keeps pushing.  One stream
 * (or host-buffer call) per context at a time: open fails with IBU_ERR_ARG while another is
 * active.  finish() flushes the last partial chunk and returns the merged result; trailing
 * bytes that do not make a whole record give IBU_ERR_TRUNCATED_RECORD with a = byte position
 * of the incomplete record (reader.rs:232-237), a stream shorter than a header IBU_ERR_IO. */
typedef struct ibu_gpu_stream ibu_gpu_stream_t;
int ibu_gpu_stream_open(ibu_gpu_ctx_t *ctx, ibu_gpu_stream_t **out, ibu_error_t *err);
int ibu_gpu_stream_push(ibu_gpu_stream_t *st, const void *bytes, size_t len, ibu_error_t *err);
/* valid once 32 bytes have been pushed; returns IBU_ERR_ARG before that */
int ibu_gpu_stream_header(const ibu_gpu_stream_t *st, ibu_header_t *header, ibu_error_t *err);
int ibu_gpu_stream_finish(ibu_gpu_stream_t *st, ibu_reduce_result_t *h_result, ibu_error_t *err);
void ibu_gpu_stream_close(ibu_gpu_stream_t *st);

/* Device path of load_to_vec (src/io/reader.rs:510-535): header validated, size
 * checked, records [start,end) of the file land in one device allocation
 * (*d_records, release with ibu_gpu_free).  end = UINT64_MAX means "to the end". */
int ibu_gpu_load_to_device(ibu_gpu_ctx_t *ctx, const char *path, uint64_t start, uint64_t end,
                           ibu_header_t *header, ibu_record_t **d_records, uint64_t *n,
                           ibu_error_t *err);

/* Writer device path (src/io/writer.rs:315-351 write_batch fed from HBM): device-resident
 * records are brought back through the pinned slot buffers (D2H on alternating streams) and
 * appended with ibu_writer_write_batch semantics while the next chunk is in flight.  With
 * ibu_gpu_sort_records this turns any file into a truthfully `sorted` one. */
int ibu_gpu_write_records(ibu_gpu_ctx_t *ctx, ibu_writer_t *writer, const ibu_record_t *d_records,
                          uint64_t n, ibu_error_t *err);

/* End-to-end unpack: host records -> host ASCII, pipelined H2D / K2 / D2H. */
int ibu_gpu_unpack_host(ibu_gpu_ctx_t *ctx, const ibu_record_t *h_records, uint64_t n,
                        uint32_t bc_len, uint32_t umi_len, uint8_t *h_bc_ascii,
                        uint8_t *h_umi_ascii, uint8_t *h_flags, ibu_reduce_result_t *h_result,
                        ibu_error_t *err);
/* End-to-end pack: host ASCII -> host records (the Writer::write_batch source). */
int ibu_gpu_pack_host(ibu_gpu_ctx_t *ctx, const uint8_t *h_bc_ascii, const uint8_t *h_umi_ascii,
                      const uint64_t *h_index, uint64_t index_base, uint64_t n, uint32_t bc_len,
                      uint32_t umi_len, ibu_record_t *h_records, uint8_t *h_flags,
                      ibu_reduce_result_t *h_result, ibu_error_t *err);

#ifdef __cplusplus
}
#endif
#endif /* IBU_B200_H */
