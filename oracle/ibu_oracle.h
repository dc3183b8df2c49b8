/*
 * ibu_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's bulk record path (noamteyssier/ibu),
 * used as the parity checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under ibu_b200/
 * may include, link or call this.
 *
 * Parity status:
 *   PINNED by the reference's own tests (see tests/test_oracle_kat.py): layout,
 *   header validation, file size rule, MmapReader::new/len/header/slice,
 *   process_parallel partitioning/batching and the count+sum processor,
 *   load_to_vec, Writer byte stream.
 *   PARITY UNPINNED: 2-bit pack/unpack (bitnuc is not a dependency of the
 *   reference and its source is not available offline; the convention is the
 *   one documented at src/constructs/record.rs:19-27 plus bitnuc's published
 *   LSB-first order), record-word validation and the per-barcode distinct-UMI
 *   table (new semantics defined by BASELINE.json / SURVEY.md §8a).
 */
#ifndef IBU_ORACLE_H
#define IBU_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t magic, version, bc_len, umi_len;
    uint64_t flags;
    uint8_t reserved[8];
} orc_header_t;
typedef struct {
    uint64_t barcode, umi, index;
} orc_record_t;

/* error codes numbered like the IbuError variants (src/error.rs:56-128) */
enum {
    ORC_OK = 0, ORC_IO = 1, ORC_NIFFLER = 2, ORC_INVALID_MAGIC = 3, ORC_TRUNCATED = 4,
    ORC_INVALID_VERSION = 5, ORC_INVALID_BC_LEN = 6, ORC_INVALID_UMI_LEN = 7,
    ORC_INVALID_MAP_SIZE = 8, ORC_INVALID_INDEX = 9, ORC_PROCESS = 10
};
typedef struct {
    int32_t code, sys;
    uint64_t a, b;
} orc_error_t;

typedef struct {
    uint64_t n_records, sum_barcode, sum_umi, sum_index, xor_all;
    uint64_t n_bad_barcode, n_bad_umi, n_bad_records;
} orc_reduce_t;

typedef struct {
    uint64_t barcode, n_records, n_distinct_umi;
} orc_barcode_row_t;

void orc_header_new(orc_header_t *h, uint32_t bc_len, uint32_t umi_len);
int orc_header_validate(const orc_header_t *h, orc_error_t *err);

typedef struct orc_mmap orc_mmap_t;
int orc_mmap_open(const char *path, orc_mmap_t **out, orc_error_t *err);
void orc_mmap_close(orc_mmap_t *m);
size_t orc_mmap_len(const orc_mmap_t *m);
orc_header_t orc_mmap_header(const orc_mmap_t *m);
int orc_mmap_slice(const orc_mmap_t *m, size_t start, size_t end, const orc_record_t **out,
                   size_t *n, orc_error_t *err);
int orc_load_to_vec(const char *path, orc_header_t *h, orc_record_t **records, size_t *n,
                    orc_error_t *err);
void orc_free(void *p);
/* first element of Reader::next on a byte buffer: 0 ok, ORC_TRUNCATED with pos (reader.rs:218-242) */
int orc_stream_first(const uint8_t *bytes, size_t len, orc_record_t *rec, orc_error_t *err);

/* Writer restatement; mode 0 = write_record per record, 1 = one write_batch */
int orc_write_file(const char *path, const orc_header_t *h, const orc_record_t *recs, size_t n,
                   int mode, orc_error_t *err);

int orc_num_cpus(void);
/* process_parallel (mmap.rs:286-332) with the built-in processors;
 * per-thread trace (nullable, capacity max_threads): records seen and
 * on_batch_complete calls of each spawned thread, n_threads_out = spawned. */
typedef struct {
    uint64_t start, end, records, batches;
} orc_thread_trace_t;
int orc_process_parallel_reduce(const orc_mmap_t *m, size_t num_threads, orc_reduce_t *out,
                                orc_thread_trace_t *trace, size_t max_threads,
                                size_t *n_threads_out, orc_error_t *err);
/* ErrorProcessor of parallel.rs:338-352: fails on record.index == fail_index */
int orc_process_parallel_fail(const orc_mmap_t *m, size_t num_threads, uint64_t fail_index,
                              orc_error_t *err);
/* user-callback form (generic ParallelProcessor); callbacks may be NULL */
typedef struct {
    void *(*clone)(void *self);
    int (*process_record)(void *self, const orc_record_t *rec);
    int (*on_batch_complete)(void *self);
    void (*drop)(void *self);
} orc_processor_vtable_t;
int orc_process_parallel(const orc_mmap_t *m, const orc_processor_vtable_t *vt, void *proc,
                         size_t num_threads, orc_error_t *err);
/* barcode histogram processor (parallel.rs:79-98) + distinct UMIs; rows sorted by barcode */
int orc_process_parallel_barcodes(const orc_mmap_t *m, size_t num_threads,
                                  orc_barcode_row_t **rows, size_t *n_rows, orc_error_t *err);

/* the same reductions over an in-memory array, same partition/batch rule, n threads */
void orc_reduce_records(const orc_record_t *recs, size_t n, uint32_t bc_len, uint32_t umi_len,
                        size_t num_threads, orc_reduce_t *out);
int orc_barcode_table(const orc_record_t *recs, size_t n, orc_barcode_row_t **rows,
                      size_t *n_rows, uint64_t *n_pairs);

/* scalar codec (record.rs:19-27; bitnuc LSB-first) */
int orc_valid_word(uint64_t w, uint32_t len);
void orc_unpack_word(uint64_t w, uint32_t len, uint8_t *out);
/* returns 0 when every byte is one of ACGTacgt, 1 otherwise; *w always written */
int orc_pack_word(const uint8_t *s, uint32_t len, uint64_t *w);
/* batch forms run through the process_parallel partition rule with num_threads threads */
void orc_unpack_records(const orc_record_t *recs, size_t n, uint32_t bc_len, uint32_t umi_len,
                        uint8_t *bc_ascii, uint8_t *umi_ascii, uint8_t *flags,
                        size_t num_threads, orc_reduce_t *out);
void orc_pack_records(const uint8_t *bc_ascii, const uint8_t *umi_ascii, const uint64_t *index,
                      uint64_t index_base, size_t n, uint32_t bc_len, uint32_t umi_len,
                      orc_record_t *recs, uint8_t *flags, size_t num_threads, orc_reduce_t *out);

/* synthetic generators (spec in DESIGN.md §Synthetic data) */
uint64_t orc_splitmix64(uint64_t x);
void orc_generate_records(orc_record_t *recs, uint64_t first, uint64_t n, uint32_t bc_len,
                          uint32_t umi_len, int mode, uint64_t param, uint64_t seed,
                          size_t num_threads);
void orc_generate_ascii(uint8_t *ascii, uint64_t first_row, uint64_t n_rows, uint32_t len,
                        uint64_t dirty_ppm, uint64_t lower_ppm, uint64_t seed,
                        size_t num_threads);

#ifdef __cplusplus
}
#endif
#endif
