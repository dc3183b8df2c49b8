// barcode_count.cu — K4: per-barcode record / distinct-UMI table (placeholder until the
// streaming segmented reduction lands; the symbols exist so the ABI is complete).
#include "ctx.h"
#include "kernels.cuh"

using namespace ibu;

extern "C" {

int ibu_gpu_barcode_count(ibu_gpu_ctx_t *ctx, const ibu_record_t *d_records, uint64_t n, int mode,
                          ibu_barcode_table_t *table, void *stream, ibu_error_t *err) {
    clear_error(err);
    return set_error(err, IBU_ERR_ARG, 0, 0, 0, "ibu_gpu_barcode_count: not implemented yet");
}

void ibu_gpu_table_free(ibu_gpu_ctx_t *ctx, ibu_barcode_table_t *table) {
    if (!ctx || !table) return;
    if (table->d_rows) ibu_gpu_free(ctx, table->d_rows);
    table->d_rows = nullptr;
    table->n_rows = 0;
}

}  // extern "C"
