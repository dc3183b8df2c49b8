// barcode_agg.cu — K4 for unsorted inputs: partition, then aggregate on chip.
//
// The per-barcode table (src/parallel.rs:79-98 HashMap<barcode, count>, extended with the number
// of distinct UMI words) needs every (barcode, umi) pair de-duplicated.  A global hash table is
// one random DRAM access per record once it outgrows L2, and a full sort moves every record
// once per digit.  This path reads a record once and its 8-byte key two or three times:
//
//   k_sample        ~256 K records at hashed positions: bit widths of the two words, distinct pairs
//                   and distinct barcodes among them (sizes every table below; tiny)
//   k_part1         24 B read + 8 B written per record.  A record whose words fit the key layout
//                   (barcode < 2^bb, umi < 2^ub, bb + ub <= 64) becomes ONE u64:
//                   key = mix64(barcode << ub | umi), mix64 a bijection, so the key's top bits
//                   are a uniform hash and the pair can be recovered from it.  A CTA groups the
//                   keys of its 4096-record tile by their top pb1 (<= 8) bits in shared memory
//                   (rank = shared-memory atomicAdd(+1) on the tile's histogram) and appends each
//                   group to its level-1 bucket as one run (one global atomicAdd per tile and
//                   bucket) — 16 keys = a full 128-byte line at fan-out 256.  Records that do
//                   not fit (`wide`: the reference's own generator writes them,
//                   examples/random.rs:46) are copied to a side list.
//   k_part2         8 B read + 8 B written per key: the same on the keys of each level-1 bucket
//                   with the next pb2 (<= 9) bits, down to 2^17 buckets of ~763 keys per 10^8
//                   records (three levels above 2^17 buckets).  One level straight to 2^17 buckets
//                   was the first form of this path: one L2 write transaction per 8-byte key,
//                   2.1 ms per 10^8 records against 0.63 + 0.46 ms for the two staged levels.
//   k_bucket_dedup2 8 B read per key.  One CTA per bucket folds the bucket's keys into a
//                   shared-memory hash table (distinct keys per bucket are balanced by the hash,
//                   however skewed the barcodes are; duplicates only make a bucket longer) and
//                   adds every distinct pair to the per-barcode table — a global open-addressing
//                   table sized from the sample, L2 resident for realistic barcode counts — or,
//                   for pair tables, appends (barcode, umi, multiplicity).
//   wide list       de-duplicated by the legacy path, then added to the same table.
//   k_table_rows    occupied slots -> rows, radix sorted by barcode (#barcodes rows, not #records).
//
// Inputs this does not suit fall back to the legacy path (barcode_count.cu): key layouts wider
// than 64 bits, almost no distinct keys (a tiny global table is L2 resident anyway), nearly as
// many barcodes as records (the rows would need a full-size sort at the end).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <utility>

#include "k4.h"
#include "k4_partition.cuh"
#include "kernels.cuh"

namespace ibu {

using namespace k4p;

namespace {

// ---------------------------------------------------------------------------------------- host

uint64_t pow2_ceil(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
uint32_t log2_of(uint64_t pow2) {
    uint32_t l = 0;
    while ((1ull << l) < pow2) l++;
    return l;
}

// Chao1 lower bound on the number of distinct values behind a sample that showed `seen` of them,
// f1 exactly once and f2 exactly twice.  A sample without repeats says nothing: `unknown`.
double chao1(double seen, double f1, double f2, double unknown) {
    if (f2 < 1.0) return f1 >= seen && seen > 64 ? unknown : seen + f1 * (f1 - 1.0) / 2.0;
    return seen + f1 * f1 / (2.0 * f2);
}

// smallest width (in bits, rounded up to even) that holds all but `slack` words of the sample
uint32_t width_of(const uint32_t *hist, uint64_t slack) {
    uint64_t above = 0;
    uint32_t w = 64;
    for (; w > 0; w--) {
        above += hist[w];
        if (above > slack) break;
    }
    w = std::max(w, 1u);
    return (w + 1u) & ~1u;
}

template <class K>
int set_max_smem(K kern, int device, size_t bytes, ibu_error_t *err) {
    // once per kernel and device (the attribute is per function, not per launch)
    // (all instantiations share this function's statics: the key is the kernel's address)
    static std::mutex m;
    static std::set<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lock(m);
    const auto key = std::make_pair((const void *)kern, device);
    if (done.count(key)) return IBU_OK;
    IBU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done.insert(key);
    return IBU_OK;
}

#define IBU_LAUNCHED(name)                                                   \
    do {                                                                     \
        g_launches.fetch_add(1, std::memory_order_relaxed);                  \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) return ibu::cuda_fail(err, e__, name);       \
    } while (0)

// IBU_B200_TRACE=1: device time of each stage (synchronises after every stage: tuning only)
struct StageTimer {
    bool on;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(bool enabled, cudaStream_t stream) : on(enabled), s(stream) {
        if (on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, s);
        }
    }
    ~StageTimer() {
        if (on) {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    }
    void lap(const char *what) {
        if (!on) return;
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        fprintf(stderr, "[ibu trace]   %-26s %8.3f ms (device)\n", what, ms);
        cudaEventRecord(a, s);
    }
};

constexpr uint32_t kOrdCap = 2048;  // keys of one bucket the ordered kernels hold in shared memory
constexpr size_t kDedupMaxSmem = 8192 * 18;  // 8 Ki slots x (key + 64-bit count + list entry)

template <bool W, bool P>
int launch_dedup(ibu_gpu_ctx *ctx, const DedupArgs &a, cudaStream_t s, ibu_error_t *err) {
    const size_t smem = ((size_t)1 << a.s_bits) * (8 + (W ? 8 : 4) + 2);  // keys, counts, claimed-slot list
    if (int rc = set_max_smem(k_bucket_dedup2<W, P>, ctx->device, kDedupMaxSmem, err)) return rc;
    const int per_sm = std::max<int>(1, std::min<size_t>(5, (200u << 10) / (smem + 1024)));  // 48 registers: 5 CTAs by registers
    const uint32_t grid = (uint32_t)std::min<uint64_t>(a.n_buckets, (uint64_t)ctx->sm_count * per_sm);
    k_bucket_dedup2<W, P><<<grid, kBlockThreads, smem, s>>>(a);
    IBU_LAUNCHED("k_bucket_dedup2");
    return IBU_OK;
}

}  // namespace

int k4_sample(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, cudaStream_t s, K4Sample *out, ibu_error_t *err) {
    *out = K4Sample{};
    if (n == 0) return IBU_OK;
    PoolScratch sc(s);
    const uint64_t m = std::min<uint64_t>(n, 1u << 18);
    const uint64_t slots = pow2_ceil(2 * m);
    // [pair fingerprints | barcode fingerprints] 0xFF, then [pair counts | barcode counts | outputs] zero
    uint8_t *buf;
    const size_t fp_bytes = slots * 16, zero_bytes = slots * 8 + 1024;
    IBU_CUDA(sc.alloc(&buf, fp_bytes + zero_bytes));
    IBU_CUDA(cudaMemsetAsync(buf, 0xFF, fp_bytes, s));
    IBU_CUDA(cudaMemsetAsync(buf + fp_bytes, 0, zero_bytes, s));
    uint64_t *ptab = reinterpret_cast<uint64_t *>(buf);
    uint32_t *pcnt = reinterpret_cast<uint32_t *>(buf + fp_bytes);
    unsigned long long *sout = reinterpret_cast<unsigned long long *>(buf + fp_bytes + slots * 8);
    SampleArgs a{recs, n, m, ptab, ptab + slots, pcnt, pcnt + slots, slots - 1, sout, reinterpret_cast<uint32_t *>(sout + kSmpWords)};
    const uint32_t grid = (uint32_t)std::min<uint64_t>((m + kBlockThreads - 1) / kBlockThreads, (uint64_t)ctx->sm_count * 8);
    k_sample<<<grid, kBlockThreads, 0, s>>>(a);
    IBU_LAUNCHED("k_sample");
    unsigned long long *mail = ctx->h_mail;
    IBU_CUDA(cudaMemcpyAsync(mail, sout, 1024, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    out->valid = true;
    out->m = m;
    out->pairs = (double)mail[kSmpPairs];
    out->barcodes = (double)mail[kSmpBarcodes];
    out->unordered = mail[kSmpUnordered];
    out->pair_coll = (double)mail[kSmpPairColl];
    out->pair_f1 = (double)(long long)mail[kSmpPairF1];
    out->pair_f2 = (double)(long long)mail[kSmpPairF2];
    out->bc_f1 = (double)(long long)mail[kSmpBcF1];
    out->bc_f2 = (double)(long long)mail[kSmpBcF2];
    out->bc_max = std::max<uint64_t>(1, mail[kSmpBcMax]);
    memcpy(out->hist, mail + kSmpWords, sizeof(out->hist));
    return IBU_OK;
}

// What the sample says about the whole input: distinct (barcode, umi) pairs and distinct barcodes
// (Chao1 lower bounds, capped at n).
void k4_estimates(const K4Sample &smp, uint64_t n, double *d_est, double *r_est) {
    *d_est = *r_est = 0;
    if (!smp.valid || n == 0) return;
    const bool whole = smp.m >= n;
    *d_est = whole ? smp.pairs : std::min((double)n, chao1(smp.pairs, smp.pair_f1, smp.pair_f2, (double)n));
    *r_est = whole ? smp.barcodes : std::min((double)n, chao1(smp.barcodes, smp.bc_f1, smp.bc_f2, (double)n));
}

// One table build of the partition path, in three steps so that an ingest pipeline can feed it chunk by
// chunk: begin (sizes from the sample, scratch), add (level-1 partition of a chunk's keys, stream
// ordered, any stream that waited for ready()), finish (further levels, de-duplicate, rows).
struct K4Level {
    uint32_t bits_in = 0, bits = 0;  // key bits consumed before this level / by it
    uint64_t cap = 0;                // keys per output bucket (uniform layout)
    uint32_t *cursors = nullptr;     // [2^(bits_in + bits)]
    uint64_t *keys = nullptr, *wts = nullptr;
};
struct K4Job {
    ibu_gpu_ctx *ctx;
    cudaStream_t s0;
    PoolScratch sc;
    StageTimer timer;
    bool trace;
    uint64_t n = 0;       // records the job was sized for
    uint64_t added = 0;
    uint32_t bb = 0, ub = 0, pb = 0, s_bits = 11;
    uint64_t P = 0, wide_cap = 0;
    double r_est = 0;
    bool exact = false, weighted = false, packed = false;
    bool ordered = false;  // order-preserving keys and k_bucket_rows instead of the hash table (k4_ordered.cuh)
    uint64_t *sort_out = nullptr;  // record sort (k4_sort_records_msd): the weights are indices, the result is the sorted records
    std::vector<K4Level> levels;  // levels[0] is filled by add(); the last one feeds the de-duplication
    // chunked mode: add() takes the records in pieces of `chunk` records, partitions a piece into a
    // small level-0 area of its stream (written and read back while still in L2) and straight on
    // into the global level 1; finish() then starts at level 2 (or at the de-duplication)
    bool chunked = false, overflowed = false;
    uint64_t chunk = 0;
    struct Tmp {
        cudaStream_t s = nullptr;
        bool taken = false;
        uint32_t *cursors = nullptr;
        uint64_t *keys = nullptr, *wts = nullptr;
    };
    std::vector<Tmp> tmps;
    uint64_t *wide = nullptr, *bases = nullptr;
    unsigned long long *ctr = nullptr;
    cudaEvent_t ready_ev = nullptr;
    K4Job(ibu_gpu_ctx *c, cudaStream_t s, bool tr) : ctx(c), s0(s), sc(s), timer(tr, s), trace(tr) {}
    ~K4Job() {
        if (ready_ev) cudaEventDestroy(ready_ev);
    }
};

void k4_job_destroy(K4Job *job) { delete job; }
bool k4_job_overflowed(const K4Job *job) { return job->overflowed; }
cudaEvent_t k4_job_ready(K4Job *job) { return job->ready_ev; }

namespace {

template <bool COUNT_ONLY>
int launch_part2(ibu_gpu_ctx *ctx, const Part2Args &a, uint64_t n_in_buckets, bool weighted, cudaStream_t s, ibu_error_t *err) {
    const size_t smem = (size_t)kPartTile * 8 * (weighted ? 2 : 1);
    const uint64_t grid = n_in_buckets * a.tiles_per_bucket;
    if (grid == 0 || grid > 0x7fffffffull) return set_error(err, IBU_ERR_ARG, 0, grid, 0, "partition grid out of range");
    if (weighted) {
        if (int rc = set_max_smem(k_part2<true, COUNT_ONLY>, ctx->device, (size_t)kPartTile * 16, err)) return rc;
        k_part2<true, COUNT_ONLY><<<(uint32_t)grid, kBlockThreads, smem, s>>>(a);
    } else {
        if (int rc = set_max_smem(k_part2<false, COUNT_ONLY>, ctx->device, (size_t)kPartTile * 16, err)) return rc;
        k_part2<false, COUNT_ONLY><<<(uint32_t)grid, kBlockThreads, smem, s>>>(a);
    }
    IBU_LAUNCHED("k_part2");
    return IBU_OK;
}

}  // namespace

int k4_job_begin(ibu_gpu_ctx *ctx, uint64_t n, const K4Hints &hints, const K4Sample &smp, bool pair_mode,
                 bool weighted, const K4Chunking &chunking, cudaStream_t s, K4Job **out, ibu_error_t *err) {
    *out = nullptr;
    const bool forced = hints.force_path == kPathPartition;
    if (!smp.valid || n == 0 || n >= (1ull << 36)) return IBU_OK;
    static const bool trace = getenv("IBU_B200_TRACE") != nullptr;

    // ---- what the sample says ----
    const double m = (double)smp.m;
    uint32_t bb, ub;
    const bool sorting = hints.sort_records;
    if (sorting) {  // every word the sample saw must fit: a sort has no side list
        bb = width_of(smp.hist, 0);
        ub = width_of(smp.hist + 65, 0);
    } else if (hints.bc_len && hints.umi_len) {
        bb = 2 * hints.bc_len;
        ub = 2 * hints.umi_len;
    } else {  // widths that hold 97 % of the sample; what does not fit goes to the wide list
        bb = width_of(smp.hist, smp.m / 32);
        ub = width_of(smp.hist + 65, smp.m / 32);
    }
    if (bb + ub > 64 || bb >= 64 || ub >= 64) return IBU_OK;  // no single-word key: legacy path
    const bool whole = smp.m >= n;  // the sample is the input
    const double d_est = whole ? smp.pairs : std::min((double)n, chao1(smp.pairs, smp.pair_f1, smp.pair_f2, (double)n));
    const double r_est = whole ? smp.barcodes : std::min((double)n, chao1(smp.barcodes, smp.bc_f1, smp.bc_f2, (double)n));
    // sum over keys of (occurrences)^2, from the colliding pairs of the sample: spread of the bucket loads
    const double scale = (double)n / m;
    const double sum_sq = (double)n + scale * scale * 2.0 * smp.pair_coll;
    if (trace)
        fprintf(stderr, "[ibu trace] sample: m=%llu pairs %.0f (f1 %.0f f2 %.0f coll %.0f) barcodes %.0f (f1 %.0f f2 %.0f) -> D~%.3g R~%.3g bb=%u ub=%u\n",
                (unsigned long long)smp.m, smp.pairs, smp.pair_f1, smp.pair_f2, smp.pair_coll, smp.barcodes, smp.bc_f1,
                smp.bc_f2, d_est, r_est, bb, ub);
    if (!forced && !sorting && d_est < 65536.0) return IBU_OK;  // a tiny global table is L2 resident: legacy hash path
    // about as many barcodes as records: a table far outside L2 and a full-size sort of its rows
    // afterwards — the buckets are cut by the barcode's own top bits instead and finished in
    // barcode order (k4_ordered.cuh); IBU_B200_K4_ORDERED=0/1 forces the choice (tests, tuning)
    bool ordered = !pair_mode && r_est > 8.0e6 && r_est > 0.05 * (double)n;
    if (const char *e = getenv("IBU_B200_K4_ORDERED")) ordered = !pair_mode && atoi(e) != 0;
    if (sorting) ordered = true;
    if (ordered && weighted && !sorting) {
        if (!forced) return IBU_OK;  // (the multiplicities would have to travel through the bucket sort: sort fallback)
        ordered = false;
    }

    // ---- sizes ----
    // ~512-1024 records per bucket, whatever the data: a bucket's distinct keys then always fit the
    // shared-memory table (duplicates only lengthen a bucket)
    std::unique_ptr<K4Job> job(new K4Job(ctx, s, trace));
    job->n = n;
    job->bb = bb;
    job->ub = ub;
    job->r_est = r_est;
    job->weighted = weighted;
    static const uint64_t per_bucket = getenv("IBU_B200_K4_BUCKET") ? std::max(64, atoi(getenv("IBU_B200_K4_BUCKET"))) : 1024;  // tuning
    job->P = std::min<uint64_t>(std::max<uint64_t>(pow2_ceil((n + per_bucket - 1) / per_bucket), 2), 1u << 21);
    job->pb = log2_of(job->P);
    if (sorting && job->pb > bb + ub) return IBU_OK;  // more buckets than keys
    // one barcode's records share a final bucket (the bucket bits lie within the barcode): a barcode the sample
    // met far more often than one that just fills a bucket would be met (Poisson: its mean + 6 sigma + 6) ends the
    // attempt before it starts — it would partition everything first and fail in the last kernel.  (Cell
    // barcodes with 10^4 - 10^5 reads each: the LSD sort's case.)
    if (sorting && job->pb <= bb && !whole) {
        const double lam = (double)kOrdCap * m / (double)n;
        if ((double)smp.bc_max > lam + 6.0 * std::sqrt(lam) + 6.0) return IBU_OK;
    }
    // k_bucket_sort_records ranks a pair among the pairs of its bin — the pairs that share the key's top
    // pb + 10 bits — so its cost grows with the bin: 3.7 ms per 10^8 records at 0.75 pairs per bin, 25 ms at 100
    // (10^6 barcodes with 100 records each: all of a barcode's records in one bin), against 9 - 11 ms for the
    // LSD sort.  Distinct prefixes of that length, from the sample's estimates: the barcodes, times what the
    // prefix sees of the UMI, at most the distinct pairs.
    if (sorting) {
        const uint32_t L = job->pb + 10;
        double prefixes = std::ldexp(1.0, (int)std::min<uint32_t>(L, 62));
        prefixes = std::min(prefixes, L > bb ? std::min(d_est, r_est * std::ldexp(1.0, (int)std::min<uint32_t>(L - bb, 40))) : r_est);
        if ((double)n / std::max(prefixes, 1.0) > 24.0) return IBU_OK;
    }
    if (ordered && !sorting && job->pb + 1 > bb) {  // fewer barcode bits than bucket bits: a barcode would span buckets
        if (!forced) return IBU_OK;
        ordered = false;
    }
    job->ordered = ordered;
    // fan-out per level: at most 2^8 from the records, 2^9 from keys (a 4096-key tile then leaves
    // runs of 16 / 8 keys per bucket); up to 2^9 buckets the records are only turned into keys first
    const uint32_t pb = job->pb;
    std::vector<uint32_t> bits;
    if (pb <= 9) bits = {0, pb};
    else if (pb <= 17) bits = {pb / 2, pb - pb / 2};
    else bits = {pb / 3, (pb - pb / 3) / 2, pb - pb / 3 - (pb - pb / 3) / 2};
    uint64_t bytes = 0;
    uint32_t used = 0;
    for (size_t l = 0; l < bits.size(); l++) {
        K4Level lv;
        lv.bits_in = used;
        lv.bits = bits[l];
        used += bits[l];
        // uniform layout with mean + 6 sigma room per bucket
        const double buckets = (double)(1ull << used);
        const double mean = (double)n / buckets, sigma = std::sqrt(sum_sq / buckets);
        // (ordered keys are only as uniform as the barcodes' top bits: a quarter more room)
        lv.cap = ((uint64_t)((ordered ? 1.25 : 1.0) * mean + 6.0 * sigma) + 64 + 15) & ~15ull;
        // (a sort has no second chance for a level before the last — the last one can be laid out exactly — so
        // those get room for key ranges twice as full as the mean where memory allows: the reference's example
        // pattern puts 3 records on the first 30 % of its barcodes and 2 on the rest)
        if (sorting && l + 1 < bits.size() && n <= 400000000ull)
            lv.cap = ((uint64_t)(2.0 * mean + 6.0 * sigma) + 64 + 15) & ~15ull;
        if (l + 1 < bits.size()) {
            // a level whose loads are this uneven (a few keys hold most of the records) is not worth
            // laying out: the global hash of the legacy path keeps such keys in L2
            if ((double)lv.cap > 2.0 * mean + 65536.0) return IBU_OK;
            bytes += (uint64_t)buckets * lv.cap * (weighted ? 16 : 8);
        }
        job->levels.push_back(lv);
    }
    const K4Level &last = job->levels.back();
    const double mean = (double)n / (double)job->P;
    // when duplicates make the final loads too uneven for the uniform layout (or the first attempt
    // overflows) the final buckets are laid out exactly from a histogram of the level before
    job->exact = chunking.force_exact || (double)last.cap > 3.0 * mean + 256.0 || (ordered && last.cap > kOrdCap);
    bytes += job->exact ? n * (weighted ? 16 : 8) : job->P * last.cap * (weighted ? 16 : 8);
    if (bytes > (64ull << 30)) return IBU_OK;
    // shared-memory table: 1.6 slots per key of the fullest bucket the uniform layout admits (all of
    // them distinct at worst); duplicate-heavy data (exact layout) has far fewer distinct keys than
    // records per bucket.  Smaller tables = more resident CTAs = more buckets in flight per SM.
    while (job->s_bits < 13 &&
           (double)(1u << job->s_bits) < 1.6 * (job->exact ? mean + 6.0 * std::sqrt(mean) : (double)last.cap))
        job->s_bits++;
    if (job->pb + job->s_bits > 60) return IBU_OK;
    job->wide_cap = n / 8 + 4096;
    job->packed = !weighted && n < (1ull << 28);

    // chunked unless the final level must be laid out exactly and there are only two levels (the
    // exact layout needs the whole level before it)
    // (device-resident input in one piece: pieces of 1 - 8 Mi records measured 2.8 - 4.0 ms against 2.33 ms
    // per 10^8 records, profiles/r2_k4bench_chunk_sweep.txt — the level-0 area does not stay in L2 long
    // enough to pay for the extra launches; the ingest pipeline chunks because its pieces arrive that way)
    static const char *env_chunk = getenv("IBU_B200_K4_CHUNK");  // records per piece for resident input (tuning)
    uint64_t chunk = chunking.chunk_records ? chunking.chunk_records : (env_chunk ? strtoull(env_chunk, nullptr, 10) : 0);
    job->chunked = chunk != 0 && !chunking.force_exact && !(job->levels.size() == 2 && job->exact) && n > chunk;
    K4Level &l0 = job->levels[0];
    if (job->chunked) {
        chunk = (chunk + kPartTile - 1) / kPartTile * kPartTile;
        job->chunk = chunk;
        const double buckets0 = (double)(1ull << l0.bits), cscale = (double)chunk / m;
        const double csum_sq = (double)chunk + cscale * cscale * 2.0 * smp.pair_coll;
        l0.cap = ((uint64_t)((double)chunk / buckets0 + 6.0 * std::sqrt(csum_sq / buckets0)) + 64 + 15) & ~15ull;
        job->tmps.resize(std::max<uint32_t>(1, chunking.n_streams));
        for (K4Job::Tmp &t : job->tmps) {
            IBU_CUDA(job->sc.alloc(&t.cursors, ((size_t)4 << l0.bits) + 256));
            IBU_CUDA(job->sc.alloc(&t.keys, ((size_t)l0.cap << l0.bits) * 8));
            if (weighted) IBU_CUDA(job->sc.alloc(&t.wts, ((size_t)l0.cap << l0.bits) * 8));
        }
        K4Level &l1 = job->levels[1];
        const size_t n1 = (size_t)1 << (l1.bits_in + l1.bits);
        IBU_CUDA(job->sc.alloc(&l1.cursors, n1 * 4 + 256));
        IBU_CUDA(cudaMemsetAsync(l1.cursors, 0, n1 * 4, s));
        IBU_CUDA(job->sc.alloc(&l1.keys, n1 * l1.cap * 8));
        if (weighted) IBU_CUDA(job->sc.alloc(&l1.wts, n1 * l1.cap * 8));
    } else {
        IBU_CUDA(job->sc.alloc(&l0.cursors, ((size_t)4 << l0.bits) + 256));
        IBU_CUDA(job->sc.alloc(&l0.keys, ((size_t)l0.cap << l0.bits) * 8));
        if (weighted) IBU_CUDA(job->sc.alloc(&l0.wts, ((size_t)l0.cap << l0.bits) * 8));
        IBU_CUDA(cudaMemsetAsync(l0.cursors, 0, (size_t)4 << l0.bits, s));
    }
    IBU_CUDA(job->sc.alloc(&job->wide, job->wide_cap * 24));
    IBU_CUDA(job->sc.alloc(&job->ctr, kCtrWords * 8));
    IBU_CUDA(cudaMemsetAsync(job->ctr, 0, kCtrWords * 8, s));
    if (int rc = set_max_smem(k_part1<false>, ctx->device, (size_t)kPartTile * 8, err)) return rc;
    if (int rc = set_max_smem(k_part1<true>, ctx->device, (size_t)kPartTile * 16, err)) return rc;
    IBU_CUDA(cudaEventCreateWithFlags(&job->ready_ev, cudaEventDisableTiming));
    IBU_CUDA(cudaEventRecord(job->ready_ev, s));
    *out = job.release();
    return IBU_OK;
}

// Partition the keys of `cnt` records on stream s, which must be the job's stream or have waited
// for k4_job_ready().  recs must be 32-byte aligned.
int k4_job_add(K4Job *job, const uint64_t *recs, uint64_t cnt, cudaStream_t s, ibu_error_t *err) {
    if (cnt == 0) return IBU_OK;
    job->added += cnt;
    const K4Level &l0 = job->levels[0];
    auto part1 = [&](const uint64_t *r, uint64_t c, uint32_t *cursors, uint64_t *keys, uint64_t *wts) -> int {
        Part1Args a{r, c, job->bb, job->ub, l0.bits, l0.cap, cursors, keys, wts, job->wide, job->wide_cap, job->ctr,
                    job->ordered ? 1u : 0u};
        const uint32_t grid = (uint32_t)((c + kPartTile - 1) / kPartTile);
        if (job->weighted) k_part1<true><<<grid, kBlockThreads, (size_t)kPartTile * 16, s>>>(a);
        else k_part1<false><<<grid, kBlockThreads, (size_t)kPartTile * 8, s>>>(a);
        IBU_LAUNCHED("k_part1");
        return IBU_OK;
    };
    if (!job->chunked) return part1(recs, cnt, l0.cursors, l0.keys, l0.wts);
    // the level-0 area of this stream (pieces on one stream follow each other; streams do not share)
    K4Job::Tmp *tmp = nullptr;
    for (K4Job::Tmp &t : job->tmps)
        if (t.taken && t.s == s) tmp = &t;
    if (!tmp)
        for (K4Job::Tmp &t : job->tmps)
            if (!t.taken) {
                t.taken = true;
                t.s = s;
                tmp = &t;
                break;
            }
    if (!tmp) return set_error(err, IBU_ERR_ARG, 0, 0, 0, "k4_job_add: more streams than the job was set up for");
    const K4Level &l1 = job->levels[1];
    const size_t L = job->levels.size();
    for (uint64_t off = 0; off < cnt; off += job->chunk) {
        const uint64_t c = std::min<uint64_t>(job->chunk, cnt - off);
        IBU_CUDA(cudaMemsetAsync(tmp->cursors, 0, (size_t)4 << l0.bits, s));
        if (int rc = part1(recs + 3 * off, c, tmp->cursors, tmp->keys, tmp->wts)) return rc;
        Part2Args a{tmp->cursors, tmp->keys, tmp->wts, l0.cap, (uint32_t)((l0.cap + kPartTile - 1) / kPartTile), l1.bits_in, l1.bits,
                    l1.cap, nullptr, l1.cursors, l1.keys, l1.wts, job->ctr, L == 2 ? (uint32_t)kFlagBucket : (uint32_t)kFlagLevel};
        if (int rc = launch_part2<false>(job->ctx, a, 1ull << l1.bits_in, job->weighted, s, err)) return rc;
    }
    return IBU_OK;
}

// The ordered form of the last stage: every final bucket sorted in shared memory, its rows written in
// place (k4_ordered.cuh).  The job's last level is laid out exactly (job->bases) when this runs.
static int finish_ordered(K4Job *job, const std::function<int()> &layout_exact, uint64_t **rows_out, uint64_t *n_rows,
                          uint64_t *n_pairs, bool *handled, ibu_error_t *err) {
    ibu_gpu_ctx *ctx = job->ctx;
    cudaStream_t s = job->s0;
    PoolScratch &sc = job->sc;
    unsigned long long *mail = ctx->h_mail, *ctr = job->ctr;
    const uint64_t n = job->added, P = job->P;
    K4Level &last = job->levels.back();
    // IBU_B200_K4_ORDERED=2 (tests): an input this path has to hand to the sort fallback is an error
    auto give_up = [&]() -> int {
        const char *e = getenv("IBU_B200_K4_ORDERED");
        if (e && atoi(e) == 2) return set_error(err, IBU_ERR_ARG, 0, mail[kCtrFlags], 0, "ordered table path gave up");
        return IBU_OK;
    };
    // the wide list is complete: its size (and the overflow flags of the levels) before the buckets are read
    IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    const uint64_t n_wide = mail[kCtrWide];
    if (job->trace)
        fprintf(stderr, "[ibu trace] ordered: %zu levels to P=2^%u, wide %llu flags %llx\n", job->levels.size(), job->pb,
                (unsigned long long)n_wide, mail[kCtrFlags]);
    if ((mail[kCtrFlags] & kFlagBucket) && !job->exact && !(job->chunked && job->levels.size() == 2)) {
        // the barcodes' top bits are not uniform enough for the uniform layout of the last level: the
        // level before is intact, lay the buckets out exactly from a histogram and go on
        const unsigned long long keep[3] = {mail[kCtrWide], mail[kCtrFlags] & ~(unsigned long long)kFlagBucket, mail[kCtrSpecial]};
        IBU_CUDA(cudaMemcpyAsync(ctr, keep, sizeof(keep), cudaMemcpyHostToDevice, s));
        sc.free_now(last.keys);
        last.keys = nullptr;
        job->exact = true;
        const int rc = layout_exact();
        if (rc < 0) return give_up();
        if (rc) return rc;
        IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
    }
    if ((mail[kCtrFlags] & (kFlagLevel | kFlagWide | kFlagBucket)) || mail[kCtrSpecial]) {
        job->overflowed = false;  // (nothing a second attempt of this path would change)
        return give_up();
    }
    uint64_t *wrows = nullptr, wn = 0, wp = 0;
    uint32_t *wstart = nullptr;
    struct FreeRows {  // rows of the wide list: cudaMallocAsync'ed by the legacy path
        uint64_t *&p;
        cudaStream_t s;
        ~FreeRows() {
            if (p && cudaFreeAsync(p, s) != cudaSuccess) cudaGetLastError();
        }
    } free_wrows{wrows, s};
    if (n_wide) {
        if (int rc = k4_legacy_unsorted(ctx, job->wide, n_wide, s, false, true, &wrows, &wn, &wp, err)) return rc;
        if (wn >= (1ull << 32)) return give_up();
        IBU_CUDA(sc.alloc(&wstart, (P + 1) * 4));
        k_wide_starts<<<(uint32_t)((P + 1 + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, s>>>(
            wrows, (uint32_t)wn, (uint32_t)P, job->bb - job->pb, wstart);
        IBU_LAUNCHED("k_wide_starts");
        job->timer.lap("wide list");
    }
    uint32_t *rows_of;
    uint64_t *row_base;
    IBU_CUDA(sc.alloc(&rows_of, P * 4));
    IBU_CUDA(sc.alloc(&row_base, (P + 1) * 8));
    IBU_CUDA(cudaMemsetAsync(rows_of, 0, P * 4, s));
    IBU_CUDA(cudaMemsetAsync(ctr + kCtrClaimed, 0, (kCtrWords - kCtrClaimed) * 8, s));
    // rows: at most one per record; handed to the caller as it is unless far too large
    uint64_t *out = nullptr;
    const uint64_t rows_cap = std::max<uint64_t>(n, 1);
    IBU_CUDA(alloc_result_rows(ctx, &out, rows_cap * 24, s));
    auto fail = [&](int rc) {
        if (cudaFreeAsync(out, s) != cudaSuccess) cudaGetLastError();
        return rc;
    };
    // the mean bucket has 512 - 1024 keys; k_bucket_sort holds two key arrays, k_bucket_emit one and the row starts
    const uint32_t cap = kOrdCap;
    if (int rc = set_max_smem(k_bucket_sort, ctx->device, 4096 * 16, err)) return fail(rc);
    if (int rc = set_max_smem(k_bucket_emit, ctx->device, 4096 * 12 + 8, err)) return fail(rc);
    OrdArgs a{job->exact ? job->bases : nullptr, last.cursors, last.cap, last.keys, (uint32_t)P, job->pb, job->bb, cap, wrows, wstart,
              rows_of, row_base, ctr, out, rows_cap};
    k_bucket_sort<<<(uint32_t)std::min<uint64_t>(P, (uint64_t)ctx->sm_count * 5), kBlockThreads, (size_t)cap * 16, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e = cudaGetLastError()) return fail(cuda_fail(err, e, "k_bucket_sort"));
    job->timer.lap("k_bucket_sort");
    k_bucket_bases<<<1, 1024, 0, s>>>(rows_of, (uint32_t)P, row_base);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_bucket_emit<<<(uint32_t)std::min<uint64_t>(P, (uint64_t)ctx->sm_count * 6), kBlockThreads, (size_t)cap * 12 + 8, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e = cudaGetLastError()) return fail(cuda_fail(err, e, "k_bucket_emit"));
    if (wn) {
        k_wide_tail<<<ctx->sm_count, kBlockThreads, 0, s>>>(wrows, (uint32_t)wn, wstart, (uint32_t)P, row_base, out, rows_cap, ctr);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaError_t e = cudaGetLastError()) return fail(cuda_fail(err, e, "k_wide_tail"));
    }
    job->timer.lap("k_bucket_emit");
    cudaError_t e = cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mail + kCtrWords, row_base + P, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(cuda_fail(err, e, "k_bucket_emit"));
    const uint64_t R = mail[kCtrWords] + mail[kCtrTail];
    if (job->trace)
        fprintf(stderr, "[ibu trace] ordered: rows %llu (+%llu beyond the layout) pairs %llu flags %llx\n", mail[kCtrWords],
                mail[kCtrTail], mail[kCtrPairs] + (unsigned long long)wp, mail[kCtrFlags]);
    if (mail[kCtrFlags] & (kFlagSmem | kFlagPairsOut)) return fail(give_up());  // a bucket that does not fit: sort fallback
    if (R * 2 < rows_cap) {  // far fewer barcodes than records after all: a block of the right size
        uint64_t *fit = nullptr;
        e = alloc_result_rows(ctx, &fit, R * 24, s);
        if (e == cudaSuccess && R) e = cudaMemcpyAsync(fit, out, R * 24, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) {
            if (fit && cudaFreeAsync(fit, s) != cudaSuccess) cudaGetLastError();
            return fail(cuda_fail(err, e, "barcode rows"));
        }
        if (cudaFreeAsync(out, s) != cudaSuccess) cudaGetLastError();
        out = fit;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(cuda_fail(err, e, "barcode rows"));
    }
    *rows_out = out;
    *n_rows = R;
    *n_pairs = mail[kCtrPairs] + wp;
    *handled = true;
    return IBU_OK;
}

// The last stage of a record sort: the final buckets' sizes -> where each bucket's records go, then every
// bucket sorted by (key, index) in shared memory and written out as records (k_bucket_sort_records).
static int finish_sort(K4Job *job, const std::function<int()> &layout_exact, bool *handled, ibu_error_t *err) {
    ibu_gpu_ctx *ctx = job->ctx;
    cudaStream_t s = job->s0;
    PoolScratch &sc = job->sc;
    unsigned long long *mail = ctx->h_mail, *ctr = job->ctr;
    const uint64_t n = job->added, P = job->P;
    K4Level &last = job->levels.back();
    IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    if (job->trace)
        fprintf(stderr, "[ibu trace] sort: %zu levels to P=2^%u, bb=%u ub=%u, wide %llu flags %llx\n", job->levels.size(), job->pb,
                job->bb, job->ub, mail[kCtrWide], mail[kCtrFlags]);
    if ((mail[kCtrFlags] & kFlagBucket) && !job->exact && !mail[kCtrWide] && !(mail[kCtrFlags] & kFlagLevel)) {
        // the keys' top bits are not uniform enough for the uniform layout of the last level: the level
        // before is intact, lay the buckets out exactly from a histogram and go on
        const unsigned long long keep[3] = {0ull, mail[kCtrFlags] & ~(unsigned long long)kFlagBucket, mail[kCtrSpecial]};
        IBU_CUDA(cudaMemcpyAsync(ctr, keep, sizeof(keep), cudaMemcpyHostToDevice, s));
        sc.free_now(last.keys);
        sc.free_now(last.wts);
        last.keys = last.wts = nullptr;
        job->exact = true;
        const int rc = layout_exact();
        if (rc < 0) return IBU_OK;
        if (rc) return rc;
        IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
    }
    // a record wider than the sample's words, the one key that doubles as the empty marker, a level that
    // overflowed: the LSD sort takes the input
    if ((mail[kCtrFlags] & (kFlagLevel | kFlagWide | kFlagBucket)) || mail[kCtrSpecial] || mail[kCtrWide]) return IBU_OK;
    uint64_t *out_base;
    IBU_CUDA(sc.alloc(&out_base, (P + 1) * 8));
    k_bucket_bases<<<1, 1024, 0, s>>>(last.cursors, (uint32_t)P, out_base);
    IBU_LAUNCHED("k_bucket_bases");
    // buckets of the uniform layout are bounded by its slab: 1536 pairs = 30 KB of shared memory leave room for
    // 5 CTAs per SM instead of 4
    const uint32_t cap = (!job->exact && last.cap <= 1536) ? 1536u : kOrdCap;
    const uint32_t per_sm = cap == 1536u ? 5u : 4u;
    if (int rc = set_max_smem(k_bucket_sort_records, ctx->device, (size_t)kOrdCap * 20, err)) return rc;
    SortRecArgs a{job->exact ? job->bases : nullptr, last.cursors, last.cap, last.keys, last.wts, (uint32_t)P, job->pb, job->bb,
                  job->ub, cap, out_base, job->sort_out, ctr};
    k_bucket_sort_records<<<(uint32_t)std::min<uint64_t>(P, (uint64_t)ctx->sm_count * per_sm), kBlockThreads, (size_t)cap * 20, s>>>(a);
    IBU_LAUNCHED("k_bucket_sort_records");
    job->timer.lap("k_bucket_sort_records");
    IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaMemcpyAsync(mail + kCtrWords, out_base + P, 8, cudaMemcpyDeviceToHost, s));
    IBU_CUDA(cudaStreamSynchronize(s));
    if (job->trace) fprintf(stderr, "[ibu trace] sort: %llu of %llu records placed, flags %llx\n", mail[kCtrWords], (unsigned long long)n, mail[kCtrFlags]);
    if ((mail[kCtrFlags] & kFlagSmem) || mail[kCtrWords] != n) return IBU_OK;  // a bucket that does not fit: LSD sort
    *handled = true;
    return IBU_OK;
}

// Everything after the first level, on the job's stream (which must have waited for every stream
// that added).  *handled = false: nothing produced, take another path.
int k4_job_finish(K4Job *job, const uint64_t *all_recs, bool pair_mode, bool pairs_sorted, uint64_t **rows_out,
                  uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err) {
    (void)all_recs;
    *handled = false;
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    ibu_gpu_ctx *ctx = job->ctx;
    cudaStream_t s = job->s0;
    PoolScratch &sc = job->sc;
    StageTimer &timer = job->timer;
    const bool trace = job->trace, weighted = job->weighted, packed = job->packed;
    unsigned long long *mail = ctx->h_mail;  // pinned: device -> host read-backs without a staged copy
    const uint64_t n = job->added, P = job->P;
    const uint32_t pb = job->pb, bb = job->bb, ub = job->ub, s_bits = job->s_bits;
    unsigned long long *ctr = job->ctr;
    uint64_t *wide = job->wide, *pairs = nullptr;
    if (n > job->n) return IBU_OK;  // more records than the job was sized for
    timer.lap(job->chunked ? "k_part1 + k_part2 (chunked)" : "k_part1");

    // ---- the further levels; the last one uniform first, exact if told so or after an overflow ----
    const size_t L = job->levels.size();
    auto run_level = [&](size_t l, bool count_only, const uint64_t *bases) -> int {
        const K4Level &in = job->levels[l - 1];
        K4Level &lv = job->levels[l];
        Part2Args a{in.cursors, in.keys, in.wts, in.cap, (uint32_t)((in.cap + kPartTile - 1) / kPartTile), lv.bits_in, lv.bits,
                    lv.cap, bases, lv.cursors, lv.keys, lv.wts, ctr, l + 1 == L ? (uint32_t)kFlagBucket : (uint32_t)kFlagLevel};
        const uint64_t n_in = 1ull << lv.bits_in;
        return count_only ? launch_part2<true>(ctx, a, n_in, weighted, s, err) : launch_part2<false>(ctx, a, n_in, weighted, s, err);
    };
    for (size_t l = job->chunked ? 2 : 1; l < L; l++) {
        K4Level &lv = job->levels[l];
        const size_t n_out = (size_t)1 << (lv.bits_in + lv.bits);
        IBU_CUDA(sc.alloc(&lv.cursors, n_out * 4 + 256));
        IBU_CUDA(cudaMemsetAsync(lv.cursors, 0, n_out * 4, s));
        if (l + 1 == L && job->exact) break;  // laid out below
        IBU_CUDA(sc.alloc(&lv.keys, n_out * lv.cap * 8));
        if (weighted) IBU_CUDA(sc.alloc(&lv.wts, n_out * lv.cap * 8));
        if (int rc = run_level(l, false, nullptr)) return rc;
        // (the level before the previous one is no longer needed, but it is NOT released here: a block
        // released in the middle of a build is handed, remapped, to the next larger request — the result
        // rows — and the following build then finds neither block as it left it.  The pool moved
        // gigabytes of mappings back and forth on every call: 100 - 900 ms per allocation at 1.5 - 2 x 10^8
        // records, with the pool's reserved size constant.  Every block now lives to the end of the build,
        // so a build's requests repeat exactly and each finds its own block again.)
        timer.lap("k_part2");
    }
    K4Level &last = job->levels[L - 1];
    auto layout_exact = [&]() -> int {
        // cursors of the last level: histogram -> bases -> the real pass
        if (n * (weighted ? 16 : 8) > (64ull << 30)) return -1;
        IBU_CUDA(cudaMemsetAsync(last.cursors, 0, (size_t)P * 4, s));
        if (int rc = run_level(L - 1, true, nullptr)) return rc;
        if (!job->bases) IBU_CUDA(sc.alloc(&job->bases, (P + 1) * 8));
        k_bucket_bases<<<1, 1024, 0, s>>>(last.cursors, (uint32_t)P, job->bases);
        IBU_LAUNCHED("k_bucket_bases");
        IBU_CUDA(cudaMemsetAsync(last.cursors, 0, (size_t)P * 4, s));
        IBU_CUDA(sc.alloc(&last.keys, n * 8));
        if (weighted) IBU_CUDA(sc.alloc(&last.wts, n * 8));
        if (int rc = run_level(L - 1, false, job->bases)) return rc;
        timer.lap("k_part2 (histogram + exact)");
        return IBU_OK;
    };
    if (job->exact) {
        const int rc = layout_exact();
        if (rc < 0) return IBU_OK;
        if (rc) return rc;
    }

    if (job->sort_out) return finish_sort(job, layout_exact, handled, err);
    if (job->ordered) return finish_ordered(job, layout_exact, rows_out, n_rows, n_pairs, handled, err);

    // ---- per-bucket de-duplication into the barcode table (grown if the estimate was short) ----
    // barcode table: generous (skewed barcode frequencies make the estimate a lower bound); only the
    // slots that are hit occupy L2
    uint64_t t_slots = pair_mode ? 0 : std::max<uint64_t>(1u << 14, pow2_ceil((uint64_t)(8.0 * std::min(job->r_est, (double)n))));
    const uint64_t pairs_cap = pair_mode ? n : 0;
    const uint32_t slot_words = packed ? 2 : 4;
    if (pair_mode) IBU_CUDA(sc.alloc(&pairs, pairs_cap * 24));
    uint64_t *slots = nullptr;
    for (int attempt = 0;; attempt++) {
        if (!pair_mode) {
            if (slots) sc.free_now(slots);
            IBU_CUDA(sc.alloc(&slots, t_slots * slot_words * 8));
            IBU_CUDA(cudaMemsetAsync(slots, 0xFF, t_slots * slot_words * 8, s));
        }
        DedupArgs a{last.cursors, job->exact ? job->bases : nullptr, last.keys, last.wts, (uint32_t)P, (uint32_t)last.cap, pb, ub,
                    s_bits, TableRef{slots, t_slots - 1, ctr, packed ? 1u : 0u}, pairs, pairs_cap};
        int rc = weighted ? (pair_mode ? launch_dedup<true, true>(ctx, a, s, err) : launch_dedup<true, false>(ctx, a, s, err))
                          : (pair_mode ? launch_dedup<false, true>(ctx, a, s, err) : launch_dedup<false, false>(ctx, a, s, err));
        if (rc) return rc;
        timer.lap("k_bucket_dedup2");
        IBU_CUDA(cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s));
        IBU_CUDA(cudaStreamSynchronize(s));
        const uint64_t flags = mail[kCtrFlags];
        if (trace)
            fprintf(stderr, "[ibu trace] partition: %zu levels to P=2^%u %s cap=%llu S=2^%u table=%llu x %u B -> wide %llu pairs %llu rows %llu flags %llx\n",
                    L, pb, job->exact ? "exact" : "uniform", (unsigned long long)last.cap, s_bits, (unsigned long long)t_slots,
                    slot_words * 8, mail[kCtrWide], mail[kCtrPairs], mail[kCtrClaimed], (unsigned long long)flags);
        if (flags & (kFlagLevel | kFlagWide | kFlagSmem | kFlagPairsOut)) return IBU_OK;  // legacy path
        const bool crowded = !pair_mode && mail[kCtrClaimed] > t_slots / 10 * 6;
        if (!(flags & (kFlagTable | kFlagBucket)) && !crowded) break;
        if (attempt == 3 || t_slots >= (1ull << 31)) return IBU_OK;
        const unsigned long long keep[3] = {mail[kCtrWide], 0ull, mail[kCtrSpecial]};
        IBU_CUDA(cudaMemsetAsync(ctr, 0, kCtrWords * 8, s));
        IBU_CUDA(cudaMemcpyAsync(ctr, keep, sizeof(keep), cudaMemcpyHostToDevice, s));
        if (flags & kFlagBucket) {
            // a final bucket overflowed the uniform layout: the level before is intact, lay the
            // buckets out exactly and repeat from there
            job->overflowed = true;
            if (job->exact || (job->chunked && L == 2)) return IBU_OK;  // (chunked: the level before is gone)
            if (trace) fprintf(stderr, "[ibu trace] a bucket overflowed (cap %llu): exact layout\n", (unsigned long long)last.cap);
            sc.free_now(last.keys);
            if (last.wts) sc.free_now(last.wts);
            last.keys = last.wts = nullptr;
            job->exact = true;
            const int rc2 = layout_exact();
            if (rc2 < 0) return IBU_OK;
            if (rc2) return rc2;
        } else {
            // more barcodes than estimated: the buckets are intact, only this stage is repeated
            t_slots *= 8;
        }
    }
    const uint64_t n_wide = mail[kCtrWide];
    for (K4Level &lv : job->levels) {
        if (lv.keys) sc.free_now(lv.keys);
        if (lv.wts) sc.free_now(lv.wts);
        lv.keys = lv.wts = nullptr;
    }

    // ---- 5. wide records and the special key ----
    uint64_t wide_pairs = 0;
    if (n_wide || mail[kCtrSpecial]) {
        uint64_t *wrows = nullptr, wn = 0, wp = 0;
        if (n_wide)
            if (int rc = k4_legacy_unsorted(ctx, wide, n_wide, s, true, true, &wrows, &wn, &wp, err)) return rc;
        ExtraArgs a{wrows, wn, ub, TableRef{slots, t_slots - 1, ctr, packed ? 1u : 0u}, pairs, pairs_cap};
        const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((wn + kBlockThreads - 1) / kBlockThreads, (uint64_t)ctx->sm_count * 8));
        k_extra_pairs<<<grid, kBlockThreads, 0, s>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(mail, ctr, kCtrWords * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (wrows && cudaFreeAsync(wrows, s) != cudaSuccess) cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(err, e, "k_extra_pairs");
        if (mail[kCtrFlags] & (kFlagTable | kFlagPairsOut)) return IBU_OK;  // (rows already added are discarded with the table)
        wide_pairs = wn;
        timer.lap("wide list");
    }
    sc.free_now(wide);
    const uint64_t total_pairs = mail[kCtrPairs] + wide_pairs;

    // ---- 6. rows ----
    uint64_t *out = nullptr;
    if (pair_mode) {
        const uint64_t k = mail[kCtrCursor];
        IBU_CUDA(cudaMallocAsync((void **)&out, k ? k * 24 : 256, s));
        cudaError_t e = cudaSuccess;
        int rc = IBU_OK;
        if (k && pairs_sorted) {
            const uint64_t vary[3] = {n_wide ? ~0ull : (1ull << bb) - 1, n_wide ? ~0ull : (1ull << ub) - 1, 0};
            static const int order[2] = {1, 0};
            rc = k4_sort_rows(ctx, pairs, k, vary, order, 2, s, out, err);
        } else if (k) {
            e = cudaMemcpyAsync(out, pairs, k * 24, cudaMemcpyDeviceToDevice, s);
        }
        if (rc == IBU_OK && e != cudaSuccess) rc = cuda_fail(err, e, "pair rows");
        if (rc == IBU_OK && (e = cudaStreamSynchronize(s)) != cudaSuccess) rc = cuda_fail(err, e, "pair rows");
        if (rc != IBU_OK) {
            cudaFreeAsync(out, s);
            return rc;
        }
        *rows_out = out;
        *n_rows = k;
        *n_pairs = k;
        *handled = true;
        return IBU_OK;
    }
    const uint64_t R = mail[kCtrClaimed], ones = mail[kCtrOnesRec] ? 1 : 0;
    uint64_t *unsorted;
    IBU_CUDA(sc.alloc(&unsorted, R ? R * 24 : 256));
    IBU_CUDA(cudaMemsetAsync(ctr + kCtrCursor, 0, 8, s));
    {
        const uint64_t blocks = (t_slots + 1023) / 1024;
        k_table_rows<<<(uint32_t)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 16), kBlockThreads, 0, s>>>(
            slots, t_slots, packed ? 1u : 0u, unsorted, ctr);
        IBU_LAUNCHED("k_table_rows");
        timer.lap("k_table_rows");
    }
    IBU_CUDA(alloc_result_rows(ctx, &out, (R + ones) * 24, s));
    int rc = IBU_OK;
    if (R) {
        const uint64_t vary[3] = {n_wide ? ~0ull : (1ull << bb) - 1, 0, 0};
        static const int order[1] = {0};
        rc = k4_sort_rows(ctx, unsorted, R, vary, order, 1, s, out, err);
        timer.lap("sort rows by barcode");
    }
    if (rc == IBU_OK && ones) {  // barcode 2^64 - 1 sorts last
        const unsigned long long row[3] = {kEmpty, mail[kCtrOnesRec], mail[kCtrOnesDist]};
        cudaError_t e = cudaMemcpyAsync(out + 3 * R, row, sizeof(row), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "cudaMemcpyAsync");
    }
    if (rc == IBU_OK) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(err, e, "barcode rows");
    }
    if (rc != IBU_OK) {
        cudaFreeAsync(out, s);
        return rc;
    }
    *rows_out = out;
    *n_rows = R + ones;
    *n_pairs = total_pairs;
    *handled = true;
    return IBU_OK;
}

int k4_partition_table(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, const K4Hints &hints, const K4Sample &smp,
                       bool pair_mode, bool pairs_sorted, bool weighted, cudaStream_t s, uint64_t **rows_out,
                       uint64_t *n_rows, uint64_t *n_pairs, bool *handled, ibu_error_t *err) {
    *handled = false;
    *rows_out = nullptr;
    *n_rows = *n_pairs = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        K4Job *job = nullptr;
        K4Chunking ch;
        ch.force_exact = attempt == 1;  // the chunked uniform layout overflowed: whole levels, exact final layout
        if (int rc = k4_job_begin(ctx, n, hints, smp, pair_mode, weighted, ch, s, &job, err)) return rc;
        if (!job) return IBU_OK;
        int rc = k4_job_add(job, recs, n, s, err);
        if (rc == IBU_OK) rc = k4_job_finish(job, recs, pair_mode, pairs_sorted, rows_out, n_rows, n_pairs, handled, err);
        const bool again = rc == IBU_OK && !*handled && k4_job_overflowed(job) && attempt == 0;
        k4_job_destroy(job);
        if (!again) return rc;
    }
    return IBU_OK;
}

int k4_sort_records_msd(ibu_gpu_ctx *ctx, const uint64_t *recs, uint64_t n, const K4Sample &smp, uint64_t *out,
                        cudaStream_t s, bool *handled, ibu_error_t *err) {
    *handled = false;
    K4Hints hints;
    hints.sort_records = true;
    for (int attempt = 0; attempt < 2; attempt++) {
        K4Job *job = nullptr;
        K4Chunking ch;
        ch.force_exact = attempt == 1;
        if (int rc = k4_job_begin(ctx, n, hints, smp, false, true, ch, s, &job, err)) return rc;
        if (!job) return IBU_OK;
        job->sort_out = out;
        uint64_t *rows = nullptr, n_rows = 0, n_pairs = 0;
        int rc = k4_job_add(job, recs, n, s, err);
        if (rc == IBU_OK) rc = k4_job_finish(job, recs, false, false, &rows, &n_rows, &n_pairs, handled, err);
        const bool again = rc == IBU_OK && !*handled && k4_job_overflowed(job) && attempt == 0;
        k4_job_destroy(job);
        if (!again) return rc;
    }
    return IBU_OK;
}

}  // namespace ibu
