// Internal declarations shared by the CUDA kernels (kernels.cu) and the host runtime (db.cpp, session.cpp).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <string>

#include "../../include/ganon_b200.h"

namespace gnb
{

// ---------------------------------------------------------------------------------------------------------------
// Arithmetic of the path (identical on host and device).
// seqan3 IBF hash seeds and multiplier: IBF.hpp:160-164, 173-187.
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define GNB_HD __host__ __device__ __forceinline__
#else
#define GNB_HD inline
#endif

GNB_HD uint64_t ibf_seed(uint32_t i) // IBF.hpp:160-164
{
    switch (i)
    {
    case 0: return 13572355802537770549ULL;
    case 1: return 13043817825332782213ULL;
    case 2: return 10650232656628343401ULL;
    case 3: return 16499269484942379435ULL;
    default: return 4893150838803335377ULL;
    }
}
constexpr uint64_t kIbfMul      = 11400714819323198485ULL;
constexpr uint64_t kMinimiserSeed = 0x8F3F73B5CF1C9ADEULL; // raptor::adjust_seed, adjust_seed.hpp:33-37

GNB_HD uint64_t mulhi64(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// hash_and_fit without the final "* technical_bins": the row of the bitvector (IBF.hpp:173-187)
GNB_HD uint64_t ibf_row(uint64_t v, uint64_t seed, uint32_t hash_shift, uint64_t bin_size)
{
    v *= seed;
    v ^= v >> hash_shift;
    v *= kIbfMul;
    return mulhi64(v, bin_size);
}

GNB_HD uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// ---------------------------------------------------------------------------------------------------------------
// Device view of one flat IBF (or one shard of it) plus the per-session bin -> node tables.
// Storage keeps the reference's bit layout: row-major [row][bin-word], 64-bit little-endian words, LSB = lowest bin
// (sdsl::bit_vector, int_vector.hpp:2029-2035; IBF.hpp:238-240).
// A "chunk" is 64 consecutive bin-words of a row = 4096 bins = one 512-byte warp-wide load (lane l owns words
// 2l, 2l+1 of the chunk = 4 x 32 bins).
// ---------------------------------------------------------------------------------------------------------------
struct Seg
{
    uint32_t mask;     // bins of the segment inside one 32-bin register
    uint32_t node;     // node (target / user bin) the bins belong to
    uint16_t reg;      // 0..3: which of the lane's four 32-bin registers
    uint16_t complete; // 1: these are ALL bins of the node -> decide here; 0: emit a partial sum
};

struct IbfDev
{
    const uint64_t *data;
    uint64_t        bin_size;   // rows
    uint32_t        hash_shift; // countl_zero(bin_size)
    uint32_t        hash_funs;
    uint32_t        row_words; // words per row held here (row stride)
    uint32_t        n_chunks;  // ceil(row_words / 64)
    // tables, indexed relative to this shard; nullptr for the dense test mode
    const uint32_t *single_mask; // [n_chunks*128] bit set: the bin alone is a whole node
    const uint32_t *bin_node;    // [n_chunks*4096]
    const uint32_t *seg_off;     // [n_chunks*32+1] CSR of multi-bin segments per (chunk, lane); nullptr: none
    const Seg      *segs;
};

// sparse result tuple: [63:40] read, [39:17] node, [16] partial flag, [15:0] count
constexpr int      kTupleReadShift = 40, kTupleNodeShift = 17;
constexpr uint32_t kMaxReadsPerBatch = 1u << 24, kMaxNodes = 1u << 23;
GNB_HD uint64_t make_tuple64(uint32_t read, uint32_t node, uint32_t partial, uint32_t count)
{
    return ((uint64_t)read << kTupleReadShift) | ((uint64_t)node << kTupleNodeShift) | ((uint64_t)partial << 16) | count;
}

// GC.cpp:492-495 + 720-724: max(1, ceil(n_hashes * rel_cutoff)) in IEEE double (exact on host and device)
GNB_HD uint32_t threshold_cutoff(uint32_t n_hashes, double rel_cutoff)
{
#ifdef __CUDA_ARCH__
    double t = ceil(__dmul_rn((double)n_hashes, rel_cutoff));
#else
    double t = __builtin_ceil((double)n_hashes * rel_cutoff);
#endif
    uint32_t c = (uint32_t)t;
    return c == 0 ? 1u : c;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel launchers (kernels.cu).  All asynchronous on `st`.
// ---------------------------------------------------------------------------------------------------------------
// K2: counts pass (write=false -> counts[n_reads]) and write pass (write=true -> hashes at hash_off).
void launch_minimisers(const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2,
                       const uint32_t *off2, const uint32_t *len2, uint32_t n_reads, uint32_t k, uint32_t w, int mode,
                       uint32_t *counts, const uint64_t *hash_off, uint64_t *hashes, uint32_t *max_count, unsigned long long *sum_count,
                       cudaStream_t st);
// ub[i] = windows of read (pair) i; optional: items[i] = its segments (K2t over segments), *max_windows = most windows of one mate
void launch_hash_upper_bounds(const uint32_t *len1, const uint32_t *len2, uint32_t n, uint32_t w, uint32_t *ub, cudaStream_t st, uint32_t *items = nullptr,
                              uint32_t *max_windows = nullptr);
// K2t over segments of long reads (k2_thread.cuh): whether a batch takes it; an upper bound of its items; the three launches
// (segments, compaction, flagged reads + totals).  hash_off is the upper-bound layout, item_off the exclusive scan of items.
bool     minimisers_segmented(uint32_t k, uint32_t w, uint32_t n_reads, uint32_t max_windows);
uint64_t minimiser_segments_bound(uint64_t total_windows, uint32_t n_reads);
void     launch_minimisers_segmented(const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2, const uint32_t *off2,
                                     const uint32_t *len2, uint32_t n_reads, uint32_t k, uint32_t w, const uint64_t *item_off, uint64_t item_bound,
                                     uint32_t *seg_cnt, uint8_t *flags, uint32_t *counts, const uint64_t *hash_off, uint64_t *hashes, uint32_t *max_count,
                                     unsigned long long *sum_count, cudaStream_t st);
// exclusive scan of counts (values > 65535 are kept in the offsets; K3 skips such reads) -> hash_off[n+1]
void   launch_scan_counts(const uint32_t *counts, uint64_t *hash_off, uint32_t n_reads, void *tmp, size_t tmp_bytes, cudaStream_t st);
size_t scan_tmp_bytes(uint32_t n_reads);
// K3 sparse: tuples of (read, node, count) for nodes reaching the cutoff (+ partial sums of multi-segment nodes)
void launch_ibf_count(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, const uint32_t *counts, const uint8_t *active,
                      uint32_t n_reads, uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor,
                      uint64_t cap, cudaStream_t st);
// K3 dense (test hook): counts[n_reads][row_words*64]
void launch_ibf_count_dense(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, uint32_t n_reads,
                            uint32_t max_hashes, uint16_t *counts, cudaStream_t st);
// K3h: one traversal round of an HIBF (items = (read, sub-IBF) pairs); see kernels.cu
void launch_hibf_round(const IbfDev *table, uint32_t hash_funs, const uint2 *items, uint32_t n_items, const uint64_t *hashes, const uint64_t *hash_off,
                       const uint32_t *counts, uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap, uint2 *items_out,
                       unsigned long long *items_cursor, uint64_t items_cap, unsigned long long *bytes, uint32_t lanes_per_item, cudaStream_t st);
// first worklist of the traversal, built on the device: (read, 0) for every active read with 1..65535 minimisers
void launch_hibf_seed_items(const uint8_t *active, const uint32_t *counts, uint32_t n_reads, uint2 *items, unsigned long long *cursor, cudaStream_t st);
constexpr uint32_t kMergedBinFlag = 0x80000000u;
// sort tuples by (read, node)
size_t sort_tmp_bytes(uint64_t n);
void   launch_sort_tuples(const uint64_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t st);
// K4: the finishing stage of one hierarchy level on the device (levels with a single filter): cross-run sums, rel-filter,
// --fpr-query, unique / LCA, report accounting and the text of .all/.one/.unc (GC.cpp:504-541, 579-627, 766-803, 1289-1322)
struct FinishSizes // per read; exclusive-scanned into offsets
{
    unsigned long long kept, all_bytes, one_bytes, unc_bytes;
};
enum
{
    kFtProcessed = 0, kFtSkippedBig, kFtSkippedSmall, kFtLength, kFtKmers, kFtClassified, kFtKmersMatches, kFtKmersClassified, kFtMatches,
    kFtUnique, kFtDiscFilter, kFtDiscFpr, kFtActiveHashesNext, kFtAmbiguous, kFinishTotals = 16
};
struct FinishParams
{
    // batch
    const uint64_t *tuples;  // sorted by (read, node)
    uint64_t        n_tuples;
    uint64_t       *entries; // scratch, same size as tuples: accepted (node, status, count) per read at the read's tuple range
    uint32_t       *tuple_start; // [n_reads] first tuple of the read, 0xFFFFFFFF = none
    const uint32_t *n_hashes, *len1, *len2, *id_off, *id_len;
    const uint8_t  *blk1;
    uint8_t        *active, *read_level;
    uint32_t        n_reads;
    // per-read outputs of pass A
    uint32_t    *n_acc;
    FinishSizes *sizes, *offs; // offs = exclusive scan of sizes, [n_reads + 1]
    uint2       *one;          // (node, count) of the .one line
    unsigned long long *totals; // [kFinishTotals]
    // outputs of pass B
    uint64_t *match_off; // [n_reads + 1]
    uint32_t *match_target, *match_count;
    char     *all_text, *one_text, *unc_text;
    // level
    const double   *node_fpr;     // [n_filters][n_nodes]: fpr of the filter's target for the node
    const uint32_t *node_class;   // [n_filters][n_nodes]: index of that fpr among the level's distinct fpr values
    unsigned long long *fpr_memo; // direct-mapped cache of --fpr-query values: [slot] = (key, bits of q)
    uint32_t        fpr_memo_mask;
    const int32_t  *parent;
    const uint32_t *depth, *name_off;
    const char     *names;
    unsigned long long *rep; // [n_nodes][5]: matches, seqs_lca, seqs_unique, discarded_matches_filter, discarded_matches_fprquery
    int32_t  root;
    double   rel_cutoffs[16];     // per filter of the level (--ibf order)
    uint32_t filter_bits, n_nodes; // low bits of a tuple's node field that hold the filter index; nodes of the level
    double   rel_filter, fpr_query, fpr_band;
    uint32_t w, level;
    uint8_t  is_hibf, skip_lca, output_lca, output_all, output_unc, first, last;
};
void   launch_finish_select(const FinishParams &p, cudaStream_t st);                         // boundaries + pass A
size_t finish_scan_tmp_bytes(uint32_t n_reads);
void   launch_finish_scan(const FinishParams &p, void *tmp, size_t tmp_bytes, cudaStream_t st); // sizes -> offs
void   launch_finish_write(const FinishParams &p, cudaStream_t st);                          // pass B
// EM reassignment of multi-matching reads from the matches kept in HBM (src/ganon/reassign.py)
struct EmSizes
{
    unsigned long long reads, id_bytes; // per read of a batch: 1 / id length if the level classified it
};
struct EmStoreDev // matches of all classified reads of a run (one EM group), CSR over reads
{
    uint64_t *off;    // [n_reads + 1] into tgt / cnt
    uint32_t *tgt;    // run-wide target id (by name)
    uint32_t *cnt;
    uint64_t *id_off; // [n_reads + 1] into ids
    char     *ids;
};
size_t em_scan_tmp_bytes(uint32_t n_reads);
// per read of the batch: {kept > 0, id_len} -> exclusive scan -> offs [n_reads + 1]
void launch_em_sizes(const FinishSizes *sizes, const uint32_t *id_len, uint32_t n_reads, EmSizes *es, EmSizes *offs, void *tmp, size_t tmp_bytes, cudaStream_t st);
// append the batch's classified reads (level-local CSR match_off/match_target/match_count) to the store
void launch_em_append(const FinishSizes *sizes, const EmSizes *offs, const uint64_t *match_off, const uint32_t *match_target, const uint32_t *match_count,
                      uint64_t n_matches, const uint32_t *id_off, const uint32_t *id_len, const uint8_t *blk, uint32_t n_reads, const uint32_t *node_to_target,
                      EmStoreDev store, uint64_t base_reads, uint64_t base_matches, uint64_t base_ids, cudaStream_t st);
// first position of every target in the store (the reference numbers targets by first appearance in the .all file)
void launch_em_first_pos(EmStoreDev store, uint64_t n_matches, unsigned long long *first_pos, cudaStream_t st);
// weights of reads with exactly one match
void launch_em_initial(EmStoreDev store, uint64_t n_reads, unsigned long long *initial, cudaStream_t st);
// one EM iteration: counts (preset to the initial weights) += 1 at the top match of every multi-matching read
void launch_em_assign(EmStoreDev store, uint64_t n_reads, const unsigned long long *weight, unsigned long long *counts, cudaStream_t st);
// `.one` lines: sizes pass (out == nullptr: line_len[r]) and write pass (line_off = exclusive scan of line_len)
void launch_em_one(EmStoreDev store, uint64_t n_reads, const unsigned long long *weight, const uint32_t *name_off, const char *names, uint64_t *line_len,
                   const uint64_t *line_off, char *out, unsigned long long *n_multi, cudaStream_t st);
// equal read ids in the store: *n_equal = pairs of equal neighbours among the sorted 64-bit hashes of the ids (0: all ids differ)
size_t em_equal_ids_tmp_bytes(uint64_t n_reads);
void   launch_em_equal_ids(EmStoreDev store, uint64_t n_reads, uint64_t *keys_a, uint64_t *keys_b, void *tmp, size_t tmp_bytes, unsigned long long *n_equal,
                           cudaStream_t st);
size_t em_scan64_tmp_bytes(uint64_t n);
void   launch_scan64(const uint64_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t st);
// build-side
void launch_fill_random(uint64_t *data, uint64_t rows, uint32_t row_words, uint32_t w0, uint32_t total_words, uint64_t bins, uint64_t seed,
                        int and_terms, cudaStream_t st);
void launch_emplace(uint64_t *data, uint64_t bin_size, uint32_t hash_shift, uint32_t hash_funs, uint32_t row_words, uint32_t w0,
                    const uint64_t *hashes, const uint32_t *bins, uint64_t n, cudaStream_t st);
// K1: FASTQ record index on the device (strict 4-line records); see kernels.cu
struct FastqIndexOut
{
    uint32_t *id_off, *id_len, *seq_off, *seq_len; // [cap_reads]
    uint32_t *status;                              // [4]: unused, n_malformed_records, first_bad_record, n_bad_alphabet_records
};
size_t fastq_index_tmp_bytes(uint64_t n_bytes);
void   launch_fastq_count(const uint8_t *blk, uint64_t n_bytes, uint32_t *n_lines_dev, void *tmp, size_t tmp_bytes, cudaStream_t st);
void   launch_fastq_line_starts(const uint8_t *blk, uint64_t n_bytes, uint32_t *line_start, uint32_t cap_lines, void *tmp, cudaStream_t st);
void   launch_fastq_records(const uint8_t *blk, const uint32_t *line_start, uint32_t n_records, FastqIndexOut out, cudaStream_t st);
void   launch_fasta_records(const uint8_t *blk, uint64_t n_bytes, const uint32_t *line_start, uint32_t n_records, FastqIndexOut out, cudaStream_t st); // 2 lines per record

// error plumbing
void        set_error(const std::string &msg);
int         fail(int code, const std::string &msg);
const char *cuda_err(cudaError_t e);
#define GNB_CUDA(call)                                                                              \
    do                                                                                              \
    {                                                                                               \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return gnb::fail(GNB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));    \
    } while (0)

#define GNB_TRY(call)        \
    do                       \
    {                        \
        int _rc = (call);    \
        if (_rc != GNB_OK)   \
            return _rc;      \
    } while (0)

} // namespace gnb
