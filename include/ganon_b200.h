/*
 * ganon_b200.h -- C ABI of libganon_b200.so: the ganon-classify hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (pirovc/ganon) has no FFI for this path: its boundary is the process boundary between the
 * Python wrapper (src/ganon/classify.py:29-64 builds an argv, util.py:9-39 runs it) and the C++ binary
 * `ganon-classify`, whose in-process entry is `bool GanonClassify::run(Config)` (src/ganon-classify/include/
 * ganon-classify/GanonClassify.hpp:8; Config fields Config.hpp:22-49).  This header is what a binding for that
 * entry would bind: plain pointers and sizes, no C++/torch types.  Each entry point names the reference code it
 * replaces.  "GC.cpp" = src/ganon-classify/GanonClassify.cpp, "IBF.hpp" = libs/seqan3/include/seqan3/search/
 * dream_index/interleaved_bloom_filter.hpp, "HIBF.hpp" = src/ganon-classify/include/ganon-classify/
 * hierarchical_interleaved_bloom_filter.hpp.
 *
 * Conventions: every function returns 0 on success or a negative gnb_status; gnb_last_error() gives the message of
 * the last failure on the calling thread.  Handles are opaque and owned by the library until *_free.  Buffers passed
 * in are owned by the caller; buffers handed out in result structs are owned by the handle and stay valid until the
 * next call on that handle.  One host thread per session at a time; different sessions are independent.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with GNB_ERR_CUDA.
 */
#ifndef GANON_B200_H
#define GANON_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNB_ABI_VERSION 2

typedef enum
{
    GNB_OK          = 0,
    GNB_ERR_ARG     = -1, /* invalid argument                                   */
    GNB_ERR_IO      = -2, /* file missing / unreadable / truncated              */
    GNB_ERR_FORMAT  = -3, /* not a ganon .ibf/.hibf, or inconsistent header     */
    GNB_ERR_CUDA    = -4, /* no device, allocation failure, kernel failure      */
    GNB_ERR_CONFIG  = -5, /* Config::validate-style error (Config.hpp:71-245)   */
    GNB_ERR_PARSE   = -6, /* reads file parse error (seqan3::parse_error)       */
    GNB_ERR_LIMIT   = -7  /* batch exceeds an implementation limit              */
} gnb_status;

typedef struct gnb_db      gnb_db;      /* one database (flat IBF or HIBF) resident in HBM        */
typedef struct gnb_session gnb_session; /* one classification run: all hierarchy levels + reports */
typedef struct gnb_comm    gnb_comm;    /* this process's place in a bin-sharded multi-GPU run    */

const char *gnb_last_error(void);
int         gnb_abi_version(void);
int         gnb_device_count(int *n_devices);

/* ------------------------------------------------------------------------------------------------------------------
 * Databases.
 * gnb_db_open replaces load_filter(TIBF) GC.cpp:949-986 / load_filter(THIBF) GC.cpp:875-938 (cereal layouts:
 * IBFConfig.hpp:18-40, IBF.hpp:561-571, sdsl int_vector.hpp:2029-2035, HIBF.hpp:163-169,293-298): the header is parsed
 * by hand, the bitvector is streamed file -> pinned staging -> HBM in chunks (never a second full host copy).
 * shard/n_shards: bin-block (column) sharding for multi-GPU; shard s keeps bin-words [s*bw/n, (s+1)*bw/n) of every
 * row (n_shards <= bin_words: a shard is at least one 64-bin word wide).  n_shards = 1 loads everything.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct
{
    int      is_hibf;
    uint32_t kmer_size;
    uint32_t window_size;
    uint32_t hash_functions;
    uint64_t bins;           /* user-visible bins of the (top-level) IBF                         */
    uint64_t technical_bins; /* 64 * bin_words                                                  */
    uint64_t bin_size_bits;  /* rows                                                            */
    uint64_t bin_words;      /* 64-bit words per row (whole filter, before sharding)            */
    uint64_t shard_word_begin, shard_word_end; /* columns held by this handle                   */
    uint64_t max_hashes_bin;
    double   max_fp;         /* IBFConfig.max_fp / HIBF fpr                                      */
    uint64_t n_targets;
    uint64_t n_ibfs;         /* 1 for a flat IBF                                                 */
    uint64_t device_bytes;   /* HBM held by the bitvector(s)                                     */
    int      device;
    uint64_t n_pages, n_resident_pages; /* host-resident tier: column pages / those kept in HBM (0 = not paged) */
    uint64_t host_bytes;     /* page-locked host memory held by streamed pages                   */
} gnb_db_info_t;

int  gnb_db_open(const char *path, int is_hibf, int device, int shard, int n_shards, gnb_db **out);
/* Host-resident tier (databases larger than the HBM they may use; the reference keeps every filter whole in host RAM,
 * GC.cpp:949-965, real ones reach 501 GB, docs/default_databases.md:75): a flat filter above `hbm_budget_bytes` is cut
 * into column pages (bin-word ranges of every row).  As many pages as fit stay in HBM, the others live in page-locked host
 * memory and pass through two staging buffers while K3 counts the page before them -- the same K3, the same tuples, the
 * same results, at PCIe speed for the streamed part.  gnb_db_open applies $GANON_B200_HBM_BUDGET_GB when set;
 * gnb_db_page_out converts a filter that is whole in HBM (benchmark / tests).  A paged filter serves sessions on one GPU. */
int  gnb_db_open_paged(const char *path, int device, uint64_t hbm_budget_bytes, gnb_db **out);
int  gnb_db_page_out(gnb_db *db, uint64_t hbm_budget_bytes);
int  gnb_db_info(const gnb_db *db, gnb_db_info_t *info);
/* target i: name, per-target false-positive rate (GC.cpp:969-982 / 932) and number of technical bins */
int  gnb_db_target(const gnb_db *db, uint64_t i, const char **name, double *fpr, uint64_t *n_bins);
void gnb_db_free(gnb_db *db);

/* Build-side helpers (the subset of ganon-build needed to make databases directly in HBM: IBF constructor
 * IBF.hpp:222-246, emplace IBF.hpp:271-286, bin map GanonBuild.cpp:619-653).  Used by the benchmark and tests. */
int gnb_db_create(uint64_t bins, uint64_t bin_size_bits, uint32_t hash_functions, uint32_t kmer_size,
                  uint32_t window_size, int device, gnb_db **out);
/* the same filter, but only the bin-word columns of shard `shard` of `n_shards` are held (gnb_db_fill_random and
 * gnb_db_emplace then produce exactly that slice of the whole filter) */
int gnb_db_create_sharded(uint64_t bins, uint64_t bin_size_bits, uint32_t hash_functions, uint32_t kmer_size,
                          uint32_t window_size, int device, int shard, int n_shards, gnb_db **out);
/* word(i) = AND of `and_terms` splitmix64 draws of (seed, i, term): bit density 2^-and_terms; padding bins cleared */
int gnb_db_fill_random(gnb_db *db, uint64_t seed, int and_terms);
int gnb_db_emplace(gnb_db *db, const uint64_t *hashes, const uint32_t *bins, uint64_t n); /* host arrays */
/* bin_target[bins]: target index of every bin; target_hashes[n_targets]: hashes_count_std entries */
int gnb_db_set_targets(gnb_db *db, uint64_t n_targets, const char *const *names, const uint32_t *bin_target,
                       const uint64_t *target_hashes, uint64_t max_hashes_bin);
/* IBFConfig.max_fp / true_max_fp / true_avg_fp of the file gnb_db_save writes (GanonBuild.cpp:251-288) */
int gnb_db_set_fp(gnb_db *db, double max_fp, double true_max_fp, double true_avg_fp);
int gnb_db_read_words(const gnb_db *db, uint64_t ibf_index, uint64_t word_offset, uint64_t n_words, uint64_t *out);
/* flat: .ibf in the reference layout (save_filter GanonBuild.cpp:251-288); HIBF: raptor 3.0.1 index layout as read by
 * load_filter(THIBF) GC.cpp:875-938 (one path "/db/<target>.minimiser" per user bin) */
int gnb_db_save(const gnb_db *db, const char *path);
/* An HIBF directly in HBM (tests / benchmark; raptor, the HIBF builder, is not part of the reference tree): n_ibfs
 * sub-IBFs (0 = top level) of bins[i] technical bins x bin_size_bits[i] rows, tables in the layout of HIBF.hpp:124-136,
 * 176-188 flattened over the sub-IBFs' bins -- next_ibf[j]: child IBF of a merged bin (else the IBF's own index),
 * bin_to_user[j]: user bin, or < 0 for a merged bin.  Target names = user_bin_names; per-target fpr = fpr (GC.cpp:932).
 * gnb_db_fill_random fills every sub-IBF (seed + index); gnb_db_emplace_ibf inserts into one sub-IBF. */
int gnb_db_create_hibf(uint64_t n_ibfs, const uint64_t *bins, const uint64_t *bin_size_bits, uint32_t hash_functions,
                       uint32_t kmer_size, uint32_t window_size, const int64_t *next_ibf, const int64_t *bin_to_user,
                       uint64_t n_user_bins, const char *const *user_bin_names, double fpr, int device, gnb_db **out);
int gnb_db_emplace_ibf(gnb_db *db, uint64_t ibf_index, const uint64_t *hashes, const uint32_t *bins, uint64_t n);

/* ganon-build's count_hashes for ONE input file (src/ganon-build/GanonBuild.cpp:184-249): the file (plain / gzip, FASTA /
 * FASTQ, read like the read files) goes through K2 in segments and the distinct minimisers of all its sequences of at
 * least min_length bases -- the set the reference builds per file before counting and storing it -- come out of a sort +
 * unique in HBM.  A parse error drops the file's hashes but keeps the counts (GanonBuild.cpp:241-245).  The hash set
 * (ascending) stays in page-locked host memory until gnb_hash_set_free. */
typedef struct gnb_hash_set gnb_hash_set;
typedef struct
{
    uint64_t n_sequences, n_skipped, n_bases; /* Total of GanonBuild.cpp:63-70                       */
    uint64_t n_hashes_total, n_unique;        /* minimisers emitted / distinct                       */
    int      parse_error;
} gnb_build_file_stats;
int  gnb_build_file_hashes(int device, const char *path, uint32_t k, uint32_t w, uint64_t min_length, int io_threads, gnb_hash_set **out,
                           gnb_build_file_stats *stats);
int  gnb_hash_set_data(const gnb_hash_set *s, const uint64_t **hashes, uint64_t *n);
void gnb_hash_set_free(gnb_hash_set *s);

/* ------------------------------------------------------------------------------------------------------------------
 * Bin-sharded runs over several GPUs (one process per GPU).  The reference has no multi-device form: GanonClassify::run
 * (GC.cpp:1676) is one process whose threads share the filter in host RAM (load_filter GC.cpp:949-965); a database larger
 * than one GPU's HBM is split here by bin-word columns (gnb_db_open(..., shard, n_shards)), every rank classifies the
 * same reads on its columns and the sparse per-read tuples -- not the per-bin count vectors -- are exchanged inside HBM.
 * The exchange (NCCL, found at run time: the copy already loaded in the process, else $GANON_B200_NCCL, else the
 * system's libnccl.so.2) lives inside the library: a session created with gnb_session_config.comm set behaves on every
 * rank like the unsharded session (classify / stage+run / submit+collect), provided all ranks make the same calls in the
 * same order on the same blocks.  Every rank returns the identical, complete result.
 * gnb_comm_unique_id: called on one rank; the GNB_COMM_ID_BYTES bytes are carried to the other ranks by the caller
 * (torch.distributed, MPI, a pipe ...).  gnb_comm_create: collective over all ranks.
 * ---------------------------------------------------------------------------------------------------------------- */
#define GNB_COMM_ID_BYTES 256
int  gnb_comm_unique_id(void *id, uint64_t cap);
int  gnb_comm_create(const void *id, int rank, int n_ranks, int device, gnb_comm **out);
int  gnb_comm_info(const gnb_comm *c, int *rank, int *n_ranks, int *device, int *nccl_version);
void gnb_comm_free(gnb_comm *c);

/* ------------------------------------------------------------------------------------------------------------------
 * Test hooks for single kernels.
 * ---------------------------------------------------------------------------------------------------------------- */
/* K2: seqan3::views::minimiser_hash(ungapped{k}, window_size{w}, adjust_seed(k)) of one sequence, emitted order
 * (minimiser_hash.hpp:76-108, minimiser.hpp:398-472, kmer_hash.hpp:618-640, adjust_seed.hpp:33-37). */
int gnb_minimisers(int device, uint32_t k, uint32_t w, const char *seq, uint64_t len, uint64_t *out, uint64_t cap,
                   uint64_t *n_out);
/* K2 over many sequences: seqs = concatenated text, seq_off[n+1]; hash_off[n+1] is always filled, hashes only if the
 * total fits in cap (call once with cap = 0 to size the buffer). */
int gnb_minimisers_batch(int device, uint32_t k, uint32_t w, const char *seqs, const uint64_t *seq_off, uint64_t n,
                         uint64_t *hash_off, uint64_t *hashes, uint64_t cap);
/* K3: counting_agent::bulk_count (IBF.hpp:1027-1042) for n_reads hash lists; counts[n_reads][technical_bins] (host).
 * ibf_index selects the sub-IBF of an HIBF (0 for flat). */
int gnb_db_bulk_count(const gnb_db *db, uint64_t ibf_index, const uint64_t *hashes, const uint64_t *hash_off,
                      uint64_t n_reads, uint16_t *counts);

/* ------------------------------------------------------------------------------------------------------------------
 * Sessions: GanonClassify::run (GC.cpp:1676) minus file opening -- the caller streams read blocks in and writes
 * the returned text out.  parse_hierarchy (GC.cpp:353-401) and Config::validate_hierarchy (Config.hpp:175-245)
 * semantics apply to the per-filter / per-level arrays below.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct
{
    uint32_t            n_filters;        /* databases, in --ibf order                                               */
    gnb_db *const      *dbs;              /* [n_filters]                                                             */
    const char *const  *hierarchy_labels; /* [n_filters] or NULL (all "H1")                                          */
    const double       *rel_cutoff;       /* [n_filters]                                                             */
    const char *const  *tax_files;        /* [n_filters] or NULL: no taxonomy => LCA skipped (Config.hpp:168-170)    */
    uint32_t            n_levels;         /* unique labels                                                           */
    const double       *rel_filter;       /* [n_levels] in order of first appearance of each label                   */
    const double       *fpr_query;        /* [n_levels] idem                                                         */
    int                 skip_lca;
    const char         *tax_root_node;    /* default "1"                                                             */
    int                 output_lca, output_all, output_unclassified, output_single;
    int                 device;
    int                 host_threads;     /* threads for the host finishing stage (0 = hardware concurrency)         */
    int                 n_reads_chunk;    /* --n-reads (only observable in the parse-error truncation rule); 0=400   */
    int                 quiet;            /* suppress WARNING lines on stderr (--quiet)                              */
    void               *cuda_stream;      /* cudaStream_t to run on (e.g. a framework's stream); NULL = own stream   */
    gnb_comm           *comm;             /* bin-sharded run: dbs[] are this rank's column shards; NULL = one GPU   */
    int                 sliced_ingest;    /* with comm: each rank copies only bytes [r*S, (r+1)*S), S = ceil(len/n  */
                                          /* rounded up to 16), of a read block to its GPU and the slices are       */
                                          /* all-gathered over NVLink; the other bytes of the block are never read   */
                                          /* on this rank (they need not be valid memory contents)                   */
} gnb_session_config;

typedef struct
{
    /* input accounting */
    uint64_t n_reads;              /* records (pairs) taken from the blocks                                        */
    uint64_t consumed1, consumed2; /* bytes of block1/block2 that formed complete records                           */
    int      parse_error;          /* !=0: a record was malformed; reads before it (n-reads rule) were classified   */
    /* structured result (CSR over reads, final matches after cutoff, rel-filter and fpr-query) */
    const uint64_t *match_off;    /* [n_reads+1]                                                                    */
    const uint32_t *match_target; /* index into gnb_session_node_name() of the level the read was classified at     */
    const uint32_t *match_count;
    const uint8_t  *read_level;   /* [n_reads] index of the level that classified the read, 0xFF = unclassified     */
    const uint32_t *n_hashes;     /* [n_reads] minimisers of the read (pair); 0 = skipped (shorter than window)     */
    uint64_t        n_classified;
    /* text, in the reference's output-file formats (GC.cpp:1289-1322) */
    uint32_t           n_levels;
    const char *const *all_text; /* [n_levels] `.all` lines  readid \t target \t count                              */
    const uint64_t    *all_len;
    const char *const *one_text; /* [n_levels] `.one` lines                                                         */
    const uint64_t    *one_len;
    const char        *unc_text; /* `.unc` lines                                                                    */
    uint64_t           unc_len;
    /* device timings of this batch (CUDA events on the session stream), milliseconds */
    float ms_h2d, ms_index, ms_minimiser, ms_count, ms_sort, ms_d2h;
    double ms_host_index, ms_host_finish, ms_total;
    uint64_t n_minimisers;       /* sum of n_hashes over processed reads                                            */
    uint64_t count_kernel_bytes; /* algorithmic bytes of the IBF-count launches: sum n_hashes * h * bin_words * 8   */
    uint64_t n_kernel_launches;  /* kernels of this library launched for the batch                                  */
    uint64_t h2d_bytes, d2h_bytes; /* bytes copied host->device / device->host for the batch                        */
    float    ms_finish_device;   /* K4 (finishing stage kernels: select, scan, write) of the batch                  */
    uint32_t levels_on_device;   /* hierarchy levels of the batch finished by K4 (the others by the host stage)     */
    float    ms_exchange;        /* bin-sharded runs: tuple exchange between the ranks (counts + lists), device time  */
    uint64_t exchanged_bytes;    /* tuples of all ranks received by this one                                         */
} gnb_batch_result;

int  gnb_session_create(const gnb_session_config *cfg, gnb_session **out);
void gnb_session_free(gnb_session *s);

/* Classify one block of reads end to end (classify<TFilter>() GC.cpp:630-832 for every hierarchy level, plus the
 * reader GC.cpp:1220-1287 for the block).  block1/block2: raw FASTQ or FASTA text in HOST memory (block2 NULL for
 * single-end); the library takes as many complete records as both blocks hold (final != 0: the blocks end the file,
 * a missing trailing newline is tolerated).  prefix_id selects the accounting scope (--batch-reads prefixes). */
int gnb_session_classify(gnb_session *s, uint32_t prefix_id, const char *block1, uint64_t len1, const char *block2,
                         uint64_t len2, int final, gnb_batch_result *out);

/* Split form of the same call for timing with inputs resident in HBM: stage = index + copy to the device,
 * run = kernels only (K1/K2/K3/sort) on the staged batch, finish = device->host + host finishing stage. */
int gnb_session_stage(gnb_session *s, const char *block1, uint64_t len1, const char *block2, uint64_t len2, int final,
                      uint64_t *n_reads);
int gnb_session_run_staged(gnb_session *s, gnb_batch_result *timings);
int gnb_session_finish_staged(gnb_session *s, uint32_t prefix_id, gnb_batch_result *out);

/* Asynchronous form of gnb_session_classify for streaming a file: submit indexes the block and copies it to the device
 * in the calling thread (staged_info->n_reads / consumed1 / consumed2 / parse_error are valid on return, so the caller
 * can cut the next block; a block that does not end the file may leave up to 2 x --n-reads complete records unconsumed --
 * they must come back at the front of the next block: the parse-error rule of GC.cpp:1240-1283 retracts whole --n-reads
 * chunks, and nothing already classified can be taken back), then kernels and host finishing run in a worker thread on the slot's own CUDA stream while the
 * next block is staged.  collect returns the oldest submitted batch (submission order).  The blocks passed to submit
 * must stay untouched until their batch has been collected; a collected result stays valid until the next collect.
 * gnb_session_in_flight: batches submitted and not collected, and how many may be in flight at once. */
int gnb_session_submit(gnb_session *s, uint32_t prefix_id, const char *block1, uint64_t len1, const char *block2,
                       uint64_t len2, int final, gnb_batch_result *staged_info);
int gnb_session_collect(gnb_session *s, gnb_batch_result *out);
int gnb_session_in_flight(const gnb_session *s, uint32_t *n, uint32_t *capacity);

/* Level-wise form for bin-sharded multi-GPU runs (every rank holds a column shard of the filter and stages the same
 * block): run_level = K2/K3 of one hierarchy level on this rank's shard; level_tuples exposes the sparse result of one
 * filter -- uint64 [63:40] read | [39:17] node | [16] partial-sum flag | [15:0] count, sorted by (read, node);
 * the caller concatenates the ranks' tuples (all-gather) and hands them back with set_level_tuples; finish_level runs
 * the host finishing stage of that level; collect_staged merges the result after the last level. */
int gnb_session_run_level(gnb_session *s, uint32_t level);
int gnb_session_level_tuples(gnb_session *s, uint32_t level, uint32_t filter, const uint64_t **tuples, uint64_t *n);
int gnb_session_set_level_tuples(gnb_session *s, uint32_t level, uint32_t filter, const uint64_t *tuples, uint64_t n);
int gnb_session_finish_level(gnb_session *s, uint32_t level);
int gnb_session_collect_staged(gnb_session *s, uint32_t prefix_id, gnb_batch_result *out);
/* The same exchange without leaving HBM (levels with one filter): run_level_device keeps the sorted tuples on the
 * device, level_tuples_device exposes them (device pointer, valid until the next call on the session),
 * set_level_tuples_device takes the concatenation of all ranks' tuples from device memory (any order; the work that
 * produced it must be complete, e.g. the NCCL all-gather synchronised) and sorts it, finish_level_device runs K4 on it
 * (rel-filter, fpr-query, LCA, accounting under prefix_id, output text) or, if K4 declines the level, the host
 * finishing stage.  collect_staged then returns the batch result as usual. */
int gnb_session_run_level_device(gnb_session *s, uint32_t level);
int gnb_session_level_tuples_device(gnb_session *s, uint32_t level, const uint64_t **dev_tuples, uint64_t *n);
int gnb_session_set_level_tuples_device(gnb_session *s, uint32_t level, const uint64_t *dev_tuples, uint64_t n);
int gnb_session_finish_level_device(gnb_session *s, uint32_t level, uint32_t prefix_id);
/* Measurement: the traversal rounds of the last HIBF filter run of the staged forms (gnb_session_run_staged ...): kernel time,
 * algorithmic bytes (sum over the round's (read, sub-IBF) items of n_hashes x h x row bytes) and worklist length per round. */
int gnb_session_hibf_rounds(gnb_session *s, uint32_t cap, float *ms, uint64_t *bytes, uint64_t *items, uint32_t *n_rounds);
/* device timings / byte counts of the staged batch so far (level-wise forms; the batch stays staged) */
int gnb_session_staged_timings(gnb_session *s, gnb_batch_result *timings);

/* Read files as the record reader sees them (parse_reads GC.cpp:1220-1287 opens them through seqan3's transparent
 * decompression, seqan3/io/detail/misc_input.hpp): plain, or gzip by magic number -- single-member, multi-member and BGZF
 * alike are inflated by `io_threads` host threads at once (0 = all, at most 16; csrc/gzstream.h), CRC-32 and length of every
 * member verified; bzip2 ("BZh") block-parallel through libbz2 (dlopen; csrc/bz2stream.cpp), block and stream CRCs verified.
 * gnb_reads_file_read: the next bytes of the decompressed stream (> 0), 0 at the end, < 0 = gnb_status.
 * Files named .embl / .genbank / .gb / .gbk / .sam (seqan3's other sequence formats, format_embl.hpp / format_genbank.hpp /
 * format_sam.hpp; compression suffix stripped first) come out rewritten record by record as two-line FASTA (">id\nSEQ\n").
 * No device is involved. */
typedef struct gnb_reads_file gnb_reads_file;
int     gnb_reads_file_open(const char *path, int io_threads, gnb_reads_file **out);
int64_t gnb_reads_file_read(gnb_reads_file *f, void *dst, uint64_t cap);
int     gnb_reads_file_is_gzip(const gnb_reads_file *f);
void    gnb_reads_file_close(gnb_reads_file *f);

/* One read file (file2 NULL / "") or one pair of files, start to end: the reader / classify / writer loop of the reference
 * (parse_reads GC.cpp:1220-1287 -> classify GC.cpp:630-832 -> write_classified / write_unclassified GC.cpp:1289-1322)
 * with the block ring, the prefetch and the writer thread inside the library: file blocks (plain: parallel preads; gzip:
 * parallel inflate, see gnb_reads_file_*) in page-locked buffers -> gnb_session_submit -> batches in flight ->
 * gnb_session_collect -> text written to the given file descriptors (-1 = discard; every rank of a bin-sharded run holds
 * the full result, so only one passes descriptors).  A parse error ends the file as in the reference (GC.cpp:1278-1283;
 * parse_error is set, the call succeeds).  block_bytes 0 = 64 MiB, io_threads 0 = all (at most 16). */
typedef struct
{
    uint32_t   n_levels; /* entries of all_fd / one_fd                                 */
    const int *all_fd;   /* [n_levels] `.all` lines of every hierarchy level, or NULL  */
    const int *one_fd;   /* [n_levels] `.one` lines, or NULL                           */
    int        unc_fd;   /* `.unc` lines                                               */
} gnb_output_fds;
typedef struct
{
    uint64_t n_records, n_classified, n_blocks, bytes_read1, bytes_read2;
    int      parse_error, is_gzip;
    double   ms_open, ms_read_wait, ms_submit, ms_collect, ms_write; /* host pipeline: where the calling thread waited */
} gnb_files_result;
int gnb_session_classify_files(gnb_session *s, uint32_t prefix_id, const char *file1, const char *file2, const gnb_output_fds *out,
                               uint64_t block_bytes, int io_threads, gnb_files_result *res);

/* Page-lock / unlock a host buffer that will be passed as a read block (cudaHostRegister): faster, truly asynchronous
 * host->device copies. */
int gnb_host_register(void *ptr, uint64_t bytes);
int gnb_host_unregister(void *ptr);

/* Node (target / taxonomy node) names of a level, as used by match_target. */
int gnb_session_level_count(const gnb_session *s, uint32_t *n_levels);
int gnb_session_level_label(const gnb_session *s, uint32_t level, const char **label);
int gnb_session_node_name(const gnb_session *s, uint32_t level, uint32_t node, const char **name);

/* Reports accumulated so far for one prefix: `.rep` (write_report GC.cpp:834-853 + write_report_totals 855-863) and
 * `.sta` (write_stats GC.cpp:1167-1218).  The text is owned by the session. */
int gnb_session_report(gnb_session *s, uint32_t prefix_id, const char **text, uint64_t *len);
int gnb_session_stats(gnb_session *s, uint32_t prefix_id, const char *prefix_name, const char **text, uint64_t *len);

/* EM reassignment of reads with several matches (`ganon classify --multiple-matches em`, the default: src/ganon/
 * classify.py:76-88 runs src/ganon/reassign.py on the `.all` / `.rep` files the binary wrote).  Here the matches of the
 * run stay in HBM: gnb_session_keep_matches(s, 1) before the first batch makes every batch append its classified reads
 * (ids, targets, counts) to a store on the device; gnb_session_reassign runs the iterations there (reassign.py:72-141:
 * initial weights from unique matches, each multi-matching read to its most probable target, until the summed change of
 * the probabilities is <= threshold or max_iter iterations; 0 = until convergence) and returns the `.one` lines
 * (reassign.py:149-179) per EM group -- one group per hierarchy level, or a single one for one level / --output-single,
 * as the reference finds its `.all` files (reassign.py:37-60) -- and the new `.rep` (reassign.py:188-220).
 * Reads that share an id are one read, as in the reference (its dictionary is keyed by read id, reassign.py:78-85): equal ids
 * are looked for on the device and the store is regrouped before the iterations only when there are any. */
typedef struct
{
    uint32_t           n_groups;
    const char *const *group_label;      /* hierarchy label of the group; "" for the single group                  */
    const char *const *one_text;         /* [n_groups] readid \t target \t count                                   */
    const uint64_t    *one_len;
    const uint32_t    *iterations;       /* [n_groups] iterations run                                               */
    const uint64_t    *reassigned_reads; /* [n_groups] reads with several matches                                   */
    const char        *rep_text;         /* new `.rep`: label, target, matches, unique, reassigned - unique, rank, name */
    uint64_t           rep_len;
} gnb_reassign_result;
int gnb_session_keep_matches(gnb_session *s, int enable);
int gnb_session_reassign(gnb_session *s, uint32_t prefix_id, double threshold, uint32_t max_iter, gnb_reassign_result *out);

typedef struct
{
    uint64_t input_seqs, seqs_processed, seqs_skipped_big, seqs_skipped_small, length_processed, kmers_processed,
        seqs_classified, kmers_matches, kmers_from_classified_seqs, matches, seqs_unique, discarded_matches_filter,
        discarded_matches_fprquery;
} gnb_totals; /* struct Total GC.cpp:162-177 */
int gnb_session_totals(const gnb_session *s, uint32_t prefix_id, int level /* -1 = all */, gnb_totals *out);

#ifdef __cplusplus
}
#endif
#endif /* GANON_B200_H */
