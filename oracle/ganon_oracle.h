/*
 * ganon_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the ganon-classify hot path, used only as
 * the checker for the CUDA implementation (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline leg).  Nothing in ganon_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  The functions below are checked (tests/test_oracle.py) against
 *   - seqan3's own known-answer tests for minimiser_hash / kmer_hash / IBF,
 *   - outputs of the unmodified reference binary (oracle/_ref/ganon-classify, compiled from
 *     /root/reference by oracle/Makefile) on the reference's test genomes + simulated reads
 *     and on adversarial synthetic inputs (committed as fixtures under tests/golden/).
 *
 * Every function cites the reference file:line it restates.  Paths are relative to the
 * reference root; "seqan3/" = libs/seqan3/include/seqan3/.
 */
#ifndef GANON_ORACLE_H
#define GANON_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* raptor::adjust_seed, src/utils/include/utils/adjust_seed.hpp:33-37 */
uint64_t go_adjust_seed(unsigned k);

/* seqan3 dna4 char_to_rank incl. IUPAC conversion, seqan3/alphabet/nucleotide/dna4.hpp:166-205.
 * Returns 0..3 (never fails: every byte maps to a rank; unknown -> 0 'A'). */
unsigned go_dna4_rank(unsigned char c);

/* seqan3 dna15 char validity, as enforced by the FASTA/FASTQ readers with dna4_traits
 * (seqan3/alphabet/nucleotide/dna15.hpp:95, nucleotide_base.hpp:147-168). 1 = legal. */
int go_dna15_valid(unsigned char c);

/* seqan3::views::minimiser_hash with an ungapped shape of size k, window w and the given
 * (already adjusted) seed: seqan3/search/views/minimiser_hash.hpp:76-108,
 * minimiser.hpp:290-302,398-472, kmer_hash.hpp:618-640.
 * `out` must hold at least (len >= k ? len-k+1 : 0) values.  Returns the number emitted. */
size_t go_minimiser_hash(const char *seq, size_t len, unsigned k, unsigned w, uint64_t seed, uint64_t *out);

/* seqan3::interleaved_bloom_filter (uncompressed), seqan3/search/dream_index/interleaved_bloom_filter.hpp */
typedef struct
{
    uint64_t        bins;           /* user-visible bin count                     */
    uint64_t        technical_bins; /* 64 * bin_words                             */
    uint64_t        bin_size;       /* rows (bits per bin)                        */
    uint64_t        hash_shift;     /* countl_zero(bin_size)                      */
    uint64_t        bin_words;      /* 64-bit words per row                       */
    uint64_t        hash_funs;      /* 1..5                                       */
    const uint64_t *data;           /* row-major [row][bin_word], LSB-first bits  */
} go_ibf;

/* hash_and_fit without the final "*technical_bins": IBF.hpp:173-187.  Returns the row. */
uint64_t go_ibf_row(const go_ibf *ibf, uint64_t value, unsigned fn);

/* counting_agent::bulk_count, IBF.hpp:1027-1042 (bulk_contains 639-664, counting_vector+= 926-953).
 * counts[technical_bins], zeroed here; counter type is uint16 and wraps like the reference's TIntCount. */
void go_ibf_bulk_count(const go_ibf *ibf, const uint64_t *hashes, size_t n, uint16_t *counts);

/* emplace: IBF.hpp:271-286 -- sets the h bits of `value` in `bin` (used to build synthetic filters;
 * `data` is written through a non-const alias). */
void go_ibf_emplace(go_ibf *ibf, uint64_t *data, uint64_t value, uint64_t bin);

/* threshold_rel + the "0 -> 1" reset, GanonClassify.cpp:492-495,720-724 */
uint64_t go_threshold_cutoff(uint64_t n_hashes, double rel_cutoff);
/* threshold_filter, GanonClassify.cpp:757-758 */
uint64_t go_threshold_filter(uint64_t max_count, uint64_t min_count, double rel_filter);

/* select_matches for a flat IBF, GanonClassify.cpp:504-541.
 *   target_off[n_targets+1], target_bins[] : CSR target -> technical bins (filter.map)
 *   target_gid[n_targets]                  : id of the target in the per-level table
 *   target_fpr[n_targets]                  : filter_config.target_fpr[target]
 *   best_count[], best_fpr[] (per-level)   : the TMatches map, 0 = absent
 *   max_count / min_count                  : running max_count_read / min_count_read
 *   counts                                  : scratch uint16[technical_bins]                    */
void go_select_matches_ibf(const go_ibf *ibf, const uint64_t *target_off, const uint64_t *target_bins,
                           const uint32_t *target_gid, const double *target_fpr, size_t n_targets,
                           const uint64_t *hashes, size_t n_hashes, uint64_t threshold_cutoff, uint64_t *best_count,
                           double *best_fpr, uint64_t *max_count, uint64_t *min_count, uint16_t *counts);

/* raptor HIBF as vendored by ganon: src/ganon-classify/include/ganon-classify/hierarchical_interleaved_bloom_filter.hpp */
typedef struct
{
    size_t         n_ibf;
    const go_ibf  *ibfs;
    const int64_t *const *next_ibf_id;  /* [ibf][technical bin of ibf (bins entries)]            */
    const int64_t *const *bin_to_user;  /* ibf_bin_to_filename_position [ibf][bin]; <0 = merged  */
    size_t         n_user_bins;
} go_hibf;

/* counting_agent_type::bulk_count / bulk_count_impl, HIBF.hpp:433-460,506-523.
 * result[n_user_bins] zeroed here; running sums are uint16 and wrap (TIntCount). */
void go_hibf_bulk_count(const go_hibf *hibf, const uint64_t *hashes, size_t n, uint64_t threshold, uint16_t *result);

/* The --fpr-query test of filter_matches, GanonClassify.cpp:588-601 (binom 498-501).
 * Returns q; the match is discarded when q > fpr_query. */
double go_fpr_query_q(uint64_t n_hashes, uint64_t count, double target_fpr);

/* false_positive + per-target fpr, GanonClassify.cpp:940-947,969-982 */
double go_target_fpr(uint64_t bin_size_bits, unsigned hash_functions, uint64_t max_hashes_bin, uint64_t target_hashes);

#ifdef __cplusplus
}
#endif
#endif
