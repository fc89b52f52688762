/*
 * ganon_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See ganon_oracle.h.
 *
 * Written from the behaviour of the reference (file:line cited per function), deliberately
 * in the most literal scalar form (deque-like window, per-bit counting) so that it shares no
 * structure with the CUDA kernels it checks.
 */
#include "ganon_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/utils/include/utils/adjust_seed.hpp:33-37 */
uint64_t go_adjust_seed(unsigned k)
{
    return 0x8F3F73B5CF1C9ADEULL >> (64u - 2u * k);
}

/* seqan3/alphabet/nucleotide/dna4.hpp:166-205: rank table A,C,G,T = 0..3 (+lower case), U=T,
 * IUPAC: R,W,M,D,H,V -> A; Y,S,B -> C; K -> G; everything else (incl. N) -> 0. */
unsigned go_dna4_rank(unsigned char c)
{
    switch (c)
    {
    case 'C': case 'c': case 'Y': case 'y': case 'S': case 's': case 'B': case 'b':
        return 1;
    case 'G': case 'g': case 'K': case 'k':
        return 2;
    case 'T': case 't': case 'U': case 'u':
        return 3;
    default:
        return 0;
    }
}

/* seqan3/alphabet/nucleotide/dna15.hpp:95 ("ABCDGHKMNRSTVWY") + nucleotide_base.hpp:147-168
 * (char_is_valid: any rank-table char, its lower case, and U/u). */
int go_dna15_valid(unsigned char c)
{
    static const char legal[] = "ABCDGHKMNRSTVWYUabcdghkmnrstvwyu";
    return c != 0 && strchr(legal, (int)c) != NULL;
}

/* minimiser_hash.hpp:91-107 builds  forward = kmer_hash(seq)^seed,
 * reverse = reverse(kmer_hash(reverse(complement(seq))))^seed ; minimiser.hpp:398-472 slides a
 * window of (w-k+1) values of min(forward,reverse).  kmer_hash.hpp:618-640: first base is the most
 * significant digit, base 4. */
size_t go_minimiser_hash(const char *seq, size_t len, unsigned k, unsigned w, uint64_t seed, uint64_t *out)
{
    if (k == 0 || len < k)
        return 0;
    size_t    nk = len - k + 1;
    uint64_t *v  = (uint64_t *)malloc(nk * sizeof(uint64_t));
    for (size_t i = 0; i < nk; ++i)
    {
        uint64_t f = 0, r = 0;
        for (unsigned j = 0; j < k; ++j)
        {
            f = f * 4 + go_dna4_rank((unsigned char)seq[i + j]);
            r = r * 4 + (3 - go_dna4_rank((unsigned char)seq[i + k - 1 - j])); /* complement = rank^3, dna4.hpp:95-98 */
        }
        f ^= seed;
        r ^= seed;
        v[i] = f < r ? f : r; /* window_value(): std::min(*urng1, *urng2), minimiser.hpp:408-414 */
    }

    size_t win = (size_t)(w - k + 1);
    if (win > nk)
        win = nk; /* minimiser.hpp:298-299 */
    size_t n_out = 0;

    /* window_first, minimiser.hpp:422-436: min_element with less_equal -> LAST of the equal minima */
    size_t off = 0; /* minimiser_position_offset within the current window */
    for (size_t x = 0; x < win; ++x)
        if (v[x] <= v[off])
            off = x;
    uint64_t cur = v[off];
    out[n_out++] = cur;

    /* next_minimiser, minimiser.hpp:444-472, repeated until the end of the range */
    for (size_t e = win; e < nk; ++e)
    {
        size_t   s       = e - win + 1; /* window is v[s..e] after pop_front/push_back */
        uint64_t new_val = v[e];
        if (off == 0)
        {
            size_t m = 0;
            for (size_t x = 0; x < win; ++x)
                if (v[s + x] <= v[s + m])
                    m = x;
            off = m;
            cur = v[s + m];
            out[n_out++] = cur; /* returns true -> value is emitted even if equal to the previous one */
        }
        else if (new_val < cur)
        {
            cur = new_val;
            off = win - 1;
            out[n_out++] = cur;
        }
        else
        {
            --off;
        }
    }
    free(v);
    return n_out;
}

static const uint64_t GO_SEEDS[5] = {13572355802537770549ULL, 13043817825332782213ULL, 10650232656628343401ULL,
                                     16499269484942379435ULL, 4893150838803335377ULL}; /* IBF.hpp:160-164 */

/* IBF.hpp:173-187 */
uint64_t go_ibf_row(const go_ibf *ibf, uint64_t h, unsigned fn)
{
    h *= GO_SEEDS[fn];
    h ^= h >> ibf->hash_shift;
    h *= 11400714819323198485ULL;
    h = (uint64_t)(((__uint128_t)h * (__uint128_t)ibf->bin_size) >> 64);
    return h;
}

static inline int go_bit(const uint64_t *data, uint64_t pos)
{
    return (int)((data[pos >> 6] >> (pos & 63)) & 1u); /* sdsl::bit_vector: LSB-first 64-bit words */
}

/* IBF.hpp:1027-1042 / 639-664 / 926-953, restated bit by bit */
void go_ibf_bulk_count(const go_ibf *ibf, const uint64_t *hashes, size_t n, uint16_t *counts)
{
    memset(counts, 0, ibf->technical_bins * sizeof(uint16_t));
    for (size_t x = 0; x < n; ++x)
    {
        uint64_t row[5];
        for (unsigned i = 0; i < ibf->hash_funs; ++i)
            row[i] = go_ibf_row(ibf, hashes[x], i) * ibf->technical_bins;
        for (uint64_t b = 0; b < ibf->technical_bins; ++b)
        {
            int all = 1;
            for (unsigned i = 0; i < ibf->hash_funs && all; ++i)
                all = go_bit(ibf->data, row[i] + b);
            if (all)
                counts[b] = (uint16_t)(counts[b] + 1);
        }
    }
}

/* IBF.hpp:271-286 */
void go_ibf_emplace(go_ibf *ibf, uint64_t *data, uint64_t value, uint64_t bin)
{
    for (unsigned i = 0; i < ibf->hash_funs; ++i)
    {
        uint64_t pos = go_ibf_row(ibf, value, i) * ibf->technical_bins + bin;
        data[pos >> 6] |= (uint64_t)1 << (pos & 63);
    }
}

/* GanonClassify.cpp:492-495 + 720-724 */
uint64_t go_threshold_cutoff(uint64_t n_hashes, double rel_cutoff)
{
    uint64_t t = (uint64_t)ceil((double)n_hashes * rel_cutoff);
    return t == 0 ? 1 : t;
}

/* GanonClassify.cpp:757-758 */
uint64_t go_threshold_filter(uint64_t max_count, uint64_t min_count, double rel_filter)
{
    return max_count - (uint64_t)ceil((double)(max_count - min_count) * rel_filter);
}

/* GanonClassify.cpp:504-541 */
void go_select_matches_ibf(const go_ibf *ibf, const uint64_t *target_off, const uint64_t *target_bins,
                           const uint32_t *target_gid, const double *target_fpr, size_t n_targets,
                           const uint64_t *hashes, size_t n_hashes, uint64_t threshold_cutoff, uint64_t *best_count,
                           double *best_fpr, uint64_t *max_count, uint64_t *min_count, uint16_t *counts)
{
    go_ibf_bulk_count(ibf, hashes, n_hashes, counts);
    for (size_t t = 0; t < n_targets; ++t)
    {
        uint64_t summed = 0;
        for (uint64_t j = target_off[t]; j < target_off[t + 1]; ++j)
            summed += counts[target_bins[j]];
        if (summed > n_hashes)
            summed = n_hashes;
        if (summed >= threshold_cutoff)
        {
            uint32_t g = target_gid[t];
            if (summed > best_count[g])
            {
                best_count[g] = summed;
                best_fpr[g]   = target_fpr[t];
                if (summed > *max_count)
                    *max_count = summed;
                if (summed < *min_count)
                    *min_count = summed;
            }
        }
    }
}

/* HIBF.hpp:433-460 */
static void go_hibf_impl(const go_hibf *hibf, const uint64_t *hashes, size_t n, int64_t idx, uint64_t threshold,
                         uint16_t *result)
{
    const go_ibf *ibf    = &hibf->ibfs[idx];
    uint16_t     *counts = (uint16_t *)malloc(ibf->technical_bins * sizeof(uint16_t));
    go_ibf_bulk_count(ibf, hashes, n, counts);
    uint16_t sum = 0; /* value_t = TIntCount: wraps */
    for (uint64_t bin = 0; bin < ibf->bins; ++bin)
    {
        sum                 = (uint16_t)(sum + counts[bin]);
        int64_t const fidx = hibf->bin_to_user[idx][bin];
        if (fidx < 0)
        {
            if (sum >= threshold)
                go_hibf_impl(hibf, hashes, n, hibf->next_ibf_id[idx][bin], threshold, result);
            sum = 0;
        }
        else if (bin + 1 == ibf->bins || fidx != hibf->bin_to_user[idx][bin + 1])
        {
            if (sum >= threshold)
                result[fidx] = sum;
            sum = 0;
        }
    }
    free(counts);
}

/* HIBF.hpp:506-523 */
void go_hibf_bulk_count(const go_hibf *hibf, const uint64_t *hashes, size_t n, uint64_t threshold, uint16_t *result)
{
    memset(result, 0, hibf->n_user_bins * sizeof(uint16_t));
    go_hibf_impl(hibf, hashes, n, 0, threshold, result);
}

/* GanonClassify.cpp:498-501 */
static double go_binom(double n, double k)
{
    return exp(lgamma(n + 1) - lgamma(n - k + 1) - lgamma(k + 1));
}

/* GanonClassify.cpp:588-601 */
double go_fpr_query_q(uint64_t n_hashes, uint64_t count, double target_fpr)
{
    double q = 1;
    for (uint64_t i = 0; i <= count; i++)
        q -= go_binom((double)n_hashes, (double)i) * pow(target_fpr, (double)i)
             * pow(1 - target_fpr, (double)(n_hashes - i));
    return q;
}

/* GanonClassify.cpp:940-947 and 969-982 */
double go_target_fpr(uint64_t bin_size_bits, unsigned hash_functions, uint64_t max_hashes_bin, uint64_t count)
{
    uint64_t n_bins_target = (uint64_t)ceil((double)count / (double)max_hashes_bin);
    uint64_t n_hashes_bin  = (uint64_t)ceil((double)count / (double)n_bins_target);
    double   fp = pow(1 - exp(-(double)hash_functions / ((double)bin_size_bits / (double)n_hashes_bin)), (double)hash_functions);
    return 1.0 - pow(1.0 - fp, (double)n_bins_target);
}
