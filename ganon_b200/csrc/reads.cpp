// Host-side record reader for read blocks that the device indexer (K1) does not take: FASTA, multi-line FASTQ,
// records with blanks.  Restates what ganon-classify gets from seqan3::sequence_file_input with dna4_traits
// (parse_reads GC.cpp:1220-1287; format_fastq.hpp:105-267; format_fasta.hpp:150-330; legality = dna15,
// nucleotide_base.hpp:147-168).  Produces spans into the block; sequences that are not contiguous in the block are
// compacted into `aux`, addressed as offsets >= block length.
#include <cstring>

#include "reads.h"

namespace gnb
{
namespace
{
struct Lut
{
    bool legal[256], space[256], digit[256];
    Lut()
    {
        memset(legal, 0, sizeof legal);
        memset(space, 0, sizeof space);
        memset(digit, 0, sizeof digit);
        for (const char *p = "ABCDGHKMNRSTVWYUabcdghkmnrstvwyu"; *p; ++p)
            legal[(uint8_t)*p] = true;
        for (const char *p = " \t\n\v\f\r"; *p; ++p)
            space[(uint8_t)*p] = true;
        for (char c = '0'; c <= '9'; ++c)
            digit[(uint8_t)c] = true;
    }
};
const Lut kLut;

// length of the run of legal letters at s
inline uint64_t legal_run(const uint8_t *s, uint64_t n)
{
    uint64_t i = 0;
    for (; i + 8 <= n; i += 8)
        if (!(kLut.legal[s[i]] & kLut.legal[s[i + 1]] & kLut.legal[s[i + 2]] & kLut.legal[s[i + 3]] & kLut.legal[s[i + 4]] & kLut.legal[s[i + 5]] & kLut.legal[s[i + 6]] &
              kLut.legal[s[i + 7]]))
            break;
    while (i < n && kLut.legal[s[i]])
        ++i;
    return i;
}

inline bool all_legal(const uint8_t *s, uint64_t n)
{
    bool ok = true;
    for (uint64_t i = 0; i < n; ++i)
        ok &= kLut.legal[s[i]];
    return ok;
}
} // namespace

bool block_is_fasta(const char *b, uint64_t len) { return len > 0 && (b[0] == '>' || b[0] == ';'); }

void index_reads_host(const char *bc, uint64_t len, bool final, uint64_t max_records, RecTable &t)
{
    const uint8_t *b = reinterpret_cast<const uint8_t *>(bc);
    t.clear();
    if (len == 0)
        return;
    const bool fasta = block_is_fasta(bc, len);
    uint64_t   p     = 0;
    auto       error = [&](const char *msg) {
        t.parse_error  = true;
        t.error_record = t.size();
        t.error_msg    = msg;
    };
    auto push = [&](uint64_t id_off, uint64_t id_len, uint64_t seq_off, uint64_t seq_len) {
        t.id_off.push_back((uint32_t)id_off);
        t.id_len.push_back((uint32_t)id_len);
        t.seq_off.push_back((uint32_t)seq_off);
        t.seq_len.push_back((uint32_t)seq_len);
    };
    while (p < len && t.size() < max_records)
    {
        const uint64_t rec_start = p;
        if (fasta)
        {
            if (b[p] != '>' && b[p] != ';')
            {
                error("Expected to be on beginning of ID");
                break;
            }
            const uint8_t *nl = (const uint8_t *)memchr(b + p, '\n', len - p);
            if (!nl)
            {
                if (final)
                    error("FASTA ID line did not end in newline.");
                break;
            }
            uint64_t id0 = p + 1;
            while (id0 < (uint64_t)(nl - b) && (b[id0] == ' ' || b[id0] == '\t'))
                ++id0; // fasta_ignore_blanks_before_id
            const uint64_t id1 = (uint64_t)(nl - b);
            uint64_t       q   = id1 + 1;
            if (q >= len)
            {
                if (final)
                    error("No sequence information given!");
                break;
            }
            // sequence: up to the next '>' / ';' (anywhere), blanks and digits skipped.  One pass over runs of letters: a record
            // whose letters form a single run stays where it is; from the second run on (wrapped lines) the runs are
            // gathered in aux.
            const uint64_t s0 = q;
            uint64_t       n  = legal_run(b + q, len - q);
            uint64_t       e  = q + n;
            const uint64_t a0 = t.aux.size();
            bool           gathered = false, bad = false;
            for (;;)
            {
                while (e < len && (kLut.space[b[e]] || kLut.digit[b[e]]))
                    ++e;
                if (e >= len || b[e] == '>' || b[e] == ';')
                    break;
                if (!kLut.legal[b[e]])
                {
                    bad = true;
                    break;
                }
                if (!gathered)
                {
                    gathered = true;
                    if (t.aux.capacity() < a0 + (len - s0))
                        t.aux.reserve(a0 + (len - s0));
                    t.aux.insert(t.aux.end(), b + s0, b + s0 + n);
                }
                const uint64_t r = legal_run(b + e, len - e);
                t.aux.insert(t.aux.end(), b + e, b + e + r);
                n += r;
                e += r;
            }
            if (bad || (e >= len && !final))
            {
                t.aux.resize(a0);
                if (bad)
                    error("Encountered an unexpected letter");
                break; // (not final: the record may continue in the next block)
            }
            if (!gathered)
                push(id0, id1 - id0, s0, n);
            else
                push(id0, id1 - id0, len + a0, n);
            p = e;
        }
        else
        {
            if (b[p] != '@')
            {
                error("Expected '@' on beginning of ID line");
                break;
            }
            const uint8_t *nl = (const uint8_t *)memchr(b + p, '\n', len - p);
            if (!nl)
            {
                if (final)
                    error("Expected end of ID-line, got end-of-file.");
                break;
            }
            const uint64_t id0 = p + 1, id1 = (uint64_t)(nl - b);
            const uint64_t s0 = id1 + 1;
            // fast path: one sequence line of legal letters, next line starts with '+'
            uint64_t seq_off = 0, seq_len = 0, plus = 0;
            bool     have = false;
            {
                const uint8_t *nl2 = s0 < len ? (const uint8_t *)memchr(b + s0, '\n', len - s0) : nullptr;
                if (nl2 && (uint64_t)(nl2 - b) + 1 < len && nl2[1] == '+' && all_legal(b + s0, (uint64_t)(nl2 - b) - s0))
                {
                    seq_off = s0;
                    seq_len = (uint64_t)(nl2 - b) - s0;
                    plus    = (uint64_t)(nl2 - b) + 1;
                    have    = true;
                }
            }
            if (!have)
            {
                // general path: letters up to the first '+', blanks skipped
                uint64_t e = s0;
                uint64_t n = 0;
                bool     bad = false;
                for (; e < len && b[e] != '+'; ++e)
                {
                    const uint8_t c = b[e];
                    if (kLut.space[c])
                        continue;
                    if (!kLut.legal[c])
                    {
                        bad = true;
                        break;
                    }
                    ++n;
                }
                if (bad)
                {
                    error("Encountered bad letter for seq");
                    break;
                }
                if (e >= len)
                {
                    if (final)
                        error("Expected second ID-line, got end-of-file.");
                    break;
                }
                const uint64_t a0 = t.aux.size();
                for (uint64_t i = s0; i < e; ++i)
                    if (!kLut.space[b[i]])
                        t.aux.push_back(b[i]);
                seq_off = len + a0;
                seq_len = n;
                plus    = e;
                t.aux_records++;
            }
            const uint8_t *nl3 = (const uint8_t *)memchr(b + plus, '\n', len - plus);
            if (!nl3)
            {
                if (final)
                    error("Expected end of second ID-line, got end-of-file.");
                else if (!have)
                    t.aux.resize(t.aux.size() - seq_len), t.aux_records--;
                break;
            }
            uint64_t q = (uint64_t)(nl3 - b) + 1;
            // qualities: seq_len non-blank characters
            uint64_t need = seq_len;
            if (q + need <= len && memchr(b + q, '\n', need) == nullptr)
                q += need, need = 0;
            else
                for (; q < len && need; ++q)
                    if (!kLut.space[b[q]])
                        --need;
            if (need)
            {
                if (final)
                    error("File ended before expected number of qualities could be read.");
                else if (!have)
                    t.aux.resize(t.aux.size() - seq_len), t.aux_records--;
                break;
            }
            if (q < len)
            {
                if (b[q] != '\n')
                {
                    error("Qualitites longer than sequence.");
                    break;
                }
                ++q;
            }
            else if (!final)
            { // cannot tell yet whether the quality line is complete
                if (!have)
                    t.aux.resize(t.aux.size() - seq_len), t.aux_records--;
                break;
            }
            push(id0, id1 - id0, seq_off, seq_len);
            p = q;
        }
        (void)rec_start;
        t.rec_end.push_back((uint32_t)p);
        t.consumed = p;
    }
    if (t.parse_error)
    {
        // drop a partially compacted sequence of the failed record (aux may hold its prefix)
    }
}

} // namespace gnb
