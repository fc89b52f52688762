// seqan3's other sequence formats as a byte stream for the record reader (A0, parse_reads GC.cpp:1220-1287).
// seqan3::sequence_file_input reads EMBL, FASTA, FASTQ, GenBank and SAM, chosen by the file name
// (sequence_file/input.hpp `valid_formats`, format_embl.hpp:84 / format_genbank.hpp:85 / format_sam.hpp:126 `file_extensions`).
// FASTA and FASTQ are what the record index (K1, csrc/reads.cpp) takes; the three others are rewritten HERE, record by
// record, into the unwrapped two-line FASTA form (">id\nSEQUENCE\n") with the id and the sequence seqan3 would return:
//   EMBL     format_embl.hpp:101-201     "ID" <blanks> id up to ';' ... first "SQ", rest of that line, letters up to '/'
//                                        (blanks and digits skipped), then three characters ("//\n")
//   GenBank  format_genbank.hpp:102-181  "LOCUS" <blanks> id up to a control character, lines up to one that starts with
//                                        'O' ("ORIGIN"), letters up to '/', then every '/', blank and digit that follows
//   SAM      format_sam.hpp:270-309, 395-540  '@' lines skipped, QNAME = field 1, SEQ = field 10 ('*' = none = parse error),
//                                        FLAG and POS must be numbers
// with the legality test of the FASTA / FASTQ readers (dna15).  A record seqan3 would throw on becomes a record the FASTA
// reader fails on at the same position (an illegal letter), so the chunk-loss rule of a parse error is the reference's, and
// the stream ends there.  Where the reference does not survive at all -- seqan3 throws something ganon-classify does not
// catch (unexpected_end_of_input on a truncated record, format_error on a malformed SAM number) -- the stream ends with a
// parse error as well.  Known difference: an EMBL id that contains a line break (an ID line without ';') is a parse error
// here; the reference keeps the line break inside the id.
#include <algorithm>
#include <cstring>
#include <vector>

#include "gzstream.h"

#include "../../include/ganon_b200.h"

namespace gnb
{

int format_of_extension(std::string name)
{
    auto lower = [](std::string x) {
        for (auto &c : x)
            c = (char)tolower((unsigned char)c);
        return x;
    };
    auto ext_of = [&](const std::string &x) {
        const size_t sl = x.find_last_of('/'), dot = x.find_last_of('.');
        return dot == std::string::npos || (sl != std::string::npos && dot < sl) ? std::string() : lower(x.substr(dot + 1));
    };
    std::string e = ext_of(name);
    if (e == "gz" || e == "bgzf" || e == "bz2" || e == "zst")
    {
        name = name.substr(0, name.size() - e.size() - 1);
        e    = ext_of(name);
    }
    for (const char *x : {"fasta", "fa", "fna", "ffn", "faa", "frn", "fas"})
        if (e == x)
            return kFormatFasta;
    for (const char *x : {"fastq", "fq"})
        if (e == x)
            return kFormatFastq;
    if (e == "embl")
        return kFormatEmbl;
    for (const char *x : {"genbank", "gb", "gbk"})
        if (e == x)
            return kFormatGenbank;
    if (e == "sam")
        return kFormatSam;
    return kFormatUnknown;
}

namespace
{
struct Classes
{
    bool legal[256], space[256], digit[256], cntrl[256];
    Classes()
    {
        memset(legal, 0, sizeof legal);
        memset(space, 0, sizeof space);
        memset(digit, 0, sizeof digit);
        memset(cntrl, 0, sizeof cntrl);
        for (const char *p = "ABCDGHKMNRSTVWYUabcdghkmnrstvwyu"; *p; ++p)
            legal[(uint8_t)*p] = true;
        for (const char *p = " \t\n\v\f\r"; *p; ++p)
            space[(uint8_t)*p] = true;
        for (char c = '0'; c <= '9'; ++c)
            digit[(uint8_t)c] = true;
        for (int c = 0; c < 32; ++c)
            cntrl[c] = true;
        cntrl[127] = true;
    }
};
const Classes kC;

inline bool is_blank(uint8_t c) { return c == ' ' || c == '\t'; }

enum Outcome
{
    kRecord, // one record appended to out, `used` bytes consumed
    kMore,   // the record is not complete in the buffer (and the buffer is not the end of the file)
    kBad     // seqan3 throws on this record
};

// the record's id must fit one FASTA header line
inline bool id_ok(const uint8_t *id, size_t n) { return memchr(id, '\n', n) == nullptr; }

inline void put_record(std::string &out, const uint8_t *id, size_t id_len, const std::string &seq)
{
    out.push_back('>');
    out.append(reinterpret_cast<const char *>(id), id_len);
    out.push_back('\n');
    out.append(seq);
    out.push_back('\n');
}

// letters up to '/' with blanks and digits skipped (run by run); p ends ON the '/'
Outcome take_sequence(const uint8_t *b, size_t n, bool final, size_t &p, std::string &seq)
{
    seq.clear();
    while (p < n)
    {
        const uint8_t c = b[p];
        if (kC.space[c] || kC.digit[c])
        {
            ++p;
            continue;
        }
        if (c == '/')
            return kRecord;
        if (!kC.legal[c])
            return kBad;
        size_t q = p + 1;
        while (q < n && kC.legal[b[q]])
            ++q;
        seq.append(reinterpret_cast<const char *>(b + p), q - p);
        p = q;
    }
    return final ? kBad : kMore;
}

// through the next '\n' (detail::take_line_or_throw)
inline Outcome take_line(const uint8_t *b, size_t n, bool final, size_t &p)
{
    const void *nl = p < n ? memchr(b + p, '\n', n - p) : nullptr;
    if (!nl)
        return final ? kBad : kMore;
    p = (size_t)(static_cast<const uint8_t *>(nl) - b) + 1;
    return kRecord;
}

Outcome embl_record(const uint8_t *b, size_t n, bool final, size_t &used, std::string &out, std::string &seq)
{
    const Outcome open = final ? kBad : kMore;
    size_t        p    = 0;
    while (p < n && !(kC.cntrl[b[p]] || is_blank(b[p])))
        ++p;
    if (p == n)
        return open;
    if (p != 2 || b[0] != 'I' || b[1] != 'D')
        return kBad; // "An entry has to start with the code word ID."
    while (p < n && is_blank(b[p]))
        ++p;
    const uint8_t *semi = p < n ? static_cast<const uint8_t *>(memchr(b + p, ';', n - p)) : nullptr;
    if (!semi)
        return open;
    const uint8_t *id     = b + p;
    const size_t   id_len = (size_t)(semi - id);
    p                     = (size_t)(semi - b);
    // the first 'S' that is followed by 'Q'
    for (;;)
    {
        const uint8_t *s = p < n ? static_cast<const uint8_t *>(memchr(b + p, 'S', n - p)) : nullptr;
        if (!s)
            return open;
        p = (size_t)(s - b) + 1;
        if (p >= n)
            return open;
        if (b[p] == 'Q')
            break;
    }
    Outcome o = take_line(b, n, final, p);
    if (o != kRecord)
        return o;
    o = take_sequence(b, n, final, p, seq);
    if (o != kRecord)
        return o;
    if (p + 3 > n && !final)
        return kMore;
    if (!id_ok(id, id_len))
        return kBad;
    put_record(out, id, id_len, seq);
    used = std::min(n, p + 3); // "//" and the character after it
    return kRecord;
}

Outcome genbank_record(const uint8_t *b, size_t n, bool final, size_t &used, std::string &out, std::string &seq)
{
    const Outcome open = final ? kBad : kMore;
    size_t        p    = 0;
    while (p < n && !(kC.cntrl[b[p]] || is_blank(b[p])))
        ++p;
    if (p == n)
        return open;
    if (p != 5 || memcmp(b, "LOCUS", 5) != 0)
        return kBad; // "An entry has to start with the code word LOCUS."
    while (p < n && is_blank(b[p]))
        ++p;
    const uint8_t *id = b + p;
    while (p < n && !kC.cntrl[b[p]])
        ++p;
    if (p == n)
        return open;
    const size_t id_len = (size_t)(b + p - id);
    Outcome      o      = take_line(b, n, final, p);
    if (o != kRecord)
        return o;
    for (;;)
    {
        if (p >= n)
            return open;
        const bool origin = b[p] == 'O';
        if ((o = take_line(b, n, final, p)) != kRecord)
            return o;
        if (origin)
            break;
    }
    o = take_sequence(b, n, final, p, seq);
    if (o != kRecord)
        return o;
    // the end marker is consumed through the blank / digit filter: every '/', blank and digit up to the next other character
    while (p < n && (b[p] == '/' || kC.space[b[p]] || kC.digit[b[p]]))
        ++p;
    if (p == n && !final)
        return kMore;
    put_record(out, id, id_len, seq);
    used = p;
    return kRecord;
}

// decimal number with an optional sign that std::from_chars<int32_t> would take whole
inline bool sam_number(const uint8_t *s, size_t n, bool allow_minus, int64_t lo, int64_t hi)
{
    size_t i   = 0;
    bool   neg = false;
    if (allow_minus && i < n && s[i] == '-')
        neg = true, ++i;
    if (i == n)
        return false;
    int64_t v = 0;
    for (; i < n; ++i)
    {
        if (!kC.digit[s[i]])
            return false;
        v = v * 10 + (s[i] - '0');
        if (v > (1ll << 40))
            return false;
    }
    v = neg ? -v : v;
    return v >= lo && v <= hi;
}

Outcome sam_record(const uint8_t *b, size_t n, bool final, size_t &used, std::string &out, std::string &seq)
{
    size_t p = 0;
    // header lines, wherever a record starts with '@' (format_sam.hpp:395-401)
    while (p < n && b[p] == '@')
    {
        const void *nl = memchr(b + p, '\n', n - p);
        if (!nl)
        {
            if (!final)
                return kMore;
            p = n;
            break;
        }
        p = (size_t)(static_cast<const uint8_t *>(nl) - b) + 1;
    }
    if (p == n)
        return final ? kBad : kMore; // a header without a record after it: "The sequence information must not be empty."
    const uint8_t *nl  = static_cast<const uint8_t *>(memchr(b + p, '\n', n - p));
    if (!nl && !final)
        return kMore;
    const size_t   end = nl ? (size_t)(nl - b) : n;
    const uint8_t *f[11];
    size_t         fl[11];
    size_t         q = p;
    for (int i = 0; i < 10; ++i)
    {
        const uint8_t *tab = q < end ? static_cast<const uint8_t *>(memchr(b + q, '\t', end - q)) : nullptr;
        if (!tab)
            return kBad; // fewer than eleven fields
        f[i]  = b + q;
        fl[i] = (size_t)(tab - (b + q));
        q     = (size_t)(tab - b) + 1;
    }
    if (fl[0] == 0)
        return kBad; // "The id information must not be empty."
    if (!sam_number(f[1], fl[1], false, 0, 65535) || !sam_number(f[3], fl[3], true, 0, 2147483647))
        return kBad; // FLAG (uint16_t), POS (int32_t, not negative)
    if (fl[9] == 0 || f[9][0] == '*')
        return kBad; // "The sequence information must not be empty."
    seq.assign(reinterpret_cast<const char *>(f[9]), fl[9]);
    for (size_t i = 0; i < fl[9]; ++i)
        if (!kC.legal[f[9][i]])
            return kBad;
    put_record(out, f[0], fl[0], seq);
    used = nl ? end + 1 : n;
    return kRecord;
}

class TranscodeSource : public ByteSource
{
  public:
    TranscodeSource(std::unique_ptr<ByteSource> inner, int format) : in_(std::move(inner)), format_(format) {}
    uint64_t size() const override { return in_->size(); } // an upper bound: the FASTA form is never longer than the record
    bool     is_gzip() const override { return in_->is_gzip(); }
    int64_t  read(char *dst, size_t cap) override
    {
        size_t got = 0;
        while (got < cap)
        {
            if (out_off_ < out_.size())
            {
                const size_t n = std::min(cap - got, out_.size() - out_off_);
                memcpy(dst + got, out_.data() + out_off_, n);
                out_off_ += n;
                got += n;
                continue;
            }
            if (done_)
                break;
            out_.clear();
            out_off_ = 0;
            const int rc = produce();
            if (rc < 0)
                return got ? (int64_t)got : (int64_t)rc;
        }
        return (int64_t)got;
    }

  private:
    // more input (unless the file has ended), then every complete record of the buffer
    int produce()
    {
        if (!eof_)
        {
            if (pos_ > 0)
            {
                buf_.erase(buf_.begin(), buf_.begin() + (long)pos_);
                pos_ = 0;
            }
            const size_t have = buf_.size(), want = std::max<size_t>(8u << 20, have);
            buf_.resize(have + want);
            size_t fill = have;
            while (fill < buf_.size())
            {
                const int64_t n = in_->read(reinterpret_cast<char *>(buf_.data()) + fill, buf_.size() - fill);
                if (n < 0)
                {
                    err_ = in_->error();
                    return (int)n;
                }
                if (n == 0)
                {
                    eof_ = true;
                    break;
                }
                fill += (size_t)n;
            }
            buf_.resize(fill);
        }
        while (pos_ < buf_.size())
        {
            size_t        used = 0;
            const uint8_t *b = buf_.data() + pos_;
            const size_t   n = buf_.size() - pos_;
            const Outcome  o = format_ == kFormatEmbl      ? embl_record(b, n, eof_, used, out_, seq_)
                               : format_ == kFormatGenbank ? genbank_record(b, n, eof_, used, out_, seq_)
                                                           : sam_record(b, n, eof_, used, out_, seq_);
            if (o == kMore)
                break;
            if (o == kBad)
            {
                out_.append(">\n!\n"); // a record the FASTA reader fails on: the parse error, at this record
                done_ = true;
                return 0;
            }
            pos_ += used;
        }
        if (eof_ && pos_ >= buf_.size())
            done_ = true;
        return 0;
    }
    std::unique_ptr<ByteSource> in_;
    int                         format_;
    std::vector<uint8_t>        buf_;
    size_t                      pos_ = 0;
    std::string                 out_, seq_;
    size_t                      out_off_ = 0;
    bool                        eof_ = false, done_ = false;
};
} // namespace

std::unique_ptr<ByteSource> wrap_sequence_format(std::unique_ptr<ByteSource> src, const std::string &path)
{
    const int f = format_of_extension(path);
    if (f == kFormatEmbl || f == kFormatGenbank || f == kFormatSam)
        return std::unique_ptr<ByteSource>(new TranscodeSource(std::move(src), f));
    return src;
}

} // namespace gnb
