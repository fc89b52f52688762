// See gzstream.h.  Deflate (RFC 1951) and gzip (RFC 1952) decoding written for chunk-parallel inflation; zlib is used for
// crc32 / crc32_combine only.
#include "gzstream.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ganon_b200.h"

namespace gnb
{
namespace
{

// ---------------------------------------------------------------------------------------------------------------------
// a small fork-join pool: parallel_for(n, fn) runs fn(0..n-1) on the workers and the calling thread
// ---------------------------------------------------------------------------------------------------------------------
class Pool
{
  public:
    explicit Pool(int n_threads)
    {
        for (int i = 1; i < std::max(1, n_threads); ++i)
            workers_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_)
            t.join();
    }
    int  size() const { return (int)workers_.size() + 1; }
    void parallel_for(size_t n, const std::function<void(size_t)> &fn)
    {
        if (n == 0)
            return;
        if (n == 1 || workers_.empty())
        {
            for (size_t i = 0; i < n; ++i)
                fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_   = &fn;
            n_    = n;
            next_ = 0;
            done_ = 0;
            ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(mu_);
        done_cv_.wait(l, [&] { return done_ == n_; });
        fn_ = nullptr;
    }

  private:
    void work()
    {
        for (;;)
        {
            size_t i;
            {
                std::lock_guard<std::mutex> l(mu_);
                if (!fn_ || next_ >= n_)
                    return;
                i = next_++;
            }
            (*fn_)(i);
            {
                std::lock_guard<std::mutex> l(mu_);
                if (++done_ == n_)
                    done_cv_.notify_all();
            }
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || gen_ != seen; });
                if (stop_)
                    return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread>              workers_;
    std::mutex                            mu_;
    std::condition_variable               cv_, done_cv_;
    const std::function<void(size_t)>    *fn_ = nullptr;
    size_t                                n_ = 0, next_ = 0, done_ = 0;
    uint64_t                              gen_ = 0;
    bool                                  stop_ = false;
};

inline bool pread_all(int fd, char *dst, size_t n, uint64_t off)
{
    size_t done = 0;
    while (done < n)
    {
        ssize_t r = pread(fd, dst + done, n - done, (off_t)(off + done));
        if (r <= 0)
            return false;
        done += (size_t)r;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// plain files: parallel preads (page-cache copies are what bounds a single reader thread)
// ---------------------------------------------------------------------------------------------------------------------
class PlainSource : public ByteSource
{
  public:
    PlainSource(int fd, uint64_t size, int threads) : fd_(fd), size_(size), pool_(threads) {}
    ~PlainSource() override { close(fd_); }
    int64_t read(char *dst, size_t cap) override
    {
        const int64_t n = read_at(dst, cap, pos_);
        if (n > 0)
            pos_ += (uint64_t)n;
        return n;
    }
    bool    seekable() const override { return true; }
    uint64_t size() const override { return size_; }
    int64_t read_at(char *dst, size_t cap, uint64_t offset) override
    {
        if (offset >= size_)
            return 0;
        const size_t want  = (size_t)std::min<uint64_t>(cap, size_ - offset);
        const size_t slice = 4u << 20;
        const size_t parts = (want + slice - 1) / slice;
        std::atomic<bool> ok{true};
        pool_.parallel_for(parts, [&](size_t i) {
            const size_t o = i * slice, n = std::min(slice, want - o);
            if (!pread_all(fd_, dst + o, n, offset + o))
                ok = false;
        });
        if (!ok)
        {
            err_ = "short read";
            return GNB_ERR_IO;
        }
        return (int64_t)want;
    }

  private:
    int      fd_;
    uint64_t size_, pos_ = 0;
    Pool     pool_;
};

// ---------------------------------------------------------------------------------------------------------------------
// deflate
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint16_t kLenBase[29]  = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
constexpr uint8_t  kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
constexpr uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
constexpr uint8_t  kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
constexpr uint8_t  kClOrder[19]  = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
constexpr uint32_t kWindow       = 32768;
constexpr uint16_t kMarker       = 0x8000; // symbol >= kMarker: byte (symbol - kMarker) of the unknown 32 KiB history

inline uint64_t load64(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}

// canonical Huffman code: a direct table over the next `fast` stream bits, the bit-serial canonical walk beyond it
struct Huff
{
    uint16_t count[16];
    uint16_t symbol[288];
    uint16_t table[1 << 11]; // (symbol << 4) | code length; 0 = longer than `fast` bits
    int      fast = 0, max_len = 0;

    // 0 = complete, 1 = incomplete, -1 = over-subscribed
    int build(const uint8_t *len, int n, int fast_bits)
    {
        fast = fast_bits;
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i)
            count[len[i]]++;
        max_len = 0;
        for (int l = 15; l >= 1; --l)
            if (count[l])
            {
                max_len = l;
                break;
            }
        int left = 1;
        for (int l = 1; l <= 15; ++l)
        {
            left <<= 1;
            left -= count[l];
            if (left < 0)
                return -1;
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; ++l)
            offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; ++i)
            if (len[i])
                symbol[offs[len[i]]++] = (uint16_t)i;
        memset(table, 0, sizeof(uint16_t) << fast);
        uint32_t code = 0;
        uint32_t idx  = 0;
        for (int l = 1; l <= 15 && l <= fast; ++l)
        {
            for (uint32_t k = 0; k < count[l]; ++k, ++code, ++idx)
            {
                uint32_t r = 0;
                for (int b = 0; b < l; ++b)
                    r |= ((code >> b) & 1u) << (l - 1 - b);
                const uint16_t e = (uint16_t)((symbol[idx] << 4) | l);
                for (uint32_t t = r; t < (1u << fast); t += 1u << l)
                    table[t] = e;
            }
            code <<= 1;
        }
        count[0] = 0;
        return left > 0 ? 1 : 0;
    }
    // code of more than `fast` bits (or none): bit-serial walk; returns the symbol and sets used, or -1
    int slow(uint64_t bits, int &used) const
    {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; ++l)
        {
            code |= (int)((bits >> (l - 1)) & 1);
            const int c = count[l];
            if (code - c < first)
            {
                used = l;
                return symbol[index + (code - first)];
            }
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

struct BitIn
{
    const uint8_t *base = nullptr;
    uint64_t       total_bits = 0; // valid bits of the buffer (the buffer is padded with >= 16 readable bytes)
    uint64_t       pos = 0;        // absolute bit position of the next unread bit
    uint64_t peek() const { return load64(base + (pos >> 3)) >> (pos & 7); } // >= 57 bits
    uint32_t get(int n)
    {
        const uint32_t v = (uint32_t)(peek() & ((1ull << n) - 1));
        pos += (unsigned)n;
        return v;
    }
};

// Packed tables for the decode loop, indexed by the next 11 (literal/length) or 9 (distance) stream bits:
//   literal/length: bits 0-3 code length (0 = longer code: bit-serial walk), bits 4-5 kind (0 literal, 1 end of block,
//                   2 length, 3 invalid symbol), bits 8-23 literal byte or length base, bits 24-27 extra bits
//   distance:       bits 0-3 code length, bits 4-7 extra bits, bits 8-23 base, bit 31 invalid symbol
struct DynHeader
{
    Huff     lit, dist;
    bool     dist_empty = false;
    uint32_t ltab[1 << 11], dtab[1 << 9];
    static uint32_t pack_lit(int sym, int len)
    {
        if (sym < 256)
            return (uint32_t)len | ((uint32_t)sym << 8);
        if (sym == 256)
            return (uint32_t)len | (1u << 4);
        if (sym > 285)
            return (uint32_t)len | (3u << 4);
        return (uint32_t)len | (2u << 4) | ((uint32_t)kLenBase[sym - 257] << 8) | ((uint32_t)kLenExtra[sym - 257] << 24);
    }
    static uint32_t pack_dist(int sym, int len)
    {
        if (sym > 29)
            return (uint32_t)len | (1u << 31);
        return (uint32_t)len | ((uint32_t)kDistExtra[sym] << 4) | ((uint32_t)kDistBase[sym] << 8);
    }
    void pack()
    {
        for (int i = 0; i < (1 << 11); ++i)
        {
            const uint16_t e = lit.table[i];
            ltab[i]          = (e & 15) ? pack_lit(e >> 4, e & 15) : 0;
        }
        for (int i = 0; i < (1 << 9); ++i)
        {
            const uint16_t e = dist.table[i];
            dtab[i]          = (e & 15) ? pack_dist(e >> 4, e & 15) : 0;
        }
    }
};

// Parses a dynamic-Huffman block header (after the 3 block bits) with zlib's validity rules (inflate.c / inftrees.c):
// the code-length code must be complete, the literal/length code must contain the end-of-block symbol and be complete
// (or be a single 1-bit code), the distance code likewise or empty.  Returns false on any violation.
bool read_dynamic_header(BitIn &in, DynHeader &h)
{
    uint64_t  w     = in.peek();
    const int hlit  = (int)(w & 31) + 257;
    const int hdist = (int)((w >> 5) & 31) + 1;
    const int hclen = (int)((w >> 10) & 15) + 4;
    if (hlit > 286 || hdist > 30)
        return false;
    in.pos += 14;
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i)
        cl[kClOrder[i]] = (uint8_t)in.get(3);
    Huff clh;
    if (clh.build(cl, 19, 7) != 0)
        return false;
    uint8_t len[286 + 30 + 8];
    int     n = 0;
    while (n < hlit + hdist)
    {
        if (in.pos > in.total_bits)
            return false;
        const uint64_t bits = in.peek();
        const uint16_t e    = clh.table[bits & 127];
        if ((e & 15) == 0)
            return false; // code-length codes are at most 7 bits: every complete code fills the table
        in.pos += e & 15;
        const int sym = e >> 4;
        if (sym < 16)
            len[n++] = (uint8_t)sym;
        else
        {
            int     rep;
            uint8_t v = 0;
            if (sym == 16)
            {
                if (n == 0)
                    return false;
                v   = len[n - 1];
                rep = 3 + (int)in.get(2);
            }
            else if (sym == 17)
                rep = 3 + (int)in.get(3);
            else
                rep = 11 + (int)in.get(7);
            if (n + rep > hlit + hdist)
                return false;
            while (rep--)
                len[n++] = v;
        }
    }
    if (len[256] == 0)
        return false;
    int rc = h.lit.build(len, hlit, 11);
    if (rc < 0 || (rc > 0 && h.lit.max_len != 1))
        return false;
    rc = h.dist.build(len + hlit, hdist, 9);
    h.dist_empty = h.dist.max_len == 0;
    if (rc < 0 || (rc > 0 && h.dist.max_len > 1))
        return false;
    return in.pos <= in.total_bits;
}

const DynHeader &fixed_header()
{
    static const DynHeader h = [] {
        DynHeader f;
        uint8_t   l[288];
        for (int i = 0; i < 144; ++i)
            l[i] = 8;
        for (int i = 144; i < 256; ++i)
            l[i] = 9;
        for (int i = 256; i < 280; ++i)
            l[i] = 7;
        for (int i = 280; i < 288; ++i)
            l[i] = 8;
        f.lit.build(l, 288, 11);
        uint8_t d[30];
        for (int i = 0; i < 30; ++i)
            d[i] = 5;
        f.dist.build(d, 30, 9); // 30 of 32 five-bit codes: incomplete by design
        f.pack();
        return f;
    }();
    return h;
}

struct MemberEnd
{
    uint64_t out_off; // output bytes of the chunk up to the member's end
    uint32_t crc, isize;
};

struct Chunk
{
    std::vector<uint16_t> sym; // kWindow marker symbols, then the output
    size_t                n_out = 0;
    uint64_t              start_bit = 0, end_bit = 0;
    bool                  ok = false, eos = false, found = false;
    bool                  ran_out = false; // failed because the buffered data ended (not an error if the file goes on)
    std::string           err;
    std::vector<MemberEnd> ends;
};

// a decoded chunk on its way to the consumer: symbols with markers, the 32 KiB history that resolves them
struct Piece
{
    std::vector<uint16_t>  sym;
    size_t                 n_out = 0;
    std::vector<uint8_t>   window;
    std::vector<MemberEnd> ends;
    std::vector<std::pair<uint64_t, uint32_t>> crcs; // (length, crc) of the runs between member ends, filled by the consumer
};

// gzip member header (RFC 1952 2.3); in.pos must be byte aligned.  false = not a gzip header / truncated
bool read_gzip_header(BitIn &in)
{
    const uint64_t nbytes = in.total_bits >> 3;
    uint64_t       p      = in.pos >> 3;
    if (p + 10 > nbytes)
        return false;
    const uint8_t *b = in.base;
    if (b[p] != 0x1f || b[p + 1] != 0x8b || b[p + 2] != 8)
        return false;
    const uint8_t flg = b[p + 3];
    p += 10;
    if (flg & 4)
    {
        if (p + 2 > nbytes)
            return false;
        p += 2 + (b[p] | (b[p + 1] << 8));
    }
    for (int f : {8, 16})
        if (flg & f)
        {
            while (p < nbytes && b[p])
                ++p;
            ++p;
        }
    if (flg & 2)
        p += 2;
    if (p > nbytes)
        return false;
    in.pos = p << 3;
    return true;
}

// Decodes deflate blocks from in.pos into c.sym (16-bit symbols, unknown history = markers) until
//   * a block boundary that equals one of `cands` (ascending bit positions; the ones run past are false positives), or
//   * the first block boundary at or after limit_bit, or
//   * the end of the data (after the last member's trailer).
// at_header: in.pos is at a gzip member header (the start of the file).
void decode_chunk(BitIn in, bool at_header, const uint64_t *cands, size_t n_cands, uint64_t limit_bit, bool last_data, Chunk &c)
{
    if (c.sym.size() < kWindow + (4u << 20))
        c.sym.resize(kWindow + (4u << 20));
    for (uint32_t i = 0; i < kWindow; ++i)
        c.sym[i] = (uint16_t)(kMarker + i);
    uint16_t *out = c.sym.data() + kWindow;
    size_t    o = 0, cap = c.sym.size() - kWindow;
    c.start_bit = in.pos;
    size_t ci   = 0;
    auto   fail = [&](const char *m) {
        c.ok      = false;
        c.err     = m;
        c.n_out   = o;
        c.ran_out = strncmp(m, "unexpected end", 14) == 0 || strncmp(m, "truncated", 9) == 0 || strstr(m, "edge of the buffered") != nullptr;
    };
    auto room = [&](size_t need) {
        if (o + need > cap)
        {
            c.sym.resize(kWindow + std::max(cap * 2, o + need + (1u << 20)));
            out = c.sym.data() + kWindow;
            cap = c.sym.size() - kWindow;
        }
    };
    bool need_header = at_header;
    for (;;)
    {
        if (need_header)
        {
            if (!read_gzip_header(in))
                return fail("not a gzip member header");
            need_header = false;
        }
        // ---- block boundary: hand over? ----
        while (ci < n_cands && cands[ci] < in.pos)
            ++ci;
        if (in.pos != c.start_bit && ((ci < n_cands && cands[ci] == in.pos) || in.pos >= limit_bit))
            break;
        if (in.pos + 3 > in.total_bits)
            return fail("unexpected end of the gzip stream");
        const uint32_t hdr    = in.get(3);
        const bool     bfinal = hdr & 1;
        const uint32_t btype  = hdr >> 1;
        if (btype == 3)
            return fail("invalid deflate block type");
        if (btype == 0)
        {
            in.pos = (in.pos + 7) & ~7ull;
            if (in.pos + 32 > in.total_bits)
                return fail("unexpected end of the gzip stream");
            const uint32_t len = in.get(16), nlen = in.get(16);
            if ((len ^ nlen) != 0xffff)
                return fail("invalid stored block lengths");
            if (in.pos + (uint64_t)len * 8 > in.total_bits)
                return fail("unexpected end of the gzip stream");
            room(len);
            const uint8_t *src = in.base + (in.pos >> 3);
            for (uint32_t i = 0; i < len; ++i)
                out[o + i] = src[i];
            o += len;
            in.pos += (uint64_t)len * 8;
        }
        else
        {
            DynHeader        dyn;
            const DynHeader *H = &fixed_header();
            if (btype == 2)
            {
                if (!read_dynamic_header(in, dyn))
                    return fail("invalid dynamic Huffman header");
                H = &dyn;
            }
            if (btype == 2)
                dyn.pack();
            const Huff     &L = H->lit, &D = H->dist;
            const uint32_t *ltab = H->ltab, *dtab = H->dtab;
            // bit buffer: `cnt` valid bits in `bb`, next unread byte at p; stream position = (p - base) * 8 - cnt
            const uint8_t *p   = in.base + (in.pos >> 3);
            uint64_t       bb  = load64(p) >> (in.pos & 7);
            int            cnt = 64 - (int)(in.pos & 7);
            p += 8;
            const uint8_t *const p_end = in.base + (in.total_bits >> 3) + 16; // the buffer's padding keeps loads legal
#define GNB_REFILL()                       \
    do                                     \
    {                                      \
        bb |= load64(p) << cnt;            \
        p += (63 - cnt) >> 3;              \
        cnt |= 56;                         \
    } while (0)
            // (the first load above took 8 bytes; cnt may be 57..64: bring it into the refill's invariant cnt <= 63)
            if (cnt == 64)
            {
                cnt = 56;
                p -= 1;
                bb &= (1ull << 56) - 1;
            }
            for (;;)
            {
                if (o + 1024 > cap)
                    room(1024 + (1u << 16));
                if (p > p_end)
                    return fail("unexpected end of the gzip stream");
                GNB_REFILL();
                uint32_t e = ltab[bb & 2047];
                // up to three literals per refill (3 x 15 bits <= 56)
                if ((e & 0x3f) != 0 && (e & 0x30) == 0)
                {
                    bb >>= e & 15;
                    cnt -= (int)(e & 15);
                    out[o++] = (uint16_t)(e >> 8);
                    e        = ltab[bb & 2047];
                    if ((e & 0x3f) != 0 && (e & 0x30) == 0)
                    {
                        bb >>= e & 15;
                        cnt -= (int)(e & 15);
                        out[o++] = (uint16_t)(e >> 8);
                        e        = ltab[bb & 2047];
                        if ((e & 0x3f) != 0 && (e & 0x30) == 0)
                        {
                            bb >>= e & 15;
                            cnt -= (int)(e & 15);
                            out[o++] = (uint16_t)(e >> 8);
                            continue;
                        }
                    }
                    GNB_REFILL();
                }
                int used = (int)(e & 15);
                if (used == 0)
                {
                    const int sym = L.slow(bb, used);
                    if (sym < 0)
                        return fail("invalid literal/length code");
                    e = DynHeader::pack_lit(sym, used);
                }
                bb >>= used;
                cnt -= used;
                const uint32_t kind = (e >> 4) & 3;
                if (kind == 0)
                {
                    out[o++] = (uint16_t)(e >> 8);
                    continue;
                }
                if (kind == 1)
                    break;
                if (kind == 3)
                    return fail("invalid literal/length symbol");
                const int      xb  = (int)(e >> 24) & 15;
                const uint32_t len = ((e >> 8) & 0xffff) + (uint32_t)(bb & ((1u << xb) - 1));
                bb >>= xb;
                cnt -= xb;
                uint32_t d = dtab[bb & 511];
                used       = (int)(d & 15);
                if (used == 0)
                {
                    const int dsym = H->dist_empty ? -1 : D.slow(bb, used);
                    if (dsym < 0)
                        return fail("invalid distance code");
                    d = DynHeader::pack_dist(dsym, used);
                }
                if (d >> 31)
                    return fail("invalid distance symbol");
                bb >>= used;
                cnt -= used;
                const int      dx   = (int)(d >> 4) & 15;
                const uint32_t dist = ((d >> 8) & 0xffff) + (uint32_t)(bb & ((1u << dx) - 1));
                bb >>= dx;
                cnt -= dx;
                // dist <= 32768 <= o + kWindow: the marker prefix makes every legal distance addressable
                const uint16_t *sp = out + o - dist;
                uint16_t       *dp = out + o;
                if (dist >= len)
                    memcpy(dp, sp, (size_t)len * 2);
                else if (dist == 1)
                {
                    const uint16_t v = sp[0];
                    for (uint32_t i = 0; i < len; ++i)
                        dp[i] = v;
                }
                else
                    for (uint32_t i = 0; i < len; ++i)
                        dp[i] = sp[i];
                o += len;
            }
#undef GNB_REFILL
            const uint64_t pos = (uint64_t)(p - in.base) * 8 - (uint64_t)cnt;
            in.pos = pos;
            if (in.pos > in.total_bits)
                return fail("unexpected end of the gzip stream");
        }
        if (bfinal)
        {
            in.pos = (in.pos + 7) & ~7ull;
            if (in.pos + 64 > in.total_bits)
                return fail("truncated gzip trailer");
            const uint32_t crc = in.get(32), isize = in.get(32);
            c.ends.push_back(MemberEnd{o, crc, isize});
            // another member, padding, or the end
            const uint64_t p = in.pos >> 3, nbytes = in.total_bits >> 3;
            if (p + 2 <= nbytes && in.base[p] == 0x1f && in.base[p + 1] == 0x8b)
                need_header = true;
            else if (last_data || p >= nbytes)
            {
                // end of the data, or trailing bytes that are not a member (gzip ignores them as well)
                if (!last_data && p >= nbytes)
                    return fail("gzip member ends at the edge of the buffered data");
                c.eos     = true;
                in.pos    = in.total_bits;
                break;
            }
            else
                return fail("data after a gzip member is not a gzip member");
        }
    }
    c.end_bit = in.pos;
    c.n_out   = o;
    c.ok      = true;
}

// First bit position in [from_bit, to_bit) where a block with a valid dynamic-Huffman header starts and decodes cleanly
// to its end-of-block symbol, followed by a plausible next block.  ~0 = none.
uint64_t find_block_start(const uint8_t *base, uint64_t total_bits, uint64_t from_bit, uint64_t to_bit)
{
    DynHeader h;
    for (uint64_t b = from_bit; b < to_bit && b + 64 < total_bits; ++b)
    {
        const uint64_t w = load64(base + (b >> 3)) >> (b & 7);
        // BTYPE = 10 (dynamic), HLIT <= 29, HDIST <= 29
        if (((w >> 1) & 3) != 2 || ((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29)
            continue;
        // quick look at the code-length code: complete (Kraft sum exactly 1)
        const int hclen = (int)((w >> 13) & 15) + 4;
        int       left  = 128;
        {
            uint64_t v = w >> 17;
            int      i = 0;
            // 57 bits are available in w: 17 header bits + 13 code lengths; the rest comes from a second load
            for (; i < hclen && i < 13 && left >= 0; ++i, v >>= 3)
            {
                const int l = (int)(v & 7);
                if (l)
                    left -= 128 >> l;
            }
            if (left > 0 && i < hclen)
            {
                const uint64_t b2 = b + 17 + 39;
                uint64_t       v2 = load64(base + (b2 >> 3)) >> (b2 & 7);
                for (; i < hclen; ++i, v2 >>= 3)
                {
                    const int l = (int)(v2 & 7);
                    if (l)
                        left -= 128 >> l;
                }
            }
        }
        if (left != 0)
            continue;
        BitIn in;
        in.base       = base;
        in.total_bits = total_bits;
        in.pos        = b + 3;
        if (!read_dynamic_header(in, h))
            continue;
        // trial decode of the block (symbols only)
        const Huff &L = h.lit, &D = h.dist;
        uint64_t    pos = in.pos;
        bool        good = false;
        uint64_t    n_sym = 0;
        for (;;)
        {
            if (pos + 48 > total_bits)
                break;
            uint64_t bits = load64(base + (pos >> 3)) >> (pos & 7);
            uint32_t e    = L.table[bits & 2047];
            int      used = (int)(e & 15), sym = (int)(e >> 4);
            if (used == 0)
            {
                sym = L.slow(bits, used);
                if (sym < 0)
                    break;
            }
            pos += (unsigned)used;
            ++n_sym;
            if (sym < 256)
                continue;
            if (sym == 256)
            {
                good = true;
                break;
            }
            if (sym > 285)
                break;
            pos += kLenExtra[sym - 257];
            bits = load64(base + (pos >> 3)) >> (pos & 7);
            e    = D.table[bits & 511];
            used = (int)(e & 15);
            int dsym = (int)(e >> 4);
            if (used == 0)
            {
                dsym = h.dist_empty ? -1 : D.slow(bits, used);
                if (dsym < 0)
                    break;
            }
            if (dsym > 29)
                break;
            pos += (unsigned)(used + kDistExtra[dsym]);
        }
        if (!good || n_sym < 16)
            continue;
        // the block that follows must not be of the reserved type (a final block is followed by a trailer instead)
        if (!(w & 1))
        {
            const uint64_t nx = load64(base + (pos >> 3)) >> (pos & 7);
            if (((nx >> 1) & 3) == 3)
                continue;
        }
        return b;
    }
    return ~0ull;
}

// symbols -> bytes: literals as they are, markers through the (now known) 32 KiB history.  Groups of 16 symbols without a
// marker are narrowed with SSE2; the others go through a 64 Ki-entry table (symbol -> byte), one load per symbol.
inline void resolve_markers(const uint16_t *s, size_t n, const uint8_t *w, uint8_t *d, std::vector<uint8_t> &lut)
{
    lut.resize(65536);
    for (int i = 0; i < 256; ++i)
        lut[i] = (uint8_t)i;
    memcpy(lut.data() + kMarker, w, kWindow);
    const uint8_t *t = lut.data();
    size_t         i = 0;
#if defined(__SSE2__)
    for (; i + 16 <= n; i += 16)
    {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 8));
        if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA)
            for (size_t j = i; j < i + 16; ++j)
                d[j] = t[s[j]];
        else
            _mm_storeu_si128(reinterpret_cast<__m128i *>(d + i), _mm_packus_epi16(a, b));
    }
#endif
    for (; i < n; ++i)
        d[i] = t[s[i]];
}

// ---------------------------------------------------------------------------------------------------------------------
// gzip files
// ---------------------------------------------------------------------------------------------------------------------
class GzSource : public ByteSource
{
  public:
    GzSource(int fd, uint64_t size, int threads) : fd_(fd), size_(size), pool_(threads), pool2_(std::max(1, std::min(threads, 8)))
    {
        chunk_bytes_ = 2u << 20;
        if (const char *e = getenv("GANON_B200_GZ_CHUNK"))
            chunk_bytes_ = std::max<uint64_t>(1u << 12, strtoull(e, nullptr, 10));
        wave_chunks_ = (size_t)pool_.size() * 2;
        window_.assign(kWindow, 0);
        producer_ = std::thread([this] { produce(); });
    }
    ~GzSource() override
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            cancel_ = true;
        }
        cv_.notify_all();
        producer_.join();
        close(fd_);
    }
    bool     is_gzip() const override { return true; }
    uint64_t size() const override { return size_; }
    // Markers are replaced HERE, by the consumer, straight into the caller's buffer (in parallel over the queued pieces):
    // the decoded bytes are written once, where they are needed.  Only a piece that straddles the end of the caller's
    // buffer goes through a side buffer.
    int64_t read(char *dst, size_t cap) override
    {
        size_t got = 0;
        if (carry_off_ < carry_.size())
        {
            const size_t n = std::min(cap, carry_.size() - carry_off_);
            memcpy(dst, carry_.data() + carry_off_, n);
            carry_off_ += n;
            got = n;
        }
        while (got < cap)
        {
            std::vector<Piece> batch;
            size_t             room = cap - got;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return !ready_.empty() || finished_; });
                if (ready_.empty())
                {
                    if (perr_.empty() && len_run_ != 0)
                        perr_ = "gzip stream ends inside a member";
                    if (!perr_.empty())
                    {
                        err_ = perr_;
                        return got ? (int64_t)got : (int64_t)GNB_ERR_IO;
                    }
                    break;
                }
                while (!ready_.empty() && batch.size() < 64 && (batch.empty() || ready_.front().n_out <= room))
                {
                    room -= std::min(room, ready_.front().n_out);
                    queued_ -= ready_.front().n_out;
                    batch.emplace_back(std::move(ready_.front()));
                    ready_.pop_front();
                }
                cv_.notify_all();
            }
            // where every piece goes: the caller's buffer, or (a first piece larger than the room left) the side buffer
            std::vector<uint8_t *> where(batch.size());
            size_t                 off = got;
            bool                   to_carry = false;
            for (size_t i = 0; i < batch.size(); ++i)
            {
                if (batch[i].n_out <= cap - off)
                {
                    where[i] = reinterpret_cast<uint8_t *>(dst) + off;
                    off += batch[i].n_out;
                }
                else
                {
                    carry_.resize(batch[i].n_out);
                    carry_off_ = 0;
                    where[i]   = carry_.data();
                    to_carry   = true;
                }
            }
            pool2_.parallel_for(batch.size(), [&](size_t i) {
                Piece               &pc = batch[i];
                std::vector<uint8_t> lut;
                resolve_markers(pc.sym.data() + kWindow, pc.n_out, pc.window.data(), where[i], lut);
                uint64_t from = 0;
                for (size_t m = 0; m <= pc.ends.size(); ++m)
                {
                    const uint64_t to = m < pc.ends.size() ? pc.ends[m].out_off : pc.n_out;
                    pc.crcs.emplace_back(to - from, (uint32_t)crc32_z(0, where[i] + from, (size_t)(to - from)));
                    from = to;
                }
            });
            // trailers of the members that end in these pieces (sequential: crc32_combine)
            for (auto &pc : batch)
                for (size_t m = 0; m < pc.crcs.size(); ++m)
                {
                    crc_run_ = (uint32_t)crc32_combine(crc_run_, pc.crcs[m].second, (z_off_t)pc.crcs[m].first);
                    len_run_ += pc.crcs[m].first;
                    if (m < pc.ends.size())
                    {
                        if (crc_run_ != pc.ends[m].crc || (uint32_t)len_run_ != pc.ends[m].isize)
                        {
                            std::lock_guard<std::mutex> l(mu_);
                            perr_ = "gzip CRC / length check failed";
                        }
                        crc_run_ = 0;
                        len_run_ = 0;
                    }
                }
            {
                std::lock_guard<std::mutex> l(mu_);
                for (auto &pc : batch)
                    if (sym_pool_.size() < 128)
                        sym_pool_.emplace_back(std::move(pc.sym));
                if (!perr_.empty())
                {
                    err_ = perr_;
                    return GNB_ERR_IO; // never hand out bytes of a member that failed its check
                }
            }
            got = off;
            if (to_carry)
            {
                const size_t n = std::min(cap - got, carry_.size());
                memcpy(dst + got, carry_.data(), n);
                carry_off_ = n;
                got += n;
            }
        }
        return (int64_t)got;
    }

  private:
    // one wave: chunk starts (finder), decoding, history propagation, marker replacement + CRCs; pieces are queued in order
    void produce()
    {
        std::string err;
        uint64_t    file_off  = 0;    // compressed bytes of the file before buf_
        uint64_t    start_bit = 0;    // where the next wave's first chunk starts, relative to buf_
        bool        at_header = true; // the very first chunk starts at the gzip header
        std::vector<uint8_t> buf;
        uint64_t    buf_valid = 0; // compressed bytes in buf (without padding)
        bool        file_done = false;
        const uint64_t wave_bytes = chunk_bytes_ * wave_chunks_;
        uint64_t       look_ahead = 0; // extra waves of compressed data kept buffered (grows if one deflate block needs it)
        for (;;)
        {
            // ---- keep two waves of compressed data buffered: the last chunk's decoder runs into the next wave ----
            {
                const uint64_t drop = std::min<uint64_t>(start_bit >> 3, buf_valid);
                if (drop)
                {
                    memmove(buf.data(), buf.data() + drop, buf_valid - drop);
                    buf_valid -= drop;
                    file_off += drop;
                    start_bit -= drop * 8;
                }
                const uint64_t want = std::min<uint64_t>((2 + look_ahead) * wave_bytes, size_ - file_off);
                buf.resize(want + 64);
                if (want > buf_valid)
                {
                    const uint64_t lo = buf_valid, n = want - buf_valid;
                    const size_t   slice = 4u << 20, parts = (size_t)((n + slice - 1) / slice);
                    std::atomic<bool> ok{true};
                    pool_.parallel_for(parts, [&](size_t i) {
                        const uint64_t o = lo + (uint64_t)i * slice, m = std::min<uint64_t>(slice, want - o);
                        if (!pread_all(fd_, (char *)buf.data() + o, m, file_off + o))
                            ok = false;
                    });
                    if (!ok)
                    {
                        err = "short read";
                        break;
                    }
                    buf_valid = want;
                }
                memset(buf.data() + buf_valid, 0, 64);
                file_done = file_off + buf_valid >= size_;
            }
            const uint64_t total_bits = buf_valid * 8;
            if (start_bit >= total_bits)
                break; // everything decoded (the previous wave ended at the end of the data)
            // ---- chunk starts: chunk 0 at start_bit, the others where the finder sees a block start ----
            // this wave: wave_bytes of compressed data from the start position; the rest of the buffer is look-ahead
            const uint64_t wave_end_byte = std::min<uint64_t>(buf_valid, (start_bit >> 3) + wave_bytes);
            const uint64_t wave_end_bit  = wave_end_byte * 8;
            const size_t   n_chunks     = (size_t)std::max<uint64_t>(1, (wave_end_bit - start_bit + chunk_bytes_ * 8 - 1) / (chunk_bytes_ * 8));
            std::vector<Chunk> chunks(n_chunks);
            {
                std::lock_guard<std::mutex> l(mu_);
                for (auto &c : chunks)
                    if (!sym_pool_.empty())
                    {
                        c.sym.swap(sym_pool_.back());
                        sym_pool_.pop_back();
                    }
            }
            chunks[0].start_bit = start_bit;
            chunks[0].found     = true;
            pool_.parallel_for(n_chunks - 1, [&](size_t k) {
                const size_t   i    = k + 1;
                const uint64_t from = start_bit + (uint64_t)i * chunk_bytes_ * 8, to = std::min<uint64_t>(from + chunk_bytes_ * 8, wave_end_bit);
                const uint64_t b    = find_block_start(buf.data(), total_bits, from, to);
                chunks[i].found     = b != ~0ull;
                chunks[i].start_bit = b;
            });
            std::vector<uint64_t> cands;
            for (size_t i = 1; i < n_chunks; ++i)
                if (chunks[i].found)
                    cands.push_back(chunks[i].start_bit);
            // ---- decode ----
            const bool last_data = file_done;
            pool_.parallel_for(n_chunks, [&](size_t i) {
                if (!chunks[i].found)
                    return;
                BitIn in;
                in.base       = buf.data();
                in.total_bits = total_bits;
                in.pos        = chunks[i].start_bit;
                const size_t first = std::upper_bound(cands.begin(), cands.end(), chunks[i].start_bit) - cands.begin();
                decode_chunk(in, at_header && i == 0, cands.data() + first, cands.size() - first, wave_end_bit, last_data, chunks[i]);
            });
            // ---- the chain of chunks that really follow one another ----
            std::vector<size_t> chain;
            bool                need_more = false;
            {
                size_t i = 0;
                for (;;)
                {
                    if (!chunks[i].ok)
                    {
                        if (chunks[i].ran_out && !file_done)
                            need_more = true; // a block reaches beyond the buffered data: buffer more and redo the wave
                        else
                            err = chunks[i].err.empty() ? "gzip decoding failed" : chunks[i].err;
                        break;
                    }
                    chain.push_back(i);
                    if (chunks[i].eos || chunks[i].end_bit >= wave_end_bit)
                        break;
                    size_t j = i + 1;
                    while (j < n_chunks && !(chunks[j].found && chunks[j].start_bit == chunks[i].end_bit))
                        ++j;
                    if (j == n_chunks)
                    {
                        err = "gzip decoding lost the block chain";
                        break;
                    }
                    i = j;
                }
            }
            if (need_more)
            {
                look_ahead = look_ahead ? look_ahead * 2 : 2;
                std::lock_guard<std::mutex> l(mu_);
                for (auto &c : chunks)
                    if (c.sym.capacity() && sym_pool_.size() < 128)
                        sym_pool_.emplace_back(std::move(c.sym));
                continue;
            }
            if (!err.empty())
                break;
            at_header = false;
            // ---- histories: window before chunk k = last 32 KiB of everything before it (sequential, 32 KiB each) ----
            std::vector<std::vector<uint8_t>> win(chain.size());
            for (size_t k = 0; k < chain.size(); ++k)
            {
                win[k] = window_;
                const Chunk    &c   = chunks[chain[k]];
                const uint16_t *s   = c.sym.data() + kWindow;
                const size_t    n   = c.n_out;
                std::vector<uint8_t> nw(kWindow);
                if (n >= kWindow)
                    for (uint32_t i = 0; i < kWindow; ++i)
                    {
                        const uint16_t v = s[n - kWindow + i];
                        nw[i]            = v < kMarker ? (uint8_t)v : window_[v - kMarker];
                    }
                else
                {
                    memcpy(nw.data(), window_.data() + n, kWindow - n);
                    for (size_t i = 0; i < n; ++i)
                    {
                        const uint16_t v      = s[i];
                        nw[kWindow - n + i] = v < kMarker ? (uint8_t)v : window_[v - kMarker];
                    }
                }
                window_.swap(nw);
            }
            // ---- hand the chunks over in order; the consumer replaces the markers (read()) ----
            bool eos = false;
            for (size_t k = 0; k < chain.size(); ++k)
            {
                Chunk &c = chunks[chain[k]];
                eos |= c.eos;
                if (c.n_out == 0 && c.ends.empty())
                    continue;
                Piece pc;
                pc.sym.swap(c.sym);
                pc.n_out = c.n_out;
                pc.window.swap(win[k]);
                pc.ends.swap(c.ends);
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return cancel_ || queued_ < max_queued_; });
                if (cancel_)
                    return;
                queued_ += pc.n_out;
                ready_.emplace_back(std::move(pc));
                cv_.notify_all();
            }
            start_bit = chunks[chain.back()].end_bit;
            {
                std::lock_guard<std::mutex> l(mu_);
                for (auto &c : chunks)
                    if (c.sym.capacity() && sym_pool_.size() < 128)
                        sym_pool_.emplace_back(std::move(c.sym));
            }
            if (eos)
                break;
            if (file_done && start_bit >= total_bits)
            {
                err = "unexpected end of the gzip stream";
                break;
            }
        }
        std::lock_guard<std::mutex> l(mu_);
        perr_     = err;
        finished_ = true;
        cv_.notify_all();
    }
    int      fd_;
    uint64_t size_;
    Pool     pool_, pool2_; // producer (find + decode) / consumer (marker replacement + CRC)
    uint64_t chunk_bytes_;
    size_t   wave_chunks_;
    std::vector<uint8_t> window_;
    std::thread          producer_;
    std::mutex           mu_;
    std::condition_variable cv_;
    std::deque<Piece>    ready_;
    size_t               queued_ = 0, max_queued_ = 512u << 20; // output bytes waiting in ready_
    std::vector<std::vector<uint16_t>> sym_pool_;                // symbol buffers for reuse (guarded by mu_)
    std::vector<uint8_t> carry_;                                 // a piece that straddled the end of the caller's buffer
    size_t               carry_off_ = 0;
    uint32_t             crc_run_ = 0; // CRC-32 / length of the current member so far (consumer)
    uint64_t             len_run_ = 0;
    bool                 finished_ = false, cancel_ = false;
    std::string          perr_;
};

} // namespace

std::unique_ptr<ByteSource> open_byte_source(const std::string &path, int threads, std::string &err)
{
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0)
    {
        err = "file not found/unreadable: " + path;
        return nullptr;
    }
    struct stat st;
    if (fstat(fd, &st) != 0)
    {
        close(fd);
        err = "cannot stat " + path;
        return nullptr;
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    unsigned char magic[2] = {0, 0};
    const ssize_t got      = pread(fd, magic, 2, 0);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)
        return std::unique_ptr<ByteSource>(new GzSource(fd, (uint64_t)st.st_size, threads > 0 ? threads : (int)std::min(16u, hw)));
    // page-cache copies: a few threads saturate them, more only take cores from the rest of the pipeline (measured: 4-8)
    return std::unique_ptr<ByteSource>(new PlainSource(fd, (uint64_t)st.st_size, threads > 0 ? threads : (int)std::min(6u, std::max(2u, hw / 2))));
}

} // namespace gnb

// ---------------------------------------------------------------------------------------------------------------------
// C ABI: the reader as a stream of bytes (no device involved)
// ---------------------------------------------------------------------------------------------------------------------
namespace gnb
{
int fail(int code, const std::string &msg);
}
struct gnb_reads_file
{
    std::unique_ptr<gnb::ByteSource> src;
};

extern "C" int gnb_reads_file_open(const char *path, int io_threads, gnb_reads_file **out)
{
    if (!path || !out)
        return gnb::fail(GNB_ERR_ARG, "gnb_reads_file_open: bad arguments");
    std::string err;
    auto        src = gnb::open_byte_source(path, io_threads, err);
    if (!src)
        return gnb::fail(GNB_ERR_IO, err);
    *out        = new gnb_reads_file();
    (*out)->src = std::move(src);
    return GNB_OK;
}

extern "C" int64_t gnb_reads_file_read(gnb_reads_file *f, void *dst, uint64_t cap)
{
    if (!f || (!dst && cap))
        return gnb::fail(GNB_ERR_ARG, "gnb_reads_file_read: bad arguments");
    const int64_t n = f->src->read(static_cast<char *>(dst), (size_t)cap);
    if (n < 0)
        return gnb::fail((int)n, f->src->error());
    return n;
}

extern "C" int gnb_reads_file_is_gzip(const gnb_reads_file *f) { return f && f->src->is_gzip() ? 1 : 0; }

extern "C" void gnb_reads_file_close(gnb_reads_file *f) { delete f; }
