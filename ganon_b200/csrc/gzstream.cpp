// See gzstream.h.  Deflate (RFC 1951) and gzip (RFC 1952) decoding written for chunk-parallel inflation; zlib is used for
// crc32 / crc32_combine only.
#include "gzstream.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
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
//                   2 length, 3 invalid symbol), bits 8-23 literal byte or length base, bits 24-27 extra bits;
//                   bit 6: TWO literals (bytes in bits 8-15 and 16-23, bits 0-3 = both code lengths together)
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
            // two literals whose codes both lie inside the 11 index bits: one entry, one step (bit 6; second byte in 16-23)
            const int l1 = e & 15;
            if (l1 != 0 && l1 < 11 && (e >> 4) < 256)
            {
                const uint16_t e2 = lit.table[i >> l1];
                const int      l2 = e2 & 15;
                if (l2 != 0 && l1 + l2 <= 11 && (e2 >> 4) < 256)
                    ltab[i] = (uint32_t)(l1 + l2) | (1u << 6) | ((uint32_t)(e >> 4) << 8) | ((uint32_t)(e2 >> 4) << 16);
            }
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

// The symbols of one Huffman block from the bit buffer (bb, cnt, p) into out[o..]: 0 at the end-of-block symbol, 1 when o
// came within a match of `stop` (the caller makes room and calls again: the state carries over), < 0 = -(index + 1) into
// kInflateErrors.  Everything the loop touches lives in locals of this function (the caller's variables are captured by its
// lambdas, which would keep them in memory and put a store-to-load round trip on every symbol).
const char *const kInflateErrors[] = {"unexpected end of the gzip stream", "invalid literal/length code", "invalid literal/length symbol", "invalid distance code",
                                      "invalid distance symbol"};

__attribute__((noinline)) int inflate_symbols(const DynHeader &H, const uint8_t *&p_io, uint64_t &bb_io, int &cnt_io, const uint8_t *const p_end,
                                              uint16_t *const out, size_t &o_io, size_t stop)
{
    const Huff     &L = H.lit, &D = H.dist;
    const uint32_t *const ltab = H.ltab, *const dtab = H.dtab;
    const uint8_t *p   = p_io;
    uint64_t       bb  = bb_io;
    int            cnt = cnt_io;
    uint16_t      *wp  = out + o_io;
    uint16_t *const wstop = out + stop - 300; // a match is at most 258 symbols (+ 7 of a copy step)
    int            rc  = 0;
    // one or two literals of a table entry: both slots are written, the second one counts only for a double entry
#define GNB_PUT_LITERALS(e)                                                         \
    do                                                                              \
    {                                                                               \
        const uint32_t two_ = (((e) >> 8) & 0xffu) | ((((e) >> 16) & 0xffu) << 16); \
        memcpy(wp, &two_, 4);                                                       \
        wp += 1 + (((e) >> 6) & 1u);                                                \
    } while (0)
#define GNB_REFILL()                       \
    do                                     \
    {                                      \
        bb |= load64(p) << cnt;            \
        p += (63 - cnt) >> 3;              \
        cnt |= 56;                         \
    } while (0)
#define GNB_LEAVE(code) \
    do                  \
    {                   \
        rc = (code);    \
        goto leave;     \
    } while (0)
    for (;;)
    {
        if (wp >= wstop)
            GNB_LEAVE(1);
        if (p > p_end)
            GNB_LEAVE(-1);
        GNB_REFILL();
        uint32_t e = ltab[bb & 2047];
        // up to three table entries of literals per refill (3 x 11 bits <= 56)
        if ((e & 0x3f) != 0 && (e & 0x30) == 0)
        {
            bb >>= e & 15;
            cnt -= (int)(e & 15);
            GNB_PUT_LITERALS(e);
            e = ltab[bb & 2047];
            if ((e & 0x3f) != 0 && (e & 0x30) == 0)
            {
                bb >>= e & 15;
                cnt -= (int)(e & 15);
                GNB_PUT_LITERALS(e);
                e = ltab[bb & 2047];
                if ((e & 0x3f) != 0 && (e & 0x30) == 0)
                {
                    bb >>= e & 15;
                    cnt -= (int)(e & 15);
                    GNB_PUT_LITERALS(e);
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
                GNB_LEAVE(-2);
            e = DynHeader::pack_lit(sym, used);
        }
        bb >>= used;
        cnt -= used;
        const uint32_t kind = (e >> 4) & 3;
        if (kind == 0)
        {
            *wp++ = (uint16_t)((e >> 8) & 0xffu);
            continue;
        }
        if (kind == 1)
            break;
        if (kind == 3)
            GNB_LEAVE(-3);
        const int      xb  = (int)(e >> 24) & 15;
        const uint32_t len = ((e >> 8) & 0xffff) + (uint32_t)(bb & ((1u << xb) - 1));
        bb >>= xb;
        cnt -= xb;
        uint32_t d = dtab[bb & 511];
        used       = (int)(d & 15);
        if (used == 0)
        {
            const int dsym = H.dist_empty ? -1 : D.slow(bb, used);
            if (dsym < 0)
                GNB_LEAVE(-4);
            d = DynHeader::pack_dist(dsym, used);
        }
        if (d >> 31)
            GNB_LEAVE(-5);
        bb >>= used;
        cnt -= used;
        const int      dx   = (int)(d >> 4) & 15;
        const uint32_t dist = ((d >> 8) & 0xffff) + (uint32_t)(bb & ((1u << dx) - 1));
        bb >>= dx;
        cnt -= dx;
        // dist <= 32768 <= symbols before wp: the marker prefix makes every legal distance addressable
        const uint16_t *sp = wp - dist;
        uint16_t       *dp = wp;
        wp += len;
        // 8 symbols (16 bytes) per step; a step may write up to 7 symbols past the match (the next symbols overwrite them).
        // Source and destination of one step do not overlap from distance 8 on.
        if (dist >= 8)
        {
            do
            {
                memcpy(dp, sp, 16);
                dp += 8;
                sp += 8;
            } while (dp < wp);
        }
        else if (dist == 1)
        {
            uint64_t v = sp[0];
            v |= v << 16;
            v |= v << 32;
            do
            {
                memcpy(dp, &v, 8);
                memcpy(dp + 4, &v, 8);
                dp += 8;
            } while (dp < wp);
        }
        else
            for (uint32_t i = 0; i < len; ++i)
                dp[i] = sp[i];
    }
leave:
#undef GNB_LEAVE
#undef GNB_REFILL
#undef GNB_PUT_LITERALS
    p_io   = p;
    bb_io  = bb;
    cnt_io = cnt;
    o_io   = (size_t)(wp - out);
    return rc;
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
            // bit buffer: `cnt` valid bits in `bb`, next unread byte at p; stream position = (p - base) * 8 - cnt
            const uint8_t *p   = in.base + (in.pos >> 3);
            uint64_t       bb  = load64(p) >> (in.pos & 7);
            int            cnt = 64 - (int)(in.pos & 7);
            p += 8;
            const uint8_t *const p_end = in.base + (in.total_bits >> 3) + 16; // the buffer's padding keeps loads legal
            // (the first load above took 8 bytes; cnt may be 57..64: bring it into the refill's invariant cnt <= 63)
            if (cnt == 64)
            {
                cnt = 56;
                p -= 1;
                bb &= (1ull << 56) - 1;
            }
            for (;;)
            {
                if (o + 1024 + (1u << 16) > cap)
                    room(1024 + (1u << 17));
                const int rc = inflate_symbols(*H, p, bb, cnt, p_end, out, o, cap - 1024);
                if (rc == 0)
                    break;
                if (rc < 0)
                    return fail(kInflateErrors[-rc - 1]);
            }
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
    // Kraft sums (in 128ths) of four 3-bit code lengths at once
    static const std::vector<uint16_t> kraft4_table = [] {
        std::vector<uint16_t> t(4096);
        for (int i = 0; i < 4096; ++i)
        {
            int sum = 0;
            for (int j = 0; j < 4; ++j)
            {
                const int l = (i >> (3 * j)) & 7;
                if (l)
                    sum += 128 >> l;
            }
            t[i] = (uint16_t)sum;
        }
        return t;
    }();
    const uint16_t *const kraft4 = kraft4_table.data();
    static const std::vector<uint8_t> head_table = [] {
        std::vector<uint8_t> t(8192);
        for (uint32_t i = 0; i < 8192; ++i)
            t[i] = ((i >> 1) & 3) == 2 && ((i >> 3) & 31) <= 29 && ((i >> 8) & 31) <= 29;
        return t;
    }();
    const uint8_t *const head_ok = head_table.data();
    DynHeader h;
    for (uint64_t b = from_bit; b < to_bit && b + 64 < total_bits; ++b)
    {
        const uint64_t w = load64(base + (b >> 3)) >> (b & 7);
        // Without branches up to the one that almost never passes (a data-dependent branch per bit position would be
        // mispredicted a fifth of the time): BTYPE = 10 (dynamic), HLIT <= 29, HDIST <= 29 from a table over the first 13
        // bits; the code-length code complete (Kraft sum exactly 1 = 128/128), four 3-bit lengths per table lookup.
        // 57 bits are available in w: 17 header bits + 13 code lengths (39 bits); the rest comes from a second load.
        const int      hclen = (int)((w >> 13) & 15) + 4;
        const int      n1    = hclen < 13 ? hclen : 13;
        const uint64_t v     = (w >> 17) & ((1ull << (3 * n1)) - 1);
        uint32_t       sum   = (uint32_t)kraft4[v & 4095] + kraft4[(v >> 12) & 4095] + kraft4[(v >> 24) & 4095] + kraft4[v >> 36];
        const uint32_t maybe = head_ok[w & 8191] & (uint32_t)((sum == 128) | ((hclen > 13) & (sum < 128)));
        if (!maybe)
            continue;
        if (hclen > 13)
        {
            const uint64_t b2 = b + 17 + 39;
            const uint64_t v2 = (load64(base + (b2 >> 3)) >> (b2 & 7)) & ((1ull << (3 * (hclen - 13))) - 1);
            sum += kraft4[v2 & 4095] + kraft4[v2 >> 12];
        }
        if (sum != 128)
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

// CRC-32 (gzip polynomial) by folding with carry-less multiplication (Gopal et al., "Fast CRC computation for generic
// polynomials using PCLMULQDQ"): four 128-bit lanes folded by 512 bits per step, then reduced to 32 bits (Barrett).  Used when
// the CPU has PCLMULQDQ and a start-up comparison with zlib's crc32 over a test pattern agrees; zlib's otherwise.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(uint32_t crc, const uint8_t *buf, size_t len) // len >= 64, multiple of 16
{
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
    const __m128i k5k0 = _mm_set_epi64x(0x0000000000, 0x0163cd6124);
    const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = k1k2;
    buf += 64;
    len -= 64;
    while (len >= 64)
    {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
        y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
        y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    // four lanes into one
    x0 = k3k4;
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16)
    {
        x2 = _mm_loadu_si128((const __m128i *)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    // 128 -> 64 bits
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = k5k0;
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    // Barrett reduction to 32 bits
    x0 = poly;
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

// crc32_z(0, p, n) for a run of output bytes
inline uint32_t crc32_of(const uint8_t *p, size_t n)
{
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool use_clmul = [] {
        if (getenv("GANON_B200_GZ_ZLIB_CRC") || !__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1"))
            return false;
        uint8_t t[1024 + 48];
        for (size_t i = 0; i < sizeof t; ++i)
            t[i] = (uint8_t)(i * 131 + (i >> 3) * 7 + 5);
        for (size_t n : {(size_t)64, (size_t)80, (size_t)1024, (size_t)1072})
            if ((uint32_t)crc32_z(0, t, n) != ~crc32_clmul(~0u, t, n))
                return false;
        if (getenv("GANON_B200_GZ_TRACE"))
            fprintf(stderr, "[gz] CRC-32 by carry-less multiplication\n");
        return true;
    }();
    if (use_clmul && n >= 64)
    {
        const size_t body = n & ~(size_t)15;
        const uint32_t c  = ~crc32_clmul(~0u, p, body);
        return body == n ? c : (uint32_t)crc32_z(c, p + body, n - body);
    }
#endif
    return (uint32_t)crc32_z(0, p, n);
}

// symbols -> bytes: literals as they are, markers through the (now known) 32 KiB history.  Groups of 16 symbols without a
// marker are narrowed with SSE2; the others go through a 64 Ki-entry table (symbol -> byte), one load per symbol.
inline void build_marker_table(const uint8_t *w, std::vector<uint8_t> &lut)
{
    lut.resize(65536);
    for (int i = 0; i < 256; ++i)
        lut[i] = (uint8_t)i;
    memcpy(lut.data() + kMarker, w, kWindow);
}

inline void resolve_markers(const uint16_t *s, size_t n, const uint8_t *t, uint8_t *d)
{
    size_t i = 0;
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
// The compressed file is cut into ranges of chunk_bytes_.  Workers take tasks in stream order from one scheduler:
//   finder r   the first block start inside range r (range 0: the gzip header), a little ahead of the decoders;
//   decoder d  from the start found in range d up to the block boundary where a later range's decoder starts;
//   slices     (ahead of everything else) marker replacement + CRC of decoded pieces for the consumer, read().
// A sequencer thread follows the chain of chunks that really follow one another (position by position: the chunk found in
// the range of the current position must start exactly there), propagates the 32 KiB history, and queues the pieces.  Where
// the chain breaks -- the finder missed a boundary (stored / fixed-Huffman blocks, a block longer than its view), a candidate
// was a false positive, a decoder ran out of its view -- the sequencer decodes from the known position itself.
// Compressed bytes live in an anonymous mapping as large as the file, filled range by range with pread (I/O errors stay
// error codes) and given back to the system behind the sequencer.
class GzSource : public ByteSource
{
  public:
    GzSource(int fd, uint64_t size, int threads) : fd_(fd), size_(size)
    {
        chunk_bytes_ = 1u << 20;
        if (const char *e = getenv("GANON_B200_GZ_CHUNK"))
            chunk_bytes_ = std::max<uint64_t>(1u << 12, strtoull(e, nullptr, 10));
        n_ranges_ = (size_t)((size_ + chunk_bytes_ - 1) / chunk_bytes_);
        const int T  = std::max(1, threads);
        in_flight_  = (size_t)T * 2 + kAhead;

        img_bytes_  = (size_t)size_ + (64u << 10);
        void *m     = mmap(nullptr, img_bytes_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED)
        {
            perr_     = "cannot map the compressed file's image";
            finished_ = true;
            return;
        }
        img_    = static_cast<uint8_t *>(m);
        loaded_ = std::vector<std::atomic<uint8_t>>(n_ranges_);
        slots_.resize(n_ranges_);
        window_.assign(kWindow, 0);
        for (int i = 0; i < T; ++i)
            workers_.emplace_back([this] { work(); });
        sequencer_ = std::thread([this] { sequence(); });
    }
    ~GzSource() override
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            cancel_ = true;
        }
        {
            std::lock_guard<std::mutex> l(smu_);
            stop_ = true;
        }
        cv_.notify_all();
        scv_.notify_all();
        if (sequencer_.joinable())
            sequencer_.join();
        for (auto &t : workers_)
            t.join();
        if (img_)
            munmap(img_, img_bytes_);
        close(fd_);
    }
    bool     is_gzip() const override { return true; }
    uint64_t size() const override { return size_; }
    // Markers are replaced HERE, for the consumer, straight into the caller's buffer (slices of the queued pieces on the
    // workers): the decoded bytes are written once, where they are needed.  Only a piece that straddles the end of the
    // caller's buffer goes through a side buffer.
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
                while (!ready_.empty() && batch.size() < 256 && (batch.empty() || ready_.front().n_out <= room))
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
            // slices of at most 1 MiB that do not cross a member's end: (piece, from, to), CRC of each
            struct Slice
            {
                size_t   piece;
                uint64_t from, to;
                uint32_t crc;
                bool     member_ends; // a gzip member ends where the slice does
                size_t   end_index;
            };
            std::vector<Slice> slices;
            for (size_t i = 0; i < batch.size(); ++i)
            {
                uint64_t from = 0;
                for (size_t m = 0; m <= batch[i].ends.size(); ++m)
                {
                    const uint64_t to = m < batch[i].ends.size() ? batch[i].ends[m].out_off : batch[i].n_out;
                    uint64_t       a  = from;
                    do
                    {
                        const uint64_t z = std::min<uint64_t>(to, a + (1u << 20));
                        slices.push_back(Slice{i, a, z, 0, m < batch[i].ends.size() && z == to, m});
                        a = z;
                    } while (a < to);
                    from = to;
                }
            }
            std::vector<std::vector<uint8_t>> luts(batch.size());
            run_hi(batch.size(), [&](size_t i) { build_marker_table(batch[i].window.data(), luts[i]); });
            run_hi(slices.size(), [&](size_t k) {
                Slice &sl = slices[k];
                Piece &pc = batch[sl.piece];
                resolve_markers(pc.sym.data() + kWindow + sl.from, (size_t)(sl.to - sl.from), luts[sl.piece].data(), where[sl.piece] + sl.from);
                sl.crc = crc32_of(where[sl.piece] + sl.from, (size_t)(sl.to - sl.from));
            });
            // trailers of the members that end in these pieces (sequential: crc32_combine)
            for (auto const &sl : slices)
            {
                crc_run_ = (uint32_t)crc32_combine(crc_run_, sl.crc, (z_off_t)(sl.to - sl.from));
                len_run_ += sl.to - sl.from;
                if (sl.member_ends)
                {
                    const MemberEnd &me = batch[sl.piece].ends[sl.end_index];
                    if (crc_run_ != me.crc || (uint32_t)len_run_ != me.isize)
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
                    recycle(pc.sym);
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
    static constexpr size_t kAhead = 3; // a decoder knows the starts found in this many ranges after its own

    struct Slot
    {
        int      find_state = 0, dec_state = 0; // 0: not started, 1: running, 2: done (smu_)
        bool     found = false, discard = false;
        uint64_t start_bit = 0;
        Chunk    chunk;
    };

    void recycle(std::vector<uint16_t> &sym) // mu_ held
    {
        if (sym.capacity() && sym_pool_.size() < 256)
            sym_pool_.emplace_back(std::move(sym));
        sym = std::vector<uint16_t>();
    }
    void take_buffer(Chunk &c)
    {
        std::lock_guard<std::mutex> l(mu_);
        if (!sym_pool_.empty())
        {
            c.sym.swap(sym_pool_.back());
            sym_pool_.pop_back();
        }
    }
    uint64_t file_bits() const { return size_ * 8; }
    uint64_t range_bits() const { return chunk_bytes_ * 8; }

    // compressed bytes of range r in the image
    bool ensure_loaded(size_t r)
    {
        if (r >= n_ranges_)
            return true;
        uint8_t st = loaded_[r].load(std::memory_order_acquire);
        if (st == 2)
            return true;
        uint8_t expect = 0;
        if (st == 0 && loaded_[r].compare_exchange_strong(expect, 1))
        {
            const uint64_t off = (uint64_t)r * chunk_bytes_, n = std::min<uint64_t>(chunk_bytes_, size_ - off);
            const bool     ok  = pread_all(fd_, reinterpret_cast<char *>(img_) + off, (size_t)n, off);
            loaded_[r].store(ok ? 2 : 3, std::memory_order_release);
            return ok;
        }
        while ((st = loaded_[r].load(std::memory_order_acquire)) == 1)
            std::this_thread::yield();
        return st == 2;
    }
    // the view [.., end of range hi]: its length in bits, clipped to the file
    uint64_t view_bits(size_t hi) const { return std::min<uint64_t>(size_, ((uint64_t)hi + 1) * chunk_bytes_) * 8; }

    void find(size_t r)
    {
        bool     found = false;
        uint64_t start = 0;
        bool     io_ok = ensure_loaded(r) && ensure_loaded(r + 1);
        if (io_ok)
        {
            if (r == 0)
                found = true; // the gzip header
            else
            {
                const uint64_t b = find_block_start(img_, view_bits(r + 1), (uint64_t)r * range_bits(), std::min(file_bits(), ((uint64_t)r + 1) * range_bits()));
                found = b != ~0ull;
                start = b;
            }
        }
        std::lock_guard<std::mutex> l(smu_);
        Slot &s      = slots_[r];
        s.found      = found;
        s.start_bit  = start;
        s.find_state = 2;
        if (!io_ok)
            io_failed_ = true;
        scv_.notify_all();
    }

    void decode(size_t d, std::vector<uint64_t> cands)
    {
        Slot        &s  = slots_[d];
        const size_t hi = std::min(n_ranges_ - 1, d + kAhead + 1);
        bool         io_ok = true;
        for (size_t r = d; r <= hi && io_ok; ++r)
            io_ok = ensure_loaded(r);
        if (io_ok)
        {
            take_buffer(s.chunk);
            BitIn in;
            in.base       = img_;
            in.total_bits = view_bits(hi);
            in.pos        = s.start_bit;
            const uint64_t limit = d + kAhead + 1 < n_ranges_ ? ((uint64_t)d + kAhead + 1) * range_bits() : ~0ull;
            decode_chunk(in, d == 0, cands.data(), cands.size(), limit, in.total_bits == file_bits(), s.chunk);
        }
        else
        {
            s.chunk.ok  = false;
            s.chunk.err = "short read";
        }
        std::unique_lock<std::mutex> l(smu_);
        s.dec_state = 2;
        if (!io_ok)
            io_failed_ = true;
        const bool drop = s.discard;
        scv_.notify_all();
        l.unlock();
        if (drop)
        {
            std::lock_guard<std::mutex> lm(mu_);
            recycle(s.chunk.sym);
        }
    }

    // the next task in stream order (smu_ held); false: nothing to do right now
    bool pick(std::function<void()> &job)
    {
        for (;;)
        {
            if (next_dec_ < n_ranges_)
            {
                const size_t need = std::min(n_ranges_ - 1, next_dec_ + kAhead);
                if (next_find_ <= need)
                { // the finders a decoder waits for come first
                    const size_t r         = next_find_++;
                    slots_[r].find_state = 1;
                    job                  = [this, r] { find(r); };
                    return true;
                }
                if (next_dec_ < seq_range_ + in_flight_)
                {
                    bool ready = true;
                    for (size_t j = next_dec_; j <= need && ready; ++j)
                        ready = slots_[j].find_state == 2;
                    if (ready)
                    {
                        const size_t d = next_dec_++;
                        Slot        &s = slots_[d];
                        if (!s.found || s.discard || d < seq_range_)
                        { // no start in this range, or the sequencer is past it already
                            s.dec_state = 2;
                            scv_.notify_all();
                            continue;
                        }
                        std::vector<uint64_t> cands;
                        for (size_t j = d + 1; j <= need; ++j)
                            if (slots_[j].found)
                                cands.push_back(slots_[j].start_bit);
                        s.dec_state = 1;
                        job         = [this, d, cands] { decode(d, cands); };
                        return true;
                    }
                }
            }
            if (next_find_ < n_ranges_ && next_find_ < next_dec_ + kAhead + 1 + workers_.size())
            {
                const size_t r         = next_find_++;
                slots_[r].find_state = 1;
                job                  = [this, r] { find(r); };
                return true;
            }
            return false;
        }
    }

    void work()
    {
        for (;;)
        {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> l(smu_);
                for (;;)
                {
                    if (stop_)
                        return;
                    if (!hi_.empty())
                    {
                        job = std::move(hi_.front());
                        hi_.pop_front();
                        break;
                    }
                    if (!seq_done_ && pick(job))
                        break;
                    scv_.wait(l);
                }
            }
            job();
        }
    }

    // fn(0..n-1) ahead of the decoding tasks, on the workers and the calling thread
    void run_hi(size_t n, const std::function<void(size_t)> &fn)
    {
        if (n == 0)
            return;
        if (n == 1)
            return fn(0);
        std::mutex              m;
        std::condition_variable c;
        size_t                  left = n;
        {
            std::lock_guard<std::mutex> l(smu_);
            for (size_t i = 0; i < n; ++i)
                hi_.emplace_back([&, i] {
                    fn(i);
                    std::lock_guard<std::mutex> lm(m);
                    if (--left == 0)
                        c.notify_all();
                });
        }
        scv_.notify_all();
        for (;;)
        {
            std::function<void()> job;
            {
                std::lock_guard<std::mutex> l(smu_);
                if (hi_.empty())
                    break;
                job = std::move(hi_.front());
                hi_.pop_front();
            }
            job();
        }
        std::unique_lock<std::mutex> lm(m);
        c.wait(lm, [&] { return left == 0; });
    }

    // the chunk that starts at `pos`, decoded by the sequencer itself with a view that grows until the chunk fits
    bool decode_here(uint64_t pos, size_t r, Chunk &c, std::string &err)
    {
        for (size_t ahead = kAhead;; ahead *= 2)
        {
            const size_t hi = std::min(n_ranges_ - 1, r + ahead + 1);
            for (size_t j = r; j <= hi; ++j)
                if (!ensure_loaded(j))
                {
                    err = "short read";
                    return false;
                }
            std::vector<uint64_t> cands;
            {
                std::lock_guard<std::mutex> l(smu_);
                for (size_t j = r; j <= std::min(n_ranges_ - 1, r + ahead); ++j)
                    if (slots_[j].find_state == 2 && slots_[j].found && !slots_[j].discard && slots_[j].start_bit > pos)
                        cands.push_back(slots_[j].start_bit);
            }
            if (c.sym.empty())
                take_buffer(c);
            c.ends.clear();
            BitIn in;
            in.base       = img_;
            in.total_bits = view_bits(hi);
            in.pos        = pos;
            const uint64_t limit = r + ahead + 1 < n_ranges_ ? ((uint64_t)r + ahead + 1) * range_bits() : ~0ull;
            decode_chunk(in, pos == 0, cands.data(), cands.size(), limit, in.total_bits == file_bits(), c);
            if (c.ok)
                return true;
            if (!(c.ran_out && in.total_bits < file_bits()))
            {
                err = c.err.empty() ? "gzip decoding failed" : c.err;
                return false;
            }
        }
    }

    void sequence()
    {
        std::string err;
        uint64_t    pos        = 0; // where the next chunk starts: certain
        size_t      freed_upto = 0; // ranges [0, freed_upto) of the image were given back
        size_t      prev_r     = 0;
        bool        eos        = false;
        while (!eos)
        {
            if (pos >= file_bits())
            {
                err = "unexpected end of the gzip stream";
                break;
            }
            const size_t r = (size_t)(pos / range_bits());
            Slot        &s = slots_[r];
            bool         spec;
            {
                std::unique_lock<std::mutex> l(smu_);
                seq_range_ = r;
                // chunks the chain skipped: their buffers go back once their decoders are done
                for (size_t j = prev_r; j < r; ++j)
                    drop_slot(j);
                scv_.notify_all();
                scv_.wait(l, [&] { return stop_ || io_failed_ || s.find_state == 2; });
                if (stop_)
                    return;
                spec = s.found && s.start_bit == pos && !io_failed_;
                if (spec)
                    scv_.wait(l, [&] { return stop_ || s.dec_state == 2; });
                if (stop_)
                    return;
                if (!spec && !(s.found && s.start_bit > pos))
                    drop_slot(r); // (a start found further on in this range stays a candidate: the chain may come back to it)
            }
            if (io_failed_)
            {
                err = "short read";
                break;
            }
            Chunk  own;
            Chunk *c = &s.chunk;
            if (spec && !s.chunk.ok && !(s.chunk.ran_out && view_bits(std::min(n_ranges_ - 1, r + kAhead + 1)) < file_bits()))
            {
                err = s.chunk.err.empty() ? "gzip decoding failed" : s.chunk.err; // the start was certain: a real error
                break;
            }
            if (!spec || !s.chunk.ok)
            {
                if (spec)
                {
                    std::lock_guard<std::mutex> l(mu_);
                    own.sym.swap(s.chunk.sym); // ran out of its view: again, with more of the file
                }
                if (!decode_here(pos, r, own, err))
                    break;
                c = &own;
            }
            // ---- history: window before the next chunk = last 32 KiB of everything so far ----
            std::vector<uint8_t> win = window_;
            {
                const uint16_t *sy = c->sym.data() + kWindow;
                const size_t    n  = c->n_out;
                std::vector<uint8_t> nw(kWindow);
                if (n >= kWindow)
                    for (uint32_t i = 0; i < kWindow; ++i)
                    {
                        const uint16_t v = sy[n - kWindow + i];
                        nw[i]            = v < kMarker ? (uint8_t)v : window_[v - kMarker];
                    }
                else
                {
                    memcpy(nw.data(), window_.data() + n, kWindow - n);
                    for (size_t i = 0; i < n; ++i)
                    {
                        const uint16_t v      = sy[i];
                        nw[kWindow - n + i] = v < kMarker ? (uint8_t)v : window_[v - kMarker];
                    }
                }
                window_.swap(nw);
            }
            eos = c->eos;
            pos = c->end_bit;
            if (c->n_out != 0 || !c->ends.empty())
            {
                Piece pc;
                pc.sym.swap(c->sym);
                pc.n_out = c->n_out;
                pc.window.swap(win);
                pc.ends.swap(c->ends);
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return cancel_ || queued_ < max_queued_; });
                if (cancel_)
                    return;
                queued_ += pc.n_out;
                ready_.emplace_back(std::move(pc));
                cv_.notify_all();
            }
            else
            {
                std::lock_guard<std::mutex> l(mu_);
                recycle(c->sym);
            }
            prev_r = r + 1;
            // compressed bytes behind the position are not needed again
            const size_t keep_from = (size_t)(pos / range_bits());
            if (keep_from > freed_upto + 8)
            { // whole pages inside [freed_upto, keep_from) only: madvise rounds the length up
                const uint64_t page = 4096, a = ((uint64_t)freed_upto * chunk_bytes_ + page - 1) & ~(page - 1), b = ((uint64_t)keep_from * chunk_bytes_) & ~(page - 1);
                if (b > a)
                    madvise(img_ + a, (size_t)(b - a), MADV_DONTNEED);
                freed_upto = keep_from;
            }
        }
        {
            std::lock_guard<std::mutex> l(smu_);
            seq_done_ = true; // no more decoding tasks; the workers stay for the consumer's slices
            for (size_t j = prev_r; j < n_ranges_ && j < next_dec_; ++j)
                drop_slot(j);
        }
        std::lock_guard<std::mutex> l(mu_);
        perr_     = err;
        finished_ = true;
        cv_.notify_all();
    }
    void drop_slot(size_t j) // smu_ held
    {
        Slot &s = slots_[j];
        if (s.discard)
            return;
        s.discard = true;
        if (s.dec_state == 2)
        {
            std::lock_guard<std::mutex> l(mu_);
            recycle(s.chunk.sym);
        }
    }

    int      fd_;
    uint64_t size_;
    uint64_t chunk_bytes_;
    size_t   n_ranges_ = 0, in_flight_ = 0;
    uint8_t *img_ = nullptr;
    size_t   img_bytes_ = 0;
    std::vector<std::atomic<uint8_t>> loaded_; // 0: not read, 1: being read, 2: in the image, 3: read error
    std::vector<Slot>    slots_;
    std::vector<uint8_t> window_;
    std::vector<std::thread> workers_;
    std::thread              sequencer_;
    // scheduler (smu_): positions of the finder / decoder fronts, the range the sequencer is at, consumer jobs
    std::mutex              smu_;
    std::condition_variable scv_;
    size_t                  next_find_ = 0, next_dec_ = 0, seq_range_ = 0;
    std::deque<std::function<void()>> hi_;
    bool                    stop_ = false, seq_done_ = false, io_failed_ = false;
    // queue to the consumer (mu_)
    std::mutex              mu_;
    std::condition_variable cv_;
    std::deque<Piece>       ready_;
    size_t                  queued_ = 0, max_queued_ = 128u << 20; // output bytes waiting in ready_
    std::vector<std::vector<uint16_t>> sym_pool_;                   // symbol buffers for reuse (mu_)
    std::vector<uint8_t> carry_;                                    // a piece that straddled the end of the caller's buffer
    size_t               carry_off_ = 0;
    uint32_t             crc_run_ = 0; // CRC-32 / length of the current member so far (consumer)
    uint64_t             len_run_ = 0;
    bool                 finished_ = false, cancel_ = false;
    std::string          perr_;
};

} // namespace

std::unique_ptr<ByteSource> open_byte_source(const std::string &path, int threads, std::string &err, int share)
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
    unsigned char magic[3] = {0, 0, 0};
    const ssize_t got      = pread(fd, magic, 3, 0);
    if (got == 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h')
    {
        const unsigned automatic = std::max(1u, std::min(16u, hw > 2 ? hw - 2 : 1u) / (unsigned)std::max(1, share));
        return wrap_sequence_format(open_bz2_source(fd, (uint64_t)st.st_size, threads > 0 ? threads : (int)automatic), path);
    }
    if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b)
    {
        // all host threads but two (at most 16), split between the files read at the same time: a file run also has its
        // staging thread (copies, kernel launches, short waits), a writer and the driver's threads, and a worker per core
        // starves them (16-CPU host, c2 reads from gzip: 16 workers 0.68-1.5 s, 14 workers 0.55 s)
        const unsigned automatic = std::max(1u, std::min(16u, hw > 2 ? hw - 2 : 1u) / (unsigned)std::max(1, share));
        return wrap_sequence_format(std::unique_ptr<ByteSource>(new GzSource(fd, (uint64_t)st.st_size, threads > 0 ? threads : (int)automatic)), path);
    }
    // page-cache copies: a few threads saturate them, more only take cores from the rest of the pipeline (measured: 4-8)
    return wrap_sequence_format(std::unique_ptr<ByteSource>(new PlainSource(fd, (uint64_t)st.st_size, threads > 0 ? threads : (int)std::min(6u, std::max(2u, hw / 2)))), path);
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
