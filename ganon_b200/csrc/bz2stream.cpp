// bzip2-compressed read files (A0, parse_reads GC.cpp:1220-1287).  The reference cannot be built without libbz2
// (CMakeLists.txt:114 `find_package( BZip2 REQUIRED )`): seqan3's transparent decompression layer
// (seqan3/io/detail/misc_input.hpp:145-153, magic "BZh") hands a `.bz2` read file to libbz2 on one thread.  Here the blocks
// of the file are decoded by all host threads at once:
//   * a bzip2 block starts with the 48-bit number 0x314159265359 and a stream ends with 0x177245385090 + its CRC, at any
//     BIT position; blocks do not depend on each other.  A scanner looks for the two numbers (a table over byte pairs says
//     at which of the eight bit offsets the two bytes can belong to one of them, the 48 bits are then compared);
//   * every block becomes a bzip2 stream of its own -- "BZh9", the block's bits moved to a byte boundary, the end-of-stream
//     number and the block's CRC as the stream's CRC -- which libbz2 decodes and verifies (the CRC of every block); the
//     scanner checks the stream's own CRC (the blocks' CRCs folded together) and that every stream's first block stands right
//     behind its "BZh1".."BZh9" header.  libbz2 is reached through dlopen (the image has the library but not its header),
//     so the product links without it;
//   * a scanner thread runs ahead of the decoders (at most three blocks per worker are queued), the workers decode, read()
//     hands the bytes on in order.  A block that fails to decode is retried together with the segment(s) after it (its end
//     was a chance occurrence of the number inside compressed data: once per ~10^14 bits) before it is an error.
// Concatenated streams (pbzip2, `cat a.bz2 b.bz2`) are just more blocks.  A file that ends without an end-of-stream number, or
// inside a block, is an error (GNB_ERR_IO), as are damaged blocks: never wrong bytes.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "gzstream.h"

#include "../../include/ganon_b200.h"

namespace gnb
{
namespace
{
// ---- libbz2 through dlopen (bzlib.h: bz_stream, BZ2_bzDecompress*) ----
struct BzStream
{
    char        *next_in;
    unsigned int avail_in, total_in_lo32, total_in_hi32;
    char        *next_out;
    unsigned int avail_out, total_out_lo32, total_out_hi32;
    void        *state;
    void *(*bzalloc)(void *, int, int);
    void (*bzfree)(void *, void *);
    void *opaque;
};
struct BzLib
{
    int (*init)(BzStream *, int, int) = nullptr;
    int (*run)(BzStream *)            = nullptr;
    int (*end)(BzStream *)            = nullptr;
    bool ok() const { return init && run && end; }
};
const BzLib &bzlib()
{
    static const BzLib lib = [] {
        BzLib l;
        for (const char *name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"})
            if (void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL))
            {
                l.init = reinterpret_cast<int (*)(BzStream *, int, int)>(dlsym(h, "BZ2_bzDecompressInit"));
                l.run  = reinterpret_cast<int (*)(BzStream *)>(dlsym(h, "BZ2_bzDecompress"));
                l.end  = reinterpret_cast<int (*)(BzStream *)>(dlsym(h, "BZ2_bzDecompressEnd"));
                if (l.ok())
                    break;
            }
        return l;
    }();
    return lib;
}
constexpr int kBzOk = 0, kBzStreamEnd = 4;

constexpr uint64_t kBlockMagic = 0x314159265359ull, kEndMagic = 0x177245385090ull, kMask48 = (1ull << 48) - 1;

// 48 bits from bit position `bit` (most significant bit first, as bzip2 writes them); the buffer has 8 readable bytes past
// the last position asked for
inline uint64_t bits48(const uint8_t *b, uint64_t bit)
{
    const uint8_t *p = b + (bit >> 3);
    uint64_t       v = 0;
    for (int i = 0; i < 8; ++i)
        v = (v << 8) | p[i];
    return (v >> (16 - (bit & 7))) & kMask48;
}

// pair_mask[b0 << 8 | b1]: bit s set when, for a number that starts s bits into the byte BEFORE b0, the bytes b0 b1 are
// what its bits 8-s .. 23-s must be (both numbers)
struct PairTable
{
    std::vector<uint8_t> mask;
    PairTable() : mask(65536, 0)
    {
        for (uint64_t magic : {kBlockMagic, kEndMagic})
            for (int s = 0; s < 8; ++s)
            {
                // the number occupies bits [s, s + 48) of a 7-byte span; its bytes 1 and 2 are fully determined
                const unsigned __int128 span = (unsigned __int128)magic << (56 - 48 - s);
                const unsigned          b1 = (unsigned)(span >> 40) & 0xff, b2 = (unsigned)(span >> 32) & 0xff;
                mask[b1 << 8 | b2] |= (uint8_t)(1u << s);
            }
    }
};
const PairTable kPairs;

struct Segment // one block (or the tail of the data), as a copy of the bytes that hold its bits
{
    std::vector<uint8_t> bytes;
    unsigned             first_bit = 0; // position of the block number in bytes[0]
    uint64_t             n_bits = 0;    // up to the next number
    std::vector<uint8_t> out;
    bool                 ok = false;
};

// the block as a stream of its own: "BZh9", its bits, the end-of-stream number, its CRC (the 32 bits after its number)
bool decode_segment(const uint8_t *bytes, size_t n_bytes, unsigned first_bit, uint64_t n_bits, std::vector<uint8_t> &out)
{
    const BzLib &L = bzlib();
    if (n_bits < 80)
        return false;
    std::vector<uint8_t> s;
    s.reserve((size_t)(n_bits >> 3) + 32);
    s.insert(s.end(), {'B', 'Z', 'h', '9'});
    // move the bits to a byte boundary
    const size_t whole = (size_t)(n_bits >> 3);
    const unsigned rest = (unsigned)(n_bits & 7);
    auto byte_at = [&](size_t i) -> uint8_t { // 8 bits from bit first_bit + 8 i
        const unsigned a = bytes[i], b = i + 1 < n_bytes ? bytes[i + 1] : 0u;
        return (uint8_t)(((a << 8 | b) >> (8 - first_bit)) & 0xff);
    };
    for (size_t i = 0; i < whole; ++i)
        s.push_back(byte_at(i));
    const uint32_t crc = (uint32_t)s[4 + 6] << 24 | (uint32_t)s[4 + 7] << 16 | (uint32_t)s[4 + 8] << 8 | s[4 + 9];
    // the last `rest` bits, then 48 + 32 bits of trailer
    uint64_t acc   = rest ? (uint64_t)(byte_at(whole) >> (8 - rest)) : 0;
    int      n_acc = (int)rest;
    auto     put   = [&](uint64_t v, int n) {
        for (int i = n - 1; i >= 0; --i)
        {
            acc = (acc << 1) | ((v >> i) & 1);
            if (++n_acc == 8)
            {
                s.push_back((uint8_t)acc);
                acc   = 0;
                n_acc = 0;
            }
        }
    };
    put(kEndMagic, 48);
    put(crc, 32);
    if (n_acc)
        s.push_back((uint8_t)(acc << (8 - n_acc)));
    BzStream z{};
    if (L.init(&z, 0, 0) != kBzOk)
        return false;
    out.clear();
    out.resize(1u << 20);
    z.next_in  = reinterpret_cast<char *>(s.data());
    z.avail_in = (unsigned)s.size();
    size_t done = 0;
    bool   good = false;
    for (;;)
    {
        z.next_out  = reinterpret_cast<char *>(out.data()) + done;
        z.avail_out = (unsigned)std::min<size_t>(out.size() - done, 1u << 30);
        const unsigned before = z.avail_out;
        const int      rc     = L.run(&z);
        done += before - z.avail_out;
        if (rc == kBzStreamEnd)
        {
            good = true;
            break;
        }
        if (rc != kBzOk || (z.avail_in == 0 && z.avail_out != 0))
            break; // damaged, or the data ended inside the block
        if (z.avail_out == 0)
            out.resize(out.size() * 2);
    }
    L.end(&z);
    out.resize(good ? done : 0);
    return good;
}

class Bz2Source : public ByteSource
{
  public:
    Bz2Source(int fd, uint64_t size, int threads) : fd_(fd), size_(size), threads_(std::max(1, threads))
    {
        if (!bzlib().ok())
        {
            perr_     = "bzip2-compressed file, but libbz2.so.1.0 cannot be loaded";
            finished_ = true;
            return;
        }
        for (int i = 0; i < threads_; ++i)
            workers_.emplace_back([this] { work(); });
        producer_ = std::thread([this] { produce(); });
    }
    ~Bz2Source() override
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        if (producer_.joinable())
            producer_.join();
        for (auto &t : workers_)
            t.join();
        close(fd_);
    }
    bool     is_gzip() const override { return true; } // "compressed": the stream's length is not the file's
    uint64_t size() const override { return size_; }
    int64_t  read(char *dst, size_t cap) override
    {
        size_t got = 0;
        while (got < cap)
        {
            if (cur_off_ < cur_.size())
            {
                const size_t n = std::min(cap - got, cur_.size() - cur_off_);
                memcpy(dst + got, cur_.data() + cur_off_, n);
                cur_off_ += n;
                got += n;
                continue;
            }
            std::unique_lock<std::mutex> l(mu_);
            cv_.wait(l, [&] { return (!ready_.empty() && ready_.front()->done) || (finished_ && ready_.empty()); });
            if (!ready_.empty())
            {
                std::shared_ptr<Task> t = ready_.front();
                ready_.pop_front();
                queued_ -= 1;
                // a segment that does not decode: its end may be a chance occurrence of a block number inside compressed
                // data -- decode it together with the segment(s) after it before calling the file damaged
                for (int merged = 0; !t->seg.ok && merged < 3; ++merged)
                {
                    cv_.wait(l, [&] { return (!ready_.empty() && ready_.front()->done) || (finished_ && ready_.empty()); });
                    if (ready_.empty())
                        break;
                    std::shared_ptr<Task> nx = ready_.front();
                    ready_.pop_front();
                    queued_ -= 1;
                    l.unlock();
                    cv_.notify_all();
                    const uint64_t end_bit = t->seg.first_bit + t->seg.n_bits; // relative to t's first byte
                    if (end_bit & 7)
                        t->seg.bytes.pop_back(); // the byte shared with the successor
                    t->seg.bytes.insert(t->seg.bytes.end(), nx->seg.bytes.begin(), nx->seg.bytes.end());
                    t->seg.n_bits += nx->seg.n_bits;
                    t->seg.ok = decode_segment(t->seg.bytes.data(), t->seg.bytes.size(), t->seg.first_bit, t->seg.n_bits, t->seg.out);
                    l.lock();
                }
                if (!t->seg.ok)
                {
                    if (perr_.empty() || perr_ == "bzip2 stream ends inside a block")
                        perr_ = "damaged bzip2 block (or the data ends inside one)";
                    ready_.clear();
                    err_ = perr_;
                    l.unlock();
                    cv_.notify_all();
                    return got ? (int64_t)got : (int64_t)GNB_ERR_IO;
                }
                l.unlock();
                cv_.notify_all();
                cur_.swap(t->seg.out);
                cur_off_ = 0;
                continue;
            }
            if (!perr_.empty())
            {
                err_ = perr_;
                return got ? (int64_t)got : (int64_t)GNB_ERR_IO;
            }
            break;
        }
        return (int64_t)got;
    }

  private:
    struct Task
    {
        Segment seg;
        bool    done = false;
    };

    void work()
    {
        for (;;)
        {
            std::shared_ptr<Task> t;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || !todo_.empty(); });
                if (stop_)
                    return;
                t = todo_.front();
                todo_.pop_front();
            }
            t->seg.ok = decode_segment(t->seg.bytes.data(), t->seg.bytes.size(), t->seg.first_bit, t->seg.n_bits, t->seg.out);
            {
                std::lock_guard<std::mutex> l(mu_);
                t->done = true;
            }
            cv_.notify_all();
        }
    }

    // compressed bytes [lo_, lo_ + buf_.size()) of the file; more() appends, drop() forgets what lies before a position
    bool more()
    {
        const uint64_t at = lo_ + buf_.size() - kPad;
        if (at >= size_)
            return false;
        const size_t n = (size_t)std::min<uint64_t>(4u << 20, size_ - at);
        buf_.resize(buf_.size() + n);
        size_t done = 0;
        while (done < n)
        {
            const ssize_t r = pread(fd_, buf_.data() + buf_.size() - kPad - n + done, n - done, (off_t)(at + done));
            if (r <= 0)
            {
                io_failed_ = true;
                buf_.resize(buf_.size() - (n - done));
                return false;
            }
            done += (size_t)r;
        }
        memset(buf_.data() + buf_.size() - kPad, 0, kPad);
        return true;
    }

    void fail(const char *m)
    {
        std::lock_guard<std::mutex> l(mu_);
        if (perr_.empty())
            perr_ = m;
    }

    // a segment [from, to) (absolute bit positions) to the workers, in order; waits while too many are queued
    bool submit(uint64_t from, uint64_t to)
    {
        // test aid: every segment cut in two at a bit position that is no block start, as a chance occurrence of a block
        // number inside compressed data would cut it (read() then has to put the halves together again)
        static const bool split = getenv("GANON_B200_BZ2_SPLIT") != nullptr;
        if (split && to - from > 4096)
        {
            const uint64_t mid = from + (to - from) / 2 + 3;
            return submit_one(from, mid) && submit_one(mid, to);
        }
        return submit_one(from, to);
    }
    bool submit_one(uint64_t from, uint64_t to)
    {
        auto t = std::make_shared<Task>();
        const uint64_t b0 = from >> 3, b1 = (to + 7) >> 3;
        t->seg.bytes.assign(buf_.begin() + (long)(b0 - lo_), buf_.begin() + (long)(b1 - lo_));
        t->seg.first_bit = (unsigned)(from & 7);
        t->seg.n_bits    = to - from;
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || queued_ < (size_t)threads_ * 3; });
        if (stop_)
            return false;
        ready_.push_back(t);
        todo_.push_back(t);
        ++queued_;
        l.unlock();
        cv_.notify_all();
        return true;
    }

    // the scanner: reads the file front to back, cuts it at the block / end-of-stream numbers, queues the blocks
    void produce()
    {
        buf_.assign(kPad, 0);
        uint64_t pos       = 0;     // scan position (bytes)
        uint64_t start     = ~0ull; // bit position of the number that opens the segment being assembled
        bool     in_stream = false, ended = false;
        uint64_t expect    = 32; // where the first number of the (next) stream must start
        uint32_t combined  = 0;  // CRC of the open stream so far
        auto     number_at = [&](uint64_t bit) { return bits48(buf_.data(), bit - lo_ * 8); }; // absolute bit position
        auto     stream_header_at = [&](uint64_t byte) { // "BZh1" .. "BZh9"
            const uint8_t *h = buf_.data() + (byte - lo_);
            return byte >= lo_ && h[0] == 'B' && h[1] == 'Z' && h[2] == 'h' && h[3] >= '1' && h[3] <= '9';
        };
        for (;;)
        {
            // keep 16 bytes of look-ahead behind the scan position
            while (lo_ + buf_.size() - kPad < std::min<uint64_t>(size_, pos + 64) && more())
            {
            }
            if (io_failed_)
            {
                fail("short read");
                break;
            }
            const uint64_t have = lo_ + buf_.size() - kPad; // bytes of the file in the buffer
            if (pos + 8 > have)
            {
                if (have < size_)
                    continue;
                break; // the end of the file
            }
            // candidates: the number starts s bits into byte pos (its bytes 1, 2 are at pos + 1, pos + 2)
            const uint8_t *b = buf_.data() + (pos - lo_);
            const uint8_t  m = kPairs.mask[(unsigned)b[1] << 8 | b[2]];
            if (m)
                for (int s = 0; s < 8; ++s)
                    if (m >> s & 1)
                    {
                        const uint64_t bit = pos * 8 + (unsigned)s;
                        const uint64_t v   = number_at(bit);
                        if (v != kBlockMagic && v != kEndMagic)
                            continue;
                        if (bit + 48 > size_ * 8)
                            continue;
                        if (start != ~0ull)
                        {
                            if (!submit(start, bit))
                                return;
                        }
                        else if (bit != expect || !stream_header_at(bit / 8 - 4))
                        { // the first number of a stream stands right behind its 4-byte header: bytes were skipped (a
                          // damaged block number in front of this one), or this is something after the end of the data
                            if (!ended)
                                fail("damaged bzip2 stream (no block where the first one should start)");
                            goto done; // (after a complete stream: trailing bytes that are not a stream are ignored, as bzip2 does)
                        }
                        start     = v == kBlockMagic ? bit : ~0ull;
                        in_stream = v == kBlockMagic || bit + 80 > size_ * 8; // (the stream's CRC must follow its end number)
                        ended     = true;
                        // the 32 bits behind a number: the block's CRC, or the stream's = the blocks' CRCs folded together
                        const uint32_t crc = bit + 80 <= size_ * 8 ? (uint32_t)(number_at(bit + 32) & 0xffffffffu) : 0;
                        if (v == kBlockMagic)
                            combined = ((combined << 1) | (combined >> 31)) ^ crc;
                        else
                        {
                            if (!in_stream && crc != combined)
                            {
                                fail("bzip2 stream CRC mismatch");
                                goto done;
                            }
                            combined = 0;
                            expect   = ((bit + 80 + 7) >> 3) * 8 + 32; // a further stream starts at the next byte boundary
                        }
                        // forget the bytes in front of the open segment (or of the scan position)
                        const uint64_t keep = (start != ~0ull ? start >> 3 : pos);
                        if (keep > lo_ + (8u << 20))
                        {
                            buf_.erase(buf_.begin(), buf_.begin() + (long)(keep - lo_));
                            lo_ = keep;
                        }
                    }
            ++pos;
        }
    done:
        if (start != ~0ull || in_stream || !ended)
            fail("bzip2 stream ends inside a block"); // (keeps an earlier message)
        std::lock_guard<std::mutex> l(mu_);
        finished_ = true;
        cv_.notify_all();
    }

    static constexpr size_t kPad = 16;
    int                     fd_;
    uint64_t                size_;
    int                     threads_;
    std::vector<std::thread> workers_;
    std::thread              producer_;
    std::vector<uint8_t>     buf_;
    uint64_t                 lo_ = 0;
    bool                     io_failed_ = false;
    std::mutex               mu_;
    std::condition_variable  cv_;
    std::deque<std::shared_ptr<Task>> ready_, todo_;
    size_t                   queued_ = 0;
    std::vector<uint8_t>     cur_;
    size_t                   cur_off_ = 0;
    bool                     finished_ = false, stop_ = false;
    std::string              perr_;
};
} // namespace

std::unique_ptr<ByteSource> open_bz2_source(int fd, uint64_t size, int threads) { return std::unique_ptr<ByteSource>(new Bz2Source(fd, size, threads)); }

} // namespace gnb
