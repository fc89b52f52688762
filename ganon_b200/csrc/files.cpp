// gnb_session_classify_files: the reader / classify / writer loop of one read file (or pair) -- what ganon-classify's
// parse_reads (GC.cpp:1220-1287), classify threads (GC.cpp:630-832) and write_classified / write_unclassified threads
// (GC.cpp:1289-1322) do with their queues, here as: block ring in page-locked memory <- ByteSource (parallel preads or
// parallel inflate, gzstream.h) -> gnb_session_submit (H2D + K1) -> batches in flight on the GPU -> gnb_session_collect ->
// writer thread.  A block that is being classified is never touched: buffers are only replaced after every batch in
// flight has been collected.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gnb_internal.h"
#include "gzstream.h"

namespace gnb
{
// session.cpp: how the session takes its blocks (bin-sharded runs with sliced ingest read only this rank's slice)
void session_ingest_mode(const gnb_session *s, int *sliced, int *rank, int *n_ranks);
int  set_blocking_waits(int on);

namespace
{
using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

// Page-locked block buffers are expensive to make (~70 ms per 64 MiB) and every file needs a ring of them: buffers of
// finished files are kept and handed to the next file of the process.
class PinnedPool
{
  public:
    static PinnedPool &instance()
    {
        static PinnedPool *p = new PinnedPool(); // lives until the process ends (the driver frees the pages)
        return *p;
    }
    char *get(size_t bytes)
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].second >= bytes && free_[i].second <= bytes + bytes / 2)
                {
                    char *p = free_[i].first;
                    sizes_.emplace_back(p, free_[i].second);
                    free_.erase(free_.begin() + (long)i);
                    return p;
                }
        }
        char *p = nullptr;
        if (cudaMallocHost((void **)&p, bytes) != cudaSuccess)
            return nullptr;
        std::lock_guard<std::mutex> l(mu_);
        sizes_.emplace_back(p, bytes);
        return p;
    }
    void put(char *p)
    {
        std::lock_guard<std::mutex> l(mu_);
        for (size_t i = 0; i < sizes_.size(); ++i)
            if (sizes_[i].first == p)
            {
                free_.push_back(sizes_[i]);
                sizes_.erase(sizes_.begin() + (long)i);
                break;
            }
        // keep at most 16 idle buffers
        while (free_.size() > 16)
        {
            cudaFreeHost(free_.front().first);
            free_.erase(free_.begin());
        }
    }

  private:
    std::mutex                             mu_;
    std::vector<std::pair<char *, size_t>> free_, sizes_;
};

// One read file as a sequence of blocks.  Stream form (gzip, or any file of an unsliced session): every ring buffer has a
// headroom in front of the fresh bytes; the unconsumed tail of a block (whole records the session held back, plus the
// partial record at its end) is copied right-aligned into the next buffer's headroom, so fresh bytes never move and the
// next buffer is filled by a background thread while the current block is staged.  Positional form (plain file, sliced
// ingest): the block is the file range [pos, pos + len) of which only this rank's slice is read; pos advances by the
// consumed bytes.
class BlockStream
{
  public:
    BlockStream(std::unique_ptr<ByteSource> src, size_t block_bytes, int n_buffers, bool positional, int rank, int n_ranks)
        : src_(std::move(src)), block_(block_bytes), positional_(positional), rank_(rank), n_ranks_(n_ranks), bufs_((size_t)n_buffers, nullptr)
    {
        worker_ = std::thread([this] { prefetch_loop(); });
    }
    ~BlockStream()
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        worker_.join();
        release();
    }
    // Buffers are page-locked when they are first used (pinning 64 MiB takes tens of milliseconds: a ring pinned up front
    // cost 0.3 s before the first byte was read, and a small file never needs most of it); the prefetch thread pins the
    // buffer it fills and, once idle, the one after it.
    int alloc()
    {
        release();
        cur_ = -1;
        return GNB_OK;
    }
    void release()
    {
        std::lock_guard<std::mutex> l(alloc_mu_);
        ++gen_;
        for (auto &b : bufs_)
            if (b)
            {
                PinnedPool::instance().put(b);
                b = nullptr;
            }
    }
    void resize(size_t head, size_t block)
    {
        std::lock_guard<std::mutex> l(alloc_mu_);
        head_  = head;
        block_ = block;
        ++gen_;
    }
    char *buffer(int idx)
    {
        for (;;)
        {
            size_t   bytes;
            uint64_t gen;
            {
                std::lock_guard<std::mutex> l(alloc_mu_);
                if (bufs_[idx] != nullptr)
                    return bufs_[idx];
                bytes = head_ + block_ + 64;
                gen   = gen_;
            }
            char *p = PinnedPool::instance().get(bytes); // not under the lock: the other thread keeps using its buffers
            if (p == nullptr)
                return nullptr;
            std::lock_guard<std::mutex> l(alloc_mu_);
            if (gen == gen_ && bufs_[idx] == nullptr)
                return bufs_[idx] = p;
            PinnedPool::instance().put(p); // the ring was rebuilt with other sizes meanwhile, or the other thread was first
        }
    }
    const char *ptr() const { return positional_ ? bufs_[cur_] : bufs_[cur_] + start_; }
    // first byte of the current block (a rank of a sliced run holds only its slice of the block: read it from the file)
    bool first_byte(char *c)
    {
        if (!positional_ || rank_ == 0)
        {
            *c = ptr()[0];
            return true;
        }
        return src_->read_at(c, 1, pos_) == 1;
    }
    size_t      fill() const { return fill_; }
    bool        eof() const { return eof_; }
    uint64_t    bytes_in() const { return bytes_in_; }

    // the next block; < 0 on an I/O error.  Callers drain the batches in flight before a call that may rebuild buffers
    // (needs_rebuild()).
    bool needs_rebuild() const { return !positional_ && tail_.size() > head_; }
    int  next_block()
    {
        const int idx = (cur_ + 1) % (int)bufs_.size();
        if (positional_)
        {
            cur_             = idx;
            const uint64_t n = src_->size() > pos_ ? std::min<uint64_t>(block_, src_->size() - pos_) : 0;
            const uint64_t N = (uint64_t)n_ranks_, r = (uint64_t)rank_;
            const uint64_t slice = (((n + N - 1) / N) + 15) & ~15ull; // the session's slicing rule (stage(), session.cpp)
            const uint64_t lo = std::min(n, r * slice), hi = std::min(n, (r + 1) * slice);
            if (buffer(idx) == nullptr)
                return fail(GNB_ERR_CUDA, "cannot allocate page-locked read buffers");
            if (hi > lo)
            {
                const int64_t got = src_->read_at(bufs_[idx] + lo, hi - lo, pos_ + lo);
                if (got != (int64_t)(hi - lo))
                    return fail(GNB_ERR_IO, "short read: " + src_->error());
                bytes_in_ += hi - lo;
            }
            fill_ = n;
            eof_  = pos_ + n >= src_->size();
            return GNB_OK;
        }
        if (needs_rebuild())
        {
            // a tail longer than the headroom (a record of more than 1 MiB): more room in front; the caller has drained
            wait_prefetch();
            std::string fresh;
            if (pf_idx_ == idx && pf_n_ > 0)
                fresh.assign(bufs_[idx] + head_, (size_t)pf_n_);
            const bool had = pf_idx_ == idx;
            const int64_t n_prev = pf_n_;
            resize(2 * tail_.size(), block_);
            GNB_TRY(alloc());
            if (had)
            {
                if (n_prev < 0)
                    return fail(GNB_ERR_IO, src_->error());
                if (buffer(0) == nullptr)
                    return fail(GNB_ERR_CUDA, "cannot allocate page-locked read buffers");
                memcpy(bufs_[0] + head_, fresh.data(), fresh.size());
                pf_idx_ = 0;
            }
            return place(0);
        }
        return place(idx);
    }
    void consume(uint64_t n)
    {
        if (positional_)
            pos_ += n;
        else
            tail_.assign(ptr() + n, fill_ - n);
    }
    // not one complete record in the block: double the block size (the caller has drained the batches in flight)
    int grow()
    {
        if (positional_)
        {
            resize(head_, block_ * 2);
            return alloc();
        }
        wait_prefetch();
        tail_.assign(ptr(), fill_);
        if (pf_idx_ >= 0 && pf_n_ > 0)
            tail_.append(bufs_[pf_idx_] + head_, (size_t)pf_n_); // bytes already taken from the file stay in order
        if (pf_idx_ >= 0 && pf_n_ < 0)
            return fail(GNB_ERR_IO, src_->error());
        pf_idx_ = -1;
        resize(std::max(head_, 2 * tail_.size()), block_ * 2);
        return alloc();
    }

  private:
    int place(int idx)
    {
        if (pf_idx_ != idx && !raw_eof_)
            request(idx);
        int64_t fresh = 0;
        if (pf_idx_ == idx)
        {
            wait_prefetch();
            if (pf_n_ < 0)
                return fail(pf_n_ == GNB_ERR_CUDA ? GNB_ERR_CUDA : GNB_ERR_IO, pf_n_ == GNB_ERR_CUDA ? "cannot allocate page-locked read buffers" : src_->error());
            fresh   = pf_n_;
            pf_idx_ = -1;
        }
        cur_ = idx;
        if (buffer(idx) == nullptr)
            return fail(GNB_ERR_CUDA, "cannot allocate page-locked read buffers");
        memcpy(bufs_[idx] + head_ - tail_.size(), tail_.data(), tail_.size());
        start_ = head_ - tail_.size();
        fill_  = tail_.size() + (size_t)fresh;
        tail_.clear();
        eof_ = raw_eof_;
        if (!raw_eof_)
            request((idx + 1) % (int)bufs_.size()); // overlaps with the staging of this block
        return GNB_OK;
    }
    void request(int idx)
    {
        std::lock_guard<std::mutex> l(mu_);
        pf_idx_  = idx;
        pf_busy_ = true;
        cv_.notify_all();
    }
    void wait_prefetch()
    {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return !pf_busy_; });
    }
    void prefetch_loop()
    {
        for (;;)
        {
            int idx;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || pf_busy_; });
                if (stop_)
                    return;
                idx = pf_idx_;
            }
            int64_t total = 0;
            char   *buf   = buffer(idx);
            if (buf == nullptr)
                total = GNB_ERR_CUDA;
            while (buf != nullptr && (size_t)total < block_)
            {
                const int64_t got = src_->read(buf + head_ + total, block_ - (size_t)total);
                if (got < 0)
                {
                    total = got;
                    break;
                }
                if (got == 0)
                {
                    raw_eof_ = true;
                    break;
                }
                total += got;
            }
            if (total > 0)
                bytes_in_ += (uint64_t)total;
            {
                std::lock_guard<std::mutex> l(mu_);
                pf_n_    = total;
                pf_busy_ = false;
                cv_.notify_all();
            }
            if (total > 0 && !raw_eof_)
                buffer((idx + 1) % (int)bufs_.size()); // idle until the next request: pin the buffer it will name
        }
    }

    std::unique_ptr<ByteSource> src_;
    size_t                      block_, head_ = 1u << 20;
    bool                        positional_;
    int                         rank_, n_ranks_;
    std::vector<char *>         bufs_;
    uint64_t                    gen_ = 0; // counts rebuilds of the ring (alloc_mu_)
    int                         cur_ = -1;
    size_t                      start_ = 0, fill_ = 0;
    std::string                 tail_;
    bool                        eof_ = false;
    volatile bool               raw_eof_ = false;
    uint64_t                    pos_ = 0, bytes_in_ = 0;
    std::thread                 worker_;
    std::mutex                  mu_, alloc_mu_; // alloc_mu_: bufs_ entries
    std::condition_variable     cv_;
    int                         pf_idx_ = -1;
    int64_t                     pf_n_ = 0;
    bool                        pf_busy_ = false, stop_ = false;
};

// output text goes through one writer thread (the reference has one per output kind)
class Writer
{
  public:
    Writer() { th_ = std::thread([this] { loop(); }); }
    ~Writer() { finish(); }
    void put(int fd, const char *p, uint64_t n)
    {
        if (fd < 0 || n == 0)
            return;
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return queued_ < (512u << 20); });
        q_.emplace_back(fd, std::string(p, (size_t)n));
        queued_ += n;
        cv_.notify_all();
    }
    bool finish()
    {
        if (th_.joinable())
        {
            {
                std::lock_guard<std::mutex> l(mu_);
                done_ = true;
            }
            cv_.notify_all();
            th_.join();
        }
        return !failed_;
    }
    double ms_write = 0;

  private:
    void loop()
    {
        for (;;)
        {
            std::pair<int, std::string> item;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return done_ || !q_.empty(); });
                if (q_.empty())
                    return;
                item = std::move(q_.front());
                q_.pop_front();
            }
            auto        t0 = Clock::now();
            const char *p  = item.second.data();
            size_t      left = item.second.size();
            while (left && !failed_)
            {
                const ssize_t w = write(item.first, p, left);
                if (w <= 0)
                {
                    failed_ = true;
                    break;
                }
                p += w;
                left -= (size_t)w;
            }
            ms_write += ms_since(t0);
            std::lock_guard<std::mutex> l(mu_);
            queued_ -= item.second.size();
            cv_.notify_all();
        }
    }
    std::thread                             th_;
    std::mutex                              mu_;
    std::condition_variable                 cv_;
    std::deque<std::pair<int, std::string>> q_;
    uint64_t                                queued_ = 0;
    bool                                    done_ = false;
    volatile bool                           failed_ = false;
};

} // namespace
} // namespace gnb

using namespace gnb;

extern "C" int gnb_session_classify_files(gnb_session *s, uint32_t prefix_id, const char *file1, const char *file2, const gnb_output_fds *out,
                                          uint64_t block_bytes, int io_threads, gnb_files_result *res)
{
    if (!s || !file1)
        return fail(GNB_ERR_ARG, "gnb_session_classify_files: bad arguments");
    if (block_bytes == 0)
        block_bytes = 64ull << 20;
    if (block_bytes >= (1ull << 31))
        return fail(GNB_ERR_LIMIT, "read blocks are limited to 2 GiB");
    gnb_files_result R{};
    uint32_t         n_in_flight = 0, capacity = 1;
    GNB_TRY(gnb_session_in_flight(s, &n_in_flight, &capacity));
    if (n_in_flight)
        return fail(GNB_ERR_ARG, "gnb_session_classify_files: batches are in flight, collect first");
    int sliced = 0, rank = 0, n_ranks = 1;
    session_ingest_mode(s, &sliced, &rank, &n_ranks);
    struct BlockingWaits
    {
        int prev = set_blocking_waits(1);
        ~BlockingWaits() { set_blocking_waits(prev); }
    } blocking_waits;
    const bool paired = file2 && file2[0];
    auto       t_open = Clock::now();
    std::unique_ptr<BlockStream> st[2];
    for (int k = 0; k < (paired ? 2 : 1); ++k)
    {
        std::string err;
        auto        src = open_byte_source(k ? file2 : file1, io_threads, err, paired ? 2 : 1);
        if (!src)
            return fail(GNB_ERR_IO, err);
        R.is_gzip |= src->is_gzip() ? 1 : 0;
        const bool positional = sliced && src->seekable();
        st[k].reset(new BlockStream(std::move(src), (size_t)block_bytes, (int)capacity + 2, positional, rank, n_ranks));
        GNB_TRY(st[k]->alloc());
    }
    R.ms_open = ms_since(t_open);
    // the reference reads a file in the format its NAME says: an unknown extension ends its run (uncaught exception), content
    // of the other format is a parse error on the first record -- nothing of the file is classified (GC.cpp:1278-1283)
    for (int k = 0; k < (paired ? 2 : 1); ++k)
        if (format_of_extension(k ? file2 : file1) == 0)
            return fail(GNB_ERR_PARSE, std::string("unknown file extension (the reference reads .fasta/.fa/.fna/.ffn/.faa/.frn/.fas, .fastq/.fq, .embl, "
                                                   ".genbank/.gb/.gbk and .sam, optionally compressed): ") +
                                           (k ? file2 : file1));
    Writer   writer;
    uint32_t pending = 0;
    int      rc = GNB_OK;
    bool     first_block = true;
    auto     collect_one = [&]() -> int {
        gnb_batch_result r{};
        auto             t0 = Clock::now();
        const int        c  = gnb_session_collect(s, &r);
        R.ms_collect += ms_since(t0);
        --pending;
        if (c != GNB_OK)
            return c;
        R.n_classified += r.n_classified;
        if (out)
        {
            for (uint32_t li = 0; li < r.n_levels && li < out->n_levels; ++li)
            {
                if (out->all_fd)
                    writer.put(out->all_fd[li], r.all_text[li], r.all_len[li]);
                if (out->one_fd)
                    writer.put(out->one_fd[li], r.one_text[li], r.one_len[li]);
            }
            writer.put(out->unc_fd, r.unc_text, r.unc_len);
        }
        return GNB_OK;
    };
    auto drain = [&]() -> int {
        int first = GNB_OK;
        while (pending)
        {
            const int c = collect_one();
            if (c != GNB_OK && first == GNB_OK)
                first = c;
        }
        return first;
    };
    for (;;)
    {
        auto t0 = Clock::now();
        if (st[0]->needs_rebuild() || (paired && st[1]->needs_rebuild()))
            if ((rc = drain()) != GNB_OK)
                break;
        if ((rc = st[0]->next_block()) != GNB_OK || (paired && (rc = st[1]->next_block()) != GNB_OK))
            break;
        R.ms_read_wait += ms_since(t0);
        const bool final = st[0]->eof() && (!paired || st[1]->eof());
        if (st[0]->fill() == 0 || (paired && st[1]->fill() == 0))
            break; // nothing (more) to pair
        if (first_block)
        {
            first_block = false;
            bool mismatch = false;
            for (int k = 0; k < (paired ? 2 : 1); ++k)
            {
                const int fmt = format_of_extension(k ? file2 : file1); // EMBL / GenBank / SAM arrive rewritten as FASTA
                char      c   = 0;
                if (!st[k]->first_byte(&c))
                {
                    rc = fail(GNB_ERR_IO, "short read");
                    break;
                }
                mismatch |= (fmt != kFormatFastq && c != '>' && c != ';') || (fmt == kFormatFastq && c != '@');
            }
            if (rc != GNB_OK)
                break;
            if (mismatch)
            {
                fprintf(stderr, "Error parsing file(s) [%s, %s]: the content is not in the format of the file extension\n", file1, paired ? file2 : "");
                R.parse_error = 1;
                break;
            }
        }
        ++R.n_blocks;
        gnb_batch_result info{};
        t0 = Clock::now();
        rc = gnb_session_submit(s, prefix_id, st[0]->ptr(), st[0]->fill(), paired ? st[1]->ptr() : nullptr, paired ? st[1]->fill() : 0, final ? 1 : 0, &info);
        R.ms_submit += ms_since(t0);
        if (rc != GNB_OK)
            break;
        ++pending;
        R.n_records += info.n_reads;
        const bool stop = info.parse_error || final || info.n_reads == 0;
        while (pending > (stop ? 0u : capacity - 1))
            if ((rc = collect_one()) != GNB_OK)
                break;
        if (rc != GNB_OK)
            break;
        if (info.parse_error)
        {
            R.parse_error = 1;
            break; // the rest of the file is skipped (GC.cpp:1278-1283)
        }
        if (info.n_reads == 0 && !final)
        {
            if (block_bytes * 2 >= (1ull << 31))
            {
                rc = fail(GNB_ERR_LIMIT, "a single record does not fit a 1 GiB block");
                break;
            }
            block_bytes *= 2;
            if ((rc = st[0]->grow()) != GNB_OK || (paired && (rc = st[1]->grow()) != GNB_OK))
                break;
            continue;
        }
        st[0]->consume(info.consumed1);
        if (paired)
            st[1]->consume(info.consumed2);
        if (final)
            break;
    }
    const std::string err = rc != GNB_OK ? gnb_last_error() : "";
    (void)drain();
    const bool wrote = writer.finish();
    R.ms_write       = writer.ms_write;
    R.bytes_read1    = st[0]->bytes_in();
    R.bytes_read2    = paired ? st[1]->bytes_in() : 0;
    if (res)
        *res = R;
    if (rc != GNB_OK)
        return fail(rc, err);
    if (!wrote)
        return fail(GNB_ERR_IO, "writing output files failed");
    return GNB_OK;
}
