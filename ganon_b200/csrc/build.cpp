// ganon-build's count_hashes for one input file (src/ganon-build/GanonBuild.cpp:184-249) on the device: the file is read
// as the record reader sees it (plain or gzip, FASTA / FASTQ: reads.cpp, gzstream.cpp), every sequence of at least
// --min-length bases is cut into segments of kSegWindows windows (overlapping by w - 1 bases, so every window of the
// sequence lies in exactly one segment), K2 hashes the segments as if they were reads, and the file's DISTINCT minimisers
// -- what the reference collects in a robin_hood::unordered_set -- come from a radix sort + unique in HBM.  The set of
// minimiser values of a sequence is the union over its windows, so cutting changes nothing in the result.
// A sequence shorter than the window gets a window of its own length (seqan3 minimiser.hpp:298-299), one shorter than k
// has no k-mer.  A parse error drops the file's hashes (GanonBuild.cpp:241-245) but keeps the sequence counts.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "gnb_internal.h"
#include "gzstream.h"
#include "reads.h"

namespace gnb
{
size_t unique_tmp_bytes(uint64_t n);
// sorts `in` (n values) into `tmp_keys`, writes the distinct values to `out` and their number to *d_n_out
void launch_sort_unique(const uint64_t *in, uint64_t *tmp_keys, uint64_t *out, uint64_t n, unsigned long long *d_n_out, void *tmp, size_t tmp_bytes, cudaStream_t st);
} // namespace gnb

using namespace gnb;

struct gnb_hash_set
{
    int       device = 0;
    uint64_t *h = nullptr; // host copy of the distinct hashes (ascending)
    uint64_t  n = 0;
    ~gnb_hash_set() { free(h); }
};

namespace
{
constexpr uint32_t kSegWindows = 2048;

struct DevMem
{
    void  *p = nullptr;
    size_t cap = 0;
    int    ensure(size_t bytes, bool keep = false, size_t used = 0)
    {
        if (bytes <= cap)
            return GNB_OK;
        void        *q    = nullptr;
        const size_t want = bytes + bytes / 2 + 256;
        GNB_CUDA(cudaMalloc(&q, want));
        if (keep && p && used)
            GNB_CUDA(cudaMemcpy(q, p, used, cudaMemcpyDeviceToDevice));
        if (p)
            cudaFree(p);
        p   = q;
        cap = want;
        return GNB_OK;
    }
    ~DevMem()
    {
        if (p)
            cudaFree(p);
    }
    template <typename T>
    T *as() const
    {
        return reinterpret_cast<T *>(p);
    }
};

// Device and host buffers of one gnb_build_file_hashes call; finished calls hand them to the next one (a build makes one
// call per input file, possibly from several host threads at once: allocations and the zeroing of a fresh 256 MiB host buffer
// would otherwise dominate small genomes).
struct Builder
{
    int      device = 0;
    uint32_t k = 0, w = 0;
    cudaStream_t st = nullptr;
    DevMem   d_blk, d_off, d_len, d_cnt, d_hoff, d_tmp, d_all, d_sorted, d_uniq, d_n;
    uint64_t n_all = 0;
    std::unique_ptr<char[]> host; // file block (not zero-initialised)
    size_t   host_cap = 0;
    ~Builder()
    {
        if (st)
            cudaStreamDestroy(st);
    }
    int ensure_host(size_t bytes, size_t keep)
    {
        if (bytes <= host_cap)
            return GNB_OK;
        std::unique_ptr<char[]> nb(new (std::nothrow) char[bytes]);
        if (!nb)
            return fail(GNB_ERR_LIMIT, "out of host memory for the file block");
        if (keep)
            memcpy(nb.get(), host.get(), keep);
        host.swap(nb);
        host_cap = bytes;
        return GNB_OK;
    }

    // K2 over segments (off / len relative to the block text already in d_blk) with window w_eff; appends to d_all
    int hash_segments(const std::vector<uint32_t> &off, const std::vector<uint32_t> &len, uint32_t w_eff)
    {
        size_t done = 0;
        while (done < off.size())
        {
            const size_t n = std::min<size_t>(off.size() - done, kMaxReadsPerBatch - 2);
            GNB_TRY(d_off.ensure(n * 4));
            GNB_TRY(d_len.ensure(n * 4));
            GNB_TRY(d_cnt.ensure(n * 4));
            GNB_TRY(d_hoff.ensure((n + 1) * 8));
            GNB_TRY(d_tmp.ensure(scan_tmp_bytes((uint32_t)n)));
            GNB_CUDA(cudaMemcpyAsync(d_off.p, off.data() + done, n * 4, cudaMemcpyHostToDevice, st));
            GNB_CUDA(cudaMemcpyAsync(d_len.p, len.data() + done, n * 4, cudaMemcpyHostToDevice, st));
            launch_minimisers(d_blk.as<uint8_t>(), d_off.as<uint32_t>(), d_len.as<uint32_t>(), nullptr, nullptr, nullptr, (uint32_t)n, k, w_eff, 0, d_cnt.as<uint32_t>(),
                              nullptr, nullptr, nullptr, nullptr, st);
            launch_scan_counts(d_cnt.as<uint32_t>(), d_hoff.as<uint64_t>(), (uint32_t)n, d_tmp.p, d_tmp.cap, st);
            uint64_t total = 0;
            GNB_CUDA(cudaMemcpyAsync(&total, d_hoff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
            GNB_CUDA(cudaStreamSynchronize(st));
            if (total)
            {
                GNB_TRY(d_all.ensure((n_all + total) * 8, true, n_all * 8));
                // the write pass places read i at hash_off[i]: hand it the tail of d_all as its output array
                launch_minimisers(d_blk.as<uint8_t>(), d_off.as<uint32_t>(), d_len.as<uint32_t>(), nullptr, nullptr, nullptr, (uint32_t)n, k, w_eff, 1, nullptr,
                                  d_hoff.as<uint64_t>(), d_all.as<uint64_t>() + n_all, nullptr, nullptr, st);
                n_all += total;
            }
            GNB_CUDA(cudaStreamSynchronize(st)); // off / len are reused by the next piece
            GNB_CUDA(cudaGetLastError());
            done += n;
        }
        return GNB_OK;
    }
};
std::mutex                            g_pool_mu;
std::vector<std::unique_ptr<Builder>> g_pool;

std::unique_ptr<Builder> take_builder(int device)
{
    {
        std::lock_guard<std::mutex> l(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i]->device == device)
            {
                std::unique_ptr<Builder> b = std::move(g_pool[i]);
                g_pool.erase(g_pool.begin() + (long)i);
                return b;
            }
    }
    std::unique_ptr<Builder> b(new Builder);
    b->device = device;
    return b;
}

void give_builder(std::unique_ptr<Builder> b)
{
    std::lock_guard<std::mutex> l(g_pool_mu);
    if (g_pool.size() < 16)
        g_pool.push_back(std::move(b));
}
} // namespace

extern "C" int gnb_build_file_hashes(int device, const char *path, uint32_t k, uint32_t w, uint64_t min_length, int io_threads, gnb_hash_set **out,
                                     gnb_build_file_stats *stats)
{
    if (!path || !out || k < 1 || k > 32 || w < k || w - k + 1 > 256)
        return fail(GNB_ERR_ARG, "gnb_build_file_hashes: bad arguments");
    *out = nullptr;
    gnb_build_file_stats S{};
    GNB_CUDA(cudaSetDevice(device));
    std::string err;
    auto        src = open_byte_source(path, io_threads, err);
    if (!src)
        return fail(GNB_ERR_IO, err);
    std::unique_ptr<Builder> Bp = take_builder(device);
    struct Return
    {
        std::unique_ptr<Builder> &b;
        ~Return() { give_builder(std::move(b)); }
    } give_back{Bp};
    Builder &B = *Bp;
    B.k     = k;
    B.w     = w;
    B.n_all = 0;
    if (!B.st)
        GNB_CUDA(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
    // block size: the whole file when it is a small plain one, else 256 MiB (grown for longer records)
    size_t block = 256u << 20;
    if (!src->is_gzip() && src->size() + 64 < block)
        block = (size_t)src->size() + 64;
    GNB_TRY(B.ensure_host(block, 0));
    struct
    {
        Builder &B;
        size_t   n;
        char    *data() { return B.host.get(); }
        size_t   size() const { return n; }
    } buf{B, block};
    size_t            have = 0;
    bool              eof = false, parse_error = false;
    RecTable          t;
    std::vector<uint32_t> off, len, s_off, s_len;
    while (!parse_error)
    {
        while (!eof && have < buf.size())
        {
            const int64_t got = src->read(buf.data() + have, buf.size() - have);
            if (got < 0)
                return fail(GNB_ERR_IO, src->error());
            if (got == 0)
                eof = true;
            have += (size_t)got;
        }
        if (have == 0)
            break;
        t.clear();
        index_reads_host(buf.data(), have, eof, kMaxReadsPerBatch - 1, t);
        size_t n = t.size();
        if (t.parse_error && t.error_record <= n)
        {
            parse_error = true;
            n           = std::min<size_t>(n, (size_t)t.error_record);
        }
        if (n == 0 && !eof && !parse_error)
        {
            if (buf.size() >= (1ull << 31) - (1u << 20))
                return fail(GNB_ERR_LIMIT, "a single sequence record does not fit a 2 GiB block");
            buf.n = std::min<size_t>(buf.size() * 2, (1ull << 31) - (1u << 20));
            GNB_TRY(B.ensure_host(buf.n, have));
            continue;
        }
        // ---- segments of this block's records ----
        off.clear();
        len.clear();
        s_off.clear();
        s_len.clear();
        for (size_t r = 0; r < n; ++r)
        {
            const uint32_t L = t.seq_len[r];
            if (L < min_length)
            {
                ++S.n_skipped;
                continue;
            }
            ++S.n_sequences;
            S.n_bases += L;
            if (L < k)
                continue;
            if (L < w)
            { // the window shrinks to the sequence: one minimiser
                s_off.push_back(t.seq_off[r]);
                s_len.push_back(L);
                continue;
            }
            const uint32_t n_win = L - w + 1;
            for (uint32_t a = 0; a < n_win; a += kSegWindows)
            {
                off.push_back(t.seq_off[r] + a);
                len.push_back(std::min(kSegWindows, n_win - a) + w - 1);
            }
        }
        if (!off.empty() || !s_off.empty())
        {
            // block text, then the reader's side buffer (sequences assembled from wrapped lines): offsets >= have point there
            GNB_TRY(B.d_blk.ensure(have + t.aux.size() + 64));
            GNB_CUDA(cudaMemcpyAsync(B.d_blk.p, buf.data(), have, cudaMemcpyHostToDevice, B.st));
            if (!t.aux.empty())
                GNB_CUDA(cudaMemcpyAsync(B.d_blk.as<char>() + have, t.aux.data(), t.aux.size(), cudaMemcpyHostToDevice, B.st));
            GNB_TRY(B.hash_segments(off, len, w));
            // short sequences: grouped by length, each group with its own (clamped) window
            std::vector<size_t> order(s_off.size());
            for (size_t i = 0; i < order.size(); ++i)
                order[i] = i;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return s_len[a] < s_len[b]; });
            for (size_t i = 0; i < order.size();)
            {
                size_t j = i;
                off.clear();
                len.clear();
                while (j < order.size() && s_len[order[j]] == s_len[order[i]])
                {
                    off.push_back(s_off[order[j]]);
                    len.push_back(s_len[order[j]]);
                    ++j;
                }
                GNB_TRY(B.hash_segments(off, len, s_len[order[i]]));
                i = j;
            }
        }
        if (parse_error)
            break;
        const uint64_t used = t.consumed_for(n);
        if (eof && (used >= have || n == 0))
            break;
        memmove(buf.data(), buf.data() + used, have - used);
        have -= (size_t)used;
        if (eof && have == 0)
            break;
    }
    S.parse_error    = parse_error ? 1 : 0;
    S.n_hashes_total = B.n_all;
    std::unique_ptr<gnb_hash_set> hs(new gnb_hash_set);
    hs->device = device;
    if (!parse_error && B.n_all)
    {
        GNB_TRY(B.d_sorted.ensure(B.n_all * 8));
        GNB_TRY(B.d_uniq.ensure(B.n_all * 8));
        GNB_TRY(B.d_n.ensure(8));
        GNB_TRY(B.d_tmp.ensure(unique_tmp_bytes(B.n_all)));
        launch_sort_unique(B.d_all.as<uint64_t>(), B.d_sorted.as<uint64_t>(), B.d_uniq.as<uint64_t>(), B.n_all, B.d_n.as<unsigned long long>(), B.d_tmp.p, B.d_tmp.cap, B.st);
        unsigned long long nu = 0;
        GNB_CUDA(cudaMemcpyAsync(&nu, B.d_n.p, 8, cudaMemcpyDeviceToHost, B.st));
        GNB_CUDA(cudaStreamSynchronize(B.st));
        GNB_CUDA(cudaGetLastError());
        hs->n = nu;
        if (nu)
        {
            hs->h = static_cast<uint64_t *>(malloc(nu * 8));
            if (!hs->h)
                return fail(GNB_ERR_LIMIT, "out of host memory for the hash set");
            GNB_CUDA(cudaMemcpyAsync(hs->h, B.d_uniq.p, nu * 8, cudaMemcpyDeviceToHost, B.st));
            GNB_CUDA(cudaStreamSynchronize(B.st));
        }
    }
    S.n_unique = hs->n;
    if (stats)
        *stats = S;
    *out = hs.release();
    return GNB_OK;
}

extern "C" int gnb_hash_set_data(const gnb_hash_set *s, const uint64_t **hashes, uint64_t *n)
{
    if (!s || !hashes || !n)
        return fail(GNB_ERR_ARG, "gnb_hash_set_data: bad arguments");
    *hashes = s->h;
    *n      = s->n;
    return GNB_OK;
}

extern "C" void gnb_hash_set_free(gnb_hash_set *s) { delete s; }
