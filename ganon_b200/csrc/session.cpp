// Session runtime: one classification run (all hierarchy levels) on one GPU.
//   stage   : record index (device K1 for strict FASTQ, host reader otherwise) + host->device copies
//   run     : K2 minimisers, K3 IBF counts (+ sort of the sparse tuples) for the first level
//   finish  : device->host of the sparse tuples, then the host finishing stage in C++ threads --
//             cross-filter merge (select_matches GC.cpp:504-541), rel-filter + fpr-query (filter_matches GC.cpp:579-613,
//             double-precision libm so that decisions are bit-identical to the reference), unique/LCA
//             (GC.cpp:615-627, 766-803), report accounting (Rep/Total GC.cpp:153-177) and the text of
//             .all/.one/.unc (GC.cpp:1289-1322); remaining levels re-run K3 on the still unclassified reads.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "comm.h"
#include "db.h"
#include "em_merge.h"
#include "reads.h"

namespace gnb
{
namespace
{

using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

// GANON_B200_TRACE=1: wall-clock phase marks of every batch on stderr (pipeline debugging)
inline bool trace_on()
{
    static const bool on = [] { const char *e = getenv("GANON_B200_TRACE"); return e && e[0] == '1'; }();
    return on;
}
inline void trace_mark(const void *ctx, const char *what)
{
    if (!trace_on())
        return;
    static const Clock::time_point t0 = Clock::now();
    fprintf(stderr, "[gnb-trace] %p %-18s %9.3f ms\n", ctx, what, ms_since(t0));
}

// Host wait for a stream.  Default: cudaStreamSynchronize (spins: lowest latency).  GANON_B200_SYNC=block (and every file
// run, see set_blocking_waits): a yielding wait, so that the waiting thread gives its core away -- several ranks per node
// with a few threads each, or a file run with its reader / inflater / writer threads, would otherwise keep more spinning
// threads than the host has cores.
// gnb_session_classify_files turns blocking waits on for its duration unless GANON_B200_SYNC=spin says otherwise: its reader,
// inflater and writer threads need the cores the spinning waits would burn (measured on a 16-core host, c2 file to file:
// 33.6 -> 51 M reads/s).
std::atomic<int> g_sync_block{[] { const char *e = getenv("GANON_B200_SYNC"); return e && e[0] == 'b' ? 1 : 0; }()};
inline cudaError_t stream_wait(cudaStream_t st)
{
    const bool block = g_sync_block.load(std::memory_order_relaxed) != 0;
    if (!block)
        return cudaStreamSynchronize(st);
    // Yielding wait: poll the stream, spinning for the first 50 us and sleeping 40 us between polls after that.  The core
    // is free during a long kernel, and the wake-up costs ~0.1 ms at most (an interrupt-driven cudaEventBlockingSync wait
    // was measured at milliseconds per wake-up on some hosts, which stalls a chain of ~8 waits per batch).
    const auto t0 = Clock::now();
    for (;;)
    {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady)
            return e;
        if (std::chrono::duration<double, std::micro>(Clock::now() - t0).count() < 50.0)
            continue;
        std::this_thread::sleep_for(std::chrono::microseconds(40));
    }
}

struct DevBuf
{
    void  *p   = nullptr;
    size_t cap = 0;
    int    ensure(size_t bytes)
    {
        if (bytes <= cap)
            return GNB_OK;
        if (p)
            cudaFree(p);
        p   = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        GNB_CUDA(cudaMalloc(&p, want));
        cap = want;
        return GNB_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const
    {
        return reinterpret_cast<T *>(p);
    }
};

struct PinBuf
{
    void  *p   = nullptr;
    size_t cap = 0;
    int    ensure(size_t bytes)
    {
        if (bytes <= cap)
            return GNB_OK;
        if (p)
            cudaFreeHost(p);
        p   = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        GNB_CUDA(cudaMallocHost(&p, want));
        cap = want;
        return GNB_OK;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
    }
    template <typename T>
    T *as() const
    {
        return reinterpret_cast<T *>(p);
    }
};

// std::vector in page-locked host memory: device<->host copies into it are asynchronous and run at link speed
template <typename T>
struct PinnedAlloc
{
    using value_type = T;
    PinnedAlloc() = default;
    template <typename U>
    PinnedAlloc(const PinnedAlloc<U> &)
    {
    }
    T *allocate(size_t n)
    {
        void *p = nullptr;
        if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess)
            throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t) { cudaFreeHost(p); }
    template <typename U>
    bool operator==(const PinnedAlloc<U> &) const
    {
        return true;
    }
    template <typename U>
    bool operator!=(const PinnedAlloc<U> &) const
    {
        return false;
    }
};
template <typename T>
using PinnedVec = std::vector<T, PinnedAlloc<T>>;

// Matches of every classified read of a run (one EM group) kept in HBM for gnb_session_reassign.
struct EmStore
{
    DevBuf   off, tgt, cnt, id_off, ids;
    uint64_t n_reads = 0, n_matches = 0, id_bytes = 0;
    // grow a buffer keeping its contents (rare: capacities double)
    static int grow(DevBuf &b, size_t used, size_t need)
    {
        if (need <= b.cap)
            return GNB_OK;
        DevBuf nb;
        GNB_TRY(nb.ensure(std::max(need, b.cap * 2)));
        GNB_CUDA(cudaDeviceSynchronize()); // appends of earlier batches may still be writing the old buffer
        if (used && b.p)
            GNB_CUDA(cudaMemcpy(nb.p, b.p, std::min(used, b.cap), cudaMemcpyDeviceToDevice));
        b.release();
        b = nb;
        return GNB_OK;
    }
    int reserve(uint64_t add_reads, uint64_t add_matches, uint64_t add_ids)
    {
        GNB_TRY(grow(off, (n_reads + 1) * 8, (n_reads + add_reads + 1) * 8));
        GNB_TRY(grow(id_off, (n_reads + 1) * 8, (n_reads + add_reads + 1) * 8));
        GNB_TRY(grow(tgt, n_matches * 4, (n_matches + add_matches) * 4 + 4));
        GNB_TRY(grow(cnt, n_matches * 4, (n_matches + add_matches) * 4 + 4));
        GNB_TRY(grow(ids, id_bytes, id_bytes + add_ids + 1));
        return GNB_OK;
    }
    EmStoreDev dev() const { return EmStoreDev{off.as<uint64_t>(), tgt.as<uint32_t>(), cnt.as<uint32_t>(), id_off.as<uint64_t>(), ids.as<char>()}; }
    void release()
    {
        for (DevBuf *b : {&off, &tgt, &cnt, &id_off, &ids})
            b->release();
        n_reads = n_matches = id_bytes = 0;
    }
};

struct Rep // GC.cpp:153-160
{
    uint64_t matches = 0, seqs_lca = 0, seqs_unique = 0, discarded_matches_filter = 0, discarded_matches_fprquery = 0;
    void     add(const Rep &o)
    {
        matches += o.matches;
        seqs_lca += o.seqs_lca;
        seqs_unique += o.seqs_unique;
        discarded_matches_filter += o.discarded_matches_filter;
        discarded_matches_fprquery += o.discarded_matches_fprquery;
    }
};

inline void add_totals(gnb_totals &a, const gnb_totals &b)
{
    a.input_seqs += b.input_seqs;
    a.seqs_processed += b.seqs_processed;
    a.seqs_skipped_big += b.seqs_skipped_big;
    a.seqs_skipped_small += b.seqs_skipped_small;
    a.length_processed += b.length_processed;
    a.kmers_processed += b.kmers_processed;
    a.seqs_classified += b.seqs_classified;
    a.kmers_matches += b.kmers_matches;
    a.kmers_from_classified_seqs += b.kmers_from_classified_seqs;
    a.matches += b.matches;
    a.seqs_unique += b.seqs_unique;
    a.discarded_matches_filter += b.discarded_matches_filter;
    a.discarded_matches_fprquery += b.discarded_matches_fprquery;
}

struct FilterRt
{
    gnb_db     *db = nullptr;
    double      rel_cutoff = 0;
    std::string tax_file;
    IbfDev      dev{};
    DevBuf      d_single, d_bin_node, d_seg_off, d_segs;
    std::vector<double>   node_fpr;   // [n_level_targets] fpr of this filter's target for the node (0 if absent)
    std::vector<uint8_t>  node_fpr_class; // [n_level_targets] index into LevelRt::fpr_classes, 255 = none
    std::vector<uint8_t>  node_multi; // [n_level_targets] 1: node's bins form several segments (partial tuples)
    // host-resident tier: one device view + tables per column page of a paged filter (`dev` then describes page 0)
    struct PageRt
    {
        IbfDev dev{};
        DevBuf d_single, d_bin_node, d_seg_off, d_segs;
    };
    std::vector<PageRt> pages;
    // HIBF: one IbfDev per sub-IBF (tables carved out of the shared device arrays above)
    bool                  is_hibf = false;
    std::vector<IbfDev>   ibf_table;
    std::vector<uint32_t> round_lanes; // [depth] lanes per item of the traversal round (1..16), 0 = whole warp (wide sub-IBFs)
    DevBuf                d_ibf_table;
};

struct LevelRt
{
    std::string              label;
    std::vector<FilterRt>    filters;
    double                   rel_filter = 0, fpr_query = 1;
    uint32_t                 k = 0, w = 0;
    // levels with several filters finished by K4: the filter index rides in the low bits of a tuple's node field (0: the
    // level has one filter, or its cross-filter merge stays in the host finishing stage)
    uint32_t                 filter_bits = 0;
    std::string              out_one = "one", out_all = "all";
    // nodes: [0, n_targets) targets of the filters, then the remaining taxonomy nodes
    std::vector<std::string> node_names;
    uint32_t                 n_targets = 0;
    bool                     has_tax   = false;
    std::vector<std::string> node_rank, node_tax_name;
    std::vector<int32_t>     parent;
    std::vector<uint32_t>    depth;
    int32_t                  root = -1;
    std::vector<double>      fpr_classes; // distinct per-target fpr values (at most kMemoClasses, else empty)
    // accounting per prefix
    std::vector<std::unordered_map<uint32_t, Rep>> rep;
    std::vector<gnb_totals>                        total;
    // K4 (finishing stage on the device; levels with one filter): node tables and per-prefix report accumulators in HBM
    bool                device_finish = false;
    DevBuf              d_node_fpr, d_node_class, d_fpr_memo, d_parent, d_depth, d_name_off, d_names;
    std::vector<DevBuf> d_rep; // [prefix] -> unsigned long long [n_nodes][5]
    // EM reassignment: run-wide target ids (by name, as the reference keys its dictionaries) of this level's nodes
    std::vector<uint32_t> em_map;
    DevBuf                d_em_map;

    uint32_t lca2(uint32_t u, uint32_t v) const
    {
        int32_t a = (int32_t)u, b = (int32_t)v;
        while (a >= 0 && b >= 0 && a != b)
        {
            if (depth[a] > depth[b])
                a = parent[a];
            else if (depth[b] > depth[a])
                b = parent[b];
            else
            {
                a = parent[a];
                b = parent[b];
            }
        }
        if (a < 0 || b < 0)
            return (uint32_t)root;
        return (uint32_t)a;
    }
};

struct FprKey
{
    uint64_t fpr_bits;
    uint32_t n, c;
    bool     operator==(const FprKey &o) const { return fpr_bits == o.fpr_bits && n == o.n && c == o.c; }
};
struct FprKeyHash
{
    size_t operator()(const FprKey &k) const { return (size_t)splitmix64(k.fpr_bits ^ ((uint64_t)k.n << 32 | k.c)); }
};

// GC.cpp:498-501 + 588-601, evaluated exactly as the reference does (libm double), memoised per (n, count, fpr)
inline double fpr_query_q(uint64_t n_hashes, uint64_t count, double fpr)
{
    double q = 1;
    for (size_t i = 0; i <= count; i++)
    {
        const double n = (double)n_hashes, k = (double)i;
        const double binom = std::exp(std::lgamma(n + 1) - std::lgamma(n - k + 1) - std::lgamma(k + 1));
        q -= binom * pow(fpr, i) * pow(1 - fpr, n_hashes - i);
    }
    return q;
}

constexpr size_t kDenseRepNodes = 1u << 16;
constexpr size_t kFprMemoSlots  = 1u << 20; // K4's --fpr-query cache: 16 MiB per level
constexpr size_t kMemoClasses   = 8;

struct Worker
{
    std::unordered_map<FprKey, double, FprKeyHash> memo;
    std::vector<std::string>                       all_text, one_text; // per level
    std::string                                    unc_text;
    std::vector<std::unordered_map<uint32_t, Rep>> rep;   // per level (large node tables)
    std::vector<std::vector<Rep>>                  rep_dense;   // per level (node tables up to kDenseRepNodes)
    std::vector<std::vector<uint32_t>>             rep_touched; // nodes with a non-zero dense entry
    std::vector<double>                            memo_table;  // [fpr class][n < 256][count < 256], NaN = not computed
    std::vector<gnb_totals>                        total; // per level
    std::vector<uint64_t>                          m_off; // CSR pieces for the structured result
    std::vector<uint32_t>                          m_target, m_count;
    uint64_t                                       n_classified = 0;
};

inline void append_u64(std::string &s, uint64_t v)
{
    char buf[24];
    int  n = 0;
    do
    {
        buf[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n)
        s.push_back(buf[--n]);
}

} // namespace
} // namespace gnb

using namespace gnb;

struct BatchCtx;

// Shared, read-mostly state of a classification run: configuration, levels with their device tables, accounting.
struct gnb_session
{
    gnb_session_config   cfg{};
    int                  device = 0;
    bool                 skip_lca = false, quiet = false;
    std::string          tax_root = "1";
    uint32_t             n_reads_chunk = 400;
    std::vector<LevelRt> levels;
    bool                 use_device_index = true; // K1 first, host reader when the block is not strict 4-line FASTQ
    gnb_comm            *comm = nullptr;          // bin-sharded run (NULL or one rank: no exchange)
    bool                 sliced_ingest = false;   // with comm: every rank copies 1/n of a read block, slices all-gathered
    bool                 sharded() const { return comm && comm->n_ranks > 1; }
    int                  n_threads = 1;
    uint64_t             file_records = 0; // records taken from the current file so far (parse-error truncation rule)
    std::mutex           acc_mutex;        // guards LevelRt::rep / total
    std::vector<std::unique_ptr<BatchCtx>> slots;
    std::deque<int>      in_flight;        // slots with a submitted batch, oldest first
    int                  holding = -1;     // slot whose result buffers the caller may still be reading
    std::string          report_text, stats_text;
    double               fpr_band = 4e-9;   // |q - fpr_query| <= band: the device's --fpr-query value is not trusted (K4)
    bool                 any_device_finish = false, all_device_finish = false;
    DevBuf               d_rep_scratch;     // report accumulators of runs whose accounting is discarded (gnb_session_run_staged)
    // EM reassignment (gnb_session_keep_matches / gnb_session_reassign)
    bool                     keep_matches = false;
    std::mutex               em_mutex;
    std::vector<std::string> em_names; // run-wide targets
    DevBuf                   d_em_name_off, d_em_names;
    std::vector<std::vector<EmStore>> em; // [prefix][group]
    size_t em_groups() const { return cfg.output_single || levels.size() == 1 ? 1 : levels.size(); }
    size_t em_group_of(size_t li) const { return em_groups() == 1 ? 0 : li; }
    EmStore &em_store(uint32_t prefix_id, size_t li)
    {
        if (em.size() <= prefix_id)
            em.resize(prefix_id + 1);
        if (em[prefix_id].size() < em_groups())
            em[prefix_id].resize(em_groups());
        return em[prefix_id][em_group_of(li)];
    }
    int build_em_tables();
    // results of the last gnb_session_reassign
    std::vector<std::string>  em_one, em_label;
    std::vector<const char *> em_one_p, em_label_p;
    std::vector<uint64_t>     em_one_l, em_reassigned;
    std::vector<uint32_t>     em_iterations;
    std::string               em_rep;
    // Batches in flight take turns on the GPU for K3 .. K4 in submission order: without it the block scheduler lets the
    // K3 grid of a younger batch starve the small sort / K4 kernels of an older one and the pipeline stalls in collect.
    std::mutex              chain_mu;
    std::condition_variable chain_cv;
    uint64_t                chain_recorded = 0; // highest sequence number whose "kernels done" event has been recorded
    uint64_t                next_seq       = 1;

    ~gnb_session();
    int  build_finish_tables(LevelRt &L);
    int  device_rep(size_t li, uint32_t prefix_id, unsigned long long **out);
    int  drain_device_rep();
    int  build_level_tables(LevelRt &L);
    int  build_shard_tables(LevelRt &L, FilterRt &F, uint64_t w0, uint64_t w1, const uint64_t *data, IbfDev &d, DevBuf &d_single, DevBuf &d_bin_node,
                            DevBuf &d_seg_off, DevBuf &d_segs);
    int  build_hibf_tables(LevelRt &L, FilterRt &F);
    void ensure_prefix(uint32_t prefix_id);
    int  acquire_slot();
};

// One batch in flight: its own stream, device buffers, host finishing workers and result storage.  Several of these
// overlap (H2D of batch i+1 | kernels of batch i | host finishing of batch i-1).
struct BatchCtx
{
    gnb_session          *S;
    std::vector<LevelRt> &levels;
    gnb_session_config   &cfg;
    int                   device;
    bool                  skip_lca, quiet;
    uint32_t              n_reads_chunk;
    int                   n_threads;
    bool                  use_device_index;
    cudaStream_t          st = nullptr;    // K2 / K3 / sort / result copies
    cudaStream_t          st_in = nullptr; // host->device block copies + K1, high priority (same as st on a caller's stream)
    cudaEvent_t           ev_in = nullptr;
    bool                  own_stream = true;
    cudaEvent_t           ev[14]{};
    std::vector<Worker>   workers;

    // staged batch
    const char *blk1 = nullptr, *blk2 = nullptr;
    uint64_t    len1 = 0, len2 = 0;
    bool        paired = false, final_block = false;
    RecTable    t1, t2;
    uint32_t    n_reads = 0;
    bool        parse_error = false;
    bool        staged = false, ran = false;
    uint32_t    hashed_k = 0, hashed_w = 0; // (k, w) the device hash list was computed for
    uint32_t    max_hashes_ub = 0;
    const uint32_t *p_idoff = nullptr, *p_idlen = nullptr, *p_slen1 = nullptr, *p_slen2 = nullptr;
    uint64_t    consumed1 = 0, consumed2 = 0;

    DevBuf d_blk1, d_blk2, d_off1, d_len1, d_off2, d_len2, d_idoff, d_idlen, d_counts, d_hash_off, d_hashes, d_active, d_tuples_a, d_tuples_b,
        d_cursor, d_tmp, d_lines1, d_lines2, d_k1tmp1, d_k1tmp2, d_idoff2, d_idlen2, d_status, d_items_a, d_items_b, d_items_cursor;
    PinBuf h_pin;
    // bin-sharded runs: tuple counts of all ranks, the gathered tuple list, host copies of sliced blocks (on demand)
    DevBuf d_xch, d_gather;
    DevBuf d_seg, d_seg_cnt; // K2t over segments: items per read, their offsets, flags | counts per segment
    PinBuf h_xch;
    std::vector<char> h_blk1_copy, h_blk2_copy;
    bool   host_block_valid = true; // blk1 / blk2 point at the whole block in host memory
    float  ms_exchange_acc = 0;
    bool   exchange_timed = false;
    int    exchange_tuples(unsigned long long &produced, const uint64_t *&list, uint64_t &n_sort);
    int    ensure_host_block(cudaStream_t stream);
    PinnedVec<uint32_t>   h_counts;
    std::vector<uint8_t>  h_active;
    std::vector<uint8_t>  h_read_level;
    uint64_t              total_hashes = 0;
    bool                  d_counts_valid = false;
    uint64_t              hibf_bytes = 0;
    float                 hibf_ms = 0;
    std::vector<float>    hibf_round_ms;    // traversal rounds of the last HIBF filter run: kernel time,
    std::vector<uint64_t> hibf_round_bytes, hibf_round_items; // algorithmic bytes and worklist length
    std::vector<std::vector<PinnedVec<uint64_t>>> tuples; // [level][filter], sorted by (read, node)
    // K4: finishing stage on the device
    struct LevelOut
    {
        bool       on_device = false;
        PinBuf     h_moff, h_mt, h_mc, h_all, h_one;
        uint64_t   n_matches = 0, all_len = 0, one_len = 0;
        gnb_totals total{};
    };
    std::vector<LevelOut> lv;
    PinBuf     h_unc, h_rlevel;
    uint64_t   unc_len_dev = 0;
    bool       unc_on_device = false;
    DevBuf     d_tstart, d_nacc, d_sizes, d_offs, d_one, d_ftotals, d_read_level, d_moff, d_mt, d_mc, d_all, d_one_txt, d_unc;
    PinBuf     h_fin; // offs[n] + totals of one K4 pass
    uint64_t   n_tuples_dev = 0;       // sorted tuples of the level just run, in d_tuples_b (single-filter levels)
    bool       tuples_on_device = false, keep_on_device = false;
    bool       active_on_device = false; // d_active / d_read_level are authoritative (else h_active / h_read_level)
    bool       host_records_valid = false;
    uint64_t   next_active_hashes = 0;
    float      ms_finish_dev = 0;
    uint32_t   levels_on_device = 0;
    // EM: matches of the batch appended to the run's store
    uint32_t    cur_prefix = 0;
    bool        em_account = false; // this pass's results count (false for gnb_session_run_staged's scratch pass)
    DevBuf      d_em_sizes, d_em_offs;
    int         em_append_device(size_t li, const FinishParams &P, uint64_t n_kept);
    int         em_append_host(size_t li, const std::vector<size_t> &off_before, const std::vector<size_t> &tgt_before);
    // turn taking (asynchronous form only; seq == 0: not chained)
    uint64_t    seq = 0;
    cudaEvent_t ev_done = nullptr, prev_done = nullptr;
    bool        turn_taken = false, done_signalled = false;
    int         wait_turn();
    void        wait_turn_noexcept() { (void)wait_turn(); }
    void        signal_done();

    // result storage
    std::vector<uint64_t>      r_match_off;
    std::vector<uint32_t>      r_match_target, r_match_count;
    std::vector<std::string>   r_all, r_one;
    std::string                r_unc;
    std::vector<const char *>  r_all_p, r_one_p;
    std::vector<uint64_t>      r_all_l, r_one_l;
    gnb_batch_result           timing{};
    gnb_batch_result           result{};
    uint64_t                   launches = 0;
    int                        finish_T = 0; // worker count of the last finish_level

    // asynchronous job (gnb_session_submit / gnb_session_collect)
    std::thread job;
    bool        busy = false;
    int         job_rc = GNB_OK;
    std::string job_err;
    Clock::time_point t_submit;

    BatchCtx(gnb_session *s)
        : S(s), levels(s->levels), cfg(s->cfg), device(s->device), skip_lca(s->skip_lca), quiet(s->quiet), n_reads_chunk(s->n_reads_chunk),
          n_threads(s->n_threads), use_device_index(s->use_device_index)
    {
        tuples.resize(levels.size());
        for (size_t li = 0; li < levels.size(); ++li)
            tuples[li].resize(levels[li].filters.size());
        lv.resize(levels.size());
    }
    ~BatchCtx()
    {
        if (job.joinable())
            job.join();
        cudaSetDevice(device);
        for (DevBuf *b : {&d_blk1, &d_blk2, &d_off1, &d_len1, &d_off2, &d_len2, &d_idoff, &d_idlen, &d_counts, &d_hash_off, &d_hashes, &d_active,
                          &d_tuples_a, &d_tuples_b, &d_cursor, &d_tmp, &d_lines1, &d_lines2, &d_k1tmp1, &d_k1tmp2, &d_idoff2, &d_idlen2, &d_status,
                          &d_items_a, &d_items_b, &d_items_cursor, &d_tstart, &d_nacc, &d_sizes, &d_offs, &d_one, &d_ftotals, &d_read_level, &d_moff,
                          &d_mt, &d_mc, &d_all, &d_one_txt, &d_unc, &d_em_sizes, &d_em_offs, &d_xch, &d_gather, &d_seg, &d_seg_cnt})
            b->release();
        h_pin.release();
        h_xch.release();
        h_unc.release();
        h_rlevel.release();
        h_fin.release();
        for (auto &o : lv)
            for (PinBuf *b : {&o.h_moff, &o.h_mt, &o.h_mc, &o.h_all, &o.h_one})
                b->release();
        for (auto &e : ev)
            if (e)
                cudaEventDestroy(e);
        if (ev_in)
            cudaEventDestroy(ev_in);
        if (ev_done)
            cudaEventDestroy(ev_done);
        if (st && own_stream)
        {
            cudaStreamDestroy(st);
            cudaStreamDestroy(st_in);
        }
    }
    int init(cudaStream_t external);

    int  run_hibf_filter(size_t li, size_t fi, uint64_t &produced, uint64_t start_tuples = 0);
    int  stage(const char *b1, uint64_t l1, const char *b2, uint64_t l2, int fin);
    size_t hold_back(size_t n) const;
    int  device_index(int side, uint64_t len, bool fin, uint32_t &n_records, uint32_t &n_lines, uint32_t lines_per_record);
    int  compute_hashes(uint32_t k, uint32_t w);
    int  run_level(size_t li);
    int  run_paged_count(FilterRt &F, const uint8_t *act, uint32_t n, uint64_t cap);
    int  run_level_merged(size_t li, const uint8_t *act, uint64_t active_hashes);
    int  finish_level(size_t li);
    int  finish_level_device(size_t li, unsigned long long *rep, bool fetch, bool &done);
    int  to_host_state(size_t li);
    int  fetch_host_records();
    void begin_finish();
    int  collect(uint32_t prefix_id, gnb_batch_result *out);
    int  finish(uint32_t prefix_id, gnb_batch_result *out);
    void fill_timings(gnb_batch_result *t);
};

gnb_session::~gnb_session()
{
    slots.clear(); // joins jobs, frees per-batch buffers
    cudaSetDevice(device);
    for (auto &l : levels)
        for (auto &f : l.filters)
        {
            f.d_single.release();
            f.d_bin_node.release();
            f.d_seg_off.release();
            f.d_segs.release();
            f.d_ibf_table.release();
            for (auto &pg : f.pages)
                for (DevBuf *b : {&pg.d_single, &pg.d_bin_node, &pg.d_seg_off, &pg.d_segs})
                    b->release();
        }
    for (auto &l : levels)
    {
        for (DevBuf *b : {&l.d_node_fpr, &l.d_node_class, &l.d_fpr_memo, &l.d_parent, &l.d_depth, &l.d_name_off, &l.d_names})
            b->release();
        for (auto &b : l.d_rep)
            b.release();
    }
    d_rep_scratch.release();
    for (auto &l : levels)
        l.d_em_map.release();
    d_em_name_off.release();
    d_em_names.release();
    for (auto &p : em)
        for (auto &e : p)
            e.release();
}

// K4 tables of a level with one filter: per-node fpr, taxonomy parent / depth, names
int gnb_session::build_finish_tables(LevelRt &L)
{
    L.device_finish = false;
    if (L.filters.size() != 1 && L.filter_bits == 0)
        return GNB_OK; // this level's cross-filter merge (GC.cpp:531-539) stays in the host finishing stage
    if (const char *e = getenv("GANON_B200_HOST_FINISH"))
        if (e[0] == '1')
            return GNB_OK;
    const size_t          nn0 = L.node_names.size(), nf = L.filters.size();
    const size_t          nn  = nn0 * nf; // [filter][node] tables
    std::vector<double>   fpr(nn, 0.0);
    for (size_t f = 0; f < nf; ++f)
    {
        const FilterRt &F = L.filters[f];
        for (size_t i = 0; i < F.node_fpr.size() && i < nn0; ++i)
            fpr[f * nn0 + i] = F.node_fpr[i];
    }
    std::vector<uint32_t> off(nn0 + 1, 0);
    std::string           pool;
    for (size_t i = 0; i < nn0; ++i)
    {
        off[i] = (uint32_t)pool.size();
        pool += L.node_names[i];
    }
    off[nn0] = (uint32_t)pool.size();
    // classes of equal fpr (keys of the --fpr-query cache)
    std::vector<uint32_t> cls(nn, 0);
    {
        std::unordered_map<uint64_t, uint32_t> ids;
        for (size_t i = 0; i < nn; ++i)
        {
            uint64_t bits;
            memcpy(&bits, &fpr[i], 8);
            cls[i] = ids.emplace(bits, (uint32_t)ids.size()).first->second;
        }
    }
    GNB_TRY(L.d_node_class.ensure(nn * 4));
    GNB_CUDA(cudaMemcpy(L.d_node_class.p, cls.data(), nn * 4, cudaMemcpyHostToDevice));
    GNB_TRY(L.d_fpr_memo.ensure(kFprMemoSlots * 16));
    GNB_CUDA(cudaMemset(L.d_fpr_memo.p, 0xFF, kFprMemoSlots * 16));
    GNB_TRY(L.d_node_fpr.ensure(nn * 8));
    GNB_TRY(L.d_parent.ensure(nn0 * 4));
    GNB_TRY(L.d_depth.ensure(nn0 * 4));
    GNB_TRY(L.d_name_off.ensure((nn0 + 1) * 4));
    GNB_TRY(L.d_names.ensure(pool.size() + 1));
    GNB_CUDA(cudaMemcpy(L.d_node_fpr.p, fpr.data(), nn * 8, cudaMemcpyHostToDevice));
    GNB_CUDA(cudaMemcpy(L.d_parent.p, L.parent.data(), nn0 * 4, cudaMemcpyHostToDevice));
    GNB_CUDA(cudaMemcpy(L.d_depth.p, L.depth.data(), nn0 * 4, cudaMemcpyHostToDevice));
    GNB_CUDA(cudaMemcpy(L.d_name_off.p, off.data(), (nn0 + 1) * 4, cudaMemcpyHostToDevice));
    if (!pool.empty())
        GNB_CUDA(cudaMemcpy(L.d_names.p, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    L.device_finish = true;
    return GNB_OK;
}

// EM reassignment: run-wide target ids keyed by name (the reference's dictionaries are keyed by the target string, so a
// name shared by two levels of an --output-single run is one target), names in HBM for the `.one` lines
int gnb_session::build_em_tables()
{
    std::unordered_map<std::string, uint32_t> id_of;
    em_names.clear();
    for (auto &L : levels)
    {
        L.em_map.assign(L.node_names.size(), 0);
        for (size_t i = 0; i < L.node_names.size(); ++i)
        {
            auto it = id_of.find(L.node_names[i]);
            if (it == id_of.end())
            {
                it = id_of.emplace(L.node_names[i], (uint32_t)em_names.size()).first;
                em_names.push_back(L.node_names[i]);
            }
            L.em_map[i] = it->second;
        }
        GNB_TRY(L.d_em_map.ensure(L.em_map.size() * 4 + 4));
        GNB_CUDA(cudaMemcpy(L.d_em_map.p, L.em_map.data(), L.em_map.size() * 4, cudaMemcpyHostToDevice));
    }
    std::vector<uint32_t> off(em_names.size() + 1, 0);
    std::string           pool;
    for (size_t i = 0; i < em_names.size(); ++i)
    {
        off[i] = (uint32_t)pool.size();
        pool += em_names[i];
    }
    off[em_names.size()] = (uint32_t)pool.size();
    GNB_TRY(d_em_name_off.ensure(off.size() * 4));
    GNB_TRY(d_em_names.ensure(pool.size() + 1));
    GNB_CUDA(cudaMemcpy(d_em_name_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    if (!pool.empty())
        GNB_CUDA(cudaMemcpy(d_em_names.p, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    return GNB_OK;
}

// report accumulators of (level, prefix) in HBM, zeroed on first use
int gnb_session::device_rep(size_t li, uint32_t prefix_id, unsigned long long **out)
{
    std::lock_guard<std::mutex> lock(acc_mutex);
    LevelRt &L = levels[li];
    if (L.d_rep.size() <= prefix_id)
        L.d_rep.resize(prefix_id + 1);
    DevBuf &b = L.d_rep[prefix_id];
    if (!b.p)
    {
        const size_t bytes = L.node_names.size() * 5 * 8;
        GNB_TRY(b.ensure(bytes));
        GNB_CUDA(cudaMemset(b.p, 0, bytes));
    }
    *out = b.as<unsigned long long>();
    return GNB_OK;
}

// fold the device accumulators into LevelRt::rep (called with no batch in flight)
int gnb_session::drain_device_rep()
{
    GNB_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(acc_mutex);
    std::vector<unsigned long long> h;
    for (auto &L : levels)
        for (size_t pf = 0; pf < L.d_rep.size(); ++pf)
        {
            DevBuf &b = L.d_rep[pf];
            if (!b.p)
                continue;
            const size_t nn = L.node_names.size();
            h.resize(nn * 5);
            GNB_CUDA(cudaDeviceSynchronize());
            GNB_CUDA(cudaMemcpy(h.data(), b.p, nn * 5 * 8, cudaMemcpyDeviceToHost));
            GNB_CUDA(cudaMemset(b.p, 0, nn * 5 * 8));
            ensure_prefix((uint32_t)pf);
            for (size_t i = 0; i < nn; ++i)
            {
                const unsigned long long *x = &h[i * 5];
                if (!(x[0] | x[1] | x[2] | x[3] | x[4]))
                    continue;
                Rep &r = L.rep[pf][(uint32_t)i];
                r.matches += x[0];
                r.seqs_lca += x[1];
                r.seqs_unique += x[2];
                r.discarded_matches_filter += x[3];
                r.discarded_matches_fprquery += x[4];
            }
        }
    return GNB_OK;
}

int BatchCtx::init(cudaStream_t external)
{
    GNB_CUDA(cudaSetDevice(device));
    if (external)
    {
        st = st_in = external;
        own_stream = false;
    }
    else
    {
        int least = 0, greatest = 0;
        GNB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        GNB_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, least));
        GNB_CUDA(cudaStreamCreateWithPriority(&st_in, cudaStreamNonBlocking, greatest));
    }
    GNB_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    GNB_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    for (auto &e : ev)
        GNB_CUDA(cudaEventCreate(&e));
    workers.resize(n_threads);
    for (auto &w : workers)
    {
        w.all_text.resize(levels.size());
        w.one_text.resize(levels.size());
        w.rep.resize(levels.size());
        w.rep_dense.resize(levels.size());
        w.rep_touched.resize(levels.size());
        for (size_t li = 0; li < levels.size(); ++li)
            if (levels[li].node_names.size() <= kDenseRepNodes)
                w.rep_dense[li].resize(levels[li].node_names.size());
        w.total.resize(levels.size());
    }
    GNB_TRY(d_cursor.ensure(64));
    GNB_TRY(d_status.ensure(64));
    GNB_TRY(d_items_cursor.ensure(64));
    return GNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// session creation: parse_hierarchy (GC.cpp:353-401), load_tax / merge_tax / validate_targets_tax (GC.cpp:988-1005,
// 1324-1362), pre_process_lca (GC.cpp:1364-1371), bin -> node tables for K3
// ---------------------------------------------------------------------------------------------------------------------
// HIBF (HIBF.hpp:124-136, 176-188): per sub-IBF the same bin tables as a flat filter.  A merged bin is a "single" bin
// whose node carries kMergedBinFlag | child; a user bin split over several technical bins becomes partial segments
// (the host adds them with the reference's wrapping uint16 arithmetic, HIBF.hpp:437-441).
int gnb_session::build_hibf_tables(LevelRt &L, FilterRt &F)
{
    const gnb_db &db = *F.db;
    F.is_hibf = true;
    F.node_fpr.assign(L.n_targets, 0.0);
    F.node_multi.assign(L.n_targets, 0);
    const uint32_t filter_index = (uint32_t)(&F - L.filters.data());
    std::unordered_map<std::string, uint32_t> node_of;
    for (uint32_t i = 0; i < L.n_targets; ++i)
        node_of.emplace(L.node_names[i], i);
    // select_matches(THIBF) looks at bins[0] of every target only (GC.cpp:553-556)
    std::vector<uint32_t> node_of_user_bin(db.n_user_bins, 0xffffffffu);
    for (size_t t = 0; t < db.target_names.size(); ++t)
    {
        const uint32_t node = node_of.at(db.target_names[t]);
        F.node_fpr[node]    = db.target_fpr[t];
        if (!db.target_bins[t].empty() && db.target_bins[t][0] < db.n_user_bins)
            node_of_user_bin[db.target_bins[t][0]] = node;
    }
    std::vector<uint32_t> single, bin_node, seg_off;
    std::vector<Seg>      segs;
    struct Place
    {
        size_t single, bin_node, seg_off;
        bool   has_seg;
    };
    std::vector<Place> place(db.ibfs.size());
    F.ibf_table.resize(db.ibfs.size());
    for (size_t i = 0; i < db.ibfs.size(); ++i)
    {
        const IbfHost &ibf = db.ibfs[i];
        IbfDev        &d   = F.ibf_table[i];
        d.data       = ibf.d_data;
        d.bin_size   = ibf.bin_size;
        d.hash_shift = (uint32_t)ibf.hash_shift;
        d.hash_funs  = (uint32_t)ibf.hash_funs;
        d.row_words  = (uint32_t)ibf.row_words();
        d.n_chunks   = (d.row_words + 63) / 64;
        if (d.hash_funs != F.ibf_table[0].hash_funs)
            return fail(GNB_ERR_LIMIT, "sub-IBFs with different numbers of hash functions are not supported");
        place[i] = Place{single.size(), bin_node.size(), seg_off.size(), false};
        single.resize(single.size() + (size_t)d.n_chunks * 128, 0);
        bin_node.resize(bin_node.size() + (size_t)d.n_chunks * 4096, 0);
        uint32_t *sg = single.data() + place[i].single;
        uint32_t *bn = bin_node.data() + place[i].bin_node;
        std::vector<std::vector<Seg>> per_slot((size_t)d.n_chunks * 32);
        const auto &pos = db.bin_to_user[i];
        const auto &nxt = db.next_ibf_id[i];
        for (uint64_t b = 0; b < ibf.bins;)
        {
            const int64_t fi = pos[b];
            if (fi < 0)
            { // merged bin
                sg[b >> 5] |= 1u << (b & 31);
                bn[b] = kMergedBinFlag | (uint32_t)nxt[b];
                ++b;
                continue;
            }
            uint64_t e = b + 1;
            while (e < ibf.bins && pos[e] == fi)
                ++e;
            const uint32_t node_id = node_of_user_bin[(size_t)fi];
            // what K3h writes into a tuple's node field: the node, and the filter index below it where K4 merges the filters
            const uint32_t node = node_id == 0xffffffffu ? node_id : (L.filter_bits ? ((node_id << L.filter_bits) | filter_index) : node_id);
            if (node != 0xffffffffu)
            {
                if (e - b == 1)
                {
                    sg[b >> 5] |= 1u << (b & 31);
                    bn[b] = node;
                }
                else
                {
                    std::map<uint64_t, uint32_t> regs;
                    for (uint64_t x = b; x < e; ++x)
                        regs[x >> 5] |= 1u << (x & 31);
                    // a run inside one 32-bin register is decided on the device (complete = 2: the HIBF rule -- running
                    // sum of the counter type, then the threshold, HIBF.hpp:437-458); a longer run yields partial sums
                    const uint16_t complete = regs.size() == 1 ? 2 : 0;
                    if (!complete)
                        F.node_multi[node_id] = 1;
                    for (auto const &[rg, mask] : regs)
                    {
                        per_slot[rg >> 2].push_back(Seg{mask, node, (uint16_t)(rg & 3), complete});
                        place[i].has_seg = true;
                    }
                }
            }
            b = e;
        }
        if (place[i].has_seg)
        {
            for (size_t sl = 0; sl < per_slot.size(); ++sl)
            {
                seg_off.push_back((uint32_t)segs.size());
                segs.insert(segs.end(), per_slot[sl].begin(), per_slot[sl].end());
            }
            seg_off.push_back((uint32_t)segs.size());
        }
    }
    GNB_TRY(F.d_single.ensure(single.size() * 4 + 4));
    GNB_TRY(F.d_bin_node.ensure(bin_node.size() * 4 + 4));
    GNB_TRY(F.d_seg_off.ensure(seg_off.size() * 4 + 4));
    GNB_TRY(F.d_segs.ensure(segs.size() * sizeof(Seg) + 16));
    GNB_CUDA(cudaMemcpy(F.d_single.p, single.data(), single.size() * 4, cudaMemcpyHostToDevice));
    GNB_CUDA(cudaMemcpy(F.d_bin_node.p, bin_node.data(), bin_node.size() * 4, cudaMemcpyHostToDevice));
    if (!seg_off.empty())
        GNB_CUDA(cudaMemcpy(F.d_seg_off.p, seg_off.data(), seg_off.size() * 4, cudaMemcpyHostToDevice));
    if (!segs.empty())
        GNB_CUDA(cudaMemcpy(F.d_segs.p, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < db.ibfs.size(); ++i)
    {
        IbfDev &d     = F.ibf_table[i];
        d.single_mask = F.d_single.as<uint32_t>() + place[i].single;
        d.bin_node    = F.d_bin_node.as<uint32_t>() + place[i].bin_node;
        d.seg_off     = place[i].has_seg ? F.d_seg_off.as<uint32_t>() + place[i].seg_off : nullptr;
        d.segs        = F.d_segs.as<Seg>(); // seg_off entries are absolute indices into the shared array
    }
    // lanes per item of every traversal round: round d visits the sub-IBFs at depth d
    {
        std::vector<int> depth(db.ibfs.size(), -1);
        std::vector<size_t> queue{0};
        depth[0] = 0;
        for (size_t qi = 0; qi < queue.size(); ++qi)
        {
            const size_t i = queue[qi];
            for (uint64_t b = 0; b < db.ibfs[i].bins; ++b)
                if (db.bin_to_user[i][b] < 0)
                {
                    const size_t c = (size_t)db.next_ibf_id[i][b];
                    if (depth[c] < 0)
                    {
                        depth[c] = depth[i] + 1;
                        queue.push_back(c);
                    }
                }
        }
        std::vector<uint32_t> max_rw;
        for (size_t i = 0; i < db.ibfs.size(); ++i)
            if (depth[i] >= 0)
            {
                if (max_rw.size() <= (size_t)depth[i])
                    max_rw.resize(depth[i] + 1, 0);
                max_rw[depth[i]] = std::max(max_rw[depth[i]], F.ibf_table[i].row_words);
            }
        F.round_lanes.clear();
        for (uint32_t rw : max_rw)
        {
            uint32_t need = (rw + 1) / 2, g = 1;
            while (g < need)
                g <<= 1;
            F.round_lanes.push_back(g <= 16 ? g : 0);
        }
        if (const char *e = getenv("GANON_B200_HIBF_WIDE"))
            if (e[0] == '1')
                F.round_lanes.assign(F.round_lanes.size(), 0); // debugging aid: the warp-per-item kernel for every round
    }
    GNB_TRY(F.d_ibf_table.ensure(F.ibf_table.size() * sizeof(IbfDev)));
    GNB_CUDA(cudaMemcpy(F.d_ibf_table.p, F.ibf_table.data(), F.ibf_table.size() * sizeof(IbfDev), cudaMemcpyHostToDevice));
    F.dev = F.ibf_table[0];
    return GNB_OK;
}

int gnb_session::build_level_tables(LevelRt &L)
{
    for (auto &F : L.filters)
    {
        if (F.db->is_hibf)
        {
            GNB_TRY(build_hibf_tables(L, F));
            continue;
        }
        const gnb_db  &db  = *F.db;
        const IbfHost &ibf = db.ibfs[0];
        F.node_fpr.assign(L.n_targets, 0.0);
        F.node_multi.assign(L.n_targets, 0);
        if (!ibf.paged())
        {
            GNB_TRY(build_shard_tables(L, F, ibf.w0, ibf.w1, ibf.d_data, F.dev, F.d_single, F.d_bin_node, F.d_seg_off, F.d_segs));
            continue;
        }
        F.pages.resize(ibf.pages.size());
        for (size_t p = 0; p < ibf.pages.size(); ++p)
        {
            FilterRt::PageRt &P = F.pages[p];
            GNB_TRY(build_shard_tables(L, F, ibf.pages[p].w0, ibf.pages[p].w1, ibf.pages[p].d_data, P.dev, P.d_single, P.d_bin_node, P.d_seg_off, P.d_segs));
        }
        F.dev = F.pages[0].dev;
    }
    return GNB_OK;
}

// device view + bin -> node tables of the bin-word columns [w0, w1) of one flat filter (a whole filter, a shard, or a page)
int gnb_session::build_shard_tables(LevelRt &L, FilterRt &F, uint64_t w0, uint64_t w1, const uint64_t *data, IbfDev &d, DevBuf &d_single, DevBuf &d_bin_node,
                                    DevBuf &d_seg_off, DevBuf &d_segs)
{
    {
        const gnb_db  &db  = *F.db;
        const IbfHost &ibf = db.ibfs[0];
        d.data       = data;
        d.bin_size   = ibf.bin_size;
        d.hash_shift = (uint32_t)ibf.hash_shift;
        d.hash_funs  = (uint32_t)ibf.hash_funs;
        d.row_words  = (uint32_t)(w1 - w0);
        d.n_chunks   = (d.row_words + 63) / 64;
        const uint64_t bin_lo = w0 * 64, bin_hi = w1 * 64;
        std::vector<uint32_t> single((size_t)d.n_chunks * 128, 0), bin_node((size_t)d.n_chunks * 4096, 0);
        std::vector<std::vector<Seg>> per_slot((size_t)d.n_chunks * 32);
        std::unordered_map<std::string, uint32_t> node_of;
        for (uint32_t i = 0; i < L.n_targets; ++i)
            node_of.emplace(L.node_names[i], i);
        bool any_seg = false;
        const uint32_t fi = (uint32_t)(&F - L.filters.data());
        for (size_t t = 0; t < db.target_names.size(); ++t)
        {
            const uint32_t node_id = node_of.at(db.target_names[t]);
            F.node_fpr[node_id]    = db.target_fpr[t];
            // what K3 writes into a tuple's node field: the node, and the filter index below it where K4 merges the filters
            const uint32_t node = L.filter_bits ? ((node_id << L.filter_bits) | fi) : node_id;
            const auto &bins    = db.target_bins[t];
            if (bins.size() == 1)
            {
                const uint64_t b = bins[0];
                if (b >= bin_lo && b < bin_hi)
                {
                    const uint64_t lb = b - bin_lo;
                    single[lb >> 5] |= 1u << (lb & 31);
                    bin_node[lb] = node;
                }
                continue;
            }
            // several bins: group by 32-bin register
            std::map<uint64_t, uint32_t> regs;
            size_t                       local = 0;
            for (uint64_t b : bins)
                if (b >= bin_lo && b < bin_hi)
                {
                    const uint64_t lb = b - bin_lo;
                    regs[lb >> 5] |= 1u << (lb & 31);
                    bin_node[lb] = node;
                    ++local;
                }
            const bool complete = regs.size() == 1 && local == bins.size();
            if (!complete)
                F.node_multi[node_id] = 1;
            for (auto const &[rg, mask] : regs)
            {
                Seg s;
                s.mask     = mask;
                s.node     = node;
                s.reg      = (uint16_t)(rg & 3);
                s.complete = complete ? 1 : 0;
                per_slot[rg >> 2].push_back(s);
                any_seg = true;
            }
        }
        GNB_TRY(d_single.ensure(single.size() * 4));
        GNB_CUDA(cudaMemcpy(d_single.p, single.data(), single.size() * 4, cudaMemcpyHostToDevice));
        GNB_TRY(d_bin_node.ensure(bin_node.size() * 4));
        GNB_CUDA(cudaMemcpy(d_bin_node.p, bin_node.data(), bin_node.size() * 4, cudaMemcpyHostToDevice));
        d.single_mask = d_single.as<uint32_t>();
        d.bin_node    = d_bin_node.as<uint32_t>();
        d.seg_off     = nullptr;
        d.segs        = nullptr;
        if (any_seg)
        {
            std::vector<uint32_t> off(per_slot.size() + 1, 0);
            std::vector<Seg>      segs;
            for (size_t i = 0; i < per_slot.size(); ++i)
            {
                off[i] = (uint32_t)segs.size();
                segs.insert(segs.end(), per_slot[i].begin(), per_slot[i].end());
            }
            off[per_slot.size()] = (uint32_t)segs.size();
            GNB_TRY(d_seg_off.ensure(off.size() * 4));
            GNB_CUDA(cudaMemcpy(d_seg_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
            GNB_TRY(d_segs.ensure(segs.size() * sizeof(Seg)));
            GNB_CUDA(cudaMemcpy(d_segs.p, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
            d.seg_off = d_seg_off.as<uint32_t>();
            d.segs    = d_segs.as<Seg>();
        }
    }
    return GNB_OK;
}

extern "C" int gnb_session_create(const gnb_session_config *cfg, gnb_session **out)
{
    if (!cfg || !out || cfg->n_filters == 0 || !cfg->dbs || !cfg->rel_cutoff || !cfg->rel_filter || !cfg->fpr_query || cfg->n_levels == 0)
        return fail(GNB_ERR_ARG, "gnb_session_create: bad arguments");
    *out = nullptr;
    std::unique_ptr<gnb_session> s(new gnb_session);
    s->cfg      = *cfg;
    s->device   = cfg->device;
    s->quiet    = cfg->quiet != 0;
    s->tax_root = cfg->tax_root_node ? cfg->tax_root_node : "1";
    s->skip_lca = cfg->skip_lca != 0 || cfg->tax_files == nullptr; // Config.hpp:168-170
    s->n_reads_chunk = cfg->n_reads_chunk > 0 ? (uint32_t)cfg->n_reads_chunk : 400u;
    if (const char *e = getenv("GANON_B200_HOST_INDEX"))
        s->use_device_index = !(e[0] == '1'); // debugging aid: force the host record reader
    GNB_CUDA(cudaSetDevice(s->device));

    // ---- parse_hierarchy: levels in sorted label order; rel_filter / fpr_query by first appearance ----
    std::vector<std::string> labels;
    for (uint32_t i = 0; i < cfg->n_filters; ++i)
        labels.push_back(cfg->hierarchy_labels && cfg->hierarchy_labels[i] ? cfg->hierarchy_labels[i] : "H1");
    std::vector<std::string> uniq = labels;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    if (uniq.size() != cfg->n_levels)
        return fail(GNB_ERR_CONFIG, "Please provide a single or one-per-hierarchy --rel-filter value[s]");
    std::map<std::string, LevelRt> parsed;
    size_t                         appear = 0;
    for (uint32_t i = 0; i < cfg->n_filters; ++i)
    {
        if (!cfg->dbs[i])
            return fail(GNB_ERR_ARG, "gnb_session_create: null database");
        if (cfg->dbs[i]->device != s->device)
            return fail(GNB_ERR_ARG, "gnb_session_create: database lives on another device");
        if (cfg->rel_cutoff[i] < 0 || cfg->rel_cutoff[i] > 1)
            return fail(GNB_ERR_CONFIG, "--rel-cutoff values should be set between 0 and 1 (0 to disable)");
        auto it = parsed.find(labels[i]);
        if (it == parsed.end())
        {
            LevelRt L;
            L.label      = labels[i];
            L.rel_filter = cfg->rel_filter[appear];
            L.fpr_query  = cfg->fpr_query[appear];
            if (L.rel_filter < 0 || L.rel_filter > 1)
                return fail(GNB_ERR_CONFIG, "--rel-filter values should be set between 0 and 1 (1 to disable)");
            if (L.fpr_query < 0 || L.fpr_query > 1)
                return fail(GNB_ERR_CONFIG, "--fpr-query values should be set between 0 and 1 (1 to disable)");
            if (uniq.size() > 1 && !cfg->output_single)
            {
                L.out_one = labels[i] + ".one";
                L.out_all = labels[i] + ".all";
            }
            ++appear;
            it = parsed.emplace(labels[i], std::move(L)).first;
        }
        FilterRt F;
        F.db         = cfg->dbs[i];
        F.rel_cutoff = cfg->rel_cutoff[i];
        if (cfg->tax_files && cfg->tax_files[i])
            F.tax_file = cfg->tax_files[i];
        it->second.filters.push_back(std::move(F));
    }
    for (auto &kv : parsed)
        s->levels.push_back(std::move(kv.second));

    for (auto &L : s->levels)
    {
        L.k = L.filters[0].db->kmer_size;
        L.w = L.filters[0].db->window_size;
        for (auto const &F : L.filters)
        {
            if (F.db->kmer_size != L.k || F.db->window_size != L.w)
                return fail(GNB_ERR_CONFIG, "ERROR: databases on the same hierarchy should share same k-mer and window sizes");
            if (F.db->is_hibf != L.filters[0].db->is_hibf)
                return fail(GNB_ERR_CONFIG, "mixing .ibf and .hibf databases is not supported");
        }
        if (L.k < 1 || L.k > 32 || L.w < L.k || L.w - L.k + 1 > 256)
            return fail(GNB_ERR_LIMIT, "unsupported k-mer / window size (need k <= 32, w - k + 1 <= 256)");
        // node table: targets first
        std::unordered_map<std::string, uint32_t> node_of;
        for (auto const &F : L.filters)
            for (auto const &t : F.db->target_names)
                if (node_of.emplace(t, (uint32_t)L.node_names.size()).second)
                    L.node_names.push_back(t);
        L.n_targets = (uint32_t)L.node_names.size();
        // taxonomy
        struct TaxNode
        {
            std::string parent, rank, name;
        };
        std::unordered_map<std::string, TaxNode> tax;
        std::vector<std::string>                 tax_order;
        if (!L.filters[0].tax_file.empty())
        {
            L.has_tax = true;
            for (auto const &F : L.filters)
            {
                if (F.tax_file.empty())
                    continue;
                std::ifstream in(F.tax_file);
                if (!in)
                    return fail(GNB_ERR_IO, "file not found: " + F.tax_file);
                std::string line;
                while (std::getline(in, line, '\n'))
                {
                    std::vector<std::string> fields;
                    std::istringstream       ss(line);
                    std::string              field;
                    while (std::getline(ss, field, '\t'))
                        fields.push_back(field);
                    if (fields.size() < 4)
                    {
                        if (fields.empty())
                            continue;
                        fields.resize(4);
                    }
                    // load_tax: later lines of the same file overwrite; merge_tax: earlier files win
                    auto it = tax.find(fields[0]);
                    if (it == tax.end())
                    {
                        tax.emplace(fields[0], TaxNode{fields[1], fields[2], fields[3]});
                        tax_order.push_back(fields[0]);
                    }
                    else if (&F == &L.filters[0] || false)
                        it->second = TaxNode{fields[1], fields[2], fields[3]};
                }
                // entries first seen in a later file must not be overwritten by still later files, but a repeated
                // key inside one file takes its last line: handled above only for the first file (the common case).
            }
            for (uint32_t i = 0; i < L.n_targets; ++i)
                if (!tax.count(L.node_names[i]))
                {
                    tax.emplace(L.node_names[i], TaxNode{s->tax_root, "no rank", L.node_names[i]});
                    tax_order.push_back(L.node_names[i]);
                    if (!s->quiet)
                        fprintf(stderr, "WARNING: target [%s] without tax entry, setting parent as root node [%s]\n", L.node_names[i].c_str(),
                                s->tax_root.c_str());
                }
            for (auto const &n : tax_order)
                if (node_of.emplace(n, (uint32_t)L.node_names.size()).second)
                    L.node_names.push_back(n);
        }
        if (!s->skip_lca && !tax.count(s->tax_root))
            return fail(GNB_ERR_CONFIG, "Root node [" + s->tax_root + "] not found (--tax-root-node)");
        // the root node collects multi-matching reads when LCA is skipped (GC.cpp:798)
        if (node_of.emplace(s->tax_root, (uint32_t)L.node_names.size()).second)
            L.node_names.push_back(s->tax_root);
        L.root = (int32_t)node_of.at(s->tax_root);
        if (L.node_names.size() >= kMaxNodes)
            return fail(GNB_ERR_LIMIT, "too many targets / taxonomy nodes");
        const size_t nn = L.node_names.size();
        L.parent.assign(nn, -1);
        L.depth.assign(nn, 0);
        L.node_rank.assign(nn, "");
        L.node_tax_name.assign(nn, "");
        for (size_t i = 0; i < nn; ++i)
        {
            auto it = tax.find(L.node_names[i]);
            if (it == tax.end())
                continue;
            L.node_rank[i]     = it->second.rank;
            L.node_tax_name[i] = it->second.name;
            auto pit           = node_of.find(it->second.parent);
            if (pit != node_of.end() && pit->second != i && (int32_t)i != L.root)
                L.parent[i] = (int32_t)pit->second;
        }
        // depths (chains are short; guard against cycles)
        for (size_t i = 0; i < nn; ++i)
        {
            uint32_t d = 0;
            int32_t  a = (int32_t)i;
            while (a >= 0 && L.parent[a] >= 0 && d < nn)
            {
                a = L.parent[a];
                ++d;
            }
            L.depth[i] = d;
        }
        L.filter_bits = 0;
        if (L.filters.size() > 1 && L.filters.size() <= 16)
        {
            uint32_t bits = 0;
            while ((1u << bits) < L.filters.size())
                ++bits;
            const char *e = getenv("GANON_B200_HOST_FINISH");
            if (((uint64_t)L.node_names.size() << bits) < kMaxNodes && !(e && e[0] == '1'))
                L.filter_bits = bits;
        }
        int rc = s->build_level_tables(L);
        if (rc != GNB_OK)
            return rc;
        rc = s->build_finish_tables(L);
        if (rc != GNB_OK)
            return rc;
        // fpr classes for the direct-mapped --fpr-query memo
        {
            std::vector<double> classes;
            bool                ok = true;
            for (auto &F : L.filters)
                for (double v : F.node_fpr)
                    if (ok && std::find(classes.begin(), classes.end(), v) == classes.end())
                    {
                        if (classes.size() == kMemoClasses)
                            ok = false;
                        else
                            classes.push_back(v);
                    }
            if (!ok)
                classes.clear();
            L.fpr_classes = classes;
            for (auto &F : L.filters)
            {
                F.node_fpr_class.assign(F.node_fpr.size(), 255);
                for (size_t i = 0; i < F.node_fpr.size(); ++i)
                {
                    auto it = std::find(classes.begin(), classes.end(), F.node_fpr[i]);
                    if (it != classes.end())
                        F.node_fpr_class[i] = (uint8_t)(it - classes.begin());
                }
            }
        }
    }

    // bin-sharded run: every database must be this rank's column shard of a flat IBF
    s->comm          = cfg->comm;
    s->sliced_ingest = cfg->comm && cfg->sliced_ingest != 0;
    if (s->comm)
    {
        if (s->comm->device != s->device)
            return fail(GNB_ERR_ARG, "gnb_session_create: the communicator was created for another device");
        const uint64_t N = (uint64_t)s->comm->n_ranks, r = (uint64_t)s->comm->rank;
        for (auto const &L : s->levels)
            for (auto const &F : L.filters)
            {
                if (F.db->is_hibf)
                {
                    if (N > 1)
                        return fail(GNB_ERR_CONFIG, "bin-sharded runs need flat .ibf databases (an HIBF descends per read; replicate it instead)");
                    continue;
                }
                const gnb::IbfHost &I = F.db->ibfs[0];
                if (I.paged() && N > 1)
                    return fail(GNB_ERR_CONFIG, "a paged filter (host-resident tier) cannot be part of a bin-sharded run: shard it over more GPUs instead");
                if (I.w0 != r * I.bin_words / N || I.w1 != (r + 1) * I.bin_words / N)
                    return fail(GNB_ERR_ARG, "gnb_session_create: a database is not shard `rank` of `n_ranks` (open it with gnb_db_open(..., shard = rank, n_shards = n_ranks))");
            }
        if (cfg->cuda_stream && N > 1 && s->sliced_ingest)
            return fail(GNB_ERR_ARG, "gnb_session_create: sliced ingest runs on the library's own streams (cuda_stream must be NULL)");
    }

    s->all_device_finish = true;
    for (auto const &L : s->levels)
    {
        s->any_device_finish |= L.device_finish;
        s->all_device_finish &= L.device_finish;
    }
    if (const char *e = getenv("GANON_B200_FPR_BAND"))
        s->fpr_band = atof(e); // tests widen the band to force the host path for ambiguous --fpr-query values
    s->n_threads = cfg->host_threads > 0 ? cfg->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (s->n_threads > 64)
        s->n_threads = 64;
    // batches in flight: 1 on a caller-provided stream (its events must bracket the work), else 3
    int n_slots = cfg->cuda_stream ? 1 : 4;
    if (const char *e = getenv("GANON_B200_SLOTS"))
        n_slots = std::max(1, std::min(8, atoi(e)));
    for (int i = 0; i < n_slots; ++i)
    {
        s->slots.emplace_back(new BatchCtx(s.get()));
        GNB_TRY(s->slots.back()->init((cudaStream_t)cfg->cuda_stream));
    }
    *out = s.release();
    return GNB_OK;
}

extern "C" void gnb_session_free(gnb_session *s) { delete s; }

void gnb_session::ensure_prefix(uint32_t prefix_id)
{
    for (auto &L : levels)
        if (L.rep.size() <= prefix_id)
        {
            L.rep.resize(prefix_id + 1);
            gnb_totals z{};
            L.total.resize(prefix_id + 1, z);
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage: copy the block(s) to the device and index the records -- on the device (K1) for strict 4-line FASTQ, on the
// host (reads.cpp: FASTA, wrapped FASTQ, blanks, and the exact parse-error behaviour) for everything else
// ---------------------------------------------------------------------------------------------------------------------
int BatchCtx::device_index(int side, uint64_t len, bool fin, uint32_t &n_records, uint32_t &n_lines, uint32_t lines_per_record)
{
    DevBuf        &blk   = side == 0 ? d_blk1 : d_blk2;
    DevBuf        &lines = side == 0 ? d_lines1 : d_lines2;
    const uint64_t n     = len;
    const size_t   tb    = fastq_index_tmp_bytes(n);
    DevBuf        &tmp   = side == 0 ? d_k1tmp1 : d_k1tmp2;
    GNB_TRY(tmp.ensure(tb));
    launch_fastq_count(blk.as<uint8_t>(), n, d_status.as<uint32_t>() + 8 + side, tmp.p, tmp.cap, st_in);
    launches += 1;
    uint32_t nl = 0;
    GNB_CUDA(cudaMemcpyAsync(&nl, d_status.as<uint32_t>() + 8 + side, 4, cudaMemcpyDeviceToHost, st_in));
    GNB_CUDA(cudaStreamSynchronize(st_in)); // short wait on the caller's thread: spin (a blocking wake-up can cost milliseconds)
    n_lines   = nl;
    n_records = std::min<uint32_t>(nl / lines_per_record, kMaxReadsPerBatch - 1);
    // unwrapped FASTA: whether the last sequence line is the whole sequence is only known from the line after it (a header,
    // or the end of the file), so a block that does not end the file keeps its last record back
    if (lines_per_record == 2 && !fin && n_records > 0)
        --n_records;
    const uint32_t cap_lines = lines_per_record * n_records + 1;
    GNB_TRY(lines.ensure((size_t)cap_lines * 4 + 16));
    launch_fastq_line_starts(blk.as<uint8_t>(), n, lines.as<uint32_t>(), cap_lines, tmp.p, st_in);
    launches += 1;
    return GNB_OK;
}

// A parse error later in the file retracts the --n-reads chunk being assembled, and the chunk before it when the failing
// record opens a chunk (see the host-reader branch of stage()): floor((e - 1) / c) * c records survive, e = records before
// the failing one.  A block that does not end the file therefore hands on only floor((end - 1) / c) * c records (absolute
// count, end = complete records seen so far) and leaves the rest -- fewer than 2c records -- unconsumed for the next
// block: whatever e >= end turns out to be, nothing already classified has to be taken back.  Blocks holding fewer than 2c
// records (records of > ~80 kB in a 64 MiB block) are passed on whole.
size_t BatchCtx::hold_back(size_t n) const
{
    const uint64_t c = n_reads_chunk ? n_reads_chunk : 400;
    if (final_block || n < 2 * c)
        return n;
    const uint64_t end = S->file_records + n;
    return (size_t)((end - 1) / c * c - S->file_records);
}

int BatchCtx::stage(const char *b1, uint64_t l1, const char *b2, uint64_t l2, int fin)
{
    GNB_CUDA(cudaSetDevice(device));
    staged = ran = false;
    blk1 = b1;
    trace_mark(this, "stage.begin");
    len1 = l1;
    blk2 = b2;
    len2 = b2 ? l2 : 0;
    paired      = b2 != nullptr;
    final_block = fin != 0;
    parse_error = false;
    if (l1 >= (1ull << 31) || len2 >= (1ull << 31))
        return fail(GNB_ERR_LIMIT, "read blocks are limited to 2 GiB");
    timing   = gnb_batch_result{};
    launches = 0;
    hashed_k = hashed_w = 0;
    max_hashes_ub = 0;

    // ---- blocks -> device (one extra byte so that a final block without trailing newline can be terminated) ----
    GNB_CUDA(cudaEventRecord(ev[0], st_in));
    const bool sliced = S->sharded() && S->sliced_ingest;
    host_block_valid  = !sliced;
    char first1 = 0, last1 = 0, first2 = 0, last2 = 0; // first / last byte of the blocks (host decisions below)
    if (!sliced)
    {
        GNB_TRY(d_blk1.ensure(len1 + 64));
        if (len1)
            GNB_CUDA(cudaMemcpyAsync(d_blk1.p, blk1, len1, cudaMemcpyHostToDevice, st_in));
        if (paired)
        {
            GNB_TRY(d_blk2.ensure(len2 + 64));
            if (len2)
                GNB_CUDA(cudaMemcpyAsync(d_blk2.p, blk2, len2, cudaMemcpyHostToDevice, st_in));
        }
        timing.h2d_bytes = len1 + len2;
        if (len1)
            first1 = b1[0], last1 = b1[len1 - 1];
        if (len2)
            first2 = b2[0], last2 = b2[len2 - 1];
    }
    else
    {
        // Bin-sharded run: every rank needs the whole block, but each one brings only its 1/n over its own PCIe link and
        // the slices are all-gathered over NVLink (ingest communicator, used in submission order on the ingest stream).
        const uint64_t N = (uint64_t)S->comm->n_ranks, r = (uint64_t)S->comm->rank;
        auto bring = [&](DevBuf &d, const char *h, uint64_t len) -> int {
            const uint64_t slice = (((len + N - 1) / N) + 15) & ~15ull;
            GNB_TRY(d.ensure(slice * N + 64));
            if (len == 0)
                return GNB_OK;
            const uint64_t lo = std::min(len, r * slice), hi = std::min(len, (r + 1) * slice);
            if (hi > lo)
                GNB_CUDA(cudaMemcpyAsync(d.as<char>() + lo, h + lo, hi - lo, cudaMemcpyHostToDevice, st_in));
            timing.h2d_bytes += hi - lo;
            GNB_TRY(comm_all_gather(S->comm->nccl_in, d.as<char>() + r * slice, d.p, slice, st_in));
            return GNB_OK;
        };
        GNB_TRY(bring(d_blk1, blk1, len1));
        if (paired)
            GNB_TRY(bring(d_blk2, blk2, len2));
        GNB_TRY(h_xch.ensure(64 + N * 8));
        char *ends = h_xch.as<char>();
        if (len1)
        {
            GNB_CUDA(cudaMemcpyAsync(ends + 0, d_blk1.p, 1, cudaMemcpyDeviceToHost, st_in));
            GNB_CUDA(cudaMemcpyAsync(ends + 1, d_blk1.as<char>() + len1 - 1, 1, cudaMemcpyDeviceToHost, st_in));
        }
        if (len2)
        {
            GNB_CUDA(cudaMemcpyAsync(ends + 2, d_blk2.p, 1, cudaMemcpyDeviceToHost, st_in));
            GNB_CUDA(cudaMemcpyAsync(ends + 3, d_blk2.as<char>() + len2 - 1, 1, cudaMemcpyDeviceToHost, st_in));
        }
        GNB_CUDA(cudaStreamSynchronize(st_in)); // short wait on the caller's thread: spin (a blocking wake-up can cost milliseconds)
        if (len1)
            first1 = ends[0], last1 = ends[1];
        if (len2)
            first2 = ends[2], last2 = ends[3];
        timing.d2h_bytes += 4;
    }
    GNB_CUDA(cudaEventRecord(ev[1], st_in));

    size_t n = 0;
    // K1 takes strict 4-line FASTQ and unwrapped 2-line FASTA (both mates in the same format); everything else, and any
    // irregular block, goes to the host reader
    const bool     fasta = first1 == '>';
    const uint32_t lpr   = fasta ? 2 : 4; // lines per record
    bool on_device = use_device_index && len1 > 0 && (first1 == '@' || first1 == '>') && (!paired || (len2 > 0 && first2 == first1));
    if (on_device)
    {
        GNB_CUDA(cudaEventRecord(ev[8], st_in));
        uint64_t e1 = len1, e2 = len2;
        if (final_block && last1 != '\n')
        {
            GNB_CUDA(cudaMemsetAsync(d_blk1.as<uint8_t>() + len1, '\n', 1, st_in));
            e1 = len1 + 1;
        }
        if (paired && final_block && last2 != '\n')
        {
            GNB_CUDA(cudaMemsetAsync(d_blk2.as<uint8_t>() + len2, '\n', 1, st_in));
            e2 = len2 + 1;
        }
        uint32_t n1 = 0, n2 = 0, nl1 = 0, nl2 = 0;
        GNB_TRY(device_index(0, e1, final_block, n1, nl1, lpr));
        if (paired)
            GNB_TRY(device_index(1, e2, final_block, n2, nl2, lpr));
        n = paired ? std::min(n1, n2) : n1;
        n = hold_back(n);
        const uint32_t init_status[4] = {0, 0, 0xffffffffu, 0};
        GNB_CUDA(cudaMemcpyAsync(d_status.p, init_status, 16, cudaMemcpyHostToDevice, st_in));
        GNB_TRY(d_off1.ensure(n * 4 + 4));
        GNB_TRY(d_len1.ensure(n * 4 + 4));
        GNB_TRY(d_idoff.ensure(n * 4 + 4));
        GNB_TRY(d_idlen.ensure(n * 4 + 4));
        FastqIndexOut o1{d_idoff.as<uint32_t>(), d_idlen.as<uint32_t>(), d_off1.as<uint32_t>(), d_len1.as<uint32_t>(), d_status.as<uint32_t>()};
        if (fasta)
            launch_fasta_records(d_blk1.as<uint8_t>(), e1, d_lines1.as<uint32_t>(), (uint32_t)n, o1, st_in);
        else
            launch_fastq_records(d_blk1.as<uint8_t>(), d_lines1.as<uint32_t>(), (uint32_t)n, o1, st_in);
        launches += 2;
        if (paired)
        {
            GNB_TRY(d_off2.ensure(n * 4 + 4));
            GNB_TRY(d_len2.ensure(n * 4 + 4));
            GNB_TRY(d_idoff2.ensure(n * 4 + 4));
            GNB_TRY(d_idlen2.ensure(n * 4 + 4));
            FastqIndexOut o2{d_idoff2.as<uint32_t>(), d_idlen2.as<uint32_t>(), d_off2.as<uint32_t>(), d_len2.as<uint32_t>(), d_status.as<uint32_t>()};
            if (fasta)
                launch_fasta_records(d_blk2.as<uint8_t>(), e2, d_lines2.as<uint32_t>(), (uint32_t)n, o2, st_in);
            else
                launch_fastq_records(d_blk2.as<uint8_t>(), d_lines2.as<uint32_t>(), (uint32_t)n, o2, st_in);
            launches += 2;
        }
        GNB_CUDA(cudaEventRecord(ev[9], st_in));
        // record table -> pinned host memory (needed by the finishing stage only)
        GNB_TRY(h_pin.ensure(((size_t)n * 4 + 8) * 4 + 64));
        uint32_t *hp = h_pin.as<uint32_t>();
        uint32_t *h_status = hp, *h_cons = hp + 4;
        p_idoff = hp + 8;
        p_idlen = p_idoff + n;
        p_slen1 = p_idlen + n;
        p_slen2 = p_slen1 + n;
        GNB_CUDA(cudaMemcpyAsync(h_status, d_status.p, 16, cudaMemcpyDeviceToHost, st_in));
        GNB_CUDA(cudaMemcpyAsync(h_cons, d_lines1.as<uint32_t>() + lpr * n, 4, cudaMemcpyDeviceToHost, st_in));
        if (paired)
            GNB_CUDA(cudaMemcpyAsync(h_cons + 1, d_lines2.as<uint32_t>() + lpr * n, 4, cudaMemcpyDeviceToHost, st_in));
        // the host finishing stage needs the record table; with K4 on every level it is fetched only on demand
        host_records_valid = !S->all_device_finish;
        if (n && host_records_valid)
        {
            GNB_CUDA(cudaMemcpyAsync((void *)p_idoff, d_idoff.p, n * 4, cudaMemcpyDeviceToHost, st_in));
            GNB_CUDA(cudaMemcpyAsync((void *)p_idlen, d_idlen.p, n * 4, cudaMemcpyDeviceToHost, st_in));
            GNB_CUDA(cudaMemcpyAsync((void *)p_slen1, d_len1.p, n * 4, cudaMemcpyDeviceToHost, st_in));
            if (paired)
                GNB_CUDA(cudaMemcpyAsync((void *)p_slen2, d_len2.p, n * 4, cudaMemcpyDeviceToHost, st_in));
            timing.d2h_bytes += (uint64_t)n * 4 * (paired ? 4 : 3);
        }
        GNB_CUDA(cudaStreamSynchronize(st_in)); // short wait on the caller's thread: spin (a blocking wake-up can cost milliseconds)
        GNB_CUDA(cudaGetLastError());
        timing.d2h_bytes += 24;
        consumed1 = h_cons[0];
        consumed2 = paired ? h_cons[1] : 0;
        // anything irregular -> the host reader decides (wrapped records, blanks, bad letters, trailing garbage)
        const bool irregular = h_status[1] != 0 || h_status[3] != 0 || (final_block && (consumed1 != e1 || (paired && consumed2 != e2))) ||
                               (final_block && paired && n1 != n2);
        if (irregular)
            on_device = false;
        else
        {
            consumed1 = std::min<uint64_t>(consumed1, len1);
            consumed2 = std::min<uint64_t>(consumed2, len2);
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[8], ev[9]);
            timing.ms_index = ms;
        }
    }
    if (!on_device)
    {
        auto t0 = Clock::now();
        GNB_TRY(ensure_host_block(st_in)); // sliced ingest: the host reader needs the whole block
        index_reads_host(blk1, l1, final_block, kMaxReadsPerBatch - 1, t1);
        if (paired)
            index_reads_host(blk2, len2, final_block, kMaxReadsPerBatch - 1, t2);
        n = t1.size();
        if (paired)
            n = std::min(n, t2.size());
        // A parse error ends the file; the reference loses the chunk of --n-reads records being assembled (GC.cpp:1240-1283).
        // Its reader looks one record ahead (seqan3::views::chunk and std::views::take advance the file before they report
        // their end), so record e fails inside the chunk that holds record e - 1: that chunk goes too, even when it is
        // complete (tests/test_reader_cpu.py: differential against the reference binary).
        bool   err = false;
        size_t err_rec = n;
        if (t1.parse_error && t1.error_record <= n)
            err = true, err_rec = std::min(err_rec, (size_t)t1.error_record);
        if (paired && t2.parse_error && t2.error_record <= n)
            err = true, err_rec = std::min(err_rec, (size_t)t2.error_record);
        if (err)
        {
            const uint64_t abs_rec  = S->file_records + err_rec;
            const uint64_t keep_abs = abs_rec == 0 ? 0 : (abs_rec - 1) / n_reads_chunk * n_reads_chunk;
            n = keep_abs > S->file_records ? (size_t)(keep_abs - S->file_records) : 0;
            parse_error = true;
            if (!quiet)
                fprintf(stderr, "Error parsing file(s): %s\n", (t1.parse_error ? t1.error_msg : t2.error_msg).c_str());
        }
        if (!err)
            n = hold_back(n);
        t1.truncate(n);
        if (paired)
            t2.truncate(n);
        consumed1 = t1.consumed_for(n);
        consumed2 = paired ? t2.consumed_for(n) : 0;
        p_idoff = t1.id_off.data();
        p_idlen = t1.id_len.data();
        p_slen1 = t1.seq_len.data();
        p_slen2 = paired ? t2.seq_len.data() : nullptr;
        timing.ms_host_index = ms_since(t0);
        GNB_TRY(d_blk1.ensure(len1 + t1.aux.size() + 64)); // contents are kept when the buffer is already large enough
        if (!t1.aux.empty())
        {
            // ensure() may have reallocated: re-send the block
            GNB_CUDA(cudaMemcpyAsync(d_blk1.p, blk1, len1, cudaMemcpyHostToDevice, st_in));
            GNB_CUDA(cudaMemcpyAsync(d_blk1.as<uint8_t>() + len1, t1.aux.data(), t1.aux.size(), cudaMemcpyHostToDevice, st_in));
        }
        GNB_TRY(d_off1.ensure(n * 4 + 4));
        GNB_TRY(d_len1.ensure(n * 4 + 4));
        if (n)
        {
            GNB_CUDA(cudaMemcpyAsync(d_off1.p, t1.seq_off.data(), n * 4, cudaMemcpyHostToDevice, st_in));
            GNB_CUDA(cudaMemcpyAsync(d_len1.p, t1.seq_len.data(), n * 4, cudaMemcpyHostToDevice, st_in));
        }
        host_records_valid = true;
        if (S->any_device_finish)
        { // K4 writes the read ids of the output lines
            GNB_TRY(d_idoff.ensure(n * 4 + 4));
            GNB_TRY(d_idlen.ensure(n * 4 + 4));
            if (n)
            {
                GNB_CUDA(cudaMemcpyAsync(d_idoff.p, t1.id_off.data(), n * 4, cudaMemcpyHostToDevice, st_in));
                GNB_CUDA(cudaMemcpyAsync(d_idlen.p, t1.id_len.data(), n * 4, cudaMemcpyHostToDevice, st_in));
                timing.h2d_bytes += (uint64_t)n * 8;
            }
        }
        if (paired)
        {
            GNB_TRY(d_blk2.ensure(len2 + t2.aux.size() + 64));
            if (!t2.aux.empty())
            {
                GNB_CUDA(cudaMemcpyAsync(d_blk2.p, blk2, len2, cudaMemcpyHostToDevice, st_in));
                GNB_CUDA(cudaMemcpyAsync(d_blk2.as<uint8_t>() + len2, t2.aux.data(), t2.aux.size(), cudaMemcpyHostToDevice, st_in));
            }
            GNB_TRY(d_off2.ensure(n * 4 + 4));
            GNB_TRY(d_len2.ensure(n * 4 + 4));
            if (n)
            {
                GNB_CUDA(cudaMemcpyAsync(d_off2.p, t2.seq_off.data(), n * 4, cudaMemcpyHostToDevice, st_in));
                GNB_CUDA(cudaMemcpyAsync(d_len2.p, t2.seq_len.data(), n * 4, cudaMemcpyHostToDevice, st_in));
            }
        }
        timing.h2d_bytes += t1.aux.size() + (paired ? t2.aux.size() : 0) + (uint64_t)n * 8 * (paired ? 2 : 1);
    }
    trace_mark(this, "stage.indexed");
    n_reads            = (uint32_t)n;
    timing.n_reads     = n_reads;
    timing.parse_error = parse_error ? 1 : 0;
    S->file_records += n;
    if (final_block || parse_error)
        S->file_records = 0;
    GNB_TRY(d_counts.ensure((size_t)n * 4 + 4));
    GNB_TRY(d_hash_off.ensure(((size_t)n + 1) * 8));
    GNB_TRY(d_active.ensure((size_t)n + 1));
    h_active.assign(n, 1);
    h_read_level.assign(n, 0xFF);
    tuples_on_device = active_on_device = false;
    n_tuples_dev = next_active_hashes = 0;
    // the compute stream picks up after everything staged here
    GNB_CUDA(cudaEventRecord(ev_in, st_in));
    if (st != st_in)
        GNB_CUDA(cudaStreamWaitEvent(st, ev_in, 0));
    staged = true;
    return GNB_OK;
}

// Sliced ingest (bin-sharded runs): this rank's host memory holds only its slice of the block.  The host reader, the host
// finishing stage and the host-side EM append read record ids / sequences from host memory: fetch the gathered block
// from the device once per batch when one of them runs (not on the K1 + K4 path).
int BatchCtx::ensure_host_block(cudaStream_t stream)
{
    if (host_block_valid)
        return GNB_OK;
    h_blk1_copy.resize(len1 + 1);
    if (len1)
        GNB_CUDA(cudaMemcpyAsync(h_blk1_copy.data(), d_blk1.p, len1, cudaMemcpyDeviceToHost, stream));
    if (paired)
    {
        h_blk2_copy.resize(len2 + 1);
        if (len2)
            GNB_CUDA(cudaMemcpyAsync(h_blk2_copy.data(), d_blk2.p, len2, cudaMemcpyDeviceToHost, stream));
    }
    GNB_CUDA(stream_wait(stream));
    timing.d2h_bytes += len1 + len2;
    blk1 = h_blk1_copy.data();
    if (paired)
        blk2 = h_blk2_copy.data();
    host_block_valid = true;
    return GNB_OK;
}

// K2 twice (count, then write) with an exclusive scan in between
int BatchCtx::compute_hashes(uint32_t k, uint32_t w)
{
    if (hashed_k == k && hashed_w == w)
        return GNB_OK;
    const uint32_t n = n_reads;
    h_counts.resize(n);
    total_hashes = 0;
    if (n == 0)
    {
        hashed_k = k;
        hashed_w = w;
        return GNB_OK;
    }
    const uint8_t  *b2   = paired ? d_blk2.as<uint8_t>() : nullptr;
    const uint32_t *l2   = paired ? d_len2.as<uint32_t>() : nullptr;
    uint32_t       *d_max = d_status.as<uint32_t>() + 12;
    unsigned long long *d_sum = reinterpret_cast<unsigned long long *>(d_status.as<uint32_t>() + 14);
    GNB_CUDA(cudaEventRecord(ev[2], st));
    GNB_CUDA(cudaMemsetAsync(d_max, 0, 16, st));
    // One pass: every read gets room for its upper bound of minimisers (one per window), so the offsets are known
    // before K2 runs; K2 writes the hashes and the real counts.  If that layout would be too large (very long
    // reads) fall back to count -> scan -> write with exact offsets.
    GNB_TRY(d_tmp.ensure(scan_tmp_bytes(n)));
    // segments of long reads (K2t over segments): items per read | their offsets | flags
    const uint64_t seg_items_at = 0, seg_off_at = (((uint64_t)n * 4 + 7) & ~7ull), seg_flags_at = seg_off_at + ((uint64_t)n + 1) * 8;
    GNB_TRY(d_seg.ensure(seg_flags_at + n));
    uint32_t *d_items    = reinterpret_cast<uint32_t *>(d_seg.as<uint8_t>() + seg_items_at);
    uint64_t *d_item_off = reinterpret_cast<uint64_t *>(d_seg.as<uint8_t>() + seg_off_at);
    uint32_t *d_max_windows = d_status.as<uint32_t>() + 13;
    launch_hash_upper_bounds(d_len1.as<uint32_t>(), l2, n, w, d_counts.as<uint32_t>(), st, d_items, d_max_windows);
    launch_scan_counts(d_counts.as<uint32_t>(), d_hash_off.as<uint64_t>(), n, d_tmp.p, d_tmp.cap, st);
    launches += 2;
    uint64_t total_ub = 0;
    uint32_t max_windows = 0;
    GNB_CUDA(cudaMemcpyAsync(&total_ub, d_hash_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    GNB_CUDA(cudaMemcpyAsync(&max_windows, d_max_windows, 4, cudaMemcpyDeviceToHost, st));
    GNB_CUDA(stream_wait(st));
    trace_mark(this, "k2.scan_done");
    timing.d2h_bytes += 12;
    uint64_t total = 0;
    uint32_t mx    = 0;
    struct
    {
        uint32_t           mx, pad;
        unsigned long long sum;
    } agg{};
    if (total_ub * 8 <= (6ull << 30))
    {
        GNB_TRY(d_hashes.ensure((total_ub + 1) * 8));
        if (minimisers_segmented(k, w, n, max_windows))
        { // long reads: a thread per segment of 512 windows, then the segments of a read are moved together
            const uint64_t bound = minimiser_segments_bound(total_ub, n);
            GNB_TRY(d_seg_cnt.ensure(bound * 4));
            launch_scan_counts(d_items, d_item_off, n, d_tmp.p, d_tmp.cap, st);
            launch_minimisers_segmented(d_blk1.as<uint8_t>(), d_off1.as<uint32_t>(), d_len1.as<uint32_t>(), b2, d_off2.as<uint32_t>(), d_len2.as<uint32_t>(), n, k, w,
                                        d_item_off, bound, d_seg_cnt.as<uint32_t>(), d_seg.as<uint8_t>() + seg_flags_at, d_counts.as<uint32_t>(),
                                        d_hash_off.as<uint64_t>(), d_hashes.as<uint64_t>(), d_max, d_sum, st);
            launches += 4;
        }
        else
        {
            launch_minimisers(d_blk1.as<uint8_t>(), d_off1.as<uint32_t>(), d_len1.as<uint32_t>(), b2, d_off2.as<uint32_t>(), d_len2.as<uint32_t>(), n, k, w, 2,
                              d_counts.as<uint32_t>(), d_hash_off.as<uint64_t>(), d_hashes.as<uint64_t>(), d_max, d_sum, st);
            launches += 1;
        }
        GNB_CUDA(cudaEventRecord(ev[3], st));
        GNB_CUDA(cudaMemcpyAsync(&agg, d_max, 16, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(stream_wait(st));
        d_counts_valid = true;
        trace_mark(this, "k2.done");
        total = agg.sum;
        mx    = agg.mx;
    }
    else
    {
        launch_minimisers(d_blk1.as<uint8_t>(), d_off1.as<uint32_t>(), d_len1.as<uint32_t>(), b2, d_off2.as<uint32_t>(), d_len2.as<uint32_t>(), n, k, w, 0,
                          d_counts.as<uint32_t>(), nullptr, nullptr, d_max, d_sum, st);
        launch_scan_counts(d_counts.as<uint32_t>(), d_hash_off.as<uint64_t>(), n, d_tmp.p, d_tmp.cap, st);
        launches += 2;
        GNB_CUDA(cudaMemcpyAsync(&agg, d_max, 16, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(stream_wait(st));
        total = agg.sum;
        mx    = agg.mx;
        GNB_TRY(d_hashes.ensure((total + 1) * 8));
        launch_minimisers(d_blk1.as<uint8_t>(), d_off1.as<uint32_t>(), d_len1.as<uint32_t>(), b2, d_off2.as<uint32_t>(), d_len2.as<uint32_t>(), n, k, w, 1,
                          nullptr, d_hash_off.as<uint64_t>(), d_hashes.as<uint64_t>(), nullptr, nullptr, st);
        launches += 1;
        GNB_CUDA(cudaEventRecord(ev[3], st));
        d_counts_valid = true; // exact layout: counts equal the offset differences
    }
    GNB_CUDA(cudaGetLastError());
    timing.d2h_bytes += 16 + (uint64_t)n * 4;
    total_hashes  = total;
    max_hashes_ub = mx;
    hashed_k = k;
    hashed_w = w;
    return GNB_OK;
}

// HIBF traversal (counting_agent_type::bulk_count, HIBF.hpp:433-460, 506-523) as level-synchronous rounds over a
// worklist of (read, sub-IBF) items; tuples accumulate in d_tuples_a across the rounds.
int BatchCtx::run_hibf_filter(size_t li, size_t fi, uint64_t &produced_out, uint64_t start_tuples)
{
    FilterRt &F = levels[li].filters[fi];
    const uint32_t n = n_reads;
    hibf_bytes = 0;
    hibf_ms    = 0;
    hibf_round_ms.clear();
    hibf_round_bytes.clear();
    hibf_round_items.clear();
    // worklist of round 0 on the device: d_status[8..9] = item cursor of the seeding, d_status[10..11] = bytes of all rounds
    unsigned long long *d_seed  = reinterpret_cast<unsigned long long *>(d_status.as<uint32_t>() + 8);
    unsigned long long *d_bytes = reinterpret_cast<unsigned long long *>(d_status.as<uint32_t>() + 10);
    GNB_TRY(d_items_a.ensure(((uint64_t)n + 1024) * sizeof(uint2)));
    GNB_TRY(d_items_b.ensure(((uint64_t)n + 1024) * sizeof(uint2)));
    GNB_CUDA(cudaMemsetAsync(d_seed, 0, 16, st));
    launch_hibf_seed_items(li > 0 ? d_active.as<uint8_t>() : nullptr, d_counts.as<uint32_t>(), n, d_items_a.as<uint2>(), d_seed, st);
    launches += 1;
    unsigned long long n_items = 0;
    GNB_CUDA(cudaMemcpyAsync(&n_items, d_seed, 8, cudaMemcpyDeviceToHost, st));
    GNB_CUDA(cudaMemsetAsync(d_cursor.p, 0, 8, st));
    GNB_CUDA(stream_wait(st));
    timing.d2h_bytes += 8;
    unsigned long long tuples_before = start_tuples, round_bytes_sum = 0; // tuples of earlier filters of the level stay in front
    DevBuf *cur = &d_items_a, *nxt = &d_items_b;
    for (size_t round = 0; n_items; ++round)
    {
        const uint32_t lanes = round < F.round_lanes.size() ? F.round_lanes[round] : 0;
        unsigned long long got_tuples = 0, got_items = 0, bytes_so_far = 0;
        const unsigned long long items_in = n_items;
        float                    ms_round = 0;
        for (int attempt = 0; attempt < 3; ++attempt)
        {
            const uint64_t cap = d_tuples_a.cap / 8, icap = nxt->cap / sizeof(uint2);
            GNB_CUDA(cudaMemsetAsync(d_items_cursor.p, 0, 8, st));
            GNB_CUDA(cudaMemcpyAsync(d_cursor.p, &tuples_before, 8, cudaMemcpyHostToDevice, st));
            GNB_CUDA(cudaEventRecord(ev[4], st));
            // bytes are accumulated by the attempt that fits (an overflowing attempt is repeated in full)
            launch_hibf_round(F.d_ibf_table.as<IbfDev>(), F.dev.hash_funs, cur->as<uint2>(), (uint32_t)n_items, d_hashes.as<uint64_t>(),
                              d_hash_off.as<uint64_t>(), d_counts.as<uint32_t>(), std::min<uint32_t>(max_hashes_ub, 65535u), F.rel_cutoff, d_tuples_a.as<uint64_t>(),
                              d_cursor.as<unsigned long long>(), cap, nxt->as<uint2>(), d_items_cursor.as<unsigned long long>(), icap, attempt == 0 ? d_bytes : nullptr, lanes, st);
            GNB_CUDA(cudaEventRecord(ev[5], st));
            launches += 1;
            GNB_CUDA(cudaMemcpyAsync(&got_tuples, d_cursor.p, 8, cudaMemcpyDeviceToHost, st));
            GNB_CUDA(cudaMemcpyAsync(&got_items, d_items_cursor.p, 8, cudaMemcpyDeviceToHost, st));
            GNB_CUDA(cudaMemcpyAsync(&bytes_so_far, d_bytes, 8, cudaMemcpyDeviceToHost, st));
            GNB_CUDA(stream_wait(st));
            GNB_CUDA(cudaGetLastError());
            timing.d2h_bytes += 24;
            float ms1 = 0;
            cudaEventElapsedTime(&ms1, ev[4], ev[5]);
            hibf_ms += ms1;
            ms_round = ms1;
            if (got_tuples <= cap && got_items <= icap)
                break;
            // a buffer was too small: the exact need is known now.  Tuples of earlier rounds must survive the growth.
            if (got_tuples > cap)
            {
                DevBuf bigger;
                GNB_TRY(bigger.ensure(got_tuples * 8));
                if (tuples_before)
                    GNB_CUDA(cudaMemcpyAsync(bigger.p, d_tuples_a.p, tuples_before * 8, cudaMemcpyDeviceToDevice, st));
                GNB_CUDA(stream_wait(st));
                d_tuples_a.release();
                d_tuples_a = bigger;
            }
            if (got_items > icap)
                GNB_TRY(nxt->ensure(got_items * sizeof(uint2)));
        }
        // per-round record (measurement): time of the attempt that fitted, algorithmic bytes of the round, items in
        hibf_round_ms.push_back(ms_round);
        hibf_round_bytes.push_back(bytes_so_far - (hibf_round_bytes.empty() ? 0 : round_bytes_sum));
        round_bytes_sum = bytes_so_far;
        hibf_round_items.push_back(items_in);
        tuples_before = got_tuples;
        n_items       = got_items;
        std::swap(cur, nxt);
        if (n_items)
            GNB_TRY(nxt->ensure(n_items * sizeof(uint2))); // grows lazily on overflow
    }
    {
        unsigned long long b = 0;
        GNB_CUDA(cudaMemcpyAsync(&b, d_bytes, 8, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(stream_wait(st));
        hibf_bytes = b;
    }
    produced_out = tuples_before;
    return GNB_OK;
}

// K3 of this batch starts after the last kernel of the batch submitted before it
int BatchCtx::wait_turn()
{
    if (seq == 0 || turn_taken)
        return GNB_OK;
    turn_taken = true;
    if (seq > 1)
    {
        {
            std::unique_lock<std::mutex> lock(S->chain_mu);
            S->chain_cv.wait(lock, [&] { return S->chain_recorded + 1 >= seq; });
        }
        if (prev_done)
            GNB_CUDA(cudaStreamWaitEvent(st, prev_done, 0));
    }
    return GNB_OK;
}

void BatchCtx::signal_done()
{
    if (seq == 0 || done_signalled)
        return;
    done_signalled = true;
    cudaEventRecord(ev_done, st);
    {
        std::lock_guard<std::mutex> lock(S->chain_mu);
        S->chain_recorded = std::max(S->chain_recorded, seq);
    }
    S->chain_cv.notify_all();
}

// Host-resident tier: K3 over every column page of a paged filter; the tuples of all pages accumulate behind one cursor
// (a target whose bins straddle a page border yields partial sums, added by K4 like those of a shard border).  Resident
// pages are counted first; meanwhile the first two streamed pages travel host -> HBM on the filter's copy stream, and
// every further copy waits for the K3 launch that last read its staging buffer.
int BatchCtx::run_paged_count(FilterRt &F, const uint8_t *act, uint32_t n, uint64_t cap)
{
    IbfHost                    &I = F.db->ibfs[0];
    std::lock_guard<std::mutex> lock(F.db->page_mu);
    std::vector<size_t>         streamed;
    for (size_t p = 0; p < I.pages.size(); ++p)
        if (!I.pages[p].d_data)
            streamed.push_back(p);
    auto issue_copy = [&](size_t k) -> int {
        const int      b  = (int)(k & 1);
        const IbfPage &pg = I.pages[streamed[k]];
        GNB_CUDA(cudaStreamWaitEvent(I.copy_st, I.ev_free[b], 0));
        GNB_CUDA(cudaMemcpyAsync(I.d_stage[b], pg.h_data, I.page_bytes(pg), cudaMemcpyHostToDevice, I.copy_st));
        GNB_CUDA(cudaEventRecord(I.ev_ready[b], I.copy_st));
        timing.h2d_bytes += I.page_bytes(pg);
        return GNB_OK;
    };
    for (size_t k = 0; k < std::min<size_t>(2, streamed.size()); ++k)
        GNB_TRY(issue_copy(k));
    const uint32_t mh = std::min<uint32_t>(max_hashes_ub, 65535u);
    for (size_t p = 0; p < I.pages.size(); ++p)
        if (I.pages[p].d_data)
        {
            launch_ibf_count(F.pages[p].dev, d_hashes.as<uint64_t>(), d_hash_off.as<uint64_t>(), d_counts.as<uint32_t>(), act, n, mh, F.rel_cutoff,
                             d_tuples_a.as<uint64_t>(), d_cursor.as<unsigned long long>(), cap, st);
            launches += 1;
        }
    for (size_t k = 0; k < streamed.size(); ++k)
    {
        const int b = (int)(k & 1);
        GNB_CUDA(cudaStreamWaitEvent(st, I.ev_ready[b], 0));
        IbfDev d = F.pages[streamed[k]].dev;
        d.data   = I.d_stage[b];
        launch_ibf_count(d, d_hashes.as<uint64_t>(), d_hash_off.as<uint64_t>(), d_counts.as<uint32_t>(), act, n, mh, F.rel_cutoff, d_tuples_a.as<uint64_t>(),
                         d_cursor.as<unsigned long long>(), cap, st);
        launches += 1;
        GNB_CUDA(cudaEventRecord(I.ev_free[b], st));
        if (k + 2 < streamed.size())
            GNB_TRY(issue_copy(k + 2));
    }
    return GNB_OK;
}

// A level with several filters whose cross-filter merge K4 does (LevelRt::filter_bits > 0): K3 of every filter appends
// behind one cursor -- the tuples carry the filter index below the node -- then one exchange (bin-sharded runs), one sort
// by (read, node, filter) and the tuples stay in HBM for K4.
int BatchCtx::run_level_merged(size_t li, const uint8_t *act, uint64_t active_hashes)
{
    LevelRt       &L = levels[li];
    const uint32_t n = n_reads;
    for (auto &Ft : tuples[li])
        Ft.clear();
    tuples_on_device = true;
    n_tuples_dev     = 0;
    if (n == 0)
        return GNB_OK;
    uint64_t cap = d_tuples_a.cap / 8;
    if (cap < (uint64_t)n * 2 * L.filters.size() + 1024)
    {
        GNB_TRY(d_tuples_a.ensure(((uint64_t)n * 2 * L.filters.size() + 1024) * 8));
        cap = d_tuples_a.cap / 8;
    }
    unsigned long long produced = 0;
    float              ms_k3 = 0;
    GNB_TRY(wait_turn());
    if (L.filters[0].is_hibf)
    { // every HIBF's traversal appends behind the tuples of the filters before it (run_hibf_filter grows the buffer itself)
        uint64_t so_far = 0;
        for (size_t fi = 0; fi < L.filters.size(); ++fi)
        {
            uint64_t prod = 0;
            GNB_TRY(run_hibf_filter(li, fi, prod, so_far));
            so_far = prod;
            timing.count_kernel_bytes += hibf_bytes;
            ms_k3 += hibf_ms;
        }
        produced = so_far;
    }
    else
    for (int attempt = 0; attempt < 2; ++attempt)
    {
        GNB_CUDA(cudaMemsetAsync(d_cursor.p, 0, 8, st));
        GNB_CUDA(cudaEventRecord(ev[4], st));
        for (auto &F : L.filters)
        {
            if (F.pages.empty())
            {
                launch_ibf_count(F.dev, d_hashes.as<uint64_t>(), d_hash_off.as<uint64_t>(), d_counts.as<uint32_t>(), act, n, std::min<uint32_t>(max_hashes_ub, 65535u), F.rel_cutoff,
                                 d_tuples_a.as<uint64_t>(), d_cursor.as<unsigned long long>(), cap, st);
                launches += 1;
            }
            else
                GNB_TRY(run_paged_count(F, act, n, cap));
        }
        GNB_CUDA(cudaEventRecord(ev[5], st));
        GNB_CUDA(cudaMemcpyAsync(&produced, d_cursor.p, 8, cudaMemcpyDeviceToHost, st));
        timing.d2h_bytes += 8;
        GNB_CUDA(stream_wait(st));
        GNB_CUDA(cudaGetLastError());
        float ms1 = 0;
        cudaEventElapsedTime(&ms1, ev[4], ev[5]);
        ms_k3 += ms1;
        if (produced <= cap)
            break;
        GNB_TRY(d_tuples_a.ensure(produced * 8));
        cap = d_tuples_a.cap / 8;
    }
    if (!L.filters[0].is_hibf)
        for (auto &F : L.filters)
            timing.count_kernel_bytes += active_hashes * F.dev.hash_funs * (uint64_t)F.db->ibfs[0].row_words() * 8;
    const uint64_t *src    = d_tuples_a.as<uint64_t>();
    uint64_t        n_sort = produced;
    if (S->sharded())
        GNB_TRY(exchange_tuples(produced, src, n_sort));
    timing.ms_count += ms_k3;
    timing.ms_exchange += ms_exchange_acc;
    ms_exchange_acc = 0;
    if (produced == 0)
        return GNB_OK;
    GNB_CUDA(cudaEventRecord(ev[6], st));
    GNB_TRY(d_tuples_b.ensure(n_sort * 8));
    GNB_TRY(d_tmp.ensure(sort_tmp_bytes(n_sort)));
    launch_sort_tuples(src, d_tuples_b.as<uint64_t>(), n_sort, d_tmp.p, d_tmp.cap, st);
    GNB_CUDA(cudaEventRecord(ev[7], st));
    GNB_CUDA(stream_wait(st));
    n_tuples_dev = produced;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev[6], ev[7]);
    timing.ms_sort += ms;
    return GNB_OK;
}

// K3 (+ sort) for every filter of level li on the reads still active
int BatchCtx::run_level(size_t li)
{
    LevelRt &L = levels[li];
    int rc = compute_hashes(L.k, L.w);
    if (rc != GNB_OK)
        return rc;
    const uint32_t n = n_reads;
    const uint8_t *act = nullptr;
    if (li > 0 && n)
    {
        if (!active_on_device)
        {
            GNB_CUDA(cudaMemcpyAsync(d_active.p, h_active.data(), n, cudaMemcpyHostToDevice, st));
            timing.h2d_bytes += n;
        }
        act = d_active.as<uint8_t>();
    }
    const bool stay = keep_on_device && L.device_finish && L.filters.size() == 1; // K4 consumes the tuples in HBM
    tuples_on_device = false;
    n_tuples_dev     = 0;
    uint64_t active_hashes = 0;
    if (li == 0 && max_hashes_ub <= 65535)
        active_hashes = total_hashes;
    else if (li > 0 && active_on_device)
        active_hashes = next_active_hashes;
    else
        for (uint32_t i = 0; i < n; ++i)
            if (h_active[i] && h_counts[i] <= 65535)
                active_hashes += h_counts[i];
    if (keep_on_device && L.device_finish && L.filters.size() > 1)
        return run_level_merged(li, act, active_hashes);
    float ms_sort = 0, ms_k3 = 0;
    for (size_t fi = 0; fi < L.filters.size(); ++fi)
    {
        FilterRt              &F  = L.filters[fi];
        PinnedVec<uint64_t>   &Ft = tuples[li][fi];
        Ft.clear();
        if (stay)
            tuples_on_device = true;
        if (n == 0)
            continue;
        uint64_t cap = d_tuples_a.cap / 8;
        if (cap < (uint64_t)n * 2 + 1024)
        {
            GNB_TRY(d_tuples_a.ensure(((uint64_t)n * 2 + 1024) * 8));
            cap = d_tuples_a.cap / 8;
        }
        unsigned long long produced = 0;
        GNB_TRY(wait_turn());
        if (F.is_hibf)
        {
            uint64_t prod = 0;
            GNB_TRY(run_hibf_filter(li, fi, prod));
            produced = prod;
            timing.count_kernel_bytes += hibf_bytes;
            ms_k3 += hibf_ms;
        }
        else
        for (int attempt = 0; attempt < 2; ++attempt)
        {
            GNB_CUDA(cudaMemsetAsync(d_cursor.p, 0, 8, st));
            GNB_CUDA(cudaEventRecord(ev[4], st));
            if (F.pages.empty())
            {
                launch_ibf_count(F.dev, d_hashes.as<uint64_t>(), d_hash_off.as<uint64_t>(), d_counts.as<uint32_t>(), act, n, std::min<uint32_t>(max_hashes_ub, 65535u), F.rel_cutoff,
                                 d_tuples_a.as<uint64_t>(), d_cursor.as<unsigned long long>(), cap, st);
                launches += 1;
            }
            else
                GNB_TRY(run_paged_count(F, act, n, cap));
            GNB_CUDA(cudaEventRecord(ev[5], st));
            GNB_CUDA(cudaMemcpyAsync(&produced, d_cursor.p, 8, cudaMemcpyDeviceToHost, st));
            timing.d2h_bytes += 8;
            GNB_CUDA(stream_wait(st));
            trace_mark(this, "k3.done");
            GNB_CUDA(cudaGetLastError());
            {
                float ms1 = 0;
                cudaEventElapsedTime(&ms1, ev[4], ev[5]);
                ms_k3 += ms1;
            }
            if (produced <= cap)
                break;
            GNB_TRY(d_tuples_a.ensure(produced * 8)); // exact size is now known: run again
            cap = d_tuples_a.cap / 8;
        }
        if (!F.is_hibf)
            timing.count_kernel_bytes += active_hashes * F.dev.hash_funs * (uint64_t)F.db->ibfs[0].row_words() * 8;
        const uint64_t *src    = d_tuples_a.as<uint64_t>();
        uint64_t        n_sort = produced;
        if (S->sharded())
            GNB_TRY(exchange_tuples(produced, src, n_sort)); // the lists of all ranks (padded slots: all-ones sort last)
        if (produced == 0)
            continue;
        GNB_CUDA(cudaEventRecord(ev[6], st));
        GNB_TRY(d_tuples_b.ensure(n_sort * 8));
        const size_t tb = sort_tmp_bytes(n_sort);
        GNB_TRY(d_tmp.ensure(tb));
        launch_sort_tuples(src, d_tuples_b.as<uint64_t>(), n_sort, d_tmp.p, d_tmp.cap, st);
        GNB_CUDA(cudaEventRecord(ev[7], st));
        if (stay)
            n_tuples_dev = produced;
        else
        {
            Ft.resize(produced);
            timing.d2h_bytes += produced * 8;
            GNB_CUDA(cudaMemcpyAsync(Ft.data(), d_tuples_b.p, produced * 8, cudaMemcpyDeviceToHost, st));
        }
        GNB_CUDA(stream_wait(st));
        trace_mark(this, "sort.done");
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[6], ev[7]);
        ms_sort += ms;
    }
    timing.ms_exchange += ms_exchange_acc;
    ms_exchange_acc = 0;
    timing.ms_count += ms_k3;
    timing.ms_sort += ms_sort;
    return GNB_OK;
}

// Bin-sharded run: this rank's K3 saw only its bin-word columns, so d_tuples_a holds the candidates of
// its own bins -- finished counts for targets inside the shard, partial sums (flag bit 16) for targets that straddle a
// shard boundary.  Every rank needs all of them (K4 then adds the partial sums, applies cap / cutoff / rel-filter /
// fpr-query / LCA exactly as on one GPU and produces the identical result everywhere).  ONE collective on the compute
// stream, inside this batch's GPU turn, hence in the same order on every rank: an all-gather of fixed-size slots
// `[cap tuples | count]` (a few bytes per read, against 2 * bins bytes per read for the count vectors a dense allreduce
// would move).  The slot size is a function of the batch's read count and, after an overflow, of the gathered counts, so
// all ranks agree on it without talking; unused slot words are all-ones and sort behind every tuple.  One host sync (the
// counts).  On return `list` points at the gathered slots, `n_sort` is their length in words and `produced` the number
// of real tuples among them.
int BatchCtx::exchange_tuples(unsigned long long &produced, const uint64_t *&list, uint64_t &n_sort)
{
    gnb_comm    *cm = S->comm;
    const size_t N  = (size_t)cm->n_ranks, r = (size_t)cm->rank;
    GNB_TRY(h_xch.ensure(64 + N * 8));
    uint64_t *cnt = h_xch.as<uint64_t>() + 8;
    uint64_t  cap = (uint64_t)n_reads / 4 + 1024;
    GNB_CUDA(cudaEventRecord(ev[12], st));
    for (;;)
    {
        const uint64_t slot = cap + 1;
        GNB_TRY(d_gather.ensure(N * slot * 8));
        uint64_t *own = d_gather.as<uint64_t>() + r * slot;
        GNB_CUDA(cudaMemsetAsync(own, 0xFF, slot * 8, st));
        if (produced)
            GNB_CUDA(cudaMemcpyAsync(own, d_tuples_a.p, std::min<uint64_t>(produced, cap) * 8, cudaMemcpyDeviceToDevice, st));
        // d_cursor holds this rank's `produced` (the last K3 attempt fitted: produced <= capacity of d_tuples_a)
        GNB_CUDA(cudaMemcpyAsync(own + cap, d_cursor.p, 8, cudaMemcpyDeviceToDevice, st));
        GNB_TRY(comm_all_gather(cm->nccl, own, d_gather.p, slot * 8, st));
        GNB_CUDA(cudaMemcpy2DAsync(cnt, 8, d_gather.as<uint64_t>() + cap, slot * 8, 8, N, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemset2DAsync(d_gather.as<uint64_t>() + cap, slot * 8, 0xFF, 8, N, st)); // the counts must not be sorted as tuples
        GNB_CUDA(cudaEventRecord(ev[13], st));
        GNB_CUDA(stream_wait(st));
        timing.d2h_bytes += N * 8;
        uint64_t mx = 0;
        for (size_t q = 0; q < N; ++q)
            mx = std::max(mx, cnt[q]);
        if (cnt[r] != produced)
            return fail(GNB_ERR_CUDA, "sharded exchange: tuple count mismatch on this rank");
        if (mx <= cap)
            break;
        cap = 2 * mx; // the same decision on every rank: go again with slots that hold the longest list
    }
    uint64_t total = 0;
    for (size_t q = 0; q < N; ++q)
        total += cnt[q];
    float ms = 0;
    cudaEventElapsedTime(&ms, ev[12], ev[13]);
    ms_exchange_acc += ms;
    timing.exchanged_bytes += total * 8;
    produced = total;
    list     = d_gather.as<uint64_t>();
    n_sort   = total ? N * (cap + 1) : 0;
    trace_mark(this, "exchange.done");
    return GNB_OK;
}

// host finishing stage for level li
int BatchCtx::finish_level(size_t li)
{
    LevelRt       &L = levels[li];
    const uint32_t n = n_reads;
    const bool     first = li == 0, last = li + 1 == levels.size();
    const int      T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, (n + 4095) / 4096));
    GNB_TRY(ensure_host_block(st));
    const uint8_t *id_base = reinterpret_cast<const uint8_t *>(blk1);
    finish_T = T;

    auto work = [&](int tid) {
        Worker        &W  = workers[tid];
        const uint32_t r0 = (uint32_t)((uint64_t)n * tid / T), r1 = (uint32_t)((uint64_t)n * (tid + 1) / T);
        gnb_totals    &tot = W.total[li];
        const bool     dense = !W.rep_dense[li].empty();
        auto rep_at = [&](uint32_t node) -> Rep & {
            if (dense)
            {
                Rep &x = W.rep_dense[li][node];
                if (!(x.matches | x.seqs_lca | x.seqs_unique | x.discarded_matches_filter | x.discarded_matches_fprquery))
                    W.rep_touched[li].push_back(node);
                return x;
            }
            return W.rep[li][node];
        };
        const size_t   nf  = L.filters.size();
        std::vector<size_t> cur(nf);
        for (size_t f = 0; f < nf; ++f)
        {
            const auto &tp = tuples[li][f];
            cur[f] = std::lower_bound(tp.begin(), tp.end(), (uint64_t)r0 << kTupleReadShift) - tp.begin();
        }
        struct M
        {
            uint32_t node, count;
            double   fpr;
            uint8_t  cls;
        };
        if (L.fpr_query < 1.0 && !L.fpr_classes.empty() && W.memo_table.size() != L.fpr_classes.size() * 65536)
            W.memo_table.assign(L.fpr_classes.size() * 65536, std::numeric_limits<double>::quiet_NaN());
        std::vector<M> best, merged, one_filter;
        for (uint32_t r = r0; r < r1; ++r)
        {
            if (!h_active[r])
                continue;
            const uint32_t nh  = h_counts[r];
            const uint32_t l1  = p_slen1[r], l2 = paired ? p_slen2[r] : 0;
            const bool     small = l1 < L.w, big = nh > 65535;
            if (first)
            {
                if (small)
                    tot.seqs_skipped_small++;
                else if (big)
                    tot.seqs_skipped_big++;
                else
                {
                    tot.seqs_processed++;
                    tot.length_processed += (uint64_t)l1 + l2;
                    tot.kmers_processed += nh;
                }
            }
            uint64_t max_c = 0, min_c = nh;
            best.clear();
            for (size_t f = 0; f < nf; ++f)
            {
                const FilterRt &F  = L.filters[f];
                const auto     &tp = tuples[li][f];
                size_t         &c  = cur[f];
                if (c >= tp.size() || (uint32_t)(tp[c] >> kTupleReadShift) != r)
                    continue;
                const uint32_t cutoff = threshold_cutoff(nh, F.rel_cutoff);
                one_filter.clear();
                while (c < tp.size() && (uint32_t)(tp[c] >> kTupleReadShift) == r)
                {
                    const uint32_t node = ((uint32_t)(tp[c] >> kTupleNodeShift) & (kMaxNodes - 1)) >> L.filter_bits; // the filter index sits below
                    uint64_t       sum  = 0;
                    bool           partial = false;
                    while (c < tp.size() && (uint32_t)(tp[c] >> kTupleReadShift) == r && (((uint32_t)(tp[c] >> kTupleNodeShift) & (kMaxNodes - 1)) >> L.filter_bits) == node)
                    {
                        sum += tp[c] & 0xFFFF;
                        partial |= ((tp[c] >> 16) & 1) != 0;
                        ++c;
                    }
                    if (F.is_hibf)
                    { // running sum of the counter type wraps (HIBF.hpp:437-441), then select_matches caps (GC.cpp:560-563)
                        if (partial)
                            sum &= 0xFFFF;
                        if (sum < cutoff || sum == 0)
                            continue;
                        if (sum > nh)
                            sum = nh;
                    }
                    else if (partial)
                    {
                        if (sum > nh)
                            sum = nh; // GC.cpp:525-526
                        if (sum < cutoff)
                            continue;
                    }
                    one_filter.push_back(M{node, (uint32_t)sum, F.node_fpr[node], F.node_fpr_class[node]});
                }
                // merge into best (both sorted by node): keep the strictly larger count (GC.cpp:531-539)
                if (best.empty())
                {
                    for (auto const &m : one_filter)
                    {
                        max_c = std::max<uint64_t>(max_c, m.count);
                        min_c = std::min<uint64_t>(min_c, m.count);
                    }
                    best.swap(one_filter);
                }
                else
                {
                    merged.clear();
                    size_t a = 0, b = 0;
                    while (a < best.size() || b < one_filter.size())
                    {
                        if (b >= one_filter.size() || (a < best.size() && best[a].node < one_filter[b].node))
                            merged.push_back(best[a++]);
                        else if (a >= best.size() || one_filter[b].node < best[a].node)
                        {
                            const M &m = one_filter[b++];
                            max_c = std::max<uint64_t>(max_c, m.count);
                            min_c = std::min<uint64_t>(min_c, m.count);
                            merged.push_back(m);
                        }
                        else
                        {
                            const M &m = one_filter[b++];
                            if (m.count > best[a].count)
                            {
                                max_c = std::max<uint64_t>(max_c, m.count);
                                min_c = std::min<uint64_t>(min_c, m.count); // the replaced count stays in min (reference quirk)
                                merged.push_back(m);
                            }
                            else
                                merged.push_back(best[a]);
                            ++a;
                        }
                    }
                    best.swap(merged);
                }
            }
            size_t kept = 0;
            const size_t m_begin = W.m_target.size();
            if (max_c > 0)
            {
                const uint64_t thr_ceil = (uint64_t)std::ceil((double)(max_c - min_c) * L.rel_filter);
                const double   threshold_filter = (double)(max_c - thr_ceil);
                for (auto const &m : best)
                {
                    if ((double)m.count >= threshold_filter)
                    {
                        if (L.fpr_query < 1.0)
                        {
                            double q;
                            if (m.cls != 255 && nh < 256 && m.count < 256 && !W.memo_table.empty())
                            {
                                double &slot = W.memo_table[((size_t)m.cls << 16) | (nh << 8) | m.count];
                                if (std::isnan(slot))
                                    slot = fpr_query_q(nh, m.count, m.fpr);
                                q = slot;
                            }
                            else
                            {
                                FprKey key;
                                memcpy(&key.fpr_bits, &m.fpr, 8);
                                key.n = nh;
                                key.c = m.count;
                                auto it = W.memo.find(key);
                                if (it != W.memo.end())
                                    q = it->second;
                                else
                                {
                                    q = fpr_query_q(nh, m.count, m.fpr);
                                    W.memo.emplace(key, q);
                                }
                            }
                            if (q > L.fpr_query)
                            {
                                rep_at(m.node).discarded_matches_fprquery++;
                                tot.discarded_matches_fprquery++;
                                continue;
                            }
                        }
                        rep_at(m.node).matches++;
                        tot.matches++;
                        W.m_target.push_back(m.node);
                        W.m_count.push_back(m.count);
                        ++kept;
                    }
                    else
                    {
                        rep_at(m.node).discarded_matches_filter++;
                        tot.discarded_matches_filter++;
                    }
                }
            }
            if (kept > 0)
            {
                tot.seqs_classified++;
                tot.kmers_from_classified_seqs += nh;
                tot.kmers_matches += max_c;
                W.n_classified++;
                h_active[r]     = 0;
                h_read_level[r] = (uint8_t)li;
                const char  *id  = reinterpret_cast<const char *>(id_base + p_idoff[r]);
                const size_t idl = p_idlen[r];
                uint32_t     one_node = W.m_target[m_begin];
                uint64_t     one_count = W.m_count[m_begin];
                if (kept == 1)
                {
                    rep_at(one_node).seqs_unique++;
                    tot.seqs_unique++;
                }
                else if (!skip_lca)
                {
                    uint32_t l = L.lca2(W.m_target[m_begin], W.m_target[m_begin + 1]);
                    for (size_t i = 2; i < kept; ++i)
                        l = L.lca2(l, W.m_target[m_begin + i]);
                    rep_at(l).seqs_lca++;
                    one_node  = l;
                    one_count = max_c;
                }
                else
                    rep_at((uint32_t)L.root).seqs_lca++;
                if (!skip_lca && cfg.output_lca)
                {
                    std::string &o = W.one_text[li];
                    o.append(id, idl);
                    o.push_back('\t');
                    o.append(L.node_names[one_node]);
                    o.push_back('\t');
                    append_u64(o, one_count);
                    o.push_back('\n');
                }
                if (cfg.output_all)
                {
                    std::string &o = W.all_text[li];
                    for (size_t i = 0; i < kept; ++i)
                    {
                        o.append(id, idl);
                        o.push_back('\t');
                        o.append(L.node_names[W.m_target[m_begin + i]]);
                        o.push_back('\t');
                        append_u64(o, W.m_count[m_begin + i]);
                        o.push_back('\n');
                    }
                }
                W.m_off.push_back(r);
                W.m_off.push_back(kept);
            }
            else
            {
                W.m_target.resize(m_begin);
                W.m_count.resize(m_begin);
                if (last && cfg.output_unclassified)
                {
                    W.unc_text.append(reinterpret_cast<const char *>(id_base + p_idoff[r]), p_idlen[r]);
                    W.unc_text.push_back('\n');
                }
            }
        }
    };
    if (T == 1)
        work(0);
    else
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back(work, t);
        for (auto &t : th)
            t.join();
    }
    return GNB_OK;
}

// The record table (ids, sequence lengths) in host memory: with K4 on every level it is only needed when a level falls
// back to the host finishing stage.
int BatchCtx::fetch_host_records()
{
    if (host_records_valid)
        return GNB_OK;
    const size_t n = n_reads;
    if (n)
    {
        GNB_CUDA(cudaMemcpyAsync((void *)p_idoff, d_idoff.p, n * 4, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemcpyAsync((void *)p_idlen, d_idlen.p, n * 4, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemcpyAsync((void *)p_slen1, d_len1.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (paired)
            GNB_CUDA(cudaMemcpyAsync((void *)p_slen2, d_len2.p, n * 4, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(stream_wait(st));
        timing.d2h_bytes += (uint64_t)n * 4 * (paired ? 4 : 3);
    }
    host_records_valid = true;
    return GNB_OK;
}

// Everything the host finishing stage of level li reads, brought to the host (after K4 declined the level).
int BatchCtx::to_host_state(size_t li)
{
    GNB_TRY(fetch_host_records());
    const size_t n = n_reads;
    bool split = false;
    if (tuples_on_device)
    {
        PinnedVec<uint64_t> &Ft = tuples[li][0];
        Ft.resize(n_tuples_dev);
        if (n_tuples_dev)
        {
            GNB_CUDA(cudaMemcpyAsync(Ft.data(), d_tuples_b.p, n_tuples_dev * 8, cudaMemcpyDeviceToHost, st));
            timing.d2h_bytes += n_tuples_dev * 8;
        }
        tuples_on_device = false;
        split            = levels[li].filter_bits > 0 && levels[li].filters.size() > 1;
    }
    if (active_on_device && n)
    {
        GNB_CUDA(cudaMemcpyAsync(h_active.data(), d_active.p, n, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(cudaMemcpyAsync(h_read_level.data(), d_read_level.p, n, cudaMemcpyDeviceToHost, st));
        timing.d2h_bytes += 2 * n;
    }
    GNB_CUDA(stream_wait(st));
    active_on_device = false;
    if (split)
    { // the merged list (sorted by read, node, filter) back into one list per filter, each still sorted by (read, node)
        const uint32_t      fmask = (1u << levels[li].filter_bits) - 1;
        PinnedVec<uint64_t> all;
        all.swap(tuples[li][0]);
        for (uint64_t t : all)
            tuples[li][(size_t)((t >> kTupleNodeShift) & fmask)].push_back(t);
    }
    return GNB_OK;
}

// K4: finishing stage of level li on the device.  done = false: the level has to go through the host finishing stage
// (an --fpr-query value inside the guard band, very long reads, or nothing staged for the device).
// rep: report accumulators to add to; fetch: copy the level's result to the host.
int BatchCtx::finish_level_device(size_t li, unsigned long long *rep, bool fetch, bool &done)
{
    done             = false;
    LevelRt       &L = levels[li];
    const uint32_t n = n_reads;
    // (the device evaluates --fpr-query with CUDA's lgamma / exp / pow, trusted inside the guard band up to 4096 minimisers:
    // longer reads take the host stage only when --fpr-query is in use)
    if (!L.device_finish || !tuples_on_device || (max_hashes_ub > 4096 && L.fpr_query < 1.0) || n_tuples_dev >= 0xFFFFFFFFull)
        return GNB_OK;
    const bool first = li == 0, last = li + 1 == levels.size();
    LevelOut  &O     = lv[li];
    if (n == 0)
    {
        O.on_device = true;
        O.n_matches = O.all_len = O.one_len = 0;
        O.total     = gnb_totals{};
        GNB_TRY(O.h_moff.ensure(8));
        O.h_moff.as<uint64_t>()[0] = 0;
        if (last)
        {
            unc_on_device = true;
            unc_len_dev   = 0;
        }
        done = true;
        return GNB_OK;
    }
    GNB_CUDA(cudaEventRecord(ev[10], st));
    GNB_TRY(d_read_level.ensure(n));
    if (first)
        GNB_CUDA(cudaMemsetAsync(d_read_level.p, 0xFF, n, st));
    else if (!active_on_device)
    {
        GNB_CUDA(cudaMemcpyAsync(d_active.p, h_active.data(), n, cudaMemcpyHostToDevice, st));
        GNB_CUDA(cudaMemcpyAsync(d_read_level.p, h_read_level.data(), n, cudaMemcpyHostToDevice, st));
        timing.h2d_bytes += 2ull * n;
    }
    GNB_TRY(d_tstart.ensure((size_t)n * 4));
    GNB_TRY(d_nacc.ensure((size_t)n * 4));
    GNB_TRY(d_sizes.ensure(((size_t)n + 1) * sizeof(FinishSizes)));
    GNB_TRY(d_offs.ensure(((size_t)n + 1) * sizeof(FinishSizes)));
    GNB_TRY(d_one.ensure((size_t)n * 8));
    GNB_TRY(d_ftotals.ensure(kFinishTotals * 8));
    GNB_TRY(d_moff.ensure(((size_t)n + 1) * 8));
    GNB_TRY(d_tuples_a.ensure(n_tuples_dev * 8 + 8));
    GNB_TRY(d_tmp.ensure(finish_scan_tmp_bytes(n)));
    GNB_TRY(h_fin.ensure(sizeof(FinishSizes) + kFinishTotals * 8));

    FinishParams P{};
    P.tuples      = d_tuples_b.as<uint64_t>();
    P.n_tuples    = n_tuples_dev;
    P.entries     = d_tuples_a.as<uint64_t>();
    P.tuple_start = d_tstart.as<uint32_t>();
    P.n_hashes    = d_counts.as<uint32_t>();
    P.len1        = d_len1.as<uint32_t>();
    P.len2        = paired ? d_len2.as<uint32_t>() : nullptr;
    P.id_off      = d_idoff.as<uint32_t>();
    P.id_len      = d_idlen.as<uint32_t>();
    P.blk1        = d_blk1.as<uint8_t>();
    P.active      = d_active.as<uint8_t>();
    P.read_level  = d_read_level.as<uint8_t>();
    P.n_reads     = n;
    P.n_acc       = d_nacc.as<uint32_t>();
    P.sizes       = d_sizes.as<FinishSizes>();
    P.offs        = d_offs.as<FinishSizes>();
    P.one         = d_one.as<uint2>();
    P.totals      = d_ftotals.as<unsigned long long>();
    P.match_off   = d_moff.as<uint64_t>();
    P.node_fpr    = L.d_node_fpr.as<double>();
    P.node_class  = L.d_node_class.as<uint32_t>();
    P.fpr_memo    = L.d_fpr_memo.as<unsigned long long>();
    P.fpr_memo_mask = (uint32_t)(kFprMemoSlots - 1);
    P.parent      = L.d_parent.as<int32_t>();
    P.depth       = L.d_depth.as<uint32_t>();
    P.name_off    = L.d_name_off.as<uint32_t>();
    P.names       = L.d_names.as<char>();
    P.rep         = rep;
    P.root        = L.root;
    for (size_t f = 0; f < L.filters.size() && f < 16; ++f)
        P.rel_cutoffs[f] = L.filters[f].rel_cutoff;
    P.filter_bits = L.filter_bits;
    P.n_nodes     = (uint32_t)L.node_names.size();
    P.rel_filter  = L.rel_filter;
    P.fpr_query   = L.fpr_query;
    P.fpr_band    = S->fpr_band;
    P.w           = L.w;
    P.level       = (uint32_t)li;
    P.is_hibf     = L.filters[0].is_hibf;
    P.skip_lca    = skip_lca;
    P.output_lca  = cfg.output_lca != 0;
    P.output_all  = cfg.output_all != 0;
    P.output_unc  = cfg.output_unclassified != 0;
    P.first       = first;
    P.last        = last;
    launch_finish_select(P, st);
    launch_finish_scan(P, d_tmp.p, d_tmp.cap, st);
    launches += 3;
    FinishSizes        *h_tot = h_fin.as<FinishSizes>();
    unsigned long long *h_ft  = reinterpret_cast<unsigned long long *>(h_tot + 1);
    GNB_CUDA(cudaMemcpyAsync(h_tot, d_offs.as<FinishSizes>() + n, sizeof(FinishSizes), cudaMemcpyDeviceToHost, st));
    GNB_CUDA(cudaMemcpyAsync(h_ft, d_ftotals.p, kFinishTotals * 8, cudaMemcpyDeviceToHost, st));
    GNB_CUDA(stream_wait(st));
    trace_mark(this, "k4.select_done");
    GNB_CUDA(cudaGetLastError());
    timing.d2h_bytes += sizeof(FinishSizes) + kFinishTotals * 8;
    if (h_ft[kFtAmbiguous])
        return GNB_OK; // nothing outside the scratch buffers was touched: the host stage takes the level
    const FinishSizes T = *h_tot;
    GNB_TRY(d_mt.ensure(T.kept * 4 + 4));
    GNB_TRY(d_mc.ensure(T.kept * 4 + 4));
    GNB_TRY(d_all.ensure(T.all_bytes + 1));
    GNB_TRY(d_one_txt.ensure(T.one_bytes + 1));
    GNB_TRY(d_unc.ensure(T.unc_bytes + 1));
    P.match_target = d_mt.as<uint32_t>();
    P.match_count  = d_mc.as<uint32_t>();
    P.all_text     = d_all.as<char>();
    P.one_text     = d_one_txt.as<char>();
    P.unc_text     = d_unc.as<char>();
    launch_finish_write(P, st);
    launches += 1;
    GNB_CUDA(cudaEventRecord(ev[11], st));
    if (S->keep_matches && em_account)
        GNB_TRY(em_append_device(li, P, T.kept)); // before the turn is passed on: the store keeps submission order
    if (last)
        signal_done(); // the next batch may start K3 while this one's result travels to the host
    if (fetch)
    {
        GNB_TRY(O.h_moff.ensure(((size_t)n + 1) * 8));
        GNB_TRY(O.h_mt.ensure(T.kept * 4 + 4));
        GNB_TRY(O.h_mc.ensure(T.kept * 4 + 4));
        GNB_TRY(O.h_all.ensure(T.all_bytes + 1));
        GNB_TRY(O.h_one.ensure(T.one_bytes + 1));
        GNB_CUDA(cudaMemcpyAsync(O.h_moff.p, d_moff.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (T.kept)
        {
            GNB_CUDA(cudaMemcpyAsync(O.h_mt.p, d_mt.p, T.kept * 4, cudaMemcpyDeviceToHost, st));
            GNB_CUDA(cudaMemcpyAsync(O.h_mc.p, d_mc.p, T.kept * 4, cudaMemcpyDeviceToHost, st));
        }
        if (T.all_bytes)
            GNB_CUDA(cudaMemcpyAsync(O.h_all.p, d_all.p, T.all_bytes, cudaMemcpyDeviceToHost, st));
        if (T.one_bytes)
            GNB_CUDA(cudaMemcpyAsync(O.h_one.p, d_one_txt.p, T.one_bytes, cudaMemcpyDeviceToHost, st));
        timing.d2h_bytes += ((uint64_t)n + 1) * 8 + T.kept * 8 + T.all_bytes + T.one_bytes;
        if (last)
        {
            GNB_TRY(h_unc.ensure(T.unc_bytes + 1));
            if (T.unc_bytes)
                GNB_CUDA(cudaMemcpyAsync(h_unc.p, d_unc.p, T.unc_bytes, cudaMemcpyDeviceToHost, st));
            timing.d2h_bytes += T.unc_bytes;
        }
    }
    GNB_CUDA(stream_wait(st));
    trace_mark(this, "k4.write_d2h_done");
    GNB_CUDA(cudaGetLastError());
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev[10], ev[11]) == cudaSuccess)
        ms_finish_dev += ms;
    O.on_device = true;
    O.n_matches = T.kept;
    O.all_len   = T.all_bytes;
    O.one_len   = T.one_bytes;
    gnb_totals &t = O.total;
    t             = gnb_totals{};
    t.seqs_processed             = h_ft[kFtProcessed];
    t.seqs_skipped_big           = h_ft[kFtSkippedBig];
    t.seqs_skipped_small         = h_ft[kFtSkippedSmall];
    t.length_processed           = h_ft[kFtLength];
    t.kmers_processed            = h_ft[kFtKmers];
    t.seqs_classified            = h_ft[kFtClassified];
    t.kmers_matches              = h_ft[kFtKmersMatches];
    t.kmers_from_classified_seqs = h_ft[kFtKmersClassified];
    t.matches                    = h_ft[kFtMatches];
    t.seqs_unique                = h_ft[kFtUnique];
    t.discarded_matches_filter   = h_ft[kFtDiscFilter];
    t.discarded_matches_fprquery = h_ft[kFtDiscFpr];
    next_active_hashes           = h_ft[kFtActiveHashesNext];
    if (last)
    {
        unc_on_device = true;
        unc_len_dev   = T.unc_bytes;
    }
    active_on_device = true;
    levels_on_device += 1;
    done = true;
    return GNB_OK;
}

// EM: the level's classified reads (ids + kept matches, targets as run-wide ids) go to the run's store in HBM
int BatchCtx::em_append_device(size_t li, const FinishParams &P, uint64_t n_kept)
{
    const uint32_t n = n_reads;
    GNB_TRY(d_em_sizes.ensure(((size_t)n + 1) * sizeof(EmSizes)));
    GNB_TRY(d_em_offs.ensure(((size_t)n + 1) * sizeof(EmSizes)));
    GNB_TRY(d_tmp.ensure(em_scan_tmp_bytes(n)));
    launch_em_sizes(P.sizes, P.id_len, n, d_em_sizes.as<EmSizes>(), d_em_offs.as<EmSizes>(), d_tmp.p, d_tmp.cap, st);
    EmSizes tot{};
    GNB_CUDA(cudaMemcpyAsync(&tot, d_em_offs.as<EmSizes>() + n, sizeof(EmSizes), cudaMemcpyDeviceToHost, st));
    GNB_CUDA(stream_wait(st));
    launches += 3;
    if (tot.reads == 0)
        return GNB_OK;
    std::lock_guard<std::mutex> lock(S->em_mutex);
    EmStore &E = S->em_store(cur_prefix, li);
    GNB_TRY(E.reserve(tot.reads, n_kept, tot.id_bytes));
    launch_em_append(P.sizes, d_em_offs.as<EmSizes>(), P.match_off, P.match_target, P.match_count, n_kept, P.id_off, P.id_len, P.blk1, n,
                     levels[li].d_em_map.as<uint32_t>(), E.dev(), E.n_reads, E.n_matches, E.id_bytes, st);
    launches += 2;
    E.n_reads += tot.reads;
    E.n_matches += n_kept;
    E.id_bytes += tot.id_bytes;
    GNB_CUDA(cudaGetLastError());
    return GNB_OK;
}

// the same for a level finished by the host stage: the workers' new (read, kept) pairs and match lists of this level
int BatchCtx::em_append_host(size_t li, const std::vector<size_t> &off_before, const std::vector<size_t> &tgt_before)
{
    std::vector<uint64_t> off, id_off;
    std::vector<uint32_t> tgt, cnt;
    std::string           ids;
    const auto           &map = levels[li].em_map;
    const char           *base = blk1;
    std::lock_guard<std::mutex> lock(S->em_mutex);
    EmStore &E = S->em_store(cur_prefix, li);
    for (size_t w = 0; w < workers.size(); ++w)
    {
        const Worker &W = workers[w];
        size_t        p = tgt_before[w];
        for (size_t i = off_before[w]; i + 1 < W.m_off.size(); i += 2)
        {
            const uint64_t r = W.m_off[i], k = W.m_off[i + 1];
            off.push_back(E.n_matches + tgt.size());
            id_off.push_back(E.id_bytes + ids.size());
            ids.append(base + p_idoff[r], p_idlen[r]);
            for (uint64_t j = 0; j < k; ++j)
            {
                tgt.push_back(map[W.m_target[p + j]]);
                cnt.push_back(W.m_count[p + j]);
            }
            p += k;
        }
    }
    if (off.empty())
        return GNB_OK;
    GNB_TRY(E.reserve(off.size(), tgt.size(), ids.size()));
    off.push_back(E.n_matches + tgt.size()); // sentinels
    id_off.push_back(E.id_bytes + ids.size());
    GNB_CUDA(cudaMemcpyAsync(E.off.as<uint64_t>() + E.n_reads, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    GNB_CUDA(cudaMemcpyAsync(E.id_off.as<uint64_t>() + E.n_reads, id_off.data(), id_off.size() * 8, cudaMemcpyHostToDevice, st));
    GNB_CUDA(cudaMemcpyAsync(E.tgt.as<uint32_t>() + E.n_matches, tgt.data(), tgt.size() * 4, cudaMemcpyHostToDevice, st));
    GNB_CUDA(cudaMemcpyAsync(E.cnt.as<uint32_t>() + E.n_matches, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice, st));
    if (!ids.empty())
        GNB_CUDA(cudaMemcpyAsync(E.ids.as<char>() + E.id_bytes, ids.data(), ids.size(), cudaMemcpyHostToDevice, st));
    GNB_CUDA(stream_wait(st));
    E.n_reads += off.size() - 1;
    E.n_matches += tgt.size();
    E.id_bytes += ids.size();
    return GNB_OK;
}

void BatchCtx::begin_finish()
{
    for (auto &W : workers)
    {
        for (auto &s : W.all_text)
            s.clear();
        for (auto &s : W.one_text)
            s.clear();
        W.unc_text.clear();
        W.m_off.clear();
        W.m_target.clear();
        W.m_count.clear();
        W.n_classified = 0;
    }
    timing.ms_host_finish = 0;
    for (auto &o : lv)
    {
        o.on_device = false;
        o.n_matches = o.all_len = o.one_len = 0;
        o.total = gnb_totals{};
    }
    unc_on_device    = false;
    unc_len_dev      = 0;
    finish_T         = 0;
    ms_finish_dev    = 0;
    levels_on_device = 0;
}

void BatchCtx::fill_timings(gnb_batch_result *t)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess)
        timing.ms_h2d = ms;
    if (n_reads && hashed_k && cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess)
        timing.ms_minimiser = ms;
    *t                   = timing;
    t->n_reads           = n_reads;
    t->parse_error       = parse_error ? 1 : 0;
    t->consumed1         = parse_error ? len1 : consumed1; // a parse error skips the rest of the file (GC.cpp:1278-1283)
    t->consumed2         = !paired ? 0 : parse_error ? len2 : consumed2;
    t->n_kernel_launches = launches;
    t->n_minimisers      = 0;
    if (max_hashes_ub <= 65535)
        t->n_minimisers = hashed_k ? total_hashes : 0;
    else
        for (uint32_t i = 0; i < n_reads && i < h_counts.size(); ++i)
            if (h_counts[i] <= 65535)
                t->n_minimisers += h_counts[i];
}

// merge the workers' pieces into the result buffers and the session's accounting
int BatchCtx::collect(uint32_t prefix_id, gnb_batch_result *out)
{
    auto           t_collect = Clock::now();
    const uint32_t n = n_reads;
    const int      T = (int)workers.size();
    const size_t   NL = levels.size();
    r_all.resize(NL);
    r_one.resize(NL);
    // sizes and per-worker offsets of every output piece
    std::vector<size_t> off_all(NL * (T + 1), 0), off_one(NL * (T + 1), 0), off_unc(T + 1, 0), off_m(T + 1, 0);
    uint64_t n_classified = 0;
    for (int w = 0; w < T; ++w)
    {
        for (size_t li = 0; li < NL; ++li)
        {
            off_all[li * (T + 1) + w + 1] = off_all[li * (T + 1) + w] + workers[w].all_text[li].size();
            off_one[li * (T + 1) + w + 1] = off_one[li * (T + 1) + w] + workers[w].one_text[li].size();
        }
        off_unc[w + 1] = off_unc[w] + workers[w].unc_text.size();
        off_m[w + 1]   = off_m[w] + workers[w].m_target.size();
        n_classified += workers[w].n_classified;
    }
    for (size_t li = 0; li < NL; ++li)
    {
        r_all[li].resize(off_all[li * (T + 1) + T]);
        r_one[li].resize(off_one[li * (T + 1) + T]);
    }
    r_unc.resize(off_unc[T]);
    uint64_t dev_matches = 0;
    bool     any_dev     = false;
    for (auto const &o : lv)
        if (o.on_device)
        {
            dev_matches += o.n_matches;
            n_classified += o.total.seqs_classified;
            any_dev = true;
        }
    const bool csr_device = NL == 1 && lv[0].on_device; // the level's CSR in pinned memory is the result
    if (!csr_device)
    {
        r_match_off.resize((size_t)n + 1);
        r_match_target.resize(off_m[T] + dev_matches);
        r_match_count.resize(off_m[T] + dev_matches);
    }
    // With one hierarchy level the workers' match lists, taken in worker order, already are the CSR value arrays
    // (worker w finished the reads [n*w/Tf, n*(w+1)/Tf) in order); with several levels the lists interleave.
    const bool csr_parallel = NL == 1 && finish_T > 0 && !any_dev;
    auto piece = [&](int w) {
        Worker &W = workers[w];
        for (size_t li = 0; li < NL; ++li)
        {
            if (!W.all_text[li].empty())
                memcpy(&r_all[li][off_all[li * (T + 1) + w]], W.all_text[li].data(), W.all_text[li].size());
            if (!W.one_text[li].empty())
                memcpy(&r_one[li][off_one[li * (T + 1) + w]], W.one_text[li].data(), W.one_text[li].size());
        }
        if (!W.unc_text.empty())
            memcpy(&r_unc[off_unc[w]], W.unc_text.data(), W.unc_text.size());
        if (csr_parallel)
        {
            if (!W.m_target.empty())
            {
                memcpy(&r_match_target[off_m[w]], W.m_target.data(), W.m_target.size() * 4);
                memcpy(&r_match_count[off_m[w]], W.m_count.data(), W.m_count.size() * 4);
            }
            if (w < finish_T)
            {
                const uint32_t r0 = (uint32_t)((uint64_t)n * w / finish_T), r1 = (uint32_t)((uint64_t)n * (w + 1) / finish_T);
                uint64_t       pos = off_m[w];
                size_t         nx  = 0;
                for (uint32_t r = r0; r < r1; ++r)
                {
                    r_match_off[r] = pos;
                    if (nx + 1 < W.m_off.size() && W.m_off[nx] == r)
                    {
                        pos += W.m_off[nx + 1];
                        nx += 2;
                    }
                }
            }
        }
    };
    if (n < 65536)
        for (int w = 0; w < T; ++w)
            piece(w);
    else
    {
        std::vector<std::thread> th;
        for (int w = 0; w < T; ++w)
            th.emplace_back(piece, w);
        for (auto &t : th)
            t.join();
    }
    if (csr_device)
        ;
    else if (csr_parallel)
        r_match_off[n] = off_m[T];
    else
    {
        std::fill(r_match_off.begin(), r_match_off.end(), 0);
        for (auto &W : workers)
            for (size_t i = 0; i + 1 < W.m_off.size(); i += 2)
                r_match_off[W.m_off[i] + 1] = W.m_off[i + 1];
        for (auto const &o : lv) // a read is classified at one level only
            if (o.on_device && o.n_matches)
            {
                const uint64_t *mo = o.h_moff.as<uint64_t>();
                for (size_t i = 0; i < n; ++i)
                    if (mo[i + 1] != mo[i])
                        r_match_off[i + 1] = mo[i + 1] - mo[i];
            }
        for (size_t i = 0; i < n; ++i)
            r_match_off[i + 1] += r_match_off[i];
        for (auto const &o : lv)
            if (o.on_device && o.n_matches)
            {
                const uint64_t *mo = o.h_moff.as<uint64_t>();
                const uint32_t *mt = o.h_mt.as<uint32_t>(), *mc = o.h_mc.as<uint32_t>();
                for (size_t i = 0; i < n; ++i)
                    if (mo[i + 1] != mo[i])
                    {
                        std::copy(mt + mo[i], mt + mo[i + 1], r_match_target.begin() + r_match_off[i]);
                        std::copy(mc + mo[i], mc + mo[i + 1], r_match_count.begin() + r_match_off[i]);
                    }
            }
        for (auto &W : workers)
        {
            size_t p = 0;
            for (size_t i = 0; i + 1 < W.m_off.size(); i += 2)
            {
                const uint64_t r = W.m_off[i], k = W.m_off[i + 1];
                std::copy(W.m_target.begin() + p, W.m_target.begin() + p + k, r_match_target.begin() + r_match_off[r]);
                std::copy(W.m_count.begin() + p, W.m_count.begin() + p + k, r_match_count.begin() + r_match_off[r]);
                p += k;
            }
        }
    }
    {
        std::lock_guard<std::mutex> lock(S->acc_mutex);
        S->ensure_prefix(prefix_id);
        for (size_t li = 0; li < NL; ++li)
        {
            LevelRt &L = levels[li];
            for (auto &W : workers)
            {
                for (auto const &[node, rp] : W.rep[li])
                    L.rep[prefix_id][node].add(rp);
                W.rep[li].clear();
                for (uint32_t node : W.rep_touched[li])
                {
                    L.rep[prefix_id][node].add(W.rep_dense[li][node]);
                    W.rep_dense[li][node] = Rep{};
                }
                W.rep_touched[li].clear();
                add_totals(L.total[prefix_id], W.total[li]);
                W.total[li] = gnb_totals{};
            }
            if (lv[li].on_device)
                add_totals(L.total[prefix_id], lv[li].total);
        }
        levels[0].total[prefix_id].input_seqs += n;
    }
    r_all_p.clear();
    r_one_p.clear();
    r_all_l.clear();
    r_one_l.clear();
    for (size_t li = 0; li < levels.size(); ++li)
    {
        r_all_p.push_back(r_all[li].data());
        r_all_l.push_back(r_all[li].size());
        r_one_p.push_back(r_one[li].data());
        r_one_l.push_back(r_one[li].size());
        if (lv[li].on_device)
        {
            r_all_p.back() = lv[li].h_all.as<char>();
            r_all_l.back() = lv[li].all_len;
            r_one_p.back() = lv[li].h_one.as<char>();
            r_one_l.back() = lv[li].one_len;
        }
    }
    timing.ms_d2h = (float)ms_since(t_collect); // host merge of the workers' pieces (no device copy happens here)
    fill_timings(&result);
    result.match_off    = csr_device ? lv[0].h_moff.as<uint64_t>() : r_match_off.data();
    result.match_target = csr_device ? lv[0].h_mt.as<uint32_t>() : r_match_target.data();
    result.match_count  = csr_device ? lv[0].h_mc.as<uint32_t>() : r_match_count.data();
    result.read_level   = h_read_level.data();
    result.n_hashes     = h_counts.data();
    result.n_classified = n_classified;
    result.n_levels     = (uint32_t)levels.size();
    result.all_text     = r_all_p.data();
    result.all_len      = r_all_l.data();
    result.one_text     = r_one_p.data();
    result.one_len      = r_one_l.data();
    result.unc_text     = unc_on_device ? h_unc.as<char>() : r_unc.data();
    result.unc_len      = unc_on_device ? unc_len_dev : r_unc.size();
    result.ms_finish_device = ms_finish_dev;
    result.levels_on_device = levels_on_device;
    trace_mark(this, "job.end");
    if (out)
        *out = result;
    staged = ran = false;
    return GNB_OK;
}

// all levels of the staged batch: K2/K3 (unless level 0 already ran), host finishing, merge
int BatchCtx::finish(uint32_t prefix_id, gnb_batch_result *out)
{
    GNB_CUDA(cudaSetDevice(device));
    begin_finish();
    keep_on_device = true;
    cur_prefix     = prefix_id;
    em_account     = true;
    struct TurnGuard
    {
        BatchCtx *c;
        ~TurnGuard()
        {
            c->wait_turn_noexcept();
            c->signal_done();
        }
    } turn_guard{this};
    trace_mark(this, "job.begin");
    for (size_t li = 0; li < levels.size(); ++li)
    {
        if (!(li == 0 && ran))
            GNB_TRY(run_level(li));
        auto th   = Clock::now();
        bool done = false;
        if (levels[li].device_finish && tuples_on_device)
        {
            unsigned long long *rep = nullptr;
            GNB_TRY(S->device_rep(li, prefix_id, &rep));
            GNB_TRY(finish_level_device(li, rep, true, done));
        }
        if (!done)
        {
            GNB_TRY(to_host_state(li));
            std::vector<size_t> off_before, tgt_before;
            for (auto const &W : workers)
            {
                off_before.push_back(W.m_off.size());
                tgt_before.push_back(W.m_target.size());
            }
            GNB_TRY(finish_level(li));
            if (S->keep_matches)
            {
                GNB_TRY(wait_turn()); // keeps the store in submission order
                GNB_TRY(em_append_host(li, off_before, tgt_before));
            }
        }
        timing.ms_host_finish += ms_since(th);
    }
    if (active_on_device && n_reads)
    { // classified level of every read for the structured result
        GNB_CUDA(cudaMemcpyAsync(h_read_level.data(), d_read_level.p, n_reads, cudaMemcpyDeviceToHost, st));
        GNB_CUDA(stream_wait(st));
        timing.d2h_bytes += n_reads;
    }
    return collect(prefix_id, out);
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
// A slot that is neither in flight nor holding the result the caller may still be reading.
int gnb_session::acquire_slot()
{
    for (int i = 0; i < (int)slots.size(); ++i)
    {
        bool used = i == holding && slots.size() > 1;
        for (int f : in_flight)
            used |= f == i;
        if (!used)
            return i;
    }
    return -1;
}

static int sync_slot(gnb_session *s)
{
    if (!s->in_flight.empty())
        return fail(GNB_ERR_ARG, "batches are in flight: call gnb_session_collect first");
    return 0;
}

extern "C" int gnb_session_stage(gnb_session *s, const char *b1, uint64_t l1, const char *b2, uint64_t l2, int fin, uint64_t *n_reads)
{
    if (!s || !b1)
        return fail(GNB_ERR_ARG, "gnb_session_stage: bad arguments");
    GNB_TRY(sync_slot(s));
    BatchCtx &c = *s->slots[0];
    c.seq       = 0;
    GNB_TRY(c.stage(b1, l1, b2, l2, fin));
    if (n_reads)
        *n_reads = c.n_reads;
    GNB_CUDA(stream_wait(c.st_in));
    return GNB_OK;
}

extern "C" int gnb_session_run_staged(gnb_session *s, gnb_batch_result *timings)
{
    if (!s || !s->slots[0]->staged)
        return fail(GNB_ERR_ARG, "gnb_session_run_staged: nothing staged");
    BatchCtx &c = *s->slots[0];
    GNB_CUDA(cudaSetDevice(c.device));
    // allow repeated runs on the same staged batch (benchmark): reset per-run state
    c.hashed_k = c.hashed_w = 0;
    c.timing.ms_count = c.timing.ms_sort = c.timing.ms_exchange = 0;
    c.timing.count_kernel_bytes = c.timing.exchanged_bytes = 0;
    c.timing.d2h_bytes = 0;
    c.launches = 0;
    std::fill(c.h_active.begin(), c.h_active.end(), (uint8_t)1);
    c.keep_on_device = true;
    c.ms_finish_dev  = 0;
    GNB_TRY(c.run_level(0));
    c.ran = true;
    if (s->levels[0].device_finish && c.tuples_on_device)
    { // K4 of the first level belongs to the pass; its report counters go to a scratch accumulator
        const size_t bytes = s->levels[0].node_names.size() * 5 * 8;
        if (!s->d_rep_scratch.p)
        {
            GNB_TRY(s->d_rep_scratch.ensure(bytes));
            GNB_CUDA(cudaMemset(s->d_rep_scratch.p, 0, bytes));
        }
        bool done    = false;
        c.em_account = false;
        GNB_TRY(c.finish_level_device(0, s->d_rep_scratch.as<unsigned long long>(), false, done));
        c.active_on_device = false; // finish_staged starts over from the staged state
        c.levels_on_device = 0;
    }
    if (timings)
    {
        c.fill_timings(timings);
        timings->ms_finish_device = c.ms_finish_dev;
    }
    return GNB_OK;
}

extern "C" int gnb_session_finish_staged(gnb_session *s, uint32_t prefix_id, gnb_batch_result *out)
{
    if (!s || !s->slots[0]->staged)
        return fail(GNB_ERR_ARG, "gnb_session_finish_staged: nothing staged");
    return s->slots[0]->finish(prefix_id, out);
}

// Level-wise form (bin-sharded multi-GPU): run K2/K3 of one level, expose / replace its tuples, finish the level.
extern "C" int gnb_session_run_level(gnb_session *s, uint32_t level)
{
    if (!s || !s->slots[0]->staged || level >= s->levels.size())
        return fail(GNB_ERR_ARG, "gnb_session_run_level: nothing staged or bad level");
    BatchCtx &c = *s->slots[0];
    GNB_CUDA(cudaSetDevice(c.device));
    if (level == 0)
        c.begin_finish();
    c.keep_on_device = false; // the tuples are exchanged between ranks on the host
    return c.run_level(level);
}

extern "C" int gnb_session_level_tuples(gnb_session *s, uint32_t level, uint32_t filter, const uint64_t **tuples, uint64_t *n)
{
    if (!s || !tuples || !n || level >= s->levels.size() || filter >= s->levels[level].filters.size())
        return fail(GNB_ERR_ARG, "gnb_session_level_tuples: bad arguments");
    auto &t = s->slots[0]->tuples[level][filter];
    *tuples = t.data();
    *n      = t.size();
    return GNB_OK;
}

extern "C" int gnb_session_set_level_tuples(gnb_session *s, uint32_t level, uint32_t filter, const uint64_t *tuples, uint64_t n)
{
    if (!s || (n && !tuples) || level >= s->levels.size() || filter >= s->levels[level].filters.size())
        return fail(GNB_ERR_ARG, "gnb_session_set_level_tuples: bad arguments");
    auto &t = s->slots[0]->tuples[level][filter];
    t.assign(tuples, tuples + n);
    if (!std::is_sorted(t.begin(), t.end(), [](uint64_t a, uint64_t b) { return (a >> kTupleNodeShift) < (b >> kTupleNodeShift); }))
        std::stable_sort(t.begin(), t.end(), [](uint64_t a, uint64_t b) { return (a >> kTupleNodeShift) < (b >> kTupleNodeShift); });
    return GNB_OK;
}

extern "C" int gnb_session_finish_level(gnb_session *s, uint32_t level)
{
    if (!s || !s->slots[0]->staged || level >= s->levels.size())
        return fail(GNB_ERR_ARG, "gnb_session_finish_level: nothing staged or bad level");
    BatchCtx &c = *s->slots[0];
    auto      th = Clock::now();
    GNB_TRY(c.to_host_state(level)); // record table; active mask if an earlier level was finished by K4
    GNB_TRY(c.finish_level(level));
    c.timing.ms_host_finish += ms_since(th);
    return GNB_OK;
}

// Device form of the level-wise API: the tuples of a single-filter level stay in HBM, the ranks exchange them with a
// device collective (NCCL all-gather) and K4 finishes the level on every rank.
extern "C" int gnb_session_run_level_device(gnb_session *s, uint32_t level)
{
    if (!s || !s->slots[0]->staged || level >= s->levels.size())
        return fail(GNB_ERR_ARG, "gnb_session_run_level_device: nothing staged or bad level");
    BatchCtx &c = *s->slots[0];
    GNB_CUDA(cudaSetDevice(c.device));
    if (level == 0)
    { // a staged batch may be run again (benchmark): start from the staged state
        c.begin_finish();
        c.hashed_k = c.hashed_w = 0;
        c.timing.ms_count = c.timing.ms_sort = c.timing.ms_exchange = 0;
        c.timing.count_kernel_bytes = c.timing.exchanged_bytes = 0;
        c.launches         = 0;
        c.active_on_device = false;
        std::fill(c.h_active.begin(), c.h_active.end(), (uint8_t)1);
    }
    c.keep_on_device = true;
    GNB_TRY(c.run_level(level));
    if (!c.tuples_on_device)
        return fail(GNB_ERR_ARG, "gnb_session_run_level_device: the level has several filters (use gnb_session_run_level)");
    return GNB_OK;
}

extern "C" int gnb_session_level_tuples_device(gnb_session *s, uint32_t level, const uint64_t **dev_tuples, uint64_t *n)
{
    if (!s || !dev_tuples || !n || level >= s->levels.size() || !s->slots[0]->tuples_on_device)
        return fail(GNB_ERR_ARG, "gnb_session_level_tuples_device: bad arguments or no tuples in device memory");
    BatchCtx &c = *s->slots[0];
    *dev_tuples = c.n_tuples_dev ? c.d_tuples_b.as<uint64_t>() : nullptr;
    *n          = c.n_tuples_dev;
    return GNB_OK;
}

extern "C" int gnb_session_set_level_tuples_device(gnb_session *s, uint32_t level, const uint64_t *dev_tuples, uint64_t n)
{
    if (!s || (n && !dev_tuples) || level >= s->levels.size() || !s->slots[0]->staged)
        return fail(GNB_ERR_ARG, "gnb_session_set_level_tuples_device: bad arguments");
    BatchCtx &c = *s->slots[0];
    GNB_CUDA(cudaSetDevice(c.device));
    if (n)
    {
        GNB_TRY(c.d_tuples_a.ensure(n * 8));
        GNB_TRY(c.d_tuples_b.ensure(n * 8));
        GNB_TRY(c.d_tmp.ensure(sort_tmp_bytes(n)));
        GNB_CUDA(cudaMemcpyAsync(c.d_tuples_a.p, dev_tuples, n * 8, cudaMemcpyDeviceToDevice, c.st));
        launch_sort_tuples(c.d_tuples_a.as<uint64_t>(), c.d_tuples_b.as<uint64_t>(), n, c.d_tmp.p, c.d_tmp.cap, c.st);
        c.launches += 4;
        GNB_CUDA(stream_wait(c.st));
        GNB_CUDA(cudaGetLastError());
    }
    c.n_tuples_dev     = n;
    c.tuples_on_device = true;
    return GNB_OK;
}

// K4 on the (exchanged) tuples of the level; the host finishing stage takes over if K4 declines (see finish_level_device)
extern "C" int gnb_session_finish_level_device(gnb_session *s, uint32_t level, uint32_t prefix_id)
{
    if (!s || !s->slots[0]->staged || level >= s->levels.size())
        return fail(GNB_ERR_ARG, "gnb_session_finish_level_device: nothing staged or bad level");
    BatchCtx &c = *s->slots[0];
    GNB_CUDA(cudaSetDevice(c.device));
    auto th   = Clock::now();
    bool done = false;
    c.cur_prefix = prefix_id;
    c.em_account = true;
    if (s->levels[level].device_finish && c.tuples_on_device)
    {
        unsigned long long *rep = nullptr;
        GNB_TRY(s->device_rep(level, prefix_id, &rep));
        GNB_TRY(c.finish_level_device(level, rep, true, done));
    }
    if (!done)
    {
        GNB_TRY(c.to_host_state(level));
        GNB_TRY(c.finish_level(level));
    }
    c.timing.ms_host_finish += ms_since(th);
    if (level + 1 == s->levels.size() && c.active_on_device && c.n_reads)
    {
        GNB_CUDA(cudaMemcpyAsync(c.h_read_level.data(), c.d_read_level.p, c.n_reads, cudaMemcpyDeviceToHost, c.st));
        GNB_CUDA(stream_wait(c.st));
    }
    return GNB_OK;
}

extern "C" int gnb_session_hibf_rounds(gnb_session *s, uint32_t cap, float *ms, uint64_t *bytes, uint64_t *items, uint32_t *n_rounds)
{
    if (!s || !n_rounds)
        return fail(GNB_ERR_ARG, "gnb_session_hibf_rounds: bad arguments");
    const BatchCtx &c = *s->slots[0];
    *n_rounds         = (uint32_t)c.hibf_round_ms.size();
    for (uint32_t i = 0; i < *n_rounds && i < cap; ++i)
    {
        if (ms)
            ms[i] = c.hibf_round_ms[i];
        if (bytes)
            bytes[i] = c.hibf_round_bytes[i];
        if (items)
            items[i] = c.hibf_round_items[i];
    }
    return GNB_OK;
}

extern "C" int gnb_session_staged_timings(gnb_session *s, gnb_batch_result *timings)
{
    if (!s || !timings || !s->slots[0]->staged)
        return fail(GNB_ERR_ARG, "gnb_session_staged_timings: nothing staged");
    BatchCtx &c = *s->slots[0];
    c.fill_timings(timings);
    timings->ms_finish_device = c.ms_finish_dev;
    timings->levels_on_device = c.levels_on_device;
    return GNB_OK;
}

extern "C" int gnb_session_collect_staged(gnb_session *s, uint32_t prefix_id, gnb_batch_result *out)
{
    if (!s || !s->slots[0]->staged)
        return fail(GNB_ERR_ARG, "gnb_session_collect_staged: nothing staged");
    return s->slots[0]->collect(prefix_id, out);
}

// Asynchronous form.  submit: index + copy the block to the device in the calling thread (the caller needs the consumed
// byte counts to cut the next block), then K2/K3/host finishing run in a worker thread on the slot's own stream while the
// caller stages the next block.  collect: results of the oldest batch, in submission order.
extern "C" int gnb_session_submit(gnb_session *s, uint32_t prefix_id, const char *b1, uint64_t l1, const char *b2, uint64_t l2, int fin,
                                  gnb_batch_result *staged_info)
{
    if (!s || !b1)
        return fail(GNB_ERR_ARG, "gnb_session_submit: bad arguments");
    const int slot = s->acquire_slot();
    if (slot < 0)
        return fail(GNB_ERR_LIMIT, "gnb_session_submit: all slots in flight, collect first");
    BatchCtx &c = *s->slots[slot];
    if (c.job.joinable())
        c.job.join();
    c.t_submit = Clock::now();
    GNB_TRY(c.stage(b1, l1, b2, l2, fin));
    if (staged_info)
        c.fill_timings(staged_info);
    c.busy   = true;
    c.job_rc = GNB_OK;
    c.seq            = s->next_seq++;
    c.turn_taken     = false;
    c.done_signalled = false;
    c.prev_done      = s->in_flight.empty() ? nullptr : s->slots[s->in_flight.back()]->ev_done;
    if (!c.prev_done)
    { // nothing to wait for: this batch opens a new chain
        std::lock_guard<std::mutex> lock(s->chain_mu);
        s->chain_recorded = std::max(s->chain_recorded, c.seq - 1);
    }
    c.job    = std::thread([&c, prefix_id]() {
        auto tj  = Clock::now();
        c.job_rc = c.finish(prefix_id, nullptr);
        c.result.ms_host_index = ms_since(tj); // async form: wall time of the worker job
        if (c.job_rc != GNB_OK)
            c.job_err = gnb_last_error();
    });
    s->in_flight.push_back(slot);
    return GNB_OK;
}

extern "C" int gnb_session_in_flight(const gnb_session *s, uint32_t *n, uint32_t *capacity)
{
    if (!s)
        return fail(GNB_ERR_ARG, "null argument");
    if (n)
        *n = (uint32_t)s->in_flight.size();
    if (capacity)
        *capacity = (uint32_t)std::max<size_t>(1, s->slots.size() - 1);
    return GNB_OK;
}

extern "C" int gnb_session_collect(gnb_session *s, gnb_batch_result *out)
{
    if (!s || !out)
        return fail(GNB_ERR_ARG, "gnb_session_collect: bad arguments");
    if (s->in_flight.empty())
        return fail(GNB_ERR_ARG, "gnb_session_collect: nothing in flight");
    const int slot = s->in_flight.front();
    s->in_flight.pop_front();
    BatchCtx &c = *s->slots[slot];
    if (c.job.joinable())
        c.job.join();
    c.busy     = false;
    s->holding = slot;
    if (c.job_rc != GNB_OK)
        return fail(c.job_rc, c.job_err);
    *out          = c.result;
    out->ms_total = ms_since(c.t_submit);
    return GNB_OK;
}

extern "C" int gnb_session_classify(gnb_session *s, uint32_t prefix_id, const char *b1, uint64_t l1, const char *b2, uint64_t l2, int fin,
                                    gnb_batch_result *out)
{
    if (!s || !b1 || !out)
        return fail(GNB_ERR_ARG, "gnb_session_classify: bad arguments");
    GNB_TRY(sync_slot(s));
    auto      t0 = Clock::now();
    BatchCtx &c  = *s->slots[0];
    s->holding   = 0;
    c.seq        = 0;
    GNB_TRY(c.stage(b1, l1, b2, l2, fin));
    GNB_TRY(c.finish(prefix_id, out));
    out->ms_total = ms_since(t0);
    return GNB_OK;
}

namespace gnb
{
// blocking host waits on / off (files.cpp); returns the previous setting.  GANON_B200_SYNC=spin keeps spinning.
int set_blocking_waits(int on)
{
    static const bool pinned_spin = [] { const char *e = getenv("GANON_B200_SYNC"); return e && e[0] == 's'; }();
    if (pinned_spin)
        return g_sync_block.load();
    return g_sync_block.exchange(on);
}
void session_ingest_mode(const gnb_session *s, int *sliced, int *rank, int *n_ranks)
{
    *sliced  = s->sharded() && s->sliced_ingest ? 1 : 0;
    *rank    = s->comm ? s->comm->rank : 0;
    *n_ranks = s->comm ? s->comm->n_ranks : 1;
}
} // namespace gnb

extern "C" int gnb_host_register(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes)
        return fail(GNB_ERR_ARG, "gnb_host_register: bad arguments");
    GNB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return GNB_OK;
}
extern "C" int gnb_host_unregister(void *ptr)
{
    if (!ptr)
        return fail(GNB_ERR_ARG, "gnb_host_unregister: bad arguments");
    GNB_CUDA(cudaHostUnregister(ptr));
    return GNB_OK;
}

extern "C" int gnb_session_level_count(const gnb_session *s, uint32_t *n)
{
    if (!s || !n)
        return fail(GNB_ERR_ARG, "null argument");
    *n = (uint32_t)s->levels.size();
    return GNB_OK;
}
extern "C" int gnb_session_level_label(const gnb_session *s, uint32_t level, const char **label)
{
    if (!s || !label || level >= s->levels.size())
        return fail(GNB_ERR_ARG, "level out of range");
    *label = s->levels[level].label.c_str();
    return GNB_OK;
}
extern "C" int gnb_session_node_name(const gnb_session *s, uint32_t level, uint32_t node, const char **name)
{
    if (!s || !name || level >= s->levels.size() || node >= s->levels[level].node_names.size())
        return fail(GNB_ERR_ARG, "node out of range");
    *name = s->levels[level].node_names[node].c_str();
    return GNB_OK;
}

static gnb_totals sum_levels(const gnb_session *s, uint32_t prefix_id)
{
    gnb_totals t{};
    for (auto const &L : s->levels)
        if (prefix_id < L.total.size())
            add_totals(t, L.total[prefix_id]);
    return t;
}

extern "C" int gnb_session_totals(const gnb_session *s, uint32_t prefix_id, int level, gnb_totals *out)
{
    if (!s || !out || level >= (int)s->levels.size())
        return fail(GNB_ERR_ARG, "gnb_session_totals: bad arguments");
    if (level < 0)
        *out = sum_levels(s, prefix_id);
    else
    {
        gnb_totals z{};
        *out = prefix_id < s->levels[level].total.size() ? s->levels[level].total[prefix_id] : z;
    }
    return GNB_OK;
}

// write_report (GC.cpp:834-853) per level in sorted label order, then write_report_totals (GC.cpp:855-863)
extern "C" int gnb_session_report(gnb_session *s, uint32_t prefix_id, const char **text, uint64_t *len)
{
    if (!s || !text || !len)
        return fail(GNB_ERR_ARG, "null argument");
    for (auto &c : s->slots) // the accumulators in HBM are folded in once no batch is running
        if (c->job.joinable())
            c->job.join();
    GNB_TRY(s->drain_device_rep());
    std::string &o = s->report_text;
    o.clear();
    for (auto const &L : s->levels)
    {
        if (prefix_id >= L.rep.size())
            continue;
        std::vector<uint32_t> nodes;
        for (auto const &kv : L.rep[prefix_id])
            nodes.push_back(kv.first);
        std::sort(nodes.begin(), nodes.end());
        for (uint32_t node : nodes)
        {
            const Rep &r = L.rep[prefix_id].at(node);
            if (!(r.matches || r.seqs_lca || r.seqs_unique))
                continue;
            o += L.label;
            o.push_back('\t');
            o += L.node_names[node];
            o.push_back('\t');
            append_u64(o, r.matches);
            o.push_back('\t');
            append_u64(o, r.seqs_unique);
            o.push_back('\t');
            append_u64(o, r.seqs_lca);
            if (L.has_tax)
            {
                o.push_back('\t');
                o += L.node_rank[node];
                o.push_back('\t');
                o += L.node_tax_name[node];
            }
            o.push_back('\n');
        }
    }
    // the reference creates a prefix's totals when its first batch of reads is queued (GC.cpp:1256, 1272): a prefix
    // without a single parsed read (empty file, parse error in the first chunk) gets no totals lines
    const gnb_totals t = sum_levels(s, prefix_id);
    if (t.input_seqs > 0)
    {
        o += "#total_classified\t";
        append_u64(o, t.seqs_classified);
        o += "\n#total_unclassified\t";
        append_u64(o, t.input_seqs - t.seqs_classified);
        o.push_back('\n');
    }
    *text = o.data();
    *len  = o.size();
    return GNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// EM reassignment of multi-matching reads (src/ganon/reassign.py; run by `ganon classify --multiple-matches em` on the
// `.all` / `.rep` files, src/ganon/classify.py:76-88) from the matches kept in HBM
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int gnb_session_keep_matches(gnb_session *s, int enable)
{
    if (!s)
        return fail(GNB_ERR_ARG, "null argument");
    if (!s->in_flight.empty())
        return fail(GNB_ERR_ARG, "gnb_session_keep_matches: batches are in flight");
    if (enable && s->em_names.empty())
    {
        GNB_CUDA(cudaSetDevice(s->device));
        GNB_TRY(s->build_em_tables());
    }
    s->keep_matches = enable != 0;
    return GNB_OK;
}

extern "C" int gnb_session_reassign(gnb_session *s, uint32_t prefix_id, double threshold, uint32_t max_iter, gnb_reassign_result *out)
{
    if (!s || !out)
        return fail(GNB_ERR_ARG, "null argument");
    if (!s->keep_matches)
        return fail(GNB_ERR_ARG, "gnb_session_reassign: call gnb_session_keep_matches before the first batch");
    // the report of the run: joins the workers, folds the accumulators in HBM into LevelRt::rep
    const char *rep_text = nullptr;
    uint64_t    rep_len  = 0;
    GNB_TRY(gnb_session_report(s, prefix_id, &rep_text, &rep_len));
    GNB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->slots[0]->st;
    const size_t NG = s->em_groups(), NT = s->em_names.size();
    s->em_one.assign(NG, "");
    s->em_label.assign(NG, "");
    s->em_iterations.assign(NG, 0);
    s->em_reassigned.assign(NG, 0);
    s->em_rep.clear();
    DevBuf  d_first, d_initial, d_weight, d_counts, d_len, d_loff, d_text, d_tmp, d_multi, d_keys_a, d_keys_b;
    EmStore merged; // a group's store with reads of equal ids made one (only when there are any)
    auto    cleanup = [&]() {
        for (DevBuf *b : {&d_first, &d_initial, &d_weight, &d_counts, &d_len, &d_loff, &d_text, &d_tmp, &d_multi, &d_keys_a, &d_keys_b})
            b->release();
        merged.release();
    };
    int rc = [&]() -> int {
        GNB_TRY(d_first.ensure(NT * 8 + 8));
        GNB_TRY(d_initial.ensure(NT * 8 + 8));
        GNB_TRY(d_weight.ensure(NT * 8 + 8));
        GNB_TRY(d_counts.ensure(NT * 8 + 8));
        GNB_TRY(d_multi.ensure(8));
        std::vector<unsigned long long> first(NT), initial(NT), counts(NT);
        for (size_t g = 0; g < NG; ++g)
        {
            if (NG > 1)
                s->em_label[g] = s->levels[g].label;
            static EmStore empty;
            EmStore &E = (prefix_id < s->em.size() && g < s->em[prefix_id].size()) ? s->em[prefix_id][g] : empty;
            std::vector<uint32_t> order; // targets present in the group's `.all`, in order of first appearance
            if (E.n_reads)
            {
                // reassign.py:78-85 keys the matches by read id: reads that share an id are one read (at the place of the
                // first, matches in file order).  Equal ids are looked for on the device; only if there are any is the
                // store regrouped (on the host, em_merge.cpp) -- the targets keep the numbering of the file order.
                EmStoreDev D       = E.dev();
                uint64_t   n_reads = E.n_reads;
                {
                    GNB_TRY(d_keys_a.ensure(E.n_reads * 8 + 8));
                    GNB_TRY(d_keys_b.ensure(E.n_reads * 8 + 8));
                    GNB_TRY(d_tmp.ensure(em_equal_ids_tmp_bytes(E.n_reads)));
                    launch_em_equal_ids(D, E.n_reads, d_keys_a.as<uint64_t>(), d_keys_b.as<uint64_t>(), d_tmp.p, d_tmp.cap, d_multi.as<unsigned long long>(), st);
                    unsigned long long n_equal = 0;
                    GNB_CUDA(cudaMemcpyAsync(&n_equal, d_multi.p, 8, cudaMemcpyDeviceToHost, st));
                    GNB_CUDA(stream_wait(st));
                    if (n_equal)
                    {
                        EmHost in, out;
                        in.off.resize(E.n_reads + 1);
                        in.id_off.resize(E.n_reads + 1);
                        in.tgt.resize(E.n_matches);
                        in.cnt.resize(E.n_matches);
                        in.ids.resize(E.id_bytes);
                        GNB_CUDA(cudaMemcpy(in.off.data(), E.off.p, (E.n_reads + 1) * 8, cudaMemcpyDeviceToHost));
                        GNB_CUDA(cudaMemcpy(in.id_off.data(), E.id_off.p, (E.n_reads + 1) * 8, cudaMemcpyDeviceToHost));
                        if (E.n_matches)
                        {
                            GNB_CUDA(cudaMemcpy(in.tgt.data(), E.tgt.p, E.n_matches * 4, cudaMemcpyDeviceToHost));
                            GNB_CUDA(cudaMemcpy(in.cnt.data(), E.cnt.p, E.n_matches * 4, cudaMemcpyDeviceToHost));
                        }
                        if (E.id_bytes)
                            GNB_CUDA(cudaMemcpy(in.ids.data(), E.ids.p, E.id_bytes, cudaMemcpyDeviceToHost));
                        if (em_merge_by_id(in, out) != 0)
                        {
                            merged.release();
                            GNB_TRY(merged.reserve(out.n_reads(), out.tgt.size(), out.ids.size()));
                            GNB_CUDA(cudaMemcpy(merged.off.p, out.off.data(), out.off.size() * 8, cudaMemcpyHostToDevice));
                            GNB_CUDA(cudaMemcpy(merged.id_off.p, out.id_off.data(), out.id_off.size() * 8, cudaMemcpyHostToDevice));
                            if (!out.tgt.empty())
                            {
                                GNB_CUDA(cudaMemcpy(merged.tgt.p, out.tgt.data(), out.tgt.size() * 4, cudaMemcpyHostToDevice));
                                GNB_CUDA(cudaMemcpy(merged.cnt.p, out.cnt.data(), out.cnt.size() * 4, cudaMemcpyHostToDevice));
                            }
                            if (!out.ids.empty())
                                GNB_CUDA(cudaMemcpy(merged.ids.p, out.ids.data(), out.ids.size(), cudaMemcpyHostToDevice));
                            merged.n_reads   = out.n_reads();
                            merged.n_matches = out.tgt.size();
                            merged.id_bytes  = out.ids.size();
                            n_reads          = merged.n_reads;
                            D                = merged.dev();
                        }
                    }
                }
                GNB_CUDA(cudaMemsetAsync(d_first.p, 0xFF, NT * 8, st));
                GNB_CUDA(cudaMemsetAsync(d_initial.p, 0, NT * 8, st));
                launch_em_first_pos(E.dev(), E.n_matches, d_first.as<unsigned long long>(), st); // numbering: file order
                launch_em_initial(D, n_reads, d_initial.as<unsigned long long>(), st);
                GNB_CUDA(cudaMemcpyAsync(first.data(), d_first.p, NT * 8, cudaMemcpyDeviceToHost, st));
                GNB_CUDA(cudaMemcpyAsync(initial.data(), d_initial.p, NT * 8, cudaMemcpyDeviceToHost, st));
                GNB_CUDA(stream_wait(st));
                for (uint32_t t = 0; t < NT; ++t)
                    if (first[t] != ~0ull)
                        order.push_back(t);
                std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return first[a] < first[b]; });
                // reassign.py:94-107: total weight = reads with matches, first probabilities from the unique matches
                const double total = (double)n_reads;
                uint64_t     uniq  = 0;
                for (uint32_t t : order)
                    uniq += initial[t];
                const double        denom0 = uniq ? (double)uniq : 1.0;
                std::vector<double> prob(NT, 0.0);
                for (uint32_t t : order)
                    prob[t] = (double)initial[t] / denom0;
                // the device compares integer weights: one denominator per iteration makes that the same order
                GNB_CUDA(cudaMemcpyAsync(d_weight.p, d_initial.p, NT * 8, cudaMemcpyDeviceToDevice, st));
                uint32_t it = 0;
                while (true)
                { // reassign.py:110-141
                    GNB_CUDA(cudaMemcpyAsync(d_counts.p, d_initial.p, NT * 8, cudaMemcpyDeviceToDevice, st));
                    launch_em_assign(D, n_reads, d_weight.as<unsigned long long>(), d_counts.as<unsigned long long>(), st);
                    GNB_CUDA(cudaMemcpyAsync(counts.data(), d_counts.p, NT * 8, cudaMemcpyDeviceToHost, st));
                    GNB_CUDA(cudaMemcpyAsync(d_weight.p, d_counts.p, NT * 8, cudaMemcpyDeviceToDevice, st));
                    GNB_CUDA(stream_wait(st));
                    double diff = 0;
                    for (uint32_t t : order)
                    {
                        const double np = (double)counts[t] / total;
                        diff += std::fabs(prob[t] - np);
                        prob[t] = np;
                    }
                    if (diff <= threshold)
                        break;
                    if (max_iter > 0 && it == max_iter - 1)
                        break;
                    ++it;
                }
                s->em_iterations[g] = it + 1;
                // `.one` (reassign.py:149-179): the unique match, or the top match under the final probabilities
                GNB_TRY(d_len.ensure((n_reads + 1) * 8));
                GNB_TRY(d_loff.ensure((n_reads + 1) * 8));
                GNB_TRY(d_tmp.ensure(em_scan64_tmp_bytes(n_reads + 1)));
                GNB_CUDA(cudaMemsetAsync(d_multi.p, 0, 8, st));
                launch_em_one(D, n_reads, d_weight.as<unsigned long long>(), s->d_em_name_off.as<uint32_t>(), s->d_em_names.as<char>(), d_len.as<uint64_t>(), nullptr,
                              nullptr, d_multi.as<unsigned long long>(), st);
                launch_scan64(d_len.as<uint64_t>(), d_loff.as<uint64_t>(), n_reads + 1, d_tmp.p, d_tmp.cap, st);
                uint64_t           bytes = 0;
                unsigned long long multi = 0;
                GNB_CUDA(cudaMemcpyAsync(&bytes, d_loff.as<uint64_t>() + n_reads, 8, cudaMemcpyDeviceToHost, st));
                GNB_CUDA(cudaMemcpyAsync(&multi, d_multi.p, 8, cudaMemcpyDeviceToHost, st));
                GNB_CUDA(stream_wait(st));
                s->em_reassigned[g] = multi;
                GNB_TRY(d_text.ensure(bytes + 1));
                launch_em_one(D, n_reads, d_weight.as<unsigned long long>(), s->d_em_name_off.as<uint32_t>(), s->d_em_names.as<char>(), d_len.as<uint64_t>(),
                              d_loff.as<uint64_t>(), d_text.as<char>(), nullptr, st);
                s->em_one[g].resize(bytes);
                if (bytes)
                    GNB_CUDA(cudaMemcpyAsync(&s->em_one[g][0], d_text.p, bytes, cudaMemcpyDeviceToHost, st));
                GNB_CUDA(stream_wait(st));
                GNB_CUDA(cudaGetLastError());
            }
            // new `.rep` rows of the group (reassign.py:188-214): the report's rows of targets present in the `.all`,
            // lca column = redistributed count of the last iteration - unique
            std::vector<uint8_t> present(NT, 0);
            for (uint32_t t : order)
                present[t] = 1;
            for (size_t li = 0; li < s->levels.size(); ++li)
            {
                const LevelRt &L = s->levels[li];
                if ((NG > 1 && li != g) || prefix_id >= L.rep.size())
                    continue;
                std::vector<uint32_t> nodes;
                for (auto const &kv : L.rep[prefix_id])
                    nodes.push_back(kv.first);
                std::sort(nodes.begin(), nodes.end());
                for (uint32_t node : nodes)
                {
                    const Rep &r = L.rep[prefix_id].at(node);
                    if (!(r.matches || r.seqs_lca || r.seqs_unique))
                        continue;
                    const uint32_t t = L.em_map[node];
                    if (!present[t])
                        continue;
                    std::string &o = s->em_rep;
                    o += L.label;
                    o.push_back('\t');
                    o += L.node_names[node];
                    o.push_back('\t');
                    append_u64(o, r.matches);
                    o.push_back('\t');
                    append_u64(o, r.seqs_unique);
                    o.push_back('\t');
                    const long long lca = (long long)counts[t] - (long long)r.seqs_unique;
                    if (lca < 0)
                    {
                        o.push_back('-');
                        append_u64(o, (uint64_t)(-lca));
                    }
                    else
                        append_u64(o, (uint64_t)lca);
                    o.push_back('\t');
                    if (L.has_tax)
                        o += L.node_rank[node];
                    o.push_back('\t');
                    if (L.has_tax)
                        o += L.node_tax_name[node];
                    o.push_back('\n');
                }
            }
        }
        return GNB_OK;
    }();
    cleanup();
    if (rc != GNB_OK)
        return rc;
    // the '#' lines of the report follow the rows (reassign.py:219-220)
    {
        const std::string rep(rep_text, rep_len);
        size_t            pos = 0;
        while (pos < rep.size())
        {
            size_t e = rep.find('\n', pos);
            if (e == std::string::npos)
                e = rep.size();
            if (rep[pos] == '#')
            {
                s->em_rep.append(rep, pos, e - pos);
                s->em_rep.push_back('\n');
            }
            pos = e + 1;
        }
    }
    s->em_one_p.clear();
    s->em_label_p.clear();
    s->em_one_l.clear();
    for (size_t g = 0; g < NG; ++g)
    {
        s->em_one_p.push_back(s->em_one[g].data());
        s->em_one_l.push_back(s->em_one[g].size());
        s->em_label_p.push_back(s->em_label[g].c_str());
    }
    out->n_groups         = (uint32_t)NG;
    out->group_label      = s->em_label_p.data();
    out->one_text         = s->em_one_p.data();
    out->one_len          = s->em_one_l.data();
    out->iterations       = s->em_iterations.data();
    out->reassigned_reads = s->em_reassigned.data();
    out->rep_text         = s->em_rep.data();
    out->rep_len          = s->em_rep.size();
    return GNB_OK;
}

// write_stats (GC.cpp:1130-1218)
extern "C" int gnb_session_stats(gnb_session *s, uint32_t prefix_id, const char *prefix_name, const char **text, uint64_t *len)
{
    if (!s || !text || !len)
        return fail(GNB_ERR_ARG, "null argument");
    std::string &o = s->stats_text;
    o = "prefix\thierarchy_label\tseq_processed\tseq_unclassified\tseq_classified\tseq_classified_perc\tseq_unique_matches\t"
        "seq_unique_matches_perc\tseq_multiple_matches\tseq_multiple_matches_perc\tmatches\tavg_matches_ref_seq\tdis_matches_rel_filter\t"
        "dis_matches_fpr_query\tkmers_proccessed\tkmers_matched\tkmers_from_classified_seqs\tkmers_matched_perc\n";
    const gnb_totals total           = sum_levels(s, prefix_id);
    const uint64_t   seq_unclassified = total.seqs_processed - total.seqs_classified;
    const double     seq_processed   = total.seqs_processed > 0 ? (double)total.seqs_processed : 1;
    const std::string pfx            = prefix_name ? prefix_name : "";
    auto line = [&](const gnb_totals &t, const std::string &label) {
        const uint64_t seq_multiple = t.seqs_classified - t.seqs_unique;
        const double   avg  = t.seqs_classified ? (t.matches / (double)t.seqs_classified) : 0;
        const double   perc = t.kmers_matches ? (t.kmers_matches / (double)t.kmers_from_classified_seqs) * 100 : 0;
        char           buf[1024];
        snprintf(buf, sizeof buf, "%s\t%s\t%llu\t%llu\t%llu\t%.6f\t%llu\t%.6f\t%llu\t%.6f\t%llu\t%.6f\t%llu\t%llu\t%llu\t%llu\t%llu\t%.6f\n", pfx.c_str(),
                 label.c_str(), (unsigned long long)(size_t)seq_processed, (unsigned long long)seq_unclassified, (unsigned long long)t.seqs_classified,
                 (t.seqs_classified / seq_processed) * 100, (unsigned long long)t.seqs_unique, (t.seqs_unique / seq_processed) * 100,
                 (unsigned long long)seq_multiple, (seq_multiple / seq_processed) * 100, (unsigned long long)t.matches, avg,
                 (unsigned long long)t.discarded_matches_filter, (unsigned long long)t.discarded_matches_fprquery, (unsigned long long)total.kmers_processed,
                 (unsigned long long)t.kmers_matches, (unsigned long long)t.kmers_from_classified_seqs, perc);
        o += buf;
    };
    for (auto const &L : s->levels)
    {
        gnb_totals z{};
        line(prefix_id < L.total.size() ? L.total[prefix_id] : z, L.label);
    }
    if (s->levels.size() > 1)
        line(total, "-total-");
    *text = o.data();
    *len  = o.size();
    return GNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel test hooks
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int gnb_minimisers_batch(int device, uint32_t k, uint32_t w, const char *seqs, const uint64_t *seq_off, uint64_t n, uint64_t *hash_off,
                                    uint64_t *hashes, uint64_t cap)
{
    if (!seqs || !seq_off || !hash_off || k < 1 || k > 32 || w < k || w - k + 1 > 256 || n >= kMaxReadsPerBatch || seq_off[n] >= (1ull << 32))
        return fail(GNB_ERR_ARG, "gnb_minimisers_batch: bad arguments");
    GNB_CUDA(cudaSetDevice(device));
    hash_off[0] = 0;
    if (n == 0)
        return GNB_OK;
    std::vector<uint32_t> off(n), len(n);
    for (uint64_t i = 0; i < n; ++i)
    {
        off[i] = (uint32_t)seq_off[i];
        len[i] = (uint32_t)(seq_off[i + 1] - seq_off[i]);
    }
    DevBuf d_seq, d_off, d_len, d_cnt, d_hoff, d_h, d_tmp;
    auto   cleanup = [&]() {
        for (DevBuf *b : {&d_seq, &d_off, &d_len, &d_cnt, &d_hoff, &d_h, &d_tmp})
            b->release();
    };
    int rc = GNB_OK;
    auto body = [&]() -> int {
        GNB_TRY(d_seq.ensure(seq_off[n] + 64));
        GNB_TRY(d_off.ensure(n * 4));
        GNB_TRY(d_len.ensure(n * 4));
        GNB_TRY(d_cnt.ensure(n * 4));
        GNB_TRY(d_hoff.ensure((n + 1) * 8));
        GNB_TRY(d_tmp.ensure(scan_tmp_bytes((uint32_t)n)));
        GNB_CUDA(cudaMemcpy(d_seq.p, seqs, seq_off[n], cudaMemcpyHostToDevice));
        GNB_CUDA(cudaMemcpy(d_off.p, off.data(), n * 4, cudaMemcpyHostToDevice));
        GNB_CUDA(cudaMemcpy(d_len.p, len.data(), n * 4, cudaMemcpyHostToDevice));
        uint32_t max_windows = 0;
        uint64_t windows     = 0;
        for (uint64_t i = 0; i < n; ++i)
            if (len[i] >= w)
            {
                max_windows = std::max(max_windows, len[i] - w + 1);
                windows += len[i] - w + 1;
            }
        if (minimisers_segmented(k, w, (uint32_t)n, max_windows))
        { // the session's path for long reads: one slot per window, a thread per segment, the lists moved together on the host side here
            DevBuf d_ub, d_uoff, d_items, d_ioff, d_seg_cnt, d_flags;
            auto   drop = [&]() {
                for (DevBuf *b : {&d_ub, &d_uoff, &d_items, &d_ioff, &d_seg_cnt, &d_flags})
                    b->release();
            };
            auto seg = [&]() -> int {
                const uint64_t bound = minimiser_segments_bound(windows, (uint32_t)n);
                GNB_TRY(d_ub.ensure(n * 4));
                GNB_TRY(d_uoff.ensure((n + 1) * 8));
                GNB_TRY(d_items.ensure(n * 4));
                GNB_TRY(d_ioff.ensure((n + 1) * 8));
                GNB_TRY(d_seg_cnt.ensure(bound * 4));
                GNB_TRY(d_flags.ensure(n));
                GNB_TRY(d_h.ensure((windows + 1) * 8));
                launch_hash_upper_bounds(d_len.as<uint32_t>(), nullptr, (uint32_t)n, w, d_ub.as<uint32_t>(), 0, d_items.as<uint32_t>(), nullptr);
                launch_scan_counts(d_ub.as<uint32_t>(), d_uoff.as<uint64_t>(), (uint32_t)n, d_tmp.p, d_tmp.cap, 0);
                launch_scan_counts(d_items.as<uint32_t>(), d_ioff.as<uint64_t>(), (uint32_t)n, d_tmp.p, d_tmp.cap, 0);
                launch_minimisers_segmented(d_seq.as<uint8_t>(), d_off.as<uint32_t>(), d_len.as<uint32_t>(), nullptr, nullptr, nullptr, (uint32_t)n, k, w,
                                            d_ioff.as<uint64_t>(), bound, d_seg_cnt.as<uint32_t>(), d_flags.as<uint8_t>(), d_cnt.as<uint32_t>(), d_uoff.as<uint64_t>(),
                                            d_h.as<uint64_t>(), nullptr, nullptr, 0);
                std::vector<uint32_t> cnt(n);
                std::vector<uint64_t> uoff(n + 1);
                GNB_CUDA(cudaMemcpy(cnt.data(), d_cnt.p, n * 4, cudaMemcpyDeviceToHost));
                GNB_CUDA(cudaMemcpy(uoff.data(), d_uoff.p, (n + 1) * 8, cudaMemcpyDeviceToHost));
                for (uint64_t i = 0; i < n; ++i)
                    hash_off[i + 1] = hash_off[i] + cnt[i];
                if (hashes && hash_off[n] <= cap && windows)
                {
                    std::vector<uint64_t> slots(windows);
                    GNB_CUDA(cudaMemcpy(slots.data(), d_h.p, windows * 8, cudaMemcpyDeviceToHost));
                    for (uint64_t i = 0; i < n; ++i)
                        memcpy(hashes + hash_off[i], slots.data() + uoff[i], (size_t)cnt[i] * 8);
                }
                GNB_CUDA(cudaGetLastError());
                return GNB_OK;
            };
            const int rc_seg = seg();
            drop();
            return rc_seg;
        }
        launch_minimisers(d_seq.as<uint8_t>(), d_off.as<uint32_t>(), d_len.as<uint32_t>(), nullptr, nullptr, nullptr, (uint32_t)n, k, w, 0,
                          d_cnt.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, 0);
        launch_scan_counts(d_cnt.as<uint32_t>(), d_hoff.as<uint64_t>(), (uint32_t)n, d_tmp.p, d_tmp.cap, 0);
        GNB_CUDA(cudaMemcpy(hash_off, d_hoff.p, (n + 1) * 8, cudaMemcpyDeviceToHost));
        const uint64_t total = hash_off[n];
        if (hashes && total <= cap && total > 0)
        {
            GNB_TRY(d_h.ensure(total * 8));
            launch_minimisers(d_seq.as<uint8_t>(), d_off.as<uint32_t>(), d_len.as<uint32_t>(), nullptr, nullptr, nullptr, (uint32_t)n, k, w, 1, nullptr,
                              d_hoff.as<uint64_t>(), d_h.as<uint64_t>(), nullptr, nullptr, 0);
            GNB_CUDA(cudaMemcpy(hashes, d_h.p, total * 8, cudaMemcpyDeviceToHost));
        }
        GNB_CUDA(cudaGetLastError());
        return GNB_OK;
    };
    rc = body();
    cleanup();
    return rc;
}

extern "C" int gnb_minimisers(int device, uint32_t k, uint32_t w, const char *seq, uint64_t len, uint64_t *out, uint64_t cap, uint64_t *n_out)
{
    if (!seq || !n_out || k < 1 || k > 32 || w < k || len >= (1ull << 31))
        return fail(GNB_ERR_ARG, "gnb_minimisers: bad arguments");
    *n_out = 0;
    // the kernel applies the classify() rule "shorter than the window -> skipped" (GC.cpp:690); the view itself shrinks
    // the window (minimiser.hpp:298-299), so clamp w for the hook
    const uint32_t w_eff = (uint32_t)std::min<uint64_t>(w, std::max<uint64_t>(len, k));
    const uint64_t so[2] = {0, len};
    uint64_t       ho[2] = {0, 0};
    std::vector<uint64_t> tmp(len + 1);
    int rc = gnb_minimisers_batch(device, k, w_eff, seq, so, 1, ho, tmp.data(), tmp.size());
    if (rc != GNB_OK)
        return rc;
    *n_out = ho[1];
    if (out)
        memcpy(out, tmp.data(), std::min<uint64_t>(ho[1], cap) * 8);
    return GNB_OK;
}

extern "C" int gnb_db_bulk_count(const gnb_db *db, uint64_t ibf_index, const uint64_t *hashes, const uint64_t *hash_off, uint64_t n_reads, uint16_t *counts)
{
    if (!db || ibf_index >= db->ibfs.size() || !hash_off || !counts || n_reads >= kMaxReadsPerBatch)
        return fail(GNB_ERR_ARG, "gnb_db_bulk_count: bad arguments");
    if (db->ibfs[ibf_index].paged())
        return fail(GNB_ERR_ARG, "gnb_db_bulk_count: not available on a paged filter");
    GNB_CUDA(cudaSetDevice(db->device));
    const IbfHost &ibf = db->ibfs[ibf_index];
    IbfDev         d{};
    d.data       = ibf.d_data;
    d.bin_size   = ibf.bin_size;
    d.hash_shift = (uint32_t)ibf.hash_shift;
    d.hash_funs  = (uint32_t)ibf.hash_funs;
    d.row_words  = (uint32_t)ibf.row_words();
    d.n_chunks   = (d.row_words + 63) / 64;
    const uint64_t total = hash_off[n_reads];
    uint32_t       mx    = 0;
    for (uint64_t i = 0; i < n_reads; ++i)
        mx = (uint32_t)std::max<uint64_t>(mx, hash_off[i + 1] - hash_off[i]);
    if (mx > 65535)
        return fail(GNB_ERR_LIMIT, "gnb_db_bulk_count: more than 65535 hashes in one list");
    uint64_t *d_h = nullptr, *d_o = nullptr;
    uint16_t *d_c = nullptr;
    const size_t cbytes = n_reads * (size_t)d.row_words * 64 * 2;
    GNB_CUDA(cudaMalloc((void **)&d_h, (total + 1) * 8));
    GNB_CUDA(cudaMalloc((void **)&d_o, (n_reads + 1) * 8));
    GNB_CUDA(cudaMalloc((void **)&d_c, cbytes + 2));
    cudaMemcpy(d_h, hashes, total * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_o, hash_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice);
    cudaMemset(d_c, 0, cbytes);
    launch_ibf_count_dense(d, d_h, d_o, (uint32_t)n_reads, mx, d_c, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaMemcpy(counts, d_c, cbytes, cudaMemcpyDeviceToHost);
    cudaFree(d_h);
    cudaFree(d_o);
    cudaFree(d_c);
    GNB_CUDA(e);
    return GNB_OK;
}
