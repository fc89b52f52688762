// Host-side database objects: parsed .ibf / .hibf headers + bitvectors resident in HBM.
#pragma once
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "gnb_internal.h"

namespace gnb
{

// Host-resident tier: a filter larger than the HBM it may use is cut into column pages (bin-word ranges of
// every row, the layout of a shard).  Some pages stay in HBM, the others live in page-locked host memory and rotate through
// two staging buffers in HBM while K3 works on the page before them.
struct IbfPage
{
    uint64_t  w0 = 0, w1 = 0;
    uint64_t *d_data = nullptr; // resident page
    uint64_t *h_data = nullptr; // streamed page (page-locked host memory), [bin_size][w1 - w0]
};

struct IbfHost
{
    uint64_t  bins = 0, technical_bins = 0, bin_size = 0, hash_shift = 0, bin_words = 0, hash_funs = 0;
    uint64_t  w0 = 0, w1 = 0;   // bin-word columns [w0, w1) held on the device
    uint64_t *d_data = nullptr; // [bin_size][w1 - w0]; nullptr for a paged filter
    std::vector<IbfPage> pages; // empty: everything in d_data
    uint64_t    *d_stage[2] = {nullptr, nullptr};
    cudaEvent_t  ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    cudaStream_t copy_st = nullptr;
    uint64_t     stage_bytes = 0;
    bool      paged() const { return !pages.empty(); }
    uint64_t  row_words() const { return w1 - w0; }
    uint64_t  page_bytes(const IbfPage &p) const { return bin_size * (p.w1 - p.w0) * 8; }
    uint64_t  device_bytes() const
    {
        if (!paged())
            return bin_size * row_words() * 8;
        uint64_t b = 2 * stage_bytes;
        for (auto const &p : pages)
            if (p.d_data)
                b += page_bytes(p);
        return b;
    }
    uint64_t host_bytes() const
    {
        uint64_t b = 0;
        for (auto const &p : pages)
            if (p.h_data)
                b += page_bytes(p);
        return b;
    }
};

// cut into pages for `budget` bytes of HBM; false: does not even hold two one-word pages.  n_resident pages stay in HBM.
bool plan_pages(const IbfHost &t, uint64_t budget, std::vector<IbfPage> &pages, size_t &n_resident);
int  alloc_page_storage(IbfHost &t, size_t n_resident);
void free_page_storage(IbfHost &t);

} // namespace gnb

struct gnb_db
{
    bool     is_hibf = false;
    int      device  = 0;
    uint32_t kmer_size = 0, window_size = 0;
    uint64_t max_hashes_bin = 0;
    double   max_fp = 0;
    std::vector<gnb::IbfHost> ibfs; // 1 for a flat IBF
    // flat IBF (.ibf): IBFConfig + hashes_count_std + bin_map (GanonBuild.cpp:251-288)
    int                                            version[3] = {2, 4, 1};
    double                                         true_max_fp = 0, true_avg_fp = 0;
    std::vector<std::pair<std::string, uint64_t>>  hashes_count;
    std::vector<std::pair<uint64_t, std::string>>  bin_map; // (technical bin | user bin, target)
    // derived: filter.map (GC.cpp:1021-1025) + target_fpr (GC.cpp:969-982 / 932)
    std::vector<std::string>           target_names;
    std::vector<double>                target_fpr;
    std::vector<std::vector<uint64_t>> target_bins; // flat: technical bins; hibf: user bins
    // HIBF (HIBF.hpp:124-136, 176-188)
    std::vector<std::vector<int64_t>> next_ibf_id, bin_to_user;
    uint64_t                          n_user_bins = 0;
    std::mutex                        page_mu; // a paged filter's staging buffers serve one level pass at a time

    void derive_targets();
};
