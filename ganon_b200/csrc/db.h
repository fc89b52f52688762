// Host-side database objects: parsed .ibf / .hibf headers + bitvectors resident in HBM.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "gnb_internal.h"

namespace gnb
{

struct IbfHost
{
    uint64_t  bins = 0, technical_bins = 0, bin_size = 0, hash_shift = 0, bin_words = 0, hash_funs = 0;
    uint64_t  w0 = 0, w1 = 0;   // bin-word columns [w0, w1) held on the device
    uint64_t *d_data = nullptr; // [bin_size][w1 - w0]
    uint64_t  row_words() const { return w1 - w0; }
    uint64_t  device_bytes() const { return bin_size * row_words() * 8; }
};

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

    void derive_targets();
};
