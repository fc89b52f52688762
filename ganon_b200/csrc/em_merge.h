// EM reassignment: reads that share an id.  src/ganon/reassign.py:78-85 keys its dictionary of matches by the read id of
// the `.all` lines, so two classified reads with the same id are ONE read to it: it stands where the first of them stood,
// and its matches are theirs in file order.  The store in HBM keeps reads by position; gnb_session_reassign looks for equal
// ids (hashes sorted on the device) and, only when there are any, regroups the store with this function.
// Plain C++ (no CUDA types): tests/native/em_merge_host.cpp compiles it for the CPU suite.
#pragma once
#include <cstdint>
#include <vector>

namespace gnb
{
struct EmHost // CSR over reads: matches (tgt, cnt) and id bytes, as in EmStoreDev
{
    std::vector<uint64_t> off, id_off; // [n_reads + 1]
    std::vector<uint32_t> tgt, cnt;
    std::vector<char>     ids;
    uint64_t              n_reads() const { return off.empty() ? 0 : off.size() - 1; }
};
// out = in with equal ids merged (first position, matches concatenated in order); returns the number of reads merged away
uint64_t em_merge_by_id(const EmHost &in, EmHost &out);
} // namespace gnb
