#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gnb
{

struct RecTable
{
    std::vector<uint32_t> id_off, id_len, seq_off, seq_len; // seq_off >= block length: offset (minus length) into aux
    std::vector<uint32_t> rec_end;                          // one past the last byte of each record
    std::vector<uint8_t>  aux;
    uint64_t              aux_records  = 0;
    uint64_t              consumed     = 0;
    bool                  parse_error  = false;
    uint64_t              error_record = 0;
    std::string           error_msg;
    size_t                size() const { return id_off.size(); }
    void                  clear()
    {
        id_off.clear();
        id_len.clear();
        seq_off.clear();
        seq_len.clear();
        rec_end.clear();
        aux.clear();
        aux_records  = 0;
        consumed     = 0;
        parse_error  = false;
        error_record = 0;
        error_msg.clear();
    }
    void truncate(size_t n)
    {
        id_off.resize(n);
        id_len.resize(n);
        seq_off.resize(n);
        seq_len.resize(n);
        rec_end.resize(n);
    }
    uint64_t consumed_for(size_t n) const { return n ? rec_end[n - 1] : 0; }
};

bool block_is_fasta(const char *b, uint64_t len);
// Index up to max_records complete records of the block.  final: the block ends the file.
void index_reads_host(const char *b, uint64_t len, bool final, uint64_t max_records, RecTable &t);

} // namespace gnb
