// Host build of the record reader of the library (ganon_b200/csrc/reads.cpp, plain C++) for the CPU test suite: the same
// source the product links, behind a small C interface for ctypes.
#include "../../ganon_b200/csrc/reads.cpp"

extern "C"
{
    void *rh_index(const char *block, uint64_t len, int final_block, uint64_t max_records)
    {
        auto *t = new gnb::RecTable();
        gnb::index_reads_host(block, len, final_block != 0, max_records, *t);
        return t;
    }
    uint64_t    rh_size(void *h) { return static_cast<gnb::RecTable *>(h)->size(); }
    uint64_t    rh_consumed(void *h) { return static_cast<gnb::RecTable *>(h)->consumed; }
    int         rh_error(void *h) { return static_cast<gnb::RecTable *>(h)->parse_error ? 1 : 0; }
    uint64_t    rh_error_record(void *h) { return static_cast<gnb::RecTable *>(h)->error_record; }
    const char *rh_error_msg(void *h) { return static_cast<gnb::RecTable *>(h)->error_msg.c_str(); }
    // record i: id and sequence spans (the sequence lies in the block, or in the table's aux area when it was not contiguous)
    void rh_get(void *h, const char *block, uint64_t len, uint64_t i, const char **id, uint32_t *id_len, const char **seq, uint32_t *seq_len)
    {
        auto *t  = static_cast<gnb::RecTable *>(h);
        *id      = block + t->id_off[i];
        *id_len  = t->id_len[i];
        *seq_len = t->seq_len[i];
        *seq     = t->seq_off[i] >= len ? reinterpret_cast<const char *>(t->aux.data()) + (t->seq_off[i] - len) : block + t->seq_off[i];
    }
    uint64_t rh_rec_end(void *h, uint64_t i) { return static_cast<gnb::RecTable *>(h)->rec_end[i]; }
    void     rh_free(void *h) { delete static_cast<gnb::RecTable *>(h); }
}
