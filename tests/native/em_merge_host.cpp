// Host build of the EM store regrouping (ganon_b200/csrc/em_merge.cpp, plain C++) for the CPU test suite: the same source
// the product links, behind a small C interface for ctypes.
#include "../../ganon_b200/csrc/em_merge.cpp"

extern "C"
{
    // in: CSR arrays of n reads; out arrays sized like the inputs; returns the number of reads after merging
    uint64_t emh_merge(uint64_t n, const uint64_t *off, const uint32_t *tgt, const uint32_t *cnt, const uint64_t *id_off, const char *ids, uint64_t *o_off,
                       uint32_t *o_tgt, uint32_t *o_cnt, uint64_t *o_id_off, char *o_ids)
    {
        gnb::EmHost in, out;
        in.off.assign(off, off + n + 1);
        in.id_off.assign(id_off, id_off + n + 1);
        in.tgt.assign(tgt, tgt + off[n]);
        in.cnt.assign(cnt, cnt + off[n]);
        in.ids.assign(ids, ids + id_off[n]);
        const uint64_t gone = gnb::em_merge_by_id(in, out);
        const uint64_t g    = n - gone;
        memcpy(o_off, out.off.data(), (g + 1) * 8);
        memcpy(o_id_off, out.id_off.data(), (g + 1) * 8);
        memcpy(o_tgt, out.tgt.data(), out.tgt.size() * 4);
        memcpy(o_cnt, out.cnt.data(), out.cnt.size() * 4);
        memcpy(o_ids, out.ids.data(), out.ids.size());
        return g;
    }
}
