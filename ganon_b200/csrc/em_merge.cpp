// See em_merge.h.
#include "em_merge.h"

#include <cstring>
#include <string_view>
#include <unordered_map>

namespace gnb
{
uint64_t em_merge_by_id(const EmHost &in, EmHost &out)
{
    const uint64_t n = in.n_reads();
    std::unordered_map<std::string_view, uint64_t> group_of_id;
    group_of_id.reserve((size_t)n * 2);
    std::vector<uint64_t> group(n), leader, size;
    for (uint64_t r = 0; r < n; ++r)
    {
        const std::string_view id(in.ids.data() + in.id_off[r], (size_t)(in.id_off[r + 1] - in.id_off[r]));
        const auto             it = group_of_id.try_emplace(id, (uint64_t)leader.size());
        if (it.second)
        {
            leader.push_back(r);
            size.push_back(0);
        }
        group[r] = it.first->second;
        size[group[r]] += in.off[r + 1] - in.off[r];
    }
    const uint64_t g = leader.size();
    out.off.assign(g + 1, 0);
    out.id_off.assign(g + 1, 0);
    for (uint64_t i = 0; i < g; ++i)
    {
        out.off[i + 1]    = out.off[i] + size[i];
        out.id_off[i + 1] = out.id_off[i] + (in.id_off[leader[i] + 1] - in.id_off[leader[i]]);
    }
    out.tgt.resize(in.tgt.size());
    out.cnt.resize(in.cnt.size());
    out.ids.resize((size_t)out.id_off[g]);
    for (uint64_t i = 0; i < g; ++i)
        memcpy(out.ids.data() + out.id_off[i], in.ids.data() + in.id_off[leader[i]], (size_t)(out.id_off[i + 1] - out.id_off[i]));
    std::vector<uint64_t> cursor(out.off.begin(), out.off.end() - 1);
    for (uint64_t r = 0; r < n; ++r)
    {
        const uint64_t a = in.off[r], len = in.off[r + 1] - a;
        uint64_t      &c = cursor[group[r]];
        if (len)
        {
            memcpy(out.tgt.data() + c, in.tgt.data() + a, (size_t)len * 4);
            memcpy(out.cnt.data() + c, in.cnt.data() + a, (size_t)len * 4);
        }
        c += len;
    }
    return n - g;
}
} // namespace gnb
